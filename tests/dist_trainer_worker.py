"""Worker of tests/test_dist_gpu.py: one rank of a data-parallel FusedActorTrainer run (launched by torchrun, NCCL).
Checks, on the real kernels: (1) all-reduced shard gradients == gradient of the gathered batch, (2) the sharded
trainer walks the trajectory of a single-process trainer on the full batch, (3) replicas stay bit-identical,
(4) CUDA-graph replay with the captured, overlapped all-reduce == eager steps, (5) the critic update's data-parallel form
(update_critic(process_group=...)) == the single-process update on the gathered batch, replicas identical, (6) the
row-sharded action ascent with the per-iteration all-reduce of the clip norm (SURVEY 8e, H2 semantics (ii)) == the ascent of
one process on the gathered, mode-sorted batch, (7) the process group tears down with the trainer closed.
Exit code 0 = all checks passed on this rank."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import port                                   # noqa: E402
from ddiffpg_b200 import (DiffusionPolicy, DistributionalDoubleQ, FusedActorTrainer, FusedCriticTrainer,   # noqa: E402
                          q_action_ascent_segments, update_critic)
from ddiffpg_b200 import dist as ddist                    # noqa: E402


def make(p, T, dev):
    pol = DiffusionPolicy(34, 8, T, device="cuda")
    pol.load_state_dict(p)
    return pol.to(dev)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    T, Bl = 5, 1536                                       # rows per rank
    B = Bl * world
    gen = torch.Generator().manual_seed(4242)             # the same full batch on every rank
    full = (torch.randn(B, 34, generator=gen), torch.rand(B, 8, generator=gen) * 2 - 1, torch.randn(B, 8, generator=gen),
            torch.randint(0, T, (B,), generator=gen))
    full = [x.to(dev) for x in full]
    lo, hi = ddist.shard_bounds(B, world, rank)
    mine = [x[lo:hi].contiguous() for x in full]
    p = port.init_actor_params(77)

    for precision, tol in (("fp32", 1e-5), ("bf16", 2e-3)):
        # (1) reduced shard gradient == full-batch gradient
        pol = make(p, T, dev)
        inv = 1.0 / (B * 8)
        l_sh, g_sh = pol._loss_and_grads(*mine, inv_count=inv, precision=precision)
        l_sh, g_sh = l_sh.clone(), g_sh.clone()
        ddist.allreduce_sum_(g_sh, l_sh)
        l_full, g_full = pol._loss_and_grads(*full, inv_count=inv, precision=precision)
        rel = ((g_sh - g_full).norm() / g_full.norm()).item()
        assert rel <= tol, f"{precision}: reduced vs full-batch gradient rel-L2 {rel:.3e}"
        assert abs(l_sh.item() - l_full.item()) <= tol * abs(l_full.item()), (precision, l_sh.item(), l_full.item())

        # (2) sharded trainer == single-process trainer on the full batch, (3) replicas identical, (4) graph == eager
        results = {}
        for graph in (False, True):
            tr = FusedActorTrainer(make(p, T, dev), precision=precision, graph=graph)
            ref = FusedActorTrainer(make(p, T, dev), precision=precision, graph=False, process_group=False)
            for it in range(4):
                loss, gn = tr.step(*mine[:2], noise=mine[2], timesteps=mine[3])
                l_ref, gn_ref = ref.step(*full[:2], noise=full[2], timesteps=full[3], global_batch=B)
                assert abs(loss.item() - l_ref.item()) <= 10 * tol * abs(l_ref.item()), (precision, graph, it, loss.item(), l_ref.item())
                assert abs(gn.item() / gn_ref.item() - 1) <= 10 * tol, (precision, graph, it, gn.item(), gn_ref.item())
            flat = tr.flat.double()
            mine_sum = torch.stack([flat.sum(), flat.abs().sum()])
            allc = [torch.zeros_like(mine_sum) for _ in range(world)]
            dist.all_gather(allc, mine_sum)
            assert all(torch.equal(allc[0], c) for c in allc), f"{precision} graph={graph}: replicas diverged"
            # Adam normalises steps to ~lr: compare where the reference moved decisively, in units of lr
            d_tr, d_ref = tr.flat - torch.cat([p[k].reshape(-1) for k in port.ACTOR_KEYS]).to(dev), None
            d_ref = ref.flat - torch.cat([p[k].reshape(-1) for k in port.ACTOR_KEYS]).to(dev)
            big = d_ref.abs() > 0.5 * 4 * 3e-4
            agree = ((d_tr - d_ref).abs()[big] <= 0.25 * 4 * 3e-4).float().mean().item()
            assert agree >= (0.999 if precision == "fp32" else 0.98), (precision, graph, agree)
            results[graph] = tr.flat.clone()
            tr.close()
            ref.close()
        # graph replay and eager launches run the same kernels in the same order
        # (bf16 path: fp32 atomics in the dW GEMMs land in a different order from run to run, and Adam turns a sign flip
        # of a near-zero gradient into a step of lr -- a handful of elements may differ by up to the 4 steps taken)
        d = (results[True] - results[False]).abs()
        if precision == "fp32":
            assert d.max().item() <= 1e-6, (precision, d.max().item())
        else:
            assert d.max().item() <= 4 * 3e-4 * 1.05 and (d > 1e-4).float().mean().item() <= 1e-3, (precision, d.max().item())
    # (5) N1: sharded update_critic == update on the full batch (three steps of torch's AdamW on both), replicas identical
    def critic(params):
        c = DistributionalDoubleQ(29, 8, v_min=0, v_max=5, num_atoms=51, device="cuda")
        c.load_state_dict(params)
        return c.to(dev)
    gc = torch.Generator().manual_seed(99)
    cb = [torch.randn(B, 29, generator=gc), torch.rand(B, 8, generator=gc) * 2 - 1, torch.rand(B, 1, generator=gc),
          torch.randn(B, 29, generator=gc), torch.rand(B, 8, generator=gc) * 2 - 1, (torch.rand(B, 1, generator=gc) < 0.2).float()]
    cb = [x.to(dev) for x in cb]
    cmine = [x[lo:hi].contiguous() for x in cb]
    pc, pt = port.init_critic_params(5), port.init_critic_params(6)
    for precision, tol in (("fp32", 1e-5), ("bf16", 5e-3)):
        sh, ref, tgt = critic(pc), critic(pc), critic(pt).requires_grad_(False)
        sh.train_precision = ref.train_precision = precision
        o_sh, o_ref = torch.optim.AdamW(sh.parameters(), lr=5e-4), torch.optim.AdamW(ref.parameters(), lr=5e-4)
        for it in range(3):
            _, l_sh, n_sh = update_critic(sh, tgt, o_sh, *cmine, gamma_n=0.97, process_group=dist.group.WORLD)
            _, l_ref, n_ref = update_critic(ref, tgt, o_ref, *cb, gamma_n=0.97)
            assert abs(l_sh - l_ref) <= 10 * tol * abs(l_ref), (precision, it, l_sh, l_ref)
            assert abs(n_sh / n_ref - 1) <= 10 * tol, (precision, it, n_sh, n_ref)
        flat = torch.cat([q.detach().reshape(-1) for q in sh.parameters()]).double()
        mine_sum = torch.stack([flat.sum(), flat.abs().sum()])
        allc = [torch.zeros_like(mine_sum) for _ in range(world)]
        dist.all_gather(allc, mine_sum)
        assert all(torch.equal(allc[0], c) for c in allc), f"critic {precision}: replicas diverged"
        # the fused form (flat-vector clip + AdamW, graph-captured all-reduce): sharded == gathered batch, replicas identical
        ftr = FusedCriticTrainer(critic(pc), tgt, lr=5e-4, precision=precision, graph=True)
        fref = FusedCriticTrainer(critic(pc), tgt, lr=5e-4, precision=precision, graph=False, process_group=False)
        for it in range(4):
            l_f, n_f = ftr.step(*cmine, gamma_n=0.97)
            l_r, n_r = fref.step(*cb, gamma_n=0.97)
            assert abs(l_f.item() - l_r.item()) <= 10 * tol * abs(l_r.item()), (precision, it, l_f.item(), l_r.item())
            assert abs(n_f.item() / n_r.item() - 1) <= 10 * tol, (precision, it, n_f.item(), n_r.item())
        mine_sum = torch.stack([ftr.flat.double().sum(), ftr.flat.double().abs().sum()])
        allc = [torch.zeros_like(mine_sum) for _ in range(world)]
        dist.all_gather(allc, mine_sum)
        assert all(torch.equal(allc[0], c) for c in allc), f"fused critic {precision}: replicas diverged"
        ftr.close()
        fref.close()
    # (6) H2, global-batch semantics: every rank holds a slice of each mode segment; with the K-float all-reduce of sum g^2
    # per iteration the ranks take the steps one process takes on the gathered batch (clip active in the second pass: max_norm = 1e-4)
    ga = torch.Generator().manual_seed(31)
    seg_off = [0, 1000, 1000 + 1700, 1000 + 1700 + 372]     # three modes, ragged, the last shorter than a tile per rank
    Bq = seg_off[-1]
    q_obs, q_act = torch.randn(Bq, 29, generator=ga).to(dev), (torch.rand(Bq, 8, generator=ga) * 2 - 1).to(dev)
    critics = [critic(port.init_critic_params(40 + m)).requires_grad_(False) for m in range(3)]
    pairs, local_off, counts = ddist.shard_segments(seg_off, world, rank)
    rows = torch.cat([torch.arange(a, b) for a, b in pairs]).to(dev)
    for precision, tol in (("fp32", 2e-5), ("bf16", 2e-3)):
        for max_norm in (1.0, 1e-4):
            whole = q_act.clone()
            ma_ref, n_ref = q_action_ascent_segments(critics, q_obs, whole, seg_off, iters=20, max_norm=max_norm,
                                                     precision=precision, return_norms=True)
            part = q_act[rows].contiguous()
            ma_sh, n_sh = q_action_ascent_segments(critics, q_obs[rows].contiguous(), part, local_off, iters=20,
                                                   max_norm=max_norm, precision=precision, return_norms=True,
                                                   mean_counts=counts, process_group=dist.group.WORLD)
            assert (n_sh / n_ref - 1).abs().max().item() <= 10 * tol, (precision, max_norm, "norms", n_sh, n_ref)
            assert (ma_sh - ma_ref).abs().max().item() <= 10 * tol, (precision, max_norm, ma_sh, ma_ref)
            d = (part - whole[rows]).abs()
            if precision == "fp32":
                assert d.max().item() <= 1e-4, (precision, max_norm, d.max().item())
            else:       # identical arithmetic per row; only the fp32 atomics of the norm land in a different order
                assert d.mean().item() <= 1e-4 and (d > 1e-2).float().mean().item() <= 1e-3, (precision, max_norm, d.mean().item())
            # without the exchange step the shard-local norm differs (the clip is active at max_norm = 1e-4)
            if max_norm < 1.0:
                loc = q_act[rows].contiguous()
                q_action_ascent_segments(critics, q_obs[rows].contiguous(), loc, local_off, iters=20, max_norm=max_norm,
                                         precision=precision, mean_counts=counts)
                assert (loc - whole[rows]).abs().max().item() > 1e-3, "the clip was not active: the test does not discriminate"
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()                          # (7)
    print(f"rank {rank}: ok", flush=True)


if __name__ == "__main__":
    main()
