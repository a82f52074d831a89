"""world_size-2 gloo tests (CPU) of the data-parallel host logic: row sharding, mode-segment sharding and the
H3 exchange step (sum all-reduce of gradients computed with the GLOBAL 1/(B*A) equals the full-batch gradient).
The per-shard gradients come from the oracle port, standing in for the CUDA kernel (same contract:
loss partial sums and gradients scaled by inv_count)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ddiffpg_b200 import dist as ddist


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 256, 65536, 1000003):
        for world in (1, 2, 3, 8):
            spans = [ddist.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_segments():
    seg = [0, 37, 37, 42, 172]
    got = [ddist.shard_segments(seg, 2, r) for r in range(2)]
    for m in range(4):
        (lo0, hi0), (lo1, hi1) = got[0][0][m], got[1][0][m]
        assert lo0 == seg[m] and hi0 == lo1 and hi1 == seg[m + 1]
    assert got[0][2] == [37, 0, 5, 130] and got[0][1][-1] + got[1][1][-1] == 172


def _worker(rank, world, port_no, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import port
    torch.set_num_threads(2)
    T, B, h = 5, 50, 256
    gen = torch.Generator().manual_seed(5)
    p = port.init_actor_params(3, h=h)
    state, action = torch.randn(B, 34, generator=gen), torch.rand(B, 8, generator=gen) * 2 - 1
    noise, ts = torch.randn(B, 8, generator=gen), torch.randint(0, T, (B,), generator=gen)
    lo, hi = ddist.shard_bounds(B, world, rank)
    inv = ddist.global_inv_count(hi - lo, 8)
    assert abs(inv - 1.0 / (B * 8)) < 1e-15
    # per-shard loss/grads with the global normaliser == (local mean) * local_rows / B
    l_loc, g_loc = port.actor_loss_and_grads(p, state[lo:hi], action[lo:hi], noise[lo:hi], ts[lo:hi], T)
    scale = (hi - lo) / B
    flat = torch.cat([g_loc[k].reshape(-1) for k in port.ACTOR_KEYS]) * scale
    loss = (l_loc * scale).reshape(1).clone()
    ddist.allreduce_sum_(flat, loss)
    l_ref, g_ref = port.actor_loss_and_grads(p, state, action, noise, ts, T)
    ref = torch.cat([g_ref[k].reshape(-1) for k in port.ACTOR_KEYS])
    ok = bool(torch.allclose(flat, ref, rtol=1e-4, atol=1e-7) and abs(loss.item() - l_ref.item()) < 1e-6)
    open(os.path.join(tmp, f"ok{rank}"), "w").write(str(ok))
    dist.destroy_process_group()


def test_gradient_allreduce_equals_full_batch_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, 29641, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"ok{r}").read() == "True"


def _ascent_worker(rank, world, port_no, tmp):
    """The contract of ddp_q_action_ascent_sharded (SURVEY 8e, H2 semantics (ii)) restated with torch on each rank's
    shard: gradient of -sum(q_min) / B_global, sum g^2 all-reduced, the reference's clip coefficient and Adam step --
    must equal the oracle port's ascent of the gathered batch (here with the clip active)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import port
    torch.set_num_threads(2)
    gen = torch.Generator().manual_seed(17)
    B, iters, lr, eps, max_norm, lim = 23, 6, 0.03, 1e-5, 0.01, 1 - 1e-5
    p = port.init_critic_params(8, scale=2.0)
    obs, act = torch.randn(B, 29, generator=gen), torch.rand(B, 8, generator=gen) * 2 - 1
    _, a_ref, n_ref, _ = port.q_action_ascent(p, obs, act.clone(), iters=iters, lr=lr, eps=eps, max_norm=max_norm,
                                              return_trace=True)
    (lo, hi), = ddist.shard_segments([0, B], world, rank)[0]
    a = act[lo:hi].clone().clamp_(-lim, lim)
    opt = torch.optim.Adam([a], lr=lr, eps=eps)
    norms = []
    for _ in range(iters):
        a.requires_grad_(True)
        loss = -port.q_min({k: v.detach() for k, v in p.items()}, obs[lo:hi], a, 0.0, 5.0).sum() / B      # global 1/B
        opt.zero_grad(set_to_none=True)
        loss.backward()
        gsq = (a.grad.double() ** 2).sum().reshape(1)
        dist.all_reduce(gsq)                                                                        # the exchange step
        norm = gsq.sqrt().float()
        a.grad.mul_(torch.clamp(max_norm / (norm + 1e-6), max=1.0))                                 # clip_grad_norm_
        norms.append(norm)
        opt.step()
        a.requires_grad_(False)
        a.clamp_(-lim, lim)
    ok = bool(torch.allclose(a, a_ref[lo:hi], atol=2e-6) and torch.allclose(torch.cat(norms), n_ref, rtol=1e-5)
              and n_ref.min().item() > max_norm)
    open(os.path.join(tmp, f"ok{rank}"), "w").write(str(ok))
    dist.destroy_process_group()


def test_sharded_ascent_with_norm_exchange_equals_gathered_batch_gloo(tmp_path):
    world = 2
    mp.spawn(_ascent_worker, args=(world, 29643, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"ok{r}").read() == "True"


def test_host_batch_ranges_cover_the_batch_in_whole_waves():
    """Ranges of get_actions_host: contiguous cover of [0, B), one range below 8k rows per range; below three sampler
    waves (sms x 128 rows) the partial wave first and whole waves behind it, never more than `chunks` ranges; from three
    waves on exactly one wave (the only exposed upload) followed by the rest in one launch."""
    from ddiffpg_b200.models import host_batch_ranges
    sms = 148
    wave = sms * 128
    for B in (0, 1, 1000, 16383, 16384, 18944, 20000, 30000, 37888, 56831, 56832, 65536, 75776, 1 << 20):
        for chunks in (1, 2, 3, 4, 8):
            r = host_batch_ranges(B, chunks, sms)
            assert len(r) <= max(chunks, 1)
            assert sum(hi - lo for lo, hi in r) == B
            assert all(r[i][1] == r[i + 1][0] for i in range(len(r) - 1)) and (not r or (r[0][0] == 0 and r[-1][1] == B))
            if len(r) > 1 and B >= 3 * wave:
                assert r == [(0, wave), (wave, B)]
            elif len(r) > 1:
                assert all((hi - lo) % wave == 0 for lo, hi in r[1:])      # everything behind the first range: whole waves
                first = r[0][1] - r[0][0]
                assert first < wave or first % wave == 0                      # the partial wave, if any, comes first
    assert host_batch_ranges(65536, 4, sms) == [(0, wave), (wave, 65536)]
    assert host_batch_ranges(50000, 4, sms) == [(0, 50000 - 2 * wave), (50000 - 2 * wave, 50000 - wave), (50000 - wave, 50000)]


def test_trainer_bucket_plan_covers_the_gradient_buffer():
    """FusedActorTrainer._bucket_plan: every plan covers [0, n + 1) (gradient + the loss slot) exactly once, each bucket
    waits for the event of the LAST group it contains, and the default is one collective per gradient group."""
    from ddiffpg_b200.algo import FusedActorTrainer

    class Stub:
        _offsets = [0, 10, 14, 30, 34, 100, 110, 200, 208, 240, 244, 250, 252]
        _group_slices = FusedActorTrainer._group_slices
        _bucket_plan = FusedActorTrainer._bucket_plan

        def __init__(self, buckets, world):
            self.buckets, self._world = buckets, world

        def world_size(self):
            return self._world
    n = 252
    for buckets, world, want in ((4, 2, 4), (2, 2, 2), (1, 2, 1), (None, 2, 4), (None, 8, 4), (None, 1, 4)):
        plan = Stub(buckets, world)._bucket_plan(n)
        assert len(plan) == want
        cover = sorted((lo, hi) for _, lo, hi in plan)
        assert cover[0][0] == 0 and cover[-1][1] == n + 1 and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
        assert [g for g, _, _ in plan] == sorted(g for g, _, _ in plan) and plan[-1][0] == 3
        groups = Stub(4, 2)._group_slices(n)
        for g, lo, hi in plan:          # the bucket's event is the one of the last-finished group inside it
            inside = [k for k, (a, b) in enumerate(groups) if a >= lo and b <= hi]
            assert g == max(inside)


def test_available_indices_cache_follows_the_storage():
    """ReplayKernels.available_indices (pure index logic, no kernel): equals torch.where(torch.isin(...)) of
    simple_replay.py:151, is cached per id set while the storage is unchanged, and recomputed once buf_id is re-allocated
    (what the reference's add_to_buffer / remove do)."""
    import torch
    from ddiffpg_b200.replay import ReplayKernels

    class Store(ReplayKernels):
        pass
    s = Store()
    s.buf_id = torch.tensor([[0.], [1.], [1.], [2.], [5.], [7.]])
    for ids in ([1, 5], [0], [9], [7, 2, 0], []):
        ref = torch.where(torch.isin(s.buf_id, torch.tensor(ids, dtype=s.buf_id.dtype)))[0] if ids else torch.empty(0, dtype=torch.int64)
        assert torch.equal(s.available_indices(ids), ref)
        assert s.get_buffer_size(ids) == ref.shape[0]
    first = s.available_indices([1, 5])
    assert s.available_indices([1, 5]) is first
    s.buf_id = torch.cat([s.buf_id, torch.tensor([[5.]])])
    assert s.available_indices([1, 5]).tolist() == [1, 2, 4, 6]
