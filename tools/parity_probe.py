"""Measure (not assert) the bf16-path error metrics the parity tests bound, so the bounds in tests/ are set from data.
Usage on the GPU box: python tools/parity_probe.py [h1] [h3] [h2]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import port                                         # noqa: E402
from tests.util import make_critic, make_policy                 # noqa: E402

what = set(sys.argv[1:]) or {"h1", "h3", "h2"}
dev = "cuda"


def h1():
    for h in (1024, 512, 256):
        for T in (5, 20, 100):
            B, n = 65536, 4096
            gen = torch.Generator().manual_seed(1000 + h + T)
            p = port.init_actor_params(83, h=h)
            state = torch.randn(B, 34, generator=gen)
            noise = torch.randn(T, B, 8, generator=gen)
            pol = make_policy(p, T, precision="bf16", hidden=(h, h // 2, h // 4))
            out = pol.get_actions(state.to(dev), noise=noise.to(dev)).cpu()
            sl = slice(30000, 30000 + n)
            t0 = time.time()
            ref = port.actor_sample(p, state[sl], noise[:, sl], T)
            err = (out[sl] - ref).abs().flatten()
            q = torch.quantile(err, torch.tensor([0.5, 0.99, 0.9999]))
            print(f"H1 bf16 h={h} T={T}: max {err.max():.3e} mean {err.mean():.3e} p50 {q[0]:.2e} p99 {q[1]:.2e} "
                  f"p99.99 {q[2]:.2e} n>1e-2 {(err > 1e-2).sum().item()}/{err.numel()}  (oracle {time.time() - t0:.1f}s)",
                  flush=True)


def h3():
    for B, T in ((64, 5), (700, 5), (4096, 5), (1000, 20), (4096, 100)):
        gen = torch.Generator().manual_seed(900 + B)
        p = port.init_actor_params(84)
        state = torch.randn(B, 34, generator=gen)
        action = torch.rand(B, 8, generator=gen) * 2 - 1
        noise = torch.randn(B, 8, generator=gen)
        ts = torch.randint(0, T, (B,), generator=gen)
        l_ref, g_ref = port.actor_loss_and_grads(p, state, action, noise, ts, T)
        pol = make_policy(p, T)
        pol.train_precision = "bf16"
        loss = pol.get_loss(state.to(dev), action.to(dev), noise=noise.to(dev), timesteps=ts.to(dev))
        loss.backward()
        got = torch.cat([q.grad.reshape(-1) for _, q in pol.named_parameters()]).cpu()
        ref = torch.cat([g_ref[k].reshape(-1) for k in port.ACTOR_KEYS])
        rel = ((got - ref).norm() / ref.norm()).item()
        print(f"H3 bf16 B={B} T={T}: loss rel {abs(loss.item() - l_ref.item()) / l_ref.item():.2e} flat rel-L2 {rel:.3e}")
        for k, q in pol.named_parameters():
            r = g_ref[k]
            e = ((q.grad.cpu() - r).norm() / r.norm().clamp_min(1e-12)).item()
            print(f"    {k:28s} rel-L2 {e:.3e}  |ref| {r.norm().item():.3e}")
        sys.stdout.flush()


def h2():
    from ddiffpg_b200 import q_action_ascent_segments
    for B in (64, 2000, 16384):
        gen = torch.Generator().manual_seed(800 + B)
        p = port.init_critic_params(92, scale=2.0)
        obs, act = torch.randn(B, 29, generator=gen), torch.rand(B, 8, generator=gen) * 2 - 1
        m_ref, a_ref, norms, gaps, grads = port.q_action_ascent(p, obs, act.clone(), iters=20, return_trace=True,
                                                                return_grads=True)
        work = act.to(dev).clone()
        q_action_ascent_segments([make_critic(p)], obs.to(dev), work, [0, B], iters=20, precision="bf16")
        d = (work.cpu() - a_ref).abs()
        gmin = grads.abs().min(0).values                         # [B, A] smallest |g| an element saw
        grms = grads.pow(2).mean().sqrt().item()
        mingap = gaps.abs().min(0).values                        # [B]
        print(f"H2 bf16 ascent B={B}: g rms {grms:.3e} (Adam eps 1e-5)  all elements: max {d.max():.3e} mean {d.mean():.3e}")
        for gap_thr in (1e-2, 3e-2):
            rows = mingap > gap_thr
            for rel_thr in (0.0, 0.05, 0.1, 0.25, 0.5, 1.0):
                ok = rows[:, None] & (gmin > rel_thr * grms)
                if ok.any():
                    e = d[ok]
                    print(f"    gap>{gap_thr:.0e} |g|>{rel_thr:4.2f} rms: kept {ok.float().mean().item():6.1%}  max {e.max():.3e} "
                          f"mean {e.mean():.3e} frac<=1e-2 {(e <= 1e-2).float().mean().item():.4f}")
        sys.stdout.flush()


if "h1" in what:
    h1()
if "h3" in what:
    h3()
if "h2" in what:
    h2()
