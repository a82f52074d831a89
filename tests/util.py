"""Shared helpers for the parity tests (oracle = checker only)."""
import os

import numpy as np
import torch

from oracle import port

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def checksum(p, keys):
    return np.array([[float(p[k].double().sum()), float(p[k].double().abs().sum())] for k in keys])


def actor_params_for(g):
    p = port.init_actor_params(int(g["wseed"]), scale=float(g["scale"]))
    np.testing.assert_allclose(checksum(p, port.ACTOR_KEYS), g["checksum"], rtol=1e-12,
                               err_msg="seeded weight generator drifted from the one that made the fixtures")
    return p


def critic_params_for(g):
    p = port.init_critic_params(int(g["wseed"]), scale=float(g["scale"]))
    np.testing.assert_allclose(checksum(p, port.CRITIC_KEYS), g["checksum"], rtol=1e-12)
    return p


def make_policy(params, T, precision="fp32", hidden=(1024, 512, 256), S=34, A=8):
    from ddiffpg_b200 import DiffusionPolicy
    pol = DiffusionPolicy(S, A, T, device="cuda", hidden=hidden, precision=precision)
    pol.load_state_dict(params)
    return pol.to("cuda")


def make_critic(params, O=29, A=8, hidden=None):
    from ddiffpg_b200 import DistributionalDoubleQ
    c = DistributionalDoubleQ(O, A, v_min=0, v_max=5, num_atoms=51, device="cuda", hidden_layers=hidden)
    c.load_state_dict(params)
    return c.to("cuda")


def assert_close(actual, expected, rtol, atol, what=""):
    actual = torch.as_tensor(actual).detach().cpu().double()
    expected = torch.as_tensor(expected).detach().cpu().double()
    diff = (actual - expected).abs()
    bound = atol + rtol * expected.abs()
    bad = diff > bound
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.numel()} elements out of tolerance "
                           f"(rtol={rtol}, atol={atol}); max abs diff {diff.max().item():.3e}, "
                           f"worst ratio {(diff / bound).max().item():.2f}")


def assert_ascent_close(actual, expected, gaps, iters, lr, what="", rtol=1e-4, atol=5e-5, ridge=2e-5):
    """Compare Q-ascent results row by row.

    The ascent maximises min(Q1, Q2); rows are attracted to the ridge Q1 == Q2, where the arg-min -- and
    with it the whole action gradient -- flips on the last float bit (the fp32 reference itself is
    discontinuous there).  Rows whose reference gap |Q1-Q2| ever drops below ``ridge`` (a few fp32 ulps of
    Q ~ 2.5) are therefore only
    required to stay within the distance Adam can open after the first possible flip (lr per iteration);
    every other row must match to (rtol, atol)."""
    actual = torch.as_tensor(actual).detach().cpu().double()
    expected = torch.as_tensor(expected).detach().cpu().double()
    gaps = torch.as_tensor(gaps).detach().cpu().abs()                    # [iters, B]
    near = gaps < ridge
    on_ridge = near.any(0)
    assert_close(actual[~on_ridge], expected[~on_ridge], rtol, atol, what + " (smooth rows)")
    if on_ridge.any():
        first = torch.where(near, torch.arange(iters)[:, None], iters).min(0).values[on_ridge]
        bound = (iters - first).double()[:, None] * lr * 1.05 + atol
        diff = (actual[on_ridge] - expected[on_ridge]).abs()
        assert (diff <= bound).all(), f"{what}: ridge rows moved further than Adam allows ({diff.max():.3e})"
    return int(on_ridge.sum())
