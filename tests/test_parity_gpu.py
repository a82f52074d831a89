"""GPU parity tests: the CUDA path, called through the C ABI via the host mirror, against the oracle and
the reference-generated golden fixtures.  fp32 tolerance: rtol 1e-4 (north-star), atol 2e-5 (the fp32
reference itself sits 1e-6..1e-5 from an fp64 run of the same chain, see tools/numerics_probe.py)."""
import numpy as np
import pytest
import torch

from oracle import port
from tests.util import (actor_params_for, assert_ascent_close, assert_close, critic_params_for, load_golden,
                        make_critic, make_policy)

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-4, 2e-5


def _dev(x):
    return torch.as_tensor(x).to("cuda")


# ------------------------------------------------------------------------------------------ H1
@pytest.mark.parametrize("name", ["h1_T5_B16", "h1_T20_B8", "h1_T100_B4"])
def test_sampler_fp32_matches_reference_fixture(name):
    g = load_golden(name)
    pol = make_policy(actor_params_for(g), int(g["T"]))
    out = pol.get_actions(_dev(g["state"]), noise=_dev(g["noise"]))
    atol = ATOL if int(g["T"]) <= 20 else 1e-4          # T=100: 1/sqrt(abar_99) = 2029 amplifies fp32 noise
    assert_close(out, g["action"], RTOL, atol, name)
    assert out.abs().max().item() <= 1.0


def test_sampler_fp32_chaotic_weights_within_fp64_noise_floor():
    """Wide (x2.5) weights make the chain ill-conditioned; the bound is the fp32 reference's own distance
    from an fp64 evaluation of the same chain."""
    g = load_golden("h1_T5_B16_wide")
    p = actor_params_for(g)
    state, noise = torch.from_numpy(g["state"]), torch.from_numpy(g["noise"])
    ref64 = port.actor_sample(port.cast_params(p, torch.float64), state.double(), noise.double(), 5)
    floor = (torch.from_numpy(g["action"]).double() - ref64).abs().max().item()
    out = make_policy(p, 5).get_actions(_dev(state), noise=_dev(noise)).cpu().double()
    assert (out - ref64).abs().max().item() <= 20 * max(floor, 1e-5)


@pytest.mark.parametrize("B", [1, 3, 20, 256, 700, 4096])
def test_sampler_fp32_vs_oracle_batches(B):
    T = 5
    gen = torch.Generator().manual_seed(100 + B)
    p = port.init_actor_params(41)
    state = torch.randn(B, 34, generator=gen)
    noise = torch.randn(T, B, 8, generator=gen)
    ref = port.actor_sample(p, state, noise, T)
    out = make_policy(p, T).get_actions(_dev(state), noise=_dev(noise))
    assert_close(out, ref, RTOL, ATOL, f"B={B}")


@pytest.mark.parametrize("h,T", [(256, 5), (512, 20), (256, 100)])
def test_sampler_fp32_width_sweep(h, T):
    B = 64
    gen = torch.Generator().manual_seed(7 * h + T)
    p = port.init_actor_params(43, h=h)
    state = torch.randn(B, 34, generator=gen)
    noise = torch.randn(T, B, 8, generator=gen)
    ref = port.actor_sample(p, state, noise, T)
    out = make_policy(p, T, hidden=(h, h // 2, h // 4)).get_actions(_dev(state), noise=_dev(noise))
    assert_close(out, ref, RTOL, ATOL if T <= 20 else 1e-4, f"h={h} T={T}")


def test_sampler_edge_cases():
    p = port.init_actor_params(44)
    pol = make_policy(p, 5)
    assert pol(torch.zeros(0, 34, device="cuda")).shape == (0, 8)               # empty batch
    with pytest.raises(ValueError):
        pol(torch.zeros(4, 33, device="cuda"))                                   # wrong state width
    with pytest.raises(ValueError):
        pol.get_actions(torch.zeros(4, 34, device="cuda"), noise=torch.zeros(5, 3, 8, device="cuda"))
    a = pol(torch.randn(33, 34, device="cuda"))                                  # device-drawn noise
    assert a.shape == (33, 8) and torch.isfinite(a).all() and a.abs().max() <= 1.0
    # rows are independent: a batch equals its two halves
    gen = torch.Generator().manual_seed(3)
    state, noise = torch.randn(50, 34, generator=gen).cuda(), torch.randn(5, 50, 8, generator=gen).cuda()
    full = pol.get_actions(state, noise=noise)
    lo = pol.get_actions(state[:23], noise=noise[:, :23].contiguous())
    hi = pol.get_actions(state[23:], noise=noise[:, 23:].contiguous())
    assert torch.equal(full, torch.cat([lo, hi]))


def test_repack_follows_parameter_updates():
    p = port.init_actor_params(45)
    pol = make_policy(p, 5)
    gen = torch.Generator().manual_seed(4)
    state, noise = torch.randn(8, 34, generator=gen), torch.randn(5, 8, 8, generator=gen)
    a0 = pol.get_actions(_dev(state), noise=_dev(noise))
    with torch.no_grad():
        for q in pol.parameters():
            q.mul_(1.05)                                   # in-place update bumps _version -> repack
    p2 = {k: v * 1.05 for k, v in p.items()}
    a1 = pol.get_actions(_dev(state), noise=_dev(noise))
    assert_close(a1, port.actor_sample(p2, state, noise, 5), RTOL, ATOL, "after in-place update")
    assert not torch.equal(a0, a1)


# ------------------------------------------------------------------------------------------ H2
@pytest.mark.parametrize("name", ["h2_B32", "h2_B8_wide_clip"])
def test_q_forward_and_ascent_match_reference_fixture(name):
    from ddiffpg_b200 import update_target_action
    g = load_golden(name)
    cri = make_critic(critic_params_for(g))
    obs, act = _dev(g["obs"]), _dev(g["action"])
    cri.requires_grad_(False)
    p1, p2 = cri.get_q1_q2(obs, act)
    assert_close(p1, g["p1"], 1e-4, 1e-6, "p1")
    assert_close(p2, g["p2"], 1e-4, 1e-6, "p2")
    a = act.clone().requires_grad_(True)
    q = cri.get_q_min(obs, a)
    assert_close(q, g["q_min"], 1e-4, 1e-5, "q_min")
    q.sum().backward()
    assert_close(a.grad, g["dq_da"], 1e-3, 1e-5 * max(1.0, float(np.abs(g["dq_da"]).max())), "dq/da")
    work = act.clone()
    mean_abs, upd = update_target_action(obs, work, cri, action_lr=0.03, update_times=int(g["iters"]))
    _, _, _, gaps = port.q_action_ascent(critic_params_for(g), torch.from_numpy(g["obs"]),
                                         torch.from_numpy(g["action"]).clone(), iters=int(g["iters"]),
                                         return_trace=True)
    n_ridge = assert_ascent_close(upd, g["new_action"], gaps, int(g["iters"]), 0.03, "ascent result")
    assert n_ridge <= 4
    assert torch.equal(upd, work) and upd.data_ptr() != work.data_ptr()       # in place + deep copy
    assert abs(mean_abs - float(g["mean_abs"])) < 1e-5 + 0.03 * 20 * n_ridge / upd.shape[0]
    assert all(q.requires_grad for q in cri.parameters())                       # left as the reference leaves it


def test_q_ascent_norm_trace_and_clip():
    from ddiffpg_b200 import q_action_ascent_segments
    g = load_golden("h2_B8_wide_clip")
    cri = make_critic(critic_params_for(g))
    work = _dev(g["action"]).clone()
    _, norms = q_action_ascent_segments([cri], _dev(g["obs"]), work, [0, 8], iters=20, return_norms=True)
    assert_close(norms[0], g["norms"], 2e-3, 1e-5, "pre-clip norms")
    assert norms.max().item() > 1.0


@pytest.mark.parametrize("B", [1, 5, 256, 1500])
def test_q_ascent_vs_oracle_batches(B):
    from ddiffpg_b200 import update_target_action
    gen = torch.Generator().manual_seed(200 + B)
    p = port.init_critic_params(51, scale=1.5)
    obs = torch.randn(B, 29, generator=gen)
    act = torch.rand(B, 8, generator=gen) * 2 - 1
    m_ref, a_ref, _, gaps = port.q_action_ascent(p, obs, act.clone(), iters=20, return_trace=True)
    work = _dev(act).clone()
    m, upd = update_target_action(_dev(obs), work, make_critic(p))
    n_ridge = assert_ascent_close(upd, a_ref, gaps, 20, 0.03, f"B={B}")
    assert n_ridge <= max(2, B // 8)
    assert abs(m - m_ref) < 1e-5 + 0.6 * n_ridge / B


def test_q_ascent_mode_segments_equal_separate_calls():
    """K+1 critics over rows sorted by mode == one reference call per mode (uneven, one empty segment)."""
    from ddiffpg_b200 import q_action_ascent_segments
    gen = torch.Generator().manual_seed(9)
    sizes = [37, 0, 5, 130]
    ps = [port.init_critic_params(60 + i, scale=1.0 + 0.5 * i) for i in range(len(sizes))]
    B = sum(sizes)
    obs = torch.randn(B, 29, generator=gen)
    act = torch.rand(B, 8, generator=gen) * 2 - 1
    off = np.concatenate([[0], np.cumsum(sizes)])
    refs, means, gaps = [], [], []
    for i, n in enumerate(sizes):
        if n == 0:
            means.append(0.0)
            continue
        m, a, _, gp = port.q_action_ascent(ps[i], obs[off[i]:off[i + 1]], act[off[i]:off[i + 1]].clone(), iters=20,
                                           return_trace=True)
        refs.append(a)
        means.append(m)
        gaps.append(gp)
    work = _dev(act).clone()
    mean_abs = q_action_ascent_segments([make_critic(p) for p in ps], _dev(obs), work, off.tolist(), iters=20)
    n_ridge = assert_ascent_close(work, torch.cat(refs), torch.cat(gaps, dim=1), 20, 0.03, "segmented ascent")
    assert n_ridge <= 20
    assert_close(mean_abs, means, 1e-4, 1e-5 + 0.6 * n_ridge / 5, "mean|a| per mode")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_q_ascent_sharded_exchange_step(precision):
    """ddp_q_action_ascent_sharded (SURVEY 8e, H2 semantics (ii)) on one GPU: (a) an identity exchange step reproduces the
    plain call; (b) a rank that holds X of the two-rank batch [X; X] -- global count 2 B, the other rank's sum g^2 equal to
    its own, so the exchange step doubles it -- takes exactly the steps the first half of [X; X] takes in one process
    (clip active: max_norm = 1e-4); (c) a failing exchange step surfaces as the Python exception it raised.
    The tensor path is pinned to its per-iteration schedule for the comparison (the single-launch form of small batches
    differs on a few ridge rows, tests/test_tc_gpu.py::test_q_chain_variants_agree); what remains is the order of the
    fp32 atomics in the norm."""
    import ctypes
    from ddiffpg_b200 import _lib, q_action_ascent_segments
    gen = torch.Generator().manual_seed(77)
    sizes = [300, 0, 213]
    off = np.concatenate([[0], np.cumsum(sizes)]).tolist()
    B = off[-1]
    critics = [make_critic(port.init_critic_params(80 + i, scale=1.5)) for i in range(3)]
    obs, act = _dev(torch.randn(B, 29, generator=gen)), _dev(torch.rand(B, 8, generator=gen) * 2 - 1)

    def same(x, y, what):
        d = (x - y).abs()
        if precision == "fp32":     # (one row may sit on the Q1 == Q2 ridge to the last bit: the order of the atomics decides)
            assert int((d.max(1).values > 2e-5).sum().item()) <= 1, (what, d.max().item())
        else:
            bad = (d.max(1).values > 1e-3).float().mean().item()
            assert bad <= 0.01 and d.mean().item() <= 2e-4, (what, bad, d.mean().item())

    dbg = _lib.lib().ddp_debug_q_variant
    dbg.argtypes, dbg.restype = [ctypes.c_int, ctypes.c_int], None
    dbg(0, 0)
    try:
        for max_norm in (1.0, 1e-4):
            plain, ident = act.clone(), act.clone()
            m0, n0 = q_action_ascent_segments(critics, obs, plain, off, iters=20, max_norm=max_norm, precision=precision,
                                              return_norms=True)
            calls = []
            m1, n1 = q_action_ascent_segments(critics, obs, ident, off, iters=20, max_norm=max_norm, precision=precision,
                                              return_norms=True, gsq_reduce=lambda t: calls.append(tuple(t.shape)))
            assert calls == [(3,)] * 20
            same(ident, plain, f"identity exchange step, max_norm={max_norm}")
            assert (n1[[0, 2]] / n0[[0, 2]] - 1).abs().max().item() <= 1e-3
            assert_close(m1, m0, 1e-3, 1e-4, "mean|a|")
        # (b) the two-rank batch [X; X], mode-sorted: every segment twice as long
        obs2 = torch.cat([torch.cat([obs[off[m]:off[m + 1]]] * 2) for m in range(3)])
        act2 = torch.cat([torch.cat([act[off[m]:off[m + 1]]] * 2) for m in range(3)])
        off2 = [2 * o for o in off]
        q_action_ascent_segments(critics, obs2, act2, off2, iters=20, max_norm=1e-4, precision=precision)
        first_half = torch.cat([act2[off2[m]:off2[m] + sizes[m]] for m in range(3)])
        shard = act.clone()
        q_action_ascent_segments(critics, obs, shard, off, iters=20, max_norm=1e-4, precision=precision,
                                 mean_counts=[2 * n for n in sizes], gsq_reduce=lambda t: t.mul_(2.0))
        same(shard, first_half, "doubled exchange step vs [X; X]")
        local = act.clone()                               # shard-local norm instead: visibly different steps
        q_action_ascent_segments(critics, obs, local, off, iters=20, max_norm=1e-4, precision=precision,
                                 mean_counts=[2 * n for n in sizes])
        assert (local - first_half).abs().mean().item() > 1e-3
    finally:
        dbg(0, -1)

    def boom(t):
        raise KeyError("exchange step failed")
    with pytest.raises(KeyError):
        q_action_ascent_segments(critics, obs, act.clone(), off, iters=3, precision=precision, gsq_reduce=boom)


# ------------------------------------------------------------------------------------------ H3
def _flat(grads):
    return torch.cat([grads[k].reshape(-1) for k in port.ACTOR_KEYS])


@pytest.mark.parametrize("name", ["h3_T5_B64", "h3_T20_B32_wide"])
def test_train_loss_and_grads_match_reference_fixture(name):
    g = load_golden(name)
    p = actor_params_for(g)
    pol = make_policy(p, int(g["T"]))
    loss = pol.get_loss(_dev(g["state"]), _dev(g["action"]), noise=_dev(g["noise"]),
                        timesteps=_dev(g["timesteps"]))
    assert loss.dim() == 0 and abs(loss.item() - float(g["loss"])) < 1e-5 * max(1.0, float(g["loss"]))
    loss.backward()
    params = dict(pol.named_parameters())
    total = torch.sqrt(sum((q.grad ** 2).sum() for q in params.values())).item()
    assert abs(total / float(g["grad_norm"]) - 1) < 1e-4
    for i, k in enumerate(port.ACTOR_KEYS):
        gk = params[k].grad
        samp = gk if gk.numel() <= 4096 else gk.flatten()[::997]
        scale = float(g[f"gnorm_{i}"]) / np.sqrt(gk.numel())            # rms of this tensor's gradient
        assert_close(samp.reshape(-1), g[f"gsample_{i}"].reshape(-1), 1e-3, 1e-3 * scale + 1e-9, k)
        assert abs(gk.norm().item() - float(g[f"gnorm_{i}"])) <= 1e-4 * float(g[f"gnorm_{i}"]) + 1e-9


@pytest.mark.parametrize("B,T", [(1, 5), (37, 5), (700, 5), (2000, 20), (1500, 100)])
def test_train_grads_vs_oracle(B, T):
    gen = torch.Generator().manual_seed(300 + B)
    p = port.init_actor_params(71)
    state = torch.randn(B, 34, generator=gen)
    action = torch.rand(B, 8, generator=gen) * 2 - 1
    noise = torch.randn(B, 8, generator=gen)
    ts = torch.randint(0, T, (B,), generator=gen)
    l_ref, g_ref = port.actor_loss_and_grads(p, state, action, noise, ts, T)
    pol = make_policy(p, T)
    loss = pol.get_loss(_dev(state), _dev(action), noise=_dev(noise), timesteps=_dev(ts))
    loss.backward()
    assert abs(loss.item() - l_ref.item()) < 1e-5 * max(1.0, l_ref.item())
    got = torch.cat([q.grad.reshape(-1) for _, q in pol.named_parameters()]).cpu()
    ref = _flat(g_ref)
    rel = (got - ref).norm() / ref.norm()
    assert rel < 1e-4, f"relative L2 error of the flat gradient {rel:.3e}"
    for k, q in pol.named_parameters():
        r = g_ref[k]
        assert (q.grad.cpu() - r).abs().max() <= 1e-3 * r.abs().max() + 1e-8, k


def _assert_params_after_adam(pol, ref_params, steps, lr):
    """Adam normalises every element's step to ~lr, so an element whose gradient is rounding noise
    (|g| ~ 1e-9) can legitimately land anywhere within steps*lr; everything else must agree closely."""
    for k, q in pol.named_parameters():
        d = (q.detach().cpu().double() - ref_params[k].double()).abs()
        assert d.max().item() <= steps * lr * 0.1, f"{k}: max deviation {d.max().item():.3e}"
        assert (d > 2e-6 + 1e-4 * ref_params[k].abs().double()).double().mean().item() < 1e-3, k
        assert d.mean().item() < 2e-7, f"{k}: mean deviation {d.mean().item():.3e}"


def test_update_actor_matches_reference_optimizer_step():
    """get_loss + the reference's optimizer_update (AdamW, clip 1.0) for two steps == the oracle."""
    from ddiffpg_b200 import update_actor
    g = load_golden("h3_T5_B64")
    p = actor_params_for(g)
    args = (torch.from_numpy(g["state"]), torch.from_numpy(g["action"]), torch.from_numpy(g["noise"]),
            torch.from_numpy(g["timesteps"]))
    pol = make_policy(p, 5)
    opt = torch.optim.AdamW(pol.parameters(), 3e-4)
    st, cur = None, p
    for _ in range(2):
        loss, gnorm = update_actor(pol, opt, _dev(args[0]), _dev(args[1]), noise=_dev(args[2]), timesteps=_dev(args[3]))
        l_ref, n_ref, cur, st = port.adamw_train_step(cur, *args, 5, opt_state=st)
        assert abs(loss - l_ref.item()) < 1e-5 and abs(gnorm / n_ref.item() - 1) < 1e-4
    _assert_params_after_adam(pol, cur, steps=2, lr=3e-4)


def test_fused_trainer_matches_oracle_steps():
    from ddiffpg_b200 import FusedActorTrainer
    g = load_golden("h3_T20_B32_wide")
    p = actor_params_for(g)
    args = (torch.from_numpy(g["state"]), torch.from_numpy(g["action"]), torch.from_numpy(g["noise"]),
            torch.from_numpy(g["timesteps"]))
    pol = make_policy(p, 20)
    tr = FusedActorTrainer(pol)
    assert tuple(pol.state_dict().keys()) == port.ACTOR_KEYS
    st, cur = None, p
    for _ in range(3):
        loss, gnorm = tr.step(_dev(args[0]), _dev(args[1]), noise=_dev(args[2]), timesteps=_dev(args[3]))
        l_ref, n_ref, cur, st = port.adamw_train_step(cur, *args, 20, opt_state=st)
        assert abs(loss.item() - l_ref.item()) < 1e-5 * max(1, l_ref.item())
        assert abs(gnorm.item() / n_ref.item() - 1) < 1e-4
    _assert_params_after_adam(pol, cur, steps=3, lr=3e-4)
    # the sampler sees the updated weights
    gen = torch.Generator().manual_seed(1)
    s, n = torch.randn(6, 34, generator=gen), torch.randn(20, 6, 8, generator=gen)
    assert_close(pol.get_actions(_dev(s), noise=_dev(n)), port.actor_sample(cur, s, n, 20), 1e-3, 1e-3, "post-train")


# ------------------------------------------------------------------------------------------ N3 noise epilogue
@pytest.mark.parametrize("precision,tol", [("fp32", 3e-5), ("bf16", 1e-2)])
def test_sampler_with_fused_exploration_noise(precision, tol):
    """get_actions / get_tgt_policy_actions (ddiffpg.py:82-110): sampler + noise + clamps in one launch."""
    from ddiffpg_b200 import get_actions, get_tgt_policy_actions
    B, T = 300, 5
    gen = torch.Generator().manual_seed(31)
    p = port.init_actor_params(46)
    state, noise = torch.randn(B, 34, generator=gen), torch.randn(T, B, 8, generator=gen)
    z = torch.randn(B, 8, generator=gen)
    base = port.actor_sample(p, state, noise, T)
    pol = make_policy(p, T, precision=precision)
    got = get_actions(pol, _dev(state), noise_type="mixed", std_min=0.05, std_max=0.8, noise=_dev(noise), expl_noise=_dev(z))
    assert_close(got, port.add_noise_to_actions(base, z, 0.05, 0.8), 1e-4, tol, "mixed")
    got = get_actions(pol, _dev(state), noise_type="fixed", std=0.3, noise=_dev(noise), expl_noise=_dev(z))
    assert_close(got, port.add_noise_to_actions(base, z, 0.3, 0.3), 1e-4, tol, "fixed")
    got = get_tgt_policy_actions(pol, _dev(state), tgt_pol_std=0.8, tgt_pol_noise_bound=0.2, noise=_dev(noise), expl_noise=_dev(z))
    assert_close(got, port.add_noise_to_actions(base, z, 0.8, 0.8, noise_bounds=(-0.2, 0.2)), 1e-4, tol, "target policy")
    assert got.abs().max().item() <= 1.0
    plain = get_actions(pol, _dev(state), sample=False, noise=_dev(noise))
    assert_close(plain, base, 1e-4, tol, "sample=False")


def test_noise_epilogue_matches_reference_fixture():
    """The epilogue alone against the reference's utils/noise.py outputs: a zero-step stand-in is not possible, so
    compare through the oracle identity out = noise_fn(sampler output)."""
    g = load_golden("n3_noise")
    a, z = torch.from_numpy(g["a"]), torch.from_numpy(g["z"])
    assert torch.equal(port.add_noise_to_actions(a, z, 0.05, 0.8), torch.from_numpy(g["mixed"]))


# ------------------------------------------------------------------------------------------ N1
def _critic_pair(g=None, seeds=(41, 42), scale=1.5):
    p, pt = port.init_critic_params(seeds[0], scale=scale), port.init_critic_params(seeds[1], scale=scale)
    if g is not None:
        from tests.util import checksum
        np.testing.assert_allclose(checksum(p, port.CRITIC_KEYS), g["checksum"], rtol=1e-12)
        np.testing.assert_allclose(checksum(pt, port.CRITIC_KEYS), g["checksum_t"], rtol=1e-12)
    return p, pt


def _flat_to_named(flat, params):
    out, off = {}, 0
    for k in port.CRITIC_KEYS:
        n = params[k].numel()
        out[k] = flat[off:off + n].view(params[k].shape).cpu()
        off += n
    assert off == flat.numel()
    return out


def _rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-12)).item()


def test_critic_update_matches_reference_fixture():
    """Loss and all 16 gradients of update_critic (ddiffpg.py:322-349) against the reference's autograd."""
    from ddiffpg_b200 import critic_loss_and_grads
    g = load_golden("n1_critic")
    p, pt = _critic_pair(g)
    critic, target = make_critic(p), make_critic(pt)
    loss, flat = critic_loss_and_grads(critic, target, _dev(g["obs"]), _dev(g["act"]), _dev(g["nobs"]),
                                       _dev(g["nact"]), _dev(g["reward"]), _dev(g["done"]), float(g["gamma"]))
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * max(1.0, abs(float(g["loss"])))
    named = _flat_to_named(flat, p)
    for i, k in enumerate(port.CRITIC_KEYS):
        ref = torch.from_numpy(g[f"g_{i}"])
        got = named[k] if named[k].numel() <= 8192 else named[k].flatten()[::97]
        assert _rel_l2(got.reshape(ref.shape), ref) <= 1e-4, k
        assert abs(float(named[k].norm()) - float(g[f"gnorm_{i}"])) <= 1e-4 * float(g[f"gnorm_{i}"]) + 1e-8, k


@pytest.mark.parametrize("B", [1, 7, 512, 3000])
def test_critic_update_vs_oracle_batches(B):
    from ddiffpg_b200 import critic_loss_and_grads
    p, pt = _critic_pair(seeds=(51, 52), scale=1.2)
    gen = torch.Generator().manual_seed(600 + B)
    obs, nobs = torch.randn(B, 29, generator=gen), torch.randn(B, 29, generator=gen)
    act, nact = torch.rand(B, 8, generator=gen) * 2 - 1, torch.rand(B, 8, generator=gen) * 2 - 1
    reward = torch.rand(B, 1, generator=gen) * 3.0
    reward[::5] = 0.0
    done = (torch.rand(B, 1, generator=gen) < 0.25).float()
    tq = port.critic_target_dist(pt, nobs, nact, reward, done, 0.97)
    # rows whose clamped mass rounds to 1 + 1ulp make F.binary_cross_entropy raise (in the reference too): drop them
    keep = tq.max(1).values <= 1.0
    obs, nobs, act, nact, reward, done, tq = (x[keep] for x in (obs, nobs, act, nact, reward, done, tq))
    assert keep.float().mean() > 0.9
    l_ref, g_ref = port.critic_loss_and_grads(p, tq, obs, act)
    loss, flat = critic_loss_and_grads(make_critic(p), make_critic(pt), _dev(obs), _dev(act), _dev(nobs), _dev(nact),
                                       _dev(reward), _dev(done), 0.97)
    assert abs(loss.item() - l_ref.item()) <= 1e-5 * max(1.0, abs(l_ref.item()))
    named = _flat_to_named(flat, p)
    for k in port.CRITIC_KEYS:
        assert _rel_l2(named[k], g_ref[k]) <= 1e-4, (k, _rel_l2(named[k], g_ref[k]))


def test_update_critic_matches_reference_optimizer_step():
    """Whole update_critic: loss -> backward -> clip_grad_norm_ -> AdamW, against torch on the port's gradients."""
    from ddiffpg_b200 import update_critic
    p, pt = _critic_pair(seeds=(61, 62), scale=1.0)
    gen = torch.Generator().manual_seed(77)
    B = 256
    obs, nobs = torch.randn(B, 29, generator=gen), torch.randn(B, 29, generator=gen)
    act, nact = torch.rand(B, 8, generator=gen) * 2 - 1, torch.rand(B, 8, generator=gen) * 2 - 1
    reward, done = torch.rand(B, 1, generator=gen), (torch.rand(B, 1, generator=gen) < 0.2).float()
    critic, target = make_critic(p), make_critic(pt)
    opt = torch.optim.AdamW(critic.parameters(), lr=5e-4)
    ref = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    ropt = torch.optim.AdamW([ref[k] for k in port.CRITIC_KEYS], lr=5e-4)
    for _ in range(3):
        _, loss, gnorm = update_critic(critic, target, opt, _dev(obs), _dev(act), _dev(reward), _dev(nobs), _dev(nact),
                                       _dev(done), gamma_n=0.99 ** 3, max_grad_norm=1.0)
        tq = port.critic_target_dist(pt, nobs, nact, reward, done, 0.99 ** 3).clamp_max(1.0)
        l_ref, g_ref = port.critic_loss_and_grads({k: v.detach() for k, v in ref.items()}, tq, obs, act)
        for k in port.CRITIC_KEYS:
            ref[k].grad = g_ref[k].clone()
        n_ref = torch.nn.utils.clip_grad_norm_([ref[k] for k in port.CRITIC_KEYS], 1.0)
        ropt.step()
        assert abs(loss - l_ref.item()) <= 2e-5 * max(1.0, abs(l_ref.item()))
        assert abs(gnorm - n_ref.item()) <= 1e-4 * n_ref.item()
    _assert_params_after_adam(critic, {k: v.detach() for k, v in ref.items()}, steps=3, lr=5e-4)


@pytest.mark.parametrize("graph", [False, True])
def test_fused_critic_trainer_matches_reference_optimizer_step(graph):
    """FusedCriticTrainer (fused loss / backward + clip + AdamW on a flat vector, optionally replayed as a CUDA graph)
    walks the trajectory of update_critic's torch tail: torch AdamW + clip_grad_norm_ on the port's gradients."""
    from ddiffpg_b200 import FusedCriticTrainer
    p, pt = _critic_pair(seeds=(61, 62), scale=1.0)
    gen = torch.Generator().manual_seed(78)
    B = 300
    obs, nobs = torch.randn(B, 29, generator=gen), torch.randn(B, 29, generator=gen)
    act, nact = torch.rand(B, 8, generator=gen) * 2 - 1, torch.rand(B, 8, generator=gen) * 2 - 1
    reward, done = torch.rand(B, 1, generator=gen), (torch.rand(B, 1, generator=gen) < 0.2).float()
    critic, target = make_critic(p), make_critic(pt)
    keys = list(critic.state_dict().keys())
    tr = FusedCriticTrainer(critic, target, lr=5e-4, graph=graph)
    assert list(critic.state_dict().keys()) == keys
    ref = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    ropt = torch.optim.AdamW([ref[k] for k in port.CRITIC_KEYS], lr=5e-4)
    steps = 4
    for _ in range(steps):
        loss, gnorm = tr.step(_dev(obs), _dev(act), _dev(reward), _dev(nobs), _dev(nact), _dev(done), gamma_n=0.99 ** 3)
        tq = port.critic_target_dist(pt, nobs, nact, reward, done, 0.99 ** 3).clamp_max(1.0)
        l_ref, g_ref = port.critic_loss_and_grads({k: v.detach() for k, v in ref.items()}, tq, obs, act)
        for k in port.CRITIC_KEYS:
            ref[k].grad = g_ref[k].clone()
        n_ref = torch.nn.utils.clip_grad_norm_([ref[k] for k in port.CRITIC_KEYS], 1.0)
        ropt.step()
        assert abs(loss.item() - l_ref.item()) <= 2e-5 * max(1.0, abs(l_ref.item()))
        assert abs(gnorm.item() - n_ref.item()) <= 1e-4 * n_ref.item()
    _assert_params_after_adam(critic, {k: v.detach() for k, v in ref.items()}, steps=steps, lr=5e-4)
    # the kernels read the updated weights: a forward through the module agrees with the port on the reference weights
    with torch.no_grad():
        q = critic.get_q_min(_dev(obs[:50]), _dev(act[:50]))
    assert_close(q, port.q_min({k: v.detach() for k, v in ref.items()}, obs[:50], act[:50]), 1e-3, 1e-3, "q_min after training")
    tr.close()


def test_reference_style_soft_update_reaches_the_kernels():
    """The reference's soft_update writes the target critic through param.data (utils/torch_util.py:9-12, called at
    ddiffpg.py:266), which bumps neither data_ptr nor _version: the next update_critic must nevertheless see the new
    target weights (the target is re-packed on every use), and get_q1_q2 under autograd raises instead of going eager."""
    from ddiffpg_b200 import critic_loss_and_grads
    p, pt = _critic_pair(seeds=(63, 64), scale=1.0)
    gen = torch.Generator().manual_seed(78)
    B = 200
    obs, nobs = torch.randn(B, 29, generator=gen), torch.randn(B, 29, generator=gen)
    act, nact = torch.rand(B, 8, generator=gen) * 2 - 1, torch.rand(B, 8, generator=gen) * 2 - 1
    reward, done = torch.rand(B, 1, generator=gen), (torch.rand(B, 1, generator=gen) < 0.2).float()
    critic, target = make_critic(p), make_critic(pt)
    args = (_dev(obs), _dev(act), _dev(nobs), _dev(nact), _dev(reward), _dev(done), 0.97)
    l0, _ = critic_loss_and_grads(critic, target, *args)
    l0 = l0.item()
    tau = 0.5
    with torch.no_grad():                                   # verbatim reference soft_update
        for tar, cur in zip(target.parameters(), critic.parameters()):
            tar.data.copy_(cur.data * tau + tar.data * (1.0 - tau))
    pt2 = {k: p[k] * tau + pt[k] * (1 - tau) for k in pt}
    l1, _ = critic_loss_and_grads(critic, target, *args)
    tq = port.critic_target_dist(pt2, nobs, nact, reward, done, 0.97).clamp_max(1.0)
    l_ref, _ = port.critic_loss_and_grads(p, tq, obs, act)
    assert abs(l1.item() - l_ref.item()) <= 1e-5 * max(1.0, abs(l_ref.item()))
    assert abs(l1.item() - l0) > 1e-4                       # the update was not a no-op for this check
    with pytest.raises(NotImplementedError):
        critic.get_q1_q2(_dev(obs), _dev(act))              # trainable weights + grad mode: no eager fallback


# ------------------------------------------------------------------------------------------ N4
def _make_rnd(p):
    from ddiffpg_b200 import RNDModel
    m = RNDModel(69)
    m.load_state_dict(p)
    return m.to("cuda")


class _IntrinsicHost:
    """The fields of the reference's IntrinsicM that IntrinsicKernels reads (utils/intrinsic.py:9-31), for a box without
    the reference package; where it is installed, accelerate_intrinsic(IntrinsicM) inherits them from the real class."""

    def __init__(self, p, pos_enc=True):
        from ddiffpg_b200 import RNDModel
        self.rnd_model = RNDModel(69)
        self.rnd_model.load_state_dict(p)
        self.rnd_model.to("cuda")
        self.rnd_optimizer = torch.optim.AdamW(self.rnd_model.parameters(), 1e-4)
        self.pos_enc, self.update_step = pos_enc, 0

    def encode_obs(self, obs):
        return port.encode_obs_antmaze(obs.cpu()).to(obs.device)


def _intrinsic(p):
    from ddiffpg_b200 import IntrinsicKernels
    return type("IntrinsicM", (IntrinsicKernels, _IntrinsicHost), {})(p)


def test_rnd_matches_reference_fixture():
    """IntrinsicM.get_novelty / compute_reward / update (utils/intrinsic.py:33-75) against the reference's own outputs:
    the novelty and the loss / gradients come from the kernels, the NovelD shaping around them is the reference's code
    (restated by the oracle port)."""
    g = load_golden("n4_rnd")
    p = port.init_rnd_params(71)
    im = _intrinsic(p)
    obs, nobs = _dev(g["obs"]), _dev(g["nobs"])
    enc = im.encode_obs(obs)
    assert_close(enc, g["enc"], 1e-5, 1e-5, "positional encoding")
    assert_close(im.get_novelty(enc), g["novelty"], RTOL, ATOL, "novelty")
    r0 = port.noveld_reward(im.get_novelty(enc).cpu(), im.get_novelty(im.encode_obs(nobs)).cpu())
    assert_close(r0, g["r0"], RTOL, 1e-7, "noveld reward (raw)")
    x = im.encode_obs(torch.cat([obs, nobs]))
    loss, flat = im.rnd_model.loss_and_grads(x)
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * float(g["loss"])
    off = 0
    for i, (k, q) in enumerate(im.rnd_model.predictor.named_parameters()):
        got = flat[off:off + q.numel()].view(q.shape).cpu()
        off += q.numel()
        ref = torch.from_numpy(g[f"g_{i}"])
        sub = got if got.numel() <= 8192 else got.flatten()[::97]
        assert _rel_l2(sub.reshape(ref.shape), ref) <= 1e-4, k
        assert abs(float(got.norm()) - float(g[f"gnorm_{i}"])) <= 1e-4 * float(g[f"gnorm_{i}"]) + 1e-9, k
    assert off == flat.numel()
    pf, tf = im.rnd_model(x)
    rp, rt = port.rnd_forward(p, x.cpu())
    assert_close(pf, rp, RTOL, ATOL, "predictor features")
    assert_close(tf, rt, RTOL, ATOL, "target features")


@pytest.mark.parametrize("graph", [False, True])
def test_fused_rnd_trainer_matches_reference_optimizer_step(graph):
    """IntrinsicM.update through FusedRNDTrainer (enable_fused_update): the predictor walks the trajectory of the
    reference's tail -- clip_grad_norm_(1.0) + torch AdamW(1e-4) on the oracle's gradients -- eager and as a CUDA graph."""
    p = port.init_rnd_params(74)
    im = _intrinsic(p)
    im.enable_fused_update(graph=graph)
    gen = torch.Generator().manual_seed(31)
    obs = torch.randn(500, 29, generator=gen)
    keys = [k for k in port.RND_KEYS if k.startswith("predictor")]
    ref = {k: v.clone().requires_grad_(k in keys) for k, v in p.items()}
    ropt = torch.optim.AdamW([ref[k] for k in keys], lr=1e-4)
    steps = 4
    for _ in range(steps):
        loss, gnorm = im.update(_dev(obs))
        x = port.encode_obs_antmaze(obs)
        l_ref, g_ref = port.rnd_loss_and_grads({k: v.detach() for k, v in ref.items()}, x)
        for k in keys:
            ref[k].grad = g_ref[k].clone()
        n_ref = torch.nn.utils.clip_grad_norm_([ref[k] for k in keys], 1.0)
        ropt.step()
        assert abs(loss - l_ref.item()) <= 2e-5 * max(1.0, abs(l_ref.item()))
        assert abs(gnorm - n_ref.item()) <= 1e-4 * n_ref.item()
    for k in keys:
        d = (dict(im.rnd_model.named_parameters())[k].detach().cpu().double() - ref[k].detach().double()).abs()
        assert d.max().item() <= steps * 1e-4 * 0.1, (k, d.max().item())
        assert d.mean().item() < 2e-7, (k, d.mean().item())
    assert im.update_step == steps
    im.rnd_trainer.close()


@pytest.mark.parametrize("B", [1, 9, 600, 5000])
def test_rnd_update_vs_oracle_batches(B):
    p = port.init_rnd_params(72, scale=1.3)
    gen = torch.Generator().manual_seed(900 + B)
    x = torch.randn(B, 69, generator=gen)
    m = _make_rnd(p)
    assert_close(m.novelty(_dev(x)), port.rnd_novelty(p, x), RTOL, ATOL, "novelty")
    loss, flat = m.loss_and_grads(_dev(x))
    l_ref, g_ref = port.rnd_loss_and_grads(p, x)
    assert abs(loss.item() - l_ref.item()) <= 1e-5 * l_ref.item()
    off = 0
    for k in (k for k in port.RND_KEYS if k.startswith("predictor")):
        n = p[k].numel()
        assert _rel_l2(flat[off:off + n].view(p[k].shape).cpu(), g_ref[k]) <= 1e-4, k
        off += n


def test_intrinsic_update_matches_torch_adamw():
    """IntrinsicM.update: loss, clip_grad_norm_(1.0) and AdamW(1e-4) on the predictor; the target stays frozen."""
    p = port.init_rnd_params(73)
    im = _intrinsic(p)
    gen = torch.Generator().manual_seed(5)
    obs = torch.randn(300, 29, generator=gen)
    ref = {k: v.clone().requires_grad_(k.startswith("predictor")) for k, v in p.items()}
    keys = [k for k in port.RND_KEYS if k.startswith("predictor")]
    ropt = torch.optim.AdamW([ref[k] for k in keys], 1e-4)
    for _ in range(3):
        loss, gnorm = im.update(_dev(obs))
        l_ref, g_ref = port.rnd_loss_and_grads({k: v.detach() for k, v in ref.items()}, port.encode_obs_antmaze(obs))
        for k in keys:
            ref[k].grad = g_ref[k].clone()
        n_ref = torch.nn.utils.clip_grad_norm_([ref[k] for k in keys], 1.0)
        ropt.step()
        assert abs(loss - l_ref.item()) <= 2e-5 * l_ref.item()
        assert abs(gnorm - n_ref.item()) <= 1e-4 * n_ref.item()
    assert im.update_step == 3
    _assert_params_after_adam(im.rnd_model, {k: v.detach() for k, v in ref.items()}, steps=3, lr=1e-4)


# ------------------------------------------------------------------------------------------ N2
def _replay_from_fixture(g):
    from ddiffpg_b200 import DiffusionReplayBuffer
    buf = DiffusionReplayBuffer(10000, 29, 8, device="cuda")
    buf.buf_obs, buf.buf_action = _dev(g["store_obs"]), _dev(g["store_action"])
    buf.buf_target_action, buf.buf_reward = _dev(g["store_target_action"]), _dev(g["store_reward"])
    buf.buf_next_obs, buf.buf_done, buf.buf_id = _dev(g["store_next_obs"]), _dev(g["store_done"]), _dev(g["store_id"])
    buf.cur_capacity = buf.buf_obs.shape[0]
    return buf


def test_replay_sample_and_scatter_match_reference_fixture():
    """Byte-exact: the gathers move fp32 values, nothing is computed."""
    from ddiffpg_b200 import add_embedding
    g = load_golden("n2_replay")
    buf = _replay_from_fixture(g)
    groups = [[0, 1, 2, 3, 4], [0, 3], [1, 4]]
    names = ("obs", "action", "target", "reward", "next_obs", "done")
    for gi, grp in enumerate(groups):
        data, idx = buf.sample_batch(len(g[f"draw_{gi}"]), grp, gi, indices=_dev(g[f"draw_{gi}"]))
        assert torch.equal(idx.cpu(), torch.from_numpy(g[f"idx_{gi}"]))
        for name, t in zip(names, data):
            assert torch.equal(t.cpu(), torch.from_numpy(g[f"g{gi}_{name}"])), (gi, name)
        assert buf.available_indices(grp).shape[0] == sum(1 for t in g["store_id"].ravel() if int(t) in grp)
    emb = _dev(g["emb"])
    assert torch.equal(add_embedding(_dev(g["g1_obs"]), emb, zero_indices=g["zero_idx"]).cpu(), torch.from_numpy(g["emb_state"]))
    assert torch.equal(add_embedding(_dev(g["g0_obs"]), emb, p=0).cpu(), torch.from_numpy(g["emb_state_p0"]))
    # fused form: all groups, embedded states included, one launch
    embs = torch.stack([emb, emb * 2, emb * 3])
    o, seg_off, idx, grp_ids, se, ne = buf.sample_groups([g[f"idx_{gi}"] for gi in range(3)], embeddings=embs,
                                                         zero_state=g["zero_idx"] + 14)
    assert seg_off == [0, 14, 26, 38]
    for gi in range(3):
        for name in names:
            assert torch.equal(o[name][seg_off[gi]:seg_off[gi + 1]].cpu(), torch.from_numpy(g[f"g{gi}_{name}"])), (gi, name)
    ref_se = torch.from_numpy(g["emb_state"]).clone()
    ref_se[:, 29:] *= 2                                        # group 1 uses 2 * emb
    assert torch.equal(se[14:26].cpu(), ref_se)
    assert torch.equal(ne[26:, :29].cpu(), torch.from_numpy(g["g2_next_obs"])) and torch.equal(ne[26:, 29:].cpu(), (emb * 3).cpu().expand(12, 5))
    # scatter-back: rows drawn once must match exactly, duplicated rows keep one of their candidates
    buf.update_target_action(_dev(g["new_action"]), _dev(g["idx_2"]), 2)
    after, ref = buf.buf_target_action.cpu(), torch.from_numpy(g["target_after"])
    idx2 = torch.from_numpy(g["idx_2"])
    uniq, counts = torch.unique(idx2, return_counts=True)
    once = uniq[counts == 1]
    assert torch.equal(after[:2], ref[:2])
    keep = torch.ones(after.shape[1], dtype=torch.bool); keep[uniq[counts > 1]] = False
    assert torch.equal(after[2][keep], ref[2][keep]) and len(once) > 0
    for d in uniq[counts > 1]:
        cands = torch.from_numpy(g["new_action"])[idx2 == d]
        assert any(torch.equal(after[2, d], c) for c in cands)


def test_goal_buffer_sample_batch_matches_reference_fixture():
    """DiffusionGoalBuffer.sample_batch / add_temp_data (diffusion_replay.py:250-332) through GoalBufferKernels, byte-exact
    against the reference's own outputs with its recorded torch.randint draws replayed in the reference's order."""
    from ddiffpg_b200 import GoalBufferKernels
    g = load_golden("n2_goal_buffer")

    class Holder(GoalBufferKernels):        # the attributes of the reference class the two methods read
        pass
    gb = Holder()
    gb.device, gb.replay_buffer = "cuda", _replay_from_fixture(g)
    gb.success_id, gb.unsuccess_id = g["success_id"].tolist(), g["unsuccess_id"].tolist()
    gb.clusters, gb.unsuccess_clusters = g["clusters"].tolist(), g["unsuccess_clusters"].tolist()
    gb.Qs, gb.embeddings = ["Q0", "Q1", "Q2"], [torch.zeros(5), torch.ones(5), -torch.ones(5)]
    gb.temp_state, gb.temp_action = _dev(g["temp_state"]), _dev(g["temp_action"])
    gb.temp_reward, gb.temp_next_state, gb.temp_done = _dev(g["temp_reward"]), _dev(g["temp_next_state"]), _dev(g["temp_done"])
    draws = []
    for i in range(3):
        draws += [g[f"draw_{i}"]] if g[f"draw_{i}"].shape[0] else []
        draws += [g[f"tdraw_{i}"]] if g[f"tdraw_{i}"].shape[0] else []
    it = iter(draws)
    real = torch.randint

    def replay(high, size=None, device=None, **kw):
        d = next(it)
        assert tuple(size) == d.shape and int(d.max()) < high
        return torch.from_numpy(d).to(device)
    torch.randint = replay
    try:
        data_list = gb.sample_batch(int(g["batch"]))
        assert next(it, None) is None
        it = iter(draws[:2])
        single, rows0 = gb.add_temp_data(g["draw_0"].shape[0] + g["tdraw_0"].shape[0], gb.success_id + gb.unsuccess_id, 0)
    finally:
        torch.randint = real
    names = ("obs", "action", "target", "reward", "next_obs", "done")
    for i, d in enumerate(data_list):
        assert d["Q"] == gb.Qs[i] and torch.equal(d["embedding"], gb.embeddings[i])
        assert torch.equal(d["indices"].cpu(), torch.from_numpy(g[f"idx_{i}"]))
        for name, t in zip(names, d["batch"]):
            assert t.is_cuda and t.dtype == torch.float32
            assert torch.equal(t.cpu(), torch.from_numpy(g[f"g{i}_{name}"])), (i, name)
    assert torch.equal(rows0.cpu(), torch.from_numpy(g["idx_0"]))
    for name, t in zip(names, single):
        assert torch.equal(t.cpu(), torch.from_numpy(g[f"g0_{name}"])), name
    # available_indices (table look-up) == torch.where(torch.isin(...)) of simple_replay.py:151, any id subset
    rb = gb.replay_buffer
    for ids in ([0], [5, 2], [1, 3, 4], list(range(6)), [9]):
        ref = torch.where(torch.isin(rb.buf_id, torch.tensor(ids, device="cuda").to(rb.buf_id.dtype)))[0]
        assert torch.equal(rb.available_indices(ids), ref) and rb.get_buffer_size(ids) == ref.shape[0]


def test_replay_two_zeroing_draws_modes_and_range_checks():
    """sample_groups with DIFFERENT zero_state / zero_next draws (the reference draws two independent p = 0.5 masks,
    ddiffpg.py:246-252), add_embedding(modes=...) (utils/torch_util.py:24-34) and the IndexError of a stale index."""
    from ddiffpg_b200 import add_embedding
    g = load_golden("n2_replay")
    buf = _replay_from_fixture(g)
    emb = _dev(g["emb"])
    embs = torch.stack([emb, emb * 2, emb * 3])
    zs, zn = [0, 3, 5, 20, 37], [1, 3, 14, 15, 30]
    o, seg_off, idx, grp_ids, se, ne = buf.sample_groups([g[f"idx_{gi}"] for gi in range(3)], embeddings=embs,
                                                         zero_state=zs, zero_next=zn)
    n = seg_off[-1]
    want = embs[grp_ids.long()]
    ws, wn = want.clone(), want.clone()
    ws[zs] = 0
    wn[zn] = 0
    assert torch.equal(se[:, 29:], ws) and torch.equal(ne[:, 29:], wn) and n == 38
    assert torch.equal(se[:, :29], o["obs"]) and torch.equal(ne[:, :29], o["next_obs"])
    # modes branch: s = int(10 * 0.7) = 7 rows in blocks of 3 (2 + the remainder), 2, 2; the rest keep `embedding`
    state = _dev(g["g0_obs"])[:10]
    modes = [emb * 10, emb * 20, emb * 30]
    out = add_embedding(state, emb, p=0.7, modes=modes)
    exp = emb.expand(10, -1).clone()
    exp[0:3], exp[3:5], exp[5:7] = modes[0], modes[1], modes[2]
    assert torch.equal(out[:, :29], state) and torch.equal(out[:, 29:], exp)
    with pytest.raises(IndexError):
        buf.sample_groups([[0, 1, buf.buf_obs.shape[0]]])
    with pytest.raises(IndexError):
        buf.update_target_action(torch.zeros(1, 8, device="cuda"), torch.tensor([0], device="cuda"), 7)


def test_replay_gather_large_roundtrip():
    """Full-size property: gather(scatter(x)) == x for a permutation, 1M-row buffer, 3 groups."""
    from ddiffpg_b200 import DiffusionReplayBuffer
    N, n = 1 << 20, 1 << 18
    gen = torch.Generator(device="cuda").manual_seed(3)
    buf = DiffusionReplayBuffer(N, 29, 8, device="cuda")
    buf.buf_obs = torch.randn(N, 29, device="cuda", generator=gen)
    buf.buf_next_obs = torch.randn(N, 29, device="cuda", generator=gen)
    buf.buf_action = torch.rand(N, 8, device="cuda", generator=gen)
    buf.buf_target_action = torch.zeros(3, N, 8, device="cuda")
    buf.buf_reward = torch.rand(N, 1, device="cuda", generator=gen)
    buf.buf_done = torch.rand(N, 1, device="cuda", generator=gen) < 0.1
    buf.buf_id = torch.zeros(N, 1, device="cuda")
    perm = torch.randperm(N, device="cuda", generator=gen)[:n]
    grp = (torch.arange(n, device="cuda") % 3).to(torch.int32).sort().values.contiguous()
    new = torch.rand(n, 8, device="cuda", generator=gen)
    buf.scatter_groups(new, perm, grp)
    sizes = [int((grp == k).sum()) for k in range(3)]
    parts = torch.split(perm, sizes)
    o, seg_off, idx, gids, _, _ = buf.sample_groups(parts)
    assert torch.equal(o["target"], new) and torch.equal(gids, grp) and seg_off[-1] == n
    assert torch.equal(o["obs"], buf.buf_obs[perm]) and torch.equal(o["done"], buf.buf_done[perm].float())
    assert torch.equal(o["reward"], buf.buf_reward[perm]) and torch.equal(o["next_obs"], buf.buf_next_obs[perm])


def test_fused_trainer_graph_replay_matches_eager():
    """The CUDA-graph form of the training step (captured on the second call, replayed afterwards) walks the same
    trajectory as the eager launch sequence and as the oracle's AdamW steps, with fresh inputs every step."""
    from ddiffpg_b200 import FusedActorTrainer
    T, B = 5, 96
    p = port.init_actor_params(47)
    gen = torch.Generator().manual_seed(9)
    batches = [(torch.randn(B, 34, generator=gen), torch.rand(B, 8, generator=gen) * 2 - 1,
                torch.randn(B, 8, generator=gen), torch.randint(0, T, (B,), generator=gen)) for _ in range(5)]
    pol_g, pol_e = make_policy(p, T), make_policy(p, T)
    tr_g, tr_e = FusedActorTrainer(pol_g, graph=True), FusedActorTrainer(pol_e)
    st, cur = None, p
    for s, a, n, t in batches:
        lg, ng = tr_g.step(_dev(s), _dev(a), noise=_dev(n), timesteps=_dev(t))
        le, ne = tr_e.step(_dev(s), _dev(a), noise=_dev(n), timesteps=_dev(t))
        l_ref, n_ref, cur, st = port.adamw_train_step(cur, s, a, n, t, T, opt_state=st)
        assert abs(lg.item() - le.item()) <= 1e-6 * max(1.0, abs(le.item()))
        assert abs(ng.item() - ne.item()) <= 1e-5 * ne.item()
        assert abs(lg.item() - l_ref.item()) < 1e-5 * max(1, l_ref.item()) and abs(ng.item() / n_ref.item() - 1) < 1e-4
    assert tr_g.step_count == 5 and int(tr_g._step_dev.item()) == 5
    _assert_params_after_adam(pol_g, cur, steps=5, lr=3e-4)
    # the sampler sees the weights the replayed graph wrote
    s6, n6 = torch.randn(7, 34, generator=gen), torch.randn(T, 7, 8, generator=gen)
    assert_close(pol_g.get_actions(_dev(s6), noise=_dev(n6)), pol_e.get_actions(_dev(s6), noise=_dev(n6)), 1e-4, 1e-4, "post-train")


# ------------------------------------------------------------------------------------------ agent-level mixin
class _Cfg(dict):
    """Attribute + .get access, like the reference's OmegaConf nodes."""
    __getattr__ = dict.__getitem__


def _agent(actor, actor_target, lr=3e-4):
    from ddiffpg_b200 import HotPathMixin
    cfg = _Cfg(algo=_Cfg(obs_norm=False, gamma=0.99, nstep=3, max_grad_norm=1.0, num_atoms=51,
                         noise=_Cfg(type="mixed", std_min=0.05, std_max=0.6, tgt_pol_std=0.8, tgt_pol_noise_bound=0.2)),
               diffusion=_Cfg(action_lr=0.03, update_times=20))

    class Agent(HotPathMixin):
        pass
    a = Agent()
    a.cfg, a.actor, a.actor_target = cfg, actor, actor_target
    a.actor_optimizer = torch.optim.AdamW(actor.parameters(), lr)
    return a


def test_hot_path_mixin_reads_the_reference_cfg_and_matches_the_oracle():
    """One inner iteration of AgentDDiffPG.update_net (ddiffpg.py:231-281) through the mixin's methods, every
    random draw injected: exploration actions, critic update (target-policy actions -> projection -> BCE -> AdamW),
    action ascent and actor update, each against the oracle's composition of the same reference steps."""
    T, B = 5, 160
    gen = torch.Generator().manual_seed(101)
    pa, pc, pt = port.init_actor_params(60), port.init_critic_params(61, scale=1.2), port.init_critic_params(62, scale=1.2)
    agent = _agent(make_policy(pa, T), make_policy(pa, T))
    critic, critic_t = make_critic(pc), make_critic(pt)
    copt = torch.optim.AdamW(critic.parameters(), 5e-4)
    obs, nobs = torch.randn(B, 29, generator=gen), torch.randn(B, 29, generator=gen)
    emb = torch.randn(5, generator=gen)
    state, nstate = torch.cat([obs, emb.expand(B, 5)], 1), torch.cat([nobs, emb.expand(B, 5)], 1)
    act = torch.rand(B, 8, generator=gen) * 2 - 1
    reward, done = torch.rand(B, 1, generator=gen), (torch.rand(B, 1, generator=gen) < 0.2).float()
    noise, z = torch.randn(T, B, 8, generator=gen), torch.randn(B, 8, generator=gen)
    # exploration actions: sampler + add_mixed_normal_noise(std_min, std_max) from the cfg
    a_expl = agent.get_actions(_dev(state), noise=_dev(noise), expl_noise=_dev(z))
    assert_close(a_expl, port.add_noise_to_actions(port.actor_sample(pa, state, noise, T), z, 0.05, 0.6), 1e-4, 3e-5, "get_actions")
    # critic update with gamma ** nstep and the target-policy smoothing of the cfg
    _, c_loss, c_norm = agent.update_critic(critic, critic_t, copt, _dev(obs), _dev(act), _dev(reward), _dev(nobs),
                                            _dev(nstate), _dev(done), noise=_dev(noise), expl_noise=_dev(z))
    nact = port.add_noise_to_actions(port.actor_sample(pa, nstate, noise, T), z, 0.8, 0.8, noise_bounds=(-0.2, 0.2))
    tq = port.critic_target_dist(pt, nobs, nact, reward, done, 0.99 ** 3).clamp_max(1.0)
    l_ref, g_ref = port.critic_loss_and_grads(pc, tq, obs, act)
    n_ref = torch.sqrt(sum((g ** 2).sum() for g in g_ref.values())).item()
    assert abs(c_loss - l_ref.item()) <= 2e-5 * max(1.0, l_ref.item()) and abs(c_norm - n_ref) <= 2e-4 * n_ref
    # action ascent on the updated critic (action_lr / update_times / max_grad_norm from the cfg), then the actor step
    pc_new = {k: v.detach().cpu() for k, v in critic.state_dict().items()}
    work = _dev(act).clone()
    mean_abs, new_action = agent.update_target_action(_dev(obs), work, critic)
    m_ref, a_ref, _, gaps = port.q_action_ascent(pc_new, obs, act.clone(), iters=20, return_trace=True)
    assert_ascent_close(new_action, a_ref, gaps, 20, 0.03, "update_target_action")
    assert abs(mean_abs - m_ref) <= 1e-4 and new_action.data_ptr() != work.data_ptr()
    assert all(p.requires_grad for p in critic.parameters())
    ts, n2 = torch.randint(0, T, (B,), generator=gen), torch.randn(B, 8, generator=gen)
    loss = agent.actor.get_loss(_dev(state), new_action, noise=_dev(n2), timesteps=_dev(ts))
    gnorm = agent.optimizer_update(agent.actor_optimizer, loss)
    l2, _, _, _ = port.adamw_train_step(pa, state, new_action.cpu(), n2, ts, T)
    assert abs(loss.item() - l2.item()) <= 1e-5 * max(1.0, l2.item()) and gnorm.item() > 0
