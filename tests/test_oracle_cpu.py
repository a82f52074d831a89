"""CPU tests: the oracle against the reference-generated golden vectors and known-answer constants,
the C-ABI library (loads, exports every declared symbol, rejects bad arguments without a GPU) and the
host-side mirror of the reference interface."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import ddpm, port
from tests.util import actor_params_for, critic_params_for, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ schedule (SURVEY.md 8c constants)
def test_schedule_known_answers_T5():
    s = ddpm.DDPMSchedulerRestated(num_train_timesteps=5)
    np.testing.assert_allclose(s.betas.numpy(), [0.10129408, 0.27954385, 0.47363535, 0.72405237, 0.99900001],
                               rtol=2e-7)
    np.testing.assert_allclose(s.alphas_cumprod.numpy(),
                               [0.89870590, 0.64747816, 0.34080964, 0.09404562, 9.4044401e-05], rtol=3e-6)
    c = ddpm.ddpm_step_constants(5).numpy()
    table = {4: (0.99995297, 103.11776733, 0.30639073, 0.02865130, 0.95138508),
             3: (0.95181632, 3.26084924, 0.46657300, 0.38222390, 0.72583389),
             2: (0.81190538, 1.71294749, 0.57815701, 0.38798824, 0.50327992),
             1: (0.59373552, 1.24276054, 0.75174886, 0.24389443, 0.28341579),
             0: (0.31826735, 1.05485117, 1.0, 0.0, 0.0)}
    for t, row in table.items():
        np.testing.assert_allclose(c[t], row, rtol=5e-6, atol=1e-7)


@pytest.mark.parametrize("T,b0,bl,al", [(20, 0.00799272, 0.999, 6.0595662e-06), (100, 0.00063128, 0.999, 2.4285407e-07)])
def test_schedule_known_answers_long(T, b0, bl, al):
    s = ddpm.DDPMSchedulerRestated(num_train_timesteps=T)
    assert abs(s.betas[0].item() - b0) < 2e-8 * max(1, b0 / 1e-3)
    assert abs(s.betas[-1].item() - bl) < 1e-7
    assert abs(s.alphas_cumprod[-1].item() / al - 1) < 2e-5
    s.set_timesteps(T)
    assert s.timesteps.tolist() == list(range(T - 1, -1, -1))


def test_posemb_known_answers():
    e = port.sinusoidal_pos_emb(torch.tensor([1.0]), 256)[0]
    np.testing.assert_allclose(e[:3].numpy(), [0.8415, 0.8016, 0.7611], atol=5e-5)
    np.testing.assert_allclose(e[128:131].numpy(), [0.5403, 0.5978, 0.6487], atol=5e-5)


def test_param_counts():
    assert sum(v.numel() for v in port.init_actor_params(0).values()) == 1489928
    assert sum(v.numel() for v in port.init_critic_params(0).values()) == 380518


# ------------------------------------------------------------------ oracle port vs reference-made fixtures
@pytest.mark.parametrize("name", ["h1_T5_B16", "h1_T5_B16_wide", "h1_T20_B8", "h1_T100_B4"])
def test_port_sampler_matches_reference(name):
    g = load_golden(name)
    p = actor_params_for(g)
    T = int(g["T"])
    out = port.actor_sample(p, torch.from_numpy(g["state"]), torch.from_numpy(g["noise"]), T)
    np.testing.assert_allclose(out.numpy(), g["action"], rtol=0, atol=2e-6)
    # the restated diffusers scheduler against the reference's in-tree DDPM on the same chain
    tol = {5: 5e-6, 20: 2e-5, 100: 1e-4}[T]
    np.testing.assert_allclose(out.numpy(), g["action_intree_ddpm"], rtol=0, atol=tol)
    eps = port.actor_eps(p, torch.from_numpy(g["noise"][0]), torch.ones(out.shape[0]) * (T - 1),
                         torch.from_numpy(g["state"]))
    np.testing.assert_allclose(eps.detach().numpy(), g["eps_first"], rtol=0, atol=2e-6)
    assert np.abs(g["action"]).max() <= 1.0


@pytest.mark.parametrize("name", ["h3_T5_B64", "h3_T20_B32_wide"])
def test_port_loss_and_grads_match_reference(name):
    g = load_golden(name)
    p = actor_params_for(g)
    loss, grads = port.actor_loss_and_grads(p, torch.from_numpy(g["state"]), torch.from_numpy(g["action"]),
                                            torch.from_numpy(g["noise"]), torch.from_numpy(g["timesteps"]),
                                            int(g["T"]))
    assert abs(loss.item() - float(g["loss"])) < 1e-6
    total = torch.sqrt(sum((v ** 2).sum() for v in grads.values())).item()
    assert abs(total / float(g["grad_norm"]) - 1) < 1e-5
    for i, k in enumerate(port.ACTOR_KEYS):
        gk = grads[k]
        samp = gk if gk.numel() <= 4096 else gk.flatten()[::997]
        np.testing.assert_allclose(samp.numpy().reshape(-1), g[f"gsample_{i}"].reshape(-1), rtol=1e-4, atol=1e-7)
        assert abs(gk.norm().item() - float(g[f"gnorm_{i}"])) <= 1e-5 * max(1.0, float(g[f"gnorm_{i}"]))


@pytest.mark.parametrize("name", ["h2_B32", "h2_B8_wide_clip"])
def test_port_q_ascent_matches_reference(name):
    g = load_golden(name)
    p = critic_params_for(g)
    obs, act = torch.from_numpy(g["obs"]), torch.from_numpy(g["action"])
    p1, p2 = port.q1_q2(p, obs, act)
    np.testing.assert_allclose(p1.numpy(), g["p1"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(p2.numpy(), g["p2"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(port.q_min(p, obs, act).numpy(), g["q_min"], rtol=1e-6, atol=1e-6)
    a = act.clone().requires_grad_(True)
    port.q_min(p, obs, a).sum().backward()
    np.testing.assert_allclose(a.grad.numpy(), g["dq_da"], rtol=1e-4, atol=1e-6)
    mean_abs, new_a, norms, _ = port.q_action_ascent(p, obs, act.clone(), iters=int(g["iters"]), return_trace=True)
    np.testing.assert_allclose(new_a.numpy(), g["new_action"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(norms.numpy(), g["norms"], rtol=1e-5)
    assert abs(mean_abs - float(g["mean_abs"])) < 1e-6
    if name.endswith("clip"):
        assert g["norms"].max() > 1.0, "fixture must exercise the active-clip branch"


def test_port_adamw_step_matches_torch():
    """oracle.port.adamw_train_step against torch.optim.AdamW + clip_grad_norm_ (ac_base.py:52,83-92)."""
    g = load_golden("h3_T5_B64")
    p = actor_params_for(g)
    args = (torch.from_numpy(g["state"]), torch.from_numpy(g["action"]), torch.from_numpy(g["noise"]),
            torch.from_numpy(g["timesteps"]), int(g["T"]))
    q = {k: torch.nn.Parameter(v.clone()) for k, v in p.items()}
    opt = torch.optim.AdamW([q[k] for k in port.ACTOR_KEYS], 3e-4)
    st, cur = None, p
    for _ in range(2):
        loss = port.actor_loss(q, *args)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([q[k] for k in port.ACTOR_KEYS], 1.0)
        opt.step()
        _, _, cur, st = port.adamw_train_step(cur, *args, opt_state=st)
    for k in port.ACTOR_KEYS:
        np.testing.assert_allclose(cur[k].numpy(), q[k].detach().numpy(), rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------ C ABI without a GPU
def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ddiffpg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ddp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_lib):
    from ddiffpg_b200 import _lib
    handle = ctypes.CDLL(built_lib)
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/ddiffpg_b200.h but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert set(_lib.PROTOTYPES) == set(names)
    assert _lib.lib().ddp_abi_version() == 1


def test_abi_argument_errors_without_gpu(built_lib):
    from ddiffpg_b200 import _lib
    L = _lib.lib()
    shape = _lib.ActorShape(34, 8, 5, 256, 1024, 512, 256)
    n32 = L.ddp_actor_packed_bytes(shape, 0)
    n16 = L.ddp_actor_packed_bytes(shape, 1)
    assert n32 > 4 * (42 * 1024 + 1024 * 512 + 512 * 256 + 256 * 8) and n16 > n32
    assert L.ddp_actor_grad_count(shape) == 1489928
    bad = _lib.ActorShape(34, 8, 500, 256, 1024, 512, 256)          # T above the supported maximum
    assert L.ddp_actor_packed_bytes(bad, 0) == 0
    assert b"T" in L.ddp_last_error()
    assert L.ddp_actor_sample(shape, None, None, None, None, 4, 0, None, 0, None) == -2
    assert L.ddp_actor_sample(shape, None, None, None, None, 0, 0, None, 0, None) == 0      # empty batch
    assert L.ddp_actor_sample(shape, None, None, None, None, -1, 0, None, 0, None) == -1
    q = _lib.QShape(29, 8, 51, 0.0, 5.0, 3, 512, 256, 128)
    assert L.ddp_q_packed_bytes(q, 0) > 3 * 4 * 380518
    qbad = _lib.QShape(29, 8, 51, 5.0, 0.0, 1, 512, 256, 128)
    assert L.ddp_q_packed_bytes(qbad, 0) == 0
    assert L.ddp_q_forward(q, None, None, None, None, None, None, None, None, 0, 0, None, 0, None) == 0
    # round-2 entry points: the critic update and RND on both precisions (sizes only -- no compute without a GPU)
    q1 = _lib.QShape(29, 8, 51, 0.0, 5.0, 1, 512, 256, 128)
    assert L.ddp_q_packed_bytes(q1, 1) > L.ddp_q_packed_bytes(q1, 0)
    w32, w16 = L.ddp_q_critic_train_workspace_bytes(q1, 4096, 0), L.ddp_q_critic_train_workspace_bytes(q1, 4096, 1)
    assert w32 > 0 and w16 > 0 and L.ddp_q_critic_train_workspace_bytes(q1, 4096, 7) == 0
    assert L.ddp_q_critic_loss_fwd_bwd(q, None, None, None, None, None, None, None, None, 0.99, None, None, 8, 1, None, 0,
                                       None) == -1                  # one critic per call
    assert L.ddp_q_critic_loss_fwd_bwd(q1, None, None, None, None, None, None, None, None, 0.99, None, None, 8, 1, None, 0,
                                       None) == -2                  # NULL arguments
    r = _lib.RndShape(69, 128, 512, 256, 128)
    assert L.ddp_rnd_packed_bytes_p(r, 0) == L.ddp_rnd_packed_bytes(r) > 0
    assert L.ddp_rnd_packed_bytes_p(r, 1) > L.ddp_rnd_packed_bytes_p(r, 0)
    assert L.ddp_rnd_workspace_bytes_p(r, 4096, 1) > 0 and L.ddp_rnd_workspace_bytes_p(r, 4096, 0) == L.ddp_rnd_train_workspace_bytes(r, 4096)
    assert L.ddp_rnd_packed_bytes_p(_lib.RndShape(300, 128, 512, 256, 128), 1) == 0       # tensor path: D <= 256
    assert L.ddp_rnd_packed_bytes_p(_lib.RndShape(69, 128, 500, 256, 128), 1) == 0        # ... widths multiples of 64
    assert L.ddp_rnd_packed_bytes_p(_lib.RndShape(69, 128, 500, 256, 128), 0) > 0         # the FMA path takes them
    assert L.ddp_rnd_novelty_p(r, None, None, None, None, None, 4, 1, None, 0, None) == -2
    # the row-sharded ascent (exchange-step callback): argument checks are the plain entry point's
    common = (None, None, None, 20, 0.03, 0.9, 0.999, 1e-5, 1.0, 0.99999, None, None)
    assert L.ddp_q_action_ascent_sharded(q, None, None, *common, 8, 0, None, 0, None, None, None) == -2
    assert L.ddp_q_action_ascent_sharded(q, None, None, *common, 0, 0, None, 0, None, None, None) == -1
    assert L.ddp_q_action_ascent(q, None, None, *common, 8, 0, None, 0, None) == -2


# ------------------------------------------------------------------ host mirror of the reference surface
def test_state_dict_keys_match_reference_names():
    from ddiffpg_b200 import DiffusionPolicy, DistributionalDoubleQ
    pol = DiffusionPolicy(34, 8, 5, device="cpu")
    assert tuple(pol.state_dict().keys()) == port.ACTOR_KEYS
    ref = port.init_actor_params(0)
    assert all(pol.state_dict()[k].shape == ref[k].shape for k in port.ACTOR_KEYS)
    cri = DistributionalDoubleQ([29], 8, v_min=0, v_max=5, num_atoms=51, device="cpu")
    assert tuple(cri.state_dict().keys()) == port.CRITIC_KEYS
    assert cri.z_atoms.shape == (51,) and cri.v_min == 0 and cri.v_max == 5
    pol2 = DiffusionPolicy([29 + 5], 8, 5)       # Sequence state_dim like ac_base.py:29-31 passes
    assert pol2.state_dim == 34


def test_same_seed_gives_reference_init():
    """Construction order equals the reference's, so torch.manual_seed(s) yields the reference's weights."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference not mounted")
    from ddiffpg_b200 import DiffusionPolicy, DistributionalDoubleQ
    R = ref_loader.load_reference()
    torch.manual_seed(5); a = DiffusionPolicy(34, 8, 5, device="cpu")
    torch.manual_seed(5); b = R.DiffusionPolicy(34, 8, 5, device="cpu")
    for k, v in b.state_dict().items():
        assert torch.equal(a.state_dict()[k], v), k
    torch.manual_seed(6); c = DistributionalDoubleQ(29, 8, 0, 5, 51, device="cpu")
    torch.manual_seed(6); d = R.DistributionalDoubleQ(29, 8, 0, 5, 51, device="cpu")
    for k, v in d.state_dict().items():
        assert torch.equal(c.state_dict()[k], v), k


def test_unsupported_paths_fail_loudly():
    from ddiffpg_b200 import DiffusionPolicy
    pol = DiffusionPolicy(34, 8, 5, device="cpu")
    with pytest.raises(NotImplementedError):
        pol(torch.zeros(2, 34), sample=False)
    with pytest.raises(NotImplementedError):
        pol(torch.zeros(2, 34), add_noise=True)
    with pytest.raises(NotImplementedError):
        DiffusionPolicy(34, 8, 5, energy=True)
    with pytest.raises(RuntimeError, match="CUDA only"):       # no CPU fallback
        pol(torch.zeros(2, 34))


def test_port_action_noise_matches_reference_fixture():
    """N3: add_mixed_normal_noise / add_normal_noise (utils/noise.py:19-41) with injected standard normals."""
    g = load_golden("n3_noise")
    a, z = torch.from_numpy(g["a"]), torch.from_numpy(g["z"])
    np.testing.assert_array_equal(port.add_noise_to_actions(a, z, 0.05, 0.8).numpy(), g["mixed"])
    np.testing.assert_array_equal(port.add_noise_to_actions(a, z, 0.3, 0.3).numpy(), g["fixed"])
    np.testing.assert_array_equal(port.add_noise_to_actions(a, z, 0.8, 0.8, noise_bounds=(-0.2, 0.2)).numpy(), g["tgt"])


def _critic_fixture_params(g):
    p, pt = port.init_critic_params(41, scale=1.5), port.init_critic_params(42, scale=1.5)
    from tests.util import checksum
    np.testing.assert_allclose(checksum(p, port.CRITIC_KEYS), g["checksum"], rtol=1e-12)
    np.testing.assert_allclose(checksum(pt, port.CRITIC_KEYS), g["checksum_t"], rtol=1e-12)
    return p, pt


def test_port_critic_update_matches_reference_fixture():
    """N1: C51 projection (utils/distl_util.py:4-20) + BCE critic loss and its gradients (ddiffpg.py:325-349)."""
    g = load_golden("n1_critic")
    p, pt = _critic_fixture_params(g)
    t = lambda k: torch.from_numpy(g[k])
    tq = port.critic_target_dist(pt, t("nobs"), t("nact"), t("reward"), t("done"), float(g["gamma"]))
    np.testing.assert_allclose(tq.numpy(), g["target_q"], rtol=0, atol=1e-7)
    # every projected row is a distribution; the element-wise min of two may lose mass but never gains any
    np.testing.assert_allclose(g["proj1"].sum(1), 1.0, atol=1e-5)
    assert (tq.sum(1) <= 1.0 + 1e-5).all()
    loss, grads = port.critic_loss_and_grads(p, tq, t("obs"), t("act"))
    assert abs(loss.item() - float(g["loss"])) < 1e-6
    for i, k in enumerate(port.CRITIC_KEYS):
        ref = g[f"g_{i}"]
        got = grads[k] if grads[k].numel() <= 8192 else grads[k].flatten()[::97]
        np.testing.assert_allclose(got.numpy().reshape(ref.shape), ref, rtol=1e-5, atol=1e-7, err_msg=k)
        assert abs(float(grads[k].norm()) - float(g[f"gnorm_{i}"])) <= 1e-5 * max(1.0, float(g[f"gnorm_{i}"]))


def test_c51_projection_edge_cases():
    """done rows collapse onto the reward atom; integer positions put all mass on one atom; out-of-range targets
    clamp to the end atoms."""
    atoms = 51
    nd = torch.full((4, atoms), 1.0 / atoms)
    reward = torch.tensor([[1.0], [7.0], [-3.0], [0.25]])
    done = torch.tensor([[1.0], [1.0], [1.0], [1.0]])
    pr = port.c51_projection(nd, reward, done, 0.99)
    assert abs(pr[0, 10].item() - 1.0) < 1e-5                  # 1.0 / 0.1 = atom 10 exactly
    assert abs(pr[1, 50].item() - 1.0) < 1e-5                  # clamped to v_max
    assert abs(pr[2, 0].item() - 1.0) < 1e-5                   # clamped to v_min
    assert abs(pr[3, 2].item() - 0.5) < 1e-4 and abs(pr[3, 3].item() - 0.5) < 1e-4
    np.testing.assert_allclose(pr.sum(1).numpy(), 1.0, atol=1e-5)


def test_port_rnd_matches_reference_fixture():
    """N4: RNDModel / IntrinsicM (models/mlp.py:233-267, utils/intrinsic.py) -- encoding, novelty, NovelD reward,
    mse loss and predictor gradients."""
    from tests.util import checksum
    g = load_golden("n4_rnd")
    p = port.init_rnd_params(71)
    np.testing.assert_allclose(checksum(p, port.RND_KEYS), g["checksum"], rtol=1e-12)
    obs, nobs = torch.from_numpy(g["obs"]), torch.from_numpy(g["nobs"])
    enc = port.encode_obs_antmaze(obs)
    np.testing.assert_array_equal(enc.numpy(), g["enc"])
    assert enc.shape[1] == 69
    nov = port.rnd_novelty(p, enc)
    np.testing.assert_allclose(nov.numpy(), g["novelty"], rtol=1e-6)
    r0 = port.noveld_reward(nov, port.rnd_novelty(p, port.encode_obs_antmaze(nobs)))
    np.testing.assert_allclose(r0.numpy(), g["r0"], rtol=1e-6, atol=1e-9)
    loss, grads = port.rnd_loss_and_grads(p, port.encode_obs_antmaze(torch.cat([obs, nobs])))
    assert abs(loss.item() - float(g["loss"])) <= 1e-6 * float(g["loss"])
    for i, k in enumerate(k for k in port.RND_KEYS if k.startswith("predictor")):
        ref = g[f"g_{i}"]
        got = grads[k] if grads[k].numel() <= 8192 else grads[k].flatten()[::97]
        np.testing.assert_allclose(got.numpy().reshape(ref.shape), ref, rtol=1e-5, atol=1e-9, err_msg=k)


def test_port_replay_matches_reference_fixture():
    """N2: DiffusionReplayBuffer.sample_batch / update_target_action and add_embedding with recorded draws."""
    g = load_golden("n2_replay")
    store = {k[6:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("store_")}
    for gi in range(3):
        idx = torch.from_numpy(g[f"idx_{gi}"])
        data = port.replay_gather(store, idx, gi)
        for name, t in zip(("obs", "action", "target", "reward", "next_obs", "done"), data):
            np.testing.assert_array_equal(t.numpy(), g[f"g{gi}_{name}"], err_msg=f"group {gi} {name}")
    emb = torch.from_numpy(g["emb"])
    np.testing.assert_array_equal(port.add_embedding_port(torch.from_numpy(g["g1_obs"]), emb, g["zero_idx"]).numpy(), g["emb_state"])
    np.testing.assert_array_equal(port.add_embedding_port(torch.from_numpy(g["g0_obs"]), emb, []).numpy(), g["emb_state_p0"])
    port.replay_scatter(store, torch.from_numpy(g["new_action"]), torch.from_numpy(g["idx_2"]), 2)
    np.testing.assert_array_equal(store["target_action"].numpy(), g["target_after"])


def test_port_goal_buffer_matches_reference_fixture():
    """N2, caller side: DiffusionGoalBuffer.sample_batch / add_temp_data (diffusion_replay.py:250-332) with the
    reference's recorded draws -- group split, temp-buffer share of group 0, and the six tensors of every group."""
    g = load_golden("n2_goal_buffer")
    store = {k[6:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("store_")}
    temp = {k[5:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("temp_")}
    plan = port.goal_buffer_plan(int(g["batch"]), g["success_id"].tolist(), g["unsuccess_id"].tolist(),
                                 g["clusters"].tolist(), g["unsuccess_clusters"].tolist(), temp["state"].shape[0], store["id"])
    assert [(p[1], p[2]) for p in plan] == [(g[f"draw_{i}"].shape[0], g[f"tdraw_{i}"].shape[0]) for i in range(3)]
    assert plan[0][2] > 0 and sum(p[1] + p[2] for p in plan) == int(g["batch"])
    for i, (grp, b_sample, b_temp) in enumerate(plan):
        draw = torch.from_numpy(g[f"draw_{i}"]) if b_sample else None
        tdraw = torch.from_numpy(g[f"tdraw_{i}"]) if b_temp else None
        data, rows = port.goal_buffer_group(store, temp, grp, i, draw, tdraw)
        np.testing.assert_array_equal(rows.numpy(), g[f"idx_{i}"])
        for name, t in zip(("obs", "action", "target", "reward", "next_obs", "done"), data):
            np.testing.assert_array_equal(t.numpy(), g[f"g{i}_{name}"], err_msg=f"group {i} {name}")


def test_bench_reference_arm_prints_one_json_line():
    """bench.py contract on the CPU-only arm: exactly one line on stdout, valid JSON, the keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and "workload" in d["config"]


def test_soft_update_is_bit_identical_to_the_reference_expression():
    """ddiffpg_b200.soft_update issues three multi-tensor launches instead of three per parameter; the arithmetic per
    element is the reference's (utils/torch_util.py:9-12): cur * tau + tar * (1 - tau), two products and one sum."""
    import torch.nn as nn
    from ddiffpg_b200 import soft_update
    torch.manual_seed(5)
    cur = nn.Sequential(nn.Linear(37, 64), nn.ELU(), nn.Linear(64, 51))
    tar = nn.Sequential(nn.Linear(37, 64), nn.ELU(), nn.Linear(64, 51))
    want = [c.data * 0.05 + t.data * (1.0 - 0.05) for t, c in zip(tar.parameters(), cur.parameters())]
    soft_update(tar, cur, 0.05)
    for got, w in zip(tar.parameters(), want):
        assert torch.equal(got.data, w)
