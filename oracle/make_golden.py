"""Generate tests/golden/*.npz by running the REFERENCE's own modules (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference):  python -m oracle.make_golden

Weights are NOT stored (1.49 M floats per actor): they are regenerated from a seed by
``oracle.port.init_*_params`` and loaded into the reference modules with ``load_state_dict``; a
checksum of every tensor is stored so a drift of the seeded generator would be caught.
Noise is injected into the reference by (a) queueing the per-step draws inside the restated
scheduler and (b) replacing ``torch.randn`` for the single initial draw of
``DiffusionPolicy.get_actions`` (diffusion_mlp.py:222-223).
"""
import os
from copy import deepcopy

import numpy as np
import torch

from . import port, ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def param_checksum(p, keys):
    return np.array([[float(p[k].double().sum()), float(p[k].double().abs().sum())] for k in keys])


class _PatchedRandn:
    """Replace torch.randn / torch.randn_like for one call each with queued tensors."""

    def __init__(self, queue):
        self.queue = list(queue)

    def __enter__(self):
        self._randn, self._randn_like = torch.randn, torch.randn_like
        torch.randn = lambda *a, **k: self.queue.pop(0).clone()
        torch.randn_like = lambda *a, **k: self.queue.pop(0).clone()
        return self

    def __exit__(self, *exc):
        torch.randn, torch.randn_like = self._randn, self._randn_like


def ref_actor(R, params, T):
    pol = R.DiffusionPolicy(34, 8, T, device="cpu")
    pol.load_state_dict(params)
    return pol


def ref_sample(R, pol, state, noise):
    pol.noise_scheduler.noise_queue = [n for n in noise[1:]]
    with _PatchedRandn([noise[0]]):
        with torch.no_grad():
            return pol(state)                  # DiffusionPolicy.forward -> get_actions(sample=True)


def intree_ddpm_sample(R, pol, state, noise, T):
    """The same chain through the reference's in-tree DDPM (baseline_models.py:59-185)."""
    class Net(torch.nn.Module):
        def forward(self, x, t, s):
            return pol.net(x, t.float(), s)
    dif = R.Diffusion(state_dim=34, action_dim=8, model=Net(), max_action=1.0,
                      beta_schedule="cosine", n_timesteps=T)
    # x_T via randn, then one randn_like per step; p_sample also draws at t == 0 (masked by 0).
    with _PatchedRandn([n for n in noise] + [torch.zeros_like(noise[0])]):
        with torch.no_grad():
            return dif.sample(state)


def ref_q_ascent(critic, obs, action, iters, lr, max_norm=1.0):
    """update_target_action (ddiffpg.py:358-373) + optimizer_update (ac_base.py:83-92) around the
    reference's real critic.  The agent module cannot be imported (gym), so its statements are
    replayed here one for one."""
    critic.requires_grad_(False)
    lim = 1 - 1e-5
    action.clamp_(-lim, lim)
    opt = torch.optim.Adam([action], lr=lr, eps=1e-5)
    norms = []
    for _ in range(iters):
        action.requires_grad_(True)
        Q = critic.get_q_min(obs, action)
        loss = -Q.mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        norms.append(float(torch.nn.utils.clip_grad_norm_(parameters=opt.param_groups[0]["params"],
                                                          max_norm=max_norm)))
        opt.step()
        action.requires_grad_(False)
        action.clamp_(-lim, lim)
    upd = deepcopy(action.detach())
    critic.requires_grad_(True)
    return torch.abs(action).mean().item(), upd, np.array(norms)


def main():
    R = ref_loader.load_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)

    # ---------------------------------------------------------------- H1 sampler
    for name, T, B, wseed, scale in (("h1_T5_B16", 5, 16, 11, 1.0), ("h1_T5_B16_wide", 5, 16, 12, 2.5),
                                     ("h1_T20_B8", 20, 8, 13, 1.0), ("h1_T100_B4", 100, 4, 14, 1.5)):
        g = torch.Generator().manual_seed(1000 + T + B)
        params = port.init_actor_params(wseed, scale=scale)
        state = torch.randn(B, 34, generator=g)
        noise = torch.randn(T, B, 8, generator=g)
        pol = ref_actor(R, params, T)
        act_ref = ref_sample(R, pol, state, noise)
        act_port = port.actor_sample(params, state, noise, T)
        act_tree = intree_ddpm_sample(R, pol, state, noise, T)
        eps_ref = pol.net(noise[0], torch.ones(B) * (T - 1), state).detach()
        d_port = (act_ref - act_port).abs().max().item()
        d_tree = (act_ref - act_tree).abs().max().item()
        print(f"{name}: |ref-port|={d_port:.3e} |ref-intree_ddpm|={d_tree:.3e} "
              f"sat={(act_ref.abs() >= 1).float().mean().item():.2f}")
        assert d_port < 1e-5 and d_tree < 5e-4, (d_port, d_tree)
        np.savez(os.path.join(OUT, name + ".npz"), T=T, wseed=wseed, scale=scale,
                 state=state.numpy(), noise=noise.numpy(), action=act_ref.numpy(),
                 action_intree_ddpm=act_tree.numpy(), eps_first=eps_ref.numpy(),
                 checksum=param_checksum(params, port.ACTOR_KEYS))

    # ---------------------------------------------------------------- H3 loss + grads
    for name, T, B, wseed, scale in (("h3_T5_B64", 5, 64, 21, 1.0), ("h3_T20_B32_wide", 20, 32, 22, 2.0)):
        g = torch.Generator().manual_seed(2000 + T + B)
        params = port.init_actor_params(wseed, scale=scale)
        state = torch.randn(B, 34, generator=g)
        action = torch.rand(B, 8, generator=g) * 2 - 1
        noise = torch.randn(B, 8, generator=g)
        ts = torch.randint(0, T, (B,), generator=g)
        pol = ref_actor(R, params, T)
        loss = pol.get_loss(state, action, noise=noise, timesteps=ts)
        loss.backward()
        grads = {k: v.grad.detach() for k, v in pol.named_parameters()}
        gn = torch.sqrt(sum((v ** 2).sum() for v in grads.values()))
        l_port, g_port = port.actor_loss_and_grads(params, state, action, noise, ts, T)
        d = max((grads[k] - g_port[k]).abs().max().item() for k in port.ACTOR_KEYS)
        print(f"{name}: loss={loss.item():.6f} |dloss|={abs(loss.item() - l_port.item()):.2e} "
              f"max|dgrad|={d:.2e} gnorm={gn.item():.4f}")
        assert abs(loss.item() - l_port.item()) < 1e-6 and d < 1e-6
        save = dict(T=T, wseed=wseed, scale=scale, state=state.numpy(), action=action.numpy(),
                    noise=noise.numpy(), timesteps=ts.numpy(), loss=loss.item(), grad_norm=gn.item(),
                    checksum=param_checksum(params, port.ACTOR_KEYS))
        for i, k in enumerate(port.ACTOR_KEYS):
            gk = grads[k]
            save[f"gnorm_{i}"] = float(gk.norm())
            save[f"gsum_{i}"] = float(gk.double().sum())
            # biases and the 8-row head in full; a fixed strided sample of the big matrices
            save[f"gsample_{i}"] = (gk if gk.numel() <= 4096 else gk.flatten()[::997]).numpy()
        np.savez(os.path.join(OUT, name + ".npz"), **save)

    # ---------------------------------------------------------------- H2 double-Q + ascent
    for name, B, wseed, scale, iters in (("h2_B32", 32, 31, 1.0, 20), ("h2_B8_wide_clip", 8, 32, 6.0, 20)):
        g = torch.Generator().manual_seed(3000 + B)
        params = port.init_critic_params(wseed, scale=scale)
        obs = torch.randn(B, 29, generator=g)
        act = torch.rand(B, 8, generator=g) * 2.2 - 1.1      # some rows start outside the clamp
        critic = R.DistributionalDoubleQ(29, 8, v_min=0, v_max=5, num_atoms=51, device="cpu")
        critic.load_state_dict(params)
        with torch.no_grad():
            p1, p2 = critic.get_q1_q2(obs, act)
            qm = critic.get_q_min(obs, act)
        a0 = act.clone().requires_grad_(True)
        critic.get_q_min(obs, a0).sum().backward()
        mean_abs, upd, norms = ref_q_ascent(critic, obs, act.clone(), iters, 0.03)
        m2, u2, n2, gaps = port.q_action_ascent(params, obs, act.clone(), iters=iters, return_trace=True)
        d = (upd - u2).abs().max().item()
        print(f"{name}: |ref-port| ascent={d:.2e} mean|a|={mean_abs:.5f} norms[0]={norms[0]:.4e} "
              f"max norm={norms.max():.4e}")
        assert d < 1e-6
        np.savez(os.path.join(OUT, name + ".npz"), wseed=wseed, scale=scale, iters=iters,
                 obs=obs.numpy(), action=act.numpy(), p1=p1.numpy(), p2=p2.numpy(), q_min=qm.numpy(),
                 dq_da=a0.grad.numpy(), new_action=upd.numpy(), mean_abs=mean_abs, norms=norms,
                 checksum=param_checksum(params, port.CRITIC_KEYS))


def noise_fixture():
    """N3: the reference's own add_mixed_normal_noise / add_normal_noise (ddiffpg/utils/noise.py) with the
    Gaussian draw replaced by std * z."""
    import sys
    sys.path.insert(0, ref_loader.REF_ROOT)
    from ddiffpg.utils import noise as ref_noise
    g = torch.Generator().manual_seed(4000)
    a = torch.rand(37, 8, generator=g) * 2 - 1
    z = torch.randn(37, 8, generator=g)
    real_normal = torch.normal
    torch.normal = lambda mean, std: mean + std * z
    try:
        mixed = ref_noise.add_mixed_normal_noise(a, std_max=0.8, std_min=0.05, out_bounds=[-1.0, 1.0])
        fixed = ref_noise.add_normal_noise(a, std=0.3, out_bounds=[-1.0, 1.0])
        tgt = ref_noise.add_normal_noise(a, std=0.8, noise_bounds=[-0.2, 0.2], out_bounds=[-1.0, 1.0])
    finally:
        torch.normal = real_normal
    assert torch.equal(mixed, port.add_noise_to_actions(a, z, 0.05, 0.8))
    assert torch.equal(tgt, port.add_noise_to_actions(a, z, 0.8, 0.8, noise_bounds=(-0.2, 0.2)))
    np.savez(os.path.join(OUT, "n3_noise.npz"), a=a.numpy(), z=z.numpy(), mixed=mixed.numpy(), fixed=fixed.numpy(),
             tgt=tgt.numpy())
    print("n3_noise: port == reference")


def critic_fixture():
    """N1: the reference's own C51 projection (ddiffpg/utils/distl_util.py) and critic loss
    (ddiffpg/algo/ddiffpg.py:322-349) around its real DistributionalDoubleQ modules."""
    import sys
    import torch.nn.functional as F
    sys.path.insert(0, ref_loader.REF_ROOT)
    from ddiffpg.utils.distl_util import projection
    R = ref_loader.load_reference()
    g = torch.Generator().manual_seed(5000)
    B, gamma = 48, 0.99
    params = port.init_critic_params(41, scale=1.5)
    params_t = port.init_critic_params(42, scale=1.5)
    obs, nobs = torch.randn(B, 29, generator=g), torch.randn(B, 29, generator=g)
    act, nact = torch.rand(B, 8, generator=g) * 2 - 1, torch.rand(B, 8, generator=g) * 2 - 1
    reward = torch.rand(B, 1, generator=g) * 2.0
    reward[:4] = 0.0                                       # integer-b cases: l == u fix-ups
    reward[4:8] = 4.5                                      # upper atoms clamp at v_max
    done = (torch.rand(B, 1, generator=g) < 0.3).float()
    critic = R.DistributionalDoubleQ(29, 8, v_min=0, v_max=5, num_atoms=51, device="cpu"); critic.load_state_dict(params)
    target = R.DistributionalDoubleQ(29, 8, v_min=0, v_max=5, num_atoms=51, device="cpu"); target.load_state_dict(params_t)
    with torch.no_grad():
        t1, t2 = target.get_q1_q2(nobs, nact)
        pr1 = projection(next_dist=t1, reward=reward, done=done, gamma=gamma, v_min=0, v_max=5, num_atoms=51,
                         support=critic.z_atoms, device="cpu")
        pr2 = projection(next_dist=t2, reward=reward, done=done, gamma=gamma, v_min=0, v_max=5, num_atoms=51,
                         support=critic.z_atoms, device="cpu")
        tq = torch.min(pr1, pr2)
    c1, c2 = critic.get_q1_q2(obs, act)
    loss = F.binary_cross_entropy(c1, tq) + F.binary_cross_entropy(c2, tq)
    loss.backward()
    grads = {k: v.grad.detach() for k, v in critic.named_parameters()}
    tq_port = port.critic_target_dist(params_t, nobs, nact, reward, done, gamma)
    l_port, g_port = port.critic_loss_and_grads(params, tq_port, obs, act)
    d = max((grads[k] - g_port[k]).abs().max().item() for k in port.CRITIC_KEYS)
    print(f"n1_critic: |proj diff|={(tq - tq_port).abs().max().item():.2e} |dloss|={abs(loss.item() - l_port.item()):.2e} max|dgrad|={d:.2e}")
    assert (tq - tq_port).abs().max().item() < 1e-7 and d < 1e-7
    save = dict(gamma=gamma, obs=obs.numpy(), act=act.numpy(), nobs=nobs.numpy(), nact=nact.numpy(), reward=reward.numpy(),
                done=done.numpy(), proj1=pr1.numpy(), target_q=tq.numpy(), loss=loss.item(),
                checksum=param_checksum(params, port.CRITIC_KEYS), checksum_t=param_checksum(params_t, port.CRITIC_KEYS))
    for i, k in enumerate(port.CRITIC_KEYS):
        save[f"g_{i}"] = grads[k].numpy() if grads[k].numel() <= 8192 else grads[k].flatten()[::97].numpy()
        save[f"gnorm_{i}"] = float(grads[k].norm())
    np.savez(os.path.join(OUT, "n1_critic.npz"), **save)


def rnd_fixture():
    """N4: the reference's own RNDModel (ddiffpg/models/mlp.py:233-267) inside its own IntrinsicM
    (ddiffpg/utils/intrinsic.py), NovelD reward with and without running normalisation, one predictor update."""
    import sys
    R = ref_loader.load_reference()
    sys.modules["ddiffpg.models.mlp"] = sys.modules["_ref_mlp"]          # intrinsic.py imports RNDModel from there
    sys.path.insert(0, ref_loader.REF_ROOT)
    from ddiffpg.utils.intrinsic import IntrinsicM
    g = torch.Generator().manual_seed(6000)
    B = 40
    params = port.init_rnd_params(71)
    obs, nobs = torch.randn(B, 29, generator=g), torch.randn(B, 29, generator=g)
    obs[:, :2] *= 4.0; nobs[:, :2] *= 4.0                                # ant positions span several metres
    im = IntrinsicM((29,), type="noveld", env_name="antmaze-v1", normalize=True, pos_enc=True, L=10, warm_up=0, device="cpu")
    im.rnd_model.load_state_dict(params)
    enc = im.encode_obs(obs)
    nov = im.get_novelty(enc)
    r0 = im.compute_reward(obs, nobs)                                    # update_step == 0: no normalisation
    im.update_step = 1
    r1 = im.compute_reward(obs, nobs)                                    # running statistics updated, normalised
    rms = torch.stack([im.rnd_rms.mean.reshape(()), im.rnd_rms.var.reshape(())])
    x = torch.cat([obs, nobs])
    pf, tf = im.rnd_model(im.encode_obs(x))
    loss = torch.nn.functional.mse_loss(pf, tf.detach())
    loss.backward()
    keys = [k for k in port.RND_KEYS if k.startswith("predictor")]
    grads = {k: dict(im.rnd_model.named_parameters())[k].grad.detach() for k in keys}
    # the port against the real thing
    encp = port.encode_obs_antmaze(obs)
    novp = port.rnd_novelty(params, encp)
    lp, gp = port.rnd_loss_and_grads(params, port.encode_obs_antmaze(x))
    d = max((grads[k] - gp[k]).abs().max().item() for k in keys)
    r0p = port.noveld_reward(novp, port.rnd_novelty(params, port.encode_obs_antmaze(nobs)))
    print(f"n4_rnd: |enc diff|={(enc - encp).abs().max().item():.2e} |nov diff|={(nov - novp).abs().max().item():.2e} "
          f"|r0 diff|={(r0 - r0p).abs().max().item():.2e} |dloss|={abs(loss.item() - lp.item()):.2e} max|dgrad|={d:.2e}")
    assert (enc - encp).abs().max().item() < 1e-6 and (nov - novp).abs().max().item() < 1e-5 and d < 1e-7
    save = dict(obs=obs.numpy(), nobs=nobs.numpy(), enc=enc.numpy(), novelty=nov.numpy(), r0=r0.numpy(), r1=r1.numpy(),
                rms=rms.numpy(), loss=loss.item(), checksum=param_checksum(params, port.RND_KEYS))
    for i, k in enumerate(keys):
        save[f"g_{i}"] = grads[k].numpy() if grads[k].numel() <= 8192 else grads[k].flatten()[::97].numpy()
        save[f"gnorm_{i}"] = float(grads[k].norm())
    np.savez(os.path.join(OUT, "n4_rnd.npz"), **save)


def replay_fixture():
    """N2: the reference's own DiffusionReplayBuffer.sample_batch / update_target_action
    (ddiffpg/replay/simple_replay.py:98-200) and add_embedding (ddiffpg/utils/torch_util.py:17-43), with their
    random draws recorded (torch.randint replaced by a recorded draw, np.random seeded)."""
    import sys
    sys.path.insert(0, ref_loader.REF_ROOT)
    from ddiffpg.replay.simple_replay import DiffusionReplayBuffer
    from ddiffpg.utils.torch_util import add_embedding
    g = torch.Generator().manual_seed(7000)
    buf = DiffusionReplayBuffer(10000, 29, 8, device="cpu")
    lens = [17, 30, 9, 22, 40]
    for tid, n in enumerate(lens):
        tr = (torch.randn(n, 29, generator=g), torch.rand(n, 8, generator=g) * 2 - 1, torch.rand(n, 8, generator=g) * 2 - 1,
              torch.rand(n, generator=g), torch.randn(n, 29, generator=g), torch.rand(n, generator=g) < 0.2)
        buf.add_to_buffer(tr, tid)
    buf.update_target_action_dim([-1, 0])                        # three target-action slots (explore + 2 modes)
    buf.buf_target_action[1] += 0.25; buf.buf_target_action[2] -= 0.25
    store = dict(obs=buf.buf_obs.clone(), action=buf.buf_action.clone(), target_action=buf.buf_target_action.clone(),
                 reward=buf.buf_reward.clone(), next_obs=buf.buf_next_obs.clone(), done=buf.buf_done.clone(), id=buf.buf_id.clone())
    groups = [[0, 1, 2, 3, 4], [0, 3], [1, 4]]
    draws = [torch.randint(sum(lens[t] for t in grp), (sz,), generator=g) for grp, sz in zip(groups, (14, 12, 12))]
    save = {f"store_{k}": v.numpy() for k, v in store.items()}
    real_randint = torch.randint
    for gi, (grp, draw) in enumerate(zip(groups, draws)):
        torch.randint = lambda *a, **k: draw                     # replay the recorded draw inside the reference
        try:
            data, idx = buf.sample_batch(draw.shape[0], grp, gi, device="cpu")
        finally:
            torch.randint = real_randint
        port_data = port.replay_gather(store, idx, gi)
        assert all(torch.equal(a, b) for a, b in zip(data, port_data))
        save[f"draw_{gi}"] = draw.numpy(); save[f"idx_{gi}"] = idx.numpy()
        for name, t in zip(("obs", "action", "target", "reward", "next_obs", "done"), data):
            save[f"g{gi}_{name}"] = t.numpy()
    emb = torch.randn(5, generator=g)
    np.random.seed(11)
    state = save["g1_obs"]
    es = add_embedding(torch.from_numpy(state), emb)             # p = 0.5: np.random.choice(12, 6, replace=False)
    np.random.seed(11)
    zero_idx = np.random.choice(state.shape[0], size=int(state.shape[0] * 0.5), replace=False)
    assert torch.equal(es, port.add_embedding_port(torch.from_numpy(state), emb, zero_idx))
    e0 = add_embedding(torch.from_numpy(save["g0_obs"]), emb, p=0)
    new_action = torch.rand(12, 8, generator=g) * 2 - 1
    buf.update_target_action(new_action, torch.from_numpy(save["idx_2"]), 2)
    save.update(emb=emb.numpy(), zero_idx=zero_idx, emb_state=es.numpy(), emb_state_p0=e0.numpy(), new_action=new_action.numpy(),
                target_after=buf.buf_target_action.numpy())
    np.savez(os.path.join(OUT, "n2_replay.npz"), **save)
    print("n2_replay: port == reference")


def goal_buffer_fixture():
    """N2, the caller side: the reference's own DiffusionGoalBuffer.sample_batch / add_temp_data
    (ddiffpg/replay/diffusion_replay.py:250-332) on a small replay + temp buffer, with its torch.randint draws recorded.
    The module is loaded by path; its clustering dependency (dtaidistance) and the Q scheduler (pulls gym through
    ddiffpg.models) are absent here and not on this path: both are stubbed, and the buffer object is built without
    its constructor (which only sets the attributes assigned below)."""
    import sys
    import types
    sys.path.insert(0, ref_loader.REF_ROOT)
    stub = types.ModuleType("dtaidistance"); stub.dtw_ndim = None
    sys.modules.setdefault("dtaidistance", stub)
    qs = types.ModuleType("ddiffpg.utils.Q_scheduler"); qs.Q_scheduler = object
    sys.modules.setdefault("ddiffpg.utils.Q_scheduler", qs)
    from ddiffpg.replay.diffusion_replay import DiffusionGoalBuffer
    from ddiffpg.replay.simple_replay import DiffusionReplayBuffer
    g = torch.Generator().manual_seed(7100)
    rb = DiffusionReplayBuffer(10000, 29, 8, device="cpu")
    lens = [17, 30, 9, 22, 40, 13]
    for tid, n in enumerate(lens):
        tr = (torch.randn(n, 29, generator=g), torch.rand(n, 8, generator=g) * 2 - 1, torch.rand(n, 8, generator=g) * 2 - 1,
              torch.rand(n, generator=g), torch.randn(n, 29, generator=g), torch.rand(n, generator=g) < 0.2)
        rb.add_to_buffer(tr, tid)
    rb.update_target_action_dim([-1, 0])                          # explore + 2 modes
    rb.buf_target_action[1] += 0.25; rb.buf_target_action[2] -= 0.25
    # add_temp_data calls replay_buffer.sample_batch with its default device='cuda' (no GPU here): same method, device="cpu"
    import functools
    rb.sample_batch = functools.partial(DiffusionReplayBuffer.sample_batch, rb, device="cpu")
    gb = object.__new__(DiffusionGoalBuffer)
    gb.device, gb.replay_buffer = "cpu", rb
    gb.success_id, gb.unsuccess_id = [0, 1, 3, 4], [2, 5]
    gb.clusters, gb.unsuccess_clusters = [[0, 3], [1, 4]], [[2], [5]]
    gb.Qs, gb.embeddings = ["Q0", "Q1", "Q2"], [torch.zeros(5), torch.ones(5), -torch.ones(5)]
    nt = 23
    gb.temp_state, gb.temp_action = torch.randn(nt, 29, generator=g), torch.rand(nt, 8, generator=g) * 2 - 1
    gb.temp_reward, gb.temp_next_state = torch.rand(nt, 1, generator=g), torch.randn(nt, 29, generator=g)
    gb.temp_done = torch.rand(nt, 1, generator=g) < 0.3
    batch = 50                                                    # 50 = 3 * 16 + 2: group 0 takes the remainder
    store = dict(obs=rb.buf_obs.clone(), action=rb.buf_action.clone(), target_action=rb.buf_target_action.clone(),
                 reward=rb.buf_reward.clone(), next_obs=rb.buf_next_obs.clone(), done=rb.buf_done.clone(), id=rb.buf_id.clone())
    temp = dict(state=gb.temp_state, action=gb.temp_action, reward=gb.temp_reward, next_state=gb.temp_next_state, done=gb.temp_done)
    plan = port.goal_buffer_plan(batch, gb.success_id, gb.unsuccess_id, gb.clusters, gb.unsuccess_clusters, nt, rb.buf_id)
    recorded, real_randint = [], torch.randint

    def recording_randint(high, size=None, device=None, **kw):   # the reference draws on `device`; record what it drew
        t = real_randint(high, size, generator=g)
        recorded.append(t)
        return t
    torch.randint = recording_randint
    try:
        data_list = gb.sample_batch(batch, device="cpu")
    finally:
        torch.randint = real_randint
    save = {f"store_{k}": v.numpy() for k, v in store.items()}
    save.update({f"temp_{k}": v.numpy() for k, v in temp.items()})
    it = iter(recorded)
    for i, ((grp, b_sample, b_temp), d) in enumerate(zip(plan, data_list)):
        draw = next(it) if b_sample else None
        tdraw = next(it) if b_temp else None
        assert (draw is None or draw.shape[0] == b_sample) and (tdraw is None or tdraw.shape[0] == b_temp)
        pd, rows = port.goal_buffer_group(store, temp, grp, i, draw, tdraw)
        assert all(torch.equal(a.float(), b) for a, b in zip(d["batch"], pd)), i
        assert (rows is None and d["indices"] is None) or torch.equal(rows, d["indices"])
        assert d["Q"] == gb.Qs[i] and torch.equal(d["embedding"], gb.embeddings[i])
        save[f"draw_{i}"] = draw.numpy() if draw is not None else np.zeros(0, np.int64)
        save[f"tdraw_{i}"] = tdraw.numpy() if tdraw is not None else np.zeros(0, np.int64)
        save[f"idx_{i}"] = d["indices"].numpy() if d["indices"] is not None else np.zeros(0, np.int64)
        for name, t in zip(("obs", "action", "target", "reward", "next_obs", "done"), d["batch"]):
            save[f"g{i}_{name}"] = t.float().numpy()
    assert next(it, None) is None
    save.update(batch=batch, success_id=np.array(gb.success_id), unsuccess_id=np.array(gb.unsuccess_id),
                clusters=np.array(gb.clusters), unsuccess_clusters=np.array(gb.unsuccess_clusters))
    np.savez(os.path.join(OUT, "n2_goal_buffer.npz"), **save)
    print("n2_goal_buffer: port == reference;", [(len(p[0]), p[1], p[2]) for p in plan])


if __name__ == "__main__":
    argv = set(__import__("sys").argv[1:])
    only = {"--noise-only": noise_fixture, "--critic-only": critic_fixture, "--rnd-only": rnd_fixture,
            "--replay-only": replay_fixture, "--goal-buffer-only": goal_buffer_fixture}
    if argv & set(only):
        for flag in sorted(argv & set(only)):
            only[flag]()
    else:
        main()
        noise_fixture()
        critic_fixture()
        rnd_fixture()
        replay_fixture()
        goal_buffer_fixture()
