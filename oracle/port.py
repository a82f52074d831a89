"""Functional CPU restatement of the reference hot path (TEST INFRASTRUCTURE, "port" oracle).

Every function cites the reference lines it follows (paths relative to the reference repo).
Parameters are plain dicts keyed by the reference ``state_dict()`` names, so the same weights
drive the reference modules (``oracle/make_golden.py``), this port and the CUDA kernels.
Works in any float dtype (fp32 for parity, fp64 to calibrate tolerances).
"""
import math

import torch
import torch.nn.functional as F

from .ddpm import DDPMSchedulerRestated

ACTOR_KEYS = (
    "net.time_mlp.1.weight", "net.time_mlp.1.bias", "net.time_mlp.3.weight", "net.time_mlp.3.bias",
    "net.mlp.0.weight", "net.mlp.0.bias", "net.mlp.2.weight", "net.mlp.2.bias",
    "net.mlp.4.weight", "net.mlp.4.bias", "net.mlp.6.weight", "net.mlp.6.bias",
)
CRITIC_KEYS = tuple(f"net_q{j}.net.{i}.{w}" for j in (1, 2) for i in (0, 2, 4, 6)
                    for w in ("weight", "bias"))


# --------------------------------------------------------------------------- init helpers
def _linear_init(gen, out_f, in_f, dtype=torch.float32):
    """nn.Linear default init (kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(in), 1/sqrt(in)) for both)."""
    bound = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=gen, dtype=torch.float64) * 2 - 1) * bound
    b = (torch.rand(out_f, generator=gen, dtype=torch.float64) * 2 - 1) * bound
    return w.to(dtype), b.to(dtype)


def init_actor_params(seed, S=34, A=8, h=1024, D=256, scale=1.0):
    """Random actor weights with the reference shapes (diffusion_mlp.py:38-58); h=1024 is the reference.

    ``scale`` > 1 widens the weights so the chain leaves the small-output regime of a fresh init.
    """
    g = torch.Generator().manual_seed(seed)
    shapes = [(4 * D, D), (D, 4 * D), (h, D + S + A), (h // 2, h), (h // 4, h // 2), (A, h // 4)]
    p = {}
    for i, (o, k) in enumerate(shapes):
        w, b = _linear_init(g, o, k)
        p[ACTOR_KEYS[2 * i]] = w * scale
        p[ACTOR_KEYS[2 * i + 1]] = b * scale
    return p


def init_critic_params(seed, O=29, A=8, atoms=51, hidden=(512, 256, 128), scale=1.0):
    """Random double-Q weights with the reference shapes (mlp.py:23-35,131-138)."""
    g = torch.Generator().manual_seed(seed)
    dims = [O + A, *hidden, atoms]
    p = {}
    for j in (1, 2):
        for li, (k, o) in enumerate(zip(dims[:-1], dims[1:])):
            w, b = _linear_init(g, o, k)
            p[f"net_q{j}.net.{2 * li}.weight"] = w * scale
            p[f"net_q{j}.net.{2 * li}.bias"] = b * scale
    return p


def cast_params(p, dtype):
    return {k: v.detach().to(dtype).clone() for k, v in p.items()}


# --------------------------------------------------------------------------- H1 / H3 network
def sinusoidal_pos_emb(t, dim=256):
    """diffusion_mlp.py:14-21 -- [sin(t f_i) | cos(t f_i)], f_i = exp(-i ln(1e4)/(dim/2-1)).

    The reference builds f in fp32 regardless of the dtype of ``t``; a float64 ``t`` promotes.
    """
    half = dim // 2
    c = math.log(10000) / (half - 1)
    f = torch.exp(torch.arange(half) * -c)
    e = t[:, None] * f[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def time_mlp(p, t, dim=256):
    """diffusion_mlp.py:38-43 -- pos-emb -> Linear(D,4D) -> Mish -> Linear(4D,D)."""
    e = sinusoidal_pos_emb(t, dim).to(p["net.time_mlp.1.weight"].dtype)
    hdn = F.mish(F.linear(e, p["net.time_mlp.1.weight"], p["net.time_mlp.1.bias"]))
    return F.linear(hdn, p["net.time_mlp.3.weight"], p["net.time_mlp.3.bias"])


def actor_eps(p, x, t, cond):
    """DiffusionNet.forward, diffusion_mlp.py:62-73 -- cat order is [temb, cond, x] (:70)."""
    D = p["net.time_mlp.3.weight"].shape[0]
    temb = time_mlp(p, t, D)
    z = torch.cat([temb, cond, x], dim=-1)
    z = F.mish(F.linear(z, p["net.mlp.0.weight"], p["net.mlp.0.bias"]))
    z = F.mish(F.linear(z, p["net.mlp.2.weight"], p["net.mlp.2.bias"]))
    z = F.mish(F.linear(z, p["net.mlp.4.weight"], p["net.mlp.4.bias"]))
    return F.linear(z, p["net.mlp.6.weight"], p["net.mlp.6.bias"])


@torch.no_grad()
def actor_sample(p, state, noise, T, return_chain=False):
    """DiffusionPolicy.get_actions(sample=True, add_noise=False), diffusion_mlp.py:219-251.

    ``noise[0]`` replaces the initial ``torch.randn((B, A))`` (:222); ``noise[j]`` (j >= 1) replaces
    the draw inside ``scheduler.step`` at t = T - j (t > 0 only).  ``noise`` has shape [T, B, A].
    """
    B = state.shape[0]
    dtype = state.dtype
    sched = DDPMSchedulerRestated(num_train_timesteps=T, noise_queue=[n for n in noise[1:]])
    x = noise[0].clone()
    sched.set_timesteps(T)
    chain = [x.clone()]
    for k in sched.timesteps:
        tt = (torch.ones(B) * k).to(dtype)            # :229 (float timesteps in the sampler)
        eps = actor_eps(p, x, tt, state)
        x = sched.step(model_output=eps, timestep=k, sample=x).prev_sample.to(dtype)
        chain.append(x.clone())
    return (x, chain) if return_chain else x


def actor_loss(p, state, action, noise, timesteps, T):
    """DiffusionPolicy.get_loss, diffusion_mlp.py:294-321 (noise and timesteps injected)."""
    sched = DDPMSchedulerRestated(num_train_timesteps=T)
    noisy = sched.add_noise(action, noise, timesteps)
    eps = actor_eps(p, noisy, timesteps, state)       # int64 timesteps feed the pos-emb (:313)
    return F.mse_loss(eps, noise)


def actor_loss_and_grads(p, state, action, noise, timesteps, T):
    """loss + d loss / d params in ACTOR_KEYS order (what ``objective.backward()`` leaves in .grad,
    ac_base.py:83-85), before clipping."""
    q = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    loss = actor_loss(q, state, action, noise, timesteps, T)
    grads = torch.autograd.grad(loss, [q[k] for k in ACTOR_KEYS])
    return loss.detach(), {k: g for k, g in zip(ACTOR_KEYS, grads)}


def add_noise_to_actions(actions, z, std_min, std_max, noise_bounds=None, out_bounds=(-1.0, 1.0)):
    """utils/noise.py:19-41 with the Gaussian draw replaced by std * z (z standard normal, injected):
    add_mixed_normal_noise (std = linspace(std_min, std_max, B) per row) and, for std_min == std_max,
    add_normal_noise; optional noise clamp (get_tgt_policy_actions, ddiffpg.py:104-109) and output clamp."""
    B = actions.shape[0]
    std_seq = torch.linspace(std_min, std_max, B).to(actions.dtype).unsqueeze(-1).expand(actions.shape)
    noise = std_seq * z
    if noise_bounds is not None:
        noise = noise.clamp(noise_bounds[0], noise_bounds[1])
    out = actions + noise
    if out_bounds is not None:
        out = out.clamp(out_bounds[0], out_bounds[1])
    return out


def clip_coef(total_norm, max_norm=1.0):
    """torch.nn.utils.clip_grad_norm_: coef = clamp(max_norm / (norm + 1e-6), max=1)."""
    return torch.clamp(max_norm / (total_norm + 1e-6), max=1.0)


def adamw_train_step(p, state, action, noise, timesteps, T, opt_state=None, lr=3e-4,
                     betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, max_norm=1.0):
    """update_actor, ddiffpg.py:353-356 = get_loss + optimizer_update (ac_base.py:83-92) with
    AdamW(lr=actor_lr) (ac_base.py:52; torch defaults betas .9/.999, eps 1e-8, wd 1e-2).

    Returns (loss, pre-clip grad norm, new params, new optimizer state)."""
    loss, g = actor_loss_and_grads(p, state, action, noise, timesteps, T)
    total = torch.sqrt(sum((g[k].double() ** 2).sum() for k in ACTOR_KEYS)).to(loss.dtype)
    coef = clip_coef(total, max_norm)
    if opt_state is None:
        opt_state = {"step": 0, "m": {k: torch.zeros_like(p[k]) for k in ACTOR_KEYS},
                     "v": {k: torch.zeros_like(p[k]) for k in ACTOR_KEYS}}
    step = opt_state["step"] + 1
    b1, b2 = betas
    new_p, m_new, v_new = {}, {}, {}
    for k in ACTOR_KEYS:
        gk = g[k] * coef
        w = p[k] * (1 - lr * weight_decay)
        m = opt_state["m"][k] * b1 + (1 - b1) * gk
        v = opt_state["v"][k] * b2 + (1 - b2) * gk * gk
        denom = v.sqrt() / math.sqrt(1 - b2 ** step) + eps
        new_p[k] = w - (lr / (1 - b1 ** step)) * m / denom
        m_new[k], v_new[k] = m, v
    return loss, total, new_p, {"step": step, "m": m_new, "v": v_new}


# --------------------------------------------------------------------------- H2 critic
def mlp_elu(p, prefix, x):
    """MLPNet / create_simple_mlp, mlp.py:13-35 -- Linear/ELU x3 then Linear."""
    for i in (0, 2, 4):
        x = F.elu(F.linear(x, p[f"{prefix}.net.{i}.weight"], p[f"{prefix}.net.{i}.bias"]))
    return F.linear(x, p[f"{prefix}.net.6.weight"], p[f"{prefix}.net.6.bias"])


def z_atoms(v_min=0.0, v_max=5.0, atoms=51, dtype=torch.float32):
    """mlp.py:141 -- linspace(v_min, v_max, num_atoms)."""
    return torch.linspace(v_min, v_max, atoms).to(dtype)


def q1_q2(p, obs, act):
    """DistributionalDoubleQ.get_q1_q2, mlp.py:149-151."""
    x = torch.cat((obs, act), dim=1)
    return (torch.softmax(mlp_elu(p, "net_q1", x), dim=1),
            torch.softmax(mlp_elu(p, "net_q2", x), dim=1))


def q_min(p, obs, act, v_min=0.0, v_max=5.0):
    """DistributionalDoubleQ.get_q_min, mlp.py:143-147."""
    p1, p2 = q1_q2(p, obs, act)
    z = z_atoms(v_min, v_max, p1.shape[1], p1.dtype)
    return torch.min(torch.sum(p1 * z, dim=1), torch.sum(p2 * z, dim=1))


def c51_projection(next_dist, reward, done, gamma, v_min=0.0, v_max=5.0, num_atoms=51):
    """utils/distl_util.py:4-20 -- categorical (C51) projection of r + (1-done) gamma z onto the support."""
    support = z_atoms(v_min, v_max, num_atoms, next_dist.dtype)
    delta_z = (v_max - v_min) / (num_atoms - 1)
    B = reward.shape[0]
    target_z = (reward + (1 - done) * gamma * support).clamp(min=v_min, max=v_max)
    b = (target_z - v_min) / delta_z
    lo = b.floor().long()
    up = b.ceil().long()
    lo[torch.logical_and(up > 0, lo == up)] -= 1
    up[torch.logical_and(lo < (num_atoms - 1), lo == up)] += 1
    proj = torch.zeros_like(next_dist)
    offset = (torch.arange(B) * num_atoms).unsqueeze(1).expand(B, num_atoms)
    proj.view(-1).index_add_(0, (lo + offset).view(-1), (next_dist * (up.to(b.dtype) - b)).view(-1))
    proj.view(-1).index_add_(0, (up + offset).view(-1), (next_dist * (b - lo.to(b.dtype))).view(-1))
    return proj


def critic_target_dist(p_target, next_obs, next_action, reward, done, gamma, v_min=0.0, v_max=5.0):
    """AgentDDiffPG.update_critic, ddiffpg.py:325-346 (no_grad block): min of the two projected target heads.
    reward / done are [B, 1] like the replay buffer hands them over."""
    with torch.no_grad():
        t1, t2 = q1_q2(p_target, next_obs, next_action)
        n = t1.shape[1]
        return torch.min(c51_projection(t1, reward, done, gamma, v_min, v_max, n),
                         c51_projection(t2, reward, done, gamma, v_min, v_max, n))


def critic_loss_and_grads(p, target_q, obs, action):
    """ddiffpg.py:348-349: BCE(current_Q1, target_Q) + BCE(current_Q2, target_Q) and its gradients in
    CRITIC_KEYS order (what objective.backward() leaves in .grad, before clipping)."""
    q = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    c1, c2 = q1_q2(q, obs, action)
    loss = F.binary_cross_entropy(c1, target_q) + F.binary_cross_entropy(c2, target_q)
    grads = torch.autograd.grad(loss, [q[k] for k in CRITIC_KEYS])
    return loss.detach(), {k: g for k, g in zip(CRITIC_KEYS, grads)}


def q_action_ascent(p, obs, action, iters=20, lr=0.03, eps=1e-5, max_norm=1.0,
                    betas=(0.9, 0.999), v_min=0.0, v_max=5.0, return_trace=False, return_grads=False):
    """AgentDDiffPG.update_target_action, ddiffpg.py:358-373 (+ optimizer_update, ac_base.py:83-92).

    Uses the real torch.optim.Adam and clip_grad_norm_ like the reference.  Returns
    (mean |a|, new action[B, A]) and, with ``return_trace``, the per-iteration pre-clip gradient norms
    and the per-iteration, per-row gap Q1-Q2 [iters, B].  The ascent climbs min(Q1, Q2), so rows drift
    onto the ridge Q1 == Q2 where the arg-min (hence the gradient) flips on the last float bit; the
    parity tests use the gap to treat such rows with the bound the non-smooth objective allows.
    ``return_grads`` appends the per-iteration pre-clip gradients [iters, B, A] (what Adam consumes before
    the clip coefficient), so a test can tell elements with a well-defined step sign from near-zero ones.
    """
    pp = {k: v.detach() for k, v in p.items()}
    action = action.detach().clone()
    lim = 1 - 1e-5
    action.clamp_(-lim, lim)
    opt = torch.optim.Adam([action], lr=lr, eps=eps, betas=betas)
    norms, gaps, grads = [], [], []
    for _ in range(iters):
        action.requires_grad_(True)
        if return_trace:
            with torch.no_grad():
                d1, d2 = q1_q2(pp, obs, action)
                zz = z_atoms(v_min, v_max, d1.shape[1], d1.dtype)
                gaps.append(((d1 * zz).sum(1) - (d2 * zz).sum(1)).clone())
        loss = -q_min(pp, obs, action, v_min, v_max).mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if return_grads:
            grads.append(action.grad.detach().clone())
        norms.append(torch.nn.utils.clip_grad_norm_([action], max_norm=max_norm).detach().clone())
        opt.step()
        action.requires_grad_(False)
        action.clamp_(-lim, lim)
    out = action.detach().clone()
    res = (torch.abs(out).mean().item(), out)
    if return_trace:
        res = res + (torch.stack(norms), torch.stack(gaps))
    if return_grads:
        res = res + (torch.stack(grads),)
    return res


# ------------------------------------------------------------------------------------------ N4: RND / NovelD
RND_KEYS = tuple(f"{net}.{i}.{w}" for net in ("predictor", "target") for i in (0, 2, 4, 6) for w in ("weight", "bias"))


def init_rnd_params(seed, D=69, hidden=(512, 256, 128), feat=128, scale=1.0):
    """Seeded RNDModel state_dict (ddiffpg/models/mlp.py:233-260 shapes; values from a plain generator)."""
    g = torch.Generator().manual_seed(seed)
    dims = (D,) + tuple(hidden) + (feat,)
    p = {}
    for net in ("predictor", "target"):
        for li, i in enumerate((0, 2, 4, 6)):
            fan_in = dims[li]
            p[f"{net}.{i}.weight"] = (torch.rand(dims[li + 1], fan_in, generator=g) * 2 - 1) * scale * (3.0 / fan_in) ** 0.5
            p[f"{net}.{i}.bias"] = (torch.rand(dims[li + 1], generator=g) * 2 - 1) * 0.1
    return p


def _rnd_net(p, net, x):
    h = x
    for i in (0, 2, 4):
        h = F.elu(F.linear(h, p[f"{net}.{i}.weight"], p[f"{net}.{i}.bias"]))
    return F.linear(h, p[f"{net}.6.weight"], p[f"{net}.6.bias"])


def rnd_forward(p, x):
    """RNDModel.forward (mlp.py:262-266): (predict_feature, target_feature)."""
    return _rnd_net(p, "predictor", x), _rnd_net(p, "target", x)


def rnd_novelty(p, x):
    """IntrinsicM.get_novelty (utils/intrinsic.py:62-65)."""
    pf, tf = rnd_forward(p, x)
    return torch.norm(pf - tf, dim=1, p=2)


def rnd_loss_and_grads(p, x):
    """IntrinsicM.update up to the backward (utils/intrinsic.py:67-72): mse loss and predictor gradients."""
    q = {k: (v.clone().requires_grad_(True) if k.startswith("predictor") else v) for k, v in p.items()}
    pf, tf = rnd_forward(q, x)
    loss = F.mse_loss(pf, tf.detach())
    keys = [k for k in RND_KEYS if k.startswith("predictor")]
    grads = torch.autograd.grad(loss, [q[k] for k in keys])
    return loss.detach(), dict(zip(keys, grads))


def encode_obs_antmaze(obs, L=10):
    """IntrinsicM.encode_obs for antmaze (utils/intrinsic.py:85-88,122-171): NeRF encoding of the 2-d position."""
    x = obs[:, :2]
    outs = [x]
    for k in range(L):
        f = 2.0 ** k
        outs += [torch.sin(x * f), torch.cos(x * f)]
    return torch.cat(outs + [obs[:, 2:]], dim=1)


def noveld_reward(nov_obs, nov_next):
    """IntrinsicM.compute_reward, type 'noveld', before normalisation kicks in (utils/intrinsic.py:45-59)."""
    return 0.01 * torch.clamp(nov_next - 0.5 * nov_obs, min=0).unsqueeze(1)


# ------------------------------------------------------------------------------------------ N2: batch assembly
def replay_gather(bufs, indices, target_idx):
    """DiffusionReplayBuffer.sample_batch after the index draw (replay/simple_replay.py:155-162).
    bufs: dict obs, action, target_action [K,N,A], reward [N,1], next_obs, done (bool [N,1])."""
    return (bufs["obs"][indices], bufs["action"][indices], bufs["target_action"][target_idx, indices],
            bufs["reward"][indices], bufs["next_obs"][indices], bufs["done"][indices].float())


def goal_buffer_plan(batch_size, success_id, unsuccess_id, clusters, unsuccess_clusters, temp_size, buf_id):
    """The group / size split of DiffusionGoalBuffer.sample_batch (replay/diffusion_replay.py:255-266) and the
    temp-buffer share of add_temp_data (:286-292): per group (trajectory ids, rows from the replay, rows from the
    temp buffer)."""
    groups = [list(success_id) + list(unsuccess_id)] + [list(c) + list(u) for c, u in zip(clusters, unsuccess_clusters)]
    sizes = [batch_size // len(groups)] * len(groups)
    sizes[0] += batch_size % len(groups)
    plan = []
    for i, (grp, b) in enumerate(zip(groups, sizes)):
        buffer_size = int(torch.isin(buf_id, torch.tensor(grp, dtype=buf_id.dtype)).sum())
        b_temp = int((temp_size / (temp_size + buffer_size)) * b) if i == 0 else 0
        plan.append((grp, b - b_temp, b_temp))
    return plan


def goal_buffer_group(bufs, temp, grp, target_idx, draw, temp_draw):
    """One add_temp_data call (:285-332) after its two index draws.  temp: dict state, action, reward, next_state, done."""
    parts, rows = [], None
    if draw is not None:
        avail = torch.where(torch.isin(bufs["id"], torch.tensor(grp, dtype=bufs["id"].dtype)))[0]
        rows = avail[draw]
        parts.append(replay_gather(bufs, rows, target_idx))
    if temp_draw is not None:
        parts.append((temp["state"][temp_draw], temp["action"][temp_draw], temp["action"][temp_draw],
                      temp["reward"][temp_draw], temp["next_state"][temp_draw], temp["done"][temp_draw].float()))
    return tuple(torch.cat([p[k].float() for p in parts]) for k in range(6)), rows


def add_embedding_port(state, embedding, zero_indices):
    """add_embedding (utils/torch_util.py:17-43) with the np.random.choice draw passed in."""
    new_embedding = embedding.unsqueeze(0).repeat(state.shape[0], 1)
    if len(zero_indices):
        new_embedding[torch.as_tensor(zero_indices).long()] = 0.0
    return torch.cat([state, new_embedding], dim=1)


def replay_scatter(bufs, new_action, indices, i):
    """DiffusionReplayBuffer.update_target_action (replay/simple_replay.py:198-200), in place."""
    bufs["target_action"][i, indices] = new_action
