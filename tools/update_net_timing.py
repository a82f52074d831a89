"""One iteration of the reference's AgentDDiffPG.update_net (ddiffpg/algo/ddiffpg.py:205-300) at its own configuration
(batch 4 096 over the explore group + K mode groups, AntMaze shapes, T = 5, 20 ascent iterations), composed of this repo's
pieces on one GPU, next to the same sequence through the oracle port on the host CPU.

per iteration:  goal-buffer batch (all groups, one gather)  ->  NovelD reward (RND novelty of obs and next_obs)
                per group: target-policy actions (fused sampler + noise epilogue) -> critic update -> soft_update
                           -> 20-iteration action ascent
                target-action scatter-back  ->  denoiser update on the concatenated batch  ->  RND update
"""
import argparse, sys, time, torch
sys.path.insert(0, '.')
from oracle import port
from ddiffpg_b200 import (DiffusionPolicy, DistributionalDoubleQ, DiffusionReplayBuffer, FusedActorTrainer, FusedCriticTrainer,
                          FusedRNDTrainer, GoalBufferKernels, RNDModel, add_embedding, get_tgt_policy_actions,
                          q_action_ascent_segments, soft_update)

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--modes", type=int, default=2)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--cpu-iters", type=int, default=1)
args = ap.parse_args()
dev, B, K, O, A, E, T = "cuda", args.batch, args.modes, 29, 8, 5, 5
G = K + 1
gen = torch.Generator().manual_seed(0)

# ---- synthetic replay: 60 trajectories of 300 rows, half of them successful, clustered into K modes
rb = DiffusionReplayBuffer(10 ** 6, O, A, device=dev)
n_traj, L = 60, 300
N = n_traj * L
rb.buf_obs, rb.buf_next_obs = torch.randn(N, O, generator=gen).to(dev), torch.randn(N, O, generator=gen).to(dev)
rb.buf_action = (torch.rand(N, A, generator=gen) * 2 - 1).to(dev)
rb.buf_target_action = rb.buf_action.unsqueeze(0).repeat(G, 1, 1).contiguous()
rb.buf_reward, rb.buf_done = torch.rand(N, 1, generator=gen).to(dev), (torch.rand(N, 1, generator=gen) < 0.01).to(dev)
rb.buf_id = torch.arange(n_traj).repeat_interleave(L).reshape(-1, 1).float().to(dev)
rb.cur_capacity = N

class Goal(GoalBufferKernels):
    pass
gb = Goal()
gb.device, gb.replay_buffer = dev, rb
succ = list(range(0, n_traj // 2))
gb.success_id, gb.unsuccess_id = succ, list(range(n_traj // 2, n_traj))
gb.clusters = [succ[m::K] for m in range(K)]
gb.unsuccess_clusters = [gb.unsuccess_id[m::K] for m in range(K)]
gb.embeddings = [torch.randn(E, generator=gen).to(dev) for _ in range(G)]
gb.temp_state, gb.temp_action = torch.randn(200, O, generator=gen).to(dev), (torch.rand(200, A, generator=gen) * 2 - 1).to(dev)
gb.temp_reward, gb.temp_next_state = torch.rand(200, 1, generator=gen).to(dev), torch.randn(200, O, generator=gen).to(dev)
gb.temp_done = torch.zeros(200, 1, dtype=torch.bool, device=dev)

# ---- networks (reference shapes), fused trainers
torch.manual_seed(0)
actor = DiffusionPolicy(O + E, A, T, device=dev, precision=args.precision).to(dev)
actor.train_precision = args.precision
actor_tr = FusedActorTrainer(actor, precision=args.precision, graph=True, process_group=False)
Qs = []
for g in range(G):
    q = DistributionalDoubleQ(O, A, v_min=0, v_max=5, num_atoms=51, device=dev, precision=args.precision).to(dev)
    qt = DistributionalDoubleQ(O, A, v_min=0, v_max=5, num_atoms=51, device=dev, precision=args.precision).to(dev).requires_grad_(False)
    qt.load_state_dict(q.state_dict())
    q.train_precision = args.precision
    Qs.append({"Q": q, "target_Q": qt, "trainer": FusedCriticTrainer(q, qt, lr=5e-4, graph=True, process_group=False)})
gb.Qs = Qs
rnd = RNDModel(69, precision=args.precision).to(dev)
rnd_tr = FusedRNDTrainer(rnd, graph=True)

def _enc(obs, Lp=10):        # IntrinsicM.encode_obs for AntMaze (utils/intrinsic.py:85-94): NeRF encoding of (x, y), on the device
    x = obs[:, :2]
    outs = [x]
    for k in range(Lp):
        outs += [torch.sin(x * 2.0 ** k), torch.cos(x * 2.0 ** k)]
    return torch.cat(outs + [obs[:, 2:]], dim=1)

from ddiffpg_b200.models import _PackCache
ascent_cache = _PackCache()


def iteration():
    data_list = gb.sample_batch(B)
    obs = torch.cat([d["batch"][0] for d in data_list]); nobs = torch.cat([d["batch"][4] for d in data_list])
    reward = torch.cat([d["batch"][3] for d in data_list])
    nov = rnd.novelty(_enc(torch.cat([obs, nobs])))
    r_int = 0.01 * torch.clamp(nov[obs.shape[0]:] - 0.5 * nov[:obs.shape[0]], min=0).unsqueeze(1)
    rewards = reward + r_int
    prev, states = 0, []
    for i, d in enumerate(data_list):
        s, a, ta, _, ns, dn = d["batch"]
        n = s.shape[0]
        r = r_int[prev:prev + n] if i == 0 else rewards[prev:prev + n]
        es = add_embedding(s, d["embedding"], p=0 if i == 0 else 0.5)
        ens = add_embedding(ns, d["embedding"], p=0 if i == 0 else 0.5)
        nact = get_tgt_policy_actions(actor, ens)
        d["Q"]["trainer"].step(s, a, r, ns, nact, dn, gamma_n=0.99 ** 3)
        soft_update(d["Q"]["target_Q"], d["Q"]["Q"], 0.05)
        states.append(es)
        prev += n
    # the ascent of group i only reads critic i: the three reference calls (one per group, after that group's critic update)
    # are one segmented call here -- same per-group 1/B, clip norm and Adam state
    critics = [d["Q"]["Q"] for d in data_list]
    seg = [0]
    for d in data_list:
        seg.append(seg[-1] + d["batch"][0].shape[0])
    ta_all = torch.cat([d["batch"][2] for d in data_list]).contiguous()
    for c in critics:
        c.requires_grad_(False)
    q_action_ascent_segments(critics, obs, ta_all, seg, iters=20, precision=args.precision, cache=ascent_cache)
    for c in critics:
        c.requires_grad_(True)
    ascent_cache.dirty = True                 # the critics change every iteration
    actions = [ta_all]
    for i, d in enumerate(data_list):
        if d["indices"] is not None:
            rb.update_target_action(ta_all[seg[i]:seg[i] + d["indices"].shape[0]], d["indices"], i)
    actor_tr.step(torch.cat(states), torch.cat(actions))
    rnd_tr.step(_enc(torch.cat([obs, nobs])))

for _ in range(3):
    iteration()
torch.cuda.synchronize()
t0 = time.perf_counter()
n_it = 10
for _ in range(n_it):
    iteration()
torch.cuda.synchronize()
gpu_ms = (time.perf_counter() - t0) / n_it * 1e3

# ---- the same sequence through the oracle port (torch CPU fp32, all host threads)
torch.set_num_threads(torch.get_num_threads())
pa, pq, pr = port.init_actor_params(0), [port.init_critic_params(10 + g) for g in range(G)], port.init_rnd_params(3)
sizes = [B // G + (B % G if g == 0 else 0) for g in range(G)]
cpu = [dict(s=torch.randn(n, O), a=torch.rand(n, A) * 2 - 1, r=torch.rand(n, 1), ns=torch.randn(n, O), d=torch.zeros(n, 1)) for n in sizes]
def cpu_iteration():
    obs = torch.cat([c["s"] for c in cpu]); nobs = torch.cat([c["ns"] for c in cpu])
    port.rnd_novelty(pr, port.encode_obs_antmaze(torch.cat([obs, nobs])))
    states, actions = [], []
    for g, c in enumerate(cpu):
        n = c["s"].shape[0]
        emb = torch.zeros(n, E)
        nact = port.actor_sample(pa, torch.cat([c["ns"], emb], 1), torch.randn(T, n, A), T)
        tq = port.critic_target_dist(pq[g], c["ns"], nact, c["r"], c["d"], 0.99 ** 3).clamp_max(1.0)
        port.critic_loss_and_grads(pq[g], tq, c["s"], c["a"])
        _, ta = port.q_action_ascent(pq[g], c["s"], c["a"], iters=20)
        states.append(torch.cat([c["s"], emb], 1)); actions.append(ta)
    st, ac = torch.cat(states), torch.cat(actions)
    port.adamw_train_step(pa, st, ac, torch.randn(st.shape[0], A), torch.randint(0, T, (st.shape[0],)), T)
    port.rnd_loss_and_grads(pr, port.encode_obs_antmaze(torch.cat([obs, nobs])))
cpu_iteration()
t0 = time.perf_counter()
for _ in range(args.cpu_iters):
    cpu_iteration()
cpu_ms = (time.perf_counter() - t0) / args.cpu_iters * 1e3
print(f"update_net iteration, batch {B}, {G} groups, {args.precision}: GPU {gpu_ms:.2f} ms (wall clock, host launches included); "
      f"oracle port on {torch.get_num_threads()} host threads {cpu_ms:.0f} ms (optimizer steps of the critics / RND not included)")
