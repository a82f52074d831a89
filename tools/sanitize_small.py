"""Small invocations of every kernel family, for compute-sanitizer (memcheck) runs on the GPU box."""
import sys, torch
sys.path.insert(0, '.')
from oracle import port
from tests.util import make_policy, make_critic
from ddiffpg_b200 import (q_action_ascent_segments, FusedActorTrainer, RNDModel, DiffusionReplayBuffer, add_embedding,
                          critic_loss_and_grads)
dev = "cuda"
gen = torch.Generator().manual_seed(0)
T = 5
for prec in ("fp32", "bf16"):
    pol = make_policy(port.init_actor_params(1), T, precision=prec)
    for B in (1, 130, 300):
        s, n = torch.randn(B, 34, generator=gen).to(dev), torch.randn(T, B, 8, generator=gen).to(dev)
        pol.get_actions(s, noise=n, expl_std=(0.05, 0.8))
    tr = FusedActorTrainer(pol, precision=prec)
    for B in (1, 200):
        tr.step(torch.randn(B, 34, generator=gen).to(dev), (torch.rand(B, 8, generator=gen) * 2 - 1).to(dev))
    crit = [make_critic(port.init_critic_params(10 + i)) for i in range(3)]
    off = [0, 129, 129, 400]
    obs, act = torch.randn(400, 29, generator=gen).to(dev), (torch.rand(400, 8, generator=gen) * 2 - 1).to(dev)
    q_action_ascent_segments(crit, obs, act.clone(), off, iters=3, precision=prec)
    crit[0].precision = prec
    crit[0].requires_grad_(False)
    crit[0].get_q1_q2(obs[:77], act[:77])
    crit[0].get_q_min(obs[:77], act[:77].clone().requires_grad_(True)).sum().backward()
c, ct = make_critic(port.init_critic_params(20)), make_critic(port.init_critic_params(21))
B = 150
critic_loss_and_grads(c, ct, torch.randn(B, 29, generator=gen).to(dev), torch.rand(B, 8, generator=gen).to(dev),
                      torch.randn(B, 29, generator=gen).to(dev), torch.rand(B, 8, generator=gen).to(dev),
                      torch.rand(B, 1, generator=gen).to(dev), torch.zeros(B, 1).to(dev), 0.99)
critic_loss_and_grads(c, ct, torch.randn(B, 29, generator=gen).to(dev), torch.rand(B, 8, generator=gen).to(dev),
                      torch.randn(B, 29, generator=gen).to(dev), torch.rand(B, 8, generator=gen).to(dev),
                      torch.rand(B, 1, generator=gen).to(dev), torch.zeros(B, 1).to(dev), 0.99, precision="bf16")
for prec in ("fp32", "bf16"):
    rnd = RNDModel(69, precision=prec).to(dev)
    x = torch.randn(70, 69, generator=gen).to(dev)
    rnd.novelty(x); rnd.loss_and_grads(x); rnd(x)
buf = DiffusionReplayBuffer(1000, 29, 8, device=dev)
for tid in range(3):
    n = 20 + tid
    buf.add_to_buffer((torch.randn(n, 29).to(dev), torch.rand(n, 8).to(dev), torch.rand(n, 8).to(dev), torch.rand(n).to(dev),
                       torch.randn(n, 29).to(dev), (torch.rand(n) < 0.2).to(dev)), tid)
buf.update_target_action_dim([-1])
data, idx = buf.sample_batch(33, [0, 2], 1, device=dev)
buf.update_target_action(data[2] * 0.5, idx, 1)
add_embedding(data[0], torch.randn(5).to(dev))
buf.sample_groups([idx[:10], idx[10:]], embeddings=torch.randn(2, 5).to(dev))
torch.cuda.synchronize()
print("sanitize_small: all calls returned")
