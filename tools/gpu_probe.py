"""Quick device timings of the three hot-path calls (development aid, not the bench)."""
import sys, time, torch
sys.path.insert(0, '.')
from oracle import port
from tests.util import make_policy, make_critic
from ddiffpg_b200 import update_target_action

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

print(torch.cuda.get_device_name(0))
p = port.init_actor_params(1)
for prec in sys.argv[1:] or ["fp32"]:
    for T in (5,):
        pol = make_policy(p, T, precision=prec)
        for B in (256, 4096, 65536):
            s = torch.randn(B, 34, device='cuda'); n = torch.randn(T, B, 8, device='cuda')
            ms = timeit(lambda: pol.get_actions(s, noise=n))
            print(f"H1 {prec} T={T} B={B}: {ms:.3f} ms  {B/ms*1e3:.3e} actions/s  {B*6725632/ms/1e9:.2f} TFLOP/s(alg)")
cp = port.init_critic_params(2)
cri = make_critic(cp)
for B in (256, 4096, 65536):
    o = torch.randn(B, 29, device='cuda'); a = torch.rand(B, 8, device='cuda')
    ms = timeit(lambda: update_target_action(o, a, cri), n=3, warm=1)
    print(f"H2 fp32 B={B}: {ms:.3f} ms  {B/ms*1e3:.3e} states/s")
pol = make_policy(p, 5)
for B in (4096, 65536):
    s = torch.randn(B, 34, device='cuda'); a = torch.rand(B, 8, device='cuda'); n = torch.randn(B, 8, device='cuda')
    t = torch.randint(0, 5, (B,), device='cuda')
    ms = timeit(lambda: pol._loss_and_grads(s, a, n, t), n=3, warm=1)
    print(f"H3 fp32 fwd+bwd B={B}: {ms:.3f} ms  {B/ms*1e3:.3e} rows/s")
