"""CPU-side vs GPU-side time of one fused training step (bf16 path) and of its parts."""
import sys, time, torch
sys.path.insert(0, '.')
from oracle import port
from tests.util import make_policy
from ddiffpg_b200 import FusedActorTrainer
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 5
pol = make_policy(port.init_actor_params(1), T); tr = FusedActorTrainer(pol, precision="bf16")
s = torch.randn(B, 34, device='cuda'); a = torch.rand(B, 8, device='cuda'); n = torch.randn(B, 8, device='cuda'); t = torch.randint(0, T, (B,), device='cuda')
def gpu_cpu(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0 = time.perf_counter(); e0.record()
    for _ in range(reps): fn()
    e1.record(); c1 = time.perf_counter(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (c1 - c0) * 1e3 / reps
print("full step        gpu %.3f ms  cpu-submit %.3f ms" % gpu_cpu(lambda: tr.step(s, a, noise=n, timesteps=t)))
print("loss+grads only  gpu %.3f ms  cpu-submit %.3f ms" % gpu_cpu(lambda: pol._loss_and_grads(s, a, n, t, precision="bf16")))
pol.mark_dirty()
def pack(): pol.mark_dirty(); pol._packed("bf16")
print("pack only        gpu %.3f ms  cpu-submit %.3f ms" % gpu_cpu(pack))
