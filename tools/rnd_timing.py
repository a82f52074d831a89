"""N4 (RND / NovelD): novelty and predictor update, fp32 FMA path against the bf16 tensor path, at the reference's
update-batch sizes and at a large batch."""
import sys, torch
sys.path.insert(0, '.')
from oracle import port
from ddiffpg_b200 import RNDModel

def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(n): fn()
    ev[1].record(); torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n

p = port.init_rnd_params(7)
for B in (4096, 8192, 131072):
    x = torch.randn(B, 69, device="cuda")
    row = [f"rows {B:7d}"]
    for prec in ("fp32", "bf16"):
        m = RNDModel(69, precision=prec); m.load_state_dict(p); m = m.to("cuda")
        row.append(f"{prec}: novelty {timed(lambda: m.novelty(x)):.3f} ms, update {timed(lambda: m.loss_and_grads(x)):.3f} ms")
    print("   ".join(row))
