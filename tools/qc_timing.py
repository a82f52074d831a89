"""Per-phase cycle breakdown of the fused critic chain kernel (CTA 0), via the ddp_debug_qc_timing hook."""
import ctypes, sys, torch
sys.path.insert(0, '.')
from oracle import port
from tests.util import make_critic
from ddiffpg_b200 import _lib
from ddiffpg_b200.algo import q_action_ascent_segments
L = _lib.lib()
L.ddp_debug_qc_timing.argtypes = [ctypes.c_void_p]; L.ddp_debug_qc_timing.restype = None
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
iters = 20
c = make_critic(port.init_critic_params(3)); c.precision = "bf16"
obs = torch.randn(B, 29, device='cuda'); act = torch.rand(B, 8, device='cuda') * 2 - 1
run = lambda: q_action_ascent_segments([c], obs, act.clone(), [0, B], iters=iters, precision="bf16")
for _ in range(2): run()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record(); run(); ev[1].record(); torch.cuda.synchronize()
print(f"ascent B={B}: {ev[0].elapsed_time(ev[1]):.3f} ms for {iters} iterations")
buf = torch.zeros(256, dtype=torch.int64, device='cuda')
L.ddp_debug_qc_timing(buf.data_ptr())
run(); torch.cuda.synchronize()
L.ddp_debug_qc_timing(None)
names = ["A0 chunk", "wait F1", "drain F1 (8)", "F2 wait+drain (4)", "F3 wait+drain (2)", "logits", "wait B4", "B4+B3 drains", "B2 wait+drain (8)", "Ba read"]
import os
tiles = (B + 127) // 128; grid = min(tiles, int(os.environ.get('DDP_QC_GRID', 148))); my = len(range(0, tiles, grid))
tot = int(buf[:10].sum())
print(f"CTA0 ran {my} tiles x {iters} launches; total {tot} clk; per tile {tot/(my*iters):.0f} clk")
for nm, v in zip(names, buf.tolist()): print(f"  {nm:22s} {v/(my*iters):9.0f} clk/tile  {100*v/max(tot,1):5.1f}%")
fn = ["tmem wait a (+issue b)", "a_empty wait", "emit a", "tmem wait b (+issue a')", "emit b (+d prefetch)", "publish (fence, arrive)"]
for off, nm in ((16, "forward drains"), (24, "backward drains")):
    v = buf[off:off + 7].tolist(); n = max(v[6], 1)
    print(f"{nm}: {n} chunks, {sum(v[:6])/n:.0f} clk per chunk")
    for k in range(6): print(f"    {fn[k]:28s} {v[k]/n:7.0f} clk")
per = buf[32:32 + 148].tolist()
if any(per):
    import statistics
    a4 = [v for i, v in enumerate(per) if v and i < tiles - 3 * grid]; a3 = [v for i, v in enumerate(per) if v and i >= tiles - 3 * grid]
    print(f"per-CTA work cycles per pass (x{iters} passes, before the barrier wait): 4-tile CTAs n={len(a4)} min {min(a4)/iters/1e3:.0f}k med {statistics.median(a4)/iters/1e3:.0f}k max {max(a4)/iters/1e3:.0f}k;"
          f" 3-tile CTAs n={len(a3)} min {min(a3)/iters/1e3:.0f}k med {statistics.median(a3)/iters/1e3:.0f}k max {max(a3)/iters/1e3:.0f}k")
    worst = sorted(range(148), key=lambda i: -per[i])[:8]
    print("slowest CTAs:", [(i, round(per[i] / iters / 1e3)) for i in worst])
