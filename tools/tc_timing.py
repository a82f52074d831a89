"""Per-phase cycle breakdown of the fused sampler (CTA 0), via the ddp_debug_tc_timing hook."""
import ctypes, sys, torch
sys.path.insert(0, '.')
from oracle import port
from tests.util import make_policy
from ddiffpg_b200 import _lib
L = _lib.lib()
L.ddp_debug_tc_timing.argtypes = [ctypes.c_void_p]; L.ddp_debug_tc_timing.restype = None
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 5
pol = make_policy(port.init_actor_params(1), T, precision="bf16")
s = torch.randn(B, 34, device='cuda'); n = torch.randn(T, B, 8, device='cuda')
for _ in range(3): pol.get_actions(s, noise=n)
buf = torch.zeros(16, dtype=torch.int64, device='cuda')
L.ddp_debug_tc_timing(buf.data_ptr())
pol.get_actions(s, noise=n); torch.cuda.synchronize()
L.ddp_debug_tc_timing(None)
names = ["L0+Mish (16 chunks)", "wait acc1", "drain acc1 (8)", "wait acc2", "drain acc2 (4)", "wait acc3", "head+step+bar"]
raw = buf.tolist(); units = max(raw[14], 1)
tot = int(buf[:7].sum())
print(f"B={B}: CTA0 ran {units} tile-steps; {tot} clk inside steps; per tile-step {tot/units:.0f} clk; whole job list {raw[12]} clk in {raw[13]/1e3:.1f} us ({raw[12]/max(raw[13],1)*1e3:.0f} MHz)")
for nm, v in zip(names, raw): print(f"  {nm:22s} {v/units:9.0f} clk/tile-step  {100*v/tot:5.1f}%")
