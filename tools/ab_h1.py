"""A/B of sampler builds: for the library in DDP_LIB_PATH (default: the in-tree one) print parity against the oracle,
the per-phase cycles of CTA 0 and the kernel-only time at 65 536 / 75 776 rows (CUDA events, L2 flushed between calls).
  DDP_LIB_PATH=$PWD/ddiffpg_b200/libv_x.so python tools/ab_h1.py [T] [h]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import port                              # noqa: E402
from tests.util import make_policy                   # noqa: E402
from ddiffpg_b200 import _lib                        # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 5
h = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
tag = os.path.basename(os.environ.get("DDP_LIB_PATH", "in-tree"))
L = _lib.lib()
L.ddp_debug_tc_timing.argtypes = [ctypes.c_void_p]
L.ddp_debug_tc_timing.restype = None
gen = torch.Generator().manual_seed(5)
p = port.init_actor_params(83, h=h)
pol = make_policy(p, T, precision="bf16", hidden=(h, h // 2, h // 4))
n = 4096
st, nz = torch.randn(n, 34, generator=gen), torch.randn(T, n, 8, generator=gen)
ref = port.actor_sample(p, st, nz, T)
out = pol.get_actions(st.cuda(), noise=nz.cuda()).cpu()
err = (out - ref).abs()
line = f"{tag:22s} T={T} h={h} err max {err.max():.2e} mean {err.mean():.2e} |"
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B in (65536, 75776):
    s = torch.randn(B, 34, device="cuda")
    z = torch.randn(T, B, 8, device="cuda")
    for _ in range(30 if B == 65536 else 5):     # a fresh box needs a few ms of load before its clocks settle
        pol.get_actions(s, noise=z)
    if B == 65536:
        buf = torch.zeros(16, dtype=torch.int64, device="cuda")
        L.ddp_debug_tc_timing(buf.data_ptr())
        pol.get_actions(s, noise=z)
        torch.cuda.synchronize()
        L.ddp_debug_tc_timing(None)
        raw = buf.tolist()
        units = max(raw[14], 1)
        v = [x / units for x in raw[:7]]
        line += f" tile-step {sum(v):6.0f} clk (L0 {v[0]:5.0f} w1 {v[1]:4.0f} d1 {v[2]:5.0f} w2 {v[3]:4.0f} d2 {v[4]:4.0f} w3 {v[5]:4.0f} hd {v[6]:4.0f}) |"
        line += f" CTA0: {units} units, {raw[12]} clk in {raw[13] / 1e3:.1f} us = {raw[12] / max(raw[13], 1) * 1e3:.0f} MHz, outside steps {(raw[12] - sum(raw[:7])) / max(raw[12], 1):.1%}; {raw[10]} jobs: load+in0 {raw[8] / max(raw[10], 1):.0f} clk, state partial {raw[9] / max(raw[10], 1):.0f} clk per job |"
    meds = []
    for _ in range(3):
        ts = []
        for _ in range(20):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pol.get_actions(s, noise=z)
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        meds.append(sorted(ts)[len(ts) // 2])
    ms = min(meds)
    flops = 2.0 * (T * (0.625 * h * h + 10 * h) + 34 * h) * B
    line += f" B={B}: {ms:.4f} ms {flops / ms / 1e9 / 1622.3:.3f} |"
print(line, flush=True)
