"""Max / quantile action error of the tensor-path sampler against the oracle over a FULL 65 536-row batch.
  [DDP_LIB_PATH=...] python tools/err_full.py [T] [h] [seeds]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import port
from tests.util import make_policy
T = int(sys.argv[1]) if len(sys.argv) > 1 else 5
h = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
seeds = int(sys.argv[3]) if len(sys.argv) > 3 else 1
tag = os.path.basename(os.environ.get("DDP_LIB_PATH", "in-tree"))
torch.set_num_threads(os.cpu_count())
worst = 0.0
for sd in range(seeds):
    B = 65536
    gen = torch.Generator().manual_seed(100 + sd)
    p = port.init_actor_params(83 + sd, h=h)
    st, nz = torch.randn(B, 34, generator=gen), torch.randn(T, B, 8, generator=gen)
    ref = port.actor_sample(p, st, nz, T)
    pol = make_policy(p, T, precision="bf16", hidden=(h, h // 2, h // 4))
    out = pol.get_actions(st.cuda(), noise=nz.cuda()).cpu()
    err = (out - ref).abs().flatten()
    q = torch.quantile(err[:4000000], torch.tensor([0.999, 0.99999]))
    worst = max(worst, err.max().item())
    print(f"{tag:20s} T={T} h={h} seed {sd}: max {err.max():.3e} p99.9 {q[0]:.2e} p99.999 {q[1]:.2e} n>5e-3 {(err > 5e-3).sum().item()} n>1e-2 {(err > 1e-2).sum().item()} of {err.numel()}", flush=True)
print(f"{tag:20s} T={T} h={h}: worst over {seeds} seeds {worst:.3e}")
