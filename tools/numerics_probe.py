"""Probe: fp32-vs-fp64 noise floor and emulated bf16 tensor-path error of the H1 chain (CPU)."""
import sys, torch, math
import torch.nn.functional as F
sys.path.insert(0, '.')
from oracle import port
from oracle.ddpm import ddpm_step_constants

def bf(x): return x.to(torch.bfloat16).to(torch.float32)

def chain_emul(p, state, noise, T, mode):
    """mode: dict(x_split, s_split, head_fp32, w0_fp32)"""
    D=256
    cst = ddpm_step_constants(T)
    W0 = p['net.mlp.0.weight']; b0 = p['net.mlp.0.bias']
    W1,b1 = p['net.mlp.2.weight'], p['net.mlp.2.bias']
    W2,b2 = p['net.mlp.4.weight'], p['net.mlp.4.bias']
    W3,b3 = p['net.mlp.6.weight'], p['net.mlp.6.bias']
    S = state.shape[1]
    x = noise[0].clone()
    W0s, W0x = W0[:, D:D+S], W0[:, D+S:]
    def split_mm(a, W, split):
        Wb = bf(W)
        if split == 0: return bf(a) @ Wb.t()
        ah = bf(a); al = bf(a - ah)
        if split == 1: return ah @ Wb.t() + al @ Wb.t()
        Wl = bf(W - Wb)   # split==2: weights split too (3 products)
        return ah @ Wb.t() + al @ Wb.t() + ah @ Wl.t()
    for j, t in enumerate(range(T-1, -1, -1)):
        tt = torch.full((1,), float(t))
        temb = port.time_mlp(p, tt, D)
        tb0 = temb @ W0[:, :D].t() + b0     # fp32 table
        if mode.get('l0_fp32'):
            z = state @ W0s.t() + x @ W0x.t() + tb0
        else:
            z = split_mm(state, W0s, mode.get('s_split',0)) + split_mm(x, W0x, mode.get('x_split',0)) + tb0
        h = F.mish(z)
        h = F.mish(split_mm(h, W1, mode.get('h_split',0)) + b1)
        h = F.mish(split_mm(h, W2, mode.get('h_split',0)) + b2)
        if mode.get('head_fp32'): eps = h @ W3.t() + b3
        else: eps = split_mm(h, W3, mode.get('head_split',0)) + b3
        ce, ci, cx0, cxt, sg = cst[t]
        x0 = ((x - ce*eps) * ci).clamp(-1, 1)
        x = cx0*x0 + cxt*x
        if t > 0: x = x + sg*noise[j+1]
    return x

def report(name, a, ref):
    d = (a-ref).abs()
    rel = d / ref.abs().clamp_min(1e-3)
    print(f"{name:38s} max_abs={d.max():.3e} mean_abs={d.mean():.3e} frac>1e-2={(d>1e-2).float().mean():.4f} frac>1e-3={(d>1e-3).float().mean():.4f} allclose(1e-4,1e-5)={torch.allclose(a,ref,rtol=1e-4,atol=1e-5)}")

for T, B, scale in ((5, 2048, 1.0), (5, 2048, 2.5), (20, 1024, 1.0), (100, 256, 1.0)):
    g = torch.Generator().manual_seed(7)
    p = port.init_actor_params(3, scale=scale)
    state = torch.randn(B, 34, generator=g); noise = torch.randn(T, B, 8, generator=g)
    a32 = port.actor_sample(p, state, noise, T)
    p64 = port.cast_params(p, torch.float64)
    a64 = port.actor_sample(p64, state.double(), noise.double(), T).float()
    print(f"--- T={T} B={B} scale={scale} sat={(a64.abs()>=1).float().mean():.2f}")
    report("fp32 oracle vs fp64", a32, a64)
    report("fp32 hoisted-algebra emul vs fp64", chain_emul(p, state, noise, T, dict(l0_fp32=1, head_fp32=1, h_split=2)), a64)
    report("bf16 plain", chain_emul(p, state, noise, T, {}), a64)
    report("bf16 + x hi/lo", chain_emul(p, state, noise, T, dict(x_split=1)), a64)
    report("bf16 + x,s hi/lo", chain_emul(p, state, noise, T, dict(x_split=1, s_split=1)), a64)
    report("bf16 + x,s hi/lo + head fp32", chain_emul(p, state, noise, T, dict(x_split=1, s_split=1, head_fp32=1)), a64)
    report("bf16 + l0 fp32 + head fp32", chain_emul(p, state, noise, T, dict(l0_fp32=1, head_fp32=1)), a64)
    report("bf16 + l0 fp32 + head split1", chain_emul(p, state, noise, T, dict(l0_fp32=1, head_split=1)), a64)
    report("all split1 (act hi/lo)", chain_emul(p, state, noise, T, dict(x_split=1, s_split=1, h_split=1, head_split=1)), a64)
    report("all split2 (3-product)", chain_emul(p, state, noise, T, dict(x_split=2, s_split=2, h_split=2, head_split=2)), a64)
