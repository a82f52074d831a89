"""Which layer's bf16 rounding dominates the error of eps_hat (first step) and of the final action?"""
import sys, torch, itertools
import torch.nn.functional as F
sys.path.insert(0, '.')
from oracle import port
from oracle.ddpm import ddpm_step_constants
def bf(x): return x.to(torch.bfloat16).to(torch.float32)
def mm(a, W, mode):
    # mode: 0 = bf16 x bf16 ; 1 = act hi/lo x bf16 W ; 2 = act bf16 x W hi/lo ; 3 = both split (3 products) ; 4 = fp32
    if mode == 4: return a @ W.t()
    Wb = bf(W); ab = bf(a)
    if mode == 0: return ab @ Wb.t()
    if mode == 1: return ab @ Wb.t() + bf(a - ab) @ Wb.t()
    if mode == 2: return ab @ Wb.t() + ab @ bf(W - Wb).t()
    return ab @ Wb.t() + bf(a - ab) @ Wb.t() + ab @ bf(W - Wb).t()
def chain(p, state, noise, T, modes):
    D=256; cst = ddpm_step_constants(T)
    W0,b0 = p['net.mlp.0.weight'], p['net.mlp.0.bias']
    W1,b1 = p['net.mlp.2.weight'], p['net.mlp.2.bias']
    W2,b2 = p['net.mlp.4.weight'], p['net.mlp.4.bias']
    W3,b3 = p['net.mlp.6.weight'], p['net.mlp.6.bias']
    S = state.shape[1]; x = noise[0].clone(); eps_first=None
    for j, t in enumerate(range(T-1, -1, -1)):
        temb = port.time_mlp(p, torch.full((1,), float(t)), D)
        tb0 = temb @ W0[:, :D].t() + b0
        z = mm(torch.cat([state, x], 1), W0[:, D:], modes[0]) + tb0
        h = F.mish(z); h = F.mish(mm(h, W1, modes[1]) + b1); h = F.mish(mm(h, W2, modes[2]) + b2)
        eps = mm(h, W3, modes[3]) + b3
        if eps_first is None: eps_first = eps
        ce, ci, cx0, cxt, sg = cst[t]
        x0 = ((x - ce*eps) * ci).clamp(-1, 1); x = cx0*x0 + cxt*x
        if t > 0: x = x + sg*noise[j+1]
    return x, eps_first
T=5; B=16384
g = torch.Generator().manual_seed(11)
p = port.init_actor_params(81)
state = torch.randn(B, 34, generator=g); noise = torch.randn(T, B, 8, generator=g)
ref, eref = chain(p, state, noise, T, (4,4,4,4))
def rep(name, modes):
    a, e = chain(p, state, noise, T, modes)
    d = (a-ref).abs(); de = (e-eref).abs()
    print(f"{name:34s} eps: max {de.max():.2e} rms {de.pow(2).mean().sqrt():.2e} | action: max {d.max():.2e} rms {d.pow(2).mean().sqrt():.2e} n>1e-2 {int((d>1e-2).sum())} n>5e-3 {int((d>5e-3).sum())}")
rep("all bf16", (0,0,0,0))
rep("L0 exact", (4,0,0,0)); rep("L1 exact", (0,4,0,0)); rep("L2 exact", (0,0,4,0)); rep("L3 exact", (0,0,0,4))
rep("L3 act-split", (0,0,0,1)); rep("L3 both split", (0,0,0,3)); rep("L2+L3 both split", (0,0,3,3))
rep("L1 W-split", (0,2,0,0)); rep("L1 both", (0,3,0,0)); rep("L0 split3", (3,0,0,0))
rep("L0,L2,L3 split3", (3,0,3,3)); rep("all split3", (3,3,3,3)); rep("all act-split", (1,1,1,1)); rep("all W-split", (2,2,2,2))
print("--- first-step-only variants (modes per step: first, rest)")
def chain2(p, state, noise, T, modes_first, modes_rest):
    D=256; cst = ddpm_step_constants(T)
    W0,b0 = p['net.mlp.0.weight'], p['net.mlp.0.bias']; W1,b1 = p['net.mlp.2.weight'], p['net.mlp.2.bias']
    W2,b2 = p['net.mlp.4.weight'], p['net.mlp.4.bias']; W3,b3 = p['net.mlp.6.weight'], p['net.mlp.6.bias']
    x = noise[0].clone()
    for j, t in enumerate(range(T-1, -1, -1)):
        modes = modes_first if j == 0 else modes_rest
        temb = port.time_mlp(p, torch.full((1,), float(t)), D)
        tb0 = temb @ W0[:, :D].t() + b0
        z = mm(torch.cat([state, x], 1), W0[:, D:], modes[0]) + tb0
        h = F.mish(z); h = F.mish(mm(h, W1, modes[1]) + b1); h = F.mish(mm(h, W2, modes[2]) + b2)
        eps = mm(h, W3, modes[3]) + b3
        ce, ci, cx0, cxt, sg = cst[t]
        x0 = ((x - ce*eps) * ci).clamp(-1, 1); x = cx0*x0 + cxt*x
        if t > 0: x = x + sg*noise[j+1]
    return x
def rep2(name, mf, mr):
    a = chain2(p, state, noise, T, mf, mr); d = (a-ref).abs()
    print(f"{name:44s} action: max {d.max():.2e} rms {d.pow(2).mean().sqrt():.2e} n>1e-2 {int((d>1e-2).sum())} n>5e-3 {int((d>5e-3).sum())}")
rep2("first exact, rest bf16", (4,4,4,4), (0,0,0,0))
rep2("first split3, rest bf16", (3,3,3,3), (0,0,0,0))
rep2("first W-split, rest bf16", (2,2,2,2), (0,0,0,0))
rep2("first L0,L2,L3 split3 + L1 W-split, rest bf16", (3,2,3,3), (0,0,0,0))
rep2("first L0,L2,L3 split3, rest bf16", (3,0,3,3), (0,0,0,0))
rep2("first bf16, rest exact", (0,0,0,0), (4,4,4,4))
rep2("first two exact, rest bf16", (4,4,4,4), (0,0,0,0))
print("--- fp16 operand variants")
def hf(x): return x.to(torch.float16).to(torch.float32)
def mm16(a, W, mode):
    if mode == 5: return hf(a) @ hf(W).t()
    return mm_orig(a, W, mode)
mm_orig = mm
def rep3(name, mf, mr):
    global mm
    mm = mm16
    a = chain2(p, state, noise, T, mf, mr); d = (a-ref).abs()
    mm = mm_orig
    print(f"{name:44s} action: max {d.max():.2e} rms {d.pow(2).mean().sqrt():.2e} n>1e-2 {int((d>1e-2).sum())} n>5e-3 {int((d>5e-3).sum())} n>2e-3 {int((d>2e-3).sum())}")
rep3("first fp16, rest bf16", (5,5,5,5), (0,0,0,0))
rep3("all fp16", (5,5,5,5), (5,5,5,5))
rep3("first two fp16, rest bf16", (5,5,5,5), (0,0,0,0))
for seed in (1,2,3):
    g = torch.Generator().manual_seed(100+seed)
    p = port.init_actor_params(90+seed)
    state = torch.randn(B, 34, generator=g)*1.5; noise = torch.randn(T, B, 8, generator=g)
    ref = chain2(p, state, noise, T, (4,4,4,4), (4,4,4,4))
    rep3(f"seed {seed}: all bf16", (0,0,0,0), (0,0,0,0))
    rep3(f"seed {seed}: first fp16, rest bf16", (5,5,5,5), (0,0,0,0))
