"""Where does tcgen05.mma with M = 64 (cta_group::1) put D[r][n] in TMEM?  D[r][n] = (r+1) + 128 (n+1), all 128 lanes dumped."""
import ctypes, sys, torch
sys.path.insert(0, '.')
from ddiffpg_b200 import _lib
L = _lib.lib()
fn = L.ddp_debug_tc_m64_probe
fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]; fn.restype = ctypes.c_int
for N in (64, 256):
    A = torch.zeros(64, 64); A[:, 0] = torch.arange(1, 65).float(); A[:, 1] = 128.0
    B = torch.zeros(N, 64); B[:, 0] = 1.0; B[:, 1] = torch.arange(1, N + 1).float()
    raw = torch.zeros(128 * N + 512, device='cuda')
    Ad, Bd = A.cuda().bfloat16().contiguous(), B.cuda().bfloat16().contiguous()      # keep the device copies alive
    rc = fn(Ad.data_ptr(), Bd.data_ptr(), raw.data_ptr(), N, None)
    torch.cuda.synchronize()
    assert rc == 0, L.ddp_last_error()
    frag = raw[128 * N:].cpu().round().long().reshape(4, 32, 4)
    raw = raw[:128 * N].reshape(128, N).cpu()
    used = (raw >= 0)
    lanes = sorted(set(torch.nonzero(used)[:, 0].tolist()))
    print(f"N={N}: {len(lanes)} lanes hold data: {lanes}")
    v = raw.round().long()
    r = v % 128 - 1; n = v // 128 - 1
    ok = True
    for lane in lanes:
        cols = torch.nonzero(used[lane])[:, 0]
        rows = sorted(set(r[lane][cols].tolist()))
        straight = n[lane][cols].tolist() == cols.tolist()
        ok &= len(rows) == 1 and rows[0] == (lane % 32) + 16 * (lane // 32) and straight and len(cols) == N
        if lane in (0, 1, 15, 32, 33, 64, 96, 111):
            print(f"  lane {lane:3d}: row {rows}  column c holds n = c: {straight}  ({len(cols)} columns)")
    print(f"  layout D[r][n] -> lane (r % 16) + 32 (r // 16), column n: {'confirmed' if ok else 'NOT confirmed'}")
    # 16x256b.x1: which D[r][n] does register i of thread t of warp w hold?
    fr, fn_ = frag % 128 - 1, frag // 128 - 1
    exp_ok = True
    for w in range(4):
        for t in range(32):
            want = [(16 * w + t // 4, 2 * (t % 4)), (16 * w + t // 4, 2 * (t % 4) + 1), (16 * w + t // 4 + 8, 2 * (t % 4)), (16 * w + t // 4 + 8, 2 * (t % 4) + 1)]
            got = [(int(fr[w, t, i]), int(fn_[w, t, i])) for i in range(4)]
            exp_ok &= got == want
    print("  16x256b.x1 warp 1:", {t: [(int(fr[1, t, i]), int(fn_[1, t, i])) for i in range(4)] for t in (0, 1, 4, 31)})
    print(f"  16x256b.x1 = mma.m16n8 C fragment (thread t: rows t/4 and t/4+8 of the quarter, columns 2(t%4), +1): {'confirmed' if exp_ok else 'NOT confirmed'}")
