"""GPU tests of the two tcgen05 GEMM kernels the H2/H3 tensor paths are assembled from, against torch on the
same bf16 operands (fp32 accumulation): every epilogue of the row GEMM (Mish + derivative, ELU, linear, the two
backward maskings), row groups with per-group weights, ragged sizes, and the MN-major dW GEMM with row splits."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
c_void, c_int, c_long, c_size = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_size_t


def _load():
    from ddiffpg_b200 import _lib
    L = _lib.lib()
    L.ddp_debug_row_gemm.restype = c_int
    L.ddp_debug_row_gemm.argtypes = [c_void, c_int, c_void, c_int, c_long, c_int, c_int, c_int, c_void, c_void, c_void,
                                     c_void, c_void, c_int, ctypes.POINTER(c_long), c_size, c_size, c_void, c_void,
                                     c_int, c_int, c_void]
    L.ddp_debug_row_gemm_nvalid.restype = c_int
    L.ddp_debug_row_gemm_nvalid.argtypes = [c_void, c_int, c_void, c_int, c_long, c_int, c_int, c_void, c_void, c_int, c_int, c_void]
    L.ddp_debug_dw_gemm.restype = c_int
    L.ddp_debug_dw_gemm.argtypes = [c_void, c_int, c_int, c_void, c_int, c_int, c_long, c_void, c_int, c_void, c_void]
    return L, _lib


def _ptr(t):
    return None if t is None else t.data_ptr()


def row_gemm(A, W, epi, bias=None, aux=None, groups=None, tbl=None, trow=None, in_place=False):
    L, _lib = _load()
    M, K = A.shape
    ng = 1 if groups is None else len(groups) - 1
    N = W.shape[-2]
    off = (c_long * (ng + 1))(*([0, M] if groups is None else groups))
    out_a = aux if in_place else torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
    out_d = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
    out_f = torch.zeros(M, N, device="cuda")
    rc = L.ddp_debug_row_gemm(_ptr(A), A.stride(0), _ptr(W), W.stride(-2), M, N, K, epi, _ptr(bias), _ptr(aux), _ptr(out_a),
                              _ptr(out_d), _ptr(out_f), ng, off, W.stride(0) if W.dim() == 3 else 0,
                              bias.stride(0) if bias is not None and bias.dim() == 2 else 0, _ptr(tbl), _ptr(trow),
                              tbl.stride(0) if tbl is not None else 0, tbl.shape[0] if tbl is not None else 0,
                              _lib.stream_ptr())
    _lib.check(rc, "ddp_debug_row_gemm")
    torch.cuda.synchronize()
    return out_a.float(), out_d.float(), out_f


def _mk(shape, gen, scale=1.0):
    return (torch.randn(*shape, generator=gen) * scale).to(torch.bfloat16).cuda()


def _close(got, ref, rtol, atol, what):
    err = (got - ref).abs()
    bad = err > atol + rtol * ref.abs()
    assert not bad.any(), f"{what}: {int(bad.sum())} bad, max err {err.max().item():.3e}"


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 512, 1024), (1000, 1024, 48), (77, 16, 256), (513, 128, 200)])
def test_row_gemm_linear(M, N, K):
    gen = torch.Generator().manual_seed(M + N + K)
    Kp = (K + 7) // 8 * 8
    A, W = _mk((M, Kp), gen), _mk((N, Kp), gen, K ** -0.5)
    A[:, K:] = 7.0          # columns beyond K must be ignored (the tensor map bounds the contraction)
    bias = torch.randn(N, generator=gen).cuda()
    _, _, out = row_gemm(A[:, :K] if K == Kp else A.as_strided((M, K), (Kp, 1)), W.as_strided((N, K), (Kp, 1)), 2, bias=bias)
    ref = A[:, :K].float() @ W[:, :K].float().t() + bias
    _close(out, ref, 1e-3, 1e-3, "linear")


def test_row_gemm_mish_forward_with_row_table():
    M, N, K, T = 700, 512, 64, 5
    gen = torch.Generator().manual_seed(1)
    A, W = _mk((M, K), gen), _mk((N, K), gen, K ** -0.5)
    tbl = torch.randn(T, N, generator=gen).cuda()
    trow = torch.randint(0, T, (M,), generator=gen).cuda()
    a, d, _ = row_gemm(A, W, 0, tbl=tbl, trow=trow)
    z = (A.float() @ W.float().t() + tbl[trow]).requires_grad_(True)
    y = F.mish(z)
    y.sum().backward()
    _close(a, y.detach(), 1e-2, 1e-2, "mish")
    _close(d, z.grad, 1e-2, 1e-2, "mish'")


def test_row_gemm_elu_forward_and_backward_masks():
    M, N, K = 400, 256, 512
    gen = torch.Generator().manual_seed(2)
    A, W = _mk((M, K), gen), _mk((N, K), gen, K ** -0.5)
    bias = torch.randn(N, generator=gen).cuda()
    a, _, _ = row_gemm(A, W, 1, bias=bias)
    z = A.float() @ W.float().t() + bias
    _close(a, F.elu(z), 1e-2, 1e-2, "elu")
    aux = _mk((M, N), gen)
    m, _, _ = row_gemm(A, W, 3, aux=aux)
    _close(m, (A.float() @ W.float().t()) * aux.float(), 1e-2, 2e-2, "mul_d")
    act = F.elu(_mk((M, N), gen).float()).to(torch.bfloat16)
    m2, _, _ = row_gemm(A, W, 4, aux=act)
    dz = torch.where(act.float() > 0, torch.ones_like(z), act.float() + 1)
    _close(m2, (A.float() @ W.float().t()) * dz, 1e-2, 2e-2, "mul_elu_d")


@pytest.mark.parametrize("M,N,K", [(1, 64, 64), (127, 128, 64), (129, 256, 128), (400, 512, 256), (1000, 1024, 512),
                                   (20000, 256, 64), (70001, 512, 512)])
@pytest.mark.parametrize("epi", [3, 4])
def test_row_gemm_backward_epilogue_tma_equals_direct(M, N, K, epi):
    """The backward maskings through the TMA-staged aux tile (one row group: aux in by TMA, gradient out by TMA store,
    in place) against the direct row-per-lane epilogue: same arithmetic, so bit-identical -- ragged last tiles (rows
    past M are clipped by the tensor map), one to four 256-column tiles, 64- / 128-column halves; and against torch."""
    L, _ = _load()
    dbg = L.ddp_debug_row_gemm_direct_aux
    dbg.argtypes, dbg.restype = [c_int], None
    gen = torch.Generator().manual_seed(M + N + K + epi)
    A, W = _mk((M, K), gen), _mk((N, K), gen, K ** -0.5)
    aux = _mk((M, N), gen) if epi == 3 else F.elu(_mk((M, N), gen).float()).to(torch.bfloat16)
    guard = torch.full((64, N), 3.0, dtype=torch.bfloat16, device="cuda")      # rows behind the matrix must stay untouched
    buf = torch.cat([aux, guard])
    tma, _, _ = row_gemm(A, W, epi, aux=buf[:M].clone() if False else buf[:M], in_place=True)
    assert torch.equal(buf[M:], guard)
    dbg(1)
    try:
        direct, _, _ = row_gemm(A, W, epi, aux=aux.clone(), in_place=True)
    finally:
        dbg(0)
    assert torch.equal(tma, direct)
    d = aux.float() if epi == 3 else torch.where(aux.float() > 0, torch.ones_like(aux, dtype=torch.float32), aux.float() + 1)
    _close(tma, (A.float() @ W.float().t()) * d, 1e-2, 2e-2 * K ** 0.5 / 8, "masked gradient")


@pytest.mark.parametrize("M,N,K", [(1, 64, 64), (127, 128, 64), (129, 256, 128), (400, 512, 64), (1000, 1024, 256), (70001, 512, 64)])
def test_row_gemm_elu_forward_tma_store_equals_staged(M, N, K):
    """EPI_ELU_FWD of a single row group written as SWIZZLE_128B boxes and shipped by TMA store, against the per-warp
    staged epilogue (bit-identical: same arithmetic) and against torch; rows behind the matrix stay untouched."""
    L, _lib = _load()
    dbg = L.ddp_debug_row_gemm_direct_aux
    dbg.argtypes, dbg.restype = [c_int], None
    gen = torch.Generator().manual_seed(3 * M + N + K)
    A, W = _mk((M, K), gen), _mk((N, K), gen, K ** -0.5)
    bias = torch.randn(N, generator=gen).cuda()
    buf = torch.full((M + 64, N), 3.0, dtype=torch.bfloat16, device="cuda")
    tma, _, _ = row_gemm(A, W, 1, bias=bias, aux=buf[:M], in_place=True)
    assert torch.equal(buf[M:], torch.full((64, N), 3.0, dtype=torch.bfloat16, device="cuda"))
    dbg(1)
    try:
        staged, _, _ = row_gemm(A, W, 1, bias=bias)
    finally:
        dbg(0)
    assert torch.equal(tma, staged)
    _close(tma, F.elu(A.float() @ W.float().t() + bias), 1e-2, 1e-2, "elu")


@pytest.mark.parametrize("M,N,nv", [(300, 64, 51), (1000, 128, 128), (77, 16, 8)])
def test_row_gemm_linear_partial_columns(M, N, nv):
    """EPI_LINEAR_F32 with fewer valid columns than the tile (the 51 atoms of the critic head): whole 16-column pieces
    go out as float4 stores, the ragged piece element by element; columns >= n_valid stay untouched."""
    L, _lib = _load()
    gen = torch.Generator().manual_seed(M + N + nv)
    K = 128
    A, W = _mk((M, K), gen), _mk((N, K), gen, K ** -0.5)
    bias = torch.randn(N, generator=gen).cuda()
    out_f = torch.full((M, N), -7.0, device="cuda")
    off = (c_long * 2)(0, M)
    rc = L.ddp_debug_row_gemm_nvalid(_ptr(A), K, _ptr(W), K, M, N, K, _ptr(bias), _ptr(out_f), N, nv, _lib.stream_ptr())
    _lib.check(rc, "ddp_debug_row_gemm_nvalid")
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t() + bias
    _close(out_f[:, :nv], ref[:, :nv], 1e-3, 1e-3, "linear")
    assert (out_f[:, nv:] == -7.0).all()


def test_row_gemm_groups_use_their_own_weights():
    sizes = [130, 0, 5, 300]
    off = [0]
    for s in sizes:
        off.append(off[-1] + s)
    M, N, K = off[-1], 128, 64
    gen = torch.Generator().manual_seed(3)
    A = _mk((M, K), gen)
    W = _mk((len(sizes), N, K), gen, K ** -0.5)
    bias = torch.randn(len(sizes), N, generator=gen).cuda()
    _, _, out = row_gemm(A, W, 2, bias=bias, groups=off)
    for g, s in enumerate(sizes):
        sl = slice(off[g], off[g + 1])
        _close(out[sl], A[sl].float() @ W[g].float().t() + bias[g], 1e-3, 1e-3, f"group {g}")


@pytest.mark.parametrize("R,N,K", [(64, 128, 64), (1000, 256, 512), (4096, 512, 1024), (333, 8, 256), (5, 1024, 64), (70000, 256, 128)])
def test_dw_gemm(R, N, K):
    L, _lib = _load()
    gen = torch.Generator().manual_seed(R + N)
    Np = max(8, (N + 7) // 8 * 8)
    dZ, X = _mk((R, Np), gen), _mk((R, K), gen)
    C = torch.zeros(N, K + 3, device="cuda")
    colmap = torch.arange(K, dtype=torch.int32).cuda() + 3
    colmap[0] = -1
    rc = L.ddp_debug_dw_gemm(_ptr(dZ), Np, N, _ptr(X), K, K, R, _ptr(C), K + 3, _ptr(colmap), _lib.stream_ptr())
    _lib.check(rc, "ddp_debug_dw_gemm")
    torch.cuda.synchronize()
    ref = dZ[:, :N].float().t() @ X.float()
    assert C[:, :3].abs().max().item() == 0.0 and True
    err = (C[:, 4:] - ref[:, 1:]).abs().max().item()
    assert err <= 2e-3 * R ** 0.5 + 1e-3, f"max err {err:.3e}"
