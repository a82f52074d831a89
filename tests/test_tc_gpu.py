"""GPU tests of the bf16 tensor-core path: the tcgen05/TMA building blocks (self-test GEMM) and the fused
sampler against the fp32 oracle at the separately stated bf16 bound (1e-2 absolute on actions in [-1, 1])."""
import ctypes

import pytest
import torch

from oracle import port
from tests.util import actor_params_for, load_golden, make_policy

pytestmark = pytest.mark.gpu

BF16_ATOL = 1e-2


def _dev(x):
    return torch.as_tensor(x).to("cuda")


@pytest.mark.parametrize("N,K", [(256, 64), (256, 256), (128, 128), (16, 256), (64, 512)])
def test_tcgen05_blocks_selftest_gemm(N, K):
    """C = A.B^T through manual SWIZZLE_128B A chunks + TMA B tiles + tcgen05.mma + TMEM loads."""
    from ddiffpg_b200 import _lib
    L = _lib.lib()
    fn = L.ddp_debug_tc_gemm
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    gen = torch.Generator().manual_seed(N + K)
    A = torch.randn(128, K, generator=gen).to(torch.bfloat16).cuda()
    Bm = torch.randn(N, K, generator=gen).to(torch.bfloat16).cuda()
    C = torch.zeros(128, N, device="cuda")
    _lib.check(fn(A.data_ptr(), Bm.data_ptr(), C.data_ptr(), N, K, _lib.stream_ptr()), "ddp_debug_tc_gemm")
    torch.cuda.synchronize()
    ref = A.float() @ Bm.float().t()
    err = (C - ref).abs().max().item()
    assert err < 1e-3 * K ** 0.5, f"max |C - A.B^T| = {err:.3e}"


@pytest.mark.parametrize("name", ["h1_T5_B16", "h1_T20_B8"])
def test_sampler_bf16_reference_fixture(name):
    g = load_golden(name)
    pol = make_policy(actor_params_for(g), int(g["T"]), precision="bf16")
    out = pol.get_actions(_dev(g["state"]), noise=_dev(g["noise"])).cpu()
    err = (out - torch.from_numpy(g["action"])).abs()
    assert err.max().item() <= BF16_ATOL, f"max abs err {err.max().item():.3e}"
    assert out.abs().max().item() <= 1.0


@pytest.mark.parametrize("B,T", [(1, 5), (100, 5), (128, 5), (129, 5), (1000, 5), (5000, 5), (300, 20)])
def test_sampler_bf16_vs_oracle(B, T):
    gen = torch.Generator().manual_seed(500 + B + T)
    p = port.init_actor_params(81)
    state = torch.randn(B, 34, generator=gen)
    noise = torch.randn(T, B, 8, generator=gen)
    ref = port.actor_sample(p, state, noise, T)
    pol = make_policy(p, T, precision="bf16")
    out = pol.get_actions(_dev(state), noise=_dev(noise)).cpu()
    err = (out - ref).abs()
    assert err.max().item() <= BF16_ATOL, f"B={B}: max abs err {err.max().item():.3e}"
    assert err.mean().item() <= 5e-4, f"B={B}: mean abs err {err.mean().item():.3e}"
    # fp32 path on the same inputs agrees with the tensor path to the same bound
    out32 = pol.get_actions(_dev(state), noise=_dev(noise), precision="fp32").cpu()
    assert (out32 - out).abs().max().item() <= BF16_ATOL


@pytest.mark.parametrize("h", [256, 512])
def test_sampler_bf16_width_sweep(h):
    B, T = 300, 5
    gen = torch.Generator().manual_seed(600 + h)
    p = port.init_actor_params(82, h=h)
    state = torch.randn(B, 34, generator=gen)
    noise = torch.randn(T, B, 8, generator=gen)
    ref = port.actor_sample(p, state, noise, T)
    out = make_policy(p, T, precision="bf16", hidden=(h, h // 2, h // 4)).get_actions(_dev(state), noise=_dev(noise)).cpu()
    assert (out - ref).abs().max().item() <= BF16_ATOL


def test_sampler_bf16_large_batch_properties():
    """BASELINE size (65,536 rows): size-independent checks -- range, determinism, row independence, and a
    4,096-row slice against the oracle."""
    B, T = 65536, 5
    gen = torch.Generator().manual_seed(7)
    p = port.init_actor_params(83)
    state = torch.randn(B, 34, generator=gen)
    noise = torch.randn(T, B, 8, generator=gen)
    pol = make_policy(p, T, precision="bf16")
    sd, nd = _dev(state), _dev(noise)
    out = pol.get_actions(sd, noise=nd)
    assert torch.isfinite(out).all() and out.abs().max().item() <= 1.0
    assert torch.equal(out, pol.get_actions(sd, noise=nd))                      # deterministic
    sl = slice(30000, 34096)
    part = pol.get_actions(sd[sl].contiguous(), noise=nd[:, sl].contiguous())
    assert torch.equal(part, out[sl])                                           # rows independent of tiling
    ref = port.actor_sample(p, state[sl], noise[:, sl], T)
    assert (part.cpu() - ref).abs().max().item() <= BF16_ATOL


# ------------------------------------------------------------------------------------------ H3 tensor path
@pytest.mark.parametrize("B,T", [(64, 5), (700, 5), (4096, 5), (1000, 20)])
def test_train_bf16_grads_vs_oracle(B, T):
    """tcgen05 GEMM path of the training step: loss and all 12 gradients against the fp32 oracle at the bf16
    bound (relative L2 error of the flat gradient <= 2e-2, of every weight matrix <= 3e-2)."""
    gen = torch.Generator().manual_seed(900 + B)
    p = port.init_actor_params(84)
    state = torch.randn(B, 34, generator=gen)
    action = torch.rand(B, 8, generator=gen) * 2 - 1
    noise = torch.randn(B, 8, generator=gen)
    ts = torch.randint(0, T, (B,), generator=gen)
    l_ref, g_ref = port.actor_loss_and_grads(p, state, action, noise, ts, T)
    pol = make_policy(p, T)
    pol.train_precision = "bf16"
    loss = pol.get_loss(_dev(state), _dev(action), noise=_dev(noise), timesteps=_dev(ts))
    loss.backward()
    assert abs(loss.item() - l_ref.item()) <= 2e-3 * l_ref.item()
    got = torch.cat([q.grad.reshape(-1) for _, q in pol.named_parameters()]).cpu()
    ref = torch.cat([g_ref[k].reshape(-1) for k in port.ACTOR_KEYS])
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel <= 2e-2, f"flat gradient relative L2 error {rel:.3e}"
    for k, q in pol.named_parameters():
        r = g_ref[k]
        e = ((q.grad.cpu() - r).norm() / r.norm().clamp_min(1e-12)).item()
        assert e <= 3e-2, f"{k}: relative L2 error {e:.3e}"


def test_fused_trainer_bf16_tracks_fp32():
    """Three fused training steps on the tensor path stay close to the fp32 oracle trajectory."""
    from ddiffpg_b200 import FusedActorTrainer
    B, T = 2048, 5
    gen = torch.Generator().manual_seed(77)
    p = port.init_actor_params(85)
    args = (torch.randn(B, 34, generator=gen), torch.rand(B, 8, generator=gen) * 2 - 1,
            torch.randn(B, 8, generator=gen), torch.randint(0, T, (B,), generator=gen))
    pol = make_policy(p, T)
    tr = FusedActorTrainer(pol, precision="bf16")
    st, cur = None, p
    for _ in range(3):
        loss, gnorm = tr.step(_dev(args[0]), _dev(args[1]), noise=_dev(args[2]), timesteps=_dev(args[3]))
        l_ref, n_ref, cur, st = port.adamw_train_step(cur, *args, T, opt_state=st)
        assert abs(loss.item() - l_ref.item()) <= 3e-3 * l_ref.item()
        assert abs(gnorm.item() / n_ref.item() - 1) <= 2e-2
