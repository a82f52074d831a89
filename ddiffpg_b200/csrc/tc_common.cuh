// sm_100a building blocks used by the tensor-core kernels: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld) and the K-major SWIZZLE_128B shared-memory operand layout.
// Bit layouts follow the PTX ISA "tcgen05 matrix / instruction descriptors" (cross-checked with
// cute/arch/mma_sm100_desc.hpp of the CUTLASS headers vendored in the image).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ddp {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    // the last operand is the suspend-time hint: the warp sleeps in hardware instead of spinning
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    return ok != 0;
}
// non-blocking probe (no suspend hint): has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded spin: a protocol bug must surface as a trap (reported by the next CUDA call), not as a hang
// of the GPU box.  ~2^28 polls is seconds; a healthy wait is microseconds.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) { asm volatile("trap;"); }
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy writes (st.shared) -> visible to the async proxy (TMA / tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tile load: coordinates (x = innermost element index, y = row index)
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* m, uint32_t bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(x), "r"(y) : "memory");
}

// 2-D tile store (shared -> global), bulk-group completion
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src_smem, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src_smem), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk stores of this thread have finished READING their shared-memory source
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------ TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {     // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {      // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ------------------------------------------------------------------------------------ UMMA descriptors
// Shared-memory matrix descriptor, K-major operand in the canonical SWIZZLE_128B layout:
// rows of 64 bf16 (128 B), 8-row groups of 1024 B; start address in 16-byte units; LBO unused (1);
// SBO = 1024 B between 8-row groups; version 1 (Blackwell) at bit 46; layout type 2 at bits 61-63.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address  [0,14)
    d |= (uint64_t)1 << 16;                               // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset [32,46)
    d |= (uint64_t)1 << 46;                               // descriptor version
    d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
    return d;
}
// Instruction descriptor for kind::f16: D = fp32, A = B = bf16 (format 1) or fp16 (format 0), both K-major,
// dense, no negate.
__host__ __device__ constexpr uint32_t make_idesc_16(int M, int N, bool f16) {
    return (1u << 4) | ((f16 ? 0u : 1u) << 7) | ((f16 ? 0u : 1u) << 10) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) { return make_idesc_16(M, N, false); }
// D[tmem] (+)= A[smem] . B[smem]^T, issued by ONE thread on behalf of the CTA
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// mbarrier arrives once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (lane i of the warp = TMEM lane
// 32*(warp%4)+i, encoded in bits 16+ of taddr by the caller)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------ operand layout
// Byte offset of element (row, col) of a [rows][64] bf16 K-major SWIZZLE_128B tile (base 1024-aligned):
// 16-byte chunk index (col/8) is XORed with (row % 8).
__host__ __device__ __forceinline__ uint32_t sw128_offset(int row, int col) {
    return (uint32_t)row * 128u + ((((uint32_t)col >> 3) ^ ((uint32_t)row & 7u)) << 4) + (((uint32_t)col & 7u) << 1);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);      // .x = lo (low 16 bits), .y = hi
    return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    // saturate to the finite fp16 range: an overflow must not become inf -> NaN inside the accumulators
    // (one F2FP.SATFINITE.F16.F32.PACK_AB; .x = lo in the low 16 bits)
    uint32_t d;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
template <bool F16> __device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    return F16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi);
}
template <bool F16> __device__ __forceinline__ uint16_t cvt16(float v) {
    if (F16) { __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)); return *reinterpret_cast<uint16_t*>(&h); }
    __nv_bfloat16 b = __float2bfloat16(v);
    return *reinterpret_cast<uint16_t*>(&b);
}

// one m16k16 A fragment (four 8x8 b16 matrices) from shared memory; lane l supplies the address of row
// (l & 7) + 8 * ((l >> 3) & 1), k offset 8 * (l >> 4)
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t smem_addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr));
}

// four 8x8 b16 matrices from the mma.sync C-fragment layout (register j = this thread's pair of matrix j: row lane / 4,
// columns 2 (lane % 4), +1) to shared memory; lane l supplies the 16-byte row address of row l % 8 of matrix l / 8
__device__ __forceinline__ void stmatrix_x4(uint32_t smem_addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
    asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};"
                 ::"r"(smem_addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

// legacy warp-level tensor-core MMA for the K=48 first layer (register accumulators, no TMEM needed)
__device__ __forceinline__ void mma_m16n8k16_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void mma_m16n8k16_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <bool F16>
__device__ __forceinline__ void mma_m16n8k16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if (F16) mma_m16n8k16_f16(c, a, b0, b1); else mma_m16n8k16_bf16(c, a, b0, b1);
}

// Mish for the 16-bit epilogues: 7 FP32 + 2 MUFU instructions.  With e = exp(x), s = e + 1:
// tanh(softplus(x)) = (s^2 - 1)/(s^2 + 1) = 1 - 2/(s^2 + 1), so mish(x) = x - 2x / (s^2 + 1).
// ex2.approx / rcp.approx are ~2^-22 relative, far below the 2^-9 (bf16) / 2^-11 (fp16) of the output;
// the exponent is capped at x = 20 (mish(x) = x to fp32 there, s^2 stays finite).
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mish_fast(float x) {
    const float t = fminf(x * 1.4426950408889634f, 28.853900817779268f);
    const float s = ex2_approx(t) + 1.f;
    const float r = rcp_approx(fmaf(s, s, 1.f));
    return fmaf(-2.f * x, r, x);
}
// N independent Mish evaluations in explicit stages (all ex2, then all rcp): with only two epilogue warps
// per SM sub-partition the MUFU latency has to be covered by instruction-level parallelism inside the warp.
// ex2 of two values with ONE MUFU op (packed fp16 in, packed fp16 out; ~2^-11 relative).  The MUFU pipe
// (16 lanes/SM) is the tightest resource of the Mish epilogues; this halves the exp half of its load.
// Arguments must be <= 15.9 so that 2^t stays finite in fp16.
__device__ __forceinline__ void ex2_pair_f16(float t0, float t1, float& e0, float& e1) {
    __half2 h = __floats2half2_rn(t0, t1);
    uint32_t u = *reinterpret_cast<uint32_t*>(&h), v;
    asm("ex2.approx.f16x2 %0, %1;" : "=r"(v) : "r"(u));
    const float2 f = __half22float2(*reinterpret_cast<__half2*>(&v));
    e0 = f.x; e1 = f.y;
}
// RCPG = 2 or 4: one MUFU.RCP serves RCPG elements -- 1/a and 1/b from r = rcp(a*b) as r*b and r*a (extra FMULs run
// on the wide FMA pipe, the 16-lane MUFU pipe is what the Mish epilogues saturate).  The exponent cap at x = 11
// keeps the product of four (e^x+1)^2+1 terms below 2^128; mish(x) == x to fp32 precision beyond it.
#ifndef DDP_MISH_RCPG
#define DDP_MISH_RCPG 2
#endif
// Instruction count per element (RCPG = 2): FFMA, FMNMX, EX2, FADD, FFMA, 1/2 FMUL, 1/2 RCP, FMUL, FFMA.  The exponent is
// shifted by -1/2 so that v = (e^x + 1)/sqrt(2) and q = -(v^2 + 1/2) = -((e^x+1)^2 + 1)/2 come out of one FADD and one
// FFMA; then mish(x) = x + x/q needs no separate -2x product.
template <int N, bool HALF_EX2 = false, int RCPG = DDP_MISH_RCPG>
__device__ __forceinline__ void mish_fast_n(float (&x)[N]) {
    float s[N];
    constexpr float kCap = (HALF_EX2 || RCPG > 1) ? 15.37f : 28.353900817779268f;
    constexpr float kL2e = 1.4426950408889634f, kRh = 0.70710678118654752f;
    if (HALF_EX2) {
#pragma unroll
        for (int i = 0; i < N; i += 2)
            ex2_pair_f16(fminf(fmaf(x[i], kL2e, -0.5f), kCap), fminf(fmaf(x[i + 1], kL2e, -0.5f), kCap), s[i], s[i + 1]);
    } else if (RCPG == 1) {
        // no cap needed: e^x = inf -> q = -inf -> 1/q = -0 -> mish = x
#pragma unroll
        for (int i = 0; i < N; ++i) s[i] = ex2_approx(fmaf(x[i], kL2e, -0.5f));
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) s[i] = ex2_approx(fminf(fmaf(x[i], kL2e, -0.5f), kCap));
    }
#pragma unroll
    for (int i = 0; i < N; ++i) { const float v = s[i] + kRh; s[i] = fmaf(v, -v, -0.5f); }
    if (RCPG == 4) {
#pragma unroll
        for (int i = 0; i < N; i += 4) {
            const float p01 = s[i] * s[i + 1], p23 = s[i + 2] * s[i + 3];
            const float r = rcp_approx(p01 * p23);
            const float r01 = r * p23, r23 = r * p01;
            const float r0 = r01 * s[i + 1], r1 = r01 * s[i], r2 = r23 * s[i + 3], r3 = r23 * s[i + 2];
            s[i] = r0; s[i + 1] = r1; s[i + 2] = r2; s[i + 3] = r3;
        }
    } else if (RCPG == 2) {
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            const float r = rcp_approx(s[i] * s[i + 1]);
            const float r0 = r * s[i + 1], r1 = r * s[i];
            s[i] = r0; s[i + 1] = r1;
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) s[i] = rcp_approx(s[i]);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = fmaf(x[i], s[i], x[i]);
}

// Mish of two values in packed fp16 arithmetic: x (f16x2) in, mish(x) (f16x2) out -- the operand format of the next
// layer's MMA, so no conversion follows.  13 instructions per PAIR (the fp32 sequence above costs 19 per pair) and half
// the MUFU work (ex2.approx.f16x2 is two MUFU.EX2.F16; there is no reciprocal on the XU pipe at all):
//   t = min(x log2e - 1/2, 7)                     HFMA2, HMNMX2      (cap: p below stays < 2^15)
//   v = 2^t + 1/sqrt2 = (e^x + 1)/sqrt2           2 MUFU, PRMT, HADD2
//   p = v^2 + 1/2 = ((e^x + 1)^2 + 1)/2 >= 1      HFMA2
//   r ~ 1/p: r0 = bits(0x77b7) - bits(p) (7 % seed), e = 1 - p r0, r = r0 + r0 (e + e^2)   IADD, 3 HFMA2 (7.4e-4 rel.)
//   mish = x - x r                                HFMA2
// Relative error ~2^-10 for x >= 0, absolute ~|x| 2^-11 for x << 0 (1 - r cancels there; mish itself -> 0): the same
// size as rounding the activation to a 16-bit operand.  Used by the denoising steps after the first one.
__device__ __forceinline__ uint32_t h2_as_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 u32_as_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t mish_h2(uint32_t xb) {
    const __half2 x = u32_as_h2(xb);
    const __half2 kL2E = __float2half2_rn(1.4426950408889634f), kMH = __float2half2_rn(-0.5f), kCap = __float2half2_rn(7.0f);
    const __half2 kRh = __float2half2_rn(0.70710678118654752f), kHalf = __float2half2_rn(0.5f), kOne = __float2half2_rn(1.f);
    const __half2 t = __hmin2(__hfma2(x, kL2E, kMH), kCap);
    uint32_t ub;
    asm("ex2.approx.f16x2 %0, %1;" : "=r"(ub) : "r"(h2_as_u32(t)));
    const __half2 v = __hadd2(u32_as_h2(ub), kRh);
    const __half2 p = __hfma2(v, v, kHalf);
    const __half2 r0 = u32_as_h2(0x77b777b7u - h2_as_u32(p));
    const __half2 e = __hfma2(__hneg2(p), r0, kOne);
    const __half2 f = __hfma2(e, e, e);
    const __half2 r = __hfma2(r0, f, r0);
    return h2_as_u32(__hfma2(__hneg2(x), r, x));
}

// legacy warp-level MMA with fp16 accumulators: D (2 x f16x2) = A . B + C; c[0] = (row g, cols 2t, 2t+1), c[1] = row g + 8
__device__ __forceinline__ void mma_m16n8k16_f16acc(uint32_t (&c)[2], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0, %1}, {%2, %3, %4, %5}, {%6, %7}, {%0, %1};"
        : "+r"(c[0]), "+r"(c[1]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ------------------------------------------------------------------------------------ host: tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 16-bit (bf16 or fp16) row-major [rows][cols] matrix, box = [box_rows][64 cols] with the 128-byte swizzle
inline int make_tmap_bf16_sw128(CUtensorMap* map, const void* gptr, uint64_t rows, uint64_t cols, uint32_t box_rows,
                                bool f16 = false) {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return -1;
        fn = (PFN_encodeTiled)p;
    }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(gptr), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

}  // namespace tc
}  // namespace ddp
