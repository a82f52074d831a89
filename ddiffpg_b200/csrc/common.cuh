// Shared device/host helpers for the ddiffpg_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/ddiffpg_b200.h"

namespace ddp {

// ------------------------------------------------------------------------------------ errors
void set_error(const char* fmt, ...);
#define DDP_FAIL(code, ...) do { ::ddp::set_error(__VA_ARGS__); return (code); } while (0)
#define DDP_CUDA_CHECK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { \
    ::ddp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    return DDP_ERR_CUDA; } } while (0)
#define DDP_LAUNCH_CHECK(name) do { cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) { \
    ::ddp::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e)); return DDP_ERR_CUDA; } } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int pad4(int x) { return (x + 3) & ~3; }

// ------------------------------------------------------------------------------------ activations
// Mish(x) = x * tanh(softplus(x)) (nn.Mish, reference diffusion_mlp.py:30).  With e = exp(x) and
// n = e*(e+2): tanh(log(1+e)) = n/(n+2).  No cancellation; x > 20 returns x like ATen does in fp32.
__device__ __forceinline__ float mish_f(float x) {
    if (x > 20.f) return x;
    float e = expf(x);
    float n = e * (e + 2.f);
    return x * (n / (n + 2.f));
}
// value and derivative: mish'(x) = w + x * (1 - w^2) * sigmoid(x), w = tanh(softplus(x)).
__device__ __forceinline__ void mish_fd(float x, float& y, float& dy) {
    if (x > 20.f) { y = x; dy = 1.f; return; }
    float e = expf(x);
    float n = e * (e + 2.f);
    float w = n / (n + 2.f);
    float sg = e / (1.f + e);
    y = x * w;
    dy = w + x * (1.f - w * w) * sg;
}
// ELU(alpha=1) (reference mlp.py:13); derivative from the activation value a: a > 0 ? 1 : a + 1.
__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : expm1f(x); }
__device__ __forceinline__ float elu_grad_from_act(float a) { return a > 0.f ? 1.f : a + 1.f; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------ row-tile GEMM
// out[r][n] = epi( sum_k in[r][k] * Wt[k*ldw + n] ), r < RT, n < N, for one CTA-resident row tile.
//   in   : shared memory, row stride ld_in floats (16-byte aligned rows), K4*4 valid (zero padded)
//   Wt   : global memory, [K4*4][ldw] with ldw % 4 == 0 (K-major = the Linear weight transposed)
//   epi  : callable (int r, int n0, float4 acc) invoked once per (row, 4-column group)
// Thread mapping: CG = N/4 column groups; KS lanes (power of two <= 32, adjacent lanes) split K in
// blocks of 4 and are reduced with shuffles, so every warp streams KS rows x (32/KS)*16 B of Wt per
// request.  Column groups beyond the thread count are looped.  All threads must call (barrier-free;
// the caller synchronises before and after).
template <int RT, int NT, typename Epi>
__device__ __forceinline__ void tile_linear(const float* __restrict__ Wt, int ldw, int K4, int N,
                                            const float* in, int ld_in, int KS, Epi epi) {
    const int CG = N >> 2;
    const int tid = threadIdx.x;
    const int ks = tid & (KS - 1);
    const int groups_per_pass = NT / KS;
    const int passes = (CG + groups_per_pass - 1) / groups_per_pass;
    for (int pass = 0; pass < passes; ++pass) {
        // every thread runs every pass so the shuffles stay convergent; out-of-range lanes idle
        const int cg = pass * groups_per_pass + tid / KS;
        const bool active = cg < CG;
        float4 acc[RT];
#pragma unroll
        for (int r = 0; r < RT; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) {
            const float* wp = Wt + 4 * cg;
            int kb = ks;
            float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0, w2 = w0, w3 = w0;
            if (kb < K4) {
                const float* q = wp + (size_t)(4 * kb) * ldw;
                w0 = __ldg(reinterpret_cast<const float4*>(q));
                w1 = __ldg(reinterpret_cast<const float4*>(q + ldw));
                w2 = __ldg(reinterpret_cast<const float4*>(q + 2 * ldw));
                w3 = __ldg(reinterpret_cast<const float4*>(q + 3 * ldw));
            }
            while (kb < K4) {
                const int kn = kb + KS;
                float4 n0 = w0, n1 = w1, n2 = w2, n3 = w3;
                if (kn < K4) {      // register double buffer: next block's weights in flight
                    const float* q = wp + (size_t)(4 * kn) * ldw;
                    n0 = __ldg(reinterpret_cast<const float4*>(q));
                    n1 = __ldg(reinterpret_cast<const float4*>(q + ldw));
                    n2 = __ldg(reinterpret_cast<const float4*>(q + 2 * ldw));
                    n3 = __ldg(reinterpret_cast<const float4*>(q + 3 * ldw));
                }
                const float* ip = in + 4 * kb;
#pragma unroll
                for (int r = 0; r < RT; ++r) {
                    const float4 a = *reinterpret_cast<const float4*>(ip + r * ld_in);
                    acc[r].x = fmaf(a.x, w0.x, acc[r].x); acc[r].y = fmaf(a.x, w0.y, acc[r].y);
                    acc[r].z = fmaf(a.x, w0.z, acc[r].z); acc[r].w = fmaf(a.x, w0.w, acc[r].w);
                    acc[r].x = fmaf(a.y, w1.x, acc[r].x); acc[r].y = fmaf(a.y, w1.y, acc[r].y);
                    acc[r].z = fmaf(a.y, w1.z, acc[r].z); acc[r].w = fmaf(a.y, w1.w, acc[r].w);
                    acc[r].x = fmaf(a.z, w2.x, acc[r].x); acc[r].y = fmaf(a.z, w2.y, acc[r].y);
                    acc[r].z = fmaf(a.z, w2.z, acc[r].z); acc[r].w = fmaf(a.z, w2.w, acc[r].w);
                    acc[r].x = fmaf(a.w, w3.x, acc[r].x); acc[r].y = fmaf(a.w, w3.y, acc[r].y);
                    acc[r].z = fmaf(a.w, w3.z, acc[r].z); acc[r].w = fmaf(a.w, w3.w, acc[r].w);
                }
                w0 = n0; w1 = n1; w2 = n2; w3 = n3;
                kb = kn;
            }
        }
        // reduce the K split across the KS adjacent lanes (all 32 lanes participate)
        for (int o = 1; o < KS; o <<= 1) {
#pragma unroll
            for (int r = 0; r < RT; ++r) {
                acc[r].x += __shfl_xor_sync(0xffffffffu, acc[r].x, o);
                acc[r].y += __shfl_xor_sync(0xffffffffu, acc[r].y, o);
                acc[r].z += __shfl_xor_sync(0xffffffffu, acc[r].z, o);
                acc[r].w += __shfl_xor_sync(0xffffffffu, acc[r].w, o);
            }
        }
        if (active) {
#pragma unroll
            for (int r = 0; r < RT; ++r)
                if ((r & (KS - 1)) == ks) epi(r, 4 * cg, acc[r]);
        }
    }
}

// ------------------------------------------------------------------------------------ pack helpers
// Wt[k][n] = W[n][koff + k] for k < K, 0 for K <= k < Kp; n < N, 0 for N <= n < ldo.
static __global__ void transpose_pack_kernel(const float* __restrict__ W, int ldw, int koff, int K, int Kp,
                                      int N, int ldo, float* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)Kp * ldo;
    if (i >= total) return;
    int k = (int)(i / ldo), n = (int)(i % ldo);
    out[i] = (k < K && n < N) ? W[(size_t)n * ldw + koff + k] : 0.f;
}

static __global__ void copy_pad_kernel(const float* __restrict__ src, int n, int np, float* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) dst[i] = i < n ? src[i] : 0.f;
}

// K-split factor for a layer with N outputs on NT threads: the largest power of two <= NT/(N/4),
// capped at 8 (beyond that the partial-sum shuffles cost more than the idle lanes).
__host__ __device__ inline int pick_ksplit(int N, int NT) {
    int cg = N >> 2;
    int ks = 1;
    while (ks * 2 * cg <= NT && ks < 8) ks *= 2;
    return ks;
}

}  // namespace ddp
