// Micro-benchmark: throughput of the Mish epilogue arithmetic with 8 warps per SM (2 per sub-partition), registers
// only.  Variants isolate the MUFU ops, the FP32 ops and the 16-bit pack.  Prints clk per 32-lane element group.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/mish_bw tools/micro/mish_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../../ddiffpg_b200/csrc/tc_common.cuh"
using ddp::tc::mish_h2;
using ddp::tc::pack_f16x2;

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpa(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t packbf(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ uint32_t packint(float lo, float hi) {
    const uint32_t a = __float_as_uint(lo) + 0x8000u, b = __float_as_uint(hi) + 0x8000u;
    return __byte_perm(a, b, 0x7632);
}

template <int V>
__global__ void __launch_bounds__(512, 1) k(const float* in, uint32_t* out, int iters, long long* clk) {
    float x[16];
    for (int i = 0; i < 16; ++i) x[i] = in[threadIdx.x * 16 + i];
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float s[16], y[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) y[i] = x[i] + __uint_as_float(acc & 1u);    // loop-carried, cheap
        if (V == 0 || V == 2 || V == 3) {            // full Mish (pair rcp)
#pragma unroll
            for (int i = 0; i < 16; ++i) s[i] = ex2a(fminf(fmaf(y[i], 1.4426950408889634f, -0.5f), 15.37f));
#pragma unroll
            for (int i = 0; i < 16; ++i) { const float v = s[i] + 0.70710678f; s[i] = fmaf(v, -v, -0.5f); }
#pragma unroll
            for (int i = 0; i < 16; i += 2) { const float r = rcpa(s[i] * s[i + 1]); const float r0 = r * s[i + 1], r1 = r * s[i]; s[i] = r0; s[i + 1] = r1; }
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] = fmaf(y[i], s[i], y[i]);
        } else if (V == 1) {                          // MUFU only: 1 ex2 per element + 1 rcp per pair
#pragma unroll
            for (int i = 0; i < 16; ++i) s[i] = ex2a(y[i]);
#pragma unroll
            for (int i = 0; i < 16; i += 2) { const float r = rcpa(s[i]); y[i] = r; y[i + 1] = s[i + 1]; }
        } else if (V == 5) {                          // one rcp per element, no exponent cap: 4 FP + 2 MUFU
#pragma unroll
            for (int i = 0; i < 16; ++i) s[i] = ex2a(fmaf(y[i], 1.4426950408889634f, -0.5f));
#pragma unroll
            for (int i = 0; i < 16; ++i) { const float v = s[i] + 0.70710678f; s[i] = rcpa(fmaf(v, -v, -0.5f)); }
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] = fmaf(y[i], s[i], y[i]);
        } else if (V == 4) {                          // FP32 ops only (no MUFU)
#pragma unroll
            for (int i = 0; i < 16; ++i) s[i] = fminf(fmaf(y[i], 1.4426950408889634f, -0.5f), 15.37f);
#pragma unroll
            for (int i = 0; i < 16; ++i) { const float v = s[i] + 0.70710678f; s[i] = fmaf(v, -v, -0.5f); }
#pragma unroll
            for (int i = 0; i < 16; i += 2) { const float r = s[i] * s[i + 1]; const float r0 = r * s[i + 1], r1 = r * s[i]; s[i] = r0; s[i + 1] = r1; }
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] = fmaf(y[i], s[i], y[i]);
        }
        if (V == 6) {                                 // packed-fp16 Mish from fp32 accumulators (drain path): F2FP + HADD2 bias + mish_h2
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                const __half2 xb = __hadd2(ddp::tc::u32_as_h2(pack_f16x2(y[i], y[i + 1])), ddp::tc::u32_as_h2(0x2e662e66u));
                acc ^= mish_h2(ddp::tc::h2_as_u32(xb));
            }
        } else if (V == 7) {                          // packed-fp16 Mish on packed inputs (layer-0 path, fp16 accumulators)
#pragma unroll
            for (int i = 0; i < 16; i += 2) acc ^= mish_h2(__float_as_uint(y[i]) ^ (__float_as_uint(y[i + 1]) >> 16));
        }
        if (V == 0 || V == 1 || V == 4 || V == 5) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) acc ^= packbf(y[i], y[i + 1]);
        } else if (V == 2) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) acc ^= packint(y[i], y[i + 1]);
        } else if (V == 3) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc ^= __float_as_uint(y[i]);
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int V>
void run(const char* name, const float* in, uint32_t* out, long long* clk, int threads) {
    const int iters = 4096;
    k<V><<<148, threads>>>(in, out, iters, clk);
    k<V><<<148, threads>>>(in, out, iters, clk);
    cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    const int warps_per_smsp = threads / 128;
    printf("%-34s threads %3d: %7.2f clk per 16-element batch per warp, %6.2f clk per warp-element per sub-partition\n", name, threads,
           (double)h / iters, (double)h / iters / 16.0 / warps_per_smsp);
}

int main() {
    float* in; uint32_t* out; long long* clk;
    cudaMalloc(&in, 512 * 16 * 4); cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&clk, 8);
    cudaMemset(in, 0, 512 * 16 * 4);
    for (int threads : {256, 512}) {
        run<0>("mish + F2FP pack", in, out, clk, threads);
        run<6>("packed-fp16 mish (F2FP in, bias)", in, out, clk, threads);
        run<7>("packed-fp16 mish (packed in)", in, out, clk, threads);
        run<5>("mish, rcp per element + F2FP pack", in, out, clk, threads);
        run<2>("mish + integer round/pack", in, out, clk, threads);
        run<3>("mish, no pack", in, out, clk, threads);
        run<1>("MUFU only (1.5/elem) + F2FP pack", in, out, clk, threads);
        run<4>("FP32 only + F2FP pack", in, out, clk, threads);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
