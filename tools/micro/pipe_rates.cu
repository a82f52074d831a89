// Micro-benchmark: issue rate of the instructions the Mish epilogues are made of, per SM sub-partition.
// 8 independent dependency chains per thread, W warps per sub-partition; prints clk per warp-instruction per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/micro/pipe_rates tools/micro/pipe_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

enum { FFMA, FMUL, FADD, FMNMX, HFMA2, HADD2, HMNMX2, EX2, EX2H, RCP, TANH, TANHH, F2FP, IADD, PRMT, HMMA32, HMMA16, LOP, MIX_FFMA_EX2, MIX_HFMA2_EX2H, FFMAC, FFMA2R, HFMA2C, RSQ, LG2, F2FPBF, H2F, NOPS };
const char* kNames[] = {"FFMA", "FMUL", "FADD", "FMNMX", "HFMA2", "HADD2", "HMNMX2", "MUFU.EX2", "MUFU.EX2.F16 (one half)", "MUFU.RCP + FADD", "MUFU.TANH", "MUFU.TANH.F16", "F2FP.F16.F32.PACK_AB",
                        "IADD3", "PRMT", "HMMA.16816.F32 (f16 in)", "HMMA.16816.F16 (f16 acc)", "LOP3", "mix 4 FFMA + 1 EX2", "mix 4 HFMA2 + 1 EX2.F16",
                        "FFMA (2 immediates)", "FFMA (1 immediate)", "HFMA2 (same reg twice)", "MUFU.RSQ + FADD", "MUFU.LG2 + FADD", "F2FP.BF16.F32.PACK_AB", "HADD2.F32 (f16 -> f32)"};

template <int OP>
__global__ void __launch_bounds__(512, 1) k(uint32_t* out, int iters, long long* clk, uint32_t seed) {
    uint32_t r[8];
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { r[i] = seed + threadIdx.x * 8 + i; f[i] = 1.0f + 1e-3f * (float)(threadIdx.x + i); }
    uint32_t a4[4] = {seed, seed + 1, seed + 2, seed + 3};
    float c4[8][4] = {};
    uint32_t h4[8][2] = {};
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(f[(i + 1) & 7]), "f"(f[(i + 2) & 7]));
                if (OP == FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(f[(i + 1) & 7]));
                if (OP == FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(f[(i + 1) & 7]));
                if (OP == FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(f[(i + 1) & 7]));
                if (OP == HFMA2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(r[(i + 1) & 7]), "r"(r[(i + 2) & 7]));
                if (OP == HADD2) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 1) & 7]));
                if (OP == HMNMX2) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 1) & 7]));
                if (OP == EX2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
                if (OP == EX2H) asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %0; ex2.approx.f16 lo, lo; mov.b32 %0, {lo, hi};}" : "+r"(r[i]));
                if (OP == RCP) asm volatile("rcp.approx.ftz.f32 %0, %0;\n\tadd.rn.f32 %0, %0, 0f3f800000;" : "+f"(f[i]));
                if (OP == FFMAC) asm volatile("fma.rn.f32 %0, %0, 0f3fb8aa3b, 0fbf000000;" : "+f"(f[i]));
                if (OP == FFMA2R) asm volatile("fma.rn.f32 %0, %0, %1, 0fbf000000;" : "+f"(f[i]) : "f"(f[(i + 1) & 7]));
                if (OP == HFMA2C) asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(r[i]) : "r"(0x3dc53dc5u));
                if (OP == RSQ) asm volatile("rsqrt.approx.ftz.f32 %0, %0;\n\tadd.rn.f32 %0, %0, 0f3f800000;" : "+f"(f[i]));
                if (OP == LG2) asm volatile("lg2.approx.ftz.f32 %0, %0;\n\tadd.rn.f32 %0, %0, 0f40000000;" : "+f"(f[i]));
                if (OP == TANH) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f[i]));
                if (OP == TANHH) asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %0; tanh.approx.f16 lo, lo; mov.b32 %0, {lo, hi};}" : "+r"(r[i]));
                if (OP == F2FP) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r[i]) : "f"(f[i]), "f"(f[(i + 1) & 7])); f[i] = __uint_as_float(r[i]); }
                if (OP == F2FPBF) { asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r[i]) : "f"(f[i]), "f"(f[(i + 1) & 7])); f[i] = __uint_as_float(r[i]); }
                if (OP == H2F) { asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo;}" : "=f"(f[i]) : "r"(r[i])); r[i] = __float_as_uint(f[i]); }
                if (OP == IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 1) & 7]));
                if (OP == LOP) asm volatile("xor.b32 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 1) & 7]));
                if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(r[i]) : "r"(r[(i + 1) & 7]));
                if (OP == HMMA32)
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                                 : "+f"(c4[i][0]), "+f"(c4[i][1]), "+f"(c4[i][2]), "+f"(c4[i][3]) : "r"(a4[0]), "r"(a4[1]), "r"(a4[2]), "r"(a4[3]), "r"(r[0]), "r"(r[1]));
                if (OP == HMMA16)
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0, %1}, {%2, %3, %4, %5}, {%6, %7}, {%0, %1};"
                                 : "+r"(h4[i][0]), "+r"(h4[i][1]) : "r"(a4[0]), "r"(a4[1]), "r"(a4[2]), "r"(a4[3]), "r"(r[0]), "r"(r[1]));
                if (OP == MIX_FFMA_EX2) {
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(f[(i + 1) & 7]), "f"(f[(i + 2) & 7]));
                    if ((i & 3) == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c4[i][0]));
                }
                if (OP == MIX_HFMA2_EX2H) {
                    asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(r[(i + 1) & 7]), "r"(r[(i + 2) & 7]));
                    if ((i & 3) == 0) asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %0; ex2.approx.f16 lo, lo; mov.b32 %0, {lo, hi};}" : "+r"(h4[i][0]));
                }
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= r[i] ^ __float_as_uint(f[i]) ^ __float_as_uint(c4[i][0] + c4[i][1] + c4[i][2] + c4[i][3]) ^ h4[i][0] ^ h4[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int OP>
void run(uint32_t* out, long long* clk) {
    const int iters = 2048;
    for (int threads : {256, 512}) {
        k<OP><<<148, threads>>>(out, iters, clk, 0x3c003c00u);
        k<OP><<<148, threads>>>(out, iters, clk, 0x3c003c00u);
        cudaDeviceSynchronize();
        long long h = 0;
        cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
        const double per = (double)h / iters / 32.0 / (threads / 128);     // 32 instructions of OP per iteration per warp
        const double extra = (OP == MIX_FFMA_EX2 || OP == MIX_HFMA2_EX2H) ? 1.25 : 1.0;
        printf("%-28s %d warps/SMSP: %6.2f clk per warp-instruction per sub-partition%s\n", kNames[OP], threads / 128, per / extra,
               extra > 1 ? " (per instruction of the mix)" : "");
    }
}

int main() {
    uint32_t* out; long long* clk;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&clk, 8);
    run<FFMA>(out, clk); run<FMUL>(out, clk); run<FADD>(out, clk); run<FMNMX>(out, clk); run<HFMA2>(out, clk); run<HADD2>(out, clk);
    run<HMNMX2>(out, clk); run<EX2>(out, clk); run<EX2H>(out, clk); run<RCP>(out, clk); run<TANH>(out, clk); run<TANHH>(out, clk);
    run<F2FP>(out, clk); run<IADD>(out, clk); run<PRMT>(out, clk); run<LOP>(out, clk); run<HMMA32>(out, clk); run<HMMA16>(out, clk);
    run<MIX_FFMA_EX2>(out, clk); run<MIX_HFMA2_EX2H>(out, clk);
    run<FFMAC>(out, clk); run<FFMA2R>(out, clk); run<HFMA2C>(out, clk); run<RSQ>(out, clk); run<LG2>(out, clk); run<F2FPBF>(out, clk); run<H2F>(out, clk);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
