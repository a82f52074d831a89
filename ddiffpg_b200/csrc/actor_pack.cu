// ddp_actor_pack: weight re-layout, time-embedding table and DDPM schedule constants.
//
// Reference semantics folded here (paths relative to the reference repo):
//   SinusoidalPosEmb.forward            ddiffpg/models/diffusion_mlp.py:14-21
//   DiffusionNet.time_mlp               ddiffpg/models/diffusion_mlp.py:38-43,68
//   cat([t, cond, x]) @ net.mlp.0       ddiffpg/models/diffusion_mlp.py:70-71  (time columns only)
//   DDPMScheduler(squaredcos_cap_v2)    third-party diffusers ^0.18.2, ctor at diffusion_mlp.py:167-173
#include <math.h>
#include "actor_layout.cuh"

namespace ddp {

// squaredcos_cap_v2 betas in double, alphas / cumprod / step coefficients in fp32 -- the same
// precision sequence the scheduler's 0-dim fp32 tensors go through (oracle/ddpm.py).
void fill_schedule(int T, ScheduleTable& tab) {
    auto abar = [](double u) { double c = cos((u + 0.008) / 1.008 * M_PI / 2.0); return c * c; };
    float ac[kMaxT];
    float prod = 1.0f;
    for (int i = 0; i < T; ++i) {
        double b = 1.0 - abar((double)(i + 1) / T) / abar((double)i / T);
        if (b > 0.999) b = 0.999;
        volatile float beta = (float)b;
        volatile float alpha = 1.0f - beta;
        volatile float p = prod * alpha;
        prod = p;
        ac[i] = prod;
    }
    for (int t = 0; t < T; ++t) {
        volatile float a_t = ac[t];
        volatile float a_prev = t > 0 ? ac[t - 1] : 1.0f;
        volatile float b_t = 1.0f - a_t;
        volatile float b_prev = 1.0f - a_prev;
        volatile float cur_alpha = a_t / a_prev;
        volatile float cur_beta = 1.0f - cur_alpha;
        volatile float s_at = sqrtf(a_t);
        volatile float s_bt = sqrtf(b_t);
        volatile float num0 = sqrtf(a_prev) * cur_beta;
        volatile float c_x0 = num0 / b_t;
        volatile float num1 = sqrtf(cur_alpha) * b_prev;
        volatile float c_xt = num1 / b_t;
        volatile float var0 = b_prev / b_t;
        volatile float var = var0 * cur_beta;
        if (var < 1e-20f) var = 1e-20f;
        float* r = tab.v[t];
        r[CST_CEPS] = s_bt;
        r[CST_SQRT_AB] = s_at;
        r[CST_CX0] = c_x0;
        r[CST_CXT] = c_xt;
        r[CST_SIGMA] = t > 0 ? sqrtf(var) : 0.0f;
        r[CST_ADD_A] = s_at;
        r[CST_ADD_B] = s_bt;
        r[7] = 0.0f;
    }
}

// the five small jobs of a pack in one launch: blockIdx.y = b1 | b2 | b3 (zero padded to np3) | scheduler constants | posemb
__global__ void actor_small_pack_kernel(ScheduleTable tab, int T, int D, const float* __restrict__ b1, int n1,
                                        float* __restrict__ d1, const float* __restrict__ b2, int n2, float* __restrict__ d2,
                                        const float* __restrict__ b3, int n3, int np3, float* __restrict__ d3,
                                        float* __restrict__ cst, float* __restrict__ pe) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    switch (blockIdx.y) {
    case 0: if (i < n1) d1[i] = b1[i]; break;
    case 1: if (i < n2) d2[i] = b2[i]; break;
    case 2: if (i < np3) d3[i] = i < n3 ? b3[i] : 0.f; break;
    case 3: if (i < T * kCstStride) cst[i] = tab.v[i / kCstStride][i % kCstStride]; break;
    default:
        if (i < T * D) {
            const int t = i / D, k = i % D, half = D / 2;
            const float c = -(float)(log(10000.0) / (half - 1));
            const int j = k < half ? k : k - half;
            const float f = expf(__fmul_rn((float)j, c));
            const float a = __fmul_rn((float)t, f);
            pe[i] = k < half ? sinf(a) : cosf(a);
        }
    }
}

// y[t][n] = b[n] + sum_k W[n*ldw + koff + k] * x[t*ldx + k]; optional Mish; one warp per (t, n).
// zout (may be NULL) receives the pre-activation.
__global__ void rows_linear_kernel(const float* __restrict__ W, int ldw, int koff, const float* __restrict__ b,
                                   const float* __restrict__ x, int ldx, int K, int N, int apply_mish,
                                   float* __restrict__ zout, float* __restrict__ y, int ldy) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int t = blockIdx.y;
    if (warp >= N) return;
    const float* w = W + (size_t)warp * ldw + koff;
    const float* xr = x + (size_t)t * ldx;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(w[k], xr[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
        float z = acc + b[warp];
        if (zout) zout[(size_t)t * ldy + warp] = z;
        y[(size_t)t * ldy + warp] = apply_mish ? mish_f(z) : z;
    }
}

int pack_actor_fp32(const ActorLayout& L, const float* const p[12], float* out, bool fma_operands, cudaStream_t st) {
    const int D = L.D, S = L.S, A = L.A, T = L.T;
    const int ld0 = D + S + A;
    auto blocks = [](size_t n) { return (unsigned)((n + 255) / 256); };
    if (fma_operands) {
        // transposed fp32 weights: streamed by the warp-FMA kernels only (the tensor paths read their own 16-bit copies)
        transpose_pack_kernel<<<blocks((size_t)L.K0p * L.h1), 256, 0, st>>>(p[4], ld0, D, S + A, L.K0p, L.h1, L.h1, out + L.wt0);
        transpose_pack_kernel<<<blocks((size_t)L.h1 * L.h2), 256, 0, st>>>(p[6], L.h1, 0, L.h1, L.h1, L.h2, L.h2, out + L.wt1);
        transpose_pack_kernel<<<blocks((size_t)L.h2 * L.h3), 256, 0, st>>>(p[8], L.h2, 0, L.h2, L.h2, L.h3, L.h3, out + L.wt2);
        transpose_pack_kernel<<<blocks((size_t)L.h3 * L.A4), 256, 0, st>>>(p[10], L.h3, 0, L.h3, L.h3, A, L.A4, out + L.wt3);
        copy_pad_kernel<<<blocks((size_t)L.A4 * L.h3), 256, 0, st>>>(p[10], A * L.h3, L.A4 * L.h3, out + L.w3b);
    }
    // biases, scheduler constants and positional embedding: one launch (block y = job)
    ScheduleTable tab;
    fill_schedule(T, tab);
    {
        size_t most = (size_t)T * D;
        if ((size_t)L.h2 > most) most = L.h2;
        if ((size_t)T * kCstStride > most) most = (size_t)T * kCstStride;
        dim3 grid(blocks(most), 5);
        actor_small_pack_kernel<<<grid, 256, 0, st>>>(tab, T, D, p[7], L.h2, out + L.b1, p[9], L.h3, out + L.b2, p[11], A, L.A4,
                                                     out + L.b3, out + L.cst, out + L.pe);
    }
    // time path on the T distinct timesteps
    dim3 g1(blocks((size_t)4 * D * 32), T), g2(blocks((size_t)D * 32), T), g3(blocks((size_t)L.h1 * 32), T);
    rows_linear_kernel<<<g1, 256, 0, st>>>(p[0], D, 0, p[1], out + L.pe, D, D, 4 * D, 1, out + L.zmid, out + L.hmid, 4 * D);
    rows_linear_kernel<<<g2, 256, 0, st>>>(p[2], 4 * D, 0, p[3], out + L.hmid, 4 * D, 4 * D, D, 0, nullptr, out + L.temb, D);
    rows_linear_kernel<<<g3, 256, 0, st>>>(p[4], ld0, 0, p[5], out + L.temb, D, D, L.h1, 0, nullptr, out + L.tb0, L.h1);
    DDP_LAUNCH_CHECK("actor pack kernels");
    return DDP_OK;
}

}  // namespace ddp
