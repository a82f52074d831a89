// H2, bf16 tensor-core path: mode-segmented double-Q forward, its gradient w.r.t. the action and the Adam
// action-ascent loop, assembled from the grouped tcgen05 row GEMM (csrc/tc_gemm.cu; one group per mode, so all
// K+1 critics run in the same launches).
//
// Reference semantics (paths relative to the reference repo):
//   MLPNet (ELU) / DistributionalDoubleQ.get_q1_q2 / get_q_min     ddiffpg/models/mlp.py:13-35,143-151
//   AgentDDiffPG.update_target_action + optimizer_update           ddiffpg/algo/ddiffpg.py:358-373, ac_base.py:83-92
// Operands bf16, accumulation fp32; softmax / expectation / arg-min selection, the clip norm and Adam are fp32.
// Per ascent iteration: 2 nets x 4 forward GEMMs, one softmax/selection kernel, 2 nets x 4 backward GEMMs (ELU
// derivative fused in the epilogue, from the stored activation), one gradient/norm kernel, one Adam kernel.
#include <math.h>
#include "q_layout.cuh"
#include "tc_gemm.cuh"

namespace ddp {

// fused per-tile forward/backward chain (csrc/q_chain_tc.cu)
bool q_chain_shape_ok(const QLayout& L);
size_t q_chain_workspace(const QLayout& L);
struct QChainAscent {
    int iters;
    float* act;
    float *m1, *m2;
    float* gnorm_out;
    unsigned int* grid_bar;
    float lr, beta1, beta2, eps, max_norm, lim;
};
int q_chain_pass(const QLayout& L, const void* packed, const int64_t* seg_off, const float* scale, const void* xin,
                 float* g_out, float* gsq, float* qmin, float* p1, float* p2, long B, void* scratch,
                 size_t scratch_bytes, cudaStream_t st, const QChainAscent* asc);

namespace {

using bf16 = __nv_bfloat16;

// Debug switches (ddp_debug_q_variant, not in the public header): the three schedules of the same bf16 arithmetic that
// tests/test_tc_gpu.py pins against each other.  Defaults: fused chain where the shape fits; all iterations in one
// cooperative launch while the batch is at most ~half a wave of tiles.
int g_q_no_chain = 0;        // 1: layer-by-layer GEMM path even where the fused kernel applies
int g_q_fused_adam = -1;     // -1: by batch size, 0 / 1: force the per-iteration / the single-launch form
bool use_chain(const QLayout& L) { return !g_q_no_chain && q_chain_shape_ok(L); }

struct QSeg {
    long off[kMaxModes + 1];
    float inv_cnt[kMaxModes];
    int n_modes;
};

struct QTcWs {
    bf16 *xin, *a1[2], *a2[2], *a3[2], *dl[2];
    float *logits[2], *ga[2], *g, *m1, *m2, *gsq, *abs_sum;
    unsigned int* grid_bar;
    void* chain_scratch;
    size_t total, adam_bytes;
};

QTcWs carve(const QLayout& L, long B, int iters, uint8_t* base) {
    QTcWs w{};
    size_t o = 0;
    auto take = [&](size_t bytes) { uint8_t* r = base ? base + o : nullptr; o += (bytes + 255) / 256 * 256; return r; };
    const bool chain = use_chain(L);
    if (chain) w.chain_scratch = take(q_chain_workspace(L));
    w.xin = (bf16*)take((size_t)B * L.K1c * 2);
    for (int j = 0; j < 2 && !chain; ++j) {
        w.a1[j] = (bf16*)take((size_t)B * L.h1 * 2);
        w.a2[j] = (bf16*)take((size_t)B * L.h2 * 2);
        w.a3[j] = (bf16*)take((size_t)B * L.h3 * 2);
        w.dl[j] = (bf16*)take((size_t)B * 64 * 2);
        w.logits[j] = (float*)take((size_t)B * 64 * 4);
        w.ga[j] = (float*)take((size_t)B * 16 * 4);
    }
    w.g = (float*)take((size_t)B * L.A * 4);
    const size_t adam0 = o;
    w.m1 = (float*)take((size_t)B * L.A * 4);
    w.m2 = (float*)take((size_t)B * L.A * 4);
    w.gsq = (float*)take((size_t)(iters > 0 ? iters : 1) * kMaxModes * 4);
    w.abs_sum = (float*)take(kMaxModes * 4);
    w.grid_bar = (unsigned int*)take(256);
    w.adam_bytes = o - adam0;
    w.total = o;
    return w;
}

// xin[r] = [obs (O) | action (A) | 0 ...] as bf16, 64 columns; `clamp_lim` > 0 also clamps the fp32 action in
// place first (the pre-loop clamp of update_target_action, ddiffpg.py:361)
__global__ void q_tc_prep_kernel(const float* __restrict__ obs, float* __restrict__ act, int O, int A, long B,
                                 float clamp_lim, bf16* __restrict__ xin, int K1c) {
    // one thread = 8 consecutive columns of a row = one 16-byte store; K1c (a multiple of 64) columns per row
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int per_row = K1c >> 3;
    const long row = idx / per_row;
    const int c0 = (int)(idx - row * per_row) * 8;
    if (row >= B) return;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = c0 + i;
        float x = 0.f;
        if (c < O) x = obs[row * O + c];
        else if (c < O + A) {
            x = act[row * A + c - O];
            if (clamp_lim > 0.f) { x = fminf(fmaxf(x, -clamp_lim), clamp_lim); act[row * A + c - O] = x; }
        }
        v[i] = x;
    }
    if (xin) {
        __nv_bfloat162 h[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(xin + row * K1c + c0) = *reinterpret_cast<const uint4*>(h);
    }
}

// One warp per row: softmax over atoms of both nets, expectations, min / selection, d min(Q1,Q2) / d logits.
__global__ void q_tc_softmax_kernel(const float* __restrict__ lg1, const float* __restrict__ lg2,
                                    const float* __restrict__ z, int atoms, long B, float* __restrict__ qmin_out,
                                    float* __restrict__ p1_out, float* __restrict__ p2_out, bf16* __restrict__ dl1,
                                    bf16* __restrict__ dl2) {
    const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= B) return;
    const int c0 = lane, c1 = lane + 32;
    const float z0 = c0 < atoms ? z[c0] : 0.f, z1 = c1 < atoms ? z[c1] : 0.f;
    float p[2][2], q[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float* lg = (j == 0 ? lg1 : lg2) + row * 64;
        const float l0 = c0 < atoms ? lg[c0] : -INFINITY, l1 = c1 < atoms ? lg[c1] : -INFINITY;
        const float mx = warp_max(fmaxf(l0, l1));
        const float e0 = c0 < atoms ? expf(l0 - mx) : 0.f, e1 = c1 < atoms ? expf(l1 - mx) : 0.f;
        const float sum = warp_sum(e0 + e1);
        p[j][0] = e0 / sum; p[j][1] = e1 / sum;
        q[j] = warp_sum(p[j][0] * z0 + p[j][1] * z1);
    }
    if (lane == 0 && qmin_out) qmin_out[row] = fminf(q[0], q[1]);
    if (p1_out) { if (c0 < atoms) p1_out[row * atoms + c0] = p[0][0]; if (c1 < atoms) p1_out[row * atoms + c1] = p[0][1]; }
    if (p2_out) { if (c0 < atoms) p2_out[row * atoms + c0] = p[1][0]; if (c1 < atoms) p2_out[row * atoms + c1] = p[1][1]; }
    if (dl1) {
        // only the smaller net carries gradient; exact ties split evenly (torch.min backward)
        const float w1 = q[0] == q[1] ? 0.5f : (q[0] < q[1] ? 1.f : 0.f), w2 = 1.f - w1;
        dl1[row * 64 + c0] = __float2bfloat16(w1 * p[0][0] * (z0 - q[0]));
        dl1[row * 64 + c1] = __float2bfloat16(w1 * p[0][1] * (z1 - q[0]));
        dl2[row * 64 + c0] = __float2bfloat16(w2 * p[1][0] * (z0 - q[1]));
        dl2[row * 64 + c1] = __float2bfloat16(w2 * p[1][1] * (z1 - q[1]));
    }
}

// g = scale * (ga1 + ga2) per action element (+ per-mode sum of squares for the clip norm)
__global__ void q_tc_grad_kernel(QSeg seg, const float* __restrict__ ga1, const float* __restrict__ ga2, int A,
                                 long n_elems, int ascent, float* __restrict__ g, float* __restrict__ gsq) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ float bins[kMaxModes];
    if (threadIdx.x < kMaxModes) bins[threadIdx.x] = 0.f;
    __syncthreads();
    if (i < n_elems) {
        const long row = i / A;
        const int c = (int)(i % A);
        int mode = 0;
        while (mode + 1 < seg.n_modes && row >= seg.off[mode + 1]) ++mode;
        const float scale = ascent ? -seg.inv_cnt[mode] : 1.f;
        const float v = scale * (ga1[row * 16 + c] + ga2[row * 16 + c]);
        g[i] = v;
        if (ascent) atomicAdd(&bins[mode], v * v);
    }
    __syncthreads();
    if (ascent && threadIdx.x < seg.n_modes && bins[threadIdx.x] != 0.f) atomicAdd(gsq + threadIdx.x, bins[threadIdx.x]);
}

// clip_grad_norm_ + torch.optim.Adam.step + clamp_ (ac_base.py:86-91, ddiffpg.py:369); also refreshes the bf16
// action columns of the GEMM input
__global__ void q_tc_adam_kernel(QSeg seg, int O, int A, float* __restrict__ act, const float* __restrict__ g,
                                 float* __restrict__ m1, float* __restrict__ m2, const float* __restrict__ gsq,
                                 float* __restrict__ gnorm_out, int iter, int iters, float step_size, float bc2_sqrt,
                                 float b1, float b2, float eps, float max_norm, float lim, long n_elems,
                                 bf16* __restrict__ xin) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_elems) return;
    const long row = i / A;
    const int c = (int)(i % A);
    int mode = 0;
    while (mode + 1 < seg.n_modes && row >= seg.off[mode + 1]) ++mode;
    const float norm = sqrtf(gsq[mode]);
    const float coef = fminf(max_norm / (norm + 1e-6f), 1.0f);
    if (gnorm_out && i == seg.off[mode] * A) gnorm_out[mode * iters + iter] = norm;
    const float gi = g[i] * coef;
    const float ea = m1[i] * b1 + (1.f - b1) * gi;
    const float ev = m2[i] * b2 + (1.f - b2) * gi * gi;
    m1[i] = ea; m2[i] = ev;
    float v = act[i] - step_size * (ea / (sqrtf(ev) / bc2_sqrt + eps));
    v = fminf(fmaxf(v, -lim), lim);
    act[i] = v;
    if (xin) xin[row * 64 + O + c] = __float2bfloat16(v);
}

__global__ void q_tc_abs_kernel(QSeg seg, int A, const float* __restrict__ act, long n_elems, float* __restrict__ abs_sum) {
    // per-thread partial sums per mode run, one shared atomic per WARP where the warp sits inside one mode (fp32 shared
    // atomics are compare-and-swap loops: one per element on the same address was 61 us per call)
    __shared__ float bins[kMaxModes];
    if (threadIdx.x < kMaxModes) bins[threadIdx.x] = 0.f;
    __syncthreads();
    float local = 0.f;
    int cur = -1;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += (long)gridDim.x * blockDim.x) {
        const long row = i / A;
        int mode = cur < 0 ? 0 : cur;
        while (mode + 1 < seg.n_modes && row >= seg.off[mode + 1]) ++mode;
        if (mode != cur) {
            if (cur >= 0 && local != 0.f) atomicAdd(&bins[cur], local);
            cur = mode; local = 0.f;
        }
        local += fabsf(act[i]);
    }
    const unsigned peers = __match_any_sync(0xffffffffu, cur);
    if (peers == 0xffffffffu) {
        local = warp_sum(local);
        if ((threadIdx.x & 31) == 0 && cur >= 0 && local != 0.f) atomicAdd(&bins[cur], local);
    } else if (cur >= 0 && local != 0.f) {
        atomicAdd(&bins[cur], local);
    }
    __syncthreads();
    if (threadIdx.x < seg.n_modes && bins[threadIdx.x] != 0.f) atomicAdd(abs_sum + threadIdx.x, bins[threadIdx.x]);
}

__global__ void q_tc_finish_kernel(QSeg seg, int A, const float* __restrict__ abs_sum, float* __restrict__ mean_abs) {
    const int m = threadIdx.x;
    if (m < seg.n_modes) {
        const long n = (seg.off[m + 1] - seg.off[m]) * A;
        mean_abs[m] = n > 0 ? abs_sum[m] / (float)n : 0.f;
    }
}

// bf16 operands, all modes back to back per (net, layer)
__global__ void q_tc_pack_kernel(const float* __restrict__ W1, const float* __restrict__ W2, const float* __restrict__ W3,
                                 const float* __restrict__ W4, int O, int A, int atoms, int h1, int h2, int h3, int K1c,
                                 int AtP, bf16* f1, bf16* f2, bf16* f3, bf16* f4, bf16* b4, bf16* b3, bf16* b2, bf16* b1) {
    const int in1 = O + A;
    const size_t n1 = (size_t)h1 * K1c, n2 = (size_t)h2 * h1, n3 = (size_t)h3 * h2, n4 = (size_t)AtP * h3,
                 m4 = (size_t)h3 * AtP, m3 = (size_t)h2 * h3, m2 = (size_t)h1 * h2, m1 = (size_t)16 * h1;
    const size_t total = n1 + n2 + n3 + n4 + m4 + m3 + m2 + m1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t j = i;
        if (j < n1) { const int r = (int)(j / K1c), c = (int)(j % K1c); f1[j] = __float2bfloat16(c < in1 ? W1[(size_t)r * in1 + c] : 0.f); continue; }
        j -= n1;
        if (j < n2) { f2[j] = __float2bfloat16(W2[j]); continue; }
        j -= n2;
        if (j < n3) { f3[j] = __float2bfloat16(W3[j]); continue; }
        j -= n3;
        if (j < n4) { const int r = (int)(j / h3); f4[j] = __float2bfloat16(r < atoms ? W4[j] : 0.f); continue; }
        j -= n4;
        if (j < m4) { const int r = (int)(j / AtP), c = (int)(j % AtP); b4[j] = __float2bfloat16(c < atoms ? W4[(size_t)c * h3 + r] : 0.f); continue; }
        j -= m4;
        if (j < m3) { const int r = (int)(j / h3), c = (int)(j % h3); b3[j] = __float2bfloat16(W3[(size_t)c * h2 + r]); continue; }
        j -= m3;
        if (j < m2) { const int r = (int)(j / h2), c = (int)(j % h2); b2[j] = __float2bfloat16(W2[(size_t)c * h1 + r]); continue; }
        j -= m2;
        { const int r = (int)(j / h1), c = (int)(j % h1); b1[j] = __float2bfloat16(r < A ? W1[(size_t)c * in1 + O + r] : 0.f); }
    }
}

// what the 16-bit pack and the grouped GEMMs need (RND included: 69 inputs -> K1c = 128, 128 features -> AtP = 128)
bool pack_shape_ok(const QLayout& L) {
    return L.K1c <= 256 && L.A <= 16 && L.AtP <= 256 && L.h1 % 64 == 0 && L.h2 % 64 == 0 && L.h3 % 64 == 0;
}
// the critic kernels around them (softmax / BCE: one warp per row over 64 logit columns, 64-column input rows)
bool shape_ok(const QLayout& L) { return pack_shape_ok(L) && L.K1c == 64 && L.AtP == 64; }

QSeg make_seg(const QLayout& L, const int64_t* seg_off, const int64_t* seg_cnt) {
    QSeg s{};
    s.n_modes = L.n_modes;
    for (int m = 0; m <= L.n_modes; ++m) s.off[m] = seg_off[m];
    for (int m = 0; m < L.n_modes; ++m) {
        const long cnt = seg_cnt ? seg_cnt[m] : seg_off[m + 1] - seg_off[m];
        s.inv_cnt[m] = cnt > 0 ? 1.0f / (float)cnt : 0.f;
    }
    return s;
}

// forward of both nets (+ optional backward to the action) for all rows; results in w.logits / w.ga
int q_tc_pass(const QLayout& L, const uint8_t* pb, const QTcWs& w, const QSeg& seg, long B, bool backward,
              float* qmin, float* p1, float* p2, cudaStream_t st) {
    using namespace tcg;
    const float* pk = (const float*)pb;
    auto gemm = [&](const bf16* A, int lda, size_t w_off, size_t w_elems, int ldw, int N, int K, int epi,
                    const float* bias, const bf16* aux, bf16* out_a, float* out_f, int outf_ld, int n_valid) {
        RowGemm g{};
        g.A = A; g.lda = lda; g.W = (const bf16*)(pb + w_off); g.ldw = ldw; g.w_stride = w_elems;
        g.M = B; g.N = N; g.K = K; g.epi = epi; g.bias = bias; g.bias_stride = L.mode_stride;
        g.aux = aux; g.aux_ld = N; g.out_a = out_a; g.out_ld = N; g.out_f = out_f; g.outf_ld = outf_ld; g.n_valid = n_valid;
        g.groups.n_groups = seg.n_modes;
        for (int m = 0; m <= seg.n_modes; ++m) g.groups.off[m] = seg.off[m];
        return launch_row_gemm(g, st);
    };
    int rc;
    for (int j = 0; j < 2; ++j) {
        const QNetLayout& n = L.net[j];
        if ((rc = gemm(w.xin, 64, L.tc_fwd[j][0], L.tc_fwd_elems[0], 64, L.h1, 64, EPI_ELU_FWD, pk + n.b1, nullptr, w.a1[j], nullptr, 0, 0)) != DDP_OK) return rc;
        if ((rc = gemm(w.a1[j], L.h1, L.tc_fwd[j][1], L.tc_fwd_elems[1], L.h1, L.h2, L.h1, EPI_ELU_FWD, pk + n.b2, nullptr, w.a2[j], nullptr, 0, 0)) != DDP_OK) return rc;
        if ((rc = gemm(w.a2[j], L.h2, L.tc_fwd[j][2], L.tc_fwd_elems[2], L.h2, L.h3, L.h2, EPI_ELU_FWD, pk + n.b3, nullptr, w.a3[j], nullptr, 0, 0)) != DDP_OK) return rc;
        if ((rc = gemm(w.a3[j], L.h3, L.tc_fwd[j][3], L.tc_fwd_elems[3], L.h3, 64, L.h3, EPI_LINEAR_F32, pk + n.b4, nullptr, nullptr, w.logits[j], 64, L.atoms)) != DDP_OK) return rc;
    }
    const unsigned wb = (unsigned)((B * 32 + 255) / 256);
    q_tc_softmax_kernel<<<wb, 256, 0, st>>>(w.logits[0], w.logits[1], pk + L.z, L.atoms, B, qmin, p1, p2,
                                            backward ? w.dl[0] : nullptr, backward ? w.dl[1] : nullptr);
    if (!backward) return DDP_OK;
    for (int j = 0; j < 2; ++j) {
        // dZ overwrites the stored activation in place (same thread reads the activation, writes the gradient)
        if ((rc = gemm(w.dl[j], 64, L.tc_bwd[j][0], L.tc_bwd_elems[0], 64, L.h3, 64, EPI_MUL_ELU_D, nullptr, w.a3[j], w.a3[j], nullptr, 0, 0)) != DDP_OK) return rc;
        if ((rc = gemm(w.a3[j], L.h3, L.tc_bwd[j][1], L.tc_bwd_elems[1], L.h3, L.h2, L.h3, EPI_MUL_ELU_D, nullptr, w.a2[j], w.a2[j], nullptr, 0, 0)) != DDP_OK) return rc;
        if ((rc = gemm(w.a2[j], L.h2, L.tc_bwd[j][2], L.tc_bwd_elems[2], L.h2, L.h1, L.h2, EPI_MUL_ELU_D, nullptr, w.a1[j], w.a1[j], nullptr, 0, 0)) != DDP_OK) return rc;
        if ((rc = gemm(w.a1[j], L.h1, L.tc_bwd[j][3], L.tc_bwd_elems[3], L.h1, 16, L.h1, EPI_LINEAR_F32, nullptr, nullptr, nullptr, w.ga[j], 16, L.A)) != DDP_OK) return rc;
    }
    return DDP_OK;
}

}  // namespace

int pack_q_tc(const QLayout& L, const float* const p[], void* packed, cudaStream_t st) {
    if (!pack_shape_ok(L)) DDP_FAIL(DDP_ERR_UNSUPPORTED, "DDP_BF16 critic path needs O+A<=256, A<=16, atoms<=256, widths multiple of 64");
    uint8_t* b = (uint8_t*)packed;
    for (int m = 0; m < L.n_modes; ++m)
        for (int j = 0; j < 2; ++j) {
            const float* const* q = p + 16 * m + 8 * j;
            bf16* f[4]; bf16* r[4];
            for (int i = 0; i < 4; ++i) {
                f[i] = (bf16*)(b + L.tc_fwd[j][i]) + (size_t)m * L.tc_fwd_elems[i];
                r[i] = (bf16*)(b + L.tc_bwd[j][i]) + (size_t)m * L.tc_bwd_elems[i];
            }
            q_tc_pack_kernel<<<296, 256, 0, st>>>(q[0], q[2], q[4], q[6], L.O, L.A, L.atoms, L.h1, L.h2, L.h3, L.K1c, L.AtP,
                                                  f[0], f[1], f[2], f[3], r[0], r[1], r[2], r[3]);
        }
    DDP_LAUNCH_CHECK("q_tc_pack_kernel");
    return DDP_OK;
}

size_t q_tc_workspace(const QLayout& L, long B, int iters) { return carve(L, B, iters, nullptr).total; }

int q_forward_tc(const QLayout& L, const void* packed, const int64_t* seg_off, const float* obs, const float* act,
                 float* qmin, float* p1, float* p2, float* dq_da, long B, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!shape_ok(L)) DDP_FAIL(DDP_ERR_UNSUPPORTED, "DDP_BF16 critic path does not support this shape");
    if (!ws || ws_bytes < carve(L, B, 0, nullptr).total) DDP_FAIL(DDP_ERR_ARG, "critic tensor path: workspace too small");
    QTcWs w = carve(L, B, 0, (uint8_t*)ws);
    QSeg seg = make_seg(L, seg_off, nullptr);
    const unsigned eb = (unsigned)((B * 8 + 255) / 256);
    q_tc_prep_kernel<<<eb, 256, 0, st>>>(obs, const_cast<float*>(act), L.O, L.A, B, 0.f, w.xin, L.K1c);
    if (use_chain(L))
        return q_chain_pass(L, packed, seg_off, nullptr, w.xin, dq_da, nullptr, qmin, p1, p2, B, w.chain_scratch,
                            q_chain_workspace(L), st, nullptr);
    int rc = q_tc_pass(L, (const uint8_t*)packed, w, seg, B, dq_da != nullptr, qmin, p1, p2, st);
    if (rc != DDP_OK) return rc;
    if (dq_da) {
        const long n = B * L.A;
        q_tc_grad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(seg, w.ga[0], w.ga[1], L.A, n, 0, dq_da, nullptr);
    }
    DDP_LAUNCH_CHECK("q_forward_tc kernels");
    return DDP_OK;
}

int q_ascent_tc(const QLayout& L, const void* packed, const int64_t* seg_off, const int64_t* seg_cnt, const float* obs,
                float* action, int iters, float lr, float b1, float b2, float eps, float max_norm, float lim,
                float* mean_abs, float* gnorm_out, long B, void* ws, size_t ws_bytes, cudaStream_t st,
                ddp_gsq_reduce_fn reduce, void* reduce_user) {
    if (!shape_ok(L)) DDP_FAIL(DDP_ERR_UNSUPPORTED, "DDP_BF16 critic path does not support this shape");
    if (ws_bytes < carve(L, B, iters, nullptr).total) DDP_FAIL(DDP_ERR_ARG, "critic tensor path: workspace too small");
    QTcWs w = carve(L, B, iters, (uint8_t*)ws);
    QSeg seg = make_seg(L, seg_off, seg_cnt);
    const long n = B * L.A;
    const unsigned eb = (unsigned)((B * 8 + 255) / 256), nb = (unsigned)((n + 255) / 256);
    DDP_CUDA_CHECK(cudaMemsetAsync(w.m1, 0, w.adam_bytes, st));          // fresh Adam state, zeroed reductions
    q_tc_prep_kernel<<<eb, 256, 0, st>>>(obs, action, L.O, L.A, B, lim, w.xin, L.K1c);
    const bool chain = use_chain(L);
    float neg_inv[kMaxModes];
    for (int m = 0; m < L.n_modes; ++m) neg_inv[m] = -seg.inv_cnt[m];
    // All iterations in ONE cooperative launch (grid barrier on the clip norm, Adam applied by the CTA that owns the
    // rows): 8 % faster while the batch is at most ~half a wave of tiles (1.12 vs 1.22 ms at 256..4096 states: the
    // per-tile latency chain dominates and the launches in between are exposed), equal or slower beyond (4.41 vs
    // 4.33 ms at 65 536: the stream hides the launches).
    long tiles = 0;
    for (int m = 0; m < L.n_modes; ++m) tiles += (seg_off[m + 1] - seg_off[m] + 127) / 128;
    // (an exchange step between the gradient and the Adam step needs the per-iteration form)
    const bool fused = !reduce && (g_q_fused_adam >= 0 ? g_q_fused_adam != 0 : tiles <= 64);
    if (chain && fused && iters >= 1 && iters <= 32) {
        QChainAscent asc{iters, action, w.m1, w.m2, gnorm_out, w.grid_bar, lr, b1, b2, eps, max_norm, lim};
        int rc = q_chain_pass(L, packed, seg_off, neg_inv, w.xin, w.g, w.gsq, nullptr, nullptr, nullptr, B,
                              w.chain_scratch, q_chain_workspace(L), st, &asc);
        if (rc != DDP_OK) return rc;
        iters = 0;                      // nothing left for the per-iteration loop below
    }
    for (int it = 0; it < iters; ++it) {
        float* gsq = w.gsq + (size_t)it * kMaxModes;
        if (chain) {
            int rc = q_chain_pass(L, packed, seg_off, neg_inv, w.xin, w.g, gsq, nullptr, nullptr, nullptr, B,
                                  w.chain_scratch, q_chain_workspace(L), st, nullptr);
            if (rc != DDP_OK) return rc;
        } else {
            int rc = q_tc_pass(L, (const uint8_t*)packed, w, seg, B, true, nullptr, nullptr, nullptr, st);
            if (rc != DDP_OK) return rc;
            q_tc_grad_kernel<<<nb, 256, 0, st>>>(seg, w.ga[0], w.ga[1], L.A, n, 1, w.g, gsq);
        }
        // row-sharded batch: the clip norm is the norm over ALL ranks' rows of the mode (semantics (ii))
        if (reduce && reduce(gsq, L.n_modes, (void*)st, reduce_user) != 0)
            DDP_FAIL(DDP_ERR_ARG, "ddp_q_action_ascent_sharded: the reduce callback failed in iteration %d", it);
        const int step = it + 1;
        const double bc1 = 1.0 - pow((double)b1, step), bc2 = 1.0 - pow((double)b2, step);
        q_tc_adam_kernel<<<nb, 256, 0, st>>>(seg, L.O, L.A, action, w.g, w.m1, w.m2, gsq, gnorm_out, it, iters,
                                             (float)(lr / bc1), (float)sqrt(bc2), b1, b2, eps, max_norm, lim, n, w.xin);
    }
    q_tc_abs_kernel<<<nb < 592 ? nb : 592, 256, 0, st>>>(seg, L.A, action, n, w.abs_sum);
    q_tc_finish_kernel<<<1, kMaxModes, 0, st>>>(seg, L.A, w.abs_sum, mean_abs);
    DDP_LAUNCH_CHECK("q_ascent_tc kernels");
    return DDP_OK;
}


// ------------------------------------------------------------------------------------------------
// N1, bf16 tensor-core path: critic update (AgentDDiffPG.update_critic, ddiffpg/algo/ddiffpg.py:322-351).
//   target  critic_target.get_q1_q2(next_obs, next_actions) by the fused chain kernel (forward half), then the fp32 C51
//           projection of both heads and their minimum (ddiffpg/utils/distl_util.py:4-20, ddiffpg.py:340-346);
//   forward one grouped row GEMM per Linear, ELU fused, activations kept in bf16, logits fp32;
//   loss    softmax + F.binary_cross_entropy (+ d loss / d logits) in fp32, one warp per row;
//   backward per net, last layer first: dW_l = dZ_l^T . a_(l-1) (MN-major dW GEMM, bias gradient from its ones-MMA),
//           then dZ_(l-1) = (dZ_l . W_l) * elu'(a_(l-1)) written over a_(l-1) -- every activation is read by its weight
//           gradient before the backward GEMM overwrites it.
void launch_c51_projection_min(const float* p1, const float* p2, const float* reward, const float* done, float gamma,
                               float v_min, float v_max, int atoms, const float* z, long B, float* target,
                               cudaStream_t st);                              // csrc/q_fma.cu

namespace {

struct QTrainTcWs {
    void* fwd_ws; size_t fwd_bytes;
    bf16 *xin, *a1[2], *a2[2], *a3[2], *dl[2];
    float *logits[2], *p1t, *p2t, *target;
    int* colmap;
    size_t total;
};

QTrainTcWs carve_train(const QLayout& L, long B, uint8_t* base) {
    QTrainTcWs w{};
    size_t o = 0;
    auto take = [&](size_t bytes) { uint8_t* r = base ? base + o : nullptr; o += (bytes + 255) / 256 * 256; return r; };
    w.fwd_bytes = carve(L, B, 0, nullptr).total;
    w.fwd_ws = take(w.fwd_bytes);
    w.xin = (bf16*)take((size_t)B * 64 * 2);
    for (int j = 0; j < 2; ++j) {
        w.a1[j] = (bf16*)take((size_t)B * L.h1 * 2);
        w.a2[j] = (bf16*)take((size_t)B * L.h2 * 2);
        w.a3[j] = (bf16*)take((size_t)B * L.h3 * 2);
        w.dl[j] = (bf16*)take((size_t)B * 64 * 2);
        w.logits[j] = (float*)take((size_t)B * 64 * 4);
    }
    w.p1t = (float*)take((size_t)B * L.atoms * 4);
    w.p2t = (float*)take((size_t)B * L.atoms * 4);
    w.target = (float*)take((size_t)B * L.atoms * 4);
    w.colmap = (int*)take(64 * 4);
    w.total = o;
    return w;
}

__global__ void q_tc_colmap_kernel(int in1, int* __restrict__ colmap) {
    if (threadIdx.x < 64) colmap[threadIdx.x] = (int)threadIdx.x < in1 ? (int)threadIdx.x : -1;
}

// One warp per row, both nets: p = softmax(logits); loss += BCE(p, target) (logs clamped at -100 like
// F.binary_cross_entropy); dl = d loss / d logits = p (g - <p, g>), g = inv_count (p - t) / max(p (1 - p), 1e-12).
__global__ void q_tc_bce_kernel(const float* __restrict__ lg1, const float* __restrict__ lg2,
                                const float* __restrict__ target, int atoms, long B, float inv_count,
                                bf16* __restrict__ dl1, bf16* __restrict__ dl2, float* __restrict__ loss_out) {
    __shared__ float red[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + warp;
    float lsum = 0.f;
    if (row < B) {
        const int c0 = lane, c1 = lane + 32;
        const bool v0 = c0 < atoms, v1 = c1 < atoms;
        const float t0 = v0 ? target[row * atoms + c0] : 0.f, t1 = v1 ? target[row * atoms + c1] : 0.f;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float* lg = (j == 0 ? lg1 : lg2) + row * 64;
            const float l0 = v0 ? lg[c0] : -INFINITY, l1 = v1 ? lg[c1] : -INFINITY;
            // hardware exp2 / log2 / reciprocal (2^-22 relative): far inside the 1e-2 bound of this path, and the
            // kernel is bound by exactly these functions (6 transcendentals + 2 divisions per lane and net)
            const float mx = warp_max(fmaxf(l0, l1));
            const float e0 = v0 ? __expf(l0 - mx) : 0.f, e1 = v1 ? __expf(l1 - mx) : 0.f;
            const float inv = __frcp_rn(warp_sum(e0 + e1));
            const float p0 = e0 * inv, p1 = e1 * inv;
            float g0 = 0.f, g1 = 0.f;
            if (v0) {
                lsum -= t0 * fmaxf(__logf(p0), -100.f) + (1.f - t0) * fmaxf(__logf(1.f - p0), -100.f);
                g0 = __fdividef(inv_count * (p0 - t0), fmaxf((1.f - p0) * p0, 1e-12f));
            }
            if (v1) {
                lsum -= t1 * fmaxf(__logf(p1), -100.f) + (1.f - t1) * fmaxf(__logf(1.f - p1), -100.f);
                g1 = __fdividef(inv_count * (p1 - t1), fmaxf((1.f - p1) * p1, 1e-12f));
            }
            const float dot = warp_sum(p0 * g0 + p1 * g1);
            bf16* dl = (j == 0 ? dl1 : dl2) + row * 64;
            dl[c0] = __float2bfloat16(v0 ? p0 * (g0 - dot) : 0.f);
            dl[c1] = __float2bfloat16(v1 ? p1 * (g1 - dot) : 0.f);
        }
    }
    lsum = warp_sum(lsum);
    if (lane == 0) red[warp] = lsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        if (t != 0.f) atomicAdd(loss_out, t * inv_count);
    }
}

}  // namespace

size_t q_critic_train_tc_workspace(const QLayout& L, long B) { return carve_train(L, B, nullptr).total; }

int q_critic_train_tc(const QLayout& L, const void* packed, const void* packed_target, const float* obs, const float* act,
                      const float* next_obs, const float* next_act, const float* reward, const float* done, float gamma,
                      float* loss_out, float* grads, long B, void* ws, size_t ws_bytes, cudaStream_t st) {
    using namespace tcg;
    if (!shape_ok(L)) DDP_FAIL(DDP_ERR_UNSUPPORTED, "DDP_BF16 critic update does not support this shape");
    if (ws_bytes < carve_train(L, B, nullptr).total) DDP_FAIL(DDP_ERR_ARG, "critic update (tensor path): workspace too small");
    QTrainTcWs w = carve_train(L, B, (uint8_t*)ws);
    const uint8_t* pb = (const uint8_t*)packed;
    const float* pk = (const float*)packed;
    const int64_t seg_off[2] = {0, B};
    int rc = q_forward_tc(L, packed_target, seg_off, next_obs, next_act, nullptr, w.p1t, w.p2t, nullptr, B, w.fwd_ws,
                          w.fwd_bytes, st);
    if (rc != DDP_OK) return rc;
    launch_c51_projection_min(w.p1t, w.p2t, reward, done, gamma, L.v_min, L.v_max, L.atoms,
                              (const float*)packed_target + L.z, B, w.target, st);
    const size_t in1 = (size_t)L.O + L.A;
    const size_t per_net = (size_t)L.h1 * in1 + L.h1 + (size_t)L.h2 * L.h1 + L.h2 + (size_t)L.h3 * L.h2 + L.h3 +
                           (size_t)L.atoms * L.h3 + L.atoms;
    DDP_CUDA_CHECK(cudaMemsetAsync(grads, 0, 2 * per_net * sizeof(float), st));
    const unsigned eb = (unsigned)((B * 8 + 255) / 256);
    q_tc_prep_kernel<<<eb, 256, 0, st>>>(obs, const_cast<float*>(act), L.O, L.A, B, 0.f, w.xin, L.K1c);
    q_tc_colmap_kernel<<<1, 64, 0, st>>>((int)in1, w.colmap);
    auto row = [&](const bf16* A, int lda, size_t w_off, int ldw, int N, int K, int epi, const float* bias, const bf16* aux,
                   bf16* out_a, float* out_f, int outf_ld, int n_valid) {
        RowGemm g{};
        g.A = A; g.lda = lda; g.W = (const bf16*)(pb + w_off); g.ldw = ldw; g.w_stride = 0;
        g.M = B; g.N = N; g.K = K; g.epi = epi; g.bias = bias; g.bias_stride = 0;
        g.aux = aux; g.aux_ld = N; g.out_a = out_a; g.out_ld = N; g.out_f = out_f; g.outf_ld = outf_ld; g.n_valid = n_valid;
        g.groups.n_groups = 1; g.groups.off[0] = 0; g.groups.off[1] = B;
        return launch_row_gemm(g, st);
    };
    for (int j = 0; j < 2; ++j) {
        const QNetLayout& n = L.net[j];
        if ((rc = row(w.xin, 64, L.tc_fwd[j][0], 64, L.h1, 64, EPI_ELU_FWD, pk + n.b1, nullptr, w.a1[j], nullptr, 0, 0)) != DDP_OK) return rc;
        if ((rc = row(w.a1[j], L.h1, L.tc_fwd[j][1], L.h1, L.h2, L.h1, EPI_ELU_FWD, pk + n.b2, nullptr, w.a2[j], nullptr, 0, 0)) != DDP_OK) return rc;
        if ((rc = row(w.a2[j], L.h2, L.tc_fwd[j][2], L.h2, L.h3, L.h2, EPI_ELU_FWD, pk + n.b3, nullptr, w.a3[j], nullptr, 0, 0)) != DDP_OK) return rc;
        if ((rc = row(w.a3[j], L.h3, L.tc_fwd[j][3], L.h3, 64, L.h3, EPI_LINEAR_F32, pk + n.b4, nullptr, nullptr, w.logits[j], 64, L.atoms)) != DDP_OK) return rc;
    }
    q_tc_bce_kernel<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(w.logits[0], w.logits[1], w.target, L.atoms, B,
                                                              1.0f / ((float)B * (float)L.atoms), w.dl[0], w.dl[1], loss_out);
    auto dw = [&](const bf16* dz, int ldz, int N, const bf16* X, int ldx, int K, float* C, int ldc, const int* colmap,
                  float* colsum) {
        DwGemm g{};
        g.dZ = dz; g.ldz = ldz; g.N = N; g.X = X; g.ldx = ldx; g.K = K; g.R = B; g.C = C; g.ldc = ldc; g.colmap = colmap;
        g.colsum = colsum;
        return launch_dw_gemm(g, st);
    };
    for (int j = 0; j < 2; ++j) {
        // flat gradient of one net, state_dict order: W1, b1, W2, b2, W3, b3, W4, b4
        float* g1 = grads + (size_t)j * per_net;
        float* g2 = g1 + (size_t)L.h1 * in1 + L.h1;
        float* g3 = g2 + (size_t)L.h2 * L.h1 + L.h2;
        float* g4 = g3 + (size_t)L.h3 * L.h2 + L.h3;
        if ((rc = dw(w.dl[j], 64, L.atoms, w.a3[j], L.h3, L.h3, g4, L.h3, nullptr, g4 + (size_t)L.atoms * L.h3)) != DDP_OK) return rc;
        if ((rc = row(w.dl[j], 64, L.tc_bwd[j][0], 64, L.h3, 64, EPI_MUL_ELU_D, nullptr, w.a3[j], w.a3[j], nullptr, 0, 0)) != DDP_OK) return rc;
        if ((rc = dw(w.a3[j], L.h3, L.h3, w.a2[j], L.h2, L.h2, g3, L.h2, nullptr, g3 + (size_t)L.h3 * L.h2)) != DDP_OK) return rc;
        if ((rc = row(w.a3[j], L.h3, L.tc_bwd[j][1], L.h3, L.h2, L.h3, EPI_MUL_ELU_D, nullptr, w.a2[j], w.a2[j], nullptr, 0, 0)) != DDP_OK) return rc;
        if ((rc = dw(w.a2[j], L.h2, L.h2, w.a1[j], L.h1, L.h1, g2, L.h1, nullptr, g2 + (size_t)L.h2 * L.h1)) != DDP_OK) return rc;
        if ((rc = row(w.a2[j], L.h2, L.tc_bwd[j][2], L.h2, L.h1, L.h2, EPI_MUL_ELU_D, nullptr, w.a1[j], w.a1[j], nullptr, 0, 0)) != DDP_OK) return rc;
        if ((rc = dw(w.a1[j], L.h1, L.h1, w.xin, 64, 64, g1, (int)in1, w.colmap, g1 + (size_t)L.h1 * in1)) != DDP_OK) return rc;
    }
    DDP_LAUNCH_CHECK("critic update (tensor path) kernels");
    return DDP_OK;
}


// ------------------------------------------------------------------------------------------------
// N4, bf16 tensor-core path: RND / NovelD (ddiffpg/utils/intrinsic.py:62-75, ddiffpg/models/mlp.py:233-267) at the
// update-batch sizes (4 096 rows, NovelD 8 192).  L is the two-net pack of rnd_layout (net 0 = predictor, net 1 =
// target, O = input width, A = 0, atoms = feature width F).  Forward: four row GEMMs per net (ELU fused, features fp32);
// novelty / mse / d loss / d pred in fp32, one warp per row; the predictor's backward as in the critic update above
// (dW by the MN-major GEMM with its ones-MMA bias gradient, dZ written over the activation it no longer needs).
namespace {

struct RndTcWs {
    bf16 *xin, *a1[2], *a2[2], *a3[2], *dl;
    float *feat[2];
    int* colmap;
    size_t total;
};

RndTcWs carve_rnd(const QLayout& L, long B, uint8_t* base) {
    RndTcWs w{};
    size_t o = 0;
    auto take = [&](size_t bytes) { uint8_t* r = base ? base + o : nullptr; o += (bytes + 255) / 256 * 256; return r; };
    w.xin = (bf16*)take((size_t)B * L.K1c * 2);
    for (int j = 0; j < 2; ++j) {
        w.a1[j] = (bf16*)take((size_t)B * L.h1 * 2);
        w.a2[j] = (bf16*)take((size_t)B * L.h2 * 2);
        w.a3[j] = (bf16*)take((size_t)B * L.h3 * 2);
        w.feat[j] = (float*)take((size_t)B * L.AtP * 4);
    }
    w.dl = (bf16*)take((size_t)B * L.AtP * 2);
    w.colmap = (int*)take((size_t)L.K1c * 4);
    w.total = o;
    return w;
}

__global__ void rnd_tc_colmap_kernel(int in1, int K1c, int* __restrict__ colmap) {
    for (int k = threadIdx.x; k < K1c; k += blockDim.x) colmap[k] = k < in1 ? k : -1;
}

// One warp per row: novelty = ||pred - target||_2 (IntrinsicM.get_novelty), optional copies of both feature rows, and for
// the update: loss += inv_count * sum d^2, dl = 2 inv_count d (bf16, zero padded to AtP columns)
__global__ void rnd_tc_mse_kernel(const float* __restrict__ fp, const float* __restrict__ ft, int F, int AtP, long B,
                                  float inv_count, float* __restrict__ novelty, float* __restrict__ pred_out,
                                  float* __restrict__ target_out, bf16* __restrict__ dl, float* __restrict__ loss_out) {
    __shared__ float red[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + warp;
    float ss = 0.f;
    if (row < B) {
        for (int c = lane; c < AtP; c += 32) {
            const float a = c < F ? fp[row * AtP + c] : 0.f, b = c < F ? ft[row * AtP + c] : 0.f;
            const float d = a - b;
            ss = fmaf(d, d, ss);
            if (c < F && pred_out) pred_out[row * F + c] = a;
            if (c < F && target_out) target_out[row * F + c] = b;
            if (dl) dl[row * AtP + c] = __float2bfloat16(2.f * inv_count * d);
        }
    }
    ss = warp_sum(ss);
    if (row < B && lane == 0 && novelty) novelty[row] = sqrtf(ss);
    if (!loss_out) return;
    if (lane == 0) red[warp] = row < B ? ss : 0.f;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        if (t != 0.f) atomicAdd(loss_out, t * inv_count);
    }
}

// forward of `nets` nets (1: predictor only is never needed; 2: predictor + target) into w.feat
int rnd_tc_forward(const QLayout& L, const uint8_t* pb, const RndTcWs& w, const float* x, long B, cudaStream_t st) {
    using namespace tcg;
    const float* pk = (const float*)pb;
    const long per_row = L.K1c >> 3;
    q_tc_prep_kernel<<<(unsigned)((B * per_row + 255) / 256), 256, 0, st>>>(x, nullptr, L.O, 0, B, 0.f, w.xin, L.K1c);
    auto row = [&](const bf16* A, int lda, size_t w_off, int ldw, int N, int K, int epi, const float* bias, bf16* out_a,
                   float* out_f, int outf_ld, int n_valid) {
        RowGemm g{};
        g.A = A; g.lda = lda; g.W = (const bf16*)(pb + w_off); g.ldw = ldw; g.M = B; g.N = N; g.K = K; g.epi = epi;
        g.bias = bias; g.aux_ld = N; g.out_a = out_a; g.out_ld = N; g.out_f = out_f; g.outf_ld = outf_ld; g.n_valid = n_valid;
        g.groups.n_groups = 1; g.groups.off[0] = 0; g.groups.off[1] = B;
        return launch_row_gemm(g, st);
    };
    int rc;
    for (int j = 0; j < 2; ++j) {
        const QNetLayout& n = L.net[j];
        if ((rc = row(w.xin, L.K1c, L.tc_fwd[j][0], L.K1c, L.h1, L.K1c, EPI_ELU_FWD, pk + n.b1, w.a1[j], nullptr, 0, 0)) != DDP_OK) return rc;
        if ((rc = row(w.a1[j], L.h1, L.tc_fwd[j][1], L.h1, L.h2, L.h1, EPI_ELU_FWD, pk + n.b2, w.a2[j], nullptr, 0, 0)) != DDP_OK) return rc;
        if ((rc = row(w.a2[j], L.h2, L.tc_fwd[j][2], L.h2, L.h3, L.h2, EPI_ELU_FWD, pk + n.b3, w.a3[j], nullptr, 0, 0)) != DDP_OK) return rc;
        if ((rc = row(w.a3[j], L.h3, L.tc_fwd[j][3], L.h3, L.AtP, L.h3, EPI_LINEAR_F32, pk + n.b4, nullptr, w.feat[j], L.AtP, L.atoms)) != DDP_OK) return rc;
    }
    return DDP_OK;
}

}  // namespace

size_t rnd_tc_workspace(const QLayout& L, long B) { return carve_rnd(L, B, nullptr).total; }

int rnd_novelty_tc(const QLayout& L, const void* packed, const float* x, float* novelty, float* pred, float* target,
                   long B, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!pack_shape_ok(L)) DDP_FAIL(DDP_ERR_UNSUPPORTED, "DDP_BF16 RND path does not support this shape");
    if (!ws || ws_bytes < carve_rnd(L, B, nullptr).total) DDP_FAIL(DDP_ERR_ARG, "RND (tensor path): workspace too small");
    RndTcWs w = carve_rnd(L, B, (uint8_t*)ws);
    int rc = rnd_tc_forward(L, (const uint8_t*)packed, w, x, B, st);
    if (rc != DDP_OK) return rc;
    rnd_tc_mse_kernel<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(w.feat[0], w.feat[1], L.atoms, L.AtP, B, 0.f, novelty, pred,
                                                                target, nullptr, nullptr);
    DDP_LAUNCH_CHECK("RND novelty (tensor path) kernels");
    return DDP_OK;
}

int rnd_train_tc(const QLayout& L, const void* packed, const float* x, float* loss_out, float* grads, float* novelty,
                 long B, void* ws, size_t ws_bytes, cudaStream_t st) {
    using namespace tcg;
    if (!pack_shape_ok(L)) DDP_FAIL(DDP_ERR_UNSUPPORTED, "DDP_BF16 RND path does not support this shape");
    if (!ws || ws_bytes < carve_rnd(L, B, nullptr).total) DDP_FAIL(DDP_ERR_ARG, "RND (tensor path): workspace too small");
    RndTcWs w = carve_rnd(L, B, (uint8_t*)ws);
    const uint8_t* pb = (const uint8_t*)packed;
    int rc = rnd_tc_forward(L, pb, w, x, B, st);
    if (rc != DDP_OK) return rc;
    const size_t in1 = (size_t)L.O;
    const size_t n_grad = (size_t)L.h1 * in1 + L.h1 + (size_t)L.h2 * L.h1 + L.h2 + (size_t)L.h3 * L.h2 + L.h3 +
                          (size_t)L.atoms * L.h3 + L.atoms;
    DDP_CUDA_CHECK(cudaMemsetAsync(grads, 0, n_grad * sizeof(float), st));
    rnd_tc_colmap_kernel<<<1, 256, 0, st>>>((int)in1, L.K1c, w.colmap);
    rnd_tc_mse_kernel<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(w.feat[0], w.feat[1], L.atoms, L.AtP, B,
                                                                1.0f / ((float)B * (float)L.atoms), novelty, nullptr,
                                                                nullptr, w.dl, loss_out);
    auto row = [&](const bf16* A, int lda, size_t w_off, int ldw, int N, int K, bf16* aux_inout) {
        RowGemm g{};
        g.A = A; g.lda = lda; g.W = (const bf16*)(pb + w_off); g.ldw = ldw; g.M = B; g.N = N; g.K = K; g.epi = EPI_MUL_ELU_D;
        g.aux = aux_inout; g.aux_ld = N; g.out_a = aux_inout; g.out_ld = N;
        g.groups.n_groups = 1; g.groups.off[0] = 0; g.groups.off[1] = B;
        return launch_row_gemm(g, st);
    };
    auto dw = [&](const bf16* dz, int ldz, int N, const bf16* X, int ldx, int K, float* C, int ldc, const int* colmap,
                  float* colsum) {
        DwGemm g{};
        g.dZ = dz; g.ldz = ldz; g.N = N; g.X = X; g.ldx = ldx; g.K = K; g.R = B; g.C = C; g.ldc = ldc; g.colmap = colmap;
        g.colsum = colsum;
        return launch_dw_gemm(g, st);
    };
    float* g1 = grads;                                   // predictor only, state_dict order: W1, b1, ..., W4, b4
    float* g2 = g1 + (size_t)L.h1 * in1 + L.h1;
    float* g3 = g2 + (size_t)L.h2 * L.h1 + L.h2;
    float* g4 = g3 + (size_t)L.h3 * L.h2 + L.h3;
    if ((rc = dw(w.dl, L.AtP, L.atoms, w.a3[0], L.h3, L.h3, g4, L.h3, nullptr, g4 + (size_t)L.atoms * L.h3)) != DDP_OK) return rc;
    if ((rc = row(w.dl, L.AtP, L.tc_bwd[0][0], L.AtP, L.h3, L.AtP, w.a3[0])) != DDP_OK) return rc;
    if ((rc = dw(w.a3[0], L.h3, L.h3, w.a2[0], L.h2, L.h2, g3, L.h2, nullptr, g3 + (size_t)L.h3 * L.h2)) != DDP_OK) return rc;
    if ((rc = row(w.a3[0], L.h3, L.tc_bwd[0][1], L.h3, L.h2, L.h3, w.a2[0])) != DDP_OK) return rc;
    if ((rc = dw(w.a2[0], L.h2, L.h2, w.a1[0], L.h1, L.h1, g2, L.h1, nullptr, g2 + (size_t)L.h2 * L.h1)) != DDP_OK) return rc;
    if ((rc = row(w.a2[0], L.h2, L.tc_bwd[0][2], L.h2, L.h1, L.h2, w.a1[0])) != DDP_OK) return rc;
    if ((rc = dw(w.a1[0], L.h1, L.h1, w.xin, L.K1c, L.K1c, g1, (int)in1, w.colmap, g1 + (size_t)L.h1 * in1)) != DDP_OK) return rc;
    DDP_LAUNCH_CHECK("RND update (tensor path) kernels");
    return DDP_OK;
}

}  // namespace ddp

// Debug entry point (not part of the public header): choose the schedule of the tensor-core critic path explicitly
// (no_chain: 0 / 1; fused_adam: -1 = by batch size, 0, 1).  Process-wide; used by tests and A/B measurements only.
extern "C" void ddp_debug_q_variant(int no_chain, int fused_adam) {
    ddp::g_q_no_chain = no_chain;
    ddp::g_q_fused_adam = fused_adam;
}
