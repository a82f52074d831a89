// H3, bf16 tensor-core path: denoiser eps-loss forward + backward assembled from the tcgen05 row GEMM (one per
// Linear, activation / derivative fused in the TMEM epilogue) and the MN-major dW GEMM (csrc/tc_gemm.cu).
//
// Reference semantics (paths relative to the reference repo):
//   DiffusionPolicy.get_loss (add_noise, net, mse_loss)     ddiffpg/models/diffusion_mlp.py:294-321
//   objective.backward() of optimizer_update                 ddiffpg/algo/ac_base.py:83-85
// Operands bf16, accumulation fp32; loss, d loss/d eps and every gradient accumulate in fp32.  The time branch
// uses the packed fp32 [T, h1] table in the forward (added per row in the epilogue of layer 0) and, in the
// backward, G[t] = sum of dZ0 rows with timestep t obtained as dZ0^T . onehot(t) by the same dW GEMM, followed
// by the fp32 T-row products shared with the fp32 path.
#include "actor_layout.cuh"
#include "tc_gemm.cuh"

namespace ddp {

static bool g_train_no_chain = false;      // set by ddp_debug_train_no_chain (development aid, not in the public header)

void time_branch_backward(const ActorLayout& L, const float* pk, const float* const p[12], const float* G,
                          float* dtemb, float* dhmid, float* g, cudaStream_t st);
bool actor_train_chain_shape_ok(const ActorLayout& L);
int actor_train_chain_fwd(const ActorLayout& L, const void* packed, const void* xin, const int64_t* t, const float* noise,
                          float inv_count, float* loss_out, void* a0, void* d0, void* a1, void* d1, void* a2, void* d2,
                          void* deps, long B, cudaStream_t st);

namespace {

using bf16 = __nv_bfloat16;

struct TrainTcWs {
    bf16 *xin, *onehot, *a0, *d0, *a1, *d1, *a2, *d2, *deps;
    float *eps, *GT, *G, *dtemb, *dhmid;
    size_t total;
};

TrainTcWs carve(const ActorLayout& L, long B, uint8_t* base) {
    TrainTcWs w{};
    size_t o = 0;
    auto take = [&](size_t bytes) { uint8_t* r = base ? base + o : nullptr; o += (bytes + 255) / 256 * 256; return r; };
    const int Tp = (L.T + 7) / 8 * 8;
    w.xin = (bf16*)take((size_t)B * 64 * 2);
    w.onehot = (bf16*)take((size_t)B * Tp * 2);
    w.a0 = (bf16*)take((size_t)B * L.h1 * 2); w.d0 = (bf16*)take((size_t)B * L.h1 * 2);
    w.a1 = (bf16*)take((size_t)B * L.h2 * 2); w.d1 = (bf16*)take((size_t)B * L.h2 * 2);
    w.a2 = (bf16*)take((size_t)B * L.h3 * 2); w.d2 = (bf16*)take((size_t)B * L.h3 * 2);
    w.deps = (bf16*)take((size_t)B * 64 * 2);
    w.eps = (float*)take((size_t)B * 16 * 4);
    w.GT = (float*)take((size_t)L.h1 * Tp * 4);
    w.G = (float*)take((size_t)L.T * L.h1 * 4);
    w.dtemb = (float*)take((size_t)L.T * L.D * 4);
    w.dhmid = (float*)take((size_t)L.T * 4 * L.D * 4);
    w.total = o;
    return w;
}

// For T <= 8 the time branch rides on the layer-0 GEMM itself: columns 48+2t / 49+2t of the input row hold the one-hot
// of the row's timestep and the same columns of the packed W0 hold the [T, h1] time table split into a bf16 "hi" and
// a bf16 "lo" part (hi + lo reproduces the fp32 entry to 2^-17), so the forward needs no per-row table lookup and the
// layer-0 weight-gradient GEMM delivers G[t] = sum of dZ0 rows with timestep t in its "hi" columns.
constexpr int kTCol0 = 48;
__host__ __device__ inline bool time_cols(int T, int S) { return T <= 8 && S + 8 <= kTCol0; }

// xin[r] = [noisy action (A, padded to 8) | state (S) | 0 ...] (64 columns), onehot[r][t_r] = 1
__global__ void train_tc_prep_kernel(const float* __restrict__ state, const float* __restrict__ action,
                                     const float* __restrict__ noise, const int64_t* __restrict__ ts,
                                     const float* __restrict__ cst, int S, int A, int T, int Tp, long B,
                                     bf16* __restrict__ xin, bf16* __restrict__ onehot, int tcols) {
    // one thread = 8 consecutive columns of a row = one 16-byte store
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long row = idx >> 3;
    const int c0 = (int)(idx & 7) * 8;
    if (row >= B) return;
    int t = (int)ts[row];
    t = t < 0 ? 0 : (t >= T ? T - 1 : t);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = c0 + i;
        float x = 0.f;
        if (c < 8) {
            if (c < A) {
                const float* cs = cst + t * kCstStride;       // scheduler.add_noise (diffusion_mlp.py:309-310)
                x = __fadd_rn(__fmul_rn(cs[CST_ADD_A], action[row * A + c]), __fmul_rn(cs[CST_ADD_B], noise[row * A + c]));
            }
        } else if (c - 8 < S) {
            x = state[row * S + c - 8];
        } else if (tcols && c >= kTCol0) {
            x = ((c - kTCol0) >> 1) == t ? 1.f : 0.f;         // one-hot of the timestep, twice (hi / lo table columns)
        }
        v[i] = x;
    }
    __nv_bfloat162 h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(xin + row * 64 + c0) = *reinterpret_cast<const uint4*>(h);
    if (!tcols)                                               // 8 threads per row cover all Tp columns (T up to kMaxT)
        for (int cc = (int)(idx & 7); cc < Tp; cc += 8) onehot[row * Tp + cc] = __float2bfloat16(cc == t ? 1.f : 0.f);
}

__global__ void transpose_small_kernel(const float* __restrict__ src, int rows, int cols, int ld, float* __restrict__ dst) {
    // dst[c][r] = src[r][c]  (src: [rows][ld], dst: [cols][rows])
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    const int r = idx / cols, c = idx % cols;
    dst[(size_t)c * rows + r] = src[(size_t)r * ld + c];
}

// plain bf16 operands for the training GEMMs
__global__ void pack_train_tc_kernel(const float* __restrict__ W0, const float* __restrict__ W1,
                                     const float* __restrict__ W2, const float* __restrict__ W3, int D, int S, int A,
                                     int h1, int h2, int h3, bf16* __restrict__ w0, bf16* __restrict__ w3,
                                     bf16* __restrict__ w3t, bf16* __restrict__ w2t, bf16* __restrict__ w1t,
                                     int* __restrict__ colmap, const float* __restrict__ tb0, int T, int tcols) {
    const size_t n0 = (size_t)h1 * 64, n3 = (size_t)16 * h3, n3t = (size_t)h3 * 64, n2t = (size_t)h2 * h3,
                 n1t = (size_t)h1 * h2;
    const int ld0 = D + S + A;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n0 + n3 + n3t + n2t + n1t + 64;
         i += (size_t)gridDim.x * blockDim.x) {
        size_t j = i;
        if (j < n0) {                                   // w0[f][k]: k < 8 -> x column k, 8 <= k < 8+S -> state
            const int f = (int)(j / 64), k = (int)(j % 64);
            float v = 0.f;
            if (k < 8) { if (k < A) v = W0[(size_t)f * ld0 + D + S + k]; }
            else if (k - 8 < S) v = W0[(size_t)f * ld0 + D + k - 8];
            else if (tcols && k >= kTCol0 && ((k - kTCol0) >> 1) < T) {
                const float tv = tb0[(size_t)((k - kTCol0) >> 1) * h1 + f];
                const float hi = __bfloat162float(__float2bfloat16(tv));
                v = (k & 1) ? tv - hi : hi;
            }
            w0[j] = __float2bfloat16(v);
            continue;
        }
        j -= n0;
        if (j < n3) { const int r = (int)(j / h3), c = (int)(j % h3); w3[j] = __float2bfloat16(r < A ? W3[(size_t)r * h3 + c] : 0.f); continue; }
        j -= n3;
        if (j < n3t) { const int r = (int)(j / 64), c = (int)(j % 64); w3t[j] = __float2bfloat16(c < A ? W3[(size_t)c * h3 + r] : 0.f); continue; }
        j -= n3t;
        if (j < n2t) { const int r = (int)(j / h3), c = (int)(j % h3); w2t[j] = __float2bfloat16(W2[(size_t)c * h2 + r]); continue; }
        j -= n2t;
        if (j < n1t) { const int r = (int)(j / h2), c = (int)(j % h2); w1t[j] = __float2bfloat16(W1[(size_t)c * h1 + r]); continue; }
        j -= n1t;
        const int k = (int)j;
        int cm = k < 8 ? (k < A ? D + S + k : -1) : (k - 8 < S ? D + k - 8 : -1);
        if (tcols && k >= kTCol0 && !(k & 1) && ((k - kTCol0) >> 1) < T) cm = tcg::kDwCol2 + ((k - kTCol0) >> 1);
        colmap[k] = cm;
    }
}

bool shape_ok(const ActorLayout& L) { return L.A <= 8 && L.S + 8 <= 64 && L.h1 % 64 == 0 && L.h2 % 64 == 0 && L.h3 % 64 == 0; }

}  // namespace

int pack_actor_train_tc(const ActorLayout& L, const float* const p[12], void* packed, cudaStream_t st) {
    if (!shape_ok(L)) return DDP_OK;        // the sampler-only shapes simply have no training tensor path
    uint8_t* b = (uint8_t*)packed;
    pack_train_tc_kernel<<<592, 256, 0, st>>>(p[4], p[6], p[8], p[10], L.D, L.S, L.A, L.h1, L.h2, L.h3,
                                               (bf16*)(b + L.tr_w0), (bf16*)(b + L.tr_w3), (bf16*)(b + L.tr_w3t),
                                               (bf16*)(b + L.tr_w2t), (bf16*)(b + L.tr_w1t), (int*)(b + L.tr_colmap),
                                               (const float*)packed + L.tb0, L.T, time_cols(L.T, L.S) ? 1 : 0);
    DDP_LAUNCH_CHECK("pack_train_tc_kernel");
    return DDP_OK;
}

size_t actor_train_tc_workspace(const ActorLayout& L, long B) { return carve(L, B, nullptr).total; }

int actor_train_tc(const ActorLayout& L, const void* packed, const float* const p[12], const float* state,
                   const float* action, const float* noise, const int64_t* t, float inv_count, float* loss_out,
                   float* grads, long B, void* ws, size_t ws_bytes, cudaStream_t st, const cudaEvent_t* group_ev) {
    using namespace tcg;
    if (!shape_ok(L)) DDP_FAIL(DDP_ERR_UNSUPPORTED, "DDP_BF16 training path needs A<=8, S<=56, widths multiple of 64");
    if (ws_bytes < carve(L, B, nullptr).total) DDP_FAIL(DDP_ERR_ARG, "training workspace too small");
    const uint8_t* pb = (const uint8_t*)packed;
    const float* pk = (const float*)packed;
    TrainTcWs w = carve(L, B, (uint8_t*)ws);
    ddp_actor_shape shp{L.S, L.A, L.T, L.D, L.h1, L.h2, L.h3};
    const ActorGradOffsets go = actor_grad_offsets(shp);
    const int Tp = (L.T + 7) / 8 * 8, D = L.D, ld0 = D + L.S + L.A;
    DDP_CUDA_CHECK(cudaMemsetAsync(grads, 0, go.off[12] * sizeof(float), st));
    DDP_CUDA_CHECK(cudaMemsetAsync(w.GT, 0, (size_t)L.h1 * Tp * sizeof(float), st));
    const unsigned eb = (unsigned)((B * 8 + 255) / 256);
    const bool tcols = time_cols(L.T, L.S);
    train_tc_prep_kernel<<<eb, 256, 0, st>>>(state, action, noise, t, pk + L.cst, L.S, L.A, L.T, Tp, B, w.xin, w.onehot, tcols ? 1 : 0);

    auto row = [&](const bf16* A, int lda, const bf16* W, int ldw, int N, int K, int epi, const float* bias,
                   const bf16* aux, bf16* out_a, bf16* out_d, float* out_f, int outf_ld, int n_valid) {
        RowGemm g{};
        g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.M = B; g.N = N; g.K = K; g.epi = epi; g.bias = bias;
        g.aux = aux; g.aux_ld = N; g.out_a = out_a; g.out_d = out_d; g.out_ld = N;
        g.out_f = out_f; g.outf_ld = outf_ld; g.n_valid = n_valid;
        g.groups.n_groups = 1; g.groups.off[0] = 0; g.groups.off[1] = B;
        return g;
    };
    int rc;
    // ---- forward: one fused launch (all four layers on chip, activations / derivatives leave by TMA store) when the
    // shape fits the sampler's tile plan, else one row GEMM per layer (g_train_no_chain: debug switch for A/B runs and
    // for the test that pins the two forwards against each other).
    if (!g_train_no_chain && actor_train_chain_shape_ok(L)) {
        if ((rc = actor_train_chain_fwd(L, packed, w.xin, t, noise, inv_count, loss_out, w.a0, w.d0, w.a1, w.d1, w.a2, w.d2,
                                        w.deps, B, st)) != DDP_OK) return rc;
    } else {
        {
            RowGemm g = row(w.xin, 64, (const bf16*)(pb + L.tr_w0), 64, L.h1, 64, EPI_MISH_FWD, nullptr, nullptr, w.a0, w.d0, nullptr, 0, 0);
            if (!tcols) { g.tbl = pk + L.tb0; g.trow = t; g.tbl_ld = L.h1; g.tbl_rows = L.T; }   // time table (includes b0), per row
            if ((rc = launch_row_gemm(g, st)) != DDP_OK) return rc;
        }
        if ((rc = launch_row_gemm(row(w.a0, L.h1, (const bf16*)(pb + L.tc_w1), L.h1, L.h2, L.h1, EPI_MISH_FWD, pk + L.b1, nullptr, w.a1, w.d1, nullptr, 0, 0), st)) != DDP_OK) return rc;
        if ((rc = launch_row_gemm(row(w.a1, L.h2, (const bf16*)(pb + L.tc_w2), L.h2, L.h3, L.h2, EPI_MISH_FWD, pk + L.b2, nullptr, w.a2, w.d2, nullptr, 0, 0), st)) != DDP_OK) return rc;
        {   // head + mse_loss (:320) + d loss / d eps_hat in the epilogue of the last GEMM
            RowGemm g = row(w.a2, L.h3, (const bf16*)(pb + L.tr_w3), L.h3, 16, L.h3, EPI_MSE_HEAD, pk + L.b3, nullptr, w.deps, nullptr, nullptr, 0, L.A);
            g.out_ld = 64; g.target = noise; g.target_ld = L.A; g.scale = inv_count; g.loss = loss_out;
            if ((rc = launch_row_gemm(g, st)) != DDP_OK) return rc;
        }
    }
    // ---- backward, last layer first: the dZ chain (dZ_l overwrites the stored derivative d_l in place) interleaved
    // with the weight gradients, so that every gradient group is final -- and its all-reduce can start (group_ev) --
    // while the earlier layers are still being differentiated
    auto dw = [&](const bf16* dz, int ldz, int N, const bf16* X, int ldx, int K, float* C, int ldc, const int* colmap,
                  float* C2 = nullptr, int ldc2 = 0, float* colsum = nullptr) {
        DwGemm g{};
        g.dZ = dz; g.ldz = ldz; g.N = N; g.X = X; g.ldx = ldx; g.K = K; g.R = B; g.C = C; g.ldc = ldc; g.colmap = colmap;
        g.C2 = C2; g.ldc2 = ldc2; g.colsum = colsum;
        return launch_dw_gemm(g, st);
    };
    auto mark = [&](int g) -> int {
        if (group_ev && group_ev[g]) DDP_CUDA_CHECK(cudaEventRecord(group_ev[g], st));
        return DDP_OK;
    };
    // bias gradients (column sums of dZ) come out of the same launches as the weight gradients
    if ((rc = dw(w.deps, 64, L.A, w.a2, L.h3, L.h3, grads + go.off[10], L.h3, nullptr, nullptr, 0, grads + go.off[11])) != DDP_OK) return rc;
    if ((rc = mark(0)) != DDP_OK) return rc;
    if ((rc = launch_row_gemm(row(w.deps, 64, (const bf16*)(pb + L.tr_w3t), 64, L.h3, 64, EPI_MUL_D, nullptr, w.d2, w.d2, nullptr, nullptr, 0, 0), st)) != DDP_OK) return rc;
    if ((rc = dw(w.d2, L.h3, L.h3, w.a1, L.h2, L.h2, grads + go.off[8], L.h2, nullptr, nullptr, 0, grads + go.off[9])) != DDP_OK) return rc;
    if ((rc = mark(1)) != DDP_OK) return rc;
    if ((rc = launch_row_gemm(row(w.d2, L.h3, (const bf16*)(pb + L.tr_w2t), L.h3, L.h2, L.h3, EPI_MUL_D, nullptr, w.d1, w.d1, nullptr, nullptr, 0, 0), st)) != DDP_OK) return rc;
    if ((rc = dw(w.d1, L.h2, L.h2, w.a0, L.h1, L.h1, grads + go.off[6], L.h1, nullptr, nullptr, 0, grads + go.off[7])) != DDP_OK) return rc;
    if ((rc = mark(2)) != DDP_OK) return rc;
    if ((rc = launch_row_gemm(row(w.d1, L.h2, (const bf16*)(pb + L.tr_w1t), L.h2, L.h1, L.h2, EPI_MUL_D, nullptr, w.d0, w.d0, nullptr, nullptr, 0, 0), st)) != DDP_OK) return rc;
    // G^T[n][t]: from the one-hot columns of xin in the same launch (T <= 8), else from a separate one-hot operand
    if ((rc = dw(w.d0, L.h1, L.h1, w.xin, 64, 64, grads + go.off[4], ld0, (const int*)(pb + L.tr_colmap), w.GT, Tp)) != DDP_OK) return rc;
    if (!tcols && (rc = dw(w.d0, L.h1, L.h1, w.onehot, Tp, L.T, w.GT, Tp, nullptr)) != DDP_OK) return rc;
    // ---- time branch (fp32, T rows): b0's gradient comes out of it as the column sums of G
    transpose_small_kernel<<<(L.h1 * L.T + 255) / 256, 256, 0, st>>>(w.GT, L.h1, L.T, Tp, w.G);
    time_branch_backward(L, pk, p, w.G, w.dtemb, w.dhmid, grads, st);
    DDP_LAUNCH_CHECK("actor_train_tc kernels");
    return mark(3);
}

}  // namespace ddp

// Debug entry point (not part of the public header): 1 = run the forward of the tensor-core training path as one row
// GEMM per layer instead of the fused on-chip chain (A/B measurements, tests/test_tc_gpu.py).  Process-wide.
extern "C" void ddp_debug_train_no_chain(int on) { ddp::g_train_no_chain = on != 0; }
