// H3, fp32 path: denoiser eps-loss forward and hand-written backward (all 12 parameter gradients),
// plus the clip + AdamW tail of optimizer_update on a flat parameter vector.
//
// Reference semantics (paths relative to the reference repo):
//   DiffusionPolicy.get_loss (add_noise, net, mse_loss)     ddiffpg/models/diffusion_mlp.py:294-321
//   DiffusionNet.forward with per-row integer timesteps      ddiffpg/models/diffusion_mlp.py:62-73
//   objective.backward() / clip_grad_norm_ / AdamW.step      ddiffpg/algo/ac_base.py:83-92
// The time branch (pos-emb -> time_mlp -> 256 time columns of net.mlp.0) only ever sees T distinct
// inputs, so its forward comes from the packed [T, .] tables and its backward is a segment-sum of
// dZ0 by timestep followed by T-row products, instead of B-row GEMMs.
#include <math.h>
#include "actor_layout.cuh"

namespace ddp {

struct TrainWs {
    float *act0, *act1, *act2;      // [B][h1], [B][h2], [B][h3]  post-Mish activations
    float *d0, *d1, *d2;            // same shapes: mish'(z), overwritten by dZ in the backward
    float *xin;                     // [B][K0p]  layer-0 input rows [state | noisy action | 0]
    float *deps;                    // [B][A4]   d loss / d eps_hat
    float *G;                       // [T][h1]   sum of dZ0 rows per timestep
    float *dtemb, *dhmid;           // [T][D], [T][4D]
    size_t total;
};

static TrainWs carve_train_ws(const ActorLayout& L, long B, float* base) {
    TrainWs w{};
    size_t o = 0;
    auto take = [&](size_t n) { float* r = base ? base + o : nullptr; o += (n + 63) / 64 * 64; return r; };
    w.act0 = take((size_t)B * L.h1); w.act1 = take((size_t)B * L.h2); w.act2 = take((size_t)B * L.h3);
    w.d0 = take((size_t)B * L.h1); w.d1 = take((size_t)B * L.h2); w.d2 = take((size_t)B * L.h3);
    w.xin = take((size_t)B * L.K0p);
    w.deps = take((size_t)B * L.A4);
    w.G = take((size_t)L.T * L.h1);
    w.dtemb = take((size_t)L.T * L.D);
    w.dhmid = take((size_t)L.T * 4 * L.D);
    w.total = o * sizeof(float);
    return w;
}

size_t actor_train_workspace(const ActorLayout& L, long B) { return carve_train_ws(L, B, nullptr).total; }

struct TrainArgs {
    const float *wt0, *wt1, *wt2, *wt3, *b1, *b2, *b3, *tb0, *cst, *w3b;
    const float *W1, *W2;           // live net.mlp.2.weight [h2][h1], net.mlp.4.weight [h3][h2]
    int S, A, T, h1, h2, h3, K0p, A4;
    int ks0, ks1, ks2, ks3, kb2, kb1, kb0;
};

__device__ __forceinline__ void mish4(const float4 z, float4& y, float4& d) {
    mish_fd(z.x, y.x, d.x); mish_fd(z.y, y.y, d.y); mish_fd(z.z, y.z, d.z); mish_fd(z.w, y.w, d.w);
}

// -------------------------------------------------------------------------------- forward + loss
template <int RT, int NT>
__global__ void __launch_bounds__(NT) train_fwd_kernel(TrainArgs a, TrainWs w, const float* __restrict__ state,
                                                       const float* __restrict__ action,
                                                       const float* __restrict__ noise,
                                                       const int64_t* __restrict__ ts, float inv_count,
                                                       float* __restrict__ loss_out, long B) {
    extern __shared__ __align__(16) float smem[];
    const int ldA = a.h1 + 4, ldB = a.h2 + 4;
    float* in0 = smem;                      // [RT][K0p]
    float* bufA = in0 + RT * a.K0p;         // [RT][ldA]
    float* bufB = bufA + RT * ldA;          // [RT][ldB]
    __shared__ int tstep[RT];
    __shared__ float red[NT / 32];
    const long row0 = (long)blockIdx.x * RT;
    const int tid = threadIdx.x;

    if (tid < RT) {
        long row = row0 + tid;
        int t = row < B ? (int)ts[row] : 0;
        tstep[tid] = t < 0 ? 0 : (t >= a.T ? a.T - 1 : t);
    }
    __syncthreads();
    for (int i = tid; i < RT * a.K0p; i += NT) {
        int r = i / a.K0p, c = i % a.K0p;
        long row = row0 + r;
        float v = 0.f;
        if (row < B) {
            if (c < a.S) v = state[row * a.S + c];
            else if (c < a.S + a.A) {
                // scheduler.add_noise (diffusion_mlp.py:309): sqrt(abar_t)*a + sqrt(1-abar_t)*n
                const float* cs = a.cst + tstep[r] * kCstStride;
                const int q = c - a.S;
                v = __fadd_rn(__fmul_rn(cs[CST_ADD_A], action[row * a.A + q]), __fmul_rn(cs[CST_ADD_B], noise[row * a.A + q]));
            }
            w.xin[row * a.K0p + c] = v;
        }
        in0[i] = v;
    }
    __syncthreads();

    tile_linear<RT, NT>(a.wt0, a.h1, a.K0p >> 2, a.h1, in0, a.K0p, a.ks0, [&](int r, int n0, float4 v) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(a.tb0 + (size_t)tstep[r] * a.h1 + n0));
        float4 y, d;
        mish4(make_float4(v.x + bb.x, v.y + bb.y, v.z + bb.z, v.w + bb.w), y, d);
        *reinterpret_cast<float4*>(bufA + r * ldA + n0) = y;
        const long row = row0 + r;
        if (row < B) {
            *reinterpret_cast<float4*>(w.act0 + row * a.h1 + n0) = y;
            *reinterpret_cast<float4*>(w.d0 + row * a.h1 + n0) = d;
        }
    });
    __syncthreads();
    tile_linear<RT, NT>(a.wt1, a.h2, a.h1 >> 2, a.h2, bufA, ldA, a.ks1, [&](int r, int n0, float4 v) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(a.b1 + n0));
        float4 y, d;
        mish4(make_float4(v.x + bb.x, v.y + bb.y, v.z + bb.z, v.w + bb.w), y, d);
        *reinterpret_cast<float4*>(bufB + r * ldB + n0) = y;
        const long row = row0 + r;
        if (row < B) {
            *reinterpret_cast<float4*>(w.act1 + row * a.h2 + n0) = y;
            *reinterpret_cast<float4*>(w.d1 + row * a.h2 + n0) = d;
        }
    });
    __syncthreads();
    tile_linear<RT, NT>(a.wt2, a.h3, a.h2 >> 2, a.h3, bufB, ldB, a.ks2, [&](int r, int n0, float4 v) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(a.b2 + n0));
        float4 y, d;
        mish4(make_float4(v.x + bb.x, v.y + bb.y, v.z + bb.z, v.w + bb.w), y, d);
        *reinterpret_cast<float4*>(bufA + r * ldA + n0) = y;
        const long row = row0 + r;
        if (row < B) {
            *reinterpret_cast<float4*>(w.act2 + row * a.h3 + n0) = y;
            *reinterpret_cast<float4*>(w.d2 + row * a.h3 + n0) = d;
        }
    });
    __syncthreads();
    float sq = 0.f;
    tile_linear<RT, NT>(a.wt3, a.A4, a.h3 >> 2, a.A4, bufA, ldA, a.ks3, [&](int r, int n0, float4 v) {
        const long row = row0 + r;
        if (row >= B) return;
        const float e4[4] = {v.x, v.y, v.z, v.w};
        float o4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = n0 + q;
            if (c < a.A) {
                const float diff = (e4[q] + a.b3[c]) - noise[row * a.A + c];       // mse_loss (:320)
                sq = fmaf(diff, diff, sq);
                o4[q] = 2.f * diff * inv_count;
            }
        }
        *reinterpret_cast<float4*>(w.deps + row * a.A4 + n0) = make_float4(o4[0], o4[1], o4[2], o4[3]);
    });
    sq = warp_sum(sq);
    if ((tid & 31) == 0) red[tid >> 5] = sq;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int i = 0; i < NT / 32; ++i) s += red[i];
        atomicAdd(loss_out, s * inv_count);
    }
}

// -------------------------------------------------------------------------------- backward: dZ chain
template <int RT, int NT>
__global__ void __launch_bounds__(NT) train_bwd_dx_kernel(TrainArgs a, TrainWs w, const int64_t* __restrict__ ts, long B) {
    extern __shared__ __align__(16) float smem[];
    const int ldA = a.h1 + 4, ldB = a.h2 + 4;
    float* de = smem;                       // [RT][A4]
    float* bufA = de + RT * a.A4;           // [RT][ldA]   dZ2 then dZ0
    float* bufB = bufA + RT * ldA;          // [RT][ldB]   dZ1
    __shared__ int tstep[RT];
    const long row0 = (long)blockIdx.x * RT;
    const int tid = threadIdx.x;
    if (tid < RT) {
        long row = row0 + tid;
        int t = row < B ? (int)ts[row] : 0;
        tstep[tid] = t < 0 ? 0 : (t >= a.T ? a.T - 1 : t);
    }
    for (int i = tid; i < RT * a.A4; i += NT) {
        long row = row0 + i / a.A4;
        de[i] = row < B ? w.deps[row * a.A4 + i % a.A4] : 0.f;
    }
    __syncthreads();
    // dA2 = dEps . W3 ; dZ2 = dA2 * mish'(Z2)
    tile_linear<RT, NT>(a.w3b, a.h3, a.A4 >> 2, a.h3, de, a.A4, a.kb2, [&](int r, int n0, float4 v) {
        const long row = row0 + r;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < B) {
            float4* dp = reinterpret_cast<float4*>(w.d2 + row * a.h3 + n0);
            const float4 d = *dp;
            o = make_float4(v.x * d.x, v.y * d.y, v.z * d.z, v.w * d.w);
            *dp = o;
        }
        *reinterpret_cast<float4*>(bufA + r * ldA + n0) = o;
    });
    __syncthreads();
    // dA1 = dZ2 . W2 ; dZ1
    tile_linear<RT, NT>(a.W2, a.h2, a.h3 >> 2, a.h2, bufA, ldA, a.kb1, [&](int r, int n0, float4 v) {
        const long row = row0 + r;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < B) {
            float4* dp = reinterpret_cast<float4*>(w.d1 + row * a.h2 + n0);
            const float4 d = *dp;
            o = make_float4(v.x * d.x, v.y * d.y, v.z * d.z, v.w * d.w);
            *dp = o;
        }
        *reinterpret_cast<float4*>(bufB + r * ldB + n0) = o;
    });
    __syncthreads();
    // dA0 = dZ1 . W1 ; dZ0 ; per-timestep column sums feed the time branch
    tile_linear<RT, NT>(a.W1, a.h1, a.h2 >> 2, a.h1, bufB, ldB, a.kb0, [&](int r, int n0, float4 v) {
        const long row = row0 + r;
        if (row >= B) return;
        float4* dp = reinterpret_cast<float4*>(w.d0 + row * a.h1 + n0);
        const float4 d = *dp;
        const float4 o = make_float4(v.x * d.x, v.y * d.y, v.z * d.z, v.w * d.w);
        *dp = o;
        float* g = w.G + (size_t)tstep[r] * a.h1 + n0;
        atomicAdd(g, o.x); atomicAdd(g + 1, o.y); atomicAdd(g + 2, o.z); atomicAdd(g + 3, o.w);
    });
}

// -------------------------------------------------------------------------------- dW = dZ^T . X
// C[n*ldc + k] += sum_r dz[r*ldz + n] * x[r*ldx + k]  (n < N, k < K), rows split over gridDim.z;
// dbias[n] += sum_r dz[r][n] from the k-tile-0 CTAs.  64x64 output tile, 16-row chunks, 4x4 per thread.
__global__ void __launch_bounds__(256) dw_gemm_kernel(const float* __restrict__ dz, int ldz, int N,
                                                      const float* __restrict__ x, int ldx, int K,
                                                      float* __restrict__ C, int ldc, float* __restrict__ dbias,
                                                      long R, long rows_per_split) {
    __shared__ __align__(16) float sz[16][64];
    __shared__ __align__(16) float sx[16][64];
    const int n_base = blockIdx.x * 64, k_base = blockIdx.y * 64;
    const long r_begin = (long)blockIdx.z * rows_per_split;
    const long r_end = r_begin + rows_per_split < R ? r_begin + rows_per_split : R;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int lr = tid >> 4, lc = (tid & 15) * 4;        // loader: row lr of the chunk, 4 columns at lc
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
    const bool vec_z = (ldz % 4 == 0) && (n_base + 64 <= N);
    const bool vec_x = (ldx % 4 == 0) && (k_base + 64 <= K);
    for (long r0 = r_begin; r0 < r_end; r0 += 16) {
        const long r = r0 + lr;
        float4 vz = make_float4(0.f, 0.f, 0.f, 0.f), vx = vz;
        if (r < r_end) {
            const float* pz = dz + r * ldz + n_base + lc;
            const float* px = x + r * ldx + k_base + lc;
            if (vec_z) vz = *reinterpret_cast<const float4*>(pz);
            else {
                if (n_base + lc + 0 < N) vz.x = pz[0];
                if (n_base + lc + 1 < N) vz.y = pz[1];
                if (n_base + lc + 2 < N) vz.z = pz[2];
                if (n_base + lc + 3 < N) vz.w = pz[3];
            }
            if (vec_x) vx = *reinterpret_cast<const float4*>(px);
            else {
                if (k_base + lc + 0 < K) vx.x = px[0];
                if (k_base + lc + 1 < K) vx.y = px[1];
                if (k_base + lc + 2 < K) vx.z = px[2];
                if (k_base + lc + 3 < K) vx.w = px[3];
            }
        }
        __syncthreads();
        *reinterpret_cast<float4*>(&sz[lr][lc]) = vz;
        *reinterpret_cast<float4*>(&sx[lr][lc]) = vx;
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) {
            const float4 zq = *reinterpret_cast<const float4*>(&sz[rr][ty * 4]);
            const float4 xq = *reinterpret_cast<const float4*>(&sx[rr][tx * 4]);
            const float zv[4] = {zq.x, zq.y, zq.z, zq.w}, xv[4] = {xq.x, xq.y, xq.z, xq.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(zv[i], xv[j], acc[i][j]);
                bsum[i] += zv[i];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n_base + ty * 4 + i;
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k_base + tx * 4 + j;
            if (k < K) atomicAdd(C + (size_t)n * ldc + k, acc[i][j]);
        }
        if (dbias && blockIdx.y == 0 && tx == 0) atomicAdd(dbias + n, bsum[i]);
    }
}

void launch_dw(const float* dz, int ldz, int N, const float* x, int ldx, int K, float* C, int ldc,
                      float* dbias, long R, cudaStream_t st) {
    dim3 grid((N + 63) / 64, (K + 63) / 64, 1);
    // enough row splits for ~4 waves of 148 SMs, at least 64 rows per split
    long tiles = (long)grid.x * grid.y;
    long want = (592 + tiles - 1) / tiles;
    long max_splits = (R + 63) / 64;
    long splits = want < max_splits ? want : max_splits;
    if (splits < 1) splits = 1;
    long rps = ((R + splits - 1) / splits + 15) / 16 * 16;
    splits = (R + rps - 1) / rps;
    grid.z = (unsigned)splits;
    dw_gemm_kernel<<<grid, 256, 0, st>>>(dz, ldz, N, x, ldx, K, C, ldc, dbias, R, rps);
}

// y[t][i] = sum_n g[t*ldg + n] * W[n*ldw + i]  (i < M, n < N); optional multiply by mish'(zmul[t][i]).
// Block = 32 output columns x 8 warps that split n; W is read coalesced along i.
__global__ void rows_linear_t_kernel(const float* __restrict__ g, int ldg, int N, const float* __restrict__ W,
                                     int ldw, int M, const float* __restrict__ zmul, float* __restrict__ y) {
    __shared__ float part[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane, t = blockIdx.y;
    const float* gr = g + (size_t)t * ldg;
    float acc = 0.f;
    if (i < M) {
        // eight independent loads in flight per thread: the loop is a latency chain of L2 reads otherwise
        float a8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int n = warp;
        for (; n + 56 < N; n += 64) {
#pragma unroll
            for (int u = 0; u < 8; ++u) a8[u] = fmaf(gr[n + 8 * u], __ldg(W + (size_t)(n + 8 * u) * ldw + i), a8[u]);
        }
        for (; n < N; n += 8) a8[0] = fmaf(gr[n], __ldg(W + (size_t)n * ldw + i), a8[0]);
        acc = ((a8[0] + a8[1]) + (a8[2] + a8[3])) + ((a8[4] + a8[5]) + (a8[6] + a8[7]));
    }
    part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && i < M) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += part[w][lane];
        if (zmul) {
            float yv, dv;
            mish_fd(zmul[(size_t)t * M + i], yv, dv);
            s *= dv;
        }
        y[(size_t)t * M + i] = s;
    }
}

// Backward of the time branch on its T distinct rows, given G[t] = sum of dZ0 rows with timestep t:
// gradients of the 256 time columns of net.mlp.0 (+ its bias) and of both time_mlp layers.
void time_branch_backward(const ActorLayout& L, const float* pk, const float* const p[12], const float* G,
                          float* dtemb, float* dhmid, float* g, cudaStream_t st) {
    ddp_actor_shape shp{L.S, L.A, L.T, L.D, L.h1, L.h2, L.h3};
    const ActorGradOffsets go = actor_grad_offsets(shp);
    const int D = L.D, ld0 = D + L.S + L.A;
    launch_dw(G, L.h1, L.h1, pk + L.temb, D, D, g + go.off[4], ld0, g + go.off[5], L.T, st);
    dim3 gt((D + 31) / 32, L.T), gm((4 * D + 31) / 32, L.T);
    rows_linear_t_kernel<<<gt, 256, 0, st>>>(G, L.h1, L.h1, p[4], ld0, D, nullptr, dtemb);               // dtemb = G . W0[:, :D]
    launch_dw(dtemb, D, D, pk + L.hmid, 4 * D, 4 * D, g + go.off[2], 4 * D, g + go.off[3], L.T, st);
    rows_linear_t_kernel<<<gm, 256, 0, st>>>(dtemb, D, D, p[2], 4 * D, 4 * D, pk + L.zmid, dhmid);        // dZmid
    launch_dw(dhmid, 4 * D, 4 * D, pk + L.pe, D, D, g + go.off[0], D, g + go.off[1], L.T, st);
}

template <int RT>
static int launch_train(const ActorLayout& L, const TrainArgs& a, const TrainWs& w, const float* state,
                        const float* action, const float* noise, const int64_t* t, float inv_count, float* loss_out,
                        long B, cudaStream_t st) {
    constexpr int NT = 256;
    const size_t smem_f = sizeof(float) * ((size_t)RT * L.K0p + (size_t)RT * (L.h1 + 4) + (size_t)RT * (L.h2 + 4));
    const size_t smem_b = sizeof(float) * ((size_t)RT * L.A4 + (size_t)RT * (L.h1 + 4) + (size_t)RT * (L.h2 + 4));
    if (smem_f > 227 * 1024 || smem_b > 227 * 1024) DDP_FAIL(DDP_ERR_SHAPE, "training tile does not fit shared memory");
    auto kf = train_fwd_kernel<RT, NT>;
    auto kb = train_bwd_dx_kernel<RT, NT>;
    DDP_CUDA_CHECK(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
    DDP_CUDA_CHECK(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
    const unsigned grid = (unsigned)((B + RT - 1) / RT);
    kf<<<grid, NT, smem_f, st>>>(a, w, state, action, noise, t, inv_count, loss_out, B);
    kb<<<grid, NT, smem_b, st>>>(a, w, t, B);
    DDP_LAUNCH_CHECK("train fwd/bwd kernels");
    return DDP_OK;
}

int actor_train_fma(const ActorLayout& L, const float* pk, const float* const p[12], const float* state,
                    const float* action, const float* noise, const int64_t* t, float inv_count, float* loss_out,
                    float* grads, long B, void* ws, size_t ws_bytes, cudaStream_t st) {
    (void)ws_bytes;
    TrainWs w = carve_train_ws(L, B, (float*)ws);
    ddp_actor_shape shp{L.S, L.A, L.T, L.D, L.h1, L.h2, L.h3};
    const ActorGradOffsets go = actor_grad_offsets(shp);
    DDP_CUDA_CHECK(cudaMemsetAsync(grads, 0, go.off[12] * sizeof(float), st));
    DDP_CUDA_CHECK(cudaMemsetAsync(w.G, 0, (size_t)L.T * L.h1 * sizeof(float), st));

    TrainArgs a;
    a.wt0 = pk + L.wt0; a.wt1 = pk + L.wt1; a.wt2 = pk + L.wt2; a.wt3 = pk + L.wt3;
    a.b1 = pk + L.b1; a.b2 = pk + L.b2; a.b3 = pk + L.b3; a.tb0 = pk + L.tb0; a.cst = pk + L.cst; a.w3b = pk + L.w3b;
    a.W1 = p[6]; a.W2 = p[8];
    a.S = L.S; a.A = L.A; a.T = L.T; a.h1 = L.h1; a.h2 = L.h2; a.h3 = L.h3; a.K0p = L.K0p; a.A4 = L.A4;
    a.ks0 = pick_ksplit(L.h1, 256); a.ks1 = pick_ksplit(L.h2, 256); a.ks2 = pick_ksplit(L.h3, 256);
    a.ks3 = pick_ksplit(L.A4, 256);
    a.kb2 = pick_ksplit(L.h3, 256); a.kb1 = pick_ksplit(L.h2, 256); a.kb0 = pick_ksplit(L.h1, 256);
    int rc = (B <= 148 * 8) ? launch_train<4>(L, a, w, state, action, noise, t, inv_count, loss_out, B, st)
                            : launch_train<16>(L, a, w, state, action, noise, t, inv_count, loss_out, B, st);
    if (rc != DDP_OK) return rc;

    const int D = L.D, ld0 = D + L.S + L.A;
    float* g = grads;
    // trunk: dW_l = dZ_l^T . A_{l-1}, db_l = column sums of dZ_l
    launch_dw(w.deps, L.A4, L.A, w.act2, L.h3, L.h3, g + go.off[10], L.h3, g + go.off[11], B, st);
    launch_dw(w.d2, L.h3, L.h3, w.act1, L.h2, L.h2, g + go.off[8], L.h2, g + go.off[9], B, st);
    launch_dw(w.d1, L.h2, L.h2, w.act0, L.h1, L.h1, g + go.off[6], L.h1, g + go.off[7], B, st);
    launch_dw(w.d0, L.h1, L.h1, w.xin, L.K0p, L.S + L.A, g + go.off[4] + D, ld0, nullptr, B, st);
    time_branch_backward(L, pk, p, w.G, w.dtemb, w.dhmid, g, st);
    DDP_LAUNCH_CHECK("train dW kernels");
    return DDP_OK;
}

// -------------------------------------------------------------------------------- clip + AdamW
// ||g||^2 in a FIXED summation order: every block leaves its partial sum in partials[blockIdx.x] (no atomics), and
// every consumer adds the partials up in the same order (block_total).  Data-parallel replicas that hold the same
// reduced gradient therefore compute bit-identical norms, clip coefficients and parameters.
constexpr int kNormBlocks = 592;           // partial sums; scratch holds kScratchPartials + kNormBlocks floats
constexpr int kScratchPartials = 8;
__global__ void sumsq_kernel(const float* __restrict__ g, size_t n, float* __restrict__ partials) {
    __shared__ float red[8];
    float s = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        s = fmaf(g[i], g[i], s);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        partials[blockIdx.x] = t;
    }
}
// sum of `count` partials, identical in every thread of every block (256 threads)
__device__ __forceinline__ float block_total(const float* __restrict__ partials, int count) {
    __shared__ float red[8];
    __shared__ float total;
    float s = 0.f;
    for (int i = threadIdx.x; i < count; i += 256) s += partials[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        total = t;
    }
    __syncthreads();
    return total;
}

// torch.optim.AdamW single-tensor semantics: p *= 1 - lr*wd; m.lerp_(g, 1-b1); v = b2*v + (1-b2) g^2;
// p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps), after clip_grad_norm_ scaled g in place.
__global__ void clip_adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                  float* __restrict__ v, size_t n, const float* __restrict__ partials, int n_partials, float decay,
                                  float step_size, float bc2_sqrt, float b1, float b2, float eps, float max_norm,
                                  float* __restrict__ norm_out, const float* __restrict__ dev_scalars) {
    if (dev_scalars) { step_size = dev_scalars[0]; bc2_sqrt = dev_scalars[1]; }     // step count kept on the device
    const float norm = sqrtf(block_total(partials, n_partials));
    const float coef = fminf(max_norm / (norm + 1e-6f), 1.0f);
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *norm_out = norm;
    if (i >= n) return;
    const float gi = g[i] * coef;
    g[i] = gi;
    const float ea = m[i] + (1.f - b1) * (gi - m[i]);
    const float ev = v[i] * b2 + (1.f - b2) * gi * gi;
    m[i] = ea; v[i] = ev;
    const float denom = sqrtf(ev) / bc2_sqrt + eps;
    p[i] = p[i] * decay - step_size * (ea / denom);
}

int clip_adamw(float* p, float* g, float* m, float* v, size_t n, int step, float lr, float b1, float b2, float eps,
               float wd, float max_norm, float* norm_out, float* scratch, cudaStream_t st) {
    unsigned gb = (unsigned)((n + 255) / 256);
    const int nb = gb < (unsigned)kNormBlocks ? (int)gb : kNormBlocks;
    sumsq_kernel<<<nb, 256, 0, st>>>(g, n, scratch + kScratchPartials);
    const double bc1 = 1.0 - pow((double)b1, step), bc2 = 1.0 - pow((double)b2, step);
    clip_adamw_kernel<<<gb, 256, 0, st>>>(p, g, m, v, n, scratch + kScratchPartials, nb, (float)(1.0 - (double)lr * wd), (float)(lr / bc1),
                                         (float)sqrt(bc2), b1, b2, eps, max_norm, norm_out, nullptr);
    DDP_LAUNCH_CHECK("clip_adamw kernels");
    return DDP_OK;
}

// The same step with the step count on the device (incremented by the call), so that the launch sequence carries
// no per-step host value and can be captured once into a CUDA graph.  scratch: DDP_ADAMW_SCRATCH_FLOATS.
__global__ void adam_scalars_kernel(int* __restrict__ step_dev, float lr, float b1, float b2, float* __restrict__ out) {
    const int step = *step_dev + 1;
    *step_dev = step;
    const double bc1 = 1.0 - pow((double)b1, (double)step), bc2 = 1.0 - pow((double)b2, (double)step);
    out[0] = (float)((double)lr / bc1);
    out[1] = (float)sqrt(bc2);
}

int clip_adamw_dev(float* p, float* g, float* m, float* v, size_t n, int* step_dev, float lr, float b1, float b2,
                   float eps, float wd, float max_norm, float* norm_out, float* scratch, cudaStream_t st) {
    unsigned gb = (unsigned)((n + 255) / 256);
    const int nb = gb < (unsigned)kNormBlocks ? (int)gb : kNormBlocks;
    sumsq_kernel<<<nb, 256, 0, st>>>(g, n, scratch + kScratchPartials);
    adam_scalars_kernel<<<1, 1, 0, st>>>(step_dev, lr, b1, b2, scratch + 1);
    clip_adamw_kernel<<<gb, 256, 0, st>>>(p, g, m, v, n, scratch + kScratchPartials, nb, (float)(1.0 - (double)lr * wd), 0.f, 1.f, b1, b2, eps,
                                         max_norm, norm_out, scratch + 1);
    DDP_LAUNCH_CHECK("clip_adamw (device step) kernels");
    return DDP_OK;
}

}  // namespace ddp
