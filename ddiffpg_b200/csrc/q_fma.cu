// H2, fp32 warp-level FMA path: mode-segmented distributional double-Q forward, its analytic
// gradient with respect to the action, and the Adam action-ascent loop.
//
// Reference semantics (paths relative to the reference repo):
//   MLPNet / create_simple_mlp (ELU)                 ddiffpg/models/mlp.py:13-35
//   DistributionalDoubleQ.get_q1_q2 / get_q_min      ddiffpg/models/mlp.py:143-151
//   AgentDDiffPG.update_target_action                ddiffpg/algo/ddiffpg.py:358-373
//   ActorCriticBase.optimizer_update (clip + step)   ddiffpg/algo/ac_base.py:83-92
// One CTA owns a tile of rows of ONE mode segment; both nets' activations stay in shared memory
// between the forward and the hand-written backward (the critic weights are frozen, :359, so only
// dX products are needed).  The two cross-row couplings of the reference -- the 1/B of Q.mean() and
// the global L2 clip norm -- are a per-segment scalar each, reduced with one atomic per CTA.
#include <math.h>
#include "q_layout.cuh"

namespace ddp {

struct SegTable {
    long off[kMaxModes + 1];         // row offsets per mode
    long tile0[kMaxModes + 1];       // first tile index per mode
    float inv_cnt[kMaxModes];        // 1 / rows the reference's Q.mean() averages over
    int n_modes;
};

struct QArgs {
    const float* packed; size_t mode_stride;
    QNetLayout net[2]; size_t z;
    int O, A, atoms, h1, h2, h3, K1p, A4, atomsP;
    int ks1, ks2, ks3, ks4, kb4, kb3, kb2, kb1;
};

// Global buffers of the critic training step (MODE 2): everything the weight-gradient GEMMs need.
struct QTrainBufs {
    float *xin;                     // [B][K1p]  [obs | act | 0]
    float *a1[2], *a2[2], *a3[2];   // activations, then overwritten by dZ1, dZ2, dZ3 AFTER the copies below
    float *z1[2], *z2[2], *z3[2];   // dZ1, dZ2, dZ3
    float *dl[2];                   // [B][atomsP] d loss / d logits
    const float* target;            // [B][atoms]   projected target distribution
    float inv_count;                // 1 / (B * atoms): mean of binary_cross_entropy
    float* loss_out;
};

static QArgs make_qargs(const QLayout& L, const float* pk, int NT) {
    QArgs a;
    a.packed = pk; a.mode_stride = L.mode_stride; a.net[0] = L.net[0]; a.net[1] = L.net[1]; a.z = L.z;
    a.O = L.O; a.A = L.A; a.atoms = L.atoms; a.h1 = L.h1; a.h2 = L.h2; a.h3 = L.h3;
    a.K1p = L.K1p; a.A4 = L.A4; a.atomsP = L.atomsP;
    a.ks1 = pick_ksplit(L.h1, NT); a.ks2 = pick_ksplit(L.h2, NT); a.ks3 = pick_ksplit(L.h3, NT);
    a.ks4 = pick_ksplit(L.atomsP, NT);
    a.kb4 = pick_ksplit(L.h3, NT); a.kb3 = pick_ksplit(L.h2, NT); a.kb2 = pick_ksplit(L.h1, NT);
    a.kb1 = pick_ksplit(L.A4, NT);
    return a;
}

__global__ void q_support_kernel(float* __restrict__ z, int atoms, int atomsP, float v_min, float v_max) {
    // torch.linspace (mlp.py:141): filled from both ends with step = (end-start)/(steps-1)
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= atomsP) return;
    float step = (v_max - v_min) / (float)(atoms - 1);
    float v = 0.f;
    if (i < atoms) v = (i < atoms / 2) ? v_min + step * (float)i : v_max - step * (float)(atoms - i - 1);
    z[i] = v;
}

// out[r][c] = src[r*ld + off + c] for c < cols, 0 for cols <= c < ldo
__global__ void submatrix_pack_kernel(const float* __restrict__ src, int ld, int off, int rows, int cols, int ldo,
                                      float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * ldo) return;
    int r = i / ldo, c = i % ldo;
    out[i] = c < cols ? src[(size_t)r * ld + off + c] : 0.f;
}

// the fp32 operands the tensor-core path reads: the eight bias vectors of one critic and the support atoms, one launch
struct QBiasPack { const float* src[8]; size_t dst[8]; int n[8], np[8]; };
__global__ void q_bias_pack_kernel(QBiasPack bp, float* __restrict__ base, size_t z_off, int atoms, int atomsP,
                                   float v_min, float v_max) {
    const int v = blockIdx.x;
    if (v < 8) {
        for (int i = threadIdx.x; i < bp.np[v]; i += blockDim.x) base[bp.dst[v] + i] = i < bp.n[v] ? bp.src[v][i] : 0.f;
    } else {
        const float step = (v_max - v_min) / (float)(atoms - 1);       // torch.linspace, as q_support_kernel
        for (int i = threadIdx.x; i < atomsP; i += blockDim.x)
            base[z_off + i] = i < atoms ? ((i < atoms / 2) ? v_min + step * (float)i : v_max - step * (float)(atoms - i - 1)) : 0.f;
    }
}

// bias_only: the DDP_BF16 pack -- every matrix the tensor path multiplies lives in the 16-bit section (pack_q_tc), the
// fp32 section only supplies biases and atoms (1 launch per critic instead of 25)
int pack_q_fp32(const QLayout& L, const float* const p[], float* out, cudaStream_t st, bool bias_only) {
    auto blocks = [](size_t n) { return (unsigned)((n + 255) / 256); };
    const int in1 = L.O + L.A;
    for (int m = 0; m < L.n_modes && bias_only; ++m) {
        QBiasPack bp;
        for (int j = 0; j < 2; ++j) {
            const float* const* q = p + 16 * m + 8 * j;
            const QNetLayout& n = L.net[j];
            const size_t dst[4] = {n.b1, n.b2, n.b3, n.b4};
            const int cnt[4] = {L.h1, L.h2, L.h3, L.atoms}, cntp[4] = {L.h1, L.h2, L.h3, L.atomsP};
            for (int i = 0; i < 4; ++i) { bp.src[4 * j + i] = q[2 * i + 1]; bp.dst[4 * j + i] = dst[i]; bp.n[4 * j + i] = cnt[i]; bp.np[4 * j + i] = cntp[i]; }
        }
        q_bias_pack_kernel<<<9, 256, 0, st>>>(bp, out + (size_t)m * L.mode_stride, L.z, L.atoms, L.atomsP, L.v_min, L.v_max);
    }
    if (bias_only) { DDP_LAUNCH_CHECK("critic bias pack kernel"); return DDP_OK; }
    for (int m = 0; m < L.n_modes; ++m) {
        float* base = out + (size_t)m * L.mode_stride;
        for (int j = 0; j < 2; ++j) {
            const float* const* q = p + 16 * m + 8 * j;       // W1,b1,W2,b2,W3,b3,W4,b4 of net j
            const QNetLayout& n = L.net[j];
            transpose_pack_kernel<<<blocks((size_t)L.K1p * L.h1), 256, 0, st>>>(q[0], in1, 0, in1, L.K1p, L.h1, L.h1, base + n.wt1);
            transpose_pack_kernel<<<blocks((size_t)L.h1 * L.h2), 256, 0, st>>>(q[2], L.h1, 0, L.h1, L.h1, L.h2, L.h2, base + n.wt2);
            transpose_pack_kernel<<<blocks((size_t)L.h2 * L.h3), 256, 0, st>>>(q[4], L.h2, 0, L.h2, L.h2, L.h3, L.h3, base + n.wt3);
            transpose_pack_kernel<<<blocks((size_t)L.h3 * L.atomsP), 256, 0, st>>>(q[6], L.h3, 0, L.h3, L.h3, L.atoms, L.atomsP, base + n.wt4);
            copy_pad_kernel<<<blocks(L.h1), 256, 0, st>>>(q[1], L.h1, L.h1, base + n.b1);
            copy_pad_kernel<<<blocks(L.h2), 256, 0, st>>>(q[3], L.h2, L.h2, base + n.b2);
            copy_pad_kernel<<<blocks(L.h3), 256, 0, st>>>(q[5], L.h3, L.h3, base + n.b3);
            copy_pad_kernel<<<blocks(L.atomsP), 256, 0, st>>>(q[7], L.atoms, L.atomsP, base + n.b4);
            // backward operands: dX = dY . W, streamed as [contraction = out features][in features]
            copy_pad_kernel<<<blocks((size_t)L.atomsP * L.h3), 256, 0, st>>>(q[6], L.atoms * L.h3, L.atomsP * L.h3, base + n.w4b);
            copy_pad_kernel<<<blocks((size_t)L.h3 * L.h2), 256, 0, st>>>(q[4], L.h3 * L.h2, L.h3 * L.h2, base + n.w3b);
            copy_pad_kernel<<<blocks((size_t)L.h2 * L.h1), 256, 0, st>>>(q[2], L.h2 * L.h1, L.h2 * L.h1, base + n.w2b);
            if (L.A4 > 0)
                submatrix_pack_kernel<<<blocks((size_t)L.h1 * L.A4), 256, 0, st>>>(q[0], in1, L.O, L.h1, L.A, L.A4, base + n.w1a);
        }
        q_support_kernel<<<1, 64, 0, st>>>(base + L.z, L.atoms, L.atomsP, L.v_min, L.v_max);
    }
    DDP_LAUNCH_CHECK("critic pack kernels");
    return DDP_OK;
}

// ------------------------------------------------------------------------------------------------
// Forward (+ optional backward to the action) for one row tile.
//   MODE 0: inference outputs (q_min, p1, p2, dq_da) -- get_q1_q2 / get_q_min and their autograd.
//   MODE 1: one ascent iteration: g = -inv_cnt * d min(Q1,Q2)/da, plus sum g^2 into gsq[mode].
//   MODE 2: critic training step (AgentDDiffPG.update_critic, ddiffpg.py:348-349): BCE(current_Q1, target) +
//           BCE(current_Q2, target), softmax backward, dZ chain; activations and dZ go to global memory for the
//           weight-gradient GEMMs.
//   MODE 3: RNDModel (mlp.py:233-267): net 0 = predictor, net 1 = target, `atoms` = feature width.  Novelty
//           ||pred - target||_2 per row (IntrinsicM.get_novelty, utils/intrinsic.py:62-65) into qmin_out; with
//           tr.loss_out set also the predictor update of IntrinsicM.update (:67-75): mse loss, d loss / d pred and
//           the predictor's dZ chain (the target net has no gradient).
template <int RT, int NT, int MODE>
__global__ void __launch_bounds__(NT) q_tile_kernel(QArgs a, SegTable seg, const float* __restrict__ obs,
                                                    const float* __restrict__ act, float* __restrict__ qmin_out,
                                                    float* __restrict__ p1_out, float* __restrict__ p2_out,
                                                    float* __restrict__ grad_out, float* __restrict__ gsq,
                                                    QTrainBufs tr) {
    extern __shared__ __align__(16) float smem[];
    const int ld1 = a.h1 + 4, ld2 = a.h2 + 4, ld3 = a.h3 + 4, ldl = a.atomsP;
    float* in1 = smem;                                  // [RT][K1p] = [obs | act | 0]
    float* A1 = in1 + RT * a.K1p;                       // [2][RT][ld1]
    float* A2 = A1 + 2 * RT * ld1;                      // [2][RT][ld2]
    float* A3 = A2 + 2 * RT * ld2;                      // [2][RT][ld3]
    float* LG = A3 + 2 * RT * ld3;                      // [2][RT][ldl] logits -> probs -> dlogits
    float* QV = LG + 2 * RT * ldl;                      // [2][RT]
    float* GA = QV + 2 * RT;                            // [2][RT][A4] per-net action gradients
    __shared__ float red[NT / 32];

    // locate the mode segment of this tile
    int m = 0;
    while (m + 1 < seg.n_modes && (long)blockIdx.x >= seg.tile0[m + 1]) ++m;
    const long row0 = seg.off[m] + ((long)blockIdx.x - seg.tile0[m]) * RT;
    const long row_end = seg.off[m + 1];
    const float* base = a.packed + (size_t)m * a.mode_stride;
    const int tid = threadIdx.x;

    for (int i = tid; i < RT * a.K1p; i += NT) {
        int r = i / a.K1p, c = i % a.K1p;
        long row = row0 + r;
        float v = 0.f;
        if (row < row_end) {
            if (c < a.O) v = obs[row * a.O + c];
            else if (c < a.O + a.A) v = act[row * a.A + (c - a.O)];
        }
        in1[i] = v;
    }
    __syncthreads();

    for (int j = 0; j < 2; ++j) {
        const QNetLayout& n = a.net[j];
        float* a1 = A1 + j * RT * ld1; float* a2 = A2 + j * RT * ld2; float* a3 = A3 + j * RT * ld3;
        float* lg = LG + j * RT * ldl;
        const float* b1 = base + n.b1; const float* b2 = base + n.b2; const float* b3 = base + n.b3;
        const float* b4 = base + n.b4;
        tile_linear<RT, NT>(base + n.wt1, a.h1, a.K1p >> 2, a.h1, in1, a.K1p, a.ks1, [&](int r, int n0, float4 v) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(b1 + n0));
            *reinterpret_cast<float4*>(a1 + r * ld1 + n0) =
                make_float4(elu_f(v.x + bb.x), elu_f(v.y + bb.y), elu_f(v.z + bb.z), elu_f(v.w + bb.w));
        });
        __syncthreads();
        tile_linear<RT, NT>(base + n.wt2, a.h2, a.h1 >> 2, a.h2, a1, ld1, a.ks2, [&](int r, int n0, float4 v) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(b2 + n0));
            *reinterpret_cast<float4*>(a2 + r * ld2 + n0) =
                make_float4(elu_f(v.x + bb.x), elu_f(v.y + bb.y), elu_f(v.z + bb.z), elu_f(v.w + bb.w));
        });
        __syncthreads();
        tile_linear<RT, NT>(base + n.wt3, a.h3, a.h2 >> 2, a.h3, a2, ld2, a.ks3, [&](int r, int n0, float4 v) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(b3 + n0));
            *reinterpret_cast<float4*>(a3 + r * ld3 + n0) =
                make_float4(elu_f(v.x + bb.x), elu_f(v.y + bb.y), elu_f(v.z + bb.z), elu_f(v.w + bb.w));
        });
        __syncthreads();
        tile_linear<RT, NT>(base + n.wt4, a.atomsP, a.h3 >> 2, a.atomsP, a3, ld3, a.ks4, [&](int r, int n0, float4 v) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(b4 + n0));
            *reinterpret_cast<float4*>(lg + r * ldl + n0) = make_float4(v.x + bb.x, v.y + bb.y, v.z + bb.z, v.w + bb.w);
        });
        __syncthreads();
    }

    const float* zs = base + a.z;
    const int warp = tid >> 5, lane = tid & 31;
    // shared [RT][ld] tile -> global [B][ldg] rows of this tile
    auto store_tile = [&](const float* sbuf, int lds, float* gbuf, int ldg, int ncols) {
        for (int i = tid; i < RT * ncols; i += NT) {
            const int r = i / ncols, c = i % ncols;
            const long row = row0 + r;
            if (row < row_end) gbuf[row * ldg + c] = sbuf[r * lds + c];
        }
    };
    if (MODE == 2) {
        store_tile(in1, a.K1p, tr.xin, a.K1p, a.K1p);
        for (int j = 0; j < 2; ++j) {
            store_tile(A1 + j * RT * ld1, ld1, tr.a1[j], a.h1, a.h1);
            store_tile(A2 + j * RT * ld2, ld2, tr.a2[j], a.h2, a.h2);
            store_tile(A3 + j * RT * ld3, ld3, tr.a3[j], a.h3, a.h3);
        }
        // softmax, BCE against the projected target, and d loss / d logits: one warp per (net, row)
        float lsum = 0.f;
        for (int pr = warp; pr < 2 * RT; pr += NT / 32) {
            float* lg = LG + pr * ldl;
            const int j = pr / RT, r = pr % RT;
            const long row = row0 + r;
            const bool ok = row < row_end;
            const int c0 = lane, c1 = lane + 32;
            const bool v0 = c0 < a.atoms, v1 = c1 < a.atoms;
            const float l0 = v0 ? lg[c0] : -INFINITY, l1 = v1 ? lg[c1] : -INFINITY;
            const float mx = warp_max(fmaxf(l0, l1));
            const float e0 = v0 ? expf(l0 - mx) : 0.f, e1 = v1 ? expf(l1 - mx) : 0.f;
            const float sum = warp_sum(e0 + e1);
            const float p0 = e0 / sum, p1 = e1 / sum;
            const float t0 = (ok && v0) ? tr.target[row * a.atoms + c0] : 0.f, t1 = (ok && v1) ? tr.target[row * a.atoms + c1] : 0.f;
            // F.binary_cross_entropy: logs clamped at -100; backward (p - t) / max((1 - p) p, 1e-12)
            float g0 = 0.f, g1 = 0.f;
            if (ok && v0) {
                lsum -= t0 * fmaxf(logf(p0), -100.f) + (1.f - t0) * fmaxf(log1pf(-p0), -100.f);
                g0 = tr.inv_count * (p0 - t0) / fmaxf((1.f - p0) * p0, 1e-12f);
            }
            if (ok && v1) {
                lsum -= t1 * fmaxf(logf(p1), -100.f) + (1.f - t1) * fmaxf(log1pf(-p1), -100.f);
                g1 = tr.inv_count * (p1 - t1) / fmaxf((1.f - p1) * p1, 1e-12f);
            }
            const float dot = warp_sum(p0 * g0 + p1 * g1);
            const float d0 = p0 * (g0 - dot), d1 = p1 * (g1 - dot);
            if (c0 < a.atomsP) lg[c0] = d0;
            if (c1 < a.atomsP) lg[c1] = v1 ? d1 : 0.f;
            if (ok) {
                if (c0 < a.atomsP) tr.dl[j][row * a.atomsP + c0] = d0;
                if (c1 < a.atomsP) tr.dl[j][row * a.atomsP + c1] = v1 ? d1 : 0.f;
            }
        }
        lsum = warp_sum(lsum);
        if (lane == 0) red[warp] = lsum;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int w = 0; w < NT / 32; ++w) t += red[w];
            atomicAdd(tr.loss_out, t * tr.inv_count);
        }
        __syncthreads();
    }
    if (MODE == 3) {
        const bool train = tr.loss_out != nullptr;
        if (train) {
            store_tile(in1, a.K1p, tr.xin, a.K1p, a.K1p);
            store_tile(A1, ld1, tr.a1[0], a.h1, a.h1);
            store_tile(A2, ld2, tr.a2[0], a.h2, a.h2);
            store_tile(A3, ld3, tr.a3[0], a.h3, a.h3);
        }
        if (p1_out) store_tile(LG, ldl, p1_out, a.atoms, a.atoms);                 // predictor features
        if (p2_out) store_tile(LG + RT * ldl, ldl, p2_out, a.atoms, a.atoms);      // target features
        __syncthreads();
        float lsum = 0.f;
        for (int r = warp; r < RT; r += NT / 32) {            // one warp per row
            float* lp = LG + r * ldl;
            const float* lt = LG + (RT + r) * ldl;
            const long row = row0 + r;
            const bool ok = row < row_end;
            float ss = 0.f;
            for (int c = lane; c < a.atomsP; c += 32) {
                const float d = c < a.atoms ? lp[c] - lt[c] : 0.f;
                ss = fmaf(d, d, ss);
                if (train) {
                    const float g = 2.f * tr.inv_count * d;   // d mean((pred - target)^2) / d pred
                    lp[c] = g;
                    if (ok) tr.dl[0][row * a.atomsP + c] = g;
                }
            }
            ss = warp_sum(ss);
            if (ok) {
                if (lane == 0 && qmin_out) qmin_out[row] = sqrtf(ss);
                lsum += ss;
            }
        }
        if (!train) return;
        if (lane == 0) red[warp] = lsum;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int w = 0; w < NT / 32; ++w) t += red[w];
            atomicAdd(tr.loss_out, t * tr.inv_count);
        }
        __syncthreads();
    }
    // softmax over atoms and expectation: one warp per (net, row)
    for (int pr = warp; MODE < 2 && pr < 2 * RT; pr += NT / 32) {
        float* lg = LG + pr * ldl;                      // pr = j*RT + r, rows are contiguous
        const int c0 = lane, c1 = lane + 32;
        float l0 = c0 < a.atoms ? lg[c0] : -INFINITY, l1 = c1 < a.atoms ? lg[c1] : -INFINITY;
        float mx = warp_max(fmaxf(l0, l1));
        float e0 = c0 < a.atoms ? expf(l0 - mx) : 0.f, e1 = c1 < a.atoms ? expf(l1 - mx) : 0.f;
        float sum = warp_sum(e0 + e1);
        float p0 = e0 / sum, p1 = e1 / sum;
        float q = warp_sum((c0 < a.atoms ? p0 * zs[c0] : 0.f) + (c1 < a.atoms ? p1 * zs[c1] : 0.f));
        if (c0 < a.atomsP) lg[c0] = p0;
        if (c1 < a.atomsP) lg[c1] = p1;
        if (lane == 0) QV[pr] = q;
    }
    __syncthreads();

    if (MODE == 0) {
        for (int i = tid; i < RT; i += NT) {
            long row = row0 + i;
            if (row < row_end && qmin_out) qmin_out[row] = fminf(QV[i], QV[RT + i]);
        }
        for (int i = tid; i < 2 * RT * a.atoms; i += NT) {
            int j = i / (RT * a.atoms), rem = i % (RT * a.atoms), r = rem / a.atoms, c = rem % a.atoms;
            long row = row0 + r;
            float* dst = j == 0 ? p1_out : p2_out;
            if (row < row_end && dst) dst[row * a.atoms + c] = LG[(j * RT + r) * ldl + c];
        }
        if (!grad_out) return;                          // uniform: no barrier is skipped by a subset
    }

    // d min(Q1,Q2) / d logits: only the smaller net carries gradient (ties split, as torch.min does)
    for (int i = tid; MODE < 2 && i < 2 * RT * ldl; i += NT) {
        int j = i / (RT * ldl), rem = i % (RT * ldl), r = rem / ldl, c = rem % ldl;
        float q1 = QV[r], q2 = QV[RT + r];
        float w = (q1 == q2) ? 0.5f : ((j == 0) == (q1 < q2) ? 1.f : 0.f);
        float qj = j == 0 ? q1 : q2;
        float p = LG[i];
        LG[i] = (c < a.atoms && row0 + r < row_end) ? w * p * (zs[c] - qj) : 0.f;
    }
    __syncthreads();

    for (int j = 0; j < (MODE == 3 ? 1 : 2); ++j) {
        const QNetLayout& n = a.net[j];
        float* a1 = A1 + j * RT * ld1; float* a2 = A2 + j * RT * ld2; float* a3 = A3 + j * RT * ld3;
        float* lg = LG + j * RT * ldl; float* ga = GA + j * RT * a.A4;
        tile_linear<RT, NT>(base + n.w4b, a.h3, a.atomsP >> 2, a.h3, lg, ldl, a.kb4, [&](int r, int n0, float4 v) {
            float4* p = reinterpret_cast<float4*>(a3 + r * ld3 + n0);
            const float4 s = *p;
            *p = make_float4(v.x * elu_grad_from_act(s.x), v.y * elu_grad_from_act(s.y),
                             v.z * elu_grad_from_act(s.z), v.w * elu_grad_from_act(s.w));
        });
        __syncthreads();
        tile_linear<RT, NT>(base + n.w3b, a.h2, a.h3 >> 2, a.h2, a3, ld3, a.kb3, [&](int r, int n0, float4 v) {
            float4* p = reinterpret_cast<float4*>(a2 + r * ld2 + n0);
            const float4 s = *p;
            *p = make_float4(v.x * elu_grad_from_act(s.x), v.y * elu_grad_from_act(s.y),
                             v.z * elu_grad_from_act(s.z), v.w * elu_grad_from_act(s.w));
        });
        __syncthreads();
        tile_linear<RT, NT>(base + n.w2b, a.h1, a.h2 >> 2, a.h1, a2, ld2, a.kb2, [&](int r, int n0, float4 v) {
            float4* p = reinterpret_cast<float4*>(a1 + r * ld1 + n0);
            const float4 s = *p;
            *p = make_float4(v.x * elu_grad_from_act(s.x), v.y * elu_grad_from_act(s.y),
                             v.z * elu_grad_from_act(s.z), v.w * elu_grad_from_act(s.w));
        });
        __syncthreads();
        if (MODE >= 2) {
            store_tile(a3, ld3, tr.z3[j], a.h3, a.h3);
            store_tile(a2, ld2, tr.z2[j], a.h2, a.h2);
            store_tile(a1, ld1, tr.z1[j], a.h1, a.h1);
            continue;
        }
        tile_linear<RT, NT>(base + n.w1a, a.A4, a.h1 >> 2, a.A4, a1, ld1, a.kb1, [&](int r, int n0, float4 v) {
            *reinterpret_cast<float4*>(ga + r * a.A4 + n0) = v;
        });
        __syncthreads();
    }
    if (MODE >= 2) return;

    float sq = 0.f;
    const float scale = MODE == 1 ? -seg.inv_cnt[m] : 1.f;
    for (int i = tid; i < RT * a.A; i += NT) {
        int r = i / a.A, c = i % a.A;
        long row = row0 + r;
        if (row < row_end) {
            float g = scale * (GA[r * a.A4 + c] + GA[(RT + r) * a.A4 + c]);
            grad_out[row * a.A + c] = g;
            sq = fmaf(g, g, sq);
        }
    }
    if (MODE == 1) {
        sq = warp_sum(sq);
        if (lane == 0) red[warp] = sq;
        __syncthreads();
        if (tid == 0) {
            float s = 0.f;
            for (int w = 0; w < NT / 32; ++w) s += red[w];
            atomicAdd(gsq + m, s);
        }
    }
}

// Adam on the action rows of each segment, with the segment's clip coefficient
// (clip_grad_norm_ then torch.optim.Adam.step then clamp_, ac_base.py:86-91 / ddiffpg.py:369).
__global__ void q_adam_kernel(SegTable seg, int A, float* __restrict__ act, const float* __restrict__ g,
                              float* __restrict__ m1, float* __restrict__ m2, const float* __restrict__ gsq,
                              float* __restrict__ gnorm_out, int iter, int iters, float step_size, float bc2_sqrt,
                              float b1, float b2, float eps, float max_norm, float lim, long n_elems) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    int mode = 0;
    if (i < n_elems) {
        long row = i / A;
        while (mode + 1 < seg.n_modes && row >= seg.off[mode + 1]) ++mode;
        const float norm = sqrtf(gsq[mode]);
        float coef = max_norm / (norm + 1e-6f);
        coef = fminf(coef, 1.0f);
        if (gnorm_out && i == seg.off[mode] * A) gnorm_out[mode * iters + iter] = norm;
        const float gi = g[i] * coef;
        const float ea = m1[i] * b1 + (1.f - b1) * gi;          // exp_avg.lerp_(grad, 1-beta1)
        const float ev = m2[i] * b2 + (1.f - b2) * gi * gi;     // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
        m1[i] = ea; m2[i] = ev;
        const float denom = sqrtf(ev) / bc2_sqrt + eps;
        float v = act[i] - step_size * (ea / denom);
        v = fminf(fmaxf(v, -lim), lim);
        act[i] = v;
    }
}

// sum |a| per mode segment (torch.abs(action).mean(), ddiffpg.py:373)
__global__ void q_abs_sum_kernel(SegTable seg, int A, const float* __restrict__ act, long n_elems,
                                 float* __restrict__ abs_sum) {
    __shared__ float bins[kMaxModes];
    if (threadIdx.x < kMaxModes) bins[threadIdx.x] = 0.f;
    __syncthreads();
    const long stride = (long)gridDim.x * blockDim.x;
    const long n_round = (n_elems + 31) / 32 * 32;           // whole warps iterate together
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        int mode = -1;
        float v = 0.f;
        if (i < n_elems) {
            long row = i / A;
            mode = 0;
            while (mode + 1 < seg.n_modes && row >= seg.off[mode + 1]) ++mode;
            v = fabsf(act[i]);
        }
        const int m0 = __shfl_sync(0xffffffffu, mode, 0);
        if (__all_sync(0xffffffffu, mode == m0 || mode < 0) && m0 >= 0) {
            v = warp_sum(v);
            if ((threadIdx.x & 31) == 0) atomicAdd(&bins[m0], v);
        } else if (mode >= 0) {
            atomicAdd(&bins[mode], v);
        }
    }
    __syncthreads();
    if (threadIdx.x < seg.n_modes && bins[threadIdx.x] != 0.f) atomicAdd(abs_sum + threadIdx.x, bins[threadIdx.x]);
}

__global__ void q_clamp_kernel(float* __restrict__ act, long n, float lim) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) act[i] = fminf(fmaxf(act[i], -lim), lim);
}

__global__ void q_finish_kernel(SegTable seg, int A, const float* __restrict__ abs_sum, float* __restrict__ mean_abs) {
    int m = threadIdx.x;
    if (m < seg.n_modes) {
        long n = (seg.off[m + 1] - seg.off[m]) * A;
        mean_abs[m] = n > 0 ? abs_sum[m] / (float)n : 0.f;
    }
}

static size_t q_tile_smem(const QLayout& L, int RT) {
    size_t f = (size_t)RT * L.K1p + 2 * (size_t)RT * ((L.h1 + 4) + (L.h2 + 4) + (L.h3 + 4) + L.atomsP) + 2 * RT +
               2 * (size_t)RT * L.A4;
    return f * sizeof(float);
}

static long fill_segments(const QLayout& L, const int64_t* seg_off, const int64_t* seg_cnt, int RT, SegTable& st) {
    st.n_modes = L.n_modes;
    long tiles = 0;
    for (int m = 0; m < L.n_modes; ++m) {
        st.off[m] = seg_off[m];
        st.tile0[m] = tiles;
        long len = seg_off[m + 1] - seg_off[m];
        tiles += (len + RT - 1) / RT;
        long cnt = seg_cnt ? seg_cnt[m] : len;
        st.inv_cnt[m] = cnt > 0 ? 1.0f / (float)cnt : 0.f;
    }
    st.off[L.n_modes] = seg_off[L.n_modes];
    st.tile0[L.n_modes] = tiles;
    return tiles;
}

template <int RT, int MODE>
static int launch_q_tile(const QLayout& L, const float* pk, const SegTable& seg, long tiles, const float* obs,
                         const float* act, float* qmin, float* p1, float* p2, float* grad, float* gsq,
                         cudaStream_t st, const QTrainBufs* tr = nullptr) {
    constexpr int NT = 256;
    size_t smem = q_tile_smem(L, RT);
    if (smem > 227 * 1024) DDP_FAIL(DDP_ERR_SHAPE, "critic tile needs %zu B of shared memory", smem);
    auto kern = q_tile_kernel<RT, NT, MODE>;
    DDP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QArgs a = make_qargs(L, pk, NT);
    kern<<<(unsigned)tiles, NT, smem, st>>>(a, seg, obs, act, qmin, p1, p2, grad, gsq, tr ? *tr : QTrainBufs{});
    DDP_LAUNCH_CHECK("q_tile_kernel");
    return DDP_OK;
}

int q_forward_fma(const QLayout& L, const float* pk, const int64_t* seg_off, const float* obs, const float* act,
                  float* qmin, float* p1, float* p2, float* dq_da, long B, cudaStream_t st) {
    SegTable seg;
    if (B <= 148 * 8) {
        long tiles = fill_segments(L, seg_off, nullptr, 4, seg);
        return launch_q_tile<4, 0>(L, pk, seg, tiles, obs, act, qmin, p1, p2, dq_da, nullptr, st);
    }
    long tiles = fill_segments(L, seg_off, nullptr, 16, seg);
    return launch_q_tile<16, 0>(L, pk, seg, tiles, obs, act, qmin, p1, p2, dq_da, nullptr, st);
}

// workspace: g [B*A], exp_avg [B*A], exp_avg_sq [B*A], gsq [iters][kMaxModes], abs_sum [kMaxModes]
size_t q_ascent_workspace(const QLayout& L, long B, int iters) {
    size_t n = align_up((size_t)B * L.A, 64);
    return (3 * n + (size_t)iters * kMaxModes + kMaxModes) * sizeof(float);
}

int q_ascent_fma(const QLayout& L, const float* pk, const int64_t* seg_off, const int64_t* seg_cnt, const float* obs,
                 float* action, int iters, float lr, float b1, float b2, float eps, float max_norm, float lim,
                 float* mean_abs, float* gnorm_out, long B, void* ws, size_t ws_bytes, cudaStream_t st,
                 ddp_gsq_reduce_fn reduce, void* reduce_user) {
    const size_t n = align_up((size_t)B * L.A, 64);
    float* g = (float*)ws;
    float* m1 = g + n;
    float* m2 = m1 + n;
    float* gsq = m2 + n;
    float* abs_sum = gsq + (size_t)iters * kMaxModes;
    const long n_elems = B * L.A;
    // fresh Adam state (a new torch.optim.Adam per call, ddiffpg.py:362) and zeroed reductions
    DDP_CUDA_CHECK(cudaMemsetAsync(m1, 0, (2 * n + (size_t)iters * kMaxModes + kMaxModes) * sizeof(float), st));
    const unsigned eb = (unsigned)((n_elems + 255) / 256);
    q_clamp_kernel<<<eb, 256, 0, st>>>(action, n_elems, lim);                      // ddiffpg.py:361
    SegTable seg;
    const bool small = B <= 148 * 8;
    const long tiles = fill_segments(L, seg_off, seg_cnt, small ? 4 : 16, seg);
    for (int it = 0; it < iters; ++it) {
        int rc = small ? launch_q_tile<4, 1>(L, pk, seg, tiles, obs, action, nullptr, nullptr, nullptr, g,
                                            gsq + (size_t)it * kMaxModes, st)
                       : launch_q_tile<16, 1>(L, pk, seg, tiles, obs, action, nullptr, nullptr, nullptr, g,
                                             gsq + (size_t)it * kMaxModes, st);
        if (rc != DDP_OK) return rc;
        // row-sharded batch: the clip norm is the norm over ALL ranks' rows of the mode (semantics (ii))
        if (reduce && reduce(gsq + (size_t)it * kMaxModes, L.n_modes, (void*)st, reduce_user) != 0)
            DDP_FAIL(DDP_ERR_ARG, "ddp_q_action_ascent_sharded: the reduce callback failed in iteration %d", it);
        const int step = it + 1;
        const double bc1 = 1.0 - pow((double)b1, step), bc2 = 1.0 - pow((double)b2, step);
        q_adam_kernel<<<eb, 256, 0, st>>>(seg, L.A, action, g, m1, m2, gsq + (size_t)it * kMaxModes, gnorm_out, it,
                                          iters, (float)(lr / bc1), (float)sqrt(bc2), b1, b2, eps, max_norm, lim,
                                          n_elems);
    }
    q_abs_sum_kernel<<<(unsigned)((n_elems + 2047) / 2048 < 592 ? (n_elems + 2047) / 2048 : 592), 256, 0, st>>>(seg, L.A, action, n_elems, abs_sum);
    q_finish_kernel<<<1, kMaxModes, 0, st>>>(seg, L.A, abs_sum, mean_abs);
    DDP_LAUNCH_CHECK("q ascent kernels");
    return DDP_OK;
}

// ------------------------------------------------------------------------------------------------
// N1: critic update (AgentDDiffPG.update_critic, ddiffpg/algo/ddiffpg.py:322-351), fp32 path.
void launch_dw(const float* dz, int ldz, int N, const float* x, int ldx, int K, float* C, int ldc, float* dbias, long R,
               cudaStream_t st);      // fp32 dW = dZ^T . X with bias column sums (actor_train_fma.cu)

// Categorical projection of r + (1 - done) * gamma * z onto the support (ddiffpg/utils/distl_util.py:4-20) for
// both target heads, then their element-wise minimum (ddiffpg.py:346).  One warp per row.  The scatter-add
// (index_add_ in the reference) accumulates in 2^-30 fixed point with integer shared-memory atomics: fp32 atomicAdd on
// shared memory is a compare-and-swap loop that spins under the address conflicts this scatter is full of (a `done` row
// sends all atoms to the same two bins), the integer add is native -- and the sums become order-independent.  A bin
// holds at most the total mass 1 (+ rounding), the quantisation error per bin is below atoms * 2^-31 = 2.4e-8.
__global__ void c51_projection_min_kernel(const float* __restrict__ p1, const float* __restrict__ p2,
                                          const float* __restrict__ reward, const float* __restrict__ done,
                                          float gamma, float v_min, float v_max, int atoms, const float* __restrict__ z,
                                          long B, float* __restrict__ target) {
    __shared__ int bins[8][2][64];
    constexpr float kScale = 1073741824.f, kInv = 1.f / 1073741824.f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + warp;
    bins[warp][0][lane] = 0; bins[warp][0][lane + 32] = 0;
    bins[warp][1][lane] = 0; bins[warp][1][lane + 32] = 0;
    __syncwarp();
    if (row < B) {
        const float delta_z = (float)(((double)v_max - (double)v_min) / (double)(atoms - 1));
        const float r = reward[row], nd = __fmul_rn(__fsub_rn(1.f, done[row]), gamma);
        for (int j = lane; j < atoms; j += 32) {
            float tz = __fadd_rn(r, __fmul_rn(nd, z[j]));
            tz = fminf(fmaxf(tz, v_min), v_max);
            const float b = __fdiv_rn(__fsub_rn(tz, v_min), delta_z);
            int lo = (int)floorf(b), up = (int)ceilf(b);
            if (up > 0 && lo == up) lo -= 1;
            if (lo < atoms - 1 && lo == up) up += 1;
            const float wl = __fsub_rn((float)up, b), wu = __fsub_rn(b, (float)lo);
            const float a1 = p1[row * atoms + j], a2 = p2[row * atoms + j];
            atomicAdd(&bins[warp][0][lo], __float2int_rn(a1 * wl * kScale)); atomicAdd(&bins[warp][0][up], __float2int_rn(a1 * wu * kScale));
            atomicAdd(&bins[warp][1][lo], __float2int_rn(a2 * wl * kScale)); atomicAdd(&bins[warp][1][up], __float2int_rn(a2 * wu * kScale));
        }
        __syncwarp();
        for (int j = lane; j < atoms; j += 32)
            target[row * atoms + j] = (float)min(bins[warp][0][j], bins[warp][1][j]) * kInv;
    }
}

void launch_c51_projection_min(const float* p1, const float* p2, const float* reward, const float* done, float gamma,
                               float v_min, float v_max, int atoms, const float* z, long B, float* target,
                               cudaStream_t st) {
    c51_projection_min_kernel<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(p1, p2, reward, done, gamma, v_min, v_max, atoms, z,
                                                                        B, target);
}

struct QTrainWs {
    float *p1t, *p2t, *target;
    QTrainBufs b;
    size_t total;
};

static QTrainWs carve_q_train(const QLayout& L, long B, float* base) {
    QTrainWs w{};
    size_t o = 0;
    auto take = [&](size_t n) { float* r = base ? base + o : nullptr; o += (n + 63) / 64 * 64; return r; };
    w.p1t = take((size_t)B * L.atoms); w.p2t = take((size_t)B * L.atoms); w.target = take((size_t)B * L.atoms);
    w.b.xin = take((size_t)B * L.K1p);
    for (int j = 0; j < 2; ++j) {
        w.b.a1[j] = take((size_t)B * L.h1); w.b.a2[j] = take((size_t)B * L.h2); w.b.a3[j] = take((size_t)B * L.h3);
        w.b.z1[j] = take((size_t)B * L.h1); w.b.z2[j] = take((size_t)B * L.h2); w.b.z3[j] = take((size_t)B * L.h3);
        w.b.dl[j] = take((size_t)B * L.atomsP);
    }
    w.total = o * sizeof(float);
    return w;
}

size_t q_critic_train_workspace(const QLayout& L, long B) { return carve_q_train(L, B, nullptr).total; }

size_t q_grad_count(const QLayout& L) {
    const size_t in1 = L.O + L.A;
    return 2 * ((size_t)L.h1 * in1 + L.h1 + (size_t)L.h2 * L.h1 + L.h2 + (size_t)L.h3 * L.h2 + L.h3 + (size_t)L.atoms * L.h3 + L.atoms);
}

// pk / pk_target: packed critic and target critic (single mode).  grads: flat, state_dict order.
int q_critic_train_fma(const QLayout& L, const float* pk, const float* pk_target, const float* obs, const float* act,
                       const float* next_obs, const float* next_act, const float* reward, const float* done, float gamma,
                       float* loss_out, float* grads, long B, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (ws_bytes < carve_q_train(L, B, nullptr).total) DDP_FAIL(DDP_ERR_ARG, "critic training: workspace too small");
    QTrainWs w = carve_q_train(L, B, (float*)ws);
    const int64_t seg_off[2] = {0, B};
    // target distribution: critic_target.get_q1_q2(next_obs, next_actions) -> projection x2 -> min
    int rc = q_forward_fma(L, pk_target, seg_off, next_obs, next_act, nullptr, w.p1t, w.p2t, nullptr, B, st);
    if (rc != DDP_OK) return rc;
    c51_projection_min_kernel<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(w.p1t, w.p2t, reward, done, gamma, L.v_min, L.v_max,
                                                                        L.atoms, pk_target + L.z, B, w.target);
    DDP_CUDA_CHECK(cudaMemsetAsync(grads, 0, q_grad_count(L) * sizeof(float), st));
    w.b.target = w.target;
    w.b.inv_count = 1.0f / ((float)B * (float)L.atoms);
    w.b.loss_out = loss_out;
    SegTable seg;
    const bool small = B <= 148 * 8;
    const long tiles = fill_segments(L, seg_off, nullptr, small ? 4 : 16, seg);
    rc = small ? launch_q_tile<4, 2>(L, pk, seg, tiles, obs, act, nullptr, nullptr, nullptr, nullptr, nullptr, st, &w.b)
               : launch_q_tile<16, 2>(L, pk, seg, tiles, obs, act, nullptr, nullptr, nullptr, nullptr, nullptr, st, &w.b);
    if (rc != DDP_OK) return rc;
    // weight gradients, flat in state_dict order: per net W1,b1,W2,b2,W3,b3,W4,b4
    const int in1 = L.O + L.A;
    float* g = grads;
    for (int j = 0; j < 2; ++j) {
        launch_dw(w.b.z1[j], L.h1, L.h1, w.b.xin, L.K1p, in1, g, in1, g + (size_t)L.h1 * in1, B, st);
        g += (size_t)L.h1 * in1 + L.h1;
        launch_dw(w.b.z2[j], L.h2, L.h2, w.b.a1[j], L.h1, L.h1, g, L.h1, g + (size_t)L.h2 * L.h1, B, st);
        g += (size_t)L.h2 * L.h1 + L.h2;
        launch_dw(w.b.z3[j], L.h3, L.h3, w.b.a2[j], L.h2, L.h2, g, L.h2, g + (size_t)L.h3 * L.h2, B, st);
        g += (size_t)L.h3 * L.h2 + L.h3;
        launch_dw(w.b.dl[j], L.atomsP, L.atoms, w.b.a3[j], L.h3, L.h3, g, L.h3, g + (size_t)L.atoms * L.h3, B, st);
        g += (size_t)L.atoms * L.h3 + L.atoms;
    }
    DDP_LAUNCH_CHECK("critic training kernels");
    return DDP_OK;
}

// ------------------------------------------------------------------------------------------------
// RND / NovelD.  L describes RNDModel as a two-net pack: O = input width, A = 0,
// atoms = feature width, net 0 = predictor, net 1 = target.
size_t rnd_grad_count(const QLayout& L) { return q_grad_count(L) / 2; }
size_t rnd_train_workspace(const QLayout& L, long B) { return carve_q_train(L, B, nullptr).total; }

int rnd_novelty_fma(const QLayout& L, const float* pk, const float* x, float* novelty, float* pred, float* target,
                    long B, cudaStream_t st) {
    const int64_t seg_off[2] = {0, B};
    SegTable seg;
    const bool small = B <= 148 * 8;
    const long tiles = fill_segments(L, seg_off, nullptr, small ? 4 : 16, seg);
    return small ? launch_q_tile<4, 3>(L, pk, seg, tiles, x, nullptr, novelty, pred, target, nullptr, nullptr, st)
                 : launch_q_tile<16, 3>(L, pk, seg, tiles, x, nullptr, novelty, pred, target, nullptr, nullptr, st);
}

// loss_out[0] += mse(pred, target); grads = d loss / d predictor params (W1,b1,...,W4,b4), overwritten;
// novelty (optional) as in rnd_novelty_fma
int rnd_train_fma(const QLayout& L, const float* pk, const float* x, float* loss_out, float* grads, float* novelty,
                  long B, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (ws_bytes < carve_q_train(L, B, nullptr).total) DDP_FAIL(DDP_ERR_ARG, "RND update: workspace too small");
    QTrainWs w = carve_q_train(L, B, (float*)ws);
    const int64_t seg_off[2] = {0, B};
    DDP_CUDA_CHECK(cudaMemsetAsync(grads, 0, rnd_grad_count(L) * sizeof(float), st));
    w.b.target = nullptr;
    w.b.inv_count = 1.0f / ((float)B * (float)L.atoms);
    w.b.loss_out = loss_out;
    SegTable seg;
    const bool small = B <= 148 * 8;
    const long tiles = fill_segments(L, seg_off, nullptr, small ? 4 : 16, seg);
    int rc = small ? launch_q_tile<4, 3>(L, pk, seg, tiles, x, nullptr, novelty, nullptr, nullptr, nullptr, nullptr, st, &w.b)
                   : launch_q_tile<16, 3>(L, pk, seg, tiles, x, nullptr, novelty, nullptr, nullptr, nullptr, nullptr, st, &w.b);
    if (rc != DDP_OK) return rc;
    const int in1 = L.O + L.A;
    float* g = grads;
    launch_dw(w.b.z1[0], L.h1, L.h1, w.b.xin, L.K1p, in1, g, in1, g + (size_t)L.h1 * in1, B, st);
    g += (size_t)L.h1 * in1 + L.h1;
    launch_dw(w.b.z2[0], L.h2, L.h2, w.b.a1[0], L.h1, L.h1, g, L.h1, g + (size_t)L.h2 * L.h1, B, st);
    g += (size_t)L.h2 * L.h1 + L.h2;
    launch_dw(w.b.z3[0], L.h3, L.h3, w.b.a2[0], L.h2, L.h2, g, L.h2, g + (size_t)L.h3 * L.h2, B, st);
    g += (size_t)L.h3 * L.h2 + L.h3;
    launch_dw(w.b.dl[0], L.atomsP, L.atoms, w.b.a3[0], L.h3, L.h3, g, L.h3, g + (size_t)L.atoms * L.h3, B, st);
    DDP_LAUNCH_CHECK("RND update kernels");
    return DDP_OK;
}

}  // namespace ddp
