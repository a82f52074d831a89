// Layout of the packed actor weight buffer (see ddp_actor_pack in include/ddiffpg_b200.h).
#pragma once
#include "common.cuh"

namespace ddp {

constexpr int kMaxT = 128;       // schedule constants travel as a by-value kernel argument
constexpr int kCstStride = 8;    // floats per timestep in the constants table

// Per-timestep DDPM constants (diffusers DDPMScheduler.step / add_noise; reference call sites
// diffusion_mlp.py:243-247, 309-310).  Row t: c_eps=sqrt(1-abar_t), sqrt_abar_t, c_x0, c_xt, sigma_t,
// then sqrt_abar_t, sqrt(1-abar_t) again for add_noise, pad.
struct ScheduleTable { float v[kMaxT][kCstStride]; };
enum { CST_CEPS = 0, CST_SQRT_AB = 1, CST_CX0 = 2, CST_CXT = 3, CST_SIGMA = 4, CST_ADD_A = 5, CST_ADD_B = 6 };

struct ActorLayout {
    int S, A, T, D, h1, h2, h3;
    int K0p;      // pad4(S + A): layer-0 contraction over [state | x]
    int A4;       // pad4(A)
    // fp32 section (offsets in floats from the start of the buffer)
    size_t wt0;   // [K0p][h1]   net.mlp.0.weight[:, D:D+S+A]^T, zero padded rows
    size_t wt1;   // [h1][h2]    net.mlp.2.weight^T
    size_t wt2;   // [h2][h3]    net.mlp.4.weight^T
    size_t wt3;   // [h3][A4]    net.mlp.6.weight^T, zero padded columns
    size_t b1, b2, b3;            // [h2], [h3], [A4]
    size_t w3b;   // [A4][h3]    net.mlp.6.weight with zero padded rows (backward operand dX = dY.W)
    size_t tb0;   // [T][h1]     time table: W0[:, :D] . time_mlp(posemb(t)) + b0
    size_t cst;   // [T][8]      schedule constants
    size_t pe;    // [T][D]      sinusoidal embedding            (kept for the H3 backward)
    size_t zmid;  // [T][4D]     time_mlp.1 pre-activation
    size_t hmid;  // [T][4D]     mish(zmid)
    size_t temb;  // [T][D]      time_mlp output
    size_t fp32_floats;
    // bf16 tensor-core section (offsets in BYTES from the start of the buffer), precision == BF16
    size_t tc_w0, tc_w1, tc_w2, tc_w3;        // bf16 operands
    size_t tc_w0h, tc_w1h, tc_w2h, tc_w3h;    // fp16 copies for the ill-conditioned first denoising step
    // training tensor path (H3): plain K-major bf16 operands of the row GEMMs and their transposes
    size_t tr_w0, tr_w3, tr_w3t, tr_w2t, tr_w1t, tr_colmap;
    size_t total_bytes;
};

inline ActorLayout make_actor_layout(const ddp_actor_shape& s, int precision) {
    ActorLayout L{};
    L.S = s.S; L.A = s.A; L.T = s.T; L.D = s.D; L.h1 = s.h1; L.h2 = s.h2; L.h3 = s.h3;
    L.K0p = pad4(s.S + s.A);
    L.A4 = pad4(s.A);
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 63) / 64 * 64; return r; };   // 256 B aligned
    L.wt0 = take((size_t)L.K0p * s.h1);
    L.wt1 = take((size_t)s.h1 * s.h2);
    L.wt2 = take((size_t)s.h2 * s.h3);
    L.wt3 = take((size_t)s.h3 * L.A4);
    L.b1 = take(s.h2); L.b2 = take(s.h3); L.b3 = take(L.A4);
    L.w3b = take((size_t)L.A4 * s.h3);
    L.tb0 = take((size_t)s.T * s.h1);
    L.cst = take((size_t)s.T * kCstStride);
    L.pe = take((size_t)s.T * s.D);
    L.zmid = take((size_t)s.T * 4 * s.D);
    L.hmid = take((size_t)s.T * 4 * s.D);
    L.temb = take((size_t)s.T * s.D);
    L.fp32_floats = o;
    size_t bytes = align_up(o * sizeof(float), 1024);
    L.tc_w0 = L.tc_w1 = L.tc_w2 = L.tc_w3 = 0;
    L.tc_w0h = L.tc_w1h = L.tc_w2h = L.tc_w3h = 0;
    L.tr_w0 = L.tr_w3 = L.tr_w3t = L.tr_w2t = L.tr_w1t = L.tr_colmap = 0;
    if (precision == DDP_BF16) {
        // filled in by the tensor-core packer (actor_sample_tc.cu); sizes in bf16 elements
        auto takeb = [&](size_t nbytes) { size_t r = bytes; bytes += align_up(nbytes, 1024); return r; };
        L.tc_w0 = takeb((size_t)s.h1 * 64 * 2);            // [h1][64]  K = [state|x|pad] padded to 64
        L.tc_w1 = takeb((size_t)s.h2 * s.h1 * 2);          // [h2][h1]
        L.tc_w2 = takeb((size_t)s.h3 * s.h2 * 2);          // [h3][h2]
        L.tc_w3 = takeb((size_t)16 * s.h3 * 2);            // [16][h3]  A padded to 16 rows
        L.tc_w0h = takeb((size_t)s.h1 * 64 * 2);
        L.tc_w1h = takeb((size_t)s.h2 * s.h1 * 2);
        L.tc_w2h = takeb((size_t)s.h3 * s.h2 * 2);
        L.tc_w3h = takeb((size_t)16 * s.h3 * 2);
        L.tr_w0 = takeb((size_t)s.h1 * 64 * 2);            // [h1][64]   K order [x | state | 0]
        L.tr_w3 = takeb((size_t)16 * s.h3 * 2);            // [16][h3]   rows >= A are zero
        L.tr_w3t = takeb((size_t)s.h3 * 64 * 2);           // [h3][64]   W3^T, columns >= A are zero
        L.tr_w2t = takeb((size_t)s.h2 * s.h3 * 2);         // [h2][h3]   W2^T
        L.tr_w1t = takeb((size_t)s.h1 * s.h2 * 2);         // [h1][h2]   W1^T
        L.tr_colmap = takeb(64 * 4);                       // xin column -> net.mlp.0.weight column
    }
    L.total_bytes = bytes;
    return L;
}

inline int check_actor_shape(const ddp_actor_shape* s) {
    if (!s) DDP_FAIL(DDP_ERR_ARG, "actor shape is NULL");
    if (s->S <= 0 || s->A <= 0 || s->A > 32 || s->T <= 0 || s->T > kMaxT)
        DDP_FAIL(DDP_ERR_SHAPE, "actor shape: need S>0, 0<A<=32, 0<T<=%d (got S=%d A=%d T=%d)", kMaxT, s->S, s->A, s->T);
    if (s->D <= 0 || s->D % 8 || s->h1 % 16 || s->h2 % 16 || s->h3 % 16 || s->h1 <= 0 || s->h2 <= 0 || s->h3 <= 0)
        DDP_FAIL(DDP_ERR_SHAPE, "actor shape: D must be a multiple of 8 and h1,h2,h3 multiples of 16 (got D=%d h=%d,%d,%d)",
                 s->D, s->h1, s->h2, s->h3);
    if (s->h1 > 2048 || s->h2 > 2048 || s->h3 > 2048)
        DDP_FAIL(DDP_ERR_SHAPE, "actor shape: trunk widths above 2048 are not supported");
    return DDP_OK;
}

// Exploration / target-policy noise applied to the sampled action (reference: ddiffpg/utils/noise.py:19-41 as
// called by AgentDDiffPG.get_actions / get_tgt_policy_actions, ddiffpg/algo/ddiffpg.py:88-109):
//   out = clamp(a + clamp(std_r * z, -bound, +bound), -1, 1),  std_r = linspace(std_min, std_max, B)[r]
// z [B, A] is pre-drawn standard normal noise (NULL switches the epilogue off); bound <= 0 means no noise clamp.
struct ExplNoise {
    const float* z;
    float std_min, std_max, bound;
};
__device__ __forceinline__ float apply_expl_noise(const ExplNoise& e, float a, long row, long B, int A, int c) {
    if (!e.z) return a;
    // torch.linspace fills from both ends with step = (end - start) / (steps - 1)
    const float step = B > 1 ? (e.std_max - e.std_min) / (float)(B - 1) : 0.f;
    const float sd = row < B / 2 ? e.std_min + step * (float)row : e.std_max - step * (float)(B - 1 - row);
    float nz = sd * e.z[row * A + c];
    if (e.bound > 0.f) nz = fminf(fmaxf(nz, -e.bound), e.bound);
    return fminf(fmaxf(a + nz, -1.f), 1.f);
}

// offsets (in floats) of each parameter inside the flat gradient vector, state_dict order
struct ActorGradOffsets { size_t off[13]; };
inline ActorGradOffsets actor_grad_offsets(const ddp_actor_shape& s) {
    ActorGradOffsets g{};
    size_t n[12] = {(size_t)4 * s.D * s.D, (size_t)4 * s.D, (size_t)s.D * 4 * s.D, (size_t)s.D,
                    (size_t)s.h1 * (s.D + s.S + s.A), (size_t)s.h1, (size_t)s.h2 * s.h1, (size_t)s.h2,
                    (size_t)s.h3 * s.h2, (size_t)s.h3, (size_t)s.A * s.h3, (size_t)s.A};
    g.off[0] = 0;
    for (int i = 0; i < 12; ++i) g.off[i + 1] = g.off[i] + n[i];
    return g;
}

}  // namespace ddp
