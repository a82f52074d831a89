// Layout of the packed critic buffer (see ddp_q_pack in include/ddiffpg_b200.h).
#pragma once
#include "common.cuh"

namespace ddp {

constexpr int kMaxModes = 32;

// One MLPNet (reference mlp.py:23-35): forward matrices transposed (K-major) and the backward
// matrices dX = dY.W in the layout tile_linear streams ([contraction][output]).
struct QNetLayout {
    size_t wt1, wt2, wt3, wt4;   // [K1p][h1] [h1][h2] [h2][h3] [h3][atomsP]
    size_t b1, b2, b3, b4;
    size_t w4b, w3b, w2b, w1a;   // [atomsP][h3] [h3][h2] [h2][h1] [h1][A4] (action columns only)
};

struct QLayout {
    int O, A, atoms, n_modes, h1, h2, h3;
    float v_min, v_max;
    int K1p, A4, atomsP;
    int K1c, AtP;                // tensor path: input width and head width padded to whole 64-column chunks
    QNetLayout net[2];           // offsets in floats inside one mode's block
    size_t z;                    // [atomsP] support atoms (shared by both nets)
    size_t mode_stride;          // floats per mode
    // bf16 tensor-core section (precision == DDP_BF16): byte offsets; each array holds n_modes matrices back to
    // back so that the grouped row GEMM can stride over the modes.  Index [net][layer].
    size_t tc_fwd[2][4], tc_bwd[2][4];
    size_t tc_fwd_elems[4], tc_bwd_elems[4];   // elements per mode of each matrix
    size_t total_bytes;
};

inline QLayout make_q_layout(const ddp_q_shape& s, int precision) {
    QLayout L{};
    L.O = s.O; L.A = s.A; L.atoms = s.atoms; L.n_modes = s.n_modes; L.h1 = s.hid1; L.h2 = s.hid2; L.h3 = s.hid3;
    L.v_min = s.v_min; L.v_max = s.v_max;
    L.K1p = pad4(s.O + s.A); L.A4 = pad4(s.A); L.atomsP = pad4(s.atoms);
    L.K1c = (s.O + s.A + 63) / 64 * 64; L.AtP = (s.atoms + 63) / 64 * 64;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 63) / 64 * 64; return r; };
    for (int j = 0; j < 2; ++j) {
        QNetLayout& n = L.net[j];
        n.wt1 = take((size_t)L.K1p * L.h1); n.wt2 = take((size_t)L.h1 * L.h2);
        n.wt3 = take((size_t)L.h2 * L.h3);  n.wt4 = take((size_t)L.h3 * L.atomsP);
        n.b1 = take(L.h1); n.b2 = take(L.h2); n.b3 = take(L.h3); n.b4 = take(L.atomsP);
        n.w4b = take((size_t)L.atomsP * L.h3); n.w3b = take((size_t)L.h3 * L.h2);
        n.w2b = take((size_t)L.h2 * L.h1);     n.w1a = take((size_t)L.h1 * L.A4);
    }
    L.z = take(L.atomsP);
    L.mode_stride = o;
    size_t bytes = align_up(o * sizeof(float) * (size_t)s.n_modes, 1024);
    if (precision == DDP_BF16) {
        // forward: W1 [h1][K1c] (K = [obs|act|0]), W2 [h2][h1], W3 [h3][h2], W4 [AtP][h3] (rows >= atoms zero)
        // backward (dX = dY.W): W4^T [h3][AtP], W3^T [h2][h3], W2^T [h1][h2], W1[:, O:O+A]^T [16][h1]
        // (K1c = AtP = 64 for the critics; the RND nets have 69 inputs and 128 features)
        const size_t fe[4] = {(size_t)L.h1 * L.K1c, (size_t)L.h2 * L.h1, (size_t)L.h3 * L.h2, (size_t)L.AtP * L.h3};
        const size_t be[4] = {(size_t)L.h3 * L.AtP, (size_t)L.h2 * L.h3, (size_t)L.h1 * L.h2, (size_t)16 * L.h1};
        for (int i = 0; i < 4; ++i) { L.tc_fwd_elems[i] = fe[i]; L.tc_bwd_elems[i] = be[i]; }
        for (int j = 0; j < 2; ++j) {
            for (int i = 0; i < 4; ++i) { L.tc_fwd[j][i] = bytes; bytes += align_up(fe[i] * 2 * s.n_modes, 1024); }
            for (int i = 0; i < 4; ++i) { L.tc_bwd[j][i] = bytes; bytes += align_up(be[i] * 2 * s.n_modes, 1024); }
        }
    }
    L.total_bytes = bytes;
    return L;
}

inline int check_q_shape(const ddp_q_shape* s) {
    if (!s) DDP_FAIL(DDP_ERR_ARG, "critic shape is NULL");
    if (s->O <= 0 || s->A <= 0 || s->A > 32 || s->atoms < 2 || s->atoms > 64)
        DDP_FAIL(DDP_ERR_SHAPE, "critic shape: need O>0, 0<A<=32, 2<=atoms<=64 (got O=%d A=%d atoms=%d)", s->O, s->A, s->atoms);
    if (s->n_modes < 1 || s->n_modes > kMaxModes)
        DDP_FAIL(DDP_ERR_SHAPE, "critic shape: n_modes must be in [1,%d] (got %d)", kMaxModes, s->n_modes);
    if (s->hid1 <= 0 || s->hid2 <= 0 || s->hid3 <= 0 || s->hid1 % 4 || s->hid2 % 4 || s->hid3 % 4 ||
        s->hid1 > 1024 || s->hid2 > 1024 || s->hid3 > 1024)
        DDP_FAIL(DDP_ERR_SHAPE, "critic shape: hidden widths must be multiples of 4 in (0,1024]");
    if (!(s->v_max > s->v_min)) DDP_FAIL(DDP_ERR_SHAPE, "critic shape: v_max must exceed v_min");
    return DDP_OK;
}

}  // namespace ddp
