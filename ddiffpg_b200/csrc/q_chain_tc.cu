// H2, bf16 tensor-core path, fused: forward of both critic nets, softmax / expectation / arg-min selection and
// the backward to the action for one 128-row tile in ONE persistent, warp-specialised sm_100a kernel -- no
// activation ever leaves the SM (the ELU derivatives take one thread-private round trip through an L2-resident
// scratch).  One launch = one pass over all rows of all mode segments (one ascent iteration, or one inference
// call).
//
// Reference semantics (paths relative to the reference repo):
//   MLPNet (ELU) / DistributionalDoubleQ.get_q1_q2 / get_q_min     ddiffpg/models/mlp.py:13-35,143-151
//   the autograd backward of -get_q_min(obs, action).mean() w.r.t. action, ddiffpg/algo/ddiffpg.py:365-366
//
// Per tile and net (K-outer throughout: an activation chunk is consumed by the next layer as soon as it exists):
//   F1  [obs|act|0] (K=64)  . W1^T -> acc [128 x h1]   TMEM cols [0, h1)
//   F2  elu(acc+b1) chunks  . W2^T -> acc [128 x h2]   cols [0, h2)        (after those columns are drained)
//   F3  elu(..+b2)  chunks  . W3^T -> acc [128 x h3]   cols [h2, h2+h3)
//   F4  elu(..+b3)  chunks  . W4^T -> logits [128 x 64] cols [0, 64)  -> softmax, Q, dlogits = p (z - Q)
// then, once both nets' Q are known (only the smaller one carries gradient, ties split evenly):
//   B4  dlogits . W4 -> [128 x h3] cols [h2, h2+h3)  (* elu')
//   B3  dz3 chunks . W3 -> [128 x h2] cols [0, h2)   (* elu')
//   B2  dz2 chunks . W2 -> [128 x h1] cols [0, h1)   (* elu'; the column parts above h2 start before the
//                                                     previous accumulator is drained, the others after)
//   Ba  dz1 chunks . W1[:, O:O+A] -> [128 x 16] cols [0, 16)
// Weight tiles stream from L2 through a 3-stage TMA ring; all K+1 critics live in the same tensor maps (row
// offset = mode * N), a tile never straddles a mode segment.
#include <math.h>
#include <type_traits>
#include "q_layout.cuh"
#include "tc_common.cuh"
#include "scratch_cache.cuh"

namespace ddp {
using namespace tc;

// Adam action ascent carried inside the fused kernel (one cooperative launch for all iterations)
struct QChainAscent {
    int iters;
    float* act;                      // [B, A], updated in place
    float *m1, *m2;                  // zeroed Adam moments [B, A]
    float* gnorm_out;                // [n_modes, iters] or NULL
    unsigned int* grid_bar;          // zeroed
    float lr, beta1, beta2, eps, max_norm, lim;
};

namespace {

constexpr int kRows = 128;
constexpr int kChunkBytes = kRows * 128;
constexpr int kStageBytes = 256 * 128;
#ifndef DDP_QC_STAGES
#define DDP_QC_STAGES 3
#endif
constexpr int kStages = DDP_QC_STAGES;
constexpr int kASlots = 4;
#ifndef DDP_QC_EPI_WARPS
#define DDP_QC_EPI_WARPS 8
#endif
#ifndef DDP_QC_PAIR_PUBLISH
#define DDP_QC_PAIR_PUBLISH 0   // 1: two chunks per proxy fence (measured slower: the hand-off latency matters more)
#endif
constexpr bool kPairPublish = DDP_QC_PAIR_PUBLISH != 0;
constexpr int kEpiWarps = DDP_QC_EPI_WARPS;      // 8 or 16: 2 or 4 warps per SM sub-partition
constexpr int kColsPerWarp = 64 / (kEpiWarps / 4);  // columns of a chunk owned by one warp (32 or 16)
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = kEpiThreads + 64;
constexpr int kTmemCols = 512;
#ifndef DDP_QC_ABLATE
#define DDP_QC_ABLATE 0     // timing experiments only (results are wrong when set): 1 no proxy fence, 2 no ELU' scratch
#endif                      // traffic, 4 no MUFU, 8 no TMEM loads in the drains, 16 no drain arithmetic, 32 half the weight stream, 64 no bias loads, 128 no A-chunk stores
constexpr int kAblate = DDP_QC_ABLATE;     // measurements: profiles/r02/ablation_qchain.txt
constexpr int kBiasPerNet = 512 + 256 + 256 + 64;     // b1 | b2 | b3 | b4 slots (floats)
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kMaxIters = 32;               // iterations one cooperative launch can carry

struct alignas(64) QcMaps { CUtensorMap fwd[2][4]; CUtensorMap bwd[2][4]; CUtensorMap xin; };

struct QcArgs {
    const float* pk;                 // fp32 section of the packed critics (biases)
    size_t mode_stride;
    size_t b_off[2][4];
    float* g_out;                    // [B, A]  scale * d qmin / d action           (NULL: no backward)
    float* gsq;                      // [n_modes] += sum g^2 over the segment       (may be NULL)
    float* qmin;                     // [B]                                          (may be NULL)
    float* p_out[2];                 // [B, atoms] softmax of each net               (may be NULL)
    uint16_t* dscr;                  // ELU' scratch: [grid][2 nets][(h1+h2+h3)/16][2 halves][128 rows][8] bf16 (a warp's
                                     // 16-byte accesses are 512 contiguous bytes)
    long seg_off[kMaxModes + 1];
    int tile_off[kMaxModes + 1];
    float scale[kMaxModes];
    int n_modes, num_tiles;
    int O, A, atoms, h1, h2, h3, part1, nparts1;
    float v_min, dz;
    // multi-iteration ascent in one cooperative launch (iters > 1 or adam != 0): Adam / clip / clamp of
    // update_target_action applied by the CTA that owns the rows, after a grid-wide barrier on the clip norm
    int iters;                       // passes over the rows (1 = plain pass, the caller applies Adam)
    int fused_adam;
    float* act;                      // [B, A] fp32 actions, updated in place
    __nv_bfloat16* xin;              // [B, 64] bf16 rows [obs | act | 0]: action columns refreshed
    float *m1, *m2;                  // Adam moments [B, A]
    float* gnorm_out;                // [n_modes, iters] pre-clip norms (may be NULL)
    unsigned int* grid_bar;          // zeroed counter for the grid barrier
    float step_size[kMaxIters], bc2_sqrt[kMaxIters];
    float beta1, beta2, eps, max_norm, lim;
    long long* dbg;
};

struct SMQ {
    static constexpr uint32_t wring = 0;
    static constexpr uint32_t aring = wring + kStages * kStageBytes;
    static constexpr uint32_t dl = aring + kASlots * kChunkBytes;
    static constexpr uint32_t a0 = dl + 2 * kChunkBytes;                   // [obs | act | 0] chunk of the tile
    static constexpr uint32_t bias = a0 + kChunkBytes;
    static constexpr uint32_t bars = bias + 2 * kBiasPerNet * 4;
    static constexpr uint32_t tmem_ptr = bars + 8 * 24;
    static constexpr uint32_t total = tmem_ptr + 8;
};
static_assert(SMQ::total + 1024 <= 227 * 1024, "shared-memory map exceeds the 227 KB per-CTA limit");

__device__ __forceinline__ uint32_t qb_w_full(uint32_t b, int i) { return b + 8 * i; }
__device__ __forceinline__ uint32_t qb_w_empty(uint32_t b, int i) { return b + 8 * (kStages + i); }
__device__ __forceinline__ uint32_t qb_a_full(uint32_t b, int i) { return b + 8 * (2 * kStages + i); }
__device__ __forceinline__ uint32_t qb_a_empty(uint32_t b, int i) { return b + 8 * (2 * kStages + kASlots + i); }
__device__ __forceinline__ uint32_t qb_acc_full(uint32_t b, int i) { return b + 8 * (2 * kStages + 2 * kASlots + i); }
__device__ __forceinline__ uint32_t qb_lo_free(uint32_t b) { return b + 8 * (2 * kStages + 2 * kASlots + 2); }
__device__ __forceinline__ uint32_t qb_dl_full(uint32_t b) { return b + 8 * (2 * kStages + 2 * kASlots + 3); }
__device__ __forceinline__ uint32_t qb_a0_full(uint32_t b) { return b + 8 * (2 * kStages + 2 * kASlots + 4); }
__device__ __forceinline__ uint32_t qb_a0_free(uint32_t b) { return b + 8 * (2 * kStages + 2 * kASlots + 5); }
__device__ __forceinline__ uint32_t qb_acc_free(uint32_t b) { return b + 8 * (2 * kStages + 2 * kASlots + 6); }
__device__ __forceinline__ uint32_t qb_adam_done(uint32_t b) { return b + 8 * (2 * kStages + 2 * kASlots + 7); }

__device__ __forceinline__ void q_epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

struct QRing {
    int idx = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance(int n) { if (++idx == n) { idx = 0; phase ^= 1; } }
};

__device__ __forceinline__ int q_mode_of(const QcArgs& a, int tile) {
    int m = 0;
    while (m + 1 < a.n_modes && tile >= a.tile_off[m + 1]) ++m;
    return m;
}
// column part p of an h1-wide accumulator overlaps the columns [0, h2) still owned by the previous stage
__device__ __forceinline__ bool q_gated(const QcArgs& a, int p) { return p * a.part1 < a.h2; }

struct QEpi {
    uint8_t* smem;
    uint32_t bars, tmem_base, acc_cnt;
    int q, ch, lane, my_row;
    QRing as;
    uint16_t* dscr;                 // this CTA's scratch block
    long long* dbg;
    __device__ __forceinline__ void wait_acc() {
        mbar_wait(qb_acc_full(bars, acc_cnt & 1), (acc_cnt >> 1) & 1);
        ++acc_cnt;
        tc_fence_after();
    }
};

// ELU' scratch accesses (scratch_cache.cuh).  DDP_QC_SCRATCH, A/B switch measured in profiles/r02/ab_scratch_policy.txt:
// bit 0 = evict_last policy on the stores / loads (default: 185 -> 26 MB of DRAM writes per launch at no cost in time),
// bit 1 = a line is discarded from L2 once the backward has consumed it (10 MB with both, but 1.4 % slower: off).
#ifndef DDP_QC_SCRATCH
#define DDP_QC_SCRATCH 1
#endif
__device__ __forceinline__ void dscr_st(uint4* p, const uint4 w) { scratch_st<(DDP_QC_SCRATCH & 1) != 0>(p, w); }
__device__ __forceinline__ uint4 dscr_ld(const uint4* p) { return scratch_ld<(DDP_QC_SCRATCH & 1) != 0>(p); }
__device__ __forceinline__ void dscr_discard(const uint4* p, int lane, uint32_t loaded) {
    if ((DDP_QC_SCRATCH & 2) && (lane & 7) == 0) scratch_discard(p, loaded);      // 8 rows x 16 bytes = one line
}

__device__ __forceinline__ void store_chunk16(uint8_t* slot, int row, int col0, const float (&x)[16]) {
#pragma unroll
    for (int i8 = 0; i8 < 2; ++i8) {
        uint4 w;
        w.x = pack_bf16x2(x[i8 * 8 + 0], x[i8 * 8 + 1]); w.y = pack_bf16x2(x[i8 * 8 + 2], x[i8 * 8 + 3]);
        w.z = pack_bf16x2(x[i8 * 8 + 4], x[i8 * 8 + 5]); w.w = pack_bf16x2(x[i8 * 8 + 6], x[i8 * 8 + 7]);
        if (kAblate & 128) { if (w.x == 0x12345678u) *reinterpret_cast<uint4*>(slot) = w; }      // keeps the arithmetic alive
        else *reinterpret_cast<uint4*>(slot + sw128_offset(row, col0 + i8 * 8)) = w;
    }
}

// forward: 16 accumulator columns -> +bias, ELU -> A chunk (bf16); ELU' -> scratch (bf16) unless the launch has no
// backward half (KEEP_D = false: ddp_q_forward, the target heads of the critic update)
template <bool KEEP_D>
__device__ __forceinline__ void emit_fwd(const QEpi& e, uint8_t* slot, const uint32_t (&v)[16], const float* bb,
                                         int col0, uint16_t* dptr) {
    float x[16], d[16];
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) {
        const float4 b = (kAblate & 64) ? make_float4(0.1f, 0.2f, -0.1f, 0.f) : *reinterpret_cast<const float4*>(bb + i4 * 4);
        x[i4 * 4 + 0] = __uint_as_float(v[i4 * 4 + 0]) + b.x; x[i4 * 4 + 1] = __uint_as_float(v[i4 * 4 + 1]) + b.y;
        x[i4 * 4 + 2] = __uint_as_float(v[i4 * 4 + 2]) + b.z; x[i4 * 4 + 3] = __uint_as_float(v[i4 * 4 + 3]) + b.w;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] = (kAblate & 4) ? x[i] * kLog2e : ex2_approx(x[i] * kLog2e);
    // (a compare/select-free form -- t = min(x, 0), e = exp t, elu = (x - t) + (e - 1) -- measured 1.3 % slower)
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const bool neg = x[i] < 0.f;
        x[i] = neg ? d[i] - 1.f : x[i];
        d[i] = neg ? d[i] : 1.f;
    }
    store_chunk16(slot, e.my_row, col0, x);
    if (!KEEP_D) return;
    uint4 w0, w1;
    w0.x = pack_bf16x2(d[0], d[1]); w0.y = pack_bf16x2(d[2], d[3]); w0.z = pack_bf16x2(d[4], d[5]); w0.w = pack_bf16x2(d[6], d[7]);
    w1.x = pack_bf16x2(d[8], d[9]); w1.y = pack_bf16x2(d[10], d[11]); w1.z = pack_bf16x2(d[12], d[13]); w1.w = pack_bf16x2(d[14], d[15]);
    if (!(kAblate & 2)) {
        dscr_st(reinterpret_cast<uint4*>(dptr), w0);
        dscr_st(reinterpret_cast<uint4*>(dptr) + kRows, w1);
    }
}

// backward: 16 accumulator columns * ELU' -> A chunk (bf16)
__device__ __forceinline__ void emit_bwd(const QEpi& e, uint8_t* slot, const uint32_t (&v)[16], int col0, uint4 d0, uint4 d1,
                                         const uint4* dp) {
    float x[16];
    const uint32_t dw[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x[2 * i] = __uint_as_float(v[2 * i]) * __uint_as_float(dw[i] << 16);
        x[2 * i + 1] = __uint_as_float(v[2 * i + 1]) * __uint_as_float(dw[i] & 0xffff0000u);
    }
    store_chunk16(slot, e.my_row, col0, x);
    dscr_discard(dp, e.lane, d0.x);
    dscr_discard(dp + kRows, e.lane, d1.x);
}

// Drain `nchunks` 64-column chunks of the accumulator at TMEM column `col` into the A ring.  This warp owns 32
// columns of each chunk (two 16-column TMEM loads, the second in flight while the first is processed).
// FWD: +bias, ELU, derivative to scratch group `g0 + ...`; else: times the derivative read back from there.
template <bool FWD, bool KEEP_D = true>
__device__ __forceinline__ void q_drain(QEpi& e, int col, int nchunks, const float* bias, int g0, int signal_after) {
    const uint32_t tbase = e.tmem_base + ((uint32_t)(e.q * 32) << 16) + col + e.ch * kColsPerWarp;
    // scratch group of this warp's first 16 columns inside chunk 0 (4 groups per chunk)
    uint16_t* dbase = e.dscr + (size_t)(g0 + e.ch * (kColsPerWarp / 16)) * kRows * 16 + e.my_row * 8;
    uint32_t va[16], vb[16];
    uint4 d0{}, d1{}, d2{}, d3{};
    if (kAblate & 8) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { va[i] = (uint32_t)(col * 7 + i * e.lane); vb[i] = (uint32_t)(col * 5 + i * e.lane); }
    } else tmem_ld16(tbase, va);
    if (!FWD) {
        const uint4* dp = reinterpret_cast<const uint4*>(dbase);
        if (!(kAblate & 2)) {
            d0 = dscr_ld(dp); d1 = dscr_ld(dp + kRows);
            if (kColsPerWarp == 32) { d2 = dscr_ld(dp + kRows * 2); d3 = dscr_ld(dp + kRows * 3); }
        }
    }
#ifdef DDP_QC_FINE_TIMING
    long long* fine = (e.dbg && blockIdx.x == 0 && threadIdx.x == 0) ? e.dbg + (FWD ? 16 : 24) : nullptr;
    long long f0 = fine ? clock64() : 0, f1;
#define QC_FINE(slot) do { if (fine) { f1 = clock64(); fine[slot] += f1 - f0; f0 = f1; } } while (0)
#define QC_FINE_COUNT() do { if (fine) fine[6] += 1; } while (0)
#else
#define QC_FINE(slot) do { } while (0)
#define QC_FINE_COUNT() do { } while (0)
#endif
    for (int c0 = 0; c0 < nchunks; c0 += kPairPublish ? 2 : 1) {
        // two chunks per publication: the generic->async proxy fence is the expensive part of handing a chunk over
        const int n2 = (!kPairPublish || nchunks - c0 < 2) ? 1 : 2;
        QRing rs = e.as;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (u >= n2) break;
            const int c = c0 + u;
            uint16_t* dptr = dbase + (size_t)c * 4 * kRows * 16;
            if (kColsPerWarp == 32) {
                if (!(kAblate & 8)) { tmem_ld_wait(); tmem_ld16(tbase + c * 64 + 16, vb); }
                QC_FINE(0);
                mbar_wait(qb_a_empty(e.bars, rs.idx), rs.phase ^ 1);
                QC_FINE(1);
                uint8_t* slot = e.smem + SMQ::aring + rs.idx * kChunkBytes;
                rs.advance(kASlots);
                if (kAblate & 16) { }
                else if (FWD) emit_fwd<KEEP_D>(e, slot, va, bias + c * 64 + e.ch * 32, e.ch * 32, dptr);
                else emit_bwd(e, slot, va, e.ch * 32, d0, d1, reinterpret_cast<const uint4*>(dptr));
                QC_FINE(2);
                if (!(kAblate & 8)) { tmem_ld_wait(); if (c + 1 < nchunks) tmem_ld16(tbase + (c + 1) * 64, va); }
                QC_FINE(3);
                if (kAblate & 16) { }
                else if (FWD) emit_fwd<KEEP_D>(e, slot, vb, bias + c * 64 + e.ch * 32 + 16, e.ch * 32 + 16, dptr + kRows * 16);
                else {
                    emit_bwd(e, slot, vb, e.ch * 32 + 16, d2, d3, reinterpret_cast<const uint4*>(dptr) + kRows * 2);
                    if (c + 1 < nchunks) {
                        const uint4* dp = reinterpret_cast<const uint4*>(dptr + (size_t)4 * kRows * 16);
                        if (!(kAblate & 2)) {
                            d0 = dscr_ld(dp); d1 = dscr_ld(dp + kRows);
                            d2 = dscr_ld(dp + kRows * 2); d3 = dscr_ld(dp + kRows * 3);
                        }
                    }
                }
                QC_FINE(4);
            } else {
                // 16 columns per warp: the two register buffers alternate chunk by chunk (u is compile-time)
                tmem_ld_wait();
                if (c + 1 < nchunks) { if (u == 0) tmem_ld16(tbase + (c + 1) * 64, vb); else tmem_ld16(tbase + (c + 1) * 64, va); }
                uint4 e0 = d0, e1 = d1;
                if (!FWD && c + 1 < nchunks) {
                    const uint4* dp = reinterpret_cast<const uint4*>(dptr + (size_t)4 * kRows * 16);
                    d0 = dscr_ld(dp); d1 = dscr_ld(dp + kRows);
                }
                mbar_wait(qb_a_empty(e.bars, rs.idx), rs.phase ^ 1);
                uint8_t* slot = e.smem + SMQ::aring + rs.idx * kChunkBytes;
                rs.advance(kASlots);
                if (FWD) emit_fwd<KEEP_D>(e, slot, u == 0 ? va : vb, bias + c * 64 + e.ch * 16, e.ch * 16, dptr);
                else emit_bwd(e, slot, u == 0 ? va : vb, e.ch * 16, e0, e1, reinterpret_cast<const uint4*>(dptr));
            }
            QC_FINE_COUNT();
        }
        if (!(kAblate & 1)) fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (e.lane == 0) {
            mbar_arrive(qb_a_full(e.bars, e.as.idx));
            if (n2 == 2) mbar_arrive(qb_a_full(e.bars, e.as.idx + 1 == kASlots ? 0 : e.as.idx + 1));
            if (signal_after >= c0 && signal_after < c0 + n2) mbar_arrive(qb_lo_free(e.bars));
        }
        e.as = rs;
        QC_FINE(5);
    }
#undef QC_FINE
#undef QC_FINE_COUNT
}

template <bool BWD>
#if DDP_QC_EPI_WARPS >= 16
// 18 warps: ptxas would round the block up to 640 threads and cap at 96 registers
__global__ void __maxnreg__(96)
#else
__global__ void __launch_bounds__(kThreads, 1)
#endif
q_chain_tc_kernel(const __grid_constant__ QcMaps maps, const QcArgs a) {
    // BWD = false: the forward half alone (g_out == NULL) -- no ELU' scratch, none of the backward stages in the image
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    const uint32_t bars = base + SMQ::bars;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NC1 = a.h1 >> 6, NC2 = a.h2 >> 6, NC3 = a.h3 >> 6;
    const bool backward = BWD && a.g_out != nullptr;     // (runtime in the BWD image: keeps its register allocation)
    // B2 in two halves (h1 = 2 x 256 column parts): the part above h2 is complete as soon as the last dz2 chunk has been
    // multiplied, so its chunks are drained while the part that overlaps the B3 accumulator is still being multiplied.
    // The action-gradient accumulator (Ba) then lives in drained columns of the upper part, clear of the next net's B4
    // accumulator [h2, h2 + h3).
    const bool split_b2 = a.nparts1 == 2 && (NC1 & 1) == 0 && a.h2 + a.h3 + 16 <= a.h1 && a.h2 + a.h3 >= a.part1 &&
                          ((a.h2 + a.h3) & 63) == 0;
    const int ba_col = split_b2 ? a.h2 + a.h3 : 0;                    // TMEM column of the Ba accumulator
    const int ba_gate = split_b2 ? (ba_col - a.part1) >> 6 : 0;       // upper-part chunk whose drain frees those columns

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(qb_w_full(bars, i), 1); mbar_init(qb_w_empty(bars, i), 1); }
        for (int i = 0; i < kASlots; ++i) { mbar_init(qb_a_full(bars, i), kEpiWarps); mbar_init(qb_a_empty(bars, i), 1); }
        mbar_init(qb_acc_full(bars, 0), 1);
        mbar_init(qb_acc_full(bars, 1), 1);
        mbar_init(qb_lo_free(bars), kEpiWarps);
        mbar_init(qb_dl_full(bars), 4);
        mbar_init(qb_a0_full(bars), 1);
        mbar_init(qb_a0_free(bars), 1);
        mbar_init(qb_acc_free(bars), kEpiWarps);
        mbar_init(qb_adam_done(bars), kEpiWarps);
        fence_barrier_init();
    }
    if (warp == kEpiWarps + 1) tmem_alloc(base + SMQ::tmem_ptr, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + SMQ::tmem_ptr);

    if (warp == kEpiWarps) {
        // ============================================================== TMA producer (one lane)
        if (lane == 0) {
            for (int j = 0; j < 2; ++j)
                for (int i = 0; i < 4; ++i) { tma_prefetch_desc(&maps.fwd[j][i]); tma_prefetch_desc(&maps.bwd[j][i]); }
            QRing ws;
            uint32_t nload = 0;
            auto load = [&](const CUtensorMap* m, int x, int y, int rows) {
                mbar_wait(qb_w_empty(bars, ws.idx), ws.phase ^ 1);
                if ((kAblate & 32) && (nload++ & 1)) {
                    mbar_expect_tx(qb_w_full(bars, ws.idx), 0u);           // ablation: every other weight tile is not fetched
                } else {
                    mbar_expect_tx(qb_w_full(bars, ws.idx), (uint32_t)rows * 128u);
                    tma_load_2d(base + SMQ::wring + ws.idx * kStageBytes, m, qb_w_full(bars, ws.idx), x, y);
                }
                ws.advance(kStages);
            };
            tma_prefetch_desc(&maps.xin);
            // the [obs | act | 0] rows of a tile: one 16 KB box into the dedicated buffer, fetched one tile ahead
            uint32_t a0_free_phase = 0;
            auto load_input = [&](int tile, bool first) {
                if (!first) { mbar_wait(qb_a0_free(bars), a0_free_phase); a0_free_phase ^= 1; }
                const int m = q_mode_of(a, tile);
                const long row0 = a.seg_off[m] + (long)(tile - a.tile_off[m]) * kRows;
                mbar_expect_tx(qb_a0_full(bars), kChunkBytes);
                tma_load_2d(base + SMQ::a0, &maps.xin, qb_a0_full(bars), 0, (int)row0);
            };
            uint32_t adam_phase = 0;
            for (int it = 0; it < a.iters; ++it) {
            if (it > 0) {
                // the owning CTA has rewritten the action columns of its rows (generic proxy, fenced) for this pass
                mbar_wait(qb_adam_done(bars), adam_phase);
                adam_phase ^= 1;
            }
            load_input(blockIdx.x, it == 0);
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
                const int m = q_mode_of(a, tile);
                for (int j = 0; j < 2; ++j) {
                    for (int p = 0; p < a.nparts1; ++p) load(&maps.fwd[j][0], 0, m * a.h1 + p * a.part1, a.part1);
                    for (int c = 0; c < NC1; ++c) load(&maps.fwd[j][1], c * 64, m * a.h2, a.h2);
                    for (int c = 0; c < NC2; ++c) load(&maps.fwd[j][2], c * 64, m * a.h3, a.h3);
                    for (int c = 0; c < NC3; ++c) load(&maps.fwd[j][3], c * 64, m * 64, 64);
                }
                // both first layers of this tile have been issued long ago: their input buffer can take the next tile
                if (tile + (int)gridDim.x < a.num_tiles) load_input(tile + (int)gridDim.x, false);
                if (!backward) continue;
                for (int j = 0; j < 2; ++j) {
                    load(&maps.bwd[j][0], 0, m * a.h3, a.h3);
                    for (int c = 0; c < NC3; ++c) load(&maps.bwd[j][1], c * 64, m * a.h2, a.h2);
                    for (int c = 0; c < NC2; ++c)
                        for (int p = 0; p < a.nparts1; ++p)
                            if (!q_gated(a, p)) load(&maps.bwd[j][2], c * 64, m * a.h1 + p * a.part1, a.part1);
                    for (int c = 0; c < NC2; ++c)
                        for (int p = 0; p < a.nparts1; ++p)
                            if (q_gated(a, p)) load(&maps.bwd[j][2], c * 64, m * a.h1 + p * a.part1, a.part1);
                    for (int c = 0; c < NC1; ++c) {
                        const int cc = split_b2 ? (c + NC1 / 2) % NC1 : c;      // upper-part chunks are drained first
                        load(&maps.bwd[j][3], cc * 64, m * 16, 16);
                    }
                }
            }
            }
        }
    } else if (warp == kEpiWarps + 1) {
        // ============================================================== MMA issuer (one lane)
        if (lane == 0) {
            QRing ws, as;
            uint32_t lo_phase = 0, dl_phase = 0, a0_phase = 0, free_phase = 0, acc_cnt = 0;
            const uint32_t id_p1 = make_idesc_bf16(kRows, a.part1), id_h2 = make_idesc_bf16(kRows, a.h2),
                           id_h3 = make_idesc_bf16(kRows, a.h3), id_64 = make_idesc_bf16(kRows, 64),
                           id_16 = make_idesc_bf16(kRows, 16);
            // one weight stage against one A chunk: 4 K-steps of 16
            auto mma_w = [&](uint32_t tcol, uint64_t adesc, uint32_t idesc, bool first) {
                mbar_wait(qb_w_full(bars, ws.idx), ws.phase);
                tc_fence_after();
                const uint64_t bdesc = make_smem_desc_sw128(base + SMQ::wring + ws.idx * kStageBytes);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + tcol, adesc + 2 * k, bdesc + 2 * k, idesc, !(first && k == 0));
                umma_commit(qb_w_empty(bars, ws.idx));
                ws.advance(kStages);
            };
            auto a_wait = [&]() -> uint64_t {
                mbar_wait(qb_a_full(bars, as.idx), as.phase);
                tc_fence_after();
                return make_smem_desc_sw128(base + SMQ::aring + as.idx * kChunkBytes);
            };
            auto a_release = [&]() { umma_commit(qb_a_empty(bars, as.idx)); as.advance(kASlots); };
            auto acc_done = [&]() { umma_commit(qb_acc_full(bars, acc_cnt & 1)); ++acc_cnt; };
            auto lo_wait = [&]() { mbar_wait(qb_lo_free(bars), lo_phase); lo_phase ^= 1; tc_fence_after(); };
            // plain K-outer stage: nchunks ring chunks, one weight stage each
            auto stage = [&](int nchunks, uint32_t tcol, uint32_t idesc) {
                for (int c = 0; c < nchunks; ++c) {
                    const uint64_t ad = a_wait();
                    mma_w(tcol, ad, idesc, c == 0);
                    a_release();
                }
                acc_done();
            };
            for (int it = 0; it < a.iters; ++it)
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
                for (int j = 0; j < 2; ++j) {
                    // the epilogue warps have read everything they need from the accumulator columns (j == 0: the
                    // previous tile's action gradient; j == 1: the logits of net 0)
                    mbar_wait(qb_acc_free(bars), free_phase);
                    free_phase ^= 1;
                    if (j == 0) { mbar_wait(qb_a0_full(bars), a0_phase); a0_phase ^= 1; }
                    tc_fence_after();
                    {   // F1: the tile's input chunk against the column parts of W1
                        const uint64_t ad = make_smem_desc_sw128(base + SMQ::a0);
                        for (int p = 0; p < a.nparts1; ++p) mma_w(p * a.part1, ad, id_p1, true);
                        if (j == 1) umma_commit(qb_a0_free(bars));
                        acc_done();
                    }
                    lo_wait();                      // columns [0, h2) of the F1 accumulator are drained
                    stage(NC1, 0, id_h2);           // F2
                    stage(NC2, a.h2, id_h3);        // F3
                    stage(NC3, 0, id_64);           // F4
                }
                if (!backward) continue;
                mbar_wait(qb_dl_full(bars), dl_phase);
                dl_phase ^= 1;
                tc_fence_after();
                for (int j = 0; j < 2; ++j) {
                    mma_w(a.h2, make_smem_desc_sw128(base + SMQ::dl + j * kChunkBytes), id_h3, true);   // B4
                    acc_done();
                    stage(NC3, 0, id_h2);           // B3
                    {   // B2: free column parts chunk by chunk, then the parts that overlap the B3 accumulator
                        uint64_t ad[kASlots];
                        QRing look = as;
                        for (int c = 0; c < NC2; ++c) {
                            mbar_wait(qb_a_full(bars, look.idx), look.phase);
                            tc_fence_after();
                            ad[c] = make_smem_desc_sw128(base + SMQ::aring + look.idx * kChunkBytes);
                            look.advance(kASlots);
                            for (int p = 0; p < a.nparts1; ++p)
                                if (!q_gated(a, p)) mma_w(p * a.part1, ad[c], id_p1, c == 0);
                        }
                        if (split_b2) acc_done();   // the upper part is complete: its drain starts now
                        lo_wait();                  // the B3 accumulator is fully drained
                        for (int c = 0; c < NC2; ++c) {
                            for (int p = 0; p < a.nparts1; ++p)
                                if (q_gated(a, p)) mma_w(p * a.part1, ad[c], id_p1, c == 0);
                            a_release();
                        }
                        acc_done();
                    }
                    lo_wait();                      // the columns the Ba accumulator takes are drained
                    stage(NC1, ba_col, id_16);      // Ba
                }
            }
        }
    } else {
        // ============================================================== epilogue warps
        QEpi e;
        e.smem = smem; e.bars = bars; e.tmem_base = tmem_base; e.acc_cnt = 0;
        e.q = warp & 3; e.ch = warp >> 2; e.lane = lane; e.my_row = e.q * 32 + lane;
        e.dbg = a.dbg;
        const int G = (a.h1 + a.h2 + a.h3) >> 4;              // 16-column scratch groups per net
        const int g1 = 0, g2 = a.h1 >> 4, g3 = (a.h1 + a.h2) >> 4;
        float* sbias = reinterpret_cast<float*>(smem + SMQ::bias);
        const bool prof = a.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
        long long tk0 = prof ? clock64() : 0, tk1;
#define QC_TICK(slot) do { if (prof) { tk1 = clock64(); a.dbg[slot] += tk1 - tk0; tk0 = tk1; } } while (0)
        int cur_mode = -1;
        if (lane == 0) mbar_arrive(qb_acc_free(bars));          // nothing to drain before the first tile
        for (int it = 0; it < a.iters; ++it) {
        const long long iter_t0 = clock64();
        float* gsq_it = a.gsq ? a.gsq + (size_t)it * kMaxModes : nullptr;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
            const int m = q_mode_of(a, tile);
            const long row = a.seg_off[m] + (long)(tile - a.tile_off[m]) * kRows + e.my_row;
            const bool valid = row < a.seg_off[m + 1];
            if (m != cur_mode) {
                q_epi_bar_sync();                 // every warp is done with the previous mode's biases
                const float* pm = a.pk + (size_t)m * a.mode_stride;
                for (int j = 0; j < 2; ++j) {
                    float* sb = sbias + j * kBiasPerNet;
                    for (int i = threadIdx.x; i < a.h1; i += kEpiThreads) sb[i] = pm[a.b_off[j][0] + i];
                    for (int i = threadIdx.x; i < a.h2; i += kEpiThreads) sb[512 + i] = pm[a.b_off[j][1] + i];
                    for (int i = threadIdx.x; i < a.h3; i += kEpiThreads) sb[768 + i] = pm[a.b_off[j][2] + i];
                    for (int i = threadIdx.x; i < 64; i += kEpiThreads) sb[1024 + i] = i < a.atoms ? pm[a.b_off[j][3] + i] : -INFINITY;
                }
                q_epi_bar_sync();
                cur_mode = m;
            }
            float qv[2] = {0.f, 0.f};
            for (int j = 0; j < 2; ++j) {
                e.dscr = a.dscr + ((size_t)blockIdx.x * 2 + j) * G * kRows * 16;
                const float* sb = sbias + j * kBiasPerNet;
                QC_TICK(0);
                e.wait_acc();
                QC_TICK(1);
                q_drain<true, BWD>(e, 0, NC1, sb, g1, NC2 - 1);          // F1 accumulator -> a1 chunks
                QC_TICK(2);
                e.wait_acc();
                q_drain<true, BWD>(e, 0, NC2, sb + 512, g2, -1);         // F2 -> a2
                QC_TICK(3);
                e.wait_acc();
                q_drain<true, BWD>(e, a.h2, NC3, sb + 768, g3, -1);      // F3 -> a3
                QC_TICK(4);
                if (e.ch == 0) {
                    // logits -> softmax, expectation, d Q / d logits (unmasked) for this thread's row
                    e.wait_acc();
                    // columns >= atoms carry a bias of -inf (see the bias load): their exponentials are exact zeros
                    float l[64];
                    {
                        uint32_t v0[32], v1[32];
                        tmem_ld32(tmem_base + ((uint32_t)(e.q * 32) << 16), v0);
                        tmem_ld32(tmem_base + ((uint32_t)(e.q * 32) << 16) + 32, v1);
                        tmem_ld_wait();
#pragma unroll
                        for (int i4 = 0; i4 < 8; ++i4) {
                            const float4 b0 = *reinterpret_cast<const float4*>(sb + 1024 + i4 * 4);
                            const float4 b1 = *reinterpret_cast<const float4*>(sb + 1056 + i4 * 4);
                            l[i4 * 4 + 0] = __uint_as_float(v0[i4 * 4 + 0]) + b0.x; l[i4 * 4 + 1] = __uint_as_float(v0[i4 * 4 + 1]) + b0.y;
                            l[i4 * 4 + 2] = __uint_as_float(v0[i4 * 4 + 2]) + b0.z; l[i4 * 4 + 3] = __uint_as_float(v0[i4 * 4 + 3]) + b0.w;
                            l[32 + i4 * 4 + 0] = __uint_as_float(v1[i4 * 4 + 0]) + b1.x; l[32 + i4 * 4 + 1] = __uint_as_float(v1[i4 * 4 + 1]) + b1.y;
                            l[32 + i4 * 4 + 2] = __uint_as_float(v1[i4 * 4 + 2]) + b1.z; l[32 + i4 * 4 + 3] = __uint_as_float(v1[i4 * 4 + 3]) + b1.w;
                        }
                    }
                    // the accumulator columns are free again: the MMA warp may start the next net's first layer
                    tc_fence_before();
                    __syncwarp();
                    if ((j == 0 || !backward) && lane == 0) mbar_arrive(qb_acc_free(bars));
                    float mx4[4] = {l[0], l[1], l[2], l[3]};
#pragma unroll
                    for (int i = 4; i < 64; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], l[i]);
                    const float mxl = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * kLog2e;
                    float sum4[4] = {0.f, 0.f, 0.f, 0.f}, qz4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int i = 0; i < 64; ++i) {
                        const float ex = ex2_approx(fmaf(l[i], kLog2e, -mxl));
                        l[i] = ex;
                        sum4[i & 3] += ex;
                        qz4[i & 3] = fmaf(ex, (float)i, qz4[i & 3]);
                    }
                    const float inv = 1.f / ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
                    const float Qi = ((qz4[0] + qz4[1]) + (qz4[2] + qz4[3])) * inv;      // expectation of the atom index
                    qv[j] = fmaf(a.dz, Qi, a.v_min);
                    float* pout = a.p_out[j];
                    if (pout && valid) {
#pragma unroll
                        for (int at = 0; at < 64; ++at)
                            if (at < a.atoms) pout[row * a.atoms + at] = l[at] * inv;
                    }
                    uint8_t* dl = smem + SMQ::dl + j * kChunkBytes;
                    const float sdz = a.dz * inv;                                       // p (z - Q) = ex * inv * dz * (i - Qi)
#pragma unroll
                    for (int i8 = 0; i8 < 8; ++i8) {
                        float d[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) d[i] = l[i8 * 8 + i] * (sdz * ((float)(i8 * 8 + i) - Qi));
                        uint4 w;
                        w.x = pack_bf16x2(d[0], d[1]); w.y = pack_bf16x2(d[2], d[3]);
                        w.z = pack_bf16x2(d[4], d[5]); w.w = pack_bf16x2(d[6], d[7]);
                        *reinterpret_cast<uint4*>(dl + sw128_offset(e.my_row, i8 * 8)) = w;
                    }
                    tc_fence_before();
                } else {
                    ++e.acc_cnt;                   // the logits stage is read by the ch == 0 warps only
                    if ((j == 0 || !backward) && lane == 0) mbar_arrive(qb_acc_free(bars));
                }
                QC_TICK(5);
            }
            if (e.ch == 0) {
                if (a.qmin && valid) a.qmin[row] = fminf(qv[0], qv[1]);
                if (backward) {
                    // torch.min backward: the smaller head takes the gradient, exact ties split it evenly
                    const float w0 = qv[0] == qv[1] ? 0.5f : (qv[0] < qv[1] ? 1.f : 0.f);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const float w = j == 0 ? w0 : 1.f - w0;
                        if (w == 1.f) continue;
                        uint8_t* dl = smem + SMQ::dl + j * kChunkBytes;
                        for (int i8 = 0; i8 < 8; ++i8) {
                            uint4* ptr = reinterpret_cast<uint4*>(dl + sw128_offset(e.my_row, i8 * 8));
                            uint4 v = make_uint4(0, 0, 0, 0);
                            if (w != 0.f) {
                                v = *ptr;
                                uint32_t* u = reinterpret_cast<uint32_t*>(&v);
                                for (int k = 0; k < 4; ++k)
                                    u[k] = pack_bf16x2(w * __uint_as_float(u[k] << 16), w * __uint_as_float(u[k] & 0xffff0000u));
                            }
                            *ptr = v;
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(qb_dl_full(bars));
                }
            }
            if (!backward) continue;
            float da[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) da[i] = 0.f;
            for (int j = 0; j < 2; ++j) {
                e.dscr = a.dscr + ((size_t)blockIdx.x * 2 + j) * G * kRows * 16;
                e.wait_acc();
                QC_TICK(6);
                q_drain<false>(e, a.h2, NC3, nullptr, g3, -1);         // B4 accumulator * elu'(z3) -> dz3 chunks
                e.wait_acc();
                q_drain<false>(e, 0, NC2, nullptr, g2, NC2 - 1);       // B3 -> dz2 (then the low columns are free)
                QC_TICK(7);
                e.wait_acc();
                if (split_b2) {
                    q_drain<false>(e, a.part1, NC1 / 2, nullptr, g1 + (NC1 / 2) * 4, ba_gate);   // upper part -> dz1 chunks NC1/2..
                    e.wait_acc();
                    q_drain<false>(e, 0, NC1 / 2, nullptr, g1, -1);                              // lower part
                } else
                q_drain<false>(e, 0, NC1, nullptr, g1, 0);             // B2 -> dz1
                QC_TICK(8);
                if (e.ch == 0) {
                    e.wait_acc();
                    uint32_t v[16];
                    tmem_ld16(tmem_base + ((uint32_t)(e.q * 32) << 16) + ba_col, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) da[i] += __uint_as_float(v[i]);
                    tc_fence_before();
                    __syncwarp();
                    if (j == 1 && lane == 0) mbar_arrive(qb_acc_free(bars));     // the next tile may overwrite the columns
                } else {
                    ++e.acc_cnt;
                    if (j == 1 && lane == 0) mbar_arrive(qb_acc_free(bars));
                }
                QC_TICK(9);
            }
            if (e.ch == 0) {
                const float sc = a.scale[m];
                float ss = 0.f;
                if (valid) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (i < a.A) { const float g = sc * da[i]; a.g_out[row * a.A + i] = g; ss += g * g; }
                }
                if (gsq_it) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
                    if (lane == 0 && ss != 0.f) atomicAdd(gsq_it + m, ss);
                }
            }
        }
        if (a.fused_adam) {
            // ---- grid-wide barrier: every CTA has added its rows to the clip norms of this iteration
            q_epi_bar_sync();
            if (a.dbg && threadIdx.x == 0) a.dbg[32 + blockIdx.x] += clock64() - iter_t0;   // work of this pass, before the wait
            if (threadIdx.x == 0) {
                __threadfence();
                atomicAdd(a.grid_bar, 1u);
                const unsigned target = (unsigned)(it + 1) * gridDim.x;
                unsigned seen, spins = 0;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(a.grid_bar) : "memory");
                    if (seen < target) { __nanosleep(64); if (++spins > (1u << 26)) asm volatile("trap;"); }
                } while (seen < target);
            }
            q_epi_bar_sync();
            // ---- clip_grad_norm_ + Adam + clamp_ on the rows this CTA owns (ac_base.py:86-91, ddiffpg.py:362-369)
            const float step_size = a.step_size[it], bc2_sqrt = a.bc2_sqrt[it];
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
                const int m = q_mode_of(a, tile);
                const long row0 = a.seg_off[m] + (long)(tile - a.tile_off[m]) * kRows;
                const long left = a.seg_off[m + 1] - row0;
                const int nv = left < kRows ? (int)left : kRows;
                const float norm = sqrtf(__ldcg(gsq_it + m));
                const float coef = fminf(a.max_norm / (norm + 1e-6f), 1.0f);
                for (int i = threadIdx.x; i < nv * a.A; i += kEpiThreads) {
                    const long idx = row0 * a.A + i;
                    const float gi = __ldcg(a.g_out + idx) * coef;
                    const float ea = a.m1[idx] * a.beta1 + (1.f - a.beta1) * gi;
                    const float ev = a.m2[idx] * a.beta2 + (1.f - a.beta2) * gi * gi;
                    a.m1[idx] = ea; a.m2[idx] = ev;
                    float v = a.act[idx] - step_size * (ea / (sqrtf(ev) / bc2_sqrt + a.eps));
                    v = fminf(fmaxf(v, -a.lim), a.lim);
                    a.act[idx] = v;
                    const int r = i / a.A, c = i - r * a.A;
                    a.xin[(row0 + r) * 64 + a.O + c] = __float2bfloat16(v);
                }
            }
            if (blockIdx.x == 0 && a.gnorm_out && (int)threadIdx.x < a.n_modes)
                a.gnorm_out[threadIdx.x * a.iters + it] = sqrtf(__ldcg(gsq_it + threadIdx.x));
            // the refreshed rows are read back through TMA (async proxy) by this CTA's producer lane
            asm volatile("fence.proxy.async;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(qb_adam_done(bars));
        }
        }
#undef QC_TICK
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kEpiWarps + 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace

static long long* g_qc_dbg = nullptr;

bool q_chain_shape_ok(const QLayout& L) {
    return L.K1c == 64 && L.AtP == 64 && L.A <= 16 && L.A >= 1 && L.h1 % 64 == 0 && L.h2 % 64 == 0 && L.h3 % 64 == 0 &&
           L.h1 <= 512 && L.h2 <= 256 && L.h3 <= 256 && L.h2 + L.h3 <= kTmemCols && L.h2 <= L.h1 && L.h2 >= 64 &&
           (L.h1 <= 256 || L.h1 % 256 == 0);
}

static int q_chain_grid(long B, int n_modes, const int64_t* seg_off, int* tile_off) {
    int t = 0;
    for (int m = 0; m < n_modes; ++m) {
        tile_off[m] = t;
        t += (int)((seg_off[m + 1] - seg_off[m] + kRows - 1) / kRows);
    }
    tile_off[n_modes] = t;
    (void)B;
    return t;
}

size_t q_chain_workspace(const QLayout& L) {
    // scratch of the ELU derivatives: one block per CTA of the persistent grid (at most one CTA per SM; sized for
    // 160 so that the query does not need the device)
    return (size_t)160 * 2 * ((L.h1 + L.h2 + L.h3) / 16) * kRows * 16 * 2;
}

// One pass over all rows; xin = [B][64] bf16 rows [obs | action | 0].  g_out != NULL: backward to the action, g_out = scale[m] * d qmin / d action and
// gsq[m] += sum over the segment of g^2 (gsq may be NULL).  qmin / p1 / p2 optional.
int q_chain_pass(const QLayout& L, const void* packed, const int64_t* seg_off, const float* scale, const void* xin,
                 float* g_out, float* gsq, float* qmin, float* p1, float* p2, long B, void* scratch,
                 size_t scratch_bytes, cudaStream_t st, const QChainAscent* asc) {
    if (!q_chain_shape_ok(L)) DDP_FAIL(DDP_ERR_UNSUPPORTED, "fused critic kernel does not support this shape");
    if (!scratch || scratch_bytes < q_chain_workspace(L)) DDP_FAIL(DDP_ERR_ARG, "fused critic kernel: scratch too small");
    if ((DDP_QC_SCRATCH & 2) && ((uintptr_t)scratch & 127)) DDP_FAIL(DDP_ERR_ARG, "fused critic kernel: scratch must be 128-byte aligned");
    const uint8_t* pb = (const uint8_t*)packed;
    QcArgs a{};
    a.pk = (const float*)packed;
    a.mode_stride = L.mode_stride;
    for (int j = 0; j < 2; ++j) {
        a.b_off[j][0] = L.net[j].b1; a.b_off[j][1] = L.net[j].b2; a.b_off[j][2] = L.net[j].b3; a.b_off[j][3] = L.net[j].b4;
    }
    a.g_out = g_out; a.gsq = gsq; a.qmin = qmin; a.p_out[0] = p1; a.p_out[1] = p2;
    a.dscr = (uint16_t*)scratch;
    a.n_modes = L.n_modes;
    for (int m = 0; m <= L.n_modes; ++m) a.seg_off[m] = seg_off[m];
    a.num_tiles = q_chain_grid(B, L.n_modes, seg_off, a.tile_off);
    for (int m = 0; m < L.n_modes; ++m) a.scale[m] = scale ? scale[m] : 1.f;
    a.O = L.O; a.A = L.A; a.atoms = L.atoms; a.h1 = L.h1; a.h2 = L.h2; a.h3 = L.h3;
    a.nparts1 = L.h1 > 256 ? L.h1 / 256 : 1;
    a.part1 = L.h1 / a.nparts1;
    a.v_min = L.v_min;
    a.dz = (L.v_max - L.v_min) / (float)(L.atoms - 1);
    a.dbg = g_qc_dbg;
    a.iters = 1;
    if (asc) {
        if (asc->iters < 1 || asc->iters > kMaxIters) DDP_FAIL(DDP_ERR_ARG, "fused ascent carries 1..%d iterations per launch", kMaxIters);
        a.iters = asc->iters; a.fused_adam = 1;
        a.act = asc->act; a.xin = (__nv_bfloat16*)const_cast<void*>(xin); a.m1 = asc->m1; a.m2 = asc->m2;
        a.gnorm_out = asc->gnorm_out; a.grid_bar = asc->grid_bar;
        a.beta1 = asc->beta1; a.beta2 = asc->beta2; a.eps = asc->eps; a.max_norm = asc->max_norm; a.lim = asc->lim;
        for (int it = 0; it < asc->iters; ++it) {
            const double bc1 = 1.0 - pow((double)asc->beta1, it + 1), bc2 = 1.0 - pow((double)asc->beta2, it + 1);
            a.step_size[it] = (float)(asc->lr / bc1);
            a.bc2_sqrt[it] = (float)sqrt(bc2);
        }
    }
    if (a.num_tiles == 0) return DDP_OK;
    QcMaps maps;
    const uint64_t M = (uint64_t)L.n_modes;
    int bad = 0;
    for (int j = 0; j < 2; ++j) {
        bad |= make_tmap_bf16_sw128(&maps.fwd[j][0], pb + L.tc_fwd[j][0], M * L.h1, 64, a.part1);
        bad |= make_tmap_bf16_sw128(&maps.fwd[j][1], pb + L.tc_fwd[j][1], M * L.h2, L.h1, L.h2);
        bad |= make_tmap_bf16_sw128(&maps.fwd[j][2], pb + L.tc_fwd[j][2], M * L.h3, L.h2, L.h3);
        bad |= make_tmap_bf16_sw128(&maps.fwd[j][3], pb + L.tc_fwd[j][3], M * 64, L.h3, 64);
        bad |= make_tmap_bf16_sw128(&maps.bwd[j][0], pb + L.tc_bwd[j][0], M * L.h3, 64, L.h3);
        bad |= make_tmap_bf16_sw128(&maps.bwd[j][1], pb + L.tc_bwd[j][1], M * L.h2, L.h3, L.h2);
        bad |= make_tmap_bf16_sw128(&maps.bwd[j][2], pb + L.tc_bwd[j][2], M * L.h1, L.h2, a.part1);
        bad |= make_tmap_bf16_sw128(&maps.bwd[j][3], pb + L.tc_bwd[j][3], M * 16, L.h1, 16);
    }
    bad |= make_tmap_bf16_sw128(&maps.xin, xin, (uint64_t)B, 64, kRows);        // rows past B read as zeros
    if (bad) DDP_FAIL(DDP_ERR_CUDA, "cuTensorMapEncodeTiled failed for the critic weight tiles");
    int dev = 0, sms = 0;
    DDP_CUDA_CHECK(cudaGetDevice(&dev));
    DDP_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (sms > 160) sms = 160;
    const size_t smem = SMQ::total + 1024;
    const bool bwd = g_out != nullptr;
    if (asc && !bwd) DDP_FAIL(DDP_ERR_ARG, "fused ascent needs the gradient buffer");
    DDP_CUDA_CHECK(cudaFuncSetAttribute(bwd ? q_chain_tc_kernel<true> : q_chain_tc_kernel<false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = a.num_tiles < sms ? a.num_tiles : sms;
    if (asc) {
        // the grid barrier needs every CTA resident at once: cooperative launch, one CTA per SM
        void* kargs[2] = {(void*)&maps, (void*)&a};
        cudaError_t err = cudaLaunchCooperativeKernel((const void*)q_chain_tc_kernel<true>, dim3(grid), dim3(kThreads), kargs, smem, st);
        if (err != cudaSuccess) DDP_FAIL(DDP_ERR_CUDA, "cooperative launch of q_chain_tc_kernel failed: %s", cudaGetErrorString(err));
        return DDP_OK;
    }
    if (bwd) q_chain_tc_kernel<true><<<grid, kThreads, smem, st>>>(maps, a);
    else q_chain_tc_kernel<false><<<grid, kThreads, smem, st>>>(maps, a);
    DDP_LAUNCH_CHECK("q_chain_tc_kernel");
    return DDP_OK;
}

}  // namespace ddp

// Debug entry point (not part of the public header): device buffer of >= 16 int64 that CTA 0 of the next fused
// critic launches accumulates per-phase cycle counts into (NULL switches it off).
extern "C" void ddp_debug_qc_timing(long long* dev_buf) { ddp::g_qc_dbg = dev_buf; }
