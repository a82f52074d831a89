// H1, bf16 tensor-core path: the whole T-step DDPM reverse chain for a 128-row tile in one persistent,
// warp-specialised sm_100a kernel.  Operands bf16, accumulation fp32 (TMEM), epilogues fp32.
//
// Reference semantics (paths relative to the reference repo):
//   DiffusionPolicy.get_actions(sample=True)      ddiffpg/models/diffusion_mlp.py:219-251
//   DiffusionNet.forward (trunk, Mish)            ddiffpg/models/diffusion_mlp.py:50-58,62-73
//   DDPMScheduler.step                            diffusers ^0.18.2, call site diffusion_mlp.py:243-247
//
// Per CTA (one per SM, persistent over 128-row tiles), per denoising step:
//   layer 0 (K = [x|state] = 42 -> 48): warp-level mma.sync in the 8 epilogue warps, register
//            accumulators; + time table, Mish, bf16 -> 64-column A chunks in SWIZZLE_128B shared memory.
//            Only x_t changes between the steps of a tile: the contribution of state[8:] (k16 steps 1 and 2)
//            is computed once per tile, parked as fp16 in an L2-resident scratch in the owning lane's fragment
//            order, and each step adds one k16 step ([x_t | state[0:8]]) on top of it
//   layer 1 (K = h1): tcgen05.mma, A = those chunks as they appear (K-outer), B = W1 tiles streamed by
//            TMA from L2 through a 4-stage ring, D = the full [128 x h2] fp32 accumulator in TMEM
//   layer 2 (K = h2): A = Mish(acc1) chunks drained from TMEM by the epilogue warps, D = [128 x h3]
//            re-using the drained low columns of acc1
//   layer 3 (K = h3): B = W3 resident in shared memory, D = [128 x 16]
//   scheduler step on x_t: fp32 registers of the thread that owns the row; x_t never leaves the SM.
// Activations only ever exist as 16 KB chunks in a 4-slot ring; weights never leave L2/SMEM.
#include "actor_layout.cuh"
#include "tc_common.cuh"
#include "scratch_cache.cuh"

namespace ddp {
using namespace tc;

namespace {

constexpr int kRows = 128;                  // rows per tile == UMMA M
constexpr int kChunkBytes = kRows * 128;    // one A chunk: 128 rows x 64 bf16
constexpr int kStageBytes = 256 * 128;      // one weight stage: up to 256 rows x 64 bf16
#ifndef DDP_TC_STAGES
#define DDP_TC_STAGES 4
#endif
constexpr int kStages = DDP_TC_STAGES;       // 32 KB weight stages; stages + A slots / 2 = 6 fills shared memory
constexpr int kASlots = 2 * (6 - kStages);   // 16 KB A chunks
#ifndef DDP_TC_H2
#define DDP_TC_H2 1   // 1: fp16 operands in every step and, after the first step, the packed-fp16 Mish in the TMEM drains;
#endif                // 0: round-1 plan (fp16 first step, bf16 operands after it, fp32 Mish everywhere)
constexpr bool kH2 = DDP_TC_H2 != 0;
#ifndef DDP_TC_PAIR_PUBLISH
#define DDP_TC_PAIR_PUBLISH 0   // 1: two chunks per proxy fence (measured slower here: it delays the layer-1 MMAs)
#endif
constexpr bool kPairPublish = DDP_TC_PAIR_PUBLISH != 0;
#ifndef DDP_TC_KEEP_PARTIAL
#define DDP_TC_KEEP_PARTIAL 0   // 1: evict_last L2 policy on the state-partial scratch (scratch_cache.cuh)
#endif
constexpr bool kKeepPartial = DDP_TC_KEEP_PARTIAL != 0;
constexpr int kEpiWarps = 8;                 // two per SM sub-partition (16 spill under the 96-register cap: 1.28 ms)
constexpr int kColsPerWarp = 32;             // columns of a 64-column chunk owned by one warp
constexpr int kNT = kColsPerWarp / 8;        // mma.sync n8 tiles per warp in layer 0
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = kEpiThreads + 64;  // + TMA warp + MMA warp
constexpr int kIn0Stride = 56;              // bf16 per row of the layer-0 input tile (112 B: conflict-free)
constexpr int kK0 = 48;                     // layer-0 contraction: x(8) | state(<=34) | zero pad
constexpr int kTmemCols = 512;

struct TcArgs {
    const uint2* w0frag;       // layer-0 B fragments in mma.sync order (bf16)
    const uint2* w0frag_h;     // same, fp16 (first denoising step)
    const uint4* w3img;        // layer-3 B tiles, pre-swizzled shared-memory image (bf16)
    const uint4* w3img_h;      // same, fp16
    int first_f16;             // 1: step t = T-1 runs with fp16 operands (numerics)
    const float *tb0, *b1, *b2, *b3, *cst;
    const float *state, *noise;
    float* out;
    uint4* pscr;               // per-CTA scratch of the state partial sums: [h1/64][8 warps][4][32 lanes] x 16 B
    long B;
    int S, A, T, h1, h2, h3;
    int nparts1, part1;        // layer-1 output split into nparts1 parts of part1 (<= 256) columns
    int num_tiles;             // tiles of this launch (128 rows each, or 64 in the M = 64 kernel)
    long row0;                 // first row of this launch
    int split_steps;           // 1: the T steps of a tile may be shared by two neighbouring CTAs (see Job)
    uint4* pshare;             // [grid] state-partial blocks of the tiles shared by two CTAs (written by the head's CTA)
    float* handoff;            // [grid][128][8] fp32: x_t of the tile a CTA leaves unfinished, for its right neighbour
    int* handoff_flag;         // [grid] zeroed before the launch; 1 once handoff[b] is complete
    ExplNoise expl;            // optional exploration / target-policy noise epilogue
    long long* dbg;            // optional per-phase cycle counters of CTA 0 (development aid), else NULL
};

// Shared-memory map (bytes from the 1024-aligned base); sized for the largest supported shape
// (h2 <= 512, h3 <= 256) so that every offset is a compile-time constant.
struct SM {
    static constexpr uint32_t wring = 0;                                   // kStages x 32 KB weight stages
    static constexpr uint32_t aring = wring + kStages * kStageBytes;       // kASlots x 16 KB A chunks
    static constexpr uint32_t w3 = aring + kASlots * kChunkBytes;          // W3 tiles: bf16 image, fp16 image
    static constexpr uint32_t w3_f16 = w3 + 4 * 2048;
    static constexpr uint32_t b1h = w3;                                    // kH2: fp16 copies of b1 / b2 take the place of the
    static constexpr uint32_t b2h = w3 + 512 * 2;                          //      bf16 W3 image (every step runs on fp16 operands)
    static constexpr uint32_t in0 = w3 + 2 * 4 * 2048;                     // layer-0 input tile [128][56] 16-bit
    static constexpr uint32_t b1 = in0 + kRows * kIn0Stride * 2;           // fp32 biases
    static constexpr uint32_t b2 = b1 + 512 * 4;
    static constexpr uint32_t bars = b2 + 256 * 4;
    static constexpr uint32_t tmem_ptr = bars + 8 * (2 * kStages + 2 * kASlots + 2);
    static constexpr uint32_t total = tmem_ptr + 8;
};

// barrier indices inside the bars block
__device__ __forceinline__ uint32_t bar_w_full(uint32_t base, int i) { return base + 8 * i; }
__device__ __forceinline__ uint32_t bar_w_empty(uint32_t base, int i) { return base + 8 * (kStages + i); }
__device__ __forceinline__ uint32_t bar_a_full(uint32_t base, int i) { return base + 8 * (2 * kStages + i); }
__device__ __forceinline__ uint32_t bar_a_empty(uint32_t base, int i) { return base + 8 * (2 * kStages + kASlots + i); }
__device__ __forceinline__ uint32_t bar_acc_full(uint32_t base) { return base + 8 * (2 * kStages + 2 * kASlots); }
__device__ __forceinline__ uint32_t bar_lo_free(uint32_t base) { return base + 8 * (2 * kStages + 2 * kASlots + 1); }

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

struct Ring {
    int idx = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance(int n) { if (++idx == n) { idx = 0; phase ^= 1; } }
};

// Work assignment.  The unit of work is one denoising step of one tile.  Without `split_steps` CTA b owns whole tiles
// b, b + grid, ...  With it (large batches whose tile count is not a multiple of the grid: 512 tiles on 148 SMs are
// 3.46 waves that cost 4) the num_tiles * T units are dealt out in contiguous, balanced ranges of the tile-major order,
// so a CTA's range may begin and / or end inside a tile.  Such a tile is shared with the neighbouring CTA through
// x_t (128 x 8 floats in global memory, one release / acquire flag): the CTA runs the unfinished HEAD of its last tile
// FIRST, then its whole tiles, and the TAIL of its first tile LAST -- by then the left neighbour, which started with
// exactly that head, has long published it, so nobody ever waits.  The second CTA recomputes the per-tile state partial
// of layer 0 (~0.1 step).  The host only enables this when every range spans at least two tiles (no piece is both).
struct Job { int tile, j0, j1; bool head, tail; };      // steps j0..j1; head: ends before T-1 (publishes x_t); tail: starts after 0
struct Schedule {
    int b, grid, num_tiles, T, split;
    int f, js, l, je, full0, full1, njobs;
    __device__ __forceinline__ Schedule(const TcArgs& a) {
        b = blockIdx.x; grid = gridDim.x; num_tiles = a.num_tiles; T = a.T; split = a.split_steps;
        if (split) {
            const long units = (long)num_tiles * T, base = units / grid, rem = units % grid;
            const long lo = b * base + (b < rem ? b : rem), hi = lo + base + (b < rem ? 1 : 0);
            f = (int)(lo / T); js = (int)(lo % T); l = (int)((hi - 1) / T); je = (int)((hi - 1) % T);
            full0 = js > 0 ? f + 1 : f;
            full1 = je < T - 1 ? l - 1 : l;                 // inclusive
            njobs = (je < T - 1 ? 1 : 0) + (full1 - full0 + 1) + (js > 0 ? 1 : 0);
        } else {
            njobs = b < num_tiles ? (num_tiles - b + grid - 1) / grid : 0;
        }
    }
    __device__ __forceinline__ Job job(int i) const {
        if (!split) return Job{b + i * grid, 0, T - 1, false, false};
        const int has_head = je < T - 1 ? 1 : 0;
        if (has_head && i == 0) return Job{l, 0, je, true, false};
        const int k = i - has_head, nfull = full1 - full0 + 1;
        if (k < nfull) return Job{full0 + k, 0, T - 1, false, false};
        return Job{f, js, T - 1, false, true};
    }
};

// Per-thread state of an epilogue / layer-0 warp.
struct EpiCtx {
    uint8_t* smem;
    uint32_t bars, tmem_base, acc_phase;
    int q, ch, g, t4, lane, my_row;     // TMEM lane quarter, column group of the chunk, mma.sync coords, owned row
    int NC1, NC2, NC3;
    Ring as;
    uint4* pscr_base;                   // state-partial block of the current job (private, or shared with the neighbour)
    long row;
    bool valid;
    bool owner;                         // this thread owns a row of the tile (64-row tiles: lanes 0..15 only)
    float xr[8];                        // x_t of the owned row (ch == 0 threads), fp32
};

// Tile prologue form of write_in0_row: the whole [x (8) | state (S <= 40) | 0] row.  All loads of the row are issued
// before the first conversion (8-byte loads when S is even): the row costs one memory round trip, not one per element
// (the plain loop below, with its run-time trip count, measured 4.5 k clk per tile).
template <bool F16>
__device__ __forceinline__ void write_in0_row_full(const TcArgs& a, const EpiCtx& e) {
    if (!e.owner) return;
    constexpr int NS = kK0 - 8;
    float sv[NS];
    const float* sp = a.state + e.row * a.S;
    if ((a.S & 1) == 0) {
#pragma unroll
        for (int i = 0; i < NS; i += 2) {
            const float2 v = (e.valid && i + 1 < a.S) ? __ldg(reinterpret_cast<const float2*>(sp + i)) : make_float2(0.f, 0.f);
            sv[i] = v.x; sv[i + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < NS; ++i) sv[i] = (e.valid && i < a.S) ? __ldg(sp + i) : 0.f;
    }
    uint4* rp = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(e.smem + SM::in0) + e.my_row * kIn0Stride);
    rp[0] = make_uint4(pack2<F16>(e.xr[0], e.xr[1]), pack2<F16>(e.xr[2], e.xr[3]), pack2<F16>(e.xr[4], e.xr[5]), pack2<F16>(e.xr[6], e.xr[7]));
#pragma unroll
    for (int g = 0; g < NS / 8; ++g)
        rp[1 + g] = make_uint4(pack2<F16>(sv[8 * g], sv[8 * g + 1]), pack2<F16>(sv[8 * g + 2], sv[8 * g + 3]),
                               pack2<F16>(sv[8 * g + 4], sv[8 * g + 5]), pack2<F16>(sv[8 * g + 6], sv[8 * g + 7]));
}

// [x (8) | state (S) | 0 ...] of the owned row into the layer-0 input tile, in the operand format F16/bf16
template <bool F16>
__device__ __forceinline__ void write_in0_row(const TcArgs& a, const EpiCtx& e, int nstate) {
    if (!e.owner) return;
    uint16_t* rp = reinterpret_cast<uint16_t*>(e.smem + SM::in0) + e.my_row * kIn0Stride;
    uint4 xv;
    xv.x = pack2<F16>(e.xr[0], e.xr[1]); xv.y = pack2<F16>(e.xr[2], e.xr[3]);
    xv.z = pack2<F16>(e.xr[4], e.xr[5]); xv.w = pack2<F16>(e.xr[6], e.xr[7]);
    *reinterpret_cast<uint4*>(rp) = xv;
    for (int i = 0; i < nstate; ++i)
        rp[8 + i] = cvt16<F16>((e.valid && i < a.S) ? a.state[e.row * a.S + i] : 0.f);
}

// 16 accumulator columns (already in registers) of this thread's row -> +bias, Mish, 16-bit -> two
// 16-byte pieces of the A chunk slot (columns col0 .. col0+15 of the 64-column chunk)
template <bool F16, bool H2>
__device__ __forceinline__ void emit_half(const EpiCtx& e, uint8_t* slot, const uint32_t (&v)[16], const float* bb,
                                          const __half* bh, int col0) {
    if constexpr (H2) {
        // packed-fp16 path: round the accumulator pair to fp16, add the fp16 bias pair, Mish in HFMA2 arithmetic
        const uint4 b0 = *reinterpret_cast<const uint4*>(bh), b1 = *reinterpret_cast<const uint4*>(bh + 8);
        const uint32_t bp[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const __half2 x = __hadd2(u32_as_h2(pack_f16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]))), u32_as_h2(bp[i]));
            w[i] = mish_h2(h2_as_u32(x));
        }
        *reinterpret_cast<uint4*>(slot + sw128_offset(e.my_row, col0)) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(slot + sw128_offset(e.my_row, col0 + 8)) = make_uint4(w[4], w[5], w[6], w[7]);
    } else {
    float x[16];
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) {
        const float4 b = *reinterpret_cast<const float4*>(bb + i4 * 4);
        x[i4 * 4 + 0] = __uint_as_float(v[i4 * 4 + 0]) + b.x; x[i4 * 4 + 1] = __uint_as_float(v[i4 * 4 + 1]) + b.y;
        x[i4 * 4 + 2] = __uint_as_float(v[i4 * 4 + 2]) + b.z; x[i4 * 4 + 3] = __uint_as_float(v[i4 * 4 + 3]) + b.w;
    }
    mish_fast_n<16>(x);
#pragma unroll
    for (int i8 = 0; i8 < 2; ++i8) {
        uint4 w;
        w.x = pack2<F16>(x[i8 * 8 + 0], x[i8 * 8 + 1]); w.y = pack2<F16>(x[i8 * 8 + 2], x[i8 * 8 + 3]);
        w.z = pack2<F16>(x[i8 * 8 + 4], x[i8 * 8 + 5]); w.w = pack2<F16>(x[i8 * 8 + 6], x[i8 * 8 + 7]);
        *reinterpret_cast<uint4*>(slot + sw128_offset(e.my_row, col0 + i8 * 8)) = w;
    }
    }
}

// Drain `nchunks` 64-column chunks of the accumulator at TMEM column 0 into the A ring.  This warp owns
// 32 columns of each chunk, moved as 16-column TMEM loads: while one piece goes through Mish the
// next load is in flight (two 16-register buffers).  After chunk `signal_after` the thread arrives on
// lo_free (-1: never).
template <bool F16, bool H2>
__device__ __forceinline__ void drain_acc(EpiCtx& e, int nchunks, const float* bias, const __half* biash, int signal_after) {
    const uint32_t tbase = e.tmem_base + ((uint32_t)(e.q * 32) << 16) + e.ch * kColsPerWarp;
    uint32_t va[16], vb[16];
    int pending = -1;                  // ring slot written but not yet published
    tmem_ld16(tbase, va);
    for (int c = 0; c < nchunks; ++c) {
        const float* bb = bias + c * 64 + e.ch * kColsPerWarp;
        const __half* bh = biash + c * 64 + e.ch * kColsPerWarp;
        tmem_ld_wait();
        tmem_ld16(tbase + c * 64 + 16, vb);
        mbar_wait(bar_a_empty(e.bars, e.as.idx), e.as.phase ^ 1);
        uint8_t* slot = e.smem + SM::aring + e.as.idx * kChunkBytes;
        emit_half<F16, H2>(e, slot, va, bb, bh, e.ch * 32);
        tmem_ld_wait();
        if (c + 1 < nchunks) tmem_ld16(tbase + (c + 1) * 64, va);
        emit_half<F16, H2>(e, slot, vb, bb + 16, bh + 16, e.ch * 32 + 16);
        // every lane publishes its own writes to the async proxy, then one lane arrives for the warp
        // (32 lanes arriving on one mbarrier word serialise in the shared-memory pipe).  Two chunks share one
        // publication: the proxy fence is the expensive part of handing a chunk over.
        const bool flush = !kPairPublish || (c & 1) || c + 1 == nchunks;
        if (flush) {
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (e.lane == 0) {
                if (pending >= 0) mbar_arrive(bar_a_full(e.bars, pending));
                mbar_arrive(bar_a_full(e.bars, e.as.idx));
                if (c == signal_after || (pending >= 0 && c - 1 == signal_after)) mbar_arrive(bar_lo_free(e.bars));
            }
            pending = -1;
        } else {
            pending = e.as.idx;
        }
        e.as.advance(kASlots);
    }
}

// 64-row tiles (tcgen05.mma M = 64): D[r][n] sits in lane (r % 16) + 32 (r / 16), column n (tools/m64_probe.py), i.e.
// each lane quarter holds 16 rows.  tcgen05.ld.16x256b hands them to the warp as mma.m16n8 C fragments -- thread t:
// rows t/4 and t/4 + 8 of the quarter, columns 2 (t % 4), +1 of every 8-column block -- so all 32 lanes work and the
// warp's share of a chunk (16 rows x 32 columns) is half of what it is in a 128-row tile.
__device__ __forceinline__ void tmem_ld_16x256b(uint32_t taddr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
}
template <bool F16, bool H2>
__device__ __forceinline__ void drain_acc_m64(EpiCtx& e, int nchunks, const float* bias, const __half* biash, int lo_chunks) {
    const uint32_t tbase = e.tmem_base + ((uint32_t)(e.q * 32) << 16) + e.ch * kColsPerWarp;
    uint32_t va[kNT][4], vb[kNT][4];
    auto fetch = [&](int c, uint32_t (&v)[kNT][4]) {
#pragma unroll
        for (int nt = 0; nt < kNT; ++nt) tmem_ld_16x256b(tbase + c * 64 + nt * 8, v[nt]);
    };
    auto chunk = [&](int c, const uint32_t (&v)[kNT][4]) {
        uint32_t w[kNT][2];
        if constexpr (H2) {
#pragma unroll
            for (int nt = 0; nt < kNT; ++nt) {
                const __half2 b = *reinterpret_cast<const __half2*>(biash + c * 64 + e.ch * kColsPerWarp + nt * 8 + 2 * e.t4);
                w[nt][0] = mish_h2(h2_as_u32(__hadd2(u32_as_h2(pack_f16x2(__uint_as_float(v[nt][0]), __uint_as_float(v[nt][1]))), b)));
                w[nt][1] = mish_h2(h2_as_u32(__hadd2(u32_as_h2(pack_f16x2(__uint_as_float(v[nt][2]), __uint_as_float(v[nt][3]))), b)));
            }
        } else {
            float x[kNT * 4];
#pragma unroll
            for (int nt = 0; nt < kNT; ++nt) {
                const float2 b = *reinterpret_cast<const float2*>(bias + c * 64 + e.ch * kColsPerWarp + nt * 8 + 2 * e.t4);
                x[nt * 4 + 0] = __uint_as_float(v[nt][0]) + b.x; x[nt * 4 + 1] = __uint_as_float(v[nt][1]) + b.y;
                x[nt * 4 + 2] = __uint_as_float(v[nt][2]) + b.x; x[nt * 4 + 3] = __uint_as_float(v[nt][3]) + b.y;
            }
            mish_fast_n<kNT * 4>(x);
#pragma unroll
            for (int nt = 0; nt < kNT; ++nt) { w[nt][0] = pack2<F16>(x[nt * 4 + 0], x[nt * 4 + 1]); w[nt][1] = pack2<F16>(x[nt * 4 + 2], x[nt * 4 + 3]); }
        }
        mbar_wait(bar_a_empty(e.bars, e.as.idx), e.as.phase ^ 1);
        uint8_t* slot = e.smem + SM::aring + e.as.idx * kChunkBytes;
#pragma unroll
        for (int nt = 0; nt < kNT; ++nt) {
            const int r0 = e.q * 16 + e.g, col = e.ch * kColsPerWarp + nt * 8 + 2 * e.t4;
            *reinterpret_cast<uint32_t*>(slot + sw128_offset(r0, col)) = w[nt][0];
            *reinterpret_cast<uint32_t*>(slot + sw128_offset(r0 + 8, col)) = w[nt][1];
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (e.lane == 0) {
            mbar_arrive(bar_a_full(e.bars, e.as.idx));
            if (c == lo_chunks) mbar_arrive(bar_lo_free(e.bars));
        }
        e.as.advance(kASlots);
    };
    fetch(0, va);
    for (int c = 0; c < nchunks; c += 2) {
        tmem_ld_wait();
        if (c + 1 < nchunks) fetch(c + 1, vb);
        chunk(c, va);
        if (c + 1 < nchunks) {
            tmem_ld_wait();
            if (c + 2 < nchunks) fetch(c + 2, va);
            chunk(c + 1, vb);
        }
    }
}

// Scratch slot of this lane for chunk c: 4 x 16 B, [c][warp][j][lane] so that every access is one coalesced 512 B row.
// The layout depends on (warp, lane) only, so the block a CTA writes for the head of a shared tile serves the same
// threads of the neighbouring CTA that finishes the tile.
__device__ __forceinline__ uint4* pscr_slot(const TcArgs& a, const EpiCtx& e, int c) {
    return e.pscr_base + (size_t)c * (kEpiWarps * 4 * 32) + (e.q + 4 * e.ch) * (4 * 32) + e.lane;
}

// Tile prologue: P = W0[:, state[8:]] . state[8:] for the 32 rows x kColsPerWarp features this warp owns in every
// chunk (k16 steps 1 and 2 of the layer-0 contraction; operands in the first step's format), kept as fp16 in the
// lane's own fragment order.  Written and read back by the same thread: no synchronisation is involved.
static_assert(kEpiWarps == 8 && kNT == 4, "the state-partial fragment order is laid out for 8 warps x 32 columns");
// ldmatrix lane address of the m16 tile(s) this warp contracts: rows q*32 + mt*16 (128-row tile) or q*16 (64-row tile)
template <bool HM>
__device__ __forceinline__ uint32_t in0_lane_addr(const EpiCtx& e) {
    const int r = (HM ? e.q * 16 : e.q * 32) + (e.lane & 7) + ((e.lane >> 3) & 1) * 8;
    return smem_u32(e.smem + SM::in0) + (uint32_t)((r * kIn0Stride + (e.lane >> 4) * 8) * 2);
}
template <bool F16, bool HM>
__device__ __forceinline__ void state_partial(const TcArgs& a, const EpiCtx& e) {
    constexpr int MT = HM ? 1 : 2;
    const uint32_t in0_lane = in0_lane_addr<HM>(e);
    const uint2* wf = F16 ? a.w0frag_h : a.w0frag;
    uint32_t af[MT][2][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
            ldmatrix_x4(af[mt][ks], in0_lane + (uint32_t)((mt * 16 * kIn0Stride + (ks + 1) * 16) * 2));
    uint2 bfr[kNT][2], nxt[kNT][2];
    auto load_b = [&](int c, uint2 (&f)[kNT][2]) {
        const uint2* bf = wf + ((size_t)(c * 8 + e.ch * kNT) * 3) * 32 + e.lane;
#pragma unroll
        for (int nt = 0; nt < kNT; ++nt)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) f[nt][ks] = __ldg(bf + (nt * 3 + ks + 1) * 32);
    };
    load_b(0, bfr);
#pragma unroll 2
    for (int c = 0; c < e.NC1; ++c) {
        load_b(min(c + 1, e.NC1 - 1), nxt);       // next chunk's fragments in flight during this chunk's HMMAs
        uint4* dst = pscr_slot(a, e, c);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            uint32_t h[kNT][2];
#pragma unroll
            for (int nt = 0; nt < kNT; ++nt) {
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) mma_m16n8k16<F16>(acc, af[mt][ks], bfr[nt][ks].x, bfr[nt][ks].y);
                h[nt][0] = pack_f16x2(acc[0], acc[1]);
                h[nt][1] = pack_f16x2(acc[2], acc[3]);
            }
            scratch_st<kKeepPartial>(dst + (mt * 2 + 0) * 32, make_uint4(h[0][0], h[0][1], h[1][0], h[1][1]));
            scratch_st<kKeepPartial>(dst + (mt * 2 + 1) * 32, make_uint4(h[2][0], h[2][1], h[3][0], h[3][1]));
        }
#pragma unroll
        for (int nt = 0; nt < kNT; ++nt) { bfr[nt][0] = nxt[nt][0]; bfr[nt][1] = nxt[nt][1]; }
    }
}

// Layer-0 operands of one chunk that do not depend on x_t: W0 fragments of the [x_t | state[0:8]] k16 step, the
// time-table pair of each n8 tile and the parked state partial.  Chunk 0 of a step is fetched BEFORE the previous step's
// head (nothing in it depends on the new x_t), so its L2 latency hides behind the layer-3 MMA wait and the scheduler math.
struct L0Frags {
    uint2 bfr[kNT];
    float2 bias[kNT];
    uint4 pp[4];
};
template <bool F16, bool HM>
__device__ __forceinline__ void load_l0_frags(const TcArgs& a, const EpiCtx& e, int t, int c, L0Frags& fr) {
    constexpr int MT = HM ? 1 : 2;
    // packed fragment order: [chunk][32-feature half][n8 tile 0..3][k16 step][lane]; this warp's first feature inside
    // the chunk is ch * kColsPerWarp
    const float* tb = a.tb0 + (size_t)t * a.h1;
    const uint2* wf = F16 ? a.w0frag_h : a.w0frag;
    const int n8 = e.ch * kNT;                              // first n8 tile (0..7) of this warp in the chunk
    const uint2* bf = wf + ((size_t)(c * 8 + n8) * 3) * 32 + e.lane;
#pragma unroll
    for (int nt = 0; nt < kNT; ++nt) fr.bfr[nt] = __ldg(bf + (nt * 3) * 32);
#pragma unroll
    for (int nt = 0; nt < kNT; ++nt)
        fr.bias[nt] = __ldg(reinterpret_cast<const float2*>(tb + c * 64 + e.ch * kColsPerWarp + nt * 8 + 2 * e.t4));
    const uint4* ps = pscr_slot(a, e, c);
#pragma unroll
    for (int j = 0; j < 2 * MT; ++j) fr.pp[j] = scratch_ld<kKeepPartial>(ps + j * 32);
}

// One denoising step of the epilogue warps (operand format of THIS step = F16 ? fp16 : bf16; H2: the TMEM drains run
// the packed-fp16 Mish -- every step but the first, whose eps_hat error is amplified by 1/sqrt(abar_{T-1})).
// `fr` arrives holding chunk 0 of this step and leaves holding chunk 0 of the next one (`has_next`).
template <bool F16, bool HM, bool H2 = false, bool H2_ACC2 = H2>
__device__ __forceinline__ void epi_step(const TcArgs& a, EpiCtx& e, int j, L0Frags& fr, bool has_next) {
    constexpr int MT = HM ? 1 : 2;
    const int t = a.T - 1 - j;
#ifdef DDP_TC_PROFILE      // per-phase cycle counters of CTA 0 (tools/tc_timing.py builds this variant): registers the product does not pay for
    const bool prof = a.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    long long tk0 = prof ? clock64() : 0, tk1;
#define DDP_TICK(slot) do { if (prof) { tk1 = clock64(); a.dbg[slot] += tk1 - tk0; tk0 = tk1; } } while (0)
#else
#define DDP_TICK(slot) do { } while (0)
#endif
    // ---- layer 0: one 64-feature chunk at a time, straight into the A ring.  accumulator = state partial of the
    // tile (scratch, fp16) + time-table row + one k16 step over [x_t | state[0:8]]
    const uint32_t in0_lane = in0_lane_addr<HM>(e);
    uint32_t afx[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) ldmatrix_x4(afx[mt], in0_lane + (uint32_t)((mt * 16 * kIn0Stride) * 2));
    uint2 (&bfr)[kNT] = fr.bfr;
    float2 (&bias)[kNT] = fr.bias;
    uint4 (&pp)[4] = fr.pp;
    auto load_frags = [&](int c) { load_l0_frags<F16, HM>(a, e, t, c, fr); };
    int pending0 = -1;
    for (int c = 0; c < e.NC1; ++c) {
        float acc[MT][kNT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
            for (int nt = 0; nt < kNT; ++nt) {
                const uint4 pv = pp[mt * 2 + (nt >> 1)];
                const uint32_t h0 = (nt & 1) ? pv.z : pv.x, h1 = (nt & 1) ? pv.w : pv.y;
                const float2 p01 = __half22float2(*reinterpret_cast<const __half2*>(&h0));
                const float2 p23 = __half22float2(*reinterpret_cast<const __half2*>(&h1));
                acc[mt][nt][0] = p01.x + bias[nt].x; acc[mt][nt][1] = p01.y + bias[nt].y;
                acc[mt][nt][2] = p23.x + bias[nt].x; acc[mt][nt][3] = p23.y + bias[nt].y;
            }
#pragma unroll
#ifndef DDP_EXP_NO_HMMA
            for (int nt = 0; nt < kNT; ++nt) mma_m16n8k16<F16>(acc[mt][nt], afx[mt], bfr[nt].x, bfr[nt].y);
#else
            for (int nt = 0; nt < kNT; ++nt) { acc[mt][nt][0] += __uint_as_float(afx[mt][0] ^ bfr[nt].x) * 1e-30f; acc[mt][nt][3] += __uint_as_float(afx[mt][3] ^ bfr[nt].y) * 1e-30f; }
#endif
        }
        // the fragment registers are dead after the HMMAs: refill them for the next chunk now, so the
        // loads are in flight during the Mish / store phase
        if (c + 1 < e.NC1) load_frags(c + 1);
        mbar_wait(bar_a_empty(e.bars, e.as.idx), e.as.phase ^ 1);
        uint8_t* slot = e.smem + SM::aring + e.as.idx * kChunkBytes;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#ifndef DDP_EXP_NO_L0_MISH
            mish_fast_n<kNT * 4>(reinterpret_cast<float(&)[kNT * 4]>(acc[mt]));
#endif
#pragma unroll
            for (int nt = 0; nt < kNT; ++nt) {
                const int r0 = (HM ? e.q * 16 : e.q * 32 + mt * 16) + e.g, col = e.ch * kColsPerWarp + nt * 8 + 2 * e.t4;
                *reinterpret_cast<uint32_t*>(slot + sw128_offset(r0, col)) = pack2<F16>(acc[mt][nt][0], acc[mt][nt][1]);
                *reinterpret_cast<uint32_t*>(slot + sw128_offset(r0 + 8, col)) = pack2<F16>(acc[mt][nt][2], acc[mt][nt][3]);
            }
        }
        if (!kPairPublish || (c & 1) || c + 1 == e.NC1) {
            fence_proxy_async();
            __syncwarp();
            if (e.lane == 0) {
                if (pending0 >= 0) mbar_arrive(bar_a_full(e.bars, pending0));
                mbar_arrive(bar_a_full(e.bars, e.as.idx));
            }
            pending0 = -1;
        } else {
            pending0 = e.as.idx;
        }
        e.as.advance(kASlots);
    }

    DDP_TICK(0);       // layer 0 + Mish, 16 chunks
    // ---- layer-1 epilogue: acc1 (TMEM cols [0,h2)) -> +b1, Mish -> A chunks of layer 2
    const float* sb1 = reinterpret_cast<const float*>(e.smem + SM::b1);
    const float* sb2 = reinterpret_cast<const float*>(e.smem + SM::b2);
    const __half* sb1h = reinterpret_cast<const __half*>(e.smem + SM::b1h);
    const __half* sb2h = reinterpret_cast<const __half*>(e.smem + SM::b2h);
    mbar_wait(bar_acc_full(e.bars), e.acc_phase); e.acc_phase ^= 1;
    tc_fence_after();
    DDP_TICK(1);       // wait for the last layer-1 MMA
    if constexpr (HM) drain_acc_m64<F16, H2>(e, e.NC2, sb1, sb1h, e.NC3 - 1);
    else drain_acc<F16, H2>(e, e.NC2, sb1, sb1h, e.NC3 - 1);     // lo_free: TMEM cols [0, h3) are drained
    DDP_TICK(2);       // drain acc1
    // ---- layer-2 epilogue: acc2 (TMEM cols [0,h3)) -> +b2, Mish -> A chunks of layer 3
    mbar_wait(bar_acc_full(e.bars), e.acc_phase); e.acc_phase ^= 1;
    tc_fence_after();
    DDP_TICK(3);       // wait for the last layer-2 MMA
    if constexpr (HM) drain_acc_m64<F16, H2_ACC2>(e, e.NC3, sb2, sb2h, -1);
    else drain_acc<F16, H2_ACC2>(e, e.NC3, sb2, sb2h, -1);
    DDP_TICK(4);       // drain acc2

    // ---- head + scheduler step: eps_hat from acc3, x_t in registers (fp32).  Everything the head and the next step's
    // first chunk need from global memory is requested before the wait for the layer-3 MMA: the step noise and head bias
    // of the owned row, and chunk 0 of the next step's layer-0 operands (kernel-format next step: F16 again under kH2)
    float zr[8], b3r[8];
    if (e.ch == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            zr[i] = (e.valid && t > 0 && i < a.A) ? __ldg(a.noise + ((size_t)(j + 1) * a.B + e.row) * a.A + i) : 0.f;
            b3r[i] = i < a.A ? __ldg(a.b3 + i) : 0.f;
        }
    }
    if (has_next) load_l0_frags<kH2 ? true : false, HM>(a, e, t - 1, 0, fr);
    mbar_wait(bar_acc_full(e.bars), e.acc_phase); e.acc_phase ^= 1;
    tc_fence_after();
    DDP_TICK(5);       // wait for layer 3
    if (e.ch == 0) {
        uint32_t ev[8];
        tmem_ld8(e.tmem_base + ((uint32_t)(e.q * 32) << 16) + a.h3, ev);
        tmem_ld_wait();
        const float* cs = a.cst + t * kCstStride;
        const float c_eps = cs[CST_CEPS], s_ab = cs[CST_SQRT_AB], c_x0 = cs[CST_CX0], c_xt = cs[CST_CXT],
                    sigma = cs[CST_SIGMA];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float eps = __uint_as_float(ev[i]) + b3r[i];
            float x0 = __fdiv_rn(__fsub_rn(e.xr[i], __fmul_rn(c_eps, eps)), s_ab);
            x0 = fminf(fmaxf(x0, -1.f), 1.f);
            float xn = __fadd_rn(__fmul_rn(c_x0, x0), __fmul_rn(c_xt, e.xr[i]));
            if (t > 0) xn = __fadd_rn(xn, __fmul_rn(sigma, zr[i]));
            e.xr[i] = i < a.A ? xn : 0.f;
        }
        if (t > 0) {
            // kH2: every step runs on fp16 operands.  Round-1 plan: the next step runs in bf16; after the fp16 step the
            // state columns are re-written as bf16 too
            if constexpr (kH2) write_in0_row<true>(a, e, 0);
            else write_in0_row<false>(a, e, F16 ? 8 : 0);
        } else if (e.valid) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i < a.A) a.out[e.row * a.A + i] = apply_expl_noise(a.expl, e.xr[i], e.row, a.B, a.A, i);
        }
    }
    tc_fence_before();
    epi_bar_sync();         // new x visible to all layer-0 warps; acc3 reads are complete
    DDP_TICK(6);       // head + scheduler step + barrier
#undef DDP_TICK
}

// HM = false: 128-row tiles.  HM = true: 64-row tiles (tcgen05.mma M = 64) for small batches -- a separate instantiation,
// so neither pays for the other's registers or code size (one kernel with both paths inlined ran 20 % slower).
template <bool HM>
__global__ void __launch_bounds__(kThreads, 1)
actor_sample_tc_kernel(const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2,
                       const __grid_constant__ CUtensorMap map_w1h, const __grid_constant__ CUtensorMap map_w2h,
                       const TcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // SWIZZLE_128B tiles need a 1024-byte aligned base
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    const uint32_t bars = base + SM::bars;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NC1 = a.h1 >> 6, NC2 = a.h2 >> 6, NC3 = a.h3 >> 6;

    // ------------------------------------------------------------------ one-time setup
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(bar_w_full(bars, i), 1); mbar_init(bar_w_empty(bars, i), 1); }
        for (int i = 0; i < kASlots; ++i) { mbar_init(bar_a_full(bars, i), kEpiWarps); mbar_init(bar_a_empty(bars, i), 1); }
        mbar_init(bar_acc_full(bars), 1);
        mbar_init(bar_lo_free(bars), kEpiWarps);
        fence_barrier_init();
    }
    if (warp == kEpiWarps + 1) tmem_alloc(base + SM::tmem_ptr, kTmemCols);
    // resident operands: W3 tiles (pre-swizzled image), biases
    {
        const int n16 = NC3 * 2048 / 16;
        uint4* dst = reinterpret_cast<uint4*>(smem + SM::w3);
        for (int i = threadIdx.x; i < n16; i += kThreads) {
            if (!kH2) dst[i] = a.w3img[i];
            dst[(SM::w3_f16 - SM::w3) / 16 + i] = a.w3img_h[i];
        }
        float* sb1 = reinterpret_cast<float*>(smem + SM::b1);
        float* sb2 = reinterpret_cast<float*>(smem + SM::b2);
        __half* sb1h = reinterpret_cast<__half*>(smem + SM::b1h);
        __half* sb2h = reinterpret_cast<__half*>(smem + SM::b2h);
        for (int i = threadIdx.x; i < a.h2; i += kThreads) { sb1[i] = a.b1[i]; if (kH2) sb1h[i] = __float2half_rn(a.b1[i]); }
        for (int i = threadIdx.x; i < a.h3; i += kThreads) { sb2[i] = a.b2[i]; if (kH2) sb2h[i] = __float2half_rn(a.b2[i]); }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + SM::tmem_ptr);
    if (warp == kEpiWarps) {
        // ============================================================== TMA producer (one lane)
        if (lane == 0) {
            tma_prefetch_desc(&map_w1);
            tma_prefetch_desc(&map_w2);
            tma_prefetch_desc(&map_w1h);
            tma_prefetch_desc(&map_w2h);
            Ring ws;
            const Schedule sch(a);
            for (int ji = 0; ji < sch.njobs; ++ji) {
                const Job jb = sch.job(ji);
                for (int j = jb.j0; j <= jb.j1; ++j) {
                    const bool f16 = kH2 || (a.first_f16 && j == 0);
                    const CUtensorMap* m1 = f16 ? &map_w1h : &map_w1;
                    const CUtensorMap* m2 = f16 ? &map_w2h : &map_w2;
                    for (int c = 0; c < NC1; ++c)
                        for (int p = 0; p < a.nparts1; ++p) {
                            mbar_wait(bar_w_empty(bars, ws.idx), ws.phase ^ 1);
                            mbar_expect_tx(bar_w_full(bars, ws.idx), (uint32_t)a.part1 * 128u);
                            tma_load_2d(base + SM::wring + ws.idx * kStageBytes, m1, bar_w_full(bars, ws.idx),
                                        c * 64, p * a.part1);
                            ws.advance(kStages);
                        }
                    for (int c = 0; c < NC2; ++c) {
                        mbar_wait(bar_w_empty(bars, ws.idx), ws.phase ^ 1);
                        mbar_expect_tx(bar_w_full(bars, ws.idx), (uint32_t)a.h3 * 128u);
                        tma_load_2d(base + SM::wring + ws.idx * kStageBytes, m2, bar_w_full(bars, ws.idx), c * 64, 0);
                        ws.advance(kStages);
                    }
                }
            }
        }
    } else if (warp == kEpiWarps + 1) {
        // ============================================================== MMA issuer (one lane)
        if (lane == 0) {
            Ring ws, as;
            uint32_t lo_phase = 0;
            const Schedule sch(a);
            for (int ji = 0; ji < sch.njobs; ++ji) {
                const Job jb = sch.job(ji);
                for (int j = jb.j0; j <= jb.j1; ++j) {
                    const bool f16 = kH2 || (a.first_f16 && j == 0);
                    constexpr int M = HM ? 64 : kRows;
                    const uint32_t idesc1 = make_idesc_16(M, a.part1, f16);
                    const uint32_t idesc2 = make_idesc_16(M, a.h3, f16);
                    const uint32_t idesc3 = make_idesc_16(M, 16, f16);
                    const uint32_t w3base = base + (f16 ? SM::w3_f16 : SM::w3);
                    // ---- layer 1: acc1[128 x h2] (TMEM cols [0, h2)) += h0 chunk . W1 chunk^T
                    for (int c = 0; c < NC1; ++c) {
                        mbar_wait(bar_a_full(bars, as.idx), as.phase);
                        tc_fence_after();
                        const uint64_t adesc = make_smem_desc_sw128(base + SM::aring + as.idx * kChunkBytes);
                        for (int p = 0; p < a.nparts1; ++p) {
                            mbar_wait(bar_w_full(bars, ws.idx), ws.phase);
                            tc_fence_after();
                            const uint64_t bdesc = make_smem_desc_sw128(base + SM::wring + ws.idx * kStageBytes);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16(tmem_base + p * a.part1, adesc + 2 * k, bdesc + 2 * k, idesc1, (c | k) != 0);
                            umma_commit(bar_w_empty(bars, ws.idx));
                            ws.advance(kStages);
                        }
                        umma_commit(bar_a_empty(bars, as.idx));
                        as.advance(kASlots);
                    }
                    umma_commit(bar_acc_full(bars));
                    // ---- layer 2: acc2[128 x h3] re-uses TMEM cols [0, h3) once the epilogue has drained them
                    mbar_wait(bar_lo_free(bars), lo_phase);
                    lo_phase ^= 1;
                    tc_fence_after();
                    for (int c = 0; c < NC2; ++c) {
                        mbar_wait(bar_a_full(bars, as.idx), as.phase);
                        mbar_wait(bar_w_full(bars, ws.idx), ws.phase);
                        tc_fence_after();
                        const uint64_t adesc = make_smem_desc_sw128(base + SM::aring + as.idx * kChunkBytes);
                        const uint64_t bdesc = make_smem_desc_sw128(base + SM::wring + ws.idx * kStageBytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc2, (c | k) != 0);
                        umma_commit(bar_w_empty(bars, ws.idx));
                        umma_commit(bar_a_empty(bars, as.idx));
                        ws.advance(kStages);
                        as.advance(kASlots);
                    }
                    umma_commit(bar_acc_full(bars));
                    // ---- layer 3: acc3[128 x 16] at TMEM cols [h3, h3+16) (acc1 is fully drained by now)
                    for (int c = 0; c < NC3; ++c) {
                        mbar_wait(bar_a_full(bars, as.idx), as.phase);
                        tc_fence_after();
                        const uint64_t adesc = make_smem_desc_sw128(base + SM::aring + as.idx * kChunkBytes);
                        const uint64_t bdesc = make_smem_desc_sw128(w3base + c * 2048);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(tmem_base + a.h3, adesc + 2 * k, bdesc + 2 * k, idesc3, (c | k) != 0);
                        umma_commit(bar_a_empty(bars, as.idx));
                        as.advance(kASlots);
                    }
                    umma_commit(bar_acc_full(bars));
                }
            }
        }
    } else if (warp < kEpiWarps) {
        // ============================================================== epilogue / layer-0 warps
        EpiCtx e;
        e.smem = smem; e.bars = bars; e.tmem_base = tmem_base;
        e.q = warp & 3; e.ch = warp >> 2; e.g = lane >> 2; e.t4 = lane & 3; e.lane = lane;
        e.NC1 = NC1; e.NC2 = NC2; e.NC3 = NC3;
        e.acc_phase = 0;
        const Schedule sch(a);
#ifdef DDP_TC_PROFILE
        const bool prof = a.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
        long long pc0 = 0, pt0 = 0;
        if (prof) { pc0 = clock64(); asm volatile("mov.u64 %0, %globaltimer;" : "=l"(pt0)); }
#endif
        for (int ji = 0; ji < sch.njobs; ++ji) {
            const Job jb = sch.job(ji);
            const int tile = jb.tile;
            if (!HM) {
                e.my_row = e.q * 32 + lane;
                e.owner = true;
                e.row = a.row0 + (long)tile * kRows + e.my_row;
            } else {
                e.my_row = e.q * 16 + (lane & 15);
                e.owner = lane < 16;
                e.row = a.row0 + (long)tile * 64 + e.my_row;
            }
            e.valid = e.owner && e.row < a.B;
            {
                const size_t per_cta = (size_t)e.NC1 * kEpiWarps * 4 * 32;
                e.pscr_base = jb.head ? a.pshare + (size_t)blockIdx.x * per_cta
                            : jb.tail ? a.pshare + (size_t)(blockIdx.x - 1) * per_cta
                                      : a.pscr + (size_t)blockIdx.x * per_cta;
            }
#ifdef DDP_TC_PROFILE
            long long pq0 = prof ? clock64() : 0;
#endif
            // ---- tile prologue: x_T -> registers; [x | state | 0] -> in0 in the first step's operand format
            if (jb.tail) {
                // the left neighbour ran steps 0 .. j0-1 of this tile first thing: its x_t has been waiting for a while
                if (threadIdx.x == 0) {
                    const int* flag = a.handoff_flag + (blockIdx.x - 1);
                    int v, spins = 0;
                    do {
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
                        if (!v && ++spins > (1 << 26)) asm volatile("trap;");       // a protocol bug must not hang the box
                    } while (!v);
                }
                epi_bar_sync();
            }
            if (e.ch == 0) {
                const float* hx = a.handoff + ((size_t)(blockIdx.x - 1) * kRows + e.my_row) * 8;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (jb.tail) e.xr[i] = e.owner ? __ldcg(hx + i) : 0.f;
                    else e.xr[i] = (e.valid && i < a.A) ? a.noise[e.row * a.A + i] : 0.f;
                }
                if (kH2 || a.first_f16) write_in0_row_full<true>(a, e); else write_in0_row_full<false>(a, e);
            }
            epi_bar_sync();
#ifdef DDP_TC_PROFILE
            if (prof) { const long long q = clock64(); a.dbg[8] += q - pq0; pq0 = q; }      // x_T / state -> in0, barrier
#endif
            L0Frags fr;
            if constexpr (kH2) {
                if (!jb.tail) state_partial<true, HM>(a, e);       // a tail re-uses the block its left neighbour parked
#ifdef DDP_TC_PROFILE
                if (prof) { const long long q = clock64(); a.dbg[9] += q - pq0; a.dbg[10] += 1; }   // state partial; jobs
#endif
                load_l0_frags<true, HM>(a, e, a.T - 1 - jb.j0, 0, fr);
                for (int j = jb.j0; j <= jb.j1; ++j) {
                    const bool nx = j < jb.j1;
#if defined(DDP_TC_H2_FIRST) && DDP_TC_H2_FIRST == 2
                    if (j == 0) epi_step<true, HM, true, false>(a, e, j, fr, nx); else epi_step<true, HM, true>(a, e, j, fr, nx);
#elif defined(DDP_TC_H2_FIRST)
                    epi_step<true, HM, true>(a, e, j, fr, nx);
#else
                    if (j == 0) epi_step<true, HM, false>(a, e, j, fr, nx); else epi_step<true, HM, true>(a, e, j, fr, nx);
#endif
                }
            } else {
                // round-1 plan: the operand format changes after the first step, so does the fragment set: no carry-over
                if (!jb.tail) { if (a.first_f16) state_partial<true, HM>(a, e); else state_partial<false, HM>(a, e); }
                for (int j = jb.j0; j <= jb.j1; ++j) {
                    if (a.first_f16 && j == 0) { load_l0_frags<true, HM>(a, e, a.T - 1 - j, 0, fr); epi_step<true, HM>(a, e, j, fr, false); }
                    else { load_l0_frags<false, HM>(a, e, a.T - 1 - j, 0, fr); epi_step<false, HM>(a, e, j, fr, false); }
                }
            }
            if (jb.head) {
                // leave x_t (after step j1) for the right neighbour, which finishes this tile at the end of its range
                if (e.ch == 0 && e.owner) {
                    float* hx = a.handoff + ((size_t)blockIdx.x * kRows + e.my_row) * 8;
#pragma unroll
                    for (int i = 0; i < 8; ++i) __stcg(hx + i, e.xr[i]);
                }
                __threadfence();         // every thread: its part of the shared state-partial block travels with x_t
                epi_bar_sync();
                if (threadIdx.x == 0) {
                    int one = 1;
                    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.handoff_flag + blockIdx.x), "r"(one) : "memory");
                }
            }
        }
#ifdef DDP_TC_PROFILE
        if (prof) {     // slots 12..14: cycles and nanoseconds of CTA 0's whole job list, number of tile-steps it ran
            long long pt1;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(pt1));
            a.dbg[12] += clock64() - pc0; a.dbg[13] += pt1 - pt0;
            long long units = 0;
            for (int ji = 0; ji < sch.njobs; ++ji) { const Job jb = sch.job(ji); units += jb.j1 - jb.j0 + 1; }
            a.dbg[14] += units;
        }
#endif
    }

    // ------------------------------------------------------------------ teardown
    tc_fence_before();
    __syncthreads();
    if (warp == kEpiWarps + 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------- packing
// Layer-0 B fragments in mma.sync m16n8k16 order.  K order is [x (A<=8, padded to 8) | state (S) | 0 ...].
// Entry [(c*2+ch)*12 + nt*3 + ks][lane] = {b0, b1}: feature f = c*64 + ch*32 + nt*8 + lane/4,
// b0 = (k, k+1) with k = ks*16 + 2*(lane%4), b1 = (k+8, k+9).
template <bool F16>
__device__ __forceinline__ void w0_frag_pack(size_t idx, const float* __restrict__ W0, int ld0, int D, int S, int A,
                                             uint2* __restrict__ out) {
    const int lane = (int)(idx & 31), e = (int)(idx >> 5);
    const int ks = e % 3, nt = (e / 3) % 4, cc = e / 12;           // cc = c*2 + ch
    const int f = cc * 32 + nt * 8 + (lane >> 2);
    auto wk = [&](int k) -> float {                                 // weight of kernel-order input k
        if (k < 8) return k < A ? W0[(size_t)f * ld0 + D + S + k] : 0.f;
        const int s = k - 8;
        return s < S ? W0[(size_t)f * ld0 + D + s] : 0.f;
    };
    const int k0 = ks * 16 + 2 * (lane & 3);
    uint2 v;
    v.x = pack2<F16>(wk(k0), wk(k0 + 1));
    v.y = pack2<F16>(wk(k0 + 8), wk(k0 + 9));
    out[idx] = v;
}

// Layer-3 B tiles: [h3/64][16 rows][64] bf16 written as the SWIZZLE_128B shared-memory image.
template <bool F16>
__device__ __forceinline__ void w3_image_pack(size_t idx, const float* __restrict__ W3, int A, int h3, uint16_t* __restrict__ out) {
    const int col = (int)(idx & 63), r = (int)((idx >> 6) & 15), c = (int)(idx >> 10);
    const float v = r < A ? W3[(size_t)r * h3 + c * 64 + col] : 0.f;
    out[(size_t)c * 1024 + sw128_offset(r, col) / 2] = cvt16<F16>(v);
}

// Every 16-bit operand of the tensor paths in ONE launch (it used to be four to eight: the pack runs inside every training
// step, where at the reference's 4 096-row batch a launch costs as much as the work).  Index space: W1 | W2 | layer-0
// fragments | head image, in bf16; with `sampler` the same four again in fp16.
struct TcPackJob {
    const float *W0, *W1, *W2, *W3;
    uint16_t *w1, *w2, *w3, *w1h, *w2h, *w3h;
    uint2 *w0, *w0h;
    size_t n0, n1, n2, n3;
    int ld0, D, S, A, h3, sampler;
};
__global__ void actor_tc_pack_kernel(TcPackJob j) {
    const size_t per = j.n1 + j.n2 + j.n0 + j.n3;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per * (j.sampler ? 2 : 1)) return;
    const bool h = i >= per;
    if (h) i -= per;
    if (i < j.n1) { if (h) j.w1h[i] = cvt16<true>(j.W1[i]); else j.w1[i] = cvt16<false>(j.W1[i]); return; }
    i -= j.n1;
    if (i < j.n2) { if (h) j.w2h[i] = cvt16<true>(j.W2[i]); else j.w2[i] = cvt16<false>(j.W2[i]); return; }
    i -= j.n2;
    if (i < j.n0) {
        if (h) w0_frag_pack<true>(i, j.W0, j.ld0, j.D, j.S, j.A, j.w0h); else w0_frag_pack<false>(i, j.W0, j.ld0, j.D, j.S, j.A, j.w0);
        return;
    }
    i -= j.n0;
    if (h) w3_image_pack<true>(i, j.W3, j.A, j.h3, j.w3h); else w3_image_pack<false>(i, j.W3, j.A, j.h3, j.w3);
}

struct TcPacked {      // byte offsets inside the packed buffer (see ActorLayout::tc_*)
    size_t w0frag, w1, w2, w3img;
};

}  // namespace

static long long* g_tc_dbg = nullptr;   // set by ddp_debug_tc_timing (development aid)

static bool tc_shape_ok(const ActorLayout& L) {
    return L.A <= 8 && L.S + 8 <= kK0 && L.h1 % 64 == 0 && L.h2 % 64 == 0 && L.h3 % 64 == 0 && L.h2 <= 512 &&
           L.h3 <= 256 && L.h3 + 16 <= kTmemCols && L.h3 / 64 <= kASlots && L.h3 <= L.h2 &&
           (L.h2 <= 256 || L.h2 % 256 == 0);
}

int pack_actor_tc(const ActorLayout& L, const float* const p[12], void* packed, bool sampler, bool train, cudaStream_t st) {
    if (!tc_shape_ok(L))
        DDP_FAIL(DDP_ERR_UNSUPPORTED, "DDP_BF16 actor path needs A<=8, S<=40, widths multiple of 64 with h2<=512, h3<=256");
    uint8_t* base = (uint8_t*)packed;
    auto blocks = [](size_t n) { return (unsigned)((n + 255) / 256); };
    const int ld0 = L.D + L.S + L.A;
    const size_t n0 = (size_t)(L.h1 / 32) * 12 * 32, n1 = (size_t)L.h2 * L.h1, n2 = (size_t)L.h3 * L.h2,
                 n3 = (size_t)(L.h3 / 64) * 1024;
    // bf16 W1 / W2 are read by the sampler and by the training GEMMs, the bf16 layer-0 fragments and head image by the
    // sampler and by the fused training forward (csrc/actor_train_chain_tc.cu); the fp16 copies are the sampler's own
    (void)train;
    TcPackJob j{};
    j.W0 = p[4]; j.W1 = p[6]; j.W2 = p[8]; j.W3 = p[10];
    j.w0 = (uint2*)(base + L.tc_w0); j.w0h = (uint2*)(base + L.tc_w0h);
    j.w1 = (uint16_t*)(base + L.tc_w1); j.w1h = (uint16_t*)(base + L.tc_w1h);
    j.w2 = (uint16_t*)(base + L.tc_w2); j.w2h = (uint16_t*)(base + L.tc_w2h);
    j.w3 = (uint16_t*)(base + L.tc_w3); j.w3h = (uint16_t*)(base + L.tc_w3h);
    j.n0 = n0; j.n1 = n1; j.n2 = n2; j.n3 = n3;
    j.ld0 = ld0; j.D = L.D; j.S = L.S; j.A = L.A; j.h3 = L.h3; j.sampler = sampler ? 1 : 0;
    actor_tc_pack_kernel<<<blocks((n0 + n1 + n2 + n3) * (sampler ? 2 : 1)), 256, 0, st>>>(j);
    DDP_LAUNCH_CHECK("actor tensor-core pack kernels");
    return DDP_OK;
}

static int tc_sm_count() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return 0;
    return sms;
}

// Tiles of a launch.  Batches of at most half a wave of 128-row tiles run on 64-row tiles instead (tcgen05.mma M = 64, its
// own kernel instantiation): twice the SMs and 0.135 instead of 0.17 ms per call at 256 ... 9 472 states.  For the partial
// LAST wave of a large batch the same idea was built and measured -- 136 64-row tiles behind the 444 128-row tiles of a
// 65 536-state launch, also with a programmatic dependent launch so that they fill SMs as the first kernel's CTAs retire --
// and gains nothing (0.5954 vs 0.5953 ms): a 64-row tile-step costs 0.89 of a 128-row one, because the weight stream
// (1.25 MB per tile-step from L2), the MMA issue time (M = 64 runs at the M = 128 rate) and the per-chunk hand-over do not
// shrink with the rows.
struct TilePlan { int num_full, num_half, grid_full, grid_half; };
static TilePlan tile_plan(long B, int sms) {
    constexpr bool no_half = false;
    const long tiles = (B + kRows - 1) / kRows;
    TilePlan p;
    p.num_full = (int)tiles; p.num_half = 0;
    if (sms > 0 && !no_half && 2 * tiles <= sms) { p.num_full = 0; p.num_half = (int)((B + 63) / 64); }
    p.grid_full = p.num_full < sms ? p.num_full : sms;
    p.grid_half = p.num_half < sms ? p.num_half : sms;
    return p;
}

// state partial sums: one [h1/64][8 warps][4][32 lanes] x 16 B block per resident CTA (256 KB at h1 = 1024)
// + per CTA of the 128-row launch, for a tile shared with the right neighbour (split_steps): a second state-partial
// block, the x_t hand-off slot and its flag
constexpr size_t kHandoffBytes = (size_t)kRows * 8 * sizeof(float);
static size_t pscr_cta_bytes(const ActorLayout& L) { return (size_t)(L.h1 / 64) * kEpiWarps * 4 * 32 * sizeof(uint4); }
static size_t pscr_bytes(const ActorLayout& L, const TilePlan& p) { return (size_t)(p.grid_full + p.grid_half) * pscr_cta_bytes(L); }
// tile-steps are dealt out in balanced contiguous ranges when whole tiles would leave a partial last wave and every
// range still spans two tiles or more (see Schedule)
static bool split_steps(const ActorLayout& L, const TilePlan& p) {
    return p.num_full > p.grid_full && p.num_full % p.grid_full != 0 && ((long)p.num_full * L.T) / p.grid_full >= 2L * L.T;
}
size_t actor_sample_tc_workspace(const ActorLayout& L, long B) {
    const TilePlan p = tile_plan(B, tc_sm_count());
    return pscr_bytes(L, p) + (split_steps(L, p) ? (size_t)p.grid_full * (pscr_cta_bytes(L) + kHandoffBytes + 64) : 0);
}

int actor_sample_tc(const ActorLayout& L, const void* packed, const float* state, const float* noise, float* out,
                    long B, const ExplNoise& expl, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!tc_shape_ok(L)) DDP_FAIL(DDP_ERR_UNSUPPORTED, "DDP_BF16 actor path does not support this shape");
    if (!ws || ((uintptr_t)ws & 15) || ws_bytes < actor_sample_tc_workspace(L, B))
        DDP_FAIL(DDP_ERR_ARG, "ddp_actor_sample (DDP_BF16): workspace missing, misaligned or smaller than ddp_actor_sample_workspace_bytes");
    const uint8_t* base = (const uint8_t*)packed;
    const float* pk = (const float*)packed;
    TcArgs a;
    a.w0frag = (const uint2*)(base + L.tc_w0);
    a.w0frag_h = (const uint2*)(base + L.tc_w0h);
    a.w3img = (const uint4*)(base + L.tc_w3);
    a.w3img_h = (const uint4*)(base + L.tc_w3h);
    // The first reverse step divides by sqrt(abar_{T-1}) (1e2..2e3): bf16 operand rounding of eps_hat would
    // surface as ~1e-2 action errors, so that one step uses fp16 operands (same tensor rate, 3 more mantissa
    // bits).  With DDP_TC_H2 (default) every step runs on fp16 operands and this flag is moot.
    a.first_f16 = 1;
    a.tb0 = pk + L.tb0; a.b1 = pk + L.b1; a.b2 = pk + L.b2; a.b3 = pk + L.b3; a.cst = pk + L.cst;
    a.state = state; a.noise = noise; a.out = out; a.B = B;
    a.S = L.S; a.A = L.A; a.T = L.T; a.h1 = L.h1; a.h2 = L.h2; a.h3 = L.h3;
    a.nparts1 = L.h2 > 256 ? L.h2 / 256 : 1;
    a.part1 = L.h2 / a.nparts1;
    const TilePlan plan = tile_plan(B, tc_sm_count());
    a.dbg = g_tc_dbg;
    a.expl = expl;
    CUtensorMap m1, m2, m1h, m2h;
    if (make_tmap_bf16_sw128(&m1, base + L.tc_w1, L.h2, L.h1, a.part1) != 0 ||
        make_tmap_bf16_sw128(&m2, base + L.tc_w2, L.h3, L.h2, L.h3) != 0 ||
        make_tmap_bf16_sw128(&m1h, base + L.tc_w1h, L.h2, L.h1, a.part1, true) != 0 ||
        make_tmap_bf16_sw128(&m2h, base + L.tc_w2h, L.h3, L.h2, L.h3, true) != 0)
        DDP_FAIL(DDP_ERR_CUDA, "cuTensorMapEncodeTiled failed for the actor weight tiles");
    const int sms = tc_sm_count();
    if (sms <= 0) DDP_FAIL(DDP_ERR_CUDA, "cannot query the SM count");
    const size_t smem = SM::total + 1024;         // slack for the 1024-byte alignment of the base
    const size_t per_cta = (size_t)(L.h1 / 64) * kEpiWarps * 4 * 32;      // uint4 of state-partial scratch per CTA
    a.split_steps = split_steps(L, plan) ? 1 : 0;
    a.pshare = (uint4*)((uint8_t*)ws + pscr_bytes(L, plan));
    a.handoff = (float*)((uint8_t*)a.pshare + (size_t)plan.grid_full * pscr_cta_bytes(L));
    a.handoff_flag = (int*)((uint8_t*)a.handoff + (size_t)plan.grid_full * kHandoffBytes);
    if (plan.num_full > 0) {
        if (a.split_steps) DDP_CUDA_CHECK(cudaMemsetAsync(a.handoff_flag, 0, (size_t)plan.grid_full * sizeof(int), st));
        a.num_tiles = plan.num_full; a.row0 = 0; a.pscr = (uint4*)ws;
        DDP_CUDA_CHECK(cudaFuncSetAttribute(actor_sample_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        actor_sample_tc_kernel<false><<<plan.grid_full, kThreads, smem, st>>>(m1, m2, m1h, m2h, a);
    }
    if (plan.num_half > 0) {
        a.num_tiles = plan.num_half; a.row0 = (long)plan.num_full * kRows; a.pscr = (uint4*)ws + (size_t)plan.grid_full * per_cta;
        DDP_CUDA_CHECK(cudaFuncSetAttribute(actor_sample_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        actor_sample_tc_kernel<true><<<plan.grid_half, kThreads, smem, st>>>(m1, m2, m1h, m2h, a);
    }
    DDP_LAUNCH_CHECK("actor_sample_tc_kernel");
    return DDP_OK;
}

}  // namespace ddp

// ---------------------------------------------------------------------------------------- self-test
// C[128][N] = A[128][K] . B[N][K]^T with the SAME building blocks as the sampler (manual SWIZZLE_128B
// A chunks, TMA-loaded B tiles, tcgen05.mma into TMEM, 32x32b TMEM loads).  One CTA; used by
// tests/test_tc_gpu.py (test_tcgen05_blocks_selftest_gemm) to pin descriptors and layouts independently of the fused kernel.
namespace ddp {
namespace {
__global__ void __launch_bounds__(128, 1)
tc_gemm_selftest_kernel(const __grid_constant__ CUtensorMap map_b, const __nv_bfloat16* __restrict__ A,
                        float* __restrict__ C, int N, int K) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    const uint32_t a_off = 0, b_off = kChunkBytes, bar_off = kChunkBytes + kStageBytes, tp_off = bar_off + 16;
    const uint32_t bar_full = base + bar_off, bar_done = base + bar_off + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar_full, 1); mbar_init(bar_done, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(base + tp_off, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + tp_off);
    const uint32_t idesc = make_idesc_bf16(128, N);
    uint32_t phase = 0;
    for (int kc = 0; kc < K / 64; ++kc) {
        // A chunk: thread r writes its row (64 bf16) as 8 swizzled 16-byte pieces
        const uint4* src = reinterpret_cast<const uint4*>(A + (size_t)threadIdx.x * K + kc * 64);
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(smem + a_off + sw128_offset(threadIdx.x, j * 8)) = src[j];
        fence_proxy_async();
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar_full, (uint32_t)N * 128u);
            tma_load_2d(base + b_off, &map_b, bar_full, kc * 64, 0);
            mbar_wait(bar_full, phase);
            tc_fence_after();
            const uint64_t ad = make_smem_desc_sw128(base + a_off), bd = make_smem_desc_sw128(base + b_off);
            for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, ad + 2 * k, bd + 2 * k, idesc, (kc | k) != 0);
            umma_commit(bar_done);
            mbar_wait(bar_done, phase);
        }
        phase ^= 1;
        __syncthreads();
    }
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int i = 0; i < 32 && c0 + i < N; ++i) C[(size_t)(warp * 32 + lane) * N + c0 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}
}  // namespace
}  // namespace ddp

// Debug entry point (not part of the public header): A [128][K] bf16, Bm [N][K] bf16, C [128][N] fp32.
extern "C" int ddp_debug_tc_gemm(const void* A, const void* Bm, float* C, int N, int K, void* stream) {
    using namespace ddp;
    if (N % 16 || N < 16 || N > 256 || K % 64 || K <= 0) DDP_FAIL(DDP_ERR_SHAPE, "selftest: N in [16,256] step 16, K multiple of 64");
    CUtensorMap m;
    if (tc::make_tmap_bf16_sw128(&m, Bm, N, K, N) != 0) DDP_FAIL(DDP_ERR_CUDA, "selftest: tensor map encode failed");
    const size_t smem = kChunkBytes + kStageBytes + 64 + 1024;
    DDP_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_gemm_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(m, (const __nv_bfloat16*)A, C, N, K);
    DDP_LAUNCH_CHECK("tc_gemm_selftest_kernel");
    return DDP_OK;
}

// Layout probe (development aid for the 64-row tile plan): one tcgen05.mma with M = 64 (A: 64 rows x 64 bf16, manual
// SWIZZLE_128B; B: [N][64] by TMA), then ALL 128 TMEM lanes x N columns are dumped with 32x32b loads, so the host can
// see which (lane, column) each accumulator element D[r][n] landed in.  raw: [128][N] fp32.
namespace ddp {
namespace {
__global__ void __launch_bounds__(128, 1)
tc_m64_probe_kernel(const __grid_constant__ CUtensorMap map_b, const __nv_bfloat16* __restrict__ A, float* __restrict__ raw, int N) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t rawa = smem_u32(smem_raw);
    const uint32_t base = (rawa + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - rawa);
    const uint32_t a_off = 0, b_off = kChunkBytes, bar_off = kChunkBytes + kStageBytes, tp_off = bar_off + 16;
    const uint32_t bar_full = base + bar_off, bar_done = base + bar_off + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar_full, 1); mbar_init(bar_done, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(base + tp_off, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + tp_off);
    // poison the accumulator region first so that untouched lanes are recognisable
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(-1.f);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                     ::"r"(tmem_base + ((uint32_t)(warp * 32) << 16) + c0), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]),
                       "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    if (threadIdx.x < 64) {
        const uint4* src = reinterpret_cast<const uint4*>(A + (size_t)threadIdx.x * 64);
        for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(smem + a_off + sw128_offset(threadIdx.x, j * 8)) = src[j];
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
        tc_fence_after();
        mbar_expect_tx(bar_full, (uint32_t)N * 128u);
        tma_load_2d(base + b_off, &map_b, bar_full, 0, 0);
        mbar_wait(bar_full, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(64, N);
        const uint64_t ad = make_smem_desc_sw128(base + a_off), bd = make_smem_desc_sw128(base + b_off);
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, ad + 2 * k, bd + 2 * k, idesc, k != 0);
        umma_commit(bar_done);
        mbar_wait(bar_done, 0);
    }
    __syncthreads();
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) raw[(size_t)(warp * 32 + lane) * N + c0 + i] = __uint_as_float(v[i]);
    }
    // second dump: the first 8 columns of each lane quarter through the 16-lane shape (16x256b.x1: 4 registers per
    // thread), appended as frag[warp][thread][4] after the raw block
    {
        uint32_t f[4];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(f[0]), "=r"(f[1]), "=r"(f[2]), "=r"(f[3]) : "r"(tmem_base + ((uint32_t)(warp * 32) << 16)));
        tmem_ld_wait();
        for (int i = 0; i < 4; ++i) raw[(size_t)128 * N + (warp * 32 + lane) * 4 + i] = __uint_as_float(f[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}
}  // namespace
}  // namespace ddp

extern "C" int ddp_debug_tc_m64_probe(const void* A, const void* Bm, float* raw, int N, void* stream) {
    using namespace ddp;
    if (N % 16 || N < 16 || N > 256) DDP_FAIL(DDP_ERR_SHAPE, "m64 probe: N in [16,256] step 16");
    CUtensorMap m;
    if (tc::make_tmap_bf16_sw128(&m, Bm, N, 64, N) != 0) DDP_FAIL(DDP_ERR_CUDA, "m64 probe: tensor map encode failed");
    const size_t smem = kChunkBytes + kStageBytes + 64 + 1024;
    DDP_CUDA_CHECK(cudaFuncSetAttribute(tc_m64_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_m64_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(m, (const __nv_bfloat16*)A, raw, N);
    DDP_LAUNCH_CHECK("tc_m64_probe_kernel");
    return DDP_OK;
}

// Debug entry point (not part of the public header): device buffer of >= 8 int64 that CTA 0 of the next
// sampler launches accumulates per-phase cycle counts into (NULL switches it off).
extern "C" void ddp_debug_tc_timing(long long* dev_buf) { ddp::g_tc_dbg = dev_buf; }
