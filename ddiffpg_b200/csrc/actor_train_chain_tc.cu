// H3, bf16 tensor-core path, fused FORWARD: the four layers of the denoiser, the loss and d loss / d eps_hat for a
// 128-row tile in one persistent, warp-specialised sm_100a kernel (the structure of the fused sampler,
// csrc/actor_sample_tc.cu, run for one step with a per-row timestep).  Activations and Mish derivatives are needed
// again by the backward and the weight-gradient GEMMs, so every 64-column chunk that feeds the next layer's MMA from
// shared memory is ALSO shipped to global memory by a TMA store from the same buffer -- no thread ever issues a
// global store for them.
//
// Reference semantics (paths relative to the reference repo):
//   DiffusionPolicy.get_loss                     ddiffpg/models/diffusion_mlp.py:294-321
//   DiffusionNet.forward (trunk, Mish)           ddiffpg/models/diffusion_mlp.py:50-58,62-73
//
// Per CTA (one per SM, persistent over tiles):
//   layer 0 (K = [x_noisy|state] = 42 -> 48): mma.sync in the 8 epilogue warps, accumulators start from the time
//            table row of each row's timestep; Mish and Mish' -> one activation chunk (A ring) and one derivative
//            chunk (D ring) per 64 features
//   layer 1 / 2: tcgen05.mma K-outer over the chunks as they appear, weights streamed by TMA; the epilogue warps
//            drain the accumulator (TMEM), + bias, Mish / Mish' -> next chunks
//   layer 3 (head): eps_hat -> squared error (atomic partial of the loss) and d loss / d eps_hat (bf16, padded rows)
//   store lane: for every published chunk pair, TMA-stores both buffers to a_l / d_l and only then lets the slot be
//            recycled (the MMA commit is the other half of the slot's release)
#include "actor_layout.cuh"
#include "tc_common.cuh"

namespace ddp {
using namespace tc;

namespace {

using bf16 = __nv_bfloat16;

constexpr int kRows = 128;
constexpr int kChunkBytes = kRows * 128;
constexpr int kStageBytes = 256 * 128;
constexpr int kStages = 3;
constexpr int kASlots = 4;
constexpr int kDSlots = 2;
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = kEpiThreads + 96;     // + TMA-load warp + MMA warp + TMA-store warp
constexpr int kIn0Stride = 56;
constexpr int kK0 = 48;
constexpr int kTmemCols = 512;
constexpr int kNT = 4;                          // mma.sync n8 tiles per warp per chunk (32 columns)
constexpr float kLog2e = 1.4426950408889634f;

struct alignas(64) FcMaps { CUtensorMap w1, w2, a[3], d[3]; };

struct FcArgs {
    const uint2* w0frag;       // layer-0 B fragments (bf16, mma.sync order)
    const uint4* w3img;        // head weights, pre-swizzled shared-memory image
    const float *tb0, *b1, *b2, *b3;
    const bf16* xin;           // [B][64] rows [x_noisy | state | 0]
    const int64_t* t;          // [B] timesteps
    const float* noise;        // [B][A] regression target
    float* loss_out;
    bf16* deps;                // [B][64] d loss / d eps_hat, zero padded
    float inv_count;
    long B;
    int A, T, h1, h2, h3, nparts1, part1, num_tiles;
};

struct SF {
    static constexpr uint32_t wring = 0;
    static constexpr uint32_t aring = wring + kStages * kStageBytes;
    static constexpr uint32_t dring = aring + kASlots * kChunkBytes;
    static constexpr uint32_t w3 = dring + kDSlots * kChunkBytes;
    static constexpr uint32_t in0 = w3 + 4 * 2048;
    static constexpr uint32_t b1 = in0 + kRows * kIn0Stride * 2;
    static constexpr uint32_t b2 = b1 + 512 * 4;
    static constexpr uint32_t bars = b2 + 256 * 4;
    static constexpr uint32_t tmem_ptr = bars + 8 * 24;
    static constexpr uint32_t total = tmem_ptr + 8;
};
static_assert(SF::total + 1024 <= 227 * 1024, "shared-memory map exceeds the 227 KB per-CTA limit");

__device__ __forceinline__ uint32_t fb_w_full(uint32_t b, int i) { return b + 8 * i; }
__device__ __forceinline__ uint32_t fb_w_empty(uint32_t b, int i) { return b + 8 * (kStages + i); }
__device__ __forceinline__ uint32_t fb_a_full(uint32_t b, int i) { return b + 8 * (2 * kStages + i); }
__device__ __forceinline__ uint32_t fb_a_empty(uint32_t b, int i) { return b + 8 * (2 * kStages + kASlots + i); }
__device__ __forceinline__ uint32_t fb_d_full(uint32_t b, int i) { return b + 8 * (2 * kStages + 2 * kASlots + i); }
__device__ __forceinline__ uint32_t fb_d_empty(uint32_t b, int i) { return b + 8 * (2 * kStages + 2 * kASlots + kDSlots + i); }
__device__ __forceinline__ uint32_t fb_acc_full(uint32_t b) { return b + 8 * (2 * kStages + 2 * kASlots + 2 * kDSlots); }
__device__ __forceinline__ uint32_t fb_lo_free(uint32_t b) { return b + 8 * (2 * kStages + 2 * kASlots + 2 * kDSlots + 1); }

__device__ __forceinline__ void f_epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

struct FRing {
    int idx = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance(int n) { if (++idx == n) { idx = 0; phase ^= 1; } }
};

// Mish and its derivative, N values: with e = exp(x), s = e + 1, p = s^2 + 1, r = 1/p:
//   mish(x) = x (1 - 2r),  mish'(x) = (1 - 2r) + 4 x e s r^2        (one ex2 + one rcp per element)
// the exponent is capped at x = 20 (mish = x, mish' = 1 to fp32 there; e s r^2 stays finite)
// DDP_FC_MISH_V2 (default): the same two functions with the exponent shifted by +1/2 (E = sqrt2 e^x, v = (e^x + 1)/sqrt2
// and q = v^2 + 1/2 = p/2 each cost one FFMA; r = 1/q = 2/p, mish = x (1 - r), mish' = (1 - r) + x (E v) r^2):
// 11.5 FP32 + 1.5 MUFU instructions per element instead of 13.5 + 1.5 (training step 0.986 -> 0.979 ms at 131 072 rows,
// A B A B in one call; fp32 error of both functions unchanged, <= 2e-6 absolute).  0 keeps the first form below.
#ifndef DDP_FC_MISH_V2
#define DDP_FC_MISH_V2 1
#endif
template <int N>
__device__ __forceinline__ void mish_fd_n(float (&x)[N], float (&d)[N]) {
    float e[N];
    if (DDP_FC_MISH_V2) {
#pragma unroll
        for (int i = 0; i < N; ++i) e[i] = ex2_approx(fminf(fmaf(x[i], kLog2e, 0.5f), 29.353900817779268f));
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            const float v0 = fmaf(e[i], 0.5f, 0.70710678118654752f), v1 = fmaf(e[i + 1], 0.5f, 0.70710678118654752f);
            const float q0 = fmaf(v0, v0, 0.5f), q1 = fmaf(v1, v1, 0.5f);
            const float r = rcp_approx(q0 * q1);
            const float r0 = r * q1, r1 = r * q0;
            const float w0 = 1.f - r0, w1 = 1.f - r1;
            d[i] = fmaf(x[i], (e[i] * v0) * (r0 * r0), w0);
            d[i + 1] = fmaf(x[i + 1], (e[i + 1] * v1) * (r1 * r1), w1);
            x[i] *= w0;
            x[i + 1] *= w1;
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) e[i] = ex2_approx(fminf(x[i] * kLog2e, 28.853900817779268f));
    // one MUFU.RCP per two elements: 1/p0, 1/p1 = r p1, r p0 with r = rcp(p0 p1) (p <= e^40 + ..., the product stays
    // far below 2^128); the XU pipe (ex2, rcp and the two bf16 packs per element) is what this epilogue saturates
#pragma unroll
    for (int i = 0; i < N; i += 2) {
        const float s0 = e[i] + 1.f, s1 = e[i + 1] + 1.f;
        const float p0 = fmaf(s0, s0, 1.f), p1 = fmaf(s1, s1, 1.f);
        const float r = rcp_approx(p0 * p1);
        const float r0 = r * p1, r1 = r * p0;
        const float w0 = fmaf(-2.f, r0, 1.f), w1 = fmaf(-2.f, r1, 1.f);
        d[i] = fmaf(4.f * x[i], (e[i] * s0) * (r0 * r0), w0);
        d[i + 1] = fmaf(4.f * x[i + 1], (e[i + 1] * s1) * (r1 * r1), w1);
        x[i] *= w0;
        x[i + 1] *= w1;
    }
}

struct FEpi {
    uint8_t* smem;
    uint32_t bars, tmem_base, acc_phase;
    int q, ch, g, t4, lane, my_row;
    FRing as, ds;
    // both slots of the next chunk are free again (MMA + store lane released the A slot, store lane the D slot)
    __device__ __forceinline__ void acquire(uint8_t*& aslot, uint8_t*& dslot) {
        mbar_wait(fb_a_empty(bars, as.idx), as.phase ^ 1);
        mbar_wait(fb_d_empty(bars, ds.idx), ds.phase ^ 1);
        aslot = smem + SF::aring + as.idx * kChunkBytes;
        dslot = smem + SF::dring + ds.idx * kChunkBytes;
    }
    __device__ __forceinline__ void publish(bool signal_lo) {
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(fb_a_full(bars, as.idx));
            mbar_arrive(fb_d_full(bars, ds.idx));
            if (signal_lo) mbar_arrive(fb_lo_free(bars));
        }
        as.advance(kASlots);
        ds.advance(kDSlots);
    }
};

__device__ __forceinline__ void f_store16(uint8_t* slot, int row, int col0, const float (&x)[16]) {
#pragma unroll
    for (int i8 = 0; i8 < 2; ++i8) {
        uint4 w;
        w.x = pack_bf16x2(x[i8 * 8 + 0], x[i8 * 8 + 1]); w.y = pack_bf16x2(x[i8 * 8 + 2], x[i8 * 8 + 3]);
        w.z = pack_bf16x2(x[i8 * 8 + 4], x[i8 * 8 + 5]); w.w = pack_bf16x2(x[i8 * 8 + 6], x[i8 * 8 + 7]);
        *reinterpret_cast<uint4*>(slot + sw128_offset(row, col0 + i8 * 8)) = w;
    }
}

// 16 accumulator columns of this thread's row -> + bias, Mish / Mish' -> the two chunk buffers
__device__ __forceinline__ void f_emit(const FEpi& e, uint8_t* aslot, uint8_t* dslot, const uint32_t (&v)[16],
                                       const float* bb, int col0) {
    float x[16], d[16];
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) {
        const float4 b = *reinterpret_cast<const float4*>(bb + i4 * 4);
        x[i4 * 4 + 0] = __uint_as_float(v[i4 * 4 + 0]) + b.x; x[i4 * 4 + 1] = __uint_as_float(v[i4 * 4 + 1]) + b.y;
        x[i4 * 4 + 2] = __uint_as_float(v[i4 * 4 + 2]) + b.z; x[i4 * 4 + 3] = __uint_as_float(v[i4 * 4 + 3]) + b.w;
    }
    mish_fd_n<16>(x, d);
    f_store16(aslot, e.my_row, col0, x);
    f_store16(dslot, e.my_row, col0, d);
}

// accumulator at TMEM column 0 -> nchunks (activation, derivative) chunk pairs; lo_free after chunk `signal_after`
__device__ __forceinline__ void f_drain(FEpi& e, int nchunks, const float* bias, int signal_after) {
    const uint32_t tbase = e.tmem_base + ((uint32_t)(e.q * 32) << 16) + e.ch * 32;
    uint32_t va[16], vb[16];
    tmem_ld16(tbase, va);
    for (int c = 0; c < nchunks; ++c) {
        const float* bb = bias + c * 64 + e.ch * 32;
        tmem_ld_wait();
        tmem_ld16(tbase + c * 64 + 16, vb);
        uint8_t *aslot, *dslot;
        e.acquire(aslot, dslot);
        f_emit(e, aslot, dslot, va, bb, e.ch * 32);
        tmem_ld_wait();
        if (c + 1 < nchunks) tmem_ld16(tbase + (c + 1) * 64, va);
        f_emit(e, aslot, dslot, vb, bb + 16, e.ch * 32 + 16);
        e.publish(c == signal_after);
    }
}

__global__ void __launch_bounds__(kThreads, 1)
actor_train_chain_kernel(const __grid_constant__ FcMaps maps, const FcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    const uint32_t bars = base + SF::bars;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NC1 = a.h1 >> 6, NC2 = a.h2 >> 6, NC3 = a.h3 >> 6;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(fb_w_full(bars, i), 1); mbar_init(fb_w_empty(bars, i), 1); }
        for (int i = 0; i < kASlots; ++i) { mbar_init(fb_a_full(bars, i), kEpiWarps); mbar_init(fb_a_empty(bars, i), 2); }
        for (int i = 0; i < kDSlots; ++i) { mbar_init(fb_d_full(bars, i), kEpiWarps); mbar_init(fb_d_empty(bars, i), 1); }
        mbar_init(fb_acc_full(bars), 1);
        mbar_init(fb_lo_free(bars), kEpiWarps);
        fence_barrier_init();
    }
    if (warp == kEpiWarps + 1) tmem_alloc(base + SF::tmem_ptr, kTmemCols);
    {
        const int n16 = NC3 * 2048 / 16;
        uint4* dst = reinterpret_cast<uint4*>(smem + SF::w3);
        for (int i = threadIdx.x; i < n16; i += kThreads) dst[i] = a.w3img[i];
        float* sb1 = reinterpret_cast<float*>(smem + SF::b1);
        float* sb2 = reinterpret_cast<float*>(smem + SF::b2);
        for (int i = threadIdx.x; i < a.h2; i += kThreads) sb1[i] = a.b1[i];
        for (int i = threadIdx.x; i < a.h3; i += kThreads) sb2[i] = a.b2[i];
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + SF::tmem_ptr);

    if (warp == kEpiWarps) {
        // ============================================================== TMA weight loads (one lane)
        if (lane == 0) {
            tma_prefetch_desc(&maps.w1);
            tma_prefetch_desc(&maps.w2);
            FRing ws;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
                for (int c = 0; c < NC1; ++c)
                    for (int p = 0; p < a.nparts1; ++p) {
                        mbar_wait(fb_w_empty(bars, ws.idx), ws.phase ^ 1);
                        mbar_expect_tx(fb_w_full(bars, ws.idx), (uint32_t)a.part1 * 128u);
                        tma_load_2d(base + SF::wring + ws.idx * kStageBytes, &maps.w1, fb_w_full(bars, ws.idx), c * 64, p * a.part1);
                        ws.advance(kStages);
                    }
                for (int c = 0; c < NC2; ++c) {
                    mbar_wait(fb_w_empty(bars, ws.idx), ws.phase ^ 1);
                    mbar_expect_tx(fb_w_full(bars, ws.idx), (uint32_t)a.h3 * 128u);
                    tma_load_2d(base + SF::wring + ws.idx * kStageBytes, &maps.w2, fb_w_full(bars, ws.idx), c * 64, 0);
                    ws.advance(kStages);
                }
            }
        }
    } else if (warp == kEpiWarps + 1) {
        // ============================================================== MMA issuer (one lane)
        if (lane == 0) {
            FRing ws, as;
            uint32_t lo_phase = 0;
            const uint32_t idesc1 = make_idesc_bf16(kRows, a.part1), idesc2 = make_idesc_bf16(kRows, a.h3),
                           idesc3 = make_idesc_bf16(kRows, 16);
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
                for (int c = 0; c < NC1; ++c) {                     // layer 1: acc1[128 x h2] at cols [0, h2)
                    mbar_wait(fb_a_full(bars, as.idx), as.phase);
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc_sw128(base + SF::aring + as.idx * kChunkBytes);
                    for (int p = 0; p < a.nparts1; ++p) {
                        mbar_wait(fb_w_full(bars, ws.idx), ws.phase);
                        tc_fence_after();
                        const uint64_t bdesc = make_smem_desc_sw128(base + SF::wring + ws.idx * kStageBytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(tmem_base + p * a.part1, adesc + 2 * k, bdesc + 2 * k, idesc1, (c | k) != 0);
                        umma_commit(fb_w_empty(bars, ws.idx));
                        ws.advance(kStages);
                    }
                    umma_commit(fb_a_empty(bars, as.idx));
                    as.advance(kASlots);
                }
                umma_commit(fb_acc_full(bars));
                mbar_wait(fb_lo_free(bars), lo_phase);              // cols [0, h3) of acc1 are drained
                lo_phase ^= 1;
                tc_fence_after();
                for (int c = 0; c < NC2; ++c) {                     // layer 2: acc2[128 x h3] at cols [0, h3)
                    mbar_wait(fb_a_full(bars, as.idx), as.phase);
                    mbar_wait(fb_w_full(bars, ws.idx), ws.phase);
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc_sw128(base + SF::aring + as.idx * kChunkBytes);
                    const uint64_t bdesc = make_smem_desc_sw128(base + SF::wring + ws.idx * kStageBytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc2, (c | k) != 0);
                    umma_commit(fb_w_empty(bars, ws.idx));
                    umma_commit(fb_a_empty(bars, as.idx));
                    ws.advance(kStages);
                    as.advance(kASlots);
                }
                umma_commit(fb_acc_full(bars));
                for (int c = 0; c < NC3; ++c) {                     // head: acc3[128 x 16] at cols [h3, h3+16)
                    mbar_wait(fb_a_full(bars, as.idx), as.phase);
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc_sw128(base + SF::aring + as.idx * kChunkBytes);
                    const uint64_t bdesc = make_smem_desc_sw128(base + SF::w3 + c * 2048);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + a.h3, adesc + 2 * k, bdesc + 2 * k, idesc3, (c | k) != 0);
                    umma_commit(fb_a_empty(bars, as.idx));
                    as.advance(kASlots);
                }
                umma_commit(fb_acc_full(bars));
            }
        }
    } else if (warp == kEpiWarps + 2) {
        // ============================================================== TMA store lane: chunk buffers -> a_l / d_l
        if (lane == 0) {
            FRing as, ds;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
                const int row0 = tile * kRows;
                for (int layer = 0; layer < 3; ++layer) {
                    const int nc = layer == 0 ? NC1 : (layer == 1 ? NC2 : NC3);
                    for (int c = 0; c < nc; ++c) {
                        mbar_wait(fb_a_full(bars, as.idx), as.phase);
                        mbar_wait(fb_d_full(bars, ds.idx), ds.phase);
                        // (the writers fenced their st.shared towards the async proxy before arriving)
                        tma_store_2d(&maps.a[layer], base + SF::aring + as.idx * kChunkBytes, c * 64, row0);
                        tma_store_2d(&maps.d[layer], base + SF::dring + ds.idx * kChunkBytes, c * 64, row0);
                        tma_store_commit();
                        tma_store_wait_read();                      // both buffers have been read out
                        mbar_arrive(fb_a_empty(bars, as.idx));
                        mbar_arrive(fb_d_empty(bars, ds.idx));
                        as.advance(kASlots);
                        ds.advance(kDSlots);
                    }
                }
            }
            tma_store_wait_all();
        }
    } else {
        // ============================================================== epilogue / layer-0 warps
        FEpi e;
        e.smem = smem; e.bars = bars; e.tmem_base = tmem_base; e.acc_phase = 0;
        e.q = warp & 3; e.ch = warp >> 2; e.g = lane >> 2; e.t4 = lane & 3; e.lane = lane;
        e.my_row = e.q * 32 + lane;
        const float* sb1 = reinterpret_cast<const float*>(smem + SF::b1);
        const float* sb2 = reinterpret_cast<const float*>(smem + SF::b2);
        const uint32_t in0_lane = smem_u32(smem + SF::in0) +
            (uint32_t)(((e.q * 32 + (lane & 7) + ((lane >> 3) & 1) * 8) * kIn0Stride + (lane >> 4) * 8) * 2);
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
            const long row = (long)tile * kRows + e.my_row;
            const bool valid = row < a.B;
            // ---- tile prologue: this row's [x_noisy | state | 0] (48 of the 64 prepared columns) -> layer-0 input tile
            if (e.ch == 0) {
                uint4* dst = reinterpret_cast<uint4*>(smem + SF::in0 + e.my_row * kIn0Stride * 2);
                const uint4* src = reinterpret_cast<const uint4*>(a.xin + row * 64);
#pragma unroll
                for (int i = 0; i < kK0 / 8; ++i) dst[i] = valid ? __ldg(src + i) : make_uint4(0, 0, 0, 0);
            }
            // timesteps of the four rows this thread accumulates in the mma.sync fragments
            const float* tbr[2][2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const long r = (long)tile * kRows + e.q * 32 + mt * 16 + hf * 8 + e.g;
                    int t = r < a.B ? (int)__ldg(a.t + r) : 0;
                    t = t < 0 ? 0 : (t >= a.T ? a.T - 1 : t);
                    tbr[mt][hf] = a.tb0 + (size_t)t * a.h1 + e.ch * 32 + 2 * e.t4;
                }
            f_epi_bar_sync();

            // ---- layer 0: one 64-feature chunk at a time
            uint2 bfr[kNT][3];
            auto load_frags = [&](int c, uint2 (&f)[kNT][3]) {
                const uint2* bf = a.w0frag + ((size_t)(c * 8 + e.ch * kNT) * 3) * 32 + lane;
#pragma unroll
                for (int nt = 0; nt < kNT; ++nt)
#pragma unroll
                    for (int ks = 0; ks < 3; ++ks) f[nt][ks] = __ldg(bf + (nt * 3 + ks) * 32);
            };
            load_frags(0, bfr);
            // accumulators start from the time-table rows of the four rows a thread holds; the loads for chunk c + 1 are
            // issued into the (dead) accumulator registers as soon as chunk c has been stored, not at the top of c + 1
            float acc[2][kNT][4];
            auto acc_init = [&](int c) {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < kNT; ++nt) {
                        const float2 lo = __ldg(reinterpret_cast<const float2*>(tbr[mt][0] + c * 64 + nt * 8));
                        const float2 hi = __ldg(reinterpret_cast<const float2*>(tbr[mt][1] + c * 64 + nt * 8));
                        acc[mt][nt][0] = lo.x; acc[mt][nt][1] = lo.y; acc[mt][nt][2] = hi.x; acc[mt][nt][3] = hi.y;
                    }
            };
            acc_init(0);
            for (int c = 0; c < NC1; ++c) {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int ks = 0; ks < 3; ++ks) {
                        uint32_t af[4];
                        ldmatrix_x4(af, in0_lane + (uint32_t)((mt * 16 * kIn0Stride + ks * 16) * 2));
#pragma unroll
                        for (int nt = 0; nt < kNT; ++nt) mma_m16n8k16<false>(acc[mt][nt], af, bfr[nt][ks].x, bfr[nt][ks].y);
                    }
                if (c + 1 < NC1) load_frags(c + 1, bfr);
                uint8_t *aslot, *dslot;
                e.acquire(aslot, dslot);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    float d[kNT * 4];
                    mish_fd_n<kNT * 4>(reinterpret_cast<float(&)[kNT * 4]>(acc[mt]), d);
#pragma unroll
                    for (int nt = 0; nt < kNT; ++nt) {
                        const int r0 = e.q * 32 + mt * 16 + e.g, col = e.ch * 32 + nt * 8 + 2 * e.t4;
                        const uint32_t o0 = sw128_offset(r0, col), o1 = sw128_offset(r0 + 8, col);
                        *reinterpret_cast<uint32_t*>(aslot + o0) = pack_bf16x2(acc[mt][nt][0], acc[mt][nt][1]);
                        *reinterpret_cast<uint32_t*>(aslot + o1) = pack_bf16x2(acc[mt][nt][2], acc[mt][nt][3]);
                        *reinterpret_cast<uint32_t*>(dslot + o0) = pack_bf16x2(d[nt * 4 + 0], d[nt * 4 + 1]);
                        *reinterpret_cast<uint32_t*>(dslot + o1) = pack_bf16x2(d[nt * 4 + 2], d[nt * 4 + 3]);
                    }
                }
                if (c + 1 < NC1) acc_init(c + 1);
                e.publish(false);
            }
            // ---- layer-1 / layer-2 epilogues
            mbar_wait(fb_acc_full(bars), e.acc_phase); e.acc_phase ^= 1;
            tc_fence_after();
            f_drain(e, NC2, sb1, NC3 - 1);
            mbar_wait(fb_acc_full(bars), e.acc_phase); e.acc_phase ^= 1;
            tc_fence_after();
            f_drain(e, NC3, sb2, -1);
            // ---- head: eps_hat, squared error, d loss / d eps_hat
            mbar_wait(fb_acc_full(bars), e.acc_phase); e.acc_phase ^= 1;
            tc_fence_after();
            if (e.ch == 0) {
                uint32_t ev[8];
                tmem_ld8(tmem_base + ((uint32_t)(e.q * 32) << 16) + a.h3, ev);
                tmem_ld_wait();
                float sq = 0.f, d8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float diff = 0.f;
                    if (valid && i < a.A) diff = __uint_as_float(ev[i]) + __ldg(a.b3 + i) - __ldg(a.noise + row * a.A + i);
                    sq = fmaf(diff, diff, sq);
                    d8[i] = 2.f * diff * a.inv_count;
                }
                if (valid) {
                    uint4* dp = reinterpret_cast<uint4*>(a.deps + row * 64);
                    uint4 w;
                    w.x = pack_bf16x2(d8[0], d8[1]); w.y = pack_bf16x2(d8[2], d8[3]);
                    w.z = pack_bf16x2(d8[4], d8[5]); w.w = pack_bf16x2(d8[6], d8[7]);
                    dp[0] = w;
#pragma unroll
                    for (int i = 1; i < 8; ++i) dp[i] = make_uint4(0, 0, 0, 0);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                if (lane == 0 && sq != 0.f) atomicAdd(a.loss_out, sq * a.inv_count);
            }
            tc_fence_before();
            f_epi_bar_sync();         // acc3 reads complete and the input tile may be rewritten
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kEpiWarps + 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace

bool actor_train_chain_shape_ok(const ActorLayout& L) {
    return L.A <= 8 && L.S + 8 <= kK0 && L.h1 % 64 == 0 && L.h2 % 64 == 0 && L.h3 % 64 == 0 && L.h2 <= 512 &&
           L.h3 <= 256 && L.h3 + 16 <= kTmemCols && L.h3 / 64 <= kASlots && L.h3 <= L.h2 &&
           (L.h2 <= 256 || L.h2 % 256 == 0);
}

// Forward of all four layers + loss + d loss / d eps_hat.  xin: [B][64] bf16 prepared rows; a_l / d_l: [B][h_l] bf16
// activations and Mish derivatives (row-major); deps: [B][64] bf16; loss_out += sum((eps_hat - noise)^2) * inv_count.
int actor_train_chain_fwd(const ActorLayout& L, const void* packed, const void* xin, const int64_t* t, const float* noise,
                          float inv_count, float* loss_out, void* a0, void* d0, void* a1, void* d1, void* a2, void* d2,
                          void* deps, long B, cudaStream_t st) {
    if (!actor_train_chain_shape_ok(L)) DDP_FAIL(DDP_ERR_UNSUPPORTED, "fused training forward does not support this shape");
    const uint8_t* pb = (const uint8_t*)packed;
    const float* pk = (const float*)packed;
    FcArgs a{};
    a.w0frag = (const uint2*)(pb + L.tc_w0);
    a.w3img = (const uint4*)(pb + L.tc_w3);
    a.tb0 = pk + L.tb0; a.b1 = pk + L.b1; a.b2 = pk + L.b2; a.b3 = pk + L.b3;
    a.xin = (const bf16*)xin; a.t = t; a.noise = noise; a.loss_out = loss_out; a.deps = (bf16*)deps;
    a.inv_count = inv_count; a.B = B;
    a.A = L.A; a.T = L.T; a.h1 = L.h1; a.h2 = L.h2; a.h3 = L.h3;
    a.nparts1 = L.h2 > 256 ? L.h2 / 256 : 1;
    a.part1 = L.h2 / a.nparts1;
    a.num_tiles = (int)((B + kRows - 1) / kRows);
    if (a.num_tiles == 0) return DDP_OK;
    FcMaps maps;
    int bad = 0;
    bad |= make_tmap_bf16_sw128(&maps.w1, pb + L.tc_w1, L.h2, L.h1, a.part1);
    bad |= make_tmap_bf16_sw128(&maps.w2, pb + L.tc_w2, L.h3, L.h2, L.h3);
    void* outs[2][3] = {{a0, a1, a2}, {d0, d1, d2}};
    const int widths[3] = {L.h1, L.h2, L.h3};
    for (int l = 0; l < 3; ++l) {
        bad |= make_tmap_bf16_sw128(&maps.a[l], outs[0][l], (uint64_t)B, widths[l], kRows);     // rows past B are clipped
        bad |= make_tmap_bf16_sw128(&maps.d[l], outs[1][l], (uint64_t)B, widths[l], kRows);
    }
    if (bad) DDP_FAIL(DDP_ERR_CUDA, "cuTensorMapEncodeTiled failed for the fused training forward");
    int dev = 0, sms = 0;
    DDP_CUDA_CHECK(cudaGetDevice(&dev));
    DDP_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t smem = SF::total + 1024;
    DDP_CUDA_CHECK(cudaFuncSetAttribute(actor_train_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = a.num_tiles < sms ? a.num_tiles : sms;
    actor_train_chain_kernel<<<grid, kThreads, smem, st>>>(maps, a);
    DDP_LAUNCH_CHECK("actor_train_chain_kernel");
    return DDP_OK;
}

}  // namespace ddp
