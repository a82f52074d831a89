// tcgen05 GEMM kernels for the tensor-core paths of H2 (Q-ascent) and H3 (denoiser training step).
//
//  row GEMM   C[M x N] = epi(A[M x K] . W[N x K]^T)     one nn.Linear (forward) or dX = dY . W (backward) over
//             a 128-row tile, with the activation / derivative / masking fused into the TMEM epilogue
//             (reference layers: ddiffpg/models/diffusion_mlp.py:50-58, ddiffpg/models/mlp.py:13-35)
//  dW GEMM    dW[n][k] += sum_r dZ[r][n] * X[r][k]      the weight gradient autograd produces for
//             objective.backward() (ddiffpg/algo/ac_base.py:85); both operands are read MN-major straight from
//             the row-major activations, rows split over CTAs, fp32 atomics into the flat gradient
#include "tc_gemm.cuh"
#include "tc_common.cuh"

namespace ddp {
namespace tcg {
using namespace tc;

namespace {

constexpr int kStages = 4;
constexpr int kStageA = 128 * 128;          // 128 rows x 64 bf16
constexpr int kStageW = 256 * 128;          // up to 256 rows x 64 bf16
constexpr int kStageBytes = kStageA + kStageW;
#ifndef DDP_ROW_EPI_WARPS
#define DDP_ROW_EPI_WARPS 8
#endif
constexpr int kRowEpiWarps = DDP_ROW_EPI_WARPS;    // 8 or 16: warps of a lane quarter split the 16-column pieces
constexpr int kThreads = 64 + 32 * kRowEpiWarps;   // warp 0 TMA, warp 1 MMA, then the epilogue warps (row GEMM)
constexpr int kDwThreads = 192;             // dW GEMM: warps 2-5 epilogue
constexpr uint32_t kOffBars = kStages * kStageBytes;
constexpr uint32_t kOffOnes = kOffBars + 256;                  // dW GEMM: [64 reduction rows][64] bf16 ones (8 KB)
constexpr uint32_t kSmemBytes = kOffOnes + 8192 + 1024;
// row GEMM: one stage less, the space is the per-warp staging that turns the row-per-lane epilogue into
// coalesced global accesses (two [32 rows][128 B] buffers per epilogue warp)
constexpr int kRowStages = 3;
constexpr uint32_t kRowOffBars = kRowStages * kStageBytes;
constexpr uint32_t kRowOffStage = (kRowOffBars + 256 + 1023) / 1024 * 1024;      // 1024-aligned: also holds TMA boxes
constexpr uint32_t kStageWarpBytes = 2 * 4096;
constexpr uint32_t kRowSmemBytes = kRowOffStage + kRowEpiWarps * kStageWarpBytes + 1024;
static_assert(kRowSmemBytes <= 227 * 1024, "row GEMM shared memory");
// Backward epilogues with ONE row group (H3, N1, N4): the aux operand (stored derivative / activation) of a tile arrives
// by TMA into the staging area and the masked gradient leaves from the same bytes by TMA store -- two [128 x 128] bf16
// halves of the 256-column tile (four 16 KB SWIZZLE_128B boxes), so the load of the next tile's half overlaps the
// epilogue of the other half.  No thread issues a row-per-lane global access any more.
constexpr uint32_t kAuxHalfBytes = 2 * 128 * 128;
static_assert(2 * kAuxHalfBytes <= kRowEpiWarps * kStageWarpBytes, "aux halves live in the forward staging area");

struct WMaps { CUtensorMap m[kMaxGroups]; };
struct AuxMaps { CUtensorMap in, out; };

struct RowKernelArgs {
    RowGemm g;
    long tile0[kMaxGroups + 1];    // first m-tile of each group
    int n_tiles_n, BN;
    long total_tiles;
    int tma_aux;                   // 1: backward epilogue through the TMA-staged aux tile (see kAuxHalfBytes)
    int tma_fwd;                   // 1: ELU forward epilogue written as SWIZZLE_128B boxes and shipped by TMA store
};

// 16-bit (bf16) row-major matrix [rows][cols valid] with leading dimension ld; box = [box_rows][64 columns],
// 128-byte swizzle, out-of-bounds elements read as zero
int make_tmap(CUtensorMap* map, const void* gptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return -1;
        fn = (PFN_encodeTiled)p;
    }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(gptr), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

__device__ __forceinline__ void store_bf16x16(__nv_bfloat16* p, const float (&v)[16]) {
    uint4 a, b;
    a.x = pack_bf16x2(v[0], v[1]); a.y = pack_bf16x2(v[2], v[3]); a.z = pack_bf16x2(v[4], v[5]); a.w = pack_bf16x2(v[6], v[7]);
    b.x = pack_bf16x2(v[8], v[9]); b.y = pack_bf16x2(v[10], v[11]); b.z = pack_bf16x2(v[12], v[13]); b.w = pack_bf16x2(v[14], v[15]);
    reinterpret_cast<uint4*>(p)[0] = a;
    reinterpret_cast<uint4*>(p)[1] = b;
}
// ------------------------------------------------------------------------------------------ row GEMM
__global__ void __launch_bounds__(kThreads, 1)
row_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ WMaps maps_w,
                const __grid_constant__ AuxMaps maps_x, const RowKernelArgs k) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    const uint32_t bars = base + kRowOffBars;
    constexpr int kStages = kRowStages;
    auto bar_full = [&](int i) { return bars + 8 * i; };
    auto bar_empty = [&](int i) { return bars + 8 * (kStages + i); };
    auto bar_acc_full = [&](int i) { return bars + 8 * (2 * kStages + i); };
    auto bar_acc_empty = [&](int i) { return bars + 8 * (2 * kStages + 2 + i); };
    const uint32_t tmem_slot = bars + 8 * (2 * kStages + 4);
    auto bar_aux_full = [&](int i) { return bars + 8 * (2 * kStages + 5 + i); };
    auto bar_aux_empty = [&](int i) { return bars + 8 * (2 * kStages + 7 + i); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const RowGemm& g = k.g;
    const int kchunks = (g.K + 63) >> 6;
    const int halves = (k.BN + 127) >> 7;          // 128-column halves of a tile (TMA-staged backward epilogue)

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(bar_full(i), 1); mbar_init(bar_empty(i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_acc_full(i), 1); mbar_init(bar_acc_empty(i), kRowEpiWarps); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_aux_full(i), 1); mbar_init(bar_aux_empty(i), 1); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));

    // tile -> (group, first row, row limit, first column)
    auto decode = [&](long tile, int& grp, long& row0, long& row_end, int& n0) {
        const long mt = tile / k.n_tiles_n;
        n0 = (int)(tile % k.n_tiles_n) * k.BN;
        grp = 0;
        while (grp + 1 < g.groups.n_groups && mt >= k.tile0[grp + 1]) ++grp;
        row0 = g.groups.off[grp] + (mt - k.tile0[grp]) * 128;
        row_end = g.groups.off[grp + 1];
    };

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&map_a);
            if (k.tma_aux) { tma_prefetch_desc(&maps_x.in); tma_prefetch_desc(&maps_x.out); }
            if (k.tma_fwd) tma_prefetch_desc(&maps_x.out);
            int st = 0; uint32_t ph = 0, it = 0;
            for (long tile = blockIdx.x; tile < k.total_tiles; tile += gridDim.x, ++it) {
                int grp, n0; long row0, row_end;
                decode(tile, grp, row0, row_end, n0);
                // aux halves of this tile: requested as soon as the epilogue has shipped the previous tile's half out of
                // the buffer -- polled between the operand stages, so that neither stream waits for the other
                int aux_next = k.tma_aux ? 0 : halves;
                auto aux_load = [&](bool block) {
                    while (aux_next < halves) {
                        const int h = aux_next;
                        if (block) mbar_wait(bar_aux_empty(h), (it & 1) ^ 1);
                        else if (!mbar_test_wait(bar_aux_empty(h), (it & 1) ^ 1)) return;
                        const int c0 = n0 + h * 128;
                        const int boxes = (g.N - c0 + 63) / 64 < 2 ? (g.N - c0 + 63) / 64 : 2;
                        mbar_expect_tx(bar_aux_full(h), (uint32_t)boxes * 128u * 128u);
                        for (int b = 0; b < boxes; ++b)
                            tma_load_2d(base + kRowOffStage + h * kAuxHalfBytes + b * 16384, &maps_x.in, bar_aux_full(h),
                                        c0 + b * 64, (int)row0);
                        ++aux_next;
                    }
                };
                for (int c = 0; c < kchunks; ++c) {
                    aux_load(false);
                    mbar_wait(bar_empty(st), ph ^ 1);
                    mbar_expect_tx(bar_full(st), (uint32_t)(kStageA + k.BN * 128));
                    tma_load_2d(base + st * kStageBytes, &map_a, bar_full(st), c * 64, (int)row0);
                    tma_load_2d(base + st * kStageBytes + kStageA, &maps_w.m[grp], bar_full(st), c * 64, n0);
                    if (++st == kStages) { st = 0; ph ^= 1; }
                }
                aux_load(true);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_16(128, k.BN, false);
            int st = 0; uint32_t ph = 0, it = 0;
            for (long tile = blockIdx.x; tile < k.total_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                mbar_wait(bar_acc_empty(buf), ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int c = 0; c < kchunks; ++c) {
                    mbar_wait(bar_full(st), ph);
                    tc_fence_after();
                    const uint64_t ad = make_smem_desc_sw128(base + st * kStageBytes);
                    const uint64_t bd = make_smem_desc_sw128(base + st * kStageBytes + kStageA);
#pragma unroll
                    for (int q = 0; q < 4; ++q) umma_bf16(tmem_base + buf * 256, ad + 2 * q, bd + 2 * q, idesc, (c | q) != 0);
                    umma_commit(bar_empty(st));
                    if (++st == kStages) { st = 0; ph ^= 1; }
                }
                umma_commit(bar_acc_full(buf));
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue: thread = one row of the tile
        const int q = warp & 3;
        uint32_t it = 0, hb = 0;
        for (long tile = blockIdx.x; tile < k.total_tiles; tile += gridDim.x, ++it) {
            int grp, n0; long row0, row_end;
            decode(tile, grp, row0, row_end, n0);
            const int buf = it & 1;
            const long row = row0 + q * 32 + lane;
            const bool valid = row < row_end;
            const float* bias = g.bias ? g.bias + (size_t)grp * g.bias_stride : nullptr;
            const float* trow = nullptr;
            if (g.tbl && valid) {
                long t = g.trow[row];
                t = t < 0 ? 0 : (t >= g.tbl_rows ? g.tbl_rows - 1 : t);
                trow = g.tbl + t * g.tbl_ld;
            }
            const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256;
            const int pieces = k.BN >> 4;
            // Staging: the thread that owns a row writes its 16-column pieces (32 B) into a swizzled [32][128 B]
            // buffer; the warp then moves the buffer to global memory 8 lanes per 128-byte row segment, so a store
            // instruction touches 4 cache lines instead of 32 (L1 LSU wavefronts: 67 % -> 49 % of peak, forward layers
            // 144 / 93 / 48 -> 132 / 78 / 44 us at 65 536 rows).  Forward epilogues only.
            uint8_t* stg_a = smem + kRowOffStage + (warp - 2) * kStageWarpBytes;
            uint8_t* stg_d = stg_a + 4096;
            auto stg_off = [](int r, int ch) { return r * 128 + ((ch ^ (r & 7)) << 4); };
            const long wrow0 = row0 + q * 32;
            const bool staged = g.epi == EPI_MISH_FWD || g.epi == EPI_ELU_FWD;
            // piece j of the current group: accumulator columns -> this lane's row of the staging buffers
            auto process = [&](const uint32_t (&cur)[16], int p, int j) {
                const int c0 = n0 + p * 16;
                float z[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) z[i] = __uint_as_float(cur[i]);
                if (bias && c0 < g.N) {
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c0) + i4);
                        z[4 * i4] += b.x; z[4 * i4 + 1] += b.y; z[4 * i4 + 2] += b.z; z[4 * i4 + 3] += b.w;
                    }
                }
                if (trow && c0 < g.N) {
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(trow + c0) + i4);
                        z[4 * i4] += b.x; z[4 * i4 + 1] += b.y; z[4 * i4 + 2] += b.z; z[4 * i4 + 3] += b.w;
                    }
                }
                auto put = [&](uint8_t* stg, const float (&v)[16]) {
                    uint4 a, b;
                    a.x = pack_bf16x2(v[0], v[1]); a.y = pack_bf16x2(v[2], v[3]); a.z = pack_bf16x2(v[4], v[5]); a.w = pack_bf16x2(v[6], v[7]);
                    b.x = pack_bf16x2(v[8], v[9]); b.y = pack_bf16x2(v[10], v[11]); b.z = pack_bf16x2(v[12], v[13]); b.w = pack_bf16x2(v[14], v[15]);
                    *reinterpret_cast<uint4*>(stg + stg_off(lane, 2 * j)) = a;
                    *reinterpret_cast<uint4*>(stg + stg_off(lane, 2 * j + 1)) = b;
                };
                if (g.epi == EPI_MISH_FWD) {
                    float d[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        // e = exp(z), s = e + 1, p = s^2 + 1, R = 1/p:  tanh(softplus(z)) = w = 1 - 2R  and
                        // mish'(z) = w + z * sigmoid(z) * (1 - w^2) = w + 4 z s e R^2   (sigmoid = e/s, 1 - w^2 = 4 s^2 R^2)
                        const float e = ex2_approx(fminf(z[i] * 1.4426950408889634f, 28.853900817779268f));
                        const float s = e + 1.f;
                        const float R = rcp_approx(fmaf(s, s, 1.f));
                        const float w = fmaf(-2.f, R, 1.f);
                        d[i] = fmaf(4.f * z[i], (s * e) * R * R, w);
                        z[i] *= w;
                    }
                    put(stg_a, z);
                    put(stg_d, d);
                } else if (g.epi == EPI_ELU_FWD) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) z[i] = z[i] > 0.f ? z[i] : ex2_approx(z[i] * 1.4426950408889634f) - 1.f;
                    put(stg_a, z);
                } else if (g.epi == EPI_LINEAR_F32) {
                    if (valid && c0 < g.N) {
                        float* o = g.out_f + row * g.outf_ld + c0;
                        if (c0 + 16 <= g.n_valid && (g.outf_ld & 3) == 0 && ((uintptr_t)g.out_f & 15) == 0) {
#pragma unroll
                            for (int i4 = 0; i4 < 4; ++i4)
                                reinterpret_cast<float4*>(o)[i4] = make_float4(z[4 * i4], z[4 * i4 + 1], z[4 * i4 + 2], z[4 * i4 + 3]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (c0 + i < g.n_valid) o[i] = z[i];
                        }
                    }
                } else if (g.epi == EPI_MSE_HEAD) {
                    // one 16-column piece per tile: squared error of the row, its gradient as a zero-padded 64-column
                    // bf16 row (the A operand of the first backward GEMM), one loss atomic per warp
                    float sq = 0.f, d[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float e = 0.f;
                        if (valid && c0 + i < g.n_valid) e = z[i] - __ldg(g.target + row * g.target_ld + c0 + i);
                        sq = fmaf(e, e, sq);
                        d[i] = 2.f * g.scale * e;
                    }
                    if (valid && c0 == 0) {
                        uint4* o = reinterpret_cast<uint4*>(g.out_a + row * g.out_ld);
                        o[0] = make_uint4(pack_bf16x2(d[0], d[1]), pack_bf16x2(d[2], d[3]), pack_bf16x2(d[4], d[5]), pack_bf16x2(d[6], d[7]));
                        o[1] = make_uint4(pack_bf16x2(d[8], d[9]), pack_bf16x2(d[10], d[11]), pack_bf16x2(d[12], d[13]), pack_bf16x2(d[14], d[15]));
#pragma unroll
                        for (int i = 2; i < 8; ++i) o[i] = make_uint4(0u, 0u, 0u, 0u);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                    if (lane == 0 && sq != 0.f) atomicAdd(g.loss, sq * g.scale);
                }
            };
            // backward epilogues: the row-per-lane accesses go straight to global memory (staging them was measured
            // slower), but every aux operand of the warp's pieces is requested BEFORE the accumulator is waited for:
            // one exposed memory latency per tile instead of one per 16-column piece
            auto process_bwd = [&](const uint32_t (&cur)[16], int p, const uint4& a0, const uint4& a1) {
                const int c0 = n0 + p * 16;
                if (!valid || c0 >= g.N) return;
                const uint32_t w[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                float z[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float lo = __uint_as_float(w[i] << 16), hi = __uint_as_float(w[i] & 0xffff0000u);
                    if (g.epi == EPI_MUL_D) {
                        z[2 * i] = __uint_as_float(cur[2 * i]) * lo;
                        z[2 * i + 1] = __uint_as_float(cur[2 * i + 1]) * hi;
                    } else {
                        z[2 * i] = __uint_as_float(cur[2 * i]) * (lo > 0.f ? 1.f : lo + 1.f);
                        z[2 * i + 1] = __uint_as_float(cur[2 * i + 1]) * (hi > 0.f ? 1.f : hi + 1.f);
                    }
                }
                store_bf16x16(g.out_a + row * g.out_ld + c0, z);
            };
            // the warps of a lane quarter split the 16-column pieces of the tile
            constexpr int kSplit = kRowEpiWarps / 4;
            const int part = (warp - 2) >> 2;
            const int p_begin = (pieces * part + kSplit - 1) / kSplit, p_end = (pieces * (part + 1) + kSplit - 1) / kSplit;
            uint32_t va[16], vb[16];
            const bool needs_aux = g.epi == EPI_MUL_D || g.epi == EPI_MUL_ELU_D;
            static_assert(256 / 16 / (kRowEpiWarps / 4) <= 8, "a warp owns at most 8 pieces of a tile");
            if (needs_aux && k.tma_aux) {
                // thread = one row; the warps of a lane quarter split the 16-column pieces of each 128-column half
                const int r = q * 32 + lane;
                mbar_wait(bar_acc_full(buf), (it >> 1) & 1);
                tc_fence_after();
                for (int h = 0; h < halves; ++h) {
                    const int hp = (k.BN - h * 128 < 128 ? k.BN - h * 128 : 128) >> 4;      // pieces of this half
                    const int pb = (hp * part + kSplit - 1) / kSplit, pe = (hp * (part + 1) + kSplit - 1) / kSplit;
                    uint8_t* stg = smem + kRowOffStage + h * kAuxHalfBytes;
                    if (pb < pe) tmem_ld16(tb + (h * 8 + pb) * 16, va);
                    mbar_wait(bar_aux_full(h), it & 1);
                    for (int p = pb; p < pe; ++p) {
                        tmem_ld_wait();
                        uint32_t cur[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) cur[i] = va[i];
                        if (p + 1 < pe) tmem_ld16(tb + (h * 8 + p + 1) * 16, va);
                        uint8_t* bx = stg + (p >> 2) * 16384;
                        uint4* s0 = reinterpret_cast<uint4*>(bx + sw128_offset(r, (p & 3) * 16));
                        uint4* s1 = reinterpret_cast<uint4*>(bx + sw128_offset(r, (p & 3) * 16 + 8));
                        const uint4 a0 = *s0, a1 = *s1;
                        const uint32_t w[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                        float z[16];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float lo = __uint_as_float(w[i] << 16), hi = __uint_as_float(w[i] & 0xffff0000u);
                            if (g.epi == EPI_MUL_D) {
                                z[2 * i] = __uint_as_float(cur[2 * i]) * lo;
                                z[2 * i + 1] = __uint_as_float(cur[2 * i + 1]) * hi;
                            } else {
                                z[2 * i] = __uint_as_float(cur[2 * i]) * (lo > 0.f ? 1.f : lo + 1.f);
                                z[2 * i + 1] = __uint_as_float(cur[2 * i + 1]) * (hi > 0.f ? 1.f : hi + 1.f);
                            }
                        }
                        *s0 = make_uint4(pack_bf16x2(z[0], z[1]), pack_bf16x2(z[2], z[3]), pack_bf16x2(z[4], z[5]), pack_bf16x2(z[6], z[7]));
                        *s1 = make_uint4(pack_bf16x2(z[8], z[9]), pack_bf16x2(z[10], z[11]), pack_bf16x2(z[12], z[13]), pack_bf16x2(z[14], z[15]));
                    }
                    if (h == halves - 1) {
                        // every TMEM read of this tile is done: the MMA warp may start the tile after next in this buffer
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_acc_empty(buf));
                    }
                    fence_proxy_async();                                   // generic-proxy stores -> visible to the TMA store
                    asm volatile("bar.sync 1, %0;" ::"n"(32 * kRowEpiWarps) : "memory");
                    if (threadIdx.x == 64) {
                        const int c0 = n0 + h * 128;
                        for (int b = 0; b < 2 && c0 + b * 64 < g.N; ++b)
                            tma_store_2d(&maps_x.out, base + kRowOffStage + h * kAuxHalfBytes + b * 16384, c0 + b * 64, (int)row0);
                        tma_store_commit();
                        tma_store_wait_read();                             // the buffer may be refilled
                        mbar_arrive(bar_aux_empty(h));
                    }
                }
                continue;
            }
            if (needs_aux) {
                uint4 axq[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c0 = n0 + (p_begin + j) * 16;
                    axq[2 * j] = axq[2 * j + 1] = make_uint4(0u, 0u, 0u, 0u);
                    if (p_begin + j < p_end && valid && c0 < g.N) {
                        const uint4* ap = reinterpret_cast<const uint4*>(g.aux + row * g.aux_ld + c0);
                        axq[2 * j] = ap[0]; axq[2 * j + 1] = ap[1];
                    }
                }
                mbar_wait(bar_acc_full(buf), (it >> 1) & 1);
                tc_fence_after();
                if (p_begin < p_end) tmem_ld16(tb + p_begin * 16, va);
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const int p = p_begin + j;
                    if (p < p_end) {
                        tmem_ld_wait();
                        if (p + 1 < p_end) tmem_ld16(tb + (p + 1) * 16, vb);
                        process_bwd(va, p, axq[2 * j], axq[2 * j + 1]);
                        if (p + 1 < p_end) {
                            tmem_ld_wait();
                            if (p + 2 < p_end) tmem_ld16(tb + (p + 2) * 16, va);
                            process_bwd(vb, p + 1, axq[2 * j + 2], axq[2 * j + 3]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_acc_empty(buf));
                continue;
            }
            if (k.tma_fwd) {
                // ELU forward of a single row group: the thread that owns a row writes its packed pieces straight into
                // the box layout of a [128 x 128] half (two buffers, alternating), one thread ships the half by TMA
                // store -- no copy-out loop, no row-per-lane global store.  Buffer b is free again once the store
                // thread has passed its wait_group.read and everybody has met at the next named barrier.
                const int r = q * 32 + lane;
                mbar_wait(bar_acc_full(buf), (it >> 1) & 1);
                tc_fence_after();
                for (int h = 0; h < halves; ++h, ++hb) {
                    const int hp = (k.BN - h * 128 < 128 ? k.BN - h * 128 : 128) >> 4;
                    const int pb = (hp * part + kSplit - 1) / kSplit, pe = (hp * (part + 1) + kSplit - 1) / kSplit;
                    uint8_t* stg = smem + kRowOffStage + (hb & 1) * kAuxHalfBytes;
                    if (pb < pe) tmem_ld16(tb + (h * 8 + pb) * 16, va);
                    for (int p = pb; p < pe; ++p) {
                        tmem_ld_wait();
                        float z[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) z[i] = __uint_as_float(va[i]);
                        if (p + 1 < pe) tmem_ld16(tb + (h * 8 + p + 1) * 16, va);
                        const int c0 = n0 + (h * 8 + p) * 16;
                        if (bias && c0 < g.N) {
#pragma unroll
                            for (int i4 = 0; i4 < 4; ++i4) {
                                const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c0) + i4);
                                z[4 * i4] += b.x; z[4 * i4 + 1] += b.y; z[4 * i4 + 2] += b.z; z[4 * i4 + 3] += b.w;
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 16; ++i) z[i] = z[i] > 0.f ? z[i] : ex2_approx(z[i] * 1.4426950408889634f) - 1.f;
                        uint8_t* bx = stg + (p >> 2) * 16384;
                        *reinterpret_cast<uint4*>(bx + sw128_offset(r, (p & 3) * 16)) =
                            make_uint4(pack_bf16x2(z[0], z[1]), pack_bf16x2(z[2], z[3]), pack_bf16x2(z[4], z[5]), pack_bf16x2(z[6], z[7]));
                        *reinterpret_cast<uint4*>(bx + sw128_offset(r, (p & 3) * 16 + 8)) =
                            make_uint4(pack_bf16x2(z[8], z[9]), pack_bf16x2(z[10], z[11]), pack_bf16x2(z[12], z[13]), pack_bf16x2(z[14], z[15]));
                    }
                    if (h == halves - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_acc_empty(buf));
                    }
                    fence_proxy_async();
                    asm volatile("bar.sync 1, %0;" ::"n"(32 * kRowEpiWarps) : "memory");
                    if (threadIdx.x == 64) {
                        const int c0 = n0 + h * 128;
                        for (int b = 0; b < 2 && c0 + b * 64 < g.N; ++b)
                            tma_store_2d(&maps_x.out, base + kRowOffStage + (hb & 1) * kAuxHalfBytes + b * 16384, c0 + b * 64, (int)row0);
                        tma_store_commit();
                        tma_store_wait_read();
                    }
                }
                continue;
            }
            mbar_wait(bar_acc_full(buf), (it >> 1) & 1);
            tc_fence_after();
            if (p_begin < p_end) tmem_ld16(tb + p_begin * 16, va);
            for (int pg = p_begin; pg < p_end; pg += 4) {
                const int np = p_end - pg < 4 ? p_end - pg : 4;
                const int cpr = 2 * np;                              // 16-byte chunks per row of this group
                const int cg0 = n0 + pg * 16;
                for (int j = 0; j < np; j += 2) {
                    const int p = pg + j;
                    tmem_ld_wait();
                    if (p + 1 < p_end) tmem_ld16(tb + (p + 1) * 16, vb);
                    process(va, p, j);
                    if (j + 1 < np) {
                        tmem_ld_wait();
                        if (p + 2 < p_end) tmem_ld16(tb + (p + 2) * 16, va);
                        process(vb, p + 1, j + 1);
                    }
                }
                if (staged) {
                    __syncwarp();
                    for (int idx = lane; idx < 32 * cpr; idx += 32) {
                        const int r = idx / cpr, ch = idx - r * cpr;
                        if (wrow0 + r < row_end && cg0 + ch * 8 < g.N) {
                            const size_t o = (size_t)(wrow0 + r) * g.out_ld + cg0 + ch * 8;
                            *reinterpret_cast<uint4*>(g.out_a + o) = *reinterpret_cast<const uint4*>(stg_a + stg_off(r, ch));
                            if (g.epi == EPI_MISH_FWD)
                                *reinterpret_cast<uint4*>(g.out_d + o) = *reinterpret_cast<const uint4*>(stg_d + stg_off(r, ch));
                        }
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty(buf));
        }
    }
    if (threadIdx.x == 64 && (k.tma_aux || k.tma_fwd)) tma_store_wait_all();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------ dW GEMM
struct DwKernelArgs {
    DwGemm g;
    int BN;                 // X columns per tile (64..256, multiple of 64)
    int tiles_n, tiles_k;   // output tiles along dZ columns (128 each) and X columns (BN each)
    long rows_per_split;
};

// MN-major SWIZZLE_128B operand: 64-element (128 B) rows per reduction index, 64-element blocks along M/N
// `lbo` bytes apart, 8-row groups 1024 B apart (cute: ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO)))
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(kDwThreads, 1)
dw_gemm_kernel(const __grid_constant__ CUtensorMap map_z, const __grid_constant__ CUtensorMap map_x, const DwKernelArgs k) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    const uint32_t bars = base + kOffBars;
    auto bar_full = [&](int i) { return bars + 8 * i; };
    auto bar_empty = [&](int i) { return bars + 8 * (kStages + i); };
    const uint32_t bar_done = bars + 8 * (2 * kStages);
    const uint32_t tmem_slot = bars + 8 * (2 * kStages + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const DwGemm& g = k.g;

    const int tile = blockIdx.x % (k.tiles_n * k.tiles_k), split = blockIdx.x / (k.tiles_n * k.tiles_k);
    const int n0 = (tile % k.tiles_n) * 128, k0 = (tile / k.tiles_n) * k.BN;
    const long r_begin = (long)split * k.rows_per_split;
    const long r_end = r_begin + k.rows_per_split < g.R ? r_begin + k.rows_per_split : g.R;
    const int chunks = r_end > r_begin ? (int)((r_end - r_begin + 63) >> 6) : 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(bar_full(i), 1); mbar_init(bar_empty(i), 1); }
        mbar_init(bar_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    // bias gradient: only the CTAs of the first X-column tile carry it
    const bool do_colsum = g.colsum != nullptr && k0 == 0;
    if (do_colsum) {
        uint4* ones = reinterpret_cast<uint4*>(smem + kOffOnes);
        for (int i = threadIdx.x; i < 8192 / 16; i += kDwThreads) ones[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));
    const int xboxes = k.BN >> 6;

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&map_z);
            tma_prefetch_desc(&map_x);
            int st = 0; uint32_t ph = 0;
            for (int c = 0; c < chunks; ++c) {
                mbar_wait(bar_empty(st), ph ^ 1);
                mbar_expect_tx(bar_full(st), (uint32_t)((2 + xboxes) * 8192));
                const int r = (int)(r_begin + (long)c * 64);
                // rows past r_end inside the last chunk belong to the next split: they are read here too, so
                // splits are kept multiples of 64 rows on the host (only the global tail is zero-filled)
                tma_load_2d(base + st * kStageBytes, &map_z, bar_full(st), n0, r);
                tma_load_2d(base + st * kStageBytes + 8192, &map_z, bar_full(st), n0 + 64, r);
                for (int b = 0; b < xboxes; ++b)
                    tma_load_2d(base + st * kStageBytes + kStageA + b * 8192, &map_x, bar_full(st), k0 + b * 64, r);
                if (++st == kStages) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // both operands MN-major: bits 15 / 16 of the instruction descriptor
            const uint32_t idesc = make_idesc_16(128, k.BN, false) | (1u << 15) | (1u << 16);
            const uint32_t idesc1 = make_idesc_16(128, 16, false) | (1u << 15) | (1u << 16);
            int st = 0; uint32_t ph = 0;
            for (int c = 0; c < chunks; ++c) {
                mbar_wait(bar_full(st), ph);
                tc_fence_after();
                const uint32_t a0 = base + st * kStageBytes, b0 = a0 + kStageA;
#pragma unroll
                for (int q = 0; q < 4; ++q)       // 16 reduction rows per instruction = 2 KB
                    umma_bf16(tmem_base, make_smem_desc_mn_sw128(a0 + q * 2048, 8192),
                              make_smem_desc_mn_sw128(b0 + q * 2048, 8192), idesc, (c | q) != 0);
                if (do_colsum) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)   // D2[n][0..15] += sum over these 16 rows of dZ[r][n] * 1
                        umma_bf16(tmem_base + 256, make_smem_desc_mn_sw128(a0 + q * 2048, 8192),
                                  make_smem_desc_mn_sw128(base + kOffOnes + q * 2048, 8192), idesc1, (c | q) != 0);
                }
                umma_commit(bar_empty(st));
                if (++st == kStages) { st = 0; ph ^= 1; }
            }
            umma_commit(bar_done);
        }
    } else if (chunks > 0) {
        const int q = warp & 3;
        const int n = n0 + q * 32 + lane;
        mbar_wait(bar_done, 0);
        tc_fence_after();
        const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16);
        if (do_colsum) {
            uint32_t v[16];
            tmem_ld16(tb + 256, v);
            tmem_ld_wait();
            if (n < g.N) atomicAdd(g.colsum + n, __uint_as_float(v[0]));
        }
        // The accumulator arrives one output row (n) per lane; the atomics want one output row per INSTRUCTION (32 lanes
        // on 32 consecutive k): each 32 x 32 block goes through a swizzled shared-memory transpose (the pipeline stages
        // are idle by now), so a RED touches one 128-byte line instead of 32.
        uint8_t* tr = smem + (warp - 2) * 4096;                      // [32 rows][128 B], 16-byte chunks XOR-swizzled by row
        for (int p = 0; p < (k.BN >> 4); p += 2) {
            uint32_t v[32];
            tmem_ld16(tb + p * 16, reinterpret_cast<uint32_t(&)[16]>(v[0]));
            tmem_ld16(tb + p * 16 + 16, reinterpret_cast<uint32_t(&)[16]>(v[16]));
            tmem_ld_wait();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint4*>(tr + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            __syncwarp();
            const int kk = k0 + p * 16 + lane;
            int col = -1;
            if (kk < g.K) col = g.colmap ? g.colmap[kk] : kk;
            float* dst = col >= kDwCol2 ? g.C2 + (col - kDwCol2) : g.C + col;
            const size_t ld = col >= kDwCol2 ? (size_t)g.ldc2 : (size_t)g.ldc;
            const int nrows = g.N - (n0 + q * 32) < 32 ? g.N - (n0 + q * 32) : 32;
            if (col >= 0) {
                for (int r = 0; r < nrows; ++r) {
                    const float x = *reinterpret_cast<const float*>(tr + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2)));
                    atomicAdd(dst + (size_t)(n0 + q * 32 + r) * ld, x);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace

static bool g_row_no_tma_aux = false;       // ddp_debug_row_gemm_direct_aux: A/B switch for the TMA-staged epilogues (tests, measurements)

int launch_row_gemm(const RowGemm& g, cudaStream_t st) {
    if (g.M <= 0) return DDP_OK;
    if (g.N <= 0 || g.K <= 0 || g.lda % 8 || g.ldw % 8) DDP_FAIL(DDP_ERR_SHAPE, "row gemm: bad shape (N=%d K=%d lda=%d ldw=%d)", g.N, g.K, g.lda, g.ldw);
    if (g.groups.n_groups < 1 || g.groups.n_groups > kMaxGroups) DDP_FAIL(DDP_ERR_SHAPE, "row gemm: bad group count");
    RowKernelArgs k;
    k.g = g;
    const int n16 = (g.N + 15) / 16 * 16;
    k.BN = n16 < 256 ? n16 : 256;
    k.n_tiles_n = (g.N + k.BN - 1) / k.BN;
    long tiles = 0;
    for (int i = 0; i < g.groups.n_groups; ++i) {
        k.tile0[i] = tiles;
        tiles += (g.groups.off[i + 1] - g.groups.off[i] + 127) / 128;
    }
    k.tile0[g.groups.n_groups] = tiles;
    k.total_tiles = tiles * k.n_tiles_n;
    if (k.total_tiles == 0) return DDP_OK;
    CUtensorMap ma;
    WMaps mw;
    if (make_tmap(&ma, g.A, (uint64_t)g.M, (uint64_t)g.K, (uint64_t)g.lda, 128) != 0)
        DDP_FAIL(DDP_ERR_CUDA, "row gemm: tensor map (A) failed");
    for (int i = 0; i < g.groups.n_groups; ++i)
        if (make_tmap(&mw.m[i], g.W + (size_t)i * g.w_stride, (uint64_t)g.N, (uint64_t)g.K, (uint64_t)g.ldw, (uint32_t)k.BN) != 0)
            DDP_FAIL(DDP_ERR_CUDA, "row gemm: tensor map (W) failed");
    for (int i = g.groups.n_groups; i < kMaxGroups; ++i) mw.m[i] = mw.m[0];
    // backward epilogues of a single row group go through the TMA-staged aux tile (a box written past the end of a group
    // would land in the next group's rows; the tensor map only clips at M)
    AuxMaps mx;
    mx.in = ma; mx.out = ma;
    const bool needs_aux = g.epi == EPI_MUL_D || g.epi == EPI_MUL_ELU_D;
    k.tma_aux = (needs_aux && !g_row_no_tma_aux && g.groups.n_groups == 1 && g.N % 64 == 0 && g.aux_ld % 8 == 0 &&
                 g.out_ld % 8 == 0 && ((uintptr_t)g.aux & 15) == 0 && ((uintptr_t)g.out_a & 15) == 0) ? 1 : 0;
    k.tma_fwd = (g.epi == EPI_ELU_FWD && !g_row_no_tma_aux && g.groups.n_groups == 1 && g.N % 64 == 0 && g.out_ld % 8 == 0 &&
                 ((uintptr_t)g.out_a & 15) == 0 && !g.tbl) ? 1 : 0;
    if (k.tma_fwd && make_tmap(&mx.out, g.out_a, (uint64_t)g.M, (uint64_t)g.N, (uint64_t)g.out_ld, 128) != 0)
        DDP_FAIL(DDP_ERR_CUDA, "row gemm: tensor map (out) failed");
    if (k.tma_aux) {
        if (make_tmap(&mx.in, g.aux, (uint64_t)g.M, (uint64_t)g.N, (uint64_t)g.aux_ld, 128) != 0 ||
            make_tmap(&mx.out, g.out_a, (uint64_t)g.M, (uint64_t)g.N, (uint64_t)g.out_ld, 128) != 0)
            DDP_FAIL(DDP_ERR_CUDA, "row gemm: tensor map (aux) failed");
    }
    int dev = 0, sms = 0;
    DDP_CUDA_CHECK(cudaGetDevice(&dev));
    DDP_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DDP_CUDA_CHECK(cudaFuncSetAttribute(row_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowSmemBytes));
    const long grid = k.total_tiles < sms ? k.total_tiles : sms;
    row_gemm_kernel<<<(unsigned)grid, kThreads, kRowSmemBytes, st>>>(ma, mw, mx, k);
    DDP_LAUNCH_CHECK("row_gemm_kernel");
    return DDP_OK;
}

int launch_dw_gemm(const DwGemm& g, cudaStream_t st) {
    if (g.R <= 0) return DDP_OK;
    if (g.N <= 0 || g.K <= 0 || g.ldz % 8 || g.ldx % 8) DDP_FAIL(DDP_ERR_SHAPE, "dW gemm: bad shape");
    DwKernelArgs k;
    k.g = g;
    const int k64 = (g.K + 63) / 64 * 64;
    k.BN = k64 < 256 ? k64 : 256;
    k.tiles_n = (g.N + 127) / 128;
    k.tiles_k = (g.K + k.BN - 1) / k.BN;
    const long tiles = (long)k.tiles_n * k.tiles_k;
    const long chunks = (g.R + 63) / 64;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long splits = 2L * sms / tiles;                       // at most two full waves of CTAs (never a third, partial one)
    if (splits > chunks) splits = chunks;
    if (splits < 1) splits = 1;
    const long cps = (chunks + splits - 1) / splits;
    k.rows_per_split = cps * 64;
    splits = (chunks + cps - 1) / cps;
    CUtensorMap mz, mx;
    if (make_tmap(&mz, g.dZ, (uint64_t)g.R, (uint64_t)g.N, (uint64_t)g.ldz, 64) != 0 ||
        make_tmap(&mx, g.X, (uint64_t)g.R, (uint64_t)g.K, (uint64_t)g.ldx, 64) != 0)
        DDP_FAIL(DDP_ERR_CUDA, "dW gemm: tensor map failed");
    DDP_CUDA_CHECK(cudaFuncSetAttribute(dw_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    dw_gemm_kernel<<<(unsigned)(tiles * splits), kDwThreads, kSmemBytes, st>>>(mz, mx, k);
    DDP_LAUNCH_CHECK("dw_gemm_kernel");
    return DDP_OK;
}

}  // namespace tcg
}  // namespace ddp

// Debug entry point (not part of the public header): 1 = the round-1 epilogues (backward: direct row-per-lane global
// accesses; ELU forward: per-warp staging + copy-out loop) instead of the TMA-staged ones.  Process-wide; A/B
// measurements and tests only.
extern "C" void ddp_debug_row_gemm_direct_aux(int on) { ddp::tcg::g_row_no_tma_aux = on != 0; }

// Debug entry points (not part of the public header) used by tests/test_tc_gemm_gpu.py
extern "C" int ddp_debug_row_gemm(const void* A, int lda, const void* W, int ldw, long M, int N, int K, int epi,
                                  const float* bias, const void* aux, void* out_a, void* out_d, float* out_f,
                                  int n_groups, const long* group_off, size_t w_stride, size_t bias_stride,
                                  const float* tbl, const int64_t* trow, int tbl_ld, int tbl_rows, void* stream) {
    using namespace ddp::tcg;
    RowGemm g{};
    g.A = (const __nv_bfloat16*)A; g.lda = lda; g.W = (const __nv_bfloat16*)W; g.ldw = ldw; g.w_stride = w_stride;
    g.M = M; g.N = N; g.K = K; g.epi = epi; g.bias = bias; g.bias_stride = bias_stride;
    g.tbl = tbl; g.trow = trow; g.tbl_ld = tbl_ld; g.tbl_rows = tbl_rows;
    g.aux = (const __nv_bfloat16*)aux; g.aux_ld = N;
    g.out_a = (__nv_bfloat16*)out_a; g.out_d = (__nv_bfloat16*)out_d; g.out_ld = N;
    g.out_f = out_f; g.outf_ld = N; g.n_valid = N;
    g.groups.n_groups = n_groups;
    for (int i = 0; i <= n_groups; ++i) g.groups.off[i] = group_off[i];
    return launch_row_gemm(g, (cudaStream_t)stream);
}

// EPI_LINEAR_F32 with an explicit output leading dimension and valid-column count
extern "C" int ddp_debug_row_gemm_nvalid(const void* A, int lda, const void* W, int ldw, long M, int N, int K,
                                         const float* bias, float* out_f, int outf_ld, int n_valid, void* stream) {
    using namespace ddp::tcg;
    RowGemm g{};
    g.A = (const __nv_bfloat16*)A; g.lda = lda; g.W = (const __nv_bfloat16*)W; g.ldw = ldw;
    g.M = M; g.N = N; g.K = K; g.epi = EPI_LINEAR_F32; g.bias = bias;
    g.out_f = out_f; g.outf_ld = outf_ld; g.n_valid = n_valid;
    g.groups.n_groups = 1; g.groups.off[0] = 0; g.groups.off[1] = M;
    return launch_row_gemm(g, (cudaStream_t)stream);
}

extern "C" int ddp_debug_dw_gemm(const void* dZ, int ldz, int N, const void* X, int ldx, int K, long R, float* C, int ldc,
                                 const int* colmap, void* stream) {
    using namespace ddp::tcg;
    DwGemm g{};
    g.dZ = (const __nv_bfloat16*)dZ; g.ldz = ldz; g.N = N; g.X = (const __nv_bfloat16*)X; g.ldx = ldx; g.K = K;
    g.R = R; g.C = C; g.ldc = ldc; g.colmap = colmap;
    return launch_dw_gemm(g, (cudaStream_t)stream);
}
