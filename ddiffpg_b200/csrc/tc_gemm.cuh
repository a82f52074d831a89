// Host interface of the two tcgen05 GEMM kernels the tensor-core paths of H2 and H3 are assembled from
// (csrc/tc_gemm.cu).  Both are persistent, warp-specialised (TMA producer / MMA issuer / epilogue warps),
// accumulate in TMEM (two 256-column accumulators, so the epilogue of one tile overlaps the MMAs of the next)
// and read every operand through TMA with the 128-byte swizzle.
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace ddp {
namespace tcg {

constexpr int kMaxGroups = 32;

// epilogue of the row GEMM  C[M x N] = epi(A[M x K] . W[N x K]^T)
enum RowEpi {
    EPI_MISH_FWD = 0,    // z = acc + bias (+ tbl[trow[r]]);  out_a = mish(z), out_d = mish'(z)        (bf16)
    EPI_ELU_FWD = 1,     // out_a = elu(acc + bias)                                                  (bf16)
    EPI_LINEAR_F32 = 2,  // out_f[r][c] = acc + bias, c < n_valid                                    (fp32)
    EPI_MUL_D = 3,       // out_a = acc * aux[r][c]                      (aux = stored mish', bf16)  (bf16)
    EPI_MUL_ELU_D = 4,   // out_a = acc * (aux > 0 ? 1 : aux + 1)        (aux = stored ELU activation) (bf16)
    EPI_MSE_HEAD = 5     // head of the denoiser: e = acc + bias - target[r][c] (c < n_valid): *loss += scale * sum e^2,
                         // out_a[r][0..63] = 2 * scale * e, zero padded  (mse_loss forward + d loss / d eps_hat; N = 16)
};

// Rows are split into `n_groups` contiguous segments; group g uses weight matrix g (and bias block g).
struct RowGroups {
    int n_groups;
    long off[kMaxGroups + 1];
};

struct RowGemm {
    const __nv_bfloat16* A; int lda;          // [M][lda], K valid columns (lda * 2 bytes multiple of 16)
    const __nv_bfloat16* W; int ldw;          // group g: W + g * w_stride, [N][ldw] (K valid columns)
    size_t w_stride;                          // elements between the weight matrices of consecutive groups
    long M; int N, K;
    int epi;
    const float* bias; size_t bias_stride;    // [N] per group (may be NULL)
    const float* tbl; const int64_t* trow; int tbl_ld; int tbl_rows;    // optional per-row additive table
    const __nv_bfloat16* aux; int aux_ld;
    __nv_bfloat16* out_a; __nv_bfloat16* out_d; int out_ld;
    float* out_f; int outf_ld; int n_valid;
    const float* target; int target_ld; float scale; float* loss;      // EPI_MSE_HEAD
    RowGroups groups;
};
int launch_row_gemm(const RowGemm& g, cudaStream_t st);

// dW[n][colmap(k)] += sum_r dZ[r][n] * X[r][k]   (fp32 atomics; rows split over CTAs)
struct DwGemm {
    const __nv_bfloat16* dZ; int ldz; int N;  // [R][ldz], N valid columns
    const __nv_bfloat16* X; int ldx; int K;   // [R][ldx], K valid columns
    long R;
    float* C; int ldc;                        // output rows n, columns colmap[k] (or k)
    const int* colmap;                        // device array [K] (-1 = drop) or NULL
    float* C2; int ldc2;                      // optional second output: colmap[k] >= kDwCol2 goes to C2[n][colmap[k] - kDwCol2]
    float* colsum;                            // optional: colsum[n] += sum_r dZ[r][n] (bias gradient), from one extra
                                              // N = 16 MMA per reduction step against a shared-memory tile of ones
};
constexpr int kDwCol2 = 1 << 20;
int launch_dw_gemm(const DwGemm& g, cudaStream_t st);

}  // namespace tcg
}  // namespace ddp
