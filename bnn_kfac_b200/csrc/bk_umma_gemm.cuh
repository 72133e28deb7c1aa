// bk_umma_gemm.cuh — host-side argument block for the tcgen05 "NT" contraction core.
//
//   D[b][m][n] = sum_k A[b][m][k] * B[b][n][k]        (both operands K-major bf16, fp32 accumulate)
//
// Every dense contraction on the Kronecker-factored Laplace path is phrased in this form
// (see DESIGN.md "Kernels"): factor SYRK (A == B, lower tiles + mirror), matrix-normal sampling
// (triangular A), Monte-Carlo forward (bias + ReLU epilogue), kron-free quadratic forms.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace bk {

enum GemmFlags : int {
  kSyrkLower = 1,   // A == B (square): only tiles touching the lower triangle, write c <= r only
  kMirror = 2,      // with kSyrkLower: also write the transposed element (full symmetric result)
  kTriA = 4,        // A[m][k] == 0 for k > m (lower-triangular A): skip those k-blocks
  kTriB = 8,        // B[n][k] == 0 for k > tri_koff + n
  kRelu = 16,       // max(x, 0) after alpha/beta/bias
  kTriBUpper = 32,  // B[n][k] == 0 for k < n (B is the transpose of a lower-triangular matrix)
};

struct GemmArgs {
  // operands: hi parts mandatory; lo parts only for nparts == 3 (bf16x3 split precision:
  // hi*hi + hi*lo + lo*hi, ~fp32-class accuracy).  ld in elements (multiple of 8), base 16 B aligned.
  const __nv_bfloat16* A_hi = nullptr;
  const __nv_bfloat16* A_lo = nullptr;
  const __nv_bfloat16* B_hi = nullptr;
  const __nv_bfloat16* B_lo = nullptr;
  const __nv_bfloat16* A_lo2 = nullptr;  // third split (nparts == 6: hi + lo + lo2 carries 24 mantissa bits)
  const __nv_bfloat16* B_lo2 = nullptr;
  long long lda = 0, ldb = 0;
  long long strideA = 0, strideB = 0;  // per-batch element strides; 0 = operand shared by all batches
  int M = 0, N = 0, K = 0, batch = 1;
  int nparts = 1;  // 1 (bf16), 3 (bf16x3), 6 (bf16x6, fp32-class products)
  int flags = 0;
  int tri_koff = 0;  // with kTriB: B[n][k] == 0 for k > tri_koff + n (dense columns [0, tri_koff))
  // epilogue:  v = alpha * acc + beta * C + bias[n];  optional relu;  outputs: C (fp32) and/or
  // O_hi[/O_lo] (bf16 [split]) — any may be null.
  float alpha = 1.f, beta = 0.f;
  float* C = nullptr;
  long long ldc = 0, strideC = 0;
  const float* bias = nullptr;
  long long strideBias = 0;
  __nv_bfloat16* O_hi = nullptr;
  __nv_bfloat16* O_lo = nullptr;
  __nv_bfloat16* O_lo2 = nullptr;  // third split of the output (with O_lo): hi + lo + lo2 = 24 mantissa bits
  long long ldo = 0, strideO = 0;
  // Upper bound on the SMs the persistent launch occupies (0 = all).  A CTA of this kernel takes a whole SM
  // (shared memory); a background GEMM that runs beside a latency-bound kernel chain leaves SMs to the chain.
  int max_sms = 0;
};

// Returns 0 on success, negative on argument / CUDA error (see include/bk_kfac.h error codes).
int launch_umma_gemm(const GemmArgs& a, cudaStream_t stream);
// One problem of a grouped factor update: C (+)= alpha * X^T X with X^T staged K-major bf16 [d, n]
// (row pitch ldx), C fp32 [d, d] with a 16 B aligned base and pitch (ldc % 4 == 0), beta in {0, 1}.
struct SyrkGroupItem {
  const __nv_bfloat16* X_hi = nullptr;
  const __nv_bfloat16* X_lo = nullptr;  // bf16x3 only
  long long ldx = 0;
  int d = 0, n = 0;
  float alpha = 1.f, beta = 1.f;
  float* C = nullptr;
  long long ldc = 0;
};
// Up to 8 problems in one persistent CTA-pair launch (see bk_umma_gemm.cu).
// mirror = false: only the lower triangle of each C is accumulated (the caller mirrors once, when the full
// symmetric factor is read: bk_sym_finalize).
// mn_major = true: X_hi is the ROW-major activation matrix [n, d] itself (bf16, ldx = its row pitch, a multiple
// of 8): no staged K-major copy is needed; one precision pass (nparts == 1).
int launch_umma_syrk_grouped(const SyrkGroupItem* items, int count, int nparts, bool mirror,
                             cudaStream_t stream, bool mn_major = false);

// A/B switches of the grouped SYRK: bit 0 = no diagonal-tile operand dedup, bit 1 = stream-K tail on.
void set_syrk_tuning(int flags);
// Tuning / bring-up knob: force the tcgen05 cta_group of the contraction core (1 or 2; 0 = automatic).
void set_umma_cta_group(int cg);

}  // namespace bk
