// bk_kernels.cuh — internal launcher declarations (one per kernel family). The exported C ABI in
// bk_api.cu composes these; nothing here is visible outside the shared library.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace bk {

// ---- bk_prep.cu
// colsum (optional, fp32 [cols], pre-zeroed): receives sum_n scale * X[n, j] (bias row of A).
int launch_transpose_split(const float* X, long long ldx, int rows, int cols, float scale,
                           int ones_row, __nv_bfloat16* Thi, __nv_bfloat16* Tlo, long long ldt,
                           cudaStream_t stream, float* colsum = nullptr, int persistent_ctas = 0);
// Three-way bf16 splits (hi + lo + lo2: 24 mantissa bits) for the bf16x6 products of bk_chol.cu.
int launch_convert_split3(const float* X, long long ldx, int rows, int cols, __nv_bfloat16* O0,
                          __nv_bfloat16* O1, __nv_bfloat16* O2, long long ldo, cudaStream_t stream);
int launch_transpose_split3(const float* X, long long ldx, int rows, int cols, __nv_bfloat16* T0,
                            __nv_bfloat16* T1, __nv_bfloat16* T2, long long ldt,
                            cudaStream_t stream);
// Row / column d of a bias-augmented factor from the column sums (see bk_prep.cu).
int launch_colsum_bf16(const __nv_bfloat16* X, long long ldx, int rows, int cols, float scale, float* colsum,
                       cudaStream_t stream);
int launch_bias_border(float* state, long long ld, int d, const float* colsum, float alpha,
                       float beta, float n, cudaStream_t stream);
int launch_convert_split(const float* X, long long ldx, int rows, int cols, float scale,
                         int lower_only, __nv_bfloat16* Ohi, __nv_bfloat16* Olo, long long ldo,
                         cudaStream_t stream);
int launch_philox_normal(unsigned long long seed, unsigned sample0, unsigned stream_id, int rows,
                         int cols, int nsamples, float* Zf, long long ldf, long long stridef,
                         __nv_bfloat16* Zhi, __nv_bfloat16* Zlo, long long ldz, long long stridez,
                         cudaStream_t stream);

// ---- bk_factor_small.cu  (SIMT fp32 split-K SYRK for skinny factors; implicit im2col for conv)
int launch_small_syrk(float* state, long long ld_state, const float* x, long long ldx, int n, int d,
                      int has_bias, float in_scale, float alpha, float beta, cudaStream_t stream);
int launch_conv_a_syrk(float* state, long long ld_state, const float* x, int n, int c, int h, int w,
                       int kh, int kw, int pad_h, int pad_w, int stride_h, int stride_w,
                       int has_bias, float alpha, float beta, cudaStream_t stream);
int launch_conv_g_syrk(float* state, long long ld_state, const float* g, int n, int o, int hw,
                       float in_scale, float alpha, float beta, cudaStream_t stream);

// ---- bk_im2col.cu  (NCHW fp32 -> K-major bf16 patch-matrix operand of the tensor-core SYRK, wide conv factors)
int launch_im2col_split(const float* X, int n, int c, int h, int w, int kh, int kw, int ph, int pw, int sh,
                        int sw, float scale, int ones_row, __nv_bfloat16* Thi, __nv_bfloat16* Tlo,
                        long long ldt, cudaStream_t stream);

// ---- bk_syrk_fp32.cu  (full-fp32 SIMT SYRK, any d; parity mode for ill-conditioned factors)
int launch_syrk_fp32(float* state, long long ld_state, const float* x, long long ldx, int n, int d,
                     int has_bias, float in_scale, float alpha, float beta, cudaStream_t stream);

// ---- bk_diag.cu
int launch_diag_accum(float* state, const float* wgrad, const float* bgrad, int d_out, int d_in,
                      float scale, float beta, cudaStream_t stream);
int launch_diag_invert(float* inv, const float* state, long long count, float add, float multiply,
                       cudaStream_t stream);
int launch_diag_sample(float* out, const float* inv, long long count, int nsamples,
                       unsigned long long seed, unsigned sample0, unsigned stream_id,
                       const float* z_or_null, cudaStream_t stream);
int launch_diag_quadform(float* out, const float* J, long long ldj, const float* h, long long count,
                         int batch, cudaStream_t stream);

// ---- bk_chol.cu
size_t chol_inv_workspace_bytes(const int* dims, int count);
// Returns 0, the 1-based index of the first non-positive-definite factor, or a negative error.
// Synchronises `stream` (the status is computed on the device).
int chol_inv_batched(const float* const* factors, float* const* outs, const int* dims,
                     const float* add, const float* multiply, int count, void* workspace,
                     size_t workspace_bytes, cudaStream_t stream);


// 1 (default): the launch sequence of a given (workspace, dims) problem is captured once into a CUDA graph and
// replayed; 0: every call enqueues its kernels one by one.
void set_chol_graph(int enabled);
// SMs the background (FAR) outer updates of one phase may occupy together (0 = all)
void set_chol_far_sms(int sms);
// 1 (default): of each in-block trailing update only the next diagonal block stays on the chain
void set_chol_lookahead(int on);

// ---- bk_eigh.cu  (batched one-sided Jacobi eigensolver)
size_t eigh_workspace_bytes(const int* dims, int count);
// Returns 0, the 1-based index of the first factor that did not converge, or a negative error.
// Synchronises `stream` once per sweep when a factor is wider than the shared-memory path.
int eigh_batched(const float* const* factors, const long long* ldf, float* const* evals,
                 float* const* evecs, const int* dims, int count, float sym_scale, int max_sweeps,
                 void* workspace, size_t workspace_bytes, cudaStream_t stream);

// ---- bk_eigh_blocked.cu  (tensor-core block Jacobi for factors wider than the shared-memory path)
size_t eigh_blocked_workspace_bytes(int d);
// All wide factors of a batch concurrently (one pair of streams per factor, forked from / joined into `stream`).
// workspace: sum of eigh_blocked_workspace_bytes(d_i); fro2: `count` device floats of scratch;
// status[i] = 0 / 1 (not converged after max_sweeps).  Returns 0 or a negative error.
int eigh_blocked_batch(const float* const* F, const long long* ldf, const int* dims, int count, float sym_scale,
                       const float* tols, int max_sweeps, float* const* evals, float* const* evecs,
                       float* fro2, int* status, void* workspace, size_t workspace_bytes, cudaStream_t stream);
void set_eigh_mode(int mode);
void set_eigh_pair_width(int p);  // 0 = automatic, 64 or 128

// ---- bk_dense.cu
// H points at global row row0 of the P x P matrix; rows [row0, row0 + nrows) are summed.
int launch_dominance(const float* H, long long ld, int row0, int nrows, int P, float tau, const int* block_begin,
                     const int* block_end, int nblocks, double* out3, cudaStream_t stream);

int launch_ger_accum(float* state, long long ld, const float* g, int P, float alpha, float beta,
                     cudaStream_t stream);
int launch_kron(const float* a, int m, int n, const float* b, int p, int q, float* out,
                cudaStream_t stream);

// ---- bk_blockdiag.cu  (kernel-block-diagonal masks and per-component inverses of a dense Fisher)
constexpr int kBlockInvMaxDim = 160;      // one [d][d+1] fp64 matrix in one CTA's shared memory
int launch_band_mask(float* H, long long ld, int P, float tau, int in_place, const int* row_lo,
                     const int* row_hi, float* out, long long ldo, cudaStream_t stream);
int launch_block_inverse(const float* R, long long ld, int P, const int* comp_begin, const int* comp_end,
                         int ncomp, int max_dim, double scale, float* out, long long ldo, int zero_fill,
                         int* status, cudaStream_t stream);

// ---- bk_tri.cu  (lower-triangle packing of symmetric factors for the multi-GPU exchange)
int launch_tri_pack(const float* const* mats, const long long* lds, const int* dims, int count, float* packed,
                    cudaStream_t stream);
int launch_sym_finalize(float* const* mats, const long long* lds, const int* dims, int count, float scale,
                        cudaStream_t stream);
int launch_tri_unpack(float* const* mats, const long long* lds, const int* dims, int count, const float* packed,
                      float scale, int mirror, cudaStream_t stream);

// ---- bk_peer.cu  (factor exchange over peer memory: tile-packed triangles, fused pull + reduce + unpack, flags)
long long tile_packed_floats(const int* dims, int count);
int launch_tile_pack(const float* const* mats, const long long* lds, const int* dims, const long long* offs,
                     int count, float* packed, cudaStream_t stream);
int launch_tile_pack_to(const float* const* mats, const long long* lds, const int* dims, float* const* dsts,
                        int count, cudaStream_t stream);
int launch_peer_tile_unpack(float* const* mats, const long long* lds, const int* dims, int count,
                            const float* const* srcs, int nsrc, float scale, int mirror, cudaStream_t stream);
int launch_peer_copy(void* dst, const void* src, long long bytes, int vec_bytes, int ctas, cudaStream_t stream);
int launch_peer_signal(unsigned int* const* flags, int world, int me, unsigned int epoch, cudaStream_t stream);
int launch_peer_wait(const unsigned int* mine, int world, unsigned int epoch, double timeout_s, int* err,
                     cudaStream_t stream);

// ---- bk_inf.cu  (INF curvature: regularisation, fp64 pre-sampler chain, sampler tail)
int launch_inf_regularise(float* corr, long long nm, const float* lam, long long r, float add, float mult,
                          float* ric, float* rl, cudaStream_t stream);
int launch_inf_combine(float* out, const float* yl, const float* c, const float* xt, long long count,
                       cudaStream_t stream);
size_t inf_presample_workspace_bytes(int n, int a, int m, int b);
// Returns 0, 1 (vtv not positive definite), 2 (vtv + I not positive definite) or a negative error.
// Synchronises `stream`.
int inf_presample(const float* ua, long long lda, int n, int a, const float* ug, long long ldg, int m, int b,
                  const float* ric, const float* rl, float* p_out, void* workspace, size_t workspace_bytes,
                  cudaStream_t stream);

// ---- bk_metrics.cu  (calibration metrics)
int launch_calibration_rows(const float* probs, long long ld, const long long* labels, int n, int classes,
                            float* conf, float* correct, float* nll, float* ent, int* pred, double* totals,
                            cudaStream_t stream);
int launch_binned_stats(const float* x, const float* w1, const float* w2, long long n, const double* edges,
                        int nbins, int mode, double* out, cudaStream_t stream);

// ---- bk_forward.cu
int launch_sample_to_weights(const float* samples, const float* mean_w, const float* mean_b,
                             int d_out, int d_in, int has_bias, int nsamples, float* w_f32,
                             __nv_bfloat16* w_hi, __nv_bfloat16* w_lo, long long ldw, float* b_f32,
                             cudaStream_t stream);
// 1 (default): stride-1 3 x 3 / 5 x 5 layers take the register-tiled kernel; 0: always the generic one
void set_conv_fast(int on);
int launch_conv2d_relu_pool(const float* in, long long in_sample_stride, const float* w,
                            const float* b, float* out, int S, int N, int C, int H, int W, int O,
                            int KH, int KW, int SH, int SW, int PH, int PW, int relu, int pool,
                            cudaStream_t stream);
int launch_predictive_moments(const float* logits, int S, int B, int Cn, int mode, float* mean,
                              float* meansq, cudaStream_t stream);
int launch_frob_dot(float* out, const float* X, long long stride_x, const float* Y,
                    long long stride_y, long long count, int batch, int absolute, int accumulate,
                    cudaStream_t stream);

// ---- bk_small64.cu  (fp64 small-matrix path of the linearised predictive)
constexpr int kSmall64MaxDim = 112;       // two [d][d+1] fp64 buffers in one CTA's shared memory
constexpr int kSmall64MaxElems = 12544;   // d_in' * d_out of one layer (two fp64 copies in shared memory)
constexpr int kSmall64MaxBatch = 16;      // factors per launch
int launch_spd_inverse_f64(const float* const* factors, const long long* lds, const int* dims,
                           const double* add, const double* mult, double* const* outs, int count,
                           int* status, cudaStream_t stream);
// W = chol(sym_lower(F) + add I)^-1 (fp32 lower-triangular, d <= kSmall64MaxDim); status: atomicCAS(0 -> pivot).
int launch_chol_trinv_f64(const float* F, long long ldf, int d, double add, float* W, long long ldw,
                          int* status, cudaStream_t stream);
int launch_kron_quadform_f64(const float* V, long long stride_v, int batch, int dinp, int dout,
                             const double* Q, const double* H, float* out, int accumulate,
                             cudaStream_t stream);

}  // namespace bk
