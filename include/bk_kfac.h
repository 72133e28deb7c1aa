/* bk_kfac.h — C ABI of libbk_kfac.so, the sm_100a kernel library behind bnn_kfac_b200.
 *
 * Drop-in boundary: the reference (TianmingQiu/BNN_KFAC) has no FFI; its boundary is the Python class
 * API of models/curvatures.py and models/wrapper.py.  bnn_kfac_b200/curvatures.py mirrors that API and
 * is the only caller of the functions below.  Each entry point names the reference lines whose
 * arithmetic it replaces (paths relative to /root/reference).
 *
 * Conventions
 *   - plain pointers and sizes; every pointer is a DEVICE pointer unless it says "host".
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - row-major matrices; `ld*` are row pitches in ELEMENTS.
 *   - the caller owns every buffer, including workspaces (query the size first).
 *   - return value: 0 ok; > 0 1-based index of the first factor that was not positive definite;
 *     < 0 error (BK_ERR_*).  No global state besides cached function attributes.
 *   - precision: BK_PREC_BF16 (bf16 operands, fp32 accumulate) or BK_PREC_BF16X3 (hi/lo split,
 *     three tensor-core passes, ~fp32 accuracy).
 */
#ifndef BK_KFAC_H_
#define BK_KFAC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BK_OK 0
#define BK_ERR_ARG (-2)       /* bad argument / alignment */
#define BK_ERR_DRIVER (-3)    /* cuTensorMapEncodeTiled entry point unavailable */
#define BK_ERR_TMAP (-4)      /* tensor-map encode failed */
#define BK_ERR_CUDA (-5)      /* launch or runtime error */
#define BK_ERR_WORKSPACE (-6) /* workspace too small */
#define BK_ERR_ARCH (-7)      /* device is not compute capability 10.x */

#define BK_PREC_FP32 0 /* bk_syrk_accum only: full-fp32 SIMT SYRK (parity on ill-conditioned factors) */
#define BK_PREC_BF16 1
#define BK_PREC_BF16X3 3

/* flags for bk_gemm_nt (mirror bk::GemmFlags) */
#define BK_GEMM_SYRK_LOWER 1
#define BK_GEMM_MIRROR 2
#define BK_GEMM_TRI_A 4
#define BK_GEMM_TRI_B 8
#define BK_GEMM_RELU 16
#define BK_GEMM_TRI_B_UPPER 32 /* B[n][k] == 0 for k < n */
/* with BK_GEMM_TRI_B: B[n][k] == 0 only for k > koff + n (koff dense leading columns, multiple of 8) */
#define BK_GEMM_TRI_KOFF(koff) (((koff) / 8) << 8)

const char* bk_version(void);
/* Total number of kernels this library has launched in this process (monotonic counter). */
unsigned long long bk_launch_count(void);
/* 0 if the current device can run the kernels (compute capability 10.x), else BK_ERR_ARCH. */
int bk_device_check(void);

/* Tuning knob (process-wide): force the tcgen05 cta_group of the contraction core.  1 = one CTA per
 * SM (128 x 256 tiles), 2 = CTA pairs (256 x 256 tiles, tcgen05.mma.cta_group::2), 0 = automatic
 * (pairs whenever the problem has at least 192 rows and 129 columns).  Results do not depend on it. */
void bk_set_cta_group(int cta_group);
/* A/B switches of the grouped factor SYRK (profiling only): bit 0 = load both operands of diagonal tiles,
 * bit 1 = stream-K split of the last partial wave (default off: measured slower). */
void bk_set_syrk_tuning(int flags);

/* ---------------------------------------------------------------------------------------------
 * Contraction core.  D[b][m][n] = sum_k A[b][m][k] * B[b][n][k]; bf16 (uint16 storage) K-major
 * operands, optional lo parts for BK_PREC_BF16X3; v = alpha*acc + beta*C + bias[n], optional relu;
 * writes C (fp32) and/or O_hi/O_lo (bf16 split).  strideX == 0 shares the operand across batches.
 * Replaces torch.mm / `@` at models/curvatures.py:349,356,405 and the per-sample forward GEMMs
 * behind models/wrapper.py:35-44.
 */
int bk_gemm_nt(const void* a_hi, const void* a_lo, long long lda, long long stride_a,
               const void* b_hi, const void* b_lo, long long ldb, long long stride_b,
               int m, int n, int k, int batch, int precision, int flags,
               float alpha, float beta,
               float* c, long long ldc, long long stride_c,
               const float* bias, long long stride_bias,
               void* o_hi, void* o_lo, long long ldo, long long stride_o,
               void* stream);

/* fp32 [rows, cols] -> bf16 hi[/lo] [cols (+1 ones row), rows] (transposed, scaled).
 * The ones row is the reference's bias augmentation, models/curvatures.py:346-348. */
int bk_transpose_split(const float* x, long long ldx, int rows, int cols, float scale, int ones_row,
                       void* t_hi, void* t_lo, long long ldt, void* stream);
/* fp32 [rows, cols] -> bf16 hi[/lo] [rows, cols], optional lower-triangle mask. */
int bk_convert_split(const float* x, long long ldx, int rows, int cols, float scale, int lower_only,
                     void* o_hi, void* o_lo, long long ldo, void* stream);
/* Counter-based N(0,1) matrices [rows, cols]: element (r, c) of sample s of stream `stream_id` =
 * Philox4x32-10(key = seed, counter = (c/4, r, sample0 + s, stream_id)) lane c%4, Box-Muller.
 * Replaces torch.randn at models/curvatures.py:404.  Any of zf / z_hi may be null; the padding
 * columns [cols, ldz) of the bf16 outputs are zero-filled when ldz is a multiple of 4. */
int bk_philox_normal(unsigned long long seed, unsigned sample0, unsigned stream_id, int rows,
                     int cols, int nsamples, float* zf, long long ldf, long long stride_f,
                     void* z_hi, void* z_lo, long long ldz, long long stride_z, void* stream);

/* ---------------------------------------------------------------------------------------------
 * KFAC factor accumulation (models/curvatures.py:345-349, 355-356, 359-363):
 *   state = beta*state + alpha * [in_scale*x ; 1]^T [in_scale*x ; 1]      (x is [n, d] row-major)
 * state is [d+has_bias, d+has_bias] fp32, written as a full symmetric matrix.  The reference's
 * A = mm(fwd, fwd^T)/N is alpha = 1/N, in_scale = 1; its G with g = grad_output*N is in_scale = N.
 * d + has_bias > BK_SMALL_D_MAX uses the tcgen05 SYRK and needs a workspace; smaller factors use the
 * SIMT split-K kernel and need none.
 */
#define BK_SMALL_D_MAX 176
size_t bk_syrk_workspace_bytes(int n, int d, int has_bias, int precision);
int bk_syrk_accum(float* state, long long ld_state, const float* x, long long ldx, int n, int d,
                  int has_bias, float in_scale, float alpha, float beta, int precision,
                  void* workspace, size_t workspace_bytes, void* stream);
/* All factors of one KFAC.update in one call (host arrays of `count` entries, same meaning as the
 * bk_syrk_accum arguments).  Wide factors with a 16 B aligned state (base and pitch) and input are
 * staged and then accumulated by ONE persistent tensor-core launch over all of them, so wave
 * quantisation and the pipeline fill / drain are paid once per update, not once per factor; the
 * others take the bk_syrk_accum route one by one. */
size_t bk_syrk_grouped_workspace_bytes(const int* ns, const int* ds, const int* has_bias, int count,
                                       int precision);
/* flags:
 *   BK_SYRK_LOWER_ONLY  the tensor-core items accumulate the LOWER triangle only (diagonal included); the upper
 *                       triangle of those states is unspecified until bk_sym_finalize mirrors it.  Saves the
 *                       mirrored half of the epilogue's L2 reduction traffic on every update.
 *   BK_SYRK_NO_OVERLAP  stage every operand first, then run the SYRKs (default: the staging of later factors
 *                       runs on an internal per-device side stream underneath the SYRK of earlier ones; the
 *                       call is still ordered on `stream` as a whole).
 *   BK_SYRK_STAGE_PERSISTENT  A/B switch, off: the overlapped staging passes run as one persistent CTA per SM
 *                       (measured slower than the tile grid: 0.58 vs 0.49 ms per cfg5 step).
 *   BK_SYRK_ROW_MAJOR   (bk_syrk_accum_staged_grouped only) the operands are row-major bf16 activations [n, d]
 *                       (ldt = their row pitch, a multiple of 8), not staged K-major copies.
 * x_is_bf16 (nullable = all fp32): xs[i] is a bf16 matrix [n, d] (row pitch ldxs[i] % 8 == 0, 16 B aligned base).
 * Such activations (a model running under bf16 autocast) feed the tensor cores directly - no staging pass, the
 * products of bf16 values are exact in the fp32 accumulator, so one pass is the full-precision result; only wide
 * factors (d + has_bias > BK_SMALL_D_MAX, d >= 192) qualify, anything else returns BK_ERR_ARG (convert it to
 * fp32 first). */
#define BK_SYRK_LOWER_ONLY 1
#define BK_SYRK_NO_OVERLAP 2
#define BK_SYRK_STAGE_PERSISTENT 8 /* A/B switch (measured slower, off): overlapped staging passes run as one persistent CTA per SM */
#define BK_SYRK_ROW_MAJOR 4
int bk_syrk_accum_grouped(float* const* states, const long long* ld_states, const void* const* xs,
                          const int* x_is_bf16, const long long* ldxs, const int* ns, const int* ds,
                          const int* has_bias, const float* in_scales, const float* alphas,
                          const float* betas, int count, int precision, int flags, void* workspace,
                          size_t workspace_bytes, void* stream);
/* In place: lower triangle (diagonal included) *= scale, upper triangle = its mirror: turns lower-only
 * (and, for the running-average mode, lazily scaled) accumulators into the full symmetric factors the
 * reference keeps in `state` (models/curvatures.py:359-363).  Host arrays of device pointers, as bk_tri_pack. */
int bk_sym_finalize(float* const* factors_host, const long long* ld_host, const int* dims_host, int count,
                    float scale, void* stream);
/* The grouped tensor-core launch alone, on operands already staged as K-major bf16 [d, n] (no bias
 * row): up to 8 problems, states 16 B aligned with ld % 4 == 0, beta in {0, 1}; flags as above. */
int bk_syrk_accum_staged_grouped(float* const* states, const long long* ld_states,
                                 const void* const* xt_his, const void* const* xt_los,
                                 const long long* ldts, const int* ns, const int* ds,
                                 const float* alphas, const float* betas, int count, int precision,
                                 int flags, void* stream);
/* Same, operand already staged as K-major bf16 [d+has_bias, n] (e.g. by bk_transpose_split). */
int bk_syrk_accum_staged(float* state, long long ld_state, const void* xt_hi, const void* xt_lo,
                         long long ldt, int n, int dprime, float alpha, float beta, int precision,
                         void* stream);

/* Conv2d first factor with implicit im2col (models/curvatures.py:341-349): x is NCHW fp32,
 * state [c*kh*kw + has_bias]^2;  alpha is applied to the un-normalised sum (reference: 1/(N*L)). */
int bk_conv_a_accum(float* state, long long ld_state, const float* x, int n, int c, int h, int w,
                    int kh, int kw, int pad_h, int pad_w, int stride_h, int stride_w, int has_bias,
                    float alpha, float beta, void* stream);
/* Wide Conv2d factors (c*kh*kw + has_bias > BK_SMALL_D_MAX): the K-major bf16 operand of the tensor-core SYRK
 * straight from the NCHW activations, t[r][col] = scale * patch value with r = (ci*kh + i)*kw + j (unfold's row
 * order, models/curvatures.py:342-343), col = img*L + oh*OW + ow, plus a row of ones when ones_row != 0
 * (:346-348); columns [n*L, ldt) are zero.  ldt % 8 == 0, 16 B aligned outputs, t_lo nullable.  Feed the result to
 * bk_syrk_accum_staged(dprime = c*kh*kw + ones_row, n = n*L).  kh = kw = 1 regroups output gradients
 * [n, o, h'w'] into [o, n*h'w'] (:353). */
int bk_im2col_split(const float* x, int n, int c, int h, int w, int kh, int kw, int pad_h, int pad_w,
                    int stride_h, int stride_w, float scale, int ones_row, void* t_hi, void* t_lo, long long ldt,
                    void* stream);
/* Conv2d second factor (models/curvatures.py:353-356): g is [n, o, hw] fp32. */
int bk_conv_g_accum(float* state, long long ld_state, const float* g, int n, int o, int hw,
                    float in_scale, float alpha, float beta, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Diagonal curvature (models/curvatures.py:155-207).
 */
/* state[o, i] = beta*state + scale * [wgrad | bgrad]^2  (Diagonal.update :165-172) */
int bk_diag_accum(float* state, const float* wgrad, const float* bgrad, int d_out, int d_in,
                  float scale, float beta, void* stream);
/* inv = 1/sqrt(multiply*state + add)  (Diagonal.invert :202) */
int bk_diag_invert(float* inv, const float* state, long long count, float add, float multiply,
                   void* stream);
/* out[s] = z[s] * inv  (Diagonal.sample :207); z from Philox unless z_or_null is given. */
int bk_diag_sample(float* out, const float* inv, long long count, int nsamples,
                   unsigned long long seed, unsigned sample0, unsigned stream_id,
                   const float* z_or_null, void* stream);
/* out[b] = sum_j J[b, j]^2 * h[j]  (classification_ll_diagonal.py:131, regression_ll_diagonal.py:139) */
int bk_diag_quadform(float* out, const float* j, long long ldj, const float* h, long long count,
                     int batch, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Damped inversion (KFAC.invert, models/curvatures.py:381-392), batched over factors:
 *   R = sqrt(multiply)*F + sqrt(add)*I;  R <- (R + R^T)/2;  out = cholesky_lower(inverse(R))
 * computed as a reverse Cholesky + triangular inverse (no explicit inverse).  Host arrays of device
 * pointers / dims; `info` is a device int (first failing 1-based factor index, 0 if none) that the
 * call also returns after synchronising the stream.
 */
size_t bk_chol_inv_workspace_bytes(const int* dims_host, int count);
/* Process-wide switch (default 1): the step sequence of bk_damp_chol_inv_batched - ~450 dependent launches on two
 * streams for one 4097-wide factor - depends only on (workspace address, dims), so it is captured ONCE into a CUDA
 * graph and replayed by later calls (factor / output pointers and the damping scalars travel through a device
 * table that is refreshed outside the graph).  0 = enqueue every kernel on every call. */
void bk_set_chol_graph(int enabled);
/* Tuning knob (default 64): SMs that the background ("far") tensor-core updates of one inversion phase may occupy
 * together while the latency-bound diagonal / panel chain of the next outer block runs beside them; 0 = no limit. */
void bk_set_chol_far_sms(int sms);
/* Switch (default 1): look-ahead in the inner loop of the inversion - of the trailing update of step k only the next
 * 64 x 64 diagonal block is computed on the chain, the rest runs on a third stream beside the next diagonal-block
 * kernel and is awaited by the panel of step k + 1.  Same bits either way. */
void bk_set_chol_lookahead(int enabled);
int bk_damp_chol_inv_batched(const float* const* factors_host, float* const* outs_host,
                             const int* dims_host, const float* add_host,
                             const float* multiply_host, int count, void* workspace,
                             size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Symmetric eigendecomposition, batched over factors (models/utilities.py:144-159 get_eigenvectors:
 * symeig of F + F^T; :120-141 get_eigenvalues: symeig of F): S = sym_scale * (F + F^T).  d <= 164: one-sided
 * Jacobi with warp-shuffle rotations, the whole problem in the shared memory of one CTA, all such factors
 * in one launch.  Wider factors: two-sided block Jacobi (pair solves in shared memory, rotations applied as
 * batched tensor-core GEMMs), all wide factors of the batch concurrently on internal streams that are
 * forked from / joined into `stream`; accuracy: eigenvalues to ~1e-5 |lambda|_max, V diag(w) V^T = S to
 * 1e-4 .. 1e-3 |S|_F (d = 300 .. 4097).  evals[i]: [d] ascending; evecs[i]: [d, d] row-major with the
 * eigenvectors as COLUMNS (symeig / linalg.eigh layout), may be null (host array or entries).
 * Returns 0, > 0 = 1-based index of the first factor not converged after max_sweeps (<= 0: 30),
 * < 0 error.  Host arrays of device pointers / leading dimensions / dims (count <= 64).  Blocks the host
 * (one stream synchronisation per sweep).
 */
/* Tuning knob (process-wide) for factors wider than the shared-memory path (d > 164): 0 (default) = two-sided
 * block Jacobi whose rotations are applied as batched tensor-core GEMMs; 1 = element-wise one-sided
 * Jacobi streamed from L2 (one launch per round).  Set it before querying the workspace size. */
void bk_set_eigh_mode(int mode);
/* Width of a block pair of the block Jacobi (64 or 128 columns; 0 = automatic by size). */
void bk_set_eigh_pair_width(int width);
size_t bk_eigh_workspace_bytes(const int* dims_host, int count);
int bk_eigh_batched(const float* const* factors_host, const long long* ld_host,
                    float* const* evals_host, float* const* evecs_host, const int* dims_host,
                    int count, float sym_scale, int max_sweeps, void* workspace,
                    size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Dense Fisher (hessian/classification_ll_dense_kernel_diag.py:68-91, hessian/utils.py:4-23).
 * out3 (device, fp64) = { sum |diag(H + tau I)|, sum |H + tau I|, sum over the diagonal blocks
 * [block_begin[b], block_end[b]) of |H + tau I| }; blocks disjoint and ascending (device int arrays).
 * H itself is accumulated with bk_syrk_accum on the stacked flat gradients; its damped inverse is
 * bk_damp_chol_inv_batched(add = tau^2, multiply = 1).
 */
int bk_dominance(const float* h, long long ld, int p, float tau, const int* block_begin,
                 const int* block_end, int nblocks, double* out3, void* stream);
/* The same sums over the global rows [row0, row0 + nrows) only; `rows` points at row row0 (a row-block shard of
 * the dense Fisher, bnn_kfac_b200/dense_sharded.py: the ranks' partial sums are added by one all-reduce). */
int bk_dominance_rows(const float* rows, long long ld, int row0, int nrows, int p, float tau,
                      const int* block_begin, const int* block_end, int nblocks, double* out3, void* stream);
/* Rank-1 accumulation state = beta*state + alpha * g g^T, state [p, p] fp32 (row pitch ld), g [p]:
 * BlockDiagonal.update (models/curvatures.py:228-232, `torch.ger(grads, grads) * batch_size`, `+=`). */
int bk_ger_accum(float* state, long long ld, const float* g, int p, float alpha, float beta,
                 void* stream);
/* Kronecker product out[m*p, n*q] = a[m, n] (x) b[p, q], all contiguous fp32
 * (models/utilities.py:387-409, sampling_free/utils.py:279-290). */
int bk_kron(const float* a, int m, int n, const float* b, int p, int q, float* out, void* stream);

/* Kernel-block-diagonal approximation of a dense Fisher (sampling_free/utils.py:63-211,
 * generate_kernel_diag_15080 / _748 / _141 / generate_kernel_diag):
 *     H += tau I;  res = 0;  res[a:b, a:b] = H[a:b, a:b] for every block;  return res, inverse(n * res)
 * bk_band_mask: out[i][j] = H[i][j] (+ tau on the diagonal) for row_lo[i] <= j < row_hi[i], else 0; row_lo /
 * row_hi (device int [p]) are the extreme bounds of the blocks that contain row i (0, 0 for a row in no block) -
 * for interval blocks, overlapping ones included, that is exactly the union of the squares.
 * add_tau_in_place != 0 also performs the reference's in-place H += tau I.  h and out must not alias.
 * bk_block_inverse: out[a:b, a:b] = inverse(scale * res[a:b, a:b]) for each CONNECTED component [a, b) of the
 * block union (device int arrays, b - a <= BK_BLOCK_INV_MAX_DIM = max_dim bound): fp64 Gauss-Jordan with partial
 * pivoting, one CTA per component; zero_fill != 0 clears the rest of out first.  status (device int, nullable):
 * 0 or 65536 * component + (1-based step) of the first exactly singular pivot. */
#define BK_BLOCK_INV_MAX_DIM 160
int bk_band_mask(float* h, long long ld, int p, float tau, int add_tau_in_place, const int* row_lo,
                 const int* row_hi, float* out, long long ldo, void* stream);
int bk_block_inverse(const float* res, long long ld, int p, const int* comp_begin, const int* comp_end,
                     int ncomp, int max_dim, double scale, float* out, long long ldo, int zero_fill, int* status,
                     void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU factor exchange (no counterpart in the single-device reference).  The accumulated
 * factors are symmetric, so ranks exchange packed lower triangles: factor after factor in one flat fp32
 * buffer, row i of a factor = its i + 1 values at offset i (i + 1) / 2 (sum over factors of d (d + 1) / 2
 * values in total).  Host arrays of device pointers / row pitches / dims.
 */
int bk_tri_pack(const float* const* factors_host, const long long* ld_host, const int* dims_host, int count,
                float* packed, void* stream);
/* Expands the packed buffer into full [d, ld] matrices, every value multiplied by `scale`: mirror != 0 writes
 * the symmetric matrix (accumulated factors), mirror == 0 a zero upper triangle (the lower-triangular Cholesky
 * factors the owning ranks send back, models/curvatures.py:391-392). */
int bk_tri_unpack(float* const* outs_host, const long long* ld_host, const int* dims_host, int count,
                  const float* packed, float scale, int mirror, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Factor exchange over NVLink / NVSwitch PEER MEMORY (one process per GPU; no counterpart in the reference).
 * Buffers are cudaMalloc allocations exported through CUDA IPC; the 64-byte handles travel through the caller's
 * process group.  "Tile-packed" layout of a lower triangle: its 32 x 32 tiles, tile (ti, tj <= ti) at index
 * ti (ti + 1) / 2 + tj, 1024 floats each, row-major (factor after factor; bk_tile_packed_floats values in total).
 */
int bk_peer_alloc(size_t bytes, void** ptr);                    /* zero-filled device allocation */
int bk_peer_free(void* ptr);
int bk_peer_export(void* ptr, void* handle64);                  /* cudaIpcGetMemHandle: 64 bytes out */
int bk_peer_open(const void* handle64, void** ptr);             /* maps another process's allocation */
int bk_peer_close(void* ptr);
int bk_peer_read_u32(const void* ptr, unsigned int* out_host);   /* synchronous 4-byte read (error word) */
long long bk_tile_packed_floats(const int* dims_host, int count);
/* dense [d, ld] lower triangles -> tile-packed buffer (count <= 16); offsets_host[k] = first float of factor k
 * inside `packed` (a multiple of 1024), NULL = back to back */
int bk_tile_pack(const float* const* factors_host, const long long* ld_host, const int* dims_host,
                 const long long* offsets_host, int count, float* packed, void* stream);
/* The same with one destination per factor (dsts_host[k], 4 KB aligned tile-packed area of factor k): a destination
 * in another rank's exported buffer makes the pack the SEND (posted writes over NVLink). */
int bk_tile_pack_to(const float* const* factors_host, const long long* ld_host, const int* dims_host,
                    float* const* dsts_host, int count, void* stream);
/* ONE kernel = collective + unpack: out[k] = scale * sum over r < nsrc (<= 8) of the tile-packed triangle at
 * srcs_host[k * nsrc + r] (local or peer memory), summed in index order; mirror != 0 writes the symmetric matrix
 * (reduce-scatter of accumulated factors, nsrc = world), mirror == 0 a zero upper triangle (all-gather of the
 * Cholesky factors, nsrc = 1). */
int bk_peer_tile_unpack(float* const* outs_host, const long long* ld_host, const int* dims_host, int count,
                        const float* const* srcs_host, int nsrc, float scale, int mirror, void* stream);
/* Measurement aid: copy kernel with 4- or 16-byte accesses (bytes % 16 == 0) on `ctas` CTAs of 256 threads; a pull
 * or a push depending on which pointer is peer memory. */
int bk_peer_copy(void* dst, const void* src, long long bytes, int vec_bytes, int ctas, void* stream);
/* Cross-GPU ordering.  flags_host[r]: a flag array (world 32-bit slots) inside rank r's exported buffer.
 * bk_peer_signal: system-scope fence, then slot `me` of every rank's array = epoch.  bk_peer_wait: returns (in
 * stream order) once all `world` slots of the local array have reached epoch; after timeout_s seconds it stores
 * 1 + the missing rank into *err (device int) instead of spinning on. */
int bk_peer_signal(unsigned int* const* flags_host, int world, int me, unsigned int epoch, void* stream);
int bk_peer_wait(const unsigned int* flags_local, int world, unsigned int epoch, double timeout_s, int* err,
                 void* stream);

/* ---------------------------------------------------------------------------------------------
 * INF curvature: low-rank eigenbasis + diagonal correction (models/curvatures.py:476-682).
 */
/* INF.invert :537-539.  correction[correction < 0] = 0 IN PLACE (nm values);
 * reg_inv_correction = 1/sqrt(multiply*correction + add) (nm); reg_lambda = sqrt(multiply*lambda) (r). */
int bk_inf_regularise(float* correction, long long nm, const float* lambda, long long r, float add,
                      float multiply, float* reg_inv_correction, float* reg_lambda, void* stream);
/* INF.pre_sampler :565-585.  ua [n, a] / ug [m, b]: low-rank eigenvector columns (row pitches lda / ldg);
 * reg_inv_correction [n*m] (index i*m + p); reg_lambda [a*b] (index q*b + x); p_out [a*b, a*b] fp32 =
 * diag(s) (C^-1 + V^T V)^-1 diag(s) with V = c (.) kron(ua, ug) diag(s), never materialised; the r x r
 * Cholesky / triangular-inverse chain runs in fp64 (see bk_inf.cu).  Returns 0, 1 if V^T V is not
 * positive definite, 2 if V^T V + I is not, < 0 on error.  Synchronises the stream. */
size_t bk_inf_presample_workspace_bytes(int n, int a, int m, int b);
int bk_inf_presample(const float* ua, long long lda, int n, int a, const float* ug, long long ldg, int m,
                     int b, const float* reg_inv_correction, const float* reg_lambda, float* p_out,
                     void* workspace, size_t workspace_bytes, void* stream);
/* INF.sampler :611.  out = y_l - reg_inv_correction^2 * x_ps_t (all [count] fp32). */
int bk_inf_combine(float* out, const float* y_l, const float* reg_inv_correction, const float* x_ps_t,
                   long long count, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Calibration metrics of class probabilities (models/utilities.py:178-366).
 */
/* One pass over probs [n, classes] (row pitch ld), labels int64 [n] (nullable).  Per-row outputs (each
 * nullable): conf = max p, correct = (argmax == label) as 0/1, nll = -log(p[label] + 1e-12), entropy of
 * the row normalised to sum 1, pred = argmax (first maximum).  totals4 (device, fp64) receives
 * { sum correct, sum conf, sum nll, sum entropy }. */
int bk_calibration_rows(const float* probs, long long ld, const long long* labels, int n, int classes,
                        float* conf, float* correct, float* nll, float* entropy, int* pred, double* totals4,
                        void* stream);
/* out3 (device, fp64 [3, nbins]) = per bin { count, sum w1, sum w2 } of the values x[i] (w1 / w2 nullable)
 * for fp64 bin edges [nbins + 1] (device; nbins <= 256).  mode 0: (lo, hi] (expected_calibration_error
 * :322); mode 1: (lo, hi) (calibration_curve :287; edges may repeat); mode 2: [lo, hi) with the last bin
 * closed (np.histogram, binned_kl_distance :207-208). */
int bk_binned_stats(const float* x, const float* w1, const float* w2, long long n, const double* edges,
                    int nbins, int mode, double* out3, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Predictive glue (the dense contractions of these stages are bk_gemm_nt calls).
 */
/* Curvature._replace for a batch of samples (models/curvatures.py:67-82):
 *   W_s = mean_w + samples[s][:, :d_in],  b_s = mean_b + samples[s][:, d_in]
 * samples is [nsamples, d_out, d_in + has_bias] fp32.  Outputs (any weight output may be null):
 * w_f32 [nsamples, d_out, d_in], w_hi/w_lo bf16 split [nsamples, d_out, ldw], b_f32 [nsamples, d_out]. */
int bk_sample_to_weights(const float* samples, const float* mean_w, const float* mean_b, int d_out,
                         int d_in, int has_bias, int nsamples, float* w_f32, void* w_hi, void* w_lo,
                         long long ldw, float* b_f32, void* stream);
/* Per-sample conv2d (+bias, optional ReLU, optional 2x2 max-pool) for the reference CNNs
 * (models/wrapper.py:53-101): in [nsamples or 1, n, c, h, w] (in_sample_stride 0 = shared input),
 * w [nsamples, o, c, kh, kw], b [nsamples, o] or null, out [nsamples, n, o, oh', ow']. */
/* Switch (default 1) of bk_conv2d_relu_pool: stride-1 layers with 3 x 3 / 5 x 5 kernels run a register-tiled kernel
 * (weights staged once per group of 8 images, one (K + 1)^2 input patch per pooled output); 0 = always the generic
 * one-CTA-per-(sample, image) kernel.  Same values up to the order of the fp32 accumulation. */
void bk_set_conv_fast(int enabled);
int bk_conv2d_relu_pool(const float* in, long long in_sample_stride, const float* w, const float* b,
                        float* out, int nsamples, int n, int c, int h, int wd, int o, int kh, int kw,
                        int sh, int sw, int ph, int pw, int relu, int pool, void* stream);
/* Moments over posterior samples of logits [nsamples, batch, classes] (classes <= 1024):
 * mode 0: p = softmax (sampling/classification_sampling.py:74-79); mode 1: p = raw output
 * (sampling/regression_sampling.py:86-88).  mean = E_s[p]; meansq = E_s[p^2] (nullable).
 * mode 2: raw output with the CENTRED second moment, meansq = E_s[(p - mean)^2] (numpy std^2, ddof 0),
 * computed in two passes so that small variances survive |mean| >> std. */
int bk_predictive_moments(const float* logits, int nsamples, int batch, int classes, int mode,
                          float* mean, float* meansq, void* stream);
/* out[b] (+)= <x_b, y_b> (optionally absolute value): the last step of the kron-free
 * J (Q (x) H) J^T = <V, Q V H^T> (sampling_free/classification/classification_ll_block.py:131-132). */
int bk_frob_dot(float* out, const float* x, long long stride_x, const float* y, long long stride_y,
                long long count, int batch, int absolute, int accumulate, void* stream);

/* fp64 small-matrix path of the sampling-free regression predictive
 * (sampling_free/regression/regression_ll_block.py:126-138: q_inv = pinverse(N (q_i + tau I)), ...,
 * |J_i kron(q_inv, h_inv) J_i^T|).  The damped factors of that problem reach cond 1e5..5e6, where fp32
 * arithmetic resolves the quadratic form to ~1e-2 only; factors that fit one CTA are therefore inverted and
 * contracted in fp64.
 * bk_spd_inverse_f64: outs[i] (device fp64 [d, d], dense) = (multiply[i] * sym(F_i) + add[i] * I)^-1 for up to
 * BK_SMALL64_MAX_BATCH fp32 factors with d <= BK_SMALL64_MAX_DIM (host arrays of device pointers, as in
 * bk_damp_chol_inv_batched).  `status` (device int, nullable) receives 0 or 65536 * i + (1-based pivot) of the
 * first factor that is not positive definite.
 * bk_kron_quadform_f64: out[b] (+)= |<V_b, Q V_b H^T>|, V_b fp32 [d_in', d_out] at v + b * stride_v,
 * Q fp64 [d_in', d_in'], H fp64 [d_out, d_out], d_in' * d_out <= BK_SMALL64_MAX_ELEMS. */
#define BK_SMALL64_MAX_DIM 112
#define BK_SMALL64_MAX_ELEMS 12544
#define BK_SMALL64_MAX_BATCH 16
int bk_spd_inverse_f64(const float* const* factors_host, const long long* ld_host, const int* dims_host,
                       const double* add_host, const double* multiply_host, double* const* outs_host,
                       int count, int* status, void* stream);
int bk_kron_quadform_f64(const float* v, long long stride_v, int batch, int d_in_p, int d_out,
                         const double* q, const double* h, float* out, int accumulate, void* stream);

/* One diagonal block of the blocked Cholesky of the SHARDED dense Fisher (no counterpart in the
 * single-device reference, whose `pinverse(H + tau I)` is sampling_free/classification/classification_ll_dense.py:
 * 108-109): w (fp32 [d, ldw], lower triangular, zero upper part) = chol(sym_lower(f) + add I)^-1, computed in fp64
 * by one CTA, d <= BK_SMALL64_MAX_DIM; only the lower triangle of f is read.  status (device int, nullable, NOT
 * cleared by the call) receives the 1-based pivot index if the block is not positive definite. */
int bk_chol_trinv_f64(const float* f, long long ldf, int d, double add, float* w, long long ldw, int* status,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BK_KFAC_H_ */
