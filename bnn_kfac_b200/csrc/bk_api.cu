// bk_api.cu — the exported C ABI (include/bk_kfac.h).  Argument validation + composition of the
// kernel launchers; no torch types, no allocation, no hidden synchronisation except where the
// header says so (bk_damp_chol_inv_batched returns a device-computed status).
#include "../../include/bk_kfac.h"

#include "bk_common.cuh"
#include "bk_kernels.cuh"
#include "bk_umma_gemm.cuh"

namespace {

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline long long round8(long long v) { return (v + 7) / 8 * 8; }

}  // namespace

#include <atomic>
#include <mutex>

namespace bk {
static std::atomic<unsigned long long> g_launches{0};
void note_launch(int n) { g_launches.fetch_add(static_cast<unsigned long long>(n)); }
unsigned long long launch_count() { return g_launches.load(); }
}  // namespace bk

#pragma GCC visibility push(default)
extern "C" {

const char* bk_version(void) { return "bk_kfac 0.1 (sm_100a)"; }

unsigned long long bk_launch_count(void) { return bk::g_launches.load(); }

int bk_device_check(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return BK_ERR_CUDA;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
    return BK_ERR_CUDA;
  return major == 10 ? BK_OK : BK_ERR_ARCH;
}

void bk_set_cta_group(int cta_group) { bk::set_umma_cta_group(cta_group); }

void bk_set_syrk_tuning(int flags) { bk::set_syrk_tuning(flags); }

void bk_set_chol_graph(int enabled) { bk::set_chol_graph(enabled); }

void bk_set_conv_fast(int enabled) { bk::set_conv_fast(enabled); }

void bk_set_chol_far_sms(int sms) { bk::set_chol_far_sms(sms); }

void bk_set_chol_lookahead(int enabled) { bk::set_chol_lookahead(enabled); }

void bk_set_eigh_mode(int mode) { bk::set_eigh_mode(mode); }

void bk_set_eigh_pair_width(int width) { bk::set_eigh_pair_width(width); }

int bk_gemm_nt(const void* a_hi, const void* a_lo, long long lda, long long stride_a,
               const void* b_hi, const void* b_lo, long long ldb, long long stride_b, int m, int n,
               int k, int batch, int precision, int flags, float alpha, float beta, float* c,
               long long ldc, long long stride_c, const float* bias, long long stride_bias,
               void* o_hi, void* o_lo, long long ldo, long long stride_o, void* stream) {
  if (precision != BK_PREC_BF16 && precision != BK_PREC_BF16X3) return BK_ERR_ARG;
  bk::GemmArgs g;
  g.A_hi = static_cast<const __nv_bfloat16*>(a_hi);
  g.A_lo = static_cast<const __nv_bfloat16*>(a_lo);
  g.B_hi = static_cast<const __nv_bfloat16*>(b_hi);
  g.B_lo = static_cast<const __nv_bfloat16*>(b_lo);
  g.lda = lda;
  g.ldb = ldb;
  g.strideA = stride_a;
  g.strideB = stride_b;
  g.M = m;
  g.N = n;
  g.K = k;
  g.batch = batch;
  g.nparts = precision;
  g.flags = flags & 0xFF;
  g.tri_koff = ((flags >> 8) & 0xFFFFFF) * 8;
  g.alpha = alpha;
  g.beta = beta;
  g.C = c;
  g.ldc = ldc;
  g.strideC = stride_c;
  g.bias = bias;
  g.strideBias = stride_bias;
  g.O_hi = static_cast<__nv_bfloat16*>(o_hi);
  g.O_lo = static_cast<__nv_bfloat16*>(o_lo);
  g.ldo = ldo;
  g.strideO = stride_o;
  return bk::launch_umma_gemm(g, as_stream(stream));
}

int bk_transpose_split(const float* x, long long ldx, int rows, int cols, float scale, int ones_row,
                       void* t_hi, void* t_lo, long long ldt, void* stream) {
  if (x == nullptr || t_hi == nullptr || ldt < rows || ldx < cols) return BK_ERR_ARG;
  return bk::launch_transpose_split(x, ldx, rows, cols, scale, ones_row,
                                    static_cast<__nv_bfloat16*>(t_hi),
                                    static_cast<__nv_bfloat16*>(t_lo), ldt, as_stream(stream));
}

int bk_convert_split(const float* x, long long ldx, int rows, int cols, float scale, int lower_only,
                     void* o_hi, void* o_lo, long long ldo, void* stream) {
  if (x == nullptr || o_hi == nullptr || ldo < cols || ldx < cols) return BK_ERR_ARG;
  return bk::launch_convert_split(x, ldx, rows, cols, scale, lower_only,
                                  static_cast<__nv_bfloat16*>(o_hi),
                                  static_cast<__nv_bfloat16*>(o_lo), ldo, as_stream(stream));
}

int bk_philox_normal(unsigned long long seed, unsigned sample0, unsigned stream_id, int rows,
                     int cols, int nsamples, float* zf, long long ldf, long long stride_f,
                     void* z_hi, void* z_lo, long long ldz, long long stride_z, void* stream) {
  if (zf == nullptr && z_hi == nullptr) return BK_ERR_ARG;
  return bk::launch_philox_normal(seed, sample0, stream_id, rows, cols, nsamples, zf, ldf, stride_f,
                                  static_cast<__nv_bfloat16*>(z_hi),
                                  static_cast<__nv_bfloat16*>(z_lo), ldz, stride_z,
                                  as_stream(stream));
}

// ------------------------------------------------------------------------------ factor update
// Workspace layout of the tensor-core path: [X^T hi : d x ldt bf16][X^T lo (bf16x3 only)][colsum : d fp32]
// The bias row of ones is NOT staged: the SYRK runs on the d x d block and row / column d of the
// factor come from exact fp32 column sums (one extra 256-row tile row of the SYRK saved at d = 4096).
size_t bk_syrk_workspace_bytes(int n, int d, int has_bias, int precision) {
  const int dp = d + (has_bias ? 1 : 0);
  if (dp <= BK_SMALL_D_MAX || precision == BK_PREC_FP32) return 0;
  const size_t one = align_up(static_cast<size_t>(dp) * round8(n) * 2, 256);
  const size_t sums = has_bias ? align_up(static_cast<size_t>(d) * 4, 256) : 0;
  return (precision == BK_PREC_BF16X3 ? 2 * one : one) + sums;
}

int bk_syrk_accum_staged(float* state, long long ld_state, const void* xt_hi, const void* xt_lo,
                         long long ldt, int n, int dprime, float alpha, float beta, int precision,
                         void* stream) {
  if (state == nullptr || xt_hi == nullptr || ld_state < dprime) return BK_ERR_ARG;
  if (precision != BK_PREC_BF16 && precision != BK_PREC_BF16X3) return BK_ERR_ARG;
  bk::GemmArgs g;
  g.A_hi = g.B_hi = static_cast<const __nv_bfloat16*>(xt_hi);
  g.A_lo = g.B_lo = static_cast<const __nv_bfloat16*>(xt_lo);
  g.lda = g.ldb = ldt;
  g.M = g.N = dprime;
  g.K = n;
  g.batch = 1;
  g.nparts = precision;
  g.flags = bk::kSyrkLower | bk::kMirror;
  g.alpha = alpha;
  g.beta = beta;
  g.C = state;
  g.ldc = ld_state;
  return bk::launch_umma_gemm(g, as_stream(stream));
}

int bk_syrk_accum(float* state, long long ld_state, const float* x, long long ldx, int n, int d,
                  int has_bias, float in_scale, float alpha, float beta, int precision,
                  void* workspace, size_t workspace_bytes, void* stream) {
  if (state == nullptr || x == nullptr || n <= 0 || d <= 0) return BK_ERR_ARG;
  const int dp = d + (has_bias ? 1 : 0);
  if (ld_state < dp || ldx < d) return BK_ERR_ARG;
  if (dp <= BK_SMALL_D_MAX) {
    return bk::launch_small_syrk(state, ld_state, x, ldx, n, d, has_bias, in_scale, alpha, beta,
                                 as_stream(stream));
  }
  if (precision == BK_PREC_FP32) {
    return bk::launch_syrk_fp32(state, ld_state, x, ldx, n, d, has_bias, in_scale, alpha, beta,
                                as_stream(stream));
  }
  if (precision != BK_PREC_BF16 && precision != BK_PREC_BF16X3) return BK_ERR_ARG;
  const size_t need = bk_syrk_workspace_bytes(n, d, has_bias, precision);
  if (workspace == nullptr || workspace_bytes < need ||
      (reinterpret_cast<uintptr_t>(workspace) & 255) != 0)
    return BK_ERR_WORKSPACE;
  const long long ldt = round8(n);
  const size_t one = align_up(static_cast<size_t>(dp) * ldt * 2, 256);
  char* base = static_cast<char*>(workspace);
  if (has_bias && ((ldx % 4) != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0)) {
    // unaligned rows: generic staging kernel with an explicit row of ones, SYRK on d + 1 rows
    __nv_bfloat16* hi1 = reinterpret_cast<__nv_bfloat16*>(base);
    __nv_bfloat16* lo1 =
        precision == BK_PREC_BF16X3 ? reinterpret_cast<__nv_bfloat16*>(base + one) : nullptr;
    int rc1 = bk::launch_transpose_split(x, ldx, n, d, in_scale, 1, hi1, lo1, ldt, as_stream(stream));
    if (rc1) return rc1;
    return bk_syrk_accum_staged(state, ld_state, hi1, lo1, ldt, n, dp, alpha, beta, precision,
                                stream);
  }
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(base);
  __nv_bfloat16* lo =
      precision == BK_PREC_BF16X3 ? reinterpret_cast<__nv_bfloat16*>(base + one) : nullptr;
  float* colsum = has_bias
                      ? reinterpret_cast<float*>(base + (precision == BK_PREC_BF16X3 ? 2 : 1) * one)
                      : nullptr;
  cudaStream_t st = as_stream(stream);
  if (colsum != nullptr &&
      cudaMemsetAsync(colsum, 0, static_cast<size_t>(d) * 4, st) != cudaSuccess)
    return BK_ERR_CUDA;
  int rc = bk::launch_transpose_split(x, ldx, n, d, in_scale, 0, hi, lo, ldt, st, colsum);
  if (rc) return rc;
  rc = bk_syrk_accum_staged(state, ld_state, hi, lo, ldt, n, d, alpha, beta, precision, stream);
  if (rc) return rc;
  if (has_bias)
    rc = bk::launch_bias_border(state, ld_state, d, colsum, alpha, beta, static_cast<float>(n), st);
  return rc;
}

// Grouped factor update: every wide, TMA-addressable factor of the batch is staged and then
// accumulated by ONE persistent tensor-core launch (bk::launch_umma_syrk_grouped); the rest (small,
// fp32-precision or unaligned factors) go through bk_syrk_accum one by one.
size_t bk_syrk_grouped_workspace_bytes(const int* ns, const int* ds, const int* has_bias, int count,
                                       int precision) {
  size_t total = 0;
  for (int i = 0; i < count; ++i)
    total += align_up(bk_syrk_workspace_bytes(ns[i], ds[i], has_bias[i], precision), 256);
  return total;
}

namespace {

// Per-device side stream + events for the staging / SYRK pipeline of bk_syrk_accum_grouped.
struct StagePipe {
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, first = nullptr;
  cudaEvent_t staged[8] = {};
  bool ok = false;
  StagePipe() {
    ok = cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&first, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; ok && i < 8; ++i)
      ok = cudaEventCreateWithFlags(&staged[i], cudaEventDisableTiming) == cudaSuccess;
  }
};

int device_sms() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
    return 148;
  return n;
}

StagePipe* stage_pipe() {
  static std::mutex mu;
  static StagePipe* pipes[bk::kMaxDevices] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= bk::kMaxDevices) return nullptr;
  std::lock_guard<std::mutex> g(mu);
  if (pipes[dev] == nullptr) pipes[dev] = new StagePipe();
  return pipes[dev]->ok ? pipes[dev] : nullptr;
}

}  // namespace

int bk_syrk_accum_grouped(float* const* states, const long long* ld_states, const void* const* xs,
                          const int* x_is_bf16, const long long* ldxs, const int* ns, const int* ds,
                          const int* has_bias, const float* in_scales, const float* alphas,
                          const float* betas, int count, int precision, int flags, void* workspace,
                          size_t workspace_bytes, void* stream) {
  if (count <= 0) return BK_OK;
  if (states == nullptr || ld_states == nullptr || xs == nullptr || ldxs == nullptr ||
      ns == nullptr || ds == nullptr || has_bias == nullptr || in_scales == nullptr ||
      alphas == nullptr || betas == nullptr)
    return BK_ERR_ARG;
  if (bk_syrk_grouped_workspace_bytes(ns, ds, has_bias, count, precision) > workspace_bytes ||
      (workspace == nullptr && workspace_bytes > 0) ||
      (reinterpret_cast<uintptr_t>(workspace) & 255) != 0)
    return BK_ERR_WORKSPACE;
  const bool mirror = (flags & BK_SYRK_LOWER_ONLY) == 0;
  cudaStream_t st = as_stream(stream);
  char* base = static_cast<char*>(workspace);
  size_t off = 0;
  // pass 1: route every factor; the tensor-core items are collected, the others run right away
  struct Wide {
    int idx;
    char* ws;
  } wide[64], direct[64];
  int n_wide = 0, n_direct = 0;
  for (int i = 0; i < count; ++i) {
    const size_t need = bk_syrk_workspace_bytes(ns[i], ds[i], has_bias[i], precision);
    char* ws = base + off;
    off += align_up(need, 256);
    const int d = ds[i], n = ns[i], hb = has_bias[i] ? 1 : 0;
    const int dp = d + hb;
    const bool tensor = dp > BK_SMALL_D_MAX && precision != BK_PREC_FP32;
    const bool bf16_in = x_is_bf16 != nullptr && x_is_bf16[i] != 0;
    const bool aligned = (ld_states[i] % 4) == 0 && (reinterpret_cast<uintptr_t>(states[i]) & 15) == 0 &&
                         (ldxs[i] % (bf16_in ? 8 : 4)) == 0 && (reinterpret_cast<uintptr_t>(xs[i]) & 15) == 0 &&
                         (betas[i] == 0.f || betas[i] == 1.f) && d >= 192;
    if (bf16_in) {
      // bf16 activations feed the tensor cores as they are (row-major = MN-major operand of X^T X): the
      // caller converts the factors this path cannot take (narrow / unaligned) to fp32 first
      if (!tensor || !aligned || n_direct >= 64) return BK_ERR_ARG;
      if (states[i] == nullptr || xs[i] == nullptr || n <= 0 || ld_states[i] < dp || ldxs[i] < d)
        return BK_ERR_ARG;
      direct[n_direct++] = {i, ws};
      continue;
    }
    if (!tensor || !aligned || n_wide >= 64) {
      const int rc = bk_syrk_accum(states[i], ld_states[i], static_cast<const float*>(xs[i]), ldxs[i], n, d,
                                   hb, in_scales[i], alphas[i], betas[i], precision, ws, need, stream);
      if (rc) return rc;
      continue;
    }
    if (states[i] == nullptr || xs[i] == nullptr || n <= 0 || ld_states[i] < dp || ldxs[i] < d)
      return BK_ERR_ARG;
    wide[n_wide++] = {i, ws};
  }
  struct Border {
    float* state;
    long long ld;
    int d;
    const float* colsum;
    float alpha, beta, n;
  } borders[128];
  int n_borders = 0;
  StagePipe* pipe = (flags & BK_SYRK_NO_OVERLAP) ? nullptr : stage_pipe();
  // the pipeline's events are per device, not per caller: one update at a time uses them
  static std::mutex pipe_mu;
  std::unique_lock<std::mutex> pipe_lock(pipe_mu, std::defer_lock);
  if (pipe != nullptr) pipe_lock.lock();
  bool forked = false;
  auto fork_side = [&]() -> int {
    // the side stream may touch the workspace only after everything already queued on `st` (the previous
    // update's SYRKs read the same staging buffers)
    if (forked) return 0;
    if (cudaEventRecord(pipe->fork, st) != cudaSuccess ||
        cudaStreamWaitEvent(pipe->side, pipe->fork, 0) != cudaSuccess)
      return BK_ERR_CUDA;
    forked = true;
    return 0;
  };
  // pass 1b: bf16 activations — no staging at all.  The column sums for the bias row run on the side stream
  // (HBM-bound, 33 MB per 4096^2 operand) underneath the tensor-core launch; the input scale of the second
  // factor (grad_output * N, models/curvatures.py:323) moves into alpha: (s g)(s g)^T = s^2 g g^T.
  if (n_direct > 0) {
    bk::SyrkGroupItem ditems[64];
    cudaStream_t cs = pipe != nullptr ? pipe->side : st;
    bool any_sum = false;
    for (int w = 0; w < n_direct; ++w) {
      const int i = direct[w].idx;
      const int d = ds[i], n = ns[i], hb = has_bias[i] ? 1 : 0;
      bk::SyrkGroupItem& it = ditems[w];
      it.X_hi = static_cast<const __nv_bfloat16*>(xs[i]);
      it.X_lo = nullptr;
      it.ldx = ldxs[i];
      it.d = d;
      it.n = n;
      it.alpha = alphas[i] * in_scales[i] * in_scales[i];
      it.beta = betas[i];
      it.C = states[i];
      it.ldc = ld_states[i];
      if (hb) {
        float* colsum = reinterpret_cast<float*>(direct[w].ws);
        if (pipe != nullptr) {
          const int rc = fork_side();
          if (rc) return rc;
        }
        if (cudaMemsetAsync(colsum, 0, static_cast<size_t>(d) * 4, cs) != cudaSuccess) return BK_ERR_CUDA;
        const int rc = bk::launch_colsum_bf16(it.X_hi, it.ldx, n, d, in_scales[i], colsum, cs);
        if (rc) return rc;
        any_sum = true;
        borders[n_borders++] = {states[i], ld_states[i], d, colsum, alphas[i], betas[i], static_cast<float>(n)};
      }
    }
    if (any_sum && pipe != nullptr && cudaEventRecord(pipe->staged[7], pipe->side) != cudaSuccess)
      return BK_ERR_CUDA;
    for (int g0 = 0; g0 < n_direct; g0 += 8) {
      const int gc = n_direct - g0 < 8 ? n_direct - g0 : 8;
      const int rc = bk::launch_umma_syrk_grouped(ditems + g0, gc, 1, mirror, st, true);
      if (rc) return rc;
    }
    if (any_sum && pipe != nullptr && cudaStreamWaitEvent(st, pipe->staged[7], 0) != cudaSuccess)
      return BK_ERR_CUDA;
  }
  auto finish_borders = [&]() -> int {
    for (int b = 0; b < n_borders; ++b) {
      const int rc = bk::launch_bias_border(borders[b].state, borders[b].ld, borders[b].d,
                                            borders[b].colsum, borders[b].alpha, borders[b].beta,
                                            borders[b].n, st);
      if (rc) return rc;
    }
    return BK_OK;
  };
  if (n_wide == 0) return finish_borders();
  // pass 2: software pipeline over chunks of wide factors.  The staging pass (fp32 [n, d] -> K-major bf16,
  // HBM-bound) of chunk c + 1 runs on a side stream underneath the tensor-core SYRK of chunk c (the
  // lower-only SYRK leaves enough shared memory per SM for a co-resident staging CTA).  Chunk 0 is a single
  // factor so that the tensor cores start after ONE staging pass; later chunks hold up to 3.
  int chunk_begin[66];
  int n_chunks = 0;
  if (pipe != nullptr && n_wide > 1) {
    int b = 0;
    chunk_begin[n_chunks++] = 0;
    b = 1;
    while (b < n_wide) {
      chunk_begin[n_chunks++] = b;
      b += (n_wide - b) >= 6 ? 3 : (n_wide - b);   // ... 3, then whatever is left (<= 5)
      if (n_chunks >= 7) break;
    }
    if (b < n_wide) {  // more than the pipeline's events cover: last chunk takes the rest
      // (groups of 8 are split again at launch)
    }
  } else {
    chunk_begin[n_chunks++] = 0;
  }
  chunk_begin[n_chunks] = n_wide;
  bk::SyrkGroupItem items[64];
  auto stage_chunk = [&](int c, cudaStream_t s) -> int {
    for (int w = chunk_begin[c]; w < chunk_begin[c + 1]; ++w) {
      const int i = wide[w].idx;
      const int d = ds[i], n = ns[i], hb = has_bias[i] ? 1 : 0;
      const long long ldt = round8(n);
      const size_t one = align_up(static_cast<size_t>(d + hb) * ldt * 2, 256);
      char* ws = wide[w].ws;
      __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(ws);
      __nv_bfloat16* lo =
          precision == BK_PREC_BF16X3 ? reinterpret_cast<__nv_bfloat16*>(ws + one) : nullptr;
      float* colsum =
          hb ? reinterpret_cast<float*>(ws + (precision == BK_PREC_BF16X3 ? 2 : 1) * one) : nullptr;
      if (colsum != nullptr &&
          cudaMemsetAsync(colsum, 0, static_cast<size_t>(d) * 4, s) != cudaSuccess)
        return BK_ERR_CUDA;
      // A/B switch BK_SYRK_STAGE_PERSISTENT: staging passes on the side stream (= underneath a running SYRK) use
      // the persistent one-CTA-per-SM variant.  MEASURED SLOWER than the tile grid (cfg5 step 0.49 ms): 0.58 ms with
      // one tile's loads in flight per SM, 0.64 ms with three (114 registers) - slower even than no overlap
      // (0.54 - 0.57 ms); off by default, tools/gpu_time_update_ab.py.
      const int rc = bk::launch_transpose_split(static_cast<const float*>(xs[i]), ldxs[i], n, d, in_scales[i],
                                                0, hi, lo, ldt, s, colsum,
                                                (s != st && (flags & BK_SYRK_STAGE_PERSISTENT)) ? device_sms() : 0);
      if (rc) return rc;
      bk::SyrkGroupItem& it = items[w];
      it.X_hi = hi;
      it.X_lo = lo;
      it.ldx = ldt;
      it.d = d;
      it.n = n;
      it.alpha = alphas[i];
      it.beta = betas[i];
      it.C = states[i];
      it.ldc = ld_states[i];
      if (hb) borders[n_borders++] = {states[i], ld_states[i], d, colsum, alphas[i], betas[i],
                                      static_cast<float>(n)};
    }
    return 0;
  };
  auto syrk_chunk = [&](int c) -> int {
    for (int g0 = chunk_begin[c]; g0 < chunk_begin[c + 1]; g0 += 8) {
      const int gc = chunk_begin[c + 1] - g0 < 8 ? chunk_begin[c + 1] - g0 : 8;
      const int rc = bk::launch_umma_syrk_grouped(items + g0, gc, precision, mirror, st);
      if (rc) return rc;
    }
    return 0;
  };
  if (n_chunks == 1) {
    int rc = stage_chunk(0, st);
    if (rc) return rc;
    rc = syrk_chunk(0);
    if (rc) return rc;
  } else {
    int rc = fork_side();
    if (rc) return rc;
    rc = stage_chunk(0, st);
    if (rc) return rc;
    // ... and starts staging chunk 1 when chunk 0 is staged (not before: the first SYRK should not wait
    // for a staging pass that shares HBM with two others)
    if (cudaEventRecord(pipe->first, st) != cudaSuccess ||
        cudaStreamWaitEvent(pipe->side, pipe->first, 0) != cudaSuccess)
      return BK_ERR_CUDA;
    for (int c = 1; c < n_chunks; ++c) {
      rc = stage_chunk(c, pipe->side);
      if (rc) return rc;
      if (cudaEventRecord(pipe->staged[c - 1], pipe->side) != cudaSuccess) return BK_ERR_CUDA;
    }
    rc = syrk_chunk(0);
    if (rc) return rc;
    for (int c = 1; c < n_chunks; ++c) {
      if (cudaStreamWaitEvent(st, pipe->staged[c - 1], 0) != cudaSuccess) return BK_ERR_CUDA;
      rc = syrk_chunk(c);
      if (rc) return rc;
    }
  }
  return finish_borders();
}

int bk_syrk_accum_staged_grouped(float* const* states, const long long* ld_states,
                                 const void* const* xt_his, const void* const* xt_los,
                                 const long long* ldts, const int* ns, const int* ds,
                                 const float* alphas, const float* betas, int count, int precision,
                                 int flags, void* stream) {
  if (count <= 0) return BK_OK;
  if (count > 8 || states == nullptr || ld_states == nullptr || xt_his == nullptr ||
      ldts == nullptr || ns == nullptr || ds == nullptr || alphas == nullptr || betas == nullptr)
    return BK_ERR_ARG;
  if (precision != BK_PREC_BF16 && precision != BK_PREC_BF16X3) return BK_ERR_ARG;
  bk::SyrkGroupItem items[8];
  for (int i = 0; i < count; ++i) {
    items[i].X_hi = static_cast<const __nv_bfloat16*>(xt_his[i]);
    items[i].X_lo = xt_los != nullptr ? static_cast<const __nv_bfloat16*>(xt_los[i]) : nullptr;
    items[i].ldx = ldts[i];
    items[i].d = ds[i];
    items[i].n = ns[i];
    items[i].alpha = alphas[i];
    items[i].beta = betas[i];
    items[i].C = states[i];
    items[i].ldc = ld_states[i];
  }
  return bk::launch_umma_syrk_grouped(items, count, precision, (flags & BK_SYRK_LOWER_ONLY) == 0,
                                      as_stream(stream), (flags & BK_SYRK_ROW_MAJOR) != 0);
}

int bk_sym_finalize(float* const* factors_host, const long long* ld_host, const int* dims_host, int count,
                    float scale, void* stream) {
  if (count < 0 || (count > 0 && (factors_host == nullptr || ld_host == nullptr || dims_host == nullptr)))
    return BK_ERR_ARG;
  return bk::launch_sym_finalize(factors_host, ld_host, dims_host, count, scale, as_stream(stream));
}

int bk_conv_a_accum(float* state, long long ld_state, const float* x, int n, int c, int h, int w,
                    int kh, int kw, int pad_h, int pad_w, int stride_h, int stride_w, int has_bias,
                    float alpha, float beta, void* stream) {
  if (state == nullptr || x == nullptr) return BK_ERR_ARG;
  return bk::launch_conv_a_syrk(state, ld_state, x, n, c, h, w, kh, kw, pad_h, pad_w, stride_h,
                                stride_w, has_bias, alpha, beta, as_stream(stream));
}

int bk_im2col_split(const float* x, int n, int c, int h, int w, int kh, int kw, int pad_h, int pad_w,
                    int stride_h, int stride_w, float scale, int ones_row, void* t_hi, void* t_lo, long long ldt,
                    void* stream) {
  return bk::launch_im2col_split(x, n, c, h, w, kh, kw, pad_h, pad_w, stride_h, stride_w, scale, ones_row,
                                 static_cast<__nv_bfloat16*>(t_hi), static_cast<__nv_bfloat16*>(t_lo), ldt,
                                 as_stream(stream));
}

int bk_conv_g_accum(float* state, long long ld_state, const float* g, int n, int o, int hw,
                    float in_scale, float alpha, float beta, void* stream) {
  if (state == nullptr || g == nullptr) return BK_ERR_ARG;
  return bk::launch_conv_g_syrk(state, ld_state, g, n, o, hw, in_scale, alpha, beta,
                                as_stream(stream));
}

// ------------------------------------------------------------------------------------ diagonal
int bk_diag_accum(float* state, const float* wgrad, const float* bgrad, int d_out, int d_in,
                  float scale, float beta, void* stream) {
  if (state == nullptr || wgrad == nullptr) return BK_ERR_ARG;
  return bk::launch_diag_accum(state, wgrad, bgrad, d_out, d_in, scale, beta, as_stream(stream));
}
int bk_diag_invert(float* inv, const float* state, long long count, float add, float multiply,
                   void* stream) {
  if (inv == nullptr || state == nullptr) return BK_ERR_ARG;
  return bk::launch_diag_invert(inv, state, count, add, multiply, as_stream(stream));
}
int bk_diag_sample(float* out, const float* inv, long long count, int nsamples,
                   unsigned long long seed, unsigned sample0, unsigned stream_id,
                   const float* z_or_null, void* stream) {
  if (out == nullptr || inv == nullptr) return BK_ERR_ARG;
  return bk::launch_diag_sample(out, inv, count, nsamples, seed, sample0, stream_id, z_or_null,
                                as_stream(stream));
}
int bk_diag_quadform(float* out, const float* j, long long ldj, const float* h, long long count,
                     int batch, void* stream) {
  if (out == nullptr || j == nullptr || h == nullptr) return BK_ERR_ARG;
  return bk::launch_diag_quadform(out, j, ldj, h, count, batch, as_stream(stream));
}

// ------------------------------------------------------------------------------------ inversion
size_t bk_chol_inv_workspace_bytes(const int* dims_host, int count) {
  if (dims_host == nullptr || count <= 0) return 0;
  return bk::chol_inv_workspace_bytes(dims_host, count);
}

int bk_damp_chol_inv_batched(const float* const* factors_host, float* const* outs_host,
                             const int* dims_host, const float* add_host,
                             const float* multiply_host, int count, void* workspace,
                             size_t workspace_bytes, void* stream) {
  if (factors_host == nullptr || outs_host == nullptr || dims_host == nullptr ||
      add_host == nullptr || multiply_host == nullptr)
    return BK_ERR_ARG;
  return bk::chol_inv_batched(factors_host, outs_host, dims_host, add_host, multiply_host, count,
                              workspace, workspace_bytes, as_stream(stream));
}

// ------------------------------------------------------------------------- eigen / dense Fisher
size_t bk_eigh_workspace_bytes(const int* dims_host, int count) {
  if (dims_host == nullptr || count <= 0) return 0;
  return bk::eigh_workspace_bytes(dims_host, count);
}

int bk_eigh_batched(const float* const* factors_host, const long long* ld_host,
                    float* const* evals_host, float* const* evecs_host, const int* dims_host,
                    int count, float sym_scale, int max_sweeps, void* workspace,
                    size_t workspace_bytes, void* stream) {
  if (factors_host == nullptr || ld_host == nullptr || evals_host == nullptr || dims_host == nullptr)
    return BK_ERR_ARG;
  return bk::eigh_batched(factors_host, ld_host, evals_host, evecs_host, dims_host, count, sym_scale,
                          max_sweeps, workspace, workspace_bytes, as_stream(stream));
}

int bk_dominance(const float* h, long long ld, int p, float tau, const int* block_begin,
                 const int* block_end, int nblocks, double* out3, void* stream) {
  if (h == nullptr || out3 == nullptr || ld < p) return BK_ERR_ARG;
  if (nblocks > 0 && (block_begin == nullptr || block_end == nullptr)) return BK_ERR_ARG;
  return bk::launch_dominance(h, ld, 0, p, p, tau, block_begin, block_end, nblocks, out3,
                              as_stream(stream));
}

int bk_dominance_rows(const float* rows, long long ld, int row0, int nrows, int p, float tau,
                      const int* block_begin, const int* block_end, int nblocks, double* out3, void* stream) {
  if (rows == nullptr || out3 == nullptr || ld < p || row0 < 0 || nrows < 0 || row0 + nrows > p) return BK_ERR_ARG;
  if (nblocks > 0 && (block_begin == nullptr || block_end == nullptr)) return BK_ERR_ARG;
  return bk::launch_dominance(rows, ld, row0, nrows, p, tau, block_begin, block_end, nblocks, out3,
                              as_stream(stream));
}

int bk_ger_accum(float* state, long long ld, const float* g, int p, float alpha, float beta,
                 void* stream) {
  if (state == nullptr || g == nullptr || ld < p) return BK_ERR_ARG;
  return bk::launch_ger_accum(state, ld, g, p, alpha, beta, as_stream(stream));
}

int bk_kron(const float* a, int m, int n, const float* b, int p, int q, float* out, void* stream) {
  if (a == nullptr || b == nullptr || out == nullptr) return BK_ERR_ARG;
  return bk::launch_kron(a, m, n, b, p, q, out, as_stream(stream));
}

// ----------------------------------------------------------------------- multi-GPU exchange packing
int bk_tri_pack(const float* const* factors_host, const long long* ld_host, const int* dims_host, int count,
                float* packed, void* stream) {
  if (count < 0 || (count > 0 && (factors_host == nullptr || ld_host == nullptr || dims_host == nullptr ||
                                  packed == nullptr)))
    return BK_ERR_ARG;
  return bk::launch_tri_pack(factors_host, ld_host, dims_host, count, packed, as_stream(stream));
}

int bk_tri_unpack(float* const* outs_host, const long long* ld_host, const int* dims_host, int count,
                  const float* packed, float scale, int mirror, void* stream) {
  if (count < 0 || (count > 0 && (outs_host == nullptr || ld_host == nullptr || dims_host == nullptr ||
                                  packed == nullptr)))
    return BK_ERR_ARG;
  return bk::launch_tri_unpack(outs_host, ld_host, dims_host, count, packed, scale, mirror ? 1 : 0,
                               as_stream(stream));
}

// ----------------------------------------------------------------------- peer-memory exchange
int bk_peer_alloc(size_t bytes, void** ptr) {
  if (ptr == nullptr || bytes == 0) return BK_ERR_ARG;
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    return BK_ERR_CUDA;
  }
  if (cudaMemset(p, 0, bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    cudaFree(p);
    return BK_ERR_CUDA;
  }
  *ptr = p;
  return BK_OK;
}

int bk_peer_free(void* ptr) { return (ptr == nullptr || cudaFree(ptr) == cudaSuccess) ? BK_OK : BK_ERR_CUDA; }

int bk_peer_export(void* ptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (ptr == nullptr || handle64 == nullptr) return BK_ERR_ARG;
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, ptr) != cudaSuccess) {
    cudaGetLastError();
    return BK_ERR_CUDA;
  }
  memcpy(handle64, &h, 64);
  return BK_OK;
}

int bk_peer_open(const void* handle64, void** ptr) {
  if (ptr == nullptr || handle64 == nullptr) return BK_ERR_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
    cudaGetLastError();
    return BK_ERR_CUDA;
  }
  *ptr = p;
  return BK_OK;
}

int bk_peer_close(void* ptr) {
  return (ptr == nullptr || cudaIpcCloseMemHandle(ptr) == cudaSuccess) ? BK_OK : BK_ERR_CUDA;
}

int bk_peer_read_u32(const void* ptr, unsigned int* out_host) {
  if (ptr == nullptr || out_host == nullptr) return BK_ERR_ARG;
  return cudaMemcpy(out_host, ptr, 4, cudaMemcpyDeviceToHost) == cudaSuccess ? BK_OK : BK_ERR_CUDA;
}

long long bk_tile_packed_floats(const int* dims_host, int count) {
  if (count <= 0 || dims_host == nullptr) return 0;
  return bk::tile_packed_floats(dims_host, count);
}

int bk_tile_pack(const float* const* factors_host, const long long* ld_host, const int* dims_host,
                 const long long* offsets_host, int count, float* packed, void* stream) {
  if (count <= 0 || factors_host == nullptr || ld_host == nullptr || dims_host == nullptr || packed == nullptr)
    return BK_ERR_ARG;
  return bk::launch_tile_pack(factors_host, ld_host, dims_host, offsets_host, count, packed, as_stream(stream));
}

int bk_tile_pack_to(const float* const* factors_host, const long long* ld_host, const int* dims_host,
                    float* const* dsts_host, int count, void* stream) {
  if (count <= 0 || factors_host == nullptr || ld_host == nullptr || dims_host == nullptr || dsts_host == nullptr)
    return BK_ERR_ARG;
  return bk::launch_tile_pack_to(factors_host, ld_host, dims_host, dsts_host, count, as_stream(stream));
}

int bk_peer_tile_unpack(float* const* outs_host, const long long* ld_host, const int* dims_host, int count,
                        const float* const* srcs_host, int nsrc, float scale, int mirror, void* stream) {
  if (count <= 0 || outs_host == nullptr || ld_host == nullptr || dims_host == nullptr || srcs_host == nullptr)
    return BK_ERR_ARG;
  return bk::launch_peer_tile_unpack(outs_host, ld_host, dims_host, count, srcs_host, nsrc, scale, mirror ? 1 : 0,
                                     as_stream(stream));
}

int bk_peer_copy(void* dst, const void* src, long long bytes, int vec_bytes, int ctas, void* stream) {
  return bk::launch_peer_copy(dst, src, bytes, vec_bytes, ctas, as_stream(stream));
}

int bk_peer_signal(unsigned int* const* flags_host, int world, int me, unsigned int epoch, void* stream) {
  if (flags_host == nullptr) return BK_ERR_ARG;
  return bk::launch_peer_signal(flags_host, world, me, epoch, as_stream(stream));
}

int bk_peer_wait(const unsigned int* flags_local, int world, unsigned int epoch, double timeout_s, int* err,
                 void* stream) {
  return bk::launch_peer_wait(flags_local, world, epoch, timeout_s, err, as_stream(stream));
}

// ----------------------------------------------------------------------- INF curvature
int bk_inf_regularise(float* correction, long long nm, const float* lambda, long long r, float add,
                      float multiply, float* reg_inv_correction, float* reg_lambda, void* stream) {
  if (nm < 0 || r < 0) return BK_ERR_ARG;
  if (nm > 0 && (correction == nullptr || reg_inv_correction == nullptr)) return BK_ERR_ARG;
  if (r > 0 && (lambda == nullptr || reg_lambda == nullptr)) return BK_ERR_ARG;
  return bk::launch_inf_regularise(correction, nm, lambda, r, add, multiply, reg_inv_correction,
                                   reg_lambda, as_stream(stream));
}

size_t bk_inf_presample_workspace_bytes(int n, int a, int m, int b) {
  return bk::inf_presample_workspace_bytes(n, a, m, b);
}

int bk_inf_presample(const float* ua, long long lda, int n, int a, const float* ug, long long ldg, int m,
                     int b, const float* reg_inv_correction, const float* reg_lambda, float* p_out,
                     void* workspace, size_t workspace_bytes, void* stream) {
  if (ua == nullptr || ug == nullptr || reg_inv_correction == nullptr || reg_lambda == nullptr ||
      p_out == nullptr || workspace == nullptr || lda < a || ldg < b)
    return BK_ERR_ARG;
  return bk::inf_presample(ua, lda, n, a, ug, ldg, m, b, reg_inv_correction, reg_lambda, p_out, workspace,
                           workspace_bytes, as_stream(stream));
}

int bk_inf_combine(float* out, const float* y_l, const float* reg_inv_correction, const float* x_ps_t,
                   long long count, void* stream) {
  if (count > 0 && (out == nullptr || y_l == nullptr || reg_inv_correction == nullptr || x_ps_t == nullptr))
    return BK_ERR_ARG;
  return bk::launch_inf_combine(out, y_l, reg_inv_correction, x_ps_t, count, as_stream(stream));
}

// ----------------------------------------------------------------------- calibration metrics
int bk_calibration_rows(const float* probs, long long ld, const long long* labels, int n, int classes,
                        float* conf, float* correct, float* nll, float* entropy, int* pred, double* totals4,
                        void* stream) {
  if (n < 0 || classes <= 0 || ld < classes || totals4 == nullptr || (n > 0 && probs == nullptr))
    return BK_ERR_ARG;
  return bk::launch_calibration_rows(probs, ld, labels, n, classes, conf, correct, nll, entropy, pred,
                                     totals4, as_stream(stream));
}

int bk_binned_stats(const float* x, const float* w1, const float* w2, long long n, const double* edges,
                    int nbins, int mode, double* out3, void* stream) {
  if (n < 0 || edges == nullptr || out3 == nullptr || (n > 0 && x == nullptr)) return BK_ERR_ARG;
  return bk::launch_binned_stats(x, w1, w2, n, edges, nbins, mode, out3, as_stream(stream));
}

// ----------------------------------------------------------------------- predictive glue
int bk_sample_to_weights(const float* samples, const float* mean_w, const float* mean_b, int d_out,
                         int d_in, int has_bias, int nsamples, float* w_f32, void* w_hi, void* w_lo,
                         long long ldw, float* b_f32, void* stream) {
  if (samples == nullptr || mean_w == nullptr) return BK_ERR_ARG;
  if (has_bias && (mean_b == nullptr || b_f32 == nullptr)) return BK_ERR_ARG;
  if (w_hi != nullptr && ldw < d_in) return BK_ERR_ARG;
  return bk::launch_sample_to_weights(samples, mean_w, mean_b, d_out, d_in, has_bias ? 1 : 0,
                                      nsamples, w_f32, static_cast<__nv_bfloat16*>(w_hi),
                                      static_cast<__nv_bfloat16*>(w_lo), ldw, b_f32,
                                      as_stream(stream));
}

int bk_conv2d_relu_pool(const float* in, long long in_sample_stride, const float* w, const float* b,
                        float* out, int nsamples, int n, int c, int h, int wd, int o, int kh, int kw,
                        int sh, int sw, int ph, int pw, int relu, int pool, void* stream) {
  if (in == nullptr || w == nullptr || out == nullptr || sh <= 0 || sw <= 0) return BK_ERR_ARG;
  return bk::launch_conv2d_relu_pool(in, in_sample_stride, w, b, out, nsamples, n, c, h, wd, o, kh,
                                     kw, sh, sw, ph, pw, relu, pool, as_stream(stream));
}

int bk_predictive_moments(const float* logits, int nsamples, int batch, int classes, int mode,
                          float* mean, float* meansq, void* stream) {
  if (logits == nullptr || mean == nullptr) return BK_ERR_ARG;
  return bk::launch_predictive_moments(logits, nsamples, batch, classes, mode, mean, meansq,
                                       as_stream(stream));
}

int bk_frob_dot(float* out, const float* x, long long stride_x, const float* y, long long stride_y,
                long long count, int batch, int absolute, int accumulate, void* stream) {
  if (out == nullptr || x == nullptr || y == nullptr) return BK_ERR_ARG;
  return bk::launch_frob_dot(out, x, stride_x, y, stride_y, count, batch, absolute, accumulate,
                             as_stream(stream));
}

int bk_spd_inverse_f64(const float* const* factors_host, const long long* ld_host, const int* dims_host,
                       const double* add_host, const double* multiply_host, double* const* outs_host,
                       int count, int* status, void* stream) {
  if (count < 0 || (count > 0 && (factors_host == nullptr || ld_host == nullptr || dims_host == nullptr ||
                                  add_host == nullptr || multiply_host == nullptr || outs_host == nullptr)))
    return BK_ERR_ARG;
  return bk::launch_spd_inverse_f64(factors_host, ld_host, dims_host, add_host, multiply_host, outs_host,
                                    count, status, as_stream(stream));
}

int bk_kron_quadform_f64(const float* v, long long stride_v, int batch, int d_in_p, int d_out,
                         const double* q, const double* h, float* out, int accumulate, void* stream) {
  if (batch > 0 && (v == nullptr || q == nullptr || h == nullptr || out == nullptr)) return BK_ERR_ARG;
  return bk::launch_kron_quadform_f64(v, stride_v, batch, d_in_p, d_out, q, h, out, accumulate,
                                      as_stream(stream));
}

int bk_chol_trinv_f64(const float* f, long long ldf, int d, double add, float* w, long long ldw, int* status,
                      void* stream) {
  return bk::launch_chol_trinv_f64(f, ldf, d, add, w, ldw, status, as_stream(stream));
}

int bk_band_mask(float* h, long long ld, int p, float tau, int add_tau_in_place, const int* row_lo,
                 const int* row_hi, float* out, long long ldo, void* stream) {
  return bk::launch_band_mask(h, ld, p, tau, add_tau_in_place, row_lo, row_hi, out, ldo, as_stream(stream));
}

int bk_block_inverse(const float* res, long long ld, int p, const int* comp_begin, const int* comp_end,
                     int ncomp, int max_dim, double scale, float* out, long long ldo, int zero_fill, int* status,
                     void* stream) {
  return bk::launch_block_inverse(res, ld, p, comp_begin, comp_end, ncomp, max_dim, scale, out, ldo, zero_fill,
                                  status, as_stream(stream));
}

}  // extern "C"
#pragma GCC visibility pop
