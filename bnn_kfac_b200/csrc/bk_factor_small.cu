// bk_factor_small.cu — SIMT fp32 split-K SYRK for skinny Kronecker factors (d' <= 176).
//
// These factors (conv layers, small Linear layers: 5..161 wide) have a tiny output and a long
// reduction axis (conv1 of BaseNet_15k: 26 x 26 output over 147 456 patch columns), so they are
// bound by reading the activations once from HBM, not by math: no tensor cores, no reshaping into a
// GEMM.  One CTA owns a contiguous slice of the reduction axis, stages rows in shared memory through
// a loader functor (dense rows, implicit im2col gather, or the NCHW->[O, N*H*W] view of the output
// gradient), keeps a register tile of the d' x d' partial product and finishes with fp32 atomics.
//
// Reference semantics (paths relative to /root/reference):
//   Linear A / G : models/curvatures.py:345-349, 355-356
//   Conv2d A     : models/curvatures.py:341-349 (F.unfold -> [C*kh*kw(+1), N*L], / (N*L))
//   Conv2d G     : models/curvatures.py:353-356 ([O, N*H'*W'], / (N*H'*W'))
#include "bk_common.cuh"
#include "bk_kernels.cuh"

namespace bk {

namespace {

constexpr int kRowsPerTile = 32;  // reduction rows staged per iteration
constexpr int kTX = 16, kTY = 16; // thread grid over the output

struct DenseLoader {
  static constexpr bool kRowFast = false;  // consecutive threads walk feature columns (contiguous)
  const float* x;
  long long ldx;
  int d;
  float scale;
  __device__ __forceinline__ float operator()(long long row, int j) const {
    return x[row * ldx + j] * scale;
  }
};

struct ConvALoader {  // implicit im2col: row = (n, oy, ox), column j = (c, ky, kx)
  static constexpr bool kRowFast = true;  // consecutive rows = consecutive ox: contiguous pixels
  const float* x;
  int C, H, W, KH, KW, PH, PW, SH, SW, OH, OW;
  __device__ __forceinline__ float operator()(long long row, int j) const {
    const int L = OH * OW;
    const int n = static_cast<int>(row / L);
    const int p = static_cast<int>(row - static_cast<long long>(n) * L);
    const int oy = p / OW, ox = p - oy * OW;
    const int c = j / (KH * KW);
    const int r = j - c * KH * KW;
    const int ky = r / KW, kx = r - ky * KW;
    const int iy = oy * SH - PH + ky, ix = ox * SW - PW + kx;
    if (iy < 0 || iy >= H || ix < 0 || ix >= W) return 0.f;
    return x[((static_cast<long long>(n) * C + c) * H + iy) * W + ix];
  }
  // The same gather with the index arithmetic hoisted: a thread of the staging loop keeps ONE patch row (n, oy, ox)
  // per tile and walks the patch columns j = j0, j0 + 8, ...; (c, ky, kx) advance incrementally.  (Per element the
  // generic form above costs six integer divisions - more instructions than the FMAs the element feeds.)
  struct Row {
    const float* base;  // image n
    int iy0, ix0;
  };
  __device__ __forceinline__ Row row(long long r) const {
    const int L = OH * OW;
    const int n = static_cast<int>(r / L);
    const int p = static_cast<int>(r - static_cast<long long>(n) * L);
    const int oy = p / OW, ox = p - oy * OW;
    return Row{x + static_cast<long long>(n) * C * H * W, oy * SH - PH, ox * SW - PW};
  }
  __device__ __forceinline__ float at(const Row& rw, int c, int ky, int kx) const {
    const int iy = rw.iy0 + ky, ix = rw.ix0 + kx;
    if (iy < 0 || iy >= H || ix < 0 || ix >= W) return 0.f;
    return rw.base[(c * H + iy) * W + ix];
  }
};

template <class L>
struct IsConvA {
  static constexpr bool value = false;
};
template <>
struct IsConvA<ConvALoader> {
  static constexpr bool value = true;
};

struct ConvGLoader {  // g is [N, O, HW]; row = (n, p), column j = o
  static constexpr bool kRowFast = true;  // consecutive rows = consecutive p: contiguous
  const float* g;
  int O, HW;
  float scale;
  __device__ __forceinline__ float operator()(long long row, int j) const {
    const int n = static_cast<int>(row / HW);
    const int p = static_cast<int>(row - static_cast<long long>(n) * HW);
    return g[(static_cast<long long>(n) * O + j) * HW + p] * scale;
  }
};

template <int T, class Loader>
__global__ void __launch_bounds__(kTX* kTY)
small_syrk_kernel(float* __restrict__ state, long long ld_state, Loader load, long long nrows,
                  int d, int has_bias, float alpha, long long rows_per_cta) {
  extern __shared__ float xs[];  // [kRowsPerTile][dpad]
  const int dp = d + has_bias;
  const int dpad = T * 16 + 1;
  const int tx = threadIdx.x % kTX, ty = threadIdx.x / kTX;
  const long long row_begin = blockIdx.x * rows_per_cta;
  long long row_end = row_begin + rows_per_cta;
  if (row_end > nrows) row_end = nrows;

  float acc[T][T];
#pragma unroll
  for (int i = 0; i < T; ++i)
#pragma unroll
    for (int j = 0; j < T; ++j) acc[i][j] = 0.f;

  for (long long r0 = row_begin; r0 < row_end; r0 += kRowsPerTile) {
    const int nr = static_cast<int>(min(static_cast<long long>(kRowsPerTile), row_end - r0));
    // stage: consecutive threads walk whichever axis is contiguous in HBM for this loader
    if constexpr (IsConvA<Loader>::value) {
      // thread -> (patch row k = tid % 32, patch columns j = tid / 32 + 8 m): one row decode per tile, incremental
      // (c, ky, kx)
      const int k = threadIdx.x % kRowsPerTile;
      const bool live = k < nr;
      typename Loader::Row rw{};
      if (live) rw = load.row(r0 + k);
      int j = threadIdx.x / kRowsPerTile;
      int c = j / (load.KH * load.KW);
      int rem = j - c * load.KH * load.KW;
      int ky = rem / load.KW, kx = rem - ky * load.KW;
      for (; j < T * 16; j += (kTX * kTY) / kRowsPerTile) {
        float v = 0.f;
        if (live) {
          if (j < d) v = load.at(rw, c, ky, kx);
          else if (j == d && has_bias) v = 1.f;
        }
        xs[k * dpad + j] = v;
        kx += (kTX * kTY) / kRowsPerTile;
        while (kx >= load.KW) {
          kx -= load.KW;
          ++ky;
        }
        while (ky >= load.KH) {
          ky -= load.KH;
          ++c;
        }
      }
    } else
    for (int idx = threadIdx.x; idx < kRowsPerTile * (T * 16); idx += kTX * kTY) {
      int k, j;
      if (Loader::kRowFast) {
        j = idx / kRowsPerTile;
        k = idx - j * kRowsPerTile;
      } else {
        k = idx / (T * 16);
        j = idx - k * (T * 16);
      }
      float v = 0.f;
      if (k < nr) {
        if (j < d) v = load(r0 + k, j);
        else if (j == d && has_bias) v = 1.f;
      }
      xs[k * dpad + j] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < kRowsPerTile; ++k) {
      float a[T], b[T];
#pragma unroll
      for (int i = 0; i < T; ++i) a[i] = xs[k * dpad + ty + 16 * i];
#pragma unroll
      for (int j = 0; j < T; ++j) b[j] = xs[k * dpad + tx + 16 * j];
#pragma unroll
      for (int i = 0; i < T; ++i)
#pragma unroll
        for (int j = 0; j < T; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < T; ++i) {
    const int r = ty + 16 * i;
    if (r >= dp) continue;
#pragma unroll
    for (int j = 0; j < T; ++j) {
      const int c = tx + 16 * j;
      if (c < dp) atomicAdd(&state[static_cast<long long>(r) * ld_state + c], alpha * acc[i][j]);
    }
  }
}

__global__ void scale_square_kernel(float* __restrict__ s, long long ld, int d, float beta) {
  const int r = blockIdx.y;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < d; c += gridDim.x * blockDim.x) {
    float* p = s + static_cast<long long>(r) * ld + c;
    *p = (beta == 0.f) ? 0.f : (*p * beta);
  }
}

template <class Loader>
int run_small_syrk(float* state, long long ld_state, Loader load, long long nrows, int d,
                   int has_bias, float alpha, float beta, cudaStream_t stream) {
  const int dp = d + has_bias;
  if (dp <= 0 || dp > 176) return -2;
  if (beta != 1.f) {
    dim3 g((dp + 127) / 128, dp), b(128);
    scale_square_kernel<<<g, b, 0, stream>>>(state, ld_state, dp, beta);
    note_launch();
  }
  if (nrows <= 0) return cudaGetLastError() == cudaSuccess ? 0 : -5;
  // split the reduction axis: at most 4 CTAs per SM, at least 4 staged tiles per CTA
  long long ctas = (nrows + 4 * kRowsPerTile - 1) / (4 * kRowsPerTile);
  const long long cap = static_cast<long long>(kNumSMsB200) * 4;
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  long long per = (nrows + ctas - 1) / ctas;
  per = (per + kRowsPerTile - 1) / kRowsPerTile * kRowsPerTile;
  ctas = (nrows + per - 1) / per;
  const int T = (dp + 15) / 16;
#define BK_LAUNCH_T(TT)                                                                          \
  {                                                                                              \
    const size_t smem = sizeof(float) * kRowsPerTile * ((TT) * 16 + 1);                          \
    small_syrk_kernel<TT, Loader><<<static_cast<int>(ctas), kTX * kTY, smem, stream>>>(          \
        state, ld_state, load, nrows, d, has_bias, alpha, per);                                  \
    note_launch();                                                                               \
  }
  if (T <= 1) BK_LAUNCH_T(1)
  else if (T == 2) BK_LAUNCH_T(2)
  else if (T <= 4) BK_LAUNCH_T(4)
  else if (T <= 6) BK_LAUNCH_T(6)
  else if (T <= 8) BK_LAUNCH_T(8)
  else if (T <= 10) BK_LAUNCH_T(10)
  else BK_LAUNCH_T(11)
#undef BK_LAUNCH_T
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace

int launch_small_syrk(float* state, long long ld_state, const float* x, long long ldx, int n, int d,
                      int has_bias, float in_scale, float alpha, float beta, cudaStream_t stream) {
  DenseLoader l{x, ldx, d, in_scale};
  return run_small_syrk(state, ld_state, l, n, d, has_bias ? 1 : 0, alpha, beta, stream);
}

int launch_conv_a_syrk(float* state, long long ld_state, const float* x, int n, int c, int h, int w,
                       int kh, int kw, int pad_h, int pad_w, int stride_h, int stride_w,
                       int has_bias, float alpha, float beta, cudaStream_t stream) {
  if (stride_h <= 0 || stride_w <= 0) return -2;
  const int oh = (h + 2 * pad_h - kh) / stride_h + 1;
  const int ow = (w + 2 * pad_w - kw) / stride_w + 1;
  if (oh <= 0 || ow <= 0) return -2;
  ConvALoader l{x, c, h, w, kh, kw, pad_h, pad_w, stride_h, stride_w, oh, ow};
  return run_small_syrk(state, ld_state, l, static_cast<long long>(n) * oh * ow, c * kh * kw,
                        has_bias ? 1 : 0, alpha, beta, stream);
}

int launch_conv_g_syrk(float* state, long long ld_state, const float* g, int n, int o, int hw,
                       float in_scale, float alpha, float beta, cudaStream_t stream) {
  ConvGLoader l{g, o, hw, in_scale};
  return run_small_syrk(state, ld_state, l, static_cast<long long>(n) * hw, o, 0, alpha, beta,
                        stream);
}

}  // namespace bk
