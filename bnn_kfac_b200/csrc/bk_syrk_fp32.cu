// bk_syrk_fp32.cu — full-fp32 SIMT factor SYRK (precision = BK_PREC_FP32).
//
//   state = beta*state + alpha * [s*x ; 1]^T [s*x ; 1]          x: [n, d] row-major fp32
//
// Same arithmetic class as the reference's fp32 torch.mm (models/curvatures.py:349,356).  It exists
// for parity on badly conditioned factors: the damped inverse amplifies factor differences by
// cond(R) (1e4 at the reference's (0.04, 200) damping on image inputs), so a factor that is only
// bf16x3-accurate (~1e-6) cannot reproduce the reference's inverse to 1e-3 end to end, while an
// fp32 one (~1e-7) can.  The tensor-core SYRK stays the throughput path for wide layers.
//
// x is MN-major for this product (feature index contiguous), which is exactly what a register-tiled
// SIMT kernel wants: both operand panels are read coalesced with no transpose.  128x128 output tile
// per CTA (lower-triangle tiles only, mirrored on write), split-K over grid.y with fp32 atomics.
#include "bk_common.cuh"
#include "bk_kernels.cuh"

namespace bk {

namespace {

constexpr int TS = 128;  // tile side
constexpr int KC = 32;   // reduction rows staged per step
constexpr int kPadF = 4;

__global__ void __launch_bounds__(256)
syrk_fp32_kernel(float* __restrict__ state, long long ld_state, const float* __restrict__ x,
                 long long ldx, int n, int d, int has_bias, float in_scale, float alpha,
                 int rows_per_split) {
  __shared__ float As[KC][TS + kPadF];
  __shared__ float Bs[KC][TS + kPadF];
  const int dp = d + has_bias;
  // lower-triangular tile index -> (ti, tj), tj <= ti
  const int t = blockIdx.x;
  int ti = static_cast<int>((sqrtf(8.f * t + 1.f) - 1.f) * 0.5f);
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  while (ti * (ti + 1) / 2 > t) --ti;
  const int tj = t - ti * (ti + 1) / 2;
  const int r0 = ti * TS, c0 = tj * TS;
  const int k_begin = blockIdx.y * rows_per_split;
  const int k_end = min(n, k_begin + rows_per_split);
  const int tid = threadIdx.x;
  const int ty = tid / 16, tx = tid % 16;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = k_begin; k0 < k_end; k0 += KC) {
    for (int idx = tid; idx < KC * TS; idx += 256) {
      const int kk = idx / TS, i = idx - kk * TS;
      const int k = k0 + kk;
      float a = 0.f, b = 0.f;
      if (k < k_end) {
        const int ja = r0 + i, jb = c0 + i;
        if (ja < d) a = x[static_cast<long long>(k) * ldx + ja] * in_scale;
        else if (ja == d && has_bias) a = 1.f;
        if (jb < d) b = x[static_cast<long long>(k) * ldx + jb] * in_scale;
        else if (jb == d && has_bias) b = 1.f;
      }
      As[kk][i] = a;
      Bs[kk][i] = b;
    }
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < KC; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = r0 + ty * 8 + i;
    if (row >= dp) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = c0 + tx * 8 + j;
      if (col >= dp || col > row) continue;
      const float v = alpha * acc[i][j];
      atomicAdd(&state[static_cast<long long>(row) * ld_state + col], v);
      if (col != row) atomicAdd(&state[static_cast<long long>(col) * ld_state + row], v);
    }
  }
}

__global__ void scale_matrix_kernel(float* __restrict__ s, long long ld, int d, float beta) {
  const int r = blockIdx.y;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < d; c += gridDim.x * blockDim.x) {
    float* p = s + static_cast<long long>(r) * ld + c;
    *p = (beta == 0.f) ? 0.f : (*p * beta);
  }
}

}  // namespace

int launch_syrk_fp32(float* state, long long ld_state, const float* x, long long ldx, int n, int d,
                     int has_bias, float in_scale, float alpha, float beta, cudaStream_t stream) {
  const int dp = d + (has_bias ? 1 : 0);
  if (dp <= 0) return -2;
  if (beta != 1.f) {
    dim3 g((dp + 255) / 256, dp), b(256);
    scale_matrix_kernel<<<g, b, 0, stream>>>(state, ld_state, dp, beta);
  note_launch();
  }
  if (n <= 0) return cudaGetLastError() == cudaSuccess ? 0 : -5;
  const int tiles = (dp + TS - 1) / TS;
  const int ntiles = tiles * (tiles + 1) / 2;
  // split the reduction so that the grid covers the chip about twice
  int ksplit = (2 * kNumSMsB200 + ntiles - 1) / ntiles;
  const int max_split = (n + 4 * KC - 1) / (4 * KC);
  if (ksplit > max_split) ksplit = max_split;
  if (ksplit < 1) ksplit = 1;
  int rows = (n + ksplit - 1) / ksplit;
  rows = (rows + KC - 1) / KC * KC;
  ksplit = (n + rows - 1) / rows;
  syrk_fp32_kernel<<<dim3(ntiles, ksplit), 256, 0, stream>>>(state, ld_state, x, ldx, n, d,
                                                            has_bias ? 1 : 0, in_scale, alpha, rows);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace bk
