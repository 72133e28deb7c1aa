// bk_prep.cu — operand staging kernels (HBM-bound, coalesced both ways):
//   * transpose_split : fp32 [rows, cols] -> bf16 (hi[, lo]) [cols(+1), rows] K-major operand for the
//                       factor SYRK, with the reference's appended ones row for the bias
//                       (models/curvatures.py:345-348) and an input scale.
//   * convert_split   : fp32 [rows, cols] -> bf16 (hi[, lo]) same orientation (Cholesky factors,
//                       weights), optional lower-triangle mask.
//   * philox_normal   : counter-based N(0,1) generator (Philox4x32-10 + Box-Muller), one value per
//                       (seed, sample, layer, element) so results do not depend on the GPU count.
#include "bk_common.cuh"
#include "bk_kernels.cuh"

namespace bk {

namespace {

__global__ void transpose_split_kernel(const float* __restrict__ X, long long ldx, int rows,
                                       int cols, float scale, int ones_row,
                                       __nv_bfloat16* __restrict__ Thi,
                                       __nv_bfloat16* __restrict__ Tlo, long long ldt) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32;  // input column block -> output row block
  const int r0 = blockIdx.y * 32;  // input row block    -> output column block
  const int tx = threadIdx.x, ty = threadIdx.y;  // (32, 8)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + ty + 8 * k;
    const int c = c0 + tx;
    float v = 0.f;
    if (r < rows && c < cols) v = X[static_cast<long long>(r) * ldx + c] * scale;
    tile[ty + 8 * k][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int orow = c0 + ty + 8 * k;  // feature index
    const int ocol = r0 + tx;          // sample index
    if (orow < cols && ocol < rows) {
      const float v = tile[tx][ty + 8 * k];
      __nv_bfloat16 h, l;
      split_bf16(v, h, l);
      Thi[static_cast<long long>(orow) * ldt + ocol] = h;
      if (Tlo != nullptr) Tlo[static_cast<long long>(orow) * ldt + ocol] = l;
    }
  }
  // bias row of ones (one block column writes it)
  if (ones_row && blockIdx.x == 0) {
    const int ocol = r0 + ty * 32 + tx;
    if (ty == 0 && ocol < rows) {
      Thi[static_cast<long long>(cols) * ldt + ocol] = __float2bfloat16_rn(1.f);
      if (Tlo != nullptr) Tlo[static_cast<long long>(cols) * ldt + ocol] = __float2bfloat16_rn(0.f);
    }
  }
}

__global__ void convert_split_kernel(const float* __restrict__ X, long long ldx, int rows, int cols,
                                     float scale, int lower_only, __nv_bfloat16* __restrict__ Ohi,
                                     __nv_bfloat16* __restrict__ Olo, long long ldo) {
  const int r = blockIdx.y;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += gridDim.x * blockDim.x) {
    float v = X[static_cast<long long>(r) * ldx + c] * scale;
    if (lower_only && c > r) v = 0.f;
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    Ohi[static_cast<long long>(r) * ldo + c] = h;
    if (Olo != nullptr) Olo[static_cast<long long>(r) * ldo + c] = l;
  }
}

}  // namespace

int launch_transpose_split(const float* X, long long ldx, int rows, int cols, float scale,
                           int ones_row, __nv_bfloat16* Thi, __nv_bfloat16* Tlo, long long ldt,
                           cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return 0;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_split_kernel<<<grid, block, 0, stream>>>(X, ldx, rows, cols, scale, ones_row, Thi, Tlo,
                                                     ldt);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_convert_split(const float* X, long long ldx, int rows, int cols, float scale,
                         int lower_only, __nv_bfloat16* Ohi, __nv_bfloat16* Olo, long long ldo,
                         cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return 0;
  int bx = (cols + 255) / 256;
  if (bx > 64) bx = 64;
  dim3 grid(bx, rows), block(256);
  convert_split_kernel<<<grid, block, 0, stream>>>(X, ldx, rows, cols, scale, lower_only, Ohi, Olo,
                                                   ldo);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

namespace {

// Element e (row-major index into the [rows, cols] output, int64) belongs to Philox counter
// (e / 4, sample, stream_id) under key = seed; lane e % 4 of the 4 normals that counter yields.
__global__ void philox_normal_kernel(unsigned long long seed, uint32_t sample0, uint32_t stream_id,
                                     int rows, int cols, int nsamples, float* __restrict__ Zf,
                                     long long ldf, long long stridef,
                                     __nv_bfloat16* __restrict__ Zhi,
                                     __nv_bfloat16* __restrict__ Zlo, long long ldz,
                                     long long stridez) {
  const long long per = static_cast<long long>(rows) * cols;
  const long long groups = (per + 3) / 4;
  const long long total = groups * nsamples;
  for (long long g = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; g < total;
       g += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int s = static_cast<int>(g / groups);
    const long long gi = g - static_cast<long long>(s) * groups;
    uint32_t c[4] = {static_cast<uint32_t>(gi), static_cast<uint32_t>(gi >> 32), sample0 + s,
                     stream_id};
    philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    float z[4];
    box_muller(c[0], c[1], z[0], z[1]);
    box_muller(c[2], c[3], z[2], z[3]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long e = gi * 4 + j;
      if (e >= per) break;
      const int r = static_cast<int>(e / cols);
      const int cc = static_cast<int>(e - static_cast<long long>(r) * cols);
      if (Zf != nullptr) Zf[s * stridef + static_cast<long long>(r) * ldf + cc] = z[j];
      if (Zhi != nullptr) {
        __nv_bfloat16 h, l;
        split_bf16(z[j], h, l);
        const long long o = s * stridez + static_cast<long long>(r) * ldz + cc;
        Zhi[o] = h;
        if (Zlo != nullptr) Zlo[o] = l;
      }
    }
  }
}

}  // namespace

int launch_philox_normal(unsigned long long seed, unsigned sample0, unsigned stream_id, int rows,
                         int cols, int nsamples, float* Zf, long long ldf, long long stridef,
                         __nv_bfloat16* Zhi, __nv_bfloat16* Zlo, long long ldz, long long stridez,
                         cudaStream_t stream) {
  if (rows <= 0 || cols <= 0 || nsamples <= 0) return 0;
  const long long total = ((static_cast<long long>(rows) * cols + 3) / 4) * nsamples;
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(kNumSMsB200) * 16;
  if (blocks > cap) blocks = cap;
  philox_normal_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(
      seed, sample0, stream_id, rows, cols, nsamples, Zf, ldf, stridef, Zhi, Zlo, ldz, stridez);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace bk
