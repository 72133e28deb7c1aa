// bk_blockdiag.cu — kernel-block-diagonal approximations of a dense Fisher / Hessian (SURVEY §8 row f2,
// second half).
//
// Reference (paths relative to /root/reference), sampling_free/utils.py:63-211
// (`generate_kernel_diag_15080 / _748 / _141 / generate_kernel_diag`):
//     H += tau * I                                   (in place: the caller's H is modified)
//     res = 0;  for (a, b) in coords: res[a:b, a:b] = H[a:b, a:b]
//     return res, torch.inverse(n * res)
// The coordinate lists are one square block per convolution kernel / output unit plus one per bias vector.
// `generate_kernel_diag` (the regression variant) advances its first loop by 1 instead of n_hid, so for
// n_hid > 1 its first blocks OVERLAP and their union is a band, not a block diagonal; the restatement keeps
// that: the mask is the union of the squares, and the inverse is taken per CONNECTED component of the union.
//
//   band_mask_kernel       one pass: out[i][j] = H[i][j] (+ tau on the diagonal) where some block contains both
//                          i and j, else 0.  For interval blocks that is row_lo[i] <= j < row_hi[i] with
//                          row_lo / row_hi the extreme bounds of the blocks containing i (host-built, P ints).
//                          Reads only the band, writes every element once: HBM-bound, 4 P^2 bytes written.
//   block_inverse_kernel   one CTA per connected component (<= 160 rows): Gauss-Jordan with partial pivoting in
//                          shared memory, fp64 (torch.inverse is an LU with partial pivoting; a masked band of an
//                          SPD matrix need not be SPD, so no Cholesky here).  Components are a few hundred KFLOP
//                          each: latency-bound, the fp64 rate is irrelevant.
#include "bk_common.cuh"
#include "bk_kernels.cuh"

namespace bk {

namespace {

__global__ void __launch_bounds__(256)
band_mask_kernel(float* __restrict__ H, long long ld, int P, float tau, int in_place,
                 const int* __restrict__ row_lo, const int* __restrict__ row_hi, float* __restrict__ out,
                 long long ldo) {
  const bool vec = (ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  for (int i = blockIdx.x; i < P; i += gridDim.x) {
    const int lo = row_lo[i], hi = row_hi[i];
    float* hrow = H + static_cast<long long>(i) * ld;
    float* orow = out + static_cast<long long>(i) * ldo;
    auto value = [&](int j) -> float {
      if (j < lo || j >= hi) return 0.f;
      float v = hrow[j];
      if (j == i) {
        v += tau;
        if (in_place) hrow[j] = v;
      }
      return v;
    };
    if (vec) {
      const int n4 = P >> 2;
      float4* o4 = reinterpret_cast<float4*>(orow);
      for (int j4 = threadIdx.x; j4 < n4; j4 += blockDim.x) {
        const int j = 4 * j4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j + 3 >= lo && j < hi) v = make_float4(value(j), value(j + 1), value(j + 2), value(j + 3));
        __stcs(o4 + j4, v);
      }
      for (int j = (n4 << 2) + threadIdx.x; j < P; j += blockDim.x) orow[j] = value(j);
    } else {
      for (int j = threadIdx.x; j < P; j += blockDim.x) orow[j] = value(j);
    }
    // a diagonal element outside every block still receives tau in the caller's H (H += tau I is unconditional)
    if (in_place && threadIdx.x == 0 && (i < lo || i >= hi)) hrow[i] += tau;
  }
}

constexpr int kInvThreads = 256;

// shared memory: M [d][d + 1] doubles, col [d] doubles, piv [d] ints
__global__ void __launch_bounds__(kInvThreads)
block_inverse_kernel(const float* __restrict__ R, long long ld, const int* __restrict__ comp_begin,
                     const int* __restrict__ comp_end, double scale, float* __restrict__ out, long long ldo,
                     int* __restrict__ status) {
  extern __shared__ double smd[];
  const int a = comp_begin[blockIdx.x];
  const int d = comp_end[blockIdx.x] - a;
  const int ldm = d + 1;
  double* M = smd;
  double* col = M + static_cast<size_t>(d) * ldm;
  int* piv = reinterpret_cast<int*>(col + d);
  const int tid = threadIdx.x;
  __shared__ int s_p;
  __shared__ double s_pivot;
  __shared__ int bad;
  if (tid == 0) bad = 0;
  for (int e = tid; e < d * d; e += kInvThreads) {
    const int i = e / d, j = e - i * d;
    M[i * ldm + j] = scale * static_cast<double>(R[static_cast<long long>(a + i) * ld + a + j]);
  }
  __syncthreads();
  for (int k = 0; k < d; ++k) {
    // partial pivoting: largest |M[i][k]|, i >= k (first maximum, like LAPACK's idamax)
    if (tid < 32) {
      double best = -1.0;
      int bi = k;
      for (int i = k + tid; i < d; i += 32) {
        const double v = fabs(M[i * ldm + k]);
        if (v > best) {
          best = v;
          bi = i;
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double ob = __shfl_down_sync(0xffffffffu, best, off);
        const int oi = __shfl_down_sync(0xffffffffu, bi, off);
        if (ob > best || (ob == best && oi < bi)) {
          best = ob;
          bi = oi;
        }
      }
      if (tid == 0) {
        s_p = bi;
        piv[k] = bi;
        if (!(best > 0.0)) bad = k + 1;
      }
    }
    __syncthreads();
    const int p = s_p;
    if (p != k) {
      for (int j = tid; j < d; j += kInvThreads) {
        const double t = M[k * ldm + j];
        M[k * ldm + j] = M[p * ldm + j];
        M[p * ldm + j] = t;
      }
    }
    __syncthreads();
    if (tid == 0) {
      const double v = M[k * ldm + k];
      s_pivot = (v != 0.0) ? 1.0 / v : 0.0;
    }
    for (int i = tid; i < d; i += kInvThreads) col[i] = M[i * ldm + k];
    __syncthreads();
    // column k becomes e_k, row k is scaled by 1 / pivot
    const double pinv = s_pivot;
    for (int i = tid; i < d; i += kInvThreads) M[i * ldm + k] = (i == k) ? 1.0 : 0.0;
    __syncthreads();
    for (int j = tid; j < d; j += kInvThreads) M[k * ldm + j] *= pinv;
    __syncthreads();
    for (int e = tid; e < d * d; e += kInvThreads) {
      const int i = e / d, j = e - i * d;
      if (i != k) M[i * ldm + j] -= col[i] * M[k * ldm + j];
    }
    __syncthreads();
  }
  // undo the row interchanges as column interchanges, last first
  for (int k = d - 1; k >= 0; --k) {
    const int p = piv[k];
    if (p != k) {
      for (int i = tid; i < d; i += kInvThreads) {
        const double t = M[i * ldm + k];
        M[i * ldm + k] = M[i * ldm + p];
        M[i * ldm + p] = t;
      }
      __syncthreads();
    }
  }
  __syncthreads();
  for (int e = tid; e < d * d; e += kInvThreads) {
    const int i = e / d, j = e - i * d;
    out[static_cast<long long>(a + i) * ldo + a + j] = static_cast<float>(M[i * ldm + j]);
  }
  if (tid == 0 && bad != 0 && status != nullptr) atomicCAS(status, 0, blockIdx.x * 65536 + bad);
}

size_t inverse_smem(int d) {
  return static_cast<size_t>(d) * (d + 1) * sizeof(double) + static_cast<size_t>(d) * sizeof(double) +
         static_cast<size_t>(d) * sizeof(int);
}

}  // namespace

int launch_band_mask(float* H, long long ld, int P, float tau, int in_place, const int* row_lo,
                     const int* row_hi, float* out, long long ldo, cudaStream_t stream) {
  if (P <= 0) return 0;
  if (H == nullptr || out == nullptr || row_lo == nullptr || row_hi == nullptr || ld < P || ldo < P || H == out)
    return -2;
  const int grid = P < 148 * 8 ? P : 148 * 8;
  band_mask_kernel<<<grid, 256, 0, stream>>>(H, ld, P, tau, in_place, row_lo, row_hi, out, ldo);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_block_inverse(const float* R, long long ld, int P, const int* comp_begin, const int* comp_end,
                         int ncomp, int max_dim, double scale, float* out, long long ldo, int zero_fill,
                         int* status, cudaStream_t stream) {
  if (ncomp <= 0 || P <= 0) return 0;
  if (R == nullptr || out == nullptr || comp_begin == nullptr || comp_end == nullptr || ld < P || ldo < P ||
      max_dim <= 0 || max_dim > kBlockInvMaxDim)
    return -2;
  static DeviceOnce attr_once;
  if (!attr_once([] {
        return cudaFuncSetAttribute(block_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(inverse_smem(kBlockInvMaxDim))) == cudaSuccess;
      }))
    return -5;
  if (zero_fill &&
      cudaMemset2DAsync(out, static_cast<size_t>(ldo) * sizeof(float), 0, static_cast<size_t>(P) * sizeof(float),
                        static_cast<size_t>(P), stream) != cudaSuccess)
    return -5;
  if (status != nullptr && cudaMemsetAsync(status, 0, sizeof(int), stream) != cudaSuccess) return -5;
  block_inverse_kernel<<<ncomp, kInvThreads, inverse_smem(max_dim), stream>>>(R, ld, comp_begin, comp_end, scale,
                                                                             out, ldo, status);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace bk
