// bk_tri.cu — packing of symmetric factors for the multi-GPU exchange ("reduce of packed lower
// triangles").  A and G are symmetric by construction (models/curvatures.py:349,356 compute X X^T), so the
// all-reduce of the accumulated state needs d(d+1)/2 values per factor, not d^2: the NVLink / NVSwitch
// payload of the cfg5 exchange drops from 470 MB to 235 MB.
//   tri_pack    factors [d, ld] fp32 -> one flat buffer, factor after factor, row i = i + 1 values at i(i+1)/2
//   tri_unpack  flat buffer -> full [d, ld] matrices, scaled (1 / world size): mirrored (symmetric factors) or
//               with a zero upper triangle (the Cholesky factors the owners send back)
// HBM-bound; 32 x 32 tiles, the mirrored half is written through a shared-memory transpose so that both
// the direct and the mirrored stores are coalesced.  Up to 16 factors per launch (table in the kernel
// parameters, grid.z = factor).
#include "bk_common.cuh"
#include "bk_kernels.cuh"

namespace bk {

namespace {

constexpr int kMaxTri = 16;

struct TriTable {
  float* mat[kMaxTri];
  long long ld[kMaxTri];
  long long off[kMaxTri];  // first packed element of the factor
  int d[kMaxTri];
};

__global__ void __launch_bounds__(256)
tri_pack_kernel(const __grid_constant__ TriTable t, float* __restrict__ packed) {
  const int f = blockIdx.z;
  const int d = t.d[f];
  const int ti = blockIdx.y, tj = blockIdx.x;
  if (tj > ti || ti * 32 >= d) return;
  const float* m = t.mat[f];
  const long long ld = t.ld[f];
  float* out = packed + t.off[f];
  const int j = tj * 32 + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = ti * 32 + threadIdx.y + 8 * k;
    if (i < d && j <= i) out[static_cast<long long>(i) * (i + 1) / 2 + j] = m[static_cast<long long>(i) * ld + j];
  }
}

__global__ void __launch_bounds__(256)
tri_unpack_kernel(const __grid_constant__ TriTable t, const float* __restrict__ packed, float scale, int mirror) {
  __shared__ float tile[32][33];
  const int f = blockIdx.z;
  const int d = t.d[f];
  const int ti = blockIdx.y, tj = blockIdx.x;
  if (tj > ti || ti * 32 >= d) return;
  float* m = t.mat[f];
  const long long ld = t.ld[f];
  const float* in = packed + t.off[f];
  const int j = tj * 32 + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = threadIdx.y + 8 * k;
    const int i = ti * 32 + r;
    float v = 0.f;
    if (i < d && j <= i) {
      v = scale * in[static_cast<long long>(i) * (i + 1) / 2 + j];
      m[static_cast<long long>(i) * ld + j] = v;
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  // mirrored tile: element (j, i) = element (i, j), strictly above the diagonal
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = threadIdx.y + 8 * k;      // row inside the mirrored tile = column index j
    const int jj = tj * 32 + r;
    const int ii = ti * 32 + threadIdx.x;   // column inside the mirrored tile = row index i
    if (ii < d && jj < ii) m[static_cast<long long>(jj) * ld + ii] = mirror ? tile[threadIdx.x][r] : 0.f;
  }
}

// In place: lower triangle (diagonal included) *= scale, upper triangle = mirror of it.  The tensor-core factor
// update accumulates lower triangles only (bk_syrk_accum_grouped with BK_SYRK_LOWER_ONLY) and, in the
// running-average mode, in units of a lazily applied scalar; this pass produces the full symmetric factor
// the reference's callers read from `state` (models/curvatures.py:363), once per read instead of per update.
__global__ void __launch_bounds__(256)
sym_finalize_kernel(const __grid_constant__ TriTable t, float scale) {
  __shared__ float tile[32][33];
  const int f = blockIdx.z;
  const int d = t.d[f];
  const int ti = blockIdx.y, tj = blockIdx.x;
  if (tj > ti || ti * 32 >= d) return;
  float* m = t.mat[f];
  const long long ld = t.ld[f];
  const int j = tj * 32 + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = threadIdx.y + 8 * k;
    const int i = ti * 32 + r;
    float v = 0.f;
    if (i < d && j <= i) {
      v = scale * m[static_cast<long long>(i) * ld + j];
      if (scale != 1.f) m[static_cast<long long>(i) * ld + j] = v;
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = threadIdx.y + 8 * k;
    const int jj = tj * 32 + r;
    const int ii = ti * 32 + threadIdx.x;
    if (ii < d && jj < ii) m[static_cast<long long>(jj) * ld + ii] = tile[threadIdx.x][r];
  }
}

int run(bool pack, float* const* mats, const long long* lds, const int* dims, int count, float* packed,
        float scale, int mirror, cudaStream_t stream) {
  long long off = 0;
  for (int base = 0; base < count; base += kMaxTri) {
    TriTable t{};
    const int n = count - base < kMaxTri ? count - base : kMaxTri;
    int dmax = 0;
    for (int k = 0; k < n; ++k) {
      const int d = dims[base + k];
      if (d <= 0 || mats[base + k] == nullptr || lds[base + k] < d) return -2;
      t.mat[k] = mats[base + k];
      t.ld[k] = lds[base + k];
      t.d[k] = d;
      t.off[k] = off;
      off += static_cast<long long>(d) * (d + 1) / 2;
      if (d > dmax) dmax = d;
    }
    const int tiles = (dmax + 31) / 32;
    const dim3 grid(tiles, tiles, n), block(32, 8);
    if (mirror == 2) sym_finalize_kernel<<<grid, block, 0, stream>>>(t, scale);
    else if (pack) tri_pack_kernel<<<grid, block, 0, stream>>>(t, packed);
    else tri_unpack_kernel<<<grid, block, 0, stream>>>(t, packed, scale, mirror);
    note_launch();
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace

int launch_tri_pack(const float* const* mats, const long long* lds, const int* dims, int count, float* packed,
                    cudaStream_t stream) {
  return run(true, const_cast<float* const*>(mats), lds, dims, count, packed, 1.f, 1, stream);
}

int launch_sym_finalize(float* const* mats, const long long* lds, const int* dims, int count, float scale,
                        cudaStream_t stream) {
  return run(false, mats, lds, dims, count, nullptr, scale, 2, stream);
}

int launch_tri_unpack(float* const* mats, const long long* lds, const int* dims, int count, const float* packed,
                      float scale, int mirror, cudaStream_t stream) {
  return run(false, mats, lds, dims, count, const_cast<float*>(packed), scale, mirror, stream);
}

}  // namespace bk
