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

// Fast path (ldx % 4 == 0, 16 B aligned rows): 64 samples x 64 features per CTA, float4 loads,
// bf16x2 stores (one full 128 B line per warp instruction), optional per-feature column sums
// (sum_n scale * x[n, j], exact fp32) for the bias row of the first Kronecker factor: the caller then
// runs the SYRK on the d x d block only and fills row/column d from these sums
// (models/curvatures.py:346-349 without materialising the row of ones).
// Shared-memory layout of the 64 x 64 tile: sample PAIRS are stored together,
//   addr(s, f) = (s >> 1) * 129 + (s & 1) * 64 + f          (bank = (s >> 1) + f  mod 32),
// which makes both phases conflict-free: the load phase writes, per warp instruction, eight 4-float
// groups of the rows {s0, s0+2, s0+4, s0+6}; the transposed phase reads, for one feature f, the even
// (then the odd) samples 2*lane (+1).
struct Tile64Regs {
  float4 v[4];
};

// loads of one 64 x 64 tile (rows r0.., columns c0..): four independent 16-byte loads per thread
__device__ __forceinline__ void tile64_load(const float* __restrict__ X, long long ldx, int rows, int cols, int r0,
                                            int c0, Tile64Regs& t) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // unit u = pass * 8 + warp in [0, 32): (half of the 64 features, group of 8 rows, row parity)
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int u = pass * 8 + warp;
    const int half = u & 1, parity = (u >> 1) & 1, grp = u >> 2;  // grp in [0, 8)
    const int s = grp * 8 + parity + 2 * (lane >> 3);             // rows s0, s0+2, s0+4, s0+6
    const int f = half * 32 + 4 * (lane & 7);
    const int r = r0 + s, c = c0 + f;
    t.v[pass] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows) {
      const float* src = X + static_cast<long long>(r) * ldx + c;
      if (c + 3 < cols) {
        t.v[pass] = __ldcs(reinterpret_cast<const float4*>(src));
      } else {
        if (c < cols) t.v[pass].x = src[0];
        if (c + 1 < cols) t.v[pass].y = src[1];
        if (c + 2 < cols) t.v[pass].z = src[2];
      }
    }
  }
}

// registers -> shared tile -> transposed, split, stored; ends with the tile free again (two barriers)
__device__ __forceinline__ void tile64_emit(const Tile64Regs& t, float* __restrict__ tile, int rows, int cols, int r0,
                                            int c0, float scale, __nv_bfloat16* __restrict__ Thi,
                                            __nv_bfloat16* __restrict__ Tlo, long long ldt,
                                            float* __restrict__ colsum, bool persistent) {
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int u = pass * 8 + warp;
    const int half = u & 1, parity = (u >> 1) & 1, grp = u >> 2;
    const int srow = grp * 8 + parity + 2 * (lane >> 3);
    const int fcol = half * 32 + 4 * (lane & 7);
    float* d = tile + (srow >> 1) * 129 + (srow & 1) * 64 + fcol;
    d[0] = t.v[pass].x * scale;
    d[1] = t.v[pass].y * scale;
    d[2] = t.v[pass].z * scale;
    d[3] = t.v[pass].w * scale;
  }
  __syncthreads();
  const int oc = r0 + 2 * lane;  // sample pair written by this lane
#pragma unroll
  for (int pass = 0; pass < 8; ++pass) {
    const int f = warp + 8 * pass;
    const int orow = c0 + f;
    if (orow >= cols) continue;
    const float a = tile[lane * 129 + f], b = tile[lane * 129 + 64 + f];
    __nv_bfloat16 ah, al, bh, bl;
    split_bf16(a, ah, al);
    split_bf16(b, bh, bl);
    __nv_bfloat16* dh = Thi + static_cast<long long>(orow) * ldt + oc;
    if (oc + 1 < rows) {
      *reinterpret_cast<__nv_bfloat162*>(dh) = __halves2bfloat162(ah, bh);
      if (Tlo != nullptr)
        *reinterpret_cast<__nv_bfloat162*>(Tlo + static_cast<long long>(orow) * ldt + oc) =
            __halves2bfloat162(al, bl);
    } else if (oc < rows) {
      dh[0] = ah;
      if (Tlo != nullptr) Tlo[static_cast<long long>(orow) * ldt + oc] = al;
    }
  }
  if (colsum != nullptr && tid < 64 && c0 + tid < cols) {
    float sacc = 0.f;
#pragma unroll 8
    for (int k = 0; k < 32; ++k)  // rows beyond `rows` hold zeros
      sacc += tile[k * 129 + tid] + tile[k * 129 + 64 + tid];
    atomicAdd(&colsum[c0 + tid], sacc);
  }
  if (persistent) __syncthreads();
}

__global__ void __launch_bounds__(256)
transpose_split64_kernel(const float* __restrict__ X, long long ldx, int rows, int cols, float scale,
                         __nv_bfloat16* __restrict__ Thi, __nv_bfloat16* __restrict__ Tlo,
                         long long ldt, float* __restrict__ colsum) {
  __shared__ float tile[32 * 129];
  const int c0 = blockIdx.x * 64;
  const int r0 = blockIdx.y * 64;
  Tile64Regs t;
  tile64_load(X, ldx, rows, cols, r0, c0, t);
  tile64_emit(t, tile, rows, cols, r0, c0, scale, Thi, Tlo, ldt, colsum, false);
}

// PERSISTENT variant for the staging passes that run UNDER a tensor-core SYRK (bk_syrk_accum_grouped, side
// stream): the SYRK's persistent CTAs own every SM and leave room for exactly one more CTA of this size, so a
// conventional grid of 4096 tile-CTAs trickles through one slot per SM, each exposing a full load latency before
// its first store (measured: 62 us of staging hid only 18 us under a 153 us SYRK).  Here one CTA per SM walks the
// tiles and keeps the NEXT tile's loads in flight while the current one is transposed and stored.
__global__ void __launch_bounds__(256, 1)
transpose_split64_persistent_kernel(const float* __restrict__ X, long long ldx, int rows, int cols, float scale,
                                    __nv_bfloat16* __restrict__ Thi, __nv_bfloat16* __restrict__ Tlo,
                                    long long ldt, float* __restrict__ colsum) {
  __shared__ float tile[32 * 129];
  const int tx = (cols + 63) / 64, ty = (rows + 63) / 64;
  const int ntiles = tx * ty;
  const int g = gridDim.x;
  // three tiles' loads (48 KB per SM) in flight while a fourth is transposed and stored: register ring A..D
  Tile64Regs A, B, C, D;
  auto ld = [&](int t, Tile64Regs& r) {
    if (t < ntiles) tile64_load(X, ldx, rows, cols, (t / tx) * 64, (t % tx) * 64, r);
  };
  auto em = [&](int t, const Tile64Regs& r) {
    if (t < ntiles) tile64_emit(r, tile, rows, cols, (t / tx) * 64, (t % tx) * 64, scale, Thi, Tlo, ldt, colsum, true);
  };
  int t = blockIdx.x;
  ld(t, A);
  ld(t + g, B);
  ld(t + 2 * g, C);
  for (; t < ntiles; t += 4 * g) {
    ld(t + 3 * g, D);
    em(t, A);
    ld(t + 4 * g, A);
    em(t + g, B);
    ld(t + 5 * g, B);
    em(t + 2 * g, C);
    ld(t + 6 * g, C);
    em(t + 3 * g, D);
  }
}

// colsum[j] += scale * sum_n X[n][j] for row-major bf16 activations (the bias row / column of the first
// Kronecker factor when the SYRK consumes the activations directly, models/curvatures.py:346-349).
// Block: 8 warps x 32 lanes, lane = column PAIR (one bf16x2 load), warps stride over the rows of the block's
// row chunk; fp32 partial sums, one atomic per column per block.
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ X, long long ldx, int rows, int cols, float scale,
                   int rows_per_block, float* __restrict__ colsum) {
  __shared__ float part[8][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + 2 * lane;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  float s0 = 0.f, s1 = 0.f;
  if (c + 1 < cols) {
    for (int r = r0 + warp; r < r1; r += 8) {
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(X + static_cast<long long>(r) * ldx + c);
      s0 += __bfloat162float(v.x);
      s1 += __bfloat162float(v.y);
    }
  } else if (c < cols) {
    for (int r = r0 + warp; r < r1; r += 8) s0 += __bfloat162float(X[static_cast<long long>(r) * ldx + c]);
  }
  part[warp][2 * lane] = s0;
  part[warp][2 * lane + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64 && blockIdx.x * 64 + threadIdx.x < cols) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += part[w][threadIdx.x];
    atomicAdd(&colsum[blockIdx.x * 64 + threadIdx.x], scale * t);
  }
}

__global__ void fill_ones_row_kernel(__nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                     int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    hi[i] = __float2bfloat16_rn(1.f);
    if (lo != nullptr) lo[i] = __float2bfloat16_rn(0.f);
  }
}

// state[d][j] = state[j][d] = beta*old + alpha*colsum[j];  state[d][d] = beta*old + alpha*n.
__global__ void bias_border_kernel(float* __restrict__ state, long long ld, int d,
                                   const float* __restrict__ colsum, float alpha, float beta,
                                   float n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < d) {
    float* row = state + static_cast<long long>(d) * ld + j;
    float* col = state + static_cast<long long>(j) * ld + d;
    const float v = (beta == 0.f ? 0.f : beta * *row) + alpha * colsum[j];
    *row = v;
    *col = v;
  } else if (j == d) {
    float* c = state + static_cast<long long>(d) * ld + d;
    *c = (beta == 0.f ? 0.f : beta * *c) + alpha * n;
  }
}

// fp32 [rows, cols] -> bf16 hi[/lo] same orientation.  kVec: 4 columns per thread (float4 load,
// 8-byte bf16x4 stores), 4 independent loads in flight per thread.
template <bool kVec>
__global__ void __launch_bounds__(256)
convert_split_kernel(const float* __restrict__ X, long long ldx, int rows, int cols, float scale,
                     int lower_only, __nv_bfloat16* __restrict__ Ohi,
                     __nv_bfloat16* __restrict__ Olo, long long ldo) {
  constexpr int kU = 4;
  for (int r = blockIdx.y; r < rows; r += gridDim.y) {
    const float* src = X + static_cast<long long>(r) * ldx;
    __nv_bfloat16* dh = Ohi + static_cast<long long>(r) * ldo;
    __nv_bfloat16* dl = Olo != nullptr ? Olo + static_cast<long long>(r) * ldo : nullptr;
    if (kVec) {
      const int n4 = cols >> 2;
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x * kU) {
        float4 v[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int ii = i + u * gridDim.x * blockDim.x;
          if (ii < n4) v[u] = __ldcs(reinterpret_cast<const float4*>(src) + ii);
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int ii = i + u * gridDim.x * blockDim.x;
          if (ii >= n4) continue;
          const int c = ii << 2;
          float e[4] = {v[u].x * scale, v[u].y * scale, v[u].z * scale, v[u].w * scale};
          __nv_bfloat16 h[4], l[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (lower_only && c + j > r) e[j] = 0.f;
            split_bf16(e[j], h[j], l[j]);
          }
          uint2 ph;
          ph.x = (static_cast<uint32_t>(__bfloat16_as_ushort(h[1])) << 16) | __bfloat16_as_ushort(h[0]);
          ph.y = (static_cast<uint32_t>(__bfloat16_as_ushort(h[3])) << 16) | __bfloat16_as_ushort(h[2]);
          *reinterpret_cast<uint2*>(dh + c) = ph;
          if (dl != nullptr) {
            uint2 pl;
            pl.x = (static_cast<uint32_t>(__bfloat16_as_ushort(l[1])) << 16) | __bfloat16_as_ushort(l[0]);
            pl.y = (static_cast<uint32_t>(__bfloat16_as_ushort(l[3])) << 16) | __bfloat16_as_ushort(l[2]);
            *reinterpret_cast<uint2*>(dl + c) = pl;
          }
        }
      }
      for (int c = (n4 << 2) + blockIdx.x * blockDim.x + threadIdx.x; c < cols;
           c += gridDim.x * blockDim.x) {
        float v = src[c] * scale;
        if (lower_only && c > r) v = 0.f;
        __nv_bfloat16 h, l;
        split_bf16(v, h, l);
        dh[c] = h;
        if (dl != nullptr) dl[c] = l;
      }
    } else {
      for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += gridDim.x * blockDim.x) {
        float v = src[c] * scale;
        if (lower_only && c > r) v = 0.f;
        __nv_bfloat16 h, l;
        split_bf16(v, h, l);
        dh[c] = h;
        if (dl != nullptr) dl[c] = l;
      }
    }
  }
}

// ---- three-way bf16 splits (hi + lo + lo2 = 24 mantissa bits) for the fp32-class bf16x6 products of
// the blocked Cholesky trailing updates.
__device__ __forceinline__ void split3_bf16(float x, __nv_bfloat16& h, __nv_bfloat16& m,
                                            __nv_bfloat16& l) {
  h = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(h);
  m = __float2bfloat16_rn(r1);
  l = __float2bfloat16_rn(r1 - __bfloat162float(m));
}

// fp32 [rows, cols] (pitch ldx, 16 B aligned rows, cols % 4 == 0) -> three bf16 [rows, ldo] parts.
__global__ void __launch_bounds__(256)
convert_split3_kernel(const float* __restrict__ X, long long ldx, int rows, int cols,
                      __nv_bfloat16* __restrict__ O0, __nv_bfloat16* __restrict__ O1,
                      __nv_bfloat16* __restrict__ O2, long long ldo) {
  const int n4 = cols >> 2;
  for (int r = blockIdx.y; r < rows; r += gridDim.y) {
    const float4* src = reinterpret_cast<const float4*>(X + static_cast<long long>(r) * ldx);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
      const float4 v = __ldg(src + i);
      const float e[4] = {v.x, v.y, v.z, v.w};
      __nv_bfloat16 h[4], m[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) split3_bf16(e[j], h[j], m[j], l[j]);
      const long long o = static_cast<long long>(r) * ldo + (i << 2);
      uint2 p;
      p.x = (static_cast<uint32_t>(__bfloat16_as_ushort(h[1])) << 16) | __bfloat16_as_ushort(h[0]);
      p.y = (static_cast<uint32_t>(__bfloat16_as_ushort(h[3])) << 16) | __bfloat16_as_ushort(h[2]);
      *reinterpret_cast<uint2*>(O0 + o) = p;
      p.x = (static_cast<uint32_t>(__bfloat16_as_ushort(m[1])) << 16) | __bfloat16_as_ushort(m[0]);
      p.y = (static_cast<uint32_t>(__bfloat16_as_ushort(m[3])) << 16) | __bfloat16_as_ushort(m[2]);
      *reinterpret_cast<uint2*>(O1 + o) = p;
      p.x = (static_cast<uint32_t>(__bfloat16_as_ushort(l[1])) << 16) | __bfloat16_as_ushort(l[0]);
      p.y = (static_cast<uint32_t>(__bfloat16_as_ushort(l[3])) << 16) | __bfloat16_as_ushort(l[2]);
      *reinterpret_cast<uint2*>(O2 + o) = p;
    }
  }
}

// fp32 [rows, cols] -> three bf16 [cols, ldt] parts (transposed).  32 x 32 tiles through shared memory.
__global__ void transpose_split3_kernel(const float* __restrict__ X, long long ldx, int rows, int cols,
                                        __nv_bfloat16* __restrict__ T0,
                                        __nv_bfloat16* __restrict__ T1,
                                        __nv_bfloat16* __restrict__ T2, long long ldt) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // (32, 8)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + ty + 8 * k, c = c0 + tx;
    tile[ty + 8 * k][tx] = (r < rows && c < cols) ? X[static_cast<long long>(r) * ldx + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int orow = c0 + ty + 8 * k, ocol = r0 + tx;
    if (orow < cols && ocol < rows) {
      __nv_bfloat16 h, m, l;
      split3_bf16(tile[tx][ty + 8 * k], h, m, l);
      const long long o = static_cast<long long>(orow) * ldt + ocol;
      T0[o] = h;
      T1[o] = m;
      T2[o] = l;
    }
  }
}

}  // namespace

int launch_convert_split3(const float* X, long long ldx, int rows, int cols, __nv_bfloat16* O0,
                          __nv_bfloat16* O1, __nv_bfloat16* O2, long long ldo, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return 0;
  if ((cols & 3) || (ldx & 3) || (ldo & 3) || (reinterpret_cast<uintptr_t>(X) & 15)) return -2;
  int bx = (cols / 4 + 255) / 256;
  if (bx > 16) bx = 16;
  dim3 grid(bx, rows < 65535 ? rows : 65535);
  convert_split3_kernel<<<grid, 256, 0, stream>>>(X, ldx, rows, cols, O0, O1, O2, ldo);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_transpose_split3(const float* X, long long ldx, int rows, int cols, __nv_bfloat16* T0,
                            __nv_bfloat16* T1, __nv_bfloat16* T2, long long ldt,
                            cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return 0;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_split3_kernel<<<grid, block, 0, stream>>>(X, ldx, rows, cols, T0, T1, T2, ldt);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_transpose_split(const float* X, long long ldx, int rows, int cols, float scale,
                           int ones_row, __nv_bfloat16* Thi, __nv_bfloat16* Tlo, long long ldt,
                           cudaStream_t stream, float* colsum, int persistent_ctas) {
  if (rows <= 0 || cols <= 0) return 0;
  const bool fast = (ldx % 4 == 0) && (reinterpret_cast<uintptr_t>(X) % 16 == 0) && (ldt % 2 == 0) &&
                    (reinterpret_cast<uintptr_t>(Thi) % 4 == 0) &&
                    (Tlo == nullptr || reinterpret_cast<uintptr_t>(Tlo) % 4 == 0);
  if (colsum != nullptr && !fast) return -2;  // callers only request sums on the aligned path
  if (fast) {
    dim3 grid((cols + 63) / 64, (rows + 63) / 64);
    if (persistent_ctas > 0)
      transpose_split64_persistent_kernel<<<persistent_ctas, 256, 0, stream>>>(X, ldx, rows, cols, scale, Thi, Tlo,
                                                                               ldt, colsum);
    else
      transpose_split64_kernel<<<grid, 256, 0, stream>>>(X, ldx, rows, cols, scale, Thi, Tlo, ldt, colsum);
    note_launch();
    if (ones_row) {  // generic operand with an explicit row of ones (bk_transpose_split ABI)
      fill_ones_row_kernel<<<(rows + 255) / 256, 256, 0, stream>>>(
          Thi + static_cast<long long>(cols) * ldt,
          Tlo != nullptr ? Tlo + static_cast<long long>(cols) * ldt : nullptr, rows);
      note_launch();
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -5;
  }
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_split_kernel<<<grid, block, 0, stream>>>(X, ldx, rows, cols, scale, ones_row, Thi, Tlo,
                                                     ldt);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_colsum_bf16(const __nv_bfloat16* X, long long ldx, int rows, int cols, float scale, float* colsum,
                       cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return 0;
  if ((ldx & 1) != 0 || (reinterpret_cast<uintptr_t>(X) & 3) != 0) return -2;
  int by = (rows + 127) / 128;
  if (by > 32) by = 32;
  const int rpb = (rows + by - 1) / by;
  colsum_bf16_kernel<<<dim3((cols + 63) / 64, by), 256, 0, stream>>>(X, ldx, rows, cols, scale, rpb, colsum);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_bias_border(float* state, long long ld, int d, const float* colsum, float alpha,
                       float beta, float n, cudaStream_t stream) {
  bias_border_kernel<<<(d + 1 + 255) / 256, 256, 0, stream>>>(state, ld, d, colsum, alpha, beta, n);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_convert_split(const float* X, long long ldx, int rows, int cols, float scale,
                         int lower_only, __nv_bfloat16* Ohi, __nv_bfloat16* Olo, long long ldo,
                         cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return 0;
  const bool vec = (ldx % 4 == 0) && (ldo % 4 == 0) && (reinterpret_cast<uintptr_t>(X) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(Ohi) % 8 == 0) &&
                   (Olo == nullptr || reinterpret_cast<uintptr_t>(Olo) % 8 == 0);
  // one block column covers 256 threads x 4 float4 x 4 columns = 4096 columns per sweep
  int bx = vec ? (cols + 4095) / 4096 : (cols + 255) / 256;
  if (bx > 64) bx = 64;
  if (bx < 1) bx = 1;
  dim3 grid(bx, rows < 65535 ? rows : 65535), block(256);
  if (vec)
    convert_split_kernel<true><<<grid, block, 0, stream>>>(X, ldx, rows, cols, scale, lower_only, Ohi,
                                                           Olo, ldo);
  else
    convert_split_kernel<false><<<grid, block, 0, stream>>>(X, ldx, rows, cols, scale, lower_only,
                                                            Ohi, Olo, ldo);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

namespace {

// Matrix element (r, c) of sample s is lane c % 4 of the Philox block with counter
// (c / 4, r, sample0 + s, stream_id) under key = seed: one thread produces 8 consecutive columns of one
// row (two Philox calls, four Box-Muller pairs) and stores them as one 16-byte bf16x8 (rows are 16 B
// aligned because ldz % 8 == 0).  The values do not depend on the launch geometry.
//
// PERSISTENT launch: one CTA per SM, 512 threads, grid-stride over the (row, column-group) items of each sample.
// The generator is ALU / MUFU work (~30 instructions per normal) and its consumers are tensor-bound persistent
// GEMMs that occupy every SM with one CTA of ~225 KB shared memory: a conventional grid of 200 000 small blocks
// never shares an SM with such a CTA (each resident block reserves 1 KB of shared memory, eight of them leave no
// room for the GEMM CTA, and blocks keep arriving), so generator and GEMM ran strictly one after the other even on
// different streams (measured: tools/gpu_philox_ab.py).  ONE generator CTA per SM (no shared memory of its own:
// 1 KB reserved) fits beside the GEMM CTA in either launch order, and the two use different pipes.
__global__ void __launch_bounds__(512, 1)
philox_normal_kernel(unsigned long long seed, uint32_t sample0, uint32_t stream_id, int rows,
                     int cols, int nsamples, float* __restrict__ Zf, long long ldf,
                     long long stridef, __nv_bfloat16* __restrict__ Zhi,
                     __nv_bfloat16* __restrict__ Zlo, long long ldz, long long stridez) {
  const int groups = (cols + 3) >> 2;  // column groups (Philox blocks) per row
  const uint32_t pairs = static_cast<uint32_t>((groups + 1) >> 1);  // two adjacent groups = 8 columns per thread
  const uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  // 16-byte stores of 8 bf16 need 16 B aligned rows
  const bool vec = (Zhi != nullptr) && ((ldz & 7) == 0) && ((stridez & 7) == 0) &&
                   ((reinterpret_cast<uintptr_t>(Zhi) & 15) == 0) &&
                   (Zlo == nullptr || (reinterpret_cast<uintptr_t>(Zlo) & 15) == 0);
  const uint32_t items = static_cast<uint32_t>(rows) * pairs;       // per sample (host checks < 2^32)
  const uint32_t first = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  // The (sample, row, column group) index of an item advances INCREMENTALLY by the grid stride: a few divisions per
  // thread, none per item, and small matrices with many samples (conv layers, S = 100) fill the grid because the
  // samples are part of the flattened index instead of a serial outer loop.
  const uint32_t ds = stride / items, rem = stride - ds * items;
  const uint32_t dr = rem / pairs, dt = rem - dr * pairs;
  const bool hi_only = vec && Zlo == nullptr && Zf == nullptr;      // single-pass bf16 operand: no lo parts at all
  {
    {
      uint32_t s_u = first / items;
      const uint32_t q0 = first - s_u * items;
      uint32_t r_u = q0 / pairs, t_u = q0 - r_u * pairs;
      for (; s_u < static_cast<uint32_t>(nsamples); s_u += ds, r_u += dr, t_u += dt) {
        if (t_u >= pairs) {
          t_u -= pairs;
          ++r_u;
        }
        if (r_u >= static_cast<uint32_t>(rows)) {
          r_u -= static_cast<uint32_t>(rows);
          if (++s_u >= static_cast<uint32_t>(nsamples)) break;
        }
        const int s = static_cast<int>(s_u);
        const int r = static_cast<int>(r_u);
        const int t = static_cast<int>(t_u);
        // two independent Philox blocks per thread: the 10-round dependency chains interleave
        uint32_t ca[4] = {static_cast<uint32_t>(2 * t), static_cast<uint32_t>(r), sample0 + s, stream_id};
        uint32_t cb[4] = {static_cast<uint32_t>(2 * t + 1), static_cast<uint32_t>(r), sample0 + s,
                          stream_id};
        philox4x32_10(ca, k0, k1);
        philox4x32_10(cb, k0, k1);
        float z[8];
        box_muller(ca[0], ca[1], z[0], z[1]);
        box_muller(ca[2], ca[3], z[2], z[3]);
        box_muller(cb[0], cb[1], z[4], z[5]);
        box_muller(cb[2], cb[3], z[6], z[7]);
        const int c0 = t << 3;
        const int nv = min(8, cols - c0);
        if (Zf != nullptr) {
          float* dst = Zf + s * stridef + static_cast<long long>(r) * ldf + c0;
          for (int j = 0; j < nv; ++j) dst[j] = z[j];
        }
        if (hi_only && c0 + 7 < ldz) {
          // padding columns [cols, ldz) are written as zeros
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j >= nv) z[j] = 0.f;
          const __nv_bfloat162 p0 = __floats2bfloat162_rn(z[0], z[1]), p1 = __floats2bfloat162_rn(z[2], z[3]);
          const __nv_bfloat162 p2 = __floats2bfloat162_rn(z[4], z[5]), p3 = __floats2bfloat162_rn(z[6], z[7]);
          uint4 ph;
          ph.x = *reinterpret_cast<const uint32_t*>(&p0);
          ph.y = *reinterpret_cast<const uint32_t*>(&p1);
          ph.z = *reinterpret_cast<const uint32_t*>(&p2);
          ph.w = *reinterpret_cast<const uint32_t*>(&p3);
          *reinterpret_cast<uint4*>(Zhi + s * stridez + static_cast<long long>(r) * ldz + c0) = ph;
          continue;
        }
        if (Zhi != nullptr) {
          __nv_bfloat16 h[8], l[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) split_bf16(z[j], h[j], l[j]);
          const long long o = s * stridez + static_cast<long long>(r) * ldz + c0;
          if (vec && c0 + 7 < ldz) {
            // the padding columns [cols, ldz) of the operand are written as zeros
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j >= nv) h[j] = l[j] = __float2bfloat16_rn(0.f);
            uint4 ph;
            ph.x = (static_cast<uint32_t>(__bfloat16_as_ushort(h[1])) << 16) | __bfloat16_as_ushort(h[0]);
            ph.y = (static_cast<uint32_t>(__bfloat16_as_ushort(h[3])) << 16) | __bfloat16_as_ushort(h[2]);
            ph.z = (static_cast<uint32_t>(__bfloat16_as_ushort(h[5])) << 16) | __bfloat16_as_ushort(h[4]);
            ph.w = (static_cast<uint32_t>(__bfloat16_as_ushort(h[7])) << 16) | __bfloat16_as_ushort(h[6]);
            *reinterpret_cast<uint4*>(Zhi + o) = ph;
            if (Zlo != nullptr) {
              uint4 pl;
              pl.x = (static_cast<uint32_t>(__bfloat16_as_ushort(l[1])) << 16) | __bfloat16_as_ushort(l[0]);
              pl.y = (static_cast<uint32_t>(__bfloat16_as_ushort(l[3])) << 16) | __bfloat16_as_ushort(l[2]);
              pl.z = (static_cast<uint32_t>(__bfloat16_as_ushort(l[5])) << 16) | __bfloat16_as_ushort(l[4]);
              pl.w = (static_cast<uint32_t>(__bfloat16_as_ushort(l[7])) << 16) | __bfloat16_as_ushort(l[6]);
              *reinterpret_cast<uint4*>(Zlo + o) = pl;
            }
          } else {
            for (int j = 0; j < nv; ++j) {
              Zhi[o + j] = h[j];
              if (Zlo != nullptr) Zlo[o + j] = l[j];
            }
          }
        }
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
  const long long pairs = ((cols + 3) / 4 + 1) / 2;
  if (pairs * rows >= (1ll << 32)) return -2;
  static DeviceOnce carve_once;
  // same shared-memory carveout as the GEMM CTAs it is meant to run beside (no reconfiguration between the two)
  carve_once([] {
    cudaFuncSetAttribute(philox_normal_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    return true;
  });
  const long long total = pairs * rows * nsamples;
  int grid = static_cast<int>((total + 511) / 512);
  if (grid > kNumSMsB200) grid = kNumSMsB200;
  philox_normal_kernel<<<grid, 512, 0, stream>>>(seed, sample0, stream_id, rows, cols, nsamples, Zf,
                                                 ldf, stridef, Zhi, Zlo, ldz, stridez);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace bk
