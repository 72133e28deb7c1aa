// bk_dense.cu — dense-Fisher helpers (BASELINE config 3).
//
//   dominance_kernel   one pass over H: sum |diag|, sum |all|, sum |per-kernel diagonal blocks| of
//                      H + tau*I — hessian/utils.py:4-23 (`calculateDominance`), which does the same
//                      with 109 slice reductions and one host sync each.  HBM-bound, coalesced rows,
//                      fp64 accumulation (P^2 = 2.3e8 addends at P = 15 080).
// The dense Fisher itself (H = sum_b g_b g_b^T / n, hessian/classification_ll_dense_kernel_diag.py:
// 85-89) is the factor SYRK on the stacked flat gradients, its damped inverse the batched Cholesky
// path, and J H^-1 J^T two calls of the contraction core: no extra kernels.
#include "bk_common.cuh"
#include "bk_kernels.cuh"

namespace bk {

namespace {

__global__ void __launch_bounds__(256)
dominance_kernel(const float* __restrict__ H, long long ld, int row0, int nrows, int P, float tau,
                 const int* __restrict__ block_begin, const int* __restrict__ block_end, int nblocks,
                 double* __restrict__ out) {
  __shared__ double red[3][8];
  double s_diag = 0.0, s_all = 0.0, s_blk = 0.0;
  const bool vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(H) & 15) == 0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // H points at global row `row0`; this launch covers the global rows [row0, row0 + nrows)
  for (int i = row0 + blockIdx.x; i < row0 + nrows; i += gridDim.x) {
    // diagonal block that contains row i (blocks are disjoint, ascending): binary search
    int lo = 0, hi = nblocks - 1, b0 = 0, b1 = 0;
    while (lo <= hi) {
      const int mid = (lo + hi) >> 1;
      const int bb = block_begin[mid], be = block_end[mid];
      if (i < bb) hi = mid - 1;
      else if (i >= be) lo = mid + 1;
      else {
        b0 = bb;
        b1 = be;
        break;
      }
    }
    const float* row = H + static_cast<long long>(i - row0) * ld;
    float a_all = 0.f, a_blk = 0.f;  // fp32 per-thread partials over <= P/256 addends, fp64 above
    auto take = [&](float x, int jj) {
      if (jj == i) {
        x += tau;
        s_diag += fabs(static_cast<double>(x));
      }
      const float av = fabsf(x);
      a_all += av;
      if (jj >= b0 && jj < b1) a_blk += av;
    };
    if (vec) {
      const float4* row4 = reinterpret_cast<const float4*>(row);
      const int n4 = P >> 2;
      for (int j = threadIdx.x; j < n4; j += 4 * blockDim.x) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int jj = j + u * blockDim.x;
          v[u] = (jj < n4) ? __ldcs(row4 + jj) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int jj = j + u * blockDim.x;
          if (jj >= n4) continue;
          take(v[u].x, 4 * jj);
          take(v[u].y, 4 * jj + 1);
          take(v[u].z, 4 * jj + 2);
          take(v[u].w, 4 * jj + 3);
        }
      }
      for (int jj = (n4 << 2) + threadIdx.x; jj < P; jj += blockDim.x) take(row[jj], jj);
    } else {
      for (int j = threadIdx.x; j < P; j += 4 * blockDim.x) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int jj = j + u * blockDim.x;
          v[u] = (jj < P) ? __ldcs(row + jj) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int jj = j + u * blockDim.x;
          if (jj < P) take(v[u], jj);
        }
      }
    }
    s_all += a_all;
    s_blk += a_blk;
  }
  s_diag = warp_sum(s_diag);
  s_all = warp_sum(s_all);
  s_blk = warp_sum(s_blk);
  if (lane == 0) {
    red[0][warp] = s_diag;
    red[1][warp] = s_all;
    red[2][warp] = s_blk;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    atomicAdd(&out[threadIdx.x], t);
  }
}

// out[(i*p + k), (j*q + l)] = a[i, j] * b[k, l]  (models/utilities.py:387-409, einsum "ab,cd->acbd")
__global__ void kron_kernel(const float* __restrict__ a, int m, int n, const float* __restrict__ b,
                            int p, int q, float* __restrict__ out) {
  const long long cols = static_cast<long long>(n) * q;
  const long long total = static_cast<long long>(m) * p * cols;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = idx / cols, c = idx - r * cols;
    const int i = static_cast<int>(r / p), k = static_cast<int>(r - static_cast<long long>(i) * p);
    const int j = static_cast<int>(c / q), l = static_cast<int>(c - static_cast<long long>(j) * q);
    out[idx] = a[static_cast<long long>(i) * n + j] * b[static_cast<long long>(k) * q + l];
  }
}

// state[i][j] = beta*state[i][j] + alpha*g[i]*g[j]   (BlockDiagonal.update: torch.ger(grads, grads) *
// batch_size accumulated with +=, models/curvatures.py:228-232).  One pass, float4 rows when aligned.
__global__ void __launch_bounds__(256)
ger_accum_kernel(float* __restrict__ state, long long ld, const float* __restrict__ g, int P,
                 float alpha, float beta) {
  for (int i = blockIdx.x; i < P; i += gridDim.x) {
    const float gi = alpha * g[i];
    float* row = state + static_cast<long long>(i) * ld;
    for (int j = threadIdx.x; j < P; j += 4 * blockDim.x) {
      float old[4], gj[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int jj = j + u * blockDim.x;
        old[u] = (jj < P && beta != 0.f) ? row[jj] : 0.f;
        gj[u] = (jj < P) ? __ldg(g + jj) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int jj = j + u * blockDim.x;
        if (jj < P) row[jj] = fmaf(gi, gj[u], beta * old[u]);
      }
    }
  }
}

}  // namespace

int launch_ger_accum(float* state, long long ld, const float* g, int P, float alpha, float beta,
                     cudaStream_t stream) {
  if (P <= 0) return 0;
  const int grid = P < kNumSMsB200 * 16 ? P : kNumSMsB200 * 16;
  ger_accum_kernel<<<grid, 256, 0, stream>>>(state, ld, g, P, alpha, beta);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_kron(const float* a, int m, int n, const float* b, int p, int q, float* out,
                cudaStream_t stream) {
  const long long total = static_cast<long long>(m) * n * p * q;
  if (total <= 0) return 0;
  long long blocks = (total + 255) / 256;
  if (blocks > kNumSMsB200 * 16) blocks = kNumSMsB200 * 16;
  kron_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(a, m, n, b, p, q, out);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_dominance(const float* H, long long ld, int row0, int nrows, int P, float tau, const int* block_begin,
                     const int* block_end, int nblocks, double* out3, cudaStream_t stream) {
  if (cudaMemsetAsync(out3, 0, 3 * sizeof(double), stream) != cudaSuccess) return -5;
  if (P <= 0 || nrows <= 0) return 0;
  int grid = nrows < kNumSMsB200 * 16 ? nrows : kNumSMsB200 * 16;
  dominance_kernel<<<grid, 256, 0, stream>>>(H, ld, row0, nrows, P, tau, block_begin, block_end, nblocks, out3);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace bk
