// bk_eigh.cu — batched symmetric eigendecomposition of Kronecker factors by one-sided (Hestenes)
// Jacobi with a round-robin parallel ordering and warp-shuffle reductions.
//
// Reference (models/utilities.py:144-159 get_eigenvectors, :120-141 get_eigenvalues): torch.symeig of
// F + F^T (eigenvectors) / of F (eigenvalues), ascending order.  symeig no longer exists in torch;
// the oracle restates it with torch.linalg.eigh (oracle/kfac_oracle.py factor_eigenvectors).
//
// Algorithm.  S = sym_scale * (F + F^T) is symmetric.  Keep two d x d row-major workspaces whose ROWS
// are the vectors: U (initially S, rows u_i = S e_i) and V (initially I).  A rotation of the pair
// (p, q) orthogonalises u_p and u_q and is applied to the rows of both, so U = V S always holds.
// At convergence the u_i are mutually orthogonal, hence u_i = lambda_i v_i: the v_i are the
// eigenvectors and lambda_i = <u_i, v_i> (sign included: indefinite inputs are fine).
// One round = floor(d/2) disjoint pairs processed in parallel; d-1 (d even) or d (d odd) rounds = one
// sweep; quadratic convergence, typically 6-10 sweeps in fp32.
//
//   d <= kSmallMax : ONE CTA per factor, U and V live in shared memory for the whole solve, one warp
//                    per pair, dot products by warp shuffles; all factors of a batch in one launch.
//   larger d       : one launch per round for all factors; one CTA per pair, rows streamed from L2
//                    (U and V of a 1025-wide factor are 8.4 MB: L2 resident), block-level reductions.
//                    HBM/L2-bound rank-2 updates — deliberately not reshaped into GEMMs.
// Output: eigenvalues ascending + eigenvectors as COLUMNS of a row-major [d, d] matrix (the layout of
// torch.symeig / linalg.eigh).
#include "bk_common.cuh"
#include "bk_kernels.cuh"

namespace bk {

namespace {

constexpr int kSmallMax = 164;  // 2 * d * (d + 1) * 4 B <= 216 KB of shared memory
// A pair is converged when |<u_p,u_q>| <= tol * |u_p| |u_q| with tol = max(2e-7, 1e-7 sqrt(d)) (the
// rounding level of a length-d fp32 dot product), or when both vectors are numerically zero:
// |u_p| |u_q| <= kNullRel * |S|_F^2 (null space of a rank-deficient factor: any orthonormal basis is
// an eigenbasis there, and V stays orthogonal by construction).
constexpr float kNullRel = 1e-10f;

struct EighProb {
  const float* F;   // [d, ldf] input factor
  long long ldf;
  float* U;         // [d, d] workspace (rows = S-images of the eigenvector estimates)
  float* V;         // [d, d] workspace (rows = eigenvector estimates)
  float* evals;     // [d] out, ascending
  float* evecs;     // [d, d] out (nullable): eigenvectors as columns
  float* lam;       // [d] workspace: unsorted Rayleigh quotients
  float* scale2;    // device scalar: |S|_F^2 (accumulated by the init step)
  int d;
  float sym_scale;  // S = sym_scale * (F + F^T)
  float tol;
};

// Round-robin tournament: n players (n even, player >= d is a bye), round r in [0, n-1).
__device__ __forceinline__ void rr_pair(int n, int r, int k, int& p, int& q) {
  if (k == 0) {
    p = n - 1;
    q = r;
  } else {
    p = (r + k) % (n - 1);
    q = (r - k + (n - 1)) % (n - 1);
  }
  if (p > q) {
    const int t = p;
    p = q;
    q = t;
  }
}

__device__ __forceinline__ bool rotation(float alpha, float beta, float gamma, float tol,
                                         float null2, float& c, float& s) {
  // returns false if the pair is already orthogonal to working precision
  const float ab = sqrtf(alpha * beta);
  if (fabsf(gamma) <= tol * ab || ab <= null2 || gamma == 0.f) return false;
  const float zeta = (beta - alpha) / (2.f * gamma);
  const float t = copysignf(1.f, zeta) / (fabsf(zeta) + sqrtf(1.f + zeta * zeta));
  c = 1.f / sqrtf(1.f + t * t);  // correctly rounded: keeps c^2 + s^2 = 1 unbiased over many rotations
  s = c * t;
  return true;
}

// ------------------------------------------------------------------------------ shared-memory path
__global__ void __launch_bounds__(512)
eigh_small_kernel(const EighProb* __restrict__ tab, int max_sweeps) {
  extern __shared__ float sm[];
  const EighProb P = tab[blockIdx.x];
  const int d = P.d;
  if (d <= 0 || d > kSmallMax) return;
  const int ld = d + 1;
  float* U = sm;
  float* V = sm + d * ld;
  __shared__ int s_rot;
  __shared__ float s_fro[16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  float fro = 0.f;
  for (int idx = tid; idx < d * d; idx += blockDim.x) {
    const int i = idx / d, j = idx - i * d;
    const float v = P.sym_scale * (P.F[static_cast<long long>(i) * P.ldf + j] +
                                   P.F[static_cast<long long>(j) * P.ldf + i]);
    U[i * ld + j] = v;
    V[i * ld + j] = (i == j) ? 1.f : 0.f;
    fro = fmaf(v, v, fro);
  }
  fro = warp_sum(fro);
  if (lane == 0) s_fro[warp] = fro;
  __syncthreads();
  fro = 0.f;
  for (int w = 0; w < nwarps; ++w) fro += s_fro[w];
  const float null2 = kNullRel * fro;
  const int n = (d + 1) & ~1;
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    if (tid == 0) s_rot = 0;
    __syncthreads();
    for (int r = 0; r < n - 1; ++r) {
      for (int k = warp; k < n / 2; k += nwarps) {
        int p, q;
        rr_pair(n, r, k, p, q);
        if (q >= d) continue;  // bye
        float* up = U + p * ld;
        float* uq = U + q * ld;
        float a = 0.f, b = 0.f, g = 0.f;
        for (int j = lane; j < d; j += 32) {
          const float x = up[j], y = uq[j];
          a = fmaf(x, x, a);
          b = fmaf(y, y, b);
          g = fmaf(x, y, g);
        }
        a = warp_sum(a);
        b = warp_sum(b);
        g = warp_sum(g);
        float c, s;
        if (!rotation(a, b, g, P.tol, null2, c, s)) continue;  // warp-uniform
        if (lane == 0) s_rot = 1;
        float* vp = V + p * ld;
        float* vq = V + q * ld;
        for (int j = lane; j < d; j += 32) {
          const float x = up[j], y = uq[j];
          up[j] = c * x - s * y;
          uq[j] = s * x + c * y;
          const float vx = vp[j], vy = vq[j];
          vp[j] = c * vx - s * vy;
          vq[j] = s * vx + c * vy;
        }
      }
      __syncthreads();
    }
    const int any = s_rot;
    __syncthreads();
    if (!any) break;
  }
  // Rayleigh quotients, then ranks (ascending, ties by index) and the permuted, transposed write.
  for (int i = warp; i < d; i += nwarps) {
    float acc = 0.f;
    for (int j = lane; j < d; j += 32) acc = fmaf(U[i * ld + j], V[i * ld + j], acc);
    acc = warp_sum(acc);
    if (lane == 0) P.lam[i] = acc;
  }
  __syncthreads();
  for (int i = tid; i < d; i += blockDim.x) {
    const float li = P.lam[i];
    int rank = 0;
    for (int j = 0; j < d; ++j) {
      const float lj = P.lam[j];
      rank += (lj < li) || (lj == li && j < i);
    }
    P.evals[rank] = li;
    U[i * ld + d] = __int_as_float(rank);  // the padding column carries the rank
  }
  __syncthreads();
  if (P.evecs != nullptr) {
    for (int idx = tid; idx < d * d; idx += blockDim.x) {
      const int j = idx / d, i = idx - j * d;  // consecutive threads -> consecutive source vectors i
      const int rank = __float_as_int(U[i * ld + d]);
      P.evecs[static_cast<long long>(j) * d + rank] = V[i * ld + j];
    }
  }
}

// ------------------------------------------------------------------------------ global-memory path
__global__ void eigh_init_kernel(const EighProb* __restrict__ tab) {
  const EighProb P = tab[blockIdx.z];
  const int d = P.d;
  if (d <= kSmallMax) return;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  if (i0 >= d || j0 >= d) return;
  __shared__ float tr[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;  // (32, 8)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int j = j0 + ty + 8 * k, i = i0 + tx;  // F[j][i], i fastest
    tr[ty + 8 * k][tx] = (i < d && j < d) ? P.F[static_cast<long long>(j) * P.ldf + i] : 0.f;
  }
  __syncthreads();
  float fro = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = i0 + ty + 8 * k, j = j0 + tx;
    if (i < d && j < d) {
      const float f = P.F[static_cast<long long>(i) * P.ldf + j];
      const float v = P.sym_scale * (f + tr[tx][ty + 8 * k]);
      P.U[static_cast<long long>(i) * d + j] = v;
      P.V[static_cast<long long>(i) * d + j] = (i == j) ? 1.f : 0.f;
      fro = fmaf(v, v, fro);
    }
  }
  fro = warp_sum(fro);
  if (tx == 0 && fro != 0.f) atomicAdd(P.scale2, fro);
}

// One CTA per pair of one round; rows cached in registers between the reduction and the rotation.
constexpr int kRoundThreads = 256;
constexpr int kRegElems = 20;  // 256 * 20 = 5120 >= 4097: rows up to 5120 wide stay in registers

__global__ void __launch_bounds__(kRoundThreads)
eigh_round_kernel(const EighProb* __restrict__ tab, int round, int* __restrict__ rotated) {
  const EighProb P = tab[blockIdx.y];
  const int d = P.d;
  if (d <= kSmallMax) return;
  const int n = (d + 1) & ~1;
  if (round >= n - 1 || static_cast<int>(blockIdx.x) >= n / 2) return;
  int p, q;
  rr_pair(n, round, blockIdx.x, p, q);
  if (q >= d) return;
  float* up = P.U + static_cast<long long>(p) * d;
  float* uq = P.U + static_cast<long long>(q) * d;
  const int tid = threadIdx.x;
  float x[kRegElems], y[kRegElems];
  float a = 0.f, b = 0.f, g = 0.f;
#pragma unroll
  for (int e = 0; e < kRegElems; ++e) {
    const int j = tid + e * kRoundThreads;
    x[e] = (j < d) ? up[j] : 0.f;
    y[e] = (j < d) ? uq[j] : 0.f;
    a = fmaf(x[e], x[e], a);
    b = fmaf(y[e], y[e], b);
    g = fmaf(x[e], y[e], g);
  }
  for (int j = tid + kRegElems * kRoundThreads; j < d; j += kRoundThreads) {  // very wide rows
    const float xx = up[j], yy = uq[j];
    a = fmaf(xx, xx, a);
    b = fmaf(yy, yy, b);
    g = fmaf(xx, yy, g);
  }
  __shared__ float red[3][kRoundThreads / 32];
  a = warp_sum(a);
  b = warp_sum(b);
  g = warp_sum(g);
  if ((tid & 31) == 0) {
    red[0][tid >> 5] = a;
    red[1][tid >> 5] = b;
    red[2][tid >> 5] = g;
  }
  __syncthreads();
  a = b = g = 0.f;
#pragma unroll
  for (int w = 0; w < kRoundThreads / 32; ++w) {
    a += red[0][w];
    b += red[1][w];
    g += red[2][w];
  }
  float c, s;
  if (!rotation(a, b, g, P.tol, kNullRel * *P.scale2, c, s)) return;  // block-uniform
  if (tid == 0) atomicOr(&rotated[blockIdx.y], 1);
  float* vp = P.V + static_cast<long long>(p) * d;
  float* vq = P.V + static_cast<long long>(q) * d;
#pragma unroll
  for (int e = 0; e < kRegElems; ++e) {
    const int j = tid + e * kRoundThreads;
    if (j < d) {
      up[j] = c * x[e] - s * y[e];
      uq[j] = s * x[e] + c * y[e];
      const float vx = vp[j], vy = vq[j];
      vp[j] = c * vx - s * vy;
      vq[j] = s * vx + c * vy;
    }
  }
  for (int j = tid + kRegElems * kRoundThreads; j < d; j += kRoundThreads) {
    const float xx = up[j], yy = uq[j];
    up[j] = c * xx - s * yy;
    uq[j] = s * xx + c * yy;
    const float vx = vp[j], vy = vq[j];
    vp[j] = c * vx - s * vy;
    vq[j] = s * vx + c * vy;
  }
}

// lambda_i = <u_i, v_i>, one warp per vector.
__global__ void eigh_rayleigh_kernel(const EighProb* __restrict__ tab) {
  const EighProb P = tab[blockIdx.y];
  const int d = P.d;
  if (d <= kSmallMax) return;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= d) return;
  const int lane = threadIdx.x & 31;
  const float* u = P.U + static_cast<long long>(i) * d;
  const float* v = P.V + static_cast<long long>(i) * d;
  float acc = 0.f;
  for (int j = lane; j < d; j += 32) acc = fmaf(u[j], v[j], acc);
  acc = warp_sum(acc);
  if (lane == 0) P.lam[i] = acc;
}

// rank by counting + permuted, transposed write of the eigenvectors (tile transposed through smem).
__global__ void eigh_sort_kernel(const EighProb* __restrict__ tab) {
  const EighProb P = tab[blockIdx.y];
  const int d = P.d;
  if (d <= kSmallMax) return;
  const int i0 = blockIdx.x * 32;  // 32 source vectors per block
  if (i0 >= d) return;
  __shared__ int ranks[32];
  __shared__ float tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;  // (32, 8)
  if (ty == 0) {
    const int i = i0 + tx;
    int rank = -1;
    if (i < d) {
      const float li = P.lam[i];
      rank = 0;
      for (int j = 0; j < d; ++j) {
        const float lj = P.lam[j];
        rank += (lj < li) || (lj == li && j < i);
      }
      P.evals[rank] = li;
    }
    ranks[tx] = rank;
  }
  __syncthreads();
  if (P.evecs == nullptr) return;
  for (int j0 = 0; j0 < d; j0 += 32) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = i0 + ty + 8 * k, j = j0 + tx;
      tile[ty + 8 * k][tx] = (i < d && j < d) ? P.V[static_cast<long long>(i) * d + j] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = j0 + ty + 8 * k;  // component index = output row
      const int rank = ranks[tx];     // output column
      if (j < d && rank >= 0) P.evecs[static_cast<long long>(j) * d + rank] = tile[tx][ty + 8 * k];
    }
    __syncthreads();
  }
}

inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

// 0: wide factors by the tensor-core block Jacobi (bk_eigh_blocked.cu); 1: streamed element-wise Jacobi;
// 2 (bring-up): block Jacobi for every size
int g_eigh_mode = 0;
inline bool use_blocked(int d) { return g_eigh_mode == 2 || (g_eigh_mode == 0 && d > kSmallMax); }

}  // namespace

void set_eigh_mode(int mode) { g_eigh_mode = (mode == 1 || mode == 2) ? mode : 0; }

size_t eigh_workspace_bytes(const int* dims, int count) {
  size_t total = align256(sizeof(EighProb) * static_cast<size_t>(count)) + 2 * align256(4 * count);
  size_t blocked = 0;
  for (int i = 0; i < count; ++i) {
    const size_t d = static_cast<size_t>(dims[i]);
    total += align256(d * 4);  // lam
    if (use_blocked(dims[i])) {
      blocked += eigh_blocked_workspace_bytes(dims[i]);  // wide factors are solved concurrently
    } else if (dims[i] > kSmallMax) {
      total += 2 * align256(d * d * 4);
    }
  }
  return total + blocked;
}

// Returns 0, or the 1-based index of the first factor that did not converge in max_sweeps, or < 0.
int eigh_batched(const float* const* factors, const long long* ldf, float* const* evals,
                 float* const* evecs, const int* dims, int count, float sym_scale, int max_sweeps,
                 void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (count <= 0) return 0;
  if (workspace == nullptr || workspace_bytes < eigh_workspace_bytes(dims, count) ||
      (reinterpret_cast<uintptr_t>(workspace) & 255) != 0)
    return -6;
  if (max_sweeps <= 0) max_sweeps = 30;
  char* base = static_cast<char*>(workspace);
  EighProb* d_tab = reinterpret_cast<EighProb*>(base);
  size_t off = align256(sizeof(EighProb) * static_cast<size_t>(count));
  int* d_rot = reinterpret_cast<int*>(base + off);
  off += align256(4 * count);
  float* d_scale2 = reinterpret_cast<float*>(base + off);
  off += align256(4 * count);
  if (cudaMemsetAsync(d_scale2, 0, 4 * count, stream) != cudaSuccess) return -5;
  EighProb h_tab[64];
  if (count > 64) return -2;
  int max_small = 0, max_large = 0, n_large = 0, n_blocked = 0;
  for (int i = 0; i < count; ++i) {
    EighProb& P = h_tab[i];
    const size_t d = static_cast<size_t>(dims[i]);
    if (dims[i] <= 0 || factors[i] == nullptr || evals[i] == nullptr) return -2;
    P.F = factors[i];
    P.ldf = ldf[i];
    P.evals = evals[i];
    P.evecs = evecs != nullptr ? evecs[i] : nullptr;
    P.d = dims[i];
    P.sym_scale = sym_scale;
    P.scale2 = d_scale2 + i;
    {
      const float t = 1e-7f * sqrtf(static_cast<float>(dims[i]));
      P.tol = t > 2e-7f ? t : 2e-7f;
    }
    P.lam = reinterpret_cast<float*>(base + off);
    off += align256(d * 4);
    P.U = P.V = nullptr;
    if (use_blocked(dims[i])) {
      ++n_blocked;
      P.d = 0;  // the element-wise kernels skip this entry
    } else if (dims[i] > kSmallMax) {
      P.U = reinterpret_cast<float*>(base + off);
      off += align256(d * d * 4);
      P.V = reinterpret_cast<float*>(base + off);
      off += align256(d * d * 4);
      if (dims[i] > max_large) max_large = dims[i];
      ++n_large;
    } else if (dims[i] > max_small) {
      max_small = dims[i];
    }
  }
  if (cudaMemcpyAsync(d_tab, h_tab, sizeof(EighProb) * count, cudaMemcpyHostToDevice, stream) !=
      cudaSuccess)
    return -5;
  int status = 0;
  if (max_small > 0) {
    const size_t smem = 2ull * max_small * (max_small + 1) * 4;
    static DeviceOnce attr_once;
    if (!attr_once([] {
          return cudaFuncSetAttribute(eigh_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      2 * kSmallMax * (kSmallMax + 1) * 4) == cudaSuccess;
        }))
      return -5;
    eigh_small_kernel<<<count, 512, smem, stream>>>(d_tab, max_sweeps);
    note_launch();
  }
  if (n_blocked > 0) {
    // wide factors: tensor-core block Jacobi, all of them concurrently on their own streams
    const float* bf[64];
    long long bld[64];
    int bd[64], bidx[64], bstat[64];
    float btol[64];
    float* bw[64];
    float* bv[64];
    int nbk = 0;
    for (int i = 0; i < count; ++i) {
      if (!use_blocked(dims[i])) continue;
      bf[nbk] = factors[i];
      bld[nbk] = ldf[i];
      bd[nbk] = dims[i];
      btol[nbk] = h_tab[i].tol;
      bw[nbk] = evals[i];
      bv[nbk] = evecs != nullptr ? evecs[i] : nullptr;
      bidx[nbk] = i;
      ++nbk;
    }
    const int rc = eigh_blocked_batch(bf, bld, bd, nbk, sym_scale, btol, max_sweeps, bw, bv, d_scale2, bstat,
                                      base + off, workspace_bytes - off, stream);
    if (rc < 0) return rc;
    for (int k = 0; k < nbk && status == 0; ++k)
      if (bstat[k] != 0) status = bidx[k] + 1;
  }
  if (n_large > 0) {
    const dim3 tb(32, 8);
    const int t32 = (max_large + 31) / 32;
    eigh_init_kernel<<<dim3(t32, t32, count), tb, 0, stream>>>(d_tab);
    note_launch();
    const int n = (max_large + 1) & ~1;
    int h_rot[64];
    bool converged = false;
    for (int sweep = 0; sweep < max_sweeps && !converged; ++sweep) {
      if (cudaMemsetAsync(d_rot, 0, 4 * count, stream) != cudaSuccess) return -5;
      for (int r = 0; r < n - 1; ++r) {
        eigh_round_kernel<<<dim3(n / 2, count), kRoundThreads, 0, stream>>>(d_tab, r, d_rot);
      }
      note_launch(n - 1);
      if (cudaMemcpyAsync(h_rot, d_rot, 4 * count, cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
          cudaStreamSynchronize(stream) != cudaSuccess)
        return -5;
      converged = true;
      for (int i = 0; i < count; ++i)
        if (dims[i] > kSmallMax && h_rot[i] != 0) converged = false;
    }
    if (!converged) {
      for (int i = 0; i < count && status == 0; ++i)
        if (dims[i] > kSmallMax && h_rot[i] != 0) status = i + 1;
    }
    eigh_rayleigh_kernel<<<dim3((max_large + 7) / 8, count), 256, 0, stream>>>(d_tab);
    note_launch();
    eigh_sort_kernel<<<dim3(t32, count), tb, 0, stream>>>(d_tab);
    note_launch();
  }
  if (cudaGetLastError() != cudaSuccess) return -5;
  return status;
}

}  // namespace bk
