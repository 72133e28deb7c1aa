// bk_small64.cu — fp64 small-matrix path of the sampling-free linearised predictive (SURVEY §8 row a14,
// BASELINE config 2: toy regression MLP, factors 2 x 2 ... 51 x 51).
//
// Reference (paths relative to /root/reference), sampling_free/regression/regression_ll_block.py:126-138:
//     q_inv = pinverse(N * (q_i + tau I));  h_inv = pinverse(N * (h_i + tau I))
//     std_j += | J_i kron(q_inv, h_inv) J_i^T |
// The damped factors of this problem have cond(R) = 1e5 .. 5e6 (measured with the reference itself, see
// tests/golden/make_golden_cfg2.py): the quadratic form is dominated by the SMALL eigen-directions of R, and
// fp32 arithmetic (the reference's own fp32 run included) resolves it to cond * 6e-8 = 1e-2 only.  BASELINE.json
// asks for 1e-3, so for factors that fit one CTA the inverse and the quadratic form run in fp64 on the SIMT
// pipes: they are a few hundred KFLOP, latency-bound, and B200's fp64 rate is irrelevant at this size.
//   spd_inverse_f64   R = multiply * (F + F^T)/2 + add * I  ->  R^-1 (fp64, full symmetric), one CTA per
//                     factor: Cholesky in shared memory, forward substitution for L^-1, then L^-T L^-1.
//                     (R is SPD for add > 0, so the script's pinverse is the inverse.)
//   kron_quadform_f64 out[b] (+)= | <V_b, Q V_b H^T> | with V_b = J_b.view(d_in', d_out) — the Kronecker-free
//                     identity of SURVEY §8 a14 — one CTA per test point, everything in shared memory.
#include "bk_common.cuh"
#include "bk_kernels.cuh"

namespace bk {

namespace {

constexpr int kInvThreads = 256;

struct Inv64Batch {
  const float* f[kSmall64MaxBatch];
  long long ld[kSmall64MaxBatch];
  double* out[kSmall64MaxBatch];
  int d[kSmall64MaxBatch];
  double add[kSmall64MaxBatch], mult[kSmall64MaxBatch];
};

// In shared memory: A (SPD, [d][ld]) -> L in its lower triangle; X (zero-filled) -> L^-1 (lower triangular).
// *bad receives the 1-based index of the first non-positive pivot.  All threads of the CTA call it.
__device__ void chol_trinv_shared(double* A, double* X, int d, int ld, int tid, int* bad) {
  // right-looking Cholesky, one column per step
  for (int k = 0; k < d; ++k) {
    if (tid == 0) {
      const double v = A[k * ld + k];
      if (!(v > 0.0) && *bad == 0) *bad = k + 1;
      A[k * ld + k] = sqrt(v > 0.0 ? v : 1.0);
    }
    __syncthreads();
    const double inv = 1.0 / A[k * ld + k];
    for (int i = k + 1 + tid; i < d; i += kInvThreads) A[i * ld + k] *= inv;
    __syncthreads();
    const int rem = d - k - 1;
    for (int e = tid; e < rem * rem; e += kInvThreads) {
      const int i = k + 1 + e / rem, j = k + 1 + e % rem;
      if (j <= i) A[i * ld + j] -= A[i * ld + k] * A[j * ld + k];
    }
    __syncthreads();
  }
  // X = L^-1: thread per column c, forward substitution down the rows
  for (int c = tid; c < d; c += kInvThreads) {
    for (int i = c; i < d; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int t = c; t < i; ++t) s -= A[i * ld + t] * X[t * ld + c];
      X[i * ld + c] = s / A[i * ld + i];
    }
  }
  __syncthreads();
}

// shared memory: A [d][d+1] doubles (R, then L in the lower triangle), X [d][d+1] (L^-1)
__global__ void __launch_bounds__(kInvThreads)
spd_inverse_f64_kernel(const __grid_constant__ Inv64Batch p, int* __restrict__ status) {
  extern __shared__ double sm64[];
  const int b = blockIdx.x;
  const int d = p.d[b];
  const int ld = d + 1;
  double* A = sm64;
  double* X = sm64 + static_cast<size_t>(d) * ld;
  const float* F = p.f[b];
  const long long ldf = p.ld[b];
  const int tid = threadIdx.x;
  __shared__ int bad;
  if (tid == 0) bad = 0;
  for (int e = tid; e < d * d; e += kInvThreads) {
    const int i = e / d, j = e - i * d;
    const double s = 0.5 * (static_cast<double>(F[i * ldf + j]) + static_cast<double>(F[j * ldf + i]));
    A[i * ld + j] = p.mult[b] * s + (i == j ? p.add[b] : 0.0);
    X[i * ld + j] = 0.0;
  }
  __syncthreads();
  chol_trinv_shared(A, X, d, ld, tid, &bad);
  // R^-1 = X^T X  (X lower triangular: sum over t >= max(i, j))
  double* out = p.out[b];
  for (int e = tid; e < d * d; e += kInvThreads) {
    const int i = e / d, j = e - i * d;
    double s = 0.0;
    for (int t = (i > j ? i : j); t < d; ++t) s += X[t * ld + i] * X[t * ld + j];
    out[e] = s;
  }
  if (tid == 0 && bad != 0 && status != nullptr) atomicCAS(status, 0, b * 65536 + bad);
}

// One diagonal block of the sharded dense-Fisher Cholesky (distributed.dense_fisher_sharded): R = sym(F) + add I,
// W = chol(R)^-1 written as an fp32 lower-triangular matrix (zero upper triangle), one CTA.
__global__ void __launch_bounds__(kInvThreads)
chol_trinv_f64_kernel(const float* __restrict__ F, long long ldf, int d, double add, float* __restrict__ W,
                      long long ldw, int* __restrict__ status) {
  extern __shared__ double sm64[];
  const int ld = d + 1;
  double* A = sm64;
  double* X = sm64 + static_cast<size_t>(d) * ld;
  const int tid = threadIdx.x;
  __shared__ int bad;
  if (tid == 0) bad = 0;
  for (int e = tid; e < d * d; e += kInvThreads) {
    const int i = e / d, j = e - i * d;
    // lower triangle is authoritative (the trailing updates of the blocked factorisation touch it only)
    const double s = static_cast<double>(i >= j ? F[i * ldf + j] : F[j * ldf + i]);
    A[i * ld + j] = s + (i == j ? add : 0.0);
    X[i * ld + j] = 0.0;
  }
  __syncthreads();
  chol_trinv_shared(A, X, d, ld, tid, &bad);
  for (int e = tid; e < d * d; e += kInvThreads) {
    const int i = e / d, j = e - i * d;
    W[i * ldw + j] = (j <= i) ? static_cast<float>(X[i * ld + j]) : 0.f;
  }
  if (tid == 0 && bad != 0 && status != nullptr) atomicCAS(status, 0, bad);
}

// one CTA per test point: Vs [dinp][dout], U = Q V [dinp][dout] in shared memory
__global__ void __launch_bounds__(256)
kron_quadform_f64_kernel(const float* __restrict__ V, long long stride_v, int dinp, int dout,
                         const double* __restrict__ Q, const double* __restrict__ H,
                         float* __restrict__ out, int accumulate) {
  extern __shared__ double sm64[];
  double* Vs = sm64;
  double* U = sm64 + dinp * dout;
  const int tid = threadIdx.x;
  const float* v = V + blockIdx.x * stride_v;
  for (int e = tid; e < dinp * dout; e += 256) Vs[e] = static_cast<double>(v[e]);
  __syncthreads();
  for (int e = tid; e < dinp * dout; e += 256) {
    const int i = e / dout, l = e - i * dout;
    double s = 0.0;
    for (int j = 0; j < dinp; ++j) s += Q[i * dinp + j] * Vs[j * dout + l];
    U[e] = s;
  }
  __syncthreads();
  double acc = 0.0;
  for (int e = tid; e < dinp * dout; e += 256) {
    const int i = e / dout, k = e - i * dout;
    double s = 0.0;
    for (int l = 0; l < dout; ++l) s += U[i * dout + l] * H[k * dout + l];
    acc += Vs[e] * s;
  }
  __shared__ double part[8];
  acc = warp_sum(acc);
  if ((tid & 31) == 0) part[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += part[w];
    const float r = static_cast<float>(fabs(t));
    out[blockIdx.x] = accumulate ? out[blockIdx.x] + r : r;
  }
}

}  // namespace

int launch_spd_inverse_f64(const float* const* factors, const long long* lds, const int* dims,
                           const double* add, const double* mult, double* const* outs, int count,
                           int* status, cudaStream_t stream) {
  if (count <= 0) return 0;
  if (count > kSmall64MaxBatch) return -2;
  Inv64Batch p{};
  int dmax = 0;
  for (int i = 0; i < count; ++i) {
    if (factors[i] == nullptr || outs[i] == nullptr || dims[i] <= 0 || dims[i] > kSmall64MaxDim ||
        lds[i] < dims[i])
      return -2;
    p.f[i] = factors[i];
    p.ld[i] = lds[i];
    p.out[i] = outs[i];
    p.d[i] = dims[i];
    p.add[i] = add[i];
    p.mult[i] = mult[i];
    if (dims[i] > dmax) dmax = dims[i];
  }
  const size_t smem = 2ull * dmax * (dmax + 1) * sizeof(double);
  static DeviceOnce attr_once;
  if (!attr_once([] {
        return cudaFuncSetAttribute(spd_inverse_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(2ull * kSmall64MaxDim * (kSmall64MaxDim + 1) *
                                                     sizeof(double))) == cudaSuccess;
      }))
    return -5;
  if (status != nullptr && cudaMemsetAsync(status, 0, sizeof(int), stream) != cudaSuccess) return -5;
  spd_inverse_f64_kernel<<<count, kInvThreads, smem, stream>>>(p, status);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_chol_trinv_f64(const float* F, long long ldf, int d, double add, float* W, long long ldw,
                          int* status, cudaStream_t stream) {
  if (d <= 0) return 0;
  if (F == nullptr || W == nullptr || d > kSmall64MaxDim || ldf < d || ldw < d) return -2;
  const size_t smem = 2ull * d * (d + 1) * sizeof(double);
  static DeviceOnce attr_once;
  if (!attr_once([] {
        return cudaFuncSetAttribute(chol_trinv_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(2ull * kSmall64MaxDim * (kSmall64MaxDim + 1) *
                                                     sizeof(double))) == cudaSuccess;
      }))
    return -5;
  chol_trinv_f64_kernel<<<1, kInvThreads, smem, stream>>>(F, ldf, d, add, W, ldw, status);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_kron_quadform_f64(const float* V, long long stride_v, int batch, int dinp, int dout,
                             const double* Q, const double* H, float* out, int accumulate,
                             cudaStream_t stream) {
  if (batch <= 0) return 0;
  if (dinp <= 0 || dout <= 0 || static_cast<long long>(dinp) * dout > kSmall64MaxElems) return -2;
  const size_t smem = 2ull * dinp * dout * sizeof(double);
  static DeviceOnce attr_once;
  if (!attr_once([] {
        return cudaFuncSetAttribute(kron_quadform_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(2ull * kSmall64MaxElems * sizeof(double))) == cudaSuccess;
      }))
    return -5;
  kron_quadform_f64_kernel<<<batch, 256, smem, stream>>>(V, stride_v, dinp, dout, Q, H, out, accumulate);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace bk
