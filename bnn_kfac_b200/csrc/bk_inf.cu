// bk_inf.cu — kernels of the INF curvature (low-rank eigenbasis + diagonal correction), SURVEY §8(f)
// row f4.  Reference: models/curvatures.py:476-682 (paths relative to /root/reference).
//
//   inf_regularise   INF.invert :537-539    correction[correction < 0] = 0 (in place);
//                                           reg_lambda = sqrt(s*lambda); reg_inv_correction = 1/sqrt(s*corr + n)
//   presample chain  INF.pre_sampler :565-585
//        V_s  = c (.) kron(U_A, U_G) diag(s)              ((n*m) x r, r = a*b — never materialised here)
//        vtv  = V_s^T V_s ;  A_c = chol(vtv) ;  B_c = chol(vtv + I)
//        C    = A_c^-T (B_c - I) A_c^-1 ;  L_c = (C^-1 + vtv)^-1 ;  P_c = diag(s) L_c diag(s)
//      The reference forms the (n*m) x r Kronecker matrix (663 MB for one 126 x 10 x 5 000 layer) and five
//      r x r LU inverses.  Here:
//        vtv[(q,x),(q',y)] = s s' sum_i U_A[i,q] U_A[i,q'] W_i[x,y],  W_i = U_G^T diag(c_i^2) U_G
//      (n*m*b^2 + n*a^2*b^2 flops instead of n*m*a^2*b^2), and with T = B_c - I lower triangular
//        C^-1 + vtv = A_c (T^-1 + I) A_c^T,  (T^-1 + I)^-1 = (I + T)^-1 T = I - B_c^-1
//        =>  L_c = A_c^-T (I - B_c^-1) A_c^-1
//      i.e. two Cholesky factorisations, two triangular inverses and two products.  A_c^-1 has norm
//      1/sqrt(lambda_min(vtv)) and is applied from both sides of a difference that is O(vtv): fp32 would
//      lose cond(vtv) * 6e-8 (cond is 1e5..1e7 at rank 10..30 on the reference's own MLP).  The whole
//      r x r chain therefore runs in fp64 on the SIMT pipes; it is called once per invert() and r is
//      small (<= rank^2), so it is latency-, not throughput-bound.
//   inf_combine      INF.sampler :611        Y_l - c^2 (.) X_p_s^T  (the dense products of the sampler are
//                                           calls of the tensor-core contraction core)
#include "bk_common.cuh"
#include "bk_kernels.cuh"

namespace bk {

namespace {

constexpr int kNB = 32;  // Cholesky block size

__global__ void __launch_bounds__(256)
inf_regularise_kernel(float* __restrict__ corr, long long nm, const float* __restrict__ lam, long long r,
                      float add, float mult, float* __restrict__ ric, float* __restrict__ rl) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long total = nm > r ? nm : r;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += stride) {
    if (i < nm) {
      float v = corr[i];
      if (v < 0.f) v = 0.f;
      corr[i] = v;
      ric[i] = sqrtf(1.0f / fmaf(mult, v, add));
    }
    if (i < r) rl[i] = sqrtf(mult * lam[i]);
  }
}

__global__ void __launch_bounds__(256)
inf_combine_kernel(float* __restrict__ out, const float* __restrict__ yl, const float* __restrict__ c,
                   const float* __restrict__ xt, long long count) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < count; i += stride) {
    const float ci = c[i];
    out[i] = fmaf(-ci * ci, xt[i], yl[i]);
  }
}

// W[i][x][y] = sum_p c[i*m + p]^2 G[p][x] G[p][y]   (fp64; grid.z = i, 16 x 16 output tile per CTA)
__global__ void __launch_bounds__(256)
inf_w_kernel(const float* __restrict__ G, long long ldg, int m, int b, const float* __restrict__ c,
             double* __restrict__ W) {
  __shared__ double sx[16][17], sy[16][17], sc[16];
  const int i = blockIdx.z, x0 = blockIdx.y * 16, y0 = blockIdx.x * 16;
  const int tx = threadIdx.x, ty = threadIdx.y;
  double acc = 0.0;
  for (int p0 = 0; p0 < m; p0 += 16) {
    const int p = p0 + ty;
    sx[ty][tx] = (p < m && x0 + tx < b) ? static_cast<double>(G[p * ldg + x0 + tx]) : 0.0;
    sy[ty][tx] = (p < m && y0 + tx < b) ? static_cast<double>(G[p * ldg + y0 + tx]) : 0.0;
    if (ty == 0) {
      const double cv = (p0 + tx < m) ? static_cast<double>(c[static_cast<long long>(i) * m + p0 + tx]) : 0.0;
      sc[tx] = cv * cv;
    }
    __syncthreads();
#pragma unroll
    for (int pp = 0; pp < 16; ++pp) acc = fma(sc[pp] * sx[pp][ty], sy[pp][tx], acc);
    __syncthreads();
  }
  const int x = x0 + ty, y = y0 + tx;
  if (x < b && y < b) W[(static_cast<long long>(i) * b + x) * b + y] = acc;
}

// M1[X][Y] = s[X] s[Y] sum_i A[i][q] A[i][q'] W[i][x][y],  X = q*b + x, Y = q'*b + y;  M2 = M1 + I.
// Lower triangle only (the Cholesky kernels read nothing else).
__global__ void __launch_bounds__(256)
inf_vtv_kernel(const float* __restrict__ A, long long lda, int n, int a, int b,
               const double* __restrict__ W, const float* __restrict__ s, double* __restrict__ M1,
               double* __restrict__ M2) {
  const int r = a * b;
  const int Y = blockIdx.x * 16 + threadIdx.x, X = blockIdx.y * 16 + threadIdx.y;
  if (X >= r || Y >= r || Y > X) return;
  const int q = X / b, x = X - q * b, q2 = Y / b, y = Y - q2 * b;
  double acc = 0.0;
  const double* w = W + static_cast<long long>(x) * b + y;
  const long long wstride = static_cast<long long>(b) * b;
  for (int i = 0; i < n; ++i)
    acc = fma(static_cast<double>(A[i * lda + q]) * static_cast<double>(A[i * lda + q2]), w[i * wstride], acc);
  acc *= static_cast<double>(s[X]) * static_cast<double>(s[Y]);
  M1[static_cast<long long>(X) * r + Y] = acc;
  M2[static_cast<long long>(X) * r + Y] = acc + (X == Y ? 1.0 : 0.0);
}

// ---- blocked right-looking fp64 Cholesky (lower, in place), blockIdx.z / .y = matrix of the batch
__global__ void __launch_bounds__(1024)
chol64_diag_kernel(double* __restrict__ M, int r, int k0, int* __restrict__ info) {
  __shared__ double s[kNB][kNB + 1];
  double* mat = M + static_cast<long long>(blockIdx.x) * r * r;
  const int kb = min(kNB, r - k0);
  const int i = threadIdx.y, j = threadIdx.x;
  const bool in = i < kb && j < kb;
  s[i][j] = (in && j <= i) ? mat[static_cast<long long>(k0 + i) * r + k0 + j] : 0.0;
  __syncthreads();
  for (int jj = 0; jj < kb; ++jj) {
    if (i == jj && j == jj) {
      const double d = s[jj][jj];
      if (!(d > 0.0)) {
        atomicCAS(info, 0, static_cast<int>(blockIdx.x) + 1);
        s[jj][jj] = 1.0;
      } else {
        s[jj][jj] = sqrt(d);
      }
    }
    __syncthreads();
    if (j == jj && i > jj && i < kb) s[i][jj] /= s[jj][jj];
    __syncthreads();
    if (in && j > jj && j <= i) s[i][j] -= s[i][jj] * s[j][jj];
    __syncthreads();
  }
  if (in && j <= i) mat[static_cast<long long>(k0 + i) * r + k0 + j] = s[i][j];
}

// rows below a full kNB block: X L_kk^T = A  (thread per row, L_kk in shared memory)
__global__ void __launch_bounds__(128)
chol64_panel_kernel(double* __restrict__ M, int r, int k0) {
  __shared__ double l[kNB][kNB + 1];
  double* mat = M + static_cast<long long>(blockIdx.y) * r * r;
  for (int t = threadIdx.x; t < kNB * kNB; t += blockDim.x) {
    const int i = t / kNB, j = t % kNB;
    l[i][j] = mat[static_cast<long long>(k0 + i) * r + k0 + j];
  }
  __syncthreads();
  const int row = k0 + kNB + blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= r) return;
  double* a = mat + static_cast<long long>(row) * r + k0;
  double x[kNB];
#pragma unroll
  for (int j = 0; j < kNB; ++j) x[j] = a[j];
#pragma unroll
  for (int j = 0; j < kNB; ++j) {
    double v = x[j];
#pragma unroll
    for (int t = 0; t < j; ++t) v = fma(-x[t], l[j][t], v);
    x[j] = v / l[j][j];
  }
#pragma unroll
  for (int j = 0; j < kNB; ++j) a[j] = x[j];
}

// trailing update: C[i][j] -= sum_t L[i][k0+t] L[j][k0+t] for the lower 32 x 32 tiles of [k0+32, r)
__global__ void __launch_bounds__(256)
chol64_trail_kernel(double* __restrict__ M, int r, int k0) {
  if (blockIdx.x > blockIdx.y) return;
  __shared__ double li[kNB][kNB + 1], lj[kNB][kNB + 1];
  double* mat = M + static_cast<long long>(blockIdx.z) * r * r;
  const int base = k0 + kNB;
  const int i0 = base + blockIdx.y * kNB, j0 = base + blockIdx.x * kNB;
  const int tid = threadIdx.y * 16 + threadIdx.x;
  for (int t = tid; t < kNB * kNB; t += 256) {
    const int rr = t / kNB, cc = t % kNB;
    li[rr][cc] = (i0 + rr < r) ? mat[static_cast<long long>(i0 + rr) * r + k0 + cc] : 0.0;
    lj[rr][cc] = (j0 + rr < r) ? mat[static_cast<long long>(j0 + rr) * r + k0 + cc] : 0.0;
  }
  __syncthreads();
  double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
  for (int t = 0; t < kNB; ++t) {
    const double a0 = li[threadIdx.y][t], a1 = li[threadIdx.y + 16][t];
    const double b0 = lj[threadIdx.x][t], b1 = lj[threadIdx.x + 16][t];
    acc[0][0] = fma(a0, b0, acc[0][0]);
    acc[0][1] = fma(a0, b1, acc[0][1]);
    acc[1][0] = fma(a1, b0, acc[1][0]);
    acc[1][1] = fma(a1, b1, acc[1][1]);
  }
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const int i = i0 + threadIdx.y + 16 * u, j = j0 + threadIdx.x + 16 * v;
      if (i < r && j <= i) mat[static_cast<long long>(i) * r + j] -= acc[u][v];
    }
}

// X = L^-1 (lower triangular, upper part written as zero): one thread per column, forward substitution.
// Threads of a warp walk the same rows of L (broadcast loads) and adjacent columns of X (coalesced).
__global__ void __launch_bounds__(128)
trtri64_kernel(const double* __restrict__ Lm, double* __restrict__ Xm, int r) {
  const double* L = Lm + static_cast<long long>(blockIdx.y) * r * r;
  double* X = Xm + static_cast<long long>(blockIdx.y) * r * r;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= r) return;
  for (int i = 0; i < j; ++i) X[static_cast<long long>(i) * r + j] = 0.0;
  X[static_cast<long long>(j) * r + j] = 1.0 / L[static_cast<long long>(j) * r + j];
  for (int i = j + 1; i < r; ++i) {
    const double* lrow = L + static_cast<long long>(i) * r;
    double s0 = 0.0, s1 = 0.0;
    int k = j;
    for (; k + 1 < i; k += 2) {
      s0 = fma(lrow[k], X[static_cast<long long>(k) * r + j], s0);
      s1 = fma(lrow[k + 1], X[static_cast<long long>(k + 1) * r + j], s1);
    }
    if (k < i) s0 = fma(lrow[k], X[static_cast<long long>(k) * r + j], s0);
    X[static_cast<long long>(i) * r + j] = -(s0 + s1) / lrow[i];
  }
}

// C = alpha * op(X) Y + beta * C0 (r x r, fp64), optional symmetric scaling s[i] * C[i][j] * s[j] and fp32 output.
template <bool TA>
__global__ void __launch_bounds__(256)
gemm64_kernel(const double* __restrict__ X, const double* __restrict__ Y, int r, double alpha,
              const double* __restrict__ C0, double beta, const float* __restrict__ scale,
              double* __restrict__ out64, float* __restrict__ out32) {
  __shared__ double sa[16][33], sb[16][33];  // sa[k][i], sb[k][j]
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tid = threadIdx.y * 16 + threadIdx.x;
  double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  for (int k0 = 0; k0 < r; k0 += 16) {
    for (int t = tid; t < 16 * 32; t += 256) {
      int kk, ii;
      if (TA) {  // op(X)[i][k] = X[k][i]: consecutive threads walk i
        kk = t / 32;
        ii = t % 32;
        sa[kk][ii] = (k0 + kk < r && i0 + ii < r) ? X[static_cast<long long>(k0 + kk) * r + i0 + ii] : 0.0;
      } else {   // X[i][k]: consecutive threads walk k
        ii = t / 16;
        kk = t % 16;
        sa[kk][ii] = (k0 + kk < r && i0 + ii < r) ? X[static_cast<long long>(i0 + ii) * r + k0 + kk] : 0.0;
      }
      const int kb = t / 32, jj = t % 32;
      sb[kb][jj] = (k0 + kb < r && j0 + jj < r) ? Y[static_cast<long long>(k0 + kb) * r + j0 + jj] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const double a0 = sa[kk][threadIdx.y], a1 = sa[kk][threadIdx.y + 16];
      const double b0 = sb[kk][threadIdx.x], b1 = sb[kk][threadIdx.x + 16];
      acc[0][0] = fma(a0, b0, acc[0][0]);
      acc[0][1] = fma(a0, b1, acc[0][1]);
      acc[1][0] = fma(a1, b0, acc[1][0]);
      acc[1][1] = fma(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const int i = i0 + threadIdx.y + 16 * u, j = j0 + threadIdx.x + 16 * v;
      if (i >= r || j >= r) continue;
      const long long o = static_cast<long long>(i) * r + j;
      double val = alpha * acc[u][v];
      if (C0 != nullptr) val += beta * C0[o];
      if (scale != nullptr) val *= static_cast<double>(scale[i]) * static_cast<double>(scale[j]);
      if (out64 != nullptr) out64[o] = val;
      if (out32 != nullptr) out32[o] = static_cast<float>(val);
    }
}

inline int ok() { return cudaGetLastError() == cudaSuccess ? 0 : -5; }
inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace

int launch_inf_regularise(float* corr, long long nm, const float* lam, long long r, float add, float mult,
                          float* ric, float* rl, cudaStream_t stream) {
  const long long total = nm > r ? nm : r;
  if (total <= 0) return 0;
  long long blocks = (total + 255) / 256;
  if (blocks > kNumSMsB200 * 16) blocks = kNumSMsB200 * 16;
  inf_regularise_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(corr, nm, lam, r, add, mult, ric, rl);
  note_launch();
  return ok();
}

int launch_inf_combine(float* out, const float* yl, const float* c, const float* xt, long long count,
                       cudaStream_t stream) {
  if (count <= 0) return 0;
  long long blocks = (count + 255) / 256;
  if (blocks > kNumSMsB200 * 16) blocks = kNumSMsB200 * 16;
  inf_combine_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(out, yl, c, xt, count);
  note_launch();
  return ok();
}

// workspace: [info: 256 B][W: n*b*b][M1, M2: 2 r^2][X1, X2: 2 r^2][T: r^2]  (doubles)
size_t inf_presample_workspace_bytes(int n, int a, int m, int b) {
  (void)m;
  const size_t r = static_cast<size_t>(a) * b;
  return 256 + up256(static_cast<size_t>(n) * b * b * 8) + 5 * up256(r * r * 8);
}

int inf_presample(const float* ua, long long lda, int n, int a, const float* ug, long long ldg, int m, int b,
                  const float* ric, const float* rl, float* p_out, void* workspace, size_t workspace_bytes,
                  cudaStream_t stream) {
  if (n <= 0 || a <= 0 || m <= 0 || b <= 0 || n > 65535) return -2;
  if (workspace_bytes < inf_presample_workspace_bytes(n, a, m, b)) return -6;
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return -6;
  const int r = a * b;
  const size_t rr = up256(static_cast<size_t>(r) * r * 8);
  char* base = static_cast<char*>(workspace);
  int* info = reinterpret_cast<int*>(base);
  double* W = reinterpret_cast<double*>(base + 256);
  char* p = base + 256 + up256(static_cast<size_t>(n) * b * b * 8);
  // M1 and M2 (and X1, X2) must be r*r doubles apart for the batched kernels: carve them unpadded
  double* M1 = reinterpret_cast<double*>(p);
  double* M2 = M1 + static_cast<size_t>(r) * r;
  double* X1 = reinterpret_cast<double*>(p + 2 * rr);
  double* X2 = X1 + static_cast<size_t>(r) * r;
  double* T = reinterpret_cast<double*>(p + 4 * rr);
  if (cudaMemsetAsync(info, 0, 256, stream) != cudaSuccess) return -5;

  const dim3 b16(16, 16);
  const int tb = (b + 15) / 16, tr16 = (r + 15) / 16, tr32 = (r + 31) / 32;
  inf_w_kernel<<<dim3(tb, tb, n), b16, 0, stream>>>(ug, ldg, m, b, ric, W);
  inf_vtv_kernel<<<dim3(tr16, tr16), b16, 0, stream>>>(ua, lda, n, a, b, W, rl, M1, M2);
  note_launch(2);
  for (int k0 = 0; k0 < r; k0 += kNB) {
    chol64_diag_kernel<<<2, dim3(kNB, kNB), 0, stream>>>(M1, r, k0, info);
    note_launch();
    const int rest = r - k0 - kNB;
    if (rest > 0) {
      chol64_panel_kernel<<<dim3((rest + 127) / 128, 2), 128, 0, stream>>>(M1, r, k0);
      const int tt = (rest + kNB - 1) / kNB;
      chol64_trail_kernel<<<dim3(tt, tt, 2), b16, 0, stream>>>(M1, r, k0);
      note_launch(2);
    }
  }
  trtri64_kernel<<<dim3((r + 127) / 128, 2), 128, 0, stream>>>(M1, X1, r);
  // T = A^-1 - B^-1 A^-1 ;  P = diag(s) (A^-T T) diag(s)
  gemm64_kernel<false><<<dim3(tr32, tr32), b16, 0, stream>>>(X2, X1, r, -1.0, X1, 1.0, nullptr, T, nullptr);
  gemm64_kernel<true><<<dim3(tr32, tr32), b16, 0, stream>>>(X1, T, r, 1.0, nullptr, 0.0, rl, nullptr, p_out);
  note_launch(3);
  if (cudaGetLastError() != cudaSuccess) return -5;
  int h = 0;
  if (cudaMemcpyAsync(&h, info, sizeof(int), cudaMemcpyDeviceToHost, stream) != cudaSuccess) return -5;
  if (cudaStreamSynchronize(stream) != cudaSuccess) return -5;
  return h;
}

}  // namespace bk
