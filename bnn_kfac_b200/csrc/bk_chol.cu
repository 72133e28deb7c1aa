// bk_chol.cu — damped inversion of Kronecker factors, batched over factors.
//
// Reference (models/curvatures.py:381-392):
//     R = sqrt(s)*F + sqrt(n)*I ;  R <- (R + R^T)/2 ;  L = R.inverse().cholesky()      (L lower)
//
// L is obtained without forming R^-1:  R^-1 = L L^T  <=>  R = U U^T with U = L^-T upper.  With the
// index flip P (i -> d-1-i):  C = cholesky_lower(P R P),  U = P C P,  L = P C^-T P, i.e.
//     L[i][j] = Cinv[d-1-j][d-1-i].
// Same unique factor as the reference (positive diagonal), 2/3 d^3 flops instead of 7/3 d^3.
//
// Blocked algorithm, NB = 64, every step one launch for ALL factors of the batch (device-side
// problem table; CTAs of factors that are already finished exit immediately):
//   phase 1 (right-looking Cholesky)   potrf_diag -> panel (A21 L11^-T) -> trailing A22 -= L21 L21^T
//   phase 2 (triangular inverse, X = C^-1 built by block eliminations from X = I)
//                                       rowscale X_k = L_kk^-1 X_k -> X_below -= L_below,k X_k
// Two-level blocking for wide factors (padded size >= 2048): the rank-64 fp32 SIMT updates
// (128x128 register-tiled) only reach to the end of the current 256-wide outer block; everything
// beyond it receives the whole outer block at once from the tcgen05 contraction core
// (bk_umma_gemm.cu) with K = 256 and the TMA reduce-add epilogue.  Precision: cond(R) reaches 1e4-1e6
// at the reference's damping values and the tolerance on the inverse is 1e-3, which rules out bf16
// (1e-3 per product) and bf16x3 (1e-5) trailing updates; the operands are therefore split THREE ways
// (hi + lo + lo2 = 24 mantissa bits) and accumulated in six tensor-core passes (bf16x6), which is
// fp32-class and still ~4x the SIMT fp32 rate.
#include "bk_common.cuh"
#include "bk_kernels.cuh"
#include "bk_umma_gemm.cuh"

#include <mutex>
#include <vector>

namespace bk {

namespace {

constexpr int NB = 64;
constexpr int kPad = 4;

struct CholProb {
  const float* F;  // [d, d] contiguous input factor
  float* out;      // [d, d] contiguous output (lower-triangular L)
  float* R;        // [dpad, dpad] workspace: flipped damped matrix, then its Cholesky factor C
  float* X;        // [dpad, dpad] workspace: C^-1
  float* Dinv;     // [nb][64][64] inverses of the diagonal blocks of C
  int d, dpad, nb;
  float sqrt_s, sqrt_n;
};

// Programmatic dependent launch: the steps of both phases are chains of small dependent kernels, and the
// launch latency between two of them (~2-3 us) is comparable to the kernels themselves.  Every chain kernel
// first waits for the grids it depends on (complete and flushed) and then lets its own dependents be
// scheduled; launched with cudaLaunchAttributeProgrammaticStreamSerialization the next kernel's CTAs are
// resident and parked in their wait when this grid finishes.  Without the attribute both are no-ops.
__device__ __forceinline__ void pdl_wait_then_trigger() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... Args>
inline void launch_chained(void (*kernel)(Args...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                           Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, args...);
}

// ------------------------------------------------------------------ damping + flip + padding
// R[i][j] = sqrt_s * (F[fi][fj] + F[fj][fi]) / 2 + sqrt_n * (i == j),  fi = d-1-i, fj = d-1-j;
// identity on the padding; X = I.
__global__ void damp_flip_kernel(const CholProb* __restrict__ tab) {
  const CholProb p = tab[blockIdx.z];
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  if (i0 >= p.dpad || j0 >= p.dpad) return;
  __shared__ float tr[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;  // (32, 8)
  // transposed source tile: F[fj][fi], read with fi (i) fastest -> coalesced
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int j = j0 + ty + 8 * r, i = i0 + tx;
    float v = 0.f;
    if (i < p.d && j < p.d)
      v = p.F[static_cast<long long>(p.d - 1 - j) * p.d + (p.d - 1 - i)];
    tr[ty + 8 * r][tx] = v;  // tr[j_local][i_local]
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty + 8 * r, j = j0 + tx;
    float v, x = (i == j) ? 1.f : 0.f;
    if (i < p.d && j < p.d) {
      const float a = p.F[static_cast<long long>(p.d - 1 - i) * p.d + (p.d - 1 - j)];
      const float b = tr[tx][ty + 8 * r];
      v = p.sqrt_s * a + ((i == j) ? p.sqrt_n : 0.f);
      const float vt = p.sqrt_s * b + ((i == j) ? p.sqrt_n : 0.f);
      v = (v + vt) * 0.5f;
    } else {
      v = x;
    }
    p.R[static_cast<long long>(i) * p.dpad + j] = v;
    p.X[static_cast<long long>(i) * p.dpad + j] = x;
  }
}

// ------------------------------------------------------------------ diagonal block: potrf + inverse
// One CTA factorises the 64 x 64 diagonal block AND inverts the factor, in REGISTERS and in the same sweep.
// Thread (c, h) owns column c, rows 32 h .. 32 h + 31, in ONE register array v[]: column c of A until step
// j = c (right-looking Cholesky), column c of X = L^-1 afterwards (forward substitution on I) - X[j][c] is zero
// for c > j and columns c < j of A are final, so a column is in exactly one of the two states:
//     step j:  owners publish column j of A and row j of X (double-buffered shared vectors), ONE barrier, then
//     c > j :  A[i][c] -= L[i][j] L[c][j]              (i > j)     L[i][j] L[c][j] = colj[i] colj[c] / piv
//     c == j:  L[:, j] goes to a shared tile; X[i][j] = -colj[i] / piv, X[j][j] = 1 / L[j][j]
//     c < j :  X[i][c] -= L[i][j] X[j][c]              (i > j)
// i.e. v[r] = fma(-colj[i], f, v[r]) with a per-thread scalar f for everybody.  Rows are blocked, not cyclic: the
// warps of the upper half have nothing left to do in the second half of the sweep and fall through to the barrier.
// The sweep is a chain of 64 dependent steps, so what counts is the latency of one step: four warps (one per
// scheduler) with 32 independent FMAs each behind every published column, shared memory addressed through
// precomputed 32-bit offsets (the compiler re-derived generic addresses from S2R SR_CgaCtaId / SR_TID inside
// every step of the C++ version, on the critical path in front of the publishing stores and the broadcast loads),
// a bare MUFU.RSQ (pivots are >= the damping term, no denormal fix-up), no divergent A / X paths.
// History: block in shared memory, two barriers per step, inverse afterwards: 64-75 us per block; registers,
// 8 warps x 16 rows, separate A and X arrays: 23.3 us (~670 cycles per step, 70 instructions per warp per
// step); this version: see profiles/r02_potrf_diag.md.
constexpr int kDiagThreads = 128;
constexpr int kLsPitch = NB + 1;

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float rsqrt_fast(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kDiagThreads)
potrf_diag_kernel(const CholProb* __restrict__ tab, int k, int* __restrict__ info, int fuse_prev) {
  pdl_wait_then_trigger();
  const CholProb p = tab[blockIdx.x];
  if (k >= p.nb) return;
  // [0, 512): colj[2][64]   [512, 1024): xrow[2][64]   [1024, ...): L tile [64][65]
  __shared__ __align__(16) float sh[4 * NB + NB * kLsPitch];
  __shared__ int bad;
  uint32_t sbase = static_cast<uint32_t>(__cvta_generic_to_shared(sh));
  asm volatile("mov.u32 %0, %0;" : "+r"(sbase));  // opaque: keep it in a register (no S2R re-derivation per step)
  const int tid = threadIdx.x;
  float* blk = p.R + static_cast<long long>(k) * NB * p.dpad + k * NB;
  if (tid == 0) bad = 0;
  const int c = tid % NB, h = tid / NB;  // column, row half (0..1)
  constexpr int R = NB / 2;
  const int i0 = R * h;
  float v[R];
#pragma unroll
  for (int r = 0; r < R; ++r) v[r] = blk[static_cast<long long>(i0 + r) * p.dpad + c];
  if (fuse_prev) {
    // Look-ahead: the update of THIS block by the previous step, A_kk -= L_k,k-1 L_k,k-1^T, is applied here instead
    // of by a launch of its own between the panel and this kernel.  L_k,k-1 (64 x 64, just written by the panel
    // kernel) goes transposed into the (still unused) L tile: Pt[kk][i], so that a thread reads its 32 rows as
    // eight broadcast float4 and its own column c as one conflict-free word per kk.  Same accumulation order as
    // rank64_kernel (kk ascending, one fma each, then old - acc): the result has the same bits.
    float* Pt = sh + 4 * NB;  // [64][64]
    {
      const int i = tid % NB, kh = tid / NB;  // row of L_k,k-1, half of its 64 columns
      const float* src = blk - NB + static_cast<long long>(i) * p.dpad + 32 * kh;
      float4 q[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) q[u] = *reinterpret_cast<const float4*>(src + 4 * u);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        Pt[(32 * kh + 4 * u + 0) * NB + i] = q[u].x;
        Pt[(32 * kh + 4 * u + 1) * NB + i] = q[u].y;
        Pt[(32 * kh + 4 * u + 2) * NB + i] = q[u].z;
        Pt[(32 * kh + 4 * u + 3) * NB + i] = q[u].w;
      }
    }
    __syncthreads();
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
#pragma unroll 2
    for (int kk = 0; kk < NB; ++kk) {
      const float pc = Pt[kk * NB + c];
#pragma unroll
      for (int r = 0; r < R; r += 4) {
        const float4 pi = *reinterpret_cast<const float4*>(&Pt[kk * NB + i0 + r]);
        acc[r] = fmaf(pi.x, pc, acc[r]);
        acc[r + 1] = fmaf(pi.y, pc, acc[r + 1]);
        acc[r + 2] = fmaf(pi.z, pc, acc[r + 2]);
        acc[r + 3] = fmaf(pi.w, pc, acc[r + 3]);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = fmaf(-1.f, acc[r], v[r]);
    __syncthreads();  // the L tile is written from the first column step on
  }
  const uint32_t a_col = sbase + 4u * static_cast<uint32_t>(i0);  // + 256 * buffer: this thread's rows of colj
  const uint32_t a_c = sbase + 4u * static_cast<uint32_t>(c);     // + 256 * buffer: colj[c]; + 512: xrow[c]
  const uint32_t a_ls = sbase + 4u * (4 * NB + static_cast<uint32_t>(c) * kLsPitch);  // L tile row c
  bool nonpos = false;
  // The column loop is unrolled by the 32 rows a thread owns: row j of X is then the STATICALLY indexed register
  // v[jj] of its owners (a rolled loop makes the compiler index a local-memory copy of v[] dynamically).
#pragma unroll 1
  for (int jb = 0; jb < NB / R; ++jb) {
    const bool row_owner = jb == h;  // this thread holds rows 32 jb .. 32 jb + 31 (warp-uniform)
    // row i0 + r lies below row j = 32 jb + jj  <=>  h > jb (warp-uniform) or, for the owners of half jb,
    // r > jj - a compile-time fact in the unrolled body: no per-element compare
    const bool below = h > jb;
#pragma unroll
  for (int jj = 0; jj < R; ++jj) {
    const int j = jb * R + jj;
    const uint32_t buf = 256u * (jj & 1);
    if (h >= jb) {
      if (c == j) {
#pragma unroll
        for (int r = 0; r < R; r += 4) sts128(a_col + buf + 4u * r, v[r], v[r + 1], v[r + 2], v[r + 3]);
      }
      if (row_owner) sts32(a_c + 512u + buf, v[jj]);  // X[j][c] before the division by L[j][j] (c < j)
    }
    __syncthreads();
    if (h < jb) continue;  // all rows of this warp are final (warp-uniform)
    float piv = lds32(sbase + buf + 4u * j);
    const float fc = lds32(a_c + buf);          // A[c][j]  (meaningful for c > j)
    const float fx = lds32(a_c + 512u + buf);   // X[j][c]  (meaningful for c < j)
    float cv[R];
#pragma unroll
    for (int r = 0; r < R; r += 4) {
      const float4 q = lds128(a_col + buf + 4u * r);
      cv[r] = q.x; cv[r + 1] = q.y; cv[r + 2] = q.z; cv[r + 3] = q.w;
    }
    if (!(piv > 0.f)) {  // not positive definite (or NaN)
      nonpos = true;
      piv = 1.f;
    }
    const float rs = rsqrt_fast(piv);  // 1 / L[j][j]
    const float ipiv = rs * rs;
    float f = ((c > j) ? fc : fx) * ipiv;
    if (c == j) {  // the column changes state: X[:, j] starts from zero, f = X[j][j] / piv = 1 / piv
      f = ipiv;
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = 0.f;
    }
    if (h == 1) {  // column j of L (rows c >= j) into the shared tile; these two warps are active to the end
      const float lcj = (c == j) ? piv * rs : fc * rs;
      if (c >= j) sts32(a_ls + 4u * j, lcj);
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (below || r > jj) v[r] = fmaf(-cv[r], f, v[r]);
    if (row_owner && c <= j) v[jj] = (c == j) ? rs : fx * rs;  // final X[j][c]
  }
  }
  if (nonpos) bad = 1;
  __syncthreads();
  float* di = p.Dinv + static_cast<long long>(k) * NB * NB;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i = i0 + r;
    if (c <= i) blk[static_cast<long long>(i) * p.dpad + c] = sh[4 * NB + i * kLsPitch + c];
    di[i * NB + c] = (c <= i) ? v[r] : 0.f;
  }
  if (tid == 0 && bad) atomicCAS(&info[blockIdx.x], 0, k + 1);
}

// ------------------------------------------------------------------ rank-64 update
enum Mode : int { kPanel = 0, kTrail = 1, kRowScale = 2, kXUpdate = 3 };
// kTrail only (look-ahead): the first 64 x 64 block of the region - the NEXT diagonal block - is updated by a
// one-CTA launch on the chain (kDiagOnly); everything else (kSkipDiag) runs beside the next diagonal-block kernel.
constexpr int kModeMask = 3, kDiagOnly = 16, kSkipDiag = 32;

// C[m x n] = beta*C + alpha * A[m x 64] * (NT ? B[n x 64]^T : B[64 x n]);  one T x T tile / CTA, T = 128 (8 x 8
// outputs per thread) or 64 (4 x 4).  The steps are latency-bound chains: whenever the 128-wide tiling would leave
// most SMs without a CTA the host picks T = 64 - four times the CTAs, a quarter of the loads and FMAs in front of
// each CTA's stores (panel of a 4097-wide factor: 32 CTAs x ~15 us -> 64+ CTAs x ~6 us).
// `limit` (two-level blocking, 0 = none): the rank-64 update only reaches up to row / column `limit`
// of the matrix (the end of the current 256-wide outer block); everything beyond it receives the
// whole outer block at once from the tensor-core GEMM (see chol_inv_batched).
template <int T>
__global__ void __launch_bounds__(256)
rank64_kernel(const CholProb* __restrict__ tab, int k, int mode_flags, int limit) {
  pdl_wait_then_trigger();
  const int mode = mode_flags & kModeMask;
  const bool diag_only = (mode_flags & kDiagOnly) != 0, skip_diag = (mode_flags & kSkipDiag) != 0;
  constexpr int TT = T / 16;  // outputs per thread and dimension; also 128-bit loads per thread and operand
  constexpr int V = TT / 4;
  const CholProb p = tab[blockIdx.y];
  if (k >= p.nb) return;
  const long long ld = p.dpad;
  const int k0 = k * NB, k1 = (k + 1) * NB;
  float* C;
  const float *A, *B;
  long long lda, ldb;
  int m, n;
  bool nt, lower = false;
  float alpha, beta;
  if (mode == kPanel) {
    m = p.dpad - k1; n = NB;
    C = p.R + k1 * ld + k0; A = C; lda = ld;
    B = p.Dinv + static_cast<long long>(k) * NB * NB; ldb = NB; nt = true;
    alpha = 1.f; beta = 0.f;
  } else if (mode == kTrail) {
    m = n = p.dpad - k1;
    if (limit > 0) n = min(limit, p.dpad) - k1;  // columns inside the outer block only
    C = p.R + k1 * ld + k1; A = p.R + k1 * ld + k0; lda = ld; B = A; ldb = ld; nt = true;
    lower = true; alpha = -1.f; beta = 1.f;
  } else if (mode == kRowScale) {
    m = NB; n = k1;
    C = p.X + k0 * ld; A = p.Dinv + static_cast<long long>(k) * NB * NB; lda = NB;
    B = C; ldb = ld; nt = false;
    alpha = 1.f; beta = 0.f;
  } else {
    m = p.dpad - k1; n = k1;
    if (limit > 0) m = min(limit, p.dpad) - k1;  // rows inside the outer block only
    C = p.X + k1 * ld; A = p.R + k1 * ld + k0; lda = ld; B = p.X + k0 * ld; ldb = ld; nt = false;
    alpha = -1.f; beta = 1.f;
  }
  if (m <= 0 || n <= 0) return;
  const int tiles_m = (m + T - 1) / T, tiles_n = (n + T - 1) / T;
  int ti, tj;
  if (lower && limit > 0) {
    // rectangular m x n region (n <= 192) of the lower triangle: plain enumeration + masking
    ti = blockIdx.x / tiles_n;
    tj = blockIdx.x - ti * tiles_n;
    if (ti >= tiles_m || tj * T > ti * T + T - 1) return;
  } else if (lower) {
    const int t = blockIdx.x;
    ti = static_cast<int>((sqrtf(8.f * t + 1.f) - 1.f) * 0.5f);
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    while (ti * (ti + 1) / 2 > t) --ti;
    tj = t - ti * (ti + 1) / 2;
    if (ti >= tiles_m) return;
  } else {
    ti = blockIdx.x / tiles_n;
    tj = blockIdx.x - ti * tiles_n;
    if (ti >= tiles_m) return;
  }
  if (diag_only && (ti != 0 || tj != 0)) return;
  if (skip_diag && T == NB && ti == 0 && tj == 0) return;
  const int r0 = ti * T, c0 = tj * T;

  extern __shared__ float sm[];
  float(*As)[T + kPad] = reinterpret_cast<float(*)[T + kPad]>(sm);
  float(*Bs)[T + kPad] = reinterpret_cast<float(*)[T + kPad]>(sm + NB * (T + kPad));
  const int tid = threadIdx.x;
  // Operand tiles: every thread issues ALL of its 128-bit loads (TT per operand) before the first
  // shared-memory store, so one DRAM / L2 round trip is exposed per tile instead of 4 TT.
  {
    float4 va[TT];
#pragma unroll
    for (int u = 0; u < TT; ++u) {
      const int q = tid + 256 * u;
      const int i = q >> 4, k4 = (q & 15) << 2;
      va[u] = (r0 + i < m) ? __ldg(reinterpret_cast<const float4*>(A + (r0 + i) * lda + k4))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float4 vb[TT];
    if (nt) {
#pragma unroll
      for (int u = 0; u < TT; ++u) {
        const int q = tid + 256 * u;
        const int j = q >> 4, k4 = (q & 15) << 2;
        vb[u] = (c0 + j < n) ? __ldg(reinterpret_cast<const float4*>(B + (c0 + j) * ldb + k4))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int u = 0; u < TT; ++u) {
        const int q = tid + 256 * u;
        const int kk = q / (T / 4), j4 = (q % (T / 4)) << 2;
        // n and c0 are multiples of 64 here (k1 = (k+1)*64), so a float4 is all-in or all-out
        vb[u] = (c0 + j4 < n) ? __ldg(reinterpret_cast<const float4*>(B + kk * ldb + c0 + j4))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < TT; ++u) {
      const int q = tid + 256 * u;
      const int i = q >> 4, k4 = (q & 15) << 2;
      As[k4 + 0][i] = va[u].x;
      As[k4 + 1][i] = va[u].y;
      As[k4 + 2][i] = va[u].z;
      As[k4 + 3][i] = va[u].w;
    }
    if (nt) {
#pragma unroll
      for (int u = 0; u < TT; ++u) {
        const int q = tid + 256 * u;
        const int j = q >> 4, k4 = (q & 15) << 2;
        Bs[k4 + 0][j] = vb[u].x;
        Bs[k4 + 1][j] = vb[u].y;
        Bs[k4 + 2][j] = vb[u].z;
        Bs[k4 + 3][j] = vb[u].w;
      }
    } else {
#pragma unroll
      for (int u = 0; u < TT; ++u) {
        const int q = tid + 256 * u;
        const int kk = q / (T / 4), j4 = (q % (T / 4)) << 2;
        *reinterpret_cast<float4*>(&Bs[kk][j4]) = vb[u];
      }
    }
  }
  __syncthreads();
  const int ty = tid / 16, tx = tid % 16;
  float acc[TT][TT];
#pragma unroll
  for (int i = 0; i < TT; ++i)
#pragma unroll
    for (int j = 0; j < TT; ++j) acc[i][j] = 0.f;
#pragma unroll 4
  for (int kk = 0; kk < NB; ++kk) {
    float a[TT], b[TT];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * TT + 4 * v]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * TT + 4 * v]);
      a[4 * v] = av.x; a[4 * v + 1] = av.y; a[4 * v + 2] = av.z; a[4 * v + 3] = av.w;
      b[4 * v] = bv.x; b[4 * v + 1] = bv.y; b[4 * v + 2] = bv.z; b[4 * v + 3] = bv.w;
    }
#pragma unroll
    for (int i = 0; i < TT; ++i)
#pragma unroll
      for (int j = 0; j < TT; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
  // Read-modify-write of the TT x TT register tile: ALL old values are loaded (as float4) before
  // the first store - a load / store pair per element would expose one L2 round trip TT^2 times.
  // Row starts are 16 B aligned (ld and c0 are multiples of 64, tx * TT floats = 16 / 32 B).
  float4 old[TT][V];
#pragma unroll
  for (int i = 0; i < TT; ++i) {
    const int row = r0 + ty * TT + i;
    const int col = c0 + tx * TT;
    const bool live = beta != 0.f && row < m && col < n && !(lower && col > row);
    const float4* src = reinterpret_cast<const float4*>(C + row * ld + col);
#pragma unroll
    for (int v = 0; v < V; ++v)
      old[i][v] = (live && col + 4 * v < n) ? src[v] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < TT; ++i) {
    const int row = r0 + ty * TT + i;
    if (row >= m) continue;
    float* crow = C + row * ld;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float o[4] = {old[i][v].x, old[i][v].y, old[i][v].z, old[i][v].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = c0 + tx * TT + 4 * v + j;
        if (col >= n || (lower && col > row)) continue;
        if (diag_only ? (row >= NB || col >= NB) : (skip_diag && row < NB && col < NB)) continue;
        crow[col] = fmaf(alpha, acc[i][4 * v + j], beta * o[j]);
      }
    }
  }
}

constexpr int rank64_smem(int T) { return static_cast<int>(sizeof(float)) * NB * 2 * (T + kPad); }

// Tile edge for a rank-64 launch: 128 when the grid fills the GPU anyway, 64 when the launch is a latency-bound
// link of the chain (few tiles).
inline int pick_tile(long long tiles128_total) { return tiles128_total >= 2 * 148 ? 128 : 64; }

// grid.x of a rank-64 launch over an m x n region with tile edge T
inline int rank64_tiles(int m, int n, int T, bool lower_full) {
  const int tm = (m + T - 1) / T, tn = (n + T - 1) / T;
  return lower_full ? tm * (tm + 1) / 2 : tm * tn;
}

inline void launch_rank64(int m, int n, bool lower_full, int count, cudaStream_t stream, const CholProb* tab, int k,
                          int mode, int limit) {
  const bool diag_only = (mode & kDiagOnly) != 0;
  const int T = diag_only ? 64 : pick_tile(static_cast<long long>(rank64_tiles(m, n, 128, lower_full)) * count);
  const dim3 grid(diag_only ? 1 : rank64_tiles(m, n, T, lower_full), count);
  if (T == 128)
    launch_chained(rank64_kernel<128>, grid, dim3(256), static_cast<size_t>(rank64_smem(128)), stream, tab, k, mode, limit);
  else
    launch_chained(rank64_kernel<64>, grid, dim3(256), static_cast<size_t>(rank64_smem(64)), stream, tab, k, mode, limit);
  note_launch();
}

// ------------------------------------------------------------------ un-flip + transpose to output
// out[i][j] = (i >= j) ? X[d-1-j][d-1-i] : 0
__global__ void flip_out_kernel(const CholProb* __restrict__ tab) {
  const CholProb p = tab[blockIdx.z];
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  if (i0 >= p.d || j0 >= p.d) return;
  __shared__ float tr[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int j = j0 + ty + 8 * r, i = i0 + tx;  // i fastest -> X column index contiguous
    float v = 0.f;
    if (i < p.d && j < p.d && i >= j)
      v = p.X[static_cast<long long>(p.d - 1 - j) * p.dpad + (p.d - 1 - i)];
    tr[ty + 8 * r][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty + 8 * r, j = j0 + tx;
    if (i < p.d && j < p.d) p.out[static_cast<long long>(i) * p.d + j] = tr[tx][ty + 8 * r];
  }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int pad_dim(int d) { return (d + NB - 1) / NB * NB; }

constexpr int kOuter = 256;       // outer block width of the two-level algorithm (4 inner blocks)
constexpr int kTwoLevelMin = 2048;  // padded size from which the tensor-core trailing updates pay off

inline size_t staging_bytes(int max_pad, int count) {
  // per factor (the outer updates of different factors run concurrently on auxiliary streams) and per parity of
  // the outer block (the FAR part of one block's update is still reading its operands while the next block's
  // are staged): three operands (L21 for the Cholesky phase; L21 again and the transposed X block for the
  // inverse phase, which runs concurrently on its own stream), three bf16 parts each, [max_pad, kOuter]
  return max_pad >= kTwoLevelMin
             ? static_cast<size_t>(count) * 2 * 3 * 3 * align_up(static_cast<size_t>(max_pad) * kOuter * 2, 256)
             : 0;
}

// The triangular-inverse phase runs one step behind the Cholesky phase on a second stream: its step k needs
// block column k of C (final after the panel of Cholesky step k) and its own step k - 1, nothing else.
struct Pipeline {
  cudaStream_t side = nullptr;
  cudaStream_t side2 = nullptr;  // wide part of the in-block trailing updates (beside the next diagonal block)
  cudaStream_t cap = nullptr;  // origin stream of graph captures (the caller's stream may be the legacy default)
  cudaEvent_t fork = nullptr, join = nullptr, join2 = nullptr;
  // outer (tensor-core) updates of the factors of a batch: up to kAux streams per phase, forked from / joined into
  // the phase's stream, so that the partial waves of one factor's GEMM are filled by another factor's
  static constexpr int kAux = 4;
  cudaStream_t aux[2][kAux] = {};
  cudaEvent_t aux_fork[2] = {}, aux_join[2][kAux] = {};
  std::vector<cudaEvent_t> panel_done, trail_done;
  bool ok = false;
  explicit Pipeline(bool) {}  // inert instance (no device)
  Pipeline() {
    ok = cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) == cudaSuccess &&
         cudaStreamCreateWithFlags(&side2, cudaStreamNonBlocking) == cudaSuccess &&
         cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&join, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&join2, cudaEventDisableTiming) == cudaSuccess;
    for (int ph = 0; ph < 2 && ok; ++ph) {
      ok = cudaEventCreateWithFlags(&aux_fork[ph], cudaEventDisableTiming) == cudaSuccess;
      for (int i = 0; i < kAux && ok; ++i)
        ok = cudaStreamCreateWithFlags(&aux[ph][i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&aux_join[ph][i], cudaEventDisableTiming) == cudaSuccess;
    }
  }
  bool reserve(int n) {
    while (ok && static_cast<int>(panel_done.size()) < n) {
      cudaEvent_t e;
      ok = cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
      if (ok) panel_done.push_back(e);
    }
    while (ok && static_cast<int>(trail_done.size()) < n) {
      cudaEvent_t e;
      ok = cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
      if (ok) trail_done.push_back(e);
    }
    return ok;
  }
};

// one side stream + event set per device (created on first use with that device current)
Pipeline& device_pipeline() {
  static std::mutex mu;
  static Pipeline* pipes[kMaxDevices] = {};
  static Pipeline none{false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return none;
  std::lock_guard<std::mutex> g(mu);
  if (pipes[dev] == nullptr) pipes[dev] = new Pipeline();
  return *pipes[dev];
}

}  // namespace

size_t chol_inv_workspace_bytes(const int* dims, int count) {
  size_t bytes = align_up(sizeof(CholProb) * static_cast<size_t>(count), 256) +
                 align_up(sizeof(int) * static_cast<size_t>(count), 256);
  int max_pad = 0;
  for (int f = 0; f < count; ++f) {
    const size_t dp = pad_dim(dims[f]);
    bytes += align_up(dp * dp * 4, 256) * 2 + align_up((dp / NB) * NB * NB * 4, 256);
    if (static_cast<int>(dp) > max_pad) max_pad = static_cast<int>(dp);
  }
  return bytes + staging_bytes(max_pad, count);
}

namespace {

// One 256-wide outer block of factor p is final in R: hand its contribution to everything beyond the block to the
// tensor cores as fp32-class bf16x6 GEMMs (hi/lo/lo2 splits carry 24 mantissa bits, six accumulating passes; a
// plain bf16 or bf16x3 update would perturb the Schur complement by 1e-3 .. 1e-5 and the inverse by cond(R)
// times that).
//   phase 1:  R[c1:, c1:]  -= L21 L21^T          (SYRK, lower tiles only)
//   phase 2:  X[c1:, :c1]  -= L21 X[c0:c1, :c1]
// Each update is issued in two parts.  NEAR = what the inner steps of the NEXT outer block touch (phase 1: the
// next 256 columns; phase 2: the next 256 rows): it stays on the chain.  FAR = the rest: it only has to be
// complete before the outer block after next starts, so it runs beside the next block's (latency-bound, nearly
// empty) inner steps.  Both parts add into their target through the TMA reduction, so a FAR part and the next
// block's updates of the same region commute.
struct OuterStage {
  __nv_bfloat16* a;  // L21 as three bf16 parts, [m2, kOuter] each
  __nv_bfloat16* b;  // phase 2: X[c0:c1, :c1]^T as three parts, [c1, kOuter] each
  size_t part_stride;
};

int outer_stage(const CholProb& p, int c0, bool phase2, const OuterStage& st, cudaStream_t stream) {
  const int c1 = c0 + kOuter;
  if (c1 >= p.dpad) return 0;
  const int m2 = p.dpad - c1;
  const long long ld = p.dpad;
  int rc = launch_convert_split3(p.R + static_cast<long long>(c1) * ld + c0, ld, m2, kOuter, st.a,
                                 st.a + st.part_stride, st.a + 2 * st.part_stride, kOuter, stream);
  if (rc == 0 && phase2)
    rc = launch_transpose_split3(p.X + static_cast<long long>(c0) * ld, ld, kOuter, c1, st.b, st.b + st.part_stride,
                                 st.b + 2 * st.part_stride, kOuter, stream);
  return rc;
}

int g_chol_far_sms = 64;  // SMs all concurrently running FAR parts of one phase may occupy together

int outer_gemm(const CholProb& p, int c0, bool phase2, bool far, const OuterStage& st, int concurrent,
               cudaStream_t stream) {
  const int c1 = c0 + kOuter;
  if (c1 >= p.dpad) return 0;
  const int m2 = p.dpad - c1;
  const int near_rows = m2 < kOuter ? m2 : kOuter;
  if (far && m2 <= kOuter) return 0;
  const long long ld = p.dpad;
  const long long skip = far ? static_cast<long long>(kOuter) * kOuter : 0;  // FAR: rows from kOuter of L21 on
  GemmArgs g;
  g.A_hi = st.a + skip;
  g.A_lo = st.a + st.part_stride + skip;
  g.A_lo2 = st.a + 2 * st.part_stride + skip;
  g.lda = kOuter;
  g.K = kOuter;
  g.batch = 1;
  g.nparts = 6;
  g.alpha = -1.f;
  g.beta = 1.f;
  g.ldc = ld;
  g.ldb = kOuter;
  // a FAR part is background work: it must leave SMs to the chain kernels it runs beside (a CTA of the
  // contraction core owns its SM's shared memory; the chain would queue behind whole tiles otherwise)
  if (far && g_chol_far_sms > 0) g.max_sms = (g_chol_far_sms / (2 * concurrent)) * 2;
  if (!phase2) {
    if (far) {  // R[c1 + 256:, c1 + 256:] -= L21' L21'^T, lower tiles
      g.B_hi = g.A_hi;
      g.B_lo = g.A_lo;
      g.B_lo2 = g.A_lo2;
      g.M = g.N = m2 - kOuter;
      g.flags = kSyrkLower;
      g.C = p.R + static_cast<long long>(c1 + kOuter) * ld + (c1 + kOuter);
    } else {  // R[c1:, c1 : c1 + 256] -= L21 L21[:256]^T (the upper part of the first tile is never read)
      g.B_hi = st.a;
      g.B_lo = st.a + st.part_stride;
      g.B_lo2 = st.a + 2 * st.part_stride;
      g.M = m2;
      g.N = near_rows;
      g.flags = 0;
      g.C = p.R + static_cast<long long>(c1) * ld + c1;
    }
  } else {
    g.B_hi = st.b;
    g.B_lo = st.b + st.part_stride;
    g.B_lo2 = st.b + 2 * st.part_stride;
    g.N = c1;
    g.flags = 0;
    g.M = far ? m2 - kOuter : near_rows;
    g.C = p.X + static_cast<long long>(far ? c1 + kOuter : c1) * ld;
  }
  return launch_umma_gemm(g, stream);
}

}  // namespace

namespace {

int g_chol_graph = 1;
int g_chol_lookahead = 1;

struct GraphKey {
  void* workspace;
  int dev, far_sms;  // far_sms also carries the look-ahead switch (bit 30)
  std::vector<int> dims;
  bool operator==(const GraphKey& o) const {
    return workspace == o.workspace && dev == o.dev && far_sms == o.far_sms && dims == o.dims;
  }
};
struct CachedGraph {
  GraphKey key;
  cudaGraphExec_t exec = nullptr;
  int launches = 0;
};
std::mutex g_graph_mu;
std::vector<CachedGraph> g_graphs;

// The whole step sequence (damp / flip, both phases, flip-out) enqueued on `stream` (+ the pipeline's side stream,
// forked from and joined back into it).  Depends on the workspace layout and the dims only.
int enqueue_steps(const std::vector<CholProb>& h_tab2, int count, int max_pad, CholProb* d_tab, int* d_info,
                  __nv_bfloat16* stage_a, Pipeline& pipe, bool pipelined, int /*smem*/, cudaStream_t stream) {
  const int max_nb = max_pad / NB;
  const CholProb* cd_tab = d_tab;
  const dim3 tb(32, 8);
  const dim3 tg((max_pad + 31) / 32, (max_pad + 31) / 32, count);
  damp_flip_kernel<<<tg, tb, 0, stream>>>(d_tab);
  note_launch();
  // Two-level blocking for wide factors: rank-64 SIMT updates stay inside the current 256-wide outer
  // block; the update of everything beyond it is ONE fp32-class (bf16x6) tensor-core GEMM per factor.
  const bool two_level = max_pad >= kTwoLevelMin && pipelined;
  const int inner_per_outer = kOuter / NB;
  const size_t part_stride = align_up(static_cast<size_t>(max_pad) * kOuter * 2, 256) / 2;
  // all factors' outer updates for the 256-wide block at c0, issued from stream `s` of phase `ph` (0: Cholesky,
  // 1: inverse).  Staging + NEAR parts: on the auxiliary streams of the phase (factor f on stream f mod nst),
  // joined back into `s`.  FAR parts: queued behind them on the same auxiliary streams and NOT joined - the next
  // boundary's staging / NEAR parts line up behind them in stream order, which is exactly the dependency the
  // outer block after next has on them.
  auto outer_all = [&](int c0, int ph, cudaStream_t s) -> int {
    const int nst = count < Pipeline::kAux ? count : Pipeline::kAux;
    const int parity = (c0 / kOuter) & 1;
    if (cudaEventRecord(pipe.aux_fork[ph], s) != cudaSuccess) return -5;
    for (int i = 0; i < nst; ++i)
      if (cudaStreamWaitEvent(pipe.aux[ph][i], pipe.aux_fork[ph], 0) != cudaSuccess) return -5;
    auto stage_of = [&](int f) {
      // per factor: [parity][ 0..3: L21 (Cholesky phase) | 3..6: X^T block | 6..9: L21 (inverse phase) ]
      __nv_bfloat16* base = stage_a + (static_cast<size_t>(f) * 2 + parity) * 9 * part_stride;
      return OuterStage{ph == 1 ? base + 6 * part_stride : base, base + 3 * part_stride, part_stride};
    };
    for (int f = 0; f < count; ++f) {
      const OuterStage st = stage_of(f);
      int rc = outer_stage(h_tab2[f], c0, ph == 1, st, pipe.aux[ph][f % nst]);
      if (rc == 0) rc = outer_gemm(h_tab2[f], c0, ph == 1, false, st, nst, pipe.aux[ph][f % nst]);
      if (rc) return rc;
    }
    for (int i = 0; i < nst; ++i)
      if (cudaEventRecord(pipe.aux_join[ph][i], pipe.aux[ph][i]) != cudaSuccess ||
          cudaStreamWaitEvent(s, pipe.aux_join[ph][i], 0) != cudaSuccess)
        return -5;
    for (int f = 0; f < count; ++f) {
      const int rc = outer_gemm(h_tab2[f], c0, ph == 1, true, stage_of(f), nst, pipe.aux[ph][f % nst]);
      if (rc) return rc;
    }
    return 0;
  };
  // the FAR parts still in flight on the auxiliary streams of phase `ph` must land before `s` continues
  auto outer_drain = [&](int ph, cudaStream_t s) -> int {
    const int nst = count < Pipeline::kAux ? count : Pipeline::kAux;
    for (int i = 0; i < nst; ++i)
      if (cudaEventRecord(pipe.aux_join[ph][i], pipe.aux[ph][i]) != cudaSuccess ||
          cudaStreamWaitEvent(s, pipe.aux_join[ph][i], 0) != cudaSuccess)
        return -5;
    return 0;
  };
  // Wide problems: the inverse phase (X = C^-1 by block forward substitution on the identity) is pipelined one
  // step behind the Cholesky phase on a side stream; both are chains of latency-bound steps that leave most
  // of the GPU idle on their own.
  cudaStream_t s2 = pipelined ? pipe.side : stream;
  // Look-ahead inside the inner loop: of the trailing update of step k only the NEXT diagonal block sits on the chain
  // (one CTA); the rest runs on a third stream beside the next diagonal-block kernel and is awaited by panel k + 1.
  const bool lookahead = pipelined && g_chol_lookahead != 0;
  cudaStream_t s3 = lookahead ? pipe.side2 : stream;
  std::vector<char> trail_issued(static_cast<size_t>(max_nb), 0);
  if (pipelined) {
    if (cudaEventRecord(pipe.fork, stream) != cudaSuccess || cudaStreamWaitEvent(s2, pipe.fork, 0) != cudaSuccess)
      return -5;
    if (lookahead && cudaStreamWaitEvent(s3, pipe.fork, 0) != cudaSuccess) return -5;
  }
  // one step of the inverse phase (enqueued on s2)
  auto inverse_step = [&](int k) -> int {
    const int limit = two_level ? (k / inner_per_outer + 1) * kOuter : 0;
    const int n = (k + 1) * NB;
    launch_rank64(NB, n, false, count, s2, cd_tab, k, static_cast<int>(kRowScale), 0);
    int m = max_pad - (k + 1) * NB;
    if (two_level) m = (limit < max_pad ? limit : max_pad) - (k + 1) * NB;
    if (m > 0) launch_rank64(m, n, false, count, s2, cd_tab, k, static_cast<int>(kXUpdate), limit);
    if (two_level && (k + 1) % inner_per_outer == 0) {
      const int rc = outer_all((k + 1 - inner_per_outer) * NB, 1, s2);
      if (rc) return rc;
    }
    return 0;
  };
  // ---- phase 1: right-looking Cholesky of the flipped damped matrix (+ the pipelined inverse steps)
  for (int k = 0; k < max_nb; ++k) {
    const int limit = two_level ? (k / inner_per_outer + 1) * kOuter : 0;
    // look-ahead: step k - 1 left the update of this diagonal block to this kernel's prologue
    const int fuse_prev = (lookahead && k >= 1 && (!two_level || k % inner_per_outer != 0)) ? 1 : 0;
    launch_chained(potrf_diag_kernel, dim3(count), dim3(kDiagThreads), 0, stream, cd_tab, k, d_info, fuse_prev);
    note_launch();
    const int m = max_pad - (k + 1) * NB;
    if (m > 0) {
      // the panel reads block column k below the diagonal block: the wide trailing update of step k - 1 wrote it
      if (lookahead && k >= 1 && trail_issued[k - 1] &&
          cudaStreamWaitEvent(stream, pipe.trail_done[k - 1], 0) != cudaSuccess)
        return -5;
      launch_rank64(m, NB, false, count, stream, cd_tab, k, static_cast<int>(kPanel), 0);
    }
    if (pipelined) {
      // block column k of C and Dinv[k] are final: inverse step k may run
      if (cudaEventRecord(pipe.panel_done[k], stream) != cudaSuccess ||
          cudaStreamWaitEvent(s2, pipe.panel_done[k], 0) != cudaSuccess)
        return -5;
      const int rc = inverse_step(k);
      if (rc) return rc;
    }
    if (m > 0) {
      const int ncols = two_level ? limit - (k + 1) * NB : m;  // two-level: columns of the outer block right of k
      if (ncols > 0) {
        const int lim = two_level ? limit : 0;
        const bool full_lower = !two_level;
        if (!lookahead) {
          launch_rank64(m, ncols, full_lower, count, stream, cd_tab, k, static_cast<int>(kTrail), lim);
        } else {
          // (the next diagonal block's share of this update is fused into the next diagonal-block kernel)
          if (m > NB) {
            if (cudaStreamWaitEvent(s3, pipe.panel_done[k], 0) != cudaSuccess) return -5;
            launch_rank64(m, ncols, full_lower, count, s3, cd_tab, k, static_cast<int>(kTrail) | kSkipDiag, lim);
            if (cudaEventRecord(pipe.trail_done[k], s3) != cudaSuccess) return -5;
            trail_issued[k] = 1;
          }
        }
      }
    }
    if (two_level && (k + 1) % inner_per_outer == 0) {
      const int rc = outer_all((k + 1 - inner_per_outer) * NB, 0, stream);
      if (rc) return rc;
    }
  }
  // ---- phase 2 (not pipelined: small problems): X = C^-1 by block forward substitution on the identity
  if (!pipelined) {
    for (int k = 0; k < max_nb; ++k) {
      const int rc = inverse_step(k);
      if (rc) return rc;
    }
  } else {
    if (two_level && (outer_drain(0, stream) != 0 || outer_drain(1, s2) != 0)) return -5;
    if (lookahead && (cudaEventRecord(pipe.join2, s3) != cudaSuccess ||
                      cudaStreamWaitEvent(stream, pipe.join2, 0) != cudaSuccess))
      return -5;
    if (cudaEventRecord(pipe.join, s2) != cudaSuccess || cudaStreamWaitEvent(stream, pipe.join, 0) != cudaSuccess)
      return -5;
  }
  flip_out_kernel<<<tg, tb, 0, stream>>>(d_tab);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace

void set_chol_graph(int enabled) { g_chol_graph = enabled; }
void set_chol_far_sms(int sms) { g_chol_far_sms = sms < 0 ? g_chol_far_sms : sms; }
void set_chol_lookahead(int on) { g_chol_lookahead = on; }

int chol_inv_batched(const float* const* factors, float* const* outs, const int* dims,
                     const float* add, const float* multiply, int count, void* workspace,
                     size_t workspace_bytes, cudaStream_t stream) {
  if (count <= 0) return 0;
  if (workspace == nullptr || workspace_bytes < chol_inv_workspace_bytes(dims, count) ||
      (reinterpret_cast<uintptr_t>(workspace) & 255) != 0)
    return -6;
  static DeviceOnce attr_once;
  const int smem = rank64_smem(128);
  if (!attr_once([&] {
        return cudaFuncSetAttribute(rank64_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    rank64_smem(128)) == cudaSuccess &&
               cudaFuncSetAttribute(rank64_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    rank64_smem(64)) == cudaSuccess;
      }))
    return -5;
  char* w = static_cast<char*>(workspace);
  CholProb* d_tab = reinterpret_cast<CholProb*>(w);
  w += align_up(sizeof(CholProb) * static_cast<size_t>(count), 256);
  int* d_info = reinterpret_cast<int*>(w);
  w += align_up(sizeof(int) * static_cast<size_t>(count), 256);
  CholProb* h_tab = new CholProb[count];
  int max_pad = 0;
  for (int f = 0; f < count; ++f) {
    if (dims[f] <= 0 || factors[f] == nullptr || outs[f] == nullptr || add[f] < 0.f ||
        multiply[f] < 0.f) {
      delete[] h_tab;
      return -2;
    }
    CholProb& p = h_tab[f];
    p.F = factors[f];
    p.out = outs[f];
    p.d = dims[f];
    p.dpad = pad_dim(dims[f]);
    p.nb = p.dpad / NB;
    p.sqrt_s = sqrtf(multiply[f]);
    p.sqrt_n = sqrtf(add[f]);
    const size_t mat = align_up(static_cast<size_t>(p.dpad) * p.dpad * 4, 256);
    p.R = reinterpret_cast<float*>(w);
    w += mat;
    p.X = reinterpret_cast<float*>(w);
    w += mat;
    p.Dinv = reinterpret_cast<float*>(w);
    w += align_up(static_cast<size_t>(p.nb) * NB * NB * 4, 256);
    if (p.dpad > max_pad) max_pad = p.dpad;
  }
  cudaError_t e = cudaMemcpyAsync(d_tab, h_tab, sizeof(CholProb) * count, cudaMemcpyHostToDevice,
                                  stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_info, 0, sizeof(int) * count, stream);
  // the table copy is from pageable memory: it has been staged when the call returns
  std::vector<CholProb> h_tab2(h_tab, h_tab + count);
  delete[] h_tab;
  if (e != cudaSuccess) return -5;

  Pipeline& pipe = device_pipeline();
  const int max_nb = max_pad / NB;
  const bool pipelined = max_nb >= 4 && pipe.ok && pipe.reserve(max_nb);
  __nv_bfloat16* stage_a = reinterpret_cast<__nv_bfloat16*>(w);
  // Every kernel argument below is derived from the workspace address and the dims (factor / output pointers and
  // the damping scalars are read from the device table just refreshed): the sequence is captured once per
  // (workspace, dims) into a CUDA graph and replayed - one graph launch instead of ~450 (one 4097-wide factor) to
  // ~1200 (the 8 factors of the wide MLP) host-issued dependent launches on two streams.
  int dev = 0;
  cudaGetDevice(&dev);
  const bool use_graph = g_chol_graph != 0 && pipe.ok && max_nb >= 4;
  if (use_graph) {
    GraphKey key{workspace, dev, g_chol_far_sms | (g_chol_lookahead ? 1 << 30 : 0), std::vector<int>(dims, dims + count)};
    std::lock_guard<std::mutex> guard(g_graph_mu);
    CachedGraph* hit = nullptr;
    for (auto& c : g_graphs)
      if (c.key == key) hit = &c;
    if (hit == nullptr) {
      CachedGraph fresh;
      fresh.key = key;
      const unsigned long long before = launch_count();
      bool ok = cudaStreamBeginCapture(pipe.cap, cudaStreamCaptureModeRelaxed) == cudaSuccess;
      int rc_enq = -5;
      if (ok) {
        rc_enq = enqueue_steps(h_tab2, count, max_pad, d_tab, d_info, stage_a, pipe, pipelined, smem, pipe.cap);
        cudaGraph_t graph = nullptr;
        ok = cudaStreamEndCapture(pipe.cap, &graph) == cudaSuccess && rc_enq == 0 && graph != nullptr;
        if (ok) ok = cudaGraphInstantiate(&fresh.exec, graph, 0) == cudaSuccess;
        if (graph != nullptr) cudaGraphDestroy(graph);
      }
      fresh.launches = static_cast<int>(launch_count() - before);
      note_launch(-fresh.launches);  // counted per replay below
      if (ok) {
        if (g_graphs.size() >= 16) {
          for (auto& c : g_graphs) cudaGraphExecDestroy(c.exec);
          g_graphs.clear();
        }
        g_graphs.push_back(fresh);
        hit = &g_graphs.back();
      } else {
        cudaGetLastError();  // capture unsupported here: fall through to the direct path
        if (rc_enq < 0 && rc_enq != -5) return rc_enq;
      }
    }
    if (hit != nullptr) {
      if (cudaGraphLaunch(hit->exec, stream) != cudaSuccess) return -5;
      note_launch(hit->launches);
    } else {
      const int rc_enq = enqueue_steps(h_tab2, count, max_pad, d_tab, d_info, stage_a, pipe, pipelined, smem, stream);
      if (rc_enq) return rc_enq;
    }
  } else {
    const int rc_enq = enqueue_steps(h_tab2, count, max_pad, d_tab, d_info, stage_a, pipe, pipelined, smem, stream);
    if (rc_enq) return rc_enq;
  }
  if (cudaGetLastError() != cudaSuccess) return -5;

  int* h_info = new int[count];
  e = cudaMemcpyAsync(h_info, d_info, sizeof(int) * count, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  int rc = 0;
  if (e != cudaSuccess) {
    rc = -5;
  } else {
    for (int f = 0; f < count; ++f)
      if (h_info[f] != 0) {
        rc = f + 1;
        break;
      }
  }
  delete[] h_info;
  return rc;
}

}  // namespace bk
