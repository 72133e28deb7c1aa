// bk_eigh_blocked.cu — symmetric eigendecomposition of WIDE Kronecker factors (d > 164) by a two-sided
// BLOCK Jacobi method whose O(d^3) work runs on the tensor cores.
//
// Reference: models/utilities.py:144-159 / :120-141 (torch.symeig of F + F^T / of F; LAPACK syevd on the
// CPU, cuSOLVER on CUDA).  The element-wise one-sided Jacobi of bk_eigh.cu streams the whole matrix once
// per round of d/2 rotations (d - 1 launches per sweep): L2/latency-bound, 283 ms at d = 2049.  Here the
// rotations are aggregated into 128 x 128 orthogonal blocks and applied as GEMMs:
//
//   S (d_pad x d_pad, symmetric) is cut into nb blocks of b = 64 rows / columns (d padded to a multiple of
//   128; the pad carries a diagonal M > |lambda|_max that never couples, see init).  One ROUND pairs
//   adjacent blocks — (0,1)(2,3).. in even rounds, (1,2)(3,4).. in odd rounds, the two blocks of a pair
//   SWAPPING places afterwards (odd-even transposition ordering: after nb rounds every pair of blocks has
//   met exactly once = one sweep).  Because partners are always adjacent, every pair is a contiguous
//   128-wide slab at a uniform stride and a whole round is three BATCHED GEMMs:
//     inner   per pair: two-sided Jacobi sweep(s) on the 128 x 128 diagonal block G = S[pair, pair] in
//             shared memory (rotation angles straight from G: no dot products), accumulating Q^T
//     step 1  T[pair rows, :]  = Q^T S[pair rows, :]     (B operand = S[:, pair cols] by symmetry)
//     step 2  S[:, pair cols]  = T[:, pair cols] Q
//     step 3  V[:, pair cols]  = V[:, pair cols] Q       (eigenvector accumulation, columns = vectors)
//   All three are K-major "NT" contractions of the tcgen05 core; operands live as three-way bf16 splits
//   (hi + lo + lo2 = 24 mantissa bits, six MMA passes = fp32-class products) that the GEMM epilogue emits
//   directly, so no separate staging pass exists.  Per round the matrix is read and written a constant
//   number of times (HBM-bound, ~0.5 GB at d = 4097) instead of once per 2-column rotation.
// Convergence: a round-robin pass of the inner kernel applies a rotation only where
// |g_pq| > tol sqrt(|g_pp g_qq|) and |g_pq| > max(5e-7, 1.2e-7 sqrt(d)) |S|_F (rounding noise; null space of
// rank-deficient factors); the iteration ends when the off-norm estimate of a sweep stops falling (rounding
// floor reached) or a sweep rotated nothing.
#include "bk_common.cuh"
#include "bk_kernels.cuh"
#include "bk_umma_gemm.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

namespace bk {

namespace {

struct Parts {
  __nv_bfloat16* p[3];
};

// ------------------------------------------------------------------------------ init
// S_pad = [[sym_scale (F + F^T), 0], [0, 0]] (pad diagonal set by pad_diag_kernel); V_pad = I and
// fro2 += |S|_F^2 when V / fro2 are given (the final Rayleigh pass re-creates S only).
__global__ void blk_init_kernel(const float* __restrict__ F, long long ldf, int d, int dp, float sym_scale,
                                float* __restrict__ S, float* __restrict__ V, float* __restrict__ fro2) {
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  __shared__ float tr[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;  // (32, 8)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int j = j0 + ty + 8 * k, i = i0 + tx;  // F[j][i], i fastest
    tr[ty + 8 * k][tx] = (i < d && j < d) ? F[static_cast<long long>(j) * ldf + i] : 0.f;
  }
  __syncthreads();
  float fro = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = i0 + ty + 8 * k, j = j0 + tx;
    if (i < dp && j < dp) {
      float v = 0.f;
      if (i < d && j < d) v = sym_scale * (F[static_cast<long long>(i) * ldf + j] + tr[tx][ty + 8 * k]);
      S[static_cast<long long>(i) * dp + j] = v;
      if (V != nullptr) V[static_cast<long long>(i) * dp + j] = (i == j) ? 1.f : 0.f;
      fro = fmaf(v, v, fro);
    }
  }
  if (fro2 != nullptr) {
    fro = warp_sum(fro);
    if (tx == 0 && fro != 0.f) atomicAdd(fro2, fro);
  }
}

// pad diagonal M = 2 |S|_F (1 if S == 0): larger than every |lambda|, so the pad eigenpairs sort last;
// S[real, pad] is exactly 0 and every product with it stays exactly 0, so the pad never mixes in.
__global__ void blk_pad_diag_kernel(float* __restrict__ S, int d, int dp, const float* __restrict__ fro2) {
  const int i = d + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dp) return;
  const float f = sqrtf(*fro2);
  S[static_cast<long long>(i) * dp + i] = f > 0.f ? 2.f * f : 1.f;
}

__global__ void blk_identity_kernel(int b, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                    __nv_bfloat16* __restrict__ lo2) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= b * b) return;
  hi[idx] = __float2bfloat16_rn((idx / b == idx % b) ? 1.f : 0.f);
  lo[idx] = __float2bfloat16_rn(0.f);
  lo2[idx] = __float2bfloat16_rn(0.f);
}

// ------------------------------------------------------------------------------ inner two-sided Jacobi
// Circle-method round robin on P positions laid out [t_0 .. t_{h-1} | b_0 .. b_{h-1}] (h = P / 2): the pairs of
// a round are always (t_k, b_k) = positions (k, h + k); between rounds every position but t_0 moves one step
// along the ring t_1, .., t_{h-1}, b_{h-1}, .., b_0 (-> t_1).  ring_shift(x, m) = position of x after m steps.
template <int P>
__device__ __forceinline__ int ring_shift(int x, int m) {
  constexpr int h = P / 2, R = P - 1;
  if (x == 0) return 0;
  int idx = x < h ? x - 1 : 3 * h - 2 - x;
  idx = (idx + m) % R;
  if (idx < 0) idx += R;
  return idx <= h - 2 ? idx + 1 : 3 * h - 2 - idx;
}

// One CTA per pair of blocks (P = pair width = 2 b).  G = sym(S[o:o+P, o:o+P]) rebuilt exactly from the three
// bf16 splits of S, Q^T = I; `sweeps` cyclic two-sided Jacobi sweeps; writes Q^T (rows in final position order,
// halves exchanged: the blocks of the pair swap places) as bf16 splits.  stats[0] |= 1 if any rotation was
// applied, stats[1] = largest eliminated off-diagonal |g_pq| / |S|_F (as float bits).
//
// Shared-memory traffic is what bounds this kernel (one SM per pair), so a round touches every element once:
//   * the two-sided update J^T G J decomposes into independent 2 x 2 blocks, one per (row pair a, column pair b):
//     thread (a, b) loads G[{a, h+a}][{b, h+b}], applies J_a^T . J_b in registers and stores the block once —
//     no separate row and column passes;
//   * pairs never move, the DATA does: the block is stored into the second G buffer at the positions of the
//     next round (ring_shift(., +1)), so loads and stores are consecutive across the lanes of a warp (lanes walk
//     b) and conflict-free apart from the two wrap points of the ring;
//   * Q^T is only rotated in place (rows); the position -> Q^T-row map moves instead of the rows;
//   * a round without any rotation above the threshold (most rounds of the late sweeps) costs one barrier: the
//     shift is deferred (`pending`) and folded into the addresses of the next active round.
template <int P>
__global__ void __launch_bounds__(P * 8, 1)
blk_inner_kernel(Parts ss, int dp, int off, float tol, float floor_rel, const float* __restrict__ fro2,
                 int sweeps, Parts qt, int* __restrict__ stats) {
  constexpr int kLd = P + 1, h = P / 2, kThreads = P * 8;
  constexpr int kShift = P == 128 ? 7 : 6;
  constexpr int kWarps = kThreads / 32;          // row pairs per pass
  constexpr int kPasses = h / kWarps;            // 2
  static_assert(P == 128 || P == 64, "pair width");
  extern __shared__ float sm[];
  float* const G0 = sm;  // buffer `cur` is G0 + cur * P * kLd
  float* Q = sm + 2 * P * kLd;
  __shared__ float cs_c[h], cs_s[h];
  __shared__ int qrow[2][P];  // physical position -> row of Q holding that position's vector
  // position tables of the round (the ring arithmetic costs ~15 integer instructions per call and the kernel
  // is as much issue- as bandwidth-bound): src[x] = where the data of logical position x sits,
  // dst[x] = where it goes (the logical position of the next round)
  __shared__ int src[P], dst[P];
  __shared__ int s_any, s_flag[2];  // s_flag[r & 1]: a rotation of round r is above the threshold
  __shared__ float s_max, s_asym, s_off2;
  const int tid = threadIdx.x;
  const int o = off + blockIdx.x * P;
  const long long sb = static_cast<long long>(o) * dp + o;
  for (int idx = tid; idx < P * P; idx += kThreads) {
    const int i = idx >> kShift, j = idx & (P - 1);
    const long long a = sb + static_cast<long long>(i) * dp + j;
    G0[i * kLd + j] = __bfloat162float(ss.p[0][a]) + __bfloat162float(ss.p[1][a]) + __bfloat162float(ss.p[2][a]);
    Q[i * kLd + j] = (i == j) ? 1.f : 0.f;
  }
  if (tid < P) {
    qrow[0][tid] = tid;
    dst[tid] = ring_shift<P>(tid, 1);
  }
  if (tid == 0) {
    s_any = 0;
    s_max = 0.f;
    s_asym = 0.f;
    s_off2 = 0.f;
    s_flag[0] = s_flag[1] = 0;
  }
  __syncthreads();
  // symmetrise (S is symmetric up to the rounding of two different summation orders); the largest asymmetry
  // is a direct measurement of the rounding noise the slab products leave on the elements of this block
  float asym = 0.f;
  for (int idx = tid; idx < P * P; idx += kThreads) {
    const int i = idx >> kShift, j = idx & (P - 1);
    if (j < i) {
      const float x = G0[i * kLd + j], y = G0[j * kLd + i];
      const float v = 0.5f * (x + y);
      asym = fmaxf(asym, fabsf(x - y));
      G0[i * kLd + j] = v;
      G0[j * kLd + i] = v;
    }
  }
  {
    const int m = __reduce_max_sync(0xffffffffu, __float_as_int(asym));
    if ((tid & 31) == 0 && m != 0) atomicMax(reinterpret_cast<int*>(&s_asym), m);
  }
  __syncthreads();
  const float abs_floor = floor_rel * sqrtf(*fro2);
  const int warp = tid >> 5, lane = tid & 31;
  // this lane's column pairs b = lane + 32 k never change: their destination positions live in registers, and
  // so do the source positions as long as no shift is pending (the common case while rotations are dense)
  int dcp_r[h / 32], dcq_r[h / 32];
#pragma unroll
  for (int k = 0; k < h / 32; ++k) {
    dcp_r[k] = ring_shift<P>(lane + 32 * k, 1);
    dcq_r[k] = ring_shift<P>(h + lane + 32 * k, 1);
  }
  int cur = 0;      // G buffer / qrow map holding the current data
  int pending = 0;  // ring steps taken by the schedule since the data was last moved (block-uniform)
  const int rounds = sweeps * (P - 1);
  for (int r = 0; r < rounds; ++r) {
    const float* G = G0 + cur * (P * kLd);
    __syncthreads();  // the data of the previous round has landed
    if (tid < h) {  // whole warps (h = 32 or 64)
      // logical pair (tid, h + tid) sits at the physical positions shifted back by `pending`
      const int p = pending ? ring_shift<P>(tid, -pending) : tid;
      const int q = pending ? ring_shift<P>(h + tid, -pending) : h + tid;
      const float gpp = G[p * kLd + p], gqq = G[q * kLd + q], gpq = G[p * kLd + q];
      float c = 1.f, s = 0.f;
      const float av = fabsf(gpq);
      const bool rot = av > tol * sqrtf(fabsf(gpp * gqq)) && av > abs_floor;
      if (rot) {
        const float zeta = (gqq - gpp) / (2.f * gpq);
        const float t = copysignf(1.f, zeta) / (fabsf(zeta) + sqrtf(1.f + zeta * zeta));
        c = 1.f / sqrtf(1.f + t * t);
        s = c * t;
      }
      cs_c[tid] = c;
      cs_s[tid] = s;
      // one shared-memory atomic per warp, not per rotation (64 serialised atomics per round otherwise)
      const int m = __reduce_max_sync(0xffffffffu, rot ? __float_as_int(av) : 0);
      const float sq = warp_sum(av * av);  // every off-diagonal of the pair block is inspected once per sweep
      if (lane == 0) {
        atomicAdd(&s_off2, sq);
        if (m != 0) {
          s_flag[r & 1] = 1;
          atomicMax(reinterpret_cast<int*>(&s_max), m);
        }
      }
    }
    if (tid < P) src[tid] = ring_shift<P>(tid, -pending);
    __syncthreads();
    // the flag of the NEXT round is cleared here: nobody reads it before that round's second barrier, and
    // nobody writes this round's flag again before the second barrier of the round after next
    if (tid == 0) s_flag[(r + 1) & 1] = 0;
    if (s_flag[r & 1] == 0) {  // nothing to rotate: only the schedule advances (block-uniform branch)
      ++pending;
      continue;
    }
    if (tid == 0) s_any = 1;
    float* Gn = G0 + (cur ^ 1) * (P * kLd);
    // a warp owns ONE row pair per pass and its lanes walk 32 consecutive column pairs: with the odd pitch
    // every load and store below is bank-conflict free (two row pairs per warp collide on 15 of 16 banks)
#pragma unroll
    for (int pass = 0; pass < kPasses; ++pass) {
      const int a = warp + pass * kWarps;
      const float ca = cs_c[a], sa = cs_s[a];
      const int rp = src[a], rq = src[h + a];  // source rows
      const int np = dst[a], nq = dst[h + a];  // destination rows
      // G: 2 x 2 blocks  [x00 x01; x10 x11] = G[{rp, rq}][{cp, cq}]  ->  J_a^T [..] J_b at the next positions
#pragma unroll
      for (int k = 0; k < h / 32; ++k) {
        const int b = lane + 32 * k;
        const float cb = cs_c[b], sbv = cs_s[b];
        const int cp = pending ? src[b] : b, cq = pending ? src[h + b] : h + b;
        const float x00 = G[rp * kLd + cp], x01 = G[rp * kLd + cq];
        const float x10 = G[rq * kLd + cp], x11 = G[rq * kLd + cq];
        // rows: [t0; t1] = J_a^T [x0; x1]  (row_p' = c row_p - s row_q, row_q' = s row_p + c row_q)
        const float t00 = ca * x00 - sa * x10, t01 = ca * x01 - sa * x11;
        const float t10 = sa * x00 + ca * x10, t11 = sa * x01 + ca * x11;
        // columns: [y_p y_q] = [t_p t_q] J_b  (col_p' = c col_p - s col_q, col_q' = s col_p + c col_q)
        const int dcp = dcp_r[k], dcq = dcq_r[k];
        Gn[np * kLd + dcp] = cb * t00 - sbv * t01;
        Gn[np * kLd + dcq] = sbv * t00 + cb * t01;
        Gn[nq * kLd + dcp] = cb * t10 - sbv * t11;
        Gn[nq * kLd + dcq] = sbv * t10 + cb * t11;
      }
      // Q^T rows of the pair, in place; the position -> row map moves with the positions
      const int q0 = qrow[cur][rp], q1 = qrow[cur][rq];
      if (lane == 0) {
        qrow[cur ^ 1][np] = q0;
        qrow[cur ^ 1][nq] = q1;
      }
      if (sa != 0.f) {
#pragma unroll
        for (int k = 0; k < P / 32; ++k) {
          const int j = lane + 32 * k;
          const float u = Q[q0 * kLd + j], v = Q[q1 * kLd + j];
          Q[q0 * kLd + j] = ca * u - sa * v;
          Q[q1 * kLd + j] = sa * u + ca * v;
        }
      }
    }
    cur ^= 1;
    pending = 0;
  }
  __syncthreads();
  if (tid == 0) {
    const float inv = rsqrtf(fmaxf(*fro2, 1e-37f));
    if (s_any) {
      atomicOr(&stats[0], 1);
      atomicMax(&stats[1], __float_as_int(s_max * inv));
    }
    atomicMax(&stats[2], __float_as_int(s_asym * inv));
    atomicAdd(reinterpret_cast<float*>(&stats[3]), s_off2 * inv * inv);
  }
  // Q^T out: output row i = vector at physical position (i + P/2) mod P (halves exchanged)
  const long long base = static_cast<long long>(blockIdx.x) * P * P;
  for (int idx = tid; idx < P * P; idx += kThreads) {
    const int i = idx >> kShift, j = idx & (P - 1);
    const float v = Q[qrow[cur][(i + h) & (P - 1)] * kLd + j];
    const __nv_bfloat16 hh = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(hh);
    const __nv_bfloat16 l = __float2bfloat16_rn(r1);
    qt.p[0][base + idx] = hh;
    qt.p[1][base + idx] = l;
    qt.p[2][base + idx] = __float2bfloat16_rn(r1 - __bfloat162float(l));
  }
}

// ------------------------------------------------------------------------------ output
// num[c] += sum over this block's rows of P[i][c] * V[i][c], den[c] += sum V[i][c]^2
// (Rayleigh quotients v_c^T S v_c / v_c^T v_c with P = S V)
__global__ void blk_coldot_kernel(const float* __restrict__ Pm, const float* __restrict__ V, int dp,
                                  int rows_per_block, float* __restrict__ num, float* __restrict__ den) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= dp) return;
  const int i0 = blockIdx.y * rows_per_block;
  const int i1 = min(dp, i0 + rows_per_block);
  float acc = 0.f, nrm = 0.f;
  for (int i = i0; i < i1; ++i) {
    const float v = V[static_cast<long long>(i) * dp + c];
    acc = fmaf(Pm[static_cast<long long>(i) * dp + c], v, acc);
    nrm = fmaf(v, v, nrm);
  }
  atomicAdd(&num[c], acc);
  atomicAdd(&den[c], nrm);
}

__global__ void blk_rank_kernel(const float* __restrict__ num, const float* __restrict__ den, int d, int dp,
                                float* __restrict__ evals, int* __restrict__ ranks) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= dp) return;
  const float lc = num[c] / den[c];
  int rank = 0;
  for (int j = 0; j < dp; ++j) {
    const float lj = num[j] / den[j];
    rank += (lj < lc) || (lj == lc && j < c);
  }
  ranks[c] = rank;
  if (rank < d) evals[rank] = lc;
}

// evecs[j][rank[c]] = V[j][c]
__global__ void blk_gather_kernel(const float* __restrict__ V, int d, int dp, const int* __restrict__ ranks,
                                  float* __restrict__ evecs) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int rank = c < dp ? ranks[c] : d;
  for (int j = blockIdx.y; j < d; j += gridDim.y)
    if (rank < d) evecs[static_cast<long long>(j) * d + rank] = V[static_cast<long long>(j) * dp + c];
}

inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

// pair width: 64-column blocks (128-wide pairs) halve the number of rounds and matrix passes, 32-column
// blocks make the shared-memory Jacobi of a pair 8x cheaper and double the pairs in flight.  Measured on
// B200 (tools/gpu_eigh_blocked.py): see DESIGN.md.
int g_pair = 0;  // 0 = automatic
inline int pair_width(int d) {
  if (g_pair == 64 || g_pair == 128) return g_pair;
  return d > 1500 ? 128 : 64;
}
inline int pad_dim(int d, int P) { return (d + P - 1) / P * P; }

// One batched contraction of a round: out[:, blocks] (op) with operands at `off` and stride `bstride`.
int gemm6(const Parts& A, long long a_off, long long lda, long long strideA, const Parts& B, long long b_off,
          long long ldb, long long strideB, int M, int N, int K, int batch, float* C, long long c_off,
          long long ldc, long long strideC, const Parts* O, long long o_off, long long ldo, long long strideO,
          cudaStream_t stream) {
  GemmArgs g;
  g.A_hi = A.p[0] + a_off;
  g.A_lo = A.p[1] + a_off;
  g.A_lo2 = A.p[2] + a_off;
  g.lda = lda;
  g.strideA = strideA;
  g.B_hi = B.p[0] + b_off;
  g.B_lo = B.p[1] + b_off;
  g.B_lo2 = B.p[2] + b_off;
  g.ldb = ldb;
  g.strideB = strideB;
  g.M = M;
  g.N = N;
  g.K = K;
  g.batch = batch;
  g.nparts = 6;
  g.alpha = 1.f;
  g.beta = 0.f;
  if (C != nullptr) {
    g.C = C + c_off;
    g.ldc = ldc;
    g.strideC = strideC;
  }
  if (O != nullptr) {
    g.O_hi = O->p[0] + o_off;
    g.O_lo = O->p[1] + o_off;
    g.O_lo2 = O->p[2] + o_off;
    g.ldo = ldo;
    g.strideO = strideO;
  }
  return launch_umma_gemm(g, stream);
}

// Streams of one factor's solve: `main` carries the S path, `side` the eigenvector updates.  Pooled (grow-only,
// per process): a batch of wide factors advances concurrently, each factor on its own pair of streams.
struct Lane {
  cudaStream_t main = nullptr, side = nullptr;
  cudaEvent_t q_ready[2] = {nullptr, nullptr}, v_done[2] = {nullptr, nullptr}, fork = nullptr, join = nullptr;
  int* h_stats = nullptr;  // pinned: the per-sweep read-back must not block the host (other lanes are being fed)
  bool ok = false;
  Lane() {
    ok = cudaMallocHost(reinterpret_cast<void**>(&h_stats), 16) == cudaSuccess &&
         cudaStreamCreateWithFlags(&main, cudaStreamNonBlocking) == cudaSuccess &&
         cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&join, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i)
      ok = cudaEventCreateWithFlags(&q_ready[i], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&v_done[i], cudaEventDisableTiming) == cudaSuccess;
  }
};

Lane* get_lane(int i) {
  static std::vector<Lane*> pool;
  while (static_cast<int>(pool.size()) <= i) pool.push_back(new Lane());
  return pool[i]->ok ? pool[i] : nullptr;
}

template <int P>
int launch_inner(int npairs, const Parts& Ss, int dp, int off, float tol, float floor_rel, const float* fro2,
                 int sweeps, const Parts& Qs, int* stats, cudaStream_t stream) {
  constexpr size_t smem = 3ull * P * (P + 1) * 4;  // two G buffers + Q^T
  static DeviceOnce attr_once;
  if (!attr_once([] {
        return cudaFuncSetAttribute(blk_inner_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(smem)) == cudaSuccess;
      }))
    return -5;
  blk_inner_kernel<P><<<npairs, P * 8, smem, stream>>>(Ss, dp, off, tol, floor_rel, fro2, sweeps, Qs, stats);
  note_launch();
  return 0;
}

}  // namespace

void set_eigh_pair_width(int p) { g_pair = p; }

size_t eigh_blocked_workspace_bytes(int d) {
  const int P = pair_width(d);
  const size_t dp = static_cast<size_t>(pad_dim(d, P));
  const size_t f32 = align256(dp * dp * 4), b16 = align256(dp * dp * 2);
  const size_t npairs = dp / P;
  // P_f32 (Rayleigh products), V_f32, 3 x (S, T, V0, V1) splits, 2 x Q^T splits, identity splits, lam, ranks, stats
  return 2 * f32 + 12 * b16 + 6 * align256(npairs * P * P * 2) + 3 * align256(static_cast<size_t>(P / 2) * (P / 2) * 2) +
         3 * align256(dp * 4) + 256;
}

namespace {

// One wide factor: buffers carved from its slice of the workspace, work enqueued sweep by sweep on its lane.
struct BlockedSolve {
  const float* F;
  long long ldf;
  int d, kP, kB, dp, nb;
  float sym_scale, tol;
  float *evals, *evecs, *fro2;
  float *Pm, *V, *lam, *den;
  Parts Ss, Ts, Vs[2], Qs[2], Is;
  int *ranks, *stats;
  Lane* lane;
  int cur = 0;
  long long round = 0;
  int sweeps_done = 0;
  bool converged = false;

  // Rotations are skipped below this fraction of |S|_F (null space of rank-deficient factors).  It has to stay
  // this low: raising it to the rounding-noise level of the WORST elements (those coupled to a dominant
  // eigenvalue, ~5 eps sqrt(d) |S|_F) leaves off-diagonals that matter for the bulk of the spectrum
  // (reconstruction error 5e-4 -> 8e-3 at d = 4097).
  float floor_rel() const { return 5e-7f; }
  float prev_off = 1e30f;

  int setup(void* workspace) {
    kP = pair_width(d);
    kB = kP / 2;
    dp = pad_dim(d, kP);
    nb = dp / kB;  // even
    const size_t f32 = align256(static_cast<size_t>(dp) * dp * 4), b16 = align256(static_cast<size_t>(dp) * dp * 2);
    char* w = static_cast<char*>(workspace);
    Pm = reinterpret_cast<float*>(w);  // fp32 scratch: S at init, S V in the final Rayleigh pass
    w += f32;
    V = reinterpret_cast<float*>(w);
    w += f32;
    auto carve = [&](Parts& X, size_t bytes) {
      for (int i = 0; i < 3; ++i) {
        X.p[i] = reinterpret_cast<__nv_bfloat16*>(w);
        w += bytes;
      }
    };
    carve(Ss, b16);
    carve(Ts, b16);
    carve(Vs[0], b16);
    carve(Vs[1], b16);
    carve(Qs[0], align256(static_cast<size_t>(dp / kP) * kP * kP * 2));
    carve(Qs[1], align256(static_cast<size_t>(dp / kP) * kP * kP * 2));
    carve(Is, align256(static_cast<size_t>(kB) * kB * 2));
    lam = reinterpret_cast<float*>(w);  // [2][dp]: Rayleigh numerators, squared norms
    w += 2 * align256(static_cast<size_t>(dp) * 4);
    den = lam + align256(static_cast<size_t>(dp) * 4) / 4;
    ranks = reinterpret_cast<int*>(w);
    w += align256(static_cast<size_t>(dp) * 4);
    stats = reinterpret_cast<int*>(w);
    return 0;
  }

  int begin() {
    cudaStream_t stream = lane->main;
    if (cudaMemsetAsync(fro2, 0, 4, stream) != cudaSuccess) return -5;
    const int t32 = (dp + 31) / 32;
    blk_init_kernel<<<dim3(t32, t32), dim3(32, 8), 0, stream>>>(F, ldf, d, dp, sym_scale, Pm, V, fro2);
    if (dp > d) blk_pad_diag_kernel<<<(dp - d + 127) / 128, 128, 0, stream>>>(Pm, d, dp, fro2);
    blk_identity_kernel<<<(kB * kB + 255) / 256, 256, 0, stream>>>(kB, Is.p[0], Is.p[1], Is.p[2]);
    note_launch(dp > d ? 3 : 2);
    int rc = launch_convert_split3(Pm, dp, dp, dp, Ss.p[0], Ss.p[1], Ss.p[2], dp, stream);
    if (rc) return rc;
    return launch_convert_split3(V, dp, dp, dp, Vs[0].p[0], Vs[0].p[1], Vs[0].p[2], dp, stream);
  }

  // one sweep = nb rounds; ends with an asynchronous read-back of the sweep's statistics
  int enqueue_sweep() {
    cudaStream_t stream = lane->main;
    const long long ld = dp;
    const long long qstride = static_cast<long long>(kP) * kP;
    const long long bs = static_cast<long long>(nb - 1) * kB;  // blocks 0 and nb-1 sit odd rounds out: carried
                                                               // through unchanged (identity Q)
    int rc = 0;
    if (cudaMemsetAsync(stats, 0, 16, stream) != cudaSuccess) return -5;
    int inner_sweeps = sweeps_done == 0 ? 2 : 1;
    if (const char* e = getenv("BK_EIGH_INNER")) {  // bring-up: "2,2,1" = pair-solve sweeps per outer sweep
      int k = 0;
      for (const char* c = e; *c; ++c) {
        if (*c == ',') { ++k; continue; }
        if (k <= sweeps_done) inner_sweeps = *c - '0';
      }
    }
    for (int t = 0; t < nb; ++t, ++round) {
      const int odd = t & 1;
      const int off = odd ? kB : 0;
      const int npairs = odd ? nb / 2 - 1 : nb / 2;
      const int slot = static_cast<int>(round & 1);
      const Parts& Q = Qs[slot];
      const Parts& Vin = Vs[cur];
      const Parts& Vout = Vs[cur ^ 1];
      // Q^T slot `slot` was last read by the eigenvector update of round - 2 (side stream)
      if (round >= 2 && cudaStreamWaitEvent(stream, lane->v_done[slot], 0) != cudaSuccess) return -5;
      if (npairs > 0) {
        rc = kP == 128 ? launch_inner<128>(npairs, Ss, dp, off, tol, floor_rel(), fro2, inner_sweeps, Q, stats, stream)
                       : launch_inner<64>(npairs, Ss, dp, off, tol, floor_rel(), fro2, inner_sweeps, Q, stats, stream);
        if (rc) return rc;
      }
      if (cudaEventRecord(lane->q_ready[slot], stream) != cudaSuccess) return -5;
      // step 3 on the side stream (overlaps the S path of this round and the next pair solves):
      // V[:, cols] = V[:, cols] Q
      if (cudaStreamWaitEvent(lane->side, lane->q_ready[slot], 0) != cudaSuccess) return -5;
      if (npairs > 0) {
        rc = gemm6(Vin, off, ld, kP, Q, 0, kP, qstride, dp, kP, kP, npairs, V, off, ld, kP, &Vout, off, ld, kP,
                   lane->side);
        if (rc) return rc;
      }
      if (odd) {
        rc = gemm6(Vin, 0, ld, bs, Is, 0, kB, 0, dp, kB, kB, 2, V, 0, ld, bs, &Vout, 0, ld, bs, lane->side);
        if (rc) return rc;
      }
      if (cudaEventRecord(lane->v_done[slot], lane->side) != cudaSuccess) return -5;
      // step 1 (all rows of T before any column slab of it is read): T[rows, :] = Q^T S[rows, :]
      if (npairs > 0) {
        rc = gemm6(Q, 0, kP, qstride, Ss, off, ld, kP, kP, dp, kP, npairs, nullptr, 0, 0, 0, &Ts,
                   static_cast<long long>(off) * ld, ld, static_cast<long long>(kP) * ld, stream);
        if (rc) return rc;
      }
      if (odd) {
        rc = gemm6(Is, 0, kB, 0, Ss, 0, ld, bs, kB, dp, kB, 2, nullptr, 0, 0, 0, &Ts, 0, ld, bs * ld, stream);
        if (rc) return rc;
      }
      // step 2: S[:, cols] = T[:, cols] Q
      if (npairs > 0) {
        rc = gemm6(Ts, off, ld, kP, Q, 0, kP, qstride, dp, kP, kP, npairs, nullptr, 0, 0, 0, &Ss, off, ld, kP,
                   stream);
        if (rc) return rc;
      }
      if (odd) {
        rc = gemm6(Ts, 0, ld, bs, Is, 0, kB, 0, dp, kB, kB, 2, nullptr, 0, 0, 0, &Ss, 0, ld, bs, stream);
        if (rc) return rc;
      }
      cur ^= 1;
    }
    ++sweeps_done;
    return cudaMemcpyAsync(lane->h_stats, stats, 16, cudaMemcpyDeviceToHost, stream) == cudaSuccess ? 0 : -5;
  }

  // after the lane's main stream has been synchronised
  void check() {
    float smax;
    memcpy(&smax, &lane->h_stats[1], 4);
    // off = sqrt(2 sum g_pq^2) / |S|_F over the off-diagonals inspected in this sweep (every element of a pair
    // block once per visit): the classical off(S) measure.  With one pair-solve sweep per visit it falls by a
    // factor 0.5 - 0.65 per outer sweep until it reaches the rounding floor of the fp32 slab products (measured:
    // 1.8e-4 at d = 1025, 6e-4 at d = 2049, 7e-4 .. 1.2e-3 at d = 4097) and then stays put.  Stop after a sweep
    // that rotated nothing, or when off has stopped falling near that floor.  (Neither the largest rotation
    // angle nor the largest eliminated element is a usable measure: near-degenerate pairs at the noise floor of
    // a rank-deficient factor turn by 45 degrees for ever, and a maximum says little about 1.7e7 elements.)
    float off2;
    memcpy(&off2, &lane->h_stats[3], 4);
    const float off = sqrtf(2.f * off2);
    converged = (lane->h_stats[0] == 0) || (off < 1.5e-3f && off > 0.75f * prev_off);
    prev_off = off;
    if (getenv("BK_EIGH_DEBUG") != nullptr) {  // bring-up trace
      float asym;
      memcpy(&asym, &lane->h_stats[2], 4);
      fprintf(stderr, "[bk_eigh] d=%d sweep %d: rotated=%d max|g_pq|/|S|=%.3e asym/|S|=%.3e off/|S|=%.3e\n", d,
              sweeps_done, lane->h_stats[0], smax, asym, off);
    }
  }

  int finish() {
    cudaStream_t stream = lane->main;
    const long long ld = dp;
    // the last eigenvector updates run on the side stream
    if (round >= 1 && cudaStreamWaitEvent(stream, lane->v_done[(round - 1) & 1], 0) != cudaSuccess) return -5;
    if (round >= 2 && cudaStreamWaitEvent(stream, lane->v_done[round & 1], 0) != cudaSuccess) return -5;
    // eigenvalues as Rayleigh quotients with the ORIGINAL matrix: lambda_c = v_c^T S v_c / v_c^T v_c (second-order
    // accurate in the eigenvector error, free of the rounding drift of the iterated S): P = S V
    const int t32 = (dp + 31) / 32;
    blk_init_kernel<<<dim3(t32, t32), dim3(32, 8), 0, stream>>>(F, ldf, d, dp, sym_scale, Pm, nullptr, nullptr);
    if (dp > d) blk_pad_diag_kernel<<<(dp - d + 127) / 128, 128, 0, stream>>>(Pm, d, dp, fro2);
    note_launch(dp > d ? 2 : 1);
    int rc = launch_convert_split3(Pm, dp, dp, dp, Ss.p[0], Ss.p[1], Ss.p[2], dp, stream);
    if (rc) return rc;
    rc = launch_transpose_split3(V, dp, dp, dp, Ts.p[0], Ts.p[1], Ts.p[2], dp, stream);
    if (rc) return rc;
    rc = gemm6(Ss, 0, ld, 0, Ts, 0, ld, 0, dp, dp, dp, 1, Pm, 0, ld, 0, nullptr, 0, 0, 0, stream);
    if (rc) return rc;
    if (cudaMemsetAsync(lam, 0, 2 * align256(static_cast<size_t>(dp) * 4), stream) != cudaSuccess) return -5;
    const int rows_per_block = 64;
    blk_coldot_kernel<<<dim3((dp + 127) / 128, (dp + rows_per_block - 1) / rows_per_block), 128, 0, stream>>>(
        Pm, V, dp, rows_per_block, lam, den);
    blk_rank_kernel<<<(dp + 127) / 128, 128, 0, stream>>>(lam, den, d, dp, evals, ranks);
    note_launch(2);
    if (evecs != nullptr) {
      const int gy = d < 1024 ? d : 1024;
      blk_gather_kernel<<<dim3((dp + 127) / 128, gy), 128, 0, stream>>>(V, d, dp, ranks, evecs);
      note_launch();
    }
    return 0;
  }
};

}  // namespace

// Batch of wide factors, all advancing concurrently (one lane = two streams per factor; the pair solves of
// a single factor occupy only d / pair_width SMs and its slab GEMMs are latency-bound, so several factors
// fit side by side).  status[i] = 0 / 1 (not converged after max_sweeps).  Returns 0 or a negative error.
// The caller's stream is forked into the lanes and joined again; the host blocks once per sweep.
int eigh_blocked_batch(const float* const* F, const long long* ldf, const int* dims, int count, float sym_scale,
                       const float* tols, int max_sweeps, float* const* evals, float* const* evecs,
                       float* fro2, int* status, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (count <= 0) return 0;
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return -6;
  std::vector<BlockedSolve> solves(count);
  size_t off = 0;
  for (int i = 0; i < count; ++i) {
    BlockedSolve& s = solves[i];
    if (dims[i] <= 0 || F[i] == nullptr || evals[i] == nullptr) return -2;
    s.F = F[i];
    s.ldf = ldf[i];
    s.d = dims[i];
    s.sym_scale = sym_scale;
    s.tol = tols[i];
    s.evals = evals[i];
    s.evecs = evecs != nullptr ? evecs[i] : nullptr;
    s.fro2 = fro2 + i;
    s.lane = get_lane(i);
    if (s.lane == nullptr) return -5;
    const size_t need = eigh_blocked_workspace_bytes(dims[i]);
    if (off + need > workspace_bytes) return -6;
    s.setup(static_cast<char*>(workspace) + off);
    off += need;
  }
  // fork: every lane starts after the work already queued on the caller's stream
  Lane* l0 = solves[0].lane;
  if (cudaEventRecord(l0->fork, stream) != cudaSuccess) return -5;
  for (auto& s : solves) {
    if (cudaStreamWaitEvent(s.lane->main, l0->fork, 0) != cudaSuccess) return -5;
    if (cudaStreamWaitEvent(s.lane->side, l0->fork, 0) != cudaSuccess) return -5;
    const int rc = s.begin();
    if (rc) return rc;
  }
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    bool any = false;
    for (auto& s : solves) {
      if (s.converged) continue;
      const int rc = s.enqueue_sweep();
      if (rc) return rc;
      any = true;
    }
    if (!any) break;
    for (auto& s : solves) {
      if (s.converged || s.sweeps_done != sweep + 1) continue;
      if (cudaStreamSynchronize(s.lane->main) != cudaSuccess) return -5;
      s.check();
    }
  }
  for (int i = 0; i < count; ++i) {
    BlockedSolve& s = solves[i];
    const int rc = s.finish();
    if (rc) return rc;
    status[i] = s.converged ? 0 : 1;
    // join: the caller's stream continues after this factor's outputs are complete
    if (cudaEventRecord(s.lane->join, s.lane->main) != cudaSuccess) return -5;
    if (cudaStreamWaitEvent(stream, s.lane->join, 0) != cudaSuccess) return -5;
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace bk
