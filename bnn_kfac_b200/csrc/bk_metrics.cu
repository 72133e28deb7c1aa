// bk_metrics.cu — calibration metrics of predicted class probabilities, SURVEY §8(f) row f3.
// Reference: models/utilities.py:178-366 (numpy on the host, one Python loop per bin).
//
//   calibration_rows   one pass over probs [n, classes]: per row confidence = max p (:232-233),
//                      prediction = argmax (first maximum, np.argmax), correct = (pred == label) (:189),
//                      -log(p[label] + 1e-12) (:247), scipy entropy of the row (p normalised to sum 1,
//                      natural log, 0 log 0 = 0; :349-351) — plus their sums over rows in fp64.
//   binned_stats       per bin count / sum of two weights for three interval conventions:
//                        mode 0  (lo, hi]   expected_calibration_error :322
//                        mode 1  (lo, hi)   calibration_curve :287
//                        mode 2  [lo, hi), last bin [lo, hi]   np.histogram (binned_kl_distance :207-208)
//                      Edges are fp64 and values are compared in fp64, as numpy does when it compares a
//                      float32 confidence with a linspace edge.  Edges may repeat (calibration_curve takes
//                      them from the sorted confidences): binary search for the number of edges below the
//                      value, then the interval rule; warp-aggregated shared-memory accumulation.
// HBM-bound single passes; the per-bin tail (<= 256 numbers) is combined on the host.
#include "bk_common.cuh"
#include "bk_kernels.cuh"

namespace bk {

namespace {

// Per-row reduction shared by both layouts.  acc: running (best, arg, sum, sum p log p); fp32 logf per
// element and per row (1 ulp: absolute error < 4e-6 on an NLL term, far below the statistical error of a mean over
// rows); the sums over rows are carried in fp64.
struct RowAcc {
  float best = -INFINITY;
  int arg = 0x7fffffff;
  // per-thread partial sums in fp32 (<= 64 addends per thread on either path), fp64 across threads / rows
  float sum = 0.f, plogp = 0.f;
  __device__ __forceinline__ void take(float v, int c) {
    if (v > best || (v == best && c < arg)) {
      best = v;
      arg = c;
    }
    sum += v;
    if (v > 0.f) plogp = fmaf(v, logf(v), plogp);
  }
};

struct RowOut {
  float* conf;
  float* correct;
  float* nll;
  float* ent;
  int* pred;
};

__device__ __forceinline__ void finish_row(float best, int arg, double sum, double plogp, int row, const float* p,
                                           const long long* labels, int classes, const RowOut& o,
                                           double (&t)[4]) {
  struct {
    float best;
    int arg;
    double sum, plogp;
  } a{best, arg, sum, plogp};
  const long long lab = labels != nullptr ? labels[row] : -1;
  const float ok = (lab == a.arg) ? 1.f : 0.f;
  float nl = 0.f;
  if (lab >= 0 && lab < classes) nl = -logf(static_cast<float>(static_cast<double>(p[lab]) + 1e-12));
  // entropy of p / S:  -(1/S) sum p log p + log S
  const float e = a.sum > 0.0 ? static_cast<float>(-a.plogp / a.sum) + logf(static_cast<float>(a.sum)) : 0.f;
  if (o.conf != nullptr) o.conf[row] = a.best;
  if (o.correct != nullptr) o.correct[row] = ok;
  if (o.nll != nullptr) o.nll[row] = nl;
  if (o.ent != nullptr) o.ent[row] = e;
  if (o.pred != nullptr) o.pred[row] = a.arg;
  t[0] += ok;
  t[1] += a.best;
  t[2] += nl;
  t[3] += e;
}

__device__ __forceinline__ void block_totals(double (&t)[4], double* __restrict__ totals) {
  __shared__ double red[4][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double v = warp_sum(t[k]);
    if (lane == 0) red[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    if (s != 0.0) atomicAdd(&totals[threadIdx.x], s);
  }
}

// few classes (<= 64, the reference's 10-class nets): a block stages 256 rows with coalesced loads into shared
// memory (pitch classes + 1: conflict-free), then one thread per row.
constexpr int kNarrowMax = 64;
__global__ void __launch_bounds__(256)
calibration_rows_narrow_kernel(const float* __restrict__ probs, long long ld, const long long* __restrict__ labels,
                               int n, int classes, RowOut o, double* __restrict__ totals) {
  extern __shared__ float tile[];  // [256][classes + 1]
  const int pitch = classes + 1;
  double t[4] = {0.0, 0.0, 0.0, 0.0};
  const bool dense = ld == classes;
  for (int row0 = blockIdx.x * 256; row0 < n; row0 += gridDim.x * 256) {
    const int rows = min(256, n - row0);
    // (r, c) of element idx = threadIdx.x + 256 k advanced incrementally: one division per tile, not per element
    int r = threadIdx.x / classes, c = threadIdx.x - r * classes;
    const int dr = 256 / classes, dc = 256 - dr * classes;
    const float* src = probs + static_cast<long long>(row0) * ld;
    const int total = rows * classes;
    for (int idx = threadIdx.x; idx < total; idx += 4 * 256) {
      float v[4];
      int rr[4], cc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        rr[u] = r;
        cc[u] = c;
        v[u] = (idx + u * 256 < total) ? __ldcs(dense ? src + idx + u * 256 : src + static_cast<long long>(r) * ld + c)
                                       : 0.f;
        r += dr;
        c += dc;
        if (c >= classes) {
          c -= classes;
          ++r;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (idx + u * 256 < total) tile[rr[u] * pitch + cc[u]] = v[u];
    }
    __syncthreads();
    if (static_cast<int>(threadIdx.x) < rows) {
      RowAcc a;
      const float* p = tile + threadIdx.x * pitch;
      for (int c = 0; c < classes; ++c) a.take(p[c], c);
      finish_row(a.best, a.arg, a.sum, a.plogp, row0 + threadIdx.x, p, labels, classes, o, t);
    }
    __syncthreads();
  }
  block_totals(t, totals);
}

// many classes: one warp per row, lanes stride the classes
__global__ void __launch_bounds__(256)
calibration_rows_kernel(const float* __restrict__ probs, long long ld, const long long* __restrict__ labels,
                        int n, int classes, RowOut o, double* __restrict__ totals) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double t[4] = {0.0, 0.0, 0.0, 0.0};
  for (int row = blockIdx.x * 8 + warp; row < n; row += gridDim.x * 8) {
    const float* p = probs + static_cast<long long>(row) * ld;
    RowAcc a;
    for (int c = lane; c < classes; c += 4 * 32) {  // four independent loads in flight per lane
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (c + 32 * u < classes) ? __ldcs(p + c + 32 * u) : -INFINITY;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (c + 32 * u < classes) a.take(v[u], c + 32 * u);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, a.best, off);
      const int oa = __shfl_xor_sync(0xffffffffu, a.arg, off);
      if (ob > a.best || (ob == a.best && oa < a.arg)) {
        a.best = ob;
        a.arg = oa;
      }
    }
    const double sum = warp_sum(static_cast<double>(a.sum));
    const double plogp = warp_sum(static_cast<double>(a.plogp));
    if (lane == 0) finish_row(a.best, a.arg, sum, plogp, row, p, labels, classes, o, t);
  }
  block_totals(t, totals);
}

constexpr int kMaxBins = 256;

__global__ void __launch_bounds__(256)
binned_stats_kernel(const float* __restrict__ x, const float* __restrict__ w1, const float* __restrict__ w2,
                    long long n, const double* __restrict__ edges, int nbins, int mode,
                    double* __restrict__ out) {
  __shared__ double se[kMaxBins + 1];
  __shared__ double acc[3][kMaxBins];
  for (int t = threadIdx.x; t <= nbins; t += blockDim.x) se[t] = edges[t];
  for (int t = threadIdx.x; t < 3 * kMaxBins; t += blockDim.x) (&acc[0][0])[t] = 0.0;
  __syncthreads();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long tid0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long n_up = (n + 31) / 32 * 32;  // whole warps stay in the loop (shuffles below)
  for (long long i = tid0; i < n_up; i += stride) {
    int bin = -1;
    float a1 = 0.f, a2 = 0.f;
    if (i < n) {
      const double v = static_cast<double>(x[i]);
      // edges ascending (repeats allowed): k = number of edges below v (modes 0 / 1) or not above v (mode 2)
      int lo = 0, hi = nbins + 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const bool below = mode == 2 ? se[mid] <= v : se[mid] < v;
        if (below) lo = mid + 1;
        else hi = mid;
      }
      const int k = lo;  // edges[k - 1] < (<=) v, edges[k] >= (>) v
      if (mode == 0) {
        if (k >= 1 && k <= nbins) bin = k - 1;                       // (lo, hi]: edges[k] >= v
      } else if (mode == 1) {
        if (k >= 1 && k <= nbins && se[k] > v) bin = k - 1;          // (lo, hi): an edge value is in no bin
      } else {
        if (k >= 1 && k <= nbins) bin = k - 1;                       // [lo, hi)
        else if (k == nbins + 1 && v == se[nbins]) bin = nbins - 1;  // last bin closed on the right
      }
      if (bin >= 0) {
        if (w1 != nullptr) a1 = w1[i];
        if (w2 != nullptr) a2 = w2[i];
      }
    }
    // one shared-memory atomic per (warp, distinct bin) instead of one per element: calibrated nets put most
    // confidences into one or two bins
    unsigned todo = __ballot_sync(0xffffffffu, bin >= 0);
    const int lane = threadIdx.x & 31;
    while (todo) {
      const int leader = __ffs(todo) - 1;
      const int b = __shfl_sync(0xffffffffu, bin, leader);
      const bool mine = bin == b;
      const unsigned peers = __ballot_sync(0xffffffffu, mine);
      float s1 = mine ? a1 : 0.f, s2 = mine ? a2 : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if (lane == leader) {
        atomicAdd(&acc[0][b], static_cast<double>(__popc(peers)));
        if (w1 != nullptr) atomicAdd(&acc[1][b], static_cast<double>(s1));
        if (w2 != nullptr) atomicAdd(&acc[2][b], static_cast<double>(s2));
      }
      todo &= ~peers;
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 3 * nbins; t += blockDim.x) {
    const int which = t / nbins, bn = t - which * nbins;
    const double v = acc[which][bn];
    if (v != 0.0) atomicAdd(&out[which * nbins + bn], v);
  }
}

}  // namespace

int launch_calibration_rows(const float* probs, long long ld, const long long* labels, int n, int classes,
                            float* conf, float* correct, float* nll, float* ent, int* pred, double* totals,
                            cudaStream_t stream) {
  if (cudaMemsetAsync(totals, 0, 4 * sizeof(double), stream) != cudaSuccess) return -5;
  if (n <= 0) return 0;
  const RowOut o{conf, correct, nll, ent, pred};
  if (classes <= kNarrowMax) {
    int grid = (n + 255) / 256;
    if (grid > kNumSMsB200 * 8) grid = kNumSMsB200 * 8;
    const size_t smem = 256ull * (classes + 1) * 4;  // <= 66.6 KB
    static DeviceOnce attr_once;
    if (!attr_once([] {
          return cudaFuncSetAttribute(calibration_rows_narrow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      256 * (kNarrowMax + 1) * 4) == cudaSuccess;
        }))
      return -5;
    calibration_rows_narrow_kernel<<<grid, 256, smem, stream>>>(probs, ld, labels, n, classes, o, totals);
  } else {
    int grid = (n + 7) / 8;
    if (grid > kNumSMsB200 * 16) grid = kNumSMsB200 * 16;
    calibration_rows_kernel<<<grid, 256, 0, stream>>>(probs, ld, labels, n, classes, o, totals);
  }
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_binned_stats(const float* x, const float* w1, const float* w2, long long n, const double* edges,
                        int nbins, int mode, double* out, cudaStream_t stream) {
  if (nbins <= 0 || nbins > kMaxBins || mode < 0 || mode > 2) return -2;
  if (cudaMemsetAsync(out, 0, 3 * sizeof(double) * nbins, stream) != cudaSuccess) return -5;
  if (n <= 0) return 0;
  long long blocks = (n + 255) / 256;
  if (blocks > kNumSMsB200 * 4) blocks = kNumSMsB200 * 4;
  binned_stats_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(x, w1, w2, n, edges, nbins, mode, out);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace bk
