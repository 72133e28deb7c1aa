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
//                      them from the sorted confidences), so every element tests every bin.
// HBM-bound single passes; the per-bin tail (<= 256 numbers) is combined on the host.
#include "bk_common.cuh"
#include "bk_kernels.cuh"

namespace bk {

namespace {

// one warp per row
__global__ void __launch_bounds__(256)
calibration_rows_kernel(const float* __restrict__ probs, long long ld, const long long* __restrict__ labels,
                        int n, int classes, float* __restrict__ conf, float* __restrict__ correct,
                        float* __restrict__ nll, float* __restrict__ ent, int* __restrict__ pred,
                        double* __restrict__ totals) {
  __shared__ double red[4][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double t_correct = 0.0, t_conf = 0.0, t_nll = 0.0, t_ent = 0.0;
  for (int row = blockIdx.x * 8 + warp; row < n; row += gridDim.x * 8) {
    const float* p = probs + static_cast<long long>(row) * ld;
    float best = -INFINITY;
    int arg = 0x7fffffff;
    double sum = 0.0, plogp = 0.0;
    for (int c = lane; c < classes; c += 32) {
      const float v = p[c];
      if (v > best || (v == best && c < arg)) {
        best = v;
        arg = c;
      }
      sum += v;
      if (v > 0.f) plogp += static_cast<double>(v) * log(static_cast<double>(v));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ob > best || (ob == best && oa < arg)) {
        best = ob;
        arg = oa;
      }
    }
    sum = warp_sum(sum);
    plogp = warp_sum(plogp);
    if (lane == 0) {
      const long long lab = labels != nullptr ? labels[row] : -1;
      const float ok = (lab == arg) ? 1.f : 0.f;
      float nl = 0.f;
      if (lab >= 0 && lab < classes) nl = static_cast<float>(-log(static_cast<double>(p[lab]) + 1e-12));
      // entropy of p / sum:  -(1/S) sum p log p + log S
      const float e = sum > 0.0 ? static_cast<float>(-plogp / sum + log(sum)) : 0.f;
      if (conf != nullptr) conf[row] = best;
      if (correct != nullptr) correct[row] = ok;
      if (nll != nullptr) nll[row] = nl;
      if (ent != nullptr) ent[row] = e;
      if (pred != nullptr) pred[row] = arg;
      t_correct += ok;
      t_conf += best;
      t_nll += nl;
      t_ent += e;
    }
  }
  if (lane == 0) {
    red[0][warp] = t_correct;
    red[1][warp] = t_conf;
    red[2][warp] = t_nll;
    red[3][warp] = t_ent;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    atomicAdd(&totals[threadIdx.x], t);
  }
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
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride) {
    const double v = static_cast<double>(x[i]);
    for (int bn = 0; bn < nbins; ++bn) {
      const double lo = se[bn], hi = se[bn + 1];
      bool in;
      if (mode == 0) in = v > lo && v <= hi;
      else if (mode == 1) in = v > lo && v < hi;
      else in = v >= lo && (v < hi || (bn == nbins - 1 && v == hi));
      if (in) {
        atomicAdd(&acc[0][bn], 1.0);
        if (w1 != nullptr) atomicAdd(&acc[1][bn], static_cast<double>(w1[i]));
        if (w2 != nullptr) atomicAdd(&acc[2][bn], static_cast<double>(w2[i]));
        if (mode != 1) break;  // modes 0 / 2 have ascending, disjoint bins
      }
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
  int grid = (n + 7) / 8;
  if (grid > kNumSMsB200 * 8) grid = kNumSMsB200 * 8;
  calibration_rows_kernel<<<grid, 256, 0, stream>>>(probs, ld, labels, n, classes, conf, correct, nll, ent,
                                                    pred, totals);
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
