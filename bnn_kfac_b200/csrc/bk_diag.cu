// bk_diag.cu — diagonal-curvature kernels.  All are single-pass, HBM-bound, coalesced and
// grid-sized in multiples of the SM count; there is no reuse to tile for.
//
// Reference semantics (paths relative to /root/reference):
//   diag_accum    models/curvatures.py:165-172   state += [W.grad | b.grad]^2 * batch_size
//   diag_invert   models/curvatures.py:202       inv = 1/sqrt(s*state + n)
//   diag_sample   models/curvatures.py:207       normal_() * inv
//   diag_quadform sampling_free/classification/classification_ll_diagonal.py:131,
//                 sampling_free/regression/regression_ll_diagonal.py:139   sum_j J_j^2 h_j
#include "bk_common.cuh"
#include "bk_kernels.cuh"

namespace bk {

namespace {

constexpr int kBlock = 256;

inline int grid_for(long long work_items) {
  long long b = (work_items + kBlock - 1) / kBlock;
  const long long cap = static_cast<long long>(kNumSMsB200) * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

__global__ void diag_accum_kernel(float* __restrict__ state, const float* __restrict__ wgrad,
                                  const float* __restrict__ bgrad, int d_out, int d_in, float scale,
                                  float beta) {
  const int dp = d_in + (bgrad != nullptr ? 1 : 0);
  const long long total = static_cast<long long>(d_out) * dp;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int o = static_cast<int>(e / dp);
    const int i = static_cast<int>(e - static_cast<long long>(o) * dp);
    const float g = (i < d_in) ? wgrad[static_cast<long long>(o) * d_in + i] : bgrad[o];
    const float prev = (beta == 0.f) ? 0.f : beta * state[e];
    state[e] = fmaf(g * g, scale, prev);
  }
}

__global__ void diag_invert_kernel(float* __restrict__ inv, const float* __restrict__ state,
                                   long long count, float add, float multiply) {
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < count;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    // reference: torch.reciprocal(s * value + n).sqrt()
    inv[e] = sqrtf(1.0f / fmaf(multiply, state[e], add));
  }
}

__global__ void diag_sample_kernel(float* __restrict__ out, const float* __restrict__ inv,
                                   long long count, int nsamples, unsigned long long seed,
                                   uint32_t sample0, uint32_t stream_id,
                                   const float* __restrict__ z_ext) {
  const long long groups = (count + 3) / 4;
  const long long total = groups * nsamples;
  for (long long g = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; g < total;
       g += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int s = static_cast<int>(g / groups);
    const long long gi = g - static_cast<long long>(s) * groups;
    float z[4];
    if (z_ext == nullptr) {
      uint32_t c[4] = {static_cast<uint32_t>(gi), static_cast<uint32_t>(gi >> 32), sample0 + s,
                       stream_id};
      philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
      box_muller(c[0], c[1], z[0], z[1]);
      box_muller(c[2], c[3], z[2], z[3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long e = gi * 4 + j;
      if (e >= count) break;
      const float zz = (z_ext != nullptr) ? z_ext[s * count + e] : z[j];
      out[s * count + e] = zz * inv[e];
    }
  }
}

// one CTA per Jacobian row b: out[b] = sum_j J[b][j]^2 * h[j]   (fp32 loads, fp64 block reduce)
__global__ void diag_quadform_kernel(float* __restrict__ out, const float* __restrict__ J,
                                     long long ldj, const float* __restrict__ h, long long count) {
  const float* row = J + blockIdx.x * ldj;
  double acc = 0.0;
  for (long long j = threadIdx.x; j < count; j += blockDim.x) {
    const float v = row[j];
    acc += static_cast<double>(v * v * h[j]);
  }
  __shared__ double part[kBlock / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < kBlock / 32) ? part[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) out[blockIdx.x] = static_cast<float>(v);
  }
}

}  // namespace

int launch_diag_accum(float* state, const float* wgrad, const float* bgrad, int d_out, int d_in,
                      float scale, float beta, cudaStream_t stream) {
  const long long total = static_cast<long long>(d_out) * (d_in + (bgrad ? 1 : 0));
  if (total <= 0) return 0;
  diag_accum_kernel<<<grid_for(total), kBlock, 0, stream>>>(state, wgrad, bgrad, d_out, d_in, scale,
                                                            beta);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_diag_invert(float* inv, const float* state, long long count, float add, float multiply,
                       cudaStream_t stream) {
  if (count <= 0) return 0;
  diag_invert_kernel<<<grid_for(count), kBlock, 0, stream>>>(inv, state, count, add, multiply);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_diag_sample(float* out, const float* inv, long long count, int nsamples,
                       unsigned long long seed, unsigned sample0, unsigned stream_id,
                       const float* z_or_null, cudaStream_t stream) {
  if (count <= 0 || nsamples <= 0) return 0;
  diag_sample_kernel<<<grid_for(((count + 3) / 4) * nsamples), kBlock, 0, stream>>>(
      out, inv, count, nsamples, seed, sample0, stream_id, z_or_null);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_diag_quadform(float* out, const float* J, long long ldj, const float* h, long long count,
                         int batch, cudaStream_t stream) {
  if (batch <= 0) return 0;
  diag_quadform_kernel<<<batch, kBlock, 0, stream>>>(out, J, ldj, h, count);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace bk
