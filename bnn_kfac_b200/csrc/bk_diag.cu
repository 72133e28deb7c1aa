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
  const long long cap = static_cast<long long>(kNumSMsB200) * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

// Every kernel below is one streaming pass; what decides its bandwidth is bytes in flight: each thread
// issues 4 independent (128-bit where the row alignment allows) loads before the first use, grids
// are sized to cover the chip several times (Little: ~5 MB must be in flight for 6.5 TB/s).
constexpr int kUnroll = 4;

// 2-D launch: blockIdx.y = row o, blockIdx.x = column chunk of kBlock * kUnroll columns.
__global__ void __launch_bounds__(kBlock)
diag_accum_kernel(float* __restrict__ state, const float* __restrict__ wgrad,
                  const float* __restrict__ bgrad, int d_out, int d_in, float scale, float beta) {
  const int dp = d_in + (bgrad != nullptr ? 1 : 0);
  for (int o = blockIdx.y; o < d_out; o += gridDim.y) {
    const float* w = wgrad + static_cast<long long>(o) * d_in;
    float* st = state + static_cast<long long>(o) * dp;
    const int c0 = blockIdx.x * (kBlock * kUnroll) + threadIdx.x;
    float g[kUnroll], old[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int c = c0 + u * kBlock;
      g[u] = (c < d_in) ? w[c] : 0.f;
      old[u] = (c < d_in && beta != 0.f) ? st[c] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int c = c0 + u * kBlock;
      if (c < d_in) st[c] = fmaf(g[u] * g[u], scale, beta * old[u]);
    }
    if (bgrad != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
      const float gb = bgrad[o];
      const float prev = (beta == 0.f) ? 0.f : beta * st[d_in];
      st[d_in] = fmaf(gb * gb, scale, prev);
    }
  }
}

__global__ void __launch_bounds__(kBlock)
diag_invert_kernel(float* __restrict__ inv, const float* __restrict__ state, long long count,
                   float add, float multiply) {
  // reference: torch.reciprocal(s * value + n).sqrt()
  const bool vec = ((reinterpret_cast<uintptr_t>(inv) | reinterpret_cast<uintptr_t>(state)) & 15) == 0;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (vec) {
    const long long n4 = count >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(state);
    float4* o4 = reinterpret_cast<float4*>(inv);
    for (long long i = tid; i < n4; i += stride * kUnroll) {
      float4 v[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
        if (i + u * stride < n4) v[u] = __ldcs(s4 + i + u * stride);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        if (i + u * stride < n4) {
          float4 r;
          // 1/sqrt(x) by one MUFU.RSQ (2 ulp; the division + sqrt pair made this pass ALU-bound)
          r.x = rsqrtf(fmaf(multiply, v[u].x, add));
          r.y = rsqrtf(fmaf(multiply, v[u].y, add));
          r.z = rsqrtf(fmaf(multiply, v[u].z, add));
          r.w = rsqrtf(fmaf(multiply, v[u].w, add));
          o4[i + u * stride] = r;
        }
      }
    }
    for (long long e = (n4 << 2) + tid; e < count; e += stride)
      inv[e] = rsqrtf(fmaf(multiply, state[e], add));
  } else {
    for (long long e = tid; e < count; e += stride)
      inv[e] = rsqrtf(fmaf(multiply, state[e], add));
  }
}

__global__ void __launch_bounds__(kBlock)
diag_sample_kernel(float* __restrict__ out, const float* __restrict__ inv, long long count,
                   int nsamples, unsigned long long seed, uint32_t sample0, uint32_t stream_id,
                   const float* __restrict__ z_ext) {
  // element e of sample s: lane e % 4 of Philox counter (e / 4, sample0 + s, stream_id)
  const long long groups = (count + 3) / 4;
  const bool vec = (count & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(inv) |
                     reinterpret_cast<uintptr_t>(z_ext)) & 15) == 0;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (int s = blockIdx.y; s < nsamples; s += gridDim.y) {
    for (long long gi = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; gi < groups;
         gi += stride) {
      float z[4];
      if (z_ext == nullptr) {
        uint32_t c[4] = {static_cast<uint32_t>(gi), static_cast<uint32_t>(gi >> 32), sample0 + s,
                         stream_id};
        philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
        box_muller(c[0], c[1], z[0], z[1]);
        box_muller(c[2], c[3], z[2], z[3]);
      }
      if (vec) {
        const float4 h = __ldg(reinterpret_cast<const float4*>(inv) + gi);
        if (z_ext != nullptr) {
          const float4 ze = __ldcs(reinterpret_cast<const float4*>(z_ext + s * count) + gi);
          z[0] = ze.x; z[1] = ze.y; z[2] = ze.z; z[3] = ze.w;
        }
        reinterpret_cast<float4*>(out + s * count)[gi] =
            make_float4(z[0] * h.x, z[1] * h.y, z[2] * h.z, z[3] * h.w);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const long long e = gi * 4 + j;
          if (e >= count) break;
          const float zz = (z_ext != nullptr) ? z_ext[s * count + e] : z[j];
          out[s * count + e] = zz * inv[e];
        }
      }
    }
  }
}

// out[b] += sum_j J[b][j]^2 * h[j] over this CTA's column chunk (out zero-filled by the launcher):
// grid (chunks, rows) so that a handful of Jacobian rows still fills the chip; fp64 partial per CTA.
__global__ void __launch_bounds__(kBlock)
diag_quadform_kernel(float* __restrict__ out, const float* __restrict__ J, long long ldj,
                     const float* __restrict__ h, long long count, long long chunk) {
  const float* row = J + blockIdx.y * ldj;
  const long long j0 = blockIdx.x * chunk;
  const long long j1 = (j0 + chunk < count) ? j0 + chunk : count;
  double acc = 0.0;
  for (long long j = j0 + threadIdx.x; j < j1; j += kBlock * kUnroll) {
    float v[kUnroll], hv[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long jj = j + u * kBlock;
      v[u] = (jj < j1) ? __ldcs(row + jj) : 0.f;
      hv[u] = (jj < j1) ? __ldg(h + jj) : 0.f;
    }
    float part = 0.f;
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) part = fmaf(v[u] * v[u], hv[u], part);
    acc += static_cast<double>(part);
  }
  __shared__ double part_s[kBlock / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part_s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < kBlock / 32) ? part_s[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(&out[blockIdx.y], static_cast<float>(v));
  }
}

}  // namespace

int launch_diag_accum(float* state, const float* wgrad, const float* bgrad, int d_out, int d_in,
                      float scale, float beta, cudaStream_t stream) {
  const long long total = static_cast<long long>(d_out) * (d_in + (bgrad ? 1 : 0));
  if (total <= 0) return 0;
  const int chunks = (d_in + kBlock * kUnroll - 1) / (kBlock * kUnroll);
  const dim3 grid(chunks, d_out < 65535 ? d_out : 65535);
  diag_accum_kernel<<<grid, kBlock, 0, stream>>>(state, wgrad, bgrad, d_out, d_in, scale, beta);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_diag_invert(float* inv, const float* state, long long count, float add, float multiply,
                       cudaStream_t stream) {
  if (count <= 0) return 0;
  diag_invert_kernel<<<grid_for((count + 15) / 16), kBlock, 0, stream>>>(inv, state, count, add,
                                                                        multiply);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_diag_sample(float* out, const float* inv, long long count, int nsamples,
                       unsigned long long seed, unsigned sample0, unsigned stream_id,
                       const float* z_or_null, cudaStream_t stream) {
  if (count <= 0 || nsamples <= 0) return 0;
  const dim3 grid(grid_for((count + 3) / 4), nsamples < 64 ? nsamples : 64);
  diag_sample_kernel<<<grid, kBlock, 0, stream>>>(out, inv, count, nsamples, seed, sample0,
                                                  stream_id, z_or_null);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_diag_quadform(float* out, const float* J, long long ldj, const float* h, long long count,
                         int batch, cudaStream_t stream) {
  if (batch <= 0) return 0;
  if (cudaMemsetAsync(out, 0, sizeof(float) * batch, stream) != cudaSuccess) return -5;
  // ~16 K elements per CTA, but never more CTAs than ~32 per SM
  long long chunk = 16384;
  const long long max_ctas = static_cast<long long>(kNumSMsB200) * 32;
  while ((count + chunk - 1) / chunk * batch > max_ctas) chunk *= 2;
  const dim3 grid(static_cast<unsigned>((count + chunk - 1) / chunk), batch);
  diag_quadform_kernel<<<grid, kBlock, 0, stream>>>(out, J, ldj, h, count, chunk);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace bk
