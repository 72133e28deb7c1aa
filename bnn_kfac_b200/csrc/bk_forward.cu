// bk_forward.cu — glue kernels of the Monte-Carlo / linearised predictive (HBM- or latency-bound;
// the dense contractions themselves run in bk_umma_gemm.cu).
//
//   sample_to_weights   W_s = M + S_s[:, :-1], b_s = m_b + S_s[:, -1]     (Curvature._replace,
//                       models/curvatures.py:67-82) for a whole batch of samples, emitted directly as
//                       the bf16 (hi, lo) K-major operand of the forward GEMM (+ fp32 copies)
//   conv2d_relu_pool    per-sample conv + bias + ReLU + 2x2 max-pool for the reference CNNs
//                       (models/wrapper.py:53-101), weights differ per sample
//   predictive_moments  mean over samples of softmax(logits) (sampling/classification_sampling.py:
//                       74-79) or mean / mean-of-squares of raw outputs (regression_sampling.py:86-88)
//   frob_dot            out[b] = |<X_b, Y_b>|  — last step of the kron-free quadratic form
//                       (classification_ll_block.py:131-132 without materialising kron(Q, H))
#include "bk_common.cuh"
#include "bk_kernels.cuh"

namespace bk {

namespace {

__global__ void sample_to_weights_kernel(const float* __restrict__ samples,
                                         const float* __restrict__ mean_w,
                                         const float* __restrict__ mean_b, int d_out, int d_in,
                                         int has_bias, int nsamples, float* __restrict__ w_f32,
                                         __nv_bfloat16* __restrict__ w_hi,
                                         __nv_bfloat16* __restrict__ w_lo, long long ldw,
                                         float* __restrict__ b_f32) {
  const int dp = d_in + has_bias;
  const long long per = static_cast<long long>(d_out) * dp;
  const long long total = per * nsamples;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int s = static_cast<int>(e / per);
    const long long r = e - static_cast<long long>(s) * per;
    const int o = static_cast<int>(r / dp);
    const int i = static_cast<int>(r - static_cast<long long>(o) * dp);
    const float sv = samples[e];
    if (i < d_in) {
      const float v = mean_w[static_cast<long long>(o) * d_in + i] + sv;
      if (w_f32 != nullptr) w_f32[(static_cast<long long>(s) * d_out + o) * d_in + i] = v;
      if (w_hi != nullptr) {
        __nv_bfloat16 h, l;
        split_bf16(v, h, l);
        const long long q = (static_cast<long long>(s) * d_out + o) * ldw + i;
        w_hi[q] = h;
        if (w_lo != nullptr) w_lo[q] = l;
      }
    } else {
      b_f32[static_cast<long long>(s) * d_out + o] = mean_b[o] + sv;
    }
  }
}

// one CTA per (sample s, image n): weights of s and image n staged in shared memory
__global__ void conv2d_relu_pool_kernel(const float* __restrict__ in, long long in_sample_stride,
                                        const float* __restrict__ w, const float* __restrict__ b,
                                        float* __restrict__ out, int N, int C, int H, int W, int O,
                                        int KH, int KW, int SH, int SW, int PH, int PW, int relu,
                                        int pool) {
  extern __shared__ float sm[];
  const int s = blockIdx.y, n = blockIdx.x;
  const int OH = (H + 2 * PH - KH) / SH + 1, OW = (W + 2 * PW - KW) / SW + 1;
  const int QH = pool ? OH / 2 : OH, QW = pool ? OW / 2 : OW;
  float* ws = sm;                       // [O][C][KH][KW]
  float* bs = ws + O * C * KH * KW;     // [O]
  float* xs = bs + O;                   // [C][H][W]
  const float* wsrc = w + static_cast<long long>(s) * O * C * KH * KW;
  for (int i = threadIdx.x; i < O * C * KH * KW; i += blockDim.x) ws[i] = wsrc[i];
  for (int i = threadIdx.x; i < O; i += blockDim.x) bs[i] = b ? b[static_cast<long long>(s) * O + i] : 0.f;
  const float* xsrc = in + s * in_sample_stride + static_cast<long long>(n) * C * H * W;
  for (int i = threadIdx.x; i < C * H * W; i += blockDim.x) xs[i] = xsrc[i];
  __syncthreads();
  float* dst = out + (static_cast<long long>(s) * N + n) * O * QH * QW;
  for (int e = threadIdx.x; e < O * QH * QW; e += blockDim.x) {
    const int o = e / (QH * QW);
    const int r = e - o * QH * QW;
    const int qy = r / QW, qx = r - qy * QW;
    const int np = pool ? 2 : 1;
    float best = -INFINITY;
    for (int dy = 0; dy < np; ++dy)
      for (int dx = 0; dx < np; ++dx) {
        const int oy = qy * np + dy, ox = qx * np + dx;
        float acc = bs[o];
        for (int c = 0; c < C; ++c)
          for (int ky = 0; ky < KH; ++ky) {
            const int iy = oy * SH - PH + ky;
            if (iy < 0 || iy >= H) continue;
            for (int kx = 0; kx < KW; ++kx) {
              const int ix = ox * SW - PW + kx;
              if (ix < 0 || ix >= W) continue;
              acc = fmaf(ws[((o * C + c) * KH + ky) * KW + kx], xs[(c * H + iy) * W + ix], acc);
            }
          }
        best = fmaxf(best, acc);
      }
    dst[e] = relu ? fmaxf(best, 0.f) : best;  // relu and max-pool commute
  }
}

// Fast path of the same operation for the reference CNNs' layers (stride 1, square 3 x 3 or 5 x 5 kernels):
//   * one CTA per (sample s, group of kConvGroup images): the sampled weights of s are staged ONCE per group, the
//     images of the group sit zero-padded in shared memory (no bounds checks in the tap loops);
//   * a work item is one POOLED output (image, o, qy, qx): its 2 x 2 pre-pool outputs share one (K + 1)^2 input
//     patch per input channel, held in registers - K^2 weight words (warp-broadcast: neighbouring items share o)
//     and (K + 1)^2 input words for 4 K^2 FMAs, against one weight + one input word per FMA in the generic kernel;
//   * fully unrolled taps.  (LeNet-5, S = 100, batch 256: the generic kernel ran at 3.9 TFLOP/s and was 87 % of
//     the MC predictive.)
constexpr int kConvGroup = 8;
__host__ __device__ constexpr int conv_kkp(int K) { return (K * K + 3) / 4 * 4; }  // taps padded to float4
template <int K, bool POOL>
__global__ void __launch_bounds__(256)
conv2d_fast_kernel(const float* __restrict__ in, long long in_sample_stride, const float* __restrict__ w,
                   const float* __restrict__ b, float* __restrict__ out, int N, int C, int H, int W, int O, int PH,
                   int PW, int relu) {
  extern __shared__ __align__(16) float sm[];
  const int s = blockIdx.y, n0 = blockIdx.x * kConvGroup;
  const int imgs = min(kConvGroup, N - n0);
  const int HP = H + 2 * PH, WPr = W + 2 * PW;
  const int WP = WPr + (WPr & 1);  // even row pitch: 8-byte patch loads
  const int OH = HP - K + 1, OW = WPr - K + 1;
  const int QH = POOL ? OH / 2 : OH, QW = POOL ? OW / 2 : OW;
  constexpr int NP = POOL ? 2 : 1;   // pre-pool outputs per item and dimension
  constexpr int PS = K + NP - 1;     // input patch edge
  constexpr int KK = K * K, KKP = conv_kkp(K);
  const int OPAIRS = (O + 1) / 2;    // a work item computes TWO output channels from one input patch
  float* ws = sm;                                      // [2 * OPAIRS][C][KKP] (16-byte aligned rows)
  float* bs = ws + 2 * OPAIRS * C * KKP;               // [2 * OPAIRS]
  float* xs = bs + 2 * OPAIRS;                         // [kConvGroup][C][HP][WP], zero border
  const int img_words = C * HP * WP;
  const float* wsrc = w + static_cast<long long>(s) * O * C * KK;
  for (int i = threadIdx.x; i < 2 * OPAIRS * C * KKP; i += blockDim.x) {
    const int oc = i / KKP, t = i - oc * KKP;
    ws[i] = (t < KK && oc < O * C) ? wsrc[oc * KK + t] : 0.f;
  }
  for (int i = threadIdx.x; i < 2 * OPAIRS; i += blockDim.x)
    bs[i] = (b && i < O) ? b[static_cast<long long>(s) * O + i] : 0.f;
  {
    // padded tiles row by row: a group of 16 or 32 lanes owns one row (g, c, y) - two integer divisions per ROW, the
    // lanes walk x (per element the decode costs more instructions than the FMAs the element feeds in a
    // single-channel first layer)
    const int lpr = WP <= 16 ? 16 : 32;  // lanes per row
    const int sub = threadIdx.x / lpr, lx = threadIdx.x - sub * lpr;
    const int rows = imgs * C * HP;
    for (int row = sub; row < rows; row += 256 / lpr) {
      const int g = row / (C * HP), rc = row - g * C * HP;
      const int c = rc / HP, y = rc - c * HP - PH;
      const bool yin = y >= 0 && y < H;
      const float* src = in + s * in_sample_stride + (static_cast<long long>(n0 + g) * C + c) * H * W + y * W;
      for (int xp = lx; xp < WP; xp += lpr) {
        const int x = xp - PW;
        xs[row * WP + xp] = (yin && x >= 0 && x < W) ? src[x] : 0.f;
      }
    }
  }
  __syncthreads();
  const int per_pair = QH * QW, per_img = O * per_pair;
  // item decode by float reciprocals (exact here: every dividend is below 2^20, the + 0.5 keeps the quotient away
  // from the integer boundaries) - three integer divisions per item otherwise
  const float inv_g = 1.f / static_cast<float>(OPAIRS * per_pair), inv_p = 1.f / static_cast<float>(per_pair);
  const float inv_w = 1.f / static_cast<float>(QW);
  for (int e = threadIdx.x; e < imgs * OPAIRS * per_pair; e += blockDim.x) {
    const int g = __float2int_rz((static_cast<float>(e) + 0.5f) * inv_g), r = e - g * OPAIRS * per_pair;
    const int op = __float2int_rz((static_cast<float>(r) + 0.5f) * inv_p), r2 = r - op * per_pair;
    const int qy = __float2int_rz((static_cast<float>(r2) + 0.5f) * inv_w), qx = r2 - qy * QW;
    const int o0 = 2 * op;
    float acc[2][NP][NP];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int dy = 0; dy < NP; ++dy)
#pragma unroll
        for (int dx = 0; dx < NP; ++dx) acc[u][dy][dx] = bs[o0 + u];
    const float* xg = xs + g * img_words + (qy * NP) * WP + qx * NP;
    const float* wo = ws + o0 * C * KKP;
    for (int c = 0; c < C; ++c) {
      float patch[PS][PS];
      if (POOL) {  // PS is even and the patch starts at an even word: 8-byte loads
#pragma unroll
        for (int y = 0; y < PS; ++y)
#pragma unroll
          for (int x = 0; x < PS; x += 2) {
            const float2 v2 = *reinterpret_cast<const float2*>(xg + c * HP * WP + y * WP + x);
            patch[y][x] = v2.x;
            patch[y][x + 1] = v2.y;
          }
      } else {
#pragma unroll
        for (int y = 0; y < PS; ++y)
#pragma unroll
          for (int x = 0; x < PS; ++x) patch[y][x] = xg[c * HP * WP + y * WP + x];
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float wv[KKP];
#pragma unroll
        for (int t = 0; t < KKP; t += 4) {  // warp-broadcast 16-byte loads (neighbouring items share the pair)
          const float4 q = *reinterpret_cast<const float4*>(wo + (u * C + c) * KKP + t);
          wv[t] = q.x; wv[t + 1] = q.y; wv[t + 2] = q.z; wv[t + 3] = q.w;
        }
#pragma unroll
        for (int ky = 0; ky < K; ++ky)
#pragma unroll
          for (int kx = 0; kx < K; ++kx)
#pragma unroll
            for (int dy = 0; dy < NP; ++dy)
#pragma unroll
              for (int dx = 0; dx < NP; ++dx)
                acc[u][dy][dx] = fmaf(wv[ky * K + kx], patch[ky + dy][kx + dx], acc[u][dy][dx]);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (o0 + u >= O) continue;
      float best = acc[u][0][0];
#pragma unroll
      for (int dy = 0; dy < NP; ++dy)
#pragma unroll
        for (int dx = 0; dx < NP; ++dx) best = fmaxf(best, acc[u][dy][dx]);
      out[(static_cast<long long>(s) * N + n0 + g) * per_img + (o0 + u) * per_pair + r2] =
          relu ? fmaxf(best, 0.f) : best;
    }
  }
}

// one warp per input row b.  mode 0: p = softmax(logits[s, b, :]); mode 1: p = logits[s, b, :].
// mean[b, c] = (1/S) sum_s p ; meansq[b, c] = (1/S) sum_s p^2 (optional).  C <= 32 * kMaxPerLane.
constexpr int kMaxPerLane = 32;
// kPerLane = classes per lane the register arrays are sized (and the loops unrolled) for: 1 covers the 10-class nets
// (a 32-wide unroll made every sample cost 32 x the instructions of a 10-class row: 142 us for 100 x 256 x 10 logits)
template <int kPerLane>
__global__ void predictive_moments_kernel(const float* __restrict__ logits, int S, int B, int Cn,
                                          int mode, float* __restrict__ mean,
                                          float* __restrict__ meansq) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B) return;
  const int per = (Cn + 31) / 32;
  float m1[kPerLane], m2[kPerLane];
#pragma unroll
  for (int k = 0; k < kPerLane; ++k) m1[k] = m2[k] = 0.f;
  if (mode == 2) {
    // regression, CENTRED second moment (numpy std with ddof = 0, regression_sampling.py:86-88): pass 1
    // the mean, pass 2 sum (y - mean)^2.  E[y^2] - E[y]^2 in fp32 loses variances below ~1e-7 * mean^2
    // (|mean| reaches 200 on the reference's y = x^3 task); the [S, B, C] outputs are tiny, so the second
    // read is free.
    for (int s = 0; s < S; ++s) {
      const float* row = logits + (static_cast<long long>(s) * B + warp) * Cn;
#pragma unroll
      for (int k = 0; k < kPerLane; ++k) {
        const int c = lane + 32 * k;
        if (k < per && c < Cn) m1[k] += row[c];
      }
    }
    const float invS = 1.0f / static_cast<float>(S);
#pragma unroll
    for (int k = 0; k < kPerLane; ++k) m1[k] *= invS;
    for (int s = 0; s < S; ++s) {
      const float* row = logits + (static_cast<long long>(s) * B + warp) * Cn;
#pragma unroll
      for (int k = 0; k < kPerLane; ++k) {
        const int c = lane + 32 * k;
        if (k < per && c < Cn) {
          const float dv = row[c] - m1[k];
          m2[k] = fmaf(dv, dv, m2[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kPerLane; ++k) {
      const int c = lane + 32 * k;
      if (k < per && c < Cn) {
        mean[static_cast<long long>(warp) * Cn + c] = m1[k];
        if (meansq != nullptr) meansq[static_cast<long long>(warp) * Cn + c] = m2[k] * invS;
      }
    }
    return;
  }
  for (int s = 0; s < S; ++s) {
    const float* row = logits + (static_cast<long long>(s) * B + warp) * Cn;
    float v[kPerLane];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kPerLane; ++k) {
      const int c = lane + 32 * k;
      v[k] = (k < per && c < Cn) ? row[c] : -INFINITY;
      mx = fmaxf(mx, v[k]);
    }
    if (mode == 0) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < kPerLane; ++k) {
        v[k] = (v[k] == -INFINITY) ? 0.f : expf(v[k] - mx);
        sum += v[k];
      }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
#pragma unroll
      for (int k = 0; k < kPerLane; ++k) v[k] *= inv;
    }
#pragma unroll
    for (int k = 0; k < kPerLane; ++k) {
      const float pv = (v[k] == -INFINITY) ? 0.f : v[k];
      m1[k] += pv;
      m2[k] = fmaf(pv, pv, m2[k]);
    }
  }
  const float invS = 1.0f / static_cast<float>(S);
#pragma unroll
  for (int k = 0; k < kPerLane; ++k) {
    const int c = lane + 32 * k;
    if (k < per && c < Cn) {
      mean[static_cast<long long>(warp) * Cn + c] = m1[k] * invS;
      if (meansq != nullptr) meansq[static_cast<long long>(warp) * Cn + c] = m2[k] * invS;
    }
  }
}

// out[b] (+)= <x_b, y_b> over this CTA's chunk: grid (chunks, batch), fp64 per-CTA partials combined
// with one atomic per CTA; `absolute` is applied by frob_abs_kernel after all chunks landed.
__global__ void __launch_bounds__(256)
frob_dot_kernel(double* __restrict__ acc_out, const float* __restrict__ X, long long stride_x,
                const float* __restrict__ Y, long long stride_y, long long count, long long chunk) {
  const float* x = X + blockIdx.y * stride_x;
  const float* y = Y + blockIdx.y * stride_y;
  const long long j0 = blockIdx.x * chunk;
  const long long j1 = (j0 + chunk < count) ? j0 + chunk : count;
  double acc = 0.0;
  for (long long j = j0 + threadIdx.x; j < j1; j += 4 * 256) {
    float a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long jj = j + u * 256;
      a[u] = (jj < j1) ? __ldcs(x + jj) : 0.f;
      b[u] = (jj < j1) ? __ldcs(y + jj) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += static_cast<double>(a[u]) * static_cast<double>(b[u]);
  }
  __shared__ double part[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < 8) ? part[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(&acc_out[blockIdx.y], v);
  }
}

__global__ void frob_finish_kernel(float* __restrict__ out, const double* __restrict__ acc, int batch,
                                   int absolute, int accumulate) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < batch) {
    double v = acc[b];
    if (absolute) v = fabs(v);
    out[b] = accumulate ? out[b] + static_cast<float>(v) : static_cast<float>(v);
  }
}

}  // namespace

int launch_sample_to_weights(const float* samples, const float* mean_w, const float* mean_b,
                             int d_out, int d_in, int has_bias, int nsamples, float* w_f32,
                             __nv_bfloat16* w_hi, __nv_bfloat16* w_lo, long long ldw, float* b_f32,
                             cudaStream_t stream) {
  const long long total = static_cast<long long>(d_out) * (d_in + has_bias) * nsamples;
  if (total <= 0) return 0;
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(kNumSMsB200) * 8;
  if (blocks > cap) blocks = cap;
  sample_to_weights_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(
      samples, mean_w, mean_b, d_out, d_in, has_bias, nsamples, w_f32, w_hi, w_lo, ldw, b_f32);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

template <int K, bool POOL>
int launch_conv_fast(const float* in, long long in_sample_stride, const float* w, const float* b, float* out, int S,
                     int N, int C, int H, int W, int O, int PH, int PW, int relu, size_t smem, cudaStream_t stream) {
  static DeviceOnce attr_once;  // one instance per (K, POOL)
  if (smem > 48 * 1024 && !attr_once([] {
        return cudaFuncSetAttribute(conv2d_fast_kernel<K, POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    200 * 1024) == cudaSuccess;
      }))
    return -5;
  conv2d_fast_kernel<K, POOL><<<dim3((N + kConvGroup - 1) / kConvGroup, S), 256, smem, stream>>>(
      in, in_sample_stride, w, b, out, N, C, H, W, O, PH, PW, relu);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int g_conv_fast = 1;
void set_conv_fast(int on) { g_conv_fast = on; }

int launch_conv2d_relu_pool(const float* in, long long in_sample_stride, const float* w,
                            const float* b, float* out, int S, int N, int C, int H, int W, int O,
                            int KH, int KW, int SH, int SW, int PH, int PW, int relu, int pool,
                            cudaStream_t stream) {
  if (S <= 0 || N <= 0) return 0;
  if (g_conv_fast && SH == 1 && SW == 1 && KH == KW && (KH == 3 || KH == 5) && H + 2 * PH >= KH &&
      W + 2 * PW >= KW) {
    const int wp = W + 2 * PW + ((W + 2 * PW) & 1), op2 = (O + 1) / 2 * 2;
    const size_t fsmem = sizeof(float) * (static_cast<size_t>(op2) * C * conv_kkp(KH) + op2 +
                                          static_cast<size_t>(kConvGroup) * C * (H + 2 * PH) * wp);
    if (fsmem <= 200 * 1024) {
      if (KH == 5)
        return pool ? launch_conv_fast<5, true>(in, in_sample_stride, w, b, out, S, N, C, H, W, O, PH, PW, relu, fsmem,
                                                stream)
                    : launch_conv_fast<5, false>(in, in_sample_stride, w, b, out, S, N, C, H, W, O, PH, PW, relu,
                                                 fsmem, stream);
      return pool ? launch_conv_fast<3, true>(in, in_sample_stride, w, b, out, S, N, C, H, W, O, PH, PW, relu, fsmem,
                                              stream)
                  : launch_conv_fast<3, false>(in, in_sample_stride, w, b, out, S, N, C, H, W, O, PH, PW, relu, fsmem,
                                               stream);
    }
  }
  const size_t smem = sizeof(float) * (static_cast<size_t>(O) * C * KH * KW + O +
                                       static_cast<size_t>(C) * H * W);
  if (smem > 200 * 1024) return -2;
  static DeviceOnce attr_once;  // opt in to the kernel's maximum (200 KB) once per device
  if (smem > 48 * 1024 && !attr_once([] {
        return cudaFuncSetAttribute(conv2d_relu_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    200 * 1024) == cudaSuccess;
      }))
    return -5;
  conv2d_relu_pool_kernel<<<dim3(N, S), 256, smem, stream>>>(in, in_sample_stride, w, b, out, N, C, H,
                                                             W, O, KH, KW, SH, SW, PH, PW, relu,
                                                             pool);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_predictive_moments(const float* logits, int S, int B, int Cn, int mode, float* mean,
                              float* meansq, cudaStream_t stream) {
  if (S <= 0 || B <= 0 || Cn <= 0) return 0;
  if (Cn > 32 * kMaxPerLane) return -2;
  const int warps_per_block = 8;
  const int blocks = (B + warps_per_block - 1) / warps_per_block;
  if (Cn <= 32)
    predictive_moments_kernel<1><<<blocks, warps_per_block * 32, 0, stream>>>(logits, S, B, Cn, mode, mean, meansq);
  else if (Cn <= 128)
    predictive_moments_kernel<4><<<blocks, warps_per_block * 32, 0, stream>>>(logits, S, B, Cn, mode, mean, meansq);
  else
    predictive_moments_kernel<kMaxPerLane><<<blocks, warps_per_block * 32, 0, stream>>>(logits, S, B, Cn, mode, mean,
                                                                                      meansq);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_frob_dot(float* out, const float* X, long long stride_x, const float* Y,
                    long long stride_y, long long count, int batch, int absolute, int accumulate,
                    cudaStream_t stream) {
  if (batch <= 0) return 0;
  // fp64 accumulators: a small stream-ordered scratch (batch doubles).  The default pool hands its memory back
  // to the OS at every synchronisation unless told otherwise (0.5 ms per call measured): keep it.
  // (The one allocation this library makes, and the one device-wide setting it touches: the release threshold
  // of the device's DEFAULT stream-ordered pool is raised to 64 MB, once per device; see include/bk_kfac.h.)
  static DeviceOnce pool_once;
  pool_once([] {
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = 64ull << 20;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    return true;
  });
  double* acc = nullptr;
  if (cudaMallocAsync(reinterpret_cast<void**>(&acc), sizeof(double) * batch, stream) != cudaSuccess)
    return -5;
  if (cudaMemsetAsync(acc, 0, sizeof(double) * batch, stream) != cudaSuccess) {
    cudaFreeAsync(acc, stream);
    return -5;
  }
  long long chunk = 16384;
  const long long max_ctas = static_cast<long long>(kNumSMsB200) * 32;
  while ((count + chunk - 1) / chunk * batch > max_ctas) chunk *= 2;
  const dim3 grid(static_cast<unsigned>((count + chunk - 1) / chunk), batch);
  frob_dot_kernel<<<grid, 256, 0, stream>>>(acc, X, stride_x, Y, stride_y, count, chunk);
  frob_finish_kernel<<<(batch + 255) / 256, 256, 0, stream>>>(out, acc, batch, absolute, accumulate);
  note_launch(2);
  const bool ok = cudaGetLastError() == cudaSuccess;
  cudaFreeAsync(acc, stream);
  return ok ? 0 : -5;
}

}  // namespace bk
