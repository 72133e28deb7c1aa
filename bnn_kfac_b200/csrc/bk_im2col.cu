// bk_im2col.cu — tensor-core operand staging for WIDE convolution factors (C*kh*kw + 1 > BK_SMALL_D_MAX).
//
// Reference (paths relative to /root/reference), models/curvatures.py:341-349 and :353-356:
//     forward = unfold(forward, k, padding, stride)            # [N, C*kh*kw, L]       fp32, written to HBM
//     forward = forward.contiguous().permute(1, 0, 2).view(C*kh*kw, -1)   # second fp32 copy
//     forward = cat([forward, ones(1, N*L)])                   # third copy
//     first   = forward @ forward^T / (N*L)
//     backward = g.permute(1, 0, 2, 3).view(O, -1);  second = backward @ backward^T / (N*H'*W')
// Here ONE pass reads the NCHW activations and writes the K-major bf16 (hi[, lo]) operand the tcgen05 SYRK
// consumes: T[r][col], r = (c * kh + i) * kw + j (unfold's row order), col = n * L + oh * OW + ow, plus the row
// of ones of the bias augmentation.  The fp32 patch matrix is never materialised: 2 (4) bytes written per patch
// element instead of 4 + 8 + 8 (+ the transposing staging pass of a dense operand).  Each input element is
// re-read kh * kw times, out of L1 / L2 (the working set of one output row is a few input rows).
// kh = kw = 1, stride 1, pad 0 is the [N, O, H'W'] -> [O, N*H'W'] regrouping of the output gradients.
#include "bk_common.cuh"
#include "bk_kernels.cuh"

namespace bk {

namespace {

struct Im2colShape {
  int n, c, h, w, kh, kw, ph, pw, sh, sw, oh, ow;
};

// grid: (column groups of 8 * 256, rows); each thread writes 8 consecutive columns of one row (16 B per part)
__global__ void __launch_bounds__(256)
im2col_split_kernel(const float* __restrict__ X, const __grid_constant__ Im2colShape s, float scale, int ones_row,
                    __nv_bfloat16* __restrict__ Thi, __nv_bfloat16* __restrict__ Tlo, long long ldt) {
  const int r = blockIdx.y;
  const long long col0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (col0 >= ldt) return;
  const int L = s.oh * s.ow;
  const long long cols = static_cast<long long>(s.n) * L;
  const int rows = s.c * s.kh * s.kw;
  float v[8];
  if (r == rows) {  // bias augmentation
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = (ones_row && col0 + u < cols) ? 1.f : 0.f;
  } else {
    const int c = r / (s.kh * s.kw);
    const int ij = r - c * (s.kh * s.kw);
    const int i = ij / s.kw, j = ij - i * s.kw;
    int n = static_cast<int>(col0 / L);
    int p = static_cast<int>(col0 - static_cast<long long>(n) * L);
    int oh = p / s.ow, ow = p - oh * s.ow;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      float x = 0.f;
      if (col0 + u < cols) {
        const int ih = oh * s.sh - s.ph + i, iw = ow * s.sw - s.pw + j;
        if (ih >= 0 && ih < s.h && iw >= 0 && iw < s.w)
          x = __ldg(X + ((static_cast<long long>(n) * s.c + c) * s.h + ih) * s.w + iw) * scale;
      }
      v[u] = x;
      if (++ow == s.ow) {
        ow = 0;
        if (++oh == s.oh) {
          oh = 0;
          ++n;
        }
      }
    }
  }
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    __nv_bfloat16 ah, al, bh, bl;
    split_bf16(v[2 * u], ah, al);
    split_bf16(v[2 * u + 1], bh, bl);
    const __nv_bfloat162 h2 = __halves2bfloat162(ah, bh), l2 = __halves2bfloat162(al, bl);
    hi[u] = *reinterpret_cast<const uint32_t*>(&h2);
    lo[u] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  const long long off = static_cast<long long>(r) * ldt + col0;
  *reinterpret_cast<uint4*>(Thi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  if (Tlo != nullptr) *reinterpret_cast<uint4*>(Tlo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

}  // namespace

int launch_im2col_split(const float* X, int n, int c, int h, int w, int kh, int kw, int ph, int pw, int sh,
                        int sw, float scale, int ones_row, __nv_bfloat16* Thi, __nv_bfloat16* Tlo,
                        long long ldt, cudaStream_t stream) {
  if (X == nullptr || Thi == nullptr || n <= 0 || c <= 0 || kh <= 0 || kw <= 0 || sh <= 0 || sw <= 0) return -2;
  Im2colShape s{n, c, h, w, kh, kw, ph, pw, sh, sw, (h + 2 * ph - kh) / sh + 1, (w + 2 * pw - kw) / sw + 1};
  if (s.oh <= 0 || s.ow <= 0) return -2;
  const long long cols = static_cast<long long>(n) * s.oh * s.ow;
  if (ldt < cols || (ldt & 7) != 0 || (reinterpret_cast<uintptr_t>(Thi) & 15) != 0 ||
      (Tlo != nullptr && (reinterpret_cast<uintptr_t>(Tlo) & 15) != 0))
    return -2;
  const int rows = c * kh * kw + (ones_row ? 1 : 0);
  if (rows > 65535) return -2;
  dim3 grid(static_cast<unsigned>((ldt / 8 + 255) / 256), static_cast<unsigned>(rows));
  im2col_split_kernel<<<grid, 256, 0, stream>>>(X, s, scale, ones_row, Thi, Tlo, ldt);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace bk
