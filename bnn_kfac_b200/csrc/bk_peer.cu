// bk_peer.cu — factor exchange over NVLink / NVSwitch PEER MEMORY (one process per GPU, CUDA IPC).
//
// The sharded inversion has two exchange steps around the per-owner inversion:
//   (1) every factor, summed over ranks and divided by the world size, must reach the rank that inverts it
//       (a reduce-scatter of symmetric matrices);
//   (2) every Cholesky factor must reach every rank (an all-gather of lower-triangular matrices).
// Through NCCL each is pack -> collective -> unpack (three passes, two of them HBM round trips of the whole
// payload on every rank).  Here the collective and the unpack are ONE kernel that reads the peers' packed
// tiles straight over NVLink:
//   (1) peer_tile_unpack with nsrc = world: the owner's CTAs load tile t of its factor from all `world` send
//       buffers (32 independent 128-byte-coalesced loads in flight per thread), add them in rank order (the sum
//       is the same whichever rank computes it), scale, and store the dense tile and its mirror;
//   (2) the same kernel with nsrc = 1: every rank pulls the owners' packed Cholesky factors and expands them
//       with a zero upper triangle.
// Packed layout ("tile-packed"): the 32 x 32 tiles of the lower triangle, tile (ti, tj <= ti) at index
// ti (ti + 1) / 2 + tj, 1024 floats each, row-major - every tile row is one aligned 128-byte line, whatever d.
//
// Cross-GPU ordering: flags in the exported buffers.  After its pack a rank runs peer_signal (system-scope fence,
// then one store of the epoch into its slot of every peer's flag array); consumers run peer_wait (spin until all
// slots reach the epoch, bounded by a time-out that raises an error flag instead of hanging the GPU) in front of
// the pulling kernel.  A second flag set ("done") tells a producer that its buffer may be overwritten.
//
// The buffers are plain cudaMalloc allocations exported with cudaIpcGetMemHandle; the handles travel through
// torch.distributed (plumbing), see bnn_kfac_b200/distributed.py:PeerExchange.
#include "bk_common.cuh"
#include "bk_kernels.cuh"

namespace bk {

namespace {

constexpr int kMaxTri = 16;
constexpr int kMaxPeers = 8;

struct TileTable {
  float* mat[kMaxTri];
  float* dst[kMaxTri];     // tile_pack with per-factor destinations (possibly peer memory); else nullptr
  long long ld[kMaxTri];
  long long off[kMaxTri];  // first packed float of the factor inside a source buffer
  int tile0[kMaxTri + 1];  // first CTA of the factor (tiles of all factors concatenated)
  int d[kMaxTri];
  int count;
};

struct SrcTable {
  const float* p[kMaxTri][kMaxPeers];  // per factor: the buffers that hold its tile-packed triangle
};

__device__ __forceinline__ void decode(const TileTable& t, int b, int& f, int& ti, int& tj) {
  f = 0;
  while (f + 1 < t.count && b >= t.tile0[f + 1]) ++f;
  const int r = b - t.tile0[f];
  ti = static_cast<int>((sqrtf(8.f * r + 1.f) - 1.f) * 0.5f);
  while ((ti + 1) * (ti + 2) / 2 <= r) ++ti;
  while (ti * (ti + 1) / 2 > r) --ti;
  tj = r - ti * (ti + 1) / 2;
}

// dense lower triangle -> tile-packed (entries outside the matrix or above the diagonal: zero).  kPackTiles tiles
// per CTA, all loads issued before the first store (one 4 KB tile per CTA ran at 0.41 of the HBM peak: 8385 CTAs of
// four loads per thread for one 4097-wide factor).
constexpr int kPackTiles = 4;
__global__ void __launch_bounds__(256)
tile_pack_kernel(const __grid_constant__ TileTable t, float* __restrict__ packed) {
  const int total = t.tile0[t.count];
  float v[kPackTiles][4];
  float* out[kPackTiles];
#pragma unroll
  for (int q = 0; q < kPackTiles; ++q) {
    const int b = blockIdx.x * kPackTiles + q;
    out[q] = nullptr;
    if (b >= total) continue;
    int f, ti, tj;
    decode(t, b, f, ti, tj);
    const int d = t.d[f];
    const float* m = t.mat[f];
    const long long ld = t.ld[f];
    out[q] = (t.dst[f] != nullptr ? t.dst[f] : packed + t.off[f]) + static_cast<long long>(b - t.tile0[f]) * 1024;
    const int j = tj * 32 + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = ti * 32 + threadIdx.y + 8 * k;
      v[q][k] = (i < d && j <= i) ? __ldcs(m + static_cast<long long>(i) * ld + j) : 0.f;
    }
  }
#pragma unroll
  for (int q = 0; q < kPackTiles; ++q) {
    if (out[q] == nullptr) continue;
#pragma unroll
    for (int k = 0; k < 4; ++k) out[q][(threadIdx.y + 8 * k) * 32 + threadIdx.x] = v[q][k];
  }
}

// tile-packed, summed over nsrc source buffers (local or peer memory) -> dense [d, ld], scaled; upper triangle =
// mirror (symmetric factors) or zero (Cholesky factors).  TPB tiles per CTA (all loads of all tiles first).
template <int NSRC, int TPB>
__global__ void __launch_bounds__(256)
peer_tile_unpack_kernel(const __grid_constant__ TileTable t, const __grid_constant__ SrcTable src, int nsrc,
                        float scale, int mirror) {
  __shared__ float tile[32][33];
  const int total = t.tile0[t.count];
  float v[TPB][4][NSRC];
  int fs[TPB], tis[TPB], tjs[TPB];
  // all loads first: TPB x NSRC x 4 independent 4-byte loads per thread (128 B per warp and load), L1 bypassed - the
  // lines live in another GPU's memory and were written since this SM last saw them
#pragma unroll
  for (int q = 0; q < TPB; ++q) {
    const int b = blockIdx.x * TPB + q;
    fs[q] = -1;
    if (b >= total) continue;
    decode(t, b, fs[q], tis[q], tjs[q]);
    const long long base = static_cast<long long>(b - t.tile0[fs[q]]) * 1024;
#pragma unroll
    for (int s = 0; s < NSRC; ++s)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        v[q][k][s] = (s < nsrc) ? __ldcg(src.p[fs[q]][s] + base + (threadIdx.y + 8 * k) * 32 + threadIdx.x) : 0.f;
  }
#pragma unroll
  for (int q = 0; q < TPB; ++q) {
    if (fs[q] < 0) continue;  // block-uniform
    const int f = fs[q], ti = tis[q], tj = tjs[q];
    const int d = t.d[f];
    float* m = t.mat[f];
    const long long ld = t.ld[f];
    const int j = tj * 32 + threadIdx.x;
    if (q > 0) __syncthreads();  // the previous tile's mirrored reads are done
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = threadIdx.y + 8 * k;
      const int i = ti * 32 + r;
      float a = v[q][k][0];
#pragma unroll
      for (int s = 1; s < NSRC; ++s) a += v[q][k][s];  // rank order: the same sum on whichever rank reduces
      a *= scale;
      if (i < d && j <= i) m[static_cast<long long>(i) * ld + j] = a;
      tile[r][threadIdx.x] = a;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = threadIdx.y + 8 * k;      // row inside the mirrored tile = column index j
      const int jj = tj * 32 + r;
      const int ii = ti * 32 + threadIdx.x;   // column inside the mirrored tile = row index i
      if (ii < d && jj < ii) m[static_cast<long long>(jj) * ld + ii] = mirror ? tile[threadIdx.x][r] : 0.f;
    }
  }
}

struct FlagTable {
  unsigned int* p[kMaxPeers];  // flag array (one slot per rank) in every rank's exported buffer
};

__global__ void peer_signal_kernel(const __grid_constant__ FlagTable flags, int world, int me, unsigned int epoch) {
  const int r = threadIdx.x;
  if (r >= world) return;
  __threadfence_system();  // everything this stream wrote before is visible system-wide before the flag is
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags.p[r] + me), "r"(epoch) : "memory");
}

__global__ void peer_wait_kernel(const unsigned int* __restrict__ mine, int world, unsigned int epoch,
                                 unsigned long long timeout_ns, int* __restrict__ err) {
  const int r = threadIdx.x;
  if (r >= world) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine + r) : "memory");
    if (static_cast<int>(v - epoch) >= 0) break;
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > timeout_ns) {  // a peer never arrived: report instead of spinning forever
      atomicExch(err, r + 1);
      break;
    }
    __nanosleep(200);
  }
}

// bring-up / measurement: plain copy with 4- or 16-byte accesses, `unroll` independent accesses per thread in flight;
// whether it is a pull or a push is decided by which of the two pointers is peer memory
template <typename V, int U>
__global__ void __launch_bounds__(256) peer_copy_kernel(V* __restrict__ dst, const V* __restrict__ src, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n; i += U * stride) {
    V v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldcg(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) dst[i + u * stride] = v[u];
  }
  for (; i < n; i += stride) dst[i] = __ldcg(src + i);
}

// offs == nullptr: factors back to back from float 0
int fill_table(TileTable& t, float* const* mats, const long long* lds, const int* dims, const long long* offs,
               int count) {
  if (count <= 0 || count > kMaxTri) return -2;
  long long off = 0;
  int tiles = 0;
  for (int k = 0; k < count; ++k) {
    if (dims[k] <= 0 || mats[k] == nullptr || lds[k] < dims[k]) return -2;
    const int T = (dims[k] + 31) / 32;
    t.mat[k] = mats[k];
    t.ld[k] = lds[k];
    t.d[k] = dims[k];
    t.off[k] = offs != nullptr ? offs[k] : off;
    t.tile0[k] = tiles;
    off += static_cast<long long>(T) * (T + 1) / 2 * 1024;
    tiles += T * (T + 1) / 2;
  }
  t.tile0[count] = tiles;
  t.count = count;
  return tiles;
}

}  // namespace

long long tile_packed_floats(const int* dims, int count) {
  long long n = 0;
  for (int k = 0; k < count; ++k) {
    const long long T = (dims[k] + 31) / 32;
    n += T * (T + 1) / 2 * 1024;
  }
  return n;
}

int launch_tile_pack(const float* const* mats, const long long* lds, const int* dims, const long long* offs,
                     int count, float* packed, cudaStream_t stream) {
  TileTable t{};
  const int tiles = fill_table(t, const_cast<float* const*>(mats), lds, dims, offs, count);
  if (tiles < 0 || packed == nullptr) return -2;
  tile_pack_kernel<<<(tiles + kPackTiles - 1) / kPackTiles, dim3(32, 8), 0, stream>>>(t, packed);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

// per-factor destinations: the pack IS the send when a destination is another rank's buffer (posted NVLink writes)
int launch_tile_pack_to(const float* const* mats, const long long* lds, const int* dims, float* const* dsts,
                        int count, cudaStream_t stream) {
  TileTable t{};
  const int tiles = fill_table(t, const_cast<float* const*>(mats), lds, dims, nullptr, count);
  if (tiles < 0 || dsts == nullptr) return -2;
  for (int k = 0; k < count; ++k) {
    if (dsts[k] == nullptr) return -2;
    t.dst[k] = dsts[k];
  }
  tile_pack_kernel<<<(tiles + kPackTiles - 1) / kPackTiles, dim3(32, 8), 0, stream>>>(t, nullptr);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_peer_tile_unpack(float* const* mats, const long long* lds, const int* dims, int count,
                            const float* const* srcs, int nsrc, float scale, int mirror, cudaStream_t stream) {
  TileTable t{};
  const int tiles = fill_table(t, mats, lds, dims, nullptr, count);
  if (tiles < 0 || nsrc < 1 || nsrc > kMaxPeers) return -2;
  SrcTable s{};
  for (int k = 0; k < count; ++k)
    for (int r = 0; r < nsrc; ++r) {
      if (srcs[k * nsrc + r] == nullptr) return -2;
      s.p[k][r] = srcs[k * nsrc + r];
    }
  const dim3 block(32, 8);
  // 16 - 32 loads in flight per thread whatever the number of sources
  if (nsrc == 1) peer_tile_unpack_kernel<1, 4><<<(tiles + 3) / 4, block, 0, stream>>>(t, s, nsrc, scale, mirror);
  else if (nsrc == 2) peer_tile_unpack_kernel<2, 4><<<(tiles + 3) / 4, block, 0, stream>>>(t, s, nsrc, scale, mirror);
  else if (nsrc <= 4) peer_tile_unpack_kernel<4, 2><<<(tiles + 1) / 2, block, 0, stream>>>(t, s, nsrc, scale, mirror);
  else peer_tile_unpack_kernel<8, 1><<<tiles, block, 0, stream>>>(t, s, nsrc, scale, mirror);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_peer_copy(void* dst, const void* src, long long bytes, int vec_bytes, int ctas, cudaStream_t stream) {
  if (dst == nullptr || src == nullptr || bytes <= 0 || bytes % 16 != 0 || ctas <= 0) return -2;
  if (vec_bytes == 16)
    peer_copy_kernel<float4, 8><<<ctas, 256, 0, stream>>>(static_cast<float4*>(dst), static_cast<const float4*>(src),
                                                        bytes / 16);
  else
    peer_copy_kernel<float, 8><<<ctas, 256, 0, stream>>>(static_cast<float*>(dst), static_cast<const float*>(src),
                                                       bytes / 4);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_peer_signal(unsigned int* const* flags, int world, int me, unsigned int epoch, cudaStream_t stream) {
  if (world < 1 || world > kMaxPeers || me < 0 || me >= world) return -2;
  FlagTable f{};
  for (int r = 0; r < world; ++r) {
    if (flags[r] == nullptr) return -2;
    f.p[r] = flags[r];
  }
  peer_signal_kernel<<<1, 32, 0, stream>>>(f, world, me, epoch);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int launch_peer_wait(const unsigned int* mine, int world, unsigned int epoch, double timeout_s, int* err,
                     cudaStream_t stream) {
  if (world < 1 || world > kMaxPeers || mine == nullptr || err == nullptr) return -2;
  const unsigned long long ns = static_cast<unsigned long long>((timeout_s > 0 ? timeout_s : 5.0) * 1e9);
  peer_wait_kernel<<<1, 32, 0, stream>>>(mine, world, epoch, ns, err);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace bk
