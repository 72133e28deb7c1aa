// bk_umma_gemm.cu — persistent, warp-specialised tcgen05 contraction core for sm_100a.
//
//   D[b][m][n] = sum_k A[b][m][k] * B[b][n][k]
//
// Two instantiations of one kernel (template parameter CG = tcgen05 cta_group):
//   CG = 1  one CTA per SM, 128 x 256 output tile, 4-stage ring of (A 128x64 + B 256x64) bf16.
//           Shared memory carries 48 KB in + 48 KB out per k-block = 192 B/cycle at the tensor-core
//           floor against 128 B/cycle available, so this variant tops out near 0.6 of the floor; it
//           serves the small / skinny problems.
//   CG = 2  CTA PAIR (cluster of 2 on one TPC), 256 x 256 output tile: each CTA stages its own 128
//           rows of A and only HALF of the B rows (32 KB per stage, 6-stage ring), the leader CTA
//           issues tcgen05.mma.cta_group::2 (M = 256) and releases both CTAs' ring slots with a
//           multicast commit; each CTA's epilogue drains its own 128 accumulator rows.  Halves the
//           shared-memory and L2 traffic per flop: this is the throughput path for wide layers.
// Roles (256 threads per CTA, persistent, static round-robin tile schedule over clusters):
//   warp 0   TMA producer   : cp.async.bulk.tensor (128B swizzle) into the stage ring
//   warp 1   MMA issuer     : one elected lane issues 4 x tcgen05.mma (M x 256 x 16) per stage
//   warp 2   TMEM allocator : 512 columns = two M x 256 fp32 accumulators (double buffered)
//   warps 4-7 epilogue      : tcgen05.ld -> smem transpose -> coalesced fused epilogue
//                             (alpha/beta/bias/relu, fp32 and/or split-bf16 outputs,
//                              SYRK lower-triangle masking + mirrored write)
// Pipelines: smem full/empty ring, TMEM full/empty (2 stages).
//
// Reference semantics served by this core (all /root/reference paths relative to the repo root):
//   models/curvatures.py:349,356  first/second Kronecker factor  (SYRK mode, `state +=`)
//   models/curvatures.py:404-405  matrix-normal sample L_A z L_G^T (triangular-A mode)
//   models/wrapper.py:35-44       forward over weight samples      (bias + relu epilogue, batched)
#include "bk_common.cuh"
#include "bk_umma_gemm.cuh"

#include <mutex>

namespace bk {

namespace {

constexpr int BM = 128;  // accumulator rows per CTA (TMEM lanes)
constexpr int BN = 256;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UK = 16;  // K per tcgen05.mma for 16-bit inputs
constexpr int kAccStages = 2;
constexpr int kSameAB = 1 << 20;  // internal flag: the A and B operands are the same buffers (SYRK)
constexpr int kThreads = 256;
constexpr int kEpiWarp0 = 4;
constexpr uint32_t kBytesA = BM * BK * 2;  // 16 KiB
// per epilogue warp: two 32x32 fp32 tiles (direct + mirrored) in 128B-swizzled TMA layout, 1024 B
// aligned; the register-path epilogue uses the first 32x33 floats of it as a padded transpose buffer
constexpr uint32_t kEpiStageBytes = 2 * 32 * 32 * 4;
constexpr uint32_t kTmemCols = kAccStages * BN;  // 512

template <int CG>
struct Cfg {
  static constexpr int kTileM = BM * CG;                      // output rows per cluster tile
  static constexpr int kRowsB = BN / CG;                      // B rows staged by each CTA
  static constexpr uint32_t kBytesB = kRowsB * BK * 2;        // 32 / 16 KiB
  static constexpr uint32_t kStageBytes = kBytesA + kBytesB;  // 48 / 32 KiB
  static constexpr int kStages = CG == 1 ? 4 : 6;
  static constexpr uint32_t kSmemBytes =
      kStages * kStageBytes + 4 * kEpiStageBytes + 256 /*barriers*/ + 1024 /*alignment slack*/;
  // lower-only TMA epilogue (no mirrored tile): half the epilogue staging.  For CTA pairs this leaves
  // ~22 KB of the SM's shared memory free, enough for a CTA of the operand-staging kernel to be
  // co-resident (the staging of the next factors then overlaps this kernel, see bk_api.cu).
  static constexpr uint32_t kSmemBytesLower =
      kStages * kStageBytes + 4 * (kEpiStageBytes / 2) + 256 + 1024;
};

struct KParams {
  int M, N, K, batch, nparts, flags;
  int tiles_m, tiles_n, tiles_per_batch, num_tiles;
  int a_bmul, b_bmul, c_bmul;
  int tri_koff;  // kTriB: B[n][k] == 0 for k > tri_koff + n
  float alpha, beta;
  float* C;
  long long ldc, strideC;
  const float* bias;
  long long strideBias;
  __nv_bfloat16* Ohi;
  __nv_bfloat16* Olo;
  __nv_bfloat16* Olo2;
  long long ldo, strideO;
};

__host__ __device__ inline int syrk_row_tiles(int mi, int tiles_n, int tile_m) {
  int c = (mi * tile_m + tile_m - 1) / BN + 1;
  return c < tiles_n ? c : tiles_n;
}

struct Tile {
  int b, m0, n0, kb0, nkb;  // k-blocks [kb0, kb0 + nkb) carry non-zero operand data
};

template <int CG>
__device__ __forceinline__ Tile decode_tile(const KParams& p, int t) {
  constexpr int kTileM = Cfg<CG>::kTileM;
  Tile r;
  r.b = t / p.tiles_per_batch;
  int rem = t - r.b * p.tiles_per_batch;
  int mi, nj;
  if (p.flags & kSyrkLower) {
    mi = 0;
    int c = syrk_row_tiles(0, p.tiles_n, kTileM);
    while (rem >= c) {
      rem -= c;
      ++mi;
      c = syrk_row_tiles(mi, p.tiles_n, kTileM);
    }
    nj = rem;
  } else {
    // heaviest row blocks first when the k extent grows with m (triangular A)
    mi = rem / p.tiles_n;
    nj = rem - mi * p.tiles_n;
    if (p.flags & kTriA) mi = p.tiles_m - 1 - mi;
    if (p.flags & kTriB) nj = p.tiles_n - 1 - nj;
  }
  r.m0 = mi * kTileM;
  r.n0 = nj * BN;
  int kend = p.K;
  if (p.flags & kTriA) kend = min(kend, r.m0 + kTileM);
  if (p.flags & kTriB) kend = min(kend, p.tri_koff + r.n0 + BN);
  r.kb0 = (p.flags & kTriBUpper) ? min(r.n0, p.K) / BK : 0;
  r.nkb = max((kend + BK - 1) / BK - r.kb0, 0);
  return r;
}

// ---- grouped SYRK: several independent factor updates C_g += alpha_g X_g^T X_g in ONE persistent
// launch (all layers of a KFAC.update): one tile list over all problems, so wave quantisation, the
// pipeline prologue and the exposed last-tile epilogue are paid once per update instead of once per
// factor.  CTA pairs, 256 x 256 lower-triangle tiles, TMA-reduce epilogue.
constexpr int kMaxGroup = 8;
struct GroupMaps {
  CUtensorMap a0[kMaxGroup];  // X^T hi  (A and B operand: both are 128-row boxes for CTA pairs)
  CUtensorMap a1[kMaxGroup];  // X^T lo  (bf16x3 only)
  CUtensorMap c[kMaxGroup];   // fp32 state
};
struct GroupParams {
  int count, nparts;
  int mirror;  // 1: write the transposed tile as well (full symmetric state); 0: lower triangle only
  // Schedule: tiles [0, full_tiles) are taken whole, round-robin over the P CTA pairs, so the pairs that run
  // at the same time work on neighbouring tiles and share operand panels in L2.  Default: full_tiles = all.
  // Optional (set_syrk_tuning bit 1): the last T mod P tiles, which leave pairs idle for a tile time (136 tiles
  // of a 4096-wide factor on 74 pairs: 1.84 waves), are scheduled stream-K — their main-loop iterations
  // (iters[g] = k-blocks x precision passes per tile of problem g), concatenated, are cut into P equal ranges
  // (multiples of kWorkGrain); a tile cut by a boundary is finished by two pairs, both adding their partial
  // sums to C through the TMA reduction.  MEASURED SLOWER (profiles/r02_syrk_ab.md): 7 factors 364.6 vs
  // 344.2 us, 1 factor 75.1 vs 65.0 us — kept as a switch, off.  Stream-K over the WHOLE launch is worse
  // still: every pair then walks its own region of the triangle, the L2 hit rate falls from 73 % to 53 %, and
  // a 3-factor launch reads 708 MB instead of 100 MB from DRAM.
  int iters[kMaxGroup];
  int full_tiles, tail_work;
  int no_dedup;
  int tile_begin[kMaxGroup + 1];
  int M[kMaxGroup], K[kMaxGroup];
  float alpha[kMaxGroup];
};

__device__ __forceinline__ Tile decode_group_tile(const GroupParams& gp, int t, int& g) {
  g = 0;
  while (t >= gp.tile_begin[g + 1]) ++g;
  int rem = t - gp.tile_begin[g];
  int mi = 0;
  while (rem >= mi + 1) {  // row mi of the 256-tile lower triangle holds mi + 1 tiles
    rem -= mi + 1;
    ++mi;
  }
  Tile r;
  r.b = 0;
  r.m0 = mi * 256;
  r.n0 = rem * BN;
  r.kb0 = 0;
  r.nkb = (gp.K[g] + BK - 1) / BK;
  return r;
}

// Operand maps of a single problem: up to three bf16 parts per operand (hi, lo, lo2) + the fp32 output.
struct OpMaps {
  CUtensorMap a[3];
  CUtensorMap b[3];
  CUtensorMap c;
};

// Split-precision schedules: which (A part, B part) product each pass accumulates.
//   1 pass  : hi*hi                                   (bf16)
//   3 passes: + hi*lo, lo*hi                          (bf16x3, ~2^-17 per product)
//   6 passes: + hi*lo2, lo2*hi, lo*lo                 (bf16x6, 24 mantissa bits: fp32-class)
__device__ __forceinline__ void part_pair(int part, int& ia, int& ib) {
  ia = (part == 2) ? 1 : (part == 4) ? 2 : (part == 5) ? 1 : 0;
  ib = (part == 1) ? 1 : (part == 3) ? 2 : (part == 5) ? 1 : 0;
}

constexpr int kWorkGrain = 8;  // stream-K boundaries are multiples of this many main-loop iterations

// Per-tile view of "which problem": tensor maps, extents and scale.
struct Work {
  Tile tl;
  const CUtensorMap *a[3], *b[3], *c;
  int M, N;
  float alpha;
};

template <int CG, bool kGrouped>
__device__ __forceinline__ Work get_work(int t, const KParams& p, const OpMaps* om,
                                         const GroupMaps* gm, const GroupParams* gp) {
  Work w;
  if (kGrouped) {
    int g;
    w.tl = decode_group_tile(*gp, t, g);
    w.a[0] = w.b[0] = &gm->a0[g];
    w.a[1] = w.b[1] = &gm->a1[g];
    w.a[2] = w.b[2] = &gm->a1[g];
    w.c = &gm->c[g];
    w.M = w.N = gp->M[g];
    w.alpha = gp->alpha[g];
  } else {
    w.tl = decode_tile<CG>(p, t);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      w.a[i] = &om->a[i];
      w.b[i] = &om->b[i];
    }
    w.c = &om->c;
    w.M = p.M;
    w.N = p.N;
    w.alpha = p.alpha;
  }
  return w;
}

// One unit of work of a role loop: main-loop iterations [it0, it1) of one output tile.
struct Seg {
  Work wk;
  int it0, it1;
};

// Enumerates the segments of this cluster: whole tiles round-robin (single problems) or its stream-K range
// (grouped SYRK).  Every role (producer, MMA, epilogue) walks its own copy and sees the same sequence.
template <int CG, bool kGrouped>
struct WorkIter {
  const KParams& p;
  const OpMaps* om;
  const GroupMaps* gm;
  const GroupParams* gp;
  int nparts;
  int t, step, num_tiles;  // tile striding
  int u, u_end;            // stream-K cursor over the tail tiles' iterations
  __device__ __forceinline__ WorkIter(const KParams& p_, const OpMaps* om_, const GroupMaps* gm_,
                                      const GroupParams* gp_, int nparts_, int first, int step_,
                                      int num_tiles_)
      : p(p_), om(om_), gm(gm_), gp(gp_), nparts(nparts_), t(first), step(step_), num_tiles(num_tiles_) {
    u = u_end = 0;
    if (kGrouped) {
      num_tiles = gp->full_tiles;
      const long long W = gp->tail_work;
      long long a = W * first / step, b = W * (first + 1) / step;
      a -= a % kWorkGrain;
      if (first + 1 < step) b -= b % kWorkGrain;
      u = static_cast<int>(a);
      u_end = static_cast<int>(b);
    }
  }
  __device__ __forceinline__ int group_tile_iters(int tile) const {
    int g = 0;
    while (tile >= gp->tile_begin[g + 1]) ++g;
    return gp->iters[g];
  }
  __device__ __forceinline__ bool next(Seg& s) {
    if (t < num_tiles) {
      s.wk = get_work<CG, kGrouped>(t, p, om, gm, gp);
      s.it0 = 0;
      s.it1 = kGrouped ? group_tile_iters(t) : s.wk.tl.nkb * nparts;
      t += step;
      return true;
    }
    if (!kGrouped || u >= u_end) return false;
    // tail: find the tile that holds iteration u of the concatenated tail work
    int tile = gp->full_tiles, acc = 0, iters = group_tile_iters(tile);
    while (u >= acc + iters) {
      acc += iters;
      ++tile;
      iters = group_tile_iters(tile);
    }
    s.it0 = u - acc;
    const int n = min(iters - s.it0, u_end - u);
    s.it1 = s.it0 + n;
    u += n;
    s.wk = get_work<CG, true>(tile, p, om, gm, gp);
    return true;
  }
};

// kMN (grouped SYRK only): the operand is the ROW-major activation matrix X [n, d] itself (bf16, feature index
// contiguous) — "MN-major" for X^T X.  TMA boxes of 64 features x 64 samples land as [64 k-rows][128 B]; a
// 128-row operand tile is two such boxes 8 KB apart.  The UMMA descriptor then reads: 8-k-row swizzle atoms
// 1024 B apart (SBO), 64-feature chunks 8192 B apart (LBO), and one K = 16 MMA step advances two atoms
// (2048 B).  No transposed / converted copy of the activations exists on this path.
template <int CG, bool kTmaEpi, bool kGrouped, bool kMN = false>
__device__ __forceinline__ void gemm_body(const OpMaps* om, const KParams& p, const GroupMaps* gm,
                                          const GroupParams* gp) {
  const int num_tiles = kGrouped ? gp->tile_begin[gp->count] : p.num_tiles;
  const int nparts = kGrouped ? gp->nparts : p.nparts;
  const int flags = kGrouped ? (kSyrkLower | (gp->mirror ? kMirror : 0)) : p.flags;
  // per-warp epilogue staging: direct + mirrored 32 x 32 tile, or the direct tile only
  const uint32_t epi_bytes = (kTmaEpi && (flags & kSyrkLower) && !(flags & kMirror)) ? kEpiStageBytes / 2
                                                                                    : kEpiStageBytes;
  // SYRK tile on the diagonal of a CTA-pair schedule: the A rows and the B rows staged by each CTA are
  // the SAME 128 rows of X^T, so only A is loaded and the B descriptor points at it (saves 1/17 of the
  // L2 -> SM operand traffic of a 4096-wide factor)
  const bool diag_dedup = CG == 2 && (flags & kSyrkLower) != 0 &&
                          (kGrouped ? gp->no_dedup == 0 : (flags & kSameAB) != 0);
  constexpr int kStages = Cfg<CG>::kStages;
  constexpr uint32_t kStageBytes = Cfg<CG>::kStageBytes;
  constexpr int kTileM = Cfg<CG>::kTileM;
  constexpr int kRowsB = Cfg<CG>::kRowsB;
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled operand tiles need 1024 B alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_epi = smem + kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + 4 * epi_bytes);
  uint64_t* full_bar = bars;                       // [kStages]   (CG=2: the leader's are used)
  uint64_t* empty_bar = bars + kStages;            // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;        // [kAccStages]
  uint64_t* tempty_bar = tfull_bar + kAccStages;   // [kAccStages] (CG=2: the leader's are used)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + kAccStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CTA pair: rank 0 (leader) issues the MMAs; both CTAs load, and drain their own 128 rows.
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int first_tile = (CG == 2) ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_step = (CG == 2) ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0 && !kGrouped) {
    tma_prefetch_desc(&om->a[0]);
    tma_prefetch_desc(&om->b[0]);
    if (kTmaEpi) tma_prefetch_desc(&om->c);
    if (nparts > 1) {
      tma_prefetch_desc(&om->a[1]);
      tma_prefetch_desc(&om->b[1]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4 * CG);  // one arrival per epilogue warp (of both CTAs)
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (CG == 2) tmem_alloc_2sm(tmem_ptr, kTmemCols);
    else tmem_alloc(tmem_ptr, kTmemCols);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();  // peer barriers initialised before any remote arrive / TMA
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // The two single-thread role loops below are the critical path of the kernel: one iteration per 64-wide
  // k-block, i.e. per 512 tensor-core cycles.  (Measured, profiles/r02_syrk_role_loops.md: with ~110 SASS
  // instructions per iteration — an integer division for the precision pass, operand-map lookups through
  // local memory, descriptor re-encoding — the producer and the MMA issuer were ISSUE-bound and held the
  // tensor pipe at 56 % active.)  Everything that is constant per tile or per precision pass is hoisted; the
  // hot loops work on raw shared-memory addresses.
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full_base = smem_u32(full_bar);
  const uint32_t empty_base = smem_u32(empty_bar);
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // The WHOLE warp walks the loop (warp-uniform control flow and addresses stay in uniform registers);
    // one elected lane issues.  Under `if (lane == 0)` every operand of every TMA / MMA instruction went
    // through an ELECT + R2UR.BROADCAST waterfall: ~20 extra instructions per issued instruction.
    {
      uint32_t stage = 0, phase = 0;
      // CTA pairs: both CTAs' bytes complete on the LEADER's full barrier, which expects 2 stages' worth
      const uint32_t lfull_base = (CG == 2) ? mapa_u32(full_base, 0) : full_base;
      WorkIter<CG, kGrouped> wi(p, om, gm, gp, nparts, first_tile, tile_step, num_tiles);
      Seg sg;
      while (wi.next(sg)) {
        const Tile tl = sg.wk.tl;
        // this CTA's slice of the cluster tile: its own 128 rows of A, its share of the B rows
        const int a_row = tl.m0 + static_cast<int>(cta_rank) * BM;
        const int b_row = tl.n0 + static_cast<int>(cta_rank) * kRowsB;
        const int a_b = tl.b * p.a_bmul, b_b = tl.b * p.b_bmul;
        const bool diag_tile = diag_dedup && tl.m0 == tl.n0;
        int it = sg.it0;
        int part = it / tl.nkb;            // once per segment
        int kb = it - part * tl.nkb;
        while (it < sg.it1) {
          int ia, ib;
          part_pair(part, ia, ib);
          const CUtensorMap* ma = ia == 0 ? sg.wk.a[0] : (ia == 1 ? sg.wk.a[1] : sg.wk.a[2]);
          const CUtensorMap* mb = ib == 0 ? sg.wk.b[0] : (ib == 1 ? sg.wk.b[1] : sg.wk.b[2]);
          const bool same = diag_tile && ia == ib;
          const uint32_t tx = (CG == 2) ? (same ? 2 * kBytesA : 2 * kStageBytes) : kStageBytes;
          const int kend = min(tl.nkb, kb + (sg.it1 - it));
          it += kend - kb;
          int kcoord = (tl.kb0 + kb) * BK;
          for (; kb < kend; ++kb, kcoord += BK) {     // ---- hot loop
            mbar_wait_addr(empty_base + stage * 8u, phase ^ 1u);
            const uint32_t sa = smem_base + stage * kStageBytes;
            const uint32_t fb = lfull_base + stage * 8u;
            if (elect_one()) {
              if (kMN) {
                // X [n, d]: coordinates (feature, sample); two 64-feature boxes per 128-row operand tile
                if (cta_rank == 0) mbar_arrive_expect_tx_addr(full_base + stage * 8u, tx);
                tma_load_3d_2sm_addr(sa, ma, fb, a_row, kcoord, 0);
                tma_load_3d_2sm_addr(sa + kBytesA / 2, ma, fb, a_row + 64, kcoord, 0);
                if (!same) {
                  tma_load_3d_2sm_addr(sa + kBytesA, mb, fb, b_row, kcoord, 0);
                  tma_load_3d_2sm_addr(sa + kBytesA + kBytesA / 2, mb, fb, b_row + 64, kcoord, 0);
                }
              } else if (CG == 2) {
                if (cta_rank == 0) mbar_arrive_expect_tx_addr(full_base + stage * 8u, tx);
                tma_load_3d_2sm_addr(sa, ma, fb, kcoord, a_row, a_b);
                if (!same) tma_load_3d_2sm_addr(sa + kBytesA, mb, fb, kcoord, b_row, b_b);
              } else {
                mbar_arrive_expect_tx_addr(fb, tx);
                tma_load_3d_addr(sa, ma, fb, kcoord, a_row, a_b);
                tma_load_3d_addr(sa + kBytesA, mb, fb, kcoord, b_row, b_b);
              }
            }
            __syncwarp();
            if (++stage == static_cast<uint32_t>(kStages)) {
              stage = 0;
              phase ^= 1u;
            }
          }
          kb = 0;
          ++part;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, see above)
    if (cta_rank == 0) {
      // MN-major operands: bits 15 / 16 of the instruction descriptor
      constexpr uint32_t idesc = umma_idesc_bf16_f32(kTileM, BN) | (kMN ? ((1u << 15) | (1u << 16)) : 0u);
      constexpr uint64_t kStep = kMN ? 128 : 2;  // descriptor address advance per K = 16 MMA (>> 4 units)
      uint32_t stage = 0, phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      // descriptor of stage 0's A tile; stage s and the B tile are plain offsets in the (address >> 4) field
      const uint64_t desc0 = kMN ? umma_smem_desc_mn_sw128(smem_base) : umma_smem_desc_k_sw128(smem_base);
      const uint32_t tfull_base = smem_u32(tfull_bar);
      WorkIter<CG, kGrouped> wi(p, om, gm, gp, nparts, first_tile, tile_step, num_tiles);
      Seg sg;
      while (wi.next(sg)) {
        const Tile tl = sg.wk.tl;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
        const bool diag_tile = diag_dedup && tl.m0 == tl.n0;
        uint32_t accumulate = 0;           // the first MMA of a segment overwrites the accumulator
        int it = sg.it0;
        int part = it / tl.nkb;
        int kb = it - part * tl.nkb;
        while (it < sg.it1) {
          int ia, ib;
          part_pair(part, ia, ib);
          const uint64_t b_off = (diag_tile && ia == ib) ? 0ull : static_cast<uint64_t>(kBytesA >> 4);
          const int kend = min(tl.nkb, kb + (sg.it1 - it));
          it += kend - kb;
          for (; kb < kend; ++kb) {                   // ---- hot loop
            mbar_wait_addr(full_base + stage * 8u, phase);
            tc_fence_after();
            const uint64_t da = desc0 + static_cast<uint64_t>(stage * (kStageBytes >> 4));
            const uint64_t db = da + b_off;
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < BK / UK; ++k) {
                // advance 16 elements (32 B) along K inside the swizzle row: +2 in the >>4 address field
                if (CG == 2)
                  umma_bf16_ss_2sm(tmem_d, da + static_cast<uint64_t>(k) * kStep,
                                   db + static_cast<uint64_t>(k) * kStep, idesc,
                                   (accumulate | static_cast<uint32_t>(k)) != 0u ? 1u : 0u);
                else
                  umma_bf16_ss(tmem_d, da + static_cast<uint64_t>(k) * kStep,
                               db + static_cast<uint64_t>(k) * kStep, idesc,
                               (accumulate | static_cast<uint32_t>(k)) != 0u ? 1u : 0u);
              }
              // frees the smem slot (in both CTAs of a pair) when these MMAs retire
              if (CG == 2) umma_commit_2sm_addr(empty_base + stage * 8u, 3);
              else umma_commit_addr(empty_base + stage * 8u);
            }
            __syncwarp();
            accumulate = 1;
            if (++stage == static_cast<uint32_t>(kStages)) {
              stage = 0;
              phase ^= 1u;
            }
          }
          kb = 0;
          ++part;
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (elect_one()) {
          if (CG == 2) umma_commit_2sm_addr(tfull_base + static_cast<uint32_t>(acc) * 8u, 3);
          else umma_commit_addr(tfull_base + static_cast<uint32_t>(acc) * 8u);
        }
        __syncwarp();
        if (++acc == kAccStages) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp - kEpiWarp0;  // TMEM lane quadrant == warp % 4
    float* stg = reinterpret_cast<float*>(smem_epi + q * epi_bytes);
    const bool syrk = (flags & kSyrkLower) != 0;
    const bool mirror = (flags & kMirror) != 0;
    const bool relu = (flags & kRelu) != 0;
    const bool use_beta = p.beta != 0.f;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t tempty_leader0 = (CG == 2) ? mapa_u32(smem_u32(&tempty_bar[0]), 0) : 0u;
    WorkIter<CG, kGrouped> wi(p, om, gm, gp, nparts, first_tile, tile_step, num_tiles);
    Seg sg;
    while (wi.next(sg)) {
      const Work& wk = sg.wk;
      const Tile tl = wk.tl;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      // first row of this warp's 32-row band (the peer CTA of a pair holds rows 128..255)
      const int r0 = tl.m0 + static_cast<int>(cta_rank) * BM + q * 32;
      const int my_row = r0 + lane;         // row held by this thread in TMEM layout
      float* Cb = p.C ? p.C + static_cast<long long>(tl.b) * p.strideC : nullptr;
      const float* biasb = p.bias ? p.bias + static_cast<long long>(tl.b) * p.strideBias : nullptr;
      __nv_bfloat16* Ohb = p.Ohi ? p.Ohi + static_cast<long long>(tl.b) * p.strideO : nullptr;
      __nv_bfloat16* Olb = p.Olo ? p.Olo + static_cast<long long>(tl.b) * p.strideO : nullptr;
      __nv_bfloat16* Ol2b = p.Olo2 ? p.Olo2 + static_cast<long long>(tl.b) * p.strideO : nullptr;
      const uint32_t taddr0 =
          tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        const int c0 = tl.n0 + ch * 32;
        float v[32];
        tmem_ld_32x32(taddr0 + static_cast<uint32_t>(ch * 32), v);
        tmem_ld_wait();
        if (ch == BN / 32 - 1) {
          // all TMEM reads of this accumulator are done: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) mbar_arrive_cluster(tempty_leader0 + static_cast<uint32_t>(acc) * 8u);
            else mbar_arrive(&tempty_bar[acc]);
          }
        }
        // band entirely outside the matrix, or (SYRK) entirely above the diagonal: nothing to do
        if (r0 >= wk.M || c0 >= wk.N) continue;
        if (syrk && c0 > r0 + 31) continue;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= wk.alpha;

        if (kTmaEpi) {
          // ---- accumulate through the TMA: C[tile] += alpha*acc as an L2-side reduction.  The warp
          // never loads C, so nothing here waits on DRAM; out-of-range rows / columns are clipped by
          // the tensor map.  SYRK: chunks on the diagonal (c0 == r0) are masked to c <= r (direct)
          // and c < r (mirrored) by writing zeros; chunks below it are stored whole, twice.
          uint8_t* sd = reinterpret_cast<uint8_t*>(stg);  // direct tile   [32 rows][32 cols]
          uint8_t* sm = sd + 32 * 32 * 4;                 // mirrored tile [32 cols][32 rows]
          const bool diag = syrk && c0 == r0;
          const bool do_mirror = syrk && mirror;
          // (bulk groups belong to the issuing thread: elect.sync with a full mask always elects the same lane)
          if (elect_one()) tma_store_wait_read0();  // previous chunk's stores have drained the buffer
          __syncwarp();
          // direct: lane = row, eight 16 B chunks per row at swizzled position (chunk ^ (row & 7))
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            float4 o;
            o.x = (!diag || c4 * 4 + 0 <= lane) ? v[c4 * 4 + 0] : 0.f;
            o.y = (!diag || c4 * 4 + 1 <= lane) ? v[c4 * 4 + 1] : 0.f;
            o.z = (!diag || c4 * 4 + 2 <= lane) ? v[c4 * 4 + 2] : 0.f;
            o.w = (!diag || c4 * 4 + 3 <= lane) ? v[c4 * 4 + 3] : 0.f;
            *reinterpret_cast<float4*>(sd + lane * 128 + ((c4 ^ (lane & 7)) << 4)) = o;
          }
          if (do_mirror) {
            // mirrored: tile row j = original column, element (j, lane); lanes -> consecutive floats
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float o = (!diag || j < lane) ? v[j] : 0.f;
              *reinterpret_cast<float*>(sm + j * 128 + ((((lane >> 2) ^ (j & 7)) << 4) |
                                                        ((lane & 3) << 2))) = o;
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {
            tma_reduce_add_3d(wk.c, sd, c0, r0, tl.b * p.c_bmul);
            if (do_mirror) tma_reduce_add_3d(wk.c, sm, r0, c0, tl.b * p.c_bmul);
            tma_store_commit();
          }
          continue;
        }

        // mirrored (transposed) write straight from registers: lanes = consecutive columns of
        // the transposed block -> 128 B coalesced per instruction.
        // All read-modify-writes below load a whole batch of 32 old values before the first store,
        // so 32 independent loads are in flight per thread (a load-store-load chain would expose
        // the full DRAM latency 32 times per chunk).
        if (syrk && mirror && Cb != nullptr && my_row < p.M) {
          const int jmax = my_row - c0;  // columns c0 + j < my_row (strictly below the diagonal)
          float* dst0 = Cb + static_cast<long long>(c0) * p.ldc + my_row;
          float old[32];
          if (use_beta) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              old[j] = (j < jmax) ? __ldcg(dst0 + static_cast<long long>(j) * p.ldc) : 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) old[j] = fmaf(p.beta, old[j], v[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) old[j] = v[j];
          }
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < jmax) __stcg(dst0 + static_cast<long long>(j) * p.ldc, old[j]);
        }
        // direct write: transpose through smem so that lanes = consecutive columns.
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = v[j];
        __syncwarp();
        const int gc = c0 + lane;
        const bool col_ok = gc < p.N;
        const float bv = (biasb != nullptr && col_ok) ? biasb[gc] : 0.f;
        // rows rr in [rlo, rhi) of this band are written by this lane
        const int rhi = col_ok ? min(32, p.M - r0) : 0;
        const int rlo = syrk ? max(0, gc - r0) : 0;
        float x[32];
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) x[rr] = stg[rr * 33 + lane] + bv;
        __syncwarp();
        if (Cb != nullptr) {
          float* dst0 = Cb + static_cast<long long>(r0) * p.ldc + gc;
          if (use_beta) {
            float old[32];
#pragma unroll
            for (int rr = 0; rr < 32; ++rr)
              old[rr] = (rr >= rlo && rr < rhi) ? __ldcg(dst0 + static_cast<long long>(rr) * p.ldc)
                                                : 0.f;
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) x[rr] = fmaf(p.beta, old[rr], x[rr]);
          }
          if (relu) {
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) x[rr] = fmaxf(x[rr], 0.f);
          }
#pragma unroll
          for (int rr = 0; rr < 32; ++rr)
            if (rr >= rlo && rr < rhi) __stcg(dst0 + static_cast<long long>(rr) * p.ldc, x[rr]);
        } else if (relu) {
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) x[rr] = fmaxf(x[rr], 0.f);
        }
        if (Ohb != nullptr) {
          __nv_bfloat16* o0 = Ohb + static_cast<long long>(r0) * p.ldo + gc;
          __nv_bfloat16* l0 = Olb != nullptr ? Olb + static_cast<long long>(r0) * p.ldo + gc : nullptr;
          __nv_bfloat16* m0 = (Ol2b != nullptr && l0 != nullptr)
                                  ? Ol2b + static_cast<long long>(r0) * p.ldo + gc : nullptr;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            if (rr >= rlo && rr < rhi) {
              const __nv_bfloat16 h = __float2bfloat16_rn(x[rr]);
              o0[static_cast<long long>(rr) * p.ldo] = h;
              if (l0 != nullptr) {
                const float r1 = x[rr] - __bfloat162float(h);
                const __nv_bfloat16 l = __float2bfloat16_rn(r1);
                l0[static_cast<long long>(rr) * p.ldo] = l;
                if (m0 != nullptr)
                  m0[static_cast<long long>(rr) * p.ldo] = __float2bfloat16_rn(r1 - __bfloat162float(l));
              }
            }
          }
        }
      }
      if (++acc == kAccStages) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  if (kTmaEpi && warp >= kEpiWarp0 && elect_one()) tma_store_wait_read0();  // smem outlives stores
  tc_fence_before();
  if (CG == 2) cluster_sync_all();  // the peer may still be reading its TMEM / receiving commits
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2sm(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int CG, bool kTmaEpi>
__global__ void __launch_bounds__(kThreads, 1)
umma_gemm_kernel(const __grid_constant__ OpMaps om, const KParams p) {
  gemm_body<CG, kTmaEpi, false>(&om, p, nullptr, nullptr);
}

__global__ void __launch_bounds__(kThreads, 1)
umma_syrk_grouped_kernel(const __grid_constant__ GroupMaps maps,
                         const __grid_constant__ GroupParams gp) {
  KParams p{};  // batch / bias / bf16-output fields unused by the grouped accumulate path
  gemm_body<2, true, true>(nullptr, p, &maps, &gp);
}

// Same, operands = the row-major bf16 activation matrices themselves (no staging pass): see gemm_body.
__global__ void __launch_bounds__(kThreads, 1)
umma_syrk_grouped_mn_kernel(const __grid_constant__ GroupMaps maps,
                            const __grid_constant__ GroupParams gp) {
  KParams p{};
  gemm_body<2, true, true, true>(nullptr, p, &maps, &gp);
}

// ---------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
  });
  return fn;
}

// K-major operand [batch][rows][K] bf16, row pitch ld, 128B-swizzled boxes of (64 x box_rows).
int make_operand_map(CUtensorMap* map, const __nv_bfloat16* base, int rows, int K, long long ld,
                     long long stride, int batch, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return -3;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld % 8) != 0 || (stride % 8) != 0) return -2;
  const bool shared = (stride == 0) || (batch == 1);
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows),
                        static_cast<cuuint64_t>(shared ? 1 : batch)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2,
                           static_cast<cuuint64_t>(shared ? static_cast<long long>(rows) * ld
                                                          : stride) *
                               2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                   const_cast<void*>(static_cast<const void*>(base)), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -4;
}

// Row-major activations X [n][d] bf16 (row pitch ld): 128B-swizzled boxes of 64 features x 64 samples.
int make_operand_map_mn(CUtensorMap* map, const __nv_bfloat16* base, int d, int n, long long ld) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return -3;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld % 8) != 0 || ld < d) return -2;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(d), static_cast<cuuint64_t>(n), 1};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(ld) * 2 * n};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(BK), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                   const_cast<void*>(static_cast<const void*>(base)), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -4;
}

// fp32 output [batch][rows][cols], row pitch ld elements: 128B-swizzled boxes of 32 x 32.
int make_output_map(CUtensorMap* map, float* base, int rows, int cols, long long ld,
                    long long stride, int batch) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return -3;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows),
                        static_cast<cuuint64_t>(batch)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 4,
                           static_cast<cuuint64_t>(batch == 1 ? static_cast<long long>(rows) * ld
                                                              : stride) *
                               4};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, static_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -4;
}

// bring-up / A-B timing switch for the TMA-reduce epilogue (1 = allowed)
int g_allow_tma_epilogue = 1;
int g_syrk_tuning_flags = 0;  // see set_syrk_tuning()

int sm_count() {
  static std::mutex mu;
  static int cached[kMaxDevices] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return kNumSMsB200;
  std::lock_guard<std::mutex> g(mu);
  if (cached[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = n > 0 ? n : kNumSMsB200;
  }
  return cached[dev];
}

// kTmaEpi: epilogue = TMA reduce-add of alpha*acc into C (see tma_epilogue_ok()).
template <int CG, bool kTmaEpi>
int launch_cg(const GemmArgs& a, cudaStream_t stream) {
  static DeviceOnce attr_once;
  if (!attr_once([] {
        return cudaFuncSetAttribute(umma_gemm_kernel<CG, kTmaEpi>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(Cfg<CG>::kSmemBytes)) == cudaSuccess;
      }))
    return -5;

  constexpr int kTileM = Cfg<CG>::kTileM;
  KParams p{};
  p.M = a.M;
  p.N = a.N;
  p.K = a.K;
  p.batch = a.batch;
  p.nparts = a.nparts;
  p.flags = a.flags;
  if ((a.flags & kSyrkLower) && a.A_hi == a.B_hi && a.A_lo == a.B_lo && a.A_lo2 == a.B_lo2 &&
      a.lda == a.ldb && a.strideA == a.strideB)
    p.flags |= kSameAB;
  p.tiles_m = (a.M + kTileM - 1) / kTileM;
  p.tiles_n = (a.N + BN - 1) / BN;
  if (a.flags & kSyrkLower) {
    int tot = 0;
    for (int mi = 0; mi < p.tiles_m; ++mi) tot += syrk_row_tiles(mi, p.tiles_n, kTileM);
    p.tiles_per_batch = tot;
  } else {
    p.tiles_per_batch = p.tiles_m * p.tiles_n;
  }
  p.num_tiles = p.tiles_per_batch * a.batch;
  p.a_bmul = (a.strideA == 0 || a.batch == 1) ? 0 : 1;
  p.b_bmul = (a.strideB == 0 || a.batch == 1) ? 0 : 1;
  p.c_bmul = (a.batch == 1) ? 0 : 1;
  p.tri_koff = a.tri_koff;
  p.alpha = a.alpha;
  p.beta = a.beta;
  p.C = a.C;
  p.ldc = a.ldc;
  p.strideC = a.strideC;
  p.bias = a.bias;
  p.strideBias = a.strideBias;
  p.Ohi = a.O_hi;
  p.Olo = a.O_lo;
  p.Olo2 = a.O_lo2;
  p.ldo = a.ldo;
  p.strideO = a.strideO;

  // Each CTA loads boxes of its own 128 A rows and of its share (256 / CG) of the B rows.
  OpMaps om;
  const __nv_bfloat16* ap[3] = {a.A_hi, a.A_lo, a.A_lo2};
  const __nv_bfloat16* bp[3] = {a.B_hi, a.B_lo, a.B_lo2};
  const int nsplit = a.nparts == 6 ? 3 : (a.nparts == 3 ? 2 : 1);
  int rc = 0;
  for (int i = 0; i < 3; ++i) {
    const int src = i < nsplit ? i : 0;
    rc = make_operand_map(&om.a[i], ap[src], a.M, a.K, a.lda, a.strideA, a.batch, BM);
    if (rc) return rc;
    rc = make_operand_map(&om.b[i], bp[src], a.N, a.K, a.ldb, a.strideB, a.batch, Cfg<CG>::kRowsB);
    if (rc) return rc;
  }
  om.c = om.a[0];
  // beta == 0 with the accumulating epilogue: zero-fill C first.
  if (kTmaEpi) {
    rc = make_output_map(&om.c, a.C, a.M, a.N, a.ldc, a.strideC, a.batch);
    if (rc) return rc;
    if (a.beta == 0.f) {
      for (int b = 0; b < a.batch; ++b)
        if (cudaMemset2DAsync(a.C + static_cast<long long>(b) * a.strideC,
                              static_cast<size_t>(a.ldc) * 4, 0, static_cast<size_t>(a.N) * 4,
                              static_cast<size_t>(a.M), stream) != cudaSuccess)
          return -5;
    }
  }
  int sms = sm_count();
  if (a.max_sms > 0 && a.max_sms < sms) sms = a.max_sms < CG ? CG : a.max_sms;
  const int clusters_max = sms / CG;
  const int clusters = p.num_tiles < clusters_max ? p.num_tiles : clusters_max;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(clusters * CG));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg<CG>::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CG == 2 ? 1 : 0;
  const cudaError_t err = cudaLaunchKernelEx(&cfg, umma_gemm_kernel<CG, kTmaEpi>, om, p);
  note_launch();
  return (err == cudaSuccess && cudaGetLastError() == cudaSuccess) ? 0 : -5;
}

}  // namespace

int launch_umma_syrk_grouped(const SyrkGroupItem* items, int count, int nparts, bool mirror,
                             cudaStream_t stream, bool mn_major) {
  if (count <= 0) return 0;
  if (count > kMaxGroup || (nparts != 1 && nparts != 3)) return -2;
  if (mn_major && nparts != 1) return -2;  // bf16 activations are exact in one pass: there is no lo part
  static DeviceOnce attr_once;
  if (!attr_once([] {
        return cudaFuncSetAttribute(umma_syrk_grouped_kernel,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(Cfg<2>::kSmemBytes)) == cudaSuccess &&
               cudaFuncSetAttribute(umma_syrk_grouped_mn_kernel,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(Cfg<2>::kSmemBytes)) == cudaSuccess;
      }))
    return -5;
  GroupMaps maps;
  GroupParams gp{};
  gp.count = count;
  gp.nparts = nparts;
  gp.mirror = mirror ? 1 : 0;
  int total = 0;
  for (int g = 0; g < count; ++g) {
    const SyrkGroupItem& it = items[g];
    if (it.X_hi == nullptr || it.C == nullptr || it.d <= 0 || it.n <= 0) return -2;
    if (nparts == 3 && it.X_lo == nullptr) return -2;
    if ((it.ldc % 4) != 0 || (reinterpret_cast<uintptr_t>(it.C) & 15) != 0) return -2;
    if (it.beta != 0.f && it.beta != 1.f) return -2;
    int rc = mn_major ? make_operand_map_mn(&maps.a0[g], it.X_hi, it.d, it.n, it.ldx)
                      : make_operand_map(&maps.a0[g], it.X_hi, it.d, it.n, it.ldx, 0, 1, BM);
    if (rc) return rc;
    if (nparts == 3) {
      rc = make_operand_map(&maps.a1[g], it.X_lo, it.d, it.n, it.ldx, 0, 1, BM);
      if (rc) return rc;
    } else {
      maps.a1[g] = maps.a0[g];
    }
    rc = make_output_map(&maps.c[g], it.C, it.d, it.d, it.ldc, 0, 1);
    if (rc) return rc;
    if (it.beta == 0.f &&
        cudaMemset2DAsync(it.C, static_cast<size_t>(it.ldc) * 4, 0, static_cast<size_t>(it.d) * 4,
                          static_cast<size_t>(it.d), stream) != cudaSuccess)
      return -5;
    const int tm = (it.d + 255) / 256;
    gp.tile_begin[g] = total;
    gp.iters[g] = (it.n + BK - 1) / BK * nparts;
    total += tm * (tm + 1) / 2;
    gp.M[g] = it.d;
    gp.K[g] = it.n;
    gp.alpha[g] = it.alpha;
  }
  for (int g = count; g <= kMaxGroup; ++g) gp.tile_begin[g] = total;
  for (int g = count; g < kMaxGroup; ++g) gp.iters[g] = 1;
  for (int g = count; g < kMaxGroup; ++g) {
    maps.a0[g] = maps.a0[0];
    maps.a1[g] = maps.a1[0];
    maps.c[g] = maps.c[0];
  }
  // every CTA pair of the device takes part: whole tiles round-robin, then an equal share of the tail
  const int clusters = sm_count() / 2;
  gp.full_tiles = (g_syrk_tuning_flags & 2) ? total / clusters * clusters : total;
  gp.no_dedup = (g_syrk_tuning_flags & 1) ? 1 : 0;
  long long tail = 0;
  for (int g = 0, t = 0; g < count; ++g) {
    const int tiles_g = gp.tile_begin[g + 1] - gp.tile_begin[g];
    for (int k = 0; k < tiles_g; ++k, ++t)
      if (t >= gp.full_tiles) tail += gp.iters[g];
  }
  if (tail > 0x7fffffffLL) return -2;
  gp.tail_work = static_cast<int>(tail);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(clusters * 2));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = mirror ? Cfg<2>::kSmemBytes : Cfg<2>::kSmemBytesLower;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t err = mn_major ? cudaLaunchKernelEx(&cfg, umma_syrk_grouped_mn_kernel, maps, gp)
                                   : cudaLaunchKernelEx(&cfg, umma_syrk_grouped_kernel, maps, gp);
  note_launch();
  return (err == cudaSuccess && cudaGetLastError() == cudaSuccess) ? 0 : -5;
}

// A/B switches of the grouped SYRK (bring-up / profiling): bit 0 = no diagonal-tile operand dedup,
// bit 1 = stream-K split of the last partial wave (default: whole tiles only).
void set_syrk_tuning(int flags) { g_syrk_tuning_flags = flags; }

// cta_group override for bring-up / A-B timing: 0 = automatic, 1 or 2 = forced; +16 disables the
// TMA-reduce epilogue (register read-modify-write epilogue everywhere).
static int g_force_cta_group = 0;
void set_umma_cta_group(int cg) {
  g_allow_tma_epilogue = (cg & 16) ? 0 : 1;
  cg &= 15;
  g_force_cta_group = (cg == 1 || cg == 2) ? cg : 0;
}

int launch_umma_gemm(const GemmArgs& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0 || a.batch <= 0) return 0;
  if (a.K <= 0) return -2;
  if (a.A_hi == nullptr || a.B_hi == nullptr) return -2;
  if (a.nparts != 1 && a.nparts != 3 && a.nparts != 6) return -2;
  if (a.nparts >= 3 && (a.A_lo == nullptr || a.B_lo == nullptr)) return -2;
  if (a.nparts == 6 && (a.A_lo2 == nullptr || a.B_lo2 == nullptr)) return -2;
  if ((a.flags & kSyrkLower) && a.M != a.N) return -2;
  // CTA pairs (256-row tiles) pay off once the 256-row quantisation wastes little: at least one
  // full pair tile of rows and a full 256-column tile.
  int cg = (a.M >= 192 && a.N >= 129) ? 2 : 1;
  if (g_force_cta_group) cg = g_force_cta_group;
  // Accumulating epilogue through the TMA (see the kernel): needs a 16 B aligned, 16 B pitched C,
  // no fused bias / relu / bf16 outputs, and beta in {0, 1}.
  const bool tma_epi =
      g_allow_tma_epilogue && a.C != nullptr && a.bias == nullptr && a.O_hi == nullptr &&
      !(a.flags & kRelu) && (a.beta == 1.f || a.beta == 0.f) && (a.ldc % 4) == 0 &&
      (reinterpret_cast<uintptr_t>(a.C) & 15) == 0 && (a.batch == 1 || (a.strideC % 4) == 0);
  if (cg == 2)
    return tma_epi ? launch_cg<2, true>(a, stream) : launch_cg<2, false>(a, stream);
  return tma_epi ? launch_cg<1, true>(a, stream) : launch_cg<1, false>(a, stream);
}

}  // namespace bk
