// Fused attention pooling, forward (abmil.py:36-45; clam.py:37-60,170):
//
//   uv = act(h wab^T + bab)        tanh, or tanh | sigmoid for the gated form      [n_rows, NC]   NC = D or 2D
//   s  = wc . g(uv) + bc           g = u or u*v                                     [n_rows]
//   p  = post_scale_b * softmax over the rows of bag b (s)                         [n_rows]
//   M  = sum_n p[n] h[n, :]                                                          [B, L]
//
// with h read from HBM once.  A persistent CTA takes 128-row tiles of h; 18 warps:
//   warp 16  TMA producer: h k-blocks (and, when NC > 128, weight k-blocks) through a 128B-swizzled ring, mbarrier
//            tx-count.  With NC == 128 the whole projection weight (128 x L bf16 <= 128 KB) is fetched once and stays
//            resident in shared memory: the kernel is bound by the L2 slice bandwidth (~42 B/clk/SM), and re-streaming
//            the weights for every tile would double its L2 traffic.  h loads carry an L2 evict-last hint (see pooling).
//   warp 17  MMA issuer: NC/128 column passes of tcgen05.mma 128x128x16 (bf16 in, fp32 accumulators in TMEM columns
//            [128*pass, +128)); two accumulator stages when 2*NC <= 512
//   warps 0-15 two epilogue TEAMS of 8 warps; team k owns accumulator stage k and takes every second tile, so the score
//            phase of one tile overlaps the pooling phase of the previous one.  Per tile: thread = row (TMEM lane), the two
//            warps of a lane quarter split the D columns: bias, tanh / sigmoid (MUFU), rounding to the bf16 value that is
//            saved for the backward pass (staged in 32x32 units, bulk tensor stores), gating, dot with wc -> raw score.
//            Then, per bag segment inside the tile: tile-local softmax statistics (max, sum exp) by warp shuffles and the
//            exp-weighted column sums of the tile's rows, re-read with coalesced 16-byte loads while they are still in L2
//            (evict-first on this last use) -> one record (m, l, acc[L]) per (tile, bag) incidence, at index tile + bag
//            (strictly increasing along the rows, so unique).
// attnpool_merge_kernel (one CTA per bag) folds the records of a bag (online-softmax merge), writes M, the statistics
// and the normalised weights p.  The separate score / softmax / weighted-sum kernels this replaces read uv once more
// and h a second time from HBM.
//
// Roofline: HBM.  Algorithmic bytes per row: L*2 (h) + NC*2 (uv, when saved) + 8 (s, p).  FLOPs per row: 2*L*NC + 2*L.
// (Tried and dropped: an L2 prefetch of the NEXT tile's rows by the producer, which bought the GEMMs 3-7 %, costs this
// kernel 20 % - 104 -> 125 us - because it evicts the current tile's rows before the pooling pass re-reads them.)
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

// Timeline tracing / phase skipping (MURCL_DEBUG_EPI, MURCL_DEBUG_ATTNPOOL*) is compiled in only with -DMURCL_TRACE: the
// production kernels carry no instrumentation.
#ifdef MURCL_TRACE
#define MURCL_TRACE_ON 1
#else
#define MURCL_TRACE_ON 0
#endif


namespace murcl {
namespace ap {

using namespace tc;

constexpr int TILE_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int PASS_N = 128;
constexpr int MAX_L = 512;
constexpr int MAX_KB = MAX_L / BLOCK_K;
constexpr int SLAB_BYTES = TILE_M * BLOCK_K * 2;         // 16 KB: one k-block of the h tile, or of 128 weight rows
constexpr int TEAM_WARPS = 8;
constexpr int TEAM_THREADS = 32 * TEAM_WARPS;
constexpr int MAX_TEAMS = 2;
constexpr int NUM_EPI_WARPS = MAX_TEAMS * TEAM_WARPS;
constexpr int PRODUCER_WARP = NUM_EPI_WARPS, MMA_WARP = NUM_EPI_WARPS + 1;
constexpr int NUM_THREADS = 32 * (NUM_EPI_WARPS + 2);
constexpr int MAX_NC = 512, MAX_D = 512;
constexpr int UNIT_BYTES = 32 * 32 * 2;                   // uv staging unit: 32 rows x 32 bf16 columns (64-byte rows, SWIZZLE_64B)
constexpr int STAGING_BYTES = NUM_EPI_WARPS * UNIT_BYTES;
constexpr int TEAM_FLOATS = 4 * TILE_M;                  // s_tile, sc_part[2], e_tile
constexpr int MISC_FLOATS = MAX_TEAMS * TEAM_FLOATS + MAX_NC + MAX_D;
constexpr int REC_HEAD = 4;                              // record = [m, l, -, -, acc[L]]

template <bool BSTAT>
struct Cfg {
  static constexpr int BRES_BYTES = BSTAT ? MAX_KB * SLAB_BYTES : 0;
  static constexpr int STAGE_BYTES = BSTAT ? SLAB_BYTES : 2 * SLAB_BYTES;
  static constexpr int STAGES = BSTAT ? 3 : 5;
  static constexpr int SMEM_BYTES = BRES_BYTES + STAGES * STAGE_BYTES + STAGING_BYTES + MISC_FLOATS * 4 + 256 /*barriers*/ + 1024 /*align*/;
};

struct Params {
  int64_t n_rows;
  int L, NC, D, gated, n_tiles, tmem_cols, n_acc;
  const __nv_bfloat16* h;
  const float* bab;
  const float* wc;
  const float* bc;
  const int64_t* offsets;
  const int32_t* row_seg;
  __nv_bfloat16* uv;      // may be null (inference: nothing is kept for a backward pass)
  float* s;
  float* rec;
  int reverse;                 // walk the tiles from the last rows to the first (murcl_set_row_order)
  int skip;                    // MURCL_DEBUG_ATTNPOOL_SKIP bit mask (timing experiments; results are wrong when set)
  unsigned long long* trace;   // MURCL_DEBUG_ATTNPOOL=1: %globaltimer stamps of CTA 0 / warp 0, 8 per tile
};

#define AP_STAMP(slot)                                                              \
  if (MURCL_TRACE_ON && p.trace && blockIdx.x == 0 && threadIdx.x == 0 && it < 128) {                 \
    unsigned long long ts_;                                                         \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts_));                         \
    p.trace[(it >> 1) * 8 + (slot)] = ts_;                                          \
  }

__device__ __forceinline__ void team_sync(int team) { asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(TEAM_THREADS) : "memory"); }

__device__ __forceinline__ void fma8(float (&acc)[16], int o, const uint4& q, float e) {
  const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(hh[j]);
    acc[o + 2 * j] = fmaf(e, f.x, acc[o + 2 * j]);
    acc[o + 2 * j + 1] = fmaf(e, f.y, acc[o + 2 * j + 1]);
  }
}

template <bool BSTAT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
attnpool_fwd_kernel(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_w,
                    const __grid_constant__ CUtensorMap map_uv, const Params p) {
  using C = Cfg<BSTAT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t bres = (raw + 1023u) & ~1023u;                       // SWIZZLE_128B atoms need 1024 B alignment
  const uint32_t ring = bres + C::BRES_BYTES;
  const uint32_t staging = ring + C::STAGES * C::STAGE_BYTES;         // 1024 B aligned
  const uint32_t misc = staging + STAGING_BYTES;
  float* misc_f = reinterpret_cast<float*>(smem_raw + (misc - raw));
  float* bab_s = misc_f + MAX_TEAMS * TEAM_FLOATS;                    // [NC]
  float* wc_s = bab_s + MAX_NC;                                       // [D]
  const uint32_t bars = misc + MISC_FLOATS * 4;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * C::STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * C::STAGES + 2 + a); };
  const uint32_t bres_bar = bars + 8u * (2 * C::STAGES + 4);
  const uint32_t tmem_slot = bars + 8u * (2 * C::STAGES + 5);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = p.L / BLOCK_K, n_pass = p.NC / PASS_N;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), TEAM_WARPS);
    }
    mbar_init(bres_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_h)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
  }
  for (int i = threadIdx.x; i < p.NC; i += NUM_THREADS) bab_s[i] = p.bab ? p.bab[i] : 0.f;
  for (int i = threadIdx.x; i < p.D; i += NUM_THREADS) wc_s[i] = p.wc[i];
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == PRODUCER_WARP) {
    // ================= TMA producer =================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      // the tile's rows are read again by the pooling pass a few microseconds later: keep them in L2 until then
      const uint64_t keep = l2_policy_evict_last();
      if (BSTAT && (int)blockIdx.x < p.n_tiles) {
        mbar_expect_tx(bres_bar, (uint32_t)(nkb * SLAB_BYTES));
        for (int kb = 0; kb < nkb; ++kb) tma_load_2d(bres + kb * SLAB_BYTES, &map_w, bres_bar, kb * BLOCK_K, 0);
      }
      for (int ti = blockIdx.x; ti < p.n_tiles; ti += gridDim.x) {
        const int t = p.reverse ? p.n_tiles - 1 - ti : ti;
        const int row0 = t * TILE_M;
        for (int j = 0; j < n_pass; ++j) {
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(empty_bar(s), ph ^ 1u);
            const uint32_t a_dst = ring + s * C::STAGE_BYTES;
            mbar_expect_tx(full_bar(s), C::STAGE_BYTES);
            if (j == 0) tma_load_2d_hint(a_dst, &map_h, full_bar(s), kb * BLOCK_K, row0, keep);   // rows >= n_rows arrive as zeros
            else tma_load_2d(a_dst, &map_h, full_bar(s), kb * BLOCK_K, row0);
            if (!BSTAT) tma_load_2d(a_dst + SLAB_BYTES, &map_w, full_bar(s), kb * BLOCK_K, j * PASS_N);
            if (++s == C::STAGES) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TILE_M, PASS_N, false, false);
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      if (BSTAT && (int)blockIdx.x < p.n_tiles) mbar_wait(bres_bar, 0u);   // the resident projection weights have landed
      for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        mbar_wait(tempty_bar(as), aph ^ 1u);                             // the owning team has read this accumulator stage
        tcgen05_fence_after();
        for (int j = 0; j < n_pass; ++j) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.NC + j * PASS_N);
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(full_bar(s), ph);
            tcgen05_fence_after();
            const uint32_t a_src = ring + s * C::STAGE_BYTES;
            const uint32_t b_src = BSTAT ? bres + kb * SLAB_BYTES : a_src + SLAB_BYTES;
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              tcgen05_mma_bf16(d_tmem, make_smem_desc(a_src + k * 32, 16, 1024), make_smem_desc(b_src + k * 32, 16, 1024),
                               idesc, (kb > 0 || k > 0) ? 1u : 0u);
            tcgen05_commit(empty_bar(s));                                // ring slot reusable once these MMAs retire
            if (++s == C::STAGES) { s = 0; ph ^= 1u; }
          }
        }
        tcgen05_commit(tfull_bar(as));                                   // all NC accumulator columns complete
        if (++as == p.n_acc) { as = 0; aph ^= 1u; }
      }
    }
  } else if ((warp >> 3) < p.n_acc) {
    // ================= epilogue + pooling teams (warps 0-7, 8-15) =================
    const int team = warp >> 3, wt = warp & 7;
    const int quarter = warp & 3, chalf = wt >> 2;
    const int r = quarter * 32 + lane;                                   // row of the tile = TMEM lane
    const int tid = threadIdx.x - team * TEAM_THREADS;                   // 0..255 within the team
    float* s_tile = misc_f + team * TEAM_FLOATS;                         // [128] raw scores of the tile
    float* sc_part = s_tile + TILE_M;                                    // [2][128] partial dot products of the column halves
    float* e_tile = sc_part + 2 * TILE_M;                                // [128] exp(s - m) of the current segment
    // per-warp partial column sums of the current segment live in the team's (then idle) uv staging units: 8 x 2 KB
    float* pacc = reinterpret_cast<float*>(smem_raw + (staging + (uint32_t)(team * TEAM_WARPS) * UNIT_BYTES - raw));
    const int dh = p.D >> 1;                                             // D columns per warp of a quarter (multiple of 32)
    const int D = p.D, NC = p.NC, L = p.L;
    const bool gated = p.gated != 0;
    const float bc = p.bc ? p.bc[0] : 0.f;
    const int64_t rec_stride = L + REC_HEAD;
    const uint64_t once = l2_policy_evict_first();                       // last use of the tile's rows
    const uint32_t unit = staging + (uint32_t)warp * UNIT_BYTES;
    const uint32_t my_row_off = (uint32_t)lane * 64u, swz = (uint32_t)((lane >> 1) & 3);   // SWIZZLE_64B: chunk ^= (row >> 1) & 3
    const bool act0 = 8 * lane < L, act1 = 256 + 8 * lane < L;           // pooling: lane owns columns [8*lane, +8) and [256 + 8*lane, +8)
    uint32_t aph = 0;
    for (int it = team; blockIdx.x + (int64_t)it * gridDim.x < p.n_tiles; it += p.n_acc) {
      const int ti = blockIdx.x + it * gridDim.x;
      const int t = p.reverse ? p.n_tiles - 1 - ti : ti;
      const int64_t row0 = (int64_t)t * TILE_M;
      const int64_t row = row0 + r;
      const bool valid = row < p.n_rows;
      AP_STAMP(0)
      // segment bookkeeping of the tile: dependent global loads, issued before the wait so their latency is hidden
      const int64_t last_row = (row0 + TILE_M - 1 < p.n_rows ? row0 + TILE_M - 1 : p.n_rows - 1);
      const int b_lo = __ldg(p.row_seg + row0), b_hi = __ldg(p.row_seg + last_row);
      int64_t o_next = __ldg(p.offsets + b_lo);
      mbar_wait(tfull_bar(team), aph);
      aph ^= 1u;
      tcgen05_fence_after();
      AP_STAMP(1)
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(team * NC);
      float part = 0.f;
      for (int d0 = chalf * dh; d0 < (chalf + 1) * dh; d0 += 32) {
        uint32_t vp[16];                                                 // gated: the sigmoid branch waits here for its store
#pragma unroll
        for (int k = 0; k < 2; ++k) {                                    // 16 columns at a time (register budget: 576 threads)
          const int c0 = d0 + 16 * k;
          uint32_t ra[16], rb[16], up[8];
          tmem_ld16(t_row + (uint32_t)c0, ra);
          if (gated) tmem_ld16(t_row + (uint32_t)(D + c0), rb);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            // the saved (bf16-rounded) activations are what the score uses: forward and backward see the same numbers
            const float4 ba = *reinterpret_cast<const float4*>(bab_s + c0 + i);
            const float4 w4 = *reinterpret_cast<const float4*>(wc_s + c0 + i);
            const __nv_bfloat162 u01 = __floats2bfloat162_rn(tanh_fast(__uint_as_float(ra[i]) + ba.x),
                                                             tanh_fast(__uint_as_float(ra[i + 1]) + ba.y));
            const __nv_bfloat162 u23 = __floats2bfloat162_rn(tanh_fast(__uint_as_float(ra[i + 2]) + ba.z),
                                                             tanh_fast(__uint_as_float(ra[i + 3]) + ba.w));
            up[i >> 1] = *reinterpret_cast<const uint32_t*>(&u01);
            up[(i >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&u23);
            float2 g01 = __bfloat1622float2(u01), g23 = __bfloat1622float2(u23);
            if (gated) {
              const float4 bb = *reinterpret_cast<const float4*>(bab_s + D + c0 + i);
              const __nv_bfloat162 v01 = __floats2bfloat162_rn(fmaf(0.5f, tanh_fast(0.5f * (__uint_as_float(rb[i]) + bb.x)), 0.5f),
                                                               fmaf(0.5f, tanh_fast(0.5f * (__uint_as_float(rb[i + 1]) + bb.y)), 0.5f));
              const __nv_bfloat162 v23 = __floats2bfloat162_rn(fmaf(0.5f, tanh_fast(0.5f * (__uint_as_float(rb[i + 2]) + bb.z)), 0.5f),
                                                               fmaf(0.5f, tanh_fast(0.5f * (__uint_as_float(rb[i + 3]) + bb.w)), 0.5f));
              vp[8 * k + (i >> 1)] = *reinterpret_cast<const uint32_t*>(&v01);
              vp[8 * k + (i >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&v23);
              const float2 h01 = __bfloat1622float2(v01), h23 = __bfloat1622float2(v23);
              g01.x *= h01.x; g01.y *= h01.y; g23.x *= h23.x; g23.y *= h23.y;
            }
            part = fmaf(w4.x, g01.x, part);
            part = fmaf(w4.y, g01.y, part);
            part = fmaf(w4.z, g23.x, part);
            part = fmaf(w4.w, g23.y, part);
          }
          if (p.uv != nullptr && !(MURCL_TRACE_ON && p.skip & 2)) {
            // stage the 32 x 32 unit swizzled like the TMA box and hand it to the bulk-store engine (a direct store
            // would touch 32 different 128-byte lines per instruction); rows >= n_rows are clipped by the tensor map
            if (k == 0) {
              if (lane == 0) bulk_wait_read0();
              __syncwarp();
            }
            st_shared_v4(unit + my_row_off + (((uint32_t)(2 * k) ^ swz) << 4), up[0], up[1], up[2], up[3]);
            st_shared_v4(unit + my_row_off + (((uint32_t)(2 * k + 1) ^ swz) << 4), up[4], up[5], up[6], up[7]);
          }
        }
        if (p.uv != nullptr && !(MURCL_TRACE_ON && p.skip & 2)) {
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&map_uv, unit, d0, (int)row0 + quarter * 32);
            bulk_commit();
          }
          if (gated) {
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 4; ++c)
              st_shared_v4(unit + my_row_off + (((uint32_t)c ^ swz) << 4), vp[4 * c], vp[4 * c + 1], vp[4 * c + 2], vp[4 * c + 3]);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&map_uv, unit, D + d0, (int)row0 + quarter * 32);
              bulk_commit();
            }
          }
        }
      }
      AP_STAMP(2)
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(team));                      // this warp has drained its part of the accumulator
      sc_part[chalf * TILE_M + r] = part;
      team_sync(team);
      if (tid < TILE_M) {                                                // tid == r for the warps of column half 0
        const float sv = sc_part[tid] + sc_part[TILE_M + tid] + bc;
        s_tile[tid] = sv;
        if (valid) p.s[row] = sv;
      }
      team_sync(team);
      AP_STAMP(3)
      // ---- per bag segment of the tile: softmax statistics + exp-weighted column sums of the tile's rows (L2 hits) ----
      for (int b = b_lo; b <= b_hi && !(MURCL_TRACE_ON && p.skip & 4); ++b) {
        const int64_t o0 = o_next, o1 = __ldg(p.offsets + b + 1);
        o_next = o1;
        const int r_begin = (int)(o0 > row0 ? o0 - row0 : 0);
        const int r_end = (int)(o1 - row0 < TILE_M ? o1 - row0 : TILE_M);
        if (r_end <= r_begin) continue;                                  // empty bag (uniform branch)
        float m = -INFINITY;
        for (int rr = r_begin + lane; rr < r_end; rr += 32) m = fmaxf(m, s_tile[rr]);
        m = warp_max(m);                                                 // every warp computes the same value
        if (tid < TILE_M) e_tile[tid] = (tid >= r_begin && tid < r_end) ? expf(s_tile[tid] - m) : 0.f;
        if (lane == 0) bulk_wait_read0();                                // the bulk stores no longer read this warp's staging unit
        team_sync(team);
        AP_STAMP(4)
        // warp w takes rows r_begin + w, + 8, ...; a warp reads whole rows with two 512-byte requests
        {
          float acc[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[i] = 0.f;
          if (act0 && !(MURCL_TRACE_ON && p.skip & 1)) {
            const __nv_bfloat16* base = p.h + row0 * L + 8 * lane;
            int rr = r_begin + wt;
            for (; rr + 24 < r_end; rr += 32) {                            // four rows in flight
              uint4 q[8];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const __nv_bfloat16* src = base + (int64_t)(rr + 8 * u) * L;
                q[2 * u] = ldg_v4_hint(src, once);
                if (act1) q[2 * u + 1] = ldg_v4_hint(src + 256, once);
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float e = e_tile[rr + 8 * u];
                fma8(acc, 0, q[2 * u], e);
                if (act1) fma8(acc, 8, q[2 * u + 1], e);
              }
            }
            for (; rr < r_end; rr += 8) {
              const __nv_bfloat16* src = base + (int64_t)rr * L;
              const float e = e_tile[rr];
              fma8(acc, 0, ldg_v4_hint(src, once), e);
              if (act1) fma8(acc, 8, ldg_v4_hint(src + 256, once), e);
            }
          }
          // this warp's partial sums (rows wt, wt + 8, ...) into its own unit; all 512 floats are written
          float4* dst = reinterpret_cast<float4*>(pacc + wt * MAX_L + 8 * lane);
          dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
          dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
          dst[64] = make_float4(acc[8], acc[9], acc[10], acc[11]);
          dst[65] = make_float4(acc[12], acc[13], acc[14], acc[15]);
        }
        AP_STAMP(5)
        team_sync(team);
        AP_STAMP(6)
        float* rec = p.rec + (int64_t)(t + b) * rec_stride;
        if (2 * tid < L) {                                               // fixed summation order: deterministic
          float a0 = 0.f, a1 = 0.f;
#pragma unroll
          for (int w = 0; w < TEAM_WARPS; ++w) {
            const float2 v = *reinterpret_cast<const float2*>(pacc + w * MAX_L + 2 * tid);
            a0 += v.x;
            a1 += v.y;
          }
          *reinterpret_cast<float2*>(rec + REC_HEAD + 2 * tid) = make_float2(a0, a1);
        }
        if (wt == 0) {
          float l = 0.f;
          for (int rr = r_begin + lane; rr < r_end; rr += 32) l += e_tile[rr];
          l = warp_sum(l);
          if (lane == 0) {
            rec[0] = m;
            rec[1] = l;
          }
        }
        team_sync(team);                                                 // e_tile / pacc are rewritten for the next segment
        AP_STAMP(7)
      }
    }
    if (lane == 0) bulk_wait_all();                                      // this warp's bulk stores have completed
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// One CTA per bag: merge the (tile, bag) records, write M, (m, l) and the normalised weights p.
__global__ void __launch_bounds__(256) attnpool_merge_kernel(const float* __restrict__ rec, const float* __restrict__ s,
                                                             const int64_t* __restrict__ offsets, int L, int inv_sqrt_n,
                                                             float* __restrict__ M, float* __restrict__ p,
                                                             float* __restrict__ stats) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const int64_t o0 = offsets[b], o1 = offsets[b + 1];
  const int64_t stride = L + REC_HEAD;
  if (o1 <= o0) {
    for (int c = threadIdx.x; c < L; c += blockDim.x) M[(int64_t)b * L + c] = 0.f;
    if (threadIdx.x == 0 && stats) {
      stats[2 * b] = -INFINITY;
      stats[2 * b + 1] = 0.f;
    }
    return;
  }
  const int64_t t0 = o0 / TILE_M, t1 = (o1 - 1) / TILE_M;
  float m = -INFINITY;
  for (int64_t t = t0 + threadIdx.x; t <= t1; t += blockDim.x) m = fmaxf(m, rec[(t + b) * stride]);
  m = block_max(m, red);
  float l = 0.f;
  for (int64_t t = t0 + threadIdx.x; t <= t1; t += blockDim.x) l += rec[(t + b) * stride + 1] * expf(rec[(t + b) * stride] - m);
  l = block_sum(l, red);
  const float root = inv_sqrt_n ? sqrtf((float)(o1 - o0)) : 1.f;
  for (int c = threadIdx.x; c < L; c += blockDim.x) {
    float acc = 0.f;
    for (int64_t t = t0; t <= t1; ++t) {
      const float* rt = rec + (t + b) * stride;
      acc = fmaf(expf(rt[0] - m), rt[REC_HEAD + c], acc);
    }
    float v = __fdiv_rn(acc, l);
    if (inv_sqrt_n) v = __fdiv_rn(v, root);
    M[(int64_t)b * L + c] = v;
  }
  for (int64_t n = o0 + threadIdx.x; n < o1; n += blockDim.x) {
    float v = __fdiv_rn(expf(s[n] - m), l);                              // same expression as seg_softmax_kernel
    if (inv_sqrt_n) v = __fdiv_rn(v, root);
    p[n] = v;
  }
  if (threadIdx.x == 0 && stats) {
    stats[2 * b] = m;
    stats[2 * b + 1] = l;
  }
}

static int tmem_cols_for(int nc) {
  int c = 32;
  while (c < nc) c <<= 1;
  return c;
}

}  // namespace ap
}  // namespace murcl

using namespace murcl;

extern "C" {

int murcl_attnpool_supported(int L, int D, int gated, int dtype) {
  const int nc = D * (gated ? 2 : 1);
  return dtype == MURCL_BF16 && L >= 64 && L <= ap::MAX_L && L % ap::BLOCK_K == 0 && D % 64 == 0 &&
                 nc % ap::PASS_N == 0 && nc <= ap::MAX_NC
             ? 1
             : 0;
}

int64_t murcl_attnpool_workspace(int64_t n_rows, int B, int L) {
  const int64_t tiles = (n_rows + ap::TILE_M - 1) / ap::TILE_M;
  return (tiles + B + 1) * (int64_t)(L + ap::REC_HEAD);
}

int murcl_attnpool_fwd(const void* h, const void* wab, const float* bab, const float* wc, const float* bc,
                       const int64_t* offsets, const int32_t* row_seg, int64_t n_rows, int B, int L, int D, int gated,
                       int inv_sqrt_n, int dtype, void* uv, float* s, float* p, float* M, float* stats, float* workspace,
                       void* stream) {
  MURCL_REQUIRE(h && wab && wc && offsets && row_seg && s && p && M && workspace, "attnpool_fwd: null pointer");
  MURCL_REQUIRE(n_rows >= 0 && B >= 0 && n_rows < ((int64_t)1 << 31) - 256, "attnpool_fwd: bad shape");
  MURCL_REQUIRE(murcl_attnpool_supported(L, D, gated, dtype),
                "attnpool_fwd: unsupported configuration L=%d D=%d gated=%d dtype=%d (bf16, L <= 512 and %% 64 == 0, "
                "D %% 64 == 0, D*(1+gated) %% 128 == 0 and <= 512)", L, D, gated, dtype);
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  MURCL_REQUIRE(al16(h) && al16(wab) && (uv == nullptr || al16(uv)) && al16(workspace), "attnpool_fwd: operands must be 16-byte aligned");
  if (B == 0) return MURCL_OK;
  cudaStream_t st = as_stream(stream);
  const int nc = D * (gated ? 2 : 1);
  if (n_rows > 0) {
    CUtensorMap mh, mw;
    int rc = tc::make_map(&mh, h, n_rows, L, ap::BLOCK_K, ap::TILE_M);
    if (rc != MURCL_OK) return rc;
    rc = tc::make_map(&mw, wab, nc, L, ap::BLOCK_K, ap::PASS_N);
    if (rc != MURCL_OK) return rc;
    CUtensorMap muv = mw;                                                // unused when uv == NULL
    if (uv != nullptr) {
      rc = tc::make_map(&muv, uv, n_rows, nc, 32, 32, 2, true);
      if (rc != MURCL_OK) return rc;
    }
    const bool bstat = nc == ap::PASS_N;                                 // one column pass: the weights stay in shared memory
    static PerDeviceOnce configured;
    if (const int slot = configured.pending(); slot >= 0) {
      MURCL_CUDA(cudaFuncSetAttribute(ap::attnpool_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      ap::Cfg<true>::SMEM_BYTES));
      MURCL_CUDA(cudaFuncSetAttribute(ap::attnpool_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      ap::Cfg<false>::SMEM_BYTES));
      configured.mark(slot);
    }
    ap::Params prm{};
    prm.n_rows = n_rows; prm.L = L; prm.NC = nc; prm.D = D; prm.gated = gated;
    prm.n_tiles = (int)((n_rows + ap::TILE_M - 1) / ap::TILE_M);
    prm.n_acc = (2 * nc <= 512) ? 2 : 1;
    prm.tmem_cols = ap::tmem_cols_for(prm.n_acc * nc);
    prm.h = static_cast<const __nv_bfloat16*>(h);
    prm.bab = bab; prm.wc = wc; prm.bc = bc; prm.offsets = offsets; prm.row_seg = row_seg;
    prm.uv = static_cast<__nv_bfloat16*>(uv); prm.s = s; prm.rec = workspace;
    prm.reverse = row_order_descending();
    static int debug = -1;
    static unsigned long long* trace_buf = nullptr;
    if (debug < 0) {
      const char* e = getenv("MURCL_DEBUG_ATTNPOOL");
      debug = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    {
      static int skip = -1;
      if (skip < 0) {
        const char* e = getenv("MURCL_DEBUG_ATTNPOOL_SKIP");
        skip = e ? atoi(e) : 0;
      }
      prm.skip = skip;
    }
    if (debug) {
      if (!trace_buf) cudaMalloc(&trace_buf, sizeof(unsigned long long) * 8 * 64);
      cudaMemsetAsync(trace_buf, 0, sizeof(unsigned long long) * 8 * 64, st);
      prm.trace = trace_buf;
    }
    const int grid = prm.n_tiles < sm_count() ? prm.n_tiles : sm_count();
    if (bstat) ap::attnpool_fwd_kernel<true><<<grid, ap::NUM_THREADS, ap::Cfg<true>::SMEM_BYTES, st>>>(mh, mw, muv, prm);
    else ap::attnpool_fwd_kernel<false><<<grid, ap::NUM_THREADS, ap::Cfg<false>::SMEM_BYTES, st>>>(mh, mw, muv, prm);
    int rc2 = check_launch("attnpool_fwd_kernel");
    if (rc2 != MURCL_OK) return rc2;
    if (debug) {
      static int printed = 0;
      cudaStreamSynchronize(st);
      if (printed < 2 && prm.n_tiles > 4 * grid) {
        ++printed;
        unsigned long long hbuf[8 * 12];
        cudaMemcpy(hbuf, trace_buf, sizeof(hbuf), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[attnpool trace] CTA 0 warp 0 (team 0: every second tile), ns: wait-acc | score | scores->smem | stats | pool loads+fma | sync | record+sync   (NC=%d)\n", nc);
        for (int i = 3; i < 11; ++i) {
          const unsigned long long* g = hbuf + 8 * i;
          fprintf(stderr, "  tile %2d (+%6lld): %5lld | %5lld | %5lld | %5lld | %5lld | %5lld | %5lld\n", i, (long long)(g[0] - hbuf[0]),
                  (long long)(g[1] - g[0]), (long long)(g[2] - g[1]), (long long)(g[3] - g[2]), (long long)(g[4] - g[3]),
                  (long long)(g[5] - g[4]), (long long)(g[6] - g[5]), (long long)(g[7] - g[6]));
        }
      }
    }
  }
  ap::attnpool_merge_kernel<<<B, 256, 0, st>>>(workspace, s, offsets, L, inv_sqrt_n, M, p, stats);
  return check_launch("attnpool_merge_kernel");
}

}  // extern "C"
