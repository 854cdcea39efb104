// Fused attention pooling, forward (abmil.py:36-45; clam.py:37-60,170):
//
//   uv = act(h wab^T + bab)        tanh, or tanh | sigmoid for the gated form      [n_rows, NC]   NC = D or 2D
//   s  = wc . g(uv) + bc           g = u or u*v                                     [n_rows]
//   p  = post_scale_b * softmax over the rows of bag b (s)                         [n_rows]
//   M  = sum_n p[n] h[n, :]                                                          [B, L]
//
// in ONE pass over h.  A persistent CTA takes 128-row tiles of h:
//   warp 8   TMA producer: the tile's L/64 k-blocks land in a RESIDENT 128B-swizzled buffer (128 x L bf16, <= 128 KB);
//            the projection weights stream through a 4-slot ring (one 128 x 64 block per slot); the next tile is pulled
//            into L2 meanwhile
//   warp 9   MMA issuer: NC/128 column passes of tcgen05.mma 128x128x16 (bf16 in, fp32 accumulators in TMEM columns
//            [128*pass, +128)); the A operand is re-read from the resident tile for every pass
//   warps 0-7 epilogue: thread = row (TMEM lane), the two warps of a lane quarter split the D columns: bias, tanh /
//            sigmoid (MUFU), rounding to the bf16 value that is saved for the backward pass, gating, dot with wc -> raw
//            score; then, per bag segment inside the tile, tile-local softmax statistics (max, sum exp) by warp shuffles
//            and the exp-weighted column sums of the SAME shared-memory tile (warp w owns k-block w = columns 64w..64w+63)
//            -> one record (m, l, acc[L]) per (tile, bag) incidence, at index tile + bag (strictly increasing along
//            the rows, so unique).
// attnpool_merge_kernel (one CTA per bag) folds the records of a bag (online-softmax merge), writes M, the statistics
// and the normalised weights p.  h is read from HBM exactly once; the separate score / softmax / weighted-sum kernels
// it replaces read uv once and h a second time.
//
// Roofline: HBM.  Algorithmic bytes per row: L*2 (h) + NC*2 (uv, when saved) + 8 (s, p).  FLOPs per row: 2*L*NC + 2*L.
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace murcl {
namespace ap {

using namespace tc;

constexpr int TILE_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int PASS_N = 128;
constexpr int MAX_KB = 8;                               // L <= 512
constexpr int B_STAGES = 4;
constexpr int SLAB_BYTES = TILE_M * BLOCK_K * 2;         // 16 KB: one k-block of the h tile, or of 128 weight rows
constexpr int NUM_EPI_WARPS = 8;
constexpr int EPI_THREADS = 32 * NUM_EPI_WARPS;
constexpr int PRODUCER_WARP = NUM_EPI_WARPS, MMA_WARP = NUM_EPI_WARPS + 1;
constexpr int NUM_THREADS = 32 * (NUM_EPI_WARPS + 2);
constexpr int MAX_NC = 512, MAX_D = 512;
constexpr int MISC_FLOATS = 4 * TILE_M + MAX_NC + MAX_D; // s_tile, sc_part[2], e_tile, bab, wc
constexpr int SMEM_BYTES = MAX_KB * SLAB_BYTES + B_STAGES * SLAB_BYTES + MISC_FLOATS * 4 + 256 /*barriers*/ + 1024 /*align*/;
constexpr int REC_HEAD = 4;                              // record = [m, l, -, -, acc[L]]

struct Params {
  int64_t n_rows;
  int L, NC, D, gated, n_tiles, tmem_cols;
  const float* bab;
  const float* wc;
  const float* bc;
  const int64_t* offsets;
  const int32_t* row_seg;
  __nv_bfloat16* uv;      // may be null (inference: nothing is kept for a backward pass)
  float* s;
  float* rec;
};

__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }

__global__ void __launch_bounds__(NUM_THREADS, 1)
attnpool_fwd_kernel(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_w, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t a_base = (raw + 1023u) & ~1023u;                     // SWIZZLE_128B atoms need 1024 B alignment
  const uint32_t b_base = a_base + MAX_KB * SLAB_BYTES;
  const uint32_t misc = b_base + B_STAGES * SLAB_BYTES;
  float* s_tile = reinterpret_cast<float*>(smem_raw + (misc - raw));  // [128] raw scores of the tile
  float* sc_part = s_tile + TILE_M;                                   // [2][128] partial dot products of the column halves
  float* e_tile = sc_part + 2 * TILE_M;                               // [128] exp(s - m) of the current segment
  float* bab_s = e_tile + TILE_M;                                     // [NC]
  float* wc_s = bab_s + MAX_NC;                                       // [D]
  const uint32_t bars = misc + MISC_FLOATS * 4;
  auto a_full = [&](int kb) { return bars + 8u * kb; };
  auto b_full = [&](int s) { return bars + 8u * (MAX_KB + s); };
  auto b_empty = [&](int s) { return bars + 8u * (MAX_KB + B_STAGES + s); };
  const uint32_t a_empty = bars + 8u * (MAX_KB + 2 * B_STAGES);
  const uint32_t tfull = a_empty + 8u, tempty = a_empty + 16u, tmem_slot = a_empty + 24u;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = p.L / BLOCK_K, n_pass = p.NC / PASS_N;

  if (threadIdx.x == 0) {
    for (int kb = 0; kb < MAX_KB; ++kb) mbar_init(a_full(kb), 1);
    for (int s = 0; s < B_STAGES; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    mbar_init(a_empty, NUM_EPI_WARPS);
    mbar_init(tfull, 1);
    mbar_init(tempty, NUM_EPI_WARPS);
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
      int bs = 0, it = 0;
      uint32_t bph = 0;
      for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++it) {
        const int row0 = t * TILE_M;
        if (it > 0) mbar_wait(a_empty, (uint32_t)((it - 1) & 1));      // the pooling pass is done with the previous tile
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_expect_tx(a_full(kb), SLAB_BYTES);
          tma_load_2d(a_base + kb * SLAB_BYTES, &map_h, a_full(kb), kb * BLOCK_K, row0);   // rows >= n_rows arrive as zeros
        }
        if (t + (int)gridDim.x < p.n_tiles)                              // the tile after this one: HBM -> L2 now
          for (int kb = 0; kb < nkb; ++kb) tma_prefetch_l2_2d(&map_h, kb * BLOCK_K, (t + (int)gridDim.x) * TILE_M);
        for (int j = 0; j < n_pass; ++j) {
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(b_empty(bs), bph ^ 1u);
            mbar_expect_tx(b_full(bs), SLAB_BYTES);
            tma_load_2d(b_base + bs * SLAB_BYTES, &map_w, b_full(bs), kb * BLOCK_K, j * PASS_N);
            if (++bs == B_STAGES) { bs = 0; bph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TILE_M, PASS_N, false, false);
      int bs = 0, it = 0;
      uint32_t bph = 0;
      for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++it) {
        if (it > 0) mbar_wait(tempty, (uint32_t)((it - 1) & 1));        // the epilogue has read the previous accumulators
        tcgen05_fence_after();
        for (int j = 0; j < n_pass; ++j) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(j * PASS_N);
          for (int kb = 0; kb < nkb; ++kb) {
            if (j == 0) mbar_wait(a_full(kb), (uint32_t)(it & 1));
            mbar_wait(b_full(bs), bph);
            tcgen05_fence_after();
            const uint32_t a_src = a_base + kb * SLAB_BYTES;
            const uint32_t b_src = b_base + bs * SLAB_BYTES;
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              tcgen05_mma_bf16(d_tmem, make_smem_desc(a_src + k * 32, 16, 1024), make_smem_desc(b_src + k * 32, 16, 1024),
                               idesc, (kb > 0 || k > 0) ? 1u : 0u);
            tcgen05_commit(b_empty(bs));                                 // weight slot reusable once these MMAs retire
            if (++bs == B_STAGES) { bs = 0; bph ^= 1u; }
          }
        }
        tcgen05_commit(tfull);                                           // all NC accumulator columns complete
      }
    }
  } else {
    // ================= epilogue + pooling warps (0..7) =================
    const int quarter = warp & 3, chalf = warp >> 2;
    const int r = quarter * 32 + lane;                                   // row of the tile = TMEM lane
    const int tid = threadIdx.x;                                         // 0..255
    const int dh = p.D >> 1;                                             // D columns per warp of a quarter (multiple of 32)
    const int D = p.D, NC = p.NC, L = p.L;
    const bool gated = p.gated != 0;
    const float bc = p.bc ? p.bc[0] : 0.f;
    const int64_t rec_stride = L + REC_HEAD;
    int it = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++it) {
      const int64_t row0 = (int64_t)t * TILE_M;
      const int64_t row = row0 + r;
      const bool valid = row < p.n_rows;
      mbar_wait(tfull, (uint32_t)(it & 1));
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
      float part = 0.f;
      for (int d0 = chalf * dh; d0 < (chalf + 1) * dh; d0 += 32) {
        uint32_t ra[32], rb[32];
        tmem_ld32(t_row + (uint32_t)d0, ra);
        if (gated) tmem_ld32(t_row + (uint32_t)(D + d0), rb);
        tmem_ld_wait();
        uint32_t up[16], vp[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          // the saved (bf16-rounded) activations are what the score uses: forward and backward see the same numbers
          const float u0 = tanh_fast(__uint_as_float(ra[i]) + bab_s[d0 + i]);
          const float u1 = tanh_fast(__uint_as_float(ra[i + 1]) + bab_s[d0 + i + 1]);
          const __nv_bfloat162 ub = __floats2bfloat162_rn(u0, u1);
          up[i >> 1] = *reinterpret_cast<const uint32_t*>(&ub);
          float2 g = __bfloat1622float2(ub);
          if (gated) {
            const float v0 = fmaf(0.5f, tanh_fast(0.5f * (__uint_as_float(rb[i]) + bab_s[D + d0 + i])), 0.5f);
            const float v1 = fmaf(0.5f, tanh_fast(0.5f * (__uint_as_float(rb[i + 1]) + bab_s[D + d0 + i + 1])), 0.5f);
            const __nv_bfloat162 vb = __floats2bfloat162_rn(v0, v1);
            vp[i >> 1] = *reinterpret_cast<const uint32_t*>(&vb);
            const float2 gv = __bfloat1622float2(vb);
            g.x *= gv.x;
            g.y *= gv.y;
          }
          part = fmaf(wc_s[d0 + i], g.x, part);
          part = fmaf(wc_s[d0 + i + 1], g.y, part);
        }
        if (p.uv != nullptr && valid) {
          uint4* dst = reinterpret_cast<uint4*>(p.uv + row * NC + d0);   // 64 contiguous bytes per row
#pragma unroll
          for (int c = 0; c < 4; ++c) dst[c] = make_uint4(up[4 * c], up[4 * c + 1], up[4 * c + 2], up[4 * c + 3]);
          if (gated) {
            uint4* dv = reinterpret_cast<uint4*>(p.uv + row * NC + D + d0);
#pragma unroll
            for (int c = 0; c < 4; ++c) dv[c] = make_uint4(vp[4 * c], vp[4 * c + 1], vp[4 * c + 2], vp[4 * c + 3]);
          }
        }
      }
      sc_part[chalf * TILE_M + r] = part;
      tcgen05_fence_before();
      epi_sync();
      if (lane == 0) mbar_arrive(tempty);                                // TMEM may be overwritten by the next tile's MMAs
      if (tid < TILE_M) {                                                // tid == r for the warps of column half 0
        const float sv = sc_part[tid] + sc_part[TILE_M + tid] + bc;
        s_tile[tid] = sv;
        if (valid) p.s[row] = sv;
      }
      epi_sync();
      // ---- per bag segment of the tile: softmax statistics + exp-weighted column sums from the resident tile ----
      const int64_t last_row = (row0 + TILE_M - 1 < p.n_rows ? row0 + TILE_M - 1 : p.n_rows - 1);
      const int b_lo = p.row_seg[row0], b_hi = p.row_seg[last_row];
      for (int b = b_lo; b <= b_hi; ++b) {
        const int64_t o0 = p.offsets[b], o1 = p.offsets[b + 1];
        const int r_begin = (int)(o0 > row0 ? o0 - row0 : 0);
        const int r_end = (int)(o1 - row0 < TILE_M ? o1 - row0 : TILE_M);
        if (r_end <= r_begin) continue;                                  // empty bag (uniform branch)
        float m = -INFINITY;
        for (int rr = r_begin + lane; rr < r_end; rr += 32) m = fmaxf(m, s_tile[rr]);
        m = warp_max(m);                                                 // every warp computes the same value
        if (tid < TILE_M) e_tile[tid] = (tid >= r_begin && tid < r_end) ? expf(s_tile[tid] - m) : 0.f;
        epi_sync();
        if (warp < nkb) {
          float acc0 = 0.f, acc1 = 0.f;
          const uint32_t slab = a_base + (uint32_t)warp * SLAB_BYTES + (((uint32_t)lane & 3u) << 2);
          const uint32_t chunk = (uint32_t)lane >> 2;
#pragma unroll 4
          for (int rr = r_begin; rr < r_end; ++rr) {
            uint32_t wv;
            const uint32_t addr = slab + (uint32_t)rr * 128u + ((chunk ^ ((uint32_t)rr & 7u)) << 4);
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wv) : "r"(addr) : "memory");
            const float2 f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&wv));
            const float e = e_tile[rr];
            acc0 = fmaf(e, f.x, acc0);
            acc1 = fmaf(e, f.y, acc1);
          }
          float* rec = p.rec + (int64_t)(t + b) * rec_stride;
          if (warp == 0) {
            float l = 0.f;
            for (int rr = r_begin + lane; rr < r_end; rr += 32) l += e_tile[rr];
            l = warp_sum(l);
            if (lane == 0) {
              rec[0] = m;
              rec[1] = l;
            }
          }
          *reinterpret_cast<float2*>(rec + REC_HEAD + BLOCK_K * warp + 2 * lane) = make_float2(acc0, acc1);
        }
        epi_sync();                                                      // e_tile is rewritten for the next segment
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(a_empty);                               // this warp no longer reads the resident tile
    }
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
  return dtype == MURCL_BF16 && L >= 64 && L <= ap::MAX_KB * ap::BLOCK_K && L % ap::BLOCK_K == 0 && D % 64 == 0 &&
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
    static bool configured = false;
    if (!configured) {
      MURCL_CUDA(cudaFuncSetAttribute(ap::attnpool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ap::SMEM_BYTES));
      configured = true;
    }
    ap::Params prm{};
    prm.n_rows = n_rows; prm.L = L; prm.NC = nc; prm.D = D; prm.gated = gated;
    prm.n_tiles = (int)((n_rows + ap::TILE_M - 1) / ap::TILE_M);
    prm.tmem_cols = ap::tmem_cols_for(nc);
    prm.bab = bab; prm.wc = wc; prm.bc = bc; prm.offsets = offsets; prm.row_seg = row_seg;
    prm.uv = static_cast<__nv_bfloat16*>(uv); prm.s = s; prm.rec = workspace;
    const int grid = prm.n_tiles < sm_count() ? prm.n_tiles : sm_count();
    ap::attnpool_fwd_kernel<<<grid, ap::NUM_THREADS, ap::SMEM_BYTES, st>>>(mh, mw, prm);
    int rc2 = check_launch("attnpool_fwd_kernel");
    if (rc2 != MURCL_OK) return rc2;
  }
  ap::attnpool_merge_kernel<<<B, 256, 0, st>>>(workspace, s, offsets, L, inv_sqrt_n, M, p, stats);
  return check_launch("attnpool_merge_kernel");
}

}  // extern "C"
