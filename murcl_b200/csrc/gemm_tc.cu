// tcgen05 / TMEM / TMA GEMM family for the bf16 mode of murcl_linear_{fwd,bwd_input,bwd_weight}.
//
//   C[m,n] = sum_k A(m,k) * B(n,k)      bf16 operands, fp32 accumulators in tensor memory
//
// One persistent CTA per SM (or CTA pair per SM pair), 18 warps:
//   warps 0-15 epilogue    four per TMEM lane quarter, each taking every fourth 128-byte column slab: tcgen05.ld 32 lanes x 32
//                          columns -> registers -> fused epilogue (bias/activation/ReLU bits, or pooling row term / ReLU mask /
//                          scale / column sums) -> 128B-swizzled smem slab (32 rows x 128 B) -> cp.async.bulk.tensor store.
//                          Sixteen warps (<= 112 registers each, a slab passes through the registers in two 32-column halves)
//                          give every scheduler four epilogue warps to hide the tcgen05.ld / L2 latencies behind each other.
//   warp 16    TMA producer cp.async.bulk.tensor.2d -> 128B-swizzled smem ring (STAGES deep), mbarrier tx-count
//   warp 17    MMA issuer   one lane issues tcgen05.mma.kind::f16 (128*CG x BN x 16), tcgen05.commit frees smem slots and
//                          publishes the accumulator; also owns tcgen05.alloc/dealloc
// Two TMEM accumulator stages (2 x BN columns) let the epilogue of tile i overlap the MMAs of tile i+1.
//
// Operand majors (UMMA "K-major" = reduction index contiguous in memory, "MN-major" = output index contiguous):
//   forward      y  = x   w^T        A = x  [M,K]  K-major      B = w [N,K]   K-major
//   input grad   dx = dy  w          A = dy [M,N]  K-major      B = w [N,K]   MN-major (rows = reduction index n)
//   weight grad  dw = dy^T x         A = dy [M,N]  MN-major     B = x [M,K]   MN-major (rows = reduction index m), split-K
// so no operand is ever transposed in memory.
//
// Roofline: tensor pipe.  Algorithmic FLOPs = 2*M*N*K per launch.
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

// defined in gemm_simt.cu
int launch_splitk_reduce(const float* ws, int splits, int64_t stride, float* out, int64_t n, cudaStream_t st, int accumulate);

namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;               // 64 bf16 = 128 B = one swizzle span
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 16;          // four per TMEM lane quarter, splitting the column slabs
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;          // 16 KB
constexpr int SMEM_BUDGET = 224 * 1024;

enum Epi { EPI_FWD = 0, EPI_DGRAD = 1, EPI_SPLIT = 2 };

struct Params {
  int64_t M;          // rows of C
  int N;              // cols of C
  int64_t K;          // reduction length
  int64_t ldc;
  void* C;
  const float* bias;
  int act;
  const __nv_bfloat16* relu_src;
  const float* row_scale;
  const float* row_vec;
  const int32_t* row_seg;
  float* col_sum;     // EPI_DGRAD: accumulated column sums of the stored result (may be null)
  float out_scale;    // EPI_DGRAD: final scale (1/(1-p) of a dropout that followed the masked ReLU)
  unsigned long long* bits_out;        // EPI_FWD + ReLU: 1 bit per output (y > 0), layout [N/64][M] (may be null)
  const unsigned long long* bits_in;   // EPI_DGRAD: the same bit mask instead of re-reading relu_src (may be null)
  int debug;          // MURCL_DEBUG_EPI bit mask (timing experiments only; results are wrong when set)
  unsigned long long* trace;   // debug & 8: CTA 0 writes %globaltimer stamps per tile: [tile][mma_start, mma_end, epi_start, epi_end]
  int splits;
  int64_t k_chunk;    // reduction range per split (multiple of BLOCK_K)
  int64_t split_stride;
  int m_tiles, n_tiles;
  // Split-precision mode (exact-fp32 GEMMs on the bf16 tensor cores): the operands are stacks of bf16 PLANES of an fp32
  // tensor (x = hi + mid [+ lo], planes stacked along the rows, `*_plane_rows` apart) and a tile accumulates `terms`
  // products A_plane[pa[i]] . B_plane[pb[i]] over the same reduction range into one fp32 TMEM accumulator, smallest
  // terms first.  terms <= 1: plain bf16 GEMM.
  int terms;
  int pa[6], pb[6];
  int a_plane_rows, b_plane_rows;
  // L2 prefetch distance of the TMA producer (cp.async.bulk.prefetch.tensor): B-stationary kernels pull the A k-blocks of
  // the tile `l2pf` tiles ahead into L2, streaming kernels the operand k-blocks `l2pf` k-blocks ahead.  The shared-memory
  // ring alone looks ahead ~1 us of MMA time (4 x 16 KB of A beside the resident B tile), less than the loaded HBM latency.
  int l2pf;
  int reverse;        // walk the row tiles from the last to the first (murcl_set_row_order)
  int atomic_out;     // EPI_SPLIT: add the partial tile to C with vector atomics (red.global.add.v4.f32) instead of storing it
};

// CG = cta_group: 1 = one SM per 128 x BN tile; 2 = a CTA pair shares a 256 x BN tile (each CTA holds its 128 rows of A
// and HALF of B, the MMA unit exchanges the B halves), which halves the B bytes each SM pulls through L2 and reads from
// shared memory per MMA.
// BSTAT = B-stationary: the whole reduction extent (<= 8 k-blocks) of the CTA's B tile stays resident in shared memory and
// only A streams through the ring.  The persistent schedule gives a CTA (pair) the same column tile every time
// (tile step % n_tiles == 0), so B is loaded once per launch: L2 -> SM operand traffic per MMA halves again.  The L2
// slice bandwidth (~6300 B/clk chip-wide) is what bounded the streaming kernel (measured: 8 k-blocks in 4.4 us).
template <int BN, int EPI, int CG, int TOUT_BYTES, bool BSTAT>
struct Cfg {
  static constexpr int B_BYTES = (BN / CG) * BLOCK_K * 2;
  static constexpr int MAX_RES_KB = 8;                                        // resident k-blocks (reduction <= 512)
  static constexpr int BRES_BYTES = BSTAT ? MAX_RES_KB * B_BYTES : 0;
  static constexpr int STAGE_BYTES = BSTAT ? A_BYTES : A_BYTES + B_BYTES;
  // one staging unit per epilogue warp: 32 rows x 32 columns (64-byte rows of bf16, 128-byte rows of fp32);
  // the split-K kernel stores directly
  static constexpr int UNIT_BYTES = 32 * 32 * TOUT_BYTES;
  static constexpr int STAGING_BYTES = (EPI == 2) ? 0 : NUM_EPI_WARPS * UNIT_BYTES;
  static constexpr int RING_BYTES = SMEM_BUDGET - STAGING_BYTES - BRES_BYTES;
  static constexpr int STAGES = RING_BYTES / STAGE_BYTES > 8 ? 8 : RING_BYTES / STAGE_BYTES;
  static_assert(STAGES >= 3, "operand ring too shallow");
  static constexpr int TMEM_COLS = 2 * BN;                  // 256 or 512: powers of two
  static constexpr int SMEM_BYTES = BRES_BYTES + STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ +
                                    BN * 4 /*column-sum accumulator*/;
};

#define MURCL_STAMP(slot)                                                                         \
  if (MURCL_TRACE_ON && p.trace && blockIdx.x == 0 && threadIdx.x == 0 && c0 == 0) {                              \
    unsigned long long ts_;                                                                     \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts_));                                     \
    p.trace[(int64_t)24 * 2048 + ((t - tile_first) / tile_step) * 8 + (slot)] = ts_;            \
  }

template <int BN, bool A_MN, bool B_MN, int EPI, typename TOUT, int CG, bool BSTAT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_c, const Params p) {
  using C = Cfg<BN, EPI, CG, (int)sizeof(TOUT), BSTAT>;
  static_assert(!BSTAT || (!A_MN && EPI != EPI_SPLIT), "B-stationary: forward / input-gradient kernels only");
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0u;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t bres = (raw + 1023u) & ~1023u;                        // SWIZZLE_128B atoms need 1024 B alignment
  const uint32_t tiles = bres + C::BRES_BYTES;                         // resident B k-blocks (BSTAT), then the operand ring
  const uint32_t staging = tiles + C::STAGES * C::STAGE_BYTES;        // 1024 B aligned (stage sizes are multiples of 1024)
  const uint32_t bars = staging + C::STAGING_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + 2 + s); };
  const uint32_t bres_bar = bars + 8u * (2 * C::STAGES + 4);
  const uint32_t tmem_slot = bars + 8u * (2 * C::STAGES + 5);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));
  float* colacc = reinterpret_cast<float*>(smem_raw + (bars + 256u - raw));     // [BN] per-CTA column sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Roles.  The SM's warp scheduler favours the HIGHEST warp id among ready warps, so the two latency-critical
  // single-lane roles (TMA producer, MMA issuer) get the top ids; otherwise the ALU-heavy epilogue warps sharing
  // their scheduler starve them and the tensor pipe idles.
  constexpr int PRODUCER_WARP = NUM_EPI_WARPS, MMA_WARP = NUM_EPI_WARPS + 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), NUM_EPI_WARPS * CG);     // one arrival per epilogue warp (of both CTAs of a pair)
    }
    mbar_init(bres_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
  }
  if (CG == 2) cluster_sync_all();         // barrier inits of both CTAs are visible before any remote arrive / TMA
  if (warp == MMA_WARP) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(C::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(C::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // 32-bit tile arithmetic: 64-bit div/mod is ~1000 cycles for the lone MMA-issuing lane, once per tile, on the
  // critical path (measured as a 0.6 us bubble between tiles)
  const int total_tiles = p.m_tiles * p.n_tiles * p.splits;                  // m_tiles counts 128*CG-row tiles
  const int tile_first = (int)blockIdx.x / CG, tile_step = (int)gridDim.x / CG;
  const int k_blocks_full = (int)((p.k_chunk + BLOCK_K - 1) / BLOCK_K);

  auto tile_coords = [&](int t, int& mb, int& nb, int& sp) {
    const unsigned ut = (unsigned)t, nt = (unsigned)p.n_tiles, mt = (unsigned)p.m_tiles;
    const unsigned r = ut / nt;
    nb = (int)(ut - r * nt);
    sp = (int)(r / mt);
    mb = (int)(r - (unsigned)sp * mt);
    if (p.reverse) mb = p.m_tiles - 1 - mb;
  };
  auto k_range = [&](int sp, int64_t& k0, int& nkb) {
    k0 = (int64_t)sp * p.k_chunk;
    const int64_t k1 = (k0 + p.k_chunk < p.K) ? k0 + p.k_chunk : p.K;
    nkb = (int)((k1 - k0 + BLOCK_K - 1) / BLOCK_K);
    if (nkb > k_blocks_full) nkb = k_blocks_full;
  };

  if (warp == PRODUCER_WARP) {
    // ================= TMA producer =================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      if (BSTAT && tile_first < total_tiles) {
        // the CTA's column tile never changes (host: tile step % n_tiles == 0): fetch its B k-blocks once
        int mb, nb, sp, nkb;
        int64_t k0;
        tile_coords(tile_first, mb, nb, sp);
        k_range(sp, k0, nkb);
        const int n0 = nb * BN + (int)cta_rank * (BN / CG);
        if (CG == 1 || leader) mbar_expect_tx(bres_bar, (uint32_t)(CG * nkb * C::B_BYTES));
        const uint32_t bb = (CG == 2) ? mapa(bres_bar, 0) : bres_bar;
        for (int kb = 0; kb < nkb; ++kb) {
          const uint32_t b_dst = bres + kb * C::B_BYTES;
          const int kk = kb * BLOCK_K;
          if (!B_MN) {
            if (CG == 2) tma_load_2d_2sm(b_dst, &map_b, bb, kk, n0);
            else tma_load_2d(b_dst, &map_b, bb, kk, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / CG / 64; ++j) {
              if (CG == 2) tma_load_2d_2sm(b_dst + j * 8192, &map_b, bb, n0 + 64 * j, kk);
              else tma_load_2d(b_dst + j * 8192, &map_b, bb, n0 + 64 * j, kk);
            }
          }
        }
      }
      for (int t = tile_first; t < total_tiles; t += tile_step) {
        int mb, nb, sp, nkb;
        int64_t k0;
        tile_coords(t, mb, nb, sp);
        k_range(sp, k0, nkb);
        const int m0 = (mb * CG + (int)cta_rank) * BLOCK_M;                   // this CTA's 128 rows of the tile
        const int n0 = nb * BN + (int)cta_rank * (BN / CG);                    // this CTA's share of the B rows
        const int nterms = p.terms > 1 ? p.terms : 1;
        if (BSTAT && p.l2pf > 0) {
          // the A rows of a later tile of this CTA: in L2 by the time the ring asks for them
          const int tn = t + p.l2pf * tile_step;
          if (tn < total_tiles) {
            int mbn, nbn, spn;
            tile_coords(tn, mbn, nbn, spn);
            const int m0n = (mbn * CG + (int)cta_rank) * BLOCK_M;
            for (int kb = 0; kb < nkb; ++kb) tma_prefetch_l2_2d(&map_a, kb * BLOCK_K, m0n);
          }
        }
        for (int term = 0; term < nterms; ++term) {
        const int ra = p.terms > 1 ? p.pa[term] * p.a_plane_rows : 0;         // row offset of this term's A / B plane
        const int rb = p.terms > 1 ? p.pb[term] * p.b_plane_rows : 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t a_dst = tiles + s * C::STAGE_BYTES;
          const uint32_t b_dst = a_dst + A_BYTES;
          const int kk = (int)(k0 + (int64_t)kb * BLOCK_K);
          if (BSTAT) {                                                        // only A streams
            if (CG == 2) {
              if (leader) mbar_expect_tx(full_bar(s), 2 * A_BYTES);
              tma_load_2d_2sm(a_dst, &map_a, mapa(full_bar(s), 0), kk, m0);
            } else {
              mbar_expect_tx(full_bar(s), A_BYTES);
              tma_load_2d(a_dst, &map_a, full_bar(s), kk, m0);
            }
            if (++s == C::STAGES) { s = 0; ph ^= 1u; }
            continue;
          }
          if (!BSTAT && p.l2pf > 0 && p.terms <= 1 && kb + p.l2pf < nkb) {
            const int kp = kk + p.l2pf * BLOCK_K;
            if (!A_MN) {
              tma_prefetch_l2_2d(&map_a, kp, m0);
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_M / 64; ++j) tma_prefetch_l2_2d(&map_a, m0 + 64 * j, kp);
            }
            if (!B_MN) {
              tma_prefetch_l2_2d(&map_b, kp, n0);
            } else {
#pragma unroll
              for (int j = 0; j < BN / CG / 64; ++j) tma_prefetch_l2_2d(&map_b, n0 + 64 * j, kp);
            }
          }
          if (CG == 2) {
            // both CTAs load their halves; all bytes are counted on the LEADER's full barrier (it issues the MMAs)
            if (leader) mbar_expect_tx(full_bar(s), 2 * C::STAGE_BYTES);
            const uint32_t fb = mapa(full_bar(s), 0);
            if (!A_MN) {
              tma_load_2d_2sm(a_dst, &map_a, fb, kk, m0 + ra);
            } else {                                                            // weight gradient: two 64-wide MN blocks
#pragma unroll
              for (int j = 0; j < BLOCK_M / 64; ++j) tma_load_2d_2sm(a_dst + j * 8192, &map_a, fb, m0 + 64 * j, kk + ra);
            }
            if (!B_MN) {
              tma_load_2d_2sm(b_dst, &map_b, fb, kk, n0 + rb);
            } else {
#pragma unroll
              for (int j = 0; j < BN / CG / 64; ++j) tma_load_2d_2sm(b_dst + j * 8192, &map_b, fb, n0 + 64 * j, kk + rb);
            }
            if (++s == C::STAGES) { s = 0; ph ^= 1u; }
            continue;
          }
          mbar_expect_tx(full_bar(s), C::STAGE_BYTES);
          if (!A_MN) {
            tma_load_2d(a_dst, &map_a, full_bar(s), kk, m0 + ra);               // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)                               // box {64 m, 64 k-rows}
              tma_load_2d(a_dst + j * 8192, &map_a, full_bar(s), m0 + 64 * j, kk + ra);
          }
          if (!B_MN) {
            tma_load_2d(b_dst, &map_b, full_bar(s), kk, n0 + rb);               // box {64 k, BN rows}
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(b_dst + j * 8192, &map_b, full_bar(s), n0 + 64 * j, kk + rb);
          }
          if (++s == C::STAGES) { s = 0; ph ^= 1u; }
        }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ================= MMA issuer (the leader CTA of a pair issues for both) =================
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc(BLOCK_M * CG, BN, A_MN, B_MN);
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      if (BSTAT && tile_first < total_tiles) mbar_wait(bres_bar, 0u);   // the resident B tile has landed (both halves of a pair)
      for (int t = tile_first; t < total_tiles; t += tile_step) {
        int mb, nb, sp, nkb;
        int64_t k0;
        tile_coords(t, mb, nb, sp);
        k_range(sp, k0, nkb);
        mbar_wait(tempty_bar(as), aph ^ 1u);
        tcgen05_fence_after();
        if (MURCL_TRACE_ON && p.trace && blockIdx.x == 0) {
          unsigned long long ts;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts));
          p.trace[((t - tile_first) / tile_step) * 24 + 0] = ts;
        }
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        const int nkb_all = nkb * (p.terms > 1 ? p.terms : 1);              // split precision: every term adds nkb k-blocks
        for (int kb = 0; kb < nkb_all; ++kb) {
          mbar_wait(full_bar(s), ph);
          tcgen05_fence_after();
          const uint32_t a_src = tiles + s * C::STAGE_BYTES;
          const uint32_t b_src = BSTAT ? bres + kb * C::B_BYTES : a_src + A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // K-major: 16 k-elements = 32 B inside the 128 B swizzle span; SBO = 1024 B between 8-row groups.
            // MN-major: 16 k-rows = two 8-row groups of 1024 B; LBO = 8192 B between 64-wide MN blocks.
            const uint64_t adesc = A_MN ? make_smem_desc(a_src + k * 2048, 8192, 1024) : make_smem_desc(a_src + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? make_smem_desc(b_src + k * 2048, 8192, 1024) : make_smem_desc(b_src + k * 32, 16, 1024);
            if (CG == 1) tcgen05_mma_bf16(d_tmem, adesc, bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            else tcgen05_mma_bf16_2sm(d_tmem, adesc, bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          if (CG == 1) tcgen05_commit(empty_bar(s));       // smem slot reusable once these MMAs retire
          else tcgen05_commit_2sm(empty_bar(s));           // ... in both CTAs
          if (++s == C::STAGES) { s = 0; ph ^= 1u; }
        }
        if (CG == 1) tcgen05_commit(tfull_bar(as));        // accumulator complete
        else tcgen05_commit_2sm(tfull_bar(as));
        if (MURCL_TRACE_ON && p.trace && blockIdx.x == 0) {
          unsigned long long ts;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts));
          p.trace[((t - tile_first) / tile_step) * 24 + 1] = ts;      // all MMAs of the tile ISSUED
        }
        if (++as == 2) { as = 0; aph ^= 1u; }
      }
    }
  } else {
    // ================= epilogue warps (0..15) =================
    // Work unit = 32 rows (this warp's TMEM lane quarter) x 32 columns: one tcgen05.ld.32x32b.x32, one staging unit, one
    // bulk tensor store.  The four warps of a quarter take every fourth unit.
    const int quarter = warp & 3;                          // TMEM lanes [32*quarter, +32) belong to this warp
    const int ew = warp;                                   // 0..15
    const int part = ew >> 2;
    const uint32_t out_unit = staging + (uint32_t)ew * C::UNIT_BYTES;
    // bf16: 64-byte rows, SWIZZLE_64B (16-byte chunk index ^= (row >> 1) & 3); fp32: 128-byte rows, SWIZZLE_128B (^= row & 7)
    const uint32_t my_row_off = (uint32_t)lane * (32u * (uint32_t)sizeof(TOUT));
    const uint32_t swz = sizeof(TOUT) == 2 ? (uint32_t)((lane >> 1) & 3) : (uint32_t)(lane & 7);
    constexpr int EPI_THREADS = 32 * NUM_EPI_WARPS;
    int as = 0;
    uint32_t aph = 0;
    const bool has_mask = (EPI == EPI_DGRAD) && p.relu_src != nullptr && p.bits_in == nullptr;
    const bool want_colsum = (EPI == EPI_DGRAD) && sizeof(TOUT) == 2 && p.col_sum != nullptr;
    int acc_nb = -1;
    const int etid = threadIdx.x;                           // 0..511: the epilogue threads come first
    auto epi_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); };
    // Column sums (bias gradient of the layer below) live in REGISTERS across the tiles of a column tile: the persistent
    // schedule hands a CTA the same column tile again and again, so a warp meets the same (at most NU) 32-column units
    // every tile; lane l accumulates columns 2 (l & 15), +1 of every second row.  They go to global memory (atomics) only
    // when the column tile changes and at the end of the kernel - no shared-memory atomics, no CTA barrier per tile.
    constexpr int NU = (BN + 127) / 128;                    // units of a tile per epilogue warp
    float cs[NU][2];
#pragma unroll
    for (int u = 0; u < NU; ++u) cs[u][0] = cs[u][1] = 0.f;
    auto flush_colsum = [&](int nb_flush) {
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const float s0 = cs[u][0] + __shfl_xor_sync(0xffffffffu, cs[u][0], 16);
        const float s1 = cs[u][1] + __shfl_xor_sync(0xffffffffu, cs[u][1], 16);
        const int col = nb_flush * BN + part * 32 + 128 * u + 2 * (lane & 15);
        if (lane < 16 && part * 32 + 128 * u < BN) {
          if (col < p.N && s0 != 0.f) atomicAdd(p.col_sum + col, s0);
          if (col + 1 < p.N && s1 != 0.f) atomicAdd(p.col_sum + col + 1, s1);
        }
        cs[u][0] = cs[u][1] = 0.f;
      }
    };
    for (int t = tile_first; t < total_tiles; t += tile_step) {
      int mb, nb, sp;
      tile_coords(t, mb, nb, sp);
      const int64_t row0 = ((int64_t)mb * CG + cta_rank) * BLOCK_M + quarter * 32;
      const int64_t row = row0 + lane;
      const int n0 = nb * BN;
      // per-row operands of the pooling term: two dependent global loads, issued before the accumulator wait so that their
      // latency is hidden behind it (the wait is a compiler barrier: loads placed after it would start after it)
      float rs = 0.f;
      const float* rv = nullptr;
      bool same_bag = false;                                // warp-uniform: all 32 rows of this warp belong to one bag
      float gl[NU];                                         // same_bag: lane l holds row_vec[bag][n0 + c0(u) + l]
      unsigned int wb[NU];                                  // ReLU bit word of (row, unit u)
#pragma unroll
      for (int u = 0; u < NU; ++u) { gl[u] = 0.f; wb[u] = 0u; }
      if (EPI == EPI_DGRAD) {
        if (p.row_scale != nullptr) {
          int seg = -1;
          if (row < p.M) {
            rs = __ldg(p.row_scale + row);
            seg = __ldg(p.row_seg + row);
            rv = p.row_vec + (int64_t)seg * p.N;
          }
          const int seg0 = __shfl_sync(0xffffffffu, seg, 0);
          same_bag = seg0 >= 0 && __all_sync(0xffffffffu, seg == seg0 || seg < 0);
          if (same_bag) {
            // ONE coalesced 128-byte read of the bag's row vector per unit (instead of 8 broadcast float4 reads per
            // lane after the TMEM wait); the values are handed round with shuffles in the FMA loop
            const float* rv0 = p.row_vec + (int64_t)seg0 * p.N;
#pragma unroll
            for (int u = 0; u < NU; ++u) {
              const int col = n0 + part * 32 + 128 * u + lane;
              if (part * 32 + 128 * u < BN && col < p.N) gl[u] = __ldg(rv0 + col);
            }
          }
        }
        if (p.bits_in != nullptr && row < p.M) {
          // the ReLU bit words of the tile's units are requested here as well: their L2 latency used to be exposed after
          // every TMEM load (the top stall of the kernel: R2P waiting for the word)
#pragma unroll
          for (int u = 0; u < NU; ++u) {
            const int col0 = n0 + part * 32 + 128 * u;
            if (part * 32 + 128 * u < BN && col0 < p.N)
              wb[u] = __ldg(reinterpret_cast<const unsigned int*>(p.bits_in) + ((((int64_t)(col0 >> 6) * p.M + row) << 1) + ((col0 >> 5) & 1)));
          }
        }
      }
      mbar_wait(tfull_bar(as), aph);
      tcgen05_fence_after();
      if (MURCL_TRACE_ON && p.trace && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long ts;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts));
        p.trace[((t - tile_first) / tile_step) * 24 + 2] = ts;        // accumulator complete, epilogue starts
      }
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN);
      if (EPI == EPI_SPLIT) {
        // fp32 partial sums straight to the split workspace (few tiles per launch: not worth staging)
#pragma unroll 1
        for (int c0 = part * 32; c0 < BN; c0 += 128) {
          uint32_t r[32];
          tmem_ld32(t_row + (uint32_t)c0, r);
          tmem_ld_wait();
          const int col0 = n0 + c0;
          if (row < p.M && col0 < p.N) {
            float* dst = static_cast<float*>(p.C) + (int64_t)sp * p.split_stride + row * p.ldc + col0;
            if (p.atomic_out) {                            // split_stride == 0: every split adds into the same [M, N] result
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                if (col0 + i < p.N)
                  atomicAdd(reinterpret_cast<float4*>(dst + i), make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                                            __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])));
            } else {
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                if (col0 + i < p.N)
                  *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                                    __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
            }
          }
        }
      } else {
        if (want_colsum && nb != acc_nb) {                 // new column tile: flush this warp's partial sums
          if (acc_nb >= 0) flush_colsum(acc_nb);
          acc_nb = nb;
        }
#pragma unroll
        for (int u = 0; u < NU; ++u) {
          const int c0 = part * 32 + 128 * u;
          if (c0 >= BN) break;
          const int col0 = n0 + c0;
          if (col0 >= p.N || row0 >= p.M) break;            // warp-uniform: nothing of this unit is in range
          MURCL_STAMP(0)
          // ReLU bit masks: one 64-bit word per (row, 64 columns), layout [N/64][M]; a unit owns one 32-bit half of a word.
          // Consecutive rows are consecutive words: the per-lane accesses of a warp stay within 256 contiguous bytes.
          const int64_t bit_word = (((int64_t)(col0 >> 6) * p.M + row) << 1) + ((col0 >> 5) & 1);
          // the ReLU bit word is requested BEFORE the TMEM load, whose wait is a compiler barrier - otherwise its L2 latency
          // would be exposed after it, once per unit (hoisting the 8 float4 operand loads as well costs spills: slower)
          const bool whole = col0 + 32 <= p.N;               // warp-uniform
          // the unit's 128-byte line of the bias / pooling row vector: pulled into L1 now, read after the TMEM wait
          if (EPI == EPI_FWD && p.bias != nullptr) prefetch_l1(p.bias + col0);
          if (EPI == EPI_DGRAD && rv != nullptr && !same_bag) prefetch_l1(rv + col0);
          const unsigned int wbits = wb[u];
          float v[32];
          {
            uint32_t r[32];
            tmem_ld32(t_row + (uint32_t)c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          }
          if (EPI == EPI_FWD) {
            if (p.bias != nullptr) {
              if (whole) {                                  // whole unit in range: branch-free, loads issued back to back
                float4 b4[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) b4[i] = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + i);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  v[4 * i] += b4[i].x; v[4 * i + 1] += b4[i].y; v[4 * i + 2] += b4[i].z; v[4 * i + 3] += b4[i].w;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                  if (col0 + i < p.N) {                     // N % 8 == 0
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
                    v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
                  }
                }
              }
            }
            act_slab<sizeof(TOUT) == 2, 32>(v, p.act, col0, p.N);
            if (sizeof(TOUT) == 2 && p.bits_out != nullptr && row < p.M) {
              // after the ReLU v >= +0, so (v > 0) is the sign bit of the integer negation of its bit pattern;
              // a funnel shift appends one sign bit per instruction; two independent 16-bit chains
              unsigned int q0 = 0u, q1 = 0u;
#pragma unroll
              for (int i = 15; i >= 0; --i) {
                q0 = __funnelshift_l(0u - __float_as_uint(v[i]), q0, 1);
                q1 = __funnelshift_l(0u - __float_as_uint(v[16 + i]), q1, 1);
              }
              reinterpret_cast<unsigned int*>(p.bits_out)[bit_word] = q0 | (q1 << 16);
            }
          } else if (EPI == EPI_DGRAD) {
            if (same_bag) {
              const float gu = gl[u];
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaf(rs, __shfl_sync(0xffffffffu, gu, i), v[i]);
            } else if (rv != nullptr) {
              if (whole) {
                float4 g4[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) g4[i] = __ldg(reinterpret_cast<const float4*>(rv + col0) + i);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  v[4 * i] = fmaf(rs, g4[i].x, v[4 * i]); v[4 * i + 1] = fmaf(rs, g4[i].y, v[4 * i + 1]);
                  v[4 * i + 2] = fmaf(rs, g4[i].z, v[4 * i + 2]); v[4 * i + 3] = fmaf(rs, g4[i].w, v[4 * i + 3]);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                  if (col0 + i < p.N) {
                    const float4 g4 = __ldg(reinterpret_cast<const float4*>(rv + col0 + i));
                    v[i] = fmaf(rs, g4.x, v[i]); v[i + 1] = fmaf(rs, g4.y, v[i + 1]);
                    v[i + 2] = fmaf(rs, g4.z, v[i + 2]); v[i + 3] = fmaf(rs, g4.w, v[i + 3]);
                  }
                }
              }
            }
            if (p.bits_in != nullptr) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (!((wbits >> i) & 1u)) v[i] = 0.f;
            } else if (has_mask) {
              // no bit mask from the forward pass: read the ReLU output itself (64 contiguous bytes per row)
              if (row < p.M) {
                const uint4* src = reinterpret_cast<const uint4*>(p.relu_src + row * p.ldc + col0);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  if (col0 + 8 * j < p.N) {
                    const uint4 q = __ldg(src + j);
                    const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      const float2 f = __bfloat1622float2(hh[e]);
                      if (!(f.x > 0.f)) v[8 * j + 2 * e] = 0.f;
                      if (!(f.y > 0.f)) v[8 * j + 2 * e + 1] = 0.f;
                    }
                  }
                }
              }
            }
            if ((p.bits_in != nullptr || has_mask) && p.out_scale != 1.f) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] *= p.out_scale;
            }
          }
          MURCL_STAMP(1)
          if (lane == 0) bulk_wait_read0();                 // this warp's previous store has read the staging unit
          __syncwarp();
          MURCL_STAMP(2)
          // stage (swizzled like the TMA box) and hand the unit to the bulk-store engine
          if (sizeof(TOUT) == 2) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              st_shared_v4(out_unit + my_row_off + (((uint32_t)c ^ swz) << 4), pack_bf16(v[8 * c], v[8 * c + 1]),
                           pack_bf16(v[8 * c + 2], v[8 * c + 3]), pack_bf16(v[8 * c + 4], v[8 * c + 5]),
                           pack_bf16(v[8 * c + 6], v[8 * c + 7]));
          } else {
#pragma unroll
            for (int c = 0; c < 8; ++c)
              st_shared_v4(out_unit + my_row_off + (((uint32_t)c ^ swz) << 4), __float_as_uint(v[4 * c]),
                           __float_as_uint(v[4 * c + 1]), __float_as_uint(v[4 * c + 2]), __float_as_uint(v[4 * c + 3]));
          }
          MURCL_STAMP(3)
          fence_async_smem();
          __syncwarp();
          MURCL_STAMP(4)
          if (lane == 0) {
            tma_store_2d(&map_c, out_unit, col0, (int)row0);   // rows >= M and cols >= N are clipped by the tensor map
            bulk_commit();
          }
          MURCL_STAMP(5)
          if (want_colsum) {
            // column sums of the unit as stored (bf16-rounded): lane owns the 4-byte word (lane & 15) of every second row
            float s0 = 0.f, s1 = 0.f;
            const uint32_t wi = (uint32_t)lane & 15u;
#pragma unroll 8
            for (int j = 0; j < 16; ++j) {
              const uint32_t r = 2u * (uint32_t)j + ((uint32_t)lane >> 4);
              uint32_t wv;
              const uint32_t addr = out_unit + r * 64u + ((((wi >> 2) ^ ((r >> 1) & 3u)) << 4) | ((wi & 3u) << 2));
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wv) : "r"(addr) : "memory");
              const float2 f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&wv));
              s0 += f.x;
              s1 += f.y;
            }
            cs[u][0] += s0;
            cs[u][1] += s1;
          }
        }
      }
      if (MURCL_TRACE_ON && p.trace && blockIdx.x < 2 && lane == 0 && ew < 8) {
        unsigned long long ts;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts));
        p.trace[((t - tile_first) / tile_step) * 24 + 4 + blockIdx.x * 8 + ew] = ts;   // per-warp end, CTAs 0 and 1
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {                                     // this warp has drained its part of the accumulator
        if (CG == 1) mbar_arrive(tempty_bar(as));
        else mbar_arrive_cluster(mapa(tempty_bar(as), 0)); // the leader's MMA thread waits for both CTAs
      }
      if (++as == 2) { as = 0; aph ^= 1u; }
    }
    if (want_colsum && acc_nb >= 0) flush_colsum(acc_nb);
    if (EPI != EPI_SPLIT && lane == 0) bulk_wait_all();     // all bulk stores of this warp have completed
  }

  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // pair: neither CTA may leave while the other still uses its smem/TMEM
  if (warp == MMA_WARP) {
    tcgen05_fence_after();
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// 2-D row-major tensor [rows, cols] (cols contiguous) of 2-byte (bf16) or 4-byte (fp32) elements;
// box = {box_cols, box_rows}, 128B swizzle (64B for the bf16 store units), zero OOB fill on loads / clipping on stores.
static int make_store_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int elem_bytes);
int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int box_cols, int box_rows, int elem_bytes,
             bool swizzle64) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return MURCL_ECUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for [%lld x %lld] box {%d,%d}", (int)r, (long long)rows, (long long)cols,
              box_cols, box_rows);
    return MURCL_ECUDA;
  }
  return MURCL_OK;
}

// the epilogue's staging unit: 32 rows x 32 columns
static int make_store_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int elem_bytes) {
  return make_map(map, base, rows, cols, 32, 32, elem_bytes, elem_bytes == 2);
}

static int bstat_mode() {          // MURCL_DISABLE_BSTAT=1: always stream B (A/B comparison)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MURCL_DISABLE_BSTAT");
    v = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  return v;
}

// B-stationary applies when the reduction fits the resident k-blocks and the persistent grid can be made a multiple of
// the column-tile count (so a CTA keeps its column tile).
template <int CG>
static bool bstat_ok(const Params& p) {
  if (!bstat_mode() || p.splits != 1 || p.K > 8 * BLOCK_K) return false;
  const int slots = sm_count() / CG;
  const int64_t total = (int64_t)p.m_tiles * p.n_tiles;
  // worth it only when B is reused: a CTA must see several tiles
  return p.n_tiles <= slots && total >= 2 * (int64_t)slots;
}

template <int BN, bool A_MN, bool B_MN, int EPI, typename TOUT, int CG = 1, bool BSTAT = false>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const Params& p,
                  cudaStream_t st) {
  using C = Cfg<BN, EPI, CG, (int)sizeof(TOUT), BSTAT>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, EPI, TOUT, CG, BSTAT>;
  static PerDeviceOnce configured;
  if (const int slot = configured.pending(); slot >= 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("gemm_tc: cudaFuncSetAttribute(%d B smem) failed: %s", C::SMEM_BYTES, cudaGetErrorString(e));
      return MURCL_ECUDA;
    }
    configured.mark(slot);
  }
  static int debug = -1;
  if (debug < 0) {
    const char* e = getenv("MURCL_DEBUG_EPI");
    debug = e ? atoi(e) : 0;
  }
  Params pp = p;
  pp.debug = debug;
  {
    static int pf_tiles = -1, pf_kb = -1;
    if (pf_tiles < 0) {
      const char* e = getenv("MURCL_GEMM_L2PF");          // B-stationary kernels: tiles ahead (default 1; 0 = off)
      pf_tiles = e ? atoi(e) : 1;
      const char* f = getenv("MURCL_GEMM_L2PF_KB");       // streaming kernels: k-blocks ahead (default 0 = off)
      pf_kb = f ? atoi(f) : 0;
    }
    pp.l2pf = BSTAT ? pf_tiles : pf_kb;
  }
  pp.reverse = (EPI != EPI_SPLIT && row_order_descending()) ? 1 : 0;
  static unsigned long long* trace_buf = nullptr;
  if (debug & 8) {
    if (!trace_buf) cudaMalloc(&trace_buf, sizeof(unsigned long long) * 24 * 4096);
    cudaMemsetAsync(trace_buf, 0, sizeof(unsigned long long) * 24 * 4096, st);
    pp.trace = trace_buf;
  }
  const int64_t total = (int64_t)p.m_tiles * p.n_tiles * p.splits;
  int slots = sm_count() / CG;                                // persistent: one CTA (or CTA pair) per SM (pair)
  if (BSTAT) slots -= slots % p.n_tiles;                      // a CTA keeps its column tile: step % n_tiles == 0
  const int grid = (int)(total < slots ? total : slots) * CG;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ma, mb, mc, pp);
  if (e != cudaSuccess) {
    set_error("gemm_tc_kernel launch failed: %s", cudaGetErrorString(e));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return MURCL_ECUDA;
  }
  if (debug & 8) {
    cudaStreamSynchronize(st);
    static int printed = 0;
    if (printed < 3 && (int64_t)p.m_tiles * p.n_tiles * p.splits > 500) {
      ++printed;
      unsigned long long h[24 * 12];
      cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
      fprintf(stderr, "[trace] BN=%d CG=%d EPI=%d BSTAT=%d tiles/CTA timeline (ns rel. to first):\n", BN, CG, EPI, (int)BSTAT);
      for (int i = 4; i < 10; ++i) {
        fprintf(stderr, "  tile %2d: mma_start %7lld issued %7lld epi_start %7lld | warp ends cta0:", i,
                (long long)(h[24 * i] - h[0]), (long long)(h[24 * i + 1] - h[0]), (long long)(h[24 * i + 2] - h[0]));
        for (int w = 0; w < 8; ++w) fprintf(stderr, " %lld", (long long)(h[24 * i + 4 + w] - h[0]));
        fprintf(stderr, " | cta1:");
        for (int w = 0; w < 8; ++w) fprintf(stderr, " %lld", (long long)(h[24 * i + 12 + w] - h[0]));
        fprintf(stderr, "\n");
      }
      unsigned long long g[8 * 12];
      cudaMemcpy(g, trace_buf + (size_t)24 * 2048, sizeof(g), cudaMemcpyDeviceToHost);
      for (int i = 4; i < 10; ++i)
        fprintf(stderr, "  tile %2d warp0 unit0 phases (ns): tmem-ld+math %lld  store-wait %lld  sts %lld  fence %lld  store-issue %lld\n", i,
                (long long)(g[8 * i + 1] - g[8 * i]), (long long)(g[8 * i + 2] - g[8 * i + 1]), (long long)(g[8 * i + 3] - g[8 * i + 2]),
                (long long)(g[8 * i + 4] - g[8 * i + 3]), (long long)(g[8 * i + 5] - g[8 * i + 4]));
    }
  }
  return check_launch("gemm_tc_kernel");
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace tc

using namespace tc;

// The CTA-pair kernel is used for the instance-level layers (many 256-row tiles, N a multiple of 256).
static bool pair_ok(int64_t M, int N) {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MURCL_DISABLE_CTA_PAIR");
    v = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  return v == 1 && M >= 4096 && N % 256 == 0;
}

static bool tc_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MURCL_DISABLE_TCGEN05");
    v = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

bool tc_fwd_supported(int64_t M, int N, int K, int dtype, int out_dtype) {
  (void)out_dtype;
  return tc_enabled() && dtype == MURCL_BF16 && M >= 128 && N >= 128 && N % 8 == 0 && K >= 64 && K % 8 == 0;
}
bool tc_bwd_input_supported(int64_t M, int N, int K, int dtype) {
  // C = dx [M, K]; reduction over N
  return tc_enabled() && dtype == MURCL_BF16 && M >= 128 && K >= 128 && K % 8 == 0 && N >= 64 && N % 8 == 0;
}
bool tc_bwd_weight_supported(int64_t M, int N, int K, int dtype) {
  // C = dw [N, K]; reduction over M rows
  return tc_enabled() && dtype == MURCL_BF16 && M >= 64 && N >= 128 && N % 8 == 0 && K >= 128 && K % 8 == 0;
}

int tc_linear_fwd(const void* x, const void* w, const float* bias, void* y, int64_t M, int N, int K, int act, int out_dtype,
                  unsigned long long* relu_bits, cudaStream_t st) {
  if (!aligned16(x) || !aligned16(w) || !aligned16(y)) {
    set_error("linear_fwd(tcgen05): operands must be 16-byte aligned");
    return MURCL_EINVAL;
  }
  // batch-sized layers (M = a few 128-row tiles) use 64-wide column tiles so that more SMs take part
  const int BN = (M <= 512 && N % 64 == 0) ? 64 : (N % 256 == 0) ? 256 : 128;
  const bool pair = pair_ok(M, N);
  CUtensorMap ma, mb;
  int rc = make_map(&ma, x, M, K, BLOCK_K, BLOCK_M);
  if (rc != MURCL_OK) return rc;
  rc = make_map(&mb, w, N, K, BLOCK_K, pair ? BN / 2 : BN);      // a CTA of a pair loads half of the B tile
  if (rc != MURCL_OK) return rc;
  CUtensorMap mc;
  const int eb = out_dtype == MURCL_BF16 ? 2 : 4;
  rc = make_store_map(&mc, y, M, N, eb);
  if (rc != MURCL_OK) return rc;
  if (bias != nullptr && !aligned16(bias)) {
    set_error("linear_fwd(tcgen05): bias must be 16-byte aligned");
    return MURCL_EINVAL;
  }
  Params p{};
  p.M = M; p.N = N; p.K = K; p.ldc = N; p.C = y; p.bias = bias; p.act = act; p.bits_out = relu_bits;
  p.splits = 1; p.k_chunk = ((int64_t)K + BLOCK_K - 1) / BLOCK_K * BLOCK_K;
  p.m_tiles = ceil_div(M, BLOCK_M); p.n_tiles = ceil_div(N, BN);
  if (pair) {
    p.m_tiles = ceil_div(M, 2 * BLOCK_M);
    if (out_dtype == MURCL_BF16 && bstat_ok<2>(p)) return launch<256, false, false, EPI_FWD, __nv_bfloat16, 2, true>(ma, mb, mc, p, st);
    return out_dtype == MURCL_BF16 ? launch<256, false, false, EPI_FWD, __nv_bfloat16, 2>(ma, mb, mc, p, st)
                                   : launch<256, false, false, EPI_FWD, float, 2>(ma, mb, mc, p, st);
  }
  if (BN == 128 && out_dtype == MURCL_BF16 && bstat_ok<1>(p))
    return launch<128, false, false, EPI_FWD, __nv_bfloat16, 1, true>(ma, mb, mc, p, st);
  if (BN == 64)
    return out_dtype == MURCL_BF16 ? launch<64, false, false, EPI_FWD, __nv_bfloat16>(ma, mb, mc, p, st)
                                   : launch<64, false, false, EPI_FWD, float>(ma, mb, mc, p, st);
  if (out_dtype == MURCL_BF16)
    return BN == 256 ? launch<256, false, false, EPI_FWD, __nv_bfloat16>(ma, mb, mc, p, st)
                     : launch<128, false, false, EPI_FWD, __nv_bfloat16>(ma, mb, mc, p, st);
  return BN == 256 ? launch<256, false, false, EPI_FWD, float>(ma, mb, mc, p, st)
                   : launch<128, false, false, EPI_FWD, float>(ma, mb, mc, p, st);
}

int tc_linear_bwd_input(const void* dy, const void* w, void* dx, int64_t M, int N, int K, const void* relu_src,
                        const float* row_scale, const float* row_vec, const int32_t* row_seg, float* col_sum,
                        float out_scale, const unsigned long long* relu_bits, cudaStream_t st) {
  if (!aligned16(dy) || !aligned16(w) || !aligned16(dx) || (relu_src && !aligned16(relu_src))) {
    set_error("linear_bwd_input(tcgen05): operands must be 16-byte aligned");
    return MURCL_EINVAL;
  }
  // C = dx [M, K_in]; A = dy [M, N] K-major (reduction over N); B(n'=k_in, k'=n) = w[n, k_in]: MN-major, rows = n.
  const int BN = (M <= 512 && K % 64 == 0) ? 64 : (K % 256 == 0) ? 256 : 128;
  const bool pair = pair_ok(M, K);
  CUtensorMap ma, mb;
  int rc = make_map(&ma, dy, M, N, BLOCK_K, BLOCK_M);
  if (rc != MURCL_OK) return rc;
  rc = make_map(&mb, w, N, K, 64, BLOCK_K);
  if (rc != MURCL_OK) return rc;
  Params p{};
  p.M = M; p.N = K; p.K = N; p.ldc = K; p.C = dx;
  p.relu_src = static_cast<const __nv_bfloat16*>(relu_src);
  p.row_scale = row_scale; p.row_vec = row_vec; p.row_seg = row_seg; p.col_sum = col_sum;
  p.out_scale = out_scale; p.bits_in = relu_bits;
  p.splits = 1; p.k_chunk = ((int64_t)N + BLOCK_K - 1) / BLOCK_K * BLOCK_K;
  p.m_tiles = ceil_div(M, BLOCK_M); p.n_tiles = ceil_div(K, BN);
  CUtensorMap mc;
  rc = make_store_map(&mc, dx, M, K, 2);
  if (rc != MURCL_OK) return rc;
  if (row_vec != nullptr && !aligned16(row_vec)) {
    set_error("linear_bwd_input(tcgen05): row_vec must be 16-byte aligned");
    return MURCL_EINVAL;
  }
  if (pair) {
    p.m_tiles = ceil_div(M, 2 * BLOCK_M);
    if (bstat_ok<2>(p)) return launch<256, false, true, EPI_DGRAD, __nv_bfloat16, 2, true>(ma, mb, mc, p, st);
    return launch<256, false, true, EPI_DGRAD, __nv_bfloat16, 2>(ma, mb, mc, p, st);
  }
  if (BN == 128 && bstat_ok<1>(p)) return launch<128, false, true, EPI_DGRAD, __nv_bfloat16, 1, true>(ma, mb, mc, p, st);
  if (BN == 64) return launch<64, false, true, EPI_DGRAD, __nv_bfloat16>(ma, mb, mc, p, st);
  return BN == 256 ? launch<256, false, true, EPI_DGRAD, __nv_bfloat16>(ma, mb, mc, p, st)
                   : launch<128, false, true, EPI_DGRAD, __nv_bfloat16>(ma, mb, mc, p, st);
}

// The splits add their partial tiles straight into dw with vector atomics (red.global.add.v4.f32: 18-74 adds per address,
// served by the L2 atomic units while the other CTAs still compute) - no workspace round trip and no split-K reduce launch
// (measured: [512 x 512] 134 -> 130 us, the attention projection's [128 x 512] with 74 splits 81 -> 61 us; 0.22 ms per
// step).  The price is a summation order that varies from run to run, as for the bias / pooling gradients already;
// MURCL_WGRAD_ATOMIC=0 restores the workspace + fixed-order reduce.
static bool wgrad_atomic() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MURCL_WGRAD_ATOMIC");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// dx[M, K] (fp32) += dy[M, N] . w[N, K] with the reduction split across the whole GPU and the partial tiles added with vector
// atomics.  For the batch-sized layers of the recurrence (M = 128 rows, reduction 3072): the plain input-gradient launch is 16
// CTAs that each stream 48 k-blocks (18 us); here ~100 CTAs take 2 k-blocks each, and the sum lands directly in the fp32
// buffer that already holds the other gradient term of the same state (no bf16 rounding of the carried gradient).
bool tc_bwd_input_accum_supported(int64_t M, int N, int K, int dtype) {
  return tc_enabled() && dtype == MURCL_BF16 && M >= 64 && K >= 128 && K % 8 == 0 && N >= 64 && N % 8 == 0;
}

int tc_linear_bwd_input_accum(const void* dy, const void* w, float* dx, int64_t M, int N, int K, cudaStream_t st) {
  if (!aligned16(dy) || !aligned16(w) || !aligned16(dx)) {
    set_error("linear_bwd_input_accum(tcgen05): operands must be 16-byte aligned");
    return MURCL_EINVAL;
  }
  const int BN = (K % 256 == 0) ? 256 : 128;
  CUtensorMap ma, mb;
  int rc = make_map(&ma, dy, M, N, BLOCK_K, BLOCK_M);
  if (rc != MURCL_OK) return rc;
  rc = make_map(&mb, w, N, K, 64, BLOCK_K);                       // MN-major B: rows = reduction index n
  if (rc != MURCL_OK) return rc;
  Params p{};
  p.M = M; p.N = K; p.K = N; p.ldc = K; p.C = dx; p.split_stride = 0; p.atomic_out = 1;
  p.m_tiles = ceil_div(M, BLOCK_M); p.n_tiles = ceil_div(K, BN);
  const int tiles = p.m_tiles * p.n_tiles;
  const int64_t kb_total = ((int64_t)N + BLOCK_K - 1) / BLOCK_K;
  int splits = sm_count() / (tiles > 0 ? tiles : 1);
  if (splits < 1) splits = 1;
  if (splits > kb_total) splits = (int)kb_total;
  const int64_t kb_per = (kb_total + splits - 1) / splits;
  p.k_chunk = kb_per * BLOCK_K;
  p.splits = (int)((kb_total + kb_per - 1) / kb_per);
  return BN == 256 ? launch<256, false, true, EPI_SPLIT, float>(ma, mb, ma, p, st)
                   : launch<128, false, true, EPI_SPLIT, float>(ma, mb, ma, p, st);
}

// The weight gradient is operand-load bound with one CTA per 128 x 256 tile (48 KB of operands per 512 tensor-pipe clocks):
// a CTA pair on a 256 x 256 tile pulls 32 KB per CTA for the same MMA work and fits a deeper ring.
static bool wgrad_pair(int64_t M, int N, int K) {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MURCL_WGRAD_PAIR");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1 && pair_ok(M, N) && K % 256 == 0;
}

static void wgrad_plan(int64_t M, int N, int K, int& BN, int& splits, int64_t& k_chunk) {
  BN = (K % 256 == 0) ? 256 : 128;
  const bool pair = wgrad_pair(M, N, K);
  const int tiles = ceil_div(N, pair ? 2 * BLOCK_M : BLOCK_M) * ceil_div(K, BN);
  const int slots = pair ? sm_count() / 2 : sm_count();
  splits = slots / tiles;                                          // one wave
  // (cutting the reduction finer so that every CTA takes two work items - 4 x 37 = 148 = 2 x 74 slots instead of 72 busy pairs -
  //  was measured slower: 10.36 -> 10.65 ms per step)
  if (splits < 1) splits = 1;
  const int64_t kb_total = (M + BLOCK_K - 1) / BLOCK_K;
  if (splits > kb_total) splits = (int)kb_total;
  const int64_t kb_per = (kb_total + splits - 1) / splits;
  k_chunk = kb_per * BLOCK_K;
  splits = (int)((kb_total + kb_per - 1) / kb_per);                // every split owns >= 1 k-block
}

int64_t tc_linear_bwd_weight_workspace(int64_t M, int N, int K) {
  if (M < 64 || N < 128 || K < 128) return 0;
  int BN, splits;
  int64_t kc;
  wgrad_plan(M, N, K, BN, splits, kc);
  return (int64_t)splits * N * K;
}

int tc_linear_bwd_weight(const void* dy, const void* x, float* dw, int64_t M, int N, int K, float* workspace, cudaStream_t st,
                         int accumulate) {
  if (!aligned16(dy) || !aligned16(x) || !aligned16(dw) || !aligned16(workspace)) {
    set_error("linear_bwd_weight(tcgen05): operands must be 16-byte aligned");
    return MURCL_EINVAL;
  }
  int BN, splits;
  int64_t k_chunk;
  wgrad_plan(M, N, K, BN, splits, k_chunk);
  if (workspace == nullptr) {
    set_error("linear_bwd_weight(tcgen05): workspace required");
    return MURCL_EINVAL;
  }
  // C = dw [N, K_in]; reduction over the M rows.  A(m'=n, k'=m) = dy[m, n]; B(n'=k_in, k'=m) = x[m, k_in]; both MN-major.
  CUtensorMap ma, mb;
  int rc = make_map(&ma, dy, M, N, 64, BLOCK_K);
  if (rc != MURCL_OK) return rc;
  rc = make_map(&mb, x, M, K, 64, BLOCK_K);
  if (rc != MURCL_OK) return rc;
  Params p{};
  p.M = N; p.N = K; p.K = M; p.ldc = K; p.C = workspace;
  p.splits = splits; p.k_chunk = k_chunk; p.split_stride = (int64_t)N * K;
  p.m_tiles = ceil_div(N, BLOCK_M); p.n_tiles = ceil_div(K, BN);
  const bool atomic = wgrad_atomic() && splits > 1 && K % 4 == 0;
  if (atomic) {
    if (!accumulate) MURCL_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)N * K, st));
    p.C = dw; p.split_stride = 0; p.atomic_out = 1;
  }
  if (wgrad_pair(M, N, K)) {
    p.m_tiles = ceil_div(N, 2 * BLOCK_M);
    rc = launch<256, true, true, EPI_SPLIT, float, 2>(ma, mb, ma, p, st);
  } else {
    rc = BN == 256 ? launch<256, true, true, EPI_SPLIT, float>(ma, mb, ma, p, st)
                   : launch<128, true, true, EPI_SPLIT, float>(ma, mb, ma, p, st);
  }
  if (rc != MURCL_OK || atomic) return rc;
  const int64_t n = (int64_t)N * K;
  return launch_splitk_reduce(workspace, splits, n, dw, n, st, accumulate);
}

// ---- split-precision (exact-fp32) entry points -----------------------------------------------------------------
// Operands are stacks of bf16 planes [planes][plane_rows][cols] of fp32 tensors (murcl_split_planes).  planes == 2:
// x ~ hi + mid, three products (mid.hi, hi.mid, hi.hi): every product is exact in the fp32 accumulator and the dropped
// terms are below 3 * 2^-18 of |x||y|; planes == 3: six products, all terms down to 2^-24.  Smallest terms first.
static void split_terms(Params& p, int planes, int64_t a_plane_rows, int64_t b_plane_rows) {
  static const int t2a[3] = {1, 0, 0}, t2b[3] = {0, 1, 0};
  static const int t3a[6] = {2, 0, 1, 1, 0, 0}, t3b[6] = {0, 2, 1, 0, 1, 0};
  p.terms = planes == 2 ? 3 : 6;
  for (int i = 0; i < p.terms; ++i) {
    p.pa[i] = planes == 2 ? t2a[i] : t3a[i];
    p.pb[i] = planes == 2 ? t2b[i] : t3b[i];
  }
  p.a_plane_rows = (int)a_plane_rows;
  p.b_plane_rows = (int)b_plane_rows;
}

// CTA pairs for the split-precision forward / input-gradient GEMMs (six products per tile: operand loads are what bounds
// the one-CTA kernel, and a pair halves the B bytes per MMA).  MURCL_SPLIT_PAIR=0 restores the one-CTA kernels.
static bool split_pair() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MURCL_SPLIT_PAIR");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

bool tc_split_supported(int64_t M, int N, int K) {
  return tc_enabled() && M >= 1024 && N >= 128 && N % 64 == 0 && K >= 64 && K % 64 == 0 && 3 * ((M + 63) / 64 * 64) < (1ll << 31);
}

// y[M,N] fp32 = act(x w^T + bias); xp planes [planes][xpr][K], wp planes [planes][wpr][K]
int tc_split_fwd(const void* xp, const void* wp, const float* bias, float* y, int64_t M, int N, int K, int act, int planes,
                 int64_t xpr, int64_t wpr, cudaStream_t st) {
  if (!aligned16(xp) || !aligned16(wp) || !aligned16(y) || (bias && !aligned16(bias))) {
    set_error("linear_fwd_split: operands must be 16-byte aligned");
    return MURCL_EINVAL;
  }
  const int BN = (N % 256 == 0) ? 256 : 128;
  const bool pair = split_pair() && pair_ok(M, N);
  CUtensorMap ma, mb, mc;
  int rc = make_map(&ma, xp, planes * xpr, K, BLOCK_K, BLOCK_M);
  if (rc != MURCL_OK) return rc;
  rc = make_map(&mb, wp, planes * wpr, K, BLOCK_K, pair ? BN / 2 : BN);     // a CTA of a pair loads half of the B tile
  if (rc != MURCL_OK) return rc;
  rc = make_store_map(&mc, y, M, N, 4);
  if (rc != MURCL_OK) return rc;
  Params p{};
  p.M = M; p.N = N; p.K = K; p.ldc = N; p.C = y; p.bias = bias; p.act = act;
  p.splits = 1; p.k_chunk = ((int64_t)K + BLOCK_K - 1) / BLOCK_K * BLOCK_K;
  p.m_tiles = ceil_div(M, BLOCK_M); p.n_tiles = ceil_div(N, BN);
  split_terms(p, planes, xpr, wpr);
  if (pair) {
    p.m_tiles = ceil_div(M, 2 * BLOCK_M);
    return launch<256, false, false, EPI_FWD, float, 2>(ma, mb, mc, p, st);
  }
  return BN == 256 ? launch<256, false, false, EPI_FWD, float>(ma, mb, mc, p, st)
                   : launch<128, false, false, EPI_FWD, float>(ma, mb, mc, p, st);
}

// dx[M,K] fp32 = (dy w) with the fused epilogue terms; dyp planes [planes][dpr][N], wp planes [planes][wpr][K]
int tc_split_bwd_input(const void* dyp, const void* wp, float* dx, int64_t M, int N, int K, const float* row_scale,
                       const float* row_vec, const int32_t* row_seg, float out_scale, const unsigned long long* relu_bits,
                       int planes, int64_t dpr, int64_t wpr, cudaStream_t st) {
  if (!aligned16(dyp) || !aligned16(wp) || !aligned16(dx) || (row_vec && !aligned16(row_vec))) {
    set_error("linear_bwd_input_split: operands must be 16-byte aligned");
    return MURCL_EINVAL;
  }
  const int BN = (K % 256 == 0) ? 256 : 128;
  CUtensorMap ma, mb, mc;
  int rc = make_map(&ma, dyp, planes * dpr, N, BLOCK_K, BLOCK_M);
  if (rc != MURCL_OK) return rc;
  rc = make_map(&mb, wp, planes * wpr, K, 64, BLOCK_K);                  // MN-major: rows = reduction index n (+ plane offset)
  if (rc != MURCL_OK) return rc;
  rc = make_store_map(&mc, dx, M, K, 4);
  if (rc != MURCL_OK) return rc;
  Params p{};
  p.M = M; p.N = K; p.K = N; p.ldc = K; p.C = dx;
  p.row_scale = row_scale; p.row_vec = row_vec; p.row_seg = row_seg;
  p.out_scale = out_scale; p.bits_in = relu_bits;
  p.splits = 1; p.k_chunk = ((int64_t)N + BLOCK_K - 1) / BLOCK_K * BLOCK_K;
  p.m_tiles = ceil_div(M, BLOCK_M); p.n_tiles = ceil_div(K, BN);
  split_terms(p, planes, dpr, wpr);
  if (split_pair() && pair_ok(M, K)) {
    p.m_tiles = ceil_div(M, 2 * BLOCK_M);
    return launch<256, false, true, EPI_DGRAD, float, 2>(ma, mb, mc, p, st);
  }
  return BN == 256 ? launch<256, false, true, EPI_DGRAD, float>(ma, mb, mc, p, st)
                   : launch<128, false, true, EPI_DGRAD, float>(ma, mb, mc, p, st);
}

int64_t tc_split_bwd_weight_workspace(int64_t M, int N, int K) {
  int BN, splits;
  int64_t kc;
  wgrad_plan(M, N, K, BN, splits, kc);
  return (int64_t)splits * N * K;
}

// dw[N,K] (+)= dy^T x; dyp planes [planes][pr][N], xp planes [planes][pr][K]; pr % 64 == 0 with zero padding rows
int tc_split_bwd_weight(const void* dyp, const void* xp, float* dw, int64_t M, int N, int K, int planes, int64_t pr,
                        float* workspace, cudaStream_t st, int accumulate) {
  if (!aligned16(dyp) || !aligned16(xp) || !aligned16(dw) || !aligned16(workspace) || workspace == nullptr) {
    set_error("linear_bwd_weight_split: operands must be 16-byte aligned and a workspace is required");
    return MURCL_EINVAL;
  }
  if (pr % BLOCK_K != 0) {
    set_error("linear_bwd_weight_split: plane pitch %lld must be a multiple of %d rows", (long long)pr, BLOCK_K);
    return MURCL_EINVAL;
  }
  int BN, splits;
  int64_t k_chunk;
  wgrad_plan(M, N, K, BN, splits, k_chunk);
  CUtensorMap ma, mb;
  int rc = make_map(&ma, dyp, planes * pr, N, 64, BLOCK_K);
  if (rc != MURCL_OK) return rc;
  rc = make_map(&mb, xp, planes * pr, K, 64, BLOCK_K);
  if (rc != MURCL_OK) return rc;
  Params p{};
  p.M = N; p.N = K; p.K = M; p.ldc = K; p.C = workspace;
  p.splits = splits; p.k_chunk = k_chunk; p.split_stride = (int64_t)N * K;
  p.m_tiles = ceil_div(N, BLOCK_M); p.n_tiles = ceil_div(K, BN);
  split_terms(p, planes, pr, pr);
  const bool atomic = wgrad_atomic() && splits > 1 && K % 4 == 0;
  if (atomic) {
    if (!accumulate) MURCL_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)N * K, st));
    p.C = dw; p.split_stride = 0; p.atomic_out = 1;
  }
  if (wgrad_pair(M, N, K)) {
    p.m_tiles = ceil_div(N, 2 * BLOCK_M);
    rc = launch<256, true, true, EPI_SPLIT, float, 2>(ma, mb, ma, p, st);
  } else {
    rc = BN == 256 ? launch<256, true, true, EPI_SPLIT, float>(ma, mb, ma, p, st)
                   : launch<128, true, true, EPI_SPLIT, float>(ma, mb, ma, p, st);
  }
  if (rc != MURCL_OK || atomic) return rc;
  const int64_t n = (int64_t)N * K;
  return launch_splitk_reduce(workspace, splits, n, dw, n, st, accumulate);
}

}  // namespace murcl
