// tcgen05 / TMEM / TMA GEMM family for the bf16 mode of murcl_linear_{fwd,bwd_input,bwd_weight}.
//
//   C[m,n] = sum_k A(m,k) * B(n,k)      bf16 operands, fp32 accumulators in tensor memory
//
// One persistent CTA per SM, 6 warps:
//   warp 0  TMA producer   cp.async.bulk.tensor.2d -> 128B-swizzled smem ring (STAGES deep), mbarrier tx-count
//   warp 1  MMA issuer     one lane issues tcgen05.mma.cta_group::1.kind::f16 (128 x BN x 16), tcgen05.commit frees
//                          smem slots and publishes the accumulator; also owns tcgen05.alloc/dealloc
//   warps 2-5 epilogue     tcgen05.ld 32 lanes x 32 columns -> registers -> fused epilogue -> 128B-swizzled smem slab
//                          (32 rows x 128 B per warp) -> cp.async.bulk.tensor store; the ReLU-mask operand of the
//                          input-gradient epilogue arrives the same way (TMA load of the matching slab)
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

namespace murcl {

// defined in gemm_simt.cu
int launch_splitk_reduce(const float* ws, int splits, int64_t stride, float* out, int64_t n, cudaStream_t st);

namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;               // 64 bf16 = 128 B = one swizzle span
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;           // two per TMEM lane quarter, splitting the columns
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
};

// ---- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c_inner), "r"(c_outer)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c_inner, int c_outer) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c_inner), "r"(c_outer)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Activation over a register slab.  bf16 outputs use the MUFU tanh (error ~2^-11, below bf16's 2^-9 rounding);
// fp32 outputs use the accurate libm versions.
template <bool FAST, int N>
__device__ __forceinline__ void act_slab(float (&v)[N], int act, int col0, int n_cols) {
  if (act == MURCL_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = fmaxf(v[i], 0.f);
  } else if (act == MURCL_ACT_TANH || (act == MURCL_ACT_TANH_SIGMOID && col0 + N <= (n_cols >> 1))) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = FAST ? tanh_fast(v[i]) : tanhf(v[i]);
  } else if (act == MURCL_ACT_SIGMOID || (act == MURCL_ACT_TANH_SIGMOID && col0 >= (n_cols >> 1))) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = FAST ? fmaf(0.5f, tanh_fast(0.5f * v[i]), 0.5f) : 1.f / (1.f + expf(-v[i]));
  } else if (act == MURCL_ACT_TANH_SIGMOID) {      // slab straddles the tanh | sigmoid boundary
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = (col0 + i < (n_cols >> 1)) ? tanhf(v[i]) : 1.f / (1.f + expf(-v[i]));
  }
}

// ---- cluster / cta_group::2 helpers ------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {      // same smem offset in CTA `rank` of the cluster
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Relaxed: the arrive only hands the (already drained, tcgen05.wait::ld + fence) accumulator back to the MMA issuer;
// a release at cluster scope would also wait for this warp's outstanding global stores (~1.7 us per tile, measured).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of a CTA pair: data lands in the issuing CTA's smem, the transaction bytes are counted on the barrier
// at `bar_cluster_addr` (the leader CTA's barrier).
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c_inner,
                                                int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c_inner), "r"(c_outer)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit_2sm(uint32_t bar) {             // arrive on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tcgen05_mma_bf16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (sm_100 layout: address>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout type [61,64) with 2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor for kind::f16: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), majors (bits 15, 16),
// N>>3 (bits 17-22), M>>4 (bits 24-28).
__host__ __device__ constexpr uint32_t make_idesc(int umma_m, int umma_n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(umma_n >> 3) << 17) | ((uint32_t)(umma_m >> 4) << 24);
}

constexpr int SLAB_BYTES = 32 * 128;        // one epilogue warp's staging slab: 32 rows x 128 B

// CG = cta_group: 1 = one SM per 128 x BN tile; 2 = a CTA pair shares a 256 x BN tile (each CTA holds its 128 rows of A
// and HALF of B, the MMA unit exchanges the B halves), which halves the B bytes each SM pulls through L2 and reads from
// shared memory per MMA - the limiter of the 1-CTA kernel.
template <int BN, int EPI, int CG>
struct Cfg {
  static constexpr int B_BYTES = (BN / CG) * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SLABS_PER_WARP = (EPI == 2) ? 0 : 2;   // (the split-K kernel stores directly) two output slabs (double-buffered stores); the input grad's TMA-loaded
                                             // ReLU-mask slab takes the place of the second one when it is in use
  static constexpr int STAGING_BYTES = NUM_EPI_WARPS * SLABS_PER_WARP * SLAB_BYTES;
  static constexpr int STAGES = (SMEM_BUDGET - STAGING_BYTES) / STAGE_BYTES > 8 ? 8 : (SMEM_BUDGET - STAGING_BYTES) / STAGE_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;                  // 256 or 512: powers of two
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ +
                                    BN * 4 /*column-sum accumulator*/;
};

#define MURCL_STAMP(slot)                                                                         \
  if (p.trace && blockIdx.x == 0 && threadIdx.x == 0 && c0 == 0) {                              \
    unsigned long long ts_;                                                                     \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts_));                                     \
    p.trace[(int64_t)24 * 2048 + ((t - tile_first) / tile_step) * 8 + (slot)] = ts_;            \
  }

template <int BN, bool A_MN, bool B_MN, int EPI, typename TOUT, int CG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_mask, const Params p) {
  using C = Cfg<BN, EPI, CG>;
  static_assert(CG == 1 || !A_MN, "the CTA-pair kernel takes a K-major A operand");
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0u;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;                       // SWIZZLE_128B atoms need 1024 B alignment
  const uint32_t staging = tiles + C::STAGES * C::STAGE_BYTES;        // 1024 B aligned (stage sizes are multiples of 1024)
  const uint32_t bars = staging + C::STAGING_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + 2 + s); };
  auto mask_bar = [&](int w) { return bars + 8u * (2 * C::STAGES + 4 + w); };
  const uint32_t tmem_slot = bars + 8u * (2 * C::STAGES + 4 + NUM_EPI_WARPS);
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
    for (int w = 0; w < NUM_EPI_WARPS; ++w) mbar_init(mask_bar(w), 1);
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
      for (int t = tile_first; t < total_tiles; t += tile_step) {
        int mb, nb, sp, nkb;
        int64_t k0;
        tile_coords(t, mb, nb, sp);
        k_range(sp, k0, nkb);
        const int m0 = (mb * CG + (int)cta_rank) * BLOCK_M;                   // this CTA's 128 rows of the tile
        const int n0 = nb * BN + (int)cta_rank * (BN / CG);                    // this CTA's share of the B rows
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t a_dst = tiles + s * C::STAGE_BYTES;
          const uint32_t b_dst = a_dst + A_BYTES;
          const int kk = (int)(k0 + (int64_t)kb * BLOCK_K);
          if (CG == 2) {
            // both CTAs load their halves; all bytes are counted on the LEADER's full barrier (it issues the MMAs)
            if (leader) mbar_expect_tx(full_bar(s), 2 * C::STAGE_BYTES);
            const uint32_t fb = mapa(full_bar(s), 0);
            tma_load_2d_2sm(a_dst, &map_a, fb, kk, m0);
            if (!B_MN) {
              tma_load_2d_2sm(b_dst, &map_b, fb, kk, n0);
            } else {
#pragma unroll
              for (int j = 0; j < BN / CG / 64; ++j) tma_load_2d_2sm(b_dst + j * 8192, &map_b, fb, n0 + 64 * j, kk);
            }
            if (++s == C::STAGES) { s = 0; ph ^= 1u; }
            continue;
          }
          mbar_expect_tx(full_bar(s), C::STAGE_BYTES);
          if (!A_MN) {
            tma_load_2d(a_dst, &map_a, full_bar(s), kk, m0);                    // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)                               // box {64 m, 64 k-rows}
              tma_load_2d(a_dst + j * 8192, &map_a, full_bar(s), m0 + 64 * j, kk);
          }
          if (!B_MN) {
            tma_load_2d(b_dst, &map_b, full_bar(s), kk, n0);                    // box {64 k, BN rows}
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(b_dst + j * 8192, &map_b, full_bar(s), n0 + 64 * j, kk);
          }
          if (++s == C::STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ================= MMA issuer (the leader CTA of a pair issues for both) =================
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc(BLOCK_M * CG, BN, A_MN, B_MN);
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      for (int t = tile_first; t < total_tiles; t += tile_step) {
        int mb, nb, sp, nkb;
        int64_t k0;
        tile_coords(t, mb, nb, sp);
        k_range(sp, k0, nkb);
        mbar_wait(tempty_bar(as), aph ^ 1u);
        tcgen05_fence_after();
        if (p.trace && blockIdx.x == 0) {
          unsigned long long ts;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts));
          p.trace[((t - tile_first) / tile_step) * 24 + 0] = ts;
        }
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full_bar(s), ph);
          tcgen05_fence_after();
          const uint32_t a_src = tiles + s * C::STAGE_BYTES;
          const uint32_t b_src = a_src + A_BYTES;
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
        if (p.trace && blockIdx.x == 0) {
          unsigned long long ts;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts));
          p.trace[((t - tile_first) / tile_step) * 24 + 1] = ts;      // all MMAs of the tile ISSUED
        }
        if (++as == 2) { as = 0; aph ^= 1u; }
      }
    }
  } else {
    // ================= epilogue warps (0..7) =================
    const int quarter = warp & 3;                          // TMEM lanes [32*quarter, +32) belong to this warp
    const int ew = warp;                                   // 0..7
    const int half = ew >> 2;                              // the two warps of a quarter take alternate column slabs
    const uint32_t slab_base = staging + (uint32_t)ew * (C::SLABS_PER_WARP * SLAB_BYTES);
    const uint32_t mask_slab = slab_base + SLAB_BYTES;     // second slab doubles as the ReLU-mask slab (mask mode)
    uint32_t out_slab = slab_base;
    uint32_t slab_flip = 0;
    const uint32_t my_row_off = (uint32_t)lane * 128u;
    const uint32_t swz = (uint32_t)(lane & 7);             // 128B swizzle: 16-byte chunk index ^= row & 7
    constexpr int SLAB_COLS = 128 / (int)sizeof(TOUT);      // 64 bf16 or 32 fp32 columns = 128 B per row
    int as = 0;
    uint32_t aph = 0, mph = 0;
    const bool has_mask = (EPI == EPI_DGRAD) && p.relu_src != nullptr && p.bits_in == nullptr;
    const bool want_colsum = (EPI == EPI_DGRAD) && sizeof(TOUT) == 2 && p.col_sum != nullptr;
    int acc_nb = -1;
    const int etid = threadIdx.x;                           // 0..255: the epilogue threads come first
    auto flush_colacc = [&](int nb_flush) {
      asm volatile("bar.sync 1, 256;" ::: "memory");        // all epilogue warps have added their slabs
      for (int i = etid; i < BN; i += 32 * NUM_EPI_WARPS) {
        const int col = nb_flush * BN + i;
        const float sum = colacc[i];
        if (col < p.N && sum != 0.f) atomicAdd(p.col_sum + col, sum);
        colacc[i] = 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    };
    if (want_colsum) {
      for (int i = etid; i < BN; i += 32 * NUM_EPI_WARPS) colacc[i] = 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    if (has_mask && lane == 0) {                            // mask slab of the first work item of this warp
      for (int t = tile_first; t < total_tiles; t += tile_step) {
        int mb, nb, sp;
        tile_coords(t, mb, nb, sp);
        if (((int64_t)mb * CG + cta_rank) * BLOCK_M + quarter * 32 < p.M && nb * BN + half * SLAB_COLS < p.N) {
          mbar_expect_tx(mask_bar(ew), SLAB_BYTES);
          tma_load_2d(mask_slab, &map_mask, mask_bar(ew), nb * BN + half * SLAB_COLS,
                      (mb * CG + (int)cta_rank) * BLOCK_M + quarter * 32);
          break;
        }
      }
    }
    for (int t = tile_first; t < total_tiles; t += tile_step) {
      int mb, nb, sp;
      tile_coords(t, mb, nb, sp);
      const int64_t row0 = ((int64_t)mb * CG + cta_rank) * BLOCK_M + quarter * 32;
      const int64_t row = row0 + lane;
      const int n0 = nb * BN;
      mbar_wait(tfull_bar(as), aph);
      tcgen05_fence_after();
      if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long ts;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts));
        p.trace[((t - tile_first) / tile_step) * 24 + 2] = ts;        // accumulator complete, epilogue starts
      }
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN);
      if (EPI == EPI_SPLIT) {
        // fp32 partial sums straight to the split workspace (few tiles per launch: not worth staging)
#pragma unroll 1
        for (int c0 = half * 32; c0 < BN; c0 += 64) {
          uint32_t r[32];
          tmem_ld32(t_row + (uint32_t)c0, r);
          tmem_ld_wait();
          const int col0 = n0 + c0;
          if (row < p.M && col0 < p.N) {
            float* dst = static_cast<float*>(p.C) + (int64_t)sp * p.split_stride + row * p.ldc + col0;
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              if (col0 + i < p.N)
                *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                                  __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
          }
        }
      } else {
        float rs = 0.f;
        const float* rv = nullptr;
        if (EPI == EPI_DGRAD && p.row_scale != nullptr && row < p.M) {
          rs = p.row_scale[row];
          rv = p.row_vec + (int64_t)p.row_seg[row] * p.N;
        }
        if (want_colsum && nb != acc_nb) {                 // new column tile: flush the CTA's partial sums
          if (acc_nb >= 0) flush_colacc(acc_nb);
          acc_nb = nb;
        }
#pragma unroll 1
        for (int c0 = half * SLAB_COLS; c0 < BN; c0 += 2 * SLAB_COLS) {
          const int col0 = n0 + c0;
          if (col0 >= p.N || row0 >= p.M) break;            // warp-uniform: nothing of this slab is in range
          MURCL_STAMP(0)
          float v[SLAB_COLS];
          if (p.debug & 4) {
#pragma unroll
            for (int i = 0; i < SLAB_COLS; ++i) v[i] = 1.f;
          } else {
            uint32_t r[SLAB_COLS / 32][32];
#pragma unroll
            for (int j = 0; j < SLAB_COLS / 32; ++j) tmem_ld32(t_row + (uint32_t)(c0 + 32 * j), r[j]);   // both in flight
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < SLAB_COLS / 32; ++j)
#pragma unroll
              for (int i = 0; i < 32; ++i) v[32 * j + i] = __uint_as_float(r[j][i]);
          }
          MURCL_STAMP(1)
          if (EPI == EPI_FWD && !(p.debug & 2)) {
            if (p.bias != nullptr) {
              if (col0 + SLAB_COLS <= p.N) {                // whole slab in range: branch-free, loads issued back to back
                float4 b4[SLAB_COLS / 4];
#pragma unroll
                for (int i = 0; i < SLAB_COLS / 4; ++i) b4[i] = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + i);
#pragma unroll
                for (int i = 0; i < SLAB_COLS / 4; ++i) {
                  v[4 * i] += b4[i].x; v[4 * i + 1] += b4[i].y; v[4 * i + 2] += b4[i].z; v[4 * i + 3] += b4[i].w;
                }
              } else {
#pragma unroll
                for (int i = 0; i < SLAB_COLS; i += 4) {
                  if (col0 + i < p.N) {                     // N % 8 == 0
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
                    v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
                  }
                }
              }
            }
            act_slab<sizeof(TOUT) == 2, SLAB_COLS>(v, p.act, col0, p.N);
            if (sizeof(TOUT) == 2 && p.bits_out != nullptr && row < p.M) {
              // one 64-bit word per (row, 64-column slab); consecutive rows are consecutive words -> coalesced
              // after the ReLU v >= +0, so (v > 0) is the sign bit of the integer negation of its bit pattern;
              // a funnel shift appends one sign bit per instruction (2 ops per element)
              // four independent 16-bit chains (a single 32-long dependent chain is latency bound)
              unsigned int q0 = 0u, q1 = 0u, q2 = 0u, q3 = 0u;
#pragma unroll
              for (int i = 15; i >= 0; --i) {
                q0 = __funnelshift_l(0u - __float_as_uint(v[i]), q0, 1);
                q1 = __funnelshift_l(0u - __float_as_uint(v[16 + i]), q1, 1);
                q2 = __funnelshift_l(0u - __float_as_uint(v[(32 + i) % SLAB_COLS]), q2, 1);
                q3 = __funnelshift_l(0u - __float_as_uint(v[(48 + i) % SLAB_COLS]), q3, 1);
              }
              const unsigned int lo = q0 | (q1 << 16), hi = q2 | (q3 << 16);
              p.bits_out[(int64_t)(col0 >> 6) * p.M + row] = ((unsigned long long)hi << 32) | lo;
            }
          } else if (EPI == EPI_DGRAD) {
            if (rv != nullptr) {
              if (col0 + SLAB_COLS <= p.N) {
                float4 g4[SLAB_COLS / 4];
#pragma unroll
                for (int i = 0; i < SLAB_COLS / 4; ++i) g4[i] = __ldg(reinterpret_cast<const float4*>(rv + col0) + i);
#pragma unroll
                for (int i = 0; i < SLAB_COLS / 4; ++i) {
                  v[4 * i] = fmaf(rs, g4[i].x, v[4 * i]); v[4 * i + 1] = fmaf(rs, g4[i].y, v[4 * i + 1]);
                  v[4 * i + 2] = fmaf(rs, g4[i].z, v[4 * i + 2]); v[4 * i + 3] = fmaf(rs, g4[i].w, v[4 * i + 3]);
                }
              } else {
#pragma unroll
                for (int i = 0; i < SLAB_COLS; i += 4) {
                  if (col0 + i < p.N) {
                    const float4 g4 = __ldg(reinterpret_cast<const float4*>(rv + col0 + i));
                    v[i] = fmaf(rs, g4.x, v[i]); v[i + 1] = fmaf(rs, g4.y, v[i + 1]);
                    v[i + 2] = fmaf(rs, g4.z, v[i + 2]); v[i + 3] = fmaf(rs, g4.w, v[i + 3]);
                  }
                }
              }
            }
            if (p.bits_in != nullptr) {
              const unsigned long long bits = row < p.M ? __ldg(p.bits_in + (int64_t)(col0 >> 6) * p.M + row) : 0ull;
              const unsigned int lo = (unsigned int)bits, hi = (unsigned int)(bits >> 32);
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                if (!((lo >> i) & 1u)) v[i] = 0.f;
                if (!((hi >> i) & 1u)) v[(32 + i) % SLAB_COLS] = 0.f;
              }
              if (p.out_scale != 1.f) {
#pragma unroll
                for (int i = 0; i < SLAB_COLS; ++i) v[i] *= p.out_scale;
              }
            }
            if (has_mask) {
              mbar_wait(mask_bar(ew), mph);                 // slab fetched ahead (issued one slab earlier)
              mph ^= 1u;
#pragma unroll
              for (int c = 0; c < 8; ++c) {                 // bf16 mask slab: 8 chunks of 8 columns per row
                const uint4 q = ld_shared_v4(mask_slab + my_row_off + (((uint32_t)c ^ swz) << 4));
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = __bfloat1622float2(h[j]);
                  if (!(f.x > 0.f)) v[8 * c + 2 * j] = 0.f;
                  if (!(f.y > 0.f)) v[8 * c + 2 * j + 1] = 0.f;
                }
              }
              if (p.out_scale != 1.f) {
#pragma unroll
                for (int i = 0; i < SLAB_COLS; ++i) v[i] *= p.out_scale;
              }
              __syncwarp();                                 // every lane is done with the mask slab
              if (lane == 0) {                              // prefetch the next slab this warp will process
                int nt = t;
                int nc0 = c0 + 2 * SLAB_COLS;
                bool found = (nc0 < BN) && (n0 + nc0 < p.N);
                while (!found) {
                  nt += tile_step;
                  if (nt >= total_tiles) break;
                  nc0 = half * SLAB_COLS;
                  int mb2, nb2, sp2;
                  tile_coords(nt, mb2, nb2, sp2);
                  found = (((int64_t)mb2 * CG + cta_rank) * BLOCK_M + quarter * 32 < p.M) && (nb2 * BN + nc0 < p.N);
                }
                if (found) {
                  int mb2, nb2, sp2;
                  tile_coords(nt, mb2, nb2, sp2);
                  mbar_expect_tx(mask_bar(ew), SLAB_BYTES);
                  tma_load_2d(mask_slab, &map_mask, mask_bar(ew), nb2 * BN + nc0,
                              (mb2 * CG + (int)cta_rank) * BLOCK_M + quarter * 32);
                }
              }
            }
          }
          MURCL_STAMP(2)
          if (p.debug & 1) continue;                        // timing experiment: no staging, no store
          // stage the slab (swizzled like the TMA box) and hand it to the bulk-store engine
          if (has_mask) {
            if (lane == 0) bulk_wait_read0();               // single output slab: the previous store has read it
          } else {
            out_slab = slab_base + slab_flip * SLAB_BYTES;  // two output slabs: only the store before last must be done
            slab_flip ^= 1u;
            if (lane == 0) bulk_wait_read1();
          }
          __syncwarp();
          if (sizeof(TOUT) == 2) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
              st_shared_v4(out_slab + my_row_off + (((uint32_t)c ^ swz) << 4), pack_bf16(v[8 * c], v[8 * c + 1]),
                           pack_bf16(v[8 * c + 2], v[8 * c + 3]), pack_bf16(v[8 * c + 4], v[8 * c + 5]),
                           pack_bf16(v[8 * c + 6], v[8 * c + 7]));
          } else {
#pragma unroll
            for (int c = 0; c < 8; ++c)
              st_shared_v4(out_slab + my_row_off + (((uint32_t)c ^ swz) << 4), __float_as_uint(v[4 * c % SLAB_COLS]),
                           __float_as_uint(v[(4 * c + 1) % SLAB_COLS]), __float_as_uint(v[(4 * c + 2) % SLAB_COLS]),
                           __float_as_uint(v[(4 * c + 3) % SLAB_COLS]));
          }
          MURCL_STAMP(3)
          fence_async_smem();
          __syncwarp();
          MURCL_STAMP(4)
          if (lane == 0) {
            tma_store_2d(&map_c, out_slab, col0, (int)row0);   // rows >= M and cols >= N are clipped by the tensor map
            bulk_commit();
          }
          MURCL_STAMP(5)
          if (want_colsum) {
            // column sums of the slab as stored (bf16-rounded): lane owns the 4-byte word `lane` of every row
            float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
              uint32_t wv;
              const uint32_t addr = out_slab + (uint32_t)r * 128u + (((((uint32_t)lane >> 2) ^ ((uint32_t)r & 7u)) << 4) | (((uint32_t)lane & 3u) << 2));
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wv) : "r"(addr) : "memory");
              const float2 f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&wv));
              s0 += f.x;
              s1 += f.y;
            }
            atomicAdd(&colacc[c0 + 2 * lane], s0);
            atomicAdd(&colacc[c0 + 2 * lane + 1], s1);
          }
        }
      }
      if (p.trace && blockIdx.x < 2 && lane == 0) {
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
    if (want_colsum && acc_nb >= 0) flush_colacc(acc_nb);
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
// box = {box_cols, box_rows}, 128B swizzle, zero OOB fill on loads / clipping on stores.
static int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int box_cols, int box_rows,
                    int elem_bytes = 2) {
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
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for [%lld x %lld] box {%d,%d}", (int)r, (long long)rows, (long long)cols,
              box_cols, box_rows);
    return MURCL_ECUDA;
  }
  return MURCL_OK;
}

template <int BN, bool A_MN, bool B_MN, int EPI, typename TOUT, int CG = 1>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mm, const Params& p,
                  cudaStream_t st) {
  using C = Cfg<BN, EPI, CG>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, EPI, TOUT, CG>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("gemm_tc: cudaFuncSetAttribute(%d B smem) failed: %s", C::SMEM_BYTES, cudaGetErrorString(e));
      return MURCL_ECUDA;
    }
    configured = true;
  }
  static int debug = -1;
  if (debug < 0) {
    const char* e = getenv("MURCL_DEBUG_EPI");
    debug = e ? atoi(e) : 0;
  }
  Params pp = p;
  pp.debug = debug;
  static unsigned long long* trace_buf = nullptr;
  if (debug & 8) {
    if (!trace_buf) cudaMalloc(&trace_buf, sizeof(unsigned long long) * 24 * 4096);
    cudaMemsetAsync(trace_buf, 0, sizeof(unsigned long long) * 24 * 4096, st);
    pp.trace = trace_buf;
  }
  const int64_t total = (int64_t)p.m_tiles * p.n_tiles * p.splits;
  const int slots = sm_count() / CG;                          // persistent: one CTA (or CTA pair) per SM (pair)
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ma, mb, mc, mm, pp);
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
      fprintf(stderr, "[trace] BN=%d CG=%d EPI=%d tiles/CTA timeline (ns rel. to first):\n", BN, CG, EPI);
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
        fprintf(stderr, "  tile %2d warp0 slab0 phases (ns): tmem %lld  math %lld  wait+sts %lld  fence %lld  store-issue %lld\n", i,
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
  rc = make_map(&mc, y, M, N, 128 / eb, 32, eb);                 // store slab: 32 rows x 128 B
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
    return out_dtype == MURCL_BF16 ? launch<256, false, false, EPI_FWD, __nv_bfloat16, 2>(ma, mb, mc, mc, p, st)
                                   : launch<256, false, false, EPI_FWD, float, 2>(ma, mb, mc, mc, p, st);
  }
  if (BN == 64)
    return out_dtype == MURCL_BF16 ? launch<64, false, false, EPI_FWD, __nv_bfloat16>(ma, mb, mc, mc, p, st)
                                   : launch<64, false, false, EPI_FWD, float>(ma, mb, mc, mc, p, st);
  if (out_dtype == MURCL_BF16)
    return BN == 256 ? launch<256, false, false, EPI_FWD, __nv_bfloat16>(ma, mb, mc, mc, p, st)
                     : launch<128, false, false, EPI_FWD, __nv_bfloat16>(ma, mb, mc, mc, p, st);
  return BN == 256 ? launch<256, false, false, EPI_FWD, float>(ma, mb, mc, mc, p, st)
                   : launch<128, false, false, EPI_FWD, float>(ma, mb, mc, mc, p, st);
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
  CUtensorMap mc, mm;
  rc = make_map(&mc, dx, M, K, 64, 32);
  if (rc != MURCL_OK) return rc;
  rc = make_map(&mm, relu_src ? relu_src : dx, M, K, 64, 32);
  if (rc != MURCL_OK) return rc;
  if (row_vec != nullptr && !aligned16(row_vec)) {
    set_error("linear_bwd_input(tcgen05): row_vec must be 16-byte aligned");
    return MURCL_EINVAL;
  }
  if (pair) {
    p.m_tiles = ceil_div(M, 2 * BLOCK_M);
    return launch<256, false, true, EPI_DGRAD, __nv_bfloat16, 2>(ma, mb, mc, mm, p, st);
  }
  if (BN == 64) return launch<64, false, true, EPI_DGRAD, __nv_bfloat16>(ma, mb, mc, mm, p, st);
  return BN == 256 ? launch<256, false, true, EPI_DGRAD, __nv_bfloat16>(ma, mb, mc, mm, p, st)
                   : launch<128, false, true, EPI_DGRAD, __nv_bfloat16>(ma, mb, mc, mm, p, st);
}

static void wgrad_plan(int64_t M, int N, int K, int& BN, int& splits, int64_t& k_chunk) {
  BN = (K % 256 == 0) ? 256 : 128;
  const int tiles = ceil_div(N, BLOCK_M) * ceil_div(K, BN);
  splits = sm_count() / tiles;                                     // one wave
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

int tc_linear_bwd_weight(const void* dy, const void* x, float* dw, int64_t M, int N, int K, float* workspace, cudaStream_t st) {
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
  rc = BN == 256 ? launch<256, true, true, EPI_SPLIT, float>(ma, mb, ma, ma, p, st)
                 : launch<128, true, true, EPI_SPLIT, float>(ma, mb, ma, ma, p, st);
  if (rc != MURCL_OK) return rc;
  const int64_t n = (int64_t)N * K;
  return launch_splitk_reduce(workspace, splits, n, dw, n, st);
}

}  // namespace murcl
