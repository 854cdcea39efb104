// Fused attention pooling, backward (the matching pass of attnpool.cu; math in SURVEY.md 7.3; abmil.py:36-45,
// clam.py:37-60,170 differentiated):
//
//   given dM [B, L] (gradient of the pooled vectors), the saved activations uv [n_rows, NC], the normalised weights
//   p [n_rows] (incl. the post scale alpha_b) and the pooled vectors M [B, L]:
//
//     t_n  = dM[b] . h_n                                  K_b = (dM[b] . M[b]) / alpha_b
//     ds_n = p_n (t_n - K_b)                              gradient of the raw score
//     du_n = ds_n wc (.) (1 - u^2)            [(.) v]     written over u   (gated: dv_n = ds_n wc (.) u (.) v (1 - v) over v)
//     dwc  = sum_n ds_n g_n,   dbc = sum_n ds_n,   dpre_colsum = column sums of what is written (bias gradient of the
//     attention projection)
//
// ONE pass: h is read once, uv is read and rewritten in place, ds never touches memory (optional output for tests).  The
// two kernels this replaces (pool_bwd_scores + attn_score_bwd) made a pass each and exchanged ds through HBM, and a
// third tiny launch (pool_k) produced K_b.  The remaining term of the pooling backward, dh_n += p_n dM[b], is applied
// where dh is produced: in the epilogue of the projection's input-gradient GEMM (gemm_tc.cu / gemm_simt.cu).
//
// Roofline: HBM.  Algorithmic bytes per row: L*s (h) + 2*NC*s (uv in, uv out) + 4 (p) [+ 4 (ds)], s = 2 (bf16) or 4.
// FLOPs per row: 2L + ~10 NC: three orders of magnitude under the tensor ridge - plain FMA, no MMA.
//
// Structure: a persistent CTA per SM.  The rows of a tile (rows_per_stage consecutive rows) are CONTIGUOUS in h and uv,
// so one producer lane moves a tile (and its slices of p and row_seg) with 1-D bulk copies (cp.async.bulk, mbarrier tx-count) into a shared-memory
// ring of up to 4 stages (~40 KB each): the bytes in flight do not depend on registers or occupancy.  16 consumer warps
// take the rows of a stage round-robin; lane l holds columns [8l + 256j, +8) of h (one 16-byte shared load per j for
// bf16) and columns [4l + 128i, +4) of u (and v); the dot product is a warp reduction; the gradient goes straight to
// global memory (a warp writes whole rows, contiguous).  dM[b] sits in registers and is re-fetched when the warp crosses
// into the next bag (bags are contiguous row ranges), together with K_b.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace murcl {
namespace apb {

using namespace tc;

constexpr int CONSUMER_WARPS = 16;
constexpr int NUM_THREADS = 32 * (CONSUMER_WARPS + 1);
constexpr int PRODUCER_WARP = CONSUMER_WARPS;
constexpr int MAX_STAGES = 4;
constexpr int MAX_RPW = 2;                                   // rows of one stage per consumer warp
constexpr int STAGE_TARGET_BYTES = 40 * 1024;
constexpr int SMEM_BUDGET = 200 * 1024;

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
// 8 consecutive elements of a staged row as fp32
__device__ __forceinline__ void lds8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(hh[j]);
    v[2 * j] = f.x;
    v[2 * j + 1] = f.y;
  }
}
__device__ __forceinline__ void lds8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

struct Params {
  int64_t n_rows;
  int L, D, inv_sqrt_n;
  int rows_per_stage, stages, n_tiles;
  int h_stage_bytes, ps_off, stage_bytes;  // per-stage regions (multiples of 128 bytes): [h tile | uv tile | p | row_seg]
  float q;                         // 1/(1-p_drop) when a dropout followed the activations, else 1
  const float* p;
  const float* M;
  const float* dM;
  const float* wc;
  const int64_t* offsets;
  const int32_t* row_seg;
  float* ds;                       // optional
  float* dwc;
  float* dbc;
  float* dpre_colsum;              // optional
};

// NJ = ceil(L / 256) chunks of h per lane, NI = ceil(D / 128) chunks of u (and v) per lane.
template <typename T, bool GATED, int NJ, int NI>
__global__ void __launch_bounds__(NUM_THREADS, 1) attnpool_bwd_kernel(const T* __restrict__ h, T* __restrict__ uv, const Params prm) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  // layout: [stages x (h tile | uv tile)] | partial sums [D + NC] floats | barriers
  const uint32_t base = smem_u32(smem_raw);
  const int L = prm.L, D = prm.D;
  const int ld = GATED ? 2 * D : D;
  const int S = prm.stages;
  const int sums_bytes = ((D + ld) * 4 + 15) / 16 * 16;
  float* sums = reinterpret_cast<float*>(smem_raw + (size_t)S * prm.stage_bytes);
  const uint32_t bars = base + (uint32_t)(S * prm.stage_bytes) + (uint32_t)sums_bytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (MAX_STAGES + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // a CTA takes a CONTIGUOUS range of tiles: its warps cross a bag boundary (re-fetch of dM[b], M[b]: a dependent L2
  // round trip) only where the rows really change bag, not at every tile as a round-robin assignment would
  const int small_bytes = (prm.stage_bytes - prm.ps_off) >> 1;
  // 16-byte alignment of the p / row_seg slices of every full tile (rows_per_stage % 4 == 0 and aligned arrays)
  const bool small_ok = (prm.rows_per_stage & 3) == 0 && ((reinterpret_cast<uintptr_t>(prm.p) | reinterpret_cast<uintptr_t>(prm.row_seg)) & 15) == 0;
  const int t_begin = (int)((int64_t)prm.n_tiles * blockIdx.x / gridDim.x);
  const int t_end = (int)((int64_t)prm.n_tiles * (blockIdx.x + 1) / gridDim.x);
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int d = threadIdx.x; d < D + ld; d += NUM_THREADS) sums[d] = 0.f;
  __syncthreads();

  if (warp == PRODUCER_WARP) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int64_t row0 = (int64_t)t * prm.rows_per_stage;
        const int rows = (int)min((int64_t)prm.rows_per_stage, prm.n_rows - row0);
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t dst = base + (uint32_t)(s * prm.stage_bytes);
        const uint32_t hb = (uint32_t)rows * (uint32_t)(L * sizeof(T)), ub = (uint32_t)rows * (uint32_t)(ld * sizeof(T));
        // the rows' weights and bag ids ride along when the slice is a whole number of 16-byte units (every tile but a
        // ragged last one, whose consumers read them from global memory instead)
        const uint32_t sb = (rows & 3) == 0 && small_ok ? (uint32_t)rows * 4u : 0u;
        mbar_expect_tx(full_bar(s), hb + ub + 2 * sb);
        bulk_load_1d(dst, h + row0 * L, hb, full_bar(s));
        bulk_load_1d(dst + prm.h_stage_bytes, uv + row0 * ld, ub, full_bar(s));
        if (sb) {
          bulk_load_1d(dst + prm.ps_off, prm.p + row0, sb, full_bar(s));
          bulk_load_1d(dst + prm.ps_off + small_bytes, prm.row_seg + row0, sb, full_bar(s));
        }
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    const float q = prm.q, iq = 1.f / prm.q;
    float wcr[NI][4];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int d = 4 * lane + 128 * i;
#pragma unroll
      for (int j = 0; j < 4; ++j) wcr[i][j] = (d < D) ? prm.wc[d + j] : 0.f;
    }
    float part[NI][4], csa[NI][4], csb[GATED ? NI : 1][4];
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        part[i][j] = csa[i][j] = 0.f;
        if (GATED) csb[i][j] = 0.f;
      }
    float dsum = 0.f;
    int cur_b = -1;
    float g[NJ][8];                  // dM[cur_b] columns of this lane
    float Kb = 0.f;

    int s = 0;
    uint32_t ph = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const int64_t row0 = (int64_t)t * prm.rows_per_stage;
      const int rows = (int)min((int64_t)prm.rows_per_stage, prm.n_rows - row0);
      const uint8_t* st = smem_raw + (size_t)s * prm.stage_bytes;
      const T* hs = reinterpret_cast<const T*>(st);
      const T* us = reinterpret_cast<const T*>(st + prm.h_stage_bytes);
      const float* p_s = reinterpret_cast<const float*>(st + prm.ps_off);
      const int32_t* seg_s = reinterpret_cast<const int32_t*>(st + prm.ps_off + small_bytes);
      const bool staged = (rows & 3) == 0 && small_ok;
      mbar_wait(full_bar(s), ph);
      float pr[MAX_RPW];
      int br[MAX_RPW];
#pragma unroll
      for (int k = 0; k < MAX_RPW; ++k) {
        const int r = warp + k * CONSUMER_WARPS;
        if (r < rows) {
          pr[k] = staged ? p_s[r] : __ldg(prm.p + row0 + r);
          br[k] = staged ? seg_s[r] : __ldg(prm.row_seg + row0 + r);
        } else {
          pr[k] = 0.f;
          br[k] = cur_b;
        }
      }
#pragma unroll
      for (int k = 0; k < MAX_RPW; ++k) {
        const int r = warp + k * CONSUMER_WARPS;
        if (r >= rows) break;                                    // warp-uniform
        const int b = br[k];
        if (b != cur_b) {                                        // warp-uniform: next bag -> its dM row and K_b
          cur_b = b;
          const float* gb = prm.dM + (int64_t)b * L;
          const float* mb = prm.M + (int64_t)b * L;
          float kk = 0.f;
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            if (8 * lane + 256 * j < L) {
              float mv[8];
              load8(gb + 8 * lane + 256 * j, g[j]);
              load8(mb + 8 * lane + 256 * j, mv);
#pragma unroll
              for (int e = 0; e < 8; ++e) kk = fmaf(g[j][e], mv[e], kk);
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) g[j][e] = 0.f;
            }
          }
          kk = warp_sum(kk);
          const float scale = prm.inv_sqrt_n ? sqrtf((float)(prm.offsets[b + 1] - prm.offsets[b])) : 1.f;   // 1 / alpha_b
          Kb = kk * scale;
        }
        float tacc = 0.f;
        const T* hr = hs + (size_t)r * L;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          if (8 * lane + 256 * j < L) {
            float x[8];
            lds8(hr + 8 * lane + 256 * j, x);
#pragma unroll
            for (int e = 0; e < 8; ++e) tacc = fmaf(g[j][e], x[e], tacc);
          }
        }
        tacc = warp_sum(tacc);
        const float dsv = pr[k] * (tacc - Kb);
        const int64_t row = row0 + r;
        if (lane == 0) {
          dsum += dsv;
          if (prm.ds) prm.ds[row] = dsv;
        }
        const T* ur_s = us + (size_t)r * ld;
        T* ur = uv + row * ld;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const int d = 4 * lane + 128 * i;
          if (d < D) {
            const float4 u4 = load4(ur_s + d);
            const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
            float oa[4];
            if (GATED) {
              const float4 v4 = load4(ur_s + D + d);
              const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
              float ob[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                // uu, vv are the stored (possibly dropped-and-rescaled by q) activations; ua, va the raw tanh / sigmoid
                const float gw = dsv * wcr[i][j];
                const float ua = uu[j] * iq, va = vv[j] * iq;
                part[i][j] = fmaf(dsv, uu[j] * vv[j], part[i][j]);
                oa[j] = (q == 1.f || uu[j] != 0.f) ? gw * vv[j] * q * (1.f - ua * ua) : 0.f;
                ob[j] = (q == 1.f || vv[j] != 0.f) ? gw * uu[j] * q * va * (1.f - va) : 0.f;
                csb[i][j] += ob[j];
              }
              store4(ur + D + d, make_float4(ob[0], ob[1], ob[2], ob[3]));
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float ua = uu[j] * iq;
                part[i][j] = fmaf(dsv, uu[j], part[i][j]);
                oa[j] = (q == 1.f || uu[j] != 0.f) ? dsv * wcr[i][j] * q * (1.f - ua * ua) : 0.f;
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) csa[i][j] += oa[j];
            store4(ur + d, make_float4(oa[0], oa[1], oa[2], oa[3]));
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar(s));                  // this warp no longer reads the stage
      if (++s == S) { s = 0; ph ^= 1u; }
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int d = 4 * lane + 128 * i;
      if (d < D) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          atomicAdd(&sums[d + j], part[i][j]);
          atomicAdd(&sums[D + d + j], csa[i][j]);
          if (GATED) atomicAdd(&sums[2 * D + d + j], csb[i][j]);
        }
      }
    }
    if (prm.dbc && lane == 0 && dsum != 0.f) atomicAdd(prm.dbc, dsum);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += NUM_THREADS) atomicAdd(&prm.dwc[d], sums[d]);
  if (prm.dpre_colsum)
    for (int d = threadIdx.x; d < ld; d += NUM_THREADS) atomicAdd(&prm.dpre_colsum[d], sums[D + d]);
}

template <typename T, bool GATED, int NJ, int NI>
static int launch(const void* h, void* uv, Params& prm, cudaStream_t st) {
  const int ld = GATED ? 2 * prm.D : prm.D;
  const int row_bytes = (prm.L + ld) * (int)sizeof(T);
  int rows = STAGE_TARGET_BYTES / row_bytes;
  const int max_rows = MAX_RPW * CONSUMER_WARPS;
  rows = rows >= max_rows ? max_rows : rows >= CONSUMER_WARPS ? CONSUMER_WARPS : rows < 4 ? 4 : (rows & ~3);
  auto up128 = [](int v) { return (v + 127) / 128 * 128; };
  prm.rows_per_stage = rows;
  prm.h_stage_bytes = up128(rows * prm.L * (int)sizeof(T));
  prm.ps_off = prm.h_stage_bytes + up128(rows * ld * (int)sizeof(T));
  prm.stage_bytes = prm.ps_off + 2 * up128(rows * 4);
  const int tail = ((prm.D + ld) * 4 + 15) / 16 * 16 + 2 * MAX_STAGES * 8 + 128;
  int stages = (SMEM_BUDGET - tail) / prm.stage_bytes;
  stages = stages > MAX_STAGES ? MAX_STAGES : stages;
  if (stages < 1) {
    set_error("attnpool_bwd: a stage of %d rows x %d bytes does not fit in shared memory", rows, row_bytes);
    return MURCL_EINVAL;
  }
  prm.stages = stages;
  prm.n_tiles = (int)((prm.n_rows + rows - 1) / rows);
  const int smem = stages * prm.stage_bytes + tail;
  auto kern = attnpool_bwd_kernel<T, GATED, NJ, NI>;
  static PerDeviceOnce configured;
  if (const int slot = configured.pending(); slot >= 0) {
    MURCL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET));
    configured.mark(slot);
  }
  const int grid = prm.n_tiles < sm_count() ? prm.n_tiles : sm_count();
  kern<<<grid, NUM_THREADS, smem, st>>>(static_cast<const T*>(h), static_cast<T*>(uv), prm);
  return check_launch("attnpool_bwd_kernel");
}

template <typename T, bool GATED>
static int dispatch(const void* h, void* uv, Params& prm, cudaStream_t st) {
  const int nj = (prm.L + 255) / 256, ni = (prm.D + 127) / 128;
#define APB_CASE(NJ_, NI_) \
  if (nj <= NJ_ && ni <= NI_) return launch<T, GATED, NJ_, NI_>(h, uv, prm, st);
  APB_CASE(1, 1) APB_CASE(2, 1) APB_CASE(2, 2) APB_CASE(2, 3) APB_CASE(4, 1) APB_CASE(4, 4)
#undef APB_CASE
  set_error("attnpool_bwd: unsupported shape L=%d D=%d", prm.L, prm.D);
  return MURCL_EINVAL;
}

}  // namespace apb
}  // namespace murcl

using namespace murcl;

extern "C" {

int murcl_attnpool_bwd_supported(int L, int D, int gated, int dtype) {
  (void)gated;
  return (dtype == MURCL_BF16 || dtype == MURCL_F32) && L > 0 && L <= 1024 && L % 8 == 0 && D > 0 && D <= 512 && D % 8 == 0 ? 1 : 0;
}

int murcl_attnpool_bwd(const void* h, void* uv, const float* p, const float* M, const float* dM, const float* wc,
                       const int64_t* offsets, const int32_t* row_seg, int64_t n_rows, int B, int L, int D, int gated,
                       int inv_sqrt_n, float drop_scale, int dtype, float* ds, float* dwc, float* dbc, float* dpre_colsum,
                       void* stream) {
  MURCL_REQUIRE(h && uv && p && M && dM && wc && offsets && row_seg && dwc, "attnpool_bwd: null pointer");
  MURCL_REQUIRE(n_rows >= 0 && B >= 0, "attnpool_bwd: bad shape");
  MURCL_REQUIRE(murcl_attnpool_bwd_supported(L, D, gated, dtype),
                "attnpool_bwd: unsupported configuration L=%d D=%d dtype=%d (L <= 1024 and %% 8 == 0, D <= 512 and %% 8 == 0)", L, D,
                dtype);
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  MURCL_REQUIRE(al16(h) && al16(uv) && al16(M) && al16(dM), "attnpool_bwd: h, uv, M and dM must be 16-byte aligned");
  if (n_rows == 0 || B == 0) return MURCL_OK;
  apb::Params prm{};
  prm.n_rows = n_rows; prm.L = L; prm.D = D; prm.inv_sqrt_n = inv_sqrt_n;
  prm.q = drop_scale > 0.f ? drop_scale : 1.f;
  prm.p = p; prm.M = M; prm.dM = dM; prm.wc = wc; prm.offsets = offsets; prm.row_seg = row_seg;
  prm.ds = ds; prm.dwc = dwc; prm.dbc = dbc; prm.dpre_colsum = dpre_colsum;
  cudaStream_t st = as_stream(stream);
  if (dtype == MURCL_BF16)
    return gated ? apb::dispatch<__nv_bfloat16, true>(h, uv, prm, st) : apb::dispatch<__nv_bfloat16, false>(h, uv, prm, st);
  return gated ? apb::dispatch<float, true>(h, uv, prm, st) : apb::dispatch<float, false>(h, uv, prm, st);
}

}  // extern "C"
