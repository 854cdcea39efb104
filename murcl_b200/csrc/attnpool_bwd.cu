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
// Layout of the work: a warp owns whole rows; lane l holds columns [8l + 256j, +8) of h (one 16-byte load per j for bf16,
// two for fp32) and columns [4l + 128i, +4) of u (and v).  R = 4 (or 2) rows per warp are loaded before the first is used
// (R * (L + 2 NC) * s bytes in flight per warp, ~5 KB at L = 512, NC = 128, bf16).  dM[b] sits in registers and is
// re-fetched when the warp crosses into the next bag (bags are contiguous row ranges), together with K_b.
#include "common.cuh"

namespace murcl {
namespace apb {

constexpr int WARPS = 8;
constexpr int ROWS_PER_CTA = 256;

// rows of one warp in flight: 4 when a row's registers are few (bf16, L <= 512, <= 256 activation columns), else 2
template <typename T, bool GATED, int NJ, int NI>
constexpr int rows_in_flight() { return (sizeof(T) == 2 && NJ <= 2 && NI * (GATED ? 2 : 1) <= 1) ? 4 : 2; }

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void unpack8(const uint4& q, float (&v)[8]) {
  const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(hh[j]);
    v[2 * j] = f.x;
    v[2 * j + 1] = f.y;
  }
}

// raw 8-element chunk of a row as it comes from memory (kept packed while in flight: fewer registers)
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> {
  uint4 q;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { q = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void get(float (&v)[8]) const { unpack8(q, v); }
};
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = __ldg(reinterpret_cast<const float4*>(p));
    b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  }
  __device__ __forceinline__ void get(float (&v)[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};
template <typename T> struct Raw4;
template <> struct Raw4<__nv_bfloat16> {
  uint2 q;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { q = *reinterpret_cast<const uint2*>(p); }
  __device__ __forceinline__ float4 get() const {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&q.x), b = *reinterpret_cast<const __nv_bfloat162*>(&q.y);
    const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
};
template <> struct Raw4<float> {
  float4 q;
  __device__ __forceinline__ void load(const float* p) { q = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ float4 get() const { return q; }
};

struct Params {
  int64_t n_rows;
  int L, D, inv_sqrt_n, rows_per_cta;
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
__global__ void __launch_bounds__(32 * WARPS, 2) attnpool_bwd_kernel(const T* __restrict__ h, T* __restrict__ uv, const Params prm) {
  constexpr int R = rows_in_flight<T, GATED, NJ, NI>();
  extern __shared__ float sm[];      // [D] dwc partial | [NC] column sums partial
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int L = prm.L, D = prm.D;
  const int ld = GATED ? 2 * D : D;
  for (int d = threadIdx.x; d < D + ld; d += blockDim.x) sm[d] = 0.f;
  __syncthreads();
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
  float g[NJ][8];                    // dM[cur_b] columns of this lane
  float Kb = 0.f;

  for (int64_t chunk = blockIdx.x; chunk * prm.rows_per_cta < prm.n_rows; chunk += gridDim.x) {
    const int64_t r0 = chunk * prm.rows_per_cta;
    const int64_t r1 = min(prm.n_rows, r0 + prm.rows_per_cta);
    for (int64_t base = r0 + w; base < r1; base += WARPS * R) {
      Raw8<T> hq[R][NJ];
      Raw4<T> uq[R][NI], vq[GATED ? R : 1][GATED ? NI : 1];
      float pr[R];
      int br[R];
#pragma unroll
      for (int u = 0; u < R; ++u) {
        const int64_t row = base + (int64_t)u * WARPS;
        if (row < r1) {
          const T* hr = h + row * L;
#pragma unroll
          for (int j = 0; j < NJ; ++j)
            if (8 * lane + 256 * j < L) hq[u][j].load(hr + 8 * lane + 256 * j);
          const T* ur = uv + row * ld;
#pragma unroll
          for (int i = 0; i < NI; ++i)
            if (4 * lane + 128 * i < D) {
              uq[u][i].load(ur + 4 * lane + 128 * i);
              if (GATED) vq[u][i].load(ur + D + 4 * lane + 128 * i);
            }
          pr[u] = __ldg(prm.p + row);
          br[u] = __ldg(prm.row_seg + row);
        }
      }
#pragma unroll
      for (int u = 0; u < R; ++u) {
        const int64_t row = base + (int64_t)u * WARPS;
        if (row >= r1) break;                                  // warp-uniform
        if (br[u] != cur_b) {                                  // warp-uniform: next bag -> its dM row and K_b
          cur_b = br[u];
          const float* gb = prm.dM + (int64_t)cur_b * L;
          const float* mb = prm.M + (int64_t)cur_b * L;
          float k = 0.f;
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            if (8 * lane + 256 * j < L) {
              float mv[8];
              load8(gb + 8 * lane + 256 * j, g[j]);
              load8(mb + 8 * lane + 256 * j, mv);
#pragma unroll
              for (int e = 0; e < 8; ++e) k = fmaf(g[j][e], mv[e], k);
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) g[j][e] = 0.f;
            }
          }
          k = warp_sum(k);
          const float scale = prm.inv_sqrt_n ? sqrtf((float)(prm.offsets[cur_b + 1] - prm.offsets[cur_b])) : 1.f;   // 1 / alpha_b
          Kb = k * scale;
        }
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          if (8 * lane + 256 * j < L) {
            float x[8];
            hq[u][j].get(x);
#pragma unroll
            for (int e = 0; e < 8; ++e) t = fmaf(g[j][e], x[e], t);
          }
        }
        t = warp_sum(t);
        const float dsv = pr[u] * (t - Kb);
        if (lane == 0) {
          dsum += dsv;
          if (prm.ds) prm.ds[row] = dsv;
        }
        T* ur = uv + row * ld;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const int d = 4 * lane + 128 * i;
          if (d < D) {
            const float4 u4 = uq[u][i].get();
            const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
            float oa[4];
            if (GATED) {
              const float4 v4 = vq[u][i].get();
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
    }
  }
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int d = 4 * lane + 128 * i;
    if (d < D) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        atomicAdd(&sm[d + j], part[i][j]);
        atomicAdd(&sm[D + d + j], csa[i][j]);
        if (GATED) atomicAdd(&sm[2 * D + d + j], csb[i][j]);
      }
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) atomicAdd(&prm.dwc[d], sm[d]);
  if (prm.dpre_colsum)
    for (int d = threadIdx.x; d < ld; d += blockDim.x) atomicAdd(&prm.dpre_colsum[d], sm[D + d]);
  if (prm.dbc && lane == 0 && dsum != 0.f) atomicAdd(prm.dbc, dsum);
}

template <typename T, bool GATED, int NJ, int NI>
static int launch(const void* h, void* uv, const Params& prm, cudaStream_t st) {
  const int ld = GATED ? 2 * prm.D : prm.D;
  const size_t smem = (size_t)(prm.D + ld) * sizeof(float);
  const int64_t chunks = (prm.n_rows + prm.rows_per_cta - 1) / prm.rows_per_cta;
  const int64_t cap = (int64_t)sm_count() * 4;
  const int grid = (int)(chunks < cap ? chunks : cap);
  attnpool_bwd_kernel<T, GATED, NJ, NI><<<grid, 32 * WARPS, smem, st>>>(static_cast<const T*>(h), static_cast<T*>(uv), prm);
  return check_launch("attnpool_bwd_kernel");
}

template <typename T, bool GATED>
static int dispatch(const void* h, void* uv, const Params& prm, cudaStream_t st) {
  const int nj = (prm.L + 255) / 256, ni = (prm.D + 127) / 128;
#define APB_CASE(NJ_, NI_) \
  if (nj <= NJ_ && ni <= NI_) return launch<T, GATED, NJ_, NI_>(h, uv, prm, st);
  APB_CASE(1, 1) APB_CASE(2, 1) APB_CASE(2, 2) APB_CASE(2, 3) APB_CASE(4, 1) APB_CASE(4, 2) APB_CASE(4, 4)
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
  return (dtype == MURCL_BF16 || dtype == MURCL_F32) && L > 0 && L <= 1024 && L % 8 == 0 && D > 0 && D <= 512 && D % 4 == 0 ? 1 : 0;
}

int murcl_attnpool_bwd(const void* h, void* uv, const float* p, const float* M, const float* dM, const float* wc,
                       const int64_t* offsets, const int32_t* row_seg, int64_t n_rows, int B, int L, int D, int gated,
                       int inv_sqrt_n, float drop_scale, int dtype, float* ds, float* dwc, float* dbc, float* dpre_colsum,
                       void* stream) {
  MURCL_REQUIRE(h && uv && p && M && dM && wc && offsets && row_seg && dwc, "attnpool_bwd: null pointer");
  MURCL_REQUIRE(n_rows >= 0 && B >= 0, "attnpool_bwd: bad shape");
  MURCL_REQUIRE(murcl_attnpool_bwd_supported(L, D, gated, dtype),
                "attnpool_bwd: unsupported configuration L=%d D=%d dtype=%d (L <= 1024 and %% 8 == 0, D <= 512 and %% 4 == 0)", L, D,
                dtype);
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  MURCL_REQUIRE(al16(h) && al16(uv) && al16(M) && al16(dM), "attnpool_bwd: h, uv, M and dM must be 16-byte aligned");
  if (n_rows == 0 || B == 0) return MURCL_OK;
  apb::Params prm{};
  prm.n_rows = n_rows; prm.L = L; prm.D = D; prm.inv_sqrt_n = inv_sqrt_n;
  prm.rows_per_cta = apb::ROWS_PER_CTA;
  prm.q = drop_scale > 0.f ? drop_scale : 1.f;
  prm.p = p; prm.M = M; prm.dM = dM; prm.wc = wc; prm.offsets = offsets; prm.row_seg = row_seg;
  prm.ds = ds; prm.dwc = dwc; prm.dbc = dbc; prm.dpre_colsum = dpre_colsum;
  cudaStream_t st = as_stream(stream);
  if (dtype == MURCL_BF16)
    return gated ? apb::dispatch<__nv_bfloat16, true>(h, uv, prm, st) : apb::dispatch<__nv_bfloat16, false>(h, uv, prm, st);
  return gated ? apb::dispatch<float, true>(h, uv, prm, st) : apb::dispatch<float, false>(h, uv, prm, st);
}

}  // extern "C"
