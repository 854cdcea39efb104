// Attention pooling pieces: score projection tail, segmented softmax, segmented weighted row sum
// and their backward.  Replaces abmil.py:38-42, clam.py:56-60,139-170, dsmil.py:76-78.
// All kernels are HBM-bound row streamers (one warp per instance row, 8/16-byte vector loads).
#include "common.cuh"

namespace murcl {

// ---- s[n] = wc . g(n) + bc --------------------------------------------------------------------
template <typename T, bool GATED>
__global__ void __launch_bounds__(256) attn_score_fwd_kernel(const T* __restrict__ uv, const float* __restrict__ wc,
                                                             const float* __restrict__ bc, float* __restrict__ s,
                                                             int64_t N, int D) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const int ld = GATED ? 2 * D : D;
  const T* r = uv + row * ld;
  float acc = 0.f;
  if ((D & 3) == 0) {
    for (int d = lane * 4; d < D; d += 128) {
      const float4 u = load4(r + d);
      const float4 w = *reinterpret_cast<const float4*>(wc + d);
      if (GATED) {
        const float4 v = load4(r + D + d);
        acc += w.x * (u.x * v.x) + w.y * (u.y * v.y) + w.z * (u.z * v.z) + w.w * (u.w * v.w);
      } else {
        acc += w.x * u.x + w.y * u.y + w.z * u.z + w.w * u.w;
      }
    }
  } else {
    for (int d = lane; d < D; d += 32) {
      float g = Store<T>::load(r + d);
      if (GATED) g *= Store<T>::load(r + D + d);
      acc += wc[d] * g;
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) s[row] = acc + (bc ? bc[0] : 0.f);
}

// ---- segmented softmax --------------------------------------------------------------------------
__global__ void __launch_bounds__(256) seg_softmax_kernel(const float* __restrict__ s, const int64_t* __restrict__ offsets,
                                                          int C, int inv_sqrt_n, float* __restrict__ p,
                                                          float* __restrict__ stats) {
  __shared__ float red[32];
  const int b = blockIdx.x, c = blockIdx.y;
  const int64_t lo = offsets[b], hi = offsets[b + 1];
  float m = -INFINITY;
  for (int64_t n = lo + threadIdx.x; n < hi; n += blockDim.x) m = fmaxf(m, s[n * C + c]);
  m = block_max(m, red);
  float l = 0.f;
  for (int64_t n = lo + threadIdx.x; n < hi; n += blockDim.x) l += expf(s[n * C + c] - m);
  l = block_sum(l, red);
  const float root = inv_sqrt_n ? sqrtf((float)(hi - lo)) : 1.f;
  for (int64_t n = lo + threadIdx.x; n < hi; n += blockDim.x) {
    float v = __fdiv_rn(expf(s[n * C + c] - m), l);
    if (inv_sqrt_n) v = __fdiv_rn(v, root);
    p[n * C + c] = v;
  }
  if (threadIdx.x == 0 && stats) {
    stats[((int64_t)b * C + c) * 2 + 0] = m;
    stats[((int64_t)b * C + c) * 2 + 1] = l;
  }
}

// ---- segmented weighted row sum -----------------------------------------------------------------
// grid (B, S): CTA (b, j) reduces the j-th slice of bag b's rows into ws[b, j, c, :]; a second
// kernel adds the S slices in a fixed order (deterministic).
constexpr int WSUM_MAXC = 4;

template <typename T>
__global__ void __launch_bounds__(256) seg_wsum_partial_kernel(const float* __restrict__ p, const T* __restrict__ h,
                                                               const int64_t* __restrict__ offsets, int C, int L,
                                                               float* __restrict__ ws) {
  extern __shared__ float sm[];   // [rows_par][C][L] for the cross-row-group reduce
  const int b = blockIdx.x, j = blockIdx.y, S = gridDim.y;
  const int64_t lo = offsets[b], hi = offsets[b + 1];
  const int64_t len = hi - lo;
  const int64_t r0 = lo + len * j / S, r1 = lo + len * (j + 1) / S;
  const int groups = (L + 3) / 4;                        // column groups of 4
  const int cg_par = min(groups, (int)blockDim.x);       // column groups handled concurrently
  const int rows_par = blockDim.x / cg_par;
  const int cg = threadIdx.x % cg_par, rg = threadIdx.x / cg_par;
  float* out = ws + ((int64_t)b * S + j) * C * L;
  for (int g0 = 0; g0 < groups; g0 += cg_par) {
    const int col = (g0 + cg) * 4;
    float4 acc[WSUM_MAXC];
#pragma unroll
    for (int c = 0; c < WSUM_MAXC; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rg < rows_par && g0 + cg < groups) {
      for (int64_t n = r0 + rg; n < r1; n += rows_par) {
        float4 x;
        if (col + 4 <= L && (L & 3) == 0) {
          x = load4(h + n * L + col);
        } else {
          x.x = col + 0 < L ? Store<T>::load(h + n * L + col + 0) : 0.f;
          x.y = col + 1 < L ? Store<T>::load(h + n * L + col + 1) : 0.f;
          x.z = col + 2 < L ? Store<T>::load(h + n * L + col + 2) : 0.f;
          x.w = col + 3 < L ? Store<T>::load(h + n * L + col + 3) : 0.f;
        }
#pragma unroll
        for (int c = 0; c < WSUM_MAXC; ++c) {
          if (c < C) {
            const float w = p[n * C + c];
            acc[c].x = fmaf(w, x.x, acc[c].x);
            acc[c].y = fmaf(w, x.y, acc[c].y);
            acc[c].z = fmaf(w, x.z, acc[c].z);
            acc[c].w = fmaf(w, x.w, acc[c].w);
          }
        }
      }
    }
    // reduce over row groups through shared memory
    for (int c = 0; c < C; ++c) {
      __syncthreads();
      if (rg < rows_par && g0 + cg < groups) {
        float* dst = sm + ((int64_t)rg * cg_par + cg) * 4;
        dst[0] = acc[c].x; dst[1] = acc[c].y; dst[2] = acc[c].z; dst[3] = acc[c].w;
      }
      __syncthreads();
      if (rg == 0 && g0 + cg < groups) {
        float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < rows_par; ++r) {
          const float* src = sm + ((int64_t)r * cg_par + cg) * 4;
          tot.x += src[0]; tot.y += src[1]; tot.z += src[2]; tot.w += src[3];
        }
        if (col + 0 < L) out[(int64_t)c * L + col + 0] = tot.x;
        if (col + 1 < L) out[(int64_t)c * L + col + 1] = tot.y;
        if (col + 2 < L) out[(int64_t)c * L + col + 2] = tot.z;
        if (col + 3 < L) out[(int64_t)c * L + col + 3] = tot.w;
      }
    }
  }
}

// Fast path: 16-byte loads (8 bf16 / 4 fp32 columns per thread), L/VEC column threads x 256/(L/VEC) row groups,
// four rows in flight per thread.  Requires L % VEC == 0 and L/VEC <= 256.
template <typename T, int C>
__global__ void __launch_bounds__(256) seg_wsum_partial_vec_kernel(const float* __restrict__ p, const T* __restrict__ h,
                                                                   const int64_t* __restrict__ offsets, int L,
                                                                   float* __restrict__ ws) {
  constexpr int VEC = 16 / (int)sizeof(T);
  extern __shared__ float sm[];            // [rows_par][C][L]
  const int b = blockIdx.x, j = blockIdx.y, S = gridDim.y;
  const int64_t lo = offsets[b], hi = offsets[b + 1];
  const int64_t len = hi - lo;
  const int64_t r0 = lo + len * j / S, r1 = lo + len * (j + 1) / S;
  const int col_threads = L / VEC;
  const int rows_par = blockDim.x / col_threads;
  const int ct = threadIdx.x % col_threads, rg = threadIdx.x / col_threads;
  const int col = ct * VEC;
  float acc[C][VEC];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[c][v] = 0.f;
  auto load_row = [&](int64_t n, float (&x)[VEC]) {
    const float4 a = load4(h + n * L + col);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
    if (VEC == 8) {
      const float4 b4 = load4(h + n * L + col + 4);
      x[4 % VEC] = b4.x; x[5 % VEC] = b4.y; x[6 % VEC] = b4.z; x[7 % VEC] = b4.w;
    }
  };
  auto fma_row = [&](int64_t n, const float (&x)[VEC]) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float w = p[n * C + c];
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[c][v] = fmaf(w, x[v], acc[c][v]);
    }
  };
  if (rg < rows_par) {
    int64_t n = r0 + rg;
    const int64_t step = rows_par;
    for (; n + 3 * step < r1; n += 4 * step) {
      float x0[VEC], x1[VEC], x2[VEC], x3[VEC];
      load_row(n, x0); load_row(n + step, x1); load_row(n + 2 * step, x2); load_row(n + 3 * step, x3);
      fma_row(n, x0); fma_row(n + step, x1); fma_row(n + 2 * step, x2); fma_row(n + 3 * step, x3);
    }
    for (; n < r1; n += step) {
      float x0[VEC];
      load_row(n, x0);
      fma_row(n, x0);
    }
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int v = 0; v < VEC; ++v) sm[((int64_t)rg * C + c) * L + col + v] = acc[c][v];
  }
  __syncthreads();
  float* out = ws + ((int64_t)b * S + j) * C * L;
  for (int i = threadIdx.x; i < C * L; i += blockDim.x) {
    float t = 0.f;
    for (int r = 0; r < rows_par; ++r) t += sm[(int64_t)r * C * L + i];
    out[i] = t;
  }
}

__global__ void __launch_bounds__(256) seg_wsum_final_kernel(const float* __restrict__ ws, int S, int64_t per_bag,
                                                             float* __restrict__ out, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t b = i / per_bag, r = i % per_bag;
  float s = 0.f;
  for (int j = 0; j < S; ++j) s += ws[(b * S + j) * per_bag + r];
  out[i] = s;
}

static int wsum_splits(int64_t n_rows, int B) {
  int s = (4 * sm_count() + B - 1) / B;
  const int64_t avg = n_rows / (B > 0 ? B : 1);
  const int by_rows = (int)((avg + 63) / 64);      // at least ~64 rows per slice
  if (s > by_rows) s = by_rows;
  if (s < 1) s = 1;
  if (s > 1024) s = 1024;
  return s;
}

// ---- backward w.r.t. scores -----------------------------------------------------------------------
__global__ void __launch_bounds__(128) pool_k_kernel(const float* __restrict__ dM, const float* __restrict__ M,
                                                     const int64_t* __restrict__ offsets, int C, int L, int inv_sqrt_n,
                                                     float* __restrict__ kbuf) {
  __shared__ float red[32];
  const int b = blockIdx.x, c = blockIdx.y;
  const float* a = dM + ((int64_t)b * C + c) * L;
  const float* m = M + ((int64_t)b * C + c) * L;
  float acc = 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) acc = fmaf(a[i], m[i], acc);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    const float scale = inv_sqrt_n ? sqrtf((float)(offsets[b + 1] - offsets[b])) : 1.f;   // 1 / post_scale
    kbuf[(int64_t)b * C + c] = acc * scale;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) pool_bwd_scores_kernel(const float* __restrict__ p, const T* __restrict__ h,
                                                              const float* __restrict__ dM,
                                                              const float* __restrict__ kbuf,
                                                              const int32_t* __restrict__ row_seg, int64_t n_rows, int C,
                                                              int L, float* __restrict__ ds) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int b = row_seg[row];
  const T* r = h + row * L;
  float acc[WSUM_MAXC];
#pragma unroll
  for (int c = 0; c < WSUM_MAXC; ++c) acc[c] = 0.f;
  if ((L & 3) == 0) {
    for (int d = lane * 4; d < L; d += 128) {
      const float4 x = load4(r + d);
#pragma unroll
      for (int c = 0; c < WSUM_MAXC; ++c) {
        if (c < C) {
          const float4 g = *reinterpret_cast<const float4*>(dM + ((int64_t)b * C + c) * L + d);
          acc[c] += g.x * x.x + g.y * x.y + g.z * x.z + g.w * x.w;
        }
      }
    }
  } else {
    for (int d = lane; d < L; d += 32) {
      const float x = Store<T>::load(r + d);
#pragma unroll
      for (int c = 0; c < WSUM_MAXC; ++c)
        if (c < C) acc[c] = fmaf(dM[((int64_t)b * C + c) * L + d], x, acc[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < WSUM_MAXC; ++c) {
    if (c < C) {
      const float t = warp_sum(acc[c]);
      if (lane == 0) ds[row * C + c] = p[row * C + c] * (t - kbuf[(int64_t)b * C + c]);
    }
  }
}

// Fast path for one score column (ABMIL / CLAM): 16-byte loads, two rows of a warp in flight.
template <typename T>
__global__ void __launch_bounds__(256) pool_bwd_scores_c1_kernel(const float* __restrict__ p, const T* __restrict__ h,
                                                                 const float* __restrict__ dM, const float* __restrict__ kbuf,
                                                                 const int32_t* __restrict__ row_seg, int64_t n_rows, int L,
                                                                 float* __restrict__ ds) {
  constexpr int VEC = 16 / (int)sizeof(T);
  const int lane = threadIdx.x & 31;
  const int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2;
  if (row0 >= n_rows) return;
  const bool two = row0 + 1 < n_rows;
  const int b0 = row_seg[row0], b1 = two ? row_seg[row0 + 1] : b0;
  float a0 = 0.f, a1 = 0.f;
  for (int d = lane * VEC; d < L; d += 32 * VEC) {
    float x0[8], x1[8], g0[8], g1[8];
    {
      const float4 u = load4(h + row0 * L + d);
      x0[0] = u.x; x0[1] = u.y; x0[2] = u.z; x0[3] = u.w;
      if (VEC == 8) { const float4 v = load4(h + row0 * L + d + 4); x0[4] = v.x; x0[5] = v.y; x0[6] = v.z; x0[7] = v.w; }
    }
    if (two) {
      const float4 u = load4(h + (row0 + 1) * L + d);
      x1[0] = u.x; x1[1] = u.y; x1[2] = u.z; x1[3] = u.w;
      if (VEC == 8) { const float4 v = load4(h + (row0 + 1) * L + d + 4); x1[4] = v.x; x1[5] = v.y; x1[6] = v.z; x1[7] = v.w; }
    }
#pragma unroll
    for (int i = 0; i < VEC; i += 4) {
      const float4 q0 = *reinterpret_cast<const float4*>(dM + (int64_t)b0 * L + d + i);
      g0[i] = q0.x; g0[i + 1] = q0.y; g0[i + 2] = q0.z; g0[i + 3] = q0.w;
      const float4 q1 = *reinterpret_cast<const float4*>(dM + (int64_t)b1 * L + d + i);
      g1[i] = q1.x; g1[i + 1] = q1.y; g1[i + 2] = q1.z; g1[i + 3] = q1.w;
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      a0 = fmaf(g0[i], x0[i], a0);
      if (two) a1 = fmaf(g1[i], x1[i], a1);
    }
  }
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
  if (lane == 0) {
    ds[row0] = p[row0] * (a0 - kbuf[b0]);
    if (two) ds[row0 + 1] = p[row0 + 1] * (a1 - kbuf[b1]);
  }
}

// dh[n,:] (+)= sum_c p[n,c] * dM[b,c,:]   (direct term of the pooling backward)
template <typename T>
__global__ void __launch_bounds__(256) pool_bwd_direct_kernel(const float* __restrict__ p, const float* __restrict__ dM,
                                                              const int32_t* __restrict__ row_seg, int64_t n_rows, int C,
                                                              int L, T* __restrict__ dh, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int b = row_seg[row];
  for (int d = lane; d < L; d += 32) {
    float v = accumulate ? Store<T>::load(dh + row * L + d) : 0.f;
    for (int c = 0; c < C; ++c) v = fmaf(p[row * C + c], dM[((int64_t)b * C + c) * L + d], v);
    Store<T>::store(dh + row * L + d, v);
  }
}

// ---- backward through the score projection tail -------------------------------------------------
// Lane owns columns 4*lane + 128*i (8/16-byte accesses).  Besides d(pre-activation) (written over uv) the kernel
// accumulates dwc[D] and, when dpre_colsum != NULL, the column sums of what it writes (= the bias gradient of the
// attention projection), so no separate pass over the [N, D|2D] tensor is needed.
template <typename T, bool GATED>
__global__ void __launch_bounds__(256) attn_score_bwd_kernel(T* __restrict__ uv, const float* __restrict__ wc,
                                                             const float* __restrict__ ds, float* __restrict__ dwc,
                                                             float* __restrict__ dbc, float* __restrict__ dpre_colsum,
                                                             int64_t N, int D, int rows_per_cta, float q) {
  extern __shared__ float sm[];   // [D] dwc partial, [2D] column sums partial
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int ld = GATED ? 2 * D : D;
  for (int d = threadIdx.x; d < D + ld; d += blockDim.x) sm[d] = 0.f;
  __syncthreads();
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = min(N, r0 + rows_per_cta);
  float dsum = 0.f;
  const float iq = 1.f / q;       // q = 1/(1-p) when a dropout followed the activations, else 1
  constexpr int MAXI = 4;         // D <= 512
  float part[MAXI][4], csa[MAXI][4], csb[MAXI][4];
#pragma unroll
  for (int i = 0; i < MAXI; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) part[i][j] = csa[i][j] = csb[i][j] = 0.f;
  for (int64_t row = r0 + w; row < r1; row += nw) {
    const float g = ds[row];
    if (lane == 0) dsum += g;
    T* r = uv + row * ld;
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      const int d = 4 * lane + 128 * i;
      if (d < D) {
        const float4 u4 = load4(r + d);
        const float4 w4 = *reinterpret_cast<const float4*>(wc + d);
        const float u[4] = {u4.x, u4.y, u4.z, u4.w};
        const float ww[4] = {w4.x, w4.y, w4.z, w4.w};
        float oa[4], ob[4];
        if (GATED) {
          const float4 v4 = load4(r + D + d);
          const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // u, v are the stored (possibly dropped-and-rescaled by q) activations; ua, va the raw tanh / sigmoid
            const float gw = g * ww[j];
            const float ua = u[j] * iq, va = v[j] * iq;
            part[i][j] = fmaf(g, u[j] * v[j], part[i][j]);
            oa[j] = (q == 1.f || u[j] != 0.f) ? gw * v[j] * q * (1.f - ua * ua) : 0.f;
            ob[j] = (q == 1.f || v[j] != 0.f) ? gw * u[j] * q * va * (1.f - va) : 0.f;
            csb[i][j] += ob[j];
          }
          store4(r + D + d, make_float4(ob[0], ob[1], ob[2], ob[3]));
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float ua = u[j] * iq;
            part[i][j] = fmaf(g, u[j], part[i][j]);
            oa[j] = (q == 1.f || u[j] != 0.f) ? g * ww[j] * q * (1.f - ua * ua) : 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) csa[i][j] += oa[j];
        store4(r + d, make_float4(oa[0], oa[1], oa[2], oa[3]));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAXI; ++i) {
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
  for (int d = threadIdx.x; d < D; d += blockDim.x) atomicAdd(&dwc[d], sm[d]);
  if (dpre_colsum)
    for (int d = threadIdx.x; d < ld; d += blockDim.x) atomicAdd(&dpre_colsum[d], sm[D + d]);
  if (dbc && lane == 0 && dsum != 0.f) atomicAdd(dbc, dsum);
}

// Any-D fallback (D not a multiple of 4 or > 512): one column per lane step, no fused column sums.
template <typename T, bool GATED>
__global__ void __launch_bounds__(256) attn_score_bwd_generic_kernel(T* __restrict__ uv, const float* __restrict__ wc,
                                                                     const float* __restrict__ ds, float* __restrict__ dwc,
                                                                     float* __restrict__ dbc, int64_t N, int D, float q) {
  const float iq = 1.f / q;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const int ld = GATED ? 2 * D : D;
  const float g = ds[row];
  T* r = uv + row * ld;
  for (int d = lane; d < D; d += 32) {
    const float u = Store<T>::load(r + d);
    const float gw = g * wc[d];
    if (GATED) {
      const float v = Store<T>::load(r + D + d);
      const float ua = u * iq, va = v * iq;
      atomicAdd(&dwc[d], g * u * v);
      Store<T>::store(r + d, (q == 1.f || u != 0.f) ? gw * v * q * (1.f - ua * ua) : 0.f);
      Store<T>::store(r + D + d, (q == 1.f || v != 0.f) ? gw * u * q * va * (1.f - va) : 0.f);
    } else {
      const float ua = u * iq;
      atomicAdd(&dwc[d], g * u);
      Store<T>::store(r + d, (q == 1.f || u != 0.f) ? gw * q * (1.f - ua * ua) : 0.f);
    }
  }
  if (dbc && lane == 0) atomicAdd(dbc, g);
}

}  // namespace murcl

using namespace murcl;

extern "C" {

int murcl_attn_score_fwd(const void* uv, const float* wc, const float* bc, float* s, int64_t N, int D, int gated,
                         int dtype, void* stream) {
  MURCL_REQUIRE(uv && wc && s, "attn_score_fwd: null pointer");
  MURCL_REQUIRE(N >= 0 && D > 0, "attn_score_fwd: bad shape");
  if (N == 0) return MURCL_OK;
  const int grid = ceil_div(N, 8);
  cudaStream_t st = as_stream(stream);
  if (dtype == MURCL_F32) {
    if (gated) attn_score_fwd_kernel<float, true><<<grid, 256, 0, st>>>((const float*)uv, wc, bc, s, N, D);
    else attn_score_fwd_kernel<float, false><<<grid, 256, 0, st>>>((const float*)uv, wc, bc, s, N, D);
  } else if (dtype == MURCL_BF16) {
    if (gated) attn_score_fwd_kernel<__nv_bfloat16, true><<<grid, 256, 0, st>>>((const __nv_bfloat16*)uv, wc, bc, s, N, D);
    else attn_score_fwd_kernel<__nv_bfloat16, false><<<grid, 256, 0, st>>>((const __nv_bfloat16*)uv, wc, bc, s, N, D);
  } else {
    MURCL_REQUIRE(false, "attn_score_fwd: bad dtype %d", dtype);
  }
  return check_launch("attn_score_fwd_kernel");
}

int murcl_seg_softmax(const float* s, const int64_t* offsets, int B, int C, int inv_sqrt_n, float* p, float* stats,
                      void* stream) {
  MURCL_REQUIRE(s && offsets && p, "seg_softmax: null pointer");
  MURCL_REQUIRE(B >= 0 && C > 0 && C <= 65535, "seg_softmax: bad shape");
  if (B == 0) return MURCL_OK;
  seg_softmax_kernel<<<dim3(B, C), 256, 0, as_stream(stream)>>>(s, offsets, C, inv_sqrt_n, p, stats);
  return check_launch("seg_softmax_kernel");
}

int64_t murcl_seg_wsum_workspace(int64_t n_rows, int B, int C, int L) {
  return (int64_t)B * wsum_splits(n_rows, B) * C * L;
}

int murcl_seg_wsum(const float* p, const void* h, const int64_t* offsets, int64_t n_rows, int B, int C, int L, int dtype,
                   float* out, float* workspace, void* stream) {
  MURCL_REQUIRE(p && h && offsets && out && workspace, "seg_wsum: null pointer");
  MURCL_REQUIRE(B >= 0 && C > 0 && C <= WSUM_MAXC && L > 0, "seg_wsum: bad shape B=%d C=%d L=%d (C <= %d)", B, C, L,
                WSUM_MAXC);
  if (B == 0) return MURCL_OK;
  const int S = wsum_splits(n_rows, B);
  MURCL_REQUIRE(S <= 65535, "seg_wsum: too many splits");
  cudaStream_t st = as_stream(stream);
  dim3 grid(B, S);
  MURCL_REQUIRE(dtype == MURCL_F32 || dtype == MURCL_BF16, "seg_wsum: bad dtype %d", dtype);
  const int vec = dtype == MURCL_BF16 ? 8 : 4;
  const bool fast = (L % vec == 0) && (L / vec <= 256) && (C == 1 || C == 2) &&
                    ((reinterpret_cast<uintptr_t>(h) & 15) == 0);
  if (fast) {
    const int rows_par = 256 / (L / vec);
    const size_t smem = sizeof(float) * (size_t)rows_par * C * L;          // <= 256/ (L/vec) * C * L floats <= 16 KB * C
    if (dtype == MURCL_BF16) {
      if (C == 1) seg_wsum_partial_vec_kernel<__nv_bfloat16, 1><<<grid, 256, smem, st>>>(p, (const __nv_bfloat16*)h, offsets, L, workspace);
      else seg_wsum_partial_vec_kernel<__nv_bfloat16, 2><<<grid, 256, smem, st>>>(p, (const __nv_bfloat16*)h, offsets, L, workspace);
    } else {
      if (C == 1) seg_wsum_partial_vec_kernel<float, 1><<<grid, 256, smem, st>>>(p, (const float*)h, offsets, L, workspace);
      else seg_wsum_partial_vec_kernel<float, 2><<<grid, 256, smem, st>>>(p, (const float*)h, offsets, L, workspace);
    }
  } else {
    const size_t smem = sizeof(float) * 256 * 4;
    if (dtype == MURCL_F32) seg_wsum_partial_kernel<float><<<grid, 256, smem, st>>>(p, (const float*)h, offsets, C, L, workspace);
    else seg_wsum_partial_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(p, (const __nv_bfloat16*)h, offsets, C, L, workspace);
  }
  int rc = check_launch("seg_wsum_partial_kernel");
  if (rc != MURCL_OK) return rc;
  const int64_t total = (int64_t)B * C * L;
  seg_wsum_final_kernel<<<ceil_div(total, 256), 256, 0, st>>>(workspace, S, (int64_t)C * L, out, total);
  return check_launch("seg_wsum_final_kernel");
}

int murcl_pool_bwd_scores(const float* p, const void* h, const float* dM, const float* M, const int64_t* offsets,
                          const int32_t* row_seg, int64_t n_rows, int B, int C, int L, int inv_sqrt_n, int dtype,
                          float* ds, float* kbuf, void* stream) {
  MURCL_REQUIRE(p && h && dM && M && offsets && row_seg && ds && kbuf, "pool_bwd_scores: null pointer");
  MURCL_REQUIRE(B >= 0 && C > 0 && C <= WSUM_MAXC && L > 0, "pool_bwd_scores: bad shape");
  if (B == 0 || n_rows == 0) return MURCL_OK;
  cudaStream_t st = as_stream(stream);
  pool_k_kernel<<<dim3(B, C), 128, 0, st>>>(dM, M, offsets, C, L, inv_sqrt_n, kbuf);
  int rc = check_launch("pool_k_kernel");
  if (rc != MURCL_OK) return rc;
  MURCL_REQUIRE(dtype == MURCL_F32 || dtype == MURCL_BF16, "pool_bwd_scores: bad dtype %d", dtype);
  const int vec = dtype == MURCL_BF16 ? 8 : 4;
  if (C == 1 && L % vec == 0 && ((reinterpret_cast<uintptr_t>(h) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dM) & 15) == 0)) {
    const int grid = ceil_div(n_rows, 16);
    if (dtype == MURCL_F32)
      pool_bwd_scores_c1_kernel<float><<<grid, 256, 0, st>>>(p, (const float*)h, dM, kbuf, row_seg, n_rows, L, ds);
    else
      pool_bwd_scores_c1_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p, (const __nv_bfloat16*)h, dM, kbuf, row_seg, n_rows, L, ds);
    return check_launch("pool_bwd_scores_c1_kernel");
  }
  const int grid = ceil_div(n_rows, 8);
  if (dtype == MURCL_F32)
    pool_bwd_scores_kernel<float><<<grid, 256, 0, st>>>(p, (const float*)h, dM, kbuf, row_seg, n_rows, C, L, ds);
  else
    pool_bwd_scores_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p, (const __nv_bfloat16*)h, dM, kbuf, row_seg, n_rows, C, L, ds);
  return check_launch("pool_bwd_scores_kernel");
}

int murcl_pool_bwd_direct(const float* p, const float* dM, const int32_t* row_seg, int64_t n_rows, int C, int L,
                          int dtype, void* dh, int accumulate, void* stream) {
  MURCL_REQUIRE(p && dM && row_seg && dh, "pool_bwd_direct: null pointer");
  MURCL_REQUIRE(C > 0 && L > 0, "pool_bwd_direct: bad shape");
  if (n_rows == 0) return MURCL_OK;
  const int grid = ceil_div(n_rows, 8);
  cudaStream_t st = as_stream(stream);
  if (dtype == MURCL_F32) pool_bwd_direct_kernel<float><<<grid, 256, 0, st>>>(p, dM, row_seg, n_rows, C, L, (float*)dh, accumulate);
  else if (dtype == MURCL_BF16)
    pool_bwd_direct_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p, dM, row_seg, n_rows, C, L, (__nv_bfloat16*)dh, accumulate);
  else MURCL_REQUIRE(false, "pool_bwd_direct: bad dtype %d", dtype);
  return check_launch("pool_bwd_direct_kernel");
}

int murcl_attn_score_bwd(void* uv, const float* wc, const float* ds, float* dwc, float* dbc, float* dpre_colsum, int64_t N,
                         int D, int gated, float drop_scale, int dtype, void* stream) {
  MURCL_REQUIRE(uv && wc && ds && dwc, "attn_score_bwd: null pointer");
  MURCL_REQUIRE(N >= 0 && D > 0, "attn_score_bwd: bad shape");
  MURCL_REQUIRE(dtype == MURCL_F32 || dtype == MURCL_BF16, "attn_score_bwd: bad dtype %d", dtype);
  if (N == 0) return MURCL_OK;
  const float q = drop_scale > 0.f ? drop_scale : 1.f;
  cudaStream_t st = as_stream(stream);
  const int ld = gated ? 2 * D : D;
  const bool fast = (D % 4 == 0) && D <= 512 && ((reinterpret_cast<uintptr_t>(uv) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(wc) & 15) == 0);
  if (!fast) {
    const int grid = ceil_div(N, 8);
    if (dtype == MURCL_F32) {
      if (gated) attn_score_bwd_generic_kernel<float, true><<<grid, 256, 0, st>>>((float*)uv, wc, ds, dwc, dbc, N, D, q);
      else attn_score_bwd_generic_kernel<float, false><<<grid, 256, 0, st>>>((float*)uv, wc, ds, dwc, dbc, N, D, q);
    } else {
      if (gated) attn_score_bwd_generic_kernel<__nv_bfloat16, true><<<grid, 256, 0, st>>>((__nv_bfloat16*)uv, wc, ds, dwc, dbc, N, D, q);
      else attn_score_bwd_generic_kernel<__nv_bfloat16, false><<<grid, 256, 0, st>>>((__nv_bfloat16*)uv, wc, ds, dwc, dbc, N, D, q);
    }
    int rc = check_launch("attn_score_bwd_generic_kernel");
    if (rc != MURCL_OK || dpre_colsum == nullptr) return rc;
    return murcl_colsum(uv, N, ld, dtype, dpre_colsum, stream);
  }
  int rows_per_cta = (int)((N + 4 * sm_count() - 1) / (4 * sm_count()));
  if (rows_per_cta < 64) rows_per_cta = 64;
  const int grid = ceil_div(N, rows_per_cta);
  const size_t smem = sizeof(float) * (size_t)(D + ld);
  if (dtype == MURCL_F32) {
    if (gated) attn_score_bwd_kernel<float, true><<<grid, 256, smem, st>>>((float*)uv, wc, ds, dwc, dbc, dpre_colsum, N, D, rows_per_cta, q);
    else attn_score_bwd_kernel<float, false><<<grid, 256, smem, st>>>((float*)uv, wc, ds, dwc, dbc, dpre_colsum, N, D, rows_per_cta, q);
  } else {
    if (gated) attn_score_bwd_kernel<__nv_bfloat16, true><<<grid, 256, smem, st>>>((__nv_bfloat16*)uv, wc, ds, dwc, dbc, dpre_colsum, N, D, rows_per_cta, q);
    else attn_score_bwd_kernel<__nv_bfloat16, false><<<grid, 256, smem, st>>>((__nv_bfloat16*)uv, wc, ds, dwc, dbc, dpre_colsum, N, D, rows_per_cta, q);
  }
  return check_launch("attn_score_bwd_kernel");
}

}  // extern "C"
