// Fused NT-Xent / InfoNCE loss with gradient.  Replaces utils/losses.py:24-41 (which materialises a
// [2B,2B,d] broadcast and indexes with a CPU mask) and the per-bag cosine reward of
// train_MuRCL.py:253,282.  Everything stays on one stream with no host round trip.
//
//   zn_a = z_a / max(|z_a|, 1e-8)                      s_ab = zn_a . zn_b / tau
//   L    = 1/(2B) sum_a [ LSE_{b != a} s_ab - s_{a,pos(a)} ],  pos(a) = a +/- B
//   dL/dzn_a = 1/(2B tau) sum_{b != a} [ e^{s_ab - lse_a} + e^{s_ab - lse_b} - 2 [b = pos(a)] ] zn_b
//   dL/dz_a  = (dzn_a - (dzn_a . zn_a) zn_a) / |z_a|      (|z_a| > eps)
//
// Under data parallelism 2B is the GLOBAL batch (2048 rows on 8 GPUs), so the two contractions are real GEMMs:
// the Gram matrix G = Zn Zn^T ([2B,2B], a few MB of scratch) and dZn = C Zn go through the exact-fp32 SIMT GEMM;
// the row log-sum-exp and the coefficient matrix C are coalesced row kernels over G.
#include "common.cuh"

namespace murcl {

int simt_linear_fwd(const void*, const void*, const float*, void*, int64_t, int, int, int, int, int, cudaStream_t,
                    float* ws, int64_t ws_floats);
int simt_linear_bwd_input(const void*, const void*, void*, int64_t, int, int, const void*, const float*, const float*,
                          const int32_t*, float, int, cudaStream_t, float* ws, int64_t ws_floats,
                          const unsigned long long* relu_bits);

constexpr float COS_EPS = 1e-8f;

__global__ void __launch_bounds__(256) ntx_normalize_kernel(const float* __restrict__ z, int rows, int d,
                                                            float* __restrict__ zn, float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float ss = 0.f;
  for (int j = lane; j < d; j += 32) {
    const float v = z[(int64_t)row * d + j];
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), COS_EPS);
  for (int j = lane; j < d; j += 32) zn[(int64_t)row * d + j] = z[(int64_t)row * d + j] * inv;
  if (lane == 0) inv_norm[row] = inv;
}

// One CTA per row a of the Gram matrix: masked log-sum-exp over b != a and the positive logit.
__global__ void __launch_bounds__(256) ntx_rows_kernel(const float* __restrict__ gram, int B, float inv_tau,
                                                       float* __restrict__ lse, float* __restrict__ row_loss,
                                                       float* __restrict__ cos_pair) {
  __shared__ float red[32];
  const int a = blockIdx.x, R = 2 * B;
  const int pos = a < B ? a + B : a - B;
  const float* g = gram + (int64_t)a * R;
  float m = -INFINITY;
  for (int b = threadIdx.x; b < R; b += blockDim.x)
    if (b != a) m = fmaxf(m, g[b] * inv_tau);
  m = block_max(m, red);
  float l = 0.f;
  for (int b = threadIdx.x; b < R; b += blockDim.x)
    if (b != a) l += expf(g[b] * inv_tau - m);
  l = block_sum(l, red);
  if (threadIdx.x == 0) {
    const float e = m + logf(l);
    lse[a] = e;
    row_loss[a] = e - g[pos] * inv_tau;
    if (cos_pair && a < B) cos_pair[a] = g[pos];
  }
}

__global__ void __launch_bounds__(256) ntx_loss_kernel(const float* __restrict__ row_loss, int R, float* __restrict__ loss) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < R; i += blockDim.x) s += row_loss[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) loss[0] = s / (float)R;
}

// In place: gram[a,b] -> coefficient c_ab (symmetric), 0 on the diagonal.
__global__ void __launch_bounds__(256) ntx_coef_kernel(float* __restrict__ gram, const float* __restrict__ lse, int B,
                                                       float inv_tau) {
  const int R = 2 * B;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)R * R) return;
  const int a = (int)(i / R), b = (int)(i % R);
  float c = 0.f;
  if (a != b) {
    const float s = gram[i] * inv_tau;
    const int pos = a < B ? a + B : a - B;
    c = expf(s - lse[a]) + expf(s - lse[b]) - (b == pos ? 2.f : 0.f);
  }
  gram[i] = c;
}

// Small-batch contractions (2B <= 1024, the single-GPU case): 32x32 output tiles, 256 threads, 2x2 outputs per
// thread, operands staged in shared memory.  Many small CTAs instead of a handful of 128x128 GEMM tiles.
__global__ void __launch_bounds__(256) ntx_gram_small_kernel(const float* __restrict__ zn, int R, int d, float* __restrict__ gram) {
  extern __shared__ float sm[];                       // [32][d+1] rows of tile a, [32][d+1] rows of tile b
  const int ld = d + 1;
  float* sa = sm;
  float* sb = sm + 32 * ld;
  const int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
  for (int i = threadIdx.x; i < 32 * d; i += blockDim.x) {
    const int r = i / d, c = i % d;
    sa[r * ld + c] = (a0 + r < R) ? zn[(int64_t)(a0 + r) * d + c] : 0.f;
    sb[r * ld + c] = (b0 + r < R) ? zn[(int64_t)(b0 + r) * d + c] : 0.f;
  }
  __syncthreads();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;         // 16 x 16 threads, 2 x 2 outputs each
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k = 0; k < d; ++k) {
    const float x0 = sa[(ty * 2) * ld + k], x1 = sa[(ty * 2 + 1) * ld + k];
    const float y0 = sb[(tx * 2) * ld + k], y1 = sb[(tx * 2 + 1) * ld + k];
    acc[0][0] = fmaf(x0, y0, acc[0][0]); acc[0][1] = fmaf(x0, y1, acc[0][1]);
    acc[1][0] = fmaf(x1, y0, acc[1][0]); acc[1][1] = fmaf(x1, y1, acc[1][1]);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int a = a0 + ty * 2 + i, b = b0 + tx * 2 + j;
      if (a < R && b < R) gram[(int64_t)a * R + b] = acc[i][j];
    }
}

// cz[a, j] = sum_b C[a, b] * zn[b, j]
__global__ void __launch_bounds__(256) ntx_cz_small_kernel(const float* __restrict__ coef, const float* __restrict__ zn, int R,
                                                           int d, float* __restrict__ cz) {
  __shared__ float sc[32][33];                        // C tile [a][b]
  __shared__ float sz[32][33];                        // zn tile [b][j]
  const int a0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int b0 = 0; b0 < R; b0 += 32) {
    for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
      const int r = i >> 5, c = i & 31;
      sc[r][c] = (a0 + r < R && b0 + c < R) ? coef[(int64_t)(a0 + r) * R + b0 + c] : 0.f;
      sz[r][c] = (b0 + r < R && j0 + c < d) ? zn[(int64_t)(b0 + r) * d + j0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float x0 = sc[ty * 2][k], x1 = sc[ty * 2 + 1][k];
      const float y0 = sz[k][tx * 2], y1 = sz[k][tx * 2 + 1];
      acc[0][0] = fmaf(x0, y0, acc[0][0]); acc[0][1] = fmaf(x0, y1, acc[0][1]);
      acc[1][0] = fmaf(x1, y0, acc[1][0]); acc[1][1] = fmaf(x1, y1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int a = a0 + ty * 2 + i, c = j0 + tx * 2 + j;
      if (a < R && c < d) cz[(int64_t)a * d + c] = acc[i][j];
    }
}

// dz_a from g_a = (C zn)_a * inv_tau / R through the normalisation; one warp per row.
__global__ void __launch_bounds__(256) ntx_finish_kernel(const float* __restrict__ cz, const float* __restrict__ zn,
                                                         const float* __restrict__ inv_norm, int rows, int d, float scale,
                                                         float* __restrict__ dz) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float proj = 0.f;
  for (int j = lane; j < d; j += 32) proj = fmaf(cz[(int64_t)row * d + j] * scale, zn[(int64_t)row * d + j], proj);
  proj = warp_sum(proj);
  const float inv = inv_norm[row];
  const bool clamped = inv >= 1.f / COS_EPS;       // |z| <= eps: zn = z / eps is linear in z, no projection term
  for (int j = lane; j < d; j += 32) {
    const float g = cz[(int64_t)row * d + j] * scale;
    dz[(int64_t)row * d + j] = clamped ? g * inv : (g - proj * zn[(int64_t)row * d + j]) * inv;
  }
}


// ================================================================================================================
// Two-kernel form (d <= 256, d % 4 == 0: the projection widths in use).  Nothing of size [2B, 2B] touches memory.
//
//   ntx_lse_kernel   grid (row blocks of 16 rows a, splits of the b range): streams its rows b in tiles of 64 through
//                    shared memory, normalises them on the fly, keeps per-thread online (max, sum-exp) of s_ab over
//                    b != a, records the positive logit; the last split of a row block merges the partials and writes
//                    1/|z_a|, lse_a, the row's loss term and cos(z_i, z_j); the last row block to finish sums the loss
//                    terms in a fixed order (deterministic).
//   ntx_grad_kernel  grid (row blocks of 16 rows of the caller's SLAB, splits of the b range): recomputes the 16 x 64
//                    score tile, turns it into the coefficients e^{s-lse_a} + e^{s-lse_b} - 2[b = pos(a)], multiplies
//                    them with the normalised rows of the tile, adds the partial dzn rows to a scratch accumulator; the
//                    last split of a row block takes the rows through the normalisation backward and writes dz.
//
// The slab (samples [b0, b0 + nb) of both views) is what data parallelism needs: every rank evaluates the loss over the
// global batch (lse of all rows: ntx_lse_kernel, 2B x 2B x d FMAs) but only the gradient rows of its own samples
// (2 nb x 2B x 2d FMAs) - no second collective, no redundant gradient work (utils/losses.py:24-41 differentiated).
constexpr int NTX_RA2 = 16, NTX_TB = 64, NTX_MAXD = 256;

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
  return acc;
}

// grid (row blocks of 16 rows, splits of the b range).  Every CTA merges its tiles into one partial (max, sum-exp) per
// row; the last split of a row block (ticket) merges the partials into lse / loss term / cosine; the last row block
// to finish sums the loss terms in index order (deterministic).
__global__ void __launch_bounds__(256) ntx_lse_kernel(const float* __restrict__ z, int B, int d, float inv_tau, int sb0, int nb,
                                                      int tiles_per_split, float* __restrict__ inv_norm, float* __restrict__ lse,
                                                      float* __restrict__ row_loss, float* __restrict__ cos_pair,
                                                      float* __restrict__ part /* [row blocks*16][splits][2] */,
                                                      float* __restrict__ pos_s /* [R] */, float* __restrict__ loss,
                                                      unsigned int* __restrict__ tickets /* [0]: global, [1 + rb]: per row block */) {
  extern __shared__ __align__(16) float sm[];
  const int R = 2 * B, ldb = d + 4;
  float* za = sm;                               // [16][d]   normalised rows a
  float* zb = za + NTX_RA2 * d;                 // [64][d+4] raw rows b of the tile
  float* invb = zb + NTX_TB * ldb;              // [64]
  float* red_m = invb + NTX_TB;                 // [16][64]
  float* red_l = red_m + NTX_RA2 * NTX_TB;      // [16][64]
  float* inva = red_l + NTX_RA2 * NTX_TB;       // [16]
  __shared__ float red[32];
  __shared__ bool is_last;
  __shared__ int a_of[NTX_RA2];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int a0 = blockIdx.x * NTX_RA2;                       // first LOCAL row of the block: the rows are the slab's
  const int n_splits = gridDim.y;
  if (t < NTX_RA2) {                                         // samples [sb0, sb0 + nb) of both views (nb == B: every row)
    const int r = a0 + t;
    a_of[t] = r < nb ? sb0 + r : (r < 2 * nb ? B + sb0 + (r - nb) : -1);
  }
  __syncthreads();
  for (int rr = w; rr < NTX_RA2; rr += 8) {                  // rows a: a warp normalises rows w, w + 8
    const int a = a_of[rr];
    float ss = 0.f;
    if (a >= 0)
      for (int j = lane; j < d; j += 32) { const float v = z[(int64_t)a * d + j]; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), COS_EPS);
    for (int j = lane; j < d; j += 32) za[rr * d + j] = (a >= 0) ? z[(int64_t)a * d + j] * inv : 0.f;
    if (lane == 0) inva[rr] = inv;
  }
  const int bl = t & 63, ag = t >> 6;                       // thread: column bl of the tile, rows 4ag .. 4ag + 3
  float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, l4[4] = {0.f, 0.f, 0.f, 0.f};
  const int n_tiles = (R + NTX_TB - 1) / NTX_TB;
  const int tile0 = blockIdx.y * tiles_per_split;
  for (int tile = tile0; tile < tile0 + tiles_per_split && tile < n_tiles; ++tile) {
    const int b0 = tile * NTX_TB;
    __syncthreads();
    for (int i = t; i < NTX_TB * (d >> 2); i += 256) {
      const int r = i / (d >> 2), c = i % (d >> 2);
      const float4 v = (b0 + r < R) ? *reinterpret_cast<const float4*>(z + (int64_t)(b0 + r) * d + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(zb + r * ldb + 4 * c) = v;
    }
    __syncthreads();
    for (int r = w; r < NTX_TB; r += 8) {                   // norms of the tile's rows
      float ss = 0.f;
      for (int j = lane; j < d; j += 32) { const float v = zb[r * ldb + j]; ss = fmaf(v, v, ss); }
      ss = warp_sum(ss);
      if (lane == 0) invb[r] = 1.f / fmaxf(sqrtf(ss), COS_EPS);
    }
    __syncthreads();
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* rb = zb + bl * ldb;
    const float* ra = za + (4 * ag) * d;
    for (int k = 0; k < d; k += 4) {
      const float4 vb = *reinterpret_cast<const float4*>(rb + k);
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = dot4(*reinterpret_cast<const float4*>(ra + i * d + k), vb, acc[i]);
    }
    const int b = b0 + bl;
    if (b < R) {
      const float sc = invb[bl] * inv_tau;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int a = a_of[4 * ag + i];
        const float sv = acc[i] * sc;
        if (b != a) { const float mn = fmaxf(m4[i], sv); l4[i] = l4[i] * expf(m4[i] - mn) + expf(sv - mn); m4[i] = mn; }
        if (a >= 0 && b == (a < B ? a + B : a - B)) pos_s[a] = sv;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    red_m[(4 * ag + i) * NTX_TB + bl] = m4[i];
    red_l[(4 * ag + i) * NTX_TB + bl] = l4[i];
  }
  __syncthreads();
  for (int rr = w; rr < NTX_RA2; rr += 8) {                  // a warp merges the 64 per-thread partials of a row
    const float ma = red_m[rr * NTX_TB + lane], mb = red_m[rr * NTX_TB + 32 + lane];
    const float la = red_l[rr * NTX_TB + lane], lb = red_l[rr * NTX_TB + 32 + lane];
    float m = warp_max(fmaxf(ma, mb));
    float l = (la > 0.f ? la * expf(ma - m) : 0.f) + (lb > 0.f ? lb * expf(mb - m) : 0.f);
    l = warp_sum(l);
    if (lane == 0) {
      float* pp = part + ((int64_t)(a0 + rr) * n_splits + blockIdx.y) * 2;
      pp[0] = m;
      pp[1] = l;
    }
  }
  __threadfence();
  __syncthreads();
  if (t == 0) is_last = atomicAdd(tickets + 1 + blockIdx.x, 1u) == (unsigned)n_splits - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int rr = w; rr < NTX_RA2; rr += 8) {                  // last split of the row block: merge the splits' partials
    const int a = a_of[rr];
    if (a < 0) continue;
    const int64_t lr = a0 + rr;                              // the partials are indexed by the LOCAL row
    float m = -INFINITY;
    for (int sidx = lane; sidx < n_splits; sidx += 32) m = fmaxf(m, __ldcg(part + (lr * n_splits + sidx) * 2));
    m = warp_max(m);
    float l = 0.f;
    for (int sidx = lane; sidx < n_splits; sidx += 32) {
      const float pm = __ldcg(part + (lr * n_splits + sidx) * 2), pl = __ldcg(part + (lr * n_splits + sidx) * 2 + 1);
      if (pl > 0.f) l += pl * expf(pm - m);
    }
    l = warp_sum(l);
    if (lane == 0) {
      const float e = m + logf(l), sp = __ldcg(pos_s + a);
      lse[a] = e;
      inv_norm[a] = inva[rr];
      row_loss[a] = e - sp;
      if (cos_pair && a < B) cos_pair[a] = sp / inv_tau;
    }
  }
  __threadfence();
  __syncthreads();
  if (t == 0) is_last = atomicAdd(tickets, 1u) == gridDim.x - 1;
  __syncthreads();
  if (is_last) {                                             // the last row block sums the loss terms in index order
    __threadfence();
    float sacc = 0.f;                                        // nb < B: this slab's share of the loss (the caller adds the shares)
    for (int r = t; r < 2 * nb; r += 256) sacc += __ldcg(row_loss + (r < nb ? sb0 + r : B + sb0 + (r - nb)));
    sacc = block_sum(sacc, red);
    if (t == 0) loss[0] = sacc / (float)R;
  }
}

// NC = d / 128 rounded up: float4 column chunks per thread in the accumulation phase.
template <int NC>
__global__ void __launch_bounds__(256) ntx_grad_kernel(const float* __restrict__ z, int B, int d, float inv_tau, int sb0, int nb,
                                                       int tiles_per_split, const float* __restrict__ inv_norm,
                                                       const float* __restrict__ lse, float* __restrict__ acc_ws,
                                                       unsigned int* __restrict__ tickets, float* __restrict__ dz) {
  extern __shared__ __align__(16) float sm[];
  const int R = 2 * B, ldb = d + 4, ldg = NTX_TB + 1;
  float* za = sm;                               // [16][d] normalised rows a
  float* zb = za + NTX_RA2 * d;                 // [64][d+4] normalised rows b
  float* G = zb + NTX_TB * ldb;                 // [16][65]
  float* lse_b = G + NTX_RA2 * ldg;             // [64]
  float* lse_a = lse_b + NTX_TB;                // [16]
  __shared__ int a_of[NTX_RA2];
  __shared__ bool is_last;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int r0 = blockIdx.x * NTX_RA2;                       // first local row of this block (slab rows: 2 * nb)
  if (t < NTX_RA2) {
    const int r = r0 + t;
    a_of[t] = r < nb ? sb0 + r : (r < 2 * nb ? B + sb0 + (r - nb) : -1);
  }
  __syncthreads();
  for (int i = t; i < NTX_RA2 * (d >> 2); i += 256) {
    const int r = i / (d >> 2), c = i % (d >> 2);
    const int a = a_of[r];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a >= 0) {
      v = *reinterpret_cast<const float4*>(z + (int64_t)a * d + 4 * c);
      const float inv = inv_norm[a];
      v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
    }
    *reinterpret_cast<float4*>(za + r * d + 4 * c) = v;
  }
  if (t < NTX_RA2) lse_a[t] = a_of[t] >= 0 ? lse[a_of[t]] : 0.f;
  const int bl = t & 63, ag = t >> 6;                        // phase 1: column bl, rows 4ag .. 4ag + 3
  const int j4 = t & 31, g2 = t >> 5;                        // phase 2: columns 4 j4 (+128 c), rows 2 g2, 2 g2 + 1
  float acc[2][NC][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][c][e] = 0.f;
  const int tile0 = blockIdx.y * tiles_per_split;
  const int n_tiles = (R + NTX_TB - 1) / NTX_TB;
  for (int tile = tile0; tile < tile0 + tiles_per_split && tile < n_tiles; ++tile) {
    const int b0 = tile * NTX_TB;
    __syncthreads();
    for (int i = t; i < NTX_TB * (d >> 2); i += 256) {
      const int r = i / (d >> 2), c = i % (d >> 2);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b0 + r < R) {
        v = *reinterpret_cast<const float4*>(z + (int64_t)(b0 + r) * d + 4 * c);
        const float inv = inv_norm[b0 + r];
        v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
      }
      *reinterpret_cast<float4*>(zb + r * ldb + 4 * c) = v;
    }
    if (t < NTX_TB) lse_b[t] = (b0 + t < R) ? lse[b0 + t] : 0.f;
    __syncthreads();
    {
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
      const float* rb = zb + bl * ldb;
      const float* ra = za + (4 * ag) * d;
      for (int k = 0; k < d; k += 4) {
        const float4 vb = *reinterpret_cast<const float4*>(rb + k);
#pragma unroll
        for (int i = 0; i < 4; ++i) s4[i] = dot4(*reinterpret_cast<const float4*>(ra + i * d + k), vb, s4[i]);
      }
      const int b = b0 + bl;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int a = a_of[4 * ag + i];
        float c = 0.f;
        if (a >= 0 && b < R && b != a) {
          const float s = s4[i] * inv_tau;
          const int pos = a < B ? a + B : a - B;
          c = expf(s - lse_a[4 * ag + i]) + expf(s - lse_b[bl]) - (b == pos ? 2.f : 0.f);
        }
        G[(4 * ag + i) * ldg + bl] = c;
      }
    }
    __syncthreads();
    for (int b = 0; b < NTX_TB; ++b) {
      const float c0 = G[(2 * g2) * ldg + b], c1 = G[(2 * g2 + 1) * ldg + b];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        if (4 * j4 + 128 * c < d) {
          const float4 v = *reinterpret_cast<const float4*>(zb + b * ldb + 4 * j4 + 128 * c);
          acc[0][c][0] = fmaf(c0, v.x, acc[0][c][0]); acc[0][c][1] = fmaf(c0, v.y, acc[0][c][1]);
          acc[0][c][2] = fmaf(c0, v.z, acc[0][c][2]); acc[0][c][3] = fmaf(c0, v.w, acc[0][c][3]);
          acc[1][c][0] = fmaf(c1, v.x, acc[1][c][0]); acc[1][c][1] = fmaf(c1, v.y, acc[1][c][1]);
          acc[1][c][2] = fmaf(c1, v.z, acc[1][c][2]); acc[1][c][3] = fmaf(c1, v.w, acc[1][c][3]);
        }
      }
    }
  }
  // partial dzn rows -> scratch accumulator (rows of the slab), then the last split of the row block finishes them
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = r0 + 2 * g2 + i;
    if (r < 2 * nb) {
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (4 * j4 + 128 * c < d)
#pragma unroll
          for (int e = 0; e < 4; ++e) atomicAdd(acc_ws + (int64_t)r * d + 4 * j4 + 128 * c + e, acc[i][c][e]);
    }
  }
  __threadfence();
  __syncthreads();
  if (t == 0) is_last = atomicAdd(tickets + blockIdx.x, 1u) == gridDim.y - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const float scale = inv_tau / (float)R;
  for (int rr = w; rr < NTX_RA2; rr += 8) {                  // one warp per row: projection through the normalisation
    const int a = a_of[rr];
    if (a < 0) continue;
    const int r = r0 + rr;
    float proj = 0.f;
    for (int j = lane; j < d; j += 32) proj = fmaf(__ldcg(acc_ws + (int64_t)r * d + j) * scale, za[rr * d + j], proj);
    proj = warp_sum(proj);
    const float inv = inv_norm[a];
    const bool clamped = inv >= 1.f / COS_EPS;               // |z| <= eps: zn = z / eps is linear in z, no projection term
    for (int j = lane; j < d; j += 32) {
      const float g = __ldcg(acc_ws + (int64_t)r * d + j) * scale;
      dz[(int64_t)a * d + j] = clamped ? g * inv : (g - proj * za[rr * d + j]) * inv;
    }
  }
}

static bool ntx_fused_ok(int d) { return d % 4 == 0 && d <= NTX_MAXD; }

}  // namespace murcl

using namespace murcl;

extern "C" {

int64_t murcl_ntxent_workspace(int B, int d) {
  const int64_t R = 2 * (int64_t)B;
  const int64_t legacy = R * d /*zn*/ + 4 * R /*inv_norm, lse, row_loss, pad*/ + R * R /*gram / coefficients*/ + R * d /*C zn*/ +
                         32 * R * d /*split-K scratch of the C zn product*/;
  const int64_t fused = 4 * R /*inv_norm, lse, row_loss, pad*/ + R /*pos_s*/ + (R + 16) * ((R + 63) / 64) * 2 /*(max, sum-exp) partials*/ +
                        2 * (R / 16 + 2) + 64 /*tickets, alignment*/ + R * d /*dzn accumulator*/;
  return ntx_fused_ok(d) ? fused : legacy;
}

static int ntxent_legacy(const float* z, int B, int d, float temperature, float* loss, float* dz, float* cos_pair,
                         float* workspace, cudaStream_t st) {
  const int R = 2 * B;
  float* zn = workspace;
  float* inv_norm = zn + (int64_t)R * d;
  float* lse = inv_norm + R;
  float* row_loss = lse + R;
  float* gram = row_loss + 2 * R;
  float* cz = gram + (int64_t)R * R;
  float* scratch = cz + (int64_t)R * d;
  const int64_t scratch_floats = 32 * (int64_t)R * d;
  const float inv_tau = 1.f / temperature;
  ntx_normalize_kernel<<<ceil_div(R, 8), 256, 0, st>>>(z, R, d, zn, inv_norm);
  int rc = check_launch("ntx_normalize_kernel");
  if (rc != MURCL_OK) return rc;
  // Gram matrix: [R,d] x [R,d]^T -> [R,R]
  const bool small = R <= 1024 && (size_t)(64 * (d + 1)) * sizeof(float) <= 48 * 1024;
  if (small) {
    ntx_gram_small_kernel<<<dim3(ceil_div(R, 32), ceil_div(R, 32)), 256, sizeof(float) * 64 * (d + 1), st>>>(zn, R, d, gram);
    rc = check_launch("ntx_gram_small_kernel");
  } else {
    rc = simt_linear_fwd(zn, zn, nullptr, gram, R, R, d, MURCL_ACT_NONE, MURCL_F32, MURCL_F32, st, scratch, scratch_floats);
  }
  if (rc != MURCL_OK) return rc;
  ntx_rows_kernel<<<R, 256, 0, st>>>(gram, B, inv_tau, lse, row_loss, cos_pair);
  rc = check_launch("ntx_rows_kernel");
  if (rc != MURCL_OK) return rc;
  ntx_loss_kernel<<<1, 256, 0, st>>>(row_loss, R, loss);
  rc = check_launch("ntx_loss_kernel");
  if (rc != MURCL_OK || dz == nullptr) return rc;
  ntx_coef_kernel<<<ceil_div((int64_t)R * R, 256), 256, 0, st>>>(gram, lse, B, inv_tau);
  rc = check_launch("ntx_coef_kernel");
  if (rc != MURCL_OK) return rc;
  // C zn: [R,R] x [R,d] -> [R,d]   (dx = dy . w with dy = C, w = zn)
  if (small) {
    ntx_cz_small_kernel<<<dim3(ceil_div(d, 32), ceil_div(R, 32)), 256, 0, st>>>(gram, zn, R, d, cz);
    rc = check_launch("ntx_cz_small_kernel");
  } else {
    rc = simt_linear_bwd_input(gram, zn, cz, R, R, d, nullptr, nullptr, nullptr, nullptr, 1.f, MURCL_F32, st, scratch, scratch_floats, nullptr);
  }
  if (rc != MURCL_OK) return rc;
  ntx_finish_kernel<<<ceil_div(R, 8), 256, 0, st>>>(cz, zn, inv_norm, R, d, inv_tau / (float)R, dz);
  return check_launch("ntx_finish_kernel");
}

namespace {

struct NtxPlan {
  int R, n_tiles;
  float inv_tau;
  static void split(int blocks, int n_tiles, int& splits, int& tps) {     // ~2 CTAs per SM, every split >= 1 tile
    splits = ceil_div(2 * sm_count(), blocks > 0 ? blocks : 1);
    if (splits > n_tiles) splits = n_tiles;
    if (splits < 1) splits = 1;
    tps = ceil_div(n_tiles, splits);
    splits = ceil_div(n_tiles, tps);
  }
};

int ntx_configure() {
  static PerDeviceOnce configured;
  if (const int slot = configured.pending(); slot >= 0) {
    MURCL_CUDA(cudaFuncSetAttribute(ntx_lse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    MURCL_CUDA(cudaFuncSetAttribute(ntx_grad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    MURCL_CUDA(cudaFuncSetAttribute(ntx_grad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured.mark(slot);
  }
  return MURCL_OK;
}

// Log-sum-exp pass over the rows of the samples [b0, b0 + nb) of both views (nb == B: all rows): inv_norm / lse / row_loss
// of THOSE rows (global row index), their loss share, cos of those samples.  scratch: pos_s [R] | part | tickets.
int ntx_lse_pass(const float* z, int B, int d, float inv_tau, int b0, int nb, float* inv_norm, float* lse, float* row_loss,
                 float* loss, float* cos_pair, float* scratch, cudaStream_t st) {
  const int R = 2 * B, n_tiles = ceil_div(R, NTX_TB);
  const int rb = ceil_div(2 * nb, NTX_RA2);
  int splits, tps;
  NtxPlan::split(rb, n_tiles, splits, tps);
  float* pos_s = scratch;
  float* part = pos_s + R;
  unsigned int* tickets = reinterpret_cast<unsigned int*>(part + (int64_t)rb * NTX_RA2 * n_tiles * 2);
  const int64_t n_tickets = 2 + (int64_t)rb;
  MURCL_CUDA(cudaMemsetAsync(tickets, 0, sizeof(unsigned int) * (size_t)n_tickets, st));
  int rc = ntx_configure();
  if (rc != MURCL_OK) return rc;
  const size_t smem = sizeof(float) * (size_t)(NTX_RA2 * d + NTX_TB * (d + 4) + NTX_TB + 2 * NTX_RA2 * NTX_TB + NTX_RA2);
  ntx_lse_kernel<<<dim3(rb, splits), 256, smem, st>>>(z, B, d, inv_tau, b0, nb, tps, inv_norm, lse, row_loss, cos_pair, part, pos_s,
                                                     loss, tickets);
  return check_launch("ntx_lse_kernel");
}

// Gradient rows of the samples [b0, b0 + nb) of both views, given inv_norm / lse of ALL rows.  scratch: tickets | accumulator.
int ntx_grad_pass(const float* z, int B, int d, float inv_tau, int b0, int nb, const float* inv_norm, const float* lse, float* dz,
                  float* scratch, cudaStream_t st) {
  const int R = 2 * B, n_tiles = ceil_div(R, NTX_TB);
  const int row_blocks = ceil_div(2 * nb, NTX_RA2);
  int splits, tps;
  NtxPlan::split(row_blocks, n_tiles, splits, tps);
  unsigned int* tickets = reinterpret_cast<unsigned int*>(scratch);
  const int64_t n_tickets = ((int64_t)row_blocks + 3) / 4 * 4;
  float* acc_ws = scratch + n_tickets;                                    // [2 nb, d], 16-byte aligned
  // one memset node clears the tickets and, right behind them, the gradient accumulator
  MURCL_CUDA(cudaMemsetAsync(tickets, 0, sizeof(float) * ((size_t)n_tickets + (size_t)2 * nb * d), st));
  int rc = ntx_configure();
  if (rc != MURCL_OK) return rc;
  const size_t smem = sizeof(float) * (size_t)(NTX_RA2 * d + NTX_TB * (d + 4) + NTX_RA2 * (NTX_TB + 1) + NTX_TB + NTX_RA2);
  const dim3 grid(row_blocks, splits);
  if (d <= 128)
    ntx_grad_kernel<1><<<grid, 256, smem, st>>>(z, B, d, inv_tau, b0, nb, tps, inv_norm, lse, acc_ws, tickets, dz);
  else
    ntx_grad_kernel<2><<<grid, 256, smem, st>>>(z, B, d, inv_tau, b0, nb, tps, inv_norm, lse, acc_ws, tickets, dz);
  return check_launch("ntx_grad_kernel");
}

int64_t ntx_lse_scratch(int R, int rows) {            // pos_s | part | tickets (floats)
  const int64_t rb = (rows + NTX_RA2 - 1) / NTX_RA2;
  return R + rb * NTX_RA2 * ((R + NTX_TB - 1) / NTX_TB) * 2 + rb + 2 + 8;
}
int64_t ntx_grad_scratch(int rows, int d) {           // tickets | accumulator (floats)
  const int64_t rb = (rows + NTX_RA2 - 1) / NTX_RA2;
  return (rb + 3) / 4 * 4 + (int64_t)rows * d + 8;
}

}  // namespace

int murcl_ntxent_fwd_bwd_slab(const float* z, int B, int d, float temperature, int b0, int nb, float* loss, float* dz,
                              float* cos_pair, float* workspace, void* stream) {
  MURCL_REQUIRE(z && loss && workspace, "ntxent: null pointer");
  MURCL_REQUIRE(B > 0 && d > 0 && temperature > 0.f, "ntxent: bad B=%d d=%d tau=%g", B, d, (double)temperature);
  MURCL_REQUIRE(B <= 16384, "ntxent: 2B=%d rows exceed the design", 2 * B);
  MURCL_REQUIRE(b0 >= 0 && nb >= 0 && b0 + nb <= B, "ntxent: gradient slab [%d, %d) outside the batch of %d", b0, b0 + nb, B);
  cudaStream_t st = as_stream(stream);
  if (!ntx_fused_ok(d) || (reinterpret_cast<uintptr_t>(z) & 15)) {
    MURCL_REQUIRE(b0 == 0 && nb == B, "ntxent: the gradient slab needs d %% 4 == 0, d <= %d and a 16-byte aligned z", NTX_MAXD);
    return ntxent_legacy(z, B, d, temperature, loss, dz, cos_pair, workspace, st);
  }
  const int R = 2 * B;
  const float inv_tau = 1.f / temperature;
  // workspace: inv_norm [R] | lse [R] | row_loss [R] | pad [R] | lse-pass scratch | gradient-pass scratch
  float* inv_norm = workspace;
  float* lse = inv_norm + R;
  float* row_loss = lse + R;
  float* lse_scratch = row_loss + 2 * R;
  float* grad_scratch = lse_scratch + (ntx_lse_scratch(R, R) + 3) / 4 * 4;
  int rc = ntx_lse_pass(z, B, d, inv_tau, 0, B, inv_norm, lse, row_loss, loss, cos_pair, lse_scratch, st);   // all rows
  if (rc != MURCL_OK || dz == nullptr || nb == 0) return rc;
  return ntx_grad_pass(z, B, d, inv_tau, b0, nb, inv_norm, lse, dz, grad_scratch, st);
}

int murcl_ntxent_lse_slab(const float* z, int B, int d, float temperature, int b0, int nb, float* inv_norm, float* lse,
                          float* loss_share, float* cos_pair, float* workspace, void* stream) {
  MURCL_REQUIRE(z && inv_norm && lse && loss_share && workspace, "ntxent_lse_slab: null pointer");
  MURCL_REQUIRE(B > 0 && B <= 16384 && d > 0 && temperature > 0.f, "ntxent_lse_slab: bad B=%d d=%d tau=%g", B, d, (double)temperature);
  MURCL_REQUIRE(b0 >= 0 && nb > 0 && b0 + nb <= B, "ntxent_lse_slab: slab [%d, %d) outside the batch of %d", b0, b0 + nb, B);
  MURCL_REQUIRE(ntx_fused_ok(d) && !(reinterpret_cast<uintptr_t>(z) & 15), "ntxent_lse_slab: needs d %% 4 == 0, d <= %d, 16-byte aligned z",
                NTX_MAXD);
  const int R = 2 * B;
  float* row_loss = workspace;                                             // [R] (only the slab's rows are written and read)
  return ntx_lse_pass(z, B, d, 1.f / temperature, b0, nb, inv_norm, lse, row_loss, loss_share, cos_pair, workspace + R, as_stream(stream));
}

int murcl_ntxent_grad_slab(const float* z, int B, int d, float temperature, int b0, int nb, const float* inv_norm, const float* lse,
                           float* dz, float* workspace, void* stream) {
  MURCL_REQUIRE(z && inv_norm && lse && dz && workspace, "ntxent_grad_slab: null pointer");
  MURCL_REQUIRE(B > 0 && B <= 16384 && d > 0 && temperature > 0.f, "ntxent_grad_slab: bad B=%d d=%d tau=%g", B, d, (double)temperature);
  MURCL_REQUIRE(b0 >= 0 && nb > 0 && b0 + nb <= B, "ntxent_grad_slab: slab [%d, %d) outside the batch of %d", b0, b0 + nb, B);
  MURCL_REQUIRE(ntx_fused_ok(d) && !(reinterpret_cast<uintptr_t>(z) & 15) && !(reinterpret_cast<uintptr_t>(workspace) & 15),
                "ntxent_grad_slab: needs d %% 4 == 0, d <= %d, 16-byte aligned z and workspace", NTX_MAXD);
  return ntx_grad_pass(z, B, d, 1.f / temperature, b0, nb, inv_norm, lse, dz, workspace, as_stream(stream));
}

int64_t murcl_ntxent_slab_workspace(int B, int d, int nb) {
  const int64_t R = 2 * (int64_t)B;
  const int64_t a = R + ntx_lse_scratch((int)R, 2 * nb), b = ntx_grad_scratch(2 * nb, d);
  return ((a > b ? a : b) + 3) / 4 * 4;
}

int murcl_ntxent_fwd_bwd(const float* z, int B, int d, float temperature, float* loss, float* dz, float* cos_pair,
                         float* workspace, void* stream) {
  return murcl_ntxent_fwd_bwd_slab(z, B, d, temperature, 0, B, loss, dz, cos_pair, workspace, stream);
}

}  // extern "C"
