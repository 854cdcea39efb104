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

}  // namespace murcl

using namespace murcl;

extern "C" {

int64_t murcl_ntxent_workspace(int B, int d) {
  const int64_t R = 2 * (int64_t)B;
  return R * d /*zn*/ + 4 * R /*inv_norm, lse, row_loss, pad*/ + R * R /*gram / coefficients*/ + R * d /*C zn*/ +
         32 * R * d /*split-K scratch of the C zn product*/;
}

int murcl_ntxent_fwd_bwd(const float* z, int B, int d, float temperature, float* loss, float* dz, float* cos_pair,
                         float* workspace, void* stream) {
  MURCL_REQUIRE(z && loss && workspace, "ntxent: null pointer");
  MURCL_REQUIRE(B > 0 && d > 0 && temperature > 0.f, "ntxent: bad B=%d d=%d tau=%g", B, d, (double)temperature);
  MURCL_REQUIRE(B <= 16384, "ntxent: 2B=%d rows exceed the Gram-matrix scratch design", 2 * B);
  const int R = 2 * B;
  cudaStream_t st = as_stream(stream);
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

}  // extern "C"
