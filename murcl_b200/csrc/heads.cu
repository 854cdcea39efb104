// Small recurrent-head kernels (GRU cell of Full_layer / the PPO actor, rlmil.py:47,76-90,199,213-220)
// and elementwise helpers (cast, column sums, ReLU backward, CSR row -> bag map).
#include "common.cuh"

namespace murcl {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// h' = (1 - z) * n + z * h,  r = s(gi_r + gh_r), z = s(gi_z + gh_z), n = tanh(gi_n + r * gh_n)
__global__ void __launch_bounds__(256) gru_cell_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                                                           const float* __restrict__ h_prev, float* __restrict__ h_new,
                                                           float* __restrict__ gates, int B, int H) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * H) return;
  const int b = (int)(i / H), j = (int)(i % H);
  const float* a = gi + (int64_t)b * 3 * H;
  const float* c = gh + (int64_t)b * 3 * H;
  const float r = sigmoidf_(a[j] + c[j]);
  const float z = sigmoidf_(a[H + j] + c[H + j]);
  const float n = tanhf(a[2 * H + j] + r * c[2 * H + j]);
  const float hp = h_prev ? h_prev[i] : 0.f;
  h_new[i] = (1.f - z) * n + z * hp;
  if (gates) {
    float* g = gates + (int64_t)b * 3 * H;
    g[j] = r;
    g[H + j] = z;
    g[2 * H + j] = n;
  }
}

__global__ void __launch_bounds__(256) gru_cell_bwd_kernel(const float* __restrict__ dh_new, const float* __restrict__ gates,
                                                           const float* __restrict__ gh, const float* __restrict__ h_prev,
                                                           float* __restrict__ dgi, float* __restrict__ dgh,
                                                           float* __restrict__ dh_prev, int B, int H) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * H) return;
  const int b = (int)(i / H), j = (int)(i % H);
  const float* g = gates + (int64_t)b * 3 * H;
  const float r = g[j], z = g[H + j], n = g[2 * H + j];
  const float hn = gh[(int64_t)b * 3 * H + 2 * H + j];
  const float hp = h_prev ? h_prev[i] : 0.f;
  const float dh = dh_new[i];
  const float dn = dh * (1.f - z);
  const float dz = dh * (hp - n);
  const float dan = dn * (1.f - n * n);
  const float dar = dan * hn * r * (1.f - r);
  const float daz = dz * z * (1.f - z);
  float* o = dgi + (int64_t)b * 3 * H;
  float* q = dgh + (int64_t)b * 3 * H;
  o[j] = dar;
  o[H + j] = daz;
  o[2 * H + j] = dan;
  q[j] = dar;
  q[H + j] = daz;
  q[2 * H + j] = dan * r;
  if (dh_prev) dh_prev[i] = dh * z;
}

// ---- recurrent-head tape (murcl_b200/headtape.py): the same cell, laid out for a BATCHED backward over all the patch-steps ----
// Forward: also writes h' in the GEMM storage type (the operand of the next step's W_hh product and of the output layer), so
// no cast launch sits between the cell and the GEMMs.
template <typename TS>
__global__ void __launch_bounds__(256) gru_cell_fwd_tape_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                                                                const float* __restrict__ h_prev, float* __restrict__ h_new,
                                                                TS* __restrict__ h_new_s, TS* __restrict__ h_new_s2,
                                                                float* __restrict__ gates, int B, int H) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * H) return;
  const int b = (int)(i / H), j = (int)(i % H);
  const float* a = gi + (int64_t)b * 3 * H;
  const float* c = gh + (int64_t)b * 3 * H;
  const float r = sigmoidf_(a[j] + c[j]);
  const float z = sigmoidf_(a[H + j] + c[H + j]);
  const float n = tanhf(a[2 * H + j] + r * c[2 * H + j]);
  const float hp = h_prev ? h_prev[i] : 0.f;
  const float hn = (1.f - z) * n + z * hp;
  h_new[i] = hn;
  if (h_new_s) Store<TS>::store(h_new_s + i, hn);
  if (h_new_s2) Store<TS>::store(h_new_s2 + i, hn);      // the next call's h_prev slot of the batched W_hh weight gradient
  float* g = gates + (int64_t)b * 3 * H;
  g[j] = r;
  g[H + j] = z;
  g[2 * H + j] = n;
}

// Backward: dh = dh_a + dh_b (storage type: the output layer's and the next cell's W_hh input gradients) + dh_c (fp32: the
// next cell's direct z * dh term), any of them NULL; the gate gradients leave in the storage type, ready to be the operands
// of the batched weight-gradient GEMMs and of the W_hh / W_ih input-gradient GEMMs.
template <typename TS>
__global__ void __launch_bounds__(256) gru_cell_bwd_tape_kernel(const TS* __restrict__ dh_a, const TS* __restrict__ dh_b,
                                                                const float* __restrict__ dh_c, const float* __restrict__ gates,
                                                                const float* __restrict__ gh, const float* __restrict__ h_prev,
                                                                TS* __restrict__ dgi, TS* __restrict__ dgh,
                                                                float* __restrict__ dh_prev, int B, int H) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * H) return;
  const int b = (int)(i / H), j = (int)(i % H);
  const float* g = gates + (int64_t)b * 3 * H;
  const float r = g[j], z = g[H + j], n = g[2 * H + j];
  const float hn = gh[(int64_t)b * 3 * H + 2 * H + j];
  const float hp = h_prev ? h_prev[i] : 0.f;
  float dh = 0.f;
  if (dh_a) dh += Store<TS>::load(dh_a + i);
  if (dh_b) dh += Store<TS>::load(dh_b + i);
  if (dh_c) dh += dh_c[i];
  const float dn = dh * (1.f - z);
  const float dz = dh * (hp - n);
  const float dan = dn * (1.f - n * n);
  const float dar = dan * hn * r * (1.f - r);
  const float daz = dz * z * (1.f - z);
  TS* o = dgi + (int64_t)b * 3 * H;
  TS* q = dgh + (int64_t)b * 3 * H;
  Store<TS>::store(o + j, dar);
  Store<TS>::store(o + H + j, daz);
  Store<TS>::store(o + 2 * H + j, dan);
  Store<TS>::store(q + j, dar);
  Store<TS>::store(q + H + j, daz);
  Store<TS>::store(q + 2 * H + j, dan * r);
  if (dh_prev) dh_prev[i] = dh * z;
}

__global__ void __launch_bounds__(128) actor_head_kernel(const float* __restrict__ logits, const float* __restrict__ eps,
                                                         float std, float* __restrict__ action, float* __restrict__ logprob,
                                                         float* __restrict__ mean, int B, int K) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float q = 0.f;
  for (int k = 0; k < K; ++k) {
    const float mu = sigmoidf_(logits[(int64_t)b * K + k]);
    float a = mu + std * eps[(int64_t)b * K + k];
    a = fminf(fmaxf(a, 0.f), 1.f);
    const float t = (a - mu) / std;
    q = fmaf(t, t, q);
    action[(int64_t)b * K + k] = a;
    if (mean) mean[(int64_t)b * K + k] = mu;
  }
  logprob[b] = -0.5f * q - (float)K * logf(std) - 0.5f * (float)K * 1.8378770664093453f;  // log(2 pi)
}

template <typename TS, typename TD>
__global__ void __launch_bounds__(256) cast_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 4 <= n) {
    store4(dst + i, load4(src + i));
  } else {
    for (int64_t j = i; j < n; ++j) Store<TD>::store(dst + j, Store<TS>::load(src + j));
  }
}

// fp32 -> stack of bf16 planes: hi = bf16(x), mid = bf16(x - hi), lo = bf16(x - hi - mid); plane p at rows
// [p * plane_rows, +rows), the padding rows [rows, plane_rows) are written as zeros (they take part in reductions over rows).
template <int PLANES>
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ src, int64_t rows, int cols, int64_t plane_rows,
                                                           __nv_bfloat16* __restrict__ dst) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;        // element index within a padded plane
  const int64_t n_pad = plane_rows * cols;
  if (i >= n_pad) return;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < rows * (int64_t)cols) v = *reinterpret_cast<const float4*>(src + i);
  float r[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int pl = 0; pl < PLANES; ++pl) {
    __nv_bfloat16 h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      h[e] = __float2bfloat16_rn(r[e]);
      r[e] -= __bfloat162float(h[e]);                                              // exact: the residual of an RN rounding
    }
    uint2 q;
    q.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
    q.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
    *reinterpret_cast<uint2*>(dst + (int64_t)pl * n_pad + i) = q;
  }
}

__global__ void __launch_bounds__(256) row_segments_kernel(const int64_t* __restrict__ offsets, int32_t* __restrict__ row_seg) {
  const int b = blockIdx.x;
  const int64_t lo = offsets[b], hi = offsets[b + 1];
  for (int64_t n = lo + threadIdx.x; n < hi; n += blockDim.x) row_seg[n] = b;
}

// Column sums: each lane owns 4 consecutive columns (8/16-byte loads), a warp covers 128 columns of one row per
// step; grid = (column tiles of 128, row slices); warps of a CTA walk interleaved rows, shared-memory reduce across
// warps, one atomicAdd per column per CTA.
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ a, int64_t M, int N, float* __restrict__ out) {
  __shared__ float sm[8][128];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 128 + lane * 4;
  const int64_t rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if ((N & 3) == 0) {
    if (col < N) {
      int64_t r = r0 + w;
      for (; r + 24 < r1; r += 32) {            // 4 independent loads in flight per lane
        const float4 v0 = load4(a + r * N + col), v1 = load4(a + (r + 8) * N + col);
        const float4 v2 = load4(a + (r + 16) * N + col), v3 = load4(a + (r + 24) * N + col);
        acc.x += (v0.x + v1.x) + (v2.x + v3.x);
        acc.y += (v0.y + v1.y) + (v2.y + v3.y);
        acc.z += (v0.z + v1.z) + (v2.z + v3.z);
        acc.w += (v0.w + v1.w) + (v2.w + v3.w);
      }
      for (; r < r1; r += 8) {
        const float4 v = load4(a + r * N + col);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
  } else {
    for (int64_t r = r0 + w; r < r1; r += 8) {
      if (col + 0 < N) acc.x += Store<T>::load(a + r * N + col + 0);
      if (col + 1 < N) acc.y += Store<T>::load(a + r * N + col + 1);
      if (col + 2 < N) acc.z += Store<T>::load(a + r * N + col + 2);
      if (col + 3 < N) acc.w += Store<T>::load(a + r * N + col + 3);
    }
  }
  sm[w][lane * 4 + 0] = acc.x; sm[w][lane * 4 + 1] = acc.y; sm[w][lane * 4 + 2] = acc.z; sm[w][lane * 4 + 3] = acc.w;
  __syncthreads();
  if (threadIdx.x < 128) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c < N) atomicAdd(&out[c], t);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) relu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dz,
                                                       int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 4 <= n) {
    float4 g = load4(dy + i);
    const float4 v = load4(y + i);
    g.x = v.x > 0.f ? g.x : 0.f;
    g.y = v.y > 0.f ? g.y : 0.f;
    g.z = v.z > 0.f ? g.z : 0.f;
    g.w = v.w > 0.f ? g.w : 0.f;
    store4(dz + i, g);
  } else {
    for (int64_t j = i; j < n; ++j)
      Store<T>::store(dz + j, Store<T>::load(y + j) > 0.f ? Store<T>::load(dy + j) : 0.f);
  }
}

// In-place inverted dropout (clam.py:46-48,70-71; abmil.py:15,18): x *= keep / (1 - p).  Counter-based RNG: element i
// keeps iff splitmix64(seed, i) maps to u >= p; the 64-bit seed is read from DEVICE memory so that a CUDA-graph replay
// sees a fresh seed each step.  Dropped elements become exact zeros (the backward recognises them by that).
__device__ __forceinline__ float hash_uniform(uint64_t seed, uint64_t i) {
  uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}

template <typename T>
__global__ void __launch_bounds__(256) dropout_kernel(T* __restrict__ x, int64_t n, float p, float scale,
                                                      const int64_t* __restrict__ seed_dev) {
  const uint64_t seed = (uint64_t)seed_dev[0];
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 4 <= n) {
    float4 v = load4(x + i);
    v.x = hash_uniform(seed, i) >= p ? v.x * scale : 0.f;
    v.y = hash_uniform(seed, i + 1) >= p ? v.y * scale : 0.f;
    v.z = hash_uniform(seed, i + 2) >= p ? v.z * scale : 0.f;
    v.w = hash_uniform(seed, i + 3) >= p ? v.w * scale : 0.f;
    store4(x + i, v);
  } else {
    for (int64_t j = i; j < n; ++j)
      Store<T>::store(x + j, hash_uniform(seed, j) >= p ? Store<T>::load(x + j) * scale : 0.f);
  }
}

int colsum_impl(const void* a, int64_t M, int N, int dtype, float* out, cudaStream_t st, int accumulate) {
  if (!accumulate) {                       // the kernel adds its slices' partial sums with atomics
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N, st);
    if (e != cudaSuccess) {
      set_error("colsum: memset failed: %s", cudaGetErrorString(e));
      return MURCL_ECUDA;
    }
  }
  const int col_tiles = ceil_div(N, 128);
  int slices = (8 * sm_count() + col_tiles - 1) / col_tiles;
  const int by_rows = (int)((M + 255) / 256);
  if (slices > by_rows) slices = by_rows;
  if (slices < 1) slices = 1;
  if (slices > 65535) slices = 65535;
  dim3 grid(col_tiles, slices);
  if (dtype == MURCL_F32) colsum_kernel<float><<<grid, 256, 0, st>>>((const float*)a, M, N, out);
  else colsum_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)a, M, N, out);
  return check_launch("colsum_kernel");
}

}  // namespace murcl

using namespace murcl;

extern "C" {

int murcl_gru_cell_fwd(const float* gi, const float* gh, const float* h_prev, float* h_new, float* gates_out, int B, int H,
                       void* stream) {
  MURCL_REQUIRE(gi && gh && h_new, "gru_cell_fwd: null pointer");
  MURCL_REQUIRE(B >= 0 && H > 0, "gru_cell_fwd: bad shape");
  if (B == 0) return MURCL_OK;
  gru_cell_fwd_kernel<<<ceil_div((int64_t)B * H, 256), 256, 0, as_stream(stream)>>>(gi, gh, h_prev, h_new, gates_out, B, H);
  return check_launch("gru_cell_fwd_kernel");
}

int murcl_gru_cell_bwd(const float* dh_new, const float* gates, const float* gh, const float* h_prev, float* dgi, float* dgh,
                       float* dh_prev, int B, int H, void* stream) {
  MURCL_REQUIRE(dh_new && gates && gh && dgi && dgh, "gru_cell_bwd: null pointer");
  MURCL_REQUIRE(B >= 0 && H > 0, "gru_cell_bwd: bad shape");
  if (B == 0) return MURCL_OK;
  gru_cell_bwd_kernel<<<ceil_div((int64_t)B * H, 256), 256, 0, as_stream(stream)>>>(dh_new, gates, gh, h_prev, dgi, dgh,
                                                                                     dh_prev, B, H);
  return check_launch("gru_cell_bwd_kernel");
}

int murcl_gru_cell_fwd_tape(const float* gi, const float* gh, const float* h_prev, float* h_new, void* h_new_s, void* h_new_s2,
                            float* gates, int B, int H, int dtype, void* stream) {
  MURCL_REQUIRE(gi && gh && h_new && gates, "gru_cell_fwd_tape: null pointer");
  MURCL_REQUIRE(B >= 0 && H > 0, "gru_cell_fwd_tape: bad shape");
  MURCL_REQUIRE(dtype == MURCL_F32 || dtype == MURCL_BF16, "gru_cell_fwd_tape: bad dtype %d", dtype);
  if (B == 0) return MURCL_OK;
  const int grid = ceil_div((int64_t)B * H, 256);
  if (dtype == MURCL_BF16)
    gru_cell_fwd_tape_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(gi, gh, h_prev, h_new, (__nv_bfloat16*)h_new_s,
                                                                                 (__nv_bfloat16*)h_new_s2, gates, B, H);
  else
    gru_cell_fwd_tape_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(gi, gh, h_prev, h_new, (float*)h_new_s, (float*)h_new_s2, gates, B, H);
  return check_launch("gru_cell_fwd_tape_kernel");
}

int murcl_gru_cell_bwd_tape(const void* dh_a, const void* dh_b, const float* dh_c, const float* gates, const float* gh,
                            const float* h_prev, void* dgi, void* dgh, float* dh_prev, int B, int H, int dtype, void* stream) {
  MURCL_REQUIRE(gates && gh && dgi && dgh, "gru_cell_bwd_tape: null pointer");
  MURCL_REQUIRE(B >= 0 && H > 0, "gru_cell_bwd_tape: bad shape");
  MURCL_REQUIRE(dtype == MURCL_F32 || dtype == MURCL_BF16, "gru_cell_bwd_tape: bad dtype %d", dtype);
  if (B == 0) return MURCL_OK;
  const int grid = ceil_div((int64_t)B * H, 256);
  if (dtype == MURCL_BF16)
    gru_cell_bwd_tape_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(
        (const __nv_bfloat16*)dh_a, (const __nv_bfloat16*)dh_b, dh_c, gates, gh, h_prev, (__nv_bfloat16*)dgi, (__nv_bfloat16*)dgh, dh_prev, B, H);
  else
    gru_cell_bwd_tape_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)dh_a, (const float*)dh_b, dh_c, gates, gh, h_prev,
                                                                         (float*)dgi, (float*)dgh, dh_prev, B, H);
  return check_launch("gru_cell_bwd_tape_kernel");
}

int murcl_split_planes(const float* src, int64_t rows, int cols, int planes, int64_t plane_rows, void* dst, void* stream) {
  MURCL_REQUIRE(src && dst, "split_planes: null pointer");
  MURCL_REQUIRE(rows >= 0 && cols > 0 && cols % 4 == 0 && plane_rows >= rows && (planes == 2 || planes == 3),
                "split_planes: bad shape rows=%lld cols=%d plane_rows=%lld planes=%d", (long long)rows, cols, (long long)plane_rows, planes);
  MURCL_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0, "split_planes: 16-byte alignment");
  if (plane_rows == 0) return MURCL_OK;
  const int grid = ceil_div(plane_rows * cols, 1024);
  if (planes == 2) split_planes_kernel<2><<<grid, 256, 0, as_stream(stream)>>>(src, rows, cols, plane_rows, (__nv_bfloat16*)dst);
  else split_planes_kernel<3><<<grid, 256, 0, as_stream(stream)>>>(src, rows, cols, plane_rows, (__nv_bfloat16*)dst);
  return check_launch("split_planes_kernel");
}

int murcl_actor_head(const float* logits, const float* eps, float std, float* action, float* logprob, float* mean, int B,
                     int K, void* stream) {
  MURCL_REQUIRE(logits && eps && action && logprob, "actor_head: null pointer");
  MURCL_REQUIRE(B >= 0 && K > 0 && std > 0.f, "actor_head: bad B=%d K=%d std=%g", B, K, (double)std);
  if (B == 0) return MURCL_OK;
  actor_head_kernel<<<ceil_div(B, 128), 128, 0, as_stream(stream)>>>(logits, eps, std, action, logprob, mean, B, K);
  return check_launch("actor_head_kernel");
}

int murcl_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream) {
  MURCL_REQUIRE(src && dst, "cast: null pointer");
  MURCL_REQUIRE(n >= 0, "cast: negative length");
  if (n == 0) return MURCL_OK;
  cudaStream_t st = as_stream(stream);
  const int grid = ceil_div((n + 3) / 4, 256);
  if (src_dtype == MURCL_F32 && dst_dtype == MURCL_BF16)
    cast_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>((const float*)src, (__nv_bfloat16*)dst, n);
  else if (src_dtype == MURCL_BF16 && dst_dtype == MURCL_F32)
    cast_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, (float*)dst, n);
  else if (src_dtype == MURCL_F32 && dst_dtype == MURCL_F32)
    cast_kernel<float, float><<<grid, 256, 0, st>>>((const float*)src, (float*)dst, n);
  else if (src_dtype == MURCL_BF16 && dst_dtype == MURCL_BF16)
    cast_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, n);
  else MURCL_REQUIRE(false, "cast: bad dtypes %d -> %d", src_dtype, dst_dtype);
  return check_launch("cast_kernel");
}

int murcl_row_segments(const int64_t* offsets, int B, int32_t* row_seg, void* stream) {
  MURCL_REQUIRE(offsets && row_seg, "row_segments: null pointer");
  if (B <= 0) return MURCL_OK;
  row_segments_kernel<<<B, 256, 0, as_stream(stream)>>>(offsets, row_seg);
  return check_launch("row_segments_kernel");
}

int murcl_colsum(const void* a, int64_t M, int N, int dtype, float* out, void* stream) {
  MURCL_REQUIRE(a && out, "colsum: null pointer");
  MURCL_REQUIRE(M >= 0 && N > 0 && (dtype == MURCL_F32 || dtype == MURCL_BF16), "colsum: bad arguments");
  if (M == 0) {
    MURCL_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N, as_stream(stream)));
    return MURCL_OK;
  }
  return colsum_impl(a, M, N, dtype, out, as_stream(stream), 0);
}

int murcl_dropout(void* x, int64_t n, float p, const int64_t* seed_dev, int dtype, void* stream) {
  MURCL_REQUIRE(x && seed_dev, "dropout: null pointer");
  MURCL_REQUIRE(p >= 0.f && p < 1.f, "dropout: p=%g out of [0,1)", (double)p);
  if (n <= 0 || p == 0.f) return MURCL_OK;
  cudaStream_t st = as_stream(stream);
  const int grid = ceil_div((n + 3) / 4, 256);
  const float scale = 1.f / (1.f - p);
  if (dtype == MURCL_F32) dropout_kernel<float><<<grid, 256, 0, st>>>((float*)x, n, p, scale, seed_dev);
  else if (dtype == MURCL_BF16) dropout_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((__nv_bfloat16*)x, n, p, scale, seed_dev);
  else MURCL_REQUIRE(false, "dropout: bad dtype %d", dtype);
  return check_launch("dropout_kernel");
}

int murcl_relu_bwd(const void* dy, const void* y, void* dz, int64_t n, int dtype, void* stream) {
  MURCL_REQUIRE(dy && y && dz, "relu_bwd: null pointer");
  if (n <= 0) return MURCL_OK;
  cudaStream_t st = as_stream(stream);
  const int grid = ceil_div((n + 3) / 4, 256);
  if (dtype == MURCL_F32) relu_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)dy, (const float*)y, (float*)dz, n);
  else if (dtype == MURCL_BF16)
    relu_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, (__nv_bfloat16*)dz, n);
  else MURCL_REQUIRE(false, "relu_bwd: bad dtype %d", dtype);
  return check_launch("relu_bwd_kernel");
}

}  // extern "C"
