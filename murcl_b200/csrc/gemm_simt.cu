// fp32-accumulate SIMT GEMM family behind murcl_linear_{fwd,bwd_input,bwd_weight}.
//
// This is the exact-fp32 path (1e-5 relative budget against the reference; tensor cores cannot
// meet it without split precision) and the any-shape path for operands the tcgen05 kernels in
// gemm_tc.cu do not take.  Roofline: FFMA pipe (148 SM x 128 lanes x 2 x clk), ridge ~11 FLOP/B.
//
// C[m,n] = sum_k A(m,k) * B(n,k).  Operand storage is described by two flags:
//   A_KC: A stored [M,K] row-major (k contiguous)   else stored [K,M] row-major (m contiguous)
//   B_KC: B stored [N,K] row-major (k contiguous)   else stored [K,N] row-major (n contiguous)
// 128x128x16 CTA tile, 256 threads, 8x8 register tile per thread, double-buffered smem with
// register prefetch.  gridDim.z splits K (used for the weight gradient where K = #rows).
#include "common.cuh"

namespace murcl {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

struct GemmParams {
  const void* A;
  const void* B;
  void* C;
  int64_t M;      // rows of C
  int N;          // cols of C
  int64_t K;      // reduction length
  int64_t lda, ldb, ldc;
  const float* bias;        // [N] or null
  int act;
  const void* relu_src;     // [M,N] (ld = ldc) or null: multiply by (relu_src > 0)
  const float* row_scale;   // [M] or null: C += row_scale[m] * row_vec[row_seg[m]*N + n]
  const float* row_vec;
  const int32_t* row_seg;
  const unsigned long long* relu_bits;   // [N/64][M] bit mask (bit = activation > 0), alternative to relu_src
  float out_scale;          // applied last (1/(1-p) of a dropout that followed the masked ReLU); 0 means 1
  int64_t k_chunk;          // K range per blockIdx.z
  int64_t split_stride;     // elements between split outputs (C is fp32 workspace when gridDim.z > 1)
};

template <typename T>
__device__ __forceinline__ void load8(const T* base, int64_t ld, int64_t major, int64_t minor, int64_t major_lim,
                                      int64_t minor_lim, bool vec_ok, float (&v)[8]) {
  // 8 consecutive elements along the contiguous ("minor") direction of row `major`.
  if (major < major_lim && vec_ok && minor + 8 <= minor_lim) {
    const T* p = base + major * ld + minor;
    float4 a = load4(p), b = load4(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      v[j] = (major < major_lim && minor + j < minor_lim) ? Store<T>::load(base + major * ld + minor + j) : 0.f;
  }
}

template <typename TA, typename TB, typename TC, bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256, 2) gemm_simt_kernel(GemmParams p) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const TA* A = static_cast<const TA*>(p.A);
  const TB* B = static_cast<const TB*>(p.B);
  const int t = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int64_t k_begin = (int64_t)blockIdx.z * p.k_chunk;
  const int64_t k_end = min(p.K, k_begin + p.k_chunk);
  const bool a_vec = (p.lda % 8 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool b_vec = (p.ldb % 8 == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);

  // loader coordinates
  const int a_r = A_KC ? (t & 127) : (t >> 4);          // KC: row in tile | MC: k in tile
  const int a_c = A_KC ? (t >> 7) * 8 : (t & 15) * 8;   // KC: k offset    | MC: m offset
  const int b_r = B_KC ? (t & 127) : (t >> 4);
  const int b_c = B_KC ? (t >> 7) * 8 : (t & 15) * 8;

  float ra[8], rb[8];
  auto fetch = [&](int64_t k0) {
    if (A_KC) load8<TA>(A, p.lda, m0 + a_r, k0 + a_c, p.M, k_end, a_vec, ra);
    else load8<TA>(A, p.lda, k0 + a_r, m0 + a_c, k_end, p.M, a_vec, ra);
    if (B_KC) load8<TB>(B, p.ldb, n0 + b_r, k0 + b_c, p.N, k_end, b_vec, rb);
    else load8<TB>(B, p.ldb, k0 + b_r, n0 + b_c, k_end, p.N, b_vec, rb);
  };
  auto stash = [&](int buf) {
    if (A_KC) {
#pragma unroll
      for (int j = 0; j < 8; ++j) As[buf][a_c + j][a_r] = ra[j];
    } else {
      *reinterpret_cast<float4*>(&As[buf][a_r][a_c]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
      *reinterpret_cast<float4*>(&As[buf][a_r][a_c + 4]) = make_float4(ra[4], ra[5], ra[6], ra[7]);
    }
    if (B_KC) {
#pragma unroll
      for (int j = 0; j < 8; ++j) Bs[buf][b_c + j][b_r] = rb[j];
    } else {
      *reinterpret_cast<float4*>(&Bs[buf][b_r][b_c]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
      *reinterpret_cast<float4*>(&Bs[buf][b_r][b_c + 4]) = make_float4(rb[4], rb[5], rb[6], rb[7]);
    }
  };

  const int tx = t & 15, ty = t >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  if (k_begin < k_end) {
    fetch(k_begin);
    stash(0);
  }
  __syncthreads();
  int buf = 0;
  for (int64_t k0 = k_begin; k0 < k_end; k0 += BK) {
    const bool more = k0 + BK < k_end;
    if (more) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      stash(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

  // epilogue
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
    float rs = 0.f;
    const float* rv = nullptr;
    if (p.row_scale) {
      rs = p.row_scale[m];
      rv = p.row_vec + (int64_t)p.row_seg[m] * p.N;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (split) {
        static_cast<float*>(p.C)[(int64_t)blockIdx.z * p.split_stride + m * p.ldc + n] = v;
        continue;
      }
      if (p.bias) v += p.bias[n];
      v = apply_act(v, p.act, n, p.N);
      if (rv) v = fmaf(rs, rv[n], v);
      if (p.relu_bits) {
        if (!((p.relu_bits[(int64_t)(n >> 6) * p.M + m] >> (n & 63)) & 1ull)) v = 0.f;
      } else if (p.relu_src && !(Store<TC>::load(static_cast<const TC*>(p.relu_src) + m * p.ldc + n) > 0.f)) {
        v = 0.f;
      }
      if (p.out_scale != 0.f) v *= p.out_scale;
      Store<TC>::store(static_cast<TC*>(p.C) + m * p.ldc + n, v);
    }
  }
}

// Sum split-K partials: out[i] (+)= sum_z ws[z*stride + i].  `accumulate` adds to the existing contents: weight gradients
// of the T patch-steps of an optimiser step land in ONE persistent gradient buffer without a separate add kernel.
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, int splits, int64_t stride,
                                                            float* __restrict__ out, int64_t n, int accumulate, int vec_ok) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  if (i + 4 <= n && vec_ok) {
    float4 s = accumulate ? *reinterpret_cast<const float4*>(out + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z = 0; z < splits; ++z) {
      const float4 v = *reinterpret_cast<const float4*>(ws + z * stride + i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(out + i) = s;
  } else {
    for (int64_t j = i; j < n && j < i + 4; ++j) {
      float s = accumulate ? out[j] : 0.f;
      for (int z = 0; z < splits; ++z) s += ws[z * stride + j];
      out[j] = s;
    }
  }
}

// Split-K tail for the small-M dense layers: sum the partials, then the same fused epilogue as the main kernel.
template <typename TC>
__global__ void __launch_bounds__(256) splitk_epilogue_kernel(const float* __restrict__ ws, int splits, GemmParams p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.M * p.N) return;
  const int64_t m = i / p.N;
  const int n = (int)(i % p.N);
  float v = 0.f;
  for (int z = 0; z < splits; ++z) v += ws[z * p.split_stride + m * p.ldc + n];
  if (p.bias) v += p.bias[n];
  v = apply_act(v, p.act, n, p.N);
  if (p.row_scale) v = fmaf(p.row_scale[m], p.row_vec[(int64_t)p.row_seg[m] * p.N + n], v);
  if (p.relu_bits) {
    if (!((p.relu_bits[(int64_t)(n >> 6) * p.M + m] >> (n & 63)) & 1ull)) v = 0.f;
  } else if (p.relu_src && !(Store<TC>::load(static_cast<const TC*>(p.relu_src) + m * p.ldc + n) > 0.f)) {
    v = 0.f;
  }
  if (p.out_scale != 0.f) v *= p.out_scale;
  Store<TC>::store(static_cast<TC*>(p.C) + m * p.ldc + n, v);
}

// bits[(n/64) * M + m] bit (n%64) = y[m,n] > 0   (the SIMT forward's companion to the tcgen05 epilogue's bit mask)
template <typename T>
__global__ void __launch_bounds__(256) relu_bits_kernel(const T* __restrict__ y, int64_t M, int N, unsigned long long* __restrict__ bits) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int slab = blockIdx.y;
  if (m >= M) return;
  unsigned long long b = 0ull;
  const T* r = y + m * N + slab * 64;
  for (int i = 0; i < 64; ++i)
    if (Store<T>::load(r + i) > 0.f) b |= 1ull << i;
  bits[(int64_t)slab * M + m] = b;
}

int launch_relu_bits(const void* y, int64_t M, int N, int dtype, unsigned long long* bits, cudaStream_t st) {
  dim3 grid(ceil_div(M, 256), N / 64);
  if (dtype == MURCL_F32) relu_bits_kernel<float><<<grid, 256, 0, st>>>((const float*)y, M, N, bits);
  else relu_bits_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)y, M, N, bits);
  return check_launch("relu_bits_kernel");
}

int launch_splitk_reduce(const float* ws, int splits, int64_t stride, float* out, int64_t n, cudaStream_t st, int accumulate) {
  const int vec_ok = (((reinterpret_cast<uintptr_t>(ws) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 && (stride & 3) == 0) ? 1 : 0;
  splitk_reduce_kernel<<<ceil_div(n, 1024), 256, 0, st>>>(ws, splits, stride, out, n, accumulate, vec_ok);
  return check_launch("splitk_reduce_kernel");
}

template <typename TA, typename TB, typename TC, bool A_KC, bool B_KC>
static int launch(const GemmParams& p, int splits, cudaStream_t st) {
  dim3 grid(ceil_div(p.M, BM), ceil_div(p.N, BN), splits);
  gemm_simt_kernel<TA, TB, TC, A_KC, B_KC><<<grid, 256, 0, st>>>(p);
  return check_launch("gemm_simt_kernel");
}

// Forward / input-gradient launch.  When the output has too few 128x128 tiles to fill the GPU (the heads:
// M = batch size) the reduction is split across gridDim.z into a stream-ordered scratch buffer.
template <typename TA, typename TB, typename TC, bool A_KC, bool B_KC>
static int launch_auto(GemmParams p, cudaStream_t st, float* ext_ws = nullptr, int64_t ext_ws_floats = 0) {
  const int ctas = ceil_div(p.M, BM) * ceil_div(p.N, BN);
  int splits = 1;
  if (ctas * 2 <= sm_count() && p.K >= 256) {
    splits = (2 * sm_count()) / ctas;
    const int by_k = (int)(p.K / 64);
    if (splits > by_k) splits = by_k;
    if (splits > 32) splits = 32;
  }
  if (splits <= 1) return launch<TA, TB, TC, A_KC, B_KC>(p, 1, st);
  int64_t chunk = (p.K + splits - 1) / splits;
  chunk = (chunk + BK - 1) / BK * BK;
  splits = (int)((p.K + chunk - 1) / chunk);
  float* ws = nullptr;
  const size_t bytes = sizeof(float) * (size_t)splits * p.M * p.N;
  const bool own = !(ext_ws != nullptr && (int64_t)splits * p.M * p.N <= ext_ws_floats);
  if (!own) {
    ws = ext_ws;                       // caller-provided scratch (keeps the call free of allocations)
  } else {
    // Stream-ordered scratch from the device's default pool.  By default that pool hands its memory back to the OS at every
    // synchronisation point, so a training loop that reads its loss each step paid a real allocation (~1 ms) per small-M
    // layer per step (measured: DSMIL's M = 2 bag layers 1.8 ms each): keep the pool's memory cached instead.
    static PerDeviceOnce pool_kept;
    if (const int slot = pool_kept.pending(); slot >= 0) {
      int dev = 0;
      cudaMemPool_t pool;
      if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      pool_kept.mark(slot);
    }
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&ws), bytes, st);
    if (e != cudaSuccess) {
      set_error("gemm_simt: cudaMallocAsync(%zu) failed: %s", bytes, cudaGetErrorString(e));
      return MURCL_ECUDA;
    }
  }
  GemmParams q = p;
  q.C = ws;
  q.k_chunk = chunk;
  q.split_stride = p.M * p.N;
  q.ldc = p.N;
  int rc = launch<TA, TB, float, A_KC, B_KC>(q, splits, st);
  if (rc == MURCL_OK) {
    GemmParams r = p;
    r.split_stride = p.M * p.N;
    splitk_epilogue_kernel<TC><<<ceil_div(p.M * p.N, 256), 256, 0, st>>>(ws, splits, r);
    rc = check_launch("splitk_epilogue_kernel");
  }
  if (own) cudaFreeAsync(ws, st);
  return rc;
}

// ---- skinny outputs (N <= 16): instance classifiers, actor head -----------------------------------
// One warp per row: x[m,:] is read once (HBM-bound for large M), the N weight rows stay in L1/L2.
constexpr int SKINNY_N = 16;

template <typename T, typename TO>
__global__ void __launch_bounds__(256) skinny_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                         const float* __restrict__ bias, TO* __restrict__ y, int64_t M, int N,
                                                         int K, int act) {
  const int lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  float acc[SKINNY_N];
#pragma unroll
  for (int n = 0; n < SKINNY_N; ++n) acc[n] = 0.f;
  const T* xr = x + m * K;
  if ((K & 3) == 0) {
    for (int k = lane * 4; k < K; k += 128) {
      const float4 a = load4(xr + k);
#pragma unroll
      for (int n = 0; n < SKINNY_N; ++n) {
        if (n < N) {
          const float4 b = load4(w + (int64_t)n * K + k);
          acc[n] += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
        }
      }
    }
  } else {
    for (int k = lane; k < K; k += 32) {
      const float a = Store<T>::load(xr + k);
#pragma unroll
      for (int n = 0; n < SKINNY_N; ++n)
        if (n < N) acc[n] = fmaf(a, Store<T>::load(w + (int64_t)n * K + k), acc[n]);
    }
  }
#pragma unroll
  for (int n = 0; n < SKINNY_N; ++n) {
    if (n < N) {
      float v = warp_sum(acc[n]);
      if (lane == 0) {
        if (bias) v += bias[n];
        Store<TO>::store(y + m * N + n, apply_act(v, act, n, N));
      }
    }
  }
}

// dx[m,k] = sum_n dy[m,n] * w[n,k] for a short reduction (N <= 16); same epilogue options as the tiled kernel.
template <typename T>
__global__ void __launch_bounds__(256) skinny_dgrad_kernel(const T* __restrict__ dy, const T* __restrict__ w, T* __restrict__ dx,
                                                           int64_t M, int N, int K, const T* __restrict__ relu_src,
                                                           const float* __restrict__ row_scale, const float* __restrict__ row_vec,
                                                           const int32_t* __restrict__ row_seg, float out_scale) {
  const int lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  float g[SKINNY_N];
#pragma unroll
  for (int n = 0; n < SKINNY_N; ++n) g[n] = n < N ? Store<T>::load(dy + m * N + n) : 0.f;
  const float rs = row_scale ? row_scale[m] : 0.f;
  const float* rv = row_scale ? row_vec + (int64_t)row_seg[m] * K : nullptr;
  for (int k = lane; k < K; k += 32) {
    float v = 0.f;
#pragma unroll
    for (int n = 0; n < SKINNY_N; ++n)
      if (n < N) v = fmaf(g[n], Store<T>::load(w + (int64_t)n * K + k), v);
    if (rv) v = fmaf(rs, rv[k], v);
    if (relu_src && !(Store<T>::load(relu_src + m * K + k) > 0.f)) v = 0.f;
    Store<T>::store(dx + m * K + k, v * out_scale);
  }
}

int simt_linear_fwd(const void* x, const void* w, const float* bias, void* y, int64_t M, int N, int K, int act,
                    int dtype, int out_dtype, cudaStream_t st, float* ws, int64_t ws_floats) {
  GemmParams p{};
  p.A = x; p.B = w; p.C = y; p.M = M; p.N = N; p.K = K; p.lda = K; p.ldb = K; p.ldc = N;
  p.bias = bias; p.act = act; p.k_chunk = K;
  if (N <= SKINNY_N) {
    const int grid = ceil_div(M, 8);
    if (dtype == MURCL_F32 && out_dtype == MURCL_F32)
      skinny_fwd_kernel<float, float><<<grid, 256, 0, st>>>((const float*)x, (const float*)w, bias, (float*)y, M, N, K, act);
    else if (dtype == MURCL_BF16 && out_dtype == MURCL_BF16)
      skinny_fwd_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w, bias,
                                                                           (__nv_bfloat16*)y, M, N, K, act);
    else if (dtype == MURCL_BF16 && out_dtype == MURCL_F32)
      skinny_fwd_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w, bias,
                                                                   (float*)y, M, N, K, act);
    else {
      set_error("linear_fwd(simt): unsupported dtype pair %d -> %d", dtype, out_dtype);
      return MURCL_EUNSUPPORTED;
    }
    return check_launch("skinny_fwd_kernel");
  }
  if (dtype == MURCL_F32 && out_dtype == MURCL_F32) return launch_auto<float, float, float, true, true>(p, st, ws, ws_floats);
  if (dtype == MURCL_BF16 && out_dtype == MURCL_BF16)
    return launch_auto<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16, true, true>(p, st);
  if (dtype == MURCL_BF16 && out_dtype == MURCL_F32)
    return launch_auto<__nv_bfloat16, __nv_bfloat16, float, true, true>(p, st);
  set_error("linear_fwd(simt): unsupported dtype pair %d -> %d", dtype, out_dtype);
  return MURCL_EUNSUPPORTED;
}

int simt_linear_bwd_input(const void* dy, const void* w, void* dx, int64_t M, int N, int K, const void* relu_src,
                          const float* row_scale, const float* row_vec, const int32_t* row_seg, float out_scale, int dtype,
                          cudaStream_t st, float* ws, int64_t ws_floats, const unsigned long long* relu_bits) {
  GemmParams p{};
  // C = dx [M, K]; reduction over N; A = dy [M,N] k-contiguous; B(n'=k_in, k'=n) = w[n, k_in] n'-contiguous.
  p.A = dy; p.B = w; p.C = dx; p.M = M; p.N = K; p.K = N; p.lda = N; p.ldb = K; p.ldc = K;
  p.relu_src = relu_src; p.row_scale = row_scale; p.row_vec = row_vec; p.row_seg = row_seg; p.k_chunk = N;
  p.out_scale = out_scale; p.relu_bits = relu_bits;
  if (N <= SKINNY_N && relu_bits == nullptr) {
    const int grid = ceil_div(M, 8);
    if (dtype == MURCL_F32)
      skinny_dgrad_kernel<float><<<grid, 256, 0, st>>>((const float*)dy, (const float*)w, (float*)dx, M, N, K,
                                                       (const float*)relu_src, row_scale, row_vec, row_seg, out_scale);
    else
      skinny_dgrad_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)w,
                                                               (__nv_bfloat16*)dx, M, N, K, (const __nv_bfloat16*)relu_src,
                                                               row_scale, row_vec, row_seg, out_scale);
    return check_launch("skinny_dgrad_kernel");
  }
  if (dtype == MURCL_F32) return launch_auto<float, float, float, true, false>(p, st, ws, ws_floats);
  return launch_auto<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16, true, false>(p, st);
}

static int bwd_weight_splits(int64_t M, int N, int K) {
  const int tiles = ceil_div(N, BM) * ceil_div(K, BN);
  int splits = (2 * sm_count() + tiles - 1) / tiles;
  const int max_by_rows = (int)((M + 4 * BK - 1) / (4 * BK));
  if (splits > max_by_rows) splits = max_by_rows;
  if (splits < 1) splits = 1;
  if (splits > 512) splits = 512;
  return splits;
}

int64_t simt_linear_bwd_weight_workspace(int64_t M, int N, int K) {
  const int s = bwd_weight_splits(M, N, K);
  return (int64_t)(s > 1 ? s : 1) * N * K;          // one slab even without a split: the accumulate mode goes through it
}

int simt_linear_bwd_weight(const void* dy, const void* x, float* dw, int64_t M, int N, int K, int dtype,
                           float* workspace, cudaStream_t st, int accumulate) {
  const int splits = bwd_weight_splits(M, N, K);
  const bool via_ws = splits > 1 || accumulate;
  GemmParams p{};
  // C = dw [N, K]; reduction over rows M; A(m'=n, k'=m) = dy[m, n]; B(n'=k_in, k'=m) = x[m, k_in].
  p.A = dy; p.B = x; p.M = N; p.N = K; p.K = M; p.lda = N; p.ldb = K; p.ldc = K;
  int64_t chunk = (M + splits - 1) / splits;
  chunk = (chunk + BK - 1) / BK * BK;
  p.k_chunk = chunk;
  p.split_stride = (int64_t)N * K;
  p.C = via_ws ? (void*)workspace : (void*)dw;
  if (via_ws && workspace == nullptr) {
    set_error("linear_bwd_weight: workspace required (%d splits, accumulate=%d)", splits, accumulate);
    return MURCL_EINVAL;
  }
  int rc = (dtype == MURCL_F32) ? launch<float, float, float, false, false>(p, splits, st)
                                : launch<__nv_bfloat16, __nv_bfloat16, float, false, false>(p, splits, st);
  if (rc != MURCL_OK || !via_ws) return rc;
  const int64_t n = (int64_t)N * K;
  return launch_splitk_reduce(workspace, splits, n, dw, n, st, accumulate);
}

}  // namespace murcl
