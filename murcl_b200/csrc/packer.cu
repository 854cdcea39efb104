// Ragged-bag packer: cluster-wise window selection, stream compaction, row gather, zero pad and
// mixup.  Replaces utils/datasets.py:274-308 (get_feats) and :263-271 (mixup).
//
// HBM layout: all bags concatenated row-wise (CSR): feats[n_rows, D] fp32, offsets[B+1] int64,
// per patch (cluster label, rank inside its cluster) int32, per bag cluster sizes [B, K] int32.
// The kernels are pure HBM movers: 16-byte vector accesses, one warp per gathered row.
#include "common.cuh"

namespace murcl {

// ---- ingest: rank of each patch inside its cluster -------------------------------------------
// One CTA per bag walks the bag in patch order; warp-level match + cross-warp prefix in smem.
__global__ void __launch_bounds__(256) rank_patches_kernel(const int32_t* __restrict__ patch_cluster,
                                                           const int64_t* __restrict__ offsets, int K,
                                                           int32_t* __restrict__ patch_rank,
                                                           int32_t* __restrict__ cluster_sizes) {
  extern __shared__ int32_t sm[];
  const int nw = blockDim.x >> 5;
  int32_t* base = sm;           // [K] running count per cluster
  int32_t* wcnt = sm + K;       // [nw][K] per-warp counts of the current chunk
  const int bag = blockIdx.x;
  const int64_t lo = offsets[bag], hi = offsets[bag + 1];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < K * (nw + 1); i += blockDim.x) sm[i] = 0;
  __syncthreads();
  for (int64_t start = lo; start < hi; start += blockDim.x) {
    const int64_t p = start + threadIdx.x;
    int c = (p < hi) ? patch_cluster[p] : -1;
    if (c >= K) c = -1;
    const unsigned peers = __match_any_sync(0xffffffffu, c);
    const int in_warp = __popc(peers & ((1u << lane) - 1u));
    if (c >= 0 && in_warp == 0) wcnt[w * K + c] = __popc(peers);
    __syncthreads();
    if (c >= 0) {
      int r = base[c] + in_warp;
      for (int ww = 0; ww < w; ++ww) r += wcnt[ww * K + c];
      patch_rank[p] = r;
    } else if (p < hi) {
      patch_rank[p] = -1;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      int add = 0;
      for (int ww = 0; ww < nw; ++ww) {
        add += wcnt[ww * K + k];
        wcnt[ww * K + k] = 0;
      }
      base[k] += add;
    }
    __syncthreads();
  }
  for (int k = threadIdx.x; k < K; k += blockDim.x) cluster_sizes[(int64_t)bag * K + k] = base[k];
}

// ---- selection ---------------------------------------------------------------------------------
// Window arithmetic restated from datasets.py:285-291 in the float32 / int32 types torch uses
// there, plus Python's slice clamping (:294).  No FMA contraction is possible: every expression
// is a single multiply followed by rint/floor.
__device__ __forceinline__ void cluster_window(int n, float ratio, float a, int& start, int& stop) {
  const int size = (int)rintf(__fmul_rn((float)n, ratio));
  const int l = (int)floorf(__fmul_rn(a, (float)(n - size)));
  const int r = l + size;
  start = (l >= 0) ? min(l, n) : max(n + l, 0);
  stop = (r >= 0) ? min(r, n) : max(n + r, 0);
  stop = max(stop, start);
}

__global__ void __launch_bounds__(512) pack_select_kernel(const int32_t* __restrict__ patch_cluster,
                                                          const int32_t* __restrict__ patch_rank,
                                                          const int64_t* __restrict__ offsets,
                                                          const int32_t* __restrict__ cluster_sizes,
                                                          const int32_t* __restrict__ slot_bag,
                                                          const float* __restrict__ actions, int K, int FS,
                                                          int32_t* __restrict__ sel_idx, int32_t* __restrict__ sel_cnt) {
  extern __shared__ int32_t sm[];
  int32_t* w_start = sm;        // [K]
  int32_t* w_stop = sm + K;     // [K]
  __shared__ int32_t warp_tot[16];
  __shared__ int32_t s_base;
  const int slot = blockIdx.x;
  const int bag = slot_bag ? slot_bag[slot] : slot;
  const int64_t lo = offsets[bag], hi = offsets[bag + 1];
  const int64_t n_patch = hi - lo;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float ratio = (float)((double)FS / (double)n_patch);   // python double division, cast to f32
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    int s, e;
    cluster_window(cluster_sizes[(int64_t)bag * K + k], ratio, actions[(int64_t)slot * K + k], s, e);
    w_start[k] = s;
    w_stop[k] = e;
  }
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  int32_t* out = sel_idx + (int64_t)slot * FS;
  // Stream compaction of the selection predicate in patch order.  A thread takes PER = 4 consecutive patches per
  // round (a round covers 4 * blockDim patches, so a 15k-patch slide needs 8 rounds instead of 30; every round costs
  // two CTA barriers and a dependent global load), positions come from a warp prefix over the per-thread counts plus the
  // running base.
  constexpr int PER = 4;
  for (int64_t start = lo; start < hi; start += (int64_t)blockDim.x * PER) {
    const int base = s_base;
    if (base >= FS) break;                       // uniform: s_base is read after a barrier
    const int64_t p0 = start + (int64_t)threadIdx.x * PER;
    bool keep[PER];
    int cnt = 0;
#pragma unroll
    for (int e = 0; e < PER; ++e) {
      const int64_t p = p0 + e;
      keep[e] = false;
      if (p < hi) {
        const int c = patch_cluster[p];
        if (c >= 0 && c < K) {
          const int r = patch_rank[p];
          keep[e] = (r >= w_start[c]) && (r < w_stop[c]);
        }
      }
      cnt += keep[e] ? 1 : 0;
    }
    int incl = cnt;                              // inclusive prefix of the counts within the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    int before = 0, total = 0;
    for (int ww = 0; ww < nw; ++ww) {
      const int t = warp_tot[ww];
      before += (ww < w) ? t : 0;
      total += t;
    }
    int pos = base + before + incl - cnt;
#pragma unroll
    for (int e = 0; e < PER; ++e) {
      if (keep[e]) {
        if (pos < FS) out[pos] = (int32_t)(p0 + e);
        ++pos;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base = base + total;
    __syncthreads();
  }
  const int kept = min((int)s_base, FS);
  for (int i = kept + threadIdx.x; i < FS; i += blockDim.x) out[i] = -1;
  if (threadIdx.x == 0) sel_cnt[slot] = kept;
}

// ---- gather + pad + mixup ----------------------------------------------------------------------
// TI = storage of the CSR feature buffer (fp32 as the reference holds it, or bf16 for the bf16 mode's halved
// footprint / H2D volume), TO = storage of the packed batch.  The mix is always two rounded fp32 products and
// one rounded fp32 add (datasets.py:268-270), then one rounding to TO.
// `slot_order` (may be null): the order in which the grid walks the output slots.  With mixup every slot's source rows are
// read twice - as its own and as its partner's - and with the slots in plain order the second read comes ~S/3 slots (tens of
// MB) later: mostly an L2 miss.  Walking the slots along the CYCLES of the permutation (murcl_perm_cycle_order) makes the
// partner of one slot the very next slot, so every source row comes from DRAM once.  Where a row is written does not change.
template <typename TI, typename TO, bool MIX>
__global__ void __launch_bounds__(256) pack_gather_kernel(const TI* __restrict__ feats, int D,
                                                          const int32_t* __restrict__ sel_idx, int64_t n_out_rows,
                                                          int FS, const float* __restrict__ lam,
                                                          const int32_t* __restrict__ perm,
                                                          const int32_t* __restrict__ slot_order, TO* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t vrow = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (vrow >= n_out_rows) return;
  const int pos = (int)(vrow / FS), r = (int)(vrow % FS);
  const int slot = slot_order != nullptr ? slot_order[pos] : pos;
  const int64_t row = (int64_t)slot * FS + r;
  const int32_t ia = sel_idx[row];
  const TI* pa = (ia >= 0) ? feats + (int64_t)ia * D : nullptr;
  const TI* pb = nullptr;
  float l0 = 1.f, l1 = 0.f;
  if (MIX) {
    const int32_t ib = sel_idx[(int64_t)perm[slot] * FS + r];
    pb = (ib >= 0) ? feats + (int64_t)ib * D : nullptr;
    l0 = lam[slot];
    l1 = __fsub_rn(1.f, l0);
  }
  TO* po = out + row * D;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if ((D & 3) == 0) {
    for (int c = lane * 4; c < D; c += 128) {
      float4 a = pa ? load4(pa + c) : zero4;
      if (MIX) {
        const float4 b = pb ? load4(pb + c) : zero4;
        a.x = __fadd_rn(__fmul_rn(l0, a.x), __fmul_rn(l1, b.x));
        a.y = __fadd_rn(__fmul_rn(l0, a.y), __fmul_rn(l1, b.y));
        a.z = __fadd_rn(__fmul_rn(l0, a.z), __fmul_rn(l1, b.z));
        a.w = __fadd_rn(__fmul_rn(l0, a.w), __fmul_rn(l1, b.w));
      }
      store4(po + c, a);
    }
  } else {
    for (int c = lane; c < D; c += 32) {
      float a = pa ? Store<TI>::load(pa + c) : 0.f;
      if (MIX) a = __fadd_rn(__fmul_rn(l0, a), __fmul_rn(l1, pb ? Store<TI>::load(pb + c) : 0.f));
      Store<TO>::store(po + c, a);
    }
  }
}

template <typename TI, typename TO>
static void launch_gather(const void* feats, int D, const int32_t* sel_idx, int64_t rows, int FS, const float* lam,
                          const int32_t* perm, const int32_t* order, void* out, cudaStream_t st) {
  const int grid = ceil_div(rows, 8);
  if (lam != nullptr)
    pack_gather_kernel<TI, TO, true><<<grid, 256, 0, st>>>((const TI*)feats, D, sel_idx, rows, FS, lam, perm, order, (TO*)out);
  else
    pack_gather_kernel<TI, TO, false><<<grid, 256, 0, st>>>((const TI*)feats, D, sel_idx, rows, FS, lam, perm, order, (TO*)out);
}

// One CTA per permutation of S slots: lane 0 walks the cycles (start at the lowest unvisited slot, follow perm until the
// cycle closes) and appends every slot once.  A malformed `perm` (not a bijection, entries out of range) still yields a
// valid ordering of all S slots - only the locality is lost.
__global__ void __launch_bounds__(256) perm_cycle_order_kernel(const int32_t* __restrict__ perm, int S, int32_t* __restrict__ order) {
  extern __shared__ int32_t sm[];
  int32_t* p = sm;
  unsigned char* seen = reinterpret_cast<unsigned char*>(sm + S);
  const int32_t* src = perm + (int64_t)blockIdx.x * S;
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    p[i] = src[i];
    seen[i] = 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t* dst = order + (int64_t)blockIdx.x * S;
    int n = 0;
    for (int s0 = 0; s0 < S; ++s0) {
      int cur = s0;
      while (cur >= 0 && cur < S && !seen[cur]) {
        seen[cur] = 1;
        dst[n++] = cur;
        cur = p[cur];
      }
    }
  }
}

}  // namespace murcl

using namespace murcl;

extern "C" {

int murcl_csr_rank_patches(const int32_t* patch_cluster, const int64_t* offsets, int B, int K, int32_t* patch_rank,
                           int32_t* cluster_sizes, void* stream) {
  MURCL_REQUIRE(patch_cluster && offsets && patch_rank && cluster_sizes, "csr_rank_patches: null pointer");
  MURCL_REQUIRE(B >= 0 && K > 0 && K <= 2048, "csr_rank_patches: B=%d K=%d out of range", B, K);
  if (B == 0) return MURCL_OK;
  const size_t smem = sizeof(int32_t) * (size_t)K * (256 / 32 + 1);
  rank_patches_kernel<<<B, 256, smem, as_stream(stream)>>>(patch_cluster, offsets, K, patch_rank, cluster_sizes);
  return check_launch("rank_patches_kernel");
}

int murcl_pack_select(const int32_t* patch_cluster, const int32_t* patch_rank, const int64_t* offsets,
                      const int32_t* cluster_sizes, const int32_t* slot_bag, const float* actions, int S, int K, int FS,
                      int32_t* sel_idx, int32_t* sel_cnt, void* stream) {
  MURCL_REQUIRE(patch_cluster && patch_rank && offsets && cluster_sizes && actions && sel_idx && sel_cnt,
                "pack_select: null pointer");
  MURCL_REQUIRE(S >= 0 && K > 0 && K <= 4096 && FS > 0, "pack_select: S=%d K=%d FS=%d out of range", S, K, FS);
  if (S == 0) return MURCL_OK;
  pack_select_kernel<<<S, 512, sizeof(int32_t) * 2 * (size_t)K, as_stream(stream)>>>(
      patch_cluster, patch_rank, offsets, cluster_sizes, slot_bag, actions, K, FS, sel_idx, sel_cnt);
  return check_launch("pack_select_kernel");
}

int murcl_pack_gather(const void* feats, int feat_dtype, int D, const int32_t* sel_idx, int S, int FS, const float* lam,
                      const int32_t* perm, void* out, int out_dtype, void* stream) {
  return murcl_pack_gather_ordered(feats, feat_dtype, D, sel_idx, S, FS, lam, perm, nullptr, out, out_dtype, stream);
}

int murcl_perm_cycle_order(const int32_t* perm, int n_perm, int S, int32_t* order, void* stream) {
  MURCL_REQUIRE(perm && order, "perm_cycle_order: null pointer");
  MURCL_REQUIRE(n_perm >= 0 && S > 0 && S <= 8192, "perm_cycle_order: n_perm=%d S=%d out of range", n_perm, S);
  if (n_perm == 0) return MURCL_OK;
  const size_t smem = sizeof(int32_t) * (size_t)S + ((size_t)S + 3) / 4 * 4;
  perm_cycle_order_kernel<<<n_perm, 256, smem, as_stream(stream)>>>(perm, S, order);
  return check_launch("perm_cycle_order_kernel");
}

int murcl_pack_gather_ordered(const void* feats, int feat_dtype, int D, const int32_t* sel_idx, int S, int FS, const float* lam,
                              const int32_t* perm, const int32_t* slot_order, void* out, int out_dtype, void* stream) {
  MURCL_REQUIRE(feats && sel_idx && out, "pack_gather: null pointer");
  MURCL_REQUIRE(S >= 0 && FS > 0 && D > 0, "pack_gather: S=%d FS=%d D=%d out of range", S, FS, D);
  MURCL_REQUIRE((lam == nullptr) == (perm == nullptr), "pack_gather: lam and perm must be given together");
  MURCL_REQUIRE(out_dtype == MURCL_F32 || out_dtype == MURCL_BF16, "pack_gather: bad out_dtype %d", out_dtype);
  MURCL_REQUIRE(feat_dtype == MURCL_F32 || feat_dtype == MURCL_BF16, "pack_gather: bad feat_dtype %d", feat_dtype);
  if (S == 0) return MURCL_OK;
  const int64_t rows = (int64_t)S * FS;
  cudaStream_t st = as_stream(stream);
  if (feat_dtype == MURCL_F32) {
    if (out_dtype == MURCL_F32) launch_gather<float, float>(feats, D, sel_idx, rows, FS, lam, perm, slot_order, out, st);
    else launch_gather<float, __nv_bfloat16>(feats, D, sel_idx, rows, FS, lam, perm, slot_order, out, st);
  } else {
    if (out_dtype == MURCL_F32) launch_gather<__nv_bfloat16, float>(feats, D, sel_idx, rows, FS, lam, perm, slot_order, out, st);
    else launch_gather<__nv_bfloat16, __nv_bfloat16>(feats, D, sel_idx, rows, FS, lam, perm, slot_order, out, st);
  }
  return check_launch("pack_gather_kernel");
}

}  // extern "C"
