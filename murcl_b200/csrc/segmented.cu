// Segmented (per-bag) warp/block reductions: CLAM's both-ends top-k (clam.py:107-110), DSMIL's
// critical-instance arg-max and q.q_max attention logits (dsmil.py:71-77), row gather/scatter.
#include "common.cuh"

namespace murcl {

struct KeyIdx {
  float v;
  int i;
};
// Ordering "a comes before b": larger value first, lower index first among equals.
__device__ __forceinline__ bool before(KeyIdx a, KeyIdx b) { return a.v > b.v || (a.v == b.v && a.i < b.i); }

__device__ __forceinline__ KeyIdx warp_best(KeyIdx k) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    KeyIdx other;
    other.v = __shfl_xor_sync(0xffffffffu, k.v, o);
    other.i = __shfl_xor_sync(0xffffffffu, k.i, o);
    if (other.i >= 0 && (k.i < 0 || before(other, k))) k = other;
  }
  return k;
}

__device__ __forceinline__ KeyIdx block_best(KeyIdx k, KeyIdx* red) {
  k = warp_best(k);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = k;
  __syncthreads();
  KeyIdx r;
  r.v = 0.f;
  r.i = -1;
  if (threadIdx.x < nw) r = red[threadIdx.x];
  if (w == 0) r = warp_best(r);
  if (threadIdx.x == 0) red[0] = r;
  __syncthreads();
  return red[0];
}

// Single pass (k <= 8): one CTA per (bag, end).  Every thread keeps the k best of its strided share of the bag in a
// sorted register list (insertion by compile-time shifts); the bag's k best are among those 256 x k candidates and are
// drawn by k rounds of block arg-best over the list heads (the winner's owner pops its head).  p is read ONCE per end
// (the multi-pass kernel below reads it k times); same (value desc, index asc) order, so the same indices, ties included.
template <int KMAX>
__global__ void __launch_bounds__(256) seg_topk_ends_kernel(const float* __restrict__ p, const int64_t* __restrict__ offsets, int k,
                                                            int32_t* __restrict__ top_idx, int32_t* __restrict__ bot_idx) {
  __shared__ KeyIdx red[32];
  const int b = blockIdx.x, end = blockIdx.y;
  const int64_t lo = offsets[b], hi = offsets[b + 1];
  const float sign = end == 0 ? 1.f : -1.f;
  int32_t* out = (end == 0 ? top_idx : bot_idx) + (int64_t)b * k;
  KeyIdx lst[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) { lst[j].v = 0.f; lst[j].i = -1; }
  auto consider = [&](float val, int idx) {
    KeyIdx c;
    c.v = sign * val;
    c.i = idx;
    if (lst[KMAX - 1].i < 0 || before(c, lst[KMAX - 1])) {
      // insert c, keeping the list sorted: every slot takes the better of (its left neighbour, c) once c belongs left of it
#pragma unroll
      for (int j = KMAX - 1; j > 0; --j) {
        const bool left = lst[j - 1].i < 0 || before(c, lst[j - 1]);      // c goes somewhere left of slot j
        const bool here = lst[j].i < 0 || before(c, lst[j]);              // c goes at or left of slot j
        if (left) lst[j] = lst[j - 1];
        else if (here) lst[j] = c;
      }
      if (lst[0].i < 0 || before(c, lst[0])) lst[0] = c;
    }
  };
  constexpr int UNR = 8;                                   // loads of a thread in flight (the insertions are branchy and serial)
  int64_t n = lo + threadIdx.x;
  for (; n + (int64_t)(UNR - 1) * blockDim.x < hi; n += (int64_t)UNR * blockDim.x) {
    float v[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) v[u] = __ldg(p + n + (int64_t)u * blockDim.x);
#pragma unroll
    for (int u = 0; u < UNR; ++u) consider(v[u], (int)(n + (int64_t)u * blockDim.x - lo));
  }
  for (; n < hi; n += blockDim.x) consider(__ldg(p + n), (int)(n - lo));
  for (int r = 0; r < k; ++r) {
    const KeyIdx best = block_best(lst[0], red);
    if (threadIdx.x == 0) out[r] = best.i >= 0 ? (int32_t)(lo + best.i) : -1;
    if (best.i >= 0 && lst[0].i == best.i) {                               // the owner pops its head
#pragma unroll
      for (int j = 0; j < KMAX - 1; ++j) lst[j] = lst[j + 1];
      lst[KMAX - 1].i = -1;
    }
  }
}

// Multi-pass fallback (k > 8): one CTA per (bag, end).  k rounds of block arg-best; each round only considers elements that
// come strictly after the previous winner in the (value desc, index asc) order, so no marks are
// needed and the result is deterministic.  end 0: largest p, end 1: smallest p (sign flipped).
__global__ void __launch_bounds__(256) seg_topk_ends_multipass_kernel(const float* __restrict__ p,
                                                            const int64_t* __restrict__ offsets, int k,
                                                            int32_t* __restrict__ top_idx, int32_t* __restrict__ bot_idx) {
  __shared__ KeyIdx red[32];
  const int b = blockIdx.x, end = blockIdx.y;
  const int64_t lo = offsets[b], hi = offsets[b + 1];
  const float sign = end == 0 ? 1.f : -1.f;
  int32_t* out = (end == 0 ? top_idx : bot_idx) + (int64_t)b * k;
  KeyIdx last;
  last.v = INFINITY;
  last.i = -1;
  for (int r = 0; r < k; ++r) {
    KeyIdx best;
    best.v = 0.f;
    best.i = -1;
    for (int64_t n = lo + threadIdx.x; n < hi; n += blockDim.x) {
      KeyIdx c;
      c.v = sign * p[n];
      c.i = (int)(n - lo);
      const bool eligible = (last.i < 0) || before(last, c);
      if (eligible && (best.i < 0 || before(c, best))) best = c;
    }
    best = block_best(best, red);
    if (threadIdx.x == 0) out[r] = best.i >= 0 ? (int32_t)(lo + best.i) : -1;
    last = best;
    if (best.i < 0) {
      for (int rr = r + 1 + threadIdx.x; rr < k; rr += blockDim.x) out[rr] = -1;
      break;
    }
  }
}

__global__ void __launch_bounds__(256) seg_argmax_kernel(const float* __restrict__ c, const int64_t* __restrict__ offsets,
                                                         int C, int32_t* __restrict__ idx) {
  __shared__ KeyIdx red[32];
  const int b = blockIdx.x, cls = blockIdx.y;
  const int64_t lo = offsets[b], hi = offsets[b + 1];
  KeyIdx best;
  best.v = 0.f;
  best.i = -1;
  for (int64_t n = lo + threadIdx.x; n < hi; n += blockDim.x) {
    KeyIdx k;
    k.v = c[n * C + cls];
    k.i = (int)(n - lo);
    if (best.i < 0 || before(k, best)) best = k;
  }
  best = block_best(best, red);
  if (threadIdx.x == 0) idx[(int64_t)b * C + cls] = best.i >= 0 ? (int32_t)(lo + best.i) : -1;
}

// a[n,c] = q[n].q[crit[b,c]] / sqrt(Dq)
__global__ void __launch_bounds__(256) dsmil_scores_fwd_kernel(const float* __restrict__ q, const int32_t* __restrict__ crit,
                                                               const int32_t* __restrict__ row_seg, int64_t n_rows, int C,
                                                               int Dq, float inv_sqrt, float* __restrict__ a) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int b = row_seg[row];
  for (int c = 0; c < C; ++c) {
    const float* qm = q + (int64_t)crit[(int64_t)b * C + c] * Dq;
    float acc = 0.f;
    for (int d = lane; d < Dq; d += 32) acc = fmaf(q[row * Dq + d], qm[d], acc);
    acc = warp_sum(acc);
    if (lane == 0) a[row * C + c] = acc * inv_sqrt;
  }
}

// dq[n,:] = sum_c da[n,c] * q[crit[b,c],:] * inv_sqrt   (row-owned write)
__global__ void __launch_bounds__(256) dsmil_scores_bwd_rows_kernel(const float* __restrict__ q, const float* __restrict__ da,
                                                                    const int32_t* __restrict__ crit,
                                                                    const int32_t* __restrict__ row_seg, int64_t n_rows,
                                                                    int C, int Dq, float inv_sqrt, float* __restrict__ dq) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int b = row_seg[row];
  for (int d = lane; d < Dq; d += 32) {
    float v = 0.f;
    for (int c = 0; c < C; ++c) v = fmaf(da[row * C + c], q[(int64_t)crit[(int64_t)b * C + c] * Dq + d], v);
    dq[row * Dq + d] = v * inv_sqrt;
  }
}

// dq[crit[b,c],:] += sum_n da[n,c] * q[n,:] * inv_sqrt.  grid (bag, class, row slice): thread per column, each CTA
// reduces its slice of the bag's rows and adds it atomically (several slices / classes may share the target row).
__global__ void __launch_bounds__(128) dsmil_scores_bwd_crit_kernel(const float* __restrict__ q, const float* __restrict__ da,
                                                                    const int32_t* __restrict__ crit,
                                                                    const int64_t* __restrict__ offsets, int C, int Dq,
                                                                    float inv_sqrt, float* __restrict__ dq) {
  const int b = blockIdx.x, c = blockIdx.y;
  const int64_t lo = offsets[b], hi = offsets[b + 1];
  const int64_t len = hi - lo;
  const int64_t r0 = lo + len * blockIdx.z / gridDim.z, r1 = lo + len * (blockIdx.z + 1) / gridDim.z;
  const int64_t target = crit[(int64_t)b * C + c];
  if (target < 0 || r0 >= r1) return;
  for (int d = threadIdx.x; d < Dq; d += blockDim.x) {
    float a0 = 0.f, a1 = 0.f;
    int64_t n = r0;
    for (; n + 1 < r1; n += 2) {
      a0 = fmaf(da[n * C + c], q[n * Dq + d], a0);
      a1 = fmaf(da[(n + 1) * C + c], q[(n + 1) * Dq + d], a1);
    }
    if (n < r1) a0 = fmaf(da[n * C + c], q[n * Dq + d], a0);
    atomicAdd(&dq[target * Dq + d], (a0 + a1) * inv_sqrt);
  }
}

template <typename T>
__global__ void __launch_bounds__(128) gather_rows_kernel(const T* __restrict__ h, const int32_t* __restrict__ idx, int L,
                                                          float* __restrict__ out) {
  const int i = blockIdx.x;
  const int64_t src = idx[i];
  for (int d = threadIdx.x; d < L; d += blockDim.x)
    out[(int64_t)i * L + d] = src >= 0 ? Store<T>::load(h + src * L + d) : 0.f;
}

// Thread per column, serial over the (few) indexed rows: duplicate indices are race free.
template <typename T>
__global__ void __launch_bounds__(128) scatter_add_rows_kernel(T* __restrict__ dh, const int32_t* __restrict__ idx, int n_idx,
                                                               int L, const float* __restrict__ rows) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= L) return;
  for (int i = 0; i < n_idx; ++i) {
    const int64_t dst = idx[i];
    if (dst < 0) continue;
    T* p = dh + dst * L + d;
    Store<T>::store(p, Store<T>::load(p) + rows[(int64_t)i * L + d]);
  }
}

}  // namespace murcl

using namespace murcl;

extern "C" {

int murcl_seg_topk_ends(const float* p, const int64_t* offsets, int B, int k, int32_t* top_idx, int32_t* bot_idx,
                        void* stream) {
  MURCL_REQUIRE(p && offsets && top_idx && bot_idx, "seg_topk_ends: null pointer");
  MURCL_REQUIRE(B >= 0 && k > 0 && k <= 1024, "seg_topk_ends: bad B=%d k=%d", B, k);
  if (B == 0) return MURCL_OK;
  if (k <= 8) {
    seg_topk_ends_kernel<8><<<dim3(B, 2), 256, 0, as_stream(stream)>>>(p, offsets, k, top_idx, bot_idx);
    return check_launch("seg_topk_ends_kernel");
  }
  seg_topk_ends_multipass_kernel<<<dim3(B, 2), 256, 0, as_stream(stream)>>>(p, offsets, k, top_idx, bot_idx);
  return check_launch("seg_topk_ends_multipass_kernel");
}

int murcl_seg_argmax(const float* c, const int64_t* offsets, int B, int C, int32_t* idx, void* stream) {
  MURCL_REQUIRE(c && offsets && idx, "seg_argmax: null pointer");
  MURCL_REQUIRE(B >= 0 && C > 0 && C <= 65535, "seg_argmax: bad shape");
  if (B == 0) return MURCL_OK;
  seg_argmax_kernel<<<dim3(B, C), 256, 0, as_stream(stream)>>>(c, offsets, C, idx);
  return check_launch("seg_argmax_kernel");
}

int murcl_dsmil_scores_fwd(const float* q, const int32_t* crit, const int32_t* row_seg, int64_t n_rows, int C, int Dq,
                           float* a, void* stream) {
  MURCL_REQUIRE(q && crit && row_seg && a, "dsmil_scores_fwd: null pointer");
  MURCL_REQUIRE(C > 0 && Dq > 0, "dsmil_scores_fwd: bad shape");
  if (n_rows == 0) return MURCL_OK;
  dsmil_scores_fwd_kernel<<<ceil_div(n_rows, 8), 256, 0, as_stream(stream)>>>(q, crit, row_seg, n_rows, C, Dq,
                                                                                1.f / sqrtf((float)Dq), a);
  return check_launch("dsmil_scores_fwd_kernel");
}

int murcl_dsmil_scores_bwd(const float* q, const float* da, const int32_t* crit, const int32_t* row_seg,
                           const int64_t* offsets, int64_t n_rows, int B, int C, int Dq, float* dq, void* stream) {
  MURCL_REQUIRE(q && da && crit && row_seg && offsets && dq, "dsmil_scores_bwd: null pointer");
  MURCL_REQUIRE(B >= 0 && C > 0 && C <= 65535 && Dq > 0, "dsmil_scores_bwd: bad shape");
  if (n_rows == 0 || B == 0) return MURCL_OK;
  cudaStream_t st = as_stream(stream);
  const float inv = 1.f / sqrtf((float)Dq);
  dsmil_scores_bwd_rows_kernel<<<ceil_div(n_rows, 8), 256, 0, st>>>(q, da, crit, row_seg, n_rows, C, Dq, inv, dq);
  int rc = check_launch("dsmil_scores_bwd_rows_kernel");
  if (rc != MURCL_OK) return rc;
  int slices = (int)((n_rows / (B > 0 ? B : 1) + 255) / 256);          // ~256 rows per CTA
  if (slices < 1) slices = 1;
  if (slices > 1024) slices = 1024;
  dsmil_scores_bwd_crit_kernel<<<dim3(B, C, slices), 128, 0, st>>>(q, da, crit, offsets, C, Dq, inv, dq);
  return check_launch("dsmil_scores_bwd_crit_kernel");
}

int murcl_gather_rows(const void* h, const int32_t* idx, int n_idx, int L, int dtype, float* out, void* stream) {
  MURCL_REQUIRE(h && idx && out, "gather_rows: null pointer");
  MURCL_REQUIRE(n_idx >= 0 && L > 0, "gather_rows: bad shape");
  if (n_idx == 0) return MURCL_OK;
  cudaStream_t st = as_stream(stream);
  if (dtype == MURCL_F32) gather_rows_kernel<float><<<n_idx, 128, 0, st>>>((const float*)h, idx, L, out);
  else if (dtype == MURCL_BF16) gather_rows_kernel<__nv_bfloat16><<<n_idx, 128, 0, st>>>((const __nv_bfloat16*)h, idx, L, out);
  else MURCL_REQUIRE(false, "gather_rows: bad dtype %d", dtype);
  return check_launch("gather_rows_kernel");
}

int murcl_scatter_add_rows(void* dh, const int32_t* idx, int n_idx, int L, int dtype, const float* rows, void* stream) {
  MURCL_REQUIRE(dh && idx && rows, "scatter_add_rows: null pointer");
  MURCL_REQUIRE(n_idx >= 0 && L > 0, "scatter_add_rows: bad shape");
  if (n_idx == 0) return MURCL_OK;
  cudaStream_t st = as_stream(stream);
  const int grid = ceil_div(L, 128);
  if (dtype == MURCL_F32) scatter_add_rows_kernel<float><<<grid, 128, 0, st>>>((float*)dh, idx, n_idx, L, rows);
  else if (dtype == MURCL_BF16) scatter_add_rows_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>((__nv_bfloat16*)dh, idx, n_idx, L, rows);
  else MURCL_REQUIRE(false, "scatter_add_rows: bad dtype %d", dtype);
  return check_launch("scatter_add_rows_kernel");
}

}  // extern "C"

// ---- CLAM instance-classifier tail (clam.py:112-118,126-131) --------------------------------------
// One CTA per group g (= one (bag, class) pair): rows[g0:g1, :] are the gathered top-k (+ bottom-k)
// instances, classifier cls[g] is Linear(L -> 2).  Writes the mean cross-entropy of the group, the
// arg-max predictions and dlogits = (softmax - onehot) / rows_in_group for the backward.
namespace murcl {

__global__ void __launch_bounds__(256) clam_inst_ce_fwd_kernel(const float* __restrict__ rows, const int32_t* __restrict__ targets,
                                                               const int32_t* __restrict__ group_off,
                                                               const int32_t* __restrict__ group_cls,
                                                               const float* __restrict__ w, const float* __restrict__ bias, int L,
                                                               float* __restrict__ loss, int32_t* __restrict__ preds,
                                                               float* __restrict__ dlogits) {
  __shared__ float red[32];
  const int g = blockIdx.x;
  const int r0 = group_off[g], r1 = group_off[g + 1];
  const int cls = group_cls[g];
  const float* w0 = w + (int64_t)cls * 2 * L;
  const float* w1 = w0 + L;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float ce_sum = 0.f;
  for (int r = r0 + wid; r < r1; r += nw) {
    const float* x = rows + (int64_t)r * L;
    float a0 = 0.f, a1 = 0.f;
    for (int d = lane; d < L; d += 32) {
      const float v = x[d];
      a0 = fmaf(v, w0[d], a0);
      a1 = fmaf(v, w1[d], a1);
    }
    a0 = warp_sum(a0) + bias[cls * 2 + 0];
    a1 = warp_sum(a1) + bias[cls * 2 + 1];
    if (lane == 0) {
      const float m = fmaxf(a0, a1);
      const float e0 = expf(a0 - m), e1 = expf(a1 - m);
      const float lse = m + logf(e0 + e1);
      const int t = targets[r];
      ce_sum += lse - (t == 0 ? a0 : a1);
      const float inv = 1.f / (float)(r1 - r0);
      dlogits[(int64_t)r * 2 + 0] = (e0 / (e0 + e1) - (t == 0 ? 1.f : 0.f)) * inv;
      dlogits[(int64_t)r * 2 + 1] = (e1 / (e0 + e1) - (t == 1 ? 1.f : 0.f)) * inv;
      preds[r] = a1 > a0 ? 1 : 0;
    }
  }
  ce_sum = block_sum(ce_sum, red);
  if (threadIdx.x == 0) loss[g] = r1 > r0 ? ce_sum / (float)(r1 - r0) : 0.f;
}

// drows[r,:] = gl[g] * (dl[r,0] w0 + dl[r,1] w1);  dw[cls] += gl[g] * dl^T rows;  db[cls] += gl[g] * sum dl.
__global__ void __launch_bounds__(128) clam_inst_ce_bwd_kernel(const float* __restrict__ rows, const float* __restrict__ dlogits,
                                                               const float* __restrict__ gloss, const int32_t* __restrict__ group_off,
                                                               const int32_t* __restrict__ group_cls, const float* __restrict__ w,
                                                               int L, float* __restrict__ drows, float* __restrict__ dw,
                                                               float* __restrict__ db) {
  const int g = blockIdx.x;
  const int r0 = group_off[g], r1 = group_off[g + 1];
  const int cls = group_cls[g];
  const float gl = gloss[g];
  const float* w0 = w + (int64_t)cls * 2 * L;
  const float* w1 = w0 + L;
  for (int d = threadIdx.x; d < L; d += blockDim.x) {
    float s0 = 0.f, s1 = 0.f;
    for (int r = r0; r < r1; ++r) {
      const float d0 = dlogits[(int64_t)r * 2] * gl, d1 = dlogits[(int64_t)r * 2 + 1] * gl;
      const float x = rows[(int64_t)r * L + d];
      drows[(int64_t)r * L + d] = d0 * w0[d] + d1 * w1[d];
      s0 = fmaf(d0, x, s0);
      s1 = fmaf(d1, x, s1);
    }
    atomicAdd(&dw[(int64_t)cls * 2 * L + d], s0);
    atomicAdd(&dw[(int64_t)cls * 2 * L + L + d], s1);
  }
  if (threadIdx.x < 2) {
    float s = 0.f;
    for (int r = r0; r < r1; ++r) s += dlogits[(int64_t)r * 2 + threadIdx.x] * gl;
    atomicAdd(&db[cls * 2 + threadIdx.x], s);
  }
}

}  // namespace murcl

extern "C" {

int murcl_clam_inst_ce_fwd(const float* rows, const int32_t* targets, const int32_t* group_off, const int32_t* group_cls,
                           int G, const float* w, const float* bias, int L, float* loss, int32_t* preds, float* dlogits,
                           void* stream) {
  MURCL_REQUIRE(rows && targets && group_off && group_cls && w && bias && loss && preds && dlogits,
                "clam_inst_ce_fwd: null pointer");
  MURCL_REQUIRE(G >= 0 && L > 0, "clam_inst_ce_fwd: bad shape");
  if (G == 0) return MURCL_OK;
  murcl::clam_inst_ce_fwd_kernel<<<G, 256, 0, murcl::as_stream(stream)>>>(rows, targets, group_off, group_cls, w, bias, L,
                                                                          loss, preds, dlogits);
  return murcl::check_launch("clam_inst_ce_fwd_kernel");
}

int murcl_clam_inst_ce_bwd(const float* rows, const float* dlogits, const float* gloss, const int32_t* group_off,
                           const int32_t* group_cls, int G, const float* w, int L, float* drows, float* dw, float* db,
                           void* stream) {
  MURCL_REQUIRE(rows && dlogits && gloss && group_off && group_cls && w && drows && dw && db,
                "clam_inst_ce_bwd: null pointer");
  MURCL_REQUIRE(G >= 0 && L > 0, "clam_inst_ce_bwd: bad shape");
  if (G == 0) return MURCL_OK;
  murcl::clam_inst_ce_bwd_kernel<<<G, 128, 0, murcl::as_stream(stream)>>>(rows, dlogits, gloss, group_off, group_cls, w, L,
                                                                          drows, dw, db);
  return murcl::check_launch("clam_inst_ce_bwd_kernel");
}

}  // extern "C"
