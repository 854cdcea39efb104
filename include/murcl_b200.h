/*
 * murcl_b200.h - C ABI of libmurcl_b200.so, the B200 (sm_100a) implementation of MuRCL's
 * per-slide MIL hot path.
 *
 * The reference (wwu98934/MuRCL) has no FFI of its own: its operator surface is Python
 * (SURVEY.md section 8b).  Each entry point below names the reference code it replaces
 * (file:line relative to the upstream repository).  The Python drop-ins under
 * murcl_b200/dropin/ bind these symbols with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (e.g. torch.Tensor.data_ptr());
 *     the library never allocates persistent device memory and never synchronises;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - return value: MURCL_OK or a negative MURCL_E* code; murcl_last_error() returns a
 *     thread-local message for the last failure on the calling thread;
 *   - matrices are row-major and dense unless a leading dimension is given;
 *   - `dtype` arguments take MURCL_F32 or MURCL_BF16 and describe the storage type of the
 *     large activation tensors; all reductions and accumulators are fp32.
 *   - "bags" are stored CSR style: rows of all bags concatenated, `offsets[B+1]` (int64)
 *     gives each bag's row range.
 */
#ifndef MURCL_B200_H
#define MURCL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MURCL_ABI_VERSION 1

#define MURCL_OK 0
#define MURCL_EINVAL (-1)       /* bad argument */
#define MURCL_EUNSUPPORTED (-2) /* shape / dtype combination not implemented */
#define MURCL_ECUDA (-3)        /* CUDA runtime or driver error */

#define MURCL_F32 0
#define MURCL_BF16 1

#define MURCL_ACT_NONE 0
#define MURCL_ACT_RELU 1
#define MURCL_ACT_TANH 2
#define MURCL_ACT_SIGMOID 3
#define MURCL_ACT_TANH_SIGMOID 4 /* tanh on columns [0,N/2), sigmoid on [N/2,N): gated attention */

#define MURCL_GEMM_AUTO 0  /* tcgen05 when dtype/shape allow, else SIMT */
#define MURCL_GEMM_SIMT 1  /* fp32 FFMA path (exact fp32 accumulate, any shape) */
#define MURCL_GEMM_TCGEN05 2 /* TMA + tcgen05.mma, bf16 operands, fp32 TMEM accumulators */

/* ---- library ------------------------------------------------------------------------ */
int murcl_version(void);
const char* murcl_last_error(void);
/* sm count and compute capability of the current device. */
int murcl_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Number of kernel launches issued by this library on the calling process so far. */
int64_t murcl_launch_count(void);
/* Scheduling hint, per calling thread, sticky until changed; returns the previous value.  descending != 0: the following
 * murcl_linear_fwd / murcl_linear_bwd_input (tcgen05 path) and murcl_attnpool_fwd launches walk their row tiles from the
 * LAST rows to the first.  Results do not change.  A chain of layers over an activation larger than L2 alternates the
 * direction so that every consumer starts with the rows its producer wrote last - the ones still in L2 - instead of the
 * ones written first, which have been evicted by then. */
int murcl_set_row_order(int descending);

/* ---- (1) ragged-bag packer: utils/datasets.py:274-308 (get_feats), :263-271 (mixup) ---- */

/* Ingest: per-patch rank inside its cluster and per-bag cluster sizes from the per-patch
 * cluster label (`features_cluster_indices`, wsi_processing/features_clustering.py:10-25).
 * patch_cluster[n_rows] (label in [0,K) or -1), offsets[B+1] -> patch_rank[n_rows],
 * cluster_sizes[B*K].  Equivalent to the position of the patch in the JSON inverted list. */
int murcl_csr_rank_patches(const int32_t* patch_cluster, const int64_t* offsets, int B, int K,
                           int32_t* patch_rank, int32_t* cluster_sizes, void* stream);

/* Selection (datasets.py:283-305): for output slot s reading bag slot_bag[s] with actions
 * actions[s*K..], keep patch p iff its rank lies in the window the reference's Python slice
 * selects, in ascending patch order, first FS survivors.  sel_idx[S*FS] receives GLOBAL row
 * indices into the CSR buffer (-1 = zero pad row), sel_cnt[S] the number kept.  Bit-exact. */
int murcl_pack_select(const int32_t* patch_cluster, const int32_t* patch_rank, const int64_t* offsets,
                      const int32_t* cluster_sizes, const int32_t* slot_bag, const float* actions,
                      int S, int K, int FS, int32_t* sel_idx, int32_t* sel_cnt, void* stream);

/* Gather + zero pad (+ mixup when lam != NULL): out[s,r,:] = lam[s]*x(s,r) + (1-lam[s])*x(perm[s],r)
 * with x(s,r) = feats[sel_idx[s,r]] or 0; products and sum rounded separately in fp32 like the
 * reference (datasets.py:268-270).  feats [n_rows,D] in feat_dtype (fp32 as the reference stores it; bf16 halves
 * the store and the per-step H2D volume in bf16 mode); out [S,FS,D] in out_dtype (MURCL_BF16 rounds the fp32
 * result to nearest-even). */
int murcl_pack_gather(const void* feats, int feat_dtype, int D, const int32_t* sel_idx, int S, int FS,
                      const float* lam, const int32_t* perm, void* out, int out_dtype, void* stream);

/* The same gather with the grid walking the output slots in `slot_order` (a permutation of [0,S), may be NULL = plain
 * order).  Results are identical; the order only decides which source rows are still in L2 when their second reader
 * (the slot they are mixed into) arrives.  murcl_perm_cycle_order writes, for each of n_perm permutations of S slots
 * (perm[n_perm*S], the mixup partner of every slot as passed to the gather), the slots listed cycle by cycle - partner
 * after partner - which makes every source row a DRAM read once instead of twice (datasets.py:263-271 mixes slot s with
 * slot perm[s]). */
int murcl_perm_cycle_order(const int32_t* perm, int n_perm, int S, int32_t* order, void* stream);
int murcl_pack_gather_ordered(const void* feats, int feat_dtype, int D, const int32_t* sel_idx, int S, int FS,
                              const float* lam, const int32_t* perm, const int32_t* slot_order, void* out,
                              int out_dtype, void* stream);

/* ---- dense layers: every nn.Linear on the path (abmil.py:12-32, clam.py:18-77,
 *      dsmil.py:9,54-59, rlmil.py:40-53,199-200) ----------------------------------------- */

/* y[M,N] = act(x[M,K] . w[N,K]^T + bias[N]).  x,w in `dtype`; y in `out_dtype`; bias fp32/NULL.
 * relu_bits (may be NULL; needs act = RELU, N % 64 == 0): receives one bit per output, bit (n % 64) of word
 * [(n / 64) * M + m] = (y[m,n] > 0) - 1/16 of the bytes of a bf16 activation, consumed by murcl_linear_bwd_input. */
int murcl_linear_fwd(const void* x, const void* w, const float* bias, void* y, int64_t M, int N, int K,
                     int act, int dtype, int out_dtype, int backend, uint64_t* relu_bits, void* stream);

/* dx[M,K] = dy[M,N] . w[N,K]; when relu_src != NULL (shape [M,K], storage `dtype`) the result is
 * multiplied by (relu_src > 0), i.e. it is the gradient w.r.t. the previous layer's
 * pre-activation.  When row_scale != NULL: dx += row_scale[m] * row_vec[row_seg[m]][k] before
 * masking (the direct softmax-pool term, see murcl_pool_bwd).  When col_sum != NULL (fp32 [K], zeroed by
 * the caller) the column sums of the stored result are accumulated into it: dx is the next layer's dZ, so this
 * is that layer's bias gradient, produced without another pass over dx.  out_scale (0 or 1 = none) multiplies
 * the masked result: it is 1/(1-p) when relu_src was dropped out after its ReLU (clam.py:70-71), since the zeros
 * of relu_src then mark "inactive OR dropped".  relu_bits (may be NULL, needs K % 64 == 0) is the bit mask
 * murcl_linear_fwd wrote for that activation; when given it replaces the read of relu_src. */
int murcl_linear_bwd_input(const void* dy, const void* w, void* dx, int64_t M, int N, int K,
                           const void* relu_src, const float* row_scale, const float* row_vec,
                           const int32_t* row_seg, float* col_sum, float out_scale, const uint64_t* relu_bits,
                           int dtype, int backend, void* stream);

/* dx[M,K] (fp32) += dy[M,N] . w[N,K]: the plain input gradient ADDED to an existing fp32 buffer, with the reduction over N
 * split across the GPU (vector atomics).  For batch-sized layers with a long reduction - the W_hh step of the GRU head's
 * backward recurrence (rlmil.py:208-220 differentiated: M = 128, N = 3 x 1024) - where one launch of murcl_linear_bwd_input
 * is a handful of CTAs.  bf16 operands; murcl_linear_bwd_input_accum_supported tells whether the shape is taken. */
int murcl_linear_bwd_input_accum_supported(int64_t M, int N, int K, int dtype);
int murcl_linear_bwd_input_accum(const void* dy, const void* w, float* dx, int64_t M, int N, int K, int dtype, void* stream);

/* dw[N,K] (fp32) = dy[M,N]^T . x[M,K], db[N] (fp32, may be NULL) = column sums of dy.
 * `workspace` (fp32) must hold murcl_linear_bwd_weight_workspace(M,N,K) floats.  accumulate != 0 ADDS to the existing
 * contents of dw / db instead of overwriting them: the T x 2 bag passes of one optimiser step (train_MuRCL.py:291-295:
 * one backward over all patch-steps) sum their weight gradients inside the split-K reduction / column-sum kernels, in a
 * persistent gradient buffer (murcl_b200/arena.py), with no separate accumulation kernel. */
int64_t murcl_linear_bwd_weight_workspace(int64_t M, int N, int K);
int murcl_linear_bwd_weight(const void* dy, const void* x, float* dw, float* db, int64_t M, int N, int K,
                            int dtype, int backend, float* workspace, int accumulate, void* stream);

/* ---- exact-fp32 dense layers on the bf16 tensor cores (split precision) -------------------------------------------
 * The reference computes in fp32; the 1e-5 parity mode used to run on FFMA only (25 TFLOP/s).  An fp32 value
 * is the exact sum of bf16 planes  x = hi + mid (+ lo)  (hi = bf16(x), mid = bf16(x - hi), lo = bf16(x - hi - mid)), a
 * product of two bf16 numbers is exact in the fp32 accumulator of tcgen05.mma, so  x.y = sum of plane products:  with two
 * planes the three products hi.hi + hi.mid + mid.hi drop terms up to 3 * 2^-18 |x||y| (measured: 6e-6 of the output scale
 * per layer - outside the 1e-5 budget once layers compound), with three planes six products reach 2^-24 (measured 1e-6:
 * the default of the fp32 mode).  murcl_split_planes writes the planes stacked along the
 * rows, [planes][plane_rows][cols] bf16, rows [rows, plane_rows) zero; plane_rows % 64 == 0 whenever the rows are a
 * reduction dimension (weights: always; activations: for the weight gradient).  The *_split GEMMs take such stacks for both
 * operands and write fp32; their fused epilogues are those of murcl_linear_fwd / murcl_linear_bwd_input (the ReLU bit mask
 * and the column sums are separate passes here).  murcl_linear_split_supported(M, N, K): M >= 1024, N % 64 == 0 (>= 128),
 * K % 64 == 0. */
int murcl_split_planes(const float* src, int64_t rows, int cols, int planes, int64_t plane_rows, void* dst, void* stream);
int murcl_linear_split_supported(int64_t M, int N, int K);
int murcl_linear_fwd_split(const void* xp, const void* wp, const float* bias, float* y, int64_t M, int N, int K, int act,
                           int planes, int64_t x_plane_rows, int64_t w_plane_rows, uint64_t* relu_bits, void* stream);
int murcl_linear_bwd_input_split(const void* dyp, const void* wp, float* dx, int64_t M, int N, int K, const float* row_scale,
                                 const float* row_vec, const int32_t* row_seg, float* col_sum, float out_scale,
                                 const uint64_t* relu_bits, int planes, int64_t dy_plane_rows, int64_t w_plane_rows,
                                 void* stream);
int64_t murcl_linear_bwd_weight_split_workspace(int64_t M, int N, int K);
int murcl_linear_bwd_weight_split(const void* dyp, const void* xp, float* dw, int64_t M, int N, int K, int planes,
                                  int64_t plane_rows, float* workspace, int accumulate, void* stream);

/* ---- (2) attention pooling: abmil.py:38-42, clam.py:37-60,139-170 --------------------- */

/* Raw attention score s[n] = sum_d wc[d]*g[n,d] + bc with g = u (gated=0, uv is [N,D]) or
 * u*v (gated=1, uv is [N,2D]: tanh branch in columns [0,D), sigmoid branch in [D,2D)). */
int murcl_attn_score_fwd(const void* uv, const float* wc, const float* bc, float* s, int64_t N, int D,
                         int gated, int dtype, void* stream);

/* Segmented softmax over the rows of each bag for C score columns: p[n,c] = post_scale_b *
 * softmax_n(s[n,c]); post_scale_b = 1/sqrt(N_b) when inv_sqrt_n (abmil.py:41) else 1.
 * stats[b,c,0..1] = (max, sum exp).  s,p fp32 [n_rows,C]. */
int murcl_seg_softmax(const float* s, const int64_t* offsets, int B, int C, int inv_sqrt_n, float* p,
                      float* stats, void* stream);

/* Segmented weighted row sum: out[b,c,:] = sum_{n in bag b} p[n,c] * h[n,:]  (abmil.py:42,
 * clam.py:170, dsmil.py:78).  h [n_rows,L] in dtype; out fp32 [B,C,L].  workspace fp32 of
 * murcl_seg_wsum_workspace(...) floats. */
int64_t murcl_seg_wsum_workspace(int64_t n_rows, int B, int C, int L);
int murcl_seg_wsum(const float* p, const void* h, const int64_t* offsets, int64_t n_rows, int B, int C, int L,
                   int dtype, float* out, float* workspace, void* stream);

/* Fused attention pooling, forward (abmil.py:36-45; clam.py:37-60,170): one pass over the rows h [n_rows, L] of B bags
 * computes  uv = act(h wab^T + bab)  (tanh, or tanh | sigmoid when gated; [n_rows, D*(1+gated)], bf16, saved for the
 * backward pass; NULL = do not keep),  s = wc . g(uv) + bc  (raw scores, fp32 [n_rows]),  p = post_scale_b * softmax_bag(s)
 * (fp32 [n_rows]),  M[b] = sum_n p[n] h[n]  (fp32 [B, L])  and  stats[b] = (max, sum exp).  Each 128-row tile of h is
 * staged once in shared memory by TMA, multiplied on the tensor cores (tcgen05, fp32 TMEM accumulators), gated, projected
 * to scores and pooled with tile-local online-softmax statistics from the same shared-memory tile; a per-bag merge kernel
 * folds the tile records and writes p.  Replaces murcl_linear_fwd + murcl_attn_score_fwd + murcl_seg_softmax +
 * murcl_seg_wsum for C == 1.  Supported when murcl_attnpool_supported(...) != 0: bf16, L <= 512 and L % 64 == 0,
 * D % 64 == 0, D*(1+gated) % 128 == 0 and <= 512.  workspace: murcl_attnpool_workspace(n_rows, B, L) floats.
 * offsets int64 [B+1]; row_seg int32 [n_rows] (bag of every row, murcl_row_segments). */
int murcl_attnpool_supported(int L, int D, int gated, int dtype);
int64_t murcl_attnpool_workspace(int64_t n_rows, int B, int L);
int murcl_attnpool_fwd(const void* h, const void* wab, const float* bab, const float* wc, const float* bc,
                       const int64_t* offsets, const int32_t* row_seg, int64_t n_rows, int B, int L, int D, int gated,
                       int inv_sqrt_n, int dtype, void* uv, float* s, float* p, float* M, float* stats, float* workspace,
                       void* stream);

/* Backward of p = post_scale*softmax(s), M = p^T h w.r.t. s:  ds[n,c] = p[n,c]*(dM[b,c].h[n] - k[b,c])
 * with k[b,c] = (dM[b,c].M[b,c]) / post_scale_b.  row_seg[n] = bag of row n. */
int murcl_pool_bwd_scores(const float* p, const void* h, const float* dM, const float* M, const int64_t* offsets,
                          const int32_t* row_seg, int64_t n_rows, int B, int C, int L, int inv_sqrt_n, int dtype,
                          float* ds, float* kbuf /* [B*C] scratch */, void* stream);

/* Direct term of the pooling backward: dh[n,:] (+)= sum_c p[n,c] * dM[b,c,:] (accumulate != 0 adds to
 * the existing contents).  The C == 1 case is also available fused in murcl_linear_bwd_input. */
int murcl_pool_bwd_direct(const float* p, const float* dM, const int32_t* row_seg, int64_t n_rows, int C, int L,
                          int dtype, void* dh, int accumulate, void* stream);

/* Backward through the score: given ds[N] and the saved activations uv, overwrites uv with the
 * gradient w.r.t. the pre-activations (tanh' / sigmoid' applied) and accumulates dwc[D], dbc[1] and - when
 * dpre_colsum != NULL - the column sums of the written gradient ([D] or [2D]: the bias gradient of the
 * attention projection).  All three are fp32 and must be zeroed by the caller.  drop_scale = 1/(1-p) when uv went
 * through murcl_dropout after its activations (clam.py:46-48), else 1 (or 0). */
int murcl_attn_score_bwd(void* uv, const float* wc, const float* ds, float* dwc, float* dbc, float* dpre_colsum,
                         int64_t N, int D, int gated, float drop_scale, int dtype, void* stream);

/* Fused attention pooling, backward (abmil.py:36-45, clam.py:37-60,170 differentiated; the matching pass of
 * murcl_attnpool_fwd): ONE pass over h [n_rows, L] computes ds[n] = p[n] * (dM[b].h[n] - (dM[b].M[b]) / post_scale_b),
 * overwrites the saved activations uv [n_rows, D*(1+gated)] with the gradient w.r.t. the pre-activations (tanh' /
 * sigmoid' and the gate applied) and accumulates dwc[D], dbc[1] and - when dpre_colsum != NULL - the column sums of
 * the written gradient (bias gradient of the attention projection); all three fp32, zeroed by the caller.  ds itself is
 * written only when the pointer is non-NULL.  Replaces murcl_pool_bwd_scores (C == 1) + murcl_attn_score_bwd: h is read
 * once and ds never travels through memory.  The direct term dh[n] += p[n] dM[b] stays fused in murcl_linear_bwd_input.
 * h, uv in dtype (fp32 or bf16); supported when murcl_attnpool_bwd_supported(...) != 0 (L <= 1024, L % 8 == 0,
 * D <= 512, D % 8 == 0).  drop_scale as for murcl_attn_score_bwd. */
int murcl_attnpool_bwd_supported(int L, int D, int gated, int dtype);
int murcl_attnpool_bwd(const void* h, void* uv, const float* p, const float* M, const float* dM, const float* wc,
                       const int64_t* offsets, const int32_t* row_seg, int64_t n_rows, int B, int L, int D, int gated,
                       int inv_sqrt_n, float drop_scale, int dtype, float* ds, float* dwc, float* dbc, float* dpre_colsum,
                       void* stream);

/* ---- (3) segmented reductions: clam.py:103-132 (top-k instance loss), dsmil.py:71-78 ---- */

/* Indices (global rows) of the k largest and k smallest p within each bag, ordered like
 * torch.topk (descending p / ascending p).  Fails with MURCL_EINVAL if a bag has < k rows
 * (the reference raises there too). */
int murcl_seg_topk_ends(const float* p, const int64_t* offsets, int B, int k, int32_t* top_idx, int32_t* bot_idx,
                        void* stream);

/* Per bag and class: global row of the first maximum of c[n,cls] (dsmil.py:71-73: row 0 of the
 * descending sort == arg-max). */
int murcl_seg_argmax(const float* c, const int64_t* offsets, int B, int C, int32_t* idx, void* stream);

/* DSMIL attention logits a[n,c] = q[n,:].q[crit[b,c],:] / sqrt(Dq) (dsmil.py:74-77). q fp32 [n_rows,Dq]. */
int murcl_dsmil_scores_fwd(const float* q, const int32_t* crit, const int32_t* row_seg, int64_t n_rows, int C, int Dq,
                           float* a, void* stream);
/* Backward: dq[n,:] = sum_c da[n,c]*q[crit]/sqrt(Dq), and dq[crit[b,c],:] += sum_n da[n,c]*q[n,:]/sqrt(Dq).
 * dq must be zeroed by the caller. */
int murcl_dsmil_scores_bwd(const float* q, const float* da, const int32_t* crit, const int32_t* row_seg,
                           const int64_t* offsets, int64_t n_rows, int B, int C, int Dq, float* dq, void* stream);

/* Gather rows: out[i,:] = h[idx[i],:] as fp32 (index_select, clam.py:108,110; dsmil.py:73). */
int murcl_gather_rows(const void* h, const int32_t* idx, int n_idx, int L, int dtype, float* out, void* stream);
/* Scatter-add fp32 rows into a gradient buffer of storage `dtype`: dh[idx[i],:] += rows[i,:]. */
int murcl_scatter_add_rows(void* dh, const int32_t* idx, int n_idx, int L, int dtype, const float* rows, void* stream);

/* CLAM instance-classifier tail (clam.py:112-118,126-131).  rows fp32 [R,L] are the gathered top-k
 * (+ bottom-k) instances; group g = rows [group_off[g], group_off[g+1]) scored by classifier
 * group_cls[g] (w [n_cls,2,L], bias [n_cls,2]); targets[R] in {0,1}.  loss[g] = mean CE of the group,
 * preds[R] = arg-max, dlogits[R,2] = (softmax - onehot)/rows_in_group (saved for the backward). */
int murcl_clam_inst_ce_fwd(const float* rows, const int32_t* targets, const int32_t* group_off, const int32_t* group_cls,
                           int G, const float* w, const float* bias, int L, float* loss, int32_t* preds, float* dlogits,
                           void* stream);
/* Given gloss[g] = d/d loss[g]: drows[R,L]; dw, db are ACCUMULATED (caller zeroes them). */
int murcl_clam_inst_ce_bwd(const float* rows, const float* dlogits, const float* gloss, const int32_t* group_off,
                           const int32_t* group_cls, int G, const float* w, int L, float* drows, float* dw, float* db,
                           void* stream);

/* ---- (4) NT-Xent: utils/losses.py:5-41, train_MuRCL.py:253,282 ------------------------- */

/* z fp32 [2B,d]: rows [0,B) view i, [B,2B) view j.  loss[1] = mean_a(LSE_{b!=a} s_ab - s_a,pos(a)),
 * s = cos/tau; dz[2B,d] = d loss / d z (NULL to skip); cos_pair[B] = cos(z_i[b], z_j[b]) (NULL to
 * skip).  workspace: murcl_ntxent_workspace(B, d) floats.  For d % 4 == 0, d <= 256 two kernels do everything (row
 * log-sum-exp + loss, then coefficient tiles x rows with the normalisation backward) and no [2B,2B] matrix is stored. */
int64_t murcl_ntxent_workspace(int B, int d);
int murcl_ntxent_fwd_bwd(const float* z, int B, int d, float temperature, float* loss, float* dz, float* cos_pair,
                         float* workspace, void* stream);
/* The same loss with the gradient restricted to a SLAB of samples: dz rows [b0, b0+nb) and [B+b0, B+b0+nb) are written,
 * all other rows of dz are left untouched.  loss and cos_pair cover the whole batch.  Data parallelism: every
 * rank evaluates the loss over the all-gathered global batch but needs only the gradient rows of its own samples, so the
 * O((2B)^2 d) gradient contraction shrinks by the number of ranks and no second collective is needed.  Needs d % 4 == 0,
 * d <= 256 and a 16-byte aligned z unless the slab is the whole batch. */
int murcl_ntxent_fwd_bwd_slab(const float* z, int B, int d, float temperature, int b0, int nb, float* loss, float* dz,
                              float* cos_pair, float* workspace, void* stream);

/* The two passes of the slab form as separate entry points, for ranks >= 4 where evaluating the log-sum-exp of ALL
 * 2B global rows on every rank (2B x 2B x d FMAs, redundantly) costs more than one more tiny collective:
 *   murcl_ntxent_lse_slab   log-sum-exp pass over the rows of the samples [b0, b0+nb) of both views only: writes
 *                           inv_norm[a] = 1/max(|z_a|, eps) and lse[a] at those 2*nb GLOBAL row indices a (other entries
 *                           untouched), loss_share[1] = (1/2B) sum over those rows of (lse_a - s_a,pos(a)) - the shares of
 *                           all ranks add up to the loss -, cos_pair[b] for b in [b0, b0+nb) (may be NULL);
 *   (the caller all-gathers inv_norm / lse of every rank's rows and the loss shares)
 *   murcl_ntxent_grad_slab  gradient rows of the same samples from the COMPLETE inv_norm[2B] and lse[2B].
 * workspace: murcl_ntxent_slab_workspace(B, d, nb) floats for either call, 16-byte aligned. */
int64_t murcl_ntxent_slab_workspace(int B, int d, int nb);
int murcl_ntxent_lse_slab(const float* z, int B, int d, float temperature, int b0, int nb, float* inv_norm, float* lse,
                          float* loss_share, float* cos_pair, float* workspace, void* stream);
int murcl_ntxent_grad_slab(const float* z, int B, int d, float temperature, int b0, int nb, const float* inv_norm,
                           const float* lse, float* dz, float* workspace, void* stream);

/* ---- (5) recurrent heads: rlmil.py:66-97 (actor), :187-220 (Full_layer) --------------- */

/* GRU cell from the two gate pre-activations gi = x W_ih^T + b_ih, gh = h W_hh^T + b_hh
 * (both [B,3H], gate order r,z,n).  gates_out [B,3H] saves (r,z,n) for the backward (may be NULL). */
int murcl_gru_cell_fwd(const float* gi, const float* gh, const float* h_prev, float* h_new, float* gates_out, int B,
                       int H, void* stream);
/* Given dh_new: dgi, dgh [B,3H] and dh_prev [B,H] (the direct z*dh term; the caller adds dgh.W_hh). */
int murcl_gru_cell_bwd(const float* dh_new, const float* gates, const float* gh, const float* h_prev, float* dgi,
                       float* dgh, float* dh_prev, int B, int H, void* stream);

/* The same cell for the recurrent-head tape (murcl_b200/headtape.py), which differentiates Full_layer (rlmil.py:208-220)
 * over all T x 2 calls of an optimiser step in ONE batched pass.  _fwd_tape also writes h_new in the GEMM storage type
 * `dtype` (h_new_s and a second copy h_new_s2, either may be NULL) and always writes the gates.  _bwd_tape takes dh as the sum of up to three addends
 * (dh_a, dh_b in `dtype`; dh_c fp32; any may be NULL) and writes the gate gradients dgi, dgh [B, 3H] in `dtype` - the
 * operands of the batched weight-gradient GEMMs - plus dh_prev = z * dh (fp32, may be NULL). */
int murcl_gru_cell_fwd_tape(const float* gi, const float* gh, const float* h_prev, float* h_new, void* h_new_s, void* h_new_s2,
                            float* gates, int B, int H, int dtype, void* stream);
int murcl_gru_cell_bwd_tape(const void* dh_a, const void* dh_b, const float* dh_c, const float* gates, const float* gh,
                            const float* h_prev, void* dgi, void* dgh, float* dh_prev, int B, int H, int dtype, void* stream);
/* Actor head: mean = sigmoid(logits); a = clip(mean + std*eps, 0, 1); logprob of a under
 * N(mean, std^2 I) (rlmil.py:82-90). */
int murcl_actor_head(const float* logits, const float* eps, float std, float* action, float* logprob, float* mean,
                     int B, int K, void* stream);

/* ---- helpers ---------------------------------------------------------------------------- */
int murcl_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream);
/* row_seg[n] = bag id of row n, from offsets. */
int murcl_row_segments(const int64_t* offsets, int B, int32_t* row_seg, void* stream);
/* out[n] = (fp32) column sums of a[M,N] (storage dtype). */
int murcl_colsum(const void* a, int64_t M, int N, int dtype, float* out, void* stream);
/* In-place inverted dropout: x[i] = keep_i ? x[i]/(1-p) : 0 with keep_i from a counter-based hash of (seed, i);
 * the 64-bit seed is read from device memory (fresh per CUDA-graph replay). */
int murcl_dropout(void* x, int64_t n, float p, const int64_t* seed_dev, int dtype, void* stream);
/* dz = dy * (y > 0) elementwise (ReLU backward), same storage dtype. */
int murcl_relu_bwd(const void* dy, const void* y, void* dz, int64_t n, int dtype, void* stream);

/* ---- optimiser step over the flat parameter arena: train_MuRCL.py:154-171 (torch.optim.Adam), :296 (step) ---- */

/* One fused Adam step (L2 weight decay folded into the gradient, bias correction, no amsgrad - torch.optim.Adam's
 * defaults as the reference uses them) over n contiguous fp32 parameters:
 *   g' = grad_scale * grad + weight_decay * param;  m = m + (1-beta1)(g' - m);  v = beta2 v + (1-beta2) g'^2;
 *   param -= lr / (1-beta1^t) * m / (sqrt(v) / sqrt(1-beta2^t) + eps),   t = state[0] + 1.
 * shadow_bf16 (may be NULL) receives the bf16 copy of the updated parameters that the tcgen05 GEMMs read.
 * The hyper-parameters are doubles: 1 - beta and the bias corrections are formed in double and rounded once, as torch does.
 * lr_dev (may be NULL) overrides lr from device memory (a schedule that changes between CUDA-graph replays).
 * state: int64[2] on the device, zero-initialised: [0] = steps taken (advanced by the kernel), [1] = scratch. */
int murcl_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16, int64_t n,
                    double lr, double beta1, double beta2, double eps, double weight_decay, double grad_scale,
                    const float* lr_dev, int64_t* state, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MURCL_B200_H */
