// C-ABI dispatch for the dense layers: picks the tcgen05 path (gemm_tc.cu) when the operands are
// bf16 and the shape fits its tiles, otherwise the exact-fp32 SIMT path (gemm_simt.cu).
#include "common.cuh"

namespace murcl {
int simt_linear_fwd(const void*, const void*, const float*, void*, int64_t, int, int, int, int, int, cudaStream_t,
                    float* ws = nullptr, int64_t ws_floats = 0);
int simt_linear_bwd_input(const void*, const void*, void*, int64_t, int, int, const void*, const float*, const float*,
                          const int32_t*, float, int, cudaStream_t, float* ws = nullptr, int64_t ws_floats = 0,
                          const unsigned long long* relu_bits = nullptr);
int launch_relu_bits(const void* y, int64_t M, int N, int dtype, unsigned long long* bits, cudaStream_t st);
int64_t simt_linear_bwd_weight_workspace(int64_t, int, int);
int simt_linear_bwd_weight(const void*, const void*, float*, int64_t, int, int, int, float*, cudaStream_t, int accumulate);

bool tc_fwd_supported(int64_t M, int N, int K, int dtype, int out_dtype);
bool tc_bwd_input_supported(int64_t M, int N, int K, int dtype);
bool tc_bwd_weight_supported(int64_t M, int N, int K, int dtype);
int tc_linear_fwd(const void*, const void*, const float*, void*, int64_t, int, int, int, int, unsigned long long*, cudaStream_t);
int tc_linear_bwd_input(const void*, const void*, void*, int64_t, int, int, const void*, const float*, const float*,
                        const int32_t*, float*, float, const unsigned long long*, cudaStream_t);
int64_t tc_linear_bwd_weight_workspace(int64_t, int, int);
int tc_linear_bwd_weight(const void*, const void*, float*, int64_t, int, int, float*, cudaStream_t, int accumulate);

bool tc_bwd_input_accum_supported(int64_t M, int N, int K, int dtype);
int tc_linear_bwd_input_accum(const void*, const void*, float*, int64_t, int, int, cudaStream_t);

int colsum_impl(const void* a, int64_t M, int N, int dtype, float* out, cudaStream_t st, int accumulate = 0);
bool tc_split_supported(int64_t M, int N, int K);
int tc_split_fwd(const void*, const void*, const float*, float*, int64_t, int, int, int, int, int64_t, int64_t, cudaStream_t);
int tc_split_bwd_input(const void*, const void*, float*, int64_t, int, int, const float*, const float*, const int32_t*, float,
                       const unsigned long long*, int, int64_t, int64_t, cudaStream_t);
int64_t tc_split_bwd_weight_workspace(int64_t, int, int);
int tc_split_bwd_weight(const void*, const void*, float*, int64_t, int, int, int, int64_t, float*, cudaStream_t, int);
}  // namespace murcl

using namespace murcl;

static bool valid_dtype(int d) { return d == MURCL_F32 || d == MURCL_BF16; }

extern "C" {

int murcl_linear_fwd(const void* x, const void* w, const float* bias, void* y, int64_t M, int N, int K, int act,
                     int dtype, int out_dtype, int backend, uint64_t* relu_bits, void* stream) {
  MURCL_REQUIRE(x && w && y, "linear_fwd: null pointer");
  MURCL_REQUIRE(M >= 0 && N > 0 && K > 0, "linear_fwd: bad shape M=%lld N=%d K=%d", (long long)M, N, K);
  MURCL_REQUIRE(valid_dtype(dtype) && valid_dtype(out_dtype), "linear_fwd: bad dtype");
  MURCL_REQUIRE(act >= MURCL_ACT_NONE && act <= MURCL_ACT_TANH_SIGMOID, "linear_fwd: bad activation %d", act);
  MURCL_REQUIRE(act != MURCL_ACT_TANH_SIGMOID || (N % 2) == 0, "linear_fwd: gated activation needs even N");
  MURCL_REQUIRE(relu_bits == nullptr || (act == MURCL_ACT_RELU && N % 64 == 0 && dtype == out_dtype),
                "linear_fwd: the ReLU bit mask needs act=RELU, N %% 64 == 0 and matching storage types");
  if (M == 0) return MURCL_OK;
  const bool tc_ok = tc_fwd_supported(M, N, K, dtype, out_dtype);
  if (backend == MURCL_GEMM_TCGEN05 && !tc_ok) {
    set_error("linear_fwd: tcgen05 path does not take M=%lld N=%d K=%d dtype=%d->%d", (long long)M, N, K, dtype, out_dtype);
    return MURCL_EUNSUPPORTED;
  }
  if (backend != MURCL_GEMM_SIMT && tc_ok)
    return tc_linear_fwd(x, w, bias, y, M, N, K, act, out_dtype, reinterpret_cast<unsigned long long*>(relu_bits), as_stream(stream));
  int rc = simt_linear_fwd(x, w, bias, y, M, N, K, act, dtype, out_dtype, as_stream(stream));
  if (rc != MURCL_OK || relu_bits == nullptr) return rc;
  return launch_relu_bits(y, M, N, out_dtype, reinterpret_cast<unsigned long long*>(relu_bits), as_stream(stream));
}

int murcl_linear_bwd_input(const void* dy, const void* w, void* dx, int64_t M, int N, int K, const void* relu_src,
                           const float* row_scale, const float* row_vec, const int32_t* row_seg, float* col_sum,
                           float out_scale, const uint64_t* relu_bits, int dtype, int backend, void* stream) {
  MURCL_REQUIRE(dy && w && dx, "linear_bwd_input: null pointer");
  MURCL_REQUIRE(M >= 0 && N > 0 && K > 0, "linear_bwd_input: bad shape");
  MURCL_REQUIRE(valid_dtype(dtype), "linear_bwd_input: bad dtype");
  MURCL_REQUIRE((row_scale == nullptr) == (row_vec == nullptr) && (row_scale == nullptr) == (row_seg == nullptr),
                "linear_bwd_input: row_scale, row_vec and row_seg must be given together");
  if (M == 0) return MURCL_OK;
  if (out_scale == 0.f) out_scale = 1.f;
  MURCL_REQUIRE(out_scale == 1.f || relu_src != nullptr || relu_bits != nullptr,
                "linear_bwd_input: out_scale is the dropout factor of a masked ReLU");
  MURCL_REQUIRE(relu_bits == nullptr || K % 64 == 0, "linear_bwd_input: the ReLU bit mask needs K %% 64 == 0");
  const unsigned long long* bits = reinterpret_cast<const unsigned long long*>(relu_bits);
  const bool tc_ok = tc_bwd_input_supported(M, N, K, dtype);
  if (backend == MURCL_GEMM_TCGEN05 && !tc_ok) {
    set_error("linear_bwd_input: tcgen05 path does not take M=%lld N=%d K=%d dtype=%d", (long long)M, N, K, dtype);
    return MURCL_EUNSUPPORTED;
  }
  if (backend != MURCL_GEMM_SIMT && tc_ok)
    return tc_linear_bwd_input(dy, w, dx, M, N, K, relu_src, row_scale, row_vec, row_seg, col_sum, out_scale, bits, as_stream(stream));
  int rc = simt_linear_bwd_input(dy, w, dx, M, N, K, relu_src, row_scale, row_vec, row_seg, out_scale, dtype, as_stream(stream),
                                 nullptr, 0, bits);
  if (rc != MURCL_OK || col_sum == nullptr) return rc;
  return colsum_impl(dx, M, K, dtype, col_sum, as_stream(stream), 1);    // same sums, separate pass; ADDS like the fused epilogue
}

int64_t murcl_linear_bwd_weight_workspace(int64_t M, int N, int K) {
  const int64_t a = simt_linear_bwd_weight_workspace(M, N, K), b = tc_linear_bwd_weight_workspace(M, N, K);
  return a > b ? a : b;
}

int murcl_linear_bwd_weight(const void* dy, const void* x, float* dw, float* db, int64_t M, int N, int K, int dtype,
                            int backend, float* workspace, int accumulate, void* stream) {
  MURCL_REQUIRE(dy && x && dw, "linear_bwd_weight: null pointer");
  MURCL_REQUIRE(M >= 0 && N > 0 && K > 0, "linear_bwd_weight: bad shape");
  MURCL_REQUIRE(valid_dtype(dtype), "linear_bwd_weight: bad dtype");
  cudaStream_t st = as_stream(stream);
  if (M == 0) {
    if (accumulate) return MURCL_OK;
    MURCL_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)N * K, st));
    if (db) MURCL_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * (size_t)N, st));
    return MURCL_OK;
  }
  if (db) {
    int rc = colsum_impl(dy, M, N, dtype, db, st, accumulate);
    if (rc != MURCL_OK) return rc;
  }
  const bool tc_ok = tc_bwd_weight_supported(M, N, K, dtype);
  if (backend == MURCL_GEMM_TCGEN05 && !tc_ok) {
    set_error("linear_bwd_weight: tcgen05 path does not take M=%lld N=%d K=%d dtype=%d", (long long)M, N, K, dtype);
    return MURCL_EUNSUPPORTED;
  }
  if (backend != MURCL_GEMM_SIMT && tc_ok) return tc_linear_bwd_weight(dy, x, dw, M, N, K, workspace, st, accumulate);
  return simt_linear_bwd_weight(dy, x, dw, M, N, K, dtype, workspace, st, accumulate);
}

/* ---- exact-fp32 dense layers on the bf16 tensor cores (split precision) ---- */
int murcl_linear_split_supported(int64_t M, int N, int K) { return tc_split_supported(M, N, K) ? 1 : 0; }

int murcl_linear_fwd_split(const void* xp, const void* wp, const float* bias, float* y, int64_t M, int N, int K, int act,
                           int planes, int64_t x_plane_rows, int64_t w_plane_rows, uint64_t* relu_bits, void* stream) {
  MURCL_REQUIRE(xp && wp && y, "linear_fwd_split: null pointer");
  MURCL_REQUIRE(tc_split_supported(M, N, K) && (planes == 2 || planes == 3), "linear_fwd_split: unsupported shape M=%lld N=%d K=%d",
                (long long)M, N, K);
  MURCL_REQUIRE(x_plane_rows >= M && w_plane_rows >= N, "linear_fwd_split: plane pitch smaller than the tensor");
  MURCL_REQUIRE(act >= MURCL_ACT_NONE && act <= MURCL_ACT_TANH_SIGMOID, "linear_fwd_split: bad activation %d", act);
  MURCL_REQUIRE(relu_bits == nullptr || (act == MURCL_ACT_RELU && N % 64 == 0), "linear_fwd_split: bit mask needs ReLU and N %% 64 == 0");
  int rc = tc_split_fwd(xp, wp, bias, y, M, N, K, act, planes, x_plane_rows, w_plane_rows, as_stream(stream));
  if (rc != MURCL_OK || relu_bits == nullptr) return rc;
  return launch_relu_bits(y, M, N, MURCL_F32, reinterpret_cast<unsigned long long*>(relu_bits), as_stream(stream));
}

int murcl_linear_bwd_input_split(const void* dyp, const void* wp, float* dx, int64_t M, int N, int K, const float* row_scale,
                                 const float* row_vec, const int32_t* row_seg, float* col_sum, float out_scale,
                                 const uint64_t* relu_bits, int planes, int64_t dy_plane_rows, int64_t w_plane_rows,
                                 void* stream) {
  MURCL_REQUIRE(dyp && wp && dx, "linear_bwd_input_split: null pointer");
  MURCL_REQUIRE(tc_split_supported(M, K, N) && (planes == 2 || planes == 3), "linear_bwd_input_split: unsupported shape M=%lld N=%d K=%d",
                (long long)M, N, K);
  MURCL_REQUIRE((row_scale == nullptr) == (row_vec == nullptr) && (row_scale == nullptr) == (row_seg == nullptr),
                "linear_bwd_input_split: row_scale, row_vec and row_seg must be given together");
  MURCL_REQUIRE(dy_plane_rows >= M && w_plane_rows >= N && w_plane_rows % 64 == 0, "linear_bwd_input_split: bad plane pitch");
  if (out_scale == 0.f) out_scale = 1.f;
  MURCL_REQUIRE(out_scale == 1.f || relu_bits != nullptr, "linear_bwd_input_split: out_scale is the dropout factor of a masked ReLU");
  int rc = tc_split_bwd_input(dyp, wp, dx, M, N, K, row_scale, row_vec, row_seg, out_scale,
                              reinterpret_cast<const unsigned long long*>(relu_bits), planes, dy_plane_rows, w_plane_rows,
                              as_stream(stream));
  if (rc != MURCL_OK || col_sum == nullptr) return rc;
  return colsum_impl(dx, M, K, MURCL_F32, col_sum, as_stream(stream), 1);
}

int64_t murcl_linear_bwd_weight_split_workspace(int64_t M, int N, int K) { return tc_split_bwd_weight_workspace(M, N, K); }

int murcl_linear_bwd_weight_split(const void* dyp, const void* xp, float* dw, int64_t M, int N, int K, int planes,
                                  int64_t plane_rows, float* workspace, int accumulate, void* stream) {
  MURCL_REQUIRE(dyp && xp && dw && workspace, "linear_bwd_weight_split: null pointer");
  MURCL_REQUIRE(M >= 1024 && N >= 128 && N % 64 == 0 && K >= 128 && K % 64 == 0 && (planes == 2 || planes == 3),
                "linear_bwd_weight_split: unsupported shape M=%lld N=%d K=%d", (long long)M, N, K);
  MURCL_REQUIRE(plane_rows >= M && plane_rows % 64 == 0, "linear_bwd_weight_split: plane pitch must be >= M and a multiple of 64");
  return tc_split_bwd_weight(dyp, xp, dw, M, N, K, planes, plane_rows, workspace, as_stream(stream), accumulate);
}

int murcl_linear_bwd_input_accum_supported(int64_t M, int N, int K, int dtype) {
  return tc_bwd_input_accum_supported(M, N, K, dtype) ? 1 : 0;
}

int murcl_linear_bwd_input_accum(const void* dy, const void* w, float* dx, int64_t M, int N, int K, int dtype, void* stream) {
  MURCL_REQUIRE(dy && w && dx, "linear_bwd_input_accum: null pointer");
  MURCL_REQUIRE(M >= 0 && N > 0 && K > 0, "linear_bwd_input_accum: bad shape M=%lld N=%d K=%d", (long long)M, N, K);
  if (M == 0) return MURCL_OK;
  if (!tc_bwd_input_accum_supported(M, N, K, dtype)) {
    set_error("linear_bwd_input_accum: needs bf16 operands, M >= 64, N >= 64, K >= 128 (N, K %% 8 == 0); got M=%lld N=%d K=%d dtype=%d",
              (long long)M, N, K, dtype);
    return MURCL_EUNSUPPORTED;
  }
  return tc_linear_bwd_input_accum(dy, w, dx, M, N, K, as_stream(stream));
}

}  // extern "C"
