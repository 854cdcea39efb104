// Fused Adam over the flat parameter arena (murcl_b200/arena.py): the optimiser step of train_MuRCL.py:154-171,296
// (torch.optim.Adam with L2 weight decay folded into the gradient) as ONE pass over four flat fp32 buffers, which also
// refreshes the bf16 shadow the tcgen05 GEMMs read.  torch's multi-tensor Adam is ~19 launches and ~10 passes over the same
// memory (260 us per step on the 6 M-parameter arena); this is one launch at the HBM roof:
//   per parameter: read p, g, m, v (16 B), write p, m, v (12 B) + shadow (2 B) = 30 B.
// The step counter lives on the device (the whole training step replays as a CUDA graph): every thread reads it, the LAST
// block to retire bumps it.
#include "common.cuh"

namespace murcl {

struct AdamArgs {
  double lr, beta1, beta2;                 // bias corrections are evaluated in double, as torch does on the host
  float omb1, omb2, b2, eps, weight_decay, grad_scale;     // 1 - beta rounded from the DOUBLE difference (1 - 0.999f is off by 1e-5)
};

__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, __nv_bfloat16* __restrict__ shadow, int64_t n,
                                                        AdamArgs a, const float* __restrict__ lr_dev, int64_t* __restrict__ state) {
  // state[0] = steps taken so far, state[1] = retired-block ticket of this launch
  __shared__ float s_coef[2];
  if (threadIdx.x == 0) {
    const double t = (double)(state[0] + 1);
    const double bc1 = 1.0 - pow(a.beta1, t);
    const double bc2 = 1.0 - pow(a.beta2, t);
    const double lr = lr_dev != nullptr ? (double)*lr_dev : a.lr;
    s_coef[0] = (float)(lr / bc1);                  // step size
    s_coef[1] = (float)(1.0 / sqrt(bc2));           // 1 / sqrt(bias_correction2)
  }
  __syncthreads();
  const float step_size = s_coef[0], inv_sqrt_bc2 = s_coef[1];
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 4 <= n) {
    const float4 p4 = *reinterpret_cast<const float4*>(p + i);
    float4 g4 = *reinterpret_cast<const float4*>(g + i);
    float4 m4 = *reinterpret_cast<const float4*>(m + i);
    float4 v4 = *reinterpret_cast<const float4*>(v + i);
    float pp[4] = {p4.x, p4.y, p4.z, p4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w},
          vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = fmaf(a.weight_decay, pp[j], gg[j] * a.grad_scale);
      mm[j] = fmaf(a.omb1, gr - mm[j], mm[j]);                            // lerp, as torch does
      vv[j] = fmaf(a.omb2, gr * gr, a.b2 * vv[j]);
      const float denom = sqrtf(vv[j]) * inv_sqrt_bc2 + a.eps;
      pp[j] -= step_size * (mm[j] / denom);
    }
    *reinterpret_cast<float4*>(p + i) = make_float4(pp[0], pp[1], pp[2], pp[3]);
    *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    if (shadow != nullptr) store4(shadow + i, make_float4(pp[0], pp[1], pp[2], pp[3]));
  } else {
    for (int64_t j = i; j < n; ++j) {
      const float gr = fmaf(a.weight_decay, p[j], g[j] * a.grad_scale);
      const float mj = fmaf(a.omb1, gr - m[j], m[j]);
      const float vj = fmaf(a.omb2, gr * gr, a.b2 * v[j]);
      const float pj = p[j] - step_size * (mj / (sqrtf(vj) * inv_sqrt_bc2 + a.eps));
      p[j] = pj; m[j] = mj; v[j] = vj;
      if (shadow != nullptr) shadow[j] = __float2bfloat16_rn(pj);
    }
  }
  // the last block to retire advances the step counter (every block has read it by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long ticket = atomicAdd(reinterpret_cast<unsigned long long*>(state + 1), 1ull);
    if (ticket == (unsigned long long)gridDim.x - 1ull) {
      state[1] = 0;
      state[0] = state[0] + 1;
      __threadfence();
    }
  }
}

}  // namespace murcl

extern "C" {

int murcl_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16, int64_t n, double lr,
                    double beta1, double beta2, double eps, double weight_decay, double grad_scale, const float* lr_dev,
                    int64_t* state, void* stream) {
  using namespace murcl;
  MURCL_REQUIRE(param && grad && exp_avg && exp_avg_sq && state, "adam_step: null pointer");
  MURCL_REQUIRE(n >= 0, "adam_step: negative length");
  MURCL_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
                  reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0 && (reinterpret_cast<uintptr_t>(shadow_bf16) & 7) == 0,
                "adam_step: buffers must be 16-byte aligned (shadow: 8)");
  MURCL_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0, "adam_step: bad hyper-parameters");
  if (n == 0) return MURCL_OK;
  AdamArgs a{lr, beta1, beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)beta2, (float)eps, (float)weight_decay,
             (float)grad_scale};
  const int grid = ceil_div((n + 3) / 4, 256);
  adam_step_kernel<<<grid, 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, static_cast<__nv_bfloat16*>(shadow_bf16), n,
                                                        a, lr_dev, state);
  return check_launch("adam_step_kernel");
}

}  // extern "C"
