// Fused Adam over a flat fp32 parameter array (hash table or MLP blob).
//
// Replaces the torch.optim.Adam step the reference engine runs on feat_pool and the MLPs
// (reference nerfstudio/engine/optimizers.py:125-137; lr 1e-2, eps 1e-15,
// gfnerf/config.py:132-135; betas (0.9, 0.99) Hash3DAnchored.cpp:146-150) plus, in the same
// pass, the gradient zero-fill of the next step, the DDP mean (grad / world) and the
// fp32 -> fp16 re-cast of the table the reference does at the start of every forward
// (Hash3DAnchored_cuda.cu:185).  28 B/param of HBM traffic instead of ~60.
#include "common.cuh"

namespace gf {

// torch.optim.Adam's bias corrections for step t (computed in double like torch's Python scalars)
__host__ __device__ inline void bias_corrections(long long t, float lr, float beta1, float beta2, float* step_size,
                                                 float* inv_bc2_sqrt) {
  const double bc1 = 1.0 - pow((double)beta1, (double)t);
  const double bc2 = 1.0 - pow((double)beta2, (double)t);
  *step_size = (float)((double)lr / bc1);
  *inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
}

// the optimizer's step count lives on the device when the NaN guard may skip a step: a skipped step must not advance
// the bias correction (the reference skips optimizer.step() altogether, trainer.py:416-426)
__global__ void adam_count_kernel(long long* d_step, const int* skip_flag) {
  if (!(skip_flag && *skip_flag)) *d_step += 1;
}

template <bool SHADOW>
__global__ void __launch_bounds__(256)
adam_kernel(int64_t n4, float4* __restrict__ param, float4* __restrict__ grad, float4* __restrict__ m,
            float4* __restrict__ v, uint2* __restrict__ shadow, float step_size, float beta1, float beta2,
            float inv_bc2_sqrt, float eps, float inv_div, int zero_grad, const int* __restrict__ skip_flag,
            const long long* __restrict__ d_step, float lr, float beta1_d, float beta2_d) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (skip_flag && *skip_flag) {  // a NaN gradient was found: no update (trainer.py:416-426), only the zero-fill
    if (zero_grad)
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride)
        grad[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  if (d_step) bias_corrections(*d_step + 1, lr, beta1_d, beta2_d, &step_size, &inv_bc2_sqrt);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 p = param[i], g = grad[i], mm = m[i], vv = v[i];
    float* pp = &p.x; float* gp = &g.x; float* mp = &mm.x; float* vp = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float gk = gp[k] * inv_div;
      mp[k] = mp[k] + (gk - mp[k]) * (1.f - beta1);           // exp_avg.lerp_(grad, 1-beta1)
      vp[k] = vp[k] * beta2 + (1.f - beta2) * gk * gk;         // exp_avg_sq.mul_(b2).addcmul_(g,g,1-b2)
      const float denom = sqrtf(vp[k]) * inv_bc2_sqrt + eps;   // (sqrt(v)/sqrt(bc2)).add_(eps)
      pp[k] = pp[k] - step_size * (mp[k] / denom);             // param.addcdiv_(m, denom, -lr/bc1)
    }
    param[i] = p; m[i] = mm; v[i] = vv;
    if (zero_grad) grad[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (SHADOW) {
      __half2 lo = __floats2half2_rn(p.x, p.y), hi = __floats2half2_rn(p.z, p.w);
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&lo);
      o.y = *reinterpret_cast<uint32_t*>(&hi);
      shadow[i] = o;
    }
  }
}

__global__ void adam_tail_kernel(int64_t from, int64_t n, float* param, float* grad, float* m, float* v,
                                 __half* shadow, float step_size, float beta1, float beta2, float inv_bc2_sqrt,
                                 float eps, float inv_div, int zero_grad, const int* skip_flag,
                                 const long long* d_step, float lr) {
  const int64_t i = from + threadIdx.x;
  if (i >= n) return;
  if (skip_flag && *skip_flag) {
    if (zero_grad) grad[i] = 0.f;
    return;
  }
  if (d_step) bias_corrections(*d_step + 1, lr, beta1, beta2, &step_size, &inv_bc2_sqrt);
  const float gk = grad[i] * inv_div;
  const float mk = m[i] + (gk - m[i]) * (1.f - beta1);
  const float vk = v[i] * beta2 + (1.f - beta2) * gk * gk;
  const float denom = sqrtf(vk) * inv_bc2_sqrt + eps;
  const float pk = param[i] - step_size * (mk / denom);
  param[i] = pk; m[i] = mk; v[i] = vk;
  if (zero_grad) grad[i] = 0.f;
  if (shadow) shadow[i] = __float2half_rn(pk);
}

// flag |= any(isnan(x)): the trainer's NaN-gradient scan (nerfstudio/engine/trainer.py:416-423) without its host sync
__global__ void __launch_bounds__(256) nan_scan_kernel(int64_t n, const float* __restrict__ x, int* __restrict__ flag) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad = false;
  const int64_t n4 = n / 4;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(x4 + i);
    bad |= isnan(v.x) || isnan(v.y) || isnan(v.z) || isnan(v.w);
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) bad |= isnan(x[i]);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

}  // namespace gf

using namespace gf;

extern "C" int gf_grad_nan_scan(int64_t n, const float* grad, int32_t* flag, void* stream) {
  GF_REQUIRE(n >= 0 && flag, "gf_grad_nan_scan: bad arguments");
  if (n == 0) return GF_OK;
  GF_REQUIRE(grad != nullptr, "gf_grad_nan_scan: null pointer");
  GF_REQUIRE((reinterpret_cast<uintptr_t>(grad) & 15) == 0, "gf_grad_nan_scan: grad must be 16-byte aligned");
  nan_scan_kernel<<<stride_grid(n / 4 + 1, 256, 8, 2), 256, 0, (cudaStream_t)stream>>>(n, grad, flag);
  return check_launch("nan_scan_kernel");
}

static int adam_impl(int64_t n, float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_f16, float lr,
                     float beta1, float beta2, float eps, int64_t step, float grad_div, int zero_grad,
                     const int32_t* skip_flag, void* stream, int64_t* d_step = nullptr);

extern "C" int gf_adam_step(int64_t n, float* param, float* grad, float* exp_avg, float* exp_avg_sq,
                            void* shadow_f16, float lr, float beta1, float beta2, float eps, int64_t step,
                            float grad_div, int zero_grad, void* stream) {
  return adam_impl(n, param, grad, exp_avg, exp_avg_sq, shadow_f16, lr, beta1, beta2, eps, step, grad_div, zero_grad,
                   nullptr, stream);
}

extern "C" int gf_adam_step_guarded(int64_t n, float* param, float* grad, float* exp_avg, float* exp_avg_sq,
                                    void* shadow_f16, float lr, float beta1, float beta2, float eps, int64_t step,
                                    float grad_div, int zero_grad, const int32_t* skip_flag, void* stream) {
  return adam_impl(n, param, grad, exp_avg, exp_avg_sq, shadow_f16, lr, beta1, beta2, eps, step, grad_div, zero_grad,
                   skip_flag, stream);
}

extern "C" int gf_adam_step_counted(int64_t n, float* param, float* grad, float* exp_avg, float* exp_avg_sq,
                                    void* shadow_f16, float lr, float beta1, float beta2, float eps, int64_t* d_step,
                                    float grad_div, int zero_grad, const int32_t* skip_flag, void* stream) {
  GF_REQUIRE(d_step != nullptr, "gf_adam_step_counted: null step counter");
  int rc = adam_impl(n, param, grad, exp_avg, exp_avg_sq, shadow_f16, lr, beta1, beta2, eps, 1, grad_div, zero_grad,
                     skip_flag, stream, d_step);
  if (rc) return rc;
  adam_count_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((long long*)d_step, skip_flag);
  return check_launch("adam_count_kernel");
}

static int adam_impl(int64_t n, float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_f16, float lr,
                     float beta1, float beta2, float eps, int64_t step, float grad_div, int zero_grad,
                     const int32_t* skip_flag, void* stream, int64_t* d_step_) {
  GF_REQUIRE(n >= 0 && step >= 1 && grad_div != 0.f, "gf_adam_step: bad arguments");
  if (n == 0) return GF_OK;
  GF_REQUIRE(param && grad && exp_avg && exp_avg_sq, "gf_adam_step: null pointer");
  const long long* d_step = (const long long*)d_step_;
  float step_size, inv_bc2_sqrt;
  bias_corrections(step, lr, beta1, beta2, &step_size, &inv_bc2_sqrt);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n4 = n / 4;
  if (n4 > 0) {
    const int grid = stride_grid(n4, 256, 8, 2);
    if (shadow_f16)
      adam_kernel<true><<<grid, 256, 0, st>>>(n4, (float4*)param, (float4*)grad, (float4*)exp_avg,
                                              (float4*)exp_avg_sq, (uint2*)shadow_f16, step_size, beta1, beta2,
                                              inv_bc2_sqrt, eps, 1.f / grad_div, zero_grad, skip_flag, d_step, lr, beta1,
                                              beta2);
    else
      adam_kernel<false><<<grid, 256, 0, st>>>(n4, (float4*)param, (float4*)grad, (float4*)exp_avg,
                                               (float4*)exp_avg_sq, nullptr, step_size, beta1, beta2, inv_bc2_sqrt,
                                               eps, 1.f / grad_div, zero_grad, skip_flag, d_step, lr, beta1,
                                               beta2);
    int rc = check_launch("adam_kernel");
    if (rc) return rc;
  }
  if (n4 * 4 < n) {
    adam_tail_kernel<<<1, 32, 0, st>>>(n4 * 4, n, param, grad, exp_avg, exp_avg_sq, (__half*)shadow_f16, step_size,
                                       beta1, beta2, inv_bc2_sqrt, eps, 1.f / grad_div, zero_grad, skip_flag, d_step,
                                       lr);
    return check_launch("adam_tail_kernel");
  }
  return GF_OK;
}
