// Fused Adam step over one FLAT fp32 parameter / gradient bucket (behavioural_cloning.py:63-67: th.optim.Adam(lr, weight_decay),
// i.e. L2-style decay added to the gradient, not AdamW) -- groundwork for the BC step: gradients of
// every parameter live in one flat buffer so that data parallelism is ONE NCCL all-reduce, and the optimizer is one kernel.
//   g      = grad * grad_scale + weight_decay * p          (grad_scale = 1 / world_size after a sum all-reduce)
//   m      = beta1 * m + (1 - beta1) * g
//   v      = beta2 * v + (1 - beta2) * g * g
//   p     -= lr / (1 - beta1^t) * m / (sqrt(v / (1 - beta2^t)) + eps)      (torch.optim.Adam, amsgrad = False)
#pragma once
#include "common.cuh"

namespace vpt {

__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                          float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
                                                          float weight_decay, float grad_scale, float bc1, float bc2_sqrt) {
    // bc1 = 1 - beta1^t ; bc2_sqrt = sqrt(1 - beta2^t)   (computed on the host in double precision)
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
        float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        float* pa = &pp.x;
        const float* ga = &gg.x;
        float* ma = &mm.x;
        float* va = &vv.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gr = fmaf(ga[j], grad_scale, weight_decay * pa[j]);
            ma[j] = fmaf(beta1, ma[j], (1.f - beta1) * gr);
            va[j] = fmaf(beta2, va[j], (1.f - beta2) * gr * gr);
            const float denom = sqrtf(va[j]) / bc2_sqrt + eps;
            pa[j] -= (lr / bc1) * (ma[j] / denom);
        }
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    // tail (n not a multiple of 4)
    const long long t = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) {
        const float gr = fmaf(g[t], grad_scale, weight_decay * p[t]);
        m[t] = fmaf(beta1, m[t], (1.f - beta1) * gr);
        v[t] = fmaf(beta2, v[t], (1.f - beta2) * gr * gr);
        p[t] -= (lr / bc1) * (m[t] / (sqrtf(v[t]) / bc2_sqrt + eps));
    }
}

}  // namespace vpt

extern "C" int vpt_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                             float beta2, float eps, float weight_decay, float grad_scale, int32_t step, void* stream) {
    using namespace vpt;
    VPT_CHECK(params && grads && exp_avg && exp_avg_sq && n > 0 && step >= 1, "vpt_adam_step: bad arguments");
    VPT_CHECK((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0, "vpt_adam_step: buffers must be 16-byte aligned");
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    long long blocks = (n / 4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    adam_step_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                                                        grad_scale, (float)bc1, (float)sqrt(bc2));
    VPT_LAUNCH_CHECK();
    return VPT_OK;
}
