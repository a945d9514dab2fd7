// Adam update arithmetic (fused optimizer step), host/device.
//
// Restates torch.optim.Adam's update as the reference configures it -- Adam(groups, lr=0.0, eps=1e-15), per-group
// learning rates set by the schedulers, optional L2 weight decay (OmniRe/models/trainers/base.py:190-226;
// S3Gaussian/scene/gaussian_model.py:186-200) -- i.e. torch/optim/adam.py `_single_tensor_adam` /
// `_multi_tensor_adam` with amsgrad=False, maximize=False, capturable=False:
//     g   = grad * grad_scale (+ weight_decay * p)
//     m  += (1 - beta1) * (g - m)                      exp_avg.lerp_(grad, 1 - beta1)
//     v   = v * beta2 + (1 - beta2) * g * g            exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
//     p  -= step_size * m / (sqrt(v) / bc2_sqrt + eps) step_size = lr / (1 - beta1^t), bc2_sqrt = sqrt(1 - beta2^t)
// The scalar factors are formed in double on the host like torch does in Python and rounded to fp32 once.
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef EMD_HD
#ifdef __CUDACC__
#define EMD_HD __host__ __device__ __forceinline__
#else
#define EMD_HD static inline
#endif
#endif

struct AdamScalars {
    float one_minus_beta1, beta2, one_minus_beta2, step_size, bc2_sqrt, eps, weight_decay, grad_scale;
};

static inline AdamScalars adam_scalars(double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step,
                                       double grad_scale) {
    AdamScalars s;
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    s.one_minus_beta1 = (float)(1.0 - beta1);
    s.beta2 = (float)beta2;
    s.one_minus_beta2 = (float)(1.0 - beta2);
    s.step_size = (float)(lr / bc1);
    s.bc2_sqrt = (float)sqrt(bc2);
    s.eps = (float)eps;
    s.weight_decay = (float)weight_decay;
    s.grad_scale = (float)grad_scale;
    return s;
}

EMD_HD void adam_update(float& p, float g, float& m, float& v, const AdamScalars& s) {
    if (s.grad_scale != 1.0f) g *= s.grad_scale;
    if (s.weight_decay != 0.0f) g = fmaf(s.weight_decay, p, g);
    m = fmaf(s.one_minus_beta1, g - m, m);
    v = fmaf(s.one_minus_beta2 * g, g, v * s.beta2);
    const float denom = sqrtf(v) / s.bc2_sqrt + s.eps;
    p = fmaf(-s.step_size, m / denom, p);
}
