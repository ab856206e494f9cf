// The per-element optimizer update (torch.optim semantics, reference src/shared/training.py:244-255: Adam(lr) |
// SGD(lr, momentum=0.9) | AdamW(lr)), shared by the stand-alone optimizer kernel (train_api.cu) and by the weight-gradient
// GEMM epilogues that apply it straight out of TMEM (train_tc.cu).
#pragma once
#include "train_common.cuh"

// Scalars of one optimizer step of one client: formed in double and rounded to fp32 once, like Python floats entering
// fp32 tensor ops.
struct OptScalars {
    float step_size, inv_bc2_sqrt, lr, omb1, b2, omb2, eps, decay, mu, inv_b, sigma;
    int t;
};

// t = the step being taken (tcount + 1); bsz = live samples of the client's minibatch
__device__ __forceinline__ OptScalars opt_scalars(const flb_train_args& a, int t, int bsz) {
    OptScalars c;
    // beta^t by binary exponentiation (<= 2 log2 t dependent double multiplies; libm pow() is hundreds of instructions)
    double p1 = 1.0, p2 = 1.0, q1 = a.beta1, q2 = a.beta2;
    for (int e = t; e > 0; e >>= 1) {
        if (e & 1) { p1 *= q1; p2 *= q2; }
        q1 *= q1; q2 *= q2;
    }
    const double bc1 = 1.0 - p1, bc2 = 1.0 - p2;
    c.step_size = (float)(a.lr / bc1);
    c.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
    c.lr = (float)a.lr; c.omb1 = (float)(1.0 - a.beta1); c.b2 = (float)a.beta2; c.omb2 = (float)(1.0 - a.beta2);
    c.eps = (float)a.eps; c.decay = (float)(1.0 - a.lr * a.weight_decay); c.mu = (float)a.momentum;
    c.inv_b = 1.f / (float)bsz; c.sigma = a.dp_sigma; c.t = t;
    return c;
}

__device__ __forceinline__ float sqrt_fast(float x) {        // <= 1 ulp, exact 0 -> 0 (v can be exactly zero)
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// One element.  OPT: 0 Adam, 1 SGD(momentum), 2 AdamW.  DP: g is the sum of clipped per-sample gradients, z a standard normal.
template <int OPT, bool DP>
__device__ __forceinline__ void opt_update(const OptScalars& c, float g, float z, float& w, float& m, float& v) {
    if (DP) g = (g + c.sigma * z) * c.inv_b;                // (sum clipped + N(0, sigma^2)) / B
    if (OPT == 1) {                                         // SGD with momentum, dampening 0
        const float buf = c.t == 1 ? g : fmaf(c.mu, m, g);
        m = buf;
        w = w - c.lr * buf;
    } else {
        if (OPT == 2) w = w * c.decay;                      // AdamW decoupled decay
        m = m + (g - m) * c.omb1;                           // exp_avg.lerp_(grad, 1 - beta1)
        v = v * c.b2 + c.omb2 * g * g;                      // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
        // denom = sqrt(v) / sqrt(1 - b2^t) + eps; param.addcdiv_(m, denom, -step_size).  Square root and division use
        // the hardware approximations (<= 2 ulp): Adam trajectories are compared at +-lr granularity anyway
        // (conftest.adam_trajectory_check) and the IEEE forms made this kernel instruction-bound (ncu).
        const float denom = fmaf(sqrt_fast(v), c.inv_bc2_sqrt, c.eps);
        w = w - c.step_size * __fdividef(m, denom);
    }
}
