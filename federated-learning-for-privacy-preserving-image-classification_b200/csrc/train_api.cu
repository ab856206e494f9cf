// C-ABI entry points of the batched local-training path (include/flb.h, "batched local training") and the
// model-independent kernels: optimizer step over [K, ld], epoch bookkeeping.  Model-specific launch sequences live in
// train_simplecnn.cu (model 0) and train_cifar.cu (model 1).
#include "train_common.cuh"
#include "philox.cuh"
#include <string.h>

StepProfile g_prof;

namespace {

// ------------------------------------------------------------------------------------------------
// optimizer over [K, ld]; torch.optim semantics (training.py:244-255): Adam(lr) | SGD(lr, momentum=0.9) | AdamW(lr)
// W (reference layout) -> Wt (tap-major) for every tensor-core conv layer; DIR 1: Gt -> G for the live layers
template <int DIR>
__global__ void __launch_bounds__(256) tc_repack_kernel(flb_train_args a, TcConvTab t) {
    const int k = blockIdx.y;
    const float* W = a.W + (long long)k * a.ld;
    float* G = a.G + (long long)k * a.ld;
    float* wt = t.wt + (long long)k * t.ldt;
    const float* gt = t.gt + (long long)k * t.ldt;
    for (int e = blockIdx.x * 256 + threadIdx.x; e < t.ldt; e += gridDim.x * 256) {
        int i = 0;
        while (i + 1 < t.n && e >= t.toff[i + 1]) ++i;
        const int r = e - t.toff[i], cin = t.cin[i], cout = t.cout[i];
        const int ci = r % cin, co = (r / cin) % cout, tap = r / (cin * cout);
        const int p = t.woff[i] + (co * cin + ci) * 9 + tap;
        if (DIR == 0) wt[e] = W[p];
        else if (t.gt_live[i]) G[p] = gt[e];
    }
}

__device__ __forceinline__ void optimizer_body(const flb_train_args& a, int P, const TcConvTab& tab, int k, int bsz);

// grid (blocks, K).  Also advances the minibatch counter and the clients' optimizer step counts (what
// flb_train_advance does on its own for the forward-only path).
__global__ void __launch_bounds__(256) optimizer_kernel(flb_train_args a, int P, TcConvTab tab) {
    const int k = blockIdx.y;
    const int bsz = flb_bsz(a, k);
    if (bsz > 0) optimizer_body(a, P, tab, k, bsz);
    // The last CTA to finish advances the step: every CTA has read *step_ctr / tcount before it takes its ticket.
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int ticket = atomicAdd(&a.step_ctr[1], 1);
        s_last = ticket == (int)(gridDim.x * gridDim.y) - 1;
    }
    __syncthreads();
    if (s_last) {
        for (int c = threadIdx.x; c < a.K; c += blockDim.x)
            if (flb_bsz(a, c) > 0) a.tcount[c] += 1;
        __syncthreads();
        if (threadIdx.x == 0) { a.step_ctr[0] += 1; a.step_ctr[1] = 0; __threadfence(); }
    }
}

__device__ __forceinline__ void optimizer_body(const flb_train_args& a, int P, const TcConvTab& tab, int k, int bsz) {
    const int t = a.tcount[k] + 1;
    // distinct buffers (rows of four different matrices): __restrict__ lets the loads of the next quad fly past the stores
    float* __restrict__ W = a.W + (long long)k * a.ld;
    float* __restrict__ G = a.G + (long long)k * a.ld;
    float* __restrict__ M = a.M + (long long)k * a.ld;
    float* __restrict__ V = a.V + (long long)k * a.ld;
    // scalars are formed in double and rounded to fp32 once, like Python floats entering fp32 tensor ops
    // (one thread per CTA: pow() in double is hundreds of instructions)
    __shared__ float s_sc[2];
    if (threadIdx.x == 0) {
        const double bc1d = 1.0 - pow(a.beta1, (double)t), bc2d = 1.0 - pow(a.beta2, (double)t);
        s_sc[0] = (float)(a.lr / bc1d);
        s_sc[1] = (float)(1.0 / sqrt(bc2d));
    }
    __syncthreads();
    const float step_size = s_sc[0], inv_bc2_sqrt = s_sc[1];
    const float lr = (float)a.lr, omb1 = (float)(1.0 - a.beta1), b2 = (float)a.beta2, omb2 = (float)(1.0 - a.beta2);
    const float eps = (float)a.eps, decay = (float)(1.0 - a.lr * a.weight_decay), mu = (float)a.momentum;
    const float inv_b = 1.f / (float)bsz;
    const float* zrow = a.dp_z ? a.dp_z + (long long)k * a.ld : nullptr;
    float* __restrict__ wt = tab.wt + (long long)k * tab.ldt;
    const float* __restrict__ gt = tab.gt + (long long)k * tab.ldt;
    const int P4 = (P + 3) >> 2;
    const bool adam = a.opt != 1;
    int tab_lo = 0x7fffffff, tab_hi = 0;
    for (int i = 0; i < tab.n; ++i) {
        tab_lo = min(tab_lo, tab.woff[i]);
        tab_hi = max(tab_hi, tab.woff[i] + tab.cout[i] * tab.cin[i] * 9);
    }
#pragma unroll 2
    for (int c4 = blockIdx.x * 256 + threadIdx.x; c4 < P4; c4 += gridDim.x * 256) {
        const int p0 = c4 * 4;
        const bool whole = p0 + 3 < P;           // rows are 128 B aligned (ld % 32 == 0): 16 B vector access per quad
        float g[4] = {0.f, 0.f, 0.f, 0.f}, w[4] = {0.f, 0.f, 0.f, 0.f}, m[4] = {0.f, 0.f, 0.f, 0.f}, v[4] = {0.f, 0.f, 0.f, 0.f};
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        int q[4] = {-1, -1, -1, -1};
        if (whole) {
            const float4 g4 = *reinterpret_cast<const float4*>(G + p0), w4 = *reinterpret_cast<const float4*>(W + p0);
            g[0] = g4.x; g[1] = g4.y; g[2] = g4.z; g[3] = g4.w;
            w[0] = w4.x; w[1] = w4.y; w[2] = w4.z; w[3] = w4.w;
            if (adam || t > 1) { const float4 m4 = *reinterpret_cast<const float4*>(M + p0); m[0] = m4.x; m[1] = m4.y; m[2] = m4.z; m[3] = m4.w; }
            if (adam) { const float4 v4 = *reinterpret_cast<const float4*>(V + p0); v[0] = v4.x; v[1] = v4.y; v[2] = v4.z; v[3] = v4.w; }
            if (a.dp_mode == 1 && zrow) { const float4 z4 = *reinterpret_cast<const float4*>(zrow + p0); z[0] = z4.x; z[1] = z4.y; z[2] = z4.z; z[3] = z4.w; }
        } else {
            for (int e = 0; e < 4 && p0 + e < P; ++e) {
                g[e] = G[p0 + e]; w[e] = W[p0 + e]; m[e] = M[p0 + e]; v[e] = V[p0 + e];
                if (a.dp_mode == 1 && zrow) z[e] = zrow[p0 + e];
            }
        }
        const bool mapped = tab.n && p0 + 3 >= tab_lo && p0 < tab_hi;     // quad touches a tensor-core conv weight range
        if (mapped) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int layer = 0;
                q[e] = p0 + e < P ? tc_tab_map(tab, p0 + e, layer) : -1;
                if (q[e] >= 0 && tab.gt_live[layer]) g[e] = gt[q[e]];
            }
        }
        if (p0 < tab.g_zero_upto) {
            if (whole) *reinterpret_cast<float4*>(G + p0) = make_float4(0.f, 0.f, 0.f, 0.f);
            else for (int e = 0; e < 4 && p0 + e < P; ++e) G[p0 + e] = 0.f;
        }
        if (a.dp_mode == 1 && a.dp_sigma > 0.f && !zrow) {
            const float4 zz = flb_normal4(a.seed, a.client_base + a.client_stride * k, ((unsigned long long)t << 32) + c4);
            z[0] = zz.x; z[1] = zz.y; z[2] = zz.z; z[3] = zz.w;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (a.dp_mode == 1) g[e] = (g[e] + a.dp_sigma * z[e]) * inv_b;     // (sum clipped + N(0, sigma^2)) / B
            if (!adam) {                                    // SGD with momentum, dampening 0
                const float buf = t == 1 ? g[e] : fmaf(mu, m[e], g[e]);
                m[e] = buf;
                w[e] = w[e] - lr * buf;
            } else {
                if (a.opt == 2) w[e] = w[e] * decay;        // AdamW decoupled decay
                m[e] = m[e] + (g[e] - m[e]) * omb1;         // exp_avg.lerp_(grad, 1 - beta1)
                v[e] = v[e] * b2 + omb2 * g[e] * g[e];      // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
                // denom = sqrt(v) / sqrt(1 - b2^t) + eps; param.addcdiv_(m, denom, -step_size).  The two divisions use the
                // fast reciprocal forms (<= 2 ulp): Adam trajectories are compared at +-lr granularity anyway
                // (conftest.adam_trajectory_check) and IEEE division made this kernel instruction-bound (ncu).
                const float denom = fmaf(__fsqrt_rn(v[e]), inv_bc2_sqrt, eps);
                w[e] = w[e] - step_size * __fdividef(m[e], denom);
            }
        }
        if (whole) {
            *reinterpret_cast<float4*>(W + p0) = make_float4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<float4*>(M + p0) = make_float4(m[0], m[1], m[2], m[3]);
            if (adam) *reinterpret_cast<float4*>(V + p0) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            for (int e = 0; e < 4 && p0 + e < P; ++e) { W[p0 + e] = w[e]; M[p0 + e] = m[e]; if (adam) V[p0 + e] = v[e]; }
        }
        if (mapped) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (q[e] >= 0) wt[q[e]] = w[e];
        }
    }
}

__global__ void advance_kernel(flb_train_args a) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < a.K && flb_bsz(a, k) > 0) a.tcount[k] += 1;
    __syncthreads();            // single block: every tcount update read the old step first
    if (k == 0) *a.step_ctr += 1;
}

__global__ void begin_epoch_kernel(flb_train_args a) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < a.K) { a.loss_sum[k] = 0.f; a.correct[k] = 0; a.nbatch[k] = 0; a.nseen[k] = 0; }
    if (k == 0) { a.step_ctr[0] = 0; a.step_ctr[1] = 0; }
}

int check_args(const flb_train_args* a) {
    FLB_CHECK_ARG(a != nullptr, "flb_train: null args");
    FLB_CHECK_ARG(a->model == 0 || a->model == 1, "flb_train: model %d not supported (0 = simple_cnn, 1 = cifar10_cnn)", a->model);
    FLB_CHECK_ARG(a->K >= 1 && a->K <= 1024 && a->B >= 1 && a->B <= 32, "flb_train: need 1 <= K <= 1024 and 1 <= B <= 32 (K=%d B=%d)", a->K, a->B);
    const int P = a->model == 0 ? simplecnn::num_params() : cifar::num_params();
    FLB_CHECK_ARG(a->ld >= P, "flb_train: ld %lld < %d parameters", a->ld, P);
    FLB_CHECK_ARG(a->x && a->y && a->sample_off && a->nsamples && a->step_ctr && a->W && a->G && a->M && a->V &&
                  a->tcount && a->ws && a->loss_sum && a->correct && a->nbatch && a->nseen, "flb_train: null device pointer in args");
    FLB_CHECK_ARG(a->opt >= 0 && a->opt <= 2, "flb_train: Unknown optimizer type: %d", a->opt);
    FLB_CHECK_ARG(a->drop_p >= 0.f && a->drop_p < 1.f, "flb_train: dropout probability must be in [0, 1)");
    FLB_CHECK_ARG(a->precision == 0 || a->precision == 1, "flb_train: precision must be 0 (fp32) or 1 (tf32 tensor cores)");
    FLB_CHECK_ARG(a->dp_mode == 0 || a->dp_mode == 1, "flb_train: dp_mode must be 0 or 1");
    if (a->model == 1) {
        FLB_CHECK_ARG(a->bn_running != nullptr, "flb_train: cifar10_cnn needs the bn_running buffer");
        if (a->dp_mode == 1) {
            flb_set_error("flb_train: per-sample DP is undefined for cifar10_cnn (BatchNorm couples the samples of a batch)");
            return FLB_ERR_UNSUPPORTED;
        }
    }
    return FLB_OK;
}

int num_params(const flb_train_args& a) { return a.model == 0 ? simplecnn::num_params() : cifar::num_params(); }

TcConvTab tab_of(const flb_train_args& a) {
    TcConvTab t;
    if (a.model == 0) simplecnn::tc_tab(a, &t); else cifar::tc_tab(a, &t);
    return t;
}
int repack_blocks(const flb_train_args& a, const TcConvTab& t) {
    return max(1, min(flb_cdiv(t.ldt, 256), (flb_num_sms() * 8 + a.K - 1) / a.K));
}

}  // namespace

extern "C" long long flb_train_ws_bytes(int model, int K, int B) {
    if (K < 1 || B < 1) return -1;
    return model == 0 ? simplecnn::ws_bytes(K, B) : model == 1 ? cifar::ws_bytes(K, B) : -1;
}

extern "C" long long flb_train_ws_offset(int model, int K, int B, const char* name) {
    if (K < 1 || B < 1 || !name) return -1;
    return model == 0 ? simplecnn::ws_offset(K, B, name) : model == 1 ? cifar::ws_offset(K, B, name) : -1;
}

extern "C" long long flb_train_bn_floats(int model) { return model == 1 ? cifar::bn_floats() : 0; }

extern "C" int flb_train_begin_epoch(const flb_train_args* a, void* stream) {
    if (int rc = check_args(a)) return rc;
    begin_epoch_kernel<<<flb_cdiv(a->K, 256), 256, 0, (cudaStream_t)stream>>>(*a);
    const TcConvTab t = tab_of(*a);
    if (t.n) tc_repack_kernel<0><<<dim3(repack_blocks(*a, t), a->K), 256, 0, (cudaStream_t)stream>>>(*a, t);
    if (a->model == 0)
        if (int rc = simplecnn::begin_epoch_zero(*a, (cudaStream_t)stream)) return rc;
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

extern "C" int flb_train_forward(const flb_train_args* a, void* stream) {
    if (int rc = check_args(a)) return rc;
    return a->model == 0 ? simplecnn::forward(*a, (cudaStream_t)stream) : cifar::forward(*a, (cudaStream_t)stream);
}

extern "C" int flb_train_advance(const flb_train_args* a, void* stream) {
    if (int rc = check_args(a)) return rc;
    advance_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(*a);
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

static int fwd_bwd(const flb_train_args& a, cudaStream_t st, bool zero_first) {
    return a.model == 0 ? simplecnn::forward_backward(a, st, zero_first) : cifar::forward_backward(a, st);
}

extern "C" int flb_train_forward_backward(const flb_train_args* a, void* stream) {
    if (int rc = check_args(a)) return rc;
    if (int rc = fwd_bwd(*a, (cudaStream_t)stream, true)) return rc;
    const TcConvTab t = tab_of(*a);          // the step proper never needs G in the reference layout; this entry does
    if (t.n) tc_repack_kernel<1><<<dim3(repack_blocks(*a, t), a->K), 256, 0, (cudaStream_t)stream>>>(*a, t);
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

extern "C" int flb_train_step(const flb_train_args* a, void* stream) {
    if (int rc = check_args(a)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = fwd_bwd(*a, st, false)) return rc;      // accumulators are zero: begin_epoch + the optimizer keep them so
    const int P = num_params(*a);
    // 4 CTAs per SM = one resident wave at 60 registers per thread (measured best of 2/4/8/16/32)
    const int blocks = max(1, min(flb_cdiv(P / 4, 256), (flb_num_sms() * 4 + a->K - 1) / a->K));
    optimizer_kernel<<<dim3(blocks, a->K), 256, 0, st>>>(*a, P, tab_of(*a));
    MARK("optimizer");
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

// number of kernel launches (memsets excluded) one flb_train_step issues for these args
extern "C" int flb_train_step_launches(const flb_train_args* a) {
    if (!a) return -1;
    return 1 + (a->model == 0 ? simplecnn::step_launches(*a) : cifar::step_launches(*a));
}

// One step with a CUDA event after every kernel.  Synchronises the stream (profiling aid, not the product path).
// names_out receives '\n'-separated labels; ms_out[i] = device time of labelled segment i.  Returns the segment count.
extern "C" int flb_train_step_profiled(const flb_train_args* a, void* stream, char* names_out, int names_cap,
                                       float* ms_out, int max_n) {
    if (int rc = check_args(a)) return rc;
    FLB_CHECK_ARG(names_out && ms_out && names_cap > 0 && max_n > 0, "flb_train_step_profiled: bad output buffers");
    for (int i = 0; i < 96; ++i) FLB_CUDA(cudaEventCreate(&g_prof.ev[i]));
    g_prof.n = 0;
    g_prof.on = true;
    const int rc = flb_train_step(a, stream);
    g_prof.on = false;
    int n = 0;
    if (rc == FLB_OK && cudaStreamSynchronize((cudaStream_t)stream) == cudaSuccess) {
        names_out[0] = 0;
        size_t used = 0;
        for (int i = 1; i < g_prof.n && n < max_n; ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, g_prof.ev[i - 1], g_prof.ev[i]);
            ms_out[n++] = ms;
            const size_t len = strlen(g_prof.name[i]);
            if (used + len + 2 < (size_t)names_cap) {
                memcpy(names_out + used, g_prof.name[i], len);
                used += len;
                names_out[used++] = '\n';
                names_out[used] = 0;
            }
        }
    }
    for (int i = 0; i < 96; ++i) cudaEventDestroy(g_prof.ev[i]);
    return rc == FLB_OK ? n : rc;
}
