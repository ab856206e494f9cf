// C-ABI entry points of the batched local-training path (include/flb.h, "batched local training") and the
// model-independent kernels: optimizer step over [K, ld], epoch bookkeeping.  Model-specific launch sequences live in
// train_simplecnn.cu (model 0) and train_cifar.cu (model 1).
#include "train_common.cuh"
#include "philox.cuh"
#include "opt_update.cuh"
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

// grid (blocks, K).  One specialisation per (optimizer, DP mode, tensor-core weight table) so that the loop body is
// straight-line code; the quad of the NEXT iteration is loaded before the current one is processed (the kernel is a
// single wave of ~7 iterations per thread: without the prefetch every iteration exposes a full DRAM round trip -- ncu
// showed 35 % of the DRAM peak).  Also advances the minibatch counter and the clients' optimizer step counts (what
// flb_train_advance does on its own for the forward-only path).
template <int OPT, bool DP, bool TAB>
__global__ void __launch_bounds__(256, 4) optimizer_kernel(flb_train_args a, int P, TcConvTab tab) {
    const int k = blockIdx.y, tid = threadIdx.x;
    const int bsz = flb_bsz(a, k);
    __shared__ OptScalars s_c;
    if (bsz > 0) {
        // distinct buffers (rows of four different matrices; rows are 128 B aligned: ld % 32 == 0)
        float* __restrict__ W = a.W + (long long)k * a.ld;
        float* __restrict__ G = a.G + (long long)k * a.ld;
        float* __restrict__ M = a.M + (long long)k * a.ld;
        float* __restrict__ V = a.V + (long long)k * a.ld;
        const float* __restrict__ zrow = (DP && a.dp_z) ? a.dp_z + (long long)k * a.ld : nullptr;
        const int t = a.tcount[k] + 1;
        // quads [skip_lo/4, skip_hi/4) were already updated by a weight-gradient GEMM epilogue (TcConvTab): the loop runs
        // over the remaining nq_eff quads, e4 -> c4 maps around the hole
        const int nq = P >> 2, stride = gridDim.x * 256;
        const int skip_q0 = tab.skip_lo >> 2, skip_n = (tab.skip_hi - tab.skip_lo) >> 2, nq_eff = nq - skip_n;
        const bool need_m = OPT != 1 || t > 1;
        int e4 = blockIdx.x * 256 + tid;
        int c4 = e4 < skip_q0 ? e4 : e4 + skip_n;
        float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f), w4 = g4, m4 = g4, v4 = g4, z4 = g4;
        auto load = [&](int c, float4& g, float4& w, float4& m, float4& v, float4& z) {
            g = reinterpret_cast<const float4*>(G)[c];
            w = reinterpret_cast<const float4*>(W)[c];
            if (need_m) m = reinterpret_cast<const float4*>(M)[c];
            if (OPT != 1) v = reinterpret_cast<const float4*>(V)[c];
            if (DP && zrow) z = reinterpret_cast<const float4*>(zrow)[c];
        };
        bool have = e4 < nq_eff;
        if (have) load(c4, g4, w4, m4, v4, z4);             // in flight while thread 0 forms the scalars (pow in double)
        if (tid == 0) {
            const OptScalars c = opt_scalars(a, t, bsz);
            s_c = c;
        }
        __syncthreads();
        const OptScalars c = s_c;
        float* __restrict__ wt = tab.wt + (long long)k * tab.ldt;
        const float* __restrict__ gt = tab.gt + (long long)k * tab.ldt;
        int tab_lo = 0x7fffffff, tab_hi = 0;
        if (TAB)
            for (int i = 0; i < tab.n; ++i) {
                tab_lo = min(tab_lo, tab.woff[i]);
                tab_hi = max(tab_hi, tab.woff[i] + tab.cout[i] * tab.cin[i] * 9);
            }
        while (have) {
            const int e_nxt = e4 + stride;
            const int nxt = e_nxt < skip_q0 ? e_nxt : e_nxt + skip_n;
            const bool have_n = e_nxt < nq_eff;
            float4 gn = make_float4(0.f, 0.f, 0.f, 0.f), wn = gn, mn = gn, vn = gn, zn = gn;
            if (have_n) load(nxt, gn, wn, mn, vn, zn);
            const int p0 = c4 * 4;
            float g[4] = {g4.x, g4.y, g4.z, g4.w}, w[4] = {w4.x, w4.y, w4.z, w4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w},
                  v[4] = {v4.x, v4.y, v4.z, v4.w}, z[4] = {z4.x, z4.y, z4.z, z4.w};
            int q[4] = {-1, -1, -1, -1};
            const bool mapped = TAB && p0 + 3 >= tab_lo && p0 < tab_hi;     // quad touches a tensor-core conv weight range
            if (mapped) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    int layer = 0;
                    q[e] = tc_tab_map(tab, p0 + e, layer);
                    if (q[e] >= 0 && tab.gt_live[layer]) g[e] = gt[q[e]];
                }
            }
            if (p0 < tab.g_zero_upto) reinterpret_cast<float4*>(G)[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (DP && c.sigma > 0.f && !zrow) {
                const float4 zz = flb_normal4(flb_epoch_seed(a), a.client_base + a.client_stride * k, ((unsigned long long)t << 32) + c4);
                z[0] = zz.x; z[1] = zz.y; z[2] = zz.z; z[3] = zz.w;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) opt_update<OPT, DP>(c, g[e], z[e], w[e], m[e], v[e]);
            reinterpret_cast<float4*>(W)[c4] = make_float4(w[0], w[1], w[2], w[3]);
            reinterpret_cast<float4*>(M)[c4] = make_float4(m[0], m[1], m[2], m[3]);
            if (OPT != 1) reinterpret_cast<float4*>(V)[c4] = make_float4(v[0], v[1], v[2], v[3]);
            if (mapped) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (q[e] >= 0) wt[q[e]] = w[e];
            }
            c4 = nxt; e4 = e_nxt; have = have_n;
            g4 = gn; w4 = wn; m4 = mn; v4 = vn; z4 = zn;
        }
        if (blockIdx.x == 0 && tid == 0) {                 // the P % 4 parameters after the last whole quad
            for (int p = nq * 4; p < P; ++p) {
                float g = G[p], w = W[p], m = M[p], v = V[p], z = 0.f;
                int layer = 0;
                const int q = TAB ? tc_tab_map(tab, p, layer) : -1;
                if (q >= 0 && tab.gt_live[layer]) g = gt[q];
                if (p < tab.g_zero_upto) G[p] = 0.f;
                if (DP) {
                    if (zrow) z = zrow[p];
                    else if (c.sigma > 0.f) {
                        const float4 zz = flb_normal4(flb_epoch_seed(a), a.client_base + a.client_stride * k, ((unsigned long long)t << 32) + (p >> 2));
                        const float za[4] = {zz.x, zz.y, zz.z, zz.w};
                        z = za[p & 3];
                    }
                }
                opt_update<OPT, DP>(c, g, z, w, m, v);
                W[p] = w; M[p] = m;
                if (OPT != 1) V[p] = v;
                if (q >= 0) wt[q] = w;
            }
        }
    }
    // The last CTA to finish advances the step: every CTA has read *step_ctr / tcount before it takes its ticket.
    __shared__ int s_last;
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const int ticket = atomicAdd(&a.step_ctr[1], 1);
        s_last = ticket == (int)(gridDim.x * gridDim.y) - 1;
    }
    __syncthreads();
    if (s_last) {
        for (int c = tid; c < a.K; c += blockDim.x)
            if (flb_bsz(a, c) > 0) a.tcount[c] += 1;
        __syncthreads();
        if (tid == 0) { a.step_ctr[0] += 1; a.step_ctr[1] = 0; __threadfence(); }
    }
}

// at most ONE resident wave: every CTA lives for the whole kernel, so a handful of CTAs beyond the resident set would
// double its duration (ncu: 600 CTAs on 592 slots cost +6 us of a 24 us launch)
template <int OPT, bool DP, bool TAB>
void launch_optimizer_k(const flb_train_args& a, int P, const TcConvTab& tab, cudaStream_t st) {
    static const int resident = flb_resident_ctas(optimizer_kernel<OPT, DP, TAB>, 256);
    const int blocks = max(1, min(flb_cdiv((P - (tab.skip_hi - tab.skip_lo)) / 4, 256), resident / a.K));
    optimizer_kernel<OPT, DP, TAB><<<dim3(blocks, a.K), 256, 0, st>>>(a, P, tab);
}
template <int OPT, bool DP>
void launch_optimizer_tab(const flb_train_args& a, int P, const TcConvTab& tab, cudaStream_t st) {
    if (tab.n) launch_optimizer_k<OPT, DP, true>(a, P, tab, st);
    else launch_optimizer_k<OPT, DP, false>(a, P, tab, st);
}
void launch_optimizer(const flb_train_args& a, int P, const TcConvTab& tab, cudaStream_t st) {
    const bool dp = a.dp_mode == 1;
    if (a.opt == 0) { if (dp) launch_optimizer_tab<0, true>(a, P, tab, st); else launch_optimizer_tab<0, false>(a, P, tab, st); }
    else if (a.opt == 1) { if (dp) launch_optimizer_tab<1, true>(a, P, tab, st); else launch_optimizer_tab<1, false>(a, P, tab, st); }
    else { if (dp) launch_optimizer_tab<2, true>(a, P, tab, st); else launch_optimizer_tab<2, false>(a, P, tab, st); }
}

__global__ void advance_kernel(flb_train_args a) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < a.K && flb_bsz(a, k) > 0) a.tcount[k] += 1;
    __syncthreads();            // single block: every tcount update read the old step first
    if (k == 0) *a.step_ctr += 1;
}

// Start of a federated round: every resident client takes the global model and a fresh optimizer (the reference builds a new
// LocalTrainer -- new torch.optim state -- per round and loads the downloaded weights, src/client/federated_trainer.py:367-392).
// One pass instead of a broadcast copy and three fills: W[k] = global, M[k] = V[k] = 0 over the whole padded row, tcount = 0.
__global__ void __launch_bounds__(256) round_begin_kernel(flb_train_args a, const float* __restrict__ global_row) {
    const int k = blockIdx.y;
    const long long ld4 = a.ld >> 2;
    float4* W = reinterpret_cast<float4*>(a.W + (long long)k * a.ld);
    float4* M = reinterpret_cast<float4*>(a.M + (long long)k * a.ld);
    float4* V = reinterpret_cast<float4*>(a.V + (long long)k * a.ld);
    const float4* g = reinterpret_cast<const float4*>(global_row);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long c = (long long)blockIdx.x * 256 + threadIdx.x; c < ld4; c += (long long)gridDim.x * 256) {
        W[c] = __ldg(&g[c]);
        M[c] = z;
        V[c] = z;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) a.tcount[k] = 0;
}

__global__ void begin_epoch_kernel(flb_train_args a) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < a.K) { a.loss_sum[k] = 0.f; a.correct[k] = 0; a.nbatch[k] = 0; a.nseen[k] = 0; }
    if (k == 0) {
        a.step_ctr[0] = 0; a.step_ctr[1] = 0;
        if (a.epoch_nonce) *a.epoch_nonce += 1;
    }
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
    }
    return FLB_OK;
}

int num_params(const flb_train_args& a) { return a.model == 0 ? simplecnn::num_params() : cifar::num_params(); }

TcConvTab tab_of(const flb_train_args& a, bool step = false) {
    TcConvTab t;
    if (a.model == 0) simplecnn::tc_tab(a, &t, step); else cifar::tc_tab(a, &t, step);
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

extern "C" int flb_train_begin_round(const flb_train_args* a, const float* global_row, void* stream) {
    if (int rc = check_args(a)) return rc;
    FLB_CHECK_ARG(global_row != nullptr && a->ld % 4 == 0 && ((uintptr_t)global_row % 16) == 0,
                  "flb_train_begin_round: need a 16-byte aligned global row of ld floats, ld %% 4 == 0");
    static const int resident = flb_resident_ctas(round_begin_kernel, 256);
    const int blocks = max(1, min(flb_cdiv(a->ld / 4, 256), resident / a->K));
    round_begin_kernel<<<dim3(blocks, a->K), 256, 0, (cudaStream_t)stream>>>(*a, global_row);
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

static int fwd_bwd(const flb_train_args& a, cudaStream_t st, bool zero_first, bool step = false) {
    return a.model == 0 ? simplecnn::forward_backward(a, st, zero_first, step) : cifar::forward_backward(a, st, step);
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
    if (int rc = fwd_bwd(*a, st, false, true)) return rc;      // accumulators are zero: begin_epoch + the optimizer keep them so
    const int P = num_params(*a);
    launch_optimizer(*a, P, tab_of(*a, true), st);
    MARK("optimizer");
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

// flb_train_step that also hands out the minibatch gradient it applied (LocalTrainer.get_model_gradients,
// training.py:362-371: param.grad after the last step): the rows are copied after the backward pass and before the
// optimizer consumes (and re-zeroes) them; tensor-core conv layers are first brought back from the tap-major copy.
extern "C" int flb_train_step_grads(const flb_train_args* a, float* grads_out, long long ld_out, void* stream) {
    if (int rc = check_args(a)) return rc;
    const int P = num_params(*a);
    FLB_CHECK_ARG(grads_out != nullptr && ld_out >= P, "flb_train_step_grads: need grads_out [K, ld_out >= %d]", P);
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = fwd_bwd(*a, st, false)) return rc;
    const TcConvTab t = tab_of(*a);
    if (t.n) tc_repack_kernel<1><<<dim3(repack_blocks(*a, t), a->K), 256, 0, st>>>(*a, t);
    FLB_CUDA(cudaMemcpy2DAsync(grads_out, (size_t)ld_out * sizeof(float), a->G, (size_t)a->ld * sizeof(float),
                               (size_t)P * sizeof(float), (size_t)a->K, cudaMemcpyDeviceToDevice, st));
    launch_optimizer(*a, P, t, st);
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
