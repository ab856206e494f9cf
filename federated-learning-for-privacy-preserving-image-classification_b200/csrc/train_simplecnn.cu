// SimpleCNN (reference src/shared/models_pytorch.py:59-97) local-training step, batched over all resident
// clients: conv3x3(1->32)+ReLU+pool -> conv3x3(32->64)+ReLU+pool -> fc 3136->128 + ReLU + dropout -> fc 128->10,
// mean cross-entropy (src/shared/training.py:90,193), backward, optimizer step (training.py:244-255).
//
// Kernel inventory for one step (fp32 path; the TF32 tcgen05 path swaps the GEMM-shaped ones, see gemm_tc.cu):
//   conv1_fwd_pool      direct 3x3 stencil (K = 9 is too thin for a tensor-core tile) fused with bias+ReLU+2x2 pool
//   conv2 fwd           implicit GEMM [B*256 px, 288] x [288, 64] on the zero-padded NHWC grid (no im2col)
//   pool2               bias already added; ReLU + 2x2 max-pool + argmax, writes fc1's NCHW-flattened input
//   fc1 fwd             [B, 3136] x [3136, 128], split-K
//   head_fwd_bwd        fc1 bias+ReLU+dropout, fc2, softmax cross-entropy, dlogits, dh  (one CTA per client)
//   fc1 dgrad / unpool2 / conv2 dgrad            activation gradients
//   [dp_mode 1]         per-sample gradient norms from the activation gradients (ghost norms for the linear
//                       layers, on-chip per-sample conv wgrad tiles for the convs) -> clip coefficients
//   head_wgrad, fc1 wgrad, conv2 wgrad, conv1 wgrad    weight gradients (+bias as an extra GEMM column),
//                       rows scaled by the per-sample clip coefficient in dp_mode 1 -- per-sample weight
//                       gradients are never written to HBM
//   optimizer           Adam / SGD-momentum / AdamW over [K, ld]; dp_mode 1 adds Philox noise and the 1/B here
#include "train_common.cuh"
#include "gemm_simt.cuh"
#include "philox.cuh"
#include <string.h>

namespace tc {
int conv_fwd_32_64(const flb_train_args& a, const ConvGeom& g, const float* xin, float* z, int woff, int boff, cudaStream_t st);
int conv_dgrad_32_64(const flb_train_args& a, const ConvGeom& g, const float* dz, float* dx, int woff, cudaStream_t st);
int conv_wgrad_32_64(const flb_train_args& a, const ConvGeom& g, const float* xin, const float* dz, int woff, int splits, cudaStream_t st);
int fc_fwd_3136_128(const flb_train_args& a, const float* act, float* out, int woff, int splits, cudaStream_t st);
int fc_dgrad_3136_128(const flb_train_args& a, const float* dout, float* dact, int woff, cudaStream_t st);
int fc_wgrad_3136_128(const flb_train_args& a, const float* dout, const float* act, int woff, cudaStream_t st);
}  // namespace tc

namespace {

enum : int { TC_CONV2_FWD = 1, TC_FC1_FWD = 2, TC_FC1_DGRAD = 4, TC_CONV2_DGRAD = 8, TC_FC1_WGRAD = 16, TC_CONV2_WGRAD = 32 };

// which GEMMs run on the tensor cores for these args
int tc_mask_of(const flb_train_args& a) {
    int m = a.precision == 1 ? (a.tc_mask ? a.tc_mask : 63) : 0;
    if (a.B % 8) m &= ~TC_FC1_WGRAD;        // its K extent is the batch: whole 8-row MMA steps only
    return m;
}

// per-kernel CUDA-event timing of one step (flb_train_step_profiled); inactive otherwise
struct StepProfile {
    bool on = false;
    int n = 0;
    cudaEvent_t ev[48];
    const char* name[48];
};
StepProfile g_prof;
#define MARK(label)                                                       \
    do {                                                                  \
        if (g_prof.on && g_prof.n < 48) {                                 \
            cudaEventRecord(g_prof.ev[g_prof.n], st);                     \
            g_prof.name[g_prof.n++] = label;                              \
        }                                                                 \
    } while (0)

using Off = SimpleCnnOff;
constexpr int PP2 = 256;     // conv2 runs on a 16x16 padded grid (14x14 real)
constexpr int WP2 = 16;

// ------------------------------------------------------------------------------------------------
// conv1: direct stencil, one CTA per (sample, client)
__global__ void __launch_bounds__(256) conv1_fwd_pool_kernel(flb_train_args a, SimpleCnnWs ws) {
    const int b = blockIdx.x, k = blockIdx.y;
    if (b >= flb_bsz(a, k)) return;
    __shared__ float img[30][31];
    __shared__ float w[32][9];
    __shared__ float bias[32];
    const int tid = threadIdx.x;
    const float* W = a.W + (long long)k * a.ld;
    const long long s = a.sample_off[k] + (long long)(*a.step_ctr) * a.B + b;
    const float* x = a.x + s * 784;
    for (int i = tid; i < 900; i += 256) {
        const int r = i / 30, c = i % 30;
        img[r][c] = (r >= 1 && r <= 28 && c >= 1 && c <= 28) ? x[(r - 1) * 28 + (c - 1)] : 0.f;
    }
    for (int i = tid; i < 288; i += 256) w[i / 9][i % 9] = W[Off::c1w + i];
    if (tid < 32) bias[tid] = W[Off::c1b + tid];
    __syncthreads();
    const int c = tid & 31, g = tid >> 5;
    float wr[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) wr[i] = w[c][i];
    const float bc = bias[c];
    const long long kb = (long long)k * a.B + b;
    float* outp = ws.a1p + kb * (PP2 * 32);
    uint8_t* idx = ws.idx1 + kb * (196 * 32);
    for (int pp = g; pp < 196; pp += 8) {
        const int ph = pp / 14, pw = pp % 14;
        float p[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) p[i][j] = img[2 * ph + i][2 * pw + j];
        float best = -INFINITY;
        int bi = 0;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float v = bc;
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int q = 0; q < 3; ++q) v = fmaf(wr[r * 3 + q], p[i + r][j + q], v);
                if (v > best) { best = v; bi = i * 2 + j; }
            }
        outp[(ph * WP2 + pw) * 32 + c] = fmaxf(best, 0.f);
        idx[pp * 32 + c] = (uint8_t)bi;
    }
}

// conv2 epilogue on the fp32 path: ReLU + 2x2 max-pool + argmax from z2 (bias included), NCHW-flattened output
__global__ void __launch_bounds__(256) pool2_kernel(flb_train_args a, SimpleCnnWs ws) {
    const int b = blockIdx.x, k = blockIdx.y;
    if (b >= flb_bsz(a, k)) return;
    const long long kb = (long long)k * a.B + b;
    const float* z = ws.z2 + kb * (PP2 * 64);
    float* o = ws.a2 + kb * 3136;
    uint8_t* idx = ws.idx2 + kb * 3136;
    for (int e = threadIdx.x; e < 3136; e += 256) {
        const int c = e & 63, pp = e >> 6, ph = pp / 7, pw = pp % 7;
        float best = -INFINITY;
        int bi = 0;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float v = z[((2 * ph + i) * WP2 + 2 * pw + j) * 64 + c];
                if (v > best) { best = v; bi = i * 2 + j; }
            }
        o[c * 49 + pp] = fmaxf(best, 0.f);
        idx[c * 49 + pp] = (uint8_t)bi;
    }
}

// ------------------------------------------------------------------------------------------------
// implicit-GEMM problem functors on the padded NHWC grid
struct ConvFwdProb {
    static constexpr bool A_MCONTIG = false, B_NCONTIG = false;
    flb_train_args a; ConvGeom g;
    const float* xin_all; float* z_all; int woff, boff;
    const float* xin; float* z; const float* w; const float* bias; int Mtot;
    __device__ bool setup(int client, int& M, int& N, int& Kd) {
        const int bsz = flb_bsz(a, client);
        if (bsz == 0) return false;
        const long long kb = (long long)client * a.B;
        xin = xin_all + kb * g.PP() * g.Cin;
        z = z_all + kb * g.PP() * g.Cout;
        w = a.W + (long long)client * a.ld + woff;
        bias = a.W + (long long)client * a.ld + boff;
        Mtot = a.B * g.PP();
        M = bsz * g.PP(); N = g.Cout; Kd = 9 * g.Cin;
        return true;
    }
    __device__ float loadA(int m, int k) const {
        const int tap = k / g.Cin, ci = k - tap * g.Cin;
        const int row = m + (tap / 3 - 1) * g.Wp + (tap % 3 - 1);
        return (row >= 0 && row < Mtot) ? xin[(long long)row * g.Cin + ci] : 0.f;
    }
    __device__ float loadB(int n, int k) const {
        const int tap = k / g.Cin, ci = k - tap * g.Cin;
        return __ldg(&w[(n * g.Cin + ci) * 9 + tap]);
    }
    __device__ void store(int m, int n, float acc) { z[(long long)m * g.Cout + n] = acc + bias[n]; }
    __device__ void finish() {}
};

struct ConvDgradProb {      // dx[m][ci] = sum_{tap,co} dz[m - shift(tap)][co] * W[co][ci][tap]
    static constexpr bool A_MCONTIG = false, B_NCONTIG = false;
    flb_train_args a; ConvGeom g;
    const float* dz_all; float* dx_all; int woff;
    const float* dz; float* dx; const float* w; int Mtot;
    __device__ bool setup(int client, int& M, int& N, int& Kd) {
        const int bsz = flb_bsz(a, client);
        if (bsz == 0) return false;
        const long long kb = (long long)client * a.B;
        dz = dz_all + kb * g.PP() * g.Cout;
        dx = dx_all + kb * g.PP() * g.Cin;
        w = a.W + (long long)client * a.ld + woff;
        Mtot = a.B * g.PP();
        M = bsz * g.PP(); N = g.Cin; Kd = 9 * g.Cout;
        return true;
    }
    __device__ float loadA(int m, int k) const {
        const int tap = k / g.Cout, co = k - tap * g.Cout;
        const int row = m - ((tap / 3 - 1) * g.Wp + (tap % 3 - 1));
        return (row >= 0 && row < Mtot) ? dz[(long long)row * g.Cout + co] : 0.f;
    }
    __device__ float loadB(int n, int k) const {
        const int tap = k / g.Cout, co = k - tap * g.Cout;
        return __ldg(&w[(co * g.Cin + n) * 9 + tap]);
    }
    __device__ void store(int m, int n, float acc) { dx[(long long)m * g.Cin + n] = acc; }
    __device__ void finish() {}
};

// dW[co][ci][tap] = sum_px dz[px][co] * x[px + shift(tap)][ci];  column n == 9*Cin is the bias gradient.
// In dp_mode 1 each pixel row is scaled by its sample's clip coefficient.
struct ConvWgradProb {
    static constexpr bool A_MCONTIG = true, B_NCONTIG = true;
    flb_train_args a; ConvGeom g;
    const float* dz_all; const float* xin_all; const float* coef_all; int woff, boff;
    const float* dz; const float* xin; const float* coef; float* gw; float* gb; int Mtot;
    __device__ bool setup(int client, int& M, int& N, int& Kd) {
        const int bsz = flb_bsz(a, client);
        if (bsz == 0) return false;
        const long long kb = (long long)client * a.B;
        dz = dz_all + kb * g.PP() * g.Cout;
        xin = xin_all + kb * g.PP() * g.Cin;
        coef = coef_all ? coef_all + kb : nullptr;
        gw = a.G + (long long)client * a.ld + woff;
        gb = a.G + (long long)client * a.ld + boff;
        Mtot = a.B * g.PP();
        M = g.Cout; N = 9 * g.Cin + 1; Kd = bsz * g.PP();
        return true;
    }
    __device__ float loadA(int m, int k) const {
        const float v = dz[(long long)k * g.Cout + m];
        return coef ? v * coef[k / g.PP()] : v;
    }
    __device__ float loadB(int n, int k) const {
        if (n == 9 * g.Cin) return 1.f;
        const int tap = n / g.Cin, ci = n - tap * g.Cin;
        const int row = k + (tap / 3 - 1) * g.Wp + (tap % 3 - 1);
        return (row >= 0 && row < Mtot) ? xin[(long long)row * g.Cin + ci] : 0.f;
    }
    __device__ void store(int m, int n, float acc) {
        if (n == 9 * g.Cin) { atomicAdd(&gb[m], acc); return; }
        const int tap = n / g.Cin, ci = n - tap * g.Cin;
        atomicAdd(&gw[(m * g.Cin + ci) * 9 + tap], acc);
    }
    __device__ void finish() {}
};

// per-sample conv weight-gradient norm (dp_mode 1): group = (client, sample); the [Cout, 9*Cin+1] per-sample
// gradient tile lives in registers only -- squared, reduced with warp shuffles, one atomic per CTA.
struct ConvWgradNormProb {
    static constexpr bool A_MCONTIG = true, B_NCONTIG = true;
    flb_train_args a; ConvGeom g;
    const float* dz_all; const float* xin_all; float* norm2_all;
    const float* dz; const float* xin; float* dst; float sq; int lo, hi;
    __device__ bool setup(int group, int& M, int& N, int& Kd) {
        const int client = group / a.B, b = group % a.B;
        sq = 0.f;
        if (b >= flb_bsz(a, client)) return false;
        const long long kb = (long long)client * a.B;
        dz = dz_all + (kb + b) * g.PP() * g.Cout;
        xin = xin_all + kb * g.PP() * g.Cin;
        lo = -b * g.PP(); hi = (a.B - b) * g.PP();       // row bounds relative to this sample's first pixel
        xin += (long long)b * g.PP() * g.Cin;
        dst = norm2_all + kb + b;
        M = g.Cout; N = 9 * g.Cin + 1; Kd = g.PP();
        return true;
    }
    __device__ float loadA(int m, int k) const { return dz[(long long)k * g.Cout + m]; }
    __device__ float loadB(int n, int k) const {
        if (n == 9 * g.Cin) return 1.f;
        const int tap = n / g.Cin, ci = n - tap * g.Cin;
        const int row = k + (tap / 3 - 1) * g.Wp + (tap % 3 - 1);
        return (row >= lo && row < hi) ? xin[(long long)row * g.Cin + ci] : 0.f;
    }
    __device__ void store(int, int, float acc) { sq = fmaf(acc, acc, sq); }
    __device__ void finish() {
        const float v = flb_warp_sum(sq);
        if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(dst, v);
    }
};

struct LinFwdProb {         // out[b][n] += sum_k act[b][k] * W[n][k]     (bias added by the consumer)
    static constexpr bool A_MCONTIG = false, B_NCONTIG = false;
    flb_train_args a; int In, Out, woff; const float* act_all; float* out_all;
    const float* act; float* out; const float* w;
    __device__ bool setup(int client, int& M, int& N, int& Kd) {
        const int bsz = flb_bsz(a, client);
        if (bsz == 0) return false;
        act = act_all + (long long)client * a.B * In;
        out = out_all + (long long)client * a.B * Out;
        w = a.W + (long long)client * a.ld + woff;
        M = bsz; N = Out; Kd = In;
        return true;
    }
    __device__ float loadA(int m, int k) const { return act[(long long)m * In + k]; }
    __device__ float loadB(int n, int k) const { return __ldg(&w[(long long)n * In + k]); }
    __device__ void store(int m, int n, float acc) { atomicAdd(&out[m * Out + n], acc); }
    __device__ void finish() {}
};

struct LinDgradProb {       // dact[b][n] = sum_k dout[b][k] * W[k][n]
    static constexpr bool A_MCONTIG = false, B_NCONTIG = true;
    flb_train_args a; int In, Out, woff; const float* dout_all; float* dact_all;
    const float* dout; float* dact; const float* w;
    __device__ bool setup(int client, int& M, int& N, int& Kd) {
        const int bsz = flb_bsz(a, client);
        if (bsz == 0) return false;
        dout = dout_all + (long long)client * a.B * Out;
        dact = dact_all + (long long)client * a.B * In;
        w = a.W + (long long)client * a.ld + woff;
        M = bsz; N = In; Kd = Out;
        return true;
    }
    __device__ float loadA(int m, int k) const { return dout[m * Out + k]; }
    __device__ float loadB(int n, int k) const { return __ldg(&w[(long long)k * In + n]); }
    __device__ void store(int m, int n, float acc) { dact[(long long)m * In + n] = acc; }
    __device__ void finish() {}
};

struct LinWgradProb {       // dW[m][n] = sum_b dout[b][m] * act[b][n]   (bias gradient: head_wgrad_kernel)
    static constexpr bool A_MCONTIG = true, B_NCONTIG = true;
    flb_train_args a; int In, Out, woff, boff; const float* dout_all; const float* act_all; const float* coef_all;
    const float* dout; const float* act; const float* coef; float* gw; float* gb;
    __device__ bool setup(int client, int& M, int& N, int& Kd) {
        const int bsz = flb_bsz(a, client);
        if (bsz == 0) return false;
        dout = dout_all + (long long)client * a.B * Out;
        act = act_all + (long long)client * a.B * In;
        coef = coef_all ? coef_all + (long long)client * a.B : nullptr;
        gw = a.G + (long long)client * a.ld + woff;
        gb = a.G + (long long)client * a.ld + boff;
        M = Out; N = In; Kd = bsz;
        return true;
    }
    __device__ float loadA(int m, int k) const { const float v = dout[k * Out + m]; return coef ? v * coef[k] : v; }
    __device__ float loadB(int n, int k) const { return act[(long long)k * In + n]; }
    __device__ void store(int m, int n, float acc) { gw[(long long)m * In + n] = acc; }
    __device__ void finish() {}
};

// ------------------------------------------------------------------------------------------------
// classifier head: fc1 bias + ReLU + dropout, fc2, softmax cross-entropy, dlogits, dh.  One CTA per client.
__global__ void __launch_bounds__(256) head_fwd_bwd_kernel(flb_train_args a, SimpleCnnWs ws) {
    const int k = blockIdx.x;
    const int bsz = flb_bsz(a, k);
    if (bsz == 0) return;
    __shared__ float sh[32][129];
    __shared__ float sw2[10][129];
    __shared__ float slog[32][10];
    __shared__ float sdl[32][10];
    __shared__ float red[2];
    const int tid = threadIdx.x;
    const float* W = a.W + (long long)k * a.ld;
    const long long kb = (long long)k * a.B;
    const float keep_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
    const int step = *a.step_ctr;
    // h = dropout(relu(hpre + b1)); the multiplier (0 or 1/(1-p), and 0 where ReLU is inactive) is kept in dh's
    // slot until the backward part below overwrites it
    for (int e = tid; e < bsz * 128; e += 256) {
        const int b = e >> 7, j = e & 127;
        const float pre = ws.hpre[kb * 128 + e] + W[Off::f1b + j];
        float mult = pre > 0.f ? 1.f : 0.f;
        if (a.drop_p > 0.f) {
            bool keep;
            if (a.drop_keep) keep = a.drop_keep[kb * 128 + e] != 0;
            else {
                const flb_u4 r = flb_philox_block(a.seed ^ 0xD80F0A7ull, a.client_base + a.client_stride * k,
                                                  ((unsigned long long)a.tcount[k] << 12) + (e >> 2));
                const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
                keep = flb_u01(rr[e & 3]) >= a.drop_p;
            }
            mult = keep ? mult * keep_scale : 0.f;
        }
        const float hv = pre * mult;
        sh[b][j] = hv;
        ws.h[kb * 128 + e] = hv;
        ws.dh[kb * 128 + e] = mult;
    }
    for (int e = tid; e < 1280; e += 256) sw2[e >> 7][e & 127] = W[Off::f2w + e];
    if (tid < 2) red[tid] = 0.f;
    __syncthreads();
    for (int e = tid; e < bsz * 10; e += 256) {
        const int b = e / 10, j = e % 10;
        float acc = W[Off::f2b + j];
#pragma unroll 8
        for (int i = 0; i < 128; ++i) acc = fmaf(sh[b][i], sw2[j][i], acc);
        slog[b][j] = acc;
        ws.logits[kb * 10 + e] = acc;
    }
    __syncthreads();
    if (tid < bsz) {
        const int y = a.y[a.sample_off[k] + (long long)step * a.B + tid];
        float mx = slog[tid][0];
        int am = 0;
        for (int j = 1; j < 10; ++j) if (slog[tid][j] > mx) { mx = slog[tid][j]; am = j; }
        float se = 0.f;
        for (int j = 0; j < 10; ++j) se += expf(slog[tid][j] - mx);
        const float lse = logf(se) + mx;
        const float gs = a.dp_mode == 1 ? 1.f : 1.f / (float)bsz;          // mean reduction (training.py:90)
        for (int j = 0; j < 10; ++j) {
            const float p = expf(slog[tid][j] - lse);
            const float d = (p - (j == y ? 1.f : 0.f)) * gs;
            sdl[tid][j] = d;
            ws.dlog[kb * 10 + tid * 10 + j] = d;
        }
        atomicAdd(&red[0], lse - slog[tid][y]);
        atomicAdd(&red[1], am == y ? 1.f : 0.f);
    }
    __syncthreads();
    if (tid == 0) {
        a.loss_sum[k] += red[0] / (float)bsz;          // running_loss += loss.item()   (training.py:200)
        a.correct[k] += (int)(red[1] + 0.5f);          // correct += (pred == y).sum()  (training.py:201-203)
        a.nbatch[k] += 1;
        a.nseen[k] += bsz;
    }
    for (int e = tid; e < bsz * 128; e += 256) {
        const int b = e >> 7, j = e & 127;
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 10; ++c) acc = fmaf(sdl[b][c], sw2[c][j], acc);
        ws.dh[kb * 128 + e] = acc * ws.dh[kb * 128 + e];
    }
    // rows bsz..B-1 of dh feed the tensor-core fc1 wgrad as zeros (its K extent is the whole batch)
    for (int e = bsz * 128 + tid; e < a.B * 128; e += 256) ws.dh[kb * 128 + e] = 0.f;
}

// fc2 weight/bias gradients (tiny): one CTA per client
__global__ void __launch_bounds__(256) head_wgrad_kernel(flb_train_args a, SimpleCnnWs ws, int use_coef) {
    const int k = blockIdx.x;
    const int bsz = flb_bsz(a, k);
    if (bsz == 0) return;
    __shared__ float sh[32][129];
    __shared__ float sdl[32][10];
    const int tid = threadIdx.x;
    const long long kb = (long long)k * a.B;
    float* G = a.G + (long long)k * a.ld;
    for (int e = tid; e < bsz * 128; e += 256) sh[e >> 7][e & 127] = ws.h[kb * 128 + e];
    for (int e = tid; e < bsz * 10; e += 256) {
        const int b = e / 10;
        sdl[b][e % 10] = ws.dlog[kb * 10 + e] * (use_coef ? ws.coef[kb + b] : 1.f);
    }
    __syncthreads();
    for (int e = tid; e < 1280; e += 256) {
        const int j = e >> 7, i = e & 127;
        float acc = 0.f;
        for (int b = 0; b < bsz; ++b) acc = fmaf(sdl[b][j], sh[b][i], acc);
        G[Off::f2w + e] = acc;
    }
    if (tid < 10) {
        float acc = 0.f;
        for (int b = 0; b < bsz; ++b) acc += sdl[b][tid];
        G[Off::f2b + tid] = acc;
    }
    if (tid >= 128) {                           // fc1 bias gradient = column sums of dh
        const int j = tid - 128;
        float acc = 0.f;
        for (int b = 0; b < bsz; ++b) acc = fmaf(ws.dh[kb * 128 + b * 128 + j], use_coef ? ws.coef[kb + b] : 1.f, acc);
        G[Off::f1b + j] = acc;
    }
}

// conv2 bias gradient (tensor-core path; the fp32 path gets it as an extra GEMM column): per-sample column sums of dz
__global__ void __launch_bounds__(256) conv2_bias_grad_kernel(flb_train_args a, SimpleCnnWs ws, int use_coef) {
    const int b = blockIdx.x, k = blockIdx.y;
    if (b >= flb_bsz(a, k)) return;
    const long long kb = (long long)k * a.B + b;
    const float* dz = ws.z2 + kb * (PP2 * 64);
    const int c = threadIdx.x & 63, part = threadIdx.x >> 6;
    float acc = 0.f;
    for (int px = part; px < PP2; px += 4) acc += dz[px * 64 + c];
    __shared__ float red[4][64];
    red[part][c] = acc;
    __syncthreads();
    if (threadIdx.x < 64) {
        const float v = (red[0][c] + red[1][c] + red[2][c] + red[3][c]) * (use_coef ? ws.coef[kb] : 1.f);
        atomicAdd(&a.G[(long long)k * a.ld + Off::c2b + c], v);
    }
}

// dp_mode 1, tensor-core wgrads: TMA cannot scale an operand in flight, so the activation gradients are scaled by
// their sample's clip coefficient in place (after the norms have been taken)
__global__ void __launch_bounds__(256) scale_rows_kernel(flb_train_args a, float* buf, const float* coef, int per_sample) {
    const int b = blockIdx.x, k = blockIdx.y;
    if (b >= flb_bsz(a, k)) return;
    const long long kb = (long long)k * a.B + b;
    const float c = coef[kb];
    float4* p = reinterpret_cast<float4*>(buf + kb * per_sample);
    for (int e = threadIdx.x; e < per_sample / 4; e += 256) {
        float4 v = p[e];
        v.x *= c; v.y *= c; v.z *= c; v.w *= c;
        p[e] = v;
    }
}

// max-unpool + ReLU backward into the padded NHWC dz2 grid (every position written, pads = 0)
__global__ void __launch_bounds__(256) unpool2_kernel(flb_train_args a, SimpleCnnWs ws) {
    const int b = blockIdx.x, k = blockIdx.y;
    if (b >= flb_bsz(a, k)) return;
    const long long kb = (long long)k * a.B + b;
    const float* da2 = ws.da2 + kb * 3136;
    const float* a2 = ws.a2 + kb * 3136;
    const uint8_t* idx = ws.idx2 + kb * 3136;
    float* dz = ws.z2 + kb * (PP2 * 64);
    for (int e = threadIdx.x; e < PP2 * 64; e += 256) {
        const int c = e & 63, pos = e >> 6, h = pos >> 4, w = pos & 15;
        float v = 0.f;
        if (h < 14 && w < 14) {
            const int src = c * 49 + (h >> 1) * 7 + (w >> 1);
            if (idx[src] == ((h & 1) * 2 + (w & 1)) && a2[src] > 0.f) v = da2[src];
        }
        dz[e] = v;
    }
}

// conv1 weight + bias gradient of one sample (max-unpool + ReLU backward folded in): 32 x (9 + 1) values.
// dp_mode 0: atomically added into G.  dp_mode 1 (norm pass): stored per sample in g1ps and its squared norm
// added to norm2; the clipped sum is formed later by conv1_ps_reduce_kernel.
__global__ void __launch_bounds__(256) conv1_bwd_kernel(flb_train_args a, SimpleCnnWs ws, int per_sample) {
    const int b = blockIdx.x, k = blockIdx.y;
    if (b >= flb_bsz(a, k)) return;
    __shared__ float img[30][31];
    __shared__ float part[8][32][10];
    const int tid = threadIdx.x;
    const long long s = a.sample_off[k] + (long long)(*a.step_ctr) * a.B + b;
    const float* x = a.x + s * 784;
    for (int i = tid; i < 900; i += 256) {
        const int r = i / 30, c = i % 30;
        img[r][c] = (r >= 1 && r <= 28 && c >= 1 && c <= 28) ? x[(r - 1) * 28 + (c - 1)] : 0.f;
    }
    __syncthreads();
    const int c = tid & 31, g = tid >> 5;
    const long long kb = (long long)k * a.B + b;
    const float* da1 = ws.da1p + kb * (PP2 * 32);
    const float* a1 = ws.a1p + kb * (PP2 * 32);
    const uint8_t* idx = ws.idx1 + kb * (196 * 32);
    float acc[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] = 0.f;
    for (int pp = g; pp < 196; pp += 8) {
        const int ph = pp / 14, pw = pp % 14;
        const int o = (ph * WP2 + pw) * 32 + c;
        const float gv = a1[o] > 0.f ? da1[o] : 0.f;
        const int sel = idx[pp * 32 + c];
        const int y0 = 2 * ph + (sel >> 1), x0 = 2 * pw + (sel & 1);     // top-left of the 3x3 window in img (halo 1)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int q = 0; q < 3; ++q) acc[r * 3 + q] = fmaf(gv, img[y0 + r][x0 + q], acc[r * 3 + q]);
        acc[9] += gv;
    }
#pragma unroll
    for (int i = 0; i < 10; ++i) part[g][c][i] = acc[i];
    __syncthreads();
    float sq = 0.f;
    for (int e = tid; e < 320; e += 256) {
        const int cc = e / 10, i = e % 10;
        float v = 0.f;
#pragma unroll
        for (int gg = 0; gg < 8; ++gg) v += part[gg][cc][i];
        if (per_sample) {
            ws.g1ps[kb * 320 + e] = v;
            sq = fmaf(v, v, sq);
        } else {
            float* G = a.G + (long long)k * a.ld;
            atomicAdd(i == 9 ? &G[Off::c1b + cc] : &G[Off::c1w + cc * 9 + i], v);
        }
    }
    if (per_sample) {
        sq = flb_warp_sum(sq);
        if ((tid & 31) == 0 && sq != 0.f) atomicAdd(&ws.norm2[kb], sq);
    }
}

// dp_mode 1: conv1 gradient = sum_b coef[b] * g1ps[b]
__global__ void __launch_bounds__(320) conv1_ps_reduce_kernel(flb_train_args a, SimpleCnnWs ws) {
    const int k = blockIdx.x;
    const int bsz = flb_bsz(a, k);
    if (bsz == 0) return;
    const int e = threadIdx.x, cc = e / 10, i = e % 10;
    const long long kb = (long long)k * a.B;
    float v = 0.f;
    for (int b = 0; b < bsz; ++b) v = fmaf(ws.coef[kb + b], ws.g1ps[(kb + b) * 320 + e], v);
    float* G = a.G + (long long)k * a.ld;
    if (i == 9) G[Off::c1b + cc] = v; else G[Off::c1w + cc * 9 + i] = v;
}

// dp_mode 1: ghost norms of the two linear layers: ||dW_i||^2 = ||dout_i||^2 * ||act_i||^2, ||db_i||^2 = ||dout_i||^2
__global__ void __launch_bounds__(128) linear_ghost_norm_kernel(flb_train_args a, SimpleCnnWs ws) {
    const int b = blockIdx.x, k = blockIdx.y;
    if (b >= flb_bsz(a, k)) return;
    const long long kb = (long long)k * a.B + b;
    const int tid = threadIdx.x;
    float s_a2 = 0.f, s_dh = 0.f, s_h = 0.f, s_dl = 0.f;
    for (int e = tid; e < 3136; e += 128) { const float v = ws.a2[kb * 3136 + e]; s_a2 = fmaf(v, v, s_a2); }
    { const float v = ws.dh[kb * 128 + tid]; s_dh = v * v; const float u = ws.h[kb * 128 + tid]; s_h = u * u; }
    if (tid < 10) { const float v = ws.dlog[kb * 10 + tid]; s_dl = v * v; }
    __shared__ float red[4][4];
    s_a2 = flb_warp_sum(s_a2); s_dh = flb_warp_sum(s_dh); s_h = flb_warp_sum(s_h); s_dl = flb_warp_sum(s_dl);
    if ((tid & 31) == 0) { red[0][tid >> 5] = s_a2; red[1][tid >> 5] = s_dh; red[2][tid >> 5] = s_h; red[3][tid >> 5] = s_dl; }
    __syncthreads();
    if (tid == 0) {
        float t[4];
        for (int i = 0; i < 4; ++i) t[i] = red[i][0] + red[i][1] + red[i][2] + red[i][3];
        atomicAdd(&ws.norm2[kb], t[1] * (t[0] + 1.f) + t[3] * (t[2] + 1.f));
    }
}

__global__ void clip_coef_kernel(flb_train_args a, SimpleCnnWs ws) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.K * a.B) return;
    const float n = sqrtf(ws.norm2[i]);
    ws.coef[i] = n > a.dp_clip ? a.dp_clip / n : 1.f;         // clip rule of privacy.py:127-138, per sample
}

// ------------------------------------------------------------------------------------------------
// optimizer over [K, ld]; torch.optim semantics (training.py:244-255): Adam(lr) | SGD(lr, momentum=0.9) | AdamW(lr)
__global__ void __launch_bounds__(256) optimizer_kernel(flb_train_args a, int P) {
    const int k = blockIdx.y;
    const int bsz = flb_bsz(a, k);
    if (bsz == 0) return;
    const int t = a.tcount[k] + 1;
    float* W = a.W + (long long)k * a.ld;
    float* G = a.G + (long long)k * a.ld;
    float* M = a.M + (long long)k * a.ld;
    float* V = a.V + (long long)k * a.ld;
    // scalars are formed in double and rounded to fp32 once, like Python floats entering fp32 tensor ops
    const double bc1d = 1.0 - pow(a.beta1, (double)t), bc2d = 1.0 - pow(a.beta2, (double)t);
    const float step_size = (float)(a.lr / bc1d), bc2_sqrt = (float)sqrt(bc2d);
    const float lr = (float)a.lr, omb1 = (float)(1.0 - a.beta1), b2 = (float)a.beta2, omb2 = (float)(1.0 - a.beta2);
    const float eps = (float)a.eps, decay = (float)(1.0 - a.lr * a.weight_decay), mu = (float)a.momentum;
    const float inv_b = 1.f / (float)bsz;
    const float* zrow = a.dp_z ? a.dp_z + (long long)k * a.ld : nullptr;
    const int P4 = (P + 3) >> 2;
    for (int c4 = blockIdx.x * 256 + threadIdx.x; c4 < P4; c4 += gridDim.x * 256) {
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        if (a.dp_mode == 1 && a.dp_sigma > 0.f && !zrow) {
            const float4 zz = flb_normal4(a.seed, a.client_base + a.client_stride * k, ((unsigned long long)t << 32) + c4);
            z[0] = zz.x; z[1] = zz.y; z[2] = zz.z; z[3] = zz.w;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int p = c4 * 4 + e;
            if (p >= P) break;
            float g = G[p];
            if (a.dp_mode == 1) g = (g + a.dp_sigma * (zrow ? zrow[p] : z[e])) * inv_b;   // (sum clipped + N(0, sigma^2)) / B
            float w = W[p];
            if (a.opt == 1) {                               // SGD with momentum, dampening 0
                const float buf = t == 1 ? g : fmaf(mu, M[p], g);
                M[p] = buf;
                w = w - lr * buf;
            } else {
                if (a.opt == 2) w = w * decay;      // AdamW decoupled decay
                float m = M[p], v = V[p];
                m = m + (g - m) * omb1;                           // exp_avg.lerp_(grad, 1 - beta1)
                v = v * b2 + omb2 * g * g;                  // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
                M[p] = m; V[p] = v;
                const float denom = sqrtf(v) / bc2_sqrt + eps;
                w = w - step_size * (m / denom);                             // param.addcdiv_(m, denom, -step_size)
            }
            W[p] = w;
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
    if (k == 0) *a.step_ctr = 0;
}

int check_args(const flb_train_args* a) {
    FLB_CHECK_ARG(a != nullptr, "flb_train: null args");
    FLB_CHECK_ARG(a->model == 0, "flb_train: model %d not supported by this entry (0 = simple_cnn)", a->model);
    FLB_CHECK_ARG(a->K >= 1 && a->K <= 1024 && a->B >= 1 && a->B <= 32, "flb_train: need 1 <= K <= 1024 and 1 <= B <= 32 (K=%d B=%d)", a->K, a->B);
    FLB_CHECK_ARG(a->ld >= Off::P, "flb_train: ld %lld < %d parameters", a->ld, Off::P);
    FLB_CHECK_ARG(a->x && a->y && a->sample_off && a->nsamples && a->step_ctr && a->W && a->G && a->M && a->V &&
                  a->tcount && a->ws && a->loss_sum && a->correct && a->nbatch && a->nseen, "flb_train: null device pointer in args");
    FLB_CHECK_ARG(a->opt >= 0 && a->opt <= 2, "flb_train: Unknown optimizer type: %d", a->opt);
    FLB_CHECK_ARG(a->drop_p >= 0.f && a->drop_p < 1.f, "flb_train: dropout probability must be in [0, 1)");
    FLB_CHECK_ARG(a->precision == 0 || a->precision == 1, "flb_train: precision must be 0 (fp32) or 1 (tf32 tensor cores)");
    FLB_CHECK_ARG(a->dp_mode == 0 || a->dp_mode == 1, "flb_train: dp_mode must be 0 or 1");
    return FLB_OK;
}

const ConvGeom kConv2{32, 64, 14, 14, 16, 16};

int forward(const flb_train_args& a, const SimpleCnnWs& ws, cudaStream_t st) {
    const int K = a.K, B = a.B;
    const dim3 per_sample(B, K);
    FLB_CUDA(cudaMemsetAsync(ws.hpre, 0, sizeof(float) * (size_t)K * B * 128, st));
    MARK("begin");
    conv1_fwd_pool_kernel<<<per_sample, 256, 0, st>>>(a, ws);
    MARK("conv1_fwd_pool");
    const int tcm = tc_mask_of(a);
    if (tcm & TC_CONV2_FWD) {
        if (int rc = tc::conv_fwd_32_64(a, kConv2, ws.a1p, ws.z2, Off::c2w, Off::c2b, st)) return rc;
    } else {
        ConvFwdProb p{}; p.a = a; p.g = kConv2; p.xin_all = ws.a1p; p.z_all = ws.z2; p.woff = Off::c2w; p.boff = Off::c2b;
        simt::launch(p, B * PP2, 64, 1, K, st);
    }
    MARK("conv2_fwd");
    pool2_kernel<<<per_sample, 256, 0, st>>>(a, ws);
    MARK("pool2");
    if (tcm & TC_FC1_FWD) {
        if (int rc = tc::fc_fwd_3136_128(a, ws.a2, ws.hpre, Off::f1w, 7, st)) return rc;
    } else {
        LinFwdProb p{}; p.a = a; p.In = 3136; p.Out = 128; p.woff = Off::f1w; p.act_all = ws.a2; p.out_all = ws.hpre;
        simt::launch(p, B, 128, 14, K, st);
    }
    MARK("fc1_fwd");
    head_fwd_bwd_kernel<<<K, 256, 0, st>>>(a, ws);
    MARK("head_fwd_bwd");
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

int forward_backward(const flb_train_args& a, cudaStream_t st) {
    SimpleCnnWs ws;
    simplecnn_ws_carve(a.ws, a.K, a.B, &ws);
    const int K = a.K, B = a.B;
    const dim3 per_sample(B, K);
    // gradients that are accumulated with atomics start from zero: everything except fc1.weight/fc2 (plain stores)
    FLB_CUDA(cudaMemset2DAsync(a.G, a.ld * sizeof(float), 0, Off::f1w * sizeof(float), K, st));
    if (a.dp_mode == 1) FLB_CUDA(cudaMemsetAsync(ws.norm2, 0, sizeof(float) * (size_t)K * B, st));
    if (int rc = forward(a, ws, st)) return rc;

    // ---- activation gradients ----
    const int tcm = tc_mask_of(a);
    if (tcm & TC_FC1_DGRAD) {
        if (int rc = tc::fc_dgrad_3136_128(a, ws.dh, ws.da2, Off::f1w, st)) return rc;
    } else {
        LinDgradProb p{}; p.a = a; p.In = 3136; p.Out = 128; p.woff = Off::f1w; p.dout_all = ws.dh; p.dact_all = ws.da2;
        simt::launch(p, B, 3136, 1, K, st);
    }
    MARK("fc1_dgrad");
    unpool2_kernel<<<per_sample, 256, 0, st>>>(a, ws);
    MARK("unpool2");
    if (tcm & TC_CONV2_DGRAD) {
        if (int rc = tc::conv_dgrad_32_64(a, kConv2, ws.z2, ws.da1p, Off::c2w, st)) return rc;
    } else {
        ConvDgradProb p{}; p.a = a; p.g = kConv2; p.dz_all = ws.z2; p.dx_all = ws.da1p; p.woff = Off::c2w;
        simt::launch(p, B * PP2, 32, 1, K, st);
    }
    MARK("conv2_dgrad");

    // ---- per-sample clip coefficients (dp_mode 1) ----
    const float* coef = nullptr;
    if (a.dp_mode == 1) {
        linear_ghost_norm_kernel<<<per_sample, 128, 0, st>>>(a, ws);
        {
            ConvWgradNormProb p{}; p.a = a; p.g = kConv2; p.dz_all = ws.z2; p.xin_all = ws.a1p; p.norm2_all = ws.norm2;
            simt::launch(p, 64, 289, 1, K * B, st);
        }
        conv1_bwd_kernel<<<per_sample, 256, 0, st>>>(a, ws, 1);
        clip_coef_kernel<<<flb_cdiv(K * B, 256), 256, 0, st>>>(a, ws);
        coef = ws.coef;
        MARK("per_sample_norms");
    }

    // ---- weight gradients ----
    head_wgrad_kernel<<<K, 256, 0, st>>>(a, ws, a.dp_mode == 1);
    MARK("head_wgrad");
    if (tcm & TC_FC1_WGRAD) {
        if (coef) scale_rows_kernel<<<per_sample, 256, 0, st>>>(a, ws.dh, coef, 128);
        if (int rc = tc::fc_wgrad_3136_128(a, ws.dh, ws.a2, Off::f1w, st)) return rc;
    } else {
        LinWgradProb p{}; p.a = a; p.In = 3136; p.Out = 128; p.woff = Off::f1w; p.boff = Off::f1b;
        p.dout_all = ws.dh; p.act_all = ws.a2; p.coef_all = coef;
        simt::launch(p, 128, 3136, 1, K, st);
    }
    MARK("fc1_wgrad");
    if (tcm & TC_CONV2_WGRAD) {
        conv2_bias_grad_kernel<<<per_sample, 256, 0, st>>>(a, ws, coef != nullptr);
        if (coef) scale_rows_kernel<<<per_sample, 256, 0, st>>>(a, ws.z2, coef, PP2 * 64);
        if (int rc = tc::conv_wgrad_32_64(a, kConv2, ws.a1p, ws.z2, Off::c2w, 16, st)) return rc;
    } else {
        ConvWgradProb p{}; p.a = a; p.g = kConv2; p.dz_all = ws.z2; p.xin_all = ws.a1p; p.coef_all = coef;
        p.woff = Off::c2w; p.boff = Off::c2b;
        simt::launch(p, 64, 289, 16, K, st);
    }
    MARK("conv2_wgrad");
    if (a.dp_mode == 1) conv1_ps_reduce_kernel<<<K, 320, 0, st>>>(a, ws);
    else conv1_bwd_kernel<<<per_sample, 256, 0, st>>>(a, ws, 0);
    MARK("conv1_wgrad");
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

}  // namespace

extern "C" long long flb_train_ws_bytes(int model, int K, int B) {
    if (model != 0 || K < 1 || B < 1) return -1;
    return (long long)simplecnn_ws_carve(nullptr, K, B, nullptr);
}

extern "C" long long flb_train_ws_offset(int model, int K, int B, const char* name) {
    if (model != 0 || K < 1 || B < 1 || !name) return -1;
    SimpleCnnWs ws;
    simplecnn_ws_carve((void*)0, K, B, &ws);
#define FIELD(f) if (!strcmp(name, #f)) return (long long)(uintptr_t)ws.f;
    FIELD(a1p) FIELD(idx1) FIELD(z2) FIELD(a2) FIELD(idx2) FIELD(hpre) FIELD(h) FIELD(logits) FIELD(dlog) FIELD(dh)
    FIELD(da2) FIELD(da1p) FIELD(norm2) FIELD(coef) FIELD(g1ps)
#undef FIELD
    return -1;
}

extern "C" int flb_train_begin_epoch(const flb_train_args* a, void* stream) {
    if (int rc = check_args(a)) return rc;
    begin_epoch_kernel<<<flb_cdiv(a->K, 256), 256, 0, (cudaStream_t)stream>>>(*a);
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

extern "C" int flb_train_forward(const flb_train_args* a, void* stream) {
    if (int rc = check_args(a)) return rc;
    SimpleCnnWs ws;
    simplecnn_ws_carve(a->ws, a->K, a->B, &ws);
    return forward(*a, ws, (cudaStream_t)stream);
}

extern "C" int flb_train_advance(const flb_train_args* a, void* stream) {
    if (int rc = check_args(a)) return rc;
    advance_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(*a);
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

extern "C" int flb_train_forward_backward(const flb_train_args* a, void* stream) {
    if (int rc = check_args(a)) return rc;
    return forward_backward(*a, (cudaStream_t)stream);
}

extern "C" int flb_train_step(const flb_train_args* a, void* stream) {
    if (int rc = check_args(a)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = forward_backward(*a, st)) return rc;
    const int blocks = max(1, min(flb_cdiv(Off::P / 4, 256), (flb_num_sms() * 8 + a->K - 1) / a->K));
    optimizer_kernel<<<dim3(blocks, a->K), 256, 0, st>>>(*a, Off::P);
    MARK("optimizer");
    advance_kernel<<<1, 1024, 0, st>>>(*a);
    MARK("advance");
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

// number of kernel launches (memsets excluded) one flb_train_step issues for these args
extern "C" int flb_train_step_launches(const flb_train_args* a) {
    if (!a) return -1;
    return a->dp_mode == 1 ? 18 : 14;
}

// One step with a CUDA event after every kernel.  Synchronises the stream (profiling aid, not the product path).
// names_out receives '\n'-separated labels; ms_out[i] = device time of labelled segment i.  Returns the segment count.
extern "C" int flb_train_step_profiled(const flb_train_args* a, void* stream, char* names_out, int names_cap,
                                       float* ms_out, int max_n) {
    if (int rc = check_args(a)) return rc;
    FLB_CHECK_ARG(names_out && ms_out && names_cap > 0 && max_n > 0, "flb_train_step_profiled: bad output buffers");
    for (int i = 0; i < 48; ++i) FLB_CUDA(cudaEventCreate(&g_prof.ev[i]));
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
    for (int i = 0; i < 48; ++i) cudaEventDestroy(g_prof.ev[i]);
    return rc == FLB_OK ? n : rc;
}
