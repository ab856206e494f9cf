// CIFAR10CNN (reference src/shared/models_pytorch.py:100-165) local-training step, batched over all resident clients:
//   3 x [conv3x3+bias -> BatchNorm2d (batch statistics) -> ReLU, conv3x3+bias -> BN -> ReLU, maxpool2, dropout]
//   -> flatten 2048 -> fc 512 + ReLU + dropout -> fc 256 + ReLU + dropout -> fc 10,
// mean cross-entropy (src/shared/training.py:90,193), backward, optimizer step (training.py:244-255; train_api.cu).
//
// Activations live NHWC on per-image padded grids (train_common.cuh ConvGeom): 32x32 -> 33-wide rows, 1096 rows per
// image; 16x16 -> 17 / 296; 8x8 -> 9 / 88 (multiples of 8 so that the tensor-core wgrad can walk pixels 8 at a time).
// A conv writes its pre-BN output z on the whole grid (pad positions hold don't-care values); the BN consumers reduce /
// apply over REAL pixels only and write zeros at the pads, so every conv input has zero pads.
//
// BatchNorm per client and layer (BatchNorm2d defaults eps 1e-5, momentum 0.1): the per-channel sums are reduced in
// fp32 per thread and accumulated across CTAs in double (acc[k][0..3][448] = sum z, sum z^2, sum dy*xhat, sum dy), so
// mean / biased variance / dgamma / dbeta are formed once, in double, by every consumer.  Running statistics are
// client-local buffers (never federated, models_pytorch.py:25-27) and are only used in eval mode.
#include "gemm_probs.cuh"
#include "philox.cuh"
#include <string.h>

namespace tc {
int conv_fwd(const flb_train_args& a, const ConvGeom& g, const float* xin, float* z, const float* wt, long long ldt, int boff, cudaStream_t st,
             double* bn_acc = nullptr, int bn_coff = 0, int bn_stride = 0);
bool conv_fwd_fuses_stats(int cin, int cout);
int conv_dgrad(const flb_train_args& a, const ConvGeom& g, const float* dz, float* dx, const float* wt, long long ldt, cudaStream_t st);
int conv_wgrad(const flb_train_args& a, const ConvGeom& g, const float* xin, const float* dz, float* gt, long long ldt, cudaStream_t st);
int conv_wgrad_norm(const flb_train_args& a, const ConvGeom& g, const float* xin, const float* dz, float* norm2, cudaStream_t st);
int fc_fwd(const flb_train_args& a, const float* act, float* out, int in, int outf, int woff, int splits, cudaStream_t st);
int fc_dgrad(const flb_train_args& a, const float* dout, float* dact, int in, int outf, int woff, cudaStream_t st);
int fc_wgrad(const flb_train_args& a, const float* dout, const float* act, int in, int outf, int woff, cudaStream_t st, bool adam = false);
}
#include <stdlib.h>
namespace tc {
}  // namespace tc

namespace {

constexpr int NCONV = 6;
constexpr int BN_CH = 448;                       // 32 + 32 + 64 + 64 + 128 + 128
constexpr float BN_EPS = 1e-5f, BN_MOM = 0.1f;
constexpr int DROP_PER_SAMPLE = 8192 + 4096 + 2048 + 512 + 256;     // injected keep-mask floats per sample (NCHW order)

struct Net {
    int cin[NCONV], cout[NCONV], cw[NCONV], cb[NCONV], bw[NCONV], bb[NCONV], coff[NCONV];
    int toff[NCONV];                 // tap-major copy (layers 2..6; conv1 with Cin = 3 stays on CUDA cores)
    int f1w, f1b, f2w, f2b, f3w, f3b, P, ldt;
};
constexpr Net make_net() {
    Net n{};
    const int ci[NCONV] = {3, 32, 32, 64, 64, 128}, co[NCONV] = {32, 32, 64, 64, 128, 128};
    int off = 0, c = 0, t = 0;
    for (int i = 0; i < NCONV; ++i) {
        n.toff[i] = t;
        if (i > 0) t += co[i] * ci[i] * 9;
        n.cin[i] = ci[i]; n.cout[i] = co[i];
        n.cw[i] = off; off += co[i] * ci[i] * 9;
        n.cb[i] = off; off += co[i];
        n.bw[i] = off; off += co[i];
        n.bb[i] = off; off += co[i];
        n.coff[i] = c; c += co[i];
    }
    n.f1w = off; off += 512 * 2048;
    n.f1b = off; off += 512;
    n.f2w = off; off += 256 * 512;
    n.f2b = off; off += 256;
    n.f3w = off; off += 10 * 256;
    n.f3b = off; off += 10;
    n.P = off;
    n.ldt = t;
    return n;
}
constexpr Net kNet = make_net();
static_assert(kNet.P == 1470890, "CIFAR10CNN parameter count (SURVEY.md section 2a)");

// geometry of the three resolutions; Cin / Cout are filled per layer
constexpr ConvGeom geom(int level, int cin, int cout) {
    return level == 0 ? ConvGeom{cin, cout, 32, 32, 34, 33, 1096}
         : level == 1 ? ConvGeom{cin, cout, 16, 16, 18, 17, 296}
                      : ConvGeom{cin, cout, 8, 8, 10, 9, 88};
}
constexpr int PP32 = 1096, PP16 = 296, PP8 = 88;
constexpr int kNetLdt = kNet.ldt;
constexpr int kC1W = kNet.cw[0], kC1B = kNet.cb[0];      // conv1 offsets as scalars (usable in device code)

struct CifarWs {
    float *z1, *y1, *z2, *p1;          // grid 32: conv1 out, bn1+relu, conv2 out; grid 16: pooled+dropped (conv3 input)
    float *z3, *y3, *z4, *p2;          // grid 16 ...; grid 8: p2
    float *z5, *y5, *z6, *a;           // grid 8 ...; a: [2048] NCHW-flattened fc1 input
    uint8_t *i1, *i2, *i3;             // pool argmax (bits 0-1) | dropped (bit 2)
    float *hpre1, *h1, *m1, *hpre2, *h, *logits, *dlog, *dh2, *dh1, *da;
    float *d32a, *d32b, *d16p, *d16a, *d16b, *d8p, *d8a, *d8b;
    double* acc;                       // [K][4][448]
    float *wt, *gt;                    // [K][ldt] tap-major conv weights / weight gradients (tensor-core path)
    float *norm2, *coef;               // [K][B] per-sample squared gradient norms / clip coefficients (dp_mode 1)
    float* bnps;                       // [K][B][2][448] per-sample BatchNorm (dgamma | dbeta) (dp_mode 1)
};

size_t carve(void* base, int K, int B, CifarWs* ws) {
    size_t off = 0;
    char* p = (char*)base;
    const size_t KB = (size_t)K * B;
#define CARVE(field, type, count)                                  \
    do {                                                           \
        if (ws) ws->field = (type*)(p + off);                      \
        off = flb_align256(off + sizeof(type) * (size_t)(count));  \
    } while (0)
    CARVE(z1, float, KB * PP32 * 32); CARVE(y1, float, KB * PP32 * 32); CARVE(z2, float, KB * PP32 * 32);
    CARVE(p1, float, KB * PP16 * 32);
    CARVE(z3, float, KB * PP16 * 64); CARVE(y3, float, KB * PP16 * 64); CARVE(z4, float, KB * PP16 * 64);
    CARVE(p2, float, KB * PP8 * 64);
    CARVE(z5, float, KB * PP8 * 128); CARVE(y5, float, KB * PP8 * 128); CARVE(z6, float, KB * PP8 * 128);
    CARVE(a, float, KB * 2048);
    CARVE(i1, uint8_t, KB * 8192); CARVE(i2, uint8_t, KB * 4096); CARVE(i3, uint8_t, KB * 2048);
    CARVE(hpre1, float, KB * 512); CARVE(h1, float, KB * 512); CARVE(m1, float, KB * 512);
    CARVE(hpre2, float, KB * 256); CARVE(h, float, KB * 256);
    CARVE(logits, float, KB * 10); CARVE(dlog, float, KB * 10);
    CARVE(dh2, float, KB * 256); CARVE(dh1, float, KB * 512); CARVE(da, float, KB * 2048);
    CARVE(d32a, float, KB * PP32 * 32); CARVE(d32b, float, KB * PP32 * 32);
    CARVE(d16p, float, KB * PP16 * 32); CARVE(d16a, float, KB * PP16 * 64); CARVE(d16b, float, KB * PP16 * 64);
    CARVE(d8p, float, KB * PP8 * 64); CARVE(d8a, float, KB * PP8 * 128); CARVE(d8b, float, KB * PP8 * 128);
    CARVE(acc, double, (size_t)K * 4 * BN_CH);
    CARVE(wt, float, (size_t)K * kNetLdt);
    CARVE(gt, float, (size_t)K * kNetLdt);
    CARVE(norm2, float, KB);
    CARVE(coef, float, KB);
    CARVE(bnps, float, KB * 2 * BN_CH);
#undef CARVE
    return off;
}

// ---- conv1 (Cin = 3): the input is the raw NCHW sample; K = 27 is too thin for a tensor-core tile -----------------
// Direct 3->32 stencils (one CTA per sample): thread = (output channel, strip lane); the 27 weights sit in registers,
// the zero-haloed input planes in shared memory (row pitch 36 floats so that a 4-pixel strip's 6 inputs are one 16 B and
// one 8 B load; every lane of a warp reads the same inputs: broadcast).  Each thread works on strips of 4 horizontally
// adjacent pixels, so a (channel-in, kernel-row) pair costs 2 shared loads per 12 FMAs instead of 3 per 3.
__device__ __forceinline__ void load_img3(float (*img)[34][36], const float* x, int tid) {
    for (int i = tid; i < 3 * 34 * 36; i += 256) {
        const int ci = i / (34 * 36), r = (i % (34 * 36)) / 36, c = i % 36;
        img[ci][r][c] = (r >= 1 && r <= 32 && c >= 1 && c <= 32) ? x[ci * 1024 + (r - 1) * 32 + (c - 1)] : 0.f;
    }
}

// bn_acc (optional): the sample's per-channel (sum z, sum z^2) are added to the client's BatchNorm accumulators (rows 0 / 1
// of [K][4][BN_CH], layer offset 0), which saves the separate statistics pass over z1.
__global__ void __launch_bounds__(256) conv1_fwd_kernel(flb_train_args a, float* z_all, double* bn_acc) {
    const int b = blockIdx.x, k = blockIdx.y;
    if (b >= flb_bsz(a, k)) return;
    __shared__ __align__(16) float img[3][34][36];
    const int tid = threadIdx.x;
    load_img3(img, a.x + (a.sample_off[k] + (long long)(*a.step_ctr) * a.B + b) * 3072, tid);
    const int c = tid & 31, g = tid >> 5;
    const float* W = a.W + (long long)k * a.ld;
    float w[27];
#pragma unroll
    for (int i = 0; i < 27; ++i) w[i] = W[kC1W + c * 27 + i];
    const float bias = W[kC1B + c];
    __syncthreads();
    float* z = z_all + ((long long)k * a.B + b) * PP32 * 32;
    float s0 = 0.f, s1 = 0.f;
    for (int s = g; s < 256; s += 8) {                    // strip = 4 pixels (h, w0 .. w0 + 3)
        const int h = s >> 3, w0 = (s & 7) * 4;
        float acc[4] = {bias, bias, bias, bias};
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float4 p0 = *reinterpret_cast<const float4*>(&img[ci][h + r][w0]);
                const float2 p1 = *reinterpret_cast<const float2*>(&img[ci][h + r][w0 + 4]);
                const float in[6] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y};
#pragma unroll
                for (int q = 0; q < 3; ++q)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[j] = fmaf(w[ci * 9 + r * 3 + q], in[j + q], acc[j]);
            }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            z[(h * 33 + w0 + j) * 32 + c] = acc[j];
            s0 += acc[j];
            s1 = fmaf(acc[j], acc[j], s1);
        }
    }
    if (bn_acc) {
        __syncthreads();                                   // img is dead: reuse it for the cross-warp reduction
        float (*red)[8][32] = reinterpret_cast<float (*)[8][32]>(&img[0][0][0]);
        red[0][g][c] = s0;
        red[1][g][c] = s1;
        __syncthreads();
        if (tid < 64) {
            const int which = tid >> 5;
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) t += red[which][i][c];
            atomicAdd(bn_acc + (long long)k * 4 * BN_CH + which * BN_CH + c, (double)t);
        }
    }
}

// dW[c][ci][tap] += sum_px dz[px][c] * x[ci][px + shift(tap)], db[c] += sum_px dz[px][c]
// NORM (dp_mode 1): the sample's [32][27 + 1] gradient is only squared into norm2[client, sample]
template <bool NORM>
__global__ void __launch_bounds__(256) conv1_wgrad_kernel(flb_train_args a, const float* dz_all, float* norm2) {
    const int b = blockIdx.x, k = blockIdx.y;
    if (b >= flb_bsz(a, k)) return;
    __shared__ __align__(16) float img[3][34][36];
    __shared__ float part[8][32][29];
    const int tid = threadIdx.x;
    load_img3(img, a.x + (a.sample_off[k] + (long long)(*a.step_ctr) * a.B + b) * 3072, tid);
    __syncthreads();
    const int c = tid & 31, g = tid >> 5;
    const float* dz = dz_all + ((long long)k * a.B + b) * PP32 * 32;
    float acc[28];
#pragma unroll
    for (int i = 0; i < 28; ++i) acc[i] = 0.f;
    for (int s = g; s < 256; s += 8) {
        const int h = s >> 3, w0 = (s & 7) * 4;
        float gv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) gv[j] = dz[(h * 33 + w0 + j) * 32 + c];
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float4 p0 = *reinterpret_cast<const float4*>(&img[ci][h + r][w0]);
                const float2 p1 = *reinterpret_cast<const float2*>(&img[ci][h + r][w0 + 4]);
                const float in[6] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y};
#pragma unroll
                for (int q = 0; q < 3; ++q)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[ci * 9 + r * 3 + q] = fmaf(gv[j], in[j + q], acc[ci * 9 + r * 3 + q]);
            }
        acc[27] += (gv[0] + gv[1]) + (gv[2] + gv[3]);
    }
#pragma unroll
    for (int i = 0; i < 28; ++i) part[g][c][i] = acc[i];
    __syncthreads();
    float* G = a.G + (long long)k * a.ld;
    float sq = 0.f;
    for (int e = tid; e < 32 * 28; e += 256) {
        const int cc = e / 28, i = e % 28;
        float v = 0.f;
#pragma unroll
        for (int gg = 0; gg < 8; ++gg) v += part[gg][cc][i];
        if (NORM) sq = fmaf(v, v, sq);
        else atomicAdd(i == 27 ? &G[kC1B + cc] : &G[kC1W + cc * 27 + i], v);
    }
    if (NORM) {
        sq = flb_warp_sum(sq);
        if ((tid & 31) == 0 && sq != 0.f) atomicAdd(&norm2[(long long)k * a.B + b], sq);
    }
}

// ---- BatchNorm ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool real_px(const ConvGeom& g, int r) {       // r = row within the client's [B * PP] rows
    const int rr = r % g.PP();
    const int h = rr / g.Wp;
    return h < g.H && (rr - h * g.Wp) < g.W;
}

// row (within the client's [B * PP] rows) of the q-th REAL pixel, q in [0, bsz * H * W): shifts only (H, W powers of two)
__device__ __forceinline__ int real_row(const ConvGeom& g, int q, int hw_shift, int w_shift) {
    const int b = q >> hw_shift, p = q & ((1 << hw_shift) - 1);
    return b * g.PP() + (p >> w_shift) * g.Wp + (p & (g.W - 1));
}
// row of the j-th PAD position of a client, j in [0, bsz * (PP - H*W)): the pad column of every image row, then the tail rows
__device__ __forceinline__ int pad_row(const ConvGeom& g, int j) {
    const int npad = g.PP() - g.H * g.W, b = j / npad, t = j - b * npad;
    const int rr = t < g.H * (g.Wp - g.W) ? (t / (g.Wp - g.W)) * g.Wp + g.W + t % (g.Wp - g.W) : g.H * g.Wp + (t - g.H * (g.Wp - g.W));
    return b * g.PP() + rr;
}

// per-channel mean / invstd of client k for this step (train) or from the running buffers (eval)
__device__ __forceinline__ void bn_moments(const flb_train_args& a, const double* acc, int k, int ch, int n_real,
                                           float& mean, float& invstd, float& var_b) {
    if (a.eval_mode) {
        const float* run = a.bn_running + (long long)k * 2 * BN_CH;
        mean = run[ch];
        var_b = run[BN_CH + ch];
    } else {
        const double* A = acc + (long long)k * 4 * BN_CH;
        const double m = A[ch] / n_real;
        double v = A[BN_CH + ch] / n_real - m * m;
        if (v < 0.0) v = 0.0;
        mean = (float)m;
        var_b = (float)v;
    }
    invstd = 1.0f / sqrtf(var_b + BN_EPS);
}

// column sums over the real pixels of client k: MODE 0: (sum z, sum z^2) -> acc[0], acc[1];
// MODE 1: g = relu-masked upstream gradient; (sum g * xhat, sum g) -> acc[2], acc[3].
// thread = (channel quad, row lane): 16 B loads, two rows in flight per thread.
// relu_gw >= 0 (MODE 1, layers whose dy comes from a dgrad): the ReLU mask is recomputed from z with the forward's own
// arithmetic (z * alpha + beta > 0) instead of reading the activation tensor back.
template <int C, int MODE>
__global__ void __launch_bounds__(256) bn_reduce_kernel(flb_train_args a, ConvGeom g, const float* z_all, const float* dy_all,
                                                        double* acc, int coff, int relu_gw, int relu_gb) {
    const int k = blockIdx.y;
    const int bsz = flb_bsz(a, k);
    if (bsz == 0) return;
    constexpr int C4 = C / 4, RL = 256 / C4;
    const int tid = threadIdx.x, cq = tid % C4, rl = tid / C4;
    const int PP = g.PP(), nreal = bsz * g.H * g.W;
    const int hw_shift = 31 - __clz(g.H * g.W), w_shift = 31 - __clz(g.W);
    const int per = (nreal + gridDim.x - 1) / gridDim.x, r0 = blockIdx.x * per, r1 = min(nreal, r0 + per);
    const long long base = (long long)k * a.B * PP * C;
    const float4* z4 = reinterpret_cast<const float4*>(z_all + base);
    const float4* dy4 = MODE == 1 ? reinterpret_cast<const float4*>(dy_all + base) : nullptr;
    float mean[4] = {0.f, 0.f, 0.f, 0.f}, invstd[4] = {0.f, 0.f, 0.f, 0.f}, alpha[4], beta[4];
    if (MODE == 1) {
        float vb;
        const float* W = a.W + (long long)k * a.ld;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            bn_moments(a, acc, k, coff + cq * 4 + e, bsz * g.H * g.W, mean[e], invstd[e], vb);
            alpha[e] = relu_gw >= 0 ? invstd[e] * W[relu_gw + cq * 4 + e] : 0.f;
            beta[e] = relu_gw >= 0 ? W[relu_gb + cq * 4 + e] - mean[e] * alpha[e] : 0.f;
        }
    }
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
    auto row = [&](int q_) {                               // q_: index of a real pixel
        if (q_ >= r1) return;
        const long long e = (long long)real_row(g, q_, hw_shift, w_shift) * C4 + cq;
        const float4 zv = z4[e];
        const float zz[4] = {zv.x, zv.y, zv.z, zv.w};
        if (MODE == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { s0[q] += zz[q]; s1[q] = fmaf(zz[q], zz[q], s1[q]); }
        } else {
            const float4 gv4 = dy4[e];
            float gv[4] = {gv4.x, gv4.y, gv4.z, gv4.w};
            if (relu_gw >= 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (!(__fadd_rn(__fmul_rn(zz[q], alpha[q]), beta[q]) > 0.f)) gv[q] = 0.f;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) { s0[q] = fmaf(gv[q], (zz[q] - mean[q]) * invstd[q], s0[q]); s1[q] += gv[q]; }
        }
    };
    for (int r = r0 + rl; r < r1; r += 2 * RL) { row(r); row(r + RL); }
    __shared__ float red[8][256];
#pragma unroll
    for (int q = 0; q < 4; ++q) { red[q][tid] = s0[q]; red[4 + q][tid] = s1[q]; }
    __syncthreads();
    if (tid < C) {
        const int q = tid & 3, cq2 = tid >> 2;
        float t0 = 0.f, t1 = 0.f;
        for (int i = 0; i < RL; ++i) { t0 += red[q][i * C4 + cq2]; t1 += red[4 + q][i * C4 + cq2]; }
        double* A = acc + (long long)k * 4 * BN_CH + (MODE == 0 ? 0 : 2 * BN_CH) + coff + tid;
        atomicAdd(A, (double)t0);
        atomicAdd(A + BN_CH, (double)t1);
    }
}

// y = relu(bn(z)) on the real pixels, 0 on the pads; CTA (0, k) also updates the client's running statistics
template <int C>
__global__ void __launch_bounds__(256) bn_relu_apply_kernel(flb_train_args a, ConvGeom g, const float* z_all, float* y_all,
                                                            const double* acc, int coff, int gwoff, int gboff) {
    const int k = blockIdx.y;
    const int bsz = flb_bsz(a, k);
    if (bsz == 0) return;
    __shared__ float s_scale[C], s_beta[C];
    const int tid = threadIdx.x;
    const int n_real = bsz * g.H * g.W;
    const float* W = a.W + (long long)k * a.ld;
    if (tid < C) {
        float mean, invstd, vb;
        bn_moments(a, acc, k, coff + tid, n_real, mean, invstd, vb);
        // ATen's CPU kernel (batch_norm_cpu_transform_input): alpha = invstd * weight, beta = bias - mean * alpha,
        // out = in * alpha + beta -- same association here so that near-ties in the following max-pool break alike
        s_scale[tid] = invstd * W[gwoff + tid];
        s_beta[tid] = W[gboff + tid] - mean * s_scale[tid];
        if (blockIdx.x == 0 && !a.eval_mode) {
            float* run = a.bn_running + (long long)k * 2 * BN_CH + coff + tid;
            run[0] = (1.f - BN_MOM) * run[0] + BN_MOM * mean;
            const float unb = n_real > 1 ? vb * ((float)n_real / (float)(n_real - 1)) : vb;
            run[BN_CH] = (1.f - BN_MOM) * run[BN_CH] + BN_MOM * unb;
        }
    }
    __syncthreads();
    const int PP = g.PP();
    const long long base = (long long)k * a.B * PP * C;
    const float4* z4 = reinterpret_cast<const float4*>(z_all + base);
    float4* y4 = reinterpret_cast<float4*>(y_all + base);
    constexpr int C4 = C / 4, RL = 256 / C4;
    // real pixels only: this kernel is y's only writer, so the pads keep the zeros the workspace was created with.
    // thread = (channel quad, pixel lane): its scale / shift live in registers, rows come from shifts
    const int cq = tid % C4, rl = tid / C4, c = cq * 4;
    const int hw_shift = 31 - __clz(g.H * g.W), w_shift = 31 - __clz(g.W);
    const float sc[4] = {s_scale[c], s_scale[c + 1], s_scale[c + 2], s_scale[c + 3]};
    const float sb[4] = {s_beta[c], s_beta[c + 1], s_beta[c + 2], s_beta[c + 3]};
    for (int q = blockIdx.x * RL + rl; q < n_real; q += gridDim.x * RL) {
        const long long e = (long long)real_row(g, q, hw_shift, w_shift) * C4 + cq;
        const float4 v = z4[e];
        float4 o;
        o.x = fmaxf(__fadd_rn(__fmul_rn(v.x, sc[0]), sb[0]), 0.f);
        o.y = fmaxf(__fadd_rn(__fmul_rn(v.y, sc[1]), sb[1]), 0.f);
        o.z = fmaxf(__fadd_rn(__fmul_rn(v.z, sc[2]), sb[2]), 0.f);
        o.w = fmaxf(__fadd_rn(__fmul_rn(v.w, sc[3]), sb[3]), 0.f);
        y4[e] = o;
    }
}

// dropout decision of element e of drop layer `layer` (0..4) for local client k, sample b
__device__ __forceinline__ bool drop_keep(const flb_train_args& a, int k, int b, int layer, int layer_off, int per_sample, int e_nchw) {
    if (a.drop_keep) return a.drop_keep[((long long)k * a.B + b) * DROP_PER_SAMPLE + layer_off + e_nchw] != 0;
    const unsigned long long e = (unsigned long long)b * per_sample + e_nchw;
    const flb_u4 r = flb_philox_block(flb_epoch_seed(a) ^ 0xD80F0A7ull, a.client_base + a.client_stride * k,
                                      ((unsigned long long)a.tcount[k] << 24) + ((unsigned long long)layer << 20) + (e >> 2));
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
    return flb_u01(rr[e & 3]) >= a.drop_p;
}

// the same decision for 4 consecutive elements e_nchw0 .. e_nchw0 + 3 (e_nchw0 % 4 == 0): ONE Philox block instead of four
__device__ __forceinline__ void drop_keep4(const flb_train_args& a, int k, int b, int layer, int layer_off, int per_sample, int e_nchw0,
                                           bool (&keep)[4]) {
    if (a.drop_keep) {
        const uchar4 m = *reinterpret_cast<const uchar4*>(a.drop_keep + ((long long)k * a.B + b) * DROP_PER_SAMPLE + layer_off + e_nchw0);
        keep[0] = m.x != 0; keep[1] = m.y != 0; keep[2] = m.z != 0; keep[3] = m.w != 0;
        return;
    }
    const unsigned long long e = (unsigned long long)b * per_sample + e_nchw0;
    const flb_u4 r = flb_philox_block(flb_epoch_seed(a) ^ 0xD80F0A7ull, a.client_base + a.client_stride * k,
                                      ((unsigned long long)a.tcount[k] << 24) + ((unsigned long long)layer << 20) + (e >> 2));
    keep[0] = flb_u01(r.x) >= a.drop_p; keep[1] = flb_u01(r.y) >= a.drop_p;
    keep[2] = flb_u01(r.z) >= a.drop_p; keep[3] = flb_u01(r.w) >= a.drop_p;
}

// relu(bn(z)) -> 2x2 max-pool (+argmax) -> dropout.  One CTA per (sample, client).  FLAT: the output is fc1's
// NCHW-flattened input [C * Ho * Wo]; otherwise the next resolution's padded NHWC grid.  Also updates running stats.
template <int C, bool FLAT>
__global__ void __launch_bounds__(256) bn_relu_pool_drop_kernel(flb_train_args a, ConvGeom g, ConvGeom go, const float* z_all,
                                                                float* out_all, uint8_t* idx_all, const double* acc, int coff,
                                                                int gwoff, int gboff, int drop_layer, int drop_off) {
    const int b = blockIdx.x, k = blockIdx.y;
    const int bsz = flb_bsz(a, k);
    if (b >= bsz) return;
    __shared__ float s_scale[C], s_beta[C];
    const int tid = threadIdx.x;
    const int n_real = bsz * g.H * g.W;
    const float* W = a.W + (long long)k * a.ld;
    if (tid < C) {
        float mean, invstd, vb;
        bn_moments(a, acc, k, coff + tid, n_real, mean, invstd, vb);
        s_scale[tid] = invstd * W[gwoff + tid];
        s_beta[tid] = W[gboff + tid] - mean * s_scale[tid];
        if (b == 0 && !a.eval_mode) {
            float* run = a.bn_running + (long long)k * 2 * BN_CH + coff + tid;
            run[0] = (1.f - BN_MOM) * run[0] + BN_MOM * mean;
            const float unb = n_real > 1 ? vb * ((float)n_real / (float)(n_real - 1)) : vb;
            run[BN_CH] = (1.f - BN_MOM) * run[BN_CH] + BN_MOM * unb;
        }
    }
    __syncthreads();
    const long long kb = (long long)k * a.B + b;
    const float* z = z_all + kb * g.PP() * C;
    const int Ho = g.H / 2, Wo = g.W / 2, npool = Ho * Wo;
    float* out = out_all + kb * (FLAT ? C * npool : go.PP() * C);
    uint8_t* idx = idx_all + kb * npool * C;
    const float keep_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
    // thread = (channel, group): 4 consecutive pooled positions of one channel per iteration (same pooled row: Wo % 4 == 0) --
    // their dropout decisions are one Philox block, the window rows are 2 x 8 loads that coalesce over the channel lanes,
    // and the indices come from shifts
    constexpr int GROUPS = 256 / C;
    const int c = tid % C, grp = tid / C, wshift = 31 - __clz(Wo);
    const float sc = s_scale[c], sb = s_beta[c];
    for (int qd = grp; qd < npool / 4; qd += GROUPS) {
        const int pp0 = qd * 4, ph = pp0 >> wshift, pw0 = pp0 & (Wo - 1);
        float zt[2][8];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) zt[i][j] = z[((2 * ph + i) * g.Wp + 2 * pw0 + j) * C + c];
        bool keep[4] = {true, true, true, true};
        if (a.drop_p > 0.f) drop_keep4(a, k, b, drop_layer, drop_off, npool * C, c * npool + pp0, keep);
        float vo[4];
        uint8_t bo[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float best = -INFINITY;
            int bi = 0;
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float v = __fadd_rn(__fmul_rn(zt[i][2 * t + j], sc), sb);
                    if (v > best) { best = v; bi = i * 2 + j; }
                }
            float v = fmaxf(best, 0.f);
            if (a.drop_p > 0.f) {
                if (keep[t]) v *= keep_scale;
                else { v = 0.f; bi |= 4; }
            }
            vo[t] = v; bo[t] = (uint8_t)bi;
        }
        if (FLAT) {
            *reinterpret_cast<float4*>(out + c * npool + pp0) = make_float4(vo[0], vo[1], vo[2], vo[3]);
            *reinterpret_cast<uchar4*>(idx + c * npool + pp0) = make_uchar4(bo[0], bo[1], bo[2], bo[3]);
        } else {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                out[(ph * go.Wp + pw0 + t) * C + c] = vo[t];
                idx[(pp0 + t) * C + c] = bo[t];
            }
        }
    }
}

// dz = gamma * invstd * (g - dbeta/N - xhat * dgamma/N) in place over dy (zeros on pads); g = dy masked by y > 0 when
// y_all is given.  CTA (0, k) writes dgamma / dbeta into G.
template <int C>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(flb_train_args a, ConvGeom g, const float* z_all, int relu_mask,
                                                           float* dy_all, const double* acc, int coff, int gwoff, int gboff,
                                                           int conv_boff) {
    const int k = blockIdx.y;
    const int bsz = flb_bsz(a, k);
    if (bsz == 0) return;
    __shared__ float s_mean[C], s_invstd[C], s_c0[C], s_c1[C], s_c2[C], s_alpha[C], s_beta[C];
    const int tid = threadIdx.x;
    const int n_real = bsz * g.H * g.W;
    if (tid < C) {
        float mean, invstd, vb;
        bn_moments(a, acc, k, coff + tid, n_real, mean, invstd, vb);
        const double* A = acc + (long long)k * 4 * BN_CH + 2 * BN_CH + coff + tid;
        const float dgamma = (float)A[0], dbeta = (float)A[BN_CH];
        const float gamma = a.W[(long long)k * a.ld + gwoff + tid];
        s_mean[tid] = mean; s_invstd[tid] = invstd;
        s_alpha[tid] = invstd * gamma;                                  // forward: y = relu(z * alpha + beta)
        s_beta[tid] = a.W[(long long)k * a.ld + gboff + tid] - mean * s_alpha[tid];
        s_c0[tid] = gamma * invstd;
        s_c1[tid] = dbeta / (float)n_real;
        s_c2[tid] = dgamma / (float)n_real;
        if (blockIdx.x == 0) {
            float* G = a.G + (long long)k * a.ld;
            G[gwoff + tid] = dgamma;
            G[gboff + tid] = dbeta;
        }
    }
    __syncthreads();
    const int PP = g.PP();
    const long long base = (long long)k * a.B * PP * C;
    constexpr int C4 = C / 4, RL = 256 / C4;
    float4* dy4 = reinterpret_cast<float4*>(dy_all + base);
    const float4* z4 = reinterpret_cast<const float4*>(z_all + base);
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};        // thread = (channel quad, pixel lane)
    const int cq = tid % C4, rl = tid / C4, c = cq * 4;
    const int hw_shift = 31 - __clz(g.H * g.W), w_shift = 31 - __clz(g.W);
    float k_al[4], k_be[4], k_c0[4], k_c1[4], k_mi[4], k_m2[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        k_al[q] = s_alpha[c + q]; k_be[q] = s_beta[c + q]; k_c0[q] = s_c0[c + q]; k_c1[q] = s_c1[c + q];
        k_mi[q] = s_mean[c + q]; k_m2[q] = s_invstd[c + q] * s_c2[c + q];
    }
    for (int px = blockIdx.x * RL + rl; px < n_real; px += gridDim.x * RL) {
        const long long e = (long long)real_row(g, px, hw_shift, w_shift) * C4 + cq;
        const float4 gv4 = dy4[e], zv = z4[e];
        float gv[4] = {gv4.x, gv4.y, gv4.z, gv4.w}, o[4];
        const float zz[4] = {zv.x, zv.y, zv.z, zv.w};
        if (relu_mask) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (!(__fadd_rn(__fmul_rn(zz[q], k_al[q]), k_be[q]) > 0.f)) gv[q] = 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            o[q] = k_c0[q] * (gv[q] - k_c1[q] - (zz[q] - k_mi[q]) * k_m2[q]);
            bsum[q] += o[q];
        }
        dy4[e] = make_float4(o[0], o[1], o[2], o[3]);
    }
    // the producer (a dgrad GEMM over the whole padded grid) leaves values on the pads: zero them for the next GEMMs
    const int n_pad = bsz * (PP - g.H * g.W);
    for (int j = blockIdx.x * RL + rl; j < n_pad; j += gridDim.x * RL)
        dy4[(long long)pad_row(g, j) * C4 + cq] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (conv_boff >= 0) {                     // conv bias gradient = column sums of dz (tensor-core path; the fp32 wgrad
        __shared__ float red[4][256];         // GEMM carries it as an extra column).  Exactly zero in exact arithmetic.
#pragma unroll
        for (int q = 0; q < 4; ++q) red[q][tid] = bsum[q];
        __syncthreads();
        if (tid < C) {
            const int q = tid & 3, cq2 = tid >> 2;
            float t = 0.f;
            for (int i = 0; i < 256 / C4; ++i) t += red[q][i * C4 + cq2];
            atomicAdd(&a.G[(long long)k * a.ld + conv_boff + tid], t);
        }
    }
}

// dp_mode 1 (per-sample DP-SGD; the batch statistics are constants of the per-sample backward pass):
// where a BatchNorm backward reduction runs one CTA per sample, its column sums ARE the sample's (dgamma, dbeta).  They
// are stored in bnps[client, sample] and their squares -- plus the square of the sample's conv-bias gradient
// gamma * invstd * dbeta when the layer's conv bias is not covered elsewhere -- join norm2[client, sample].
struct PsOut { float* bnps; float* norm2; int gwoff; int conv_bias; };       // bnps == nullptr: batched mode (atomics into acc)
__device__ __forceinline__ void ps_store(const flb_train_args& a, const PsOut& ps, long long kb, int k, int coff, int ch, float t0, float t1,
                                         float invstd) {
    float* row = ps.bnps + kb * 2 * BN_CH + coff + ch;
    row[0] = t0;
    row[BN_CH] = t1;
    float sq = fmaf(t0, t0, t1 * t1);
    if (ps.conv_bias) {
        const float db = a.W[(long long)k * a.ld + ps.gwoff + ch] * invstd * t1;
        sq = fmaf(db, db, sq);
    }
    sq = flb_warp_sum(sq);                                   // callers: whole warps (C is a multiple of 32)
    if ((threadIdx.x & 31) == 0 && sq != 0.f) atomicAdd(&ps.norm2[kb], sq);
}

// ---- BatchNorm backward of the POOLED layers (2, 4, 6) straight from the pooled-side arrays -----------------------------
// Upstream of a max-pool the gradient is one value per 2x2 window, so the dense dy tensor is never materialised: the
// reduction reads dpool / pooled / argmax (a quarter of the grid) and gathers z at the argmax positions; the apply kernel
// forms dz = gamma * invstd * (g - dbeta/N - xhat * dgamma/N) for every grid position directly.
template <int C, bool FLAT>
__global__ void __launch_bounds__(256) bn_pool_bwd_reduce_kernel(flb_train_args a, ConvGeom g, ConvGeom go, const float* dpool_all,
                                                                 const float* pooled_all, const uint8_t* idx_all,
                                                                 const float* z_all, double* acc, int coff, PsOut ps) {
    const int b = blockIdx.x, k = blockIdx.y;
    const int bsz = flb_bsz(a, k);
    if (b >= bsz) return;
    const int tid = threadIdx.x, c = tid % C;
    const long long kb = (long long)k * a.B + b;
    const int Wo = g.W / 2, npool = (g.H / 2) * Wo;
    const float* dpool = dpool_all + kb * (FLAT ? C * npool : go.PP() * C);
    const float* pooled = pooled_all + kb * (FLAT ? C * npool : go.PP() * C);
    const uint8_t* idx = idx_all + kb * npool * C;
    const float* z = z_all + kb * g.PP() * C;
    const float keep_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
    float mean, invstd, vb;
    bn_moments(a, acc, k, coff + c, bsz * g.H * g.W, mean, invstd, vb);
    float s0 = 0.f, s1 = 0.f;
    for (int e = tid; e < npool * C; e += 256) {              // e % C == c for every e of this thread (256 % C == 0)
        const int pp = e / C, ph = pp / Wo, pw = pp - ph * Wo;
        const int src = FLAT ? c * npool + pp : (ph * go.Wp + pw) * C + c;
        if (!(pooled[src] > 0.f)) continue;                   // dropped, or ReLU inactive: no gradient through this window
        const int code = idx[FLAT ? c * npool + pp : e] & 3;
        const float gv = dpool[src] * keep_scale;
        const float zv = z[((2 * ph + (code >> 1)) * g.Wp + 2 * pw + (code & 1)) * C + c];
        s0 = fmaf(gv, (zv - mean) * invstd, s0);
        s1 += gv;
    }
    __shared__ float red[2][256];
    red[0][tid] = s0; red[1][tid] = s1;
    __syncthreads();
    if (tid < C) {
        for (int i = 1; i < 256 / C; ++i) { s0 += red[0][tid + i * C]; s1 += red[1][tid + i * C]; }
        if (ps.bnps) { ps_store(a, ps, kb, k, coff, tid, s0, s1, invstd); return; }
        double* A = acc + (long long)k * 4 * BN_CH + 2 * BN_CH + coff + tid;
        atomicAdd(A, (double)s0);
        atomicAdd(A + BN_CH, (double)s1);
    }
}

// The same reduction for the NHWC pooled layout (layers 2, 4), vectorised: thread = (channel quad, pooled-position lane);
// 16 B loads of the pooled-side arrays, the four argmax codes as one 32-bit word, and all four window positions of z as
// 16 B loads (a warp fetches those sectors anyway, whichever position each channel picks); shifts instead of divisions.
// (Measured and dropped: the same scheme for the NCHW-flattened layer 6, and CTAs that walk several samples -- both slower.)
template <int C>
__global__ void __launch_bounds__(256) bn_pool_bwd_reduce_vec_kernel(flb_train_args a, ConvGeom g, ConvGeom go, const float* dpool_all,
                                                                     const float* pooled_all, const uint8_t* idx_all,
                                                                     const float* z_all, double* acc, int coff, PsOut ps) {
    const int b = blockIdx.x, k = blockIdx.y;
    const int bsz = flb_bsz(a, k);
    if (b >= bsz) return;
    constexpr int C4 = C / 4, PL = 256 / C4;
    const int tid = threadIdx.x, cq = tid % C4, pl = tid / C4;
    __shared__ float s_mean[C], s_invstd[C];
    __shared__ float red[8][256];
    if (tid < C) {
        float mean, invstd, vb;
        bn_moments(a, acc, k, coff + tid, bsz * g.H * g.W, mean, invstd, vb);
        s_mean[tid] = mean; s_invstd[tid] = invstd;
    }
    __syncthreads();
    const long long kb = (long long)k * a.B + b;
    const int Wo = g.W / 2, npool = (g.H / 2) * Wo, wshift = 31 - __clz(Wo);
    const float4* dpool4 = reinterpret_cast<const float4*>(dpool_all + kb * go.PP() * C);
    const float4* pooled4 = reinterpret_cast<const float4*>(pooled_all + kb * go.PP() * C);
    const uint32_t* idx32 = reinterpret_cast<const uint32_t*>(idx_all + kb * npool * C);
    const float4* z4 = reinterpret_cast<const float4*>(z_all + kb * g.PP() * C);
    const float keep_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
    float mean[4], invstd[4], s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 4; ++q) { mean[q] = s_mean[cq * 4 + q]; invstd[q] = s_invstd[cq * 4 + q]; }
    for (int pp = pl; pp < npool; pp += PL) {
        const int ph = pp >> wshift, pw = pp & (Wo - 1);
        const int src = (ph * go.Wp + pw) * C4 + cq;
        const float4 p4 = pooled4[src], d4 = dpool4[src];
        const uint32_t codes = idx32[pp * C4 + cq];
        const int r00 = (2 * ph * g.Wp + 2 * pw) * C4 + cq;
        const float4 za = z4[r00], zb = z4[r00 + C4], zc = z4[r00 + g.Wp * C4], zd = z4[r00 + (g.Wp + 1) * C4];
        const float pv[4] = {p4.x, p4.y, p4.z, p4.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
        const float zz[4][4] = {{za.x, za.y, za.z, za.w}, {zb.x, zb.y, zb.z, zb.w}, {zc.x, zc.y, zc.z, zc.w}, {zd.x, zd.y, zd.z, zd.w}};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (!(pv[q] > 0.f)) continue;                     // dropped, or ReLU inactive: no gradient through this window
            const int code = (codes >> (8 * q)) & 3;
            const float zv = code == 0 ? zz[0][q] : (code == 1 ? zz[1][q] : (code == 2 ? zz[2][q] : zz[3][q]));
            const float gv = dv[q] * keep_scale;
            s0[q] = fmaf(gv, (zv - mean[q]) * invstd[q], s0[q]);
            s1[q] += gv;
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) { red[q][tid] = s0[q]; red[4 + q][tid] = s1[q]; }
    __syncthreads();
    if (tid < C) {
        const int q = tid & 3, cq2 = tid >> 2;
        float t0 = 0.f, t1 = 0.f;
        for (int i = 0; i < PL; ++i) { t0 += red[q][i * C4 + cq2]; t1 += red[4 + q][i * C4 + cq2]; }
        if (ps.bnps) { ps_store(a, ps, kb, k, coff, tid, t0, t1, s_invstd[tid]); return; }
        double* A = acc + (long long)k * 4 * BN_CH + 2 * BN_CH + coff + tid;
        atomicAdd(A, (double)t0);
        atomicAdd(A + BN_CH, (double)t1);
    }
}

template <int C, bool FLAT>
__global__ void __launch_bounds__(256) bn_pool_bwd_apply_kernel(flb_train_args a, ConvGeom g, ConvGeom go, const float* dpool_all,
                                                                const float* pooled_all, const uint8_t* idx_all, const float* z_all,
                                                                float* dz_all, const double* acc, int coff, int gwoff, int gboff,
                                                                int conv_boff) {
    const int b = blockIdx.x, k = blockIdx.y;
    const int bsz = flb_bsz(a, k);
    if (b >= bsz) return;
    __shared__ float s_mean[C], s_invstd[C], s_c0[C], s_c1[C], s_c2[C];
    const int tid = threadIdx.x;
    const int n_real = bsz * g.H * g.W;
    if (tid < C) {
        float mean, invstd, vb;
        bn_moments(a, acc, k, coff + tid, n_real, mean, invstd, vb);
        const double* A = acc + (long long)k * 4 * BN_CH + 2 * BN_CH + coff + tid;
        const bool ps = a.dp_mode == 1;      // per-sample mode: statistics are constants, dz = gamma * invstd * g; G is written later
        const float dgamma = ps ? 0.f : (float)A[0], dbeta = ps ? 0.f : (float)A[BN_CH];
        s_mean[tid] = mean; s_invstd[tid] = invstd;
        s_c0[tid] = a.W[(long long)k * a.ld + gwoff + tid] * invstd;
        s_c1[tid] = dbeta / (float)n_real;
        s_c2[tid] = dgamma / (float)n_real;
        if (b == 0 && !ps) {
            float* G = a.G + (long long)k * a.ld;
            G[gwoff + tid] = dgamma;
            G[gboff + tid] = dbeta;
        }
    }
    __syncthreads();
    const long long kb = (long long)k * a.B + b;
    const int Wo = g.W / 2, npool = (g.H / 2) * Wo;
    const float* dpool = dpool_all + kb * (FLAT ? C * npool : go.PP() * C);
    const float* pooled = pooled_all + kb * (FLAT ? C * npool : go.PP() * C);
    const uint8_t* idx = idx_all + kb * npool * C;
    const float4* z4 = reinterpret_cast<const float4*>(z_all + kb * g.PP() * C);
    float4* dz4 = reinterpret_cast<float4*>(dz_all + kb * g.PP() * C);
    const float keep_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
    constexpr int C4 = C / 4, LANES = 256 / C4;
    const int c = (tid % C4) * 4, pl = tid / C4;           // fixed channel quad per thread, pixel lane
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
    // REAL pixels only (H, W are powers of two: no divisions): this kernel is the only writer of its dz buffer, whose pad
    // positions therefore keep the zeros of the workspace initialisation
    const int wshift = 31 - __clz(g.W);
    for (int p = pl; p < g.H * g.W; p += LANES) {
        const int h = p >> wshift, w = p & (g.W - 1), e = (h * g.Wp + w) * C4 + (c >> 2);
        const int pp = (h >> 1) * Wo + (w >> 1), pos = (h & 1) * 2 + (w & 1);
        const float4 zv = z4[e];
        const float zz[4] = {zv.x, zv.y, zv.z, zv.w};
        float gq[4];
        if (FLAT) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int src = (c + q) * npool + pp;
                gq[q] = (idx[src] == pos && pooled[src] > 0.f) ? dpool[src] * keep_scale : 0.f;     // bit 2 of the code = dropped
            }
        } else {
            const int src = ((h >> 1) * go.Wp + (w >> 1)) * C + c;
            const uchar4 code = *reinterpret_cast<const uchar4*>(idx + pp * C + c);
            const float4 pv = *reinterpret_cast<const float4*>(pooled + src), dv = *reinterpret_cast<const float4*>(dpool + src);
            gq[0] = (code.x == pos && pv.x > 0.f) ? dv.x * keep_scale : 0.f;
            gq[1] = (code.y == pos && pv.y > 0.f) ? dv.y * keep_scale : 0.f;
            gq[2] = (code.z == pos && pv.z > 0.f) ? dv.z * keep_scale : 0.f;
            gq[3] = (code.w == pos && pv.w > 0.f) ? dv.w * keep_scale : 0.f;
        }
        float o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float xhat = (zz[q] - s_mean[c + q]) * s_invstd[c + q];
            o[q] = s_c0[c + q] * (gq[q] - s_c1[c + q] - xhat * s_c2[c + q]);
            bsum[q] += o[q];
        }
        dz4[e] = make_float4(o[0], o[1], o[2], o[3]);
    }
    if (conv_boff >= 0) {
        __shared__ float red[4][256];
#pragma unroll
        for (int q = 0; q < 4; ++q) red[q][tid] = bsum[q];
        __syncthreads();
        if (tid < C) {
            const int q = tid & 3, cq2 = tid >> 2;
            float t = 0.f;
            for (int i = 0; i < 256 / C4; ++i) t += red[q][i * C4 + cq2];
            atomicAdd(&a.G[(long long)k * a.ld + conv_boff + tid], t);
        }
    }
}

// ---- classifier ------------------------------------------------------------------------------------------------------
// h1 = dropout(relu(hpre1 + b)); the multiplier (0 where ReLU is inactive or the unit is dropped, else 1/(1-p)) is kept
// in m1 for the backward pass
__global__ void __launch_bounds__(256) fc_bias_relu_drop_kernel(flb_train_args a, const float* pre_all, float* h_all, float* m_all,
                                                                int n, int boff, int drop_layer, int drop_off) {
    const int b = blockIdx.x, k = blockIdx.y;
    if (b >= flb_bsz(a, k)) return;
    const long long kb = (long long)k * a.B + b;
    const float* W = a.W + (long long)k * a.ld;
    const float keep_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
    for (int j = threadIdx.x; j < n; j += 256) {
        const float pre = pre_all[kb * n + j] + W[boff + j];
        float mult = pre > 0.f ? 1.f : 0.f;
        if (a.drop_p > 0.f) mult = drop_keep(a, k, b, drop_layer, drop_off, n, j) ? mult * keep_scale : 0.f;
        h_all[kb * n + j] = pre * mult;
        m_all[kb * n + j] = mult;
    }
}

// last two layers: h = dropout(relu(hpre2 + b2)), logits = fc3(h), softmax cross-entropy, dlogits, d(hpre2).
// One CTA per client (same structure as SimpleCNN's head kernel, 256 inputs).
__global__ void __launch_bounds__(256) head_fwd_bwd_kernel(flb_train_args a, CifarWs ws) {
    constexpr int IN = 256;
    const int k = blockIdx.x;
    const int bsz = flb_bsz(a, k);
    if (bsz == 0) return;
    extern __shared__ float smem[];
    float (*sh)[IN + 1] = reinterpret_cast<float (*)[IN + 1]>(smem);                          // [32][257]
    float (*sw)[IN + 1] = reinterpret_cast<float (*)[IN + 1]>(smem + 32 * (IN + 1));          // [10][257]
    float (*slog)[10] = reinterpret_cast<float (*)[10]>(smem + 42 * (IN + 1));                // [32][10]
    float (*sdl)[10] = reinterpret_cast<float (*)[10]>(smem + 42 * (IN + 1) + 320);           // [32][10]
    float* red = smem + 42 * (IN + 1) + 640;
    const int tid = threadIdx.x;
    const float* W = a.W + (long long)k * a.ld;
    const long long kb = (long long)k * a.B;
    const float keep_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
    const int step = *a.step_ctr;
    for (int e = tid; e < bsz * IN; e += 256) {
        const int b = e / IN, j = e % IN;
        const float pre = ws.hpre2[kb * IN + e] + W[kNet.f2b + j];
        float mult = pre > 0.f ? 1.f : 0.f;
        if (a.drop_p > 0.f) mult = drop_keep(a, k, b, 4, 8192 + 4096 + 2048 + 512, IN, j) ? mult * keep_scale : 0.f;
        const float hv = pre * mult;
        sh[b][j] = hv;
        ws.h[kb * IN + e] = hv;
        ws.dh2[kb * IN + e] = mult;
    }
    for (int e = tid; e < 10 * IN; e += 256) sw[e / IN][e % IN] = W[kNet.f3w + e];
    if (tid < 2) red[tid] = 0.f;
    __syncthreads();
    for (int e = tid; e < bsz * 10; e += 256) {
        const int b = e / 10, j = e % 10;
        float acc = W[kNet.f3b + j];
#pragma unroll 8
        for (int i = 0; i < IN; ++i) acc = fmaf(sh[b][i], sw[j][i], acc);
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
        const float gs = a.dp_mode == 1 ? 1.f : 1.f / (float)bsz;           // mean reduction (training.py:90); per-sample mode: 1/B in the optimizer
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
    for (int e = tid; e < bsz * IN; e += 256) {
        const int b = e / IN, j = e % IN;
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 10; ++c) acc = fmaf(sdl[b][c], sw[c][j], acc);
        ws.dh2[kb * IN + e] = acc * ws.dh2[kb * IN + e];
    }
    for (int e = bsz * IN + tid; e < a.B * IN; e += 256) ws.dh2[kb * IN + e] = 0.f;      // zero rows feed the TC wgrad
}
constexpr size_t kHeadSmem = (42 * 257 + 640 + 2) * sizeof(float);

// fc3 weight / bias gradients and fc2's bias gradient: one CTA per client
__global__ void __launch_bounds__(256) head_wgrad_kernel(flb_train_args a, CifarWs ws) {
    constexpr int IN = 256;
    const int k = blockIdx.x;
    const int bsz = flb_bsz(a, k);
    if (bsz == 0) return;
    __shared__ float sdl[32][10];
    const int tid = threadIdx.x;
    const long long kb = (long long)k * a.B;
    float* G = a.G + (long long)k * a.ld;
    for (int e = tid; e < bsz * 10; e += 256) sdl[e / 10][e % 10] = ws.dlog[kb * 10 + e];
    __syncthreads();
    for (int e = tid; e < 10 * IN; e += 256) {
        const int j = e / IN, i = e % IN;
        float acc = 0.f;
        for (int b = 0; b < bsz; ++b) acc = fmaf(sdl[b][j], ws.h[(kb + b) * IN + i], acc);
        G[kNet.f3w + e] = acc;
    }
    if (tid < 10) {
        float acc = 0.f;
        for (int b = 0; b < bsz; ++b) acc += sdl[b][tid];
        G[kNet.f3b + tid] = acc;
    }
    {
        float acc = 0.f;
        for (int b = 0; b < bsz; ++b) acc += ws.dh2[(kb + b) * IN + tid];
        G[kNet.f2b + tid] = acc;
    }
}

// d(hpre1) = dh1 * m1 in place (zero for the rows past the batch), and fc1's bias gradient.  One CTA per client.
__global__ void __launch_bounds__(512) fc1_mask_bias_kernel(flb_train_args a, CifarWs ws) {
    const int k = blockIdx.x;
    const int bsz = flb_bsz(a, k);
    if (bsz == 0) return;
    const int j = threadIdx.x;
    const long long kb = (long long)k * a.B;
    float acc = 0.f;
    for (int b = 0; b < a.B; ++b) {
        const long long e = (kb + b) * 512 + j;
        const float v = b < bsz ? ws.dh1[e] * ws.m1[e] : 0.f;
        ws.dh1[e] = v;
        acc += v;
    }
    a.G[(long long)k * a.ld + kNet.f1b + j] = acc;
}

// ---- per-sample DP-SGD (dp_mode 1; north-star kernel 2 -- no reference code) -------------------------
// Per-sample gradient g_i = d loss_i / d theta with every BatchNorm layer's batch statistics held constant (the forward pass
// is the reference's, batch statistics included).  The backward chain is then independent per sample:
// dz = gamma * invstd * relu'(.) * dy, dgamma_i = sum_px g * xhat, dbeta_i = sum_px g, conv-bias_i = gamma * invstd * dbeta_i.
// Order of a step: activation gradients of all layers -> per-sample squared norms (ghost norms for the linears, per-sample
// conv tiles squared on chip, BatchNorm / bias sums) -> clip coefficients (privacy.py:127-138) -> every layer's upstream
// gradient rows scaled by their sample's coefficient in place -> the ordinary batched weight-gradient kernels.

// BatchNorm + ReLU backward of the layers fed by a dgrad (1, 3, 5), one CTA per (sample, client): dz in place over dy (zeros
// on the pads), the sample's (dgamma, dbeta) -> bnps, squares -> norm2.
template <int C>
__global__ void __launch_bounds__(256) bn_ps_bwd_kernel(flb_train_args a, ConvGeom g, const float* z_all, float* dy_all, const double* acc,
                                                        int coff, int gwoff, int gboff, PsOut ps) {
    const int b = blockIdx.x, k = blockIdx.y;
    const int bsz = flb_bsz(a, k);
    if (b >= bsz) return;
    __shared__ float s_mean[C], s_invstd[C], s_alpha[C], s_beta[C];
    __shared__ float red[8][256];
    const int tid = threadIdx.x;
    if (tid < C) {
        float mean, invstd, vb;
        bn_moments(a, acc, k, coff + tid, bsz * g.H * g.W, mean, invstd, vb);
        s_mean[tid] = mean; s_invstd[tid] = invstd;
        s_alpha[tid] = invstd * a.W[(long long)k * a.ld + gwoff + tid];             // forward: y = relu(z * alpha + beta); c0 = alpha
        s_beta[tid] = a.W[(long long)k * a.ld + gboff + tid] - mean * s_alpha[tid];
    }
    __syncthreads();
    constexpr int C4 = C / 4, RL = 256 / C4;
    const int cq = tid % C4, rl = tid / C4, c = cq * 4;
    const long long kb = (long long)k * a.B + b;
    float4* dy4 = reinterpret_cast<float4*>(dy_all + kb * g.PP() * C);
    const float4* z4 = reinterpret_cast<const float4*>(z_all + kb * g.PP() * C);
    const int w_shift = 31 - __clz(g.W);
    float k_al[4], k_be[4], k_mi[4], k_is[4], s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 4; ++q) { k_al[q] = s_alpha[c + q]; k_be[q] = s_beta[c + q]; k_mi[q] = s_mean[c + q]; k_is[q] = s_invstd[c + q]; }
    for (int p = rl; p < g.H * g.W; p += RL) {
        const int e = ((p >> w_shift) * g.Wp + (p & (g.W - 1))) * C4 + cq;
        const float4 gv4 = dy4[e], zv = z4[e];
        float gv[4] = {gv4.x, gv4.y, gv4.z, gv4.w}, o[4];
        const float zz[4] = {zv.x, zv.y, zv.z, zv.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (!(__fadd_rn(__fmul_rn(zz[q], k_al[q]), k_be[q]) > 0.f)) gv[q] = 0.f;
            s0[q] = fmaf(gv[q], (zz[q] - k_mi[q]) * k_is[q], s0[q]);
            s1[q] += gv[q];
            o[q] = k_al[q] * gv[q];
        }
        dy4[e] = make_float4(o[0], o[1], o[2], o[3]);
    }
    const int n_pad = g.PP() - g.H * g.W;          // the dgrad GEMM left values on the pads
    for (int j = rl; j < n_pad; j += RL) dy4[(long long)pad_row(g, j) * C4 + cq] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 4; ++q) { red[q][tid] = s0[q]; red[4 + q][tid] = s1[q]; }
    __syncthreads();
    if (tid < C) {
        const int q = tid & 3, cq2 = tid >> 2;
        float t0 = 0.f, t1 = 0.f;
        for (int i = 0; i < RL; ++i) { t0 += red[q][i * C4 + cq2]; t1 += red[4 + q][i * C4 + cq2]; }
        ps_store(a, ps, kb, k, coff, tid, t0, t1, s_invstd[tid]);
    }
}

// ghost norms of the three linear layers: ||dW_i||^2 = ||dout_i||^2 * ||act_i||^2, ||db_i||^2 = ||dout_i||^2
__global__ void __launch_bounds__(256) linear_ghost_norm_kernel(flb_train_args a, CifarWs ws) {
    const int b = blockIdx.x, k = blockIdx.y;
    if (b >= flb_bsz(a, k)) return;
    const long long kb = (long long)k * a.B + b;
    const int tid = threadIdx.x;
    float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};          // |a|^2, |dh1|^2, |h1|^2, |dh2|^2, |h|^2, |dlog|^2
    for (int e = tid; e < 2048; e += 256) { const float v = ws.a[kb * 2048 + e]; s[0] = fmaf(v, v, s[0]); }
    for (int e = tid; e < 512; e += 256) {
        const float v = ws.dh1[kb * 512 + e], u = ws.h1[kb * 512 + e];
        s[1] = fmaf(v, v, s[1]); s[2] = fmaf(u, u, s[2]);
    }
    { const float v = ws.dh2[kb * 256 + tid], u = ws.h[kb * 256 + tid]; s[3] = v * v; s[4] = u * u; }
    if (tid < 10) { const float v = ws.dlog[kb * 10 + tid]; s[5] = v * v; }
    __shared__ float red[6][8];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const float t = flb_warp_sum(s[i]);
        if ((tid & 31) == 0) red[i][tid >> 5] = t;
    }
    __syncthreads();
    if (tid == 0) {
        float t[6];
        for (int i = 0; i < 6; ++i) { t[i] = 0.f; for (int w = 0; w < 8; ++w) t[i] += red[i][w]; }
        atomicAdd(&ws.norm2[kb], t[1] * (t[0] + 1.f) + t[3] * (t[2] + 1.f) + t[5] * (t[4] + 1.f));
    }
}

__global__ void clip_coef_kernel(flb_train_args a, CifarWs ws) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.K * a.B) return;
    const float n = sqrtf(ws.norm2[i]);
    ws.coef[i] = n > a.dp_clip ? a.dp_clip / n : 1.f;         // clip rule of privacy.py:127-138, per sample
}

// every layer's upstream-gradient rows times their sample's clip coefficient, in place (TMA cannot scale an operand in
// flight); grid (sample, client, buffer)
struct ScaleTab { float* p[9]; int n[9]; };
__global__ void __launch_bounds__(256) scale_rows_kernel(flb_train_args a, ScaleTab t, const float* coef) {
    const int b = blockIdx.x, k = blockIdx.y;
    if (b >= flb_bsz(a, k)) return;
    const long long kb = (long long)k * a.B + b;
    const float c = coef[kb];
    if (c == 1.f) return;
    const int n = t.n[blockIdx.z];
    float* buf = t.p[blockIdx.z] + kb * n;
    if (n & 3) {
        for (int e = threadIdx.x; e < n; e += 256) buf[e] *= c;
        return;
    }
    float4* p = reinterpret_cast<float4*>(buf);
    for (int e = threadIdx.x; e < n / 4; e += 256) {
        float4 v = p[e];
        v.x *= c; v.y *= c; v.z *= c; v.w *= c;
        p[e] = v;
    }
}

// fc1 bias gradient from the scaled rows (fc1_mask_bias_kernel formed it before the coefficients existed)
__global__ void __launch_bounds__(512) fc1_bias_ps_kernel(flb_train_args a, CifarWs ws) {
    const int k = blockIdx.x;
    const int bsz = flb_bsz(a, k);
    if (bsz == 0) return;
    const int j = threadIdx.x;
    float acc = 0.f;
    for (int b = 0; b < bsz; ++b) acc += ws.dh1[((long long)k * a.B + b) * 512 + j];
    a.G[(long long)k * a.ld + kNet.f1b + j] = acc;
}

// clipped sums of the per-sample BatchNorm gradients, and the conv biases of layers 2..6: gamma * invstd * sum_b coef_b dbeta_b
struct PsTab { int coff[NCONV], bw[NCONV], bb[NCONV], cb[NCONV]; };        // kNet's per-layer offsets as a kernel argument
__global__ void __launch_bounds__(BN_CH) bn_ps_reduce_kernel(flb_train_args a, CifarWs ws, PsTab t) {
    const int k = blockIdx.x;
    const int bsz = flb_bsz(a, k);
    if (bsz == 0) return;
    const int ch = threadIdx.x;
    int layer = 0;
    while (layer + 1 < NCONV && ch >= t.coff[layer + 1]) ++layer;
    const int cl = ch - t.coff[layer], hw = layer < 2 ? 1024 : (layer < 4 ? 256 : 64);
    const long long kb = (long long)k * a.B;
    float t0 = 0.f, t1 = 0.f;
    for (int b = 0; b < bsz; ++b) {
        const float c = ws.coef[kb + b];
        const float* row = ws.bnps + (kb + b) * 2 * BN_CH + ch;
        t0 = fmaf(c, row[0], t0);
        t1 = fmaf(c, row[BN_CH], t1);
    }
    float* G = a.G + (long long)k * a.ld;
    G[t.bw[layer] + cl] = t0;
    G[t.bb[layer] + cl] = t1;
    if (layer > 0) {                                   // conv1's bias comes out of its own weight-gradient kernel
        float mean, invstd, vb;
        bn_moments(a, ws.acc, k, ch, bsz * hw, mean, invstd, vb);
        G[t.cb[layer] + cl] = a.W[(long long)k * a.ld + t.bw[layer] + cl] * invstd * t1;
    }
}

// ---- orchestration ------------------------------------------------------------------------------------------------------
// Which GEMM runs on the tensor cores (precision 1): args.tc_mask selects single kernels for the per-layer parity tests
// (0 = all).  Bit 3*(layer-1) + kind for conv layers 1..5 (conv2..conv6; conv1 is a direct stencil on both paths),
// bits 15..17 fc1, 18..20 fc2; kind 0 = forward, 1 = dgrad, 2 = wgrad.
enum : int { TC_FWD = 0, TC_DGRAD = 1, TC_WGRAD = 2, TC_FC1_BIT = 15, TC_FC2_BIT = 18 };
__host__ inline bool cifar_tc(const flb_train_args& a, int bit) {
    if (a.precision != 1) return false;
    return a.tc_mask == 0 || ((a.tc_mask >> bit) & 1);
}
__host__ inline bool cifar_tc_conv(const flb_train_args& a, int layer, int kind) { return layer > 0 && cifar_tc(a, 3 * (layer - 1) + kind); }
template <int C>
void bn_stats(const flb_train_args& a, const ConvGeom& g, const float* z, double* acc, int coff, cudaStream_t st) {
    static const int resident = flb_resident_ctas(bn_reduce_kernel<C, 0>, 256);
    const int chunks = max(1, min(64, resident / a.K));                  // one resident wave (flb_resident_ctas)
    bn_reduce_kernel<C, 0><<<dim3(chunks, a.K), 256, 0, st>>>(a, g, z, nullptr, acc, coff, -1, -1);
}
// layers whose upstream gradient is a dgrad output (1, 3, 5): dy is dense and still needs the ReLU mask
template <int C>
void bn_bwd(const flb_train_args& a, const ConvGeom& g, const float* z, float* dy, double* acc, int layer, cudaStream_t st) {
    static const int resident_r = flb_resident_ctas(bn_reduce_kernel<C, 1>, 256), resident_a = flb_resident_ctas(bn_bwd_apply_kernel<C>, 256);
    const int chunks = max(1, min(64, resident_r / a.K)), chunks_a = max(1, min(64, resident_a / a.K));
    bn_reduce_kernel<C, 1><<<dim3(chunks, a.K), 256, 0, st>>>(a, g, z, dy, acc, kNet.coff[layer], kNet.bw[layer], kNet.bb[layer]);
    const int conv_boff = cifar_tc_conv(a, layer, TC_WGRAD) ? kNet.cb[layer] : -1;      // else: an extra column of the fp32 wgrad GEMM
    bn_bwd_apply_kernel<C><<<dim3(chunks_a, a.K), 256, 0, st>>>(a, g, z, 1, dy, acc, kNet.coff[layer], kNet.bw[layer], kNet.bb[layer], conv_boff);
}
// pooled layers (2, 4, 6): straight from the pooled-side gradient
template <int C, bool FLAT>
void bn_pool_bwd(const flb_train_args& a, const ConvGeom& g, const ConvGeom& go, const float* dpool, const float* pooled,
                 const uint8_t* idx, const float* z, float* dz, double* acc, int layer, cudaStream_t st, PsOut ps = PsOut{nullptr, nullptr, 0, 0}) {
    const dim3 per_sample(a.B, a.K);
    if constexpr (FLAT) bn_pool_bwd_reduce_kernel<C, FLAT><<<per_sample, 256, 0, st>>>(a, g, go, dpool, pooled, idx, z, acc, kNet.coff[layer], ps);
    else bn_pool_bwd_reduce_vec_kernel<C><<<per_sample, 256, 0, st>>>(a, g, go, dpool, pooled, idx, z, acc, kNet.coff[layer], ps);
    const int conv_boff = (!ps.bnps && cifar_tc_conv(a, layer, TC_WGRAD)) ? kNet.cb[layer] : -1;
    bn_pool_bwd_apply_kernel<C, FLAT><<<per_sample, 256, 0, st>>>(a, g, go, dpool, pooled, idx, z, dz, acc, kNet.coff[layer],
                                                                 kNet.bw[layer], kNet.bb[layer], conv_boff);
}
template <int C>
void bn_apply(const flb_train_args& a, const ConvGeom& g, const float* z, float* y, const double* acc, int layer, cudaStream_t st) {
    static const int resident = flb_resident_ctas(bn_relu_apply_kernel<C>, 256);
    const int chunks = max(1, min(64, resident / a.K));
    bn_relu_apply_kernel<C><<<dim3(chunks, a.K), 256, 0, st>>>(a, g, z, y, acc, kNet.coff[layer], kNet.bw[layer], kNet.bb[layer]);
}

struct Ctx { const flb_train_args& a; const CifarWs& ws; cudaStream_t st; int rc = FLB_OK; };

// returns true when the launch also accumulated the layer's BatchNorm statistics (sum z, sum z^2) into ws.acc
bool conv_fwd(Ctx& c, const ConvGeom& g, const float* xin, float* z, int layer, bool want_stats) {
    const flb_train_args& a = c.a; cudaStream_t st = c.st;
    if (cifar_tc_conv(a, layer, TC_FWD)) {
        const bool fuse = want_stats && tc::conv_fwd_fuses_stats(g.Cin, g.Cout);
        if (int rc = tc::conv_fwd(a, g, xin, z, c.ws.wt + kNet.toff[layer], kNet.ldt, kNet.cb[layer], st,
                                  fuse ? c.ws.acc : nullptr, kNet.coff[layer], BN_CH)) c.rc = rc;
        return fuse;
    }
    ConvFwdProb p{}; p.a = a; p.g = g; p.xin_all = xin; p.z_all = z; p.woff = kNet.cw[layer]; p.boff = kNet.cb[layer];
    simt::launch(p, a.B * g.PP(), g.Cout, 1, a.K, st);
    return false;
}
void conv_dgrad(Ctx& c, const ConvGeom& g, const float* dz, float* dx, int layer) {
    const flb_train_args& a = c.a; cudaStream_t st = c.st;
    if (cifar_tc_conv(a, layer, TC_DGRAD)) { if (int rc = tc::conv_dgrad(a, g, dz, dx, c.ws.wt + kNet.toff[layer], kNet.ldt, st)) c.rc = rc; return; }
    ConvDgradProb p{}; p.a = a; p.g = g; p.dz_all = dz; p.dx_all = dx; p.woff = kNet.cw[layer];
    simt::launch(p, a.B * g.PP(), g.Cin, 1, a.K, st);
}
void conv_wgrad(Ctx& c, const ConvGeom& g, const float* xin, const float* dz, int layer) {
    const flb_train_args& a = c.a; cudaStream_t st = c.st;
    if (cifar_tc_conv(a, layer, TC_WGRAD)) { if (int rc = tc::conv_wgrad(a, g, xin, dz, c.ws.gt + kNet.toff[layer], kNet.ldt, st)) c.rc = rc; return; }
    ConvWgradProb p{}; p.a = a; p.g = g; p.dz_all = dz; p.xin_all = xin; p.coef_all = nullptr;
    p.woff = kNet.cw[layer]; p.boff = kNet.cb[layer]; p.no_bias = a.dp_mode == 1;      // per-sample mode: bn_ps_reduce_kernel owns the conv biases
    const int tiles = ((g.Cout + 63) / 64) * ((9 * g.Cin + 1 + 63) / 64);
    const int splits = max(1, min(64, flb_num_sms() * 2 / (tiles * a.K)));
    simt::launch(p, g.Cout, 9 * g.Cin + 1, splits, a.K, st);
}
void lin_fwd(Ctx& c, const float* act, float* out, int In, int Out, int woff, int splits) {
    const flb_train_args& a = c.a; cudaStream_t st = c.st;
    if (cifar_tc(a, (In == 2048 ? TC_FC1_BIT : TC_FC2_BIT) + TC_FWD)) { if (int rc = tc::fc_fwd(a, act, out, In, Out, woff, splits, st)) c.rc = rc; return; }
    LinFwdProb p{}; p.a = a; p.In = In; p.Out = Out; p.woff = woff; p.act_all = act; p.out_all = out;
    simt::launch(p, a.B, Out, splits, a.K, st);
}
void lin_dgrad(Ctx& c, const float* dout, float* dact, int In, int Out, int woff) {
    const flb_train_args& a = c.a; cudaStream_t st = c.st;
    if (cifar_tc(a, (In == 2048 ? TC_FC1_BIT : TC_FC2_BIT) + TC_DGRAD)) { if (int rc = tc::fc_dgrad(a, dout, dact, In, Out, woff, st)) c.rc = rc; return; }
    LinDgradProb p{}; p.a = a; p.In = In; p.Out = Out; p.woff = woff; p.dout_all = dout; p.dact_all = dact;
    simt::launch(p, a.B, In, 1, a.K, st);
}
// fc1.weight (71 % of the model): optimizer step applied in the wgrad epilogue instead of a gradient round trip through
// HBM -- opt-in (FLB_FUSED_ADAM=1), training step only, after the layer's dgrad has read the old weights
bool fuse_fc1_adam(const flb_train_args& a, bool step) {
    static const bool on = getenv("FLB_FUSED_ADAM") != nullptr;
    return step && on && a.dp_mode == 0 && a.B % 8 == 0 && cifar_tc(a, TC_FC1_BIT + TC_WGRAD);
}
void lin_wgrad(Ctx& c, const float* dout, const float* act, int In, int Out, int woff, bool adam = false) {
    const flb_train_args& a = c.a; cudaStream_t st = c.st;
    if (a.B % 8 == 0 && cifar_tc(a, (In == 2048 ? TC_FC1_BIT : TC_FC2_BIT) + TC_WGRAD)) {        // its K extent is the batch: whole 8-row MMA steps
        if (int rc = tc::fc_wgrad(a, dout, act, In, Out, woff, st, adam)) c.rc = rc;
        return;
    }
    LinWgradProb p{}; p.a = a; p.In = In; p.Out = Out; p.woff = woff; p.boff = 0; p.dout_all = dout; p.act_all = act; p.coef_all = nullptr;
    simt::launch(p, Out, In, 1, a.K, st);
}

constexpr ConvGeom G1 = geom(0, 3, 32), G2 = geom(0, 32, 32), G3 = geom(1, 32, 64), G4 = geom(1, 64, 64),
                   G5 = geom(2, 64, 128), G6 = geom(2, 128, 128);

int forward_impl(const flb_train_args& a, const CifarWs& ws, cudaStream_t st) {
    const int K = a.K, B = a.B;
    const dim3 per_sample(B, K);
    const size_t KB = (size_t)K * B;
    FLB_CUDA(cudaMemsetAsync(ws.acc, 0, sizeof(double) * (size_t)K * 4 * BN_CH, st));
    FLB_CUDA(cudaMemsetAsync(ws.hpre1, 0, sizeof(float) * KB * 512, st));
    FLB_CUDA(cudaMemsetAsync(ws.hpre2, 0, sizeof(float) * KB * 256, st));
    MARK("begin");
    const bool stats = !a.eval_mode;
    Ctx cx{a, ws, st, FLB_OK};
    static_assert(kNet.coff[0] == 0, "conv1_fwd_kernel adds its statistics at channel offset 0");
    conv1_fwd_kernel<<<per_sample, 256, 0, st>>>(a, ws.z1, stats ? ws.acc : nullptr);
    MARK("conv1_fwd");
    bn_apply<32>(a, G1, ws.z1, ws.y1, ws.acc, 0, st);
    MARK("bn1");
    const bool fused1 = conv_fwd(cx, G2, ws.y1, ws.z2, 1, stats);
    MARK("conv2_fwd");
    if (stats && !fused1) bn_stats<32>(a, G2, ws.z2, ws.acc, kNet.coff[1], st);
    bn_relu_pool_drop_kernel<32, false><<<per_sample, 256, 0, st>>>(a, G2, G3, ws.z2, ws.p1, ws.i1, ws.acc, kNet.coff[1], kNet.bw[1], kNet.bb[1], 0, 0);
    MARK("bn2_pool");
    const bool fused2 = conv_fwd(cx, G3, ws.p1, ws.z3, 2, stats);
    MARK("conv3_fwd");
    if (stats && !fused2) bn_stats<64>(a, G3, ws.z3, ws.acc, kNet.coff[2], st);
    bn_apply<64>(a, G3, ws.z3, ws.y3, ws.acc, 2, st);
    MARK("bn3");
    const bool fused3 = conv_fwd(cx, G4, ws.y3, ws.z4, 3, stats);
    MARK("conv4_fwd");
    if (stats && !fused3) bn_stats<64>(a, G4, ws.z4, ws.acc, kNet.coff[3], st);
    bn_relu_pool_drop_kernel<64, false><<<per_sample, 256, 0, st>>>(a, G4, G5, ws.z4, ws.p2, ws.i2, ws.acc, kNet.coff[3], kNet.bw[3], kNet.bb[3], 1, 8192);
    MARK("bn4_pool");
    const bool fused4 = conv_fwd(cx, G5, ws.p2, ws.z5, 4, stats);
    MARK("conv5_fwd");
    if (stats && !fused4) bn_stats<128>(a, G5, ws.z5, ws.acc, kNet.coff[4], st);
    bn_apply<128>(a, G5, ws.z5, ws.y5, ws.acc, 4, st);
    MARK("bn5");
    const bool fused5 = conv_fwd(cx, G6, ws.y5, ws.z6, 5, stats);
    MARK("conv6_fwd");
    if (stats && !fused5) bn_stats<128>(a, G6, ws.z6, ws.acc, kNet.coff[5], st);
    bn_relu_pool_drop_kernel<128, true><<<per_sample, 256, 0, st>>>(a, G6, G6, ws.z6, ws.a, ws.i3, ws.acc, kNet.coff[5], kNet.bw[5], kNet.bb[5], 2, 8192 + 4096);
    MARK("bn6_pool");
    lin_fwd(cx, ws.a, ws.hpre1, 2048, 512, kNet.f1w, 8);
    MARK("fc1_fwd");
    fc_bias_relu_drop_kernel<<<per_sample, 256, 0, st>>>(a, ws.hpre1, ws.h1, ws.m1, 512, kNet.f1b, 3, 8192 + 4096 + 2048);
    lin_fwd(cx, ws.h1, ws.hpre2, 512, 256, kNet.f2w, 4);
    MARK("fc2_fwd");
    static bool configured = false;
    if (!configured) {
        FLB_CUDA(cudaFuncSetAttribute(head_fwd_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHeadSmem));
        configured = true;
    }
    head_fwd_bwd_kernel<<<K, 256, kHeadSmem, st>>>(a, ws);
    MARK("head_fwd_bwd");
    if (cx.rc) return cx.rc;
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

// per-sample squared norm of a conv layer's weight gradient (bias excluded) -> norm2
void conv_wgrad_norm(Ctx& c, const ConvGeom& g, const float* xin, const float* dz, int layer) {
    const flb_train_args& a = c.a; cudaStream_t st = c.st;
    if (cifar_tc_conv(a, layer, TC_WGRAD)) { if (int rc = tc::conv_wgrad_norm(a, g, xin, dz, c.ws.norm2, st)) c.rc = rc; return; }
    ConvWgradNormProb p{}; p.a = a; p.g = g; p.dz_all = dz; p.xin_all = xin; p.norm2_all = c.ws.norm2; p.no_bias = 1;
    simt::launch(p, g.Cout, 9 * g.Cin, 1, a.K * a.B, st);
}

// dp_mode 1: backward pass of the per-sample DP-SGD step (see the kernel section above); G ends up holding
// sum_i clip(g_i), the optimizer kernel adds sigma * z and divides by the batch size
int backward_per_sample(const flb_train_args& a, const CifarWs& ws, cudaStream_t st) {
    const int K = a.K, B = a.B;
    const dim3 per_sample(B, K);
    Ctx cx{a, ws, st, FLB_OK};
    auto ps = [&](int layer, bool conv_bias) { return PsOut{ws.bnps, ws.norm2, kNet.bw[layer], conv_bias ? 1 : 0}; };
    // ---- activation gradients of every layer (each keeps its own buffer until the weight gradients are formed) ----
    lin_dgrad(cx, ws.dh2, ws.dh1, 512, 256, kNet.f2w);
    fc1_mask_bias_kernel<<<K, 512, 0, st>>>(a, ws);
    lin_dgrad(cx, ws.dh1, ws.da, 2048, 512, kNet.f1w);
    MARK("fc_dgrads");
    bn_pool_bwd<128, true>(a, G6, G6, ws.da, ws.a, ws.i3, ws.z6, ws.d8a, ws.acc, 5, st, ps(5, true));
    conv_dgrad(cx, G6, ws.d8a, ws.d8b, 5);
    bn_ps_bwd_kernel<128><<<per_sample, 256, 0, st>>>(a, G5, ws.z5, ws.d8b, ws.acc, kNet.coff[4], kNet.bw[4], kNet.bb[4], ps(4, true));
    conv_dgrad(cx, G5, ws.d8b, ws.d8p, 4);
    MARK("block3_dgrads");
    bn_pool_bwd<64, false>(a, G4, G5, ws.d8p, ws.p2, ws.i2, ws.z4, ws.d16a, ws.acc, 3, st, ps(3, true));
    conv_dgrad(cx, G4, ws.d16a, ws.d16b, 3);
    bn_ps_bwd_kernel<64><<<per_sample, 256, 0, st>>>(a, G3, ws.z3, ws.d16b, ws.acc, kNet.coff[2], kNet.bw[2], kNet.bb[2], ps(2, true));
    conv_dgrad(cx, G3, ws.d16b, ws.d16p, 2);
    MARK("block2_dgrads");
    bn_pool_bwd<32, false>(a, G2, G3, ws.d16p, ws.p1, ws.i1, ws.z2, ws.d32a, ws.acc, 1, st, ps(1, true));
    conv_dgrad(cx, G2, ws.d32a, ws.d32b, 1);
    bn_ps_bwd_kernel<32><<<per_sample, 256, 0, st>>>(a, G1, ws.z1, ws.d32b, ws.acc, kNet.coff[0], kNet.bw[0], kNet.bb[0], ps(0, false));
    MARK("block1_dgrads");
    // ---- per-sample norms -> clip coefficients ----
    linear_ghost_norm_kernel<<<per_sample, 256, 0, st>>>(a, ws);
    conv1_wgrad_kernel<true><<<per_sample, 256, 0, st>>>(a, ws.d32b, ws.norm2);
    MARK("ghost_conv1_norms");
    conv_wgrad_norm(cx, G6, ws.y5, ws.d8a, 5);
    conv_wgrad_norm(cx, G5, ws.p2, ws.d8b, 4);
    conv_wgrad_norm(cx, G4, ws.y3, ws.d16a, 3);
    conv_wgrad_norm(cx, G3, ws.p1, ws.d16b, 2);
    conv_wgrad_norm(cx, G2, ws.y1, ws.d32a, 1);
    MARK("conv_wgrad_norms");
    clip_coef_kernel<<<flb_cdiv(K * B, 256), 256, 0, st>>>(a, ws);
    ScaleTab tab{{ws.dlog, ws.dh2, ws.dh1, ws.d8a, ws.d8b, ws.d16a, ws.d16b, ws.d32a, ws.d32b},
                 {10, 256, 512, PP8 * 128, PP8 * 128, PP16 * 64, PP16 * 64, PP32 * 32, PP32 * 32}};
    scale_rows_kernel<<<dim3(B, K, 9), 256, 0, st>>>(a, tab, ws.coef);
    MARK("clip_scale");
    // ---- the ordinary batched weight-gradient kernels on the scaled rows ----
    head_wgrad_kernel<<<K, 256, 0, st>>>(a, ws);
    lin_wgrad(cx, ws.dh2, ws.h1, 512, 256, kNet.f2w);
    lin_wgrad(cx, ws.dh1, ws.a, 2048, 512, kNet.f1w);
    fc1_bias_ps_kernel<<<K, 512, 0, st>>>(a, ws);
    MARK("fc_wgrads");
    conv_wgrad(cx, G6, ws.y5, ws.d8a, 5);
    conv_wgrad(cx, G5, ws.p2, ws.d8b, 4);
    conv_wgrad(cx, G4, ws.y3, ws.d16a, 3);
    conv_wgrad(cx, G3, ws.p1, ws.d16b, 2);
    conv_wgrad(cx, G2, ws.y1, ws.d32a, 1);
    conv1_wgrad_kernel<false><<<per_sample, 256, 0, st>>>(a, ws.d32b, nullptr);
    MARK("conv_wgrads");
    PsTab pt;
    for (int i = 0; i < NCONV; ++i) { pt.coff[i] = kNet.coff[i]; pt.bw[i] = kNet.bw[i]; pt.bb[i] = kNet.bb[i]; pt.cb[i] = kNet.cb[i]; }
    bn_ps_reduce_kernel<<<K, BN_CH, 0, st>>>(a, ws, pt);
    MARK("bn_ps_reduce");
    if (cx.rc) return cx.rc;
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

int forward_backward_impl(const flb_train_args& a, cudaStream_t st, bool step) {
    CifarWs ws;
    carve(a.ws, a.K, a.B, &ws);
    const int K = a.K, B = a.B;
    const dim3 per_sample(B, K);
    // gradients accumulated with atomics (conv weights / biases) start from zero; everything else is stored
    FLB_CUDA(cudaMemset2DAsync(a.G, a.ld * sizeof(float), 0, kNet.f1w * sizeof(float), K, st));
    if (a.dp_mode == 1) FLB_CUDA(cudaMemsetAsync(ws.norm2, 0, sizeof(float) * (size_t)K * B, st));
    if (int rc = forward_impl(a, ws, st)) return rc;
    Ctx cx{a, ws, st, FLB_OK};
    if (a.precision == 1) FLB_CUDA(cudaMemsetAsync(ws.gt, 0, sizeof(float) * (size_t)K * kNet.ldt, st));
    if (a.dp_mode == 1) return backward_per_sample(a, ws, st);

    head_wgrad_kernel<<<K, 256, 0, st>>>(a, ws);
    lin_wgrad(cx, ws.dh2, ws.h1, 512, 256, kNet.f2w);
    lin_dgrad(cx, ws.dh2, ws.dh1, 512, 256, kNet.f2w);
    fc1_mask_bias_kernel<<<K, 512, 0, st>>>(a, ws);
    MARK("fc23_bwd");
    lin_dgrad(cx, ws.dh1, ws.da, 2048, 512, kNet.f1w);                 // before the wgrad: with the fused optimizer that one overwrites W
    lin_wgrad(cx, ws.dh1, ws.a, 2048, 512, kNet.f1w, fuse_fc1_adam(a, step));
    MARK("fc1_bwd");

    // block 3 (8x8, 128 channels)
    bn_pool_bwd<128, true>(a, G6, G6, ws.da, ws.a, ws.i3, ws.z6, ws.d8a, ws.acc, 5, st);
    MARK("bn6_bwd");
    conv_wgrad(cx, G6, ws.y5, ws.d8a, 5);
    MARK("conv6_wgrad");
    conv_dgrad(cx, G6, ws.d8a, ws.d8b, 5);
    MARK("conv6_dgrad");
    bn_bwd<128>(a, G5, ws.z5, ws.d8b, ws.acc, 4, st);
    MARK("bn5_bwd");
    conv_wgrad(cx, G5, ws.p2, ws.d8b, 4);
    MARK("conv5_wgrad");
    conv_dgrad(cx, G5, ws.d8b, ws.d8p, 4);
    MARK("conv5_dgrad");

    // block 2 (16x16, 64 channels)
    bn_pool_bwd<64, false>(a, G4, G5, ws.d8p, ws.p2, ws.i2, ws.z4, ws.d16a, ws.acc, 3, st);
    MARK("bn4_bwd");
    conv_wgrad(cx, G4, ws.y3, ws.d16a, 3);
    MARK("conv4_wgrad");
    conv_dgrad(cx, G4, ws.d16a, ws.d16b, 3);
    MARK("conv4_dgrad");
    bn_bwd<64>(a, G3, ws.z3, ws.d16b, ws.acc, 2, st);
    MARK("bn3_bwd");
    conv_wgrad(cx, G3, ws.p1, ws.d16b, 2);
    MARK("conv3_wgrad");
    conv_dgrad(cx, G3, ws.d16b, ws.d16p, 2);
    MARK("conv3_dgrad");

    // block 1 (32x32, 32 channels)
    bn_pool_bwd<32, false>(a, G2, G3, ws.d16p, ws.p1, ws.i1, ws.z2, ws.d32a, ws.acc, 1, st);
    MARK("bn2_bwd");
    conv_wgrad(cx, G2, ws.y1, ws.d32a, 1);
    MARK("conv2_wgrad");
    conv_dgrad(cx, G2, ws.d32a, ws.d32b, 1);
    MARK("conv2_dgrad");
    bn_bwd<32>(a, G1, ws.z1, ws.d32b, ws.acc, 0, st);
    MARK("bn1_bwd");
    conv1_wgrad_kernel<false><<<per_sample, 256, 0, st>>>(a, ws.d32b, nullptr);
    MARK("conv1_wgrad");
    if (cx.rc) return cx.rc;
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

}  // namespace

namespace cifar {
int num_params() { return kNet.P; }
long long bn_floats() { return 2 * BN_CH; }
long long ws_bytes(int K, int B) { return (long long)carve(nullptr, K, B, nullptr); }
long long ws_offset(int K, int B, const char* name) {
    CifarWs ws;
    carve((void*)0, K, B, &ws);
#define FIELD(f) if (!strcmp(name, #f)) return (long long)(uintptr_t)ws.f;
    FIELD(z1) FIELD(y1) FIELD(z2) FIELD(p1) FIELD(z3) FIELD(y3) FIELD(z4) FIELD(p2) FIELD(z5) FIELD(y5) FIELD(z6) FIELD(a)
    FIELD(hpre1) FIELD(h1) FIELD(hpre2) FIELD(h) FIELD(logits) FIELD(dlog) FIELD(dh2) FIELD(dh1) FIELD(da) FIELD(acc) FIELD(d32a) FIELD(d32b) FIELD(d16p) FIELD(d16a) FIELD(d16b) FIELD(d8p) FIELD(d8a) FIELD(d8b)
    FIELD(norm2) FIELD(coef) FIELD(bnps)
#undef FIELD
    return -1;
}
int forward(const flb_train_args& a, cudaStream_t st) {
    CifarWs ws;
    carve(a.ws, a.K, a.B, &ws);
    return forward_impl(a, ws, st);
}
int forward_backward(const flb_train_args& a, cudaStream_t st, bool step) { return forward_backward_impl(a, st, step); }
// forward 21 + backward 29 launches (conv1 carries its own BatchNorm statistics); on the tensor-core path three more
// statistic passes ride in the conv epilogues
int step_launches(const flb_train_args& a) { return 21 + (a.dp_mode == 1 ? 37 : 29) - (a.precision == 1 ? 3 : 0); }
void tc_tab(const flb_train_args& a, TcConvTab* t, bool step) {
    if (fuse_fc1_adam(a, step)) { t->skip_lo = kNet.f1w; t->skip_hi = kNet.f1b; }
    if (a.precision != 1) return;
    CifarWs ws;
    carve(a.ws, a.K, a.B, &ws);
    t->n = NCONV - 1;
    for (int i = 1; i < NCONV; ++i) {
        t->woff[i - 1] = kNet.cw[i]; t->cin[i - 1] = kNet.cin[i]; t->cout[i - 1] = kNet.cout[i];
        t->toff[i - 1] = kNet.toff[i]; t->gt_live[i - 1] = cifar_tc_conv(a, i, TC_WGRAD) ? 1 : 0;
    }
    t->ldt = kNet.ldt; t->wt = ws.wt; t->gt = ws.gt;
}
}  // namespace cifar
