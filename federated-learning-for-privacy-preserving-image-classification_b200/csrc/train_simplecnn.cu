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
#include "gemm_probs.cuh"
#include "philox.cuh"
#include <string.h>

namespace tc {
int conv_fwd(const flb_train_args& a, const ConvGeom& g, const float* xin, float* z, const float* wt, long long ldt, int boff, cudaStream_t st,
             double* bn_acc = nullptr, int bn_coff = 0, int bn_stride = 0);
int conv_dgrad(const flb_train_args& a, const ConvGeom& g, const float* dz, float* dx, const float* wt, long long ldt, cudaStream_t st);
int conv_wgrad(const flb_train_args& a, const ConvGeom& g, const float* xin, const float* dz, float* gt, long long ldt, cudaStream_t st);
int conv_fwd_pool_32_64(const flb_train_args& a, const ConvGeom& g, const float* xin, float* pooled, uint8_t* idx, const float* wt,
                        long long ldt, int boff, cudaStream_t st);
int conv_wgrad_norm_32_64(const flb_train_args& a, const ConvGeom& g, const float* xin, const float* dz, float* norm2, cudaStream_t st);
int fc_fwd(const flb_train_args& a, const float* act, float* out, int in, int outf, int woff, int splits, cudaStream_t st);
int fc_dgrad(const flb_train_args& a, const float* dout, float* dact, int in, int outf, int woff, cudaStream_t st);
int fc_wgrad(const flb_train_args& a, const float* dout, const float* act, int in, int outf, int woff, cudaStream_t st, bool adam = false);
int fc1_fused(const flb_train_args& a, const SimpleCnnWs& ws, cudaStream_t st);
}  // namespace tc

namespace {

enum : int { TC_CONV2_FWD = 1, TC_FC1_FWD = 2, TC_FC1_DGRAD = 4, TC_CONV2_DGRAD = 8, TC_FC1_WGRAD = 16, TC_CONV2_WGRAD = 32 };

// which GEMMs run on the tensor cores for these args
int tc_mask_of(const flb_train_args& a) {
    int m = a.precision == 1 ? (a.tc_mask ? a.tc_mask : 63) : 0;
    if (a.B % 8) m &= ~TC_FC1_WGRAD;        // its K extent is the batch: whole 8-row MMA steps only
    return m;
}

// fc1.weight's optimizer step applied in its weight-gradient GEMM epilogue (FcWgradSwapT<.., ADAM>): only inside a real
// training step (not the gradient-only entries), without per-sample clipping (which needs every layer's norm before any update).
// OPT-IN (FLB_FUSED_ADAM=1).  Measured at 10 clients per GPU (profiles/r02_fusion_ab.md): the epilogue streams W, M, V at
// 3.6 TB/s (26.6 us) against 10.2 us for the plain weight-gradient GEMM plus 12.3 us saved in the stand-alone optimizer --
// no less work, and the side lane it runs on cannot overlap the persistent conv2 dgrad (shared memory co-residency), so
// the round is 2 % slower with it.
bool fuse_fc1_adam(const flb_train_args& a, bool step) {
    static const bool on = getenv("FLB_FUSED_ADAM") != nullptr;
    return step && on && a.dp_mode == 0 && (tc_mask_of(a) & TC_FC1_WGRAD);
}

// fc1 forward + classifier head + fc1 dgrad as ONE launch (fc1_fused.cu) whenever both GEMMs run on the tensor cores and a
// backward pass follows (training step or gradient-only entry; evaluation keeps the separate forward kernels) -- as long as
// every client gets its own group of 14 co-resident CTAs.  With more clients than groups (10 on 148 SMs) a group walks its
// clients one after the other, each a ~20 us chain of dependent phases with nothing overlapped: measured at 50 clients the
// fused kernel takes 99 us against 31 + 14 + 25 us for the three throughput-oriented kernels (round 7.50 -> 6.85 ms), at 10
// clients 25.5 us against 11.3 + 11.2 + 9.3 (round 1.985 vs 2.006 ms).
bool fuse_fc1_block(const flb_train_args& a, bool backward) {
    static const bool off = getenv("FLB_NO_FUSED_FC1") != nullptr;
    const int m = tc_mask_of(a);
    return backward && !off && (m & TC_FC1_FWD) && (m & TC_FC1_DGRAD) && a.K <= flb_num_sms() / 14;
}

using Off = SimpleCnnOff;
constexpr int PP2 = 256;     // conv2 runs on a 16x16 padded grid (14x14 real)
constexpr int WP2 = 16;

// ------------------------------------------------------------------------------------------------
// conv1: direct stencil fused with bias + ReLU + 2x2 max-pool; two CTAs per (sample, client), one per half image
// (pooled rows 0-6 / 7-13), so that 2 * B * K CTAs cover the GPU even at 10 clients.
// Lanes = the 32 output channels, warp g = pooled row g of the half: the 4 x 4 input window of a pooling position is the
// same for all lanes (a shared-memory broadcast) and SLIDES along the row -- only its two new columns are loaded per
// position, as four 8-byte loads.  (The first version gave each warp every 8th position and re-loaded all 16 window values
// with 4-byte loads: 208 LDS per thread for 468 FMAs made the kernel MIO-bound, 11 us of SM time for ~3 us of math.)
constexpr int C1_THREADS = 224;              // 7 warps = 7 pooled rows
__global__ void __launch_bounds__(C1_THREADS) conv1_fwd_pool_kernel(flb_train_args a, SimpleCnnWs ws) {
    const int b = blockIdx.x >> 1, half = blockIdx.x & 1, k = blockIdx.y;
    if (b >= flb_bsz(a, k)) return;
    __shared__ __align__(16) float img[16][32];   // input rows 14*half - 1 .. 14*half + 14 with a zero halo; col c holds x[.][c - 1]
    __shared__ float w[32][9];
    __shared__ float bias[32];
    const int tid = threadIdx.x;
    const float* W = a.W + (long long)k * a.ld;
    const long long s = a.sample_off[k] + (long long)(*a.step_ctr) * a.B + b;
    const float* x = a.x + s * 784;
    const int r0 = 14 * half - 1;
    for (int i = tid; i < 16 * 32; i += C1_THREADS) {
        const int r = i >> 5, c = i & 31, gr = r0 + r;
        img[r][c] = (gr >= 0 && gr < 28 && c >= 1 && c <= 28) ? x[gr * 28 + (c - 1)] : 0.f;
    }
    for (int i = tid; i < 288; i += C1_THREADS) w[i / 9][i % 9] = W[Off::c1w + i];
    if (tid < 32) bias[tid] = W[Off::c1b + tid];
    __syncthreads();
    const int c = tid & 31, phl = tid >> 5, ph = half * 7 + phl;
    float wr[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) wr[i] = w[c][i];
    const float bc = bias[c];
    const long long kb = (long long)k * a.B + b;
    float* outp = ws.a1p + kb * (PP2 * 32) + (ph * WP2) * 32 + c;
    uint8_t* idx = ws.idx1 + kb * (196 * 32) + (ph * 14) * 32 + c;
    float p[4][4];                                 // window rows 2*phl .. 2*phl + 3, columns 2*pw .. 2*pw + 3
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 v = *reinterpret_cast<const float2*>(&img[2 * phl + i][0]);
        p[i][2] = v.x; p[i][3] = v.y;
    }
#pragma unroll
    for (int pw = 0; pw < 14; ++pw) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 v = *reinterpret_cast<const float2*>(&img[2 * phl + i][2 * pw + 2]);
            p[i][0] = p[i][2]; p[i][1] = p[i][3];
            p[i][2] = v.x; p[i][3] = v.y;
        }
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
        outp[pw * 32] = fmaxf(best, 0.f);
        idx[pw * 32] = (uint8_t)bi;
    }
}

// conv2 epilogue: ReLU + 2x2 max-pool + argmax from z2 (bias included), NCHW-flattened output.
// One thread per (pooled position, 4 channels): 16 B loads along the channel axis; grid (7 pooled rows * B, K).
__global__ void __launch_bounds__(112) pool2_kernel(flb_train_args a, SimpleCnnWs ws) {
    const int b = blockIdx.x / 7, ph = blockIdx.x % 7, k = blockIdx.y;
    if (b >= flb_bsz(a, k)) return;
    const long long kb = (long long)k * a.B + b;
    const float* z = ws.z2 + kb * (PP2 * 64);
    float* o = ws.a2 + kb * 3136;
    uint8_t* idx = ws.idx2 + kb * 3136;
    const int pw = threadIdx.x >> 4, c = (threadIdx.x & 15) * 4, pp = ph * 7 + pw;
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int bi[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float4 v4 = *reinterpret_cast<const float4*>(z + ((2 * ph + i) * WP2 + 2 * pw + j) * 64 + c);
            const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (v[e] > best[e]) { best[e] = v[e]; bi[e] = i * 2 + j; }
        }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        o[(c + e) * 49 + pp] = fmaxf(best[e], 0.f);
        idx[(c + e) * 49 + pp] = (uint8_t)bi[e];
    }
}

// ------------------------------------------------------------------------------------------------
// classifier head: fc1 bias + ReLU + dropout, fc2, softmax cross-entropy, dlogits, dh.
// grid (K, HEAD_PARTS): each CTA owns a slice of the client's batch; the epoch accumulators take one atomic per CTA.
constexpr int HEAD_PARTS = 8;
__global__ void __launch_bounds__(256) head_fwd_bwd_kernel(flb_train_args a, SimpleCnnWs ws) {
    const int k = blockIdx.x;
    const int bsz = flb_bsz(a, k);
    if (bsz == 0) return;
    const int per = (a.B + gridDim.y - 1) / gridDim.y, b0 = blockIdx.y * per;
    const int nloc = max(0, min(bsz, b0 + per) - b0);          // live samples of this slice
    __shared__ float sh[32][129];
    __shared__ float sw2[10][129];
    __shared__ float sb2[10];
    __shared__ float slog[32][10];
    __shared__ float sdl[32][10];
    __shared__ float red[2];
    const int tid = threadIdx.x;
    const float* W = a.W + (long long)k * a.ld;
    const long long kb = (long long)k * a.B + b0;
    // rows bsz..B-1 of dh feed the tensor-core fc1 wgrad as zeros (its K extent is the whole batch)
    for (int e = nloc * 128 + tid; e < min(per, a.B - b0) * 128; e += 256) ws.dh[kb * 128 + e] = 0.f;
    if (nloc == 0) return;
    const float keep_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
    // This kernel is a chain of four short dependent phases: every global load whose address is known up front is
    // issued here, before the first barrier (labels, fc2 weights / bias), so that only hpre's latency is exposed.
    int label = 0;                                             // of sample tid >> 4 (the softmax phase's mapping)
    if (tid < 16 * nloc) label = a.y[a.sample_off[k] + (long long)(*a.step_ctr) * a.B + b0 + (tid >> 4)];
    for (int e = tid; e < 1280; e += 256) sw2[e >> 7][e & 127] = W[Off::f2w + e];
    if (tid < 10) sb2[tid] = W[Off::f2b + tid];
    if (tid < 2) red[tid] = 0.f;
    // h = dropout(relu(hpre + b1)); the multiplier (0 or 1/(1-p), and 0 where ReLU is inactive) stays in registers for the
    // backward part below.  A thread owns 4 consecutive elements = one Philox block and one 16 B access (the kernel is
    // bound by instruction latency at 8 warps per SM, so instructions per warp are what counts: ncu, 18 cycles each).
    constexpr int QUADS = (32 / HEAD_PARTS) * 128 / 4;         // B <= 32 samples over gridDim.y = HEAD_PARTS slices
    static_assert(32 % HEAD_PARTS == 0 && QUADS <= 256, "one quad per thread");
    float mult_r[4] = {0.f, 0.f, 0.f, 0.f};
    const bool own = tid < nloc * 32;
    if (own) {
        const int e = tid * 4, b = e >> 7, j = e & 127;
        const float4 pre4 = *reinterpret_cast<const float4*>(ws.hpre + kb * 128 + e);
        const float4 b4 = *reinterpret_cast<const float4*>(W + Off::f1b + j);
        *reinterpret_cast<float4*>(ws.hpre + kb * 128 + e) = make_float4(0.f, 0.f, 0.f, 0.f);   // consumed: ready for the next step's split-K atomics
        const float pre[4] = {pre4.x + b4.x, pre4.y + b4.y, pre4.z + b4.z, pre4.w + b4.w};
        bool keep[4] = {true, true, true, true};
        if (a.drop_p > 0.f) {
            if (a.drop_keep) {
#pragma unroll
                for (int i = 0; i < 4; ++i) keep[i] = a.drop_keep[kb * 128 + e + i] != 0;
            } else {
                const int eg = b0 * 128 + e;                   // element index within the client's [B, 128] block
                const flb_u4 r = flb_philox_block(flb_epoch_seed(a) ^ 0xD80F0A7ull, a.client_base + a.client_stride * k,
                                                  ((unsigned long long)a.tcount[k] << 12) + (eg >> 2));
                const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) keep[i] = flb_u01(rr[i]) >= a.drop_p;
            }
        }
        float hv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            mult_r[i] = (pre[i] > 0.f && keep[i]) ? keep_scale : 0.f;
            hv[i] = pre[i] * mult_r[i];
            sh[b][j + i] = hv[i];
        }
        *reinterpret_cast<float4*>(ws.h + kb * 128 + e) = make_float4(hv[0], hv[1], hv[2], hv[3]);
    }
    __syncthreads();
    for (int e = tid >> 5; e < nloc * 10; e += 8) {            // one warp per logit: 4 products per lane + shuffle reduce
        const int b = e / 10, j = e % 10, l = tid & 31;
        float acc = sh[b][l] * sw2[j][l];
        acc = fmaf(sh[b][l + 32], sw2[j][l + 32], acc);
        acc = fmaf(sh[b][l + 64], sw2[j][l + 64], acc);
        acc = fmaf(sh[b][l + 96], sw2[j][l + 96], acc);
        acc = flb_warp_sum(acc) + sb2[j];
        if (l == 0) {
            slog[b][j] = acc;
            ws.logits[kb * 10 + e] = acc;
        }
    }
    __syncthreads();
    // softmax cross-entropy: 16 lanes per sample (10 live), reductions by width-16 shuffles
    if (tid < 16 * (32 / HEAD_PARTS)) {
        const int b = tid >> 4, j = tid & 15;
        const bool live = b < nloc && j < 10;
        const int bb = b < nloc ? b : 0;
        const float v = live ? slog[bb][j] : -INFINITY;
        float mx = v;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o, 16));
        int am = (live && v == mx) ? j : 99;                   // first index of the maximum, like the sequential scan
        float ex = live ? expf(v - mx) : 0.f, se = ex;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            am = min(am, __shfl_xor_sync(0xffffffffu, am, o, 16));
            se += __shfl_xor_sync(0xffffffffu, se, o, 16);
        }
        const float lse = logf(se) + mx;
        const int y = label;
        const float gs = a.dp_mode == 1 ? 1.f : 1.f / (float)bsz;          // mean reduction (training.py:90)
        if (live) {
            const float d = (expf(v - lse) - (j == y ? 1.f : 0.f)) * gs;
            sdl[b][j] = d;
            ws.dlog[kb * 10 + b * 10 + j] = d;
            if (j == y) atomicAdd(&red[0], lse - v);
            if (j == 0) atomicAdd(&red[1], am == y ? 1.f : 0.f);
        }
    }
    __syncthreads();
    if (tid == 0) {
        atomicAdd(&a.loss_sum[k], red[0] / (float)bsz);     // running_loss += loss.item()   (training.py:200)
        atomicAdd(&a.correct[k], (int)(red[1] + 0.5f));     // correct += (pred == y).sum()  (training.py:201-203)
        if (blockIdx.y == 0) {
            a.nbatch[k] += 1;
            a.nseen[k] += bsz;
        }
    }
    if (own) {
        const int e = tid * 4, b = e >> 7, j = e & 127;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 10; ++c) {
            const float d = sdl[b][c];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = fmaf(d, sw2[c][j + i], acc[i]);
        }
        *reinterpret_cast<float4*>(ws.dh + kb * 128 + e) = make_float4(acc[0] * mult_r[0], acc[1] * mult_r[1], acc[2] * mult_r[2], acc[3] * mult_r[3]);
    }
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

// conv2 bias gradient (tensor-core path; the fp32 path gets it as an extra GEMM column): every pooling window routes
// its upstream gradient to exactly one conv2 output, so sum_px dz2[px][c] = sum_pp [a2 > 0] * da2[c*49 + pp].
// 256 threads: thread = (channel, quarter of the 49 windows); all 13 load pairs of a thread are independent.
__global__ void __launch_bounds__(256) conv2_bias_grad_kernel(flb_train_args a, SimpleCnnWs ws, int use_coef) {
    const int b = blockIdx.x, k = blockIdx.y;
    if (b >= flb_bsz(a, k)) return;
    const long long kb = (long long)k * a.B + b;
    const int c = threadIdx.x >> 2, q = threadIdx.x & 3;
    const float* da2 = ws.da2 + kb * 3136 + c * 49;
    const float* a2 = ws.a2 + kb * 3136 + c * 49;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        const int pp = q + 4 * i;
        if (pp < 49) acc += a2[pp] > 0.f ? da2[pp] : 0.f;
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (use_coef == 2) {                       // per-sample DP norm pass: || db_b ||^2 joins the sample's squared gradient norm
        float sq = flb_warp_sum(q == 0 ? acc * acc : 0.f);
        if ((threadIdx.x & 31) == 0) atomicAdd(&ws.norm2[kb], sq);
        return;
    }
    if (q == 0) atomicAdd(&a.G[(long long)k * a.ld + Off::c2b + c], acc * (use_coef ? ws.coef[kb] : 1.f));
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

// max-unpool + ReLU backward into the padded NHWC dz2 grid (every position written, pads = 0).  The pooled-side
// arrays are NCHW-flattened (fc1's input order) and the grid is NHWC: they are staged through shared memory so that
// both the global reads and the 16 B global writes are coalesced.  grid (2 * B, K): one CTA per half image (8 grid rows).
__global__ void __launch_bounds__(256) unpool2_kernel(flb_train_args a, SimpleCnnWs ws, int bias_grad) {
    const int b = blockIdx.x >> 1, half = blockIdx.x & 1, k = blockIdx.y;
    if (b >= flb_bsz(a, k)) return;
    const long long kb = (long long)k * a.B + b;
    const float* da2 = ws.da2 + kb * 3136;
    const float* a2 = ws.a2 + kb * 3136;
    const uint8_t* idx = ws.idx2 + kb * 3136;
    // pooled rows needed by grid rows [8*half, 8*half + 8): ph in [4*half, 4*half + 4) (ph = 7 does not exist)
    const int ph0 = half * 4, nph = half ? 3 : 4;
    __shared__ float s_val[64][29];          // [c][(ph - ph0) * 7 + pw], value or 0 where ReLU was inactive
    __shared__ uint8_t s_code[64][29];
    for (int e = threadIdx.x; e < 64 * nph * 7; e += 256) {
        const int c = e / (nph * 7), j = e - c * (nph * 7);
        const int src = c * 49 + ph0 * 7 + j;
        s_val[c][j] = a2[src] > 0.f ? da2[src] : 0.f;
        s_code[c][j] = idx[src];
    }
    __syncthreads();
    // conv2 bias gradient (tensor-core wgrad path, no per-sample clipping): every pooling window routes its upstream
    // gradient to exactly one conv2 output, so sum_px dz2[px][c] is the sum of this staged array over the windows
    if (bias_grad && threadIdx.x < 64) {
        float t = 0.f;
        for (int j = 0; j < nph * 7; ++j) t += s_val[threadIdx.x][j];
        atomicAdd(&a.G[(long long)k * a.ld + Off::c2b + threadIdx.x], t);
    }
    float* dz = ws.z2 + kb * (PP2 * 64) + half * (8 * WP2 * 64);
    for (int e = threadIdx.x; e < 8 * WP2 * 16; e += 256) {        // (grid position, channel quad)
        const int c = (e & 15) * 4, pos = e >> 4, hl = pos >> 4, w = pos & 15, h = half * 8 + hl;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (h < 14 && w < 14) {
            const int j = ((h >> 1) - ph0) * 7 + (w >> 1), code = (h & 1) * 2 + (w & 1);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (s_code[c + q][j] == code) v[q] = s_val[c + q][j];
        }
        *reinterpret_cast<float4*>(dz + pos * 64 + c) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// conv1 weight + bias gradient of one sample (max-unpool + ReLU backward folded in): 32 x (9 + 1) values.
// dp_mode 0 (per_sample = 0): two CTAs per sample (pooled rows 0-6 / 7-13) atomically add their halves into G.
// dp_mode 1 (norm pass, per_sample = 1): one CTA per sample; the values are stored per sample in g1ps and their squared
// norm added to norm2; the clipped sum is formed later by conv1_ps_reduce_kernel.
// Registers are capped for 5 CTAs per SM (48, no spills; 123 uncapped): at 10 clients the 640 half-sample CTAs are then one
// resident wave instead of 2.2 (measured 22.8 -> 16.4 us; a cap of 3 CTAs gives 18.4 us).
template <bool PS>
__global__ void __launch_bounds__(256, PS ? 2 : 5) conv1_bwd_kernel(flb_train_args a, SimpleCnnWs ws) {
    constexpr int per_sample = PS ? 1 : 0;
    const int nhalf = per_sample ? 1 : 2;
    const int b = blockIdx.x / nhalf, half = blockIdx.x % nhalf, k = blockIdx.y;
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
    const int c = tid & 31, g = tid >> 5;
    const long long kb = (long long)k * a.B + b;
    const float* da1 = ws.da1p + kb * (PP2 * 32);
    const float* a1 = ws.a1p + kb * (PP2 * 32);
    const uint8_t* idx = ws.idx1 + kb * (196 * 32);
    const int pp0 = per_sample ? 0 : half * 98, npp = per_sample ? 196 : 98;
    // all of this thread's gradient values first (independent loads in flight together), then the stencil updates
    constexpr int MAXIT = PS ? 25 : 13;            // ceil(196 / 8) or ceil(98 / 8) pooled positions per thread
    float gvv[MAXIT];
    unsigned long long sel = 0;                    // 2-bit argmax of every iteration
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
        const int pl = g + it * 8;
        gvv[it] = 0.f;
        if (pl < npp) {
            const int pp = pp0 + pl, o = ((pp / 14) * WP2 + pp % 14) * 32 + c;
            const float av = a1[o], dv = da1[o];
            gvv[it] = av > 0.f ? dv : 0.f;
            sel |= (unsigned long long)(idx[pp * 32 + c] & 3) << (2 * it);
        }
    }
    __syncthreads();
    float acc[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] = 0.f;
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
        const int pl = g + it * 8;
        if (pl < npp) {
            const int pp = pp0 + pl, ph = pp / 14, pw = pp % 14;
            const float gv = gvv[it];
            const int si = (int)(sel >> (2 * it)) & 3;
            const int y0 = 2 * ph + (si >> 1), x0 = 2 * pw + (si & 1);     // top-left of the 3x3 window in img (halo 1)
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int q = 0; q < 3; ++q) acc[r * 3 + q] = fmaf(gv, img[y0 + r][x0 + q], acc[r * 3 + q]);
            acc[9] += gv;
        }
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

// conv1 weight + bias gradient without per-sample clipping, one CTA per half sample (pooled rows 0-6 / 7-13), warp g = pooled
// row g, lanes = channels.  The 4 x 4 input window of a pooling position is loaded ONCE for the whole warp (broadcast 8-byte
// loads, sliding along the row like conv1_fwd_pool_kernel) and the 3 x 3 patch under each lane's own argmax is picked out of
// it with register selects.  (conv1_bwd_kernel<false> read the patch straight from shared memory at lane-dependent
// addresses: 9 conflicting LDS per position -- MIO-bound; measured 15 us for ~2 us of math.)
__global__ void __launch_bounds__(C1_THREADS, 5) conv1_bwd_rows_kernel(flb_train_args a, SimpleCnnWs ws) {
    const int b = blockIdx.x >> 1, half = blockIdx.x & 1, k = blockIdx.y;
    if (b >= flb_bsz(a, k)) return;
    __shared__ __align__(16) float img[16][32];   // input rows 14*half - 1 .. 14*half + 14 with a zero halo (col c = x column c - 1)
    __shared__ float part[7][32][10];
    const int tid = threadIdx.x;
    const long long s = a.sample_off[k] + (long long)(*a.step_ctr) * a.B + b;
    const float* x = a.x + s * 784;
    const int r0 = 14 * half - 1;
    for (int i = tid; i < 16 * 32; i += C1_THREADS) {
        const int r = i >> 5, c = i & 31, gr = r0 + r;
        img[r][c] = (gr >= 0 && gr < 28 && c >= 1 && c <= 28) ? x[gr * 28 + (c - 1)] : 0.f;
    }
    const int c = tid & 31, phl = tid >> 5, ph = half * 7 + phl;
    const long long kb = (long long)k * a.B + b;
    const float* da1 = ws.da1p + kb * (PP2 * 32) + (ph * WP2) * 32 + c;
    const float* a1 = ws.a1p + kb * (PP2 * 32) + (ph * WP2) * 32 + c;
    const uint8_t* idx = ws.idx1 + kb * (196 * 32) + (ph * 14) * 32 + c;
    // the row's 14 gradient values and argmaxes first: 42 independent coalesced loads in flight
    float gvv[14];
    unsigned sel = 0;
#pragma unroll
    for (int pw = 0; pw < 14; ++pw) {
        const float av = a1[pw * 32], dv = da1[pw * 32];
        gvv[pw] = av > 0.f ? dv : 0.f;
        sel |= (unsigned)(idx[pw * 32] & 3) << (2 * pw);
    }
    __syncthreads();
    float acc[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] = 0.f;
    float p[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 v = *reinterpret_cast<const float2*>(&img[2 * phl + i][0]);
        p[i][2] = v.x; p[i][3] = v.y;
    }
#pragma unroll
    for (int pw = 0; pw < 14; ++pw) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 v = *reinterpret_cast<const float2*>(&img[2 * phl + i][2 * pw + 2]);
            p[i][0] = p[i][2]; p[i][1] = p[i][3];
            p[i][2] = v.x; p[i][3] = v.y;
        }
        const float gv = gvv[pw];
        const bool down = (sel >> (2 * pw + 1)) & 1, right = (sel >> (2 * pw)) & 1;     // argmax = 2 * (row in window) + (col in window)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            float row[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) row[j] = down ? p[r + 1][j] : p[r][j];
#pragma unroll
            for (int q = 0; q < 3; ++q) acc[r * 3 + q] = fmaf(gv, right ? row[q + 1] : row[q], acc[r * 3 + q]);
        }
        acc[9] += gv;
    }
#pragma unroll
    for (int i = 0; i < 10; ++i) part[phl][c][i] = acc[i];
    __syncthreads();
    float* G = a.G + (long long)k * a.ld;
    for (int e = tid; e < 320; e += C1_THREADS) {
        const int cc = e / 10, i = e % 10;
        float v = 0.f;
#pragma unroll
        for (int gg = 0; gg < 7; ++gg) v += part[gg][cc][i];
        atomicAdd(i == 9 ? &G[Off::c1b + cc] : &G[Off::c1w + cc * 9 + i], v);
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

const ConvGeom kConv2{32, 64, 14, 14, 16, 16};
constexpr int kLdt = 9 * 64 * 32;          // tap-major conv2 weights per client

int forward(const flb_train_args& a, const SimpleCnnWs& ws, cudaStream_t st, bool backward = false) {
    const int K = a.K, B = a.B;
    const dim3 per_sample(B, K);
    MARK("begin");          // hpre (split-K accumulator of fc1) is kept at zero by its consumer, head_fwd_bwd_kernel
    conv1_fwd_pool_kernel<<<dim3(2 * B, K), C1_THREADS, 0, st>>>(a, ws);
    MARK("conv1_fwd_pool");
    const int tcm = tc_mask_of(a);
    if (tcm & TC_CONV2_FWD) {          // bias + ReLU + max-pool fused into the GEMM epilogue: z2 is never written
        if (int rc = tc::conv_fwd_pool_32_64(a, kConv2, ws.a1p, ws.a2, ws.idx2, ws.wt, kLdt, Off::c2b, st)) return rc;
        MARK("conv2_fwd_pool");
    } else {
        ConvFwdProb p{}; p.a = a; p.g = kConv2; p.xin_all = ws.a1p; p.z_all = ws.z2; p.woff = Off::c2w; p.boff = Off::c2b;
        simt::launch(p, B * PP2, 64, 1, K, st);
        MARK("conv2_fwd");
        pool2_kernel<<<dim3(7 * B, K), 112, 0, st>>>(a, ws);
        MARK("pool2");
    }
    if (fuse_fc1_block(a, backward)) {          // forward, head and the fc1 dgrad in one launch
        if (int rc = tc::fc1_fused(a, ws, st)) return rc;
        MARK("fc1_fused");
        FLB_LAUNCH_CHECK();
        return FLB_OK;
    }
    if (tcm & TC_FC1_FWD) {
        if (int rc = tc::fc_fwd(a, ws.a2, ws.hpre, 3136, 128, Off::f1w, 7, st)) return rc;
    } else {
        LinFwdProb p{}; p.a = a; p.In = 3136; p.Out = 128; p.woff = Off::f1w; p.act_all = ws.a2; p.out_all = ws.hpre;
        simt::launch(p, B, 128, 14, K, st);
    }
    MARK("fc1_fwd");
    head_fwd_bwd_kernel<<<dim3(K, HEAD_PARTS), 256, 0, st>>>(a, ws);
    MARK("head_fwd_bwd");
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

// gradients that are accumulated with atomics must start from zero: G[0, f1w) (everything except fc1.weight / fc2, which
// are plain stores).  Inside an epoch the optimizer kernel re-zeroes them as it consumes them.
int zero_accumulators(const flb_train_args& a, const SimpleCnnWs&, cudaStream_t st) {
    FLB_CUDA(cudaMemset2DAsync(a.G, a.ld * sizeof(float), 0, Off::f1w * sizeof(float), a.K, st));
    return FLB_OK;
}

int forward_backward(const flb_train_args& a, cudaStream_t st, bool zero_first, bool step) {
    SimpleCnnWs ws;
    simplecnn_ws_carve(a.ws, a.K, a.B, &ws);
    const int K = a.K, B = a.B;
    const dim3 per_sample(B, K);
    if (zero_first)
        if (int rc = zero_accumulators(a, ws, st)) return rc;
    if (a.dp_mode == 1) FLB_CUDA(cudaMemsetAsync(ws.norm2, 0, sizeof(float) * (size_t)K * B, st));
    if (int rc = forward(a, ws, st, true)) return rc;
    const bool fused_fc1 = fuse_fc1_block(a, true);

    // Weight gradients do not feed the activation-gradient chain: without per-sample clipping they run on a side
    // stream beside it (fork / join by events -- one graph with parallel branches when the epoch is captured).
    SideLane* lane = (a.dp_mode == 0 && !g_prof.on) ? flb_side_lane() : nullptr;
    const int tcm = tc_mask_of(a);
    const float* coef = nullptr;
    const bool fused_adam = fuse_fc1_adam(a, step);

    auto wgrad_head = [&](cudaStream_t st) -> int {
        head_wgrad_kernel<<<K, 256, 0, st>>>(a, ws, a.dp_mode == 1);
        MARK("head_wgrad");
        return FLB_OK;
    };
    auto wgrad_fc1 = [&](cudaStream_t st) -> int {
        if (tcm & TC_FC1_WGRAD) {
            if (coef) scale_rows_kernel<<<per_sample, 256, 0, st>>>(a, ws.dh, coef, 128);
            if (int rc = tc::fc_wgrad(a, ws.dh, ws.a2, 3136, 128, Off::f1w, st, fused_adam)) return rc;
            if (fused_adam) { MARK("fc1_wgrad_adam"); return FLB_OK; }
        } else {
            LinWgradProb p{}; p.a = a; p.In = 3136; p.Out = 128; p.woff = Off::f1w; p.boff = Off::f1b;
            p.dout_all = ws.dh; p.act_all = ws.a2; p.coef_all = coef;
            simt::launch(p, 128, 3136, 1, K, st);
        }
        MARK("fc1_wgrad");
        return FLB_OK;
    };
    auto wgrads_head_fc1 = [&](cudaStream_t st) -> int {
        if (int rc = wgrad_head(st)) return rc;
        return wgrad_fc1(st);
    };
    auto wgrads_conv2 = [&](cudaStream_t st) -> int {
        if (tcm & TC_CONV2_WGRAD) {
            if (coef) conv2_bias_grad_kernel<<<per_sample, 256, 0, st>>>(a, ws, 1);      // else: fused into unpool2
            if (coef) scale_rows_kernel<<<per_sample, 256, 0, st>>>(a, ws.z2, coef, PP2 * 64);
            FLB_CUDA(cudaMemsetAsync(ws.gt, 0, sizeof(float) * (size_t)K * kLdt, st));       // side lane: off the critical path
            MARK("conv2_bias_grad");
            if (int rc = tc::conv_wgrad(a, kConv2, ws.a1p, ws.z2, ws.gt, kLdt, st)) return rc;
        } else {
            ConvWgradProb p{}; p.a = a; p.g = kConv2; p.dz_all = ws.z2; p.xin_all = ws.a1p; p.coef_all = coef;
            p.woff = Off::c2w; p.boff = Off::c2b;
            simt::launch(p, 64, 289, 16, K, st);
        }
        MARK("conv2_wgrad");
        return FLB_OK;
    };

    if (lane) {
        FLB_CUDA(cudaEventRecord(lane->ev[0], st));
        FLB_CUDA(cudaStreamWaitEvent(lane->s, lane->ev[0], 0));
        if (int rc = wgrad_head(lane->s)) return rc;
        if (!fused_adam)
            if (int rc = wgrad_fc1(lane->s)) return rc;
    }

    // ---- activation gradients ----
    if (fused_fc1) {
        // da2 is already there
    } else if (tcm & TC_FC1_DGRAD) {
        if (int rc = tc::fc_dgrad(a, ws.dh, ws.da2, 3136, 128, Off::f1w, st)) return rc;
        MARK("fc1_dgrad");
    } else {
        LinDgradProb p{}; p.a = a; p.In = 3136; p.Out = 128; p.woff = Off::f1w; p.dout_all = ws.dh; p.dact_all = ws.da2;
        simt::launch(p, B, 3136, 1, K, st);
        MARK("fc1_dgrad");
    }
    if (lane && fused_adam) {                  // the fused epilogue overwrites fc1.weight: only after its last reader (the dgrad)
        FLB_CUDA(cudaEventRecord(lane->ev[3], st));
        FLB_CUDA(cudaStreamWaitEvent(lane->s, lane->ev[3], 0));
        if (int rc = wgrad_fc1(lane->s)) return rc;
    }
    const int fused_bias = ((tcm & TC_CONV2_WGRAD) && a.dp_mode == 0) ? 1 : 0;
    unpool2_kernel<<<dim3(2 * B, K), 256, 0, st>>>(a, ws, fused_bias);
    MARK("unpool2");
    // with the optimizer fused into fc1's wgrad that kernel is a long memory stream: conv2's wgrad gets a lane of its own
    cudaStream_t conv2_lane = lane ? (fused_adam ? lane->s2 : lane->s) : nullptr;
    if (lane) {
        FLB_CUDA(cudaEventRecord(lane->ev[1], st));
        FLB_CUDA(cudaStreamWaitEvent(conv2_lane, lane->ev[1], 0));
        if (int rc = wgrads_conv2(conv2_lane)) return rc;
    }
    if (tcm & TC_CONV2_DGRAD) {
        if (int rc = tc::conv_dgrad(a, kConv2, ws.z2, ws.da1p, ws.wt, kLdt, st)) return rc;
    } else {
        ConvDgradProb p{}; p.a = a; p.g = kConv2; p.dz_all = ws.z2; p.dx_all = ws.da1p; p.woff = Off::c2w;
        simt::launch(p, B * PP2, 32, 1, K, st);
    }
    MARK("conv2_dgrad");

    // ---- per-sample clip coefficients (dp_mode 1) ----
    // The four norm kernels only meet in norm2 (atomics), and the weight-gradient GEMMs only need the clip coefficients:
    // both groups are split over the main stream and a side stream (fork / join by events, capturable).
    SideLane* lane_ps = (a.dp_mode == 1 && !g_prof.on && !getenv("FLB_NO_SIDE_LANE")) ? flb_side_lane() : nullptr;
    if (a.dp_mode == 1) {
        if (lane_ps) {
            FLB_CUDA(cudaEventRecord(lane_ps->ev[0], st));
            FLB_CUDA(cudaStreamWaitEvent(lane_ps->s, lane_ps->ev[0], 0));
        }
        conv1_bwd_kernel<true><<<per_sample, 256, 0, lane_ps ? lane_ps->s : st>>>(a, ws);
        MARK("conv1_ps_grad");
        linear_ghost_norm_kernel<<<per_sample, 128, 0, st>>>(a, ws);
        MARK("linear_ghost_norm");
        if (tcm & TC_CONV2_WGRAD) {          // per-sample conv2 gradient tiles live in TMEM only (tcgen05), squared on the way out
            if (int rc = tc::conv_wgrad_norm_32_64(a, kConv2, ws.a1p, ws.z2, ws.norm2, st)) return rc;
            MARK("conv2_wgrad_norm");
            conv2_bias_grad_kernel<<<per_sample, 256, 0, st>>>(a, ws, 2);
        } else {
            ConvWgradNormProb p{}; p.a = a; p.g = kConv2; p.dz_all = ws.z2; p.xin_all = ws.a1p; p.norm2_all = ws.norm2;
            simt::launch(p, 64, 289, 1, K * B, st);
        }
        if (lane_ps) {
            FLB_CUDA(cudaEventRecord(lane_ps->ev[1], lane_ps->s));
            FLB_CUDA(cudaStreamWaitEvent(st, lane_ps->ev[1], 0));
        }
        clip_coef_kernel<<<flb_cdiv(K * B, 256), 256, 0, st>>>(a, ws);
        coef = ws.coef;
        MARK("clip_coef");
    }

    // ---- weight gradients ----
    if (lane_ps) {
        FLB_CUDA(cudaEventRecord(lane_ps->ev[2], st));
        FLB_CUDA(cudaStreamWaitEvent(lane_ps->s, lane_ps->ev[2], 0));
        if (int rc = wgrads_head_fc1(lane_ps->s)) return rc;
        if (int rc = wgrads_conv2(st)) return rc;
    } else if (!lane) {
        if (int rc = wgrads_head_fc1(st)) return rc;
        if (int rc = wgrads_conv2(st)) return rc;
    }
    if (a.dp_mode == 1) conv1_ps_reduce_kernel<<<K, 320, 0, st>>>(a, ws);
    else conv1_bwd_rows_kernel<<<dim3(2 * B, K), C1_THREADS, 0, st>>>(a, ws);
    MARK("conv1_wgrad");
    if (lane) {
        FLB_CUDA(cudaEventRecord(lane->ev[2], lane->s));
        FLB_CUDA(cudaStreamWaitEvent(st, lane->ev[2], 0));
        if (conv2_lane != lane->s) {
            FLB_CUDA(cudaEventRecord(lane->ev[4], conv2_lane));
            FLB_CUDA(cudaStreamWaitEvent(st, lane->ev[4], 0));
        }
    }
    if (lane_ps) {
        FLB_CUDA(cudaEventRecord(lane_ps->ev[3], lane_ps->s));
        FLB_CUDA(cudaStreamWaitEvent(st, lane_ps->ev[3], 0));
    }
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

}  // namespace

namespace simplecnn {
int num_params() { return Off::P; }
long long ws_bytes(int K, int B) { return (long long)simplecnn_ws_carve(nullptr, K, B, nullptr); }
long long ws_offset(int K, int B, const char* name) {
    SimpleCnnWs ws;
    simplecnn_ws_carve((void*)0, K, B, &ws);
#define FIELD(f) if (!strcmp(name, #f)) return (long long)(uintptr_t)ws.f;
    FIELD(a1p) FIELD(idx1) FIELD(z2) FIELD(a2) FIELD(idx2) FIELD(hpre) FIELD(h) FIELD(logits) FIELD(dlog) FIELD(dh)
    FIELD(da2) FIELD(da1p) FIELD(norm2) FIELD(coef) FIELD(g1ps)
#undef FIELD
    return -1;
}
int forward(const flb_train_args& a, cudaStream_t st) {
    SimpleCnnWs ws;
    simplecnn_ws_carve(a.ws, a.K, a.B, &ws);
    return ::forward(a, ws, st);
}
int forward_backward(const flb_train_args& a, cudaStream_t st, bool zero_first, bool step) { return ::forward_backward(a, st, zero_first, step); }
int begin_epoch_zero(const flb_train_args& a, cudaStream_t st) {
    SimpleCnnWs ws;
    simplecnn_ws_carve(a.ws, a.K, a.B, &ws);
    return zero_accumulators(a, ws, st);
}
int step_launches(const flb_train_args& a) {
    const int m = tc_mask_of(a);
    int n = 12 - ((m & TC_CONV2_FWD) ? 1 : 0) - (fuse_fc1_block(a, true) ? 2 : 0);    // fused pool: one kernel less; conv2 bias gradient: fused into unpool2                 // + conv2_bias_grad (an extra GEMM column on the fp32 path)
    if (a.dp_mode == 1) n += 4 + ((m & TC_FC1_WGRAD) ? 1 : 0) + ((m & TC_CONV2_WGRAD) ? 3 : 0);
    return n;
}
void tc_tab(const flb_train_args& a, TcConvTab* t, bool step) {
    const int m = tc_mask_of(a);
    t->g_zero_upto = Off::f1w;
    if (fuse_fc1_adam(a, step)) { t->skip_lo = Off::f1w; t->skip_hi = Off::f1b; }
    if (!(m & (TC_CONV2_FWD | TC_CONV2_DGRAD | TC_CONV2_WGRAD))) return;
    SimpleCnnWs ws;
    simplecnn_ws_carve(a.ws, a.K, a.B, &ws);
    t->n = 1;
    t->woff[0] = Off::c2w; t->cin[0] = 32; t->cout[0] = 64; t->toff[0] = 0;
    t->gt_live[0] = (m & TC_CONV2_WGRAD) ? 1 : 0;
    t->ldt = kLdt; t->wt = ws.wt; t->gt = ws.gt;
}
}  // namespace simplecnn
