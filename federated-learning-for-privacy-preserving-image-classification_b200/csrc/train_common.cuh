// Shared definitions for the batched-over-clients training kernels.
//
// Everything a step needs lives in HBM as client-major arrays: parameters / gradients / optimizer moments
// [K, ld] (reference layer order and layouts, see layout.py), the clients' samples as one concatenated
// [sum N_c, C*H*W] array, and an activation workspace carved by flb_train_ws_layout().  One launch of each
// kernel serves ALL clients resident on the GPU (blockIdx.z / .y = client); per-step variation (which batch,
// ragged last batch, finished clients) is derived on the device from *step_ctr, so a whole epoch is a fixed
// launch sequence that can be captured in one CUDA graph.
#pragma once
#include "flb_common.cuh"
#include "../../include/flb.h"

// conv activations are stored NHWC on a zero-padded (Hp x Wp) grid per image so that a 3x3 tap is a constant
// row shift of the flattened [pixels, C] matrix (implicit GEMM without im2col; TMA-friendly)
// Only the right column(s) and bottom row(s) need to be pad: the left / top neighbours of an image's first column /
// row are the previous row's / previous image's pad positions.  PPv (rows per image, >= (H+1)*Wp + 1) overrides Hp*Wp.
struct ConvGeom {
    int Cin, Cout, H, W, Hp, Wp;
    int PPv = 0;
    __host__ __device__ int PP() const { return PPv ? PPv : Hp * Wp; }
};

__device__ __forceinline__ int flb_bsz(const flb_train_args& a, int client) {
    const int s = *a.step_ctr;
    int r = a.nsamples[client] - s * a.B;
    return r < 0 ? 0 : (r > a.B ? a.B : r);
}

// Philox key of this epoch's dropout masks / per-sample-DP noise: the caller's seed advanced by the never-reset epoch
// counter, so the optimizer step count t (reset by every train_local_model call) can stay the in-epoch counter part.
__device__ __forceinline__ unsigned long long flb_epoch_seed(const flb_train_args& a) {
    return a.seed + (a.epoch_nonce ? *a.epoch_nonce : 0ull) * 0x9E3779B97F4A7C15ull;
}

// SimpleCNN parameter offsets in a row (reference named_parameters order, models_pytorch.py:69-80)
struct SimpleCnnOff {
    static constexpr int c1w = 0, c1b = 288, c2w = 320, c2b = 18752, f1w = 18816, f1b = 420224, f2w = 420352,
                         f2b = 421632, P = 421642;
};

// workspace carve-up (element offsets are per client; total = K * per-client size)
struct SimpleCnnWs {
    float* a1p;      // [K][B][16*16][32]   conv1 output after ReLU+pool, padded NHWC (conv2 input)
    uint8_t* idx1;   // [K][B][196][32]     pool-1 argmax (0..3)
    float* z2;       // [K][B][16*16][64]   conv2 pre-activation (fp32 path) / dz2 in backward
    float* a2;       // [K][B][3136]        conv2 output after ReLU+pool, NCHW-flattened (fc1 input)
    uint8_t* idx2;   // [K][B][3136]
    float* hpre;     // [K][B][128]         fc1 pre-activation without bias (split-K accumulated)
    float* h;        // [K][B][128]         after bias, ReLU, dropout
    float* logits;   // [K][B][10]
    float* dlog;     // [K][B][10]
    float* dh;       // [K][B][128]
    float* da2;      // [K][B][3136]
    float* da1p;     // [K][B][16*16][32]
    float* norm2;    // [K][B]              per-sample squared gradient norms (dp_mode 1)
    float* coef;     // [K][B]              per-sample clip coefficients
    float* g1ps;     // [K][B][320]         per-sample conv1 weight+bias gradients (dp_mode 1)
    float* wt;       // [K][9][64][32]      conv2 weights, tap-major (tensor-core path)
    float* gt;       // [K][9][64][32]      conv2 weight gradients, tap-major
    int* fc1_ctr;    // [K][2]              arrive / depart counters of the fused classifier kernel's client barrier (zero at rest)
};

// per-kernel CUDA-event timing of one step (flb_train_step_profiled); inactive otherwise
struct StepProfile {
    bool on = false;
    int n = 0;
    cudaEvent_t ev[96];
    const char* name[96];
};
extern StepProfile g_prof;
#define MARK(label)                                                       \
    do {                                                                  \
        if (g_prof.on && g_prof.n < 96) {                                 \
            cudaEventRecord(g_prof.ev[g_prof.n], st);                     \
            g_prof.name[g_prof.n++] = label;                              \
        }                                                                 \
    } while (0)

// Tensor-core conv layers keep a second, tap-major copy of their weights, Wt[tap][Cout][Cin] per client (both GEMM
// operands become plain TMA boxes), and accumulate their weight gradients in the same layout (Gt).  The optimizer
// kernel reads Gt / writes Wt through this table, so the reference-layout rows W / G stay the API.
struct TcConvTab {
    int n = 0;                       // tensor-core conv layers (0: fp32 path)
    int woff[5], cin[5], cout[5], toff[5];
    int gt_live[5];                  // this layer's wgrad accumulated into Gt (else into the reference-layout row G)
    int ldt = 0;
    float* wt = nullptr;             // [K, ldt]
    float* gt = nullptr;             // [K, ldt]
    // Reference-layout gradients that are ACCUMULATED with atomics, G[0, g_zero_upto), are kept at zero between steps:
    // the optimizer kernel clears what it has consumed, so no memset node sits on the step's critical path.
    // (Gt is cleared by a memset on the weight-gradient side lane: clearing it from the optimizer's scattered
    // tap-major accesses was measured 55 us slower.)  0 = the model's step zeroes its accumulators itself.
    int g_zero_upto = 0;
    // Parameters [skip_lo, skip_hi) (multiples of 4) are NOT touched by the optimizer kernel in this step: the weight-gradient
    // GEMM that produced their gradient applied the optimizer in its epilogue (SimpleCNN fc1.weight, 95 % of the model).
    int skip_lo = 0, skip_hi = 0;
};
// reference-layout offset p -> tap-major offset (or -1 when p is not a tensor-core conv weight)
__device__ __forceinline__ int tc_tab_map(const TcConvTab& t, int p, int& layer) {
    for (int i = 0; i < t.n; ++i) {
        const int r = p - t.woff[i];
        if (r >= 0 && r < t.cout[i] * t.cin[i] * 9) {
            const int tap = r % 9, ci = (r / 9) % t.cin[i], co = r / (9 * t.cin[i]);
            layer = i;
            return t.toff[i] + (tap * t.cout[i] + co) * t.cin[i] + ci;
        }
    }
    return -1;
}

// model-specific launch sequences (train_simplecnn.cu, train_cifar.cu), dispatched by train_api.cu
namespace simplecnn {
int num_params();
long long ws_bytes(int K, int B);
long long ws_offset(int K, int B, const char* name);
int forward(const flb_train_args& a, cudaStream_t st);
// step: called from flb_train_step (an optimizer step follows): layers whose optimizer update is fused into their
// weight-gradient epilogue are updated here and reported through TcConvTab::skip_* by tc_tab(a, t, true)
int forward_backward(const flb_train_args& a, cudaStream_t st, bool zero_first, bool step);
int begin_epoch_zero(const flb_train_args& a, cudaStream_t st);
int step_launches(const flb_train_args& a);
void tc_tab(const flb_train_args& a, TcConvTab* t, bool step = false);
}
namespace cifar {
int num_params();
long long bn_floats();
long long ws_bytes(int K, int B);
long long ws_offset(int K, int B, const char* name);
int forward(const flb_train_args& a, cudaStream_t st);
int forward_backward(const flb_train_args& a, cudaStream_t st, bool step = false);
int step_launches(const flb_train_args& a);
void tc_tab(const flb_train_args& a, TcConvTab* t, bool step = false);
}

static inline size_t flb_align256(size_t v) { return (v + 255) & ~(size_t)255; }

static inline size_t simplecnn_ws_carve(void* base, int K, int B, SimpleCnnWs* ws) {
    size_t off = 0;
    char* p = (char*)base;
    const size_t KB = (size_t)K * B;
#define CARVE(field, type, count)                                  \
    do {                                                           \
        if (ws) ws->field = (type*)(p + off);                      \
        off = flb_align256(off + sizeof(type) * (size_t)(count));  \
    } while (0)
    CARVE(a1p, float, KB * 256 * 32);
    CARVE(idx1, uint8_t, KB * 196 * 32);
    CARVE(z2, float, KB * 256 * 64);
    CARVE(a2, float, KB * 3136);
    CARVE(idx2, uint8_t, KB * 3136);
    CARVE(hpre, float, KB * 128);
    CARVE(h, float, KB * 128);
    CARVE(logits, float, KB * 10);
    CARVE(dlog, float, KB * 10);
    CARVE(dh, float, KB * 128);
    CARVE(da2, float, KB * 3136);
    CARVE(da1p, float, KB * 256 * 32);
    CARVE(norm2, float, KB);
    CARVE(coef, float, KB);
    CARVE(g1ps, float, KB * 320);
    CARVE(wt, float, (size_t)K * 18432);
    CARVE(gt, float, (size_t)K * 18432);
    CARVE(fc1_ctr, int, (size_t)K * 2);
#undef CARVE
    return off;
}
