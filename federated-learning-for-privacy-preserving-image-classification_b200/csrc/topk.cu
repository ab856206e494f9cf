// Top-k sparsification of client updates on the device (reference TopKSparsificationCompressor,
// src/shared/compression.py:327-365): for every (client, layer) keep the k = max(1, int(n * (1 - sparsity))) entries of
// largest |x| as (values, indices) and scatter them back into zeros on the receiving side.
//
// One CTA per (layer, client) runs an exact radix select on the 32-bit pattern of |x| (monotone for non-negative
// floats): four 8-bit histogram passes in shared memory find the k-th largest key T and how many entries equal to T are
// still needed; a final pass compacts, in INDEX order, every entry with key > T plus the first `need` entries with
// key == T (lowest index first, so ties are deterministic).  torch.topk returns the pairs sorted by magnitude and leaves
// the tie order unspecified; the dense reconstruction -- the only thing the reference consumes -- is identical.
#include "flb_common.cuh"
#include "../../include/flb.h"

namespace {

constexpr int kThreads = 1024;

__device__ __forceinline__ uint32_t abs_key(float v) { return __float_as_uint(v) & 0x7fffffffu; }

// grid (L, K).  x: [K, ld]; seg_off[L+1] layer spans; kk[L] entries to keep per layer; out_off[L+1] prefix sums of kk.
// idx_out / val_out: [K, ldk] with ldk >= out_off[L]; indices are relative to the layer start.
__global__ void __launch_bounds__(kThreads)
topk_select_kernel(const float* __restrict__ x, long long ld, const long long* __restrict__ seg_off,
                   const int* __restrict__ kk, const long long* __restrict__ out_off,
                   int* __restrict__ idx_out, float* __restrict__ val_out, long long ldk) {
    const int l = blockIdx.x, c = blockIdx.y, tid = threadIdx.x;
    const long long b = seg_off[l];
    const int n = (int)(seg_off[l + 1] - b);
    const int k = min(kk[l], n);
    if (n == 0 || k <= 0) return;
    const float* __restrict__ row = x + (long long)c * ld + b;
    int* __restrict__ io = idx_out + (long long)c * ldk + out_off[l];
    float* __restrict__ vo = val_out + (long long)c * ldk + out_off[l];

    __shared__ unsigned int hist[256];
    __shared__ uint32_t s_prefix, s_mask;
    __shared__ int s_need;
    __shared__ int s_warp[2][32];
    __shared__ int s_base_sel, s_base_eq;
    if (tid == 0) { s_prefix = 0; s_mask = 0; s_need = k; }
    __syncthreads();
    for (int pass = 3; pass >= 0; --pass) {
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix, mask = s_mask;
        for (int i = tid; i < n; i += kThreads) {
            const uint32_t key = abs_key(row[i]);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> (8 * pass)) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {                                   // walk the digits from the top: the k-th largest key lives in `bin`
            int need = s_need, bin = 255;
            for (; bin > 0; --bin) {
                const int cnt = (int)hist[bin];
                if (cnt >= need) break;
                need -= cnt;
            }
            s_need = need;
            s_prefix = prefix | ((uint32_t)bin << (8 * pass));
            s_mask = mask | (0xffu << (8 * pass));
        }
        __syncthreads();
    }
    const uint32_t T = s_prefix;                          // key of the k-th largest |x|
    const int need_eq = s_need;                           // entries equal to T that are kept (lowest indices first)
    if (tid == 0) { s_base_sel = 0; s_base_eq = 0; }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    for (int i0 = 0; i0 < n; i0 += kThreads) {
        const int i = i0 + tid;
        const float v = i < n ? row[i] : 0.f;
        const uint32_t key = i < n ? abs_key(v) : 0u;
        const bool gt = i < n && key > T, eq = i < n && key == T;
        // block-wide exclusive ranks of `eq` first (to cut the ties), then of the selected flags
        const unsigned eq_b = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) s_warp[0][warp] = __popc(eq_b);
        __syncthreads();
        int eq_before = s_base_eq + __popc(eq_b & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) eq_before += s_warp[0][w];
        const bool sel = gt || (eq && eq_before < need_eq);
        const unsigned sel_b = __ballot_sync(0xffffffffu, sel);
        if (lane == 0) s_warp[1][warp] = __popc(sel_b);
        __syncthreads();
        int pos = s_base_sel + __popc(sel_b & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) pos += s_warp[1][w];
        if (sel) { io[pos] = i; vo[pos] = v; }
        __syncthreads();
        if (tid == 0) {
            int te = 0, ts = 0;
            for (int w = 0; w < kThreads / 32; ++w) { te += s_warp[0][w]; ts += s_warp[1][w]; }
            s_base_eq += te;
            s_base_sel += ts;
        }
        __syncthreads();
    }
}

// dense[c, seg_off[l] + idx] = val  (the caller zero-fills dense)
__global__ void __launch_bounds__(256)
topk_scatter_kernel(const int* __restrict__ idx, const float* __restrict__ val, long long ldk,
                    const long long* __restrict__ seg_off, const int* __restrict__ kk,
                    const long long* __restrict__ out_off, float* __restrict__ dense, long long ld) {
    const int l = blockIdx.y, c = blockIdx.z;
    const long long b = seg_off[l];
    const int n = (int)(seg_off[l + 1] - b), k = min(kk[l], n);
    const int* __restrict__ ii = idx + (long long)c * ldk + out_off[l];
    const float* __restrict__ vv = val + (long long)c * ldk + out_off[l];
    float* __restrict__ row = dense + (long long)c * ld + b;
    for (int j = blockIdx.x * 256 + threadIdx.x; j < k; j += gridDim.x * 256) {
        const int i = ii[j];
        if (i >= 0 && i < n) row[i] = vv[j];
    }
}

}  // namespace

extern "C" int flb_topk_select(const float* x, long long ld, const long long* seg_off, const int* kk,
                               const long long* out_off, int* idx_out, float* val_out, long long ldk,
                               int K, int L, void* stream) {
    FLB_CHECK_ARG(x && seg_off && kk && out_off && idx_out && val_out, "flb_topk_select: null pointer");
    FLB_CHECK_ARG(K >= 1 && K <= 65535 && L >= 1, "flb_topk_select: need 1 <= K <= 65535 and L >= 1");
    topk_select_kernel<<<dim3(L, K), kThreads, 0, (cudaStream_t)stream>>>(x, ld, seg_off, kk, out_off, idx_out, val_out, ldk);
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

extern "C" int flb_topk_scatter(const int* idx, const float* val, long long ldk, const long long* seg_off, const int* kk,
                                const long long* out_off, float* dense, long long ld, int K, int L, long long P, void* stream) {
    FLB_CHECK_ARG(idx && val && seg_off && kk && out_off && dense, "flb_topk_scatter: null pointer");
    FLB_CHECK_ARG(K >= 1 && K <= 65535 && L >= 1 && L <= 65535 && ld >= P, "flb_topk_scatter: bad K/L/ld");
    cudaStream_t st = (cudaStream_t)stream;
    FLB_CUDA(cudaMemset2DAsync(dense, ld * sizeof(float), 0, P * sizeof(float), K, st));
    topk_scatter_kernel<<<dim3(8, L, K), 256, 0, st>>>(idx, val, ldk, seg_off, kk, out_off, dense, ld);
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}
