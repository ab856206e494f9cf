// One-pass update validation and convergence reductions.
//
// The reference validates every incoming update with three full passes and three host syncs PER TENSOR
// (torch.isnan(t).any(), torch.isinf(t).any(), torch.abs(t).max().item(); src/shared/validation.py:72-91) and measures
// convergence with two norms per layer (src/aggregation/fedavg.py:144-190, src/aggregation/convergence.py:189-217).
// Here every (client, layer) tensor is read once by one launch: NaN / Inf flags and max|x| per tensor; and
// ||new - old||^2, ||new||^2 per layer in double for the convergence metric.  Tensors are addressed through a device
// pointer table, so the same kernels serve the stacked [K, ld] client rows and updates that live as separate tensors.
#include "flb_common.cuh"
#include "../../include/flb.h"

namespace {

constexpr int kThreads = 256;

// grid (chunks, L, K).  ptrs[k*L + l]: tensor l of client k; seg_off[l+1] - seg_off[l]: its element count.
__global__ void __launch_bounds__(kThreads)
update_stats_kernel(const float* const* __restrict__ ptrs, const long long* __restrict__ seg_off,
                    unsigned int* __restrict__ maxkey, unsigned int* __restrict__ flags, int L) {
    const int l = blockIdx.y, k = blockIdx.z;
    const float* __restrict__ t = ptrs[(long long)k * L + l];
    const long long n = seg_off[l + 1] - seg_off[l];
    float mx = 0.f;
    unsigned int f = 0;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
        const float v = t[i];
        if (v != v) f |= 1u;                               // NaN  (validation.py:81)
        else if (isinf(v)) f |= 2u;                        // Inf  (validation.py:84)
        mx = fmaxf(mx, fabsf(v));                          // max |x| (validation.py:87); fmaxf drops NaN
    }
    mx = flb_warp_max(mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) f |= __shfl_xor_sync(0xffffffffu, f, o);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&maxkey[(long long)k * L + l], __float_as_uint(mx));      // non-negative floats order like their bit patterns
        if (f) atomicOr(&flags[(long long)k * L + l], f);
    }
}

// grid (chunks, L).  out[2*l] += sum (new - old)^2, out[2*l + 1] += sum new^2   (double)
__global__ void __launch_bounds__(kThreads)
delta_norms_kernel(const float* const* __restrict__ new_ptrs, const float* const* __restrict__ old_ptrs,
                   const long long* __restrict__ seg_off, double* __restrict__ out) {
    const int l = blockIdx.y;
    const float* __restrict__ a = new_ptrs[l];
    const float* __restrict__ b = old_ptrs[l];
    const long long n = seg_off[l + 1] - seg_off[l];
    double d2 = 0.0, n2 = 0.0;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
        const float x = a[i], d = __fsub_rn(x, b[i]);
        d2 += (double)d * d;
        n2 += (double)x * x;
    }
    d2 = flb_warp_sum_d(d2);
    n2 = flb_warp_sum_d(n2);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out[2 * l], d2);
        atomicAdd(&out[2 * l + 1], n2);
    }
}

int chunks_for(long long total, int groups) {
    long long c = (total / groups + (long long)kThreads * 8 - 1) / ((long long)kThreads * 8);
    const long long cap = ((long long)flb_num_sms() * 16) / groups + 1;
    if (c > cap) c = cap;
    return (int)(c < 1 ? 1 : (c > 1024 ? 1024 : c));
}

}  // namespace

extern "C" int flb_update_stats(const float* const* ptrs, const long long* seg_off, float* max_abs, unsigned int* flags,
                                int K, int L, long long P, void* stream) {
    FLB_CHECK_ARG(ptrs && seg_off && max_abs && flags, "flb_update_stats: null pointer");
    FLB_CHECK_ARG(K >= 1 && K <= 65535 && L >= 1 && L <= 65535, "flb_update_stats: bad K/L");
    cudaStream_t st = (cudaStream_t)stream;
    FLB_CUDA(cudaMemsetAsync(max_abs, 0, sizeof(float) * (size_t)K * L, st));
    FLB_CUDA(cudaMemsetAsync(flags, 0, sizeof(unsigned int) * (size_t)K * L, st));
    if (P == 0) return FLB_OK;
    update_stats_kernel<<<dim3(chunks_for(P, L * K), L, K), kThreads, 0, st>>>(
        ptrs, seg_off, reinterpret_cast<unsigned int*>(max_abs), flags, L);
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

extern "C" int flb_delta_norms(const float* const* new_ptrs, const float* const* old_ptrs, const long long* seg_off,
                               double* out, int L, long long P, void* stream) {
    FLB_CHECK_ARG(new_ptrs && old_ptrs && seg_off && out, "flb_delta_norms: null pointer");
    FLB_CHECK_ARG(L >= 1 && L <= 65535, "flb_delta_norms: bad L");
    cudaStream_t st = (cudaStream_t)stream;
    FLB_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * 2 * (size_t)L, st));
    if (P == 0) return FLB_OK;
    delta_norms_kernel<<<dim3(chunks_for(P, L), L), kThreads, 0, st>>>(new_ptrs, old_ptrs, seg_off, out);
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}
