// Shared helpers for the flb (federated-learning B200) C-ABI library.
// sm_100a only; no CPU fallback anywhere in this tree.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#define FLB_OK 0
#define FLB_ERR_ARG -1
#define FLB_ERR_CUDA -2
#define FLB_ERR_NODEV -3
#define FLB_ERR_UNSUPPORTED -4

#define FLB_NUM_SMS_B200 148

void flb_set_error(const char* fmt, ...);

#define FLB_CHECK_ARG(cond, ...)                    \
    do {                                            \
        if (!(cond)) {                              \
            flb_set_error(__VA_ARGS__);             \
            return FLB_ERR_ARG;                     \
        }                                           \
    } while (0)

#define FLB_CUDA(call)                                                             \
    do {                                                                           \
        cudaError_t _e = (call);                                                   \
        if (_e != cudaSuccess) {                                                   \
            flb_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e),  \
                          __FILE__, __LINE__);                                     \
            return FLB_ERR_CUDA;                                                   \
        }                                                                          \
    } while (0)

#define FLB_LAUNCH_CHECK()                                                         \
    do {                                                                           \
        cudaError_t _e = cudaGetLastError();                                       \
        if (_e != cudaSuccess) {                                                   \
            flb_set_error("kernel launch failed: %s (%s:%d)",                      \
                          cudaGetErrorString(_e), __FILE__, __LINE__);             \
            return FLB_ERR_CUDA;                                                   \
        }                                                                          \
    } while (0)

struct SideLane { bool ready = false; cudaStream_t s = nullptr, s2 = nullptr; cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; };
SideLane* flb_side_lane();   // per host thread and device; nullptr if it cannot be created

int flb_num_sms();   // SM count of the current device (148 on B200), cached

static inline int flb_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// CTAs of `kernel` that are resident at once on the whole device (occupancy API x SM count).  Grid-stride kernels whose
// CTAs live for the whole launch are sized to AT MOST this many: a few CTAs beyond one wave double the duration.
template <class F>
static inline int flb_resident_ctas(F kernel, int threads, size_t dyn_smem = 0) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem) != cudaSuccess || per_sm < 1) {
        (void)cudaGetLastError();
        per_sm = 1;
    }
    return per_sm * flb_num_sms();
}

__device__ __forceinline__ float flb_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double flb_warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float flb_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
