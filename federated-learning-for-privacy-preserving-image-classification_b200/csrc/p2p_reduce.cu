// FedAvg partial sum FUSED with the cross-GPU reduction over NVLink peer memory: one kernel per rank, no NCCL call.
//
// Multi-GPU FedAvg: client i lives on rank i mod G, every rank forms sum_{i in rank} (n_i / sum n) theta_i
// (reference src/aggregation/fedavg.py:267-289, weights :247-256) and the ranks' partial sums are added -- the
// client -> coordinator hop and the next round's broadcast in one.  Until now: the FedAvg kernel, then a latency-bound
// NCCL all_reduce of P fp32 (1.7 MB / 5.9 MB).  Here the exchange rides in the FedAvg kernel itself, chunk by chunk:
//
//   phase A  every CTA walks column chunks j (CHUNK floats): the weighted sum over the rank's client rows, arithmetic
//            identical to fedavg_flat_vec4_kernel (fp32 multiply and add rounded separately, client order), is stored
//            straight into the INBOX of the chunk's owner, rank j mod G (peer store over NVLink; the owner's own chunks
//            are local stores), followed by a release flag in the owner's memory.  The transfer of chunk j overlaps the
//            math of chunk j + gridDim.x.
//   phase B  the owner of chunk j waits for the G flags of that chunk, adds the G partial sums in RANK ORDER (so every
//            rank ends up with bit-identical values, and the result does not depend on arrival order) and stores the final
//            chunk into the global row of EVERY rank, again followed by a flag.
//   phase C  every rank waits until all chunks of its own global row carry this call's epoch.
//
// Flags hold a per-call epoch (strictly increasing), so nothing is ever reset; a rank can only enter call e + 1 after phase
// C of call e, i.e. after every owner has consumed its inbox of call e, so the inbox needs no double buffering.  All CTAs of
// a launch are co-resident (grid <= resident capacity) and phase A never waits, so the waits of phases B / C cannot
// deadlock; spins are bounded and trap instead of hanging the box.
//
// Each rank owns one region (cudaMalloc, exported with cudaIpcGetMemHandle, opened by the peers -- one process per GPU):
//   [flags_a: nchunks x G u32][flags_b: nchunks u32][inbox: G x ld fp32][global: ld fp32]
#include "flb_common.cuh"
#include "../../include/flb.h"
#include <string.h>

namespace {

constexpr int kThreads = 256;
constexpr int kWChunk = 1024;
constexpr unsigned kSpinLimit = 1u << 26;        // x ~40 ns: seconds, then trap

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_flag(const unsigned* p, unsigned epoch) {
    unsigned spins = 0;
    while (ld_acquire_sys(p) != epoch) {
        __nanosleep(40);
        if (++spins > kSpinLimit) __trap();
    }
}

struct Regions { char* base[FLB_P2P_MAX_RANKS]; };

__global__ void __launch_bounds__(kThreads)
fedavg_allreduce_p2p_kernel(const float* __restrict__ theta, long long ldt, const float* __restrict__ w, int K, long long P4,
                            Regions reg, int rank, int G, long long ld, long long off_a, long long off_b, long long off_inbox,
                            long long off_global, int chunk4, unsigned epoch) {
    __shared__ float sw[kWChunk];
    const int tid = threadIdx.x;
    const long long nchunks = (P4 + chunk4 - 1) / chunk4;
    const float4* __restrict__ t4 = reinterpret_cast<const float4*>(theta);
    const long long ldt4 = ldt >> 2, ld4 = ld >> 2;
    const bool w_once = K <= kWChunk;                       // the usual case: all client weights staged once
    if (w_once) {
        for (int i = tid; i < K; i += kThreads) sw[i] = w[i];
        __syncthreads();
    }

    // ---- phase A: this rank's partial sums, pushed to the chunk owners -------------------------------------------------
    for (long long j = blockIdx.x; j < nchunks; j += gridDim.x) {
        const int owner = (int)(j % G);
        float4* dst = reinterpret_cast<float4*>(reg.base[owner] + off_inbox) + (long long)rank * ld4;
        const long long c_lo = j * chunk4, c_hi = min(P4, c_lo + chunk4);
        for (long long c0 = c_lo; c0 < c_hi; c0 += 2 * kThreads) {              // two float4 columns per thread and pass
            const long long ca = c0 + tid, cb = ca + kThreads;
            const bool la = ca < c_hi, lb = cb < c_hi;
            float4 acc_a = make_float4(0.f, 0.f, 0.f, 0.f), acc_b = acc_a;
            for (int k0 = 0; k0 < K; k0 += kWChunk) {
                const int kc = min(kWChunk, K - k0);
                if (!w_once) {
                    __syncthreads();
                    for (int i = tid; i < kc; i += kThreads) sw[i] = w[k0 + i];
                    __syncthreads();
                }
#pragma unroll 4
                for (int k = 0; k < kc; ++k) {
                    const float wk = sw[k];
                    const float4* row = t4 + (long long)(k0 + k) * ldt4;
                    float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
                    if (la) va = __ldcs(&row[ca]);
                    if (lb) vb = __ldcs(&row[cb]);
                    acc_a.x = __fadd_rn(acc_a.x, __fmul_rn(wk, va.x)); acc_a.y = __fadd_rn(acc_a.y, __fmul_rn(wk, va.y));
                    acc_a.z = __fadd_rn(acc_a.z, __fmul_rn(wk, va.z)); acc_a.w = __fadd_rn(acc_a.w, __fmul_rn(wk, va.w));
                    acc_b.x = __fadd_rn(acc_b.x, __fmul_rn(wk, vb.x)); acc_b.y = __fadd_rn(acc_b.y, __fmul_rn(wk, vb.y));
                    acc_b.z = __fadd_rn(acc_b.z, __fmul_rn(wk, vb.z)); acc_b.w = __fadd_rn(acc_b.w, __fmul_rn(wk, vb.w));
                }
            }
            if (la) dst[ca] = acc_a;
            if (lb) dst[cb] = acc_b;
        }
        __threadfence_system();
        __syncthreads();
        if (tid == 0) st_release_sys(reinterpret_cast<unsigned*>(reg.base[owner] + off_a) + j * G + rank, epoch);
    }

    // ---- phase B: chunks this rank owns -> sum over ranks in rank order, final values to every rank's global row ---------
    char* mine = reg.base[rank];
    for (long long j = rank + (long long)G * blockIdx.x; j < nchunks; j += (long long)G * gridDim.x) {
        if (tid < G) wait_flag(reinterpret_cast<const unsigned*>(mine + off_a) + j * G + tid, epoch);
        __syncthreads();
        const float4* inbox = reinterpret_cast<const float4*>(mine + off_inbox);
        const long long c_lo = j * chunk4, c_hi = min(P4, c_lo + chunk4);
        for (long long c = c_lo + tid; c < c_hi; c += kThreads) {
            float4 acc = __ldcg(&inbox[c]);                                         // rank 0's partial sum (L2: peer writes land there)
            for (int s = 1; s < G; ++s) {
                const float4 v = __ldcg(&inbox[(long long)s * ld4 + c]);
                acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y); acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
            }
            for (int s = 0; s < G; ++s) reinterpret_cast<float4*>(reg.base[s] + off_global)[c] = acc;
        }
        __threadfence_system();
        __syncthreads();
        if (tid < G) st_release_sys(reinterpret_cast<unsigned*>(reg.base[tid] + off_b) + j, epoch);
    }

    // ---- phase C: the whole global row of this rank has arrived ------------------------------------------------------------------
    for (long long j = (long long)blockIdx.x * kThreads + tid; j < nchunks; j += (long long)gridDim.x * kThreads)
        wait_flag(reinterpret_cast<const unsigned*>(mine + off_b) + j, epoch);
}

}  // namespace

extern "C" long long flb_p2p_region_layout(long long ld, int world, int chunk, flb_p2p_layout* out) {
    if (ld < 4 || ld % 4 || world < 1 || world > FLB_P2P_MAX_RANKS || chunk < 4 || chunk % 4 || !out) return -1;
    const long long nchunks = (ld + chunk - 1) / chunk;
    auto align = [](long long v) { return (v + 255) & ~255ll; };
    out->off_flags_a = 0;
    out->off_flags_b = align(nchunks * world * 4);
    out->off_inbox = out->off_flags_b + align(nchunks * 4);
    out->off_global = out->off_inbox + align((long long)world * ld * 4);
    out->bytes = out->off_global + align(ld * 4);
    out->ld = ld; out->world = world; out->chunk = chunk;
    return out->bytes;
}

extern "C" int flb_p2p_alloc(void** ptr, long long bytes) {
    FLB_CHECK_ARG(ptr && bytes > 0, "flb_p2p_alloc: bad arguments");
    FLB_CUDA(cudaMalloc(ptr, (size_t)bytes));
    FLB_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
    FLB_CUDA(cudaDeviceSynchronize());
    return FLB_OK;
}
extern "C" int flb_p2p_free(void* ptr) {
    if (ptr) FLB_CUDA(cudaFree(ptr));
    return FLB_OK;
}
extern "C" int flb_p2p_export(const void* ptr, unsigned char* handle64) {
    FLB_CHECK_ARG(ptr && handle64, "flb_p2p_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == FLB_P2P_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    FLB_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
    memcpy(handle64, &h, sizeof(h));
    return FLB_OK;
}
extern "C" int flb_p2p_open(const unsigned char* handle64, void** ptr) {
    FLB_CHECK_ARG(ptr && handle64, "flb_p2p_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    FLB_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return FLB_OK;
}
extern "C" int flb_p2p_close(void* ptr) {
    if (ptr) FLB_CUDA(cudaIpcCloseMemHandle(ptr));
    return FLB_OK;
}

extern "C" int flb_fedavg_allreduce_p2p(const float* theta, long long ld_theta, const float* w, int K, long long P,
                                        const flb_p2p_layout* lay, void* const* regions, int rank, unsigned int epoch, void* stream) {
    FLB_CHECK_ARG(theta && w && lay && regions, "flb_fedavg_allreduce_p2p: null pointer");
    FLB_CHECK_ARG(K >= 1 && P >= 1 && ld_theta >= P && ld_theta % 4 == 0 && ((uintptr_t)theta % 16) == 0,
                  "flb_fedavg_allreduce_p2p: need K >= 1, 1 <= P <= ld_theta, rows 16-byte aligned (K=%d P=%lld ld=%lld)", K, P, ld_theta);
    FLB_CHECK_ARG(lay->world >= 1 && lay->world <= FLB_P2P_MAX_RANKS && rank >= 0 && rank < lay->world, "flb_fedavg_allreduce_p2p: bad rank / world");
    const long long P4 = (P + 3) / 4;                       // whole float4 columns: the row pads (ld = ceil32(P)) ride along
    FLB_CHECK_ARG(P4 * 4 <= lay->ld && P4 * 4 <= ld_theta, "flb_fedavg_allreduce_p2p: region rows too short (ld %lld, P %lld)", lay->ld, P);
    FLB_CHECK_ARG(epoch != 0, "flb_fedavg_allreduce_p2p: epoch must be >= 1 and increase with every call");
    Regions reg;
    for (int i = 0; i < lay->world; ++i) {
        FLB_CHECK_ARG(regions[i] != nullptr, "flb_fedavg_allreduce_p2p: region %d not mapped", i);
        reg.base[i] = (char*)regions[i];
    }
    const int chunk4 = lay->chunk / 4;
    const long long nchunks = (P4 + chunk4 - 1) / chunk4;
    static const int resident = flb_resident_ctas(fedavg_allreduce_p2p_kernel, kThreads);
    const int grid = (int)(nchunks < resident ? nchunks : resident);
    fedavg_allreduce_p2p_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(theta, ld_theta, w, K, P4, reg, rank, lay->world, lay->ld,
                                                                           lay->off_flags_a, lay->off_flags_b, lay->off_inbox,
                                                                           lay->off_global, chunk4, epoch);
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}
