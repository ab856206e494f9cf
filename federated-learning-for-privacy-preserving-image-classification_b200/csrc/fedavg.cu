// FedAvg weighted aggregation kernels (HBM-bound; one read of every client row, one write).
//
// Replaces the K x L python-level axpy loop of FedAvgAggregator._weighted_average
// (reference src/aggregation/fedavg.py:267-289).  The arithmetic is kept identical to that
// loop -- out = 0; out = out + fp32(w_k) * theta_k for k = 0..K-1, fp32 multiply and fp32 add
// rounded separately (no FMA contraction) -- so the result is BIT-EXACT with the reference on
// the same inputs; only the loop nest is turned inside out (each thread owns 4 columns and walks
// the client axis, so every byte of theta is read exactly once, fully coalesced, 16 B per lane).
#include "flb_common.cuh"
#include "../../include/flb.h"

namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 8;      // client rows in flight per thread (8 x 16 B loads outstanding)
constexpr int kWChunk = 1024;   // weights staged in shared memory per pass

// theta: [K, ld] fp32 row-major, rows 16 B aligned (ld % 4 == 0).  out: [P].
__global__ void __launch_bounds__(kThreads)
fedavg_flat_vec4_kernel(const float* __restrict__ theta, long long ld, const float* __restrict__ w,
                        float* __restrict__ out, int K, long long P, int accumulate) {
    // P4 = float4 columns, the last one possibly partial (P % 4 != 0): its loads stay inside the row (ld % 4 == 0, ld >= P),
    // its stores are scalar -- no second launch for a two-parameter tail (SimpleCNN: P = 421 642)
    const long long P4 = (P + 3) >> 2;
    __shared__ float sw[kWChunk];
    const long long stride = (long long)gridDim.x * kThreads;
    const float4* __restrict__ t4 = reinterpret_cast<const float4*>(theta);
    const long long ld4 = ld >> 2;
    for (long long base = (long long)blockIdx.x * kThreads; base < P4; base += stride) {
        const long long c = base + threadIdx.x;
        const bool live = c < P4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int nvalid = live ? (int)min(4ll, P - 4 * c) : 0;
        if (accumulate && nvalid == 4) acc = reinterpret_cast<const float4*>(out)[c];
        else if (accumulate && nvalid > 0) {
            acc.x = out[4 * c];
            if (nvalid > 1) acc.y = out[4 * c + 1];
            if (nvalid > 2) acc.z = out[4 * c + 2];
        }
        for (int k0 = 0; k0 < K; k0 += kWChunk) {
            const int kc = min(kWChunk, K - k0);
            __syncthreads();
            for (int i = threadIdx.x; i < kc; i += kThreads) sw[i] = w[k0 + i];
            __syncthreads();
            if (!live) continue;
            int k = 0;
            for (; k + kUnroll <= kc; k += kUnroll) {
                float4 v[kUnroll];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) v[u] = __ldcs(&t4[(long long)(k0 + k + u) * ld4 + c]);
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const float wk = sw[k + u];
                    acc.x = __fadd_rn(acc.x, __fmul_rn(wk, v[u].x));
                    acc.y = __fadd_rn(acc.y, __fmul_rn(wk, v[u].y));
                    acc.z = __fadd_rn(acc.z, __fmul_rn(wk, v[u].z));
                    acc.w = __fadd_rn(acc.w, __fmul_rn(wk, v[u].w));
                }
            }
            for (; k < kc; ++k) {
                const float4 v = __ldcs(&t4[(long long)(k0 + k) * ld4 + c]);
                const float wk = sw[k];
                acc.x = __fadd_rn(acc.x, __fmul_rn(wk, v.x));
                acc.y = __fadd_rn(acc.y, __fmul_rn(wk, v.y));
                acc.z = __fadd_rn(acc.z, __fmul_rn(wk, v.z));
                acc.w = __fadd_rn(acc.w, __fmul_rn(wk, v.w));
            }
        }
        if (nvalid == 4) reinterpret_cast<float4*>(out)[c] = acc;
        else if (nvalid > 0) {
            out[4 * c] = acc.x;
            if (nvalid > 1) out[4 * c + 1] = acc.y;
            if (nvalid > 2) out[4 * c + 2] = acc.z;
        }
    }
}

// scalar columns [p0, P): tail of the vector path, or everything when rows are not 16 B aligned
__global__ void __launch_bounds__(kThreads)
fedavg_flat_scalar_kernel(const float* __restrict__ theta, long long ld, const float* __restrict__ w,
                          float* __restrict__ out, int K, long long p0, long long P, int accumulate) {
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long p = p0 + (long long)blockIdx.x * kThreads + threadIdx.x; p < P; p += stride) {
        float acc = accumulate ? out[p] : 0.f;
        for (int k = 0; k < K; ++k) acc = __fadd_rn(acc, __fmul_rn(__ldg(&w[k]), __ldcs(&theta[(long long)k * ld + p])));
        out[p] = acc;
    }
}

// Per-client, per-layer tensors that were never stacked: ptrs[k * L + l] is tensor l of client k,
// seg_off[l] .. seg_off[l+1] its span in the flat output.  Avoids a K x P gather copy in the
// drop-in aggregator when the updates already live on the device as separate tensors.
__global__ void __launch_bounds__(kThreads)
fedavg_ptrs_kernel(const float* const* __restrict__ ptrs, const long long* __restrict__ seg_off,
                   const float* __restrict__ w, float* __restrict__ out, int K, int L, long long P) {
    extern __shared__ long long s_off[];
    for (int i = threadIdx.x; i <= L; i += kThreads) s_off[i] = seg_off[i];
    __syncthreads();
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long p = (long long)blockIdx.x * kThreads + threadIdx.x; p < P; p += stride) {
        int lo = 0, hi = L;                       // largest l with s_off[l] <= p
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_off[mid] <= p) lo = mid; else hi = mid; }
        const long long j = p - s_off[lo];
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc = __fadd_rn(acc, __fmul_rn(__ldg(&w[k]), __ldcs(&ptrs[(long long)k * L + lo][j])));
        out[p] = acc;
    }
}

// uint8 affine-quantised client rows (reference src/shared/compression.py:230-244 dequant rule
// (float(q) - zp) * scale, per client and per layer), dequantised in registers and averaged in the
// same pass: K bytes + 4 bytes of traffic per parameter instead of 4K + 4.
// Each thread owns 16 consecutive parameters (one 16 B load per client row).  float(q) - zp is formed exactly in ONE
// add through the 2^23 mantissa trick (as_float(0x4B000000 | q) = 2^23 + q; both q and zp are integers below 2^24),
// the remaining fp32 multiply, multiply, add are rounded separately in the reference's order.
// The kernel is bound by the fp32 pipe (four dependent fp32 operations per BYTE of input, none of which may be fused or
// reassociated), so the four operations run as packed f32x2 instructions (sm_100 add.rn.f32x2 / mul.rn.f32x2: two IEEE
// results per issue slot, each lane rounded exactly like the scalar instruction), and the byte -> 2^23 + q expansion is one
// PRMT per element (byte e of the word under the constant's upper three bytes) instead of shift + and + or.
__device__ __forceinline__ void q8_accum4(uint32_t word, float magic_z, float s, float wk, float* acc) {
    const float2 q01 = make_float2(__uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540)),      // 2^23 + q, exact
                                   __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7541)));
    const float2 q23 = make_float2(__uint_as_float(__byte_perm(word, 0x4B000000u, 0x7542)),
                                   __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7543)));
    const float2 nz = make_float2(-magic_z, -magic_z), s2 = make_float2(s, s), w2 = make_float2(wk, wk);
    // (q - zp) exactly, * scale, then w * (.), then += : every step rounded separately, in the reference's order
    const float2 t01 = __fmul2_rn(w2, __fmul2_rn(__fadd2_rn(q01, nz), s2));
    const float2 t23 = __fmul2_rn(w2, __fmul2_rn(__fadd2_rn(q23, nz), s2));
    // the accumulation stays scalar: ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 whatever --fmad says,
    // which would round once where the reference rounds twice (caught by test_q8_fused_dequant_average)
    acc[0] = __fadd_rn(acc[0], t01.x); acc[1] = __fadd_rn(acc[1], t01.y);
    acc[2] = __fadd_rn(acc[2], t23.x); acc[3] = __fadd_rn(acc[3], t23.y);
}

// EPT = parameters per thread: 16 (one 16 B load per client row) for long rows; 4 (one 4 B load, a warp still reads whole
// 128 B lines) when P / 16 threads would leave most of the machine idle -- the client loop cannot be split, its adds
// must stay in client order to remain bit-exact.
// streaming 16-byte load that the compiler may not sink next to its first use (volatile asm statements keep their order):
// the U loads of an unrolled group are all issued before the first dequantisation starts
__device__ __forceinline__ uint4 ldcs_u4_pinned(const void* p) {
    uint4 v;
    asm volatile("ld.global.cs.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <int EPT>
__global__ void __launch_bounds__(kThreads, 3)
fedavg_q8_kernel(const uint8_t* __restrict__ q, long long ldq, const float* __restrict__ scale,
                 const float* __restrict__ zp, const long long* __restrict__ seg_off,
                 const float* __restrict__ w, float* __restrict__ out, int K, int L, long long P) {
    extern __shared__ long long s_off[];
    for (int i = threadIdx.x; i <= L; i += kThreads) s_off[i] = seg_off[i];
    __syncthreads();
    const long long PV = (P + EPT - 1) / EPT;
    const long long stride = (long long)gridDim.x * kThreads;
    const bool vec_ok = (ldq & (EPT - 1)) == 0 && ((uintptr_t)q & (EPT - 1)) == 0;
    for (long long c = (long long)blockIdx.x * kThreads + threadIdx.x; c < PV; c += stride) {
        const long long p = c * EPT;
        int lo = 0, hi = L;                       // largest l with s_off[l] <= p
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_off[mid] <= p) lo = mid; else hi = mid; }
        const bool whole = vec_ok && (p + EPT - 1 < P) && (p + EPT - 1 < s_off[lo + 1]);
        if (whole) {
            float acc[EPT];
#pragma unroll
            for (int e = 0; e < EPT; ++e) acc[e] = 0.f;
            if (EPT == 16) {
                // ncu (K = 100, P = 10 M): with one or two client rows in flight per thread the kernel idles on the long scoreboard
                // (11 stalled warps per issue, 46 % of the DRAM peak, fma pipe 49 % busy): latency-, not ALU-bound.
                // ptxas (12.9) sinks every load of a straight-line group behind the arithmetic of the previous row (the PTX has the
                // four loads first; the SASS has one load, 150 dependent instructions, the next load ...), so the prefetch is
                // carried across the loop back-edge instead: group g + 1 is requested before group g is dequantised.
                constexpr int U = 3;
                const int Kfull = K - K % U;
                uint4 cur[U], nxt[U];
                float csc[U], cmz[U], cwk[U], nsc[U], nmz[U], nwk[U];
                auto fetch = [&](int k0, uint4 (&v)[U], float (&sc)[U], float (&mz)[U], float (&wk)[U]) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        v[u] = ldcs_u4_pinned(q + (long long)(k0 + u) * ldq + p);
                        sc[u] = __ldg(&scale[(long long)(k0 + u) * L + lo]);
                        mz[u] = __ldg(&zp[(long long)(k0 + u) * L + lo]);
                        wk[u] = __ldg(&w[k0 + u]);
                    }
                };
                if (Kfull > 0) fetch(0, cur, csc, cmz, cwk);
                int k = 0;
                for (; k < Kfull; k += U) {
                    const bool more = k + U < Kfull;
                    if (more) fetch(k + U, nxt, nsc, nmz, nwk);
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const float magic_z = 8388608.0f + cmz[u];                        // exact: zp is an integer < 2^16
                        q8_accum4(cur[u].x, magic_z, csc[u], cwk[u], acc);
                        q8_accum4(cur[u].y, magic_z, csc[u], cwk[u], acc + 4);
                        q8_accum4(cur[u].z, magic_z, csc[u], cwk[u], acc + 8);
                        q8_accum4(cur[u].w, magic_z, csc[u], cwk[u], acc + 12);
                    }
                    if (more) {
#pragma unroll
                        for (int u = 0; u < U; ++u) { cur[u] = nxt[u]; csc[u] = nsc[u]; cmz[u] = nmz[u]; cwk[u] = nwk[u]; }
                    }
                }
                for (; k < K; ++k) {
                    const float s = __ldg(&scale[(long long)k * L + lo]);
                    const float magic_z = 8388608.0f + __ldg(&zp[(long long)k * L + lo]);
                    const float wk1 = __ldg(&w[k]);
                    const uint4 v = __ldcs(reinterpret_cast<const uint4*>(q + (long long)k * ldq + p));
                    q8_accum4(v.x, magic_z, s, wk1, acc);
                    q8_accum4(v.y, magic_z, s, wk1, acc + 4);
                    q8_accum4(v.z, magic_z, s, wk1, acc + 8);
                    q8_accum4(v.w, magic_z, s, wk1, acc + 12);
                }
            } else {
#pragma unroll 8
                for (int k = 0; k < K; ++k) {
                    const float s = __ldg(&scale[(long long)k * L + lo]);
                    const float magic_z = 8388608.0f + __ldg(&zp[(long long)k * L + lo]);     // exact: zp is an integer < 2^16
                    const float wk = __ldg(&w[k]);
                    q8_accum4(__ldcs(reinterpret_cast<const uint32_t*>(q + (long long)k * ldq + p)), magic_z, s, wk, acc);
                }
            }
            float4* o4 = reinterpret_cast<float4*>(out + p);
            if (((uintptr_t)out & 15) == 0) {
#pragma unroll
                for (int e = 0; e < EPT / 4; ++e) o4[e] = make_float4(acc[4 * e], acc[4 * e + 1], acc[4 * e + 2], acc[4 * e + 3]);
            } else {
#pragma unroll
                for (int e = 0; e < EPT; ++e) out[p + e] = acc[e];
            }
        } else {
            for (long long e = p; e < min(p + EPT, P); ++e) {
                int l = lo;
                while (e >= s_off[l + 1]) ++l;
                float a = 0.f;
                for (int k = 0; k < K; ++k) {
                    const float s = __ldg(&scale[(long long)k * L + l]), z = __ldg(&zp[(long long)k * L + l]);
                    a = __fadd_rn(a, __fmul_rn(__ldg(&w[k]), __fmul_rn(__fsub_rn((float)q[(long long)k * ldq + e], z), s)));
                }
                out[e] = a;
            }
        }
    }
}

int grid_for(long long work_items) {
    const long long blocks = (work_items + kThreads - 1) / kThreads;
    const long long cap = (long long)flb_num_sms() * 16;     // 16 resident CTAs of 256 threads per SM at most
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

extern "C" int flb_fedavg_weighted_sum(const float* theta, long long ld, const float* w, float* out,
                                       int K, long long P, int accumulate, void* stream) {
    FLB_CHECK_ARG(theta && w && out, "flb_fedavg_weighted_sum: null pointer");
    FLB_CHECK_ARG(K >= 1 && P >= 0 && ld >= P, "flb_fedavg_weighted_sum: need K >= 1, P >= 0, ld >= P (K=%d P=%lld ld=%lld)", K, P, ld);
    if (P == 0) return FLB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (ld % 4 == 0) && ((uintptr_t)theta % 16 == 0) && ((uintptr_t)out % 16 == 0);
    if (vec) fedavg_flat_vec4_kernel<<<grid_for((P + 3) >> 2), kThreads, 0, st>>>(theta, ld, w, out, K, P, accumulate);
    else fedavg_flat_scalar_kernel<<<grid_for(P), kThreads, 0, st>>>(theta, ld, w, out, K, 0, P, accumulate);
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

extern "C" int flb_fedavg_weighted_sum_ptrs(const float* const* ptrs, const long long* seg_off, const float* w,
                                            float* out, int K, int L, long long P, void* stream) {
    FLB_CHECK_ARG(ptrs && seg_off && w && out, "flb_fedavg_weighted_sum_ptrs: null pointer");
    FLB_CHECK_ARG(K >= 1 && L >= 1 && L <= 4096, "flb_fedavg_weighted_sum_ptrs: need K >= 1 and 1 <= L <= 4096");
    if (P == 0) return FLB_OK;
    fedavg_ptrs_kernel<<<grid_for(P), kThreads, (L + 1) * sizeof(long long), (cudaStream_t)stream>>>(ptrs, seg_off, w, out, K, L, P);
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

extern "C" int flb_fedavg_weighted_sum_q8(const uint8_t* q, long long ldq, const float* scale, const float* zp,
                                          const long long* seg_off, const float* w, float* out,
                                          int K, int L, long long P, void* stream) {
    FLB_CHECK_ARG(q && scale && zp && seg_off && w && out, "flb_fedavg_weighted_sum_q8: null pointer");
    FLB_CHECK_ARG(K >= 1 && L >= 1 && L <= 4096 && ldq >= P, "flb_fedavg_weighted_sum_q8: bad K/L/ldq");
    if (P == 0) return FLB_OK;
    const size_t smem = (L + 1) * sizeof(long long);
    if ((P + 15) / 16 >= (long long)flb_num_sms() * 1536)        // enough 16-parameter threads to fill the machine
        fedavg_q8_kernel<16><<<grid_for((P + 15) / 16), kThreads, smem, (cudaStream_t)stream>>>(q, ldq, scale, zp, seg_off, w, out, K, L, P);
    else
        fedavg_q8_kernel<4><<<grid_for((P + 3) / 4), kThreads, smem, (cudaStream_t)stream>>>(q, ldq, scale, zp, seg_off, w, out, K, L, P);
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}
