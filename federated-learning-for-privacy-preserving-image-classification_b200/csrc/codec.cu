// Update codec on the device: per-client, per-layer affine quantisation (uint8) of the stacked
// client rows, and its inverse.  Follows the reference's QuantizationCompressor arithmetic
// (src/shared/compression.py:203-244):
//   symmetric : scale = 2*max|x| / (2^b - 1),  zp = (2^b - 1) // 2
//   asymmetric: scale = (max - min) / (2^b - 1), zp = -round(min / scale)
//   q  = clamp(round_half_even(x / fp32(scale) + fp32(zp)), 0, 2^b - 1)     (stored unpacked in uint8, b <= 8)
//   x' = (float(q) - zp) * fp32(scale)
// Pass 1 reduces max|x| (or min/max) per (client, layer) with warp shuffles + one atomic per warp on
// an order-preserving integer key; pass 2 derives scale/zp and writes the codes.  Both passes are
// batched over all K clients and L layers in one launch each (HBM-bound: 4 B read + 1 B written per
// parameter in pass 2; pass 1's read is what brings the row into L2).
#include "flb_common.cuh"
#include "../../include/flb.h"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t f2key(float f) {      // monotone float -> uint32
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void q8_init_kernel(uint32_t* __restrict__ kmin, uint32_t* __restrict__ kmax, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { kmin[i] = 0xffffffffu; kmax[i] = 0u; }
}

// grid: (chunks, L, K).  keys: kmax[k*L+l] = key(max x or max|x|), kmin = key(min x)
__global__ void __launch_bounds__(kThreads)
q8_minmax_kernel(const float* __restrict__ x, long long ld, const long long* __restrict__ seg_off,
                 uint32_t* __restrict__ kmin, uint32_t* __restrict__ kmax, int L, int symmetric) {
    const int l = blockIdx.y, k = blockIdx.z;
    const long long b = seg_off[l], e = seg_off[l + 1];
    const float* __restrict__ row = x + (long long)k * ld;
    float mx = symmetric ? 0.f : -INFINITY, mn = INFINITY;
    for (long long p = b + (long long)blockIdx.x * kThreads + threadIdx.x; p < e; p += (long long)gridDim.x * kThreads) {
        const float v = row[p];
        if (symmetric) mx = fmaxf(mx, fabsf(v));
        else { mx = fmaxf(mx, v); mn = fminf(mn, v); }
    }
    mx = flb_warp_max(mx);
    mn = -flb_warp_max(-mn);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&kmax[k * L + l], f2key(mx));
        if (!symmetric) atomicMin(&kmin[k * L + l], f2key(mn));
    }
}

__global__ void q8_params_kernel(const uint32_t* __restrict__ kmin, const uint32_t* __restrict__ kmax,
                                 float* __restrict__ scale, float* __restrict__ zp, int n, int levels, int symmetric) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (symmetric) {
        const double mx = (double)key2f(kmax[i]);
        scale[i] = (float)((2.0 * mx) / (double)(levels - 1));       // compression.py:207-210
        zp[i] = (float)((levels - 1) / 2);
    } else {
        const double mn = (double)key2f(kmin[i]), mx = (double)key2f(kmax[i]);
        const double s = (mx - mn) / (double)(levels - 1);            // compression.py:212-215
        scale[i] = (float)s;
        zp[i] = (float)(-rint(mn / s));
    }
}

// grid: (chunks, L, K)
__global__ void __launch_bounds__(kThreads)
q8_quantize_kernel(const float* __restrict__ x, long long ld, const long long* __restrict__ seg_off,
                   const float* __restrict__ scale, const float* __restrict__ zp,
                   uint8_t* __restrict__ q, long long ldq, int L, float qmax) {
    const int l = blockIdx.y, k = blockIdx.z;
    const long long b = seg_off[l], e = seg_off[l + 1];
    const float s = scale[k * L + l], z = zp[k * L + l];
    const float* __restrict__ row = x + (long long)k * ld;
    uint8_t* __restrict__ qrow = q + (long long)k * ldq;
    for (long long p = b + (long long)blockIdx.x * kThreads + threadIdx.x; p < e; p += (long long)gridDim.x * kThreads) {
        float v = __fadd_rn(__fdiv_rn(row[p], s), z);                 // compression.py:217
        v = fminf(fmaxf(rintf(v), 0.f), qmax);                        // round-half-even, clamp (:218)
        qrow[p] = (s > 0.f) ? (uint8_t)v : (uint8_t)z;                // all-zero layer: code = zero point
    }
}

// Same codes, 4 parameters per thread (16 B load, 4 B store) over the flat row; the layer of a quad comes from a binary
// search in the (shared-memory) offset table, quads that straddle a layer boundary fall back to per-element lookups.
// grid: (chunks, K)
__global__ void __launch_bounds__(kThreads)
q8_quantize_vec4_kernel(const float* __restrict__ x, long long ld, const long long* __restrict__ seg_off,
                        const float* __restrict__ scale, const float* __restrict__ zp,
                        uint8_t* __restrict__ q, long long ldq, int L, long long P, float qmax) {
    extern __shared__ long long s_off[];
    for (int i = threadIdx.x; i <= L; i += kThreads) s_off[i] = seg_off[i];
    __syncthreads();
    const int k = blockIdx.y;
    const float* __restrict__ row = x + (long long)k * ld;
    uint8_t* __restrict__ qrow = q + (long long)k * ldq;
    const float* __restrict__ sc = scale + (long long)k * L;
    const float* __restrict__ zz = zp + (long long)k * L;
    const long long P4 = (P + 3) >> 2;
    for (long long c = (long long)blockIdx.x * kThreads + threadIdx.x; c < P4; c += (long long)gridDim.x * kThreads) {
        const long long p = c << 2;
        int lo = 0, hi = L;                       // largest l with s_off[l] <= p
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_off[mid] <= p) lo = mid; else hi = mid; }
        if (p + 3 < P && p + 3 < s_off[lo + 1]) {
            const float s = sc[lo], z = zz[lo];
            const float4 v4 = *reinterpret_cast<const float4*>(row + p);
            const float v[4] = {v4.x, v4.y, v4.z, v4.w};
            uint32_t word = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float t = __fadd_rn(__fdiv_rn(v[e], s), z);                   // compression.py:217
                t = fminf(fmaxf(rintf(t), 0.f), qmax);                        // round-half-even, clamp (:218)
                word |= (uint32_t)((s > 0.f) ? (uint8_t)t : (uint8_t)z) << (8 * e);
            }
            *reinterpret_cast<uint32_t*>(qrow + p) = word;
        } else {
            for (long long e = p; e < min(p + 4, P); ++e) {
                int l = lo;
                while (e >= s_off[l + 1]) ++l;
                const float s = sc[l], z = zz[l];
                float t = __fadd_rn(__fdiv_rn(row[e], s), z);
                t = fminf(fmaxf(rintf(t), 0.f), qmax);
                qrow[e] = (s > 0.f) ? (uint8_t)t : (uint8_t)z;
            }
        }
    }
}

// symmetric scale / zero point from the bits of max|x| (flb_dp_clip_noise_absmax)
__global__ void q8_params_absmax_kernel(const unsigned int* __restrict__ absmax_bits, float* __restrict__ scale,
                                        float* __restrict__ zp, int n, int levels) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double mx = (double)__uint_as_float(absmax_bits[i]);
    scale[i] = (float)((2.0 * mx) / (double)(levels - 1));           // compression.py:207-210
    zp[i] = (float)((levels - 1) / 2);
}

__global__ void __launch_bounds__(kThreads)
q8_dequantize_kernel(const uint8_t* __restrict__ q, long long ldq, const long long* __restrict__ seg_off,
                     const float* __restrict__ scale, const float* __restrict__ zp,
                     float* __restrict__ out, long long ld, int L) {
    const int l = blockIdx.y, k = blockIdx.z;
    const long long b = seg_off[l], e = seg_off[l + 1];
    const float s = scale[k * L + l], z = zp[k * L + l];
    const uint8_t* __restrict__ qrow = q + (long long)k * ldq;
    float* __restrict__ orow = out + (long long)k * ld;
    for (long long p = b + (long long)blockIdx.x * kThreads + threadIdx.x; p < e; p += (long long)gridDim.x * kThreads)
        orow[p] = __fmul_rn(__fsub_rn((float)qrow[p], z), s);         // compression.py:238-240
}

int chunks_for(long long P, int L, int K) {
    long long c = (P / L + (long long)kThreads * 8 - 1) / ((long long)kThreads * 8);    // ~8 elements per thread on an average layer
    if (c < 1) c = 1;
    long long cap = ((long long)flb_num_sms() * 16) / ((long long)L * K) + 1;
    if (cap < 8) cap = 8;            // layer sizes are very uneven (fc1 holds 70-95 % of the parameters): keep the big ones parallel
    if (c > cap) c = cap;
    return (int)(c > 1024 ? 1024 : c);
}

// the code pass: 4 parameters per thread when rows allow 16 B / 4 B accesses, else the per-layer scalar kernel
int launch_quantize(const float* x, long long ld, const long long* seg_off, const float* scale, const float* zp, uint8_t* q,
                    long long ldq, int K, int L, long long P, int levels, cudaStream_t st) {
    const bool vec = (ld % 4 == 0) && (ldq % 4 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)q % 4 == 0);
    if (vec) {
        static const int resident = flb_resident_ctas(q8_quantize_vec4_kernel, kThreads);
        long long blocks = (P / 4 + kThreads - 1) / kThreads, cap = resident / K;
        if (cap < 1) cap = 1;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        q8_quantize_vec4_kernel<<<dim3((unsigned)blocks, K), kThreads, (L + 1) * sizeof(long long), st>>>(x, ld, seg_off, scale, zp, q, ldq, L, P, (float)(levels - 1));
    } else {
        dim3 grid(chunks_for(P, L, K), L, K);
        q8_quantize_kernel<<<grid, kThreads, 0, st>>>(x, ld, seg_off, scale, zp, q, ldq, L, (float)(levels - 1));
    }
    return FLB_OK;
}

}  // namespace

extern "C" int flb_q8_quantize(const float* x, long long ld, const long long* seg_off, uint8_t* q, long long ldq,
                               float* scale, float* zp, uint32_t* scratch, int K, int L, long long P,
                               int bits, int symmetric, void* stream) {
    FLB_CHECK_ARG(x && seg_off && q && scale && zp && scratch, "flb_q8_quantize: null pointer");
    FLB_CHECK_ARG(K >= 1 && K <= 65535 && L >= 1 && L <= 65535 && ld >= P && ldq >= P, "flb_q8_quantize: bad K/L/ld");
    FLB_CHECK_ARG(bits >= 1 && bits <= 8, "flb_q8_quantize: bits must be in 1..8 (codes are stored in uint8)");
    if (P == 0) return FLB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int n = K * L, levels = 1 << bits;
    uint32_t* kmin = scratch;
    uint32_t* kmax = scratch + n;
    q8_init_kernel<<<flb_cdiv(n, 256), 256, 0, st>>>(kmin, kmax, n);
    dim3 grid(chunks_for(P, L, K), L, K);
    q8_minmax_kernel<<<grid, kThreads, 0, st>>>(x, ld, seg_off, kmin, kmax, L, symmetric);
    q8_params_kernel<<<flb_cdiv(n, 256), 256, 0, st>>>(kmin, kmax, scale, zp, n, levels, symmetric);
    if (int rc = launch_quantize(x, ld, seg_off, scale, zp, q, ldq, K, L, P, levels, st)) return rc;
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

// Symmetric quantisation whose per-(client, layer) max|x| is already known (absmax_bits from flb_dp_clip_noise_absmax):
// one parameter kernel + the code pass -- the reduction pass over x is gone.
extern "C" int flb_q8_quantize_absmax(const float* x, long long ld, const long long* seg_off, const unsigned int* absmax_bits,
                                      uint8_t* q, long long ldq, float* scale, float* zp, int K, int L, long long P,
                                      int bits, void* stream) {
    FLB_CHECK_ARG(x && seg_off && absmax_bits && q && scale && zp, "flb_q8_quantize_absmax: null pointer");
    FLB_CHECK_ARG(K >= 1 && K <= 65535 && L >= 1 && L <= 65535 && ld >= P && ldq >= P, "flb_q8_quantize_absmax: bad K/L/ld");
    FLB_CHECK_ARG(bits >= 1 && bits <= 8, "flb_q8_quantize_absmax: bits must be in 1..8 (codes are stored in uint8)");
    if (P == 0) return FLB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int n = K * L, levels = 1 << bits;
    q8_params_absmax_kernel<<<flb_cdiv(n, 256), 256, 0, st>>>(absmax_bits, scale, zp, n, levels);
    if (int rc = launch_quantize(x, ld, seg_off, scale, zp, q, ldq, K, L, P, levels, st)) return rc;
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

extern "C" int flb_q8_dequantize(const uint8_t* q, long long ldq, const long long* seg_off, const float* scale,
                                 const float* zp, float* out, long long ld, int K, int L, long long P, void* stream) {
    FLB_CHECK_ARG(q && seg_off && scale && zp && out, "flb_q8_dequantize: null pointer");
    FLB_CHECK_ARG(K >= 1 && K <= 65535 && L >= 1 && L <= 65535 && ld >= P && ldq >= P, "flb_q8_dequantize: bad K/L/ld");
    if (P == 0) return FLB_OK;
    dim3 grid(chunks_for(P, L, K), L, K);
    q8_dequantize_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(q, ldq, seg_off, scale, zp, out, ld, L);
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}
