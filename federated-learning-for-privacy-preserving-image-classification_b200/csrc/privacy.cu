// Update-level differential privacy, batched over K clients in two launches (HBM-bound).
//
// Replaces, per client, the reference chain
//   delta = w_local - w_global                      src/client/federated_trainer.py:437-443
//   GradientClipper.clip_gradients(delta)           src/shared/privacy.py:107-144
//   GaussianNoiseGenerator.add_noise_to_gradients   src/shared/privacy.py:221-254 (sigma rule :209)
//   w_upload = w_global + noisy_delta               src/client/federated_trainer.py:454-459
// which costs ~3 passes + one host sync per tensor.  Here: pass 1 reads local+global once and
// reduces sum(delta^2) per client (warp shuffles -> one double atomic per CTA); pass 2 re-reads them
// (L2-resident at these sizes), derives norm / clip coefficient / sigma ON THE DEVICE (no host sync),
// draws the Gaussian noise from Philox4x32-10 in registers and writes the upload weights once.
// Elementwise arithmetic keeps the reference's rounding sequence (fp32 mul by fp32(coef), fp32 mul
// fp32(sigma)*z, two fp32 adds, no FMA contraction) so with an injected z the result is bit-exact
// whenever the clip coefficient agrees.
#include "flb_common.cuh"
#include "philox.cuh"
#include "../../include/flb.h"

namespace {

constexpr int kThreads = 256;

template <bool VEC>
__global__ void __launch_bounds__(kThreads)
dp_sumsq_kernel(const float* __restrict__ local, long long ld, const float* __restrict__ global_w,
                double* __restrict__ norm2, long long P) {
    const int k = blockIdx.y;
    const float* __restrict__ row = local + (long long)k * ld;
    const long long P4 = (P + 3) >> 2;
    double acc = 0.0;
    for (long long c = (long long)blockIdx.x * kThreads + threadIdx.x; c < P4; c += (long long)gridDim.x * kThreads) {
        const long long p = c << 2;
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        if (VEC && p + 3 < P) {
            const float4 a = *reinterpret_cast<const float4*>(row + p);
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (global_w) g = __ldg(reinterpret_cast<const float4*>(global_w + p));
            d[0] = __fsub_rn(a.x, g.x); d[1] = __fsub_rn(a.y, g.y); d[2] = __fsub_rn(a.z, g.z); d[3] = __fsub_rn(a.w, g.w);
        } else {
            for (int e = 0; e < 4 && p + e < P; ++e) d[e] = __fsub_rn(row[p + e], global_w ? global_w[p + e] : 0.f);
        }
        // a product of two floats is exact in double
        acc += (double)d[0] * d[0] + (double)d[1] * d[1] + (double)d[2] * d[2] + (double)d[3] * d[3];
    }
    acc = flb_warp_sum_d(acc);
    __shared__ double s[kThreads / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < kThreads / 32 ? s[threadIdx.x] : 0.0;
        v = flb_warp_sum_d(v);
        if (threadIdx.x == 0) atomicAdd(&norm2[k], v);
    }
}

// ABSMAX (update codec, src/shared/compression.py:203-210): the quantiser's per-layer max|x| of the upload is a by-product
// of this pass -- the bits of |x| are a monotone integer key for non-negative floats, reduced per thread while the layer
// stays the same, then per CTA in shared memory, then with one atomicMax per (CTA, layer).  Saves the quantiser's own
// reduction pass over the fp32 upload (4 B per parameter).
constexpr int kMaxAbsLayers = 64;

template <bool VEC, bool ABSMAX>
__global__ void __launch_bounds__(kThreads)
dp_clip_noise_kernel(const float* __restrict__ local, long long ld, const float* __restrict__ global_w,
                     const float* __restrict__ z_in, const double* __restrict__ norm2,
                     float* __restrict__ out, float* __restrict__ norms_out,
                     double max_norm, double sigma_unit, unsigned long long seed,
                     unsigned long long stream_base, unsigned long long stream_stride, long long P,
                     const long long* __restrict__ seg_off, int L, unsigned int* __restrict__ absmax_bits) {
    __shared__ unsigned int s_max[ABSMAX ? kMaxAbsLayers : 1];
    __shared__ long long s_off[ABSMAX ? kMaxAbsLayers + 1 : 1];
    if (ABSMAX) {
        for (int i = threadIdx.x; i <= L; i += kThreads) s_off[i] = seg_off[i];
        for (int i = threadIdx.x; i < L; i += kThreads) s_max[i] = 0u;
        __syncthreads();
    }
    int cur_l = 0;                       // layer of the running maximum
    unsigned int cur_max = 0u;
    const int k = blockIdx.y;
    const double n = sqrt(norm2[k]);
    const float coef = n > max_norm ? (float)(max_norm / n) : 1.0f;        // privacy.py:127-131
    const double sens = n < max_norm ? n : max_norm;                          // privacy.py:140
    const float sigma = (float)(sens * sigma_unit);                           // privacy.py:209
    const bool noisy = sigma_unit != 0.0;
    if (norms_out && blockIdx.x == 0 && threadIdx.x == 0) norms_out[k] = (float)n;
    const float* __restrict__ row = local + (long long)k * ld;
    const float* __restrict__ zrow = z_in ? z_in + (long long)k * ld : nullptr;
    float* __restrict__ orow = out + (long long)k * ld;
    const long long P4 = (P + 3) >> 2;
    for (long long c = (long long)blockIdx.x * kThreads + threadIdx.x; c < P4; c += (long long)gridDim.x * kThreads) {
        const long long p = c << 2;
        const bool whole = VEC && p + 3 < P;
        float a[4] = {0.f, 0.f, 0.f, 0.f}, g[4] = {0.f, 0.f, 0.f, 0.f}, z[4] = {0.f, 0.f, 0.f, 0.f};
        if (whole) {
            const float4 av = *reinterpret_cast<const float4*>(row + p);
            a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
            if (global_w) { const float4 gv = __ldg(reinterpret_cast<const float4*>(global_w + p)); g[0] = gv.x; g[1] = gv.y; g[2] = gv.z; g[3] = gv.w; }
            if (zrow) { const float4 zv = __ldcs(reinterpret_cast<const float4*>(zrow + p)); z[0] = zv.x; z[1] = zv.y; z[2] = zv.z; z[3] = zv.w; }
        } else {
            for (int e = 0; e < 4 && p + e < P; ++e) {
                a[e] = row[p + e];
                if (global_w) g[e] = global_w[p + e];
                if (zrow) z[e] = zrow[p + e];
            }
        }
        if (noisy && !zrow) {
            const float4 zz = flb_normal4(seed, stream_base + stream_stride * (unsigned long long)k, (unsigned long long)c);
            z[0] = zz.x; z[1] = zz.y; z[2] = zz.z; z[3] = zz.w;
        }
        float r[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float d = global_w ? __fsub_rn(a[e], g[e]) : a[e];
            d = __fmul_rn(d, coef);
            if (noisy) d = __fadd_rn(d, __fmul_rn(sigma, z[e]));
            r[e] = global_w ? __fadd_rn(g[e], d) : d;
        }
        if (whole) {
            *reinterpret_cast<float4*>(orow + p) = make_float4(r[0], r[1], r[2], r[3]);
        } else {
            for (int e = 0; e < 4 && p + e < P; ++e) orow[p + e] = r[e];
        }
        if (ABSMAX) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const long long pe = p + e;
                if (pe >= P) break;
                if (pe < s_off[cur_l] || pe >= s_off[cur_l + 1]) {           // another layer: flush, then find it
                    if (cur_max) atomicMax(&s_max[cur_l], cur_max);
                    cur_max = 0u;
                    int lo = 0, hi = L;
                    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_off[mid] <= pe) lo = mid; else hi = mid; }
                    cur_l = lo;
                }
                const unsigned int bits = __float_as_uint(fabsf(r[e]));
                if (bits <= 0x7f800000u) cur_max = max(cur_max, bits);      // NaN is ignored, like torch.max of abs() would propagate
                                                                            // it only into a scale that the validator rejects anyway
            }
        }
    }
    if (ABSMAX) {
        if (cur_max) atomicMax(&s_max[cur_l], cur_max);
        __syncthreads();
        for (int i = threadIdx.x; i < L; i += kThreads)
            if (s_max[i]) atomicMax(&absmax_bits[(long long)k * L + i], s_max[i]);
    }
}

// out = x + sigma * z  (GaussianNoiseGenerator.add_noise_to_gradients on its own, privacy.py:221-254)
__global__ void __launch_bounds__(kThreads)
dp_add_noise_kernel(const float* __restrict__ x, long long ld, const float* __restrict__ z_in, float* __restrict__ out,
                    float sigma, unsigned long long seed, unsigned long long stream_base, long long P) {
    const int k = blockIdx.y;
    const float* __restrict__ row = x + (long long)k * ld;
    const float* __restrict__ zrow = z_in ? z_in + (long long)k * ld : nullptr;
    float* __restrict__ orow = out + (long long)k * ld;
    const long long P4 = (P + 3) >> 2;
    for (long long c = (long long)blockIdx.x * kThreads + threadIdx.x; c < P4; c += (long long)gridDim.x * kThreads) {
        const long long p = c << 2;
        float z[4];
        if (zrow) { for (int e = 0; e < 4; ++e) z[e] = p + e < P ? zrow[p + e] : 0.f; }
        else { const float4 zz = flb_normal4(seed, stream_base + (unsigned long long)k, (unsigned long long)c); z[0] = zz.x; z[1] = zz.y; z[2] = zz.z; z[3] = zz.w; }
        for (int e = 0; e < 4 && p + e < P; ++e) orow[p + e] = __fadd_rn(row[p + e], __fmul_rn(sigma, z[e]));
    }
}

__global__ void philox_normal_kernel(float* __restrict__ out, long long n, unsigned long long seed,
                                     unsigned long long stream) {
    const long long nb = (n + 3) >> 2;
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (long long)gridDim.x * blockDim.x) {
        const float4 z = flb_normal4(seed, stream, (unsigned long long)b);
        const float v[4] = {z.x, z.y, z.z, z.w};
        for (int e = 0; e < 4 && 4 * b + e < n; ++e) out[4 * b + e] = v[e];
    }
}

__global__ void philox_raw_kernel(uint32_t* __restrict__ out, long long nblocks, unsigned long long seed,
                                  unsigned long long stream, unsigned long long first_block) {
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nblocks; b += (long long)gridDim.x * blockDim.x) {
        const flb_u4 r = flb_philox_block(seed, stream, first_block + (unsigned long long)b);
        out[4 * b] = r.x; out[4 * b + 1] = r.y; out[4 * b + 2] = r.z; out[4 * b + 3] = r.w;
    }
}

// grid.x of a (blocks, K) launch of a grid-stride kernel: whole waves of its resident CTA count (flb_resident_ctas) --
// one wave when the clients fit, else two; never a fraction of a wave beyond
int blocks_per_client(long long P, int K, int resident) {
    const long long want = (P / 4 + kThreads - 1) / kThreads;
    long long cap = resident / K;
    if (cap < 1) cap = 1;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace

extern "C" int flb_dp_sumsq(const float* local, long long ld, const float* global_w, double* norm2,
                            int K, long long P, void* stream) {
    FLB_CHECK_ARG(local && norm2, "flb_dp_sumsq: null pointer");
    FLB_CHECK_ARG(K >= 1 && K <= 65535 && P >= 0 && ld >= P, "flb_dp_sumsq: need 1 <= K <= 65535, ld >= P");
    cudaStream_t st = (cudaStream_t)stream;
    FLB_CUDA(cudaMemsetAsync(norm2, 0, sizeof(double) * K, st));
    if (P == 0) return FLB_OK;
    const bool vec = (ld % 4 == 0) && ((uintptr_t)local % 16 == 0) && (!global_w || (uintptr_t)global_w % 16 == 0);
    static const int resident = flb_resident_ctas(dp_sumsq_kernel<true>, kThreads);
    dim3 grid(blocks_per_client(P, K, resident), K);
    if (vec) dp_sumsq_kernel<true><<<grid, kThreads, 0, st>>>(local, ld, global_w, norm2, P);
    else dp_sumsq_kernel<false><<<grid, kThreads, 0, st>>>(local, ld, global_w, norm2, P);
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

extern "C" int flb_dp_clip_noise(const float* local, long long ld, const float* global_w, const float* z_in,
                                 const double* norm2, float* out, float* norms_out, double max_norm,
                                 double sigma_unit, unsigned long long seed, unsigned long long stream_base,
                                 unsigned long long stream_stride, int K, long long P, void* stream) {
    FLB_CHECK_ARG(local && norm2 && out, "flb_dp_clip_noise: null pointer");
    FLB_CHECK_ARG(K >= 1 && K <= 65535 && P >= 0 && ld >= P, "flb_dp_clip_noise: need 1 <= K <= 65535, ld >= P");
    FLB_CHECK_ARG(max_norm > 0.0 && sigma_unit >= 0.0, "flb_dp_clip_noise: need max_norm > 0 and sigma_unit >= 0");
    if (P == 0) return FLB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (ld % 4 == 0) && ((uintptr_t)local % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
                     (!global_w || (uintptr_t)global_w % 16 == 0) && (!z_in || (uintptr_t)z_in % 16 == 0);
    static const int resident = flb_resident_ctas(dp_clip_noise_kernel<true, false>, kThreads);
    dim3 grid(blocks_per_client(P, K, resident), K);
    if (vec) dp_clip_noise_kernel<true, false><<<grid, kThreads, 0, st>>>(local, ld, global_w, z_in, norm2, out, norms_out, max_norm, sigma_unit, seed, stream_base, stream_stride, P, nullptr, 0, nullptr);
    else dp_clip_noise_kernel<false, false><<<grid, kThreads, 0, st>>>(local, ld, global_w, z_in, norm2, out, norms_out, max_norm, sigma_unit, seed, stream_base, stream_stride, P, nullptr, 0, nullptr);
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

// Same pass, and absmax_bits[k * L + l] = bits of max |out[k, seg_off[l] .. seg_off[l+1])| (a monotone key for non-negative
// floats; zeroed here): the input of flb_q8_quantize_absmax.
extern "C" int flb_dp_clip_noise_absmax(const float* local, long long ld, const float* global_w, const float* z_in,
                                        const double* norm2, float* out, float* norms_out, double max_norm,
                                        double sigma_unit, unsigned long long seed, unsigned long long stream_base,
                                        unsigned long long stream_stride, const long long* seg_off, int L,
                                        unsigned int* absmax_bits, int K, long long P, void* stream) {
    FLB_CHECK_ARG(local && norm2 && out && seg_off && absmax_bits, "flb_dp_clip_noise_absmax: null pointer");
    FLB_CHECK_ARG(K >= 1 && K <= 65535 && P >= 0 && ld >= P, "flb_dp_clip_noise_absmax: need 1 <= K <= 65535, ld >= P");
    FLB_CHECK_ARG(L >= 1 && L <= kMaxAbsLayers, "flb_dp_clip_noise_absmax: 1 <= L <= %d layers", kMaxAbsLayers);
    FLB_CHECK_ARG(max_norm > 0.0 && sigma_unit >= 0.0, "flb_dp_clip_noise_absmax: need max_norm > 0 and sigma_unit >= 0");
    cudaStream_t st = (cudaStream_t)stream;
    FLB_CUDA(cudaMemsetAsync(absmax_bits, 0, sizeof(unsigned int) * (size_t)K * L, st));
    if (P == 0) return FLB_OK;
    const bool vec = (ld % 4 == 0) && ((uintptr_t)local % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
                     (!global_w || (uintptr_t)global_w % 16 == 0) && (!z_in || (uintptr_t)z_in % 16 == 0);
    static const int resident = flb_resident_ctas(dp_clip_noise_kernel<true, true>, kThreads);
    dim3 grid(blocks_per_client(P, K, resident), K);
    if (vec) dp_clip_noise_kernel<true, true><<<grid, kThreads, 0, st>>>(local, ld, global_w, z_in, norm2, out, norms_out, max_norm, sigma_unit, seed, stream_base, stream_stride, P, seg_off, L, absmax_bits);
    else dp_clip_noise_kernel<false, true><<<grid, kThreads, 0, st>>>(local, ld, global_w, z_in, norm2, out, norms_out, max_norm, sigma_unit, seed, stream_base, stream_stride, P, seg_off, L, absmax_bits);
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

extern "C" int flb_dp_add_noise(const float* x, long long ld, const float* z_in, float* out, double sigma,
                                unsigned long long seed, unsigned long long stream_base, int K, long long P, void* stream) {
    FLB_CHECK_ARG(x && out, "flb_dp_add_noise: null pointer");
    FLB_CHECK_ARG(K >= 1 && K <= 65535 && P >= 0 && ld >= P && sigma >= 0.0, "flb_dp_add_noise: bad K/P/ld/sigma");
    if (P == 0) return FLB_OK;
    static const int resident = flb_resident_ctas(dp_add_noise_kernel, kThreads);
    dim3 grid(blocks_per_client(P, K, resident), K);
    dp_add_noise_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(x, ld, z_in, out, (float)sigma, seed, stream_base, P);
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

extern "C" int flb_philox_normal(float* out, long long n, unsigned long long seed, unsigned long long stream_id,
                                 void* stream) {
    FLB_CHECK_ARG(out && n >= 0, "flb_philox_normal: bad arguments");
    if (n == 0) return FLB_OK;
    const long long nb = (n + 3) / 4;
    long long blocks = (nb + 255) / 256;
    const long long cap = (long long)flb_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    philox_normal_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(out, n, seed, stream_id);
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}

extern "C" int flb_philox_raw(uint32_t* out, long long nblocks, unsigned long long seed, unsigned long long stream_id,
                              unsigned long long first_block, void* stream) {
    FLB_CHECK_ARG(out && nblocks >= 0, "flb_philox_raw: bad arguments");
    if (nblocks == 0) return FLB_OK;
    long long blocks = (nblocks + 255) / 256;
    const long long cap = (long long)flb_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    philox_raw_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(out, nblocks, seed, stream_id, first_block);
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}
