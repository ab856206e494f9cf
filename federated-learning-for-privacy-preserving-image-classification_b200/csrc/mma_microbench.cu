// Profiling aid: cycles per tcgen05.mma kind::tf32 instruction as a function of the shapes and operand placements the
// training kernels use.  One CTA, operands in shared memory (contents irrelevant), `reps` back-to-back 128 x N x 8 MMAs
// issued by one elected lane into one TMEM accumulator, timed with clock64() from the first issue to the completion of the
// commit.  Answers the round-1 question "why is the tensor pipe only 15-30 % active in the N = 32 / 64 convolutions":
//   * n         output columns per MMA (32, 64, 128, 256)
//   * a_shift   the A operand starts `a_shift` 128-byte rows into its 1024-byte swizzle atom (the halo convolutions address
//               their nine taps as row-shifted windows of one staged box, train_tc.cu ConvFwdHaloT / ConvDgradHaloT)
//   * a_mn/b_mn operand major-ness (MN-major tf32 operands use the 32-byte-atom swizzle)
//   * rotate    number of distinct A windows cycled through (1 = the same 4 KB every time, 9 = nine taps of a box)
#include "tc_gemm.cuh"

namespace tc {
namespace {

struct MicroParams { int n, a_shift, a_mn, b_mn, rotate, reps; long long* cycles; };

__global__ void __launch_bounds__(128, 1) mma_microbench_kernel(MicroParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    // A region: 64 KB (room for shifted / rotated windows), B region: 32 KB after it; zero-filled so that no NaN pattern slows
    // or poisons anything
    for (int i = threadIdx.x; i < (96 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&done_bar, 1); fence_barrier_init(); }
    if (warp == 1) tmem_alloc<256>(&tmem_base);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (warp == 0) {
        const bool lead = elect_one();
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 64 * 1024;
        const uint32_t id = (1u << 4) | (2u << 7) | (2u << 10) | ((p.a_mn ? 1u : 0u) << 15) | ((p.b_mn ? 1u : 0u) << 16) |
                            ((uint32_t)(p.n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        long long t0 = 0;
        for (int pass = 0; pass < 2; ++pass) {                 // pass 0 warms the pipe, pass 1 is timed
            __syncwarp();
            if (pass == 1) t0 = clock64();
            for (int r = 0; r < p.reps; ++r) {
                const int win = p.rotate > 1 ? r % p.rotate : 0;
                const uint32_t a_addr = a0 + (uint32_t)(p.a_shift + win * (p.a_shift ? 17 : 16)) * 128u + (r & 3) * 32;
                const uint64_t ad = p.a_mn ? smem_desc_mn(a0 + (r & 3) * 1024, 4096, 512) : (p.a_shift ? smem_desc_row(a_addr) : smem_desc(a_addr, 16, 1024));
                const uint64_t bd = p.b_mn ? smem_desc_mn(b0 + (r & 3) * 1024, 4096, 512) : smem_desc(b0 + (r & 3) * 32, 16, 1024);
                if (lead) mma_tf32(tmem, ad, bd, id, r > 0);
            }
            if (lead) mma_commit(&done_bar);
            __syncwarp();
            mbar_wait(&done_bar, pass);
            tc_fence_after();
        }
        const long long t1 = clock64();
        if (lead) *p.cycles = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<256>(tmem);
}


// Variant with NO per-instruction address arithmetic: 32 MMAs per loop iteration, descriptors formed from compile-time
// offsets (what the fully unrolled tap loops of the training kernels compile to).  If this runs faster than the loop
// above, that loop (and any kernel like it) is bound by the ISSUING warp, not by the tensor core.
template <int M>
__global__ void __launch_bounds__(128, 1) mma_microbench_unrolled_kernel(MicroParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (96 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&done_bar, 1); fence_barrier_init(); }
    if (warp == 1) tmem_alloc<256>(&tmem_base);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (warp == 0) {
        const bool lead = elect_one();
        const uint32_t a0 = smem_u32(smem) + (uint32_t)p.a_shift * 128u, b0 = smem_u32(smem) + 64 * 1024;
        const uint32_t id = (1u << 4) | (2u << 7) | (2u << 10) | ((p.b_mn ? 1u : 0u) << 16) |
                            ((uint32_t)(p.n >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        long long t0 = 0;
        for (int pass = 0; pass < 2; ++pass) {
            __syncwarp();
            if (pass == 1) t0 = clock64();
            for (int r = 0; r < p.reps; r += 32) {
#pragma unroll
                for (int u = 0; u < 32; ++u) {
                    // eight "taps" (row shifts of 17 rows) x four k-steps, all offsets compile-time
                    const uint32_t a_addr = a0 + (uint32_t)((u >> 2) * 17) * 128u + (u & 3) * 32;
                    const uint64_t ad = smem_desc_row(a_addr);
                    const uint64_t bd = p.b_mn ? smem_desc_mn(b0 + (u & 3) * 1024, 4096, 512) : smem_desc(b0 + (u & 3) * 32, 16, 1024);
                    if (lead) mma_tf32(tmem, ad, bd, id, (r | u) > 0);
                }
            }
            if (lead) mma_commit(&done_bar);
            __syncwarp();
            mbar_wait(&done_bar, pass);
            tc_fence_after();
        }
        const long long t1 = clock64();
        if (lead) *p.cycles = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<256>(tmem);
}

}  // namespace
}  // namespace tc

// cycles_out (device, one int64) receives the SM cycles of `reps` MMAs including the final commit round trip; the host
// divides.  Synchronises nothing; profiling aid only.
extern "C" int flb_mma_microbench(int n, int a_shift, int a_mn, int b_mn, int rotate, int reps, long long* cycles_out, void* stream) {
    FLB_CHECK_ARG(n >= 16 && n <= 256 && n % 16 == 0, "flb_mma_microbench: n must be a multiple of 16 in 16..256");
    FLB_CHECK_ARG(a_shift >= 0 && a_shift < 64 && rotate >= 1 && rotate <= 9 && reps >= 1 && cycles_out, "flb_mma_microbench: bad arguments");
    tc::MicroParams p{n, a_shift, a_mn, b_mn, rotate, reps, cycles_out};
    const int smem = 97 * 1024 + 1024;
    if (rotate == 8 || rotate == 7) {          // unrolled variants: rotate 8 -> M = 128, rotate 7 -> M = 64 (reps rounded up to 32)
        p.reps = (reps + 31) / 32 * 32;
        if (rotate == 8) {
            FLB_CUDA(cudaFuncSetAttribute(tc::mma_microbench_unrolled_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            tc::mma_microbench_unrolled_kernel<128><<<1, 128, smem, (cudaStream_t)stream>>>(p);
        } else {
            FLB_CUDA(cudaFuncSetAttribute(tc::mma_microbench_unrolled_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            tc::mma_microbench_unrolled_kernel<64><<<1, 128, smem, (cudaStream_t)stream>>>(p);
        }
        FLB_LAUNCH_CHECK();
        return FLB_OK;
    }
    FLB_CUDA(cudaFuncSetAttribute(tc::mma_microbench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc::mma_microbench_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(p);
    FLB_LAUNCH_CHECK();
    return FLB_OK;
}
