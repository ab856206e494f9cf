// SimpleCNN classifier block in ONE launch (tcgen05 TF32 path, reference src/shared/models_pytorch.py:91-96 forward,
// src/shared/training.py:193-196 loss + backward):
//
//     hpre = a2 . W1^T  ->  h = dropout(relu(hpre + b1))  ->  logits = h . W2^T + b2  ->  mean cross-entropy, dlogits
//          ->  dh = (dlogits . W2) * relu/dropout mask  ->  da2 = dh . W1
//
// Before this kernel the step ran fc1_fwd (split-K GEMM), head_fwd_bwd and fc1_dgrad as three dependent single-wave
// launches (8.3 + 9.8 + 7.3 us at 10 clients, each paying its own launch, TMEM / barrier setup and two memory round trips
// for ~0.25 GFLOP of math).  Here a client's fc1 is cut into S = 14 slices of 224 input features; CTA (slice s, group g)
//   1. TMA-loads its [128 out x 224 in] weight slice (K-major boxes) and the matching [32 x 224] activation slice, runs the
//      28 forward MMAs into TMEM and adds its partial hpre to the client's [32, 128] accumulator with fp32 reductions,
//   2. re-loads the same weight slice as MN-major boxes (straight from L2) for the dgrad while
//   3. it waits at the client's barrier (a global arrive counter: all 14 slices have added their partial),
//   4. computes the whole classifier head REDUNDANTLY in every slice CTA (4096 activations, 320 logits: cheaper than
//      broadcasting dh through memory); slice 0 also writes h / logits / dlogits / dh and the epoch accumulators,
//   5. runs the 32 dgrad MMAs (A = W slice MN-major, B = dh from shared memory) and stores its 224 columns of da2.
// The last slice to leave the barrier re-zeroes the accumulator and the counters, so the kernel is stateless between steps.
// Co-residency of a client's 14 CTAs: the grid has at most floor(SMs / 14) * 14 CTAs at one CTA per SM (launched
// cooperatively), and groups walk their clients in the same order, so a waiting CTA only ever waits for CTAs that are
// resident.  Waits are bounded (trap instead of hanging the GPU).
#include "tc_gemm.cuh"
#include "philox.cuh"
#include <stdlib.h>

namespace tc {

namespace {

using Off = SimpleCnnOff;
constexpr int F_IN = 3136, F_OUT = 128, F_S = 14, F_COLS = F_IN / F_S, F_CH = F_COLS / 32;     // 224 columns = 7 chunks per slice
static_assert(F_S * F_CH * 32 == F_IN, "slices tile the input features");
constexpr int F_THREADS = 256;
constexpr int W_BYTES = 4 * 8 * 4096;              // MN-major view: 4 out-blocks x 8 chunk slots (7 used) x [32 x 32]; K-major view: 7 x 16 KB
constexpr int A_BYTES = F_CH * 4096;               // activation slice, K-major [32 b x 32] chunks
constexpr int DH_BYTES = 4 * 4096;                 // dh, K-major [32 b x 32 out] chunks
constexpr int F_SMEM = W_BYTES + A_BYTES + DH_BYTES + 1024;

struct Fc1FusedParams {
    CUtensorMap map_w;        // fc1.weight {in, out, K}, box {32, 128, 1}, SWIZZLE_128B          (forward A operand)
    CUtensorMap map_w_mn;     // same tensor, box {32, 32, 1}, SWIZZLE_128B_ATOM_32B               (dgrad A operand, MN-major)
    CUtensorMap map_act;      // a2 {in, K*B}, box {32, 32}, SWIZZLE_128B                         (forward B operand)
    flb_train_args a;
    SimpleCnnWs ws;
    int groups;
};

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(F_THREADS, 1) fc1_fused_kernel(const __grid_constant__ Fc1FusedParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* wbuf = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* abuf = wbuf + W_BYTES;
    uint8_t* dhbuf = abuf + A_BYTES;
    __shared__ uint64_t full_bar[F_CH], fwd_bar, w2_bar, dg_bar;
    __shared__ uint32_t tmem_base;
    __shared__ __align__(16) float sh[32][132];          // h (after bias, ReLU, dropout)
    __shared__ float sw2[10][129];
    __shared__ float sb2[10];
    __shared__ float slog[32][10];
    __shared__ float sdl[32][12];
    __shared__ int slabel[32];
    __shared__ float red[2];
    __shared__ int s_last;

    const flb_train_args& a = p.a;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int slice = blockIdx.x, group = blockIdx.y, in0 = slice * F_COLS;

    if (tid == 0) {
        for (int i = 0; i < F_CH; ++i) mbar_init(&full_bar[i], 1);
        mbar_init(&fwd_bar, 1);
        mbar_init(&w2_bar, 1);
        mbar_init(&dg_bar, 1);
        fence_barrier_init();
        tma_prefetch_desc(&p.map_w);
        tma_prefetch_desc(&p.map_w_mn);
        tma_prefetch_desc(&p.map_act);
    }
    if (warp == 1) tmem_alloc<128>(&tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    const bool lead = elect_one();                   // one fixed lane per warp issues TMA / MMA (loops stay warp-uniform)
    const float keep_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;

    uint32_t it = 0;
    for (int c = group; c < a.K; c += p.groups) {
        const int bsz = flb_bsz(a, c);
        if (bsz == 0) continue;                      // same decision in all 14 slice CTAs of the client
        const uint32_t ph = it & 1;
        ++it;
        const float* W = a.W + (long long)c * a.ld;
        const long long kb0 = (long long)c * a.B;
        float* hacc = p.ws.hpre + kb0 * F_OUT;       // [B, 128] accumulator, zero at rest
        int* ctr = p.ws.fc1_ctr + 2 * c;             // [0] arrivals, [1] departures, zero at rest

        // Everything the head needs besides hpre is requested NOW by all threads (5 + 1 independent loads each) and parked in
        // registers while the warps do their forward roles; it lands in shared memory just before the first barrier.
        // (A first version let two idle warps copy fc2.weight in a loop of 20 dependent round trips: 10 us of the kernel.)
        float w2r[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) w2r[i] = W[Off::f2w + tid + F_THREADS * i];
        const float b2r = tid < 10 ? W[Off::f2b + tid] : 0.f;
        const int labr = (tid < 32 && tid < bsz) ? a.y[a.sample_off[c] + (long long)(*a.step_ctr) * a.B + tid] : 0;

        // ---- 1. forward operands + MMAs ------------------------------------------------------------------------------
        if (warp == 0) {
            for (int i = 0; i < F_CH; ++i) {
                if (lead) mbar_expect_tx(&full_bar[i], 128 * 128 + 32 * 128);
                if (lead) tma_load_3d(&p.map_w, wbuf + i * 16384, &full_bar[i], in0 + 32 * i, 0, c);
                if (lead) tma_load_2d(&p.map_act, abuf + i * 4096, &full_bar[i], in0 + 32 * i, (int)kb0);
            }
            __syncwarp();
            mbar_wait(&fwd_bar, ph);                 // the forward MMAs have read the K-major slice: overwrite it with the
            if (lead) mbar_expect_tx(&w2_bar, 4 * F_CH * 4096);        // MN-major view for the dgrad (an L2 hit by now)
            for (int i = 0; i < 4 * F_CH; ++i) {
                const int kb = i / F_CH, mc = i % F_CH;
                if (lead) tma_load_3d(&p.map_w_mn, wbuf + (kb * 8 + mc) * 4096, &w2_bar, in0 + 32 * mc, kb * 32, c);
            }
            __syncwarp();
        } else if (warp == 1) {
            constexpr uint32_t id = idesc_tf32(128, 32, false, false);
            for (int i = 0; i < F_CH; ++i) {
                mbar_wait(&full_bar[i], ph);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (lead) mma_tf32(tmem, smem_desc(smem_u32(wbuf) + i * 16384 + k * 32, 16, 1024),
                                       smem_desc(smem_u32(abuf) + i * 4096 + k * 32, 16, 1024), id, i > 0 || k > 0);
                __syncwarp();
            }
            if (lead) mma_commit(&fwd_bar);
            __syncwarp();
        } else if (warp >= 4) {
            // ---- 2. partial hpre -> the client's accumulator (lanes = consecutive output features: 128 B per reduction)
            mbar_wait(&fwd_bar, ph);
            tc_fence_after();
            const int q = warp & 3, j = q * 32 + lane;
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16), v);
#pragma unroll
            for (int b = 0; b < 32; ++b)
                if (b < bsz) atomicAdd(hacc + b * F_OUT + j, v[b]);
            __threadfence();
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) { const int e = tid + F_THREADS * i; sw2[e >> 7][e & 127] = w2r[i]; }
        if (tid < 10) sb2[tid] = b2r;
        if (tid < 32) slabel[tid] = labr;
        if (tid < 2) red[tid] = 0.f;
        tc_fence_before();
        __syncthreads();

        // ---- 3. the client's barrier -------------------------------------------------------------------------------------
        if (tid == 0) {
            atomicAdd(&ctr[0], 1);
            uint32_t spins = 0;
            while (ld_acquire(&ctr[0]) < F_S) {
                if (++spins > SPIN_LIMIT) __trap();
            }
        }
        __syncthreads();

        // ---- 4. head (every slice CTA computes it; slice 0 publishes) --------------------------------------------------
        const bool pub = slice == 0;
        float mult_r[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = (tid + F_THREADS * i) * 4, b = e >> 7, j = e & 127;
            float hv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int r = 0; r < 4; ++r) mult_r[i][r] = 0.f;
            if (b < bsz) {
                const float4 pre4 = __ldcg(reinterpret_cast<const float4*>(hacc + e));
                const float4 b4 = *reinterpret_cast<const float4*>(W + Off::f1b + j);
                const float pre[4] = {pre4.x + b4.x, pre4.y + b4.y, pre4.z + b4.z, pre4.w + b4.w};
                bool keep[4] = {true, true, true, true};
                if (a.drop_p > 0.f) {
                    if (a.drop_keep) {
#pragma unroll
                        for (int r = 0; r < 4; ++r) keep[r] = a.drop_keep[kb0 * 128 + e + r] != 0;
                    } else {
                        const flb_u4 rr = flb_philox_block(flb_epoch_seed(a) ^ 0xD80F0A7ull, a.client_base + a.client_stride * c,
                                                           ((unsigned long long)a.tcount[c] << 12) + (e >> 2));
                        keep[0] = flb_u01(rr.x) >= a.drop_p; keep[1] = flb_u01(rr.y) >= a.drop_p;
                        keep[2] = flb_u01(rr.z) >= a.drop_p; keep[3] = flb_u01(rr.w) >= a.drop_p;
                    }
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    mult_r[i][r] = (pre[r] > 0.f && keep[r]) ? keep_scale : 0.f;
                    hv[r] = pre[r] * mult_r[i][r];
                }
                if (pub) *reinterpret_cast<float4*>(p.ws.h + kb0 * 128 + e) = make_float4(hv[0], hv[1], hv[2], hv[3]);
            }
            *reinterpret_cast<float4*>(&sh[b][j]) = make_float4(hv[0], hv[1], hv[2], hv[3]);
        }
        __syncthreads();
        if (tid == 0) {                                       // every slice has read the accumulator once all 14 got here
            __threadfence();
            s_last = atomicAdd(&ctr[1], 1) == F_S - 1;
        }
        // logits: 4 lanes per dot product (32 elements each, interleaved so that the quad reads consecutive banks); every
        // thread owns 5 (sample, class) pairs.  (One warp per logit was a chain of 40 x 5 dependent shuffles per warp.)
        {
            const int part = tid & 3;
#pragma unroll 1
            for (int r = 0; r < 5; ++r) {
                const int eidx = (tid >> 2) + 64 * r;          // 320 pairs in all
                const bool ok = eidx < bsz * 10;
                const int e = ok ? eidx : 0, b = e / 10, j = e % 10;
                float t0 = 0.f, t1 = 0.f;
#pragma unroll
                for (int u = 0; u < 32; u += 2) {
                    t0 = fmaf(sh[b][4 * u + part], sw2[j][4 * u + part], t0);
                    t1 = fmaf(sh[b][4 * u + 4 + part], sw2[j][4 * u + 4 + part], t1);
                }
                float t = t0 + t1;
                t += __shfl_xor_sync(0xffffffffu, t, 1);
                t += __shfl_xor_sync(0xffffffffu, t, 2);
                if (part == 0 && ok) {
                    t += sb2[j];
                    slog[b][j] = t;
                    if (pub) p.ws.logits[kb0 * 10 + eidx] = t;
                }
            }
        }
        __syncthreads();
        if (s_last) {                                         // stateless between steps: accumulator and counters back to zero
            for (int e = tid; e < a.B * (F_OUT / 4); e += F_THREADS) reinterpret_cast<float4*>(hacc)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tid == 0) { ctr[0] = 0; ctr[1] = 0; }
        }
        // softmax cross-entropy: 16 lanes per sample (10 live), two passes of 16 samples
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const int b = (tid >> 4) + 16 * pass, j = tid & 15;
            const bool live = b < bsz && j < 10;
            const float v = live ? slog[b < bsz ? b : 0][j] : -INFINITY;
            float mx = v;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o, 16));
            int am = (live && v == mx) ? j : 99;              // first index of the maximum, like the sequential scan
            float se = live ? expf(v - mx) : 0.f;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                am = min(am, __shfl_xor_sync(0xffffffffu, am, o, 16));
                se += __shfl_xor_sync(0xffffffffu, se, o, 16);
            }
            const float lse = logf(se) + mx;
            const float gs = a.dp_mode == 1 ? 1.f : 1.f / (float)bsz;          // mean reduction (training.py:90)
            if (live) {
                const int y = slabel[b];
                const float d = (expf(v - lse) - (j == y ? 1.f : 0.f)) * gs;
                sdl[b][j] = d;
                if (pub) {
                    p.ws.dlog[kb0 * 10 + b * 10 + j] = d;
                    if (j == y) atomicAdd(&red[0], lse - v);
                    if (j == 0) atomicAdd(&red[1], am == y ? 1.f : 0.f);
                }
            }
        }
        __syncthreads();
        if (pub && tid == 0) {
            atomicAdd(&a.loss_sum[c], red[0] / (float)bsz);     // running_loss += loss.item()   (training.py:200)
            atomicAdd(&a.correct[c], (int)(red[1] + 0.5f));     // correct += (pred == y).sum()  (training.py:201-203)
            a.nbatch[c] += 1;
            a.nseen[c] += bsz;
        }
        // dh = (dlogits . W2) * mask, as the dgrad's K-major B operand (rows >= bsz are zero) and, from slice 0, to global
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = (tid + F_THREADS * i) * 4, b = e >> 7, j = e & 127;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            if (b < bsz) {
#pragma unroll
                for (int cl = 0; cl < 10; ++cl) {
                    const float d = sdl[b][cl];
#pragma unroll
                    for (int r = 0; r < 4; ++r) acc[r] = fmaf(d, sw2[cl][j + r], acc[r]);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[r] *= mult_r[i][r];
            }
            const float4 o = make_float4(acc[0], acc[1], acc[2], acc[3]);
            *reinterpret_cast<float4*>(dhbuf + (j >> 5) * 4096 + sw128_offset(b, j & 31)) = o;     // 4 consecutive k: one 16 B chunk
            if (pub && b < a.B) *reinterpret_cast<float4*>(p.ws.dh + kb0 * 128 + e) = o;           // rows bsz..B-1 = 0 for the wgrad GEMM
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();

        // ---- 5. dgrad: D[in, b] = sum_out W[out, in] * dh[b, out], two 128-row tiles (224 = 128 + 96) ---------------------
        if (warp == 1) {
            constexpr uint32_t id = idesc_tf32(128, 32, true, false);
            mbar_wait(&w2_bar, ph);
            tc_fence_after();
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (lead) mma_tf32(tmem + 32 + mt * 32, smem_desc_mn(smem_u32(wbuf) + (kb * 8 + mt * 4) * 4096 + k * 1024, 4096, 512),
                                           smem_desc(smem_u32(dhbuf) + kb * 4096 + k * 32, 16, 1024), id, kb > 0 || k > 0);
            if (lead) mma_commit(&dg_bar);
            __syncwarp();
        }
        mbar_wait(&dg_bar, ph);
        tc_fence_after();
        {
            const int q = warp & 3, mt = warp >> 2, m = mt * 128 + q * 32 + lane;
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 32 + mt * 32, v);
            if (m < F_COLS) {
                float* d = p.ws.da2 + kb0 * F_IN + in0 + m;
#pragma unroll
                for (int b = 0; b < 32; ++b)
                    if (b < bsz) d[(long long)b * F_IN] = v[b];
            }
        }
        tc_fence_before();
        __syncthreads();                                       // TMEM, operand buffers and head scratch are free for the next client
        tc_fence_after();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<128>(tmem);
}

}  // namespace

int make_fc1_maps(const flb_train_args& a, const float* act, CUtensorMap* w, CUtensorMap* w_mn, CUtensorMap* m_act);

// One launch for fc1 forward + classifier head + fc1 dgrad of every resident client (see the header comment).
int fc1_fused(const flb_train_args& a, const SimpleCnnWs& ws, cudaStream_t st) {
    if (a.B > 32) { flb_set_error("fc1_fused: batch size %d > 32", a.B); return FLB_ERR_ARG; }
    Fc1FusedParams p;
    if (int rc = make_fc1_maps(a, ws.a2, &p.map_w, &p.map_w_mn, &p.map_act)) return rc;
    p.a = a; p.ws = ws;
    const int max_groups = flb_num_sms() / F_S;
    p.groups = a.K < max_groups ? a.K : max_groups;
    if (p.groups < 1) { flb_set_error("fc1_fused: needs at least %d SMs", F_S); return FLB_ERR_UNSUPPORTED; }
    static thread_local int attr_dev = -1;
    int dev = 0;
    FLB_CUDA(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        FLB_CUDA(cudaFuncSetAttribute(fc1_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
        attr_dev = dev;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(F_S, p.groups, 1);
    cfg.blockDim = dim3(F_THREADS, 1, 1);
    cfg.dynamicSmemBytes = F_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;           // all CTAs co-resident or the launch fails: the client barrier spins
    at[0].val.cooperative = 1;
    static const bool coop = getenv("FLB_FC1_NO_COOP") == nullptr;       // A/B switch (the grid never exceeds one CTA per SM either way)
    cfg.attrs = at;
    cfg.numAttrs = coop ? 1 : 0;
    FLB_CUDA(cudaLaunchKernelEx(&cfg, fc1_fused_kernel, p));
    return FLB_OK;
}

}  // namespace tc
