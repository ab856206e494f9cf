// tcgen05 / TMEM / TMA building blocks and the warp-specialised grouped-GEMM skeleton (sm_100a only).
//
// One CTA computes one 128-row accumulator tile (or several 128-row tiles sharing the streamed operand) for one
// client: warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator + MMA issuer (one elected lane issues
// tcgen05.mma kind::tf32, fp32 accumulate in TMEM), warps 2-5 = epilogue (tcgen05.ld 32x32b, one TMEM lane per
// thread).  Operands are staged in shared memory in the canonical 128-byte-swizzled layouts: TMA boxes with a
// 128-byte inner extent land directly in that layout; small weight operands that need a permutation are written by
// all threads ("resident" operand) before the pipeline starts.  The problem-specific parts (which boxes to load per
// k-block, which MMAs to issue, what the epilogue does with the accumulator) come from a Traits class.
#pragma once
#include <cuda.h>
#include "train_common.cuh"

namespace tc {

constexpr int THREADS = 192;
constexpr uint32_t SPIN_LIMIT = 1u << 26;      // bounded waits: a broken pipeline traps instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > SPIN_LIMIT) __trap();
    }
}
// one lane of a converged warp; the same lane every time
__device__ __forceinline__ bool elect_one() {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok));
    return ok != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, void* dst, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, void* dst, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {       // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr) {           // the allocating warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}

// shared-memory matrix descriptors, version 1 (Blackwell).  lbo / sbo in bytes.
//   K-major, SWIZZLE_128B (layout type 2): rows of 128 B (32 fp32 along K), 16-byte chunks XOR-ed with (row % 8);
//       8-row groups `sbo` apart (1024 when rows are dense); lbo unused.  Matches TMA CU_TENSOR_MAP_SWIZZLE_128B.
//   MN-major tf32, SWIZZLE_128B with 32-byte atoms (layout type 1, the only MN-major layout the tf32 MMA accepts):
//       rows of 128 B (32 fp32 along M/N), consecutive K indices 128 B apart, 32-byte chunks XOR-ed with (row % 4);
//       4-K groups `sbo` apart (512 when dense), 32-element M/N chunks `lbo` apart.
//       Matches TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
__host__ __device__ constexpr uint64_t smem_desc_raw(uint32_t addr, uint32_t lbo, uint32_t sbo, uint64_t layout_type) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (layout_type << 61);
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) { return smem_desc_raw(addr, lbo, sbo, 2); }
// K-major SW128 operand that starts at an arbitrary 128-byte row of a 1024-byte-aligned swizzled tile (the shifted
// windows of the halo convolutions).  Measured on B200: the tensor core un-swizzles with the ABSOLUTE shared-memory
// address bits (chunk bits [4:6] ^= row bits [7:9]), so a start address that is only 128-byte aligned needs nothing
// else -- setting the descriptor's "matrix base offset" field (bits 49-51) to the row phase gives WRONG results here;
// it stays 0.
__device__ __forceinline__ uint64_t smem_desc_row(uint32_t addr) { return smem_desc_raw(addr, 16, 1024, 2); }
__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t addr, uint32_t lbo, uint32_t sbo) { return smem_desc_raw(addr, lbo, sbo, 1); }

// A descriptor whose start address is `bytes` further on (bytes % 16 == 0).  The address field is the low 14 bits of
// (addr >> 4) and shared-memory addresses stay below 256 KB, so the sum never carries into the neighbouring field: one
// 32-bit add on the low word instead of rebuilding the descriptor (shift, mask, two ORs) for every MMA.  In the halo
// convolutions that rebuild was ~12 uniform-datapath instructions per tcgen05.mma -- as long as the 40-cycle MMA itself
// (N = 32; scripts/mma_microbench.py): the issuing warp, not the tensor core, set the pace.
__device__ __forceinline__ uint64_t desc_advance(uint64_t d, uint32_t bytes) {
    return (d & 0xFFFFFFFF00000000ull) | (uint64_t)((uint32_t)d + (bytes >> 4));
}

// instruction descriptor: D = fp32, A = B = tf32, dense; a_mn / b_mn select MN-major operands
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Column sums of a warp's 32 x 32 register tile (lane = row, v[i] = column i): returns in lane j the sum over all lanes
// of v[j].  Transposing butterfly: every step exchanges half of the still-live columns, 31 shuffles in all.  v is clobbered.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = lane & off;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = up ? v[i] : v[i + off];
            const float keep = up ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

// Live batch size of every resident client, read ONCE per CTA into shared memory.  The persistent skeletons call
// tile_setup() for every tile from every role; with flb_bsz() inside, that was two dependent global loads (~0.8 us) on the
// critical path of each tile for the MMA warp, the producer and the epilogue warps (ncu source view, round 2).
constexpr int BSZ_TAB = 1024;                        // flb_train_args.K <= 1024 (check_args)
__device__ __forceinline__ void fill_bsz_table(const flb_train_args& a, int* tab) {
    for (int c = threadIdx.x; c < a.K && c < BSZ_TAB; c += blockDim.x) tab[c] = flb_bsz(a, c);
    // visible to all roles after the kernel's first __syncthreads()
}

// ---- per-role timeline of one CTA (profiling aid; flb_debug_trace_enable / _read in train_api.cu) -------------------------
// When enabled, CTA 0 of the resident-weight convolution kernels appends (event id, tile, clock64) records: the only way to
// see which of the three decoupled roles a tile is actually waiting for (ncu's warp-state samples cannot tell a producer
// that waits for a free stage from one that waits for memory).
struct TraceBuf { int n; int cap; long long rec[1]; };          // rec[3*i] = event, rec[3*i+1] = tile, rec[3*i+2] = clock
static __device__ TraceBuf* g_trace = nullptr;      // per translation unit (no -rdc): set by flb_debug_trace_set in train_tc.cu
// Every recording warp owns cap / 8 consecutive slots and counts them in a register (`slot`): three plain stores per event,
// no atomic round trip in the timed code.  Unused slots stay zero (the buffer is cleared by the host).
// Compiled in only with -DFLB_TRACE=1 (FLB_TRACE=1 python -m flb200.build --force): the product build carries no trace code.
#ifndef FLB_TRACE
#define FLB_TRACE 0
#endif
__device__ __forceinline__ void trace_event(TraceBuf* tb, int& slot, int ev, int tile) {
#if FLB_TRACE
    if (!tb) return;
    // slot = (first record index of this warp's region) + count, kept in a register; the region's capacity is cap / 8 - 1
    long long* r = tb->rec + 3 * (long long)slot;
    r[0] = ev; r[1] = tile; r[2] = clock64();
    ++slot;
#endif
}

// optional Traits::finish(p, lane): called once by every epilogue warp after its last tile (persistent skeletons)
template <class T, class P>
__device__ __forceinline__ auto call_finish(T& t, const P& p, int lane, int) -> decltype(t.finish(p, lane), void()) { t.finish(p, lane); }
template <class T, class P>
__device__ __forceinline__ void call_finish(T&, const P&, int, long) {}

// byte offset of element (row, k) inside a K-major SW128 tile whose rows are 128 B (32 fp32) and dense
__device__ __forceinline__ uint32_t sw128_offset(int row, int k) {
    return (uint32_t)row * 128u + ((((uint32_t)k >> 2) ^ ((uint32_t)row & 7u)) << 4) + (((uint32_t)k & 3u) << 2);
}

// Traits interface:
//   struct Params;                                   kernel parameter block (holds CUtensorMaps by value)
//   static constexpr int STAGES, STAGE_BYTES, RESIDENT_BYTES, TMEM_COLS, MINB (CTAs per SM the kernel is built for);
//   __device__ bool setup(const Params&, int& num_kb)          CTA-uniform; false -> nothing to do
//   __device__ void stage_resident(uint8_t* res, int tid)      all threads (generic-proxy writes)
//   __device__ void load(int kb, uint8_t* stage, uint64_t* bar) producer lane: expect_tx + TMA boxes
//   __device__ void mma(int kb, uint32_t stage_addr, uint32_t res_addr, uint32_t tmem)   MMA lane
//   __device__ void epilogue(uint32_t tmem, int quarter, int lane)   epilogue warps, quarter = warp % 4
template <class T>
__global__ void __launch_bounds__(THREADS, T::MINB) gemm_kernel(const __grid_constant__ typename T::Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full_bar[T::STAGES], empty_bar[T::STAGES], accum_bar;
    __shared__ uint32_t tmem_base;
    uint8_t* stages = smem;
    uint8_t* resident = smem + (size_t)T::STAGES * T::STAGE_BYTES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T t;
    int num_kb = 0;
    if (!t.setup(p, num_kb)) return;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < T::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&accum_bar, 1);
        fence_barrier_init();
        t.prefetch(p);
    }
    if (warp == 1) tmem_alloc<T::TMEM_COLS>(&tmem_base);
    if (T::RESIDENT_BYTES > 0) {
        t.stage_resident(p, resident, threadIdx.x);
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;

    // The producer and MMA warps run their loops with ALL lanes (warp-uniform control flow) and predicate only the
    // issuing instructions on one elected lane: descriptors, coordinates and barrier addresses then live in uniform
    // registers.  (Putting the whole loop under `if (lane == 0)` made the compiler wrap every UTCHMMA / UTMALDG in an
    // ELECT + R2UR.BROADCAST waterfall loop, ~150 cycles per instruction -- measured, see profiles/.)
    if (warp == 0) {
        t.lead = elect_one();
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % T::STAGES, it = kb / T::STAGES;
            mbar_wait(&empty_bar[s], (it & 1) ^ 1);
            t.load(p, kb, stages + (size_t)s * T::STAGE_BYTES, &full_bar[s]);
            __syncwarp();
        }
    } else if (warp == 1) {
        t.lead = elect_one();
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % T::STAGES, it = kb / T::STAGES;
            mbar_wait(&full_bar[s], it & 1);
            tc_fence_after();
            t.mma(kb, smem_u32(stages + (size_t)s * T::STAGE_BYTES), smem_u32(resident), tmem);
            if (t.lead) mma_commit(&empty_bar[s]);          // frees the stage when these MMAs have read it
            __syncwarp();
        }
        if (t.lead) mma_commit(&accum_bar);                 // accumulator complete
    } else {
        mbar_wait(&accum_bar, 0);
        tc_fence_after();
        t.epilogue(p, tmem, warp & 3, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<T::TMEM_COLS>(tmem);
}

// Persistent variant for GEMMs with many independent output tiles (conv fwd / dgrad): gridDim.x CTAs each walk the
// tiles blockIdx.x, blockIdx.x + gridDim.x, ...  The three roles run decoupled: the TMA producer streams k-blocks of
// tile i+1 while the MMA lane is still on tile i, and the accumulator is double-buffered in TMEM (2 x ACC_COLS
// columns) so the epilogue warps drain tile i while the MMAs of tile i+1 are issued.
// Extra Traits members:  static constexpr int ACC_COLS;
//                        __device__ bool tile_setup(const Params&, int tile, int& num_kb)   (same answer for every role)
//                        static __host__ __device__ int num_tiles(const Params&)
template <class T>
__global__ void __launch_bounds__(THREADS, T::MINB) gemm_persistent_kernel(const __grid_constant__ typename T::Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* stages = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full_bar[T::STAGES], empty_bar[T::STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base;
    __shared__ int s_bsz[BSZ_TAB];
    fill_bsz_table(p.a, s_bsz);
    constexpr int TCOLS = 2 * T::ACC_COLS <= 32 ? 32 : (2 * T::ACC_COLS <= 64 ? 64 : (2 * T::ACC_COLS <= 128 ? 128 : (2 * T::ACC_COLS <= 256 ? 256 : 512)));
    static_assert(2 * T::ACC_COLS <= 512, "double-buffered accumulator must fit TMEM");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = T::num_tiles(p);
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < T::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 4); }
        fence_barrier_init();
        T t0;
        t0.prefetch(p);
    }
    if (warp == 1) tmem_alloc<TCOLS>(&tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;

    if (warp == 0) {                                           // all lanes run the loop; one elected lane issues (see gemm_kernel)
        T t;
        t.bsz_tab = s_bsz;
        t.lead = elect_one();
        uint32_t it = 0;                                       // k-blocks issued so far (all tiles)
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            int num_kb = 0;
            if (!t.tile_setup(p, tile, num_kb)) continue;
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const uint32_t s = it % T::STAGES, ph = (it / T::STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                t.load(p, kb, stages + (size_t)s * T::STAGE_BYTES, &full_bar[s]);
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        T t;
        t.bsz_tab = s_bsz;
        t.lead = elect_one();
        uint32_t it = 0, nt = 0;                               // k-blocks / tiles consumed so far
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            int num_kb = 0;
            if (!t.tile_setup(p, tile, num_kb)) continue;
            const uint32_t acc = nt & 1, aph = (nt >> 1) & 1;
            mbar_wait(&tempty_bar[acc], aph ^ 1);              // the epilogue has drained this accumulator
            tc_fence_after();
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const uint32_t s = it % T::STAGES, ph = (it / T::STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                t.mma(kb, smem_u32(stages + (size_t)s * T::STAGE_BYTES), 0, tmem + acc * T::ACC_COLS);
                if (t.lead) mma_commit(&empty_bar[s]);
                __syncwarp();
            }
            if (t.lead) mma_commit(&tfull_bar[acc]);
            ++nt;
        }
    } else {
        T t;
        t.bsz_tab = s_bsz;
        uint32_t nt = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            int num_kb = 0;
            if (!t.tile_setup(p, tile, num_kb)) continue;
            const uint32_t acc = nt & 1, aph = (nt >> 1) & 1;
            mbar_wait(&tfull_bar[acc], aph);
            tc_fence_after();
            t.epilogue(p, tmem + acc * T::ACC_COLS, warp & 3, lane);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            ++nt;
        }
        call_finish(t, p, lane, 0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TCOLS>(tmem);
}

// Halo convolution with RESIDENT weights: like the persistent kernel, but each CTA owns a contiguous range of tiles
// (consecutive tiles belong to the same client), keeps that client's whole tap-major weight tensor in shared memory and
// streams only the activation boxes; the weights are re-loaded (w_full / w_empty handshake) when the client changes.
// Extra Traits members: W_BYTES; EPI_WARPS (4 or 8); tile_place(p, client, m0, live_rows) -> num_kb; load_w(p, wres, bar);
// load_a(p, kb, stage, bar); mma(kb, stage_addr, wres_addr, tmem); epilogue(p, tmem, quarter, lane, part, nparts).
//
// Tile walk.  A CTA's tiles are consecutive, so (client, first row m0) advance incrementally -- no division, no global
// load and no look-ahead call per tile.  (A per-role timeline of this kernel, scripts/conv_timeline.py, showed the MMA warp
// spending ~1900 of its ~6600 cycles per tile of SimpleCNN's conv2 dgrad in tile_setup() -- an integer division by a
// run-time tile count plus constant-bank reloads, twice per tile because the weight hand-back looked one tile ahead.)
// Tiles at or past a client's live rows (ragged last batch, finished clients) are skipped by jumping to the next client.
struct TileWalk {
    int tile, t1, tpc, client, m0, rows_pc, pp;
    const int* bsz;
    __device__ __forceinline__ void init(int t0, int t1_, int rows_per_client, int rows_per_image, const int* bsz_tab) {
        tile = t0; t1 = t1_; rows_pc = rows_per_client; pp = rows_per_image; bsz = bsz_tab;
        tpc = (rows_pc + 127) >> 7;
        client = t0 / tpc;                                     // the one division of the kernel
        m0 = (t0 - client * tpc) << 7;
    }
    // positions on the next live tile; false when the range is exhausted.  live = live rows of `client`
    __device__ __forceinline__ bool seek(int& live) {
        while (tile < t1) {
            live = bsz[client] * pp;
            if (m0 < live) return true;
            tile = (client + 1) * tpc; ++client; m0 = 0;        // nothing (more) to do for this client
        }
        return false;
    }
    __device__ __forceinline__ void advance() {
        ++tile; m0 += 128;
        if (m0 >= (tpc << 7)) { m0 = 0; ++client; }
    }
};

template <class T>
__global__ void __launch_bounds__(64 + 32 * T::EPI_WARPS, 1) conv_resident_kernel(const __grid_constant__ typename T::Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* wres = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* stages = wres + T::W_BYTES;
    __shared__ uint64_t full_bar[T::STAGES], empty_bar[T::STAGES], tfull_bar[2], tempty_bar[2], wfull_bar, wempty_bar;
    __shared__ uint32_t tmem_base;
    __shared__ int s_bsz[BSZ_TAB];
    fill_bsz_table(p.a, s_bsz);
    constexpr int TCOLS = 2 * T::ACC_COLS <= 32 ? 32 : (2 * T::ACC_COLS <= 64 ? 64 : (2 * T::ACC_COLS <= 128 ? 128 : (2 * T::ACC_COLS <= 256 ? 256 : 512)));
    static_assert(T::EPI_WARPS == 4 || T::EPI_WARPS == 8, "one or two epilogue warps per TMEM lane quarter");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = T::num_tiles(p);
    const int t0 = (int)((long long)blockIdx.x * ntiles / gridDim.x), t1 = (int)((long long)(blockIdx.x + 1) * ntiles / gridDim.x);
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < T::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], T::EPI_WARPS); }
        mbar_init(&wfull_bar, 1);
        mbar_init(&wempty_bar, 1);
        fence_barrier_init();
        T t0_;
        t0_.prefetch(p);
    }
    if (warp == 1) tmem_alloc<TCOLS>(&tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    TraceBuf* tb = (FLB_TRACE && blockIdx.x == 0 && lane == 0) ? g_trace : nullptr;      // one recorder lane per warp of CTA 0
    int slot = 0, slot_base = 0;
    if (tb) {                                                  // resume after the previous traced kernel (a region's last record holds its count)
        slot_base = (threadIdx.x >> 5) * (tb->cap >> 3);
        slot = slot_base + (int)tb->rec[3 * (slot_base + (tb->cap >> 3) - 1)];
    }
    trace_event(tb, slot, 1, warp);                            // 1: role loops start

    TileWalk w;
    w.init(t0, t1, p.a.B * p.g.PP(), p.g.PP(), s_bsz);
    int live = 0;
    if (warp == 0) {                                           // all lanes run the loop; one elected lane issues (see gemm_kernel)
        T t;
        t.lead = elect_one();
        uint32_t it = 0, wuse = 0;
        int cur = -1;
        for (; w.seek(live); w.advance()) {
            const int num_kb = t.tile_place(p, w.client, w.m0, live);
            if (w.client != cur) {
                mbar_wait(&wempty_bar, (wuse & 1) ^ 1);                // every MMA that read the old weights has completed
                t.load_w(p, wres, &wfull_bar);
                __syncwarp();
                cur = w.client;
                ++wuse;
                trace_event(tb, slot, 10, w.tile);                   // 10: weight load issued
            }
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const uint32_t s = it % T::STAGES, ph = (it / T::STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                t.load_a(p, kb, stages + (size_t)s * T::STAGE_BYTES, &full_bar[s]);
                __syncwarp();
                trace_event(tb, slot, 11, w.tile);                   // 11: activation box issued (after the stage was free)
            }
        }
    } else if (warp == 1) {
        T t;
        t.lead = elect_one();
        uint32_t it = 0, nt = 0, wuse = 0;
        int cur = -1;
        for (; w.seek(live); w.advance()) {
            trace_event(tb, slot, 27, w.tile);                       // 27: next tile found
            const int num_kb = t.tile_place(p, w.client, w.m0, live);
            if (w.client != cur) {
                mbar_wait(&wfull_bar, wuse & 1);
                cur = w.client;
                ++wuse;
                trace_event(tb, slot, 20, w.tile);                   // 20: weights landed
            }
            const uint32_t acc = nt & 1, aph = (nt >> 1) & 1;
            trace_event(tb, slot, 28, w.tile * 2 + acc);             // 28: about to wait for the accumulator (tile * 2 + acc)
            mbar_wait(&tempty_bar[acc], aph ^ 1);
            trace_event(tb, slot, 29, w.tile);                       // 29: accumulator wait over
            tc_fence_after();
            trace_event(tb, slot, 21, w.tile);                       // 21: accumulator free
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const uint32_t s = it % T::STAGES, ph = (it / T::STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                trace_event(tb, slot, 22, w.tile);                   // 22: activation box landed
                t.mma(kb, smem_u32(stages + (size_t)s * T::STAGE_BYTES), smem_u32(wres), tmem + acc * T::ACC_COLS);
                if (t.lead) mma_commit(&empty_bar[s]);
                __syncwarp();
                trace_event(tb, slot, 23, w.tile);                   // 23: MMAs of the k-block issued
            }
            if (t.lead) mma_commit(&tfull_bar[acc]);
            ++nt;
            // Hand the weight buffer back after the client's last live tile (the producer waits for this only if it has
            // another client's weights to load; a hand-back nobody waits for is harmless).
            if (w.m0 + 128 >= live && t.lead) mma_commit(&wempty_bar);
            __syncwarp();
            trace_event(tb, slot, 24, w.tile);                       // 24: tile handed to the epilogue
        }
    } else {
        T t;
        uint32_t nt = 0;
        const int part = (warp - 2) >> 2;
        for (; w.seek(live); w.advance()) {
            t.tile_place(p, w.client, w.m0, live);
            const uint32_t acc = nt & 1, aph = (nt >> 1) & 1;
            mbar_wait(&tfull_bar[acc], aph);
            tc_fence_after();
            if (warp == 2) trace_event(tb, slot, 30, w.tile);        // 30: accumulator complete (MMAs done)
            t.epilogue(p, tmem + acc * T::ACC_COLS, warp & 3, lane, part, T::EPI_WARPS / 4);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (warp == 2) trace_event(tb, slot, 31, w.tile);        // 31: epilogue of the tile done
            ++nt;
        }
        call_finish(t, p, lane, 0);
    }
    trace_event(tb, slot, 2, warp);                            // 2: role done
    if (tb) tb->rec[3 * (slot_base + (tb->cap >> 3) - 1)] = slot - slot_base;
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TCOLS>(tmem);
}

}  // namespace tc
