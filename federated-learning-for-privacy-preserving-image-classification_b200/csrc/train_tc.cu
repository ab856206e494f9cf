// TF32 tensor-core (tcgen05 + TMEM, TMA-fed) versions of the GEMM-shaped training kernels, batched over clients.
//
//   conv fwd    D[px, Cout]      = sum_tap  X[px + shift(tap), Cin] * W[Cout, tap, Cin]^T       X streamed by TMA (K-major)
//   conv dgrad  D[px, Cin]       = sum_tap dZ[px - shift(tap), Cout] * W[Cout, Cin, tap]        dZ streamed by TMA (K-major)
//   conv wgrad  D[(tap,Cin), Cout] = sum_px X[px + shift(tap), Cin]^T * dZ[px, Cout]            both MN-major, split over pixels
//   fc fwd      D[Out, B]        = W[Out, In] * act[B, In]^T                                     both K-major, split-K
//   fc dgrad    D[In, B]         = W[Out, In]^T * dout[B, Out]^T                                 A MN-major, B K-major
//   fc wgrad    D[Out, In]       = dout[B, Out]^T * act[B, In]                                   both MN-major
// "px" runs over the zero-padded NHWC grid (train_common.cuh), so a 3x3 tap is a row shift of the TMA box and the
// out-of-range rows come back as zeros from the TMA unit -- implicit GEMM with no im2col buffer and no halo code.
// Conv weights are read from a per-client "tap-major" copy Wt[tap][Cout][Cin] (train_common.cuh TcConvTab) that the
// optimizer kernel keeps in sync with the reference-layout row, so BOTH operands of every GEMM are plain TMA boxes and
// no CTA spends time permuting weights; conv weight gradients are accumulated in the same layout (Gt).
#include "tc_gemm.cuh"
#include "opt_update.cuh"
#include <type_traits>
#include <stdlib.h>

namespace tc {

// ---- tensor maps ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// fp32 tensor of `rank` dims (dims[0] innermost, strides in bytes for dims 1..), box with a 32-float (128 B) inner extent
static int make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, bool mn_major) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { flb_set_error("cuTensorMapEncodeTiled is not available from the driver"); return FLB_ERR_CUDA; }
    cuuint64_t gd[3]; cuuint64_t gs[2]; cuuint32_t bx[3]; cuuint32_t es[3] = {1, 1, 1};
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { flb_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return FLB_ERR_CUDA; }
    return FLB_OK;
}

// mn_major: the box feeds an MN-major tf32 operand (rows = K indices), which needs the 32-byte-atom swizzle
static int make_map_2d(CUtensorMap* m, const float* base, uint64_t rows, uint64_t cols, uint32_t box_rows, bool mn_major = false) {
    const uint64_t dims[2] = {cols, rows}, strides[1] = {cols * sizeof(float)};
    const uint32_t box[2] = {32, box_rows};
    return make_map(m, base, 2, dims, strides, box, mn_major);
}

__device__ __forceinline__ int tap_shift(int tap, int Wp) { return (tap / 3 - 1) * Wp + (tap % 3 - 1); }


// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of the function: set it once per (kernel, device)
static int ensure_smem_attr(const void* fn, int smem) {
    struct Seen { const void* fn; int dev; };
    static thread_local Seen seen[64];
    static thread_local int n = 0;
    int dev = 0;
    FLB_CUDA(cudaGetDevice(&dev));
    for (int i = 0; i < n; ++i)
        if (seen[i].fn == fn && seen[i].dev == dev) return FLB_OK;
    FLB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (n < 64) seen[n++] = Seen{fn, dev};
    return FLB_OK;
}

template <int N> struct Pow2Cols { static constexpr int value = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : (N <= 256 ? 256 : 512))); };

// ---- conv forward: D[128 px, COUT] = sum_{tap, cin chunk} X[px + shift(tap), 32] * Wt[tap][COUT][32]^T ---------------
// POOL (SimpleCNN conv2: 16-wide grid, 256 rows per image, so a 128-row tile is 8 whole grid rows): the epilogue applies
// bias + ReLU + 2x2 max-pool (+argmax) with warp shuffles -- a warp's 32 accumulator rows are two adjacent grid rows --
// and writes fc1's NCHW-flattened input directly; the pre-pool activation never goes to HBM.
// NPART > 1: the k-loop is spread over NPART partial accumulators (column blocks NPART x COUT of TMEM) so that
// consecutive MMAs do not form one dependent chain on a single accumulator; the epilogue adds the partials.
template <int CIN, int COUT, bool POOL = false, int NPART = 1>
struct ConvFwdT {
    // bn_acc (optional): per-client BatchNorm accumulators [K][4][bn_stride] in double; the epilogue adds the column sums
    // and sums of squares of z over the real pixels of the live samples to rows 0 / 1 at channel offset bn_coff
    struct Params { CUtensorMap map_x; CUtensorMap map_w; flb_train_args a; ConvGeom g; float* z_all; int boff;
                    float* pool_out; uint8_t* pool_idx; double* bn_acc; int bn_coff, bn_stride; };
    bool lead = false;               // this lane issues the TMA / MMA instructions (skeleton sets it; the rest of the warp runs along)
    static constexpr int ACC_COLS = NPART * COUT;
    // 32 accumulator columns [c0, c0 + 32) of this thread's row, partials summed
    static __device__ __forceinline__ void ld_acc(uint32_t taddr, float* v) {
        tmem_ld32(taddr, v);
#pragma unroll
        for (int j = 1; j < NPART; ++j) {
            float u[32];
            tmem_ld32(taddr + j * COUT, u);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += u[i];
        }
    }
    static __host__ __device__ int num_tiles(const Params& p) { return ((p.a.B * p.g.PP() + 127) / 128) * p.a.K; }
    const int* bsz_tab = nullptr;    // per-CTA shared-memory copy of the clients' live batch sizes (persistent skeletons)
    __device__ bool tile_setup(const Params& p, int tile, int& num_kb) {
        const int tpc = (p.a.B * p.g.PP() + 127) / 128;
        client = tile / tpc;
        const int bsz = (bsz_tab && client < BSZ_TAB) ? bsz_tab[client] : flb_bsz(p.a, client);
        m0 = (tile - client * tpc) * 128;
        live_rows = bsz * p.g.PP();
        if (m0 >= live_rows) return false;
        row0 = client * p.a.B * p.g.PP();
        num_kb = NKB;
        return true;
    }
    // resident-weight skeleton: the tile walk (tc_gemm.cuh TileWalk) hands over the placement, nothing is recomputed
    __device__ __forceinline__ int tile_place(const Params& p, int client_, int m0_, int live_rows_) {
        client = client_; m0 = m0_; live_rows = live_rows_;
        row0 = client * p.a.B * p.g.PP();
        return NKB;
    }
    static constexpr int CH = CIN / 32, NKB = 9 * CH, A_BYTES = 128 * 128, B_BYTES = COUT * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES, STAGES = COUT > 64 ? 3 : 4, RESIDENT_BYTES = 0;
    static constexpr int TMEM_COLS = Pow2Cols<COUT>::value, MINB = 2;
    int client, m0, row0, live_rows = 0;
    int epart = 0, enparts = 1;      // this epilogue warp handles the 32-column chunks c with c % enparts == epart
    int pool_tile = 0;               // tiles this epilogue warp has pooled (parity selects the staging buffer)
    // fused BatchNorm statistics: lane j of every epilogue warp carries the partial (sum, sum of squares) of column
    // 32 * chunk + j over the rows that warp has seen for client `bclient`
    // ROWACC (COUT = 32): the sums stay per ROW LANE (2 x 32 registers) across tiles and are transposed only when they are
    // flushed -- the epilogue of the 32 -> 32 layer is as long as its MMAs, two butterflies per tile would make it the limiter
    static constexpr bool ROWACC = COUT == 32;
    float bs0[COUT / 32] = {}, bs1[COUT / 32] = {};
    float rs0[ROWACC ? 32 : 1] = {}, rs1[ROWACC ? 32 : 1] = {};
    int bclient = -1;
    __device__ void bn_flush(const Params& p, int lane) {
        if (bclient < 0) return;
        if (ROWACC) {
            bs0[0] = warp_colsum32(reinterpret_cast<float (&)[32]>(rs0), lane);
            bs1[0] = warp_colsum32(reinterpret_cast<float (&)[32]>(rs1), lane);
#pragma unroll
            for (int i = 0; i < (ROWACC ? 32 : 1); ++i) { rs0[i] = 0.f; rs1[i] = 0.f; }
        }
        double* A = p.bn_acc + (long long)bclient * 4 * p.bn_stride + p.bn_coff + lane;
#pragma unroll
        for (int ch = 0; ch < COUT / 32; ++ch) {
            if (ch % enparts != epart) continue;                // another epilogue warp of this lane quarter owns the chunk
            atomicAdd(A + ch * 32, (double)bs0[ch]);
            atomicAdd(A + p.bn_stride + ch * 32, (double)bs1[ch]);
            bs0[ch] = 0.f; bs1[ch] = 0.f;
        }
    }
    __device__ void finish(const Params& p, int lane) {
        if (!POOL && p.bn_acc) bn_flush(p, lane);
    }
    __device__ bool setup(const Params& p, int& num_kb) {
        client = blockIdx.y;
        const int bsz = flb_bsz(p.a, client);
        m0 = blockIdx.x * 128;
        live_rows = bsz * p.g.PP();
        if (m0 >= bsz * p.g.PP()) return false;
        row0 = client * p.a.B * p.g.PP();
        num_kb = NKB;
        return true;
    }
    __device__ void prefetch(const Params& p) { tma_prefetch_desc(&p.map_x); tma_prefetch_desc(&p.map_w); }
    __device__ void stage_resident(const Params&, uint8_t*, int) {}
    __device__ void load(const Params& p, int kb, uint8_t* stage, uint64_t* bar) {
        const int tap = kb / CH, c = kb % CH;
        if (this->lead) mbar_expect_tx(bar, STAGE_BYTES);
        if (this->lead) tma_load_2d(&p.map_x, stage, bar, c * 32, row0 + m0 + tap_shift(tap, p.g.Wp));
        if (this->lead) tma_load_3d(&p.map_w, stage + A_BYTES, bar, c * 32, tap * COUT, client);
    }
    __device__ void mma(int kb, uint32_t stage, uint32_t, uint32_t tmem) {
        constexpr uint32_t id = idesc_tf32(128, COUT, false, false);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (this->lead) mma_tf32(tmem, smem_desc(stage + k * 32, 16, 1024), smem_desc(stage + A_BYTES + k * 32, 16, 1024), id, kb > 0 || k > 0);
    }
    // part / nparts: with two epilogue warps per TMEM lane quarter (conv_resident_kernel, EPI_WARPS = 8) each takes every
    // second 32-column chunk -- the pooled epilogue is a chain of ~14 dependent instructions per channel and set the tile
    // rate with four warps (3850 of 4400 cycles per tile, scripts/conv_timeline.py)
    __device__ void epilogue(const Params& p, uint32_t tmem, int quarter, int lane, int part = 0, int nparts = 1) {
        epart = part; enparts = nparts;
        const int m = m0 + quarter * 32 + lane;
        const float* bias = p.a.W + (long long)client * p.a.ld + p.boff;
        if constexpr (POOL) {
            // 16-wide grid, 256 rows per image: a 128-row tile is half an image = 8 grid rows = 4 pooled rows of 7 windows.
            // The accumulator tile goes to shared memory as it is ([row][64 channels], 16-byte chunks XOR-swizzled with the
            // row so that the 8 lanes of a store phase hit 8 different bank groups); then every epilogue thread pools
            // (window, channel quad) tasks with four 16-byte loads, and the results are written out coalesced through a small
            // staging array (for a fixed channel a tile's windows are one contiguous run of fc1's NCHW-flattened input).
            // MIO instructions per thread and tile: ~60.  (The first version pooled with warp shuffles straight out of the
            // TMEM registers: 96 SHFL + 32 LDG + 64 STS per thread and 32 channels -- MIO-bound at ~3800 cycles per tile
            // against ~2200 for the tile's MMAs, unchanged by a second warp per lane quarter or by staging the stores;
            // scripts/conv_timeline.py.)
            static_assert(COUT == 64, "pooled epilogue: 64 output channels");
            __shared__ __align__(16) float s_x[128 * COUT];
            __shared__ float s_val[COUT * 28];
            __shared__ uint8_t s_idx[COUT * 28];
            const int nthr = 128 * nparts, e0 = (part * 4 + quarter) * 32 + lane;       // epilogue thread id
            const int ml = quarter * 32 + lane;                                       // row within the tile
#pragma unroll 1
            for (int c0 = part * 32; c0 < COUT; c0 += 32 * nparts) {
                float v[32];
                ld_acc(tmem + ((uint32_t)(quarter * 32) << 16) + c0, v);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    *reinterpret_cast<float4*>(&s_x[ml * COUT + (((c0 >> 2) + i) ^ (ml & 7)) * 4]) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
            asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");         // all epilogue warps (warps 0 / 1 are the producer and the MMA issuer)
            const int b = m0 >> 8, half = (m0 >> 7) & 1;
            {
                const int cq = e0 & 15;                                    // channel quad of this thread's tasks
                const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + 4 * cq));
                for (int j = e0 >> 4; j < 28; j += nthr >> 4) {            // window j = pooled row (0..3) * 7 + pooled column
                    const int pr = j / 7, pc = j - pr * 7;
                    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                    int bi[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int d = 0; d < 4; ++d) {                          // (0,0), (0,1), (1,0), (1,1): first maximum wins, like the reference scan
                        const int r = (2 * pr + (d >> 1)) * 16 + 2 * pc + (d & 1);
                        const float4 x4 = *reinterpret_cast<const float4*>(&s_x[r * COUT + ((cq ^ (r & 7)) * 4)]);
                        const float x[4] = {x4.x + bv.x, x4.y + bv.y, x4.z + bv.z, x4.w + bv.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (x[e] > best[e]) { best[e] = x[e]; bi[e] = d; }
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        s_val[(4 * cq + e) * 28 + j] = fmaxf(best[e], 0.f);
                        s_idx[(4 * cq + e) * 28 + j] = (uint8_t)bi[e];
                    }
                }
            }
            asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
            const long long kb = (long long)client * p.a.B + b;
            float* o = p.pool_out + kb * (COUT * 49) + half * 28;
            uint8_t* oi = p.pool_idx + kb * (COUT * 49) + half * 28;
            const int nvalid = half ? 21 : 28;                   // pooled rows 4..6 in the lower half (row 7 does not exist)
            for (int e = e0; e < COUT * 28; e += nthr) {
                const int c = e / 28, jj = e - c * 28;
                if (jj < nvalid) {
                    o[c * 49 + jj] = s_val[e];
                    oi[c * 49 + jj] = s_idx[e];
                }
            }
            return;                    // the next tile's first barrier orders its writes to s_x / s_val behind this write-out
        }
        float* z = p.z_all + ((long long)row0 + m) * COUT;
        const bool ok = m < p.a.B * p.g.PP();
        const bool stats = p.bn_acc != nullptr;
        bool counted = false;                    // this row is a real pixel of a live sample
        if (stats) {
            if (client != bclient) { bn_flush(p, lane); bclient = client; }
            if (m < live_rows) {
                const int rr = m % p.g.PP(), h = rr / p.g.Wp;
                counted = h < p.g.H && (rr - h * p.g.Wp) < p.g.W;
            }
        }
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += 32) {
            if ((c0 / 32) % nparts != part) continue;
            float v[32];
            ld_acc(tmem + ((uint32_t)(quarter * 32) << 16) + c0, v);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + c0 + i));
                v[i] += bv.x; v[i + 1] += bv.y; v[i + 2] += bv.z; v[i + 3] += bv.w;
                if (ok) *reinterpret_cast<float4*>(z + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
            if (stats && ROWACC) {
                if (counted) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) { rs0[ROWACC ? i : 0] += v[i]; rs1[ROWACC ? i : 0] = fmaf(v[i], v[i], rs1[ROWACC ? i : 0]); }
                }
            } else if (stats) {
                float sq[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) { v[i] = counted ? v[i] : 0.f; sq[i] = v[i] * v[i]; }
                bs0[c0 / 32] += warp_colsum32(v, lane);
                bs1[c0 / 32] += warp_colsum32(sq, lane);
            }
        }
    }
};

// ---- conv dgrad: D[128 px, CIN] = sum_{tap, cout chunk} dZ[px - shift(tap), 32] * Wt[tap][32 cout][CIN] ---------------
template <int CIN, int COUT, int NPART = 1>
struct ConvDgradT {
    struct Params { CUtensorMap map_dz; CUtensorMap map_w; flb_train_args a; ConvGeom g; float* dx_all; };
    bool lead = false;               // this lane issues the TMA / MMA instructions (skeleton sets it; the rest of the warp runs along)
    static constexpr int ACC_COLS = NPART * CIN;
    static __device__ __forceinline__ void ld_acc(uint32_t taddr, float* v) {
        tmem_ld32(taddr, v);
#pragma unroll
        for (int j = 1; j < NPART; ++j) {
            float u[32];
            tmem_ld32(taddr + j * CIN, u);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += u[i];
        }
    }
    static __host__ __device__ int num_tiles(const Params& p) { return ((p.a.B * p.g.PP() + 127) / 128) * p.a.K; }
    const int* bsz_tab = nullptr;    // per-CTA shared-memory copy of the clients' live batch sizes (persistent skeletons)
    __device__ bool tile_setup(const Params& p, int tile, int& num_kb) {
        const int tpc = (p.a.B * p.g.PP() + 127) / 128;
        client = tile / tpc;
        const int bsz = (bsz_tab && client < BSZ_TAB) ? bsz_tab[client] : flb_bsz(p.a, client);
        m0 = (tile - client * tpc) * 128;
        if (m0 >= bsz * p.g.PP()) return false;
        row0 = client * p.a.B * p.g.PP();
        num_kb = NKB;
        return true;
    }
    __device__ __forceinline__ int tile_place(const Params& p, int client_, int m0_, int) {
        client = client_; m0 = m0_;
        row0 = client * p.a.B * p.g.PP();
        return NKB;
    }
    static constexpr int CH = COUT / 32, NKB = 9 * CH, NCH = CIN / 32, A_BYTES = 128 * 128, B_BYTES = NCH * 4096;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES, STAGES = CIN > 64 ? 3 : 4, RESIDENT_BYTES = 0;
    static constexpr int TMEM_COLS = Pow2Cols<CIN>::value, MINB = 2;
    int client, m0, row0;
    __device__ bool setup(const Params& p, int& num_kb) {
        client = blockIdx.y;
        const int bsz = flb_bsz(p.a, client);
        m0 = blockIdx.x * 128;
        if (m0 >= bsz * p.g.PP()) return false;
        row0 = client * p.a.B * p.g.PP();
        num_kb = NKB;
        return true;
    }
    __device__ void prefetch(const Params& p) { tma_prefetch_desc(&p.map_dz); tma_prefetch_desc(&p.map_w); }
    __device__ void stage_resident(const Params&, uint8_t*, int) {}
    __device__ void load(const Params& p, int kb, uint8_t* stage, uint64_t* bar) {
        const int tap = kb / CH, c = kb % CH;
        if (this->lead) mbar_expect_tx(bar, STAGE_BYTES);
        if (this->lead) tma_load_2d(&p.map_dz, stage, bar, c * 32, row0 + m0 - tap_shift(tap, p.g.Wp));
#pragma unroll
        for (int nc = 0; nc < NCH; ++nc)                        // B is MN-major: rows = cout (K), 32-wide cin (N) chunks
            if (this->lead) tma_load_3d(&p.map_w, stage + A_BYTES + nc * 4096, bar, nc * 32, tap * COUT + c * 32, client);
    }
    __device__ void mma(int kb, uint32_t stage, uint32_t, uint32_t tmem) {
        constexpr uint32_t id = idesc_tf32(128, CIN, false, true);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (this->lead) mma_tf32(tmem, smem_desc(stage + k * 32, 16, 1024), smem_desc_mn(stage + A_BYTES + k * 1024, 4096, 512), id, kb > 0 || k > 0);
    }
    __device__ void epilogue(const Params& p, uint32_t tmem, int quarter, int lane, int part = 0, int nparts = 1) {
        const int m = m0 + quarter * 32 + lane;
        float* dx = p.dx_all + ((long long)row0 + m) * CIN;
        const bool ok = m < p.a.B * p.g.PP();
#pragma unroll 1
        for (int c0 = part * 32; c0 < CIN; c0 += 32 * nparts) {
            float v[32];
            ld_acc(tmem + ((uint32_t)(quarter * 32) << 16) + c0, v);
            if (ok) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dx + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
        }
    }
};

// ---- halo variants of conv fwd / dgrad ---------------------------------------------------------------------------------
// The streamed kernels above re-read every activation row 9 times from L2 (once per tap) and were measured bound by
// L2 -> SM traffic.  Here ONE TMA box of 128 + 2*(Wp+1) rows is staged per 32-channel chunk, and the nine taps are nine
// MMA groups whose A descriptors start (Wp+1 + shift) rows into that box (smem_desc_row).  Activation traffic from L2
// drops 9x -> (128 + 2*halo)/128, and the client's whole weight tensor stays resident in shared memory
// (conv_resident_kernel) instead of riding along with every tile.
constexpr int HALO_A_BYTES = 200 * 128;       // up to 128 + 2*34 rows (33-wide CIFAR grid), padded to a 1024-byte multiple

template <int CIN, int COUT, bool POOL = false>
struct ConvFwdHaloT : ConvFwdT<CIN, COUT, POOL, 1> {         // one accumulator: this kernel is epilogue-bound (measured), partials only add TMEM loads
    using Base = ConvFwdT<CIN, COUT, POOL, 1>;
    using Params = typename Base::Params;
    static constexpr int CH = CIN / 32, W_BYTES = CH * 9 * COUT * 128, STAGE_BYTES = HALO_A_BYTES;
    static constexpr int FIT = (222 * 1024 - W_BYTES) / STAGE_BYTES, STAGES = FIT > 4 ? 4 : FIT;
    static_assert(STAGES >= 2, "resident weights leave no room for the activation pipeline");
    static constexpr int EPI_WARPS = COUT >= 64 ? 8 : 4;         // two epilogue warps per TMEM lane quarter once there are two column chunks
    int wp;
    __device__ __forceinline__ int tile_place(const Params& p, int client_, int m0_, int live_rows_) {
        Base::tile_place(p, client_, m0_, live_rows_);
        wp = p.g.Wp;
        return CH;
    }
    __device__ void load_w(const Params& p, uint8_t* wres, uint64_t* bar) {
        if (this->lead) mbar_expect_tx(bar, W_BYTES);
#pragma unroll 1
        for (int i = 0; i < CH * 9; ++i)                       // tile (chunk c, tap): [COUT rows][32 cin], K-major SW128
            if (this->lead) tma_load_3d(&p.map_w, wres + i * COUT * 128, bar, (i / 9) * 32, (i % 9) * COUT, this->client);
    }
    __device__ void load_a(const Params& p, int kb, uint8_t* stage, uint64_t* bar) {
        const int halo = p.g.Wp + 1;
        if (this->lead) mbar_expect_tx(bar, (128 + 2 * halo) * 128);
        if (this->lead) tma_load_2d(&p.map_x, stage, bar, kb * 32, this->row0 + this->m0 - halo);
    }
    __device__ void mma(int kb, uint32_t stage, uint32_t wres, uint32_t tmem) {
        constexpr uint32_t id = idesc_tf32(128, COUT, false, false);
        // base descriptors once per k-block; every MMA advances them by a (mostly compile-time) byte offset
        const uint64_t a_base = smem_desc_row(stage), b_base = smem_desc(wres + (uint32_t)(kb * 9) * (COUT * 128), 16, 1024);
        const uint32_t row_m = 128u, row_0 = (uint32_t)(wp + 1) * 128u, row_p = (uint32_t)(2 * wp + 1) * 128u;   // box row of tap (r, q = 1): r * wp + 1
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const uint32_t a_off = (tap / 3 == 0 ? row_m : (tap / 3 == 1 ? row_0 : row_p)) + (uint32_t)(tap % 3 - 1) * 128u;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (this->lead) mma_tf32(tmem, desc_advance(a_base, a_off + k * 32), desc_advance(b_base, tap * (COUT * 128) + k * 32), id,
                                         kb > 0 || tap > 0 || k > 0);
        }
    }
};

template <int CIN, int COUT>
struct ConvDgradHaloT : ConvDgradT<CIN, COUT, 3> {
    using Base = ConvDgradT<CIN, COUT, 3>;
    using Params = typename Base::Params;
    static constexpr int CH = COUT / 32, NCH = CIN / 32, W_BYTES = CH * 9 * NCH * 4096, STAGE_BYTES = HALO_A_BYTES;
    static constexpr int FIT = (222 * 1024 - W_BYTES) / STAGE_BYTES, STAGES = FIT > 4 ? 4 : FIT;
    static_assert(STAGES >= 2, "resident weights leave no room for the activation pipeline");
    static constexpr int EPI_WARPS = CIN >= 64 ? 8 : 4;
    int wp;
    __device__ __forceinline__ int tile_place(const Params& p, int client_, int m0_, int live_rows_) {
        Base::tile_place(p, client_, m0_, live_rows_);
        wp = p.g.Wp;
        return CH;
    }
    __device__ void load_w(const Params& p, uint8_t* wres, uint64_t* bar) {
        if (this->lead) mbar_expect_tx(bar, W_BYTES);
#pragma unroll 1
        for (int i = 0; i < CH * 9 * NCH; ++i) {               // tile (cout chunk c, tap, cin chunk nc): MN-major, [32 cout rows][32 cin]
            const int nc = i % NCH, tap = (i / NCH) % 9, c = i / (NCH * 9);
            if (this->lead) tma_load_3d(&p.map_w, wres + i * 4096, bar, nc * 32, tap * COUT + c * 32, this->client);
        }
    }
    __device__ void load_a(const Params& p, int kb, uint8_t* stage, uint64_t* bar) {
        const int halo = p.g.Wp + 1;
        if (this->lead) mbar_expect_tx(bar, (128 + 2 * halo) * 128);
        if (this->lead) tma_load_2d(&p.map_dz, stage, bar, kb * 32, this->row0 + this->m0 - halo);
    }
    __device__ void mma(int kb, uint32_t stage, uint32_t wres, uint32_t tmem) {
        constexpr uint32_t id = idesc_tf32(128, CIN, false, true);
        const uint64_t a_base = smem_desc_row(stage), b_base = smem_desc_mn(wres + (uint32_t)(kb * 9 * NCH) * 4096u, 4096, 512);
        // box row of tap (r, q): (wp + 1) - ((r - 1) * wp + (q - 1)) = (2 - r) * wp + 2 - q
        const uint32_t row_r0 = (uint32_t)(2 * wp + 1) * 128u, row_r1 = (uint32_t)(wp + 1) * 128u, row_r2 = 128u;      // q = 1
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const uint32_t a_off = (tap / 3 == 0 ? row_r0 : (tap / 3 == 1 ? row_r1 : row_r2)) - (uint32_t)(tap % 3 - 1) * 128u;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (this->lead) mma_tf32(tmem + (tap / 3) * CIN, desc_advance(a_base, a_off + k * 32), desc_advance(b_base, tap * NCH * 4096 + k * 1024), id,
                                         kb > 0 || (tap % 3) > 0 || k > 0);
        }
    }
};

// ---- conv wgrad: Gt[tap][co][ci] += sum_px X[px + shift(tap)][ci] * dZ[px][co] ----------------------------------------
// M = (tap, ci) in 32-row chunks, N = co, K = pixels (split over blockIdx.x); blockIdx.z selects MTC of the M tiles.
// NORM (north-star kernel 2, per-sample DP-SGD): one split per SAMPLE (kb_per_split = rows per image / 32), and the
// epilogue does not write the [9*Cin, Cout] per-sample weight gradient anywhere -- it squares it straight out of TMEM,
// reduces with warp shuffles and adds one float per warp to norm2[client, sample].
template <int CIN, int COUT, int MTC, bool NORM = false>
struct ConvWgradT {
    struct Params { CUtensorMap map_x; CUtensorMap map_dz; flb_train_args a; ConvGeom g; float* gt_all; long long ldt; int kb_per_split;
                    float* norm2_all; };
    bool lead = false;               // this lane issues the TMA / MMA instructions (skeleton sets it; the rest of the warp runs along)
    static constexpr int CCH = CIN / 32, ACH = 9 * CCH, BCH = COUT / 32;
    static constexpr int A_BYTES = MTC * 4 * 4096, STAGE_BYTES = A_BYTES + BCH * 4096;
    static constexpr int STAGES = STAGE_BYTES * 3 <= 200 * 1024 ? 3 : 2, RESIDENT_BYTES = 0;
    static constexpr int TMEM_COLS = Pow2Cols<MTC * COUT>::value, MINB = 1;
    static_assert(MTC * COUT <= 512, "accumulators must fit TMEM");
    int client, row0, kb0, chunk0, nch, total_rows;
    __device__ bool setup(const Params& p, int& num_kb) {
        client = blockIdx.y;
        const int bsz = flb_bsz(p.a, client);
        total_rows = bsz * p.g.PP();
        const int total = (total_rows + 31) / 32;
        kb0 = blockIdx.x * p.kb_per_split;
        if (kb0 >= total) return false;
        num_kb = min(p.kb_per_split, total - kb0);
        row0 = client * p.a.B * p.g.PP();
        chunk0 = blockIdx.z * MTC * 4;
        nch = min(MTC * 4, ACH - chunk0);
        return nch > 0;
    }
    __device__ void prefetch(const Params& p) { tma_prefetch_desc(&p.map_x); tma_prefetch_desc(&p.map_dz); }
    __device__ void stage_resident(const Params&, uint8_t*, int) {}
    __device__ void load(const Params& p, int kb, uint8_t* stage, uint64_t* bar) {
        const int px = row0 + (kb0 + kb) * 32;
        if (this->lead) mbar_expect_tx(bar, (nch + BCH) * 4096);
#pragma unroll 1
        for (int ch = 0; ch < nch; ++ch) {                       // chunk = (tap, 32-channel slice of Cin)
            const int gc = chunk0 + ch, tap = gc / CCH, c = gc % CCH;
            if (this->lead) tma_load_2d(&p.map_x, stage + ch * 4096, bar, c * 32, px + tap_shift(tap, p.g.Wp));
        }
#pragma unroll
        for (int c = 0; c < BCH; ++c) if (this->lead) tma_load_2d(&p.map_dz, stage + A_BYTES + c * 4096, bar, c * 32, px);
    }
    __device__ void mma(int kb, uint32_t stage, uint32_t, uint32_t tmem) {
        constexpr uint32_t id = idesc_tf32(128, COUT, true, true);
        const int ksteps = min(4, (total_rows - (kb0 + kb) * 32) >> 3);      // rows per image are a multiple of 8
        const int mtiles = (nch + 3) >> 2;
        for (int k = 0; k < ksteps; ++k)
            for (int mt = 0; mt < mtiles; ++mt)
                if (this->lead) mma_tf32(tmem + mt * COUT, smem_desc_mn(stage + mt * 4 * 4096 + k * 1024, 4096, 512),
                         smem_desc_mn(stage + A_BYTES + k * 1024, 4096, 512), id, kb > 0 || k > 0);
    }
    __device__ void epilogue(const Params& p, uint32_t tmem, int quarter, int lane) {
        const int mtiles = (nch + 3) >> 2;
        if (NORM) {
            float sq = 0.f;
#pragma unroll 1
            for (int mt = 0; mt < mtiles; ++mt) {
                const bool ok = ((mt * 128 + quarter * 32 + lane) >> 5) < nch;
#pragma unroll 1
                for (int c0 = 0; c0 < COUT; c0 += 32) {
                    float v[32];
                    tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + mt * COUT + c0, v);
                    if (ok) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) sq = fmaf(v[i], v[i], sq);
                    }
                }
            }
            sq = flb_warp_sum(sq);
            if (lane == 0 && sq != 0.f) atomicAdd(&p.norm2_all[(long long)client * p.a.B + blockIdx.x], sq);     // blockIdx.x = sample
            return;
        }
        float* gt = p.gt_all + (long long)client * p.ldt;
#pragma unroll 1
        for (int mt = 0; mt < mtiles; ++mt) {
            const int row = mt * 128 + quarter * 32 + lane;          // row = local chunk * 32 + (cin within the slice)
            const int ch = row >> 5, gc = chunk0 + ch, tap = gc / CCH, ci = (gc % CCH) * 32 + (row & 31);
            const bool ok = ch < nch;
#pragma unroll 1
            for (int c0 = 0; c0 < COUT; c0 += 32) {
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + mt * COUT + c0, v);
                if (ok) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) atomicAdd(&gt[((long long)tap * COUT + c0 + i) * CIN + ci], v[i]);   // lanes = consecutive ci
                }
            }
        }
    }
};

// ---- conv wgrad, halo variant ---------------------------------------------------------------------------------------------
// The kernel above fetches nine shifted copies of every activation row (one TMA box per tap).  Here ONE box of
// 32 + 2*(Wp+1) pixel rows is staged per 32-channel slice and k-block, and the M dimension is ordered (kernel row r,
// slice c, tap-in-row j): for fixed (r, c) the three taps are the same rows shifted by one pixel, i.e. three 32-row M
// chunks whose start addresses are 128 bytes apart -- a single MN-major descriptor with LBO = 128 (chunks overlap in
// memory; the fourth chunk of the 128-row MMA tile is a don't-care).  L2 -> SM traffic per k-block drops from
// 9 * 4 KB to (32 + 2*halo) * 128 B per slice.  RPC kernel rows per CTA (3, or 1 when 3 * Cin/32 * Cout > 512 TMEM columns).
constexpr int WG_BOX_BYTES = 13 * 1024;       // up to 32 + 2*34 rows of 128 B, rounded up to a 1024-byte multiple

template <int CIN, int COUT, int RPC, bool NORM = false>
struct ConvWgradHaloT {
    struct Params { CUtensorMap map_x; CUtensorMap map_dz; flb_train_args a; ConvGeom g; float* gt_all; long long ldt; int kb_per_split;
                    float* norm2_all; };
    bool lead = false;
    static constexpr int CCH = CIN / 32, BCH = COUT / 32, TILES = RPC * CCH;
    static constexpr int A_BYTES = CCH * WG_BOX_BYTES, STAGE_BYTES = A_BYTES + BCH * 4096;
    static constexpr int STAGES = STAGE_BYTES * 4 <= 200 * 1024 ? 4 : (STAGE_BYTES * 3 <= 200 * 1024 ? 3 : 2), RESIDENT_BYTES = 0;
    static constexpr int TMEM_COLS = Pow2Cols<TILES * COUT>::value, MINB = 1;
    static_assert(TILES * COUT <= 512, "accumulators must fit TMEM");
    int client, row0, row_off, rows_end, rbase, wp;      // this CTA reduces over rows [row_off, rows_end) of its client
    __device__ bool setup(const Params& p, int& num_kb) {
        client = blockIdx.y;
        const int bsz = flb_bsz(p.a, client);
        if (NORM) {                              // one split per SAMPLE: rows per image are a multiple of 8, not of 32
            if ((int)blockIdx.x >= bsz) return false;
            row_off = blockIdx.x * p.g.PP();
            rows_end = row_off + p.g.PP();
            num_kb = (p.g.PP() + 31) / 32;
        } else {
            rows_end = bsz * p.g.PP();
            const int total = (rows_end + 31) / 32;
            const int kb0 = blockIdx.x * p.kb_per_split;
            if (kb0 >= total) return false;
            num_kb = min(p.kb_per_split, total - kb0);
            row_off = kb0 * 32;
        }
        row0 = client * p.a.B * p.g.PP();
        rbase = blockIdx.z * RPC;
        wp = p.g.Wp;
        return true;
    }
    __device__ void prefetch(const Params& p) { tma_prefetch_desc(&p.map_x); tma_prefetch_desc(&p.map_dz); }
    __device__ void stage_resident(const Params&, uint8_t*, int) {}
    __device__ void load(const Params& p, int kb, uint8_t* stage, uint64_t* bar) {
        const int px = row0 + row_off + kb * 32, halo = p.g.Wp + 1;
        if (this->lead) mbar_expect_tx(bar, CCH * (32 + 2 * halo) * 128 + BCH * 4096);
#pragma unroll
        for (int c = 0; c < CCH; ++c) if (this->lead) tma_load_2d(&p.map_x, stage + c * WG_BOX_BYTES, bar, c * 32, px - halo);
#pragma unroll
        for (int c = 0; c < BCH; ++c) if (this->lead) tma_load_2d(&p.map_dz, stage + A_BYTES + c * 4096, bar, c * 32, px);
    }
    __device__ void mma(int kb, uint32_t stage, uint32_t, uint32_t tmem) {
        constexpr uint32_t id = idesc_tf32(128, COUT, true, true);
        const int ksteps = min(4, (rows_end - (row_off + kb * 32)) >> 3);    // rows per image are a multiple of 8
        // base descriptors once per k-block, advanced per MMA (tc_gemm.cuh desc_advance)
        const uint64_t a_base = smem_desc_mn(stage + (uint32_t)(rbase * wp) * 128u, 128, 512), b_base = smem_desc_mn(stage + A_BYTES, 4096, 512);
        const uint32_t row_step = (uint32_t)wp * 128u;
        for (int k = 0; k < ksteps; ++k)
#pragma unroll
            for (int rl = 0; rl < RPC; ++rl)
#pragma unroll
                for (int c = 0; c < CCH; ++c) {
                    // box row of tap (r, j = 0) for pixel 0 of the k-block: halo + shift = (Wp + 1) + (r - 1) * Wp - 1 = r * Wp
                    if (this->lead) mma_tf32(tmem + (rl * CCH + c) * COUT, desc_advance(a_base, c * WG_BOX_BYTES + rl * row_step + k * 1024),
                                             desc_advance(b_base, k * 1024), id, kb > 0 || k > 0);
                }
    }
    __device__ void epilogue(const Params& p, uint32_t tmem, int quarter, int lane) {
        // accumulator row = (tap in kernel row j = quarter, cin within the slice = lane); quarter 3 is the don't-care chunk
        if (NORM) {
            float sq = 0.f;
            if (quarter < 3) {
#pragma unroll 1
                for (int tl = 0; tl < TILES; ++tl)
#pragma unroll 1
                    for (int c0 = 0; c0 < COUT; c0 += 32) {
                        float v[32];
                        tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + tl * COUT + c0, v);
#pragma unroll
                        for (int i = 0; i < 32; ++i) sq = fmaf(v[i], v[i], sq);
                    }
            }
            sq = flb_warp_sum(sq);
            if (lane == 0 && sq != 0.f) atomicAdd(&p.norm2_all[(long long)client * p.a.B + blockIdx.x], sq);     // blockIdx.x = sample
            return;
        }
        if (quarter >= 3) return;
        float* gt = p.gt_all + (long long)client * p.ldt;
#pragma unroll 1
        for (int tl = 0; tl < TILES; ++tl) {
            const int r = rbase + tl / CCH, c = tl % CCH, tap = 3 * r + quarter, ci = c * 32 + lane;
#pragma unroll 1
            for (int c0 = 0; c0 < COUT; c0 += 32) {
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + tl * COUT + c0, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) atomicAdd(&gt[((long long)tap * COUT + c0 + i) * CIN + ci], v[i]);   // lanes = consecutive ci
            }
        }
    }
};

// ---- linear forward (swap-AB, split-K): D[128 out, B] += W[out, k] * act[b, k] --------------------------------------
template <int IN, int OUT>
struct FcFwdT {
    struct Params { CUtensorMap map_w; CUtensorMap map_act; flb_train_args a; float* out_all; int kb_per_split; };
    bool lead = false;               // this lane issues the TMA / MMA instructions (skeleton sets it; the rest of the warp runs along)
    static_assert(OUT % 128 == 0, "whole 128-row accumulator tiles");
    static constexpr int STAGES = 6, STAGE_BYTES = 128 * 128 + 32 * 128, RESIDENT_BYTES = 0, TMEM_COLS = 32, MINB = 1;
    int client, kb0, bsz, mt;
    __device__ bool setup(const Params& p, int& num_kb) {
        client = blockIdx.y;
        mt = blockIdx.z;
        bsz = flb_bsz(p.a, client);
        if (bsz == 0) return false;
        constexpr int total = (IN + 31) / 32;
        kb0 = blockIdx.x * p.kb_per_split;
        if (kb0 >= total) return false;
        num_kb = min(p.kb_per_split, total - kb0);
        return true;
    }
    __device__ void prefetch(const Params& p) { tma_prefetch_desc(&p.map_w); tma_prefetch_desc(&p.map_act); }
    __device__ void stage_resident(const Params&, uint8_t*, int) {}
    __device__ void load(const Params& p, int kb, uint8_t* stage, uint64_t* bar) {
        const int k0 = (kb0 + kb) * 32;
        if (this->lead) mbar_expect_tx(bar, STAGE_BYTES);
        if (this->lead) tma_load_3d(&p.map_w, stage, bar, k0, mt * 128, client);
        if (this->lead) tma_load_2d(&p.map_act, stage + 128 * 128, bar, k0, client * p.a.B);
    }
    __device__ void mma(int kb, uint32_t stage, uint32_t, uint32_t tmem) {
        constexpr uint32_t id = idesc_tf32(128, 32, false, false);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (this->lead) mma_tf32(tmem, smem_desc(stage + k * 32, 16, 1024), smem_desc(stage + 128 * 128 + k * 32, 16, 1024), id, kb > 0 || k > 0);
    }
    __device__ void epilogue(const Params& p, uint32_t tmem, int quarter, int lane) {
        const int j = mt * 128 + quarter * 32 + lane;
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16), v);
        float* out = p.out_all + (long long)client * p.a.B * OUT + j;
#pragma unroll
        for (int b = 0; b < 32; ++b)
            if (b < bsz) atomicAdd(out + b * OUT, v[b]);
    }
};

// ---- linear dgrad: D[128 in, B] = sum_out W[out, in] * dout[b, out] ---------------------------------------------------
template <int IN, int OUT>
struct FcDgradT {
    struct Params { CUtensorMap map_w; CUtensorMap map_dout; flb_train_args a; float* dact_all; };
    bool lead = false;               // this lane issues the TMA / MMA instructions (skeleton sets it; the rest of the warp runs along)
    static constexpr int STAGES = 4, STAGE_BYTES = 4 * 4096 + 4096, RESIDENT_BYTES = 0, TMEM_COLS = 32, MINB = 2;
    int client, m0, bsz;
    __device__ bool setup(const Params& p, int& num_kb) {
        client = blockIdx.y;
        bsz = flb_bsz(p.a, client);
        if (bsz == 0) return false;
        m0 = blockIdx.x * 128;
        num_kb = OUT / 32;
        return true;
    }
    __device__ void prefetch(const Params& p) { tma_prefetch_desc(&p.map_w); tma_prefetch_desc(&p.map_dout); }
    __device__ void stage_resident(const Params&, uint8_t*, int) {}
    __device__ void load(const Params& p, int kb, uint8_t* stage, uint64_t* bar) {
        if (this->lead) mbar_expect_tx(bar, STAGE_BYTES);
#pragma unroll
        for (int c = 0; c < 4; ++c) if (this->lead) tma_load_3d(&p.map_w, stage + c * 4096, bar, m0 + 32 * c, kb * 32, client);   // rows = out features (K)
        if (this->lead) tma_load_2d(&p.map_dout, stage + 4 * 4096, bar, kb * 32, client * p.a.B);
    }
    __device__ void mma(int kb, uint32_t stage, uint32_t, uint32_t tmem) {
        constexpr uint32_t id = idesc_tf32(128, 32, true, false);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (this->lead) mma_tf32(tmem, smem_desc_mn(stage + k * 1024, 4096, 512), smem_desc(stage + 4 * 4096 + k * 32, 16, 1024), id, kb > 0 || k > 0);
    }
    __device__ void epilogue(const Params& p, uint32_t tmem, int quarter, int lane) {
        const int m = m0 + quarter * 32 + lane;
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16), v);
        if (m < IN) {
            float* d = p.dact_all + (long long)client * p.a.B * IN + m;
#pragma unroll
            for (int b = 0; b < 32; ++b)
                if (b < bsz) d[(long long)b * IN] = v[b];
        }
    }
};

// ---- linear wgrad: D[128 out, 256 in] = sum_b dout[b, out] * act[b, in] ------------------------------------------------
template <int IN, int OUT>
struct FcWgradT {
    struct Params { CUtensorMap map_dout; CUtensorMap map_act; flb_train_args a; int woff; };
    bool lead = false;               // this lane issues the TMA / MMA instructions (skeleton sets it; the rest of the warp runs along)
    static_assert(OUT % 128 == 0, "whole 128-row accumulator tiles");
    static constexpr int NT = 256;
    static constexpr int STAGES = 1, STAGE_BYTES = 4 * 4096 + 8 * 4096, RESIDENT_BYTES = 0, TMEM_COLS = 256, MINB = 1;
    int client, n0, ksteps, mt;
    __device__ bool setup(const Params& p, int& num_kb) {
        client = blockIdx.y;
        mt = blockIdx.z;
        if (flb_bsz(p.a, client) == 0) return false;
        n0 = blockIdx.x * 256;
        ksteps = p.a.B / 8;              // rows bsz..B-1 of dout are zero (producer kernels); rows >= B belong to the next client
        num_kb = 1;
        return true;
    }
    __device__ void prefetch(const Params& p) { tma_prefetch_desc(&p.map_dout); tma_prefetch_desc(&p.map_act); }
    __device__ void stage_resident(const Params&, uint8_t*, int) {}
    __device__ void load(const Params& p, int, uint8_t* stage, uint64_t* bar) {
        if (this->lead) mbar_expect_tx(bar, STAGE_BYTES);
#pragma unroll
        for (int c = 0; c < 4; ++c) if (this->lead) tma_load_2d(&p.map_dout, stage + c * 4096, bar, mt * 128 + 32 * c, client * p.a.B);
#pragma unroll
        for (int c = 0; c < 8; ++c) if (this->lead) tma_load_2d(&p.map_act, stage + 4 * 4096 + c * 4096, bar, n0 + 32 * c, client * p.a.B);
    }
    __device__ void mma(int, uint32_t stage, uint32_t, uint32_t tmem) {
        constexpr uint32_t id = idesc_tf32(128, 256, true, true);
        for (int k = 0; k < ksteps; ++k)
            if (this->lead) mma_tf32(tmem, smem_desc_mn(stage + k * 1024, 4096, 512), smem_desc_mn(stage + 4 * 4096 + k * 1024, 4096, 512), id, k > 0);
    }
    __device__ void epilogue(const Params& p, uint32_t tmem, int quarter, int lane) {
        const int j = mt * 128 + quarter * 32 + lane;
        float* g = p.a.G + (long long)client * p.a.ld + p.woff + (long long)j * IN;
#pragma unroll 1
        for (int c0 = 0; c0 < 256; c0 += 32) {
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + c0, v);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const int n = n0 + c0 + i;
                if (n < IN) *reinterpret_cast<float4*>(g + n) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
        }
    }
};

// ---- linear wgrad, transposed accumulator: D[128 in, 128 out] = sum_b act[b, in] * dout[b, out] ---------------------------
// Same operands as above with the MMA roles swapped, so that a TMEM LANE is an input feature: for a fixed output feature j
// the 32 lanes of a warp touch W[j][n .. n+32) -- one 128-byte line per warp instruction.  (With lanes = output features
// every lane sits in a different 12.5 KB row: 32 LSU wavefronts per instruction; measured 49 us for the fused-optimizer
// epilogue in that orientation, LSU-bound, regardless of how many loads were kept in flight.)
// ADAM: the epilogue does not store the gradient -- it applies the optimizer step (opt_update.cuh: Adam / SGD-momentum /
// AdamW by args.opt) to W, M, V straight out of TMEM.  For SimpleCNN's fc1.weight (401 408 of the 421 642 parameters) this
// removes the gradient's memory round trip (4 B written + 4 B read per parameter) and takes 95 % of the optimizer's bytes
// off the step's critical path: the stand-alone optimizer kernel skips the range (TcConvTab::skip_*).  The caller orders
// this kernel after the layer's dgrad, which still reads the old W.
template <int IN, bool ADAM>
struct FcWgradSwapT {
    struct Params { CUtensorMap map_dout; CUtensorMap map_act; flb_train_args a; int woff; };
    bool lead = false;
    static constexpr int OUT = 128;      // output features per CTA (blockIdx.z walks the layer's 128-feature groups)
    static constexpr int STAGES = 1, STAGE_BYTES = 8 * 4096, RESIDENT_BYTES = 0, TMEM_COLS = 128, MINB = 2;
    int client, n0, ksteps, j0;
    __device__ bool setup(const Params& p, int& num_kb) {
        client = blockIdx.y;
        j0 = blockIdx.z * OUT;
        if (flb_bsz(p.a, client) == 0) return false;
        n0 = blockIdx.x * 128;           // input-feature tile (the last one is partial: TMA zero-fills the columns past IN)
        ksteps = p.a.B / 8;
        num_kb = 1;
        return true;
    }
    __device__ void prefetch(const Params& p) { tma_prefetch_desc(&p.map_dout); tma_prefetch_desc(&p.map_act); }
    __device__ void stage_resident(const Params&, uint8_t*, int) {}
    __device__ void load(const Params& p, int, uint8_t* stage, uint64_t* bar) {
        if (this->lead) mbar_expect_tx(bar, STAGE_BYTES);
#pragma unroll
        for (int c = 0; c < 4; ++c) if (this->lead) tma_load_2d(&p.map_act, stage + c * 4096, bar, n0 + 32 * c, client * p.a.B);
#pragma unroll
        for (int c = 0; c < 4; ++c) if (this->lead) tma_load_2d(&p.map_dout, stage + 4 * 4096 + c * 4096, bar, j0 + 32 * c, client * p.a.B);
    }
    __device__ void mma(int, uint32_t stage, uint32_t, uint32_t tmem) {
        constexpr uint32_t id = idesc_tf32(128, 128, true, true);
        for (int k = 0; k < ksteps; ++k)
            if (this->lead) mma_tf32(tmem, smem_desc_mn(stage + k * 1024, 4096, 512), smem_desc_mn(stage + 4 * 4096 + k * 1024, 4096, 512), id, k > 0);
    }
    template <int OPT>
    __device__ __forceinline__ void epilogue_opt(const Params& p, uint32_t taddr, long long off, bool live) {
        const int t = p.a.tcount[client] + 1;                 // the stand-alone optimizer kernel advances tcount later in the step
        const OptScalars c = opt_scalars(p.a, t, flb_bsz(p.a, client));
        const bool need_m = OPT != 1 || t > 1;
        float* __restrict__ W = p.a.W + off; float* __restrict__ M = p.a.M + off; float* __restrict__ V = p.a.V + off;
#pragma unroll 1
        for (int c0 = 0; c0 < OUT; c0 += 32) {
            float g[32], w[32], m[32], v[32];
            if (live) {                                       // the chunk's 96 loads are in flight before the first update
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const long long e = (long long)(c0 + i) * IN;
                    w[i] = W[e];
                    m[i] = need_m ? M[e] : 0.f;
                    v[i] = OPT != 1 ? V[e] : 0.f;
                }
            }
            tmem_ld32(taddr + c0, g);
            if (live) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const long long e = (long long)(c0 + i) * IN;
                    opt_update<OPT, false>(c, g[i], 0.f, w[i], m[i], v[i]);
                    W[e] = w[i];
                    M[e] = m[i];
                    if (OPT != 1) V[e] = v[i];
                }
            }
        }
    }
    __device__ void epilogue(const Params& p, uint32_t tmem, int quarter, int lane) {
        const int n = n0 + quarter * 32 + lane;               // this thread's input feature
        const bool live = n < IN;
        const long long off = (long long)client * p.a.ld + p.woff + (long long)j0 * IN + n;
        const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16);
        if (ADAM) {
            if (p.a.opt == 0) epilogue_opt<0>(p, taddr, off, live);
            else if (p.a.opt == 1) epilogue_opt<1>(p, taddr, off, live);
            else epilogue_opt<2>(p, taddr, off, live);
            return;
        }
        float* __restrict__ g = p.a.G + off;
#pragma unroll 1
        for (int c0 = 0; c0 < OUT; c0 += 32) {
            float v[32];
            tmem_ld32(taddr + c0, v);
            if (live) {
#pragma unroll
                for (int i = 0; i < 32; ++i) g[(long long)(c0 + i) * IN] = v[i];
            }
        }
    }
};

template <class T>
static int launch(const typename T::Params& p, dim3 grid, cudaStream_t st) {
    constexpr size_t smem = (size_t)T::STAGES * T::STAGE_BYTES + T::RESIDENT_BYTES + 1024;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    if (int rc = ensure_smem_attr(reinterpret_cast<const void*>(&gemm_kernel<T>), (int)smem)) return rc;
    gemm_kernel<T><<<grid, THREADS, smem, st>>>(p);
    return FLB_OK;
}

template <class T>
static int launch_persistent(const typename T::Params& p, cudaStream_t st) {
    constexpr size_t smem = (size_t)T::STAGES * T::STAGE_BYTES + 1024;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    if (int rc = ensure_smem_attr(reinterpret_cast<const void*>(&gemm_persistent_kernel<T>), (int)smem)) return rc;
    const int tiles = T::num_tiles(p);
    const int grid = tiles < flb_num_sms() * T::MINB ? tiles : flb_num_sms() * T::MINB;
    gemm_persistent_kernel<T><<<grid, THREADS, smem, st>>>(p);
    return FLB_OK;
}

template <class T>
static int launch_resident(const typename T::Params& p, cudaStream_t st) {
    constexpr size_t smem = (size_t)T::W_BYTES + (size_t)T::STAGES * T::STAGE_BYTES + 1024;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    if (int rc = ensure_smem_attr(reinterpret_cast<const void*>(&conv_resident_kernel<T>), (int)smem)) return rc;
    const int tiles = T::num_tiles(p);
    const int grid = tiles < flb_num_sms() ? tiles : flb_num_sms();
    conv_resident_kernel<T><<<grid, 64 + 32 * T::EPI_WARPS, smem, st>>>(p);
    return FLB_OK;
}

// tap-major conv weights of all clients as a 3-D tensor {Cin, 9*Cout, K}; box_rows x 32 boxes
static int make_wt_map(CUtensorMap* m, const float* wt, long long ldt, int cin, int cout, int K, uint32_t box_rows, bool mn_major) {
    const uint64_t dims[3] = {(uint64_t)cin, (uint64_t)9 * cout, (uint64_t)K};
    const uint64_t strides[2] = {(uint64_t)cin * sizeof(float), (uint64_t)ldt * sizeof(float)};
    const uint32_t box[3] = {32, box_rows, 1};
    return make_map(m, wt, 3, dims, strides, box, mn_major);
}

template <int CIN, int COUT>
static int conv_fwd_t(const flb_train_args& a, const ConvGeom& g, const float* xin, float* z, const float* wt, long long ldt, int boff,
                      double* bn_acc, int bn_coff, int bn_stride, cudaStream_t st) {
    constexpr bool HALO = CIN * COUT * 36 <= 150 * 1024;          // the whole weight tensor stays resident in shared memory
    using T = typename std::conditional<HALO, ConvFwdHaloT<(HALO ? CIN : 32), (HALO ? COUT : 32)>, ConvFwdT<CIN, COUT>>::type;
    typename T::Params p;
    if (int rc = make_map_2d(&p.map_x, xin, (uint64_t)a.K * a.B * g.PP(), CIN, HALO ? 128 + 2 * (g.Wp + 1) : 128)) return rc;
    if (int rc = make_wt_map(&p.map_w, wt, ldt, CIN, COUT, a.K, COUT, false)) return rc;
    p.a = a; p.g = g; p.z_all = z; p.boff = boff; p.pool_out = nullptr; p.pool_idx = nullptr;
    p.bn_acc = HALO ? bn_acc : nullptr; p.bn_coff = bn_coff; p.bn_stride = bn_stride;
    if constexpr (HALO) return launch_resident<T>(p, st);
    else return launch_persistent<T>(p, st);
}
// SimpleCNN conv2 with the fused bias + ReLU + max-pool epilogue (16-wide grid, 256 rows per image)
int conv_fwd_pool_32_64(const flb_train_args& a, const ConvGeom& g, const float* xin, float* pooled, uint8_t* idx, const float* wt,
                        long long ldt, int boff, cudaStream_t st) {
    using T = ConvFwdHaloT<32, 64, true>;
    if (g.Wp != 16 || g.PP() != 256 || g.Cin != 32 || g.Cout != 64) { flb_set_error("conv_fwd_pool_32_64: geometry"); return FLB_ERR_ARG; }
    typename T::Params p;
    if (int rc = make_map_2d(&p.map_x, xin, (uint64_t)a.K * a.B * g.PP(), 32, 128 + 2 * (g.Wp + 1))) return rc;
    if (int rc = make_wt_map(&p.map_w, wt, ldt, 32, 64, a.K, 64, false)) return rc;
    p.a = a; p.g = g; p.z_all = nullptr; p.boff = boff; p.pool_out = pooled; p.pool_idx = idx;
    p.bn_acc = nullptr; p.bn_coff = 0; p.bn_stride = 0;
    return launch_resident<T>(p, st);
}
template <int CIN, int COUT>
static int conv_dgrad_t(const flb_train_args& a, const ConvGeom& g, const float* dz, float* dx, const float* wt, long long ldt, cudaStream_t st) {
    constexpr bool HALO = CIN * COUT * 36 <= 150 * 1024;
    using T = typename std::conditional<HALO, ConvDgradHaloT<(HALO ? CIN : 32), (HALO ? COUT : 32)>, ConvDgradT<CIN, COUT>>::type;
    typename T::Params p;
    if (int rc = make_map_2d(&p.map_dz, dz, (uint64_t)a.K * a.B * g.PP(), COUT, HALO ? 128 + 2 * (g.Wp + 1) : 128)) return rc;
    if (int rc = make_wt_map(&p.map_w, wt, ldt, CIN, COUT, a.K, 32, true)) return rc;
    p.a = a; p.g = g; p.dx_all = dx;
    if constexpr (HALO) return launch_resident<T>(p, st);
    else return launch_persistent<T>(p, st);
}
template <int CIN, int COUT, int MTC>
static int conv_wgrad_t(const flb_train_args& a, const ConvGeom& g, const float* xin, const float* dz, float* gt, long long ldt, cudaStream_t st) {
    using T = ConvWgradT<CIN, COUT, MTC>;
    typename T::Params p;
    if (int rc = make_map_2d(&p.map_x, xin, (uint64_t)a.K * a.B * g.PP(), CIN, 32, true)) return rc;
    if (int rc = make_map_2d(&p.map_dz, dz, (uint64_t)a.K * a.B * g.PP(), COUT, 32, true)) return rc;
    p.a = a; p.g = g; p.gt_all = gt; p.ldt = ldt; p.norm2_all = nullptr;
    constexpr int groups = (T::ACH + MTC * 4 - 1) / (MTC * 4);
    const int total = (a.B * g.PP() + 31) / 32;
    int splits = flb_num_sms() / (a.K * groups);             // one wave of CTAs over the GPU (two waves measured slower)
    splits = splits < 1 ? 1 : (splits > total ? total : splits);
    p.kb_per_split = (total + splits - 1) / splits;
    splits = (total + p.kb_per_split - 1) / p.kb_per_split;
    return launch<T>(p, dim3(splits, a.K, groups), st);
}

// per-sample squared norms of the conv weight gradient (bias excluded): norm2[client, b] += || dW_b ||^2.
// grid (sample, client, kernel-row group): finished samples drop out in setup(); the groups of a sample meet in norm2.
template <int CIN, int COUT, int RPC>
static int conv_wgrad_norm_t(const flb_train_args& a, const ConvGeom& g, const float* xin, const float* dz, float* norm2, cudaStream_t st) {
    using T = ConvWgradHaloT<CIN, COUT, RPC, true>;
    if (g.PP() % 8 || g.Wp + 1 > 35) { flb_set_error("conv_wgrad_norm: rows per image must be a multiple of 8, grid width <= 34"); return FLB_ERR_ARG; }
    typename T::Params p;
    if (int rc = make_map_2d(&p.map_x, xin, (uint64_t)a.K * a.B * g.PP(), CIN, 32 + 2 * (g.Wp + 1), true)) return rc;
    if (int rc = make_map_2d(&p.map_dz, dz, (uint64_t)a.K * a.B * g.PP(), COUT, 32, true)) return rc;
    p.a = a; p.g = g; p.gt_all = nullptr; p.ldt = 0; p.norm2_all = norm2;
    p.kb_per_split = (g.PP() + 31) / 32;
    return launch<T>(p, dim3(a.B, a.K, 3 / RPC), st);
}
int conv_wgrad_norm(const flb_train_args& a, const ConvGeom& g, const float* xin, const float* dz, float* norm2, cudaStream_t st) {
    if (g.Cin == 32 && g.Cout == 32) return conv_wgrad_norm_t<32, 32, 3>(a, g, xin, dz, norm2, st);
    if (g.Cin == 32 && g.Cout == 64) return conv_wgrad_norm_t<32, 64, 3>(a, g, xin, dz, norm2, st);
    if (g.Cin == 64 && g.Cout == 64) return conv_wgrad_norm_t<64, 64, 3>(a, g, xin, dz, norm2, st);
    if (g.Cin == 64 && g.Cout == 128) return conv_wgrad_norm_t<64, 128, 1>(a, g, xin, dz, norm2, st);
    if (g.Cin == 128 && g.Cout == 128) return conv_wgrad_norm_t<128, 128, 1>(a, g, xin, dz, norm2, st);
    flb_set_error("tensor-core conv wgrad norm: unsupported channels %d -> %d", g.Cin, g.Cout);
    return FLB_ERR_UNSUPPORTED;
}
int conv_wgrad_norm_32_64(const flb_train_args& a, const ConvGeom& g, const float* xin, const float* dz, float* norm2, cudaStream_t st) {
    return conv_wgrad_norm(a, g, xin, dz, norm2, st);
}

// ---- host entry points used by the step orchestrators ---------------------------------------------------------------------
bool conv_supported(int cin, int cout) {
    return (cin == 32 && (cout == 32 || cout == 64)) || (cin == 64 && (cout == 64 || cout == 128)) || (cin == 128 && cout == 128);
}
#define FLB_CONV_DISPATCH(FN, ...)                                                            \
    if (g.Cin == 32 && g.Cout == 32) return FN<32, 32>(__VA_ARGS__);                          \
    if (g.Cin == 32 && g.Cout == 64) return FN<32, 64>(__VA_ARGS__);                          \
    if (g.Cin == 64 && g.Cout == 64) return FN<64, 64>(__VA_ARGS__);                          \
    if (g.Cin == 64 && g.Cout == 128) return FN<64, 128>(__VA_ARGS__);                        \
    if (g.Cin == 128 && g.Cout == 128) return FN<128, 128>(__VA_ARGS__);                      \
    flb_set_error("tensor-core conv: unsupported channels %d -> %d", g.Cin, g.Cout);          \
    return FLB_ERR_UNSUPPORTED;

// the resident-weight (halo) forward kernels can add the BatchNorm statistics of their output in the epilogue
bool conv_fwd_fuses_stats(int cin, int cout) { return conv_supported(cin, cout) && cin * cout * 36 <= 150 * 1024; }
int conv_fwd(const flb_train_args& a, const ConvGeom& g, const float* xin, float* z, const float* wt, long long ldt, int boff, cudaStream_t st,
             double* bn_acc, int bn_coff, int bn_stride) {
    FLB_CONV_DISPATCH(conv_fwd_t, a, g, xin, z, wt, ldt, boff, bn_acc, bn_coff, bn_stride, st)
}
int conv_dgrad(const flb_train_args& a, const ConvGeom& g, const float* dz, float* dx, const float* wt, long long ldt, cudaStream_t st) {
    FLB_CONV_DISPATCH(conv_dgrad_t, a, g, dz, dx, wt, ldt, st)
}
// Split-K factor of the wgrad kernels (one CTA per SM: they are shared-memory bound).  With `base` CTAs per split the
// launch takes waves(s) = ceil(base * s / SMs) rounds, each costing 1/s of the k-loop (~0.5 us per 32-pixel k-block) plus
// one accumulator flush (`flush_elems` fp32 reductions per CTA, ~0.2 ns each as 128-byte warp-wide REDs) -- calibrated on
// B200 against 1/4/10/14/29 splits.  (SimpleCNN conv2, 10 clients: 14 splits; CIFAR conv2, 100 clients: ~10; conv6: 1-2.)
static int wgrad_splits(int base, int total_kb, int flush_elems) {
    const int sms = flb_num_sms();
    const double work = 0.5 * total_kb, flush = 2e-4 * flush_elems;
    int best = 1;
    double best_cost = 1e30;
    for (int s = 1; s <= 32 && s <= total_kb; ++s) {
        const int waves = (base * s + sms - 1) / sms;
        const double cost = waves * (work / s + flush);
        if (cost < best_cost * (1.0 - 1e-9)) { best_cost = cost; best = s; }
    }
    return best;
}

template <int CIN, int COUT, int RPC>
static int conv_wgrad_halo_t(const flb_train_args& a, const ConvGeom& g, const float* xin, const float* dz, float* gt, long long ldt, cudaStream_t st) {
    using T = ConvWgradHaloT<CIN, COUT, RPC>;
    typename T::Params p;
    if (int rc = make_map_2d(&p.map_x, xin, (uint64_t)a.K * a.B * g.PP(), CIN, 32 + 2 * (g.Wp + 1), true)) return rc;
    if (int rc = make_map_2d(&p.map_dz, dz, (uint64_t)a.K * a.B * g.PP(), COUT, 32, true)) return rc;
    p.a = a; p.g = g; p.gt_all = gt; p.ldt = ldt; p.norm2_all = nullptr;
    constexpr int groups = 3 / RPC;
    const int total = (a.B * g.PP() + 31) / 32;
    int splits = wgrad_splits(a.K * groups, total, T::TILES * 96 * COUT);
    p.kb_per_split = (total + splits - 1) / splits;
    splits = (total + p.kb_per_split - 1) / p.kb_per_split;
    return launch<T>(p, dim3(splits, a.K, groups), st);
}

int conv_wgrad(const flb_train_args& a, const ConvGeom& g, const float* xin, const float* dz, float* gt, long long ldt, cudaStream_t st) {
    if (g.Wp + 1 <= 35 && !getenv("FLB_WGRAD_STREAMED")) {
        if (g.Cin == 32 && g.Cout == 32) return conv_wgrad_halo_t<32, 32, 3>(a, g, xin, dz, gt, ldt, st);
        if (g.Cin == 32 && g.Cout == 64) return conv_wgrad_halo_t<32, 64, 3>(a, g, xin, dz, gt, ldt, st);
        if (g.Cin == 64 && g.Cout == 64) return conv_wgrad_halo_t<64, 64, 3>(a, g, xin, dz, gt, ldt, st);
        if (g.Cin == 64 && g.Cout == 128) return conv_wgrad_halo_t<64, 128, 1>(a, g, xin, dz, gt, ldt, st);
        if (g.Cin == 128 && g.Cout == 128) return conv_wgrad_halo_t<128, 128, 1>(a, g, xin, dz, gt, ldt, st);
    }
    if (g.Cin == 32 && g.Cout == 32) return conv_wgrad_t<32, 32, 3>(a, g, xin, dz, gt, ldt, st);
    if (g.Cin == 32 && g.Cout == 64) return conv_wgrad_t<32, 64, 3>(a, g, xin, dz, gt, ldt, st);
    if (g.Cin == 64 && g.Cout == 64) return conv_wgrad_t<64, 64, 5>(a, g, xin, dz, gt, ldt, st);
    if (g.Cin == 64 && g.Cout == 128) return conv_wgrad_t<64, 128, 3>(a, g, xin, dz, gt, ldt, st);
    if (g.Cin == 128 && g.Cout == 128) return conv_wgrad_t<128, 128, 3>(a, g, xin, dz, gt, ldt, st);
    flb_set_error("tensor-core conv wgrad: unsupported channels %d -> %d", g.Cin, g.Cout);
    return FLB_ERR_UNSUPPORTED;
}

static int make_w_map(CUtensorMap* m, const flb_train_args& a, int woff, int in, int out, uint32_t box_rows, bool mn_major = false) {
    const uint64_t dims[3] = {(uint64_t)in, (uint64_t)out, (uint64_t)a.K};
    const uint64_t strides[2] = {(uint64_t)in * sizeof(float), (uint64_t)a.ld * sizeof(float)};
    const uint32_t box[3] = {32, box_rows, 1};
    return make_map(m, a.W + woff, 3, dims, strides, box, mn_major);
}

template <int IN, int OUT>
static int fc_fwd_t(const flb_train_args& a, const float* act, float* out, int woff, int splits, cudaStream_t st) {
    using T = FcFwdT<IN, OUT>;
    typename T::Params p;
    if (int rc = make_w_map(&p.map_w, a, woff, IN, OUT, 128)) return rc;
    if (int rc = make_map_2d(&p.map_act, act, (uint64_t)a.K * a.B, IN, 32)) return rc;
    p.a = a; p.out_all = out;
    constexpr int total = (IN + 31) / 32;
    p.kb_per_split = (total + splits - 1) / splits;
    return launch<T>(p, dim3((total + p.kb_per_split - 1) / p.kb_per_split, a.K, OUT / 128), st);
}
template <int IN, int OUT>
static int fc_dgrad_t(const flb_train_args& a, const float* dout, float* dact, int woff, cudaStream_t st) {
    using T = FcDgradT<IN, OUT>;
    typename T::Params p;
    if (int rc = make_w_map(&p.map_w, a, woff, IN, OUT, 32, true)) return rc;
    if (int rc = make_map_2d(&p.map_dout, dout, (uint64_t)a.K * a.B, OUT, 32)) return rc;
    p.a = a; p.dact_all = dact;
    return launch<T>(p, dim3((IN + 127) / 128, a.K), st);
}
template <int IN, int OUT>
static int fc_wgrad_t(const flb_train_args& a, const float* dout, const float* act, int woff, cudaStream_t st) {
    using T = FcWgradT<IN, OUT>;
    typename T::Params p;
    if (int rc = make_map_2d(&p.map_dout, dout, (uint64_t)a.K * a.B, OUT, 32, true)) return rc;
    if (int rc = make_map_2d(&p.map_act, act, (uint64_t)a.K * a.B, IN, 32, true)) return rc;
    p.a = a; p.woff = woff;
    return launch<T>(p, dim3((IN + 255) / 256, a.K, OUT / 128), st);
}
// transposed accumulator, coalesced epilogue (out a multiple of 128); ADAM applies the optimizer step instead of storing G
template <int IN, bool ADAM>
static int fc_wgrad_swap_t(const flb_train_args& a, const float* dout, const float* act, int woff, int out, cudaStream_t st) {
    using T = FcWgradSwapT<IN, ADAM>;
    typename T::Params p;
    if (int rc = make_map_2d(&p.map_dout, dout, (uint64_t)a.K * a.B, out, 32, true)) return rc;
    if (int rc = make_map_2d(&p.map_act, act, (uint64_t)a.K * a.B, IN, 32, true)) return rc;
    p.a = a; p.woff = woff;
    return launch<T>(p, dim3((IN + 127) / 128, a.K, out / 128), st);
}

// tensor maps of SimpleCNN's fused classifier kernel (fc1_fused.cu): fc1.weight as K-major [128 x 32] boxes and as MN-major
// [32 x 32] boxes, the activations a2 as [32 x 32] boxes
int make_fc1_maps(const flb_train_args& a, const float* act, CUtensorMap* w, CUtensorMap* w_mn, CUtensorMap* m_act) {
    if (int rc = make_w_map(w, a, SimpleCnnOff::f1w, 3136, 128, 128)) return rc;
    if (int rc = make_w_map(w_mn, a, SimpleCnnOff::f1w, 3136, 128, 32, true)) return rc;
    return make_map_2d(m_act, act, (uint64_t)a.K * a.B, 3136, 32);
}

#define FLB_FC_DISPATCH(FN, ...)                                                              \
    if (in == 3136 && out == 128) return FN<3136, 128>(__VA_ARGS__);                          \
    if (in == 2048 && out == 512) return FN<2048, 512>(__VA_ARGS__);                          \
    if (in == 512 && out == 256) return FN<512, 256>(__VA_ARGS__);                            \
    flb_set_error("tensor-core linear: unsupported shape %d -> %d", in, out);                 \
    return FLB_ERR_UNSUPPORTED;

int fc_fwd(const flb_train_args& a, const float* act, float* outp, int in, int out, int woff, int splits, cudaStream_t st) {
    FLB_FC_DISPATCH(fc_fwd_t, a, act, outp, woff, splits, st)
}
int fc_dgrad(const flb_train_args& a, const float* dout, float* dact, int in, int out, int woff, cudaStream_t st) {
    FLB_FC_DISPATCH(fc_dgrad_t, a, dout, dact, woff, st)
}
int fc_wgrad(const flb_train_args& a, const float* dout, const float* act, int in, int out, int woff, cudaStream_t st, bool adam) {
    static const bool no_swap = getenv("FLB_FC_WGRAD_NO_SWAP") != nullptr;       // A/B switch for the plain (gradient-storing) variant
    if (in == 3136 && out == 128) {
        if (adam) return fc_wgrad_swap_t<3136, true>(a, dout, act, woff, out, st);
        return no_swap ? fc_wgrad_t<3136, 128>(a, dout, act, woff, st) : fc_wgrad_swap_t<3136, false>(a, dout, act, woff, out, st);
    }
    if (in == 2048 && out == 512) {
        if (adam) return fc_wgrad_swap_t<2048, true>(a, dout, act, woff, out, st);
        return no_swap ? fc_wgrad_t<2048, 512>(a, dout, act, woff, st) : fc_wgrad_swap_t<2048, false>(a, dout, act, woff, out, st);
    }
    if (adam) {
        flb_set_error("tensor-core linear wgrad with the fused optimizer: unsupported shape %d -> %d", in, out);
        return FLB_ERR_UNSUPPORTED;
    }
    if (in == 512 && out == 256) return no_swap ? fc_wgrad_t<512, 256>(a, dout, act, woff, st) : fc_wgrad_swap_t<512, false>(a, dout, act, woff, out, st);
    flb_set_error("tensor-core linear: unsupported shape %d -> %d", in, out);
    return FLB_ERR_UNSUPPORTED;
}

}  // namespace tc

// ---- per-role timeline of the resident-weight convolution kernels (tc_gemm.cuh TraceBuf; profiling aid) ----------------
// buf: device memory of 8 + 24 * cap bytes ([int n][int cap][cap x (event, tile, clock) int64]), zeroed by the caller; NULL
// disables.  Each of the CTA's warps records into its own eighth of the buffer; a region's last record holds its count.
extern "C" int flb_debug_trace_set(void* buf, int cap) {
    if (buf) {
        const int hdr[2] = {0, cap};
        FLB_CUDA(cudaMemcpy(buf, hdr, sizeof(hdr), cudaMemcpyHostToDevice));
    }
    FLB_CUDA(cudaMemcpyToSymbol(tc::g_trace, &buf, sizeof(void*)));
    return FLB_OK;
}
