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
// Weights stay in the reference's [Cout][Cin][3][3] layout in HBM; the conv weight operand (73.7 KB for conv2) is
// permuted into the canonical swizzled layout by the CTA's threads once and stays resident in shared memory.
#include "tc_gemm.cuh"

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

// ---- conv forward -------------------------------------------------------------------------------------------
template <int CIN, int COUT>
struct ConvFwdTC {
    struct Params { CUtensorMap map_x; flb_train_args a; ConvGeom g; float* z_all; int woff, boff; };
    static constexpr int CH = CIN / 32, NKB = 9 * CH;
    static constexpr int STAGES = 4, STAGE_BYTES = 128 * 128, RESIDENT_BYTES = NKB * COUT * 128, TMEM_COLS = COUT <= 32 ? 32 : (COUT <= 64 ? 64 : (COUT <= 128 ? 128 : 256));
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
    __device__ void prefetch(const Params& p) { tma_prefetch_desc(&p.map_x); }
    __device__ void stage_resident(const Params& p, uint8_t* res, int tid) {
        const float* w = p.a.W + (long long)client * p.a.ld + p.woff;
        for (int s = tid; s < COUT * CIN * 9; s += THREADS) {           // coalesced read of [Cout][Cin][9]
            const int tap = s % 9, ci = (s / 9) % CIN, n = s / (9 * CIN);
            const int kb = tap * CH + ci / 32;
            *reinterpret_cast<float*>(res + (size_t)kb * COUT * 128 + sw128_offset(n, ci & 31)) = w[s];
        }
    }
    __device__ void load(const Params& p, int kb, uint8_t* stage, uint64_t* bar) {
        const int tap = kb / CH, c = kb % CH;
        mbar_expect_tx(bar, STAGE_BYTES);
        tma_load_2d(&p.map_x, stage, bar, c * 32, row0 + m0 + tap_shift(tap, p.g.Wp));
    }
    __device__ void mma(int kb, uint32_t stage, uint32_t res, uint32_t tmem) {
        constexpr uint32_t id = idesc_tf32(128, COUT, false, false);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            mma_tf32(tmem, smem_desc(stage + k * 32, 16, 1024), smem_desc(res + kb * COUT * 128 + k * 32, 16, 1024), id, kb > 0 || k > 0);
    }
    __device__ void epilogue(const Params& p, uint32_t tmem, int quarter, int lane) {
        const int m = m0 + quarter * 32 + lane;
        const float* bias = p.a.W + (long long)client * p.a.ld + p.boff;
        float* z = p.z_all + ((long long)row0 + m) * COUT;
        const bool ok = m < p.a.B * p.g.PP();
#pragma unroll 1
        for (int c0 = 0; c0 < COUT; c0 += 32) {
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + c0, v);
            if (ok) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4*>(z + c0 + i) = make_float4(v[i] + bias[c0 + i], v[i + 1] + bias[c0 + i + 1],
                                                                         v[i + 2] + bias[c0 + i + 2], v[i + 3] + bias[c0 + i + 3]);
            }
        }
    }
};

// ---- conv dgrad ---------------------------------------------------------------------------------------------
template <int CIN, int COUT>
struct ConvDgradTC {
    struct Params { CUtensorMap map_dz; flb_train_args a; ConvGeom g; float* dx_all; int woff; };
    static constexpr int CH = COUT / 32, NKB = 9 * CH;
    static constexpr int STAGES = 4, STAGE_BYTES = 128 * 128, RESIDENT_BYTES = NKB * CIN * 128, TMEM_COLS = CIN <= 32 ? 32 : (CIN <= 64 ? 64 : 128);
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
    __device__ void prefetch(const Params& p) { tma_prefetch_desc(&p.map_dz); }
    __device__ void stage_resident(const Params& p, uint8_t* res, int tid) {
        const float* w = p.a.W + (long long)client * p.a.ld + p.woff;
        for (int s = tid; s < COUT * CIN * 9; s += THREADS) {
            const int tap = s % 9, ci = (s / 9) % CIN, co = s / (9 * CIN);
            const int kb = tap * CH + co / 32;                              // B tile rows = cin (N), k = cout
            *reinterpret_cast<float*>(res + (size_t)kb * CIN * 128 + sw128_offset(ci, co & 31)) = w[s];
        }
    }
    __device__ void load(const Params& p, int kb, uint8_t* stage, uint64_t* bar) {
        const int tap = kb / CH, c = kb % CH;
        mbar_expect_tx(bar, STAGE_BYTES);
        tma_load_2d(&p.map_dz, stage, bar, c * 32, row0 + m0 - tap_shift(tap, p.g.Wp));
    }
    __device__ void mma(int kb, uint32_t stage, uint32_t res, uint32_t tmem) {
        constexpr uint32_t id = idesc_tf32(128, CIN, false, false);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            mma_tf32(tmem, smem_desc(stage + k * 32, 16, 1024), smem_desc(res + kb * CIN * 128 + k * 32, 16, 1024), id, kb > 0 || k > 0);
    }
    __device__ void epilogue(const Params& p, uint32_t tmem, int quarter, int lane) {
        const int m = m0 + quarter * 32 + lane;
        float* dx = p.dx_all + ((long long)row0 + m) * CIN;
        const bool ok = m < p.a.B * p.g.PP();
#pragma unroll 1
        for (int c0 = 0; c0 < CIN; c0 += 32) {
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + c0, v);
            if (ok) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dx + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
        }
    }
};

// ---- conv wgrad ---------------------------------------------------------------------------------------------
template <int CIN, int COUT>
struct ConvWgradTC {
    struct Params { CUtensorMap map_x; CUtensorMap map_dz; flb_train_args a; ConvGeom g; int woff; int kb_per_split; };
    static constexpr int ACH = 9 * CIN / 32, MT = (ACH * 32 + 127) / 128, ASLOTS = MT * 4, BCH = COUT / 32;
    static constexpr int A_BYTES = ASLOTS * 4096, STAGE_BYTES = A_BYTES + BCH * 4096;
    static constexpr int STAGES = 3, RESIDENT_BYTES = 0, TMEM_COLS = MT * COUT <= 64 ? 64 : (MT * COUT <= 128 ? 128 : (MT * COUT <= 256 ? 256 : 512));
    int client, row0, kb0;
    __device__ bool setup(const Params& p, int& num_kb) {
        client = blockIdx.y;
        const int bsz = flb_bsz(p.a, client);
        const int total = bsz * p.g.PP() / 32;
        kb0 = blockIdx.x * p.kb_per_split;
        if (kb0 >= total) return false;
        num_kb = min(p.kb_per_split, total - kb0);
        row0 = client * p.a.B * p.g.PP();
        return true;
    }
    __device__ void prefetch(const Params& p) { tma_prefetch_desc(&p.map_x); tma_prefetch_desc(&p.map_dz); }
    __device__ void stage_resident(const Params&, uint8_t*, int) {}
    __device__ void load(const Params& p, int kb, uint8_t* stage, uint64_t* bar) {
        const int px = row0 + (kb0 + kb) * 32;
        mbar_expect_tx(bar, (ACH + BCH) * 4096);
#pragma unroll 1
        for (int ch = 0; ch < ACH; ++ch) {                       // chunk = (tap, 32-channel slice of Cin)
            const int tap = ch / (CIN / 32), c = ch % (CIN / 32);
            tma_load_2d(&p.map_x, stage + ch * 4096, bar, c * 32, px + tap_shift(tap, p.g.Wp));
        }
#pragma unroll
        for (int c = 0; c < BCH; ++c) tma_load_2d(&p.map_dz, stage + A_BYTES + c * 4096, bar, c * 32, px);
    }
    __device__ void mma(int kb, uint32_t stage, uint32_t, uint32_t tmem) {
        constexpr uint32_t id = idesc_tf32(128, COUT, true, true);
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
                mma_tf32(tmem + mt * COUT, smem_desc_mn(stage + mt * 4 * 4096 + k * 1024, 4096, 512),
                         smem_desc_mn(stage + A_BYTES + k * 1024, 4096, 512), id, kb > 0 || k > 0);
    }
    __device__ void epilogue(const Params& p, uint32_t tmem, int quarter, int lane) {
        float* gw = p.a.G + (long long)client * p.a.ld + p.woff;
#pragma unroll 1
        for (int mt = 0; mt < MT; ++mt) {
            const int row = mt * 128 + quarter * 32 + lane;          // row = chunk * 32 + (cin within the slice)
            const int ch = row >> 5, tap = ch / (CIN / 32), ci = (ch % (CIN / 32)) * 32 + (row & 31);
            const bool ok = ch < ACH;
#pragma unroll 1
            for (int c0 = 0; c0 < COUT; c0 += 32) {
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + mt * COUT + c0, v);
                if (ok) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) atomicAdd(&gw[((c0 + i) * CIN + ci) * 9 + tap], v[i]);
                }
            }
        }
    }
};

// ---- linear forward (split-K) --------------------------------------------------------------------------------
template <int IN, int OUT>
struct FcFwdTC {
    struct Params { CUtensorMap map_w; CUtensorMap map_act; flb_train_args a; float* out_all; int kb_per_split; };
    static_assert(OUT == 128, "one 128-row accumulator tile");
    static constexpr int STAGES = 6, STAGE_BYTES = 128 * 128 + 32 * 128, RESIDENT_BYTES = 0, TMEM_COLS = 32;
    int client, kb0, bsz;
    __device__ bool setup(const Params& p, int& num_kb) {
        client = blockIdx.y;
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
        mbar_expect_tx(bar, STAGE_BYTES);
        tma_load_3d(&p.map_w, stage, bar, k0, 0, client);
        tma_load_2d(&p.map_act, stage + 128 * 128, bar, k0, client * p.a.B);
    }
    __device__ void mma(int kb, uint32_t stage, uint32_t, uint32_t tmem) {
        constexpr uint32_t id = idesc_tf32(128, 32, false, false);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            mma_tf32(tmem, smem_desc(stage + k * 32, 16, 1024), smem_desc(stage + 128 * 128 + k * 32, 16, 1024), id, kb > 0 || k > 0);
    }
    __device__ void epilogue(const Params& p, uint32_t tmem, int quarter, int lane) {
        const int j = quarter * 32 + lane;
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16), v);
        float* out = p.out_all + (long long)client * p.a.B * OUT + j;
#pragma unroll
        for (int b = 0; b < 32; ++b)
            if (b < bsz) atomicAdd(out + b * OUT, v[b]);
    }
};

// ---- linear dgrad ---------------------------------------------------------------------------------------------
template <int IN, int OUT>
struct FcDgradTC {
    struct Params { CUtensorMap map_w; CUtensorMap map_dout; flb_train_args a; float* dact_all; };
    static constexpr int STAGES = 4, STAGE_BYTES = 4 * 4096 + 4096, RESIDENT_BYTES = 0, TMEM_COLS = 32;
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
        mbar_expect_tx(bar, STAGE_BYTES);
#pragma unroll
        for (int c = 0; c < 4; ++c) tma_load_3d(&p.map_w, stage + c * 4096, bar, m0 + 32 * c, kb * 32, client);   // rows = out features (K)
        tma_load_2d(&p.map_dout, stage + 4 * 4096, bar, kb * 32, client * p.a.B);
    }
    __device__ void mma(int kb, uint32_t stage, uint32_t, uint32_t tmem) {
        constexpr uint32_t id = idesc_tf32(128, 32, true, false);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            mma_tf32(tmem, smem_desc_mn(stage + k * 1024, 4096, 512), smem_desc(stage + 4 * 4096 + k * 32, 16, 1024), id, kb > 0 || k > 0);
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

// ---- linear wgrad ---------------------------------------------------------------------------------------------
template <int IN, int OUT>
struct FcWgradTC {
    struct Params { CUtensorMap map_dout; CUtensorMap map_act; flb_train_args a; int woff; };
    static_assert(OUT == 128, "one 128-row accumulator tile");
    static constexpr int STAGES = 1, STAGE_BYTES = 4 * 4096 + 8 * 4096, RESIDENT_BYTES = 0, TMEM_COLS = 256;
    int client, n0, ksteps;
    __device__ bool setup(const Params& p, int& num_kb) {
        client = blockIdx.y;
        if (flb_bsz(p.a, client) == 0) return false;
        n0 = blockIdx.x * 256;
        ksteps = p.a.B / 8;              // rows bsz..B-1 of dout are zero (head kernel); rows >= B belong to the next client
        num_kb = 1;
        return true;
    }
    __device__ void prefetch(const Params& p) { tma_prefetch_desc(&p.map_dout); tma_prefetch_desc(&p.map_act); }
    __device__ void stage_resident(const Params&, uint8_t*, int) {}
    __device__ void load(const Params& p, int, uint8_t* stage, uint64_t* bar) {
        mbar_expect_tx(bar, STAGE_BYTES);
#pragma unroll
        for (int c = 0; c < 4; ++c) tma_load_2d(&p.map_dout, stage + c * 4096, bar, 32 * c, client * p.a.B);
#pragma unroll
        for (int c = 0; c < 8; ++c) tma_load_2d(&p.map_act, stage + 4 * 4096 + c * 4096, bar, n0 + 32 * c, client * p.a.B);
    }
    __device__ void mma(int, uint32_t stage, uint32_t, uint32_t tmem) {
        constexpr uint32_t id = idesc_tf32(128, 256, true, true);
        for (int k = 0; k < ksteps; ++k)
            mma_tf32(tmem, smem_desc_mn(stage + k * 1024, 4096, 512), smem_desc_mn(stage + 4 * 4096 + k * 1024, 4096, 512), id, k > 0);
    }
    __device__ void epilogue(const Params& p, uint32_t tmem, int quarter, int lane) {
        const int j = quarter * 32 + lane;
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

template <class T>
static int launch(const typename T::Params& p, dim3 grid, cudaStream_t st) {
    constexpr size_t smem = (size_t)T::STAGES * T::STAGE_BYTES + T::RESIDENT_BYTES + 1024;
    static bool configured = false;
    if (!configured) {
        FLB_CUDA(cudaFuncSetAttribute(gemm_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    gemm_kernel<T><<<grid, THREADS, smem, st>>>(p);
    return FLB_OK;
}

// ---- host entry points used by the step orchestrator (train_simplecnn.cu) ------------------------------------------
int conv_fwd_32_64(const flb_train_args& a, const ConvGeom& g, const float* xin, float* z, int woff, int boff, cudaStream_t st) {
    using T = ConvFwdTC<32, 64>;
    T::Params p;
    if (int rc = make_map_2d(&p.map_x, xin, (uint64_t)a.K * a.B * g.PP(), 32, 128)) return rc;
    p.a = a; p.g = g; p.z_all = z; p.woff = woff; p.boff = boff;
    return launch<T>(p, dim3((a.B * g.PP() + 127) / 128, a.K), st);
}

int conv_dgrad_32_64(const flb_train_args& a, const ConvGeom& g, const float* dz, float* dx, int woff, cudaStream_t st) {
    using T = ConvDgradTC<32, 64>;
    T::Params p;
    if (int rc = make_map_2d(&p.map_dz, dz, (uint64_t)a.K * a.B * g.PP(), 64, 128)) return rc;
    p.a = a; p.g = g; p.dx_all = dx; p.woff = woff;
    return launch<T>(p, dim3((a.B * g.PP() + 127) / 128, a.K), st);
}

int conv_wgrad_32_64(const flb_train_args& a, const ConvGeom& g, const float* xin, const float* dz, int woff, int splits, cudaStream_t st) {
    using T = ConvWgradTC<32, 64>;
    T::Params p;
    if (int rc = make_map_2d(&p.map_x, xin, (uint64_t)a.K * a.B * g.PP(), 32, 32, true)) return rc;
    if (int rc = make_map_2d(&p.map_dz, dz, (uint64_t)a.K * a.B * g.PP(), 64, 32, true)) return rc;
    p.a = a; p.g = g; p.woff = woff;
    const int total = a.B * g.PP() / 32;
    p.kb_per_split = (total + splits - 1) / splits;
    return launch<T>(p, dim3(splits, a.K), st);
}

static int make_w_map(CUtensorMap* m, const flb_train_args& a, int woff, int in, int out, uint32_t box_rows, bool mn_major = false) {
    const uint64_t dims[3] = {(uint64_t)in, (uint64_t)out, (uint64_t)a.K};
    const uint64_t strides[2] = {(uint64_t)in * sizeof(float), (uint64_t)a.ld * sizeof(float)};
    const uint32_t box[3] = {32, box_rows, 1};
    return make_map(m, a.W + woff, 3, dims, strides, box, mn_major);
}

int fc_fwd_3136_128(const flb_train_args& a, const float* act, float* out, int woff, int splits, cudaStream_t st) {
    using T = FcFwdTC<3136, 128>;
    T::Params p;
    if (int rc = make_w_map(&p.map_w, a, woff, 3136, 128, 128)) return rc;
    if (int rc = make_map_2d(&p.map_act, act, (uint64_t)a.K * a.B, 3136, 32)) return rc;
    p.a = a; p.out_all = out;
    p.kb_per_split = (98 + splits - 1) / splits;
    return launch<T>(p, dim3(splits, a.K), st);
}

int fc_dgrad_3136_128(const flb_train_args& a, const float* dout, float* dact, int woff, cudaStream_t st) {
    using T = FcDgradTC<3136, 128>;
    T::Params p;
    if (int rc = make_w_map(&p.map_w, a, woff, 3136, 128, 32, true)) return rc;
    if (int rc = make_map_2d(&p.map_dout, dout, (uint64_t)a.K * a.B, 128, 32)) return rc;
    p.a = a; p.dact_all = dact;
    return launch<T>(p, dim3((3136 + 127) / 128, a.K), st);
}

int fc_wgrad_3136_128(const flb_train_args& a, const float* dout, const float* act, int woff, cudaStream_t st) {
    using T = FcWgradTC<3136, 128>;
    T::Params p;
    if (int rc = make_map_2d(&p.map_dout, dout, (uint64_t)a.K * a.B, 128, 32, true)) return rc;
    if (int rc = make_map_2d(&p.map_act, act, (uint64_t)a.K * a.B, 3136, 32, true)) return rc;
    p.a = a; p.woff = woff;
    return launch<T>(p, dim3((3136 + 255) / 256, a.K), st);
}

}  // namespace tc
