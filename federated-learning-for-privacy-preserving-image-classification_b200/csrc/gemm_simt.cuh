// fp32 CUDA-core grouped GEMM over clients: C_g[m][n] = sum_k A_g(m, k) * B_g(n, k), one group per client.
// This is the precision==0 ("fp32-exact") path used for tight parity against the reference's fp32 CPU
// arithmetic; the throughput path is the tcgen05 kernel in gemm_tc.cu.  Operands are described by a
// problem functor (implicit-GEMM conv taps, transposed views, ragged per-client M from the device-side
// batch size), so the same 64x64x16 register-tiled kernel serves conv fwd/dgrad/wgrad and linear fwd/dgrad/wgrad.
#pragma once
#include "train_common.cuh"

namespace simt {

constexpr int BM = 64, BN = 64, BK = 16, THREADS = 256, PAD = 4;

// Prob interface:
//   __device__ bool setup(int group, int& M, int& N, int& Kd);       false -> nothing to do for this group
//   __device__ float loadA(int m, int k) const;   __device__ float loadB(int n, int k) const;
//   __device__ void store(int m, int n, float acc);                    called once per output element
//   __device__ void finish();                                          once per thread after all stores
//   static constexpr bool A_MCONTIG, B_NCONTIG;                        which index is contiguous in memory
template <class Prob>
__global__ void __launch_bounds__(THREADS) gemm_kernel(Prob p, int tiles_n, int splits) {
    int M, N, Kd;
    if (!p.setup(blockIdx.z, M, N, Kd)) return;
    const int m0 = (blockIdx.x / tiles_n) * BM, n0 = (blockIdx.x % tiles_n) * BN;
    if (m0 >= M || n0 >= N) return;
    const int kchunk = ((Kd + splits - 1) / splits + BK - 1) / BK * BK;
    const int kb = blockIdx.y * kchunk, ke = min(Kd, kb + kchunk);
    if (kb >= ke) return;

    __shared__ float As[2][BK][BM + PAD];
    __shared__ float Bs[2][BK][BN + PAD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float ra[4], rb[4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + THREADS * i;
            const int am = Prob::A_MCONTIG ? (e & (BM - 1)) : (e >> 4), ak = Prob::A_MCONTIG ? (e >> 6) : (e & (BK - 1));
            const int bn = Prob::B_NCONTIG ? (e & (BN - 1)) : (e >> 4), bk = Prob::B_NCONTIG ? (e >> 6) : (e & (BK - 1));
            ra[i] = (m0 + am < M && k0 + ak < ke) ? p.loadA(m0 + am, k0 + ak) : 0.f;
            rb[i] = (n0 + bn < N && k0 + bk < ke) ? p.loadB(n0 + bn, k0 + bk) : 0.f;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + THREADS * i;
            const int am = Prob::A_MCONTIG ? (e & (BM - 1)) : (e >> 4), ak = Prob::A_MCONTIG ? (e >> 6) : (e & (BK - 1));
            const int bn = Prob::B_NCONTIG ? (e & (BN - 1)) : (e >> 4), bk = Prob::B_NCONTIG ? (e >> 6) : (e & (BK - 1));
            As[buf][ak][am] = ra[i];
            Bs[buf][bk][bn] = rb[i];
        }
    };

    fetch(kb);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = kb; k0 < ke; k0 += BK) {
        const bool more = k0 + BK < ke;
        if (more) fetch(k0 + BK);               // global loads for the next tile fly while this one is multiplied
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (more) {
            stash(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < M && n < N) p.store(m, n, acc[i][j]);
        }
    p.finish();
}

template <class Prob>
static inline void launch(const Prob& p, int maxM, int maxN, int splits, int groups, cudaStream_t st) {
    const int tiles_m = (maxM + BM - 1) / BM, tiles_n = (maxN + BN - 1) / BN;
    dim3 grid(tiles_m * tiles_n, splits, groups);
    gemm_kernel<Prob><<<grid, THREADS, 0, st>>>(p, tiles_n, splits);
}

}  // namespace simt
