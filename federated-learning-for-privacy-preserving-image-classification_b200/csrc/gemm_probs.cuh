// Problem functors for the fp32 CUDA-core grouped GEMM (gemm_simt.cuh): implicit-GEMM convolutions on the
// zero-padded NHWC grid and linear layers, forward / dgrad / wgrad, one group per client (or per sample for the
// per-sample-norm variant).  Shared by the SimpleCNN and CIFAR10CNN step orchestrators.
#pragma once
#include "train_common.cuh"
#include "gemm_simt.cuh"

namespace {

struct ConvFwdProb {
    static constexpr bool A_MCONTIG = false, B_NCONTIG = false;
    flb_train_args a; ConvGeom g;
    const float* xin_all; float* z_all; int woff, boff;
    const float* xin; float* z; const float* w; const float* bias; int Mtot;
    __device__ bool setup(int client, int& M, int& N, int& Kd) {
        const int bsz = flb_bsz(a, client);
        if (bsz == 0) return false;
        const long long kb = (long long)client * a.B;
        xin = xin_all + kb * g.PP() * g.Cin;
        z = z_all + kb * g.PP() * g.Cout;
        w = a.W + (long long)client * a.ld + woff;
        bias = a.W + (long long)client * a.ld + boff;
        Mtot = a.B * g.PP();
        M = bsz * g.PP(); N = g.Cout; Kd = 9 * g.Cin;
        return true;
    }
    __device__ float loadA(int m, int k) const {
        const int tap = k / g.Cin, ci = k - tap * g.Cin;
        const int row = m + (tap / 3 - 1) * g.Wp + (tap % 3 - 1);
        return (row >= 0 && row < Mtot) ? xin[(long long)row * g.Cin + ci] : 0.f;
    }
    __device__ float loadB(int n, int k) const {
        const int tap = k / g.Cin, ci = k - tap * g.Cin;
        return __ldg(&w[(n * g.Cin + ci) * 9 + tap]);
    }
    __device__ void store(int m, int n, float acc) { z[(long long)m * g.Cout + n] = acc + bias[n]; }
    __device__ void finish() {}
};

struct ConvDgradProb {      // dx[m][ci] = sum_{tap,co} dz[m - shift(tap)][co] * W[co][ci][tap]
    static constexpr bool A_MCONTIG = false, B_NCONTIG = false;
    flb_train_args a; ConvGeom g;
    const float* dz_all; float* dx_all; int woff;
    const float* dz; float* dx; const float* w; int Mtot;
    __device__ bool setup(int client, int& M, int& N, int& Kd) {
        const int bsz = flb_bsz(a, client);
        if (bsz == 0) return false;
        const long long kb = (long long)client * a.B;
        dz = dz_all + kb * g.PP() * g.Cout;
        dx = dx_all + kb * g.PP() * g.Cin;
        w = a.W + (long long)client * a.ld + woff;
        Mtot = a.B * g.PP();
        M = bsz * g.PP(); N = g.Cin; Kd = 9 * g.Cout;
        return true;
    }
    __device__ float loadA(int m, int k) const {
        const int tap = k / g.Cout, co = k - tap * g.Cout;
        const int row = m - ((tap / 3 - 1) * g.Wp + (tap % 3 - 1));
        return (row >= 0 && row < Mtot) ? dz[(long long)row * g.Cout + co] : 0.f;
    }
    __device__ float loadB(int n, int k) const {
        const int tap = k / g.Cout, co = k - tap * g.Cout;
        return __ldg(&w[(co * g.Cin + n) * 9 + tap]);
    }
    __device__ void store(int m, int n, float acc) { dx[(long long)m * g.Cin + n] = acc; }
    __device__ void finish() {}
};

// dW[co][ci][tap] = sum_px dz[px][co] * x[px + shift(tap)][ci];  column n == 9*Cin is the bias gradient.
// In dp_mode 1 each pixel row is scaled by its sample's clip coefficient.
struct ConvWgradProb {
    static constexpr bool A_MCONTIG = true, B_NCONTIG = true;
    flb_train_args a; ConvGeom g;
    const float* dz_all; const float* xin_all; const float* coef_all; int woff, boff;
    int no_bias;                 // 1: the bias column is left out (CIFAR10CNN per-sample mode: the BatchNorm pass owns it)
    const float* dz; const float* xin; const float* coef; float* gw; float* gb; int Mtot;
    __device__ bool setup(int client, int& M, int& N, int& Kd) {
        const int bsz = flb_bsz(a, client);
        if (bsz == 0) return false;
        const long long kb = (long long)client * a.B;
        dz = dz_all + kb * g.PP() * g.Cout;
        xin = xin_all + kb * g.PP() * g.Cin;
        coef = coef_all ? coef_all + kb : nullptr;
        gw = a.G + (long long)client * a.ld + woff;
        gb = a.G + (long long)client * a.ld + boff;
        Mtot = a.B * g.PP();
        M = g.Cout; N = 9 * g.Cin + (no_bias ? 0 : 1); Kd = bsz * g.PP();
        return true;
    }
    __device__ float loadA(int m, int k) const {
        const float v = dz[(long long)k * g.Cout + m];
        return coef ? v * coef[k / g.PP()] : v;
    }
    __device__ float loadB(int n, int k) const {
        if (n == 9 * g.Cin) return 1.f;
        const int tap = n / g.Cin, ci = n - tap * g.Cin;
        const int row = k + (tap / 3 - 1) * g.Wp + (tap % 3 - 1);
        return (row >= 0 && row < Mtot) ? xin[(long long)row * g.Cin + ci] : 0.f;
    }
    __device__ void store(int m, int n, float acc) {
        if (n == 9 * g.Cin) { atomicAdd(&gb[m], acc); return; }
        const int tap = n / g.Cin, ci = n - tap * g.Cin;
        atomicAdd(&gw[(m * g.Cin + ci) * 9 + tap], acc);
    }
    __device__ void finish() {}
};

// per-sample conv weight-gradient norm (dp_mode 1): group = (client, sample); the [Cout, 9*Cin+1] per-sample
// gradient tile lives in registers only -- squared, reduced with warp shuffles, one atomic per CTA.
struct ConvWgradNormProb {
    static constexpr bool A_MCONTIG = true, B_NCONTIG = true;
    flb_train_args a; ConvGeom g;
    const float* dz_all; const float* xin_all; float* norm2_all;
    int no_bias;                 // 1: weight gradient only
    const float* dz; const float* xin; float* dst; float sq; int lo, hi;
    __device__ bool setup(int group, int& M, int& N, int& Kd) {
        const int client = group / a.B, b = group % a.B;
        sq = 0.f;
        if (b >= flb_bsz(a, client)) return false;
        const long long kb = (long long)client * a.B;
        dz = dz_all + (kb + b) * g.PP() * g.Cout;
        xin = xin_all + kb * g.PP() * g.Cin;
        lo = -b * g.PP(); hi = (a.B - b) * g.PP();       // row bounds relative to this sample's first pixel
        xin += (long long)b * g.PP() * g.Cin;
        dst = norm2_all + kb + b;
        M = g.Cout; N = 9 * g.Cin + (no_bias ? 0 : 1); Kd = g.PP();
        return true;
    }
    __device__ float loadA(int m, int k) const { return dz[(long long)k * g.Cout + m]; }
    __device__ float loadB(int n, int k) const {
        if (n == 9 * g.Cin) return 1.f;
        const int tap = n / g.Cin, ci = n - tap * g.Cin;
        const int row = k + (tap / 3 - 1) * g.Wp + (tap % 3 - 1);
        return (row >= lo && row < hi) ? xin[(long long)row * g.Cin + ci] : 0.f;
    }
    __device__ void store(int, int, float acc) { sq = fmaf(acc, acc, sq); }
    __device__ void finish() {
        const float v = flb_warp_sum(sq);
        if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(dst, v);
    }
};

struct LinFwdProb {         // out[b][n] += sum_k act[b][k] * W[n][k]     (bias added by the consumer)
    static constexpr bool A_MCONTIG = false, B_NCONTIG = false;
    flb_train_args a; int In, Out, woff; const float* act_all; float* out_all;
    const float* act; float* out; const float* w;
    __device__ bool setup(int client, int& M, int& N, int& Kd) {
        const int bsz = flb_bsz(a, client);
        if (bsz == 0) return false;
        act = act_all + (long long)client * a.B * In;
        out = out_all + (long long)client * a.B * Out;
        w = a.W + (long long)client * a.ld + woff;
        M = bsz; N = Out; Kd = In;
        return true;
    }
    __device__ float loadA(int m, int k) const { return act[(long long)m * In + k]; }
    __device__ float loadB(int n, int k) const { return __ldg(&w[(long long)n * In + k]); }
    __device__ void store(int m, int n, float acc) { atomicAdd(&out[m * Out + n], acc); }
    __device__ void finish() {}
};

struct LinDgradProb {       // dact[b][n] = sum_k dout[b][k] * W[k][n]
    static constexpr bool A_MCONTIG = false, B_NCONTIG = true;
    flb_train_args a; int In, Out, woff; const float* dout_all; float* dact_all;
    const float* dout; float* dact; const float* w;
    __device__ bool setup(int client, int& M, int& N, int& Kd) {
        const int bsz = flb_bsz(a, client);
        if (bsz == 0) return false;
        dout = dout_all + (long long)client * a.B * Out;
        dact = dact_all + (long long)client * a.B * In;
        w = a.W + (long long)client * a.ld + woff;
        M = bsz; N = In; Kd = Out;
        return true;
    }
    __device__ float loadA(int m, int k) const { return dout[m * Out + k]; }
    __device__ float loadB(int n, int k) const { return __ldg(&w[(long long)k * In + n]); }
    __device__ void store(int m, int n, float acc) { dact[(long long)m * In + n] = acc; }
    __device__ void finish() {}
};

struct LinWgradProb {       // dW[m][n] = sum_b dout[b][m] * act[b][n]   (bias gradient: head_wgrad_kernel)
    static constexpr bool A_MCONTIG = true, B_NCONTIG = true;
    flb_train_args a; int In, Out, woff, boff; const float* dout_all; const float* act_all; const float* coef_all;
    const float* dout; const float* act; const float* coef; float* gw; float* gb;
    __device__ bool setup(int client, int& M, int& N, int& Kd) {
        const int bsz = flb_bsz(a, client);
        if (bsz == 0) return false;
        dout = dout_all + (long long)client * a.B * Out;
        act = act_all + (long long)client * a.B * In;
        coef = coef_all ? coef_all + (long long)client * a.B : nullptr;
        gw = a.G + (long long)client * a.ld + woff;
        gb = a.G + (long long)client * a.ld + boff;
        M = Out; N = In; Kd = bsz;
        return true;
    }
    __device__ float loadA(int m, int k) const { const float v = dout[k * Out + m]; return coef ? v * coef[k] : v; }
    __device__ float loadB(int n, int k) const { return act[(long long)k * In + n]; }
    __device__ void store(int m, int n, float acc) { gw[(long long)m * In + n] = acc; }
    __device__ void finish() {}
};

}  // namespace
