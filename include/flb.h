/* flb.h -- C ABI of libflb.so, the B200 (sm_100a) kernels behind the federated hot path.
 *
 * The reference (Prashant-ambati/Federated-Learning-for-Privacy-Preserving-Image-Classification)
 * is pure Python/PyTorch and has no FFI; the boundary it does have is the Python class surface
 * (src/shared/interfaces.py:75-182).  Each entry point below names the reference code it replaces;
 * INTEGRATION.md shows the ctypes stubs a reference maintainer would add.
 *
 * Conventions: every function returns 0 on success or a negative FLB_ERR_* code, in which case
 * flb_last_error() (thread-local) describes the failure.  All data pointers are DEVICE pointers
 * owned by the caller; `stream` is a cudaStream_t (NULL = legacy default stream); nothing here
 * synchronises the device or allocates memory unless stated.  There is no CPU fallback.
 */
#ifndef FLB_H_
#define FLB_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* flb_last_error(void);
int flb_version(void);
int flb_init(int device);            /* checks a sm_100 device exists; idempotent */

/* ---- FedAvg: src/aggregation/fedavg.py:267-289 (_weighted_average) -------------------------- */
/* out[p] = sum_k w[k] * theta[k*ld + p], sequential over k in fp32 (bit-exact with the reference loop).
 * accumulate != 0 starts from the current contents of out (chunked / multi-call aggregation). */
int flb_fedavg_weighted_sum(const float* theta, long long ld, const float* w, float* out,
                            int K, long long P, int accumulate, void* stream);
/* same, over tensors that were never stacked: ptrs[k*L + l] = layer l of client k (device array of
 * device pointers), seg_off[l]..seg_off[l+1] = that layer's span in the flat output. */
int flb_fedavg_weighted_sum_ptrs(const float* const* ptrs, const long long* seg_off, const float* w,
                                 float* out, int K, int L, long long P, void* stream);
/* same, over uint8-quantised client rows (dequant rule src/shared/compression.py:230-244) */
int flb_fedavg_weighted_sum_q8(const uint8_t* q, long long ldq, const float* scale, const float* zp,
                               const long long* seg_off, const float* w, float* out,
                               int K, int L, long long P, void* stream);

/* ---- update-level DP: src/shared/privacy.py:107-144,183-254 + src/client/federated_trainer.py:428-469 */
/* norm2[k] = sum_p (local[k*ld+p] - global[p])^2 in double; global may be NULL (local is the delta). */
int flb_dp_sumsq(const float* local, long long ld, const float* global_w, double* norm2,
                 int K, long long P, void* stream);
/* out[k] = global + clip(local[k]-global)*1 + sigma_k*z ; sigma_k = min(||delta_k||, max_norm)*sigma_unit,
 * sigma_unit = sqrt(2 ln(1.25/delta))/epsilon (0 disables noise).  z_in (may be NULL) injects the standard
 * normals; otherwise z = Philox4x32-10(seed, stream_base + k, element).  norms_out[k] (may be NULL) = ||delta_k||. */
int flb_dp_clip_noise(const float* local, long long ld, const float* global_w, const float* z_in,
                      const double* norm2, float* out, float* norms_out, double max_norm,
                      double sigma_unit, unsigned long long seed, unsigned long long stream_base,
                      int K, long long P, void* stream);
/* out[k] = x[k] + sigma * z  (GaussianNoiseGenerator.add_noise_to_gradients alone, privacy.py:221-254) */
int flb_dp_add_noise(const float* x, long long ld, const float* z_in, float* out, double sigma,
                     unsigned long long seed, unsigned long long stream_base, int K, long long P, void* stream);
/* the sampler, exposed for the distribution / known-answer tests (torch.normal at privacy.py:212) */
int flb_philox_normal(float* out, long long n, unsigned long long seed, unsigned long long stream_id, void* stream);
int flb_philox_raw(uint32_t* out, long long nblocks, unsigned long long seed, unsigned long long stream_id,
                   unsigned long long first_block, void* stream);

/* ---- update codec: src/shared/compression.py:203-244 (QuantizationCompressor) ---------------- */
/* scratch: 2*K*L uint32.  scale/zp: [K*L] outputs.  bits in 1..8, codes unpacked in uint8 as upstream. */
int flb_q8_quantize(const float* x, long long ld, const long long* seg_off, uint8_t* q, long long ldq,
                    float* scale, float* zp, uint32_t* scratch, int K, int L, long long P,
                    int bits, int symmetric, void* stream);
int flb_q8_dequantize(const uint8_t* q, long long ldq, const long long* seg_off, const float* scale,
                      const float* zp, float* out, long long ld, int K, int L, long long P, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLB_H_ */
