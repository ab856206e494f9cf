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

/* L2 residency hint (no reference counterpart; a property of this implementation's memory layout): marks [base, base + bytes)
 * as persisting in L2 for kernels launched on `stream` from now on (captured into CUDA-graph kernel nodes); bytes <= 0, or a
 * range larger than the device's L2 set-aside, clears the window.  Used for the Adam moments of the batched trainer.
 * Returns the bytes covered (0: no window). */
long long flb_l2_persist_window(const void* base, long long bytes, void* stream);

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

/* ---- multi-GPU FedAvg: the rank's partial sum fused with the cross-GPU reduction over NVLink peer memory ------------------
 * One process per GPU, client i on rank i mod G (the reference's client -> coordinator hop, src/client/grpc_client.py ->
 * src/aggregation/fedavg.py:56-124, then the coordinator -> client broadcast of the new global model).  Every rank owns one
 * device region (flb_p2p_alloc, laid out by flb_p2p_region_layout), exports it (flb_p2p_export: a 64-byte cudaIpcMemHandle the
 * caller ships to the other processes, e.g. with torch.distributed.all_gather) and maps the peers' regions (flb_p2p_open).
 * flb_fedavg_allreduce_p2p then leaves  sum_ranks sum_k w[k] * theta[k]  in the `global` row (region + off_global) of EVERY
 * rank, bit-identical on all of them (rank-ordered sums): ONE kernel per rank, the partial sums travel chunk by chunk while
 * later chunks are still being summed (csrc/p2p_reduce.cu).  `epoch` must be >= 1, the same on all ranks for one call and
 * larger for every later call.  All ranks must call it for the collective to complete (like any collective). */
#define FLB_P2P_MAX_RANKS 16
#define FLB_P2P_HANDLE_BYTES 64
typedef struct flb_p2p_layout {
    long long bytes, off_flags_a, off_flags_b, off_inbox, off_global, ld;
    int world, chunk;
} flb_p2p_layout;
/* fills *out for rows of ld floats (ld % 4 == 0), `world` ranks and chunks of `chunk` floats; returns the region size or -1 */
long long flb_p2p_region_layout(long long ld, int world, int chunk, flb_p2p_layout* out);
int flb_p2p_alloc(void** ptr, long long bytes);                 /* cudaMalloc + zero; synchronises the device */
int flb_p2p_free(void* ptr);
int flb_p2p_export(const void* ptr, unsigned char* handle64);
int flb_p2p_open(const unsigned char* handle64, void** ptr);    /* enables peer access to the exporting device */
int flb_p2p_close(void* ptr);
/* regions: HOST array of lay->world device pointers, regions[rank] being the caller's own region */
int flb_fedavg_allreduce_p2p(const float* theta, long long ld_theta, const float* w, int K, long long P,
                             const flb_p2p_layout* lay, void* const* regions, int rank, unsigned int epoch, void* stream);

/* ---- update-level DP: src/shared/privacy.py:107-144,183-254 + src/client/federated_trainer.py:428-469 */
/* norm2[k] = sum_p (local[k*ld+p] - global[p])^2 in double; global may be NULL (local is the delta). */
int flb_dp_sumsq(const float* local, long long ld, const float* global_w, double* norm2,
                 int K, long long P, void* stream);
/* out[k] = global + clip(local[k]-global)*1 + sigma_k*z ; sigma_k = min(||delta_k||, max_norm)*sigma_unit,
 * sigma_unit = sqrt(2 ln(1.25/delta))/epsilon (0 disables noise).  z_in (may be NULL) injects the standard
 * normals; otherwise z = Philox4x32-10(seed, stream_base + k*stream_stride, element).  norms_out[k] (may be NULL) = ||delta_k||. */
int flb_dp_clip_noise(const float* local, long long ld, const float* global_w, const float* z_in,
                      const double* norm2, float* out, float* norms_out, double max_norm,
                      double sigma_unit, unsigned long long seed, unsigned long long stream_base,
                      unsigned long long stream_stride, int K, long long P, void* stream);
/* flb_dp_clip_noise that also delivers the update codec's per-(client, layer) max|out| (compression.py:207-210) as a
 * by-product of the same pass: absmax_bits[k*L + l] = bit pattern of max |out[k, seg_off[l] .. seg_off[l+1])| (monotone for
 * non-negative floats; zeroed by the call).  L <= 64.  Feeds flb_q8_quantize_absmax, which then skips its reduction pass. */
int flb_dp_clip_noise_absmax(const float* local, long long ld, const float* global_w, const float* z_in,
                             const double* norm2, float* out, float* norms_out, double max_norm,
                             double sigma_unit, unsigned long long seed, unsigned long long stream_base,
                             unsigned long long stream_stride, const long long* seg_off, int L,
                             unsigned int* absmax_bits, int K, long long P, void* stream);
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
/* symmetric quantisation from a known per-(client, layer) max|x| (absmax_bits as written by flb_dp_clip_noise_absmax) */
int flb_q8_quantize_absmax(const float* x, long long ld, const long long* seg_off, const unsigned int* absmax_bits,
                           uint8_t* q, long long ldq, float* scale, float* zp, int K, int L, long long P,
                           int bits, void* stream);
int flb_q8_dequantize(const uint8_t* q, long long ldq, const long long* seg_off, const float* scale,
                      const float* zp, float* out, long long ld, int K, int L, long long P, void* stream);

/* ---- one-pass update validation / convergence reductions --------------------------------------
 * src/shared/validation.py:72-91 (isnan / isinf / abs().max() per tensor: 3 passes + 3 syncs each upstream) and
 * src/aggregation/fedavg.py:144-190, src/aggregation/convergence.py:189-217 (per-layer ||new - old||, ||new||).
 * ptrs[k*L + l] = device pointer of tensor l of client k (seg_off gives the element counts).
 * max_abs[k*L + l] = max |x| (NaN ignored), flags[k*L + l] = 1 if any NaN | 2 if any Inf. */
int flb_update_stats(const float* const* ptrs, const long long* seg_off, float* max_abs, unsigned int* flags,
                     int K, int L, long long P, void* stream);
/* out[2l] = sum (new_l - old_l)^2, out[2l+1] = sum new_l^2, in double */
int flb_delta_norms(const float* const* new_ptrs, const float* const* old_ptrs, const long long* seg_off,
                    double* out, int L, long long P, void* stream);

/* ---- device-resident shard builder: src/shared/data_loader.py:298-306,454-464 + training.py:186 ----
 * out[j, c, h, w] = ((raw[idx[j]] / 255) - mean[c]) / std[c] (ToTensor + Normalize, fp32 step by step);
 * raw is uint8 [n_raw, H, W, C] when hwc != 0 (CIFAR10.data) else [n_raw, C, H, W] (MNIST.data with C = 1). */
int flb_gather_normalize_u8(const uint8_t* raw, long long n_raw, int H, int W, int C, int hwc, const long long* idx,
                            long long M, const float* mean, const float* stdv, float* out, void* stream);
int flb_gather_labels(const long long* labels, const long long* idx, int* out, long long M, void* stream);

/* ---- top-k sparsification: src/shared/compression.py:327-365 (TopKSparsificationCompressor) ----------------
 * For every (client c, layer l): the kk[l] entries of largest |x| of x[c*ld + seg_off[l] .. seg_off[l+1]) as
 * (index relative to the layer start, value) pairs at [c*ldk + out_off[l] ..), in INDEX order; ties at the threshold keep
 * the lowest indices.  out_off = exclusive prefix sums of kk (L+1 entries), ldk >= out_off[L].  Exact radix select. */
int flb_topk_select(const float* x, long long ld, const long long* seg_off, const int* kk,
                    const long long* out_off, int* idx_out, float* val_out, long long ldk,
                    int K, int L, void* stream);
/* inverse (_desparsify_tensor :346-365): dense[c, 0..P) = 0, then dense[c, seg_off[l] + idx] = val */
int flb_topk_scatter(const int* idx, const float* val, long long ldk, const long long* seg_off, const int* kk,
                     const long long* out_off, float* dense, long long ld, int K, int L, long long P, void* stream);

/* ---- batched local training: src/shared/training.py:60-212 (LocalTrainer._train_epoch), ------------------
 *      models src/shared/models_pytorch.py:59-97 (SimpleCNN, model 0) and :100-165 (CIFAR10CNN, model 1),
 *      optimizers training.py:244-255, evaluation training.py:214-242,307-360 (eval_mode) -----------------------
 * One call advances EVERY resident client by one minibatch step (zero_grad -> forward -> mean cross-entropy ->
 * backward -> optimizer step, training.py:189-197).  All pointers are device pointers; the struct itself is
 * host memory and is read at call time only. */
typedef struct flb_train_args {
    /* data: the clients' samples concatenated; client k owns rows sample_off[k] .. sample_off[k]+nsamples[k] */
    const float* x;               /* [sum N_c, C*H*W] fp32, NCHW per sample                          */
    const int* y;                 /* [sum N_c] labels                                                */
    const long long* sample_off;  /* [K]                                                             */
    const int* nsamples;          /* [K]                                                             */
    int* step_ctr;                /* [2] [0] current minibatch index within the epoch (device-advanced);
                                     [1] scratch ticket counter of the optimizer kernel (kept at 0)      */
    /* state, client-major rows of pitch ld floats, reference layer order/layouts                    */
    float* W;                     /* [K, ld] parameters                                              */
    float* G;                     /* [K, ld] gradients (scratch)                                     */
    float* M;                     /* [K, ld] Adam exp_avg / SGD momentum buffer                      */
    float* V;                     /* [K, ld] Adam exp_avg_sq                                         */
    int* tcount;                  /* [K] optimizer steps taken (device-advanced)                     */
    void* ws;                     /* activation workspace, flb_train_ws_bytes() bytes, zeroed once   */
    /* epoch accumulators (training.py:200-212): sum of batch-mean losses, argmax hits, batches, samples */
    float* loss_sum;              /* [K] */
    int* correct;                 /* [K] */
    int* nbatch;                  /* [K] */
    int* nseen;                   /* [K] */
    const unsigned char* drop_keep; /* optional injected dropout keep-mask [K, B, 128] (NULL: Philox)  */
    const float* dp_z;            /* optional injected standard normals [K, ld] for dp_mode 1 (NULL: Philox) */
    float* bn_running;            /* cifar10_cnn: client-local BatchNorm buffers [K, flb_train_bn_floats()] =
                                     running_mean of bn1..bn6 then running_var of bn1..bn6 (never federated,
                                     models_pytorch.py:25-27); NULL for simple_cnn                            */
    unsigned long long* epoch_nonce; /* [1] device counter, +1 by every flb_train_begin_epoch and never reset: folded into
                                     the Philox key of dropout masks and per-sample-DP noise so that no epoch / round /
                                     train_local_model call ever re-draws an earlier one's randomness (NULL: 0)        */
    long long ld;
    unsigned long long seed;      /* Philox seed for dropout and per-sample-DP noise                 */
    unsigned long long client_base; /* global index of local client 0 (Philox stream = client_base + k*client_stride) */
    unsigned long long client_stride; /* global-index distance between consecutive local clients (= world size) */
    double lr, beta1, beta2, eps, weight_decay, momentum;   /* torch.optim defaults are Python doubles */
    int model;                    /* 0 = simple_cnn, 1 = cifar10_cnn (models_pytorch.py:59-97, :100-165) */
    int K;                        /* resident clients                                                */
    int B;                        /* batch size, 1..32: the kernels map one batch to one 32-column MMA operand / one
                                     warp of per-sample lanes; LocalTrainer trains larger loader batches in minibatches
                                     of 32 (and logs it), BatchedClientTrainer rejects them                          */
    int precision;                /* 0 = fp32 CUDA-core kernels, 1 = TF32 tcgen05 tensor-core kernels */
    int opt;                      /* 0 adam, 1 sgd(momentum), 2 adamw                                */
    int dp_mode;                  /* 0 none (reference behaviour), 1 per-sample clip + noise (cifar10_cnn: the BatchNorm
                                     batch statistics are constants of the per-sample backward pass) */
    int eval_mode;                /* 1: model.eval() semantics -- no dropout, BatchNorm uses the running statistics */
    int tc_mask;                  /* precision 1 only: bit set = that GEMM on tensor cores (0 = all).  simple_cnn: 1 conv2 fwd,
                                     2 fc1 fwd, 4 fc1 dgrad, 8 conv2 dgrad, 16 fc1 wgrad, 32 conv2 wgrad.  cifar10_cnn: bit 3*(l-1)+kind for conv
                                     layer l = 1..5 (conv2..conv6), bits 15..17 fc1, 18..20 fc2; kind 0 fwd, 1 dgrad, 2 wgrad (test aid) */
    float drop_p;                 /* dropout probability of SimpleCNN.dropout (0.25 upstream)        */
    float dp_clip;                /* per-sample max_grad_norm C                                      */
    float dp_sigma;               /* noise std of the summed clipped gradient (= C * sigma_unit)     */
} flb_train_args;

long long flb_train_ws_bytes(int model, int K, int B);
/* floats per client of the bn_running buffer (0 for simple_cnn) */
long long flb_train_bn_floats(int model);
/* pointers (as byte offsets into ws) of named workspace arrays, for tests: returns -1 if unknown */
long long flb_train_ws_offset(int model, int K, int B, const char* name);
/* zero the epoch accumulators and step counter (start of _train_epoch, training.py:178-182) */
int flb_train_begin_epoch(const flb_train_args* a, void* stream);
/* Start of a federated round: W[k] = global_row, M[k] = V[k] = 0 over the whole padded row, tcount[k] = 0 for every resident
 * client, one pass (a new LocalTrainer -- fresh torch.optim state -- around the downloaded weights,
 * src/client/federated_trainer.py:367-392).  global_row: ld floats, 16-byte aligned. */
int flb_train_begin_round(const flb_train_args* a, const float* global_row, void* stream);
/* one minibatch step for all K clients; clients that have run out of samples are skipped */
int flb_train_step(const flb_train_args* a, void* stream);
/* same step, and grads_out[k*ld_out + p] (p < P) receives the minibatch gradient the optimizer applied (what
 * LocalTrainer.get_model_gradients returns upstream, training.py:362-371: param.grad after the last step); in dp_mode 1
 * it is the sum of the clipped per-sample gradients before noise and 1/B. */
int flb_train_step_grads(const flb_train_args* a, float* grads_out, long long ld_out, void* stream);
/* kernels launched by one flb_train_step with these args (for launch accounting under CUDA-graph replay) */
int flb_train_step_launches(const flb_train_args* a);
/* profiling aid: one step with a CUDA event after every kernel; synchronises.  names_out: newline-separated labels,
 * ms_out[i]: device milliseconds of segment i.  Returns the number of segments (>= 0) or an error code. */
int flb_train_step_profiled(const flb_train_args* a, void* stream, char* names_out, int names_cap,
                            float* ms_out, int max_n);
/* forward + loss/accuracy accumulation only (LocalTrainer._validate_epoch / evaluate_model, training.py:214-242,
 * 307-360; also model.forward): logits land in the workspace array "logits" */
int flb_train_forward(const flb_train_args* a, void* stream);
/* advance the device-side minibatch counter (and optimizer step counts) without a step */
int flb_train_advance(const flb_train_args* a, void* stream);
/* forward + loss/gradients only (no optimizer, no counter advance): fills G and the workspace, for parity tests */
int flb_train_forward_backward(const flb_train_args* a, void* stream);

/* ---- profiling aid: per-role timeline of CTA 0 of the resident-weight convolution kernels (csrc/tc_gemm.cuh TraceBuf).
 * buf: device memory, 8 + 24*cap bytes = [int n][int cap][cap x (event, tile, SM clock) int64]; NULL switches it off.
 * Synchronous (cudaMemcpyToSymbol).  scripts/conv_timeline.py prints the trace of one conv2 forward / dgrad launch. */
int flb_debug_trace_set(void* buf, int cap);

/* ---- profiling aid: cycles of `reps` back-to-back tcgen05.mma kind::tf32 128 x n x 8 instructions issued from one CTA ----
 * a_shift: the A operand starts that many 128-byte rows into its 1024-byte swizzle atom (the row-shifted tap windows of the
 * halo convolutions); a_mn / b_mn: MN-major operands; rotate: distinct A windows cycled through.  cycles_out: one device
 * int64 (SM clocks from the first issue to the completion of the final commit).  scripts/mma_microbench.py sweeps it. */
int flb_mma_microbench(int n, int a_shift, int a_mn, int b_mn, int rotate, int reps, long long* cycles_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLB_H_ */
