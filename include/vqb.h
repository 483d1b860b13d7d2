/* vqb.h — C ABI of libvqvae_b200.so: the B200 (sm_100a) kernels under the Keras-style layer API of the
 * VQ-VAE audio hot path.
 *
 * The reference (sunzeyucmu/VAE-based-Music--Deep-Generative-Models) has no FFI of its own: the hot path sits
 * behind Keras Layer/Model classes whose arithmetic is dispatched, op by op, to TensorFlow kernels.  Each entry
 * point below replaces one such group of TF op dispatches; the reference call site it replaces is cited as
 * (file:line) into the reference tree.
 *
 * Conventions
 *   - every function returns 0 (VQB_OK) or a negative VQB_ERR_* code; nothing throws; the message of the last
 *     error on the calling thread is returned by vqb_last_error().
 *   - all pointers are DEVICE pointers owned by the caller (fp32 unless stated), alive until `stream` has passed
 *     the call; functions never allocate device memory, never synchronise and only enqueue work on `stream`
 *     (a cudaStream_t passed as void*), so every call is CUDA-graph capturable.
 *   - activation / gradient tensors must be 16-byte aligned; weights, biases and codebooks only 4-byte aligned
 *     (they may be slices of one packed parameter buffer).
 *   - activations are channels-last [B, L, C] fp32 (Keras layout); Conv1D kernels are [k, Cin, Cout],
 *     Conv1DTranspose kernels [k, Cout, Cin] (Keras layouts); the codebook is [D, K] column-per-code
 *     (VectorQuantizer.py:38-44); code indices are int64 (tf.argmin default).
 *   - padding is TF "SAME": out = ceil(L/stride), pad = max((out-1)*stride + (k-1)*dilation + 1 - L, 0),
 *     left = pad/2.  Conv1DTranspose output length is L*stride, left crop (k-stride)/2.
 *   - there is no CPU fallback: on a device that is not sm_100 every compute entry point returns VQB_ERR_ARCH.
 */
#ifndef VQB_H_
#define VQB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQB_VERSION 100

enum {
  VQB_OK = 0,
  VQB_ERR_INVALID = -1,       /* bad argument (shape, null pointer, unsupported size) */
  VQB_ERR_ARCH = -2,          /* device is not compute capability 10.x */
  VQB_ERR_CUDA = -3,          /* a CUDA runtime call failed; see vqb_last_error() */
  VQB_ERR_WORKSPACE = -4,     /* workspace too small */
  VQB_ERR_UNIMPLEMENTED = -5
};

/* arithmetic used for the contraction */
enum {
  VQB_PREC_FP32 = 0, /* fp32 FMA on CUDA cores, exact fp32 accumulation */
  VQB_PREC_TF32 = 1, /* tcgen05 kind::tf32, fp32 accumulate in TMEM */
  VQB_PREC_BF16 = 2, /* tcgen05 kind::f16 (bf16 operands), fp32 accumulate in TMEM */
  VQB_PREC_BF16X2 = 3, /* operands split into 2 bf16 pieces (hi + lo), 3 MMAs per product: ~2^-16 */
  VQB_PREC_BF16X3 = 4, /* operands split into 3 bf16 pieces (all 24 mantissa bits), 6 MMAs per product: fp32-grade
                          products, fp32 accumulation — the tensor-core path that meets the fp32 parity contract */
  VQB_PREC_FP16X2 = 5  /* fp32-grade at the cost of bf16x2: operands scaled by a power of two (per tile for activations,
                          per convolution for weights) and split into 2 fp16 pieces (11 + 11 mantissa bits), 3 piece
                          products, fp32 accumulation.  Implemented by the residual-block kernels; every other
                          tensor-core kernel runs its bf16x3 variant under this setting */
};

int vqb_version(void);
const char* vqb_last_error(void);
/* number of CUDA kernels this library has enqueued in this process (diagnostics; bench.py's gpu_launches) */
int64_t vqb_kernel_launch_count(void);
/* 0 if `device` is a compute-capability-10.x GPU, VQB_ERR_ARCH otherwise (also caches the answer). */
int vqb_device_check(int device);

/* ------------------------------------------------------------------------------------------------------------
 * Conv1D / Conv1DTranspose   (replace layers.Conv1D: resnet.py:13,17; encdec.py:33,38,60,148 and
 *                             layers.Conv1DTranspose: encdec.py:67-68, plus their tape.gradient kernels,
 *                             vqvae.py:143)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct vqb_conv_desc {
  int32_t B;         /* batch */
  int32_t L;         /* INPUT length of the forward op */
  int32_t C_in;      /* forward input channels */
  int32_t C_out;     /* forward output channels */
  int32_t k;         /* taps */
  int32_t stride;    /* forward stride (Conv1D: decimation; Conv1DTranspose: upsampling) */
  int32_t dilation;  /* Conv1D only; must be 1 for Conv1DTranspose */
  int32_t relu_in;   /* 1: the op consumes ReLU(x) (fused pre-activation, resnet.py:12,16) */
  int32_t precision; /* VQB_PREC_* */
} vqb_conv_desc;

/* 1 if d->precision has a kernel for this shape and op (0 = forward, 1 = data gradient, 2 = weight gradient); fp32: always.
 * Tensor-core (bf16 family) kernels: forward / data gradient of k = 4, stride 2, 32 -> 32 convolutions (no fused ReLU,
 * residual or dx_add); weight gradient of k = 3, stride 1, 32 -> 32 convolutions with dilation <= 32. */
int vqb_conv1d_supports(const vqb_conv_desc* d, int op);
/* the same question for Conv1DTranspose (tensor-core kernels: forward / data gradient of k = 4, stride 2, 32 -> 32) */
int vqb_conv1d_transpose_supports(const vqb_conv_desc* d, int op);
/* y[B, ceil(L/stride), C_out] = conv(act(x)) + bias (+ residual, same shape as y; may be NULL) */
int vqb_conv1d_fwd(const vqb_conv_desc* d, const float* x, const float* w, const float* bias,
                   const float* residual, float* y, void* stream);
/* dx[B, L, C_in] = conv^T(dy) (* (x > 0) if d->relu_in; x may be NULL otherwise) (+ dx_add, may be NULL) */
int vqb_conv1d_dgrad(const vqb_conv_desc* d, const float* dy, const float* w, const float* x,
                     const float* dx_add, float* dx, void* stream);
/* dw[k, C_in, C_out] = sum_{b,t} act(x)[..] dy[..];  dbias[C_out] = sum dy (dbias may be NULL) */
size_t vqb_conv1d_wgrad_workspace_bytes(const vqb_conv_desc* d);
int vqb_conv1d_wgrad(const vqb_conv_desc* d, const float* x, const float* dy, float* dw, float* dbias,
                     void* workspace, size_t workspace_bytes, void* stream);

/* Optional batching of the weight-gradient reductions: between vqb_reduce_begin() and vqb_reduce_flush(stream) (same host
 * thread) the *_wgrad calls only enqueue their partial-sum kernels and MAY leave dw / dbias unwritten (the tensor-core kernels of
 * the k = 3, 32 -> 32 convolutions reduce inside their own launch and write at once); vqb_reduce_flush launches
 * all pending fixed-order reductions as a few batched kernels.  Workspaces passed to those calls must stay alive until the
 * flush.  Without begin/flush every *_wgrad call is self-contained. */
int vqb_reduce_begin(void);
int vqb_reduce_flush(void* stream);

/* y[B, L*stride, C_out] = convT(x) + bias ; w is [k, C_out, C_in] */
int vqb_conv1d_transpose_fwd(const vqb_conv_desc* d, const float* x, const float* w, const float* bias,
                             float* y, void* stream);
int vqb_conv1d_transpose_dgrad(const vqb_conv_desc* d, const float* dy, const float* w, float* dx,
                               void* stream);
size_t vqb_conv1d_transpose_wgrad_workspace_bytes(const vqb_conv_desc* d);
int vqb_conv1d_transpose_wgrad(const vqb_conv_desc* d, const float* x, const float* dy, float* dw,
                               float* dbias, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fused pre-activation residual block  (replaces ResnetConv1DBlock.call, resnet.py:11-18,29:
 *   y = x + Conv1D_k3(ReLU(Conv1D_k3,dil(ReLU(x)))), both SAME, C -> F -> C channels)
 * h [B,L,F] receives the first conv's output (pre-ReLU) for the backward pass; on the tensor-core precisions h may be NULL
 * (inference): the intermediate then never leaves the SM (256 instead of 384 bytes per position).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct vqb_resblock_desc {
  int32_t B, L, C, F, dilation;
  int32_t precision;
} vqb_resblock_desc;

/* 1 if the (shape, precision) combination has a kernel (fp32: always; bf16/tf32 tensor-core path: C = F = 32,
 * dilation <= 32), else 0 — callers pick VQB_PREC_FP32 for the rest. */
int vqb_resblock_supports(const vqb_resblock_desc* d);
int vqb_resblock_fwd(const vqb_resblock_desc* d, const float* x, const float* w1, const float* b1,
                     const float* w2, const float* b2, float* h, float* y, void* stream);
/* given dy: dx = dy + (x>0)*conv1^T((h>0)*conv2^T(dy)); dh ([B,L,F] scratch, also an output) holds
 * (h>0)*conv2^T(dy), the gradient at conv1's output, for the weight gradients. */
int vqb_resblock_bwd_data(const vqb_resblock_desc* d, const float* x, const float* h, const float* dy,
                          const float* w1, const float* w2, float* dh, float* dx, void* stream);
/* Sign-mask variants (tensor-core precisions / shapes only, see vqb_resblock_supports): the forward additionally writes
 * xbits[B, L] and hbits[B, L] — bit c of word t is set iff channel c of x (resp. h) at position t is > 0 — and the data
 * gradient reads those 8 bytes per position instead of the fp32 tensors x and h (640 -> 392 bytes per position). */
int vqb_resblock_fwd_masks(const vqb_resblock_desc* d, const float* x, const float* w1, const float* b1, const float* w2,
                           const float* b2, float* h, float* y, uint32_t* xbits, uint32_t* hbits, void* stream);
int vqb_resblock_bwd_data_masks(const vqb_resblock_desc* d, const uint32_t* xbits, const uint32_t* hbits, const float* dy,
                                const float* w1, const float* w2, float* dh, float* dx, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fused DilatedResnet1D  (replaces DilatedResnet1D.call, resnet.py:40-59: a Sequential of n ResnetConv1DBlocks whose first
 * convolutions have dilations[0..n-1], in execution order — growing for encoders, reversed for decoders, resnet.py:44-55)
 * as ONE launch: the input tile is fetched by TMA, all 2n convolutions run on the tensor cores with the activation staying
 * in shared / tensor memory, only what the caller asks for goes back to HBM (TMA stores).
 *   vqb_resstack_fwd       y[n-1] = stack(x).  Inference: h = xbits = hbits = NULL and y[i < n-1] = NULL: 256 bytes per
 *                          position for the WHOLE stack.  Under a tape: h[i] (first convolution's output of block i), y[i]
 *                          (block outputs = inputs of the next block) and the sign masks xbits[i] / hbits[i] (as
 *                          vqb_resblock_fwd_masks) are all written — the operands of the weight gradients and of
 *                          vqb_resstack_bwd_data.
 *   vqb_resstack_bwd_data  given dy = gradient at y[n-1]: dx[i] = gradient at the input of block i, dh[i] = gradient at the
 *                          output of block i's first convolution (both needed by vqb_resblock_wgrad_batch), blocks n-1 .. 0
 *                          chained on chip.
 * vqb_resstack_supports: 1 when the fused kernel exists for (C, n_blocks, dilations, precision) — C = 32, n <= 4,
 * dilations <= 32, VQB_PREC_FP16X2 — else 0 and callers compose vqb_resblock_* (the functions below return
 * VQB_ERR_UNIMPLEMENTED).  The workspace holds the packed operand images of the 2n weight tensors.
 * ---------------------------------------------------------------------------------------------------------- */
#define VQB_RESSTACK_MAX_BLOCKS 4
typedef struct vqb_resstack_desc {
  int32_t B, L, C;
  int32_t n_blocks;
  int32_t dilations[VQB_RESSTACK_MAX_BLOCKS];
  int32_t precision;
} vqb_resstack_desc;
int vqb_resstack_supports(const vqb_resstack_desc* d);
size_t vqb_resstack_workspace_bytes(const vqb_resstack_desc* d);
int vqb_resstack_fwd(const vqb_resstack_desc* d, const float* x, const float* const* w1, const float* const* b1,
                     const float* const* w2, const float* const* b2, float* const* h, float* const* y,
                     uint32_t* const* xbits, uint32_t* const* hbits, void* workspace, size_t workspace_bytes, void* stream);
/* vqb_resstack_fwd for callers that OWN the workspace (one buffer per DilatedResnet1D, never handed to anything else, e.g. a
 * member of the layer object): no earlier kernel of the stream can be using that memory, so the weight-packing launch MAY run
 * under programmatic dependent launch next to the previous kernel (VQB_RS_EARLY_PACK=1; measured slower on B200, off by
 * default).  Same results as vqb_resstack_fwd. */
int vqb_resstack_fwd_private_ws(const vqb_resstack_desc* d, const float* x, const float* const* w1, const float* const* b1,
                                const float* const* w2, const float* const* b2, float* const* h, float* const* y,
                                uint32_t* const* xbits, uint32_t* const* hbits, void* workspace, size_t workspace_bytes,
                                void* stream);
int vqb_resstack_bwd_data(const vqb_resstack_desc* d, const float* dy, const float* const* w1, const float* const* w2,
                          const uint32_t* const* xbits, const uint32_t* const* hbits, float* const* dh, float* const* dx,
                          void* workspace, size_t workspace_bytes, void* stream);
/* The same, running from the workspace of the vqb_resstack_fwd call (under a tape) of the same blocks and weights: that call
 * packs the operand images of both directions, so the backward pass of a step needs no packing launch.  The caller keeps the
 * forward's workspace alive and unmodified until this call has run (tape.gradient, vqvae.py:142-143). */
int vqb_resstack_bwd_data_packed(const vqb_resstack_desc* d, const float* dy, const uint32_t* const* xbits,
                                 const uint32_t* const* hbits, float* const* dh, float* const* dx, void* fwd_workspace,
                                 size_t workspace_bytes, void* stream);

/* Both weight gradients of the block in one call (tensor-core precisions: one launch):
 *   dw1[3, C, F] = sum ReLU(x)[t + (j-1) dilation] dh[t],  db1[F] = sum dh   (dilated conv, resnet.py:13-15)
 *   dw2[3, F, C] = sum ReLU(h)[t + (j-1)] dy[t],           db2[C] = sum dy   (second conv, resnet.py:17)
 * Takes part in vqb_reduce_begin / vqb_reduce_flush batching like the *_wgrad calls. */
size_t vqb_resblock_wgrad_workspace_bytes(const vqb_resblock_desc* d);
int vqb_resblock_wgrad(const vqb_resblock_desc* d, const float* x, const float* h, const float* dy, const float* dh,
                       float* dw1, float* db1, float* dw2, float* db2, void* workspace, size_t workspace_bytes,
                       void* stream);
/* ... of n blocks of the same shape in one call (1 <= n <= 4 for one launch on the tensor-core paths; e.g. the blocks of a
 * DilatedResnet1D once its data-gradient chain is through).  Arrays of n pointers; d->dilation is ignored, dilations[i] is
 * block i's.  db1[i] / db2[i] may be NULL. */
size_t vqb_resblock_wgrad_batch_workspace_bytes(const vqb_resblock_desc* d, int32_t n);
int vqb_resblock_wgrad_batch(const vqb_resblock_desc* d, int32_t n, const int32_t* dilations, const float* const* x,
                             const float* const* h, const float* const* dy, const float* const* dh, float* const* dw1,
                             float* const* db1, float* const* dw2, float* const* db2, void* workspace,
                             size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Decoder tail: the last Conv1DTranspose(C_mid, k=4, s=2) (encdec.py:67-68) and the final Conv1D(1, 3) (encdec.py:148)
 * as ONE composed linear operator (no activation separates them), forward and backward, exact fp32.  The
 * [B, 2L, C_mid] intermediate — the largest tensor of the model — is never formed.
 *   x [B, L, C_in]; wt [4, C_mid, C_in], bt [C_mid] (may be NULL); wf [3, C_mid, 1], bf [1] (may be NULL)
 *   recon [B, 2L, 1];  gbuf [VQB_TAIL_GBUF floats]: composed taps written by the forward call, read by the backward call.
 * Requires C_in a multiple of 4 in [4, 32] (vqb_dec_tail_supports); other shapes run the two layers separately.
 * ---------------------------------------------------------------------------------------------------------- */
#define VQB_TAIL_GBUF 272
typedef struct vqb_tail_desc {
  int32_t B, L, C_in, C_mid;
} vqb_tail_desc;
int vqb_dec_tail_supports(const vqb_tail_desc* d);
int vqb_dec_tail_fwd(const vqb_tail_desc* d, const float* x, const float* wt, const float* bt, const float* wf,
                     const float* bf, float* gbuf, float* recon, void* stream);
size_t vqb_dec_tail_bwd_workspace_bytes(const vqb_tail_desc* d);
/* given drecon [B, 2L, 1]: dx [B, L, C_in] (may be NULL) and the gradients of both layers' parameters
 * (dwt [4, C_mid, C_in], dbt [C_mid] or NULL, dwf [3, C_mid, 1], dbf [1] or NULL), all OVERWRITTEN. */
int vqb_dec_tail_bwd(const vqb_tail_desc* d, const float* x, const float* drecon, const float* wt, const float* bt,
                     const float* wf, const float* gbuf, float* dx, float* dwt, float* dbt, float* dwf, float* dbf,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * VectorQuantizer  (replaces VectorQuantizer.call / get_code_indices, VectorQuantizer.py:75-186)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct vqb_vq_desc {
  int64_t N;         /* latents (B * T/hop) */
  int32_t D;         /* embedding_dim */
  int32_t K;         /* num_embeddings */
  float beta;        /* commitment weight (VectorQuantizer.py:97) */
  int32_t precision; /* VQB_PREC_FP32: exact fp32 search; VQB_PREC_BF16 / TF32: tensor-core search with exact
                        fp32 re-ranking of the best candidates */
} vqb_vq_desc;

size_t vqb_vq_fwd_workspace_bytes(const vqb_vq_desc* d);
/* x [N,D], E [D,K] ->
 *   idx    [N] int64            nearest code, first minimum (VectorQuantizer.py:173-185)
 *   q_st   [N,D]                x + (E[:,idx] - x)          (:86-90,114), may be NULL
 *   q      [N,D]                E[:,idx]                    (:88) may be NULL
 *   loss   [1]                  beta * mean((q - x)^2)      (:97-99)
 *   m_batch[D,K], n_batch[K]    X^T onehot, sum(onehot)     (:123-124); both may be NULL (inference)
 * m_batch / n_batch are OVERWRITTEN (zeroed inside the call). */
int vqb_vq_fwd(const vqb_vq_desc* d, const float* x, const float* E, int64_t* idx, float* q_st, float* q,
               float* loss, float* m_batch, float* n_batch, void* workspace, size_t workspace_bytes,
               void* stream);
/* The same layer with bfloat16 ACTIVATIONS (BASELINE.json configs[3] "bf16"): x, q_st and q are [N,D] bfloat16 (raw 16-bit
 * patterns); the codebook, the search, the loss and the batch statistics are the fp32 arithmetic of vqb_vq_fwd applied to
 * float(x), and q_st = bf16_rn(x + (q - x)), q = bf16_rn(E[:,idx]).  The reference layer itself is fp32 end to end
 * (VectorQuantizer.py:75-124 under Keras' default float32 policy); this entry halves the layer's HBM traffic for callers that
 * keep bf16 latents.  Tensor-core search: K <= 4096. */
int vqb_vq_fwd_bf16(const vqb_vq_desc* d, const uint16_t* x, const float* E, int64_t* idx, uint16_t* q_st, uint16_t* q,
                    float* loss, float* m_batch, float* n_batch, void* workspace, size_t workspace_bytes,
                    void* stream);
/* straight-through + commitment gradient (VectorQuantizer.py:97-99,114):
 *   dx = dq_out + loss_scale * (2*beta/(N*D)) * (x - q)   (loss_scale: d total / d commitment loss) */
int vqb_vq_bwd(const vqb_vq_desc* d, const float* dq_out, const float* x, const float* q, float loss_scale,
               float* dx, void* stream);
/* EMA codebook update with dead-code restart (VectorQuantizer.py:128-159), in place on E, m_t, N_t.
 *   restart_rows [K,D]: the first K rows of shuffle(_tile(flattened)) (:137) chosen by the caller.
 *   metrics [3] = {batch usage (:151), running usage (:153), entropy (:157-158)}.
 * Every multiply and add is separately rounded (no FMA contraction), as in eager TF.  gamma is a double because
 * the reference forms (1. - gamma) in Python double arithmetic before TF casts it to fp32 (:128,131). */
int vqb_vq_ema_update(int32_t D, int32_t K, double gamma, float threshold, const float* m_batch,
                      const float* n_batch, const float* restart_rows, float* E, float* m_t, float* N_t,
                      float* metrics, void* stream);
/* the `_tile` + shuffle[:K] row pick (VectorQuantizer.py:137,191-199) for a batch that may be sharded over ranks:
 *   r = ids[i] mod N_total;  rows[i,:] = x[r - row_offset, :] if row_offset <= r < row_offset + N_local else 0
 * (single GPU: N_total = N_local, row_offset = 0; under data parallelism the per-rank results are summed). */
int vqb_gather_rows(const float* x, int64_t N_local, int32_t D, const int64_t* ids, int32_t n_ids,
                    int64_t N_total, int64_t row_offset, float* rows, void* stream);
/* ids[i] = (a*i + c) mod Nt, i < K, with (a, c) hashed from (seed, *step_counter): K distinct pseudo-random
 * row numbers of the tiled batch, identical on every rank that shares seed and step.  Nt = max(N, K rounded up
 * to a multiple of N) as `_tile` produces. */
int vqb_restart_ids(int64_t N, int32_t K, uint64_t seed, const int64_t* step_counter, int64_t* ids,
                    void* stream);
/* out[i,:] = E[:, idx[i]]   (decode path gather, vqvae.py:248) */
int vqb_gather_codes(const float* E, int32_t D, int32_t K, const int64_t* idx, int64_t n, float* out,
                     void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Multi-scale spectral convergence loss (vqvae.py:309-326, data_utils.py:25-40): the element-wise halves around the
 * FFT (cuFFT, called by the host layer).  tf.signal.stft semantics: F = 1 + (T - win) / hop frames of `win` samples,
 * periodic Hann window, zero padding at the END up to n_fft.  Spectra are complex64 (interleaved re, im),
 * [B, per_example = F * bins] with bins = n_fft / 2 + 1.
 * ---------------------------------------------------------------------------------------------------------- */
/* frames [B, F, n_fft] = windowed, zero-padded frames of x [B, T] */
int vqb_stft_frames(const float* x, int64_t B, int32_t T, int32_t n_fft, int32_t hop, int32_t win, float* frames,
                    void* stream);
size_t vqb_spec_workspace_bytes(int64_t B, int64_t per_example);
/* mag = |S|, sums[b] = sum |S[b]|^2   (target side; cached by the caller) */
int vqb_spec_mag(const float* S, int64_t B, int64_t per_example, float* mag, float* sums, void* workspace,
                 size_t workspace_bytes, void* stream);
/* sums[b] = sum (mag_t[b] - |S[b]|)^2 */
int vqb_spec_diff(const float* S, const float* mag_t, int64_t B, int64_t per_example, float* sums, void* workspace,
                  size_t workspace_bytes, void* stream);
/* dsum, tsum [nscales, B] -> loss[0] = mean_b mean_s sqrt(dsum) / sqrt(tsum); coef [nscales, B] (may be NULL) =
 * d loss / d(sum-of-squares root) factors consumed by vqb_spec_grad */
int vqb_spec_loss(const float* dsum, const float* tsum, int32_t nscales, int32_t B, float* loss, float* coef,
                  void* stream);
/* G [B, F, bins] complex64: the UNNORMALISED inverse real FFT of G (torch.fft.irfft(G, n = n_fft, norm = "forward"), cuFFT C2R)
 * is d(upstream[0] * loss) / d frames (coef = this scale's row) */
int vqb_spec_grad(const float* S, const float* mag_t, const float* coef, const float* upstream, int64_t B,
                  int64_t per_example, int32_t bins, int32_t n_fft, float* G, void* stream);
/* dx [B, T] (+)= overlap-add of dframes [B, F, n_fft] * window (accumulate != 0: add to dx) */
int vqb_stft_frames_bwd(const float* dframes, int64_t B, int32_t T, int32_t n_fft, int32_t hop, int32_t win,
                        int32_t accumulate, float* dx, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * loss head + optimiser pieces of VQVAE.train_step (vqvae.py:125,143-144)
 * ---------------------------------------------------------------------------------------------------------- */
size_t vqb_reduce_workspace_bytes(int64_t n);
/* loss[0] = mean((x - r)^2); if dr != NULL: dr = loss_scale * 2 (r - x) / n  (+ dr_add if non-NULL) */
int vqb_mse(const float* x, const float* r, int64_t n, float loss_scale, const float* dr_add, float* loss,
            float* dr, void* workspace, size_t workspace_bytes, void* stream);
/* Keras-2.7 Adam on a flat buffer, step taken from *step_counter (device int64, 1-based step = counter+1; the
 * kernel does NOT increment it): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m += (g*gs - m)(1-b1); v += ((g*gs)^2 - v)(1-b2);
 * p -= lr_t*m/(sqrt(v)+eps) */
int vqb_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2,
                  float eps, float grad_scale, const int64_t* step_counter, void* stream);
/* The same step with the learning rate read from device memory (lr_dev[0], fp32) when the kernel runs: a captured CUDA
 * graph then follows `optimizer.learning_rate = ...` / LearningRateSchedule objects (src/callback/vae_monitor.py trains
 * with an lr scheduler) without re-capture. */
int vqb_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* lr_dev, float b1, float b2,
                      float eps, float grad_scale, const int64_t* step_counter, void* stream);
int vqb_increment(int64_t* counter, void* stream);
/* out[i] = sum_{j in [term_start[i], term_start[i+1])} term_coef[j] * term_ptr[j][0],  i < n_out <= 32, <= 96 terms: the loss
 * bookkeeping of train_step (vqvae.py:127-146: level_loss = recon + commit + spectral, the sums over levels, the metric
 * increments) as ONE launch instead of one library kernel per `+`.  term_start / term_ptr / term_coef are HOST arrays (term_ptr
 * holds device pointers); the terms of an output are added in order. */
int vqb_lincomb(int32_t n_out, const int32_t* term_start, const float* const* term_ptr, const float* term_coef, float* out,
                void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VQB_H_ */
