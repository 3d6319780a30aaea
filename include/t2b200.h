/*
 * t2b200.h -- C ABI of libt2b200.so, the B200 (sm_100a) Tacotron 2 mel-spectrogram engine.
 *
 * The reference (NVIDIA/tacotron2) has no FFI / plugin layer: its hot path is ordinary Python
 * methods on nn.Modules.  This header is therefore the boundary a
 * maintainer of the reference would bind (ctypes stub in INTEGRATION.md) to replace, one for one,
 * the bodies of
 *
 *     Encoder.inference / Encoder.forward     model.py:173-201     -> t2_encoder_forward
 *     Decoder.inference                       model.py:418-454     -> t2_decoder_run (mode INFER)
 *     Decoder.forward  (teacher forcing)      model.py:381-416     -> t2_decoder_run (mode TEACHER)
 *       Prenet.forward                        model.py:97-100         (inside, per step / hoisted)
 *       Decoder.decode                        model.py:340-379        (inside, the persistent loop)
 *       Attention.forward / LocationLayer     model.py:22-26, 43-86   (inside)
 *       initialize_decoder_states             model.py:258-289        (inside: processed_memory GEMM)
 *     Postnet.forward (+ residual add)        model.py:141-146, 511, 524  -> t2_postnet_forward
 *     Tacotron2.inference, host buffers       model.py:517-529     -> t2_infer_host
 *
 * Conventions
 *   - plain C types only; every tensor argument is a raw pointer into DEVICE memory of the current
 *     CUDA device unless its name ends in _host; all float tensors are fp32, contiguous;
 *   - the caller owns every buffer; the library allocates only inside T2Model (packed weights) and
 *     never frees caller memory; scratch comes from the caller-provided workspace (size queries);
 *   - work is enqueued on the given stream (a cudaStream_t passed as void*); no call synchronises
 *     the device unless documented (t2_infer_host does, it returns host data);
 *   - every function returns 0 on success or a negative T2_ERR_* code; t2_last_error() returns a
 *     thread-local message for the last failure.  Nothing falls back to a CPU path.
 */
#ifndef T2B200_H_
#define T2B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define T2_ABI_VERSION 1

#define T2_OK               0
#define T2_ERR_INVALID     -1   /* bad argument / unsupported shape                     */
#define T2_ERR_CUDA        -2   /* a CUDA runtime call failed (message has the detail)  */
#define T2_ERR_WORKSPACE   -3   /* workspace too small                                  */
#define T2_ERR_UNSUPPORTED -4   /* hyper-parameters outside what the kernels implement  */
#define T2_ERR_WATCHDOG    -5   /* a device-side wait timed out (kernel aborted itself)  */

typedef struct T2Model T2Model; /* opaque: configuration + packed device-side weights */

/* Hyper-parameters the model code reads (hparams.py:40-75).  The kernels are specialised for the
 * reference defaults; t2_model_create returns T2_ERR_UNSUPPORTED for anything else. */
typedef struct T2Config {
  int32_t n_mel_channels;              /* 80   */
  int32_t n_symbols;                   /* 148  */
  int32_t symbols_embedding_dim;       /* 512  */
  int32_t encoder_kernel_size;         /* 5    */
  int32_t encoder_n_convolutions;      /* 3    */
  int32_t encoder_embedding_dim;       /* 512  */
  int32_t attention_rnn_dim;           /* 1024 */
  int32_t decoder_rnn_dim;             /* 1024 */
  int32_t prenet_dim;                  /* 256  */
  int32_t attention_dim;               /* 128  */
  int32_t attention_location_n_filters;    /* 32 */
  int32_t attention_location_kernel_size;  /* 31 */
  int32_t postnet_embedding_dim;       /* 512  */
  int32_t postnet_kernel_size;         /* 5    */
  int32_t postnet_n_convolutions;      /* 5    */
  float   p_attention_dropout;         /* 0.1  */
  float   p_decoder_dropout;           /* 0.1  */
  float   bn_eps;                      /* 1e-5 */
} T2Config;

/* Number of entries of the weight table: the reference state_dict in its own order
 * (84 tensors = 60 parameters + 24 BatchNorm buffers). */
#define T2_NUM_WEIGHTS 84

/* Decoder implementations selectable at run time (both are CUDA; there is no CPU path). */
#define T2_IMPL_AUTO        0   /* persistent kernel when the shape allows, else STEPWISE */
#define T2_IMPL_STEPWISE    1   /* one fp32 kernel sequence per step (bring-up / cross-check) */
#define T2_IMPL_PERSISTENT  2   /* one persistent cooperative tcgen05 kernel for the whole loop */

#define T2_MODE_INFER    0      /* Decoder.inference: free running, prenet on the fed-back frame */
#define T2_MODE_TEACHER  1      /* Decoder.forward : teacher forced, exactly n_steps_cap steps  */

int         t2_abi_version(void);
const char* t2_last_error(void);

/* Fills {sm_count, cc_major, cc_minor, l2_bytes, max_smem_optin} of the current device. */
int t2_device_info(int32_t out[5]);

/* weights[i]: device pointer to the i-th state_dict tensor (fp32; the three int64
 * num_batches_tracked entries are ignored and may be NULL).  The model keeps the pointers (not
 * copies) of the fp32 tensors it streams directly and builds packed copies of the rest, so the
 * caller must re-create (or t2_model_refresh) the handle after the parameters change. */
int t2_model_create(T2Model** out, const T2Config* cfg, const void* const* weights,
                    int32_t n_weights, void* stream);
int t2_model_refresh(T2Model* m, const void* const* weights, int32_t n_weights, void* stream);
int t2_model_destroy(T2Model* m);

/* ---- Encoder (model.py:149-201) --------------------------------------------------------------
 * text (B, T) int64 symbol ids  ->  memory (B, T, 512).
 * lengths: NULL = Encoder.inference (every row full length); else (B) int32, sorted descending,
 * lengths[0] == T = Encoder.forward's packed-sequence semantics (zeros at padded positions).
 * training != 0: batch-statistics BatchNorm (+ running stat update into the caller's tensors)
 * and dropout(0.5) with keep masks (3, B, 512, T) uint8 or Philox(seed) when NULL. */
typedef struct T2EncoderArgs {
  const int64_t* text;                 /* (B, T) symbol ids, or NULL when `embedded` is given        */
  const float* embedded;               /* (B, T, 512) embedded inputs (model.py:503 before transpose) */
  const int32_t* lengths; int32_t B, T;
  int32_t training; const uint8_t* keep; uint64_t seed;
  float* memory;                       /* out (B, T, 512) */
  void* ws; size_t ws_bytes;
  void* stash; size_t stash_bytes;     /* optional: activations kept for t2_encoder_backward (B <= 64) */
} T2EncoderArgs;
size_t t2_encoder_workspace_bytes(const T2Model* m, int32_t B, int32_t T);
int    t2_encoder_forward(T2Model* m, const T2EncoderArgs* a, void* stream);

/* Encoder backward (autograd graph of Encoder.forward, model.py:173-190): the stash of a forward call with the same
 * text / embedded, lengths, training, keep and seed; d_memory (B, T, 512) -> gradients of the encoder parameters
 * (grads[] entries that are non-NULL are overwritten; with `text` also embedding.weight) and, if non-NULL,
 * d_embedded (B, T, 512). */
typedef struct T2EncoderBwdArgs {
  const int64_t* text; const float* embedded; const int32_t* lengths; int32_t B, T;
  int32_t training; const uint8_t* keep; uint64_t seed;
  const void* stash; size_t stash_bytes;
  const float* d_memory;
  float* d_embedded;
  float* const* grads; int32_t n_grads;
  void* ws; size_t ws_bytes;
} T2EncoderBwdArgs;
size_t t2_encoder_stash_bytes(const T2Model* m, int32_t B, int32_t T);
size_t t2_encoder_backward_workspace_bytes(const T2Model* m, int32_t B, int32_t T);
int    t2_encoder_backward(T2Model* m, const T2EncoderBwdArgs* a, void* stream);

/* ---- Decoder (model.py:204-454) --------------------------------------------------------------
 * One call runs the whole autoregressive loop.
 *   memory (B, T_enc, 512); memory_lengths (B) int32 or NULL (no masking, model.py:432).
 *   INFER  : go frame -> [prenet -> decode -> stop test] x n; per-row stop latch
 *            done[b] |= sigmoid(gate[b]) > gate_threshold  (predicate of model.py:443);
 *            the loop ends when every row has fired or after n_steps_cap (= max_decoder_steps);
 *            rows that fired keep decoding, mel_lengths[b] = first firing step + 1.
 *   TEACHER: teacher_prenet (n_steps_cap, B, 256) = prenet outputs of go frame + targets
 *            (model.py:396-399); training != 0 applies dropout(p_att / p_dec) to the recurrent
 *            hidden states (model.py:355-356, 370-371) with att_keep / dec_keep
 *            (n_steps_cap, B, 1024) uint8 or Philox when NULL.
 *   prenet_keep: (n_steps_cap, 2, B, 256) uint8 keep masks for the always-on prenet dropout
 *            (model.py:99), or NULL => in-kernel Philox4x32-10 keyed by (seed, step).
 * Outputs (row-major): mel (B, T_cap, 80), gate (B, T_cap), align (B, T_cap, T_enc) where
 * T_cap = n_steps_cap; entries at t >= *n_steps are left untouched.  mel_lengths (B) int32,
 * n_steps (1) int32 on the device. */
typedef struct T2DecoderArgs {
  int32_t mode, impl, training;
  const float* memory; const int32_t* memory_lengths; int32_t B, T_enc, n_steps_cap;
  const float* teacher_prenet;
  const uint8_t* prenet_keep; const uint8_t* att_keep; const uint8_t* dec_keep;
  uint64_t seed;
  float gate_threshold, score_mask_value;
  float* mel; float* gate; float* align; int32_t* mel_lengths; int32_t* n_steps;
  void* ws; size_t ws_bytes;
  void* stash; size_t stash_bytes;     /* TEACHER only, optional: activations kept for t2_decoder_backward
                                          (t2_decoder_stash_bytes(); opaque to the caller) */
} T2DecoderArgs;
size_t t2_decoder_workspace_bytes(const T2Model* m, int32_t B, int32_t T_enc, int32_t n_steps_cap);
int    t2_decoder_run(T2Model* m, const T2DecoderArgs* a, void* stream);

/* ---- Decoder backward (the autograd graph of Decoder.forward, model.py:381-416) ----------------
 * Reverse-time recurrence over the stash of a TEACHER run with the same memory / teacher_prenet / masks /
 * seed, then the time-batched weight gradients.  B <= 64.
 *   d_mel (B, T_mel, 80), d_gate (B, T_mel): gradients wrt the mel / gate outputs (same layout as the
 *   outputs); d_align (B, T_mel, T_enc) or NULL.
 *   d_memory (B, T_enc, 512) and d_prenet (T_mel, B, 256) (gradient wrt teacher_prenet) are written.
 *   grads: T2_NUM_WEIGHTS pointers in state_dict order; the decoder entries that are non-NULL (attention_rnn,
 *   attention_layer, decoder_rnn, linear_projection, gate_layer) are OVERWRITTEN with the gradient of the
 *   corresponding parameter.  The prenet parameters are handled by t2_prenet_backward. */
typedef struct T2DecoderBwdArgs {
  const float* memory; const int32_t* memory_lengths; int32_t B, T_enc, T_mel;
  int32_t training;                    /* same value as the forward call (hidden-state dropout on / off) */
  const float* teacher_prenet;
  const uint8_t* att_keep; const uint8_t* dec_keep; uint64_t seed;
  float score_mask_value;
  const float* align;                  /* (B, T_mel, T_enc) forward output */
  const void* stash; size_t stash_bytes;
  const float* d_mel; const float* d_gate; const float* d_align;
  float* d_memory; float* d_prenet;
  float* const* grads; int32_t n_grads;
  void* ws; size_t ws_bytes;
} T2DecoderBwdArgs;
size_t t2_decoder_stash_bytes(const T2Model* m, int32_t B, int32_t T_enc, int32_t T_mel);
size_t t2_decoder_backward_workspace_bytes(const T2Model* m, int32_t B, int32_t T_enc, int32_t T_mel);
int    t2_decoder_backward(T2Model* m, const T2DecoderBwdArgs* a, void* stream);

/* Backward of t2_prenet_forward: frames (M, 80), the same keep / seed, d_out (M, 256) ->
 * grads[prenet.layers.0 / 1] overwritten (d_frames is not needed: the frames are data, model.py:396-399). */
typedef struct T2PrenetBwdArgs {
  const float* frames; int32_t M; const uint8_t* keep; uint64_t seed;
  const float* d_out;
  float* const* grads; int32_t n_grads;
  void* ws; size_t ws_bytes;
} T2PrenetBwdArgs;
size_t t2_prenet_backward_workspace_bytes(const T2Model* m, int32_t M);
int    t2_prenet_backward(T2Model* m, const T2PrenetBwdArgs* a, void* stream);

/* Prenet over a block of frames (teacher forcing hoists it out of the loop, model.py:399):
 * frames (M, 80) -> out (M, 256); keep (2, M, 256) uint8 or NULL => Philox(seed). */
int t2_prenet_forward(T2Model* m, const float* frames, int32_t M, const uint8_t* keep,
                      uint64_t seed, float* out, void* ws, size_t ws_bytes, void* stream);

/* ---- Postnet (model.py:103-146) + residual (model.py:511 / 524) ------------------------------
 * mel (B, T, 80) time-major per row (the decoder's native storage; the reference's (B,80,T)
 * tensor is a transposed view of exactly this, model.py:336)  ->  mel_post (B, 80, T) contiguous
 * = mel^T + postnet(mel^T).  lengths (B) int32 or NULL: frames t >= lengths[b] of the INPUT are
 * treated as zero and the output there is zero (batched-inference padding, see README). */
typedef struct T2PostnetArgs {
  const float* mel;
  int64_t mel_batch_stride;            /* elements between rows b and b+1 of mel; 0 = T*80 */
  const int32_t* lengths; int32_t B, T;
  int32_t training; const uint8_t* keep; uint64_t seed;
  int32_t add_residual;                /* 1: mel_post = mel^T + postnet(mel^T) (model.py:511, 524); 0: postnet only */
  float* mel_post;
  void* ws; size_t ws_bytes;
  void* stash; size_t stash_bytes;     /* optional (lengths must be NULL): activations kept for t2_postnet_backward */
} T2PostnetArgs;
size_t t2_postnet_workspace_bytes(const T2Model* m, int32_t B, int32_t T);
int    t2_postnet_forward(T2Model* m, const T2PostnetArgs* a, void* stream);

/* Postnet backward (model.py:141-146 + the residual of :511): d_mel_post (B, 80, T) -> d_mel (B, T, 80) (gradient wrt
 * the input rows, including the residual branch when add_residual) and the postnet parameter gradients. */
typedef struct T2PostnetBwdArgs {
  int32_t B, T, training, add_residual; const uint8_t* keep; uint64_t seed;
  const int32_t* wgrad_lengths;        /* (B) or NULL: frames t >= wgrad_lengths[b] of the stashed INPUT count as zero in the
                                          first conv's weight gradient -- what the reference's autograd computes, because
                                          parse_output zeroes that tensor in place after the forward pass (model.py:492) */
  const void* stash; size_t stash_bytes;
  const float* d_mel_post;
  float* d_mel;
  float* const* grads; int32_t n_grads;
  void* ws; size_t ws_bytes;
} T2PostnetBwdArgs;
size_t t2_postnet_stash_bytes(const T2Model* m, int32_t B, int32_t T);
size_t t2_postnet_backward_workspace_bytes(const T2Model* m, int32_t B, int32_t T);
int    t2_postnet_backward(T2Model* m, const T2PostnetBwdArgs* a, void* stream);

/* ---- Fused gradient clipping + Adam (train.py:229-236: clip_grad_norm_ then torch.optim.Adam.step) -----------
 * n tensors (host arrays of device pointers + element counts).  grad_norm (device, 1 float) receives the total
 * gradient norm BEFORE clipping; the gradients are scaled in place by min(1, max_norm / (norm + 1e-6)) like
 * torch.nn.utils.clip_grad_norm_ (max_norm <= 0: no clipping), then exp_avg / exp_avg_sq / the parameters are updated
 * exactly like torch.optim.Adam (L2 weight decay, bias correction with `step` counted from 1). */
typedef struct T2AdamArgs {
  int32_t n;
  float* const* params; float* const* grads; float* const* exp_avg; float* const* exp_avg_sq; const int64_t* numel;
  double lr, beta1, beta2, eps, weight_decay, max_norm;   /* doubles: 1 - beta2^step must not be rounded through fp32 */
  int32_t step;
  float* grad_norm;
  void* ws; size_t ws_bytes;
} T2AdamArgs;
size_t t2_clip_adam_workspace_bytes(int64_t total_elements, int32_t n_tensors);
int    t2_clip_adam_step(const T2AdamArgs* a, void* stream);

/* ---- mixed-precision optimizer step (the reference trains "fp16" through Apex AMP O2, train.py:173-176, 222-236) ------
 * fp16 (or fp32) model parameters + fp32 master copies; gradients arrive in the parameter's dtype multiplied by the
 * dynamic loss scale state[0].  One call = unscale, overflow check, clip_grad_norm_ on the unscaled gradients, Adam on
 * the masters, write-back of the model copies, loss-scaler update (apex LossScaler: overflow -> skip the step, scale *
 * backoff_factor; growth_interval consecutive good steps -> scale * growth_factor).  Three multi-tensor launches, no host
 * synchronisation: `state` (4 floats on the device: loss scale, good steps since the last scale change, optimizer steps
 * taken -- skipped steps do not count --, 1.0 if this step was skipped) carries everything between calls. */
typedef struct T2AmpAdamArgs {
  int32_t n;
  void* const* model_params; const int32_t* param_is_half;   /* per tensor: storage read by the model, 1 = __half */
  const void* const* grads;  const int32_t* grad_is_half;    /* per tensor: gradient x loss scale */
  float* const* master; float* const* exp_avg; float* const* exp_avg_sq; const int64_t* numel;
  double lr, beta1, beta2, eps, weight_decay, max_norm;
  int32_t growth_interval; float growth_factor, backoff_factor;
  float* state;            /* device, 4 floats (see above) */
  float* grad_norm;        /* device out: norm of the unscaled gradients before clipping (inf / nan on overflow) */
  int32_t* skipped;        /* device out (may be NULL): 1 = overflow, nothing was updated */
  void* ws; size_t ws_bytes;
} T2AmpAdamArgs;
size_t t2_amp_adam_workspace_bytes(int64_t total_elements, int32_t n_tensors);
int    t2_amp_adam_step(const T2AmpAdamArgs* a, void* stream);

/* ---- Tacotron2Loss fused with the parse_output mask and the gradient seeds (loss_function.py:8-19, model.py:487-497) ----
 * mel / mel_post (B, C, T) contiguous fp32, gate (B, T), targets of the same shapes.  output_lengths (B) int32 or NULL: when
 * given, frames t >= output_lengths[b] of mel / mel_post are zeroed and of gate set to 1e3 IN PLACE before they enter the
 * loss (parse_output).  loss[0] = MSE(mel) + MSE(mel_post) + BCEWithLogits(gate), loss[1..3] the three terms.  d_mel /
 * d_mel_post / d_gate (each may be NULL): d loss / d output, written in the same pass. */
typedef struct T2LossArgs {
  float* mel; float* mel_post; float* gate;
  const float* mel_target; const float* gate_target;
  const int32_t* output_lengths;
  int32_t B, C, T;
  float* loss;                                   /* device, 4 floats */
  float* d_mel; float* d_mel_post; float* d_gate;
  void* ws; size_t ws_bytes;
} T2LossArgs;
size_t t2_loss_workspace_bytes(void);
int    t2_tacotron2_loss(const T2LossArgs* a, void* stream);

/* ---- TacotronSTFT.mel_spectrogram (layers.py:63-80, stft.py:69-94): batch of waveforms -> log-mel spectrograms ----------
 * y (B, n_samples) fp32 in [-1, 1] (not checked here; the Python mirror asserts it like the reference does).
 * forward_basis (2 * (filter_length / 2 + 1), filter_length): the windowed Fourier basis of stft.py:44-63;
 * mel_basis (n_mel, filter_length / 2 + 1).  mel out: (B, n_mel, n_frames) = log(max(mel_basis . |STFT|, clip_val)),
 * n_frames = n_samples / hop_length + 1 (reflect padding by filter_length / 2 on both sides). */
typedef struct T2MelSpecArgs {
  const float* y; int32_t B, n_samples;
  int32_t filter_length, hop_length, n_mel;
  const float* forward_basis; const float* mel_basis;
  float clip_val;
  float* mel;
  void* ws; size_t ws_bytes;
} T2MelSpecArgs;
int32_t t2_mel_spectrogram_frames(int32_t n_samples, int32_t hop_length);
size_t  t2_mel_spectrogram_workspace_bytes(int32_t B, int32_t n_samples, int32_t filter_length, int32_t hop_length, int32_t n_mel);
int     t2_mel_spectrogram(const T2MelSpecArgs* a, void* stream);

/* ---- TextMelCollate on the device (data_utils.py:73-111): ragged batch in HBM -> the padded, length-sorted 5-tuple ------------
 * text_flat: the B token sequences concatenated (int64), text_offsets (B + 1) their prefix offsets; mel_flat: the B
 * (n_mel, L_i) row-major spectrograms concatenated, mel_offsets (B + 1) prefix offsets in FRAMES.  T_max >= longest text,
 * L_pad >= longest mel (the caller rounds it up to a multiple of n_frames_per_step, data_utils.py:93-96).  Rows are ordered by
 * decreasing text length (ties keep the original order); order[b] = source row of output row b.  All pointers device. */
typedef struct T2CollateArgs {
  const int64_t* text_flat; const int64_t* text_offsets;
  const float* mel_flat; const int64_t* mel_offsets;
  int32_t B, n_mel, T_max, L_pad;
  int32_t* order;
  int64_t* text_padded; int64_t* input_lengths;
  float* mel_padded; float* gate_padded; int64_t* output_lengths;
} T2CollateArgs;
int t2_collate(const T2CollateArgs* a, void* stream);

/* ---- Tacotron2.inference end to end with HOST buffers (model.py:517-529) ----------------------
 * text_host (B, T_text) int64 in (pinned) host memory -> mel_post_host (B, 80, T_cap) fp32,
 * mel_lengths_host (B), n_steps_host (1).  Copies H2D, runs encoder -> decoder -> postnet on
 * `stream`, copies D2H and synchronises the stream.  ws is device memory of
 * t2_infer_workspace_bytes(). */
size_t t2_infer_workspace_bytes(const T2Model* m, int32_t B, int32_t T_text, int32_t max_steps);
int    t2_infer_host(T2Model* m, const int64_t* text_host, int32_t B, int32_t T_text,
                     int32_t max_steps, float gate_threshold, uint64_t seed, int32_t impl,
                     float* mel_post_host, int32_t* mel_lengths_host, int32_t* n_steps_host,
                     void* ws, size_t ws_bytes, void* stream);

/* ---- self tests (libt2b200_selftest.so only: the same sources built with -DT2_SELFTEST; not part of the product
 * library) -------------------------------------------------------------------------------------------------
 * t2_selftest_umma: runs the tcgen05 split-fp16 GEMM engine used by the persistent decoder on a
 * (64 x K) x (N x K)^T problem and writes C (64 x N) fp32. */
#ifdef T2_SELFTEST
int t2_selftest_umma(const float* A, const float* W, int32_t N, int32_t K, int32_t passes,
                     float* C, void* stream);
/* Micro-benchmark: SM cycles for `reps` back-to-back tcgen05.mma (M x N x 16, fp16, operands resident in
 * shared memory) -> out_host[0] = issue cycles, out_host[1] = issue + completion cycles. */
int t2_selftest_mma_rate(int32_t M, int32_t N, int32_t reps, int32_t alternate_d, int64_t* out_host);
/* `reps` groups of `group` back-to-back MMAs, each group followed by tcgen05.commit + a wait for it (one K chunk of a
 * streaming event): out_host[0] = total SM cycles. */
int t2_selftest_mma_group(int32_t M, int32_t N, int32_t group, int32_t reps, int64_t* out_host);
/* The training path's general tensor-core GEMM (gemm_tc.cu): row-major C = op(A) . op(B) + beta C, strided batch. */
int t2_selftest_gemm_tc(int32_t ta, int32_t tb, int32_t M, int32_t N, int32_t K, const float* A, int64_t lda,
                        const float* B, int64_t ldb, float* C, int64_t ldc, float beta, int32_t batch,
                        int64_t strideA, int64_t strideB, int64_t strideC, void* stream);
int t2_selftest_colsum(const float* X, int64_t ld, int64_t rows, int32_t cols, float* out, void* stream);
#endif
/* ---- instrumentation ---------------------------------------------------------------------------------- */
/* After a T2_IMPL_PERSISTENT run with the same args / workspace: SM cycles spent per phase of the
 * persistent kernel, summed over steps, on three sample CTAs (out_host[3][24]; phase list in
 * decoder_persistent.cu).  Synchronises the device. */
int t2_decoder_profile(const T2DecoderArgs* a, int64_t* out_host);
/* number of kernels this library has launched since load (for bench.py's gpu_launches) */
int64_t t2_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* T2B200_H_ */
