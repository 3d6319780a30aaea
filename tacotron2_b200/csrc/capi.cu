// C ABI of libt2b200.so (include/t2b200.h): argument checking, weight packing, dispatch.
#include <stdarg.h>
#include <string.h>

#include <vector>

#include "conv_tc.h"
#include "decoder.h"
#include "gemm_tc.h"
#include "gemm_f32.cuh"
#include "train_layers.h"

namespace t2 {

std::string& last_error() {
  static thread_local std::string e;
  return e;
}
long long g_launch_count = 0;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

int encoder_forward(T2Model* m, const T2EncoderArgs* a, cudaStream_t s);
size_t encoder_ws_bytes(int B, int T);
int postnet_forward(T2Model* m, const T2PostnetArgs* a, cudaStream_t s);
size_t postnet_ws_bytes(int B, int T);
int selftest_umma(const float* A, const float* W, int N, int K, int passes, float* C, cudaStream_t s);
int mma_rate(int M, int N, int reps, int alternate_d, long long* out_host, cudaStream_t s);
int mma_group(int M, int N, int group, int reps, long long* out_host, cudaStream_t s);

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// ---- weight packing -----------------------------------------------------------------------------
// conv weight (co, ci, k) -> (co, k, ci) so a conv is a GEMM over K = taps x Cin on channels-last rows
__global__ void permute_conv_kernel(const float* __restrict__ w, float* __restrict__ out, int co, int ci, int k) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)co * ci * k) return;
  const int kk = (int)(i % k); const long r = i / k; const int c = (int)(r % ci); const int o = (int)(r / ci);
  out[((long)o * k + kk) * ci + c] = w[i];
}
__global__ void add2_kernel(const float* a, const float* b, float* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

template <typename T>
static int dmalloc(T** p, size_t n) {
  if (*p) return T2_OK;
  T2_CUDA(cudaMalloc((void**)p, n * sizeof(T)));
  return T2_OK;
}

int pack_model(T2Model* m, cudaStream_t s) {
  for (int i = 0; i < 3; ++i) {
    T2_TRY(dmalloc(&m->enc_conv_w[i], (size_t)kEnc * kEnc * kConvK));
    const long n = (long)kEnc * kEnc * kConvK;
    permute_conv_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(m->w[W_ENC_CONV0 + 7 * i], m->enc_conv_w[i], kEnc, kEnc, kConvK);
    T2_LAUNCH_CHECK();
  }
  for (int i = 0; i < 5; ++i) {
    const int ci = i == 0 ? kMel : kPost, co = i == 4 ? kMel : kPost;
    T2_TRY(dmalloc(&m->post_conv_w[i], (size_t)co * ci * kConvK));
    const long n = (long)co * ci * kConvK;
    permute_conv_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(m->w[W_POST_CONV0 + 7 * i], m->post_conv_w[i], co, ci, kConvK);
    T2_LAUNCH_CHECK();
  }
  T2_TRY(dmalloc(&m->enc_lstm_wih, (size_t)8 * kEncH * kEnc));
  T2_TRY(dmalloc(&m->enc_lstm_b, (size_t)8 * kEncH));
  for (int d = 0; d < 2; ++d) {
    T2_CUDA(cudaMemcpyAsync(m->enc_lstm_wih + (size_t)d * 4 * kEncH * kEnc, m->w[W_ENC_LSTM + 4 * d],
                            (size_t)4 * kEncH * kEnc * 4, cudaMemcpyDeviceToDevice, s));
    add2_kernel<<<(4 * kEncH + 255) / 256, 256, 0, s>>>(m->w[W_ENC_LSTM + 4 * d + 2], m->w[W_ENC_LSTM + 4 * d + 3],
                                                        m->enc_lstm_b + d * 4 * kEncH, 4 * kEncH);
    T2_LAUNCH_CHECK();
  }
  T2_TRY(dmalloc(&m->arnn_b, (size_t)4 * kARnn));
  T2_TRY(dmalloc(&m->drnn_b, (size_t)4 * kDRnn));
  add2_kernel<<<(4 * kARnn + 255) / 256, 256, 0, s>>>(m->w[W_ARNN_BIH], m->w[W_ARNN_BHH], m->arnn_b, 4 * kARnn);
  T2_LAUNCH_CHECK();
  add2_kernel<<<(4 * kDRnn + 255) / 256, 256, 0, s>>>(m->w[W_DRNN_BIH], m->w[W_DRNN_BHH], m->drnn_b, 4 * kDRnn);
  T2_LAUNCH_CHECK();
  const int kdc = kDRnn + kEnc;
  T2_TRY(dmalloc(&m->projgate_w, (size_t)(kMel + 1) * kdc));
  T2_TRY(dmalloc(&m->projgate_b, (size_t)(kMel + 1)));
  T2_CUDA(cudaMemcpyAsync(m->projgate_w, m->w[W_PROJ_W], (size_t)kMel * kdc * 4, cudaMemcpyDeviceToDevice, s));
  T2_CUDA(cudaMemcpyAsync(m->projgate_w + (size_t)kMel * kdc, m->w[W_GATE_W], (size_t)kdc * 4, cudaMemcpyDeviceToDevice, s));
  T2_CUDA(cudaMemcpyAsync(m->projgate_b, m->w[W_PROJ_B], kMel * 4, cudaMemcpyDeviceToDevice, s));
  T2_CUDA(cudaMemcpyAsync(m->projgate_b + kMel, m->w[W_GATE_B], 4, cudaMemcpyDeviceToDevice, s));
  if (!m->ones) {
    T2_TRY(dmalloc(&m->ones, (size_t)8192));
    std::vector<float> h(8192, 1.f);
    T2_CUDA(cudaMemcpy(m->ones, h.data(), 8192 * 4, cudaMemcpyHostToDevice));
  }
  if (!m->zeros) {
    T2_TRY(dmalloc(&m->zeros, (size_t)8192));
    T2_CUDA(cudaMemsetAsync(m->zeros, 0, 8192 * 4, s));
  }
  // tensor-core conv / GEMM weight images
  for (int i = 0; i < 3; ++i) T2_TRY(tc_pack_weights(m->w[W_ENC_CONV0 + 7 * i], kEnc, kEnc, kConvK, 128, &m->tc_enc_conv[i], s));
  T2_TRY(tc_pack_weights(m->enc_lstm_wih, 8 * kEncH, kEnc, 1, 128, &m->tc_enc_wih, s));
  for (int i = 0; i < 5; ++i) {
    const int ci = i == 0 ? kMel : kPost, co = i == 4 ? kMel : kPost;
    T2_TRY(tc_pack_weights(m->w[W_POST_CONV0 + 7 * i], co, ci, kConvK, i == 4 ? 80 : 128, &m->tc_post_conv[i], s));
  }
  T2_TRY(persistent_pack_create(m, s));
  return T2_OK;
}

// ---- decoder workspace --------------------------------------------------------------------------
size_t decoder_ws_bytes(int B, int T, int cap) {
  size_t n = 0;
  n += align256((size_t)B * T * kAtt * 4);                                   // pm
  n += align256(((size_t)B * (2 * kARnn + 2 * kDRnn + kEnc) + 2 * (size_t)B * T) * 4);  // state
  n += 2 * align256((size_t)B * kPre * 4);                                   // x1 x2
  n += align256((size_t)B * 4 * kARnn * 4);                                  // gates
  n += align256((size_t)B * (kMel + 1) * 4);                                 // proj
  n += align256(sizeof(DecoderCtrl));
  n += align256(persistent_ws_bytes(B, T, cap));
  return n + 256;
}

int decoder_ws_carve(const T2DecoderArgs* a, DecoderWs* w) {
  const int B = a->B, T = a->T_enc;
  if (a->ws_bytes < decoder_ws_bytes(B, T, a->n_steps_cap)) return fail(T2_ERR_WORKSPACE, "decoder workspace too small");
  char* p = (char*)(((uintptr_t)a->ws + 255) & ~(uintptr_t)255);
  w->pm = (float*)p; p += align256((size_t)B * T * kAtt * 4);
  w->state_begin = p;
  w->state_bytes = ((size_t)B * (2 * kARnn + 2 * kDRnn + kEnc) + 2 * (size_t)B * T) * 4;
  float* f = (float*)p;
  w->ah = f; f += (size_t)B * kARnn;
  w->ac = f; f += (size_t)B * kARnn;
  w->dh = f; f += (size_t)B * kDRnn;
  w->dc = f; f += (size_t)B * kDRnn;
  w->ctx = f; f += (size_t)B * kEnc;
  w->aw = f; f += (size_t)B * T;
  w->awc = f; f += (size_t)B * T;
  p += align256(w->state_bytes);
  w->x1 = (float*)p; p += align256((size_t)B * kPre * 4);
  w->x2 = (float*)p; p += align256((size_t)B * kPre * 4);
  w->gates = (float*)p; p += align256((size_t)B * 4 * kARnn * 4);
  w->proj = (float*)p; p += align256((size_t)B * (kMel + 1) * 4);
  w->ctrl = (DecoderCtrl*)p; p += align256(sizeof(DecoderCtrl));
  w->persistent = p; w->persistent_bytes = persistent_ws_bytes(B, T, a->n_steps_cap);
  return T2_OK;
}

static int check_decoder_args(const T2Model* m, const T2DecoderArgs* a) {
  if (!m || !a) return fail(T2_ERR_INVALID, "null model / args");
  if (a->B <= 0 || a->B > kMaxBatch) return fail(T2_ERR_INVALID, "decoder: B=%d outside [1, %d]", a->B, kMaxBatch);
  if (a->T_enc <= 0) return fail(T2_ERR_INVALID, "decoder: empty encoder memory (T_enc=%d)", a->T_enc);
  if (a->n_steps_cap <= 0) return fail(T2_ERR_INVALID, "decoder: n_steps_cap=%d", a->n_steps_cap);
  if (!a->memory || !a->mel || !a->gate || !a->align || !a->mel_lengths || !a->n_steps || !a->ws)
    return fail(T2_ERR_INVALID, "decoder: null tensor pointer");
  if (a->mode == T2_MODE_TEACHER && !a->teacher_prenet) return fail(T2_ERR_INVALID, "decoder: teacher mode needs teacher_prenet");
  if (a->mode != T2_MODE_TEACHER && a->mode != T2_MODE_INFER) return fail(T2_ERR_INVALID, "decoder: bad mode %d", a->mode);
  return T2_OK;
}

}  // namespace t2

using namespace t2;

extern "C" {

int t2_abi_version(void) { return T2_ABI_VERSION; }
const char* t2_last_error(void) { return last_error().c_str(); }
int64_t t2_kernel_launch_count(void) { return g_launch_count; }

int t2_device_info(int32_t out[5]) {
  int dev = 0;
  T2_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  T2_CUDA(cudaGetDeviceProperties(&p, dev));
  out[0] = p.multiProcessorCount; out[1] = p.major; out[2] = p.minor;
  out[3] = (int32_t)(p.l2CacheSize); out[4] = (int32_t)p.sharedMemPerBlockOptin;
  return T2_OK;
}

static int check_cfg(const T2Config* c) {
  const bool ok = c->n_mel_channels == kMel && c->symbols_embedding_dim == kEnc && c->encoder_kernel_size == kConvK &&
                  c->encoder_n_convolutions == 3 && c->encoder_embedding_dim == kEnc && c->attention_rnn_dim == kARnn &&
                  c->decoder_rnn_dim == kDRnn && c->prenet_dim == kPre && c->attention_dim == kAtt &&
                  c->attention_location_n_filters == kLocF && c->attention_location_kernel_size == kLocK &&
                  c->postnet_embedding_dim == kPost && c->postnet_kernel_size == kConvK && c->postnet_n_convolutions == 5 &&
                  c->n_symbols > 0;
  if (!ok) return fail(T2_ERR_UNSUPPORTED, "hyper-parameters differ from the reference defaults the sm_100a kernels are built for");
  return T2_OK;
}

static int set_weights(T2Model* m, const void* const* weights, int32_t n) {
  if (n != T2_NUM_WEIGHTS) return fail(T2_ERR_INVALID, "expected %d weight pointers, got %d", T2_NUM_WEIGHTS, n);
  for (int i = 0; i < n; ++i) {
    const bool nbt = (i >= W_ENC_CONV0 && i < W_ENC_LSTM && (i - W_ENC_CONV0) % 7 == 6) ||
                     (i >= W_POST_CONV0 && (i - W_POST_CONV0) % 7 == 6);
    if (!nbt && weights[i] == nullptr) return fail(T2_ERR_INVALID, "weight %d is null", i);
    m->w[i] = (const float*)weights[i];
  }
  return T2_OK;
}

int t2_model_create(T2Model** out, const T2Config* cfg, const void* const* weights, int32_t n_weights, void* stream) {
  if (!out || !cfg || !weights) return fail(T2_ERR_INVALID, "null argument");
  T2_TRY(check_cfg(cfg));
  int dev = 0;
  T2_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  T2_CUDA(cudaGetDeviceProperties(&p, dev));
  if (p.major != 10) return fail(T2_ERR_UNSUPPORTED, "libt2b200 is built for sm_100a only (device is sm_%d%d)", p.major, p.minor);
  T2Model* m = new T2Model();
  m->cfg = *cfg; m->device = dev; m->sm_count = p.multiProcessorCount;
  int r = set_weights(m, weights, n_weights);
  if (r == T2_OK) r = pack_model(m, (cudaStream_t)stream);
  if (r != T2_OK) { t2_model_destroy(m); return r; }
  *out = m;
  return T2_OK;
}

int t2_model_refresh(T2Model* m, const void* const* weights, int32_t n_weights, void* stream) {
  if (!m) return fail(T2_ERR_INVALID, "null model");
  T2_TRY(set_weights(m, weights, n_weights));
  return pack_model(m, (cudaStream_t)stream);
}

int t2_model_destroy(T2Model* m) {
  if (!m) return T2_OK;
  for (int i = 0; i < 3; ++i) cudaFree(m->enc_conv_w[i]);
  for (int i = 0; i < 5; ++i) cudaFree(m->post_conv_w[i]);
  cudaFree(m->enc_lstm_wih); cudaFree(m->enc_lstm_b); cudaFree(m->arnn_b); cudaFree(m->drnn_b);
  cudaFree(m->projgate_w); cudaFree(m->projgate_b); cudaFree(m->zeros); cudaFree(m->ones); cudaFree(m->dgrad_tmp);
  for (int i = 0; i < 3; ++i) cudaFree(m->tc_dgrad_enc[i]);
  for (int i = 0; i < 5; ++i) cudaFree(m->tc_dgrad_post[i]);
  for (int i = 0; i < 3; ++i) cudaFree(m->tc_enc_conv[i]);
  for (int i = 0; i < 5; ++i) cudaFree(m->tc_post_conv[i]);
  cudaFree(m->tc_enc_wih);
  persistent_pack_destroy(m);
  gemm_tc_destroy(m);
  blas_destroy(m);
  delete m;
  return T2_OK;
}

size_t t2_encoder_workspace_bytes(const T2Model*, int32_t B, int32_t T) { return encoder_ws_bytes(B, T); }
int t2_encoder_forward(T2Model* m, const T2EncoderArgs* a, void* stream) {
  if (!m || !a || (!a->text && !a->embedded) || !a->memory || !a->ws) return fail(T2_ERR_INVALID, "encoder: null argument");
  return encoder_forward(m, a, (cudaStream_t)stream);
}

size_t t2_encoder_stash_bytes(const T2Model*, int32_t B, int32_t T) { return encoder_stash_bytes(B, T); }
size_t t2_encoder_backward_workspace_bytes(const T2Model*, int32_t B, int32_t T) { return encoder_backward_ws_bytes(B, T); }
int t2_encoder_backward(T2Model* m, const T2EncoderBwdArgs* a, void* stream) {
  if (!m || !a || (!a->text && !a->embedded) || !a->stash || !a->d_memory || !a->grads || !a->ws)
    return fail(T2_ERR_INVALID, "encoder backward: null argument");
  return encoder_backward(m, a, (cudaStream_t)stream);
}
size_t t2_postnet_stash_bytes(const T2Model*, int32_t B, int32_t T) { return postnet_stash_bytes(B, T); }
size_t t2_postnet_backward_workspace_bytes(const T2Model*, int32_t B, int32_t T) { return postnet_backward_ws_bytes(B, T); }
int t2_postnet_backward(T2Model* m, const T2PostnetBwdArgs* a, void* stream) {
  if (!m || !a || !a->stash || !a->d_mel_post || !a->grads || !a->ws) return fail(T2_ERR_INVALID, "postnet backward: null argument");
  return postnet_backward(m, a, (cudaStream_t)stream);
}

size_t t2_decoder_workspace_bytes(const T2Model*, int32_t B, int32_t T_enc, int32_t cap) { return decoder_ws_bytes(B, T_enc, cap); }
int t2_decoder_run(T2Model* m, const T2DecoderArgs* a, void* stream) {
  T2_TRY(check_decoder_args(m, a));
  int impl = a->impl;
  if (impl == T2_IMPL_AUTO) impl = persistent_supported(m, a) ? T2_IMPL_PERSISTENT : T2_IMPL_STEPWISE;
  if (impl == T2_IMPL_PERSISTENT) {
    if (!persistent_supported(m, a)) return fail(T2_ERR_UNSUPPORTED, "persistent decoder does not support this shape/mode");
    return decoder_run_persistent(m, a, (cudaStream_t)stream);
  }
  // only the persistent kernel writes the training stash: a backward pass over a stash the stepwise path left
  // untouched would silently produce garbage gradients
  if (a->stash)
    return fail(T2_ERR_UNSUPPORTED, "decoder: the training stash needs the persistent implementation (%s)",
                a->impl == T2_IMPL_STEPWISE ? "impl = STEPWISE was requested"
                                            : "this shape / device does not fit it: T_enc too large for shared memory or < 128 SMs");
  return decoder_run_stepwise(m, a, (cudaStream_t)stream);
}

size_t t2_decoder_stash_bytes(const T2Model*, int32_t B, int32_t, int32_t T_mel) { return decoder_stash_bytes(B, T_mel); }
size_t t2_decoder_backward_workspace_bytes(const T2Model*, int32_t B, int32_t T_enc, int32_t T_mel) {
  return decoder_backward_ws_bytes(B, T_enc, T_mel);
}
int t2_decoder_backward(T2Model* m, const T2DecoderBwdArgs* a, void* stream) {
  if (!m || !a || !a->memory || !a->teacher_prenet || !a->align || !a->stash || !a->d_mel || !a->d_gate || !a->d_prenet ||
      !a->grads || !a->ws)
    return fail(T2_ERR_INVALID, "decoder backward: null argument");
  return decoder_backward(m, a, (cudaStream_t)stream);
}
size_t t2_prenet_backward_workspace_bytes(const T2Model*, int32_t M) { return (size_t)4 * M * kPre * 4 + 1024; }
int t2_prenet_backward(T2Model* m, const T2PrenetBwdArgs* a, void* stream) {
  if (!m || !a || !a->frames || !a->d_out || !a->grads || !a->ws || a->M <= 0) return fail(T2_ERR_INVALID, "prenet backward: bad argument");
  return prenet_backward(m, a, (cudaStream_t)stream);
}

int t2_prenet_forward(T2Model* m, const float* frames, int32_t M, const uint8_t* keep, uint64_t seed, float* out,
                      void* ws, size_t ws_bytes, void* stream) {
  if (!m || !frames || !out || M <= 0) return fail(T2_ERR_INVALID, "prenet: bad argument");
  if (ws_bytes < (size_t)M * kPre * 4) return fail(T2_ERR_WORKSPACE, "prenet workspace too small");
  float* x1 = (float*)ws;
  GemmArgs g;                                                     // model.py:97-100
  g.seg[0] = {frames, kMel, m->w[W_PRENET0], kMel, kMel};
  g.M = M; g.N = kPre; g.C = x1; g.ldc = kPre; g.act = ACT_RELU; g.p_drop = 0.5f;
  if (keep) { g.keep = keep; g.ldkeep = kPre; } else { g.philox = 1; g.seed = seed; g.site = 0xA0; }
  T2_TRY(gemm_f32(g, (cudaStream_t)stream));
  GemmArgs h;
  h.seg[0] = {x1, kPre, m->w[W_PRENET1], kPre, kPre};
  h.M = M; h.N = kPre; h.C = out; h.ldc = kPre; h.act = ACT_RELU; h.p_drop = 0.5f;
  if (keep) { h.keep = keep + (size_t)M * kPre; h.ldkeep = kPre; } else { h.philox = 1; h.seed = seed; h.site = 0xA1; }
  return gemm_f32(h, (cudaStream_t)stream);
}

size_t t2_postnet_workspace_bytes(const T2Model*, int32_t B, int32_t T) { return postnet_ws_bytes(B, T); }
int t2_postnet_forward(T2Model* m, const T2PostnetArgs* a, void* stream) {
  if (!m || !a || !a->mel || !a->mel_post || !a->ws) return fail(T2_ERR_INVALID, "postnet: null argument");
  if (a->stash) return postnet_forward_train(m, a, (cudaStream_t)stream);
  return postnet_forward(m, a, (cudaStream_t)stream);
}

// ---- end to end with host buffers ------------------------------------------------------------------
static void infer_carve(int B, int Tt, int S, char* base, int64_t** text, float** memory, float** mel, float** gate,
                        float** align, float** post, int32_t** lens, int32_t** nsteps, char** sub, size_t* sub_bytes,
                        size_t* total) {
  char* p = base;
  auto take = [&](size_t n) { char* r = p; p += align256(n); return r; };
  *text = (int64_t*)take((size_t)B * Tt * 8);
  *memory = (float*)take((size_t)B * Tt * kEnc * 4);
  *mel = (float*)take((size_t)B * S * kMel * 4);
  *gate = (float*)take((size_t)B * S * 4);
  *align = (float*)take((size_t)B * S * Tt * 4);
  *post = (float*)take((size_t)B * S * kMel * 4);
  *lens = (int32_t*)take((size_t)B * 4);
  *nsteps = (int32_t*)take(256);
  size_t sb = encoder_ws_bytes(B, Tt);
  if (decoder_ws_bytes(B, Tt, S) > sb) sb = decoder_ws_bytes(B, Tt, S);
  if (postnet_ws_bytes(B, S) > sb) sb = postnet_ws_bytes(B, S);
  *sub = take(sb); *sub_bytes = sb;
  *total = (size_t)(p - base) + 256;
}

size_t t2_infer_workspace_bytes(const T2Model*, int32_t B, int32_t T_text, int32_t max_steps) {
  int64_t* a; float *b, *c, *d, *e, *f; int32_t *g, *h; char* s; size_t sb, total;
  infer_carve(B, T_text, max_steps, nullptr, &a, &b, &c, &d, &e, &f, &g, &h, &s, &sb, &total);
  return total;
}

int t2_infer_host(T2Model* m, const int64_t* text_host, int32_t B, int32_t T_text, int32_t max_steps,
                  float gate_threshold, uint64_t seed, int32_t impl, float* mel_post_host,
                  int32_t* mel_lengths_host, int32_t* n_steps_host, void* ws, size_t ws_bytes, void* stream) {
  if (!m || !text_host || !mel_post_host || !mel_lengths_host || !n_steps_host || !ws)
    return fail(T2_ERR_INVALID, "infer_host: null argument");
  if (ws_bytes < t2_infer_workspace_bytes(m, B, T_text, max_steps)) return fail(T2_ERR_WORKSPACE, "infer workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  int64_t* text; float *memory, *mel, *gate, *align, *post; int32_t *lens, *nsteps; char* sub; size_t sb, total;
  char* base = (char*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  infer_carve(B, T_text, max_steps, base, &text, &memory, &mel, &gate, &align, &post, &lens, &nsteps, &sub, &sb, &total);
  T2_CUDA(cudaMemcpyAsync(text, text_host, (size_t)B * T_text * 8, cudaMemcpyHostToDevice, s));
  T2EncoderArgs ea; memset(&ea, 0, sizeof(ea));
  ea.text = text; ea.B = B; ea.T = T_text; ea.memory = memory; ea.ws = sub; ea.ws_bytes = sb;
  T2_TRY(encoder_forward(m, &ea, s));
  T2DecoderArgs da; memset(&da, 0, sizeof(da));
  da.mode = T2_MODE_INFER; da.impl = impl; da.memory = memory; da.B = B; da.T_enc = T_text; da.n_steps_cap = max_steps;
  da.seed = seed; da.gate_threshold = gate_threshold; da.score_mask_value = -INFINITY;
  da.mel = mel; da.gate = gate; da.align = align; da.mel_lengths = lens; da.n_steps = nsteps; da.ws = sub; da.ws_bytes = sb;
  T2_TRY(t2_decoder_run(m, &da, s));
  // the postnet runs over the full cap; frames beyond each row's length are zeroed (lengths mask)
  T2PostnetArgs pa; memset(&pa, 0, sizeof(pa));
  pa.mel = mel; pa.lengths = lens; pa.add_residual = 1; pa.B = B; pa.T = max_steps; pa.mel_post = post; pa.ws = sub; pa.ws_bytes = sb;
  T2_TRY(postnet_forward(m, &pa, s));
  T2_CUDA(cudaMemcpyAsync(mel_post_host, post, (size_t)B * max_steps * kMel * 4, cudaMemcpyDeviceToHost, s));
  T2_CUDA(cudaMemcpyAsync(mel_lengths_host, lens, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
  T2_CUDA(cudaMemcpyAsync(n_steps_host, nsteps, 4, cudaMemcpyDeviceToHost, s));
  T2_CUDA(cudaStreamSynchronize(s));
  return T2_OK;
}

int t2_decoder_profile(const T2DecoderArgs* a, int64_t* out_host) {
  if (!a || !out_host) return fail(T2_ERR_INVALID, "decoder_profile: null argument");
  DecoderWs w;
  T2_TRY(decoder_ws_carve(a, &w));
  T2_CUDA(cudaDeviceSynchronize());
  T2_CUDA(cudaMemcpy(out_host, w.ctrl->prof, sizeof(long long) * 72, cudaMemcpyDeviceToHost));
  return T2_OK;
}

#ifdef T2_SELFTEST   // libt2b200_selftest.so only
int t2_selftest_mma_rate(int32_t M, int32_t N, int32_t reps, int32_t alternate_d, int64_t* out_host) {
  return mma_rate(M, N, reps, alternate_d, (long long*)out_host, 0);
}

int t2_selftest_mma_group(int32_t M, int32_t N, int32_t group, int32_t reps, int64_t* out_host) {
  return mma_group(M, N, group, reps, (long long*)out_host, 0);
}

int t2_selftest_umma(const float* A, const float* W, int32_t N, int32_t K, int32_t passes, float* C, void* stream) {
  return selftest_umma(A, W, N, K, passes, C, (cudaStream_t)stream);
}

// C = op(A) . op(B) + beta C through the training path's tensor-core GEMM (gemm_tc.cu); batch > 1: strided batch
int t2_selftest_gemm_tc(int32_t ta, int32_t tb, int32_t M, int32_t N, int32_t K, const float* A, int64_t lda, const float* B,
                        int64_t ldb, float* C, int64_t ldc, float beta, int32_t batch, int64_t strideA, int64_t strideB,
                        int64_t strideC, void* stream) {
  static T2Model scratch_owner;            // only its GEMM scratch is used
  GemmTc g;
  g.ta = ta != 0; g.tb = tb != 0; g.M = M; g.N = N; g.K = K; g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
  g.beta = beta; g.batch = batch; g.strideA = strideA; g.strideB = strideB; g.strideC = strideC;
  return gemm_tc(&scratch_owner, (cudaStream_t)stream, g);
}
int t2_selftest_colsum(const float* X, int64_t ld, int64_t rows, int32_t cols, float* out, void* stream) {
  static T2Model scratch_owner;
  return colsum_f32(&scratch_owner, (cudaStream_t)stream, X, ld, rows, cols, out);
}
#endif

}  // extern "C"
