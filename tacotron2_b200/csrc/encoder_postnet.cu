// Encoder (model.py:149-201) and Postnet (model.py:103-146): v0 implementation on the fp32 SIMT
// GEMM engine with channels-last activations (conv1d == GEMM over K = taps x Cin).
#include <stdlib.h>

#include "conv_tc.h"
#include "gemm_f32.cuh"
#include "train_layers.h"
#include "model.h"

namespace t2 {

// ---- small kernels ----------------------------------------------------------------------------
__global__ void embed_kernel(const int64_t* __restrict__ text, const float* __restrict__ emb,
                             float* __restrict__ out, int rows, int n_symbols) {
  const int r = blockIdx.x;
  if (r >= rows) return;
  long id = text[r];
  if (id < 0) id = 0;
  if (id >= n_symbols) id = n_symbols - 1;
  const float4* src = reinterpret_cast<const float4*>(emb + id * kEnc);
  float4* dst = reinterpret_cast<float4*>(out + (long)r * kEnc);
  for (int i = threadIdx.x; i < kEnc / 4; i += blockDim.x) dst[i] = src[i];
}

// eval-mode BatchNorm folded to y = x*scale + shift                    (model.py:118, 165)
__global__ void bn_fold_kernel(const float* g, const float* b, const float* mean, const float* var,
                               float eps, float* scale, float* shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float s = g[c] / sqrtf(var[c] + eps);
  scale[c] = s;
  shift[c] = b[c] - mean[c] * s;
}

// training-mode BatchNorm statistics over M rows (biased variance, padded positions included,
// as the reference does) + running-stat update with momentum 0.1 (unbiased variance).
__global__ void __launch_bounds__(256)
bn_batch_stats_kernel(const float* __restrict__ x, int M, int C, const float* g, const float* b,
                      float eps, float* scale, float* shift, float* run_mean, float* run_var) {
  __shared__ float red[8][33];
  const int cl = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float s = 0.f;
  if (c < C) for (int r = rg; r < M; r += 8) s += x[(long)r * C + c];
  red[rg][cl] = s;
  __syncthreads();
  float mean = 0.f;
  for (int i = 0; i < 8; ++i) mean += red[i][cl];
  mean /= (float)M;
  __syncthreads();
  float q = 0.f;
  if (c < C) for (int r = rg; r < M; r += 8) { const float d = x[(long)r * C + c] - mean; q = fmaf(d, d, q); }
  red[rg][cl] = q;
  __syncthreads();
  if (rg == 0 && c < C) {
    float var = 0.f;
    for (int i = 0; i < 8; ++i) var += red[i][cl];
    var /= (float)M;
    const float sc = g[c] / sqrtf(var + eps);
    scale[c] = sc;
    shift[c] = b[c] - mean * sc;
    if (run_mean) {
      run_mean[c] = 0.9f * run_mean[c] + 0.1f * mean;
      run_var[c] = 0.9f * run_var[c] + 0.1f * var * ((float)M / (float)(M > 1 ? M - 1 : 1));
    }
  }
}

// y = act(x*scale + shift) with optional dropout; in place on (M, C)
__global__ void bn_apply_kernel(float* x, long n, int C, const float* scale, const float* shift,
                                int act, const uint8_t* keep, long keep_ld_t, int T, int philox,
                                uint64_t seed, uint32_t site, float p_drop) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % C);
  float v = x[i] * scale[c] + shift[c];
  if (act == ACT_RELU) v = fmaxf(v, 0.f);
  else if (act == ACT_TANH) v = tanhf(v);
  if (keep) {   // reference-layout mask (B, C, T): row m = b*T + t
    const long mrow = i / C; const long b = mrow / T, t = mrow % T;
    v = keep[(b * C + c) * (long)T + t] ? v * (1.f / (1.f - p_drop)) : 0.f;
  } else if (philox) {
    v = philox_keep(seed, site, (uint64_t)i, p_drop) ? v * (1.f / (1.f - p_drop)) : 0.f;
  }
  x[i] = v;
}

__global__ void mask_rows_kernel(const float* __restrict__ in, long in_batch_stride, float* __restrict__ out,
                                 const int32_t* __restrict__ len, int B, int T, int C) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)B * T * C) return;
  const long row = i / C; const int b = (int)(row / T), t = (int)(row % T);
  const int c = (int)(i - row * C);
  out[i] = (len == nullptr || t < len[b]) ? in[(long)b * in_batch_stride + (long)t * C + c] : 0.f;
}

// (B*T, C) channels-last -> (B, C, T) with optional residual and length mask (training-mode tail of
// the postnet; the eval path fuses this into the last conv's epilogue)
__global__ void transpose_residual_kernel(const float* __restrict__ y, const float* __restrict__ R,
                                          const int32_t* __restrict__ len, float* __restrict__ out,
                                          int B, int T, int C) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)B * T * C) return;
  const int t = (int)(i % T); const long r = i / T; const int c = (int)(r % C); const int b = (int)(r / C);
  const long m = (long)b * T + t;
  float v = y[m * C + c];
  if (R) v += R[m * C + c];
  if (len && t >= len[b]) v = 0.f;
  out[i] = v;
}

// One time step of both directions of the encoder BiLSTM (model.py:169-171, 180-188).
// grid (16 unit blocks, 2 directions, batch chunks of 64); 256 threads = 64 rows x 4 unit groups.
__global__ void __launch_bounds__(256)
enc_lstm_step_kernel(const float* __restrict__ gin,   // (B, T, 2048) W_ih x + b, fwd | reverse
                     const float* __restrict__ whh_f, const float* __restrict__ whh_r,
                     const float* __restrict__ h_in, float* __restrict__ h_out,  // (2, B, 256)
                     float* __restrict__ c,                                        // (2, B, 256)
                     float* __restrict__ memory,                                   // (B, T, 512)
                     const int32_t* __restrict__ lengths, int B, int T, int step) {
  extern __shared__ float sm[];
  float* ws = sm;                    // [64 rows][256]
  float* hs = sm + 64 * kEncH;       // [64][257]
  const int ub = blockIdx.x, dir = blockIdx.y, b0 = blockIdx.z * 64;
  const int tid = threadIdx.x;
  const int t = dir == 0 ? step : T - 1 - step;
  const float* whh = dir == 0 ? whh_f : whh_r;
  // smem row r = ul*4 + gate  <-  weight_hh row gate*256 + (ub*16 + ul)
  for (int i = tid; i < 64 * kEncH / 4; i += 256) {
    const int r = i / (kEncH / 4), k4 = i % (kEncH / 4);
    const int ul = r >> 2, gate = r & 3;
    reinterpret_cast<float4*>(ws)[i] =
        reinterpret_cast<const float4*>(whh + ((long)gate * kEncH + ub * 16 + ul) * kEncH)[k4];
  }
  for (int i = tid; i < 64 * kEncH; i += 256) {
    const int bl = i / kEncH, k = i % kEncH;
    hs[bl * (kEncH + 1) + k] = (b0 + bl < B) ? h_in[((long)dir * B + b0 + bl) * kEncH + k] : 0.f;
  }
  __syncthreads();
  const int bl = tid & 63, ug = tid >> 6;
  const int b = b0 + bl;
  float acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int g = 0; g < 4; ++g) acc[u][g] = 0.f;
  const float* hrow = hs + bl * (kEncH + 1);
  const float* wbase = ws + (ug * 16) * kEncH;
  for (int k = 0; k < kEncH; ++k) {
    const float hv = hrow[k];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[u][g] = fmaf(wbase[(u * 4 + g) * kEncH + k], hv, acc[u][g]);
  }
  if (b >= B) return;
  const bool valid = lengths == nullptr || t < lengths[b];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int unit = ub * 16 + ug * 4 + u;
    const float* gp = gin + ((long)b * T + t) * (8 * kEncH) + dir * 4 * kEncH;
    const long si = ((long)dir * B + b) * kEncH + unit;
    float hn = hrow[unit], cn = c[si];
    if (valid) {
      const float gi = 1.f / (1.f + expf(-(acc[u][0] + gp[unit])));
      const float gf = 1.f / (1.f + expf(-(acc[u][1] + gp[kEncH + unit])));
      const float gg = tanhf(acc[u][2] + gp[2 * kEncH + unit]);
      const float go = 1.f / (1.f + expf(-(acc[u][3] + gp[3 * kEncH + unit])));
      cn = gf * cn + gi * gg;
      hn = go * tanhf(cn);
      c[si] = cn;
    }
    h_out[si] = hn;
    memory[((long)b * T + t) * kEnc + dir * kEncH + unit] = valid ? hn : 0.f;
  }
}

// Persistent variant of the recurrence: ONE cooperative launch runs all T steps of both directions.
// 128 CTAs = 2 directions x 64 unit blocks of 4 hidden units (16 gate rows); the CTA's W_hh slice stays
// in shared memory for the whole sequence, h is exchanged through a double-buffered global array and a
// grid-wide barrier (monotonic counter) per step.  256 threads = 64 batch rows x 4 units; each thread
// owns the 4 gates of ONE (row, unit) pair, so the cell state lives in a register.
struct EncLstmCtrl { unsigned int bar_count; unsigned int pad_[3]; };

__global__ void __launch_bounds__(256, 1)
enc_lstm_persistent_kernel(const float* __restrict__ gin, const float* __restrict__ whh_f,
                           const float* __restrict__ whh_r, float* __restrict__ hbuf,   // (2 buffers, 2 dirs, B, 256)
                           float* __restrict__ memory, const int32_t* __restrict__ lengths, int B, int T,
                           EncLstmCtrl* ctrl, float* __restrict__ st_gates, float* __restrict__ st_c) {
  extern __shared__ float sm[];
  float* ws = sm;                        // [64 k4][16 rows][4]   (row = gate*4 + unit_local)
  float* hs = sm + kEncH * 16;           // [64 k4][64 batch][4]
  const int cta = blockIdx.x, dir = cta >> 6, ub = cta & 63;
  const int tid = threadIdx.x;
  const float* whh = dir == 0 ? whh_f : whh_r;
  for (int i = tid; i < 16 * (kEncH / 4); i += 256) {
    const int r = i / (kEncH / 4), k4 = i % (kEncH / 4);
    const int gate = r >> 2, ul = r & 3;
    const float4 v = reinterpret_cast<const float4*>(whh + ((long)gate * kEncH + ub * 4 + ul) * kEncH)[k4];
    reinterpret_cast<float4*>(ws)[k4 * 16 + r] = v;
  }
  const int b = tid & 63, ul = tid >> 6;             // batch row, local unit
  const int unit = ub * 4 + ul;
  float c = 0.f, hprev = 0.f;
  unsigned int target = 0;
  __syncthreads();
  for (int step = 0; step < T; ++step) {
    const int t = dir == 0 ? step : T - 1 - step;
    const float* hin = hbuf + ((long)(step & 1) * 2 + dir) * B * kEncH;
    float* hout = hbuf + ((long)((step + 1) & 1) * 2 + dir) * B * kEncH;
    // gate pre-activations from the input projection (issued first: their latency hides behind the GEMV)
    float gi4[4] = {0.f, 0.f, 0.f, 0.f};
    const bool live = b < B;
    const bool valid = live && (lengths == nullptr || t < lengths[b]);
    if (live) {
      const float* gp = gin + ((long)b * T + t) * (8 * kEncH) + dir * 4 * kEncH + unit;
#pragma unroll
      for (int g = 0; g < 4; ++g) gi4[g] = __ldg(gp + g * kEncH);
    }
    for (int i = tid; i < 64 * (kEncH / 4); i += 256) {       // h_(t-1) of all rows -> [k4][b][4]
      const int bb = i / (kEncH / 4), k4 = i % (kEncH / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bb < B) v = __ldcg(reinterpret_cast<const float4*>(hin + (long)bb * kEncH) + k4);
      reinterpret_cast<float4*>(hs)[k4 * 64 + bb] = v;
    }
    __syncthreads();
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int k4 = 0; k4 < kEncH / 4; ++k4) {
      const float4 hv = reinterpret_cast<const float4*>(hs)[k4 * 64 + b];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float4 wv = reinterpret_cast<const float4*>(ws)[k4 * 16 + g * 4 + ul];
        acc[g] = fmaf(wv.x, hv.x, acc[g]); acc[g] = fmaf(wv.y, hv.y, acc[g]);
        acc[g] = fmaf(wv.z, hv.z, acc[g]); acc[g] = fmaf(wv.w, hv.w, acc[g]);
      }
    }
    if (live) {
      float hn = hprev;
      if (valid) {
        const float gi = 1.f / (1.f + expf(-(acc[0] + gi4[0])));
        const float gf = 1.f / (1.f + expf(-(acc[1] + gi4[1])));
        const float gg = tanhf(acc[2] + gi4[2]);
        const float go = 1.f / (1.f + expf(-(acc[3] + gi4[3])));
        c = gf * c + gi * gg;
        hn = go * tanhf(c);
        if (st_gates) {   // training stash: gate activations in the layout of gin
          float* sg = st_gates + ((long)b * T + t) * (8 * kEncH) + dir * 4 * kEncH + unit;
          sg[0] = gi; sg[kEncH] = gf; sg[2 * kEncH] = gg; sg[3 * kEncH] = go;
        }
      }
      if (st_c) st_c[((long)b * T + t) * kEnc + dir * kEncH + unit] = c;
      hprev = hn;
      hout[(long)b * kEncH + unit] = hn;
      memory[((long)b * T + t) * kEnc + dir * kEncH + unit] = valid ? hn : 0.f;
    }
    // grid barrier (monotonic counter, red.release / ld.acquire)
    __syncthreads();
    target += gridDim.x;
    if (tid == 0) {
      __threadfence();
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&ctrl->bar_count) : "memory");
      const unsigned long long t0 = clock64();
      while (true) {
        unsigned int cnt;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cnt) : "l"(&ctrl->bar_count) : "memory");
        if ((int)(cnt - target) >= 0) break;
        if (clock64() - t0 > (1ull << 32)) __trap();
      }
    }
    __syncthreads();
  }
}

// ---- host side ----------------------------------------------------------------------------------
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static bool use_tc() { const char* e = getenv("T2_CONV_IMPL"); return !(e && e[0] == 's'); }   // "simt" selects the fp32 SIMT path

size_t encoder_ws_bytes(int B, int T) {
  const size_t act = align256((size_t)B * T * kEnc * 4);
  return 2 * act + align256((size_t)B * T * 8 * kEncH * 4) + 3 * align256((size_t)2 * B * kEncH * 4) +
         2 * align256(8 * kEncH * 4) + 1024 + 2 * align256(tc_planes_bytes(B, T, kEnc));
}

static int conv_bn_layer(T2Model* m, const float* x, float* y, int B, int T, int cin, int cout,
                         const float* wpk, int wbase, int act, int training, const uint8_t* keep,
                         uint64_t seed, uint32_t site, float* scale, float* shift,
                         int out_transposed, const float* R, const int32_t* row_len,
                         cudaStream_t s) {
  // wbase: index of conv.weight in the state_dict table (conv.bias, bn.weight, bn.bias,
  // running_mean, running_var follow)
  const float* cbias = m->w[wbase + 1];
  GemmArgs g;
  g.seg[0] = {x, cin, wpk, (long)kConvK * cin, kConvK * cin};
  g.M = B * T; g.N = cout; g.C = y; g.ldc = cout; g.bias = cbias;
  g.conv_T = T; g.conv_cin = cin; g.conv_pad = (kConvK - 1) / 2;
  if (!training) {
    bn_fold_kernel<<<(cout + 127) / 128, 128, 0, s>>>(m->w[wbase + 2], m->w[wbase + 3], m->w[wbase + 4],
                                                      m->w[wbase + 5], m->cfg.bn_eps, scale, shift, cout);
    T2_LAUNCH_CHECK();
    g.scale = scale; g.shift = shift; g.act = act;
    g.out_transposed = out_transposed; g.R = R; g.ldr = cout; g.row_len = row_len;
    return gemm_f32(g, s);
  }
  // training: raw conv -> batch statistics -> normalise + activation + dropout
  if (out_transposed) return fail(T2_ERR_UNSUPPORTED, "training-mode final postnet layer uses the generic path");
  T2_TRY(gemm_f32(g, s));
  bn_batch_stats_kernel<<<(cout + 31) / 32, 256, 0, s>>>(
      y, B * T, cout, m->w[wbase + 2], m->w[wbase + 3], m->cfg.bn_eps, scale, shift,
      const_cast<float*>(m->w[wbase + 4]), const_cast<float*>(m->w[wbase + 5]));
  T2_LAUNCH_CHECK();
  const long n = (long)B * T * cout;
  bn_apply_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(y, n, cout, scale, shift, act, keep, T, T,
                                                              keep ? 0 : 1, seed, site, 0.5f);
  T2_LAUNCH_CHECK();
  return T2_OK;
}

int encoder_forward(T2Model* m, const T2EncoderArgs* a, cudaStream_t s) {
  const int B = a->B, T = a->T;
  if (B <= 0 || T <= 0) return fail(T2_ERR_INVALID, "encoder: empty batch");
  if (a->ws_bytes < encoder_ws_bytes(B, T)) return fail(T2_ERR_WORKSPACE, "encoder workspace too small");
  char* p = (char*)a->ws;
  const size_t act = align256((size_t)B * T * kEnc * 4);
  float* x0 = (float*)p; p += act;
  float* x1 = (float*)p; p += act;
  float* gin = (float*)p; p += align256((size_t)B * T * 8 * kEncH * 4);
  float* hbuf0 = (float*)p; p += align256((size_t)2 * B * kEncH * 4);
  float* hbuf1 = (float*)p; p += align256((size_t)2 * B * kEncH * 4);
  float* cbuf = (float*)p; p += align256((size_t)2 * B * kEncH * 4);
  float* scale = (float*)p; p += align256(8 * kEncH * 4);
  float* shift = (float*)p; p += align256(8 * kEncH * 4);
  EncLstmCtrl* lctrl = (EncLstmCtrl*)p; p += 256;
  p = (char*)align256((size_t)p);
  __half* pl0 = (__half*)p; p += align256(tc_planes_bytes(B, T, kEnc));
  __half* pl1 = (__half*)p; p += align256(tc_planes_bytes(B, T, kEnc));
  const bool tc = use_tc() && !a->training && !a->stash;
  float* st_gates = nullptr; float* st_c = nullptr;

  if (a->stash) {
    // autograd path: fp32 conv stack with the activations kept for the backward pass (train_layers.cu)
    const float* xl = nullptr;
    T2_TRY(encoder_convs_train(m, a, s, &xl, &st_gates, &st_c, pl0));
    GemmArgs g;
    g.seg[0] = {xl, kEnc, m->enc_lstm_wih, kEnc, kEnc};
    g.M = B * T; g.N = 8 * kEncH; g.C = gin; g.ldc = 8 * kEncH; g.bias = m->enc_lstm_b;
    T2_TRY(gemm_f32(g, s));
  } else if (tc) {
    // tensor-core path: planes -> 3 x (conv k5 + folded BN + ReLU) -> LSTM input projection (fp32 rows)
    if (a->embedded) T2_TRY(tc_rows_to_planes(a->embedded, (long)T * kEnc, kEnc, kEnc, nullptr, B, T, pl0, s));
    else T2_TRY(tc_embed_to_planes(a->text, m->w[W_EMB], m->cfg.n_symbols, B, T, pl0, s));
    __half* cur = pl0; __half* nxt = pl1;
    for (int i = 0; i < 3; ++i) {                                                           // model.py:174-175, 194
      const int wb = W_ENC_CONV0 + 7 * i;
      T2_TRY(tc_fold_bn(m->w[wb + 1], m->w[wb + 2], m->w[wb + 3], m->w[wb + 4], m->w[wb + 5], m->cfg.bn_eps, scale, shift, kEnc, s));
      TcConvArgs c; memset(&c, 0, sizeof(c));
      c.in = cur; c.cin_pad = kEnc; c.wimg = m->tc_enc_conv[i]; c.taps = kConvK; c.B = B; c.T = T; c.cout = kEnc; c.nt_rows = 128;
      c.scale = scale; c.shift = shift; c.act = 1; c.out_mode = 0; c.out_planes = nxt;
      T2_TRY(tc_conv(c, s));
      __half* tmp = cur; cur = nxt; nxt = tmp;
    }
    T2_TRY(tc_fold_bn(m->enc_lstm_b, nullptr, nullptr, nullptr, nullptr, 0.f, scale, shift, 8 * kEncH, s));
    TcConvArgs c; memset(&c, 0, sizeof(c));
    c.in = cur; c.cin_pad = kEnc; c.wimg = m->tc_enc_wih; c.taps = 1; c.B = B; c.T = T; c.cout = 8 * kEncH; c.nt_rows = 128;
    c.scale = scale; c.shift = shift; c.act = 0; c.out_mode = 1; c.out_f32 = gin; c.ldo = 8 * kEncH;
    T2_TRY(tc_conv(c, s));
  } else {
    if (a->embedded) {
      T2_CUDA(cudaMemcpyAsync(x0, a->embedded, (size_t)B * T * kEnc * 4, cudaMemcpyDeviceToDevice, s));
    } else {
      embed_kernel<<<B * T, 128, 0, s>>>(a->text, m->w[W_EMB], x0, B * T, m->cfg.n_symbols);   // model.py:503/518
      T2_LAUNCH_CHECK();
    }
    float* cur = x0; float* nxt = x1;
    for (int i = 0; i < 3; ++i) {                                                             // model.py:174-175
      const uint8_t* keep = (a->training && a->keep) ? a->keep + (size_t)i * B * kEnc * T : nullptr;
      T2_TRY(conv_bn_layer(m, cur, nxt, B, T, kEnc, kEnc, m->enc_conv_w[i], W_ENC_CONV0 + 7 * i, ACT_RELU,
                           a->training, keep, a->seed, 1000 + i, scale, shift, 0, nullptr, nullptr, s));
      float* tmp = cur; cur = nxt; nxt = tmp;
    }
    {  // W_ih x + b_ih + b_hh for every time step and both directions
      GemmArgs g;
      g.seg[0] = {cur, kEnc, m->enc_lstm_wih, kEnc, kEnc};
      g.M = B * T; g.N = 8 * kEncH; g.C = gin; g.ldc = 8 * kEncH; g.bias = m->enc_lstm_b;
      T2_TRY(gemm_f32(g, s));
    }
  }
  T2_CUDA(cudaMemsetAsync(hbuf0, 0, (size_t)2 * B * kEncH * 4, s));
  T2_CUDA(cudaMemsetAsync(cbuf, 0, (size_t)2 * B * kEncH * 4, s));
  if (B <= 64 && m->sm_count >= 128 && getenv("T2_ENC_LSTM_STEPWISE") == nullptr) {
    // persistent recurrence: one cooperative launch for all T steps of both directions
    T2_CUDA(cudaMemsetAsync(hbuf1, 0, (size_t)2 * B * kEncH * 4, s));
    T2_CUDA(cudaMemsetAsync(lctrl, 0, 256, s));
    float* hb = hbuf0;   // hbuf0 and hbuf1 are adjacent (2 x (2, B, 256) when align256 adds no padding)
    if (hbuf1 != hbuf0 + (size_t)2 * B * kEncH) return fail(T2_ERR_INVALID, "encoder: h buffers not adjacent");
    const size_t psm = (size_t)(kEncH * 16 + kEncH * 64) * sizeof(float);
    T2_CUDA(cudaFuncSetAttribute(enc_lstm_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
    const float* whf = m->w[W_ENC_LSTM + 1]; const float* whr = m->w[W_ENC_LSTM + 5];
    const int32_t* lens = a->lengths; float* mem = a->memory; int Bv = B, Tv = T;
    void* args[] = {(void*)&gin, (void*)&whf, (void*)&whr, (void*)&hb, (void*)&mem, (void*)&lens, (void*)&Bv, (void*)&Tv, (void*)&lctrl,
                    (void*)&st_gates, (void*)&st_c};
    T2_CUDA(cudaLaunchCooperativeKernel((void*)enc_lstm_persistent_kernel, dim3(128), dim3(256), args, psm, s));
    g_launch_count++;
    if (a->stash) T2_TRY(encoder_stash_output(a, s));
    return T2_OK;
  }
  if (a->stash) return fail(T2_ERR_UNSUPPORTED, "encoder: the training stash needs B <= 64 (persistent BiLSTM kernel)");
  const size_t smem = (64 * kEncH + 64 * (kEncH + 1)) * sizeof(float);
  T2_CUDA(cudaFuncSetAttribute(enc_lstm_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  float* hin = hbuf0; float* hout = hbuf1;
  for (int step = 0; step < T; ++step) {
    enc_lstm_step_kernel<<<dim3(16, 2, (B + 63) / 64), 256, smem, s>>>(
        gin, m->w[W_ENC_LSTM + 1], m->w[W_ENC_LSTM + 5], hin, hout, cbuf, a->memory, a->lengths, B, T, step);
    T2_LAUNCH_CHECK();
    float* tmp = hin; hin = hout; hout = tmp;
  }
  return T2_OK;
}

size_t postnet_ws_bytes(int B, int T) {
  return 2 * align256((size_t)B * T * kPost * 4) + align256((size_t)B * T * kMel * 4) + 2 * align256(kPost * 4) + 1024 +
         2 * align256(tc_planes_bytes(B, T, kPost));
}

int postnet_forward(T2Model* m, const T2PostnetArgs* a, cudaStream_t s) {
  const int B = a->B, T = a->T;
  if (B <= 0 || T <= 0) return fail(T2_ERR_INVALID, "postnet: empty batch");
  if (a->ws_bytes < postnet_ws_bytes(B, T)) return fail(T2_ERR_WORKSPACE, "postnet workspace too small");
  char* p = (char*)a->ws;
  const size_t act = align256((size_t)B * T * kPost * 4);
  float* y0 = (float*)p; p += act;
  float* y1 = (float*)p; p += act;
  float* xin = (float*)p; p += align256((size_t)B * T * kMel * 4);
  float* scale = (float*)p; p += align256(kPost * 4);
  float* shift = (float*)p; p += align256(kPost * 4);
  p = (char*)align256((size_t)p);
  __half* pl0 = (__half*)p; p += align256(tc_planes_bytes(B, T, kPost));
  __half* pl1 = (__half*)p; p += align256(tc_planes_bytes(B, T, kPost));
  if (use_tc() && !a->training) {
    // tensor-core path (model.py:141-146 + residual :511/:524): mel -> planes -> 4 x (conv+BN+tanh) -> conv+BN (+mel)
    const long bs = a->mel_batch_stride ? a->mel_batch_stride : (long)T * kMel;
    T2_TRY(tc_rows_to_planes(a->mel, bs, kMel, 128, a->lengths, B, T, pl0, s));
    __half* cur = pl0; __half* nxt = pl1;
    for (int i = 0; i < 5; ++i) {
      const int wb = W_POST_CONV0 + 7 * i;
      const int cout = i == 4 ? kMel : kPost;
      T2_TRY(tc_fold_bn(m->w[wb + 1], m->w[wb + 2], m->w[wb + 3], m->w[wb + 4], m->w[wb + 5], m->cfg.bn_eps, scale, shift, cout, s));
      TcConvArgs c; memset(&c, 0, sizeof(c));
      c.in = cur; c.cin_pad = i == 0 ? 128 : kPost; c.wimg = m->tc_post_conv[i]; c.taps = kConvK; c.B = B; c.T = T;
      c.cout = cout; c.nt_rows = i == 4 ? 80 : 128; c.scale = scale; c.shift = shift; c.act = i == 4 ? 0 : 2;
      if (i < 4) { c.out_mode = 0; c.out_planes = nxt; }
      else { c.out_mode = 2; c.out_f32 = a->mel_post; c.residual = a->add_residual ? a->mel : nullptr; c.res_batch_stride = bs; c.row_len = a->lengths; }
      T2_TRY(tc_conv(c, s));
      __half* tmp = cur; cur = nxt; nxt = tmp;
    }
    return T2_OK;
  }
  const long n_in = (long)B * T * kMel;
  mask_rows_kernel<<<(unsigned)((n_in + 255) / 256), 256, 0, s>>>(
      a->mel, a->mel_batch_stride ? a->mel_batch_stride : (long)T * kMel, xin, a->lengths, B, T, kMel);
  T2_LAUNCH_CHECK();
  const float* cur = xin; float* nxt = y0;
  for (int i = 0; i < 5; ++i) {                                      // model.py:141-146
    const int cin = i == 0 ? kMel : kPost, cout = i == 4 ? kMel : kPost;
    const bool last = i == 4;
    const uint8_t* keep = nullptr;
    if (a->training && a->keep) keep = a->keep + (i < 4 ? (size_t)i * B * kPost * T : (size_t)4 * B * kPost * T);   // [(B,512,T)]*4 + (B,80,T)
    if (last && !a->training) {
      T2_TRY(conv_bn_layer(m, cur, a->mel_post, B, T, cin, cout, m->post_conv_w[i], W_POST_CONV0 + 7 * i,
                           ACT_NONE, 0, nullptr, 0, 0, scale, shift, 1, a->add_residual ? xin : nullptr, a->lengths, s));
    } else {
      T2_TRY(conv_bn_layer(m, cur, nxt, B, T, cin, cout, m->post_conv_w[i], W_POST_CONV0 + 7 * i,
                           last ? ACT_NONE : ACT_TANH, a->training, keep, a->seed, 2000 + i, scale, shift, 0,
                           nullptr, nullptr, s));
      if (last) {   // training-mode last layer: separate transpose + residual
        const long n = (long)B * T * kMel;
        transpose_residual_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(
            nxt, a->add_residual ? xin : nullptr, a->lengths, a->mel_post, B, T, kMel);
        T2_LAUNCH_CHECK();
        return T2_OK;
      }
      cur = nxt; nxt = (nxt == y0) ? y1 : y0;
    }
  }
  return T2_OK;
}

}  // namespace t2
