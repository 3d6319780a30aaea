// Backward pass of the teacher-forced decoder (the autograd graph of Decoder.forward, model.py:381-416),
// fp32, derived by hand (tools/bwd_algorithm_check.py is the CPU statement of exactly this decomposition,
// checked against torch autograd through the oracle).
//
// Phase 1 (reverse time, 5 kernels per step; carries = gradients wrt the step-t states coming from t+1):
//   KA lstm_bwd   : g_dh = carry + W_P^T [d_mel_t ; d_gate_t]  -> dropout mask -> LSTMCell backward  -> dG_dec[t], g_dc
//   KB skinny_nn  : [g_ah | g_ctx | g_dh'] partials = dG_dec[t] . [W_ih^d | W_hh^d]        (model.py:366-369)
//   KC attention  : g_ctx total -> g_aw -> softmax backward -> g_s = g_e v (1 - tanh^2) -> g_q, g_pm (stashed),
//                   location layer backward -> carries for aw_{t-1}, awc_{t-1}            (model.py:43-86)
//   KD lstm_bwd   : g_ah = carry + dG_dec part + g_q W_q -> mask -> LSTMCell backward      -> dG_att[t], g_ac
//   KE skinny_nn  : [g_x2 | g_ctx' | g_ah'] partials = dG_att[t] . [W_ih^a | W_hh^a]      (model.py:352-354)
// Phase 2 (time batched, plain library GEMMs through cuBLAS): every weight gradient is dG^T . X over all
// T x B rows; d_memory = d_pm W_m + sum_t aw_t (x) g_ctx_t.
#include <cublas_v2.h>
#include <stdlib.h>
#include <string.h>

#include "decoder.h"
#include "gemm_f32.cuh"
#include "umma.cuh"
#include "gemm_tc.h"
#include "wgrad_tc.h"

namespace t2 {

// ---- cuBLAS plumbing (plain time-batched GEMMs) ------------------------------------------------
int blas_handle(T2Model* m, cudaStream_t s, cublasHandle_t* out) {
  if (!m->blas) {
    cublasHandle_t h;
    if (cublasCreate(&h) != CUBLAS_STATUS_SUCCESS) return fail(T2_ERR_CUDA, "cublasCreate failed");
    cublasSetMathMode(h, CUBLAS_DEFAULT_MATH);
    m->blas = h;
  }
  *out = (cublasHandle_t)m->blas;
  if (cublasSetStream(*out, s) != CUBLAS_STATUS_SUCCESS) return fail(T2_ERR_CUDA, "cublasSetStream failed");
  return T2_OK;
}
// Every dense product of the training path goes through gemm_rm(): our tcgen05 split-fp16 GEMM (gemm_tc.cu) by default;
// T2_GEMM=cublas selects plain cuBLAS fp32 sgemm, kept only as the independent cross-check of tests/test_gpu_backward.py.
bool use_cublas_gemm() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("T2_GEMM"); v = (e && !strcmp(e, "cublas")) ? 1 : 0; }
  return v == 1;
}
// row-major C (M x N) = op(A) . op(B) + beta C;  ta: A is stored (K x M);  tb: B is stored (N x K)
int gemm_rm(T2Model* m, cudaStream_t s, bool ta, bool tb, int M, int N, int K, const float* A, long lda, const float* B, long ldb,
            float* C, long ldc, float beta) {
  if (!use_cublas_gemm()) return gemm_tc_rm(m, s, ta, tb, M, N, K, A, lda, B, ldb, C, ldc, beta);
  cublasHandle_t h;
  T2_TRY(blas_handle(m, s, &h));
  const float alpha = 1.f;
  cublasStatus_t st = cublasSgemm(h, tb ? CUBLAS_OP_T : CUBLAS_OP_N, ta ? CUBLAS_OP_T : CUBLAS_OP_N, N, M, K, &alpha, B,
                                  (int)ldb, A, (int)lda, &beta, C, (int)ldc);
  if (st != CUBLAS_STATUS_SUCCESS) return fail(T2_ERR_CUDA, "cublasSgemm failed (%d) M=%d N=%d K=%d", (int)st, M, N, K);
  g_launch_count++;
  return T2_OK;
}
int gemm_rm_wgrad(T2Model* m, cudaStream_t s, bool ta, bool tb, int M, int N, int K, const float* A, long lda, const float* B,
                  long ldb, float* C, long ldc, float beta) {
  return gemm_rm(m, s, ta, tb, M, N, K, A, lda, B, ldb, C, ldc, beta);
}
// out[c] = sum over rows of X[r * ld + c] (bias gradients; round 1 multiplied by a vector of ones through cuBLAS)
int colsum_rm(T2Model* m, cudaStream_t s, const float* X, long ld, long rows, int cols, float* out) {
  return colsum_f32(m, s, X, ld, rows, cols, out);
}

namespace {

constexpr int kSplitB = 7;       // reduction splits of the skinny GEMMs (partials summed by the consumer):
constexpr int kSplitE = 10;      // 20 x 7 = 14 x 10 = 140 CTAs = one wave of one CTA per SM
constexpr int kSplitMax = 10;
constexpr int kPBld = 1536 + 1024;   // [g_ah (1024) | g_ctx (512) | g_dh' (1024)]
constexpr int kPEld = 768 + 1024;    // [g_x2 (256) | g_ctx' (512) | g_ah' (1024)]
constexpr int kTaps = 2 * kLocK;     // 62 taps of the fused location filter
constexpr int kColsLd = 64;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// LSTMCell backward for one step (KA / KD).  grid (1024 / 256, B), block 256: thread = one hidden unit.
// ---------------------------------------------------------------------------------------------
struct LstmBwdArgs {
  // g_h = sum over sources of sum_s src[(s * 64 + b) * ld + off + unit]  +  sum_o vec[b][o] * Wv[o][unit]
  const float* src0; int ld0, off0, ns0;
  const float* src1; int ld1, off1, ns1;
  const float* add; int ldadd;             // (B, >= 1024) direct term or null
  const float* vec; int nvec, ldvec;       // (B, nvec) rows, row stride ldvec
  const float* Wv; int ldwv;               // (nvec, >= 1024)
  const uint8_t* keep;                     // (B, 1024) of this step or null
  int dropout; uint64_t seed; uint32_t site; float p;
  const float* gates;                      // (B, 4096) activations of this step
  const float* c; const float* c_prev;     // (B, 1024)
  float* g_c;                              // (B, 1024) carry, in/out
  float* dG;                               // (B, 4096) out
  uint8_t* img; float* inv_scale;          // tensor-core path: scaled split-fp16 operand image of dG + 1/scale per row
};

__global__ void __launch_bounds__(256) lstm_bwd_kernel(const LstmBwdArgs a) {
  __shared__ float s_vec[128];
  const int b = blockIdx.y, unit = blockIdx.x * 256 + threadIdx.x;
  for (int i = threadIdx.x; i < a.nvec; i += 256) s_vec[i] = a.vec[(long)b * a.ldvec + i];
  __syncthreads();
  float g_h = a.add ? a.add[(long)b * a.ldadd + unit] : 0.f;
  if (a.src0) {
#pragma unroll 5
    for (int s = 0; s < a.ns0; ++s) g_h += a.src0[((long)s * 64 + b) * a.ld0 + a.off0 + unit];
  }
  if (a.src1) {
#pragma unroll 5
    for (int s = 0; s < a.ns1; ++s) g_h += a.src1[((long)s * 64 + b) * a.ld1 + a.off1 + unit];
  }
  {
    float p4[4] = {0.f, 0.f, 0.f, 0.f};     // nvec is a multiple of 4 (0 or 128)
#pragma unroll 4
    for (int o = 0; o < a.nvec; o += 4) {
#pragma unroll
      for (int i = 0; i < 4; ++i) p4[i] = fmaf(s_vec[o + i], __ldg(a.Wv + (long)(o + i) * a.ldwv + unit), p4[i]);
    }
    g_h += (p4[0] + p4[1]) + (p4[2] + p4[3]);
  }
  if (a.dropout) {
    const long idx = (long)b * 1024 + unit;
    const bool keep = a.keep ? a.keep[idx] != 0 : philox_keep(a.seed, a.site, (uint64_t)idx, a.p);
    g_h = keep ? g_h * (1.f / (1.f - a.p)) : 0.f;
  }
  const float* gp = a.gates + (long)b * 4096 + unit;
  const float gi = gp[0], gf = gp[1024], gg = gp[2048], go = gp[3072];
  const float c = a.c[(long)b * 1024 + unit], cp = a.c_prev[(long)b * 1024 + unit];
  const float tc = tanhf(c);
  const float d_o = g_h * tc;
  const float d_c = a.g_c[(long)b * 1024 + unit] + g_h * go * (1.f - tc * tc);
  float* dg = a.dG + (long)b * 4096 + unit;
  dg[0] = d_c * gg * gi * (1.f - gi);
  dg[1024] = d_c * cp * gf * (1.f - gf);
  dg[2048] = d_c * gi * (1.f - gg * gg);
  dg[3072] = d_o * go * (1.f - go);
  a.g_c[(long)b * 1024 + unit] = d_c * gf;
}

// Same computation with one block per batch row (1024 threads = hidden units): the block knows the row maximum of
// dG, scales the row by a power of two into [0.5, 1) and writes it as the split-fp16 operand image of the
// tensor-core GEMM (gradients span many orders of magnitude; fp16 does not) next to the fp32 copy.
__global__ void __launch_bounds__(1024) lstm_bwd_row_kernel(const LstmBwdArgs a) {
  __shared__ float s_vec[128];
  __shared__ float s_max[32];
  const int b = blockIdx.x, unit = threadIdx.x;
  for (int i = threadIdx.x; i < a.nvec; i += 1024) s_vec[i] = a.vec[(long)b * a.ldvec + i];
  __syncthreads();
  float g_h = a.add ? a.add[(long)b * a.ldadd + unit] : 0.f;
  if (a.src0) {
#pragma unroll 4
    for (int s = 0; s < a.ns0; ++s) g_h += a.src0[((long)s * 64 + b) * a.ld0 + a.off0 + unit];
  }
  if (a.src1) {
#pragma unroll 4
    for (int s = 0; s < a.ns1; ++s) g_h += a.src1[((long)s * 64 + b) * a.ld1 + a.off1 + unit];
  }
  {
    float p4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int o = 0; o < a.nvec; o += 4) {
#pragma unroll
      for (int i = 0; i < 4; ++i) p4[i] = fmaf(s_vec[o + i], __ldg(a.Wv + (long)(o + i) * a.ldwv + unit), p4[i]);
    }
    g_h += (p4[0] + p4[1]) + (p4[2] + p4[3]);
  }
  if (a.dropout) {
    const long idx = (long)b * 1024 + unit;
    const bool keep = a.keep ? a.keep[idx] != 0 : philox_keep(a.seed, a.site, (uint64_t)idx, a.p);
    g_h = keep ? g_h * (1.f / (1.f - a.p)) : 0.f;
  }
  const float* gp = a.gates + (long)b * 4096 + unit;
  const float gi = gp[0], gf = gp[1024], gg = gp[2048], go = gp[3072];
  const float c = a.c[(long)b * 1024 + unit], cp = a.c_prev[(long)b * 1024 + unit];
  const float tc = tanhf(c);
  const float d_o = g_h * tc;
  const float d_c = a.g_c[(long)b * 1024 + unit] + g_h * go * (1.f - tc * tc);
  float d[4];
  d[0] = d_c * gg * gi * (1.f - gi);
  d[1] = d_c * cp * gf * (1.f - gf);
  d[2] = d_c * gi * (1.f - gg * gg);
  d[3] = d_o * go * (1.f - go);
  float* dg = a.dG + (long)b * 4096 + unit;
#pragma unroll
  for (int g = 0; g < 4; ++g) dg[g * 1024] = d[g];
  a.g_c[(long)b * 1024 + unit] = d_c * gf;
  // row maximum -> power-of-two scale
  float mx = fmaxf(fmaxf(fabsf(d[0]), fabsf(d[1])), fmaxf(fabsf(d[2]), fabsf(d[3])));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = s_max[threadIdx.x & 31];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  int e = 0;
  if (mx > 0.f && mx < 3.0e38f) frexpf(mx, &e);           // mx = f * 2^e, f in [0.5, 1)
  e = e < -100 ? -100 : (e > 100 ? 100 : e);
  const float sc = ldexpf(1.f, -e);
  if (threadIdx.x == 0) a.inv_scale[b] = ldexpf(1.f, e);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int k = g * 1024 + unit;
    __half h, l;
    split_fp16(d[g] * sc, h, l);
    __half* hi = reinterpret_cast<__half*>(a.img + (size_t)(k >> 6) * 16384);
    __half* lo = hi + 64 * kChunkK;
    const uint32_t eo = img_elem_offset(b, k & 63);
    hi[eo] = h; lo[eo] = l;
  }
}

// ---------------------------------------------------------------------------------------------
// Skinny "NN" GEMM (KB / KE): P[s][m][col] = sum_{n in split s} A[m][n] * W[n][col], m < 64.
// The output columns are the concatenation of up to two weight matrices (row-major, N x ncols each).
// grid (total_cols / 128, kSplit), block 256, tile 64 x 128, 8 x 4 outputs per thread.
// ---------------------------------------------------------------------------------------------
struct SkinnyArgs {
  const float* A; int lda; int rows; int nred;
  const float* W0; int ldw0; int cols0;
  const float* W1; int ldw1; int cols1;
  float* P; int ldp;
  int nsplit;                               // gridDim.y; the nred / 32 chunks are divided as evenly as possible
};

__global__ void __launch_bounds__(256) skinny_nn_kernel(const SkinnyArgs a) {
  constexpr int BK = 32;
  __shared__ __align__(16) float As[BK][64 + 4];
  __shared__ __align__(16) float Bs[BK][128];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int tiles0 = a.cols0 >> 7;
  const int ct = blockIdx.x;
  const float* W; int ldw; int col_out;
  if (ct < tiles0) { W = a.W0 + ct * 128; ldw = a.ldw0; col_out = ct * 128; }
  else { W = a.W1 + (ct - tiles0) * 128; ldw = a.ldw1; col_out = ct * 128; }
  const int nchunks = a.nred / BK;
  const int n_begin = (int)((long)blockIdx.y * nchunks / a.nsplit) * BK;
  const int n_end = (int)((long)(blockIdx.y + 1) * nchunks / a.nsplit) * BK;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float4 ra[2], rb[4];
  auto load = [&](int n0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid * 2 + i, r = idx >> 3, q = idx & 7;
      ra[i] = r < a.rows ? *reinterpret_cast<const float4*>(a.A + (long)r * a.lda + n0 + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256, kk = idx >> 5, c4 = idx & 31;
      rb[i] = __ldg(reinterpret_cast<const float4*>(W + (long)(n0 + kk) * ldw + c4 * 4));
    }
  };
  load(n_begin);
  for (int n0 = n_begin; n0 < n_end; n0 += BK) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid * 2 + i, r = idx >> 3, q = idx & 7;
      As[q * 4 + 0][r] = ra[i].x; As[q * 4 + 1][r] = ra[i].y; As[q * 4 + 2][r] = ra[i].z; As[q * 4 + 3][r] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256, kk = idx >> 5, c4 = idx & 31;
      *reinterpret_cast<float4*>(&Bs[kk][c4 * 4]) = rb[i];
    }
    __syncthreads();
    if (n0 + BK < n_end) load(n0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
  float* out = a.P + ((long)blockIdx.y * 64 + ty * 8) * a.ldp + col_out + tx * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    *reinterpret_cast<float4*>(out + (long)i * a.ldp) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
}

// ---------------------------------------------------------------------------------------------
// Attention backward for one step (KC).  grid (2, B): CTA (h, b) handles attention dims [64h, 64h + 64) of row b.
// ---------------------------------------------------------------------------------------------
struct AttBwdArgs {
  int t, T, B, Te, carry;                 // carry = (t < T - 1)
  const int32_t* len;
  const float* memory; const float* pm; const float* q;   // (B,Te,512), (B,Te,128), (T,B,128)
  const float* v; const float* weff;      // (128), (128, 62)
  const float* align; const float* awc;   // (B,T,Te) forward weights / cumulative weights BEFORE step t
  const float* d_align;                   // (B,T,Te) or null
  const float* PE; const float* PB;       // partials of step t+1 (KE) and of this step (KB)
  int nsE, nsB;                           // number of partial sums in PE / PB
  const float* gproj;                     // (T,B,1536) = [d_mel_t ; d_gate_t] . W_PG, all steps (time batched)
  const float* weffT;                     // (62, 128) transposed fused location filter
  float* dctx; float* dx2; float* dq;     // (T,B,512), (T,B,256), (T,B,128)
  float* gs;                              // (T,B,Te,128)
  float* gcat;                            // (2 pingpong, 2 halves, B, 2, Te)
  float* cacc;                            // (2 pingpong, B, Te)
  float* dv;                              // (B, 128) accumulated
};

constexpr int kAttT = 512;               // threads of att_bwd_kernel: 16 warps hide the global-load latency
constexpr int kAttG = kAttT / 64;         // position groups in steps (4), (5)
__global__ void __launch_bounds__(kAttT) att_bwd_kernel(const AttBwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int t = a.t, T = a.T, B = a.B, Te = a.Te;
  const int TeP = Te + kLocK - 1;
  float* s_ctx = sm;                       // 512
  float* s_dy = s_ctx + 512;               // 96
  float* s_red = s_dy + 96;                // 64 + kAttG * 64 * 2
  float* s_aw = s_red + 64 + kAttG * 128;          // Te (padded to 4)
  const int Te4 = (Te + 3) & ~3;
  float* s_ge = s_aw + Te4;                // Te
  float* s_pad0 = s_ge + Te4;              // TeP (+ slack)
  const int TeP4 = (TeP + 7) & ~3;
  float* s_pad1 = s_pad0 + TeP4;
  float* s_gs = s_pad1 + TeP4;             // Te4 x 64
  float* s_u = s_gs + (size_t)Te4 * 64;    // Te x 65
  const int rd = (t + 1) & 1, wr = t & 1;

  // previous / cumulative attention weights of this step (zero padded by 15 each side)   model.py:358-360
  for (int i = tid; i < TeP4; i += kAttT) { s_pad0[i] = 0.f; s_pad1[i] = 0.f; }
  __syncthreads();
  for (int j = tid; j < Te; j += kAttT) {
    s_pad0[15 + j] = t > 0 ? a.align[((long)b * T + t - 1) * Te + j] : 0.f;
    s_pad1[15 + j] = a.awc[((long)b * T + t) * Te + j];
    s_aw[j] = a.align[((long)b * T + t) * Te + j];
    // carries into aw_t: location conv of step t+1 (channel 0) and the cumulative weights (channel 1 + running sum)
    float g = 0.f, cv = 0.f;
    if (a.carry) {
      const float* g0 = a.gcat + (((long)rd * 2 + 0) * B + b) * 2 * Te;
      const float* g1 = a.gcat + (((long)rd * 2 + 1) * B + b) * 2 * Te;
      g = g0[j] + g1[j];
      cv = a.cacc[((long)rd * B + b) * Te + j] + g0[Te + j] + g1[Te + j];
    }
    if (h == 0) a.cacc[((long)wr * B + b) * Te + j] = cv;
    g += cv;
    if (a.d_align) g += a.d_align[((long)b * T + t) * Te + j];
    s_ge[j] = g;
  }
  // (1) total gradient wrt ctx_t: carry from step t+1's attention LSTM input, decoder LSTM input, projection
  for (int c = tid; c < 512; c += kAttT) {
    float g = a.gproj[((long)t * B + b) * 1536 + 1024 + c];
    if (a.carry) {
#pragma unroll 5
      for (int s = 0; s < a.nsE; ++s) g += a.PE[((long)s * 64 + b) * kPEld + 256 + c];
    }
#pragma unroll 4
    for (int s = 0; s < a.nsB; ++s) g += a.PB[((long)s * 64 + b) * kPBld + 1024 + c];
    s_ctx[c] = g;
    if (h == 0) a.dctx[((long)t * B + b) * 512 + c] = g;
  }
  if (h == 0 && a.carry && tid < 256) {   // gradient wrt the prenet output of step t+1 (first 256 columns of KE's result)
    float g = 0.f;
#pragma unroll 5
    for (int s = 0; s < a.nsE; ++s) g += a.PE[((long)s * 64 + b) * kPEld + tid];
    a.dx2[((long)(t + 1) * B + b) * 256 + tid] = g;
  }
  __syncthreads();
  // (2) g_aw[j] += memory[j] . g_ctx   (4 rows per warp iteration: 16 independent 16-byte loads in flight)  model.py:83-84
  {
    float4 gc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) gc[i] = *reinterpret_cast<const float4*>(s_ctx + (i * 32 + lane) * 4);
    for (int j0 = warp * 8; j0 < Te; j0 += 8 * (kAttT / 32)) {
      float acc[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        acc[r] = 0.f;
        const int j = j0 + r < Te ? j0 + r : Te - 1;
        const float4* mr = reinterpret_cast<const float4*>(a.memory + ((long)b * Te + j) * 512);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 m = __ldg(mr + i * 32 + lane);
          acc[r] += m.x * gc[i].x + m.y * gc[i].y + m.z * gc[i].z + m.w * gc[i].w;
        }
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float v = warp_sum(acc[r]);
        if (lane == 0 && j0 + r < Te) s_ge[j0 + r] += v;
      }
    }
  }
  __syncthreads();
  // (3) softmax backward: g_e = aw * (g_aw - sum_j aw g_aw)                              model.py:82
  {
    float p = 0.f;
    for (int j = tid; j < Te; j += kAttT) p += s_aw[j] * s_ge[j];
    p = warp_sum(p);
    if (lane == 0) s_red[warp] = p;
    __syncthreads();
    float dot = 0.f;
#pragma unroll
    for (int w = 0; w < kAttT / 32; ++w) dot += s_red[w];
    __syncthreads();
    for (int j = tid; j < Te; j += kAttT) s_ge[j] = s_aw[j] * (s_ge[j] - dot);
    __syncthreads();
  }
  // (4) recompute s = q + pa + pm, g_s = g_e v (1 - tanh^2 s); thread = (attention dim, group of positions)
  {
    const int al = tid & 63, jg = tid >> 6, ag = h * 64 + al;
    float w[kTaps];
#pragma unroll
    for (int k = 0; k < kTaps; ++k) w[k] = __ldg(a.weffT + (long)k * kAtt + ag);
    const float qv = a.q[((long)t * B + b) * 128 + ag];
    const float vv = __ldg(a.v + ag);
    float gq = 0.f, dv = 0.f;
    for (int j0 = jg * 4; j0 < Te; j0 += 4 * kAttG) {
      float pa[4] = {0.f, 0.f, 0.f, 0.f}, pmv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pmv[i] = j0 + i < Te ? __ldg(a.pm + ((long)b * Te + j0 + i) * 128 + ag) : 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float* pad = (c == 0 ? s_pad0 : s_pad1) + j0;
        float w0 = pad[0], w1 = pad[1], w2 = pad[2], w3 = pad[3];
#pragma unroll
        for (int k = 0; k < kLocK; ++k) {
          const float wk = w[c * kLocK + k];
          pa[0] = fmaf(wk, w0, pa[0]); pa[1] = fmaf(wk, w1, pa[1]); pa[2] = fmaf(wk, w2, pa[2]); pa[3] = fmaf(wk, w3, pa[3]);
          w0 = w1; w1 = w2; w2 = w3; w3 = pad[k + 4];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = j0 + i;
        if (j < Te) {
          const float s = qv + pa[i] + pmv[i];
          const float th = tanhf(s);
          const float ge = s_ge[j];
          const float g = ge * vv * (1.f - th * th);
          dv = fmaf(ge, th, dv);
          gq += g;
          s_gs[j * 64 + al] = g;
          a.gs[(((long)t * B + b) * Te + j) * 128 + ag] = g;
        }
      }
    }
    s_red[64 + (jg * 64 + al) * 2 + 0] = gq;
    s_red[64 + (jg * 64 + al) * 2 + 1] = dv;
    __syncthreads();
    if (tid < 64) {
      float gq4 = 0.f, dv4 = 0.f;
#pragma unroll
      for (int g = 0; g < kAttG; ++g) { gq4 += s_red[64 + (g * 64 + tid) * 2]; dv4 += s_red[64 + (g * 64 + tid) * 2 + 1]; }
      a.dq[((long)t * B + b) * 128 + h * 64 + tid] = gq4;
      a.dv[(long)b * 128 + h * 64 + tid] += dv4;
    }
  }
  // (5) U[j][ck] = sum_a g_s[j][a] Weff[a][ck]  (this half's 64 attention dims)
  {
    const int ck = tid & 63, jg = tid >> 6;
    float wc[64];
#pragma unroll
    for (int al = 0; al < 64; ++al) wc[al] = ck < kTaps ? __ldg(a.weff + (long)(h * 64 + al) * kTaps + ck) : 0.f;
    for (int j = jg; j < Te; j += kAttG) {
      const float4* g4 = reinterpret_cast<const float4*>(s_gs + j * 64);
      float u = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 g = g4[i];
        u = fmaf(g.x, wc[i * 4], u); u = fmaf(g.y, wc[i * 4 + 1], u); u = fmaf(g.z, wc[i * 4 + 2], u); u = fmaf(g.w, wc[i * 4 + 3], u);
      }
      s_u[j * 65 + ck] = u;
    }
  }
  __syncthreads();
  // (6) transposed location conv: g_cat[c][j'] = sum_k U[j' + 15 - k][c * 31 + k]        model.py:23
  for (int i = tid; i < 2 * Te; i += kAttT) {
    const int c = i / Te, jp = i - c * Te;
    float g = 0.f;
    for (int k = 0; k < kLocK; ++k) {
      const int j = jp + 15 - k;
      if (j >= 0 && j < Te) g += s_u[j * 65 + c * kLocK + k];
    }
    a.gcat[(((long)wr * 2 + h) * B + b) * 2 * Te + i] = g;
  }
}

size_t att_bwd_smem(int Te) {
  const int Te4 = (Te + 3) & ~3, TeP4 = (Te + kLocK - 1 + 7) & ~3;
  return (size_t)(512 + 96 + 64 + kAttG * 128 + 2 * Te4 + 2 * TeP4 + (size_t)Te4 * 64 + (size_t)Te * 65 + 16) * sizeof(float);
}

// ---- small helper kernels -------------------------------------------------------------------
__global__ void dy_kernel(const float* __restrict__ d_mel, const float* __restrict__ d_gate, float* __restrict__ dY, int B, int T) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)T * B * 81) return;
  const int o = (int)(i % 81); const long r = i / 81; const int b = (int)(r % B); const int t = (int)(r / B);
  dY[i] = o < 80 ? d_mel[((long)b * T + t) * 80 + o] : d_gate[(long)b * T + t];
}
__global__ void awc_kernel(const float* __restrict__ align, float* __restrict__ awc, int B, int T, int Te) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Te) return;
  const int b = i / Te, j = i - b * Te;
  float run = 0.f;
  for (int t = 0; t < T; ++t) {
    awc[((long)b * T + t) * Te + j] = run;                 // cumulative weights BEFORE step t (model.py:365)
    run += align[((long)b * T + t) * Te + j];
  }
}
__global__ void weff_kernel(const float* __restrict__ wld, const float* __restrict__ wloc, float* __restrict__ weff,
                            float* __restrict__ weffT) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // Weff[a][c*31+k] = sum_f W_ld[a][f] W_loc[f][c][k]
  if (i >= kAtt * kTaps) return;
  const int a = i / kTaps, ck = i - a * kTaps;
  float s = 0.f;
  for (int f = 0; f < kLocF; ++f) s = fmaf(wld[a * kLocF + f], wloc[f * kTaps + ck], s);
  weff[i] = s;
  weffT[(long)ck * kAtt + a] = s;
}
__global__ void dweff_split_kernel(const float* __restrict__ dweff, const float* __restrict__ wld, const float* __restrict__ wloc,
                                   float* __restrict__ d_wld, float* __restrict__ d_wloc) {
  // dWeff (128, ld 64) -> dW_ld[a][f] = sum_ck dWeff[a][ck] W_loc[f][ck] ;  dW_loc[f][ck] = sum_a W_ld[a][f] dWeff[a][ck]
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kAtt * kLocF && d_wld) {
    const int a = i / kLocF, f = i - a * kLocF;
    float s = 0.f;
    for (int ck = 0; ck < kTaps; ++ck) s = fmaf(dweff[a * kColsLd + ck], wloc[f * kTaps + ck], s);
    d_wld[i] = s;
  }
  if (i < kLocF * kTaps && d_wloc) {
    const int f = i / kTaps, ck = i - f * kTaps;
    float s = 0.f;
    for (int a = 0; a < kAtt; ++a) s = fmaf(wld[a * kLocF + f], dweff[a * kColsLd + ck], s);
    d_wloc[i] = s;
  }
}
__global__ void im2col_kernel(const float* __restrict__ align, const float* __restrict__ awc, float* __restrict__ cols,
                              int B, int T, int Te) {
  // cols[(t, b, j)][c * 31 + k] = (c == 0 ? aw_{t-1} : awc_{t-1})[j + k - 15], zero outside; columns 62, 63 = 0
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)T * B * Te * kColsLd) return;
  const int ck = (int)(i & 63); const long r = i >> 6;
  const int j = (int)(r % Te); const long tb = r / Te; const int b = (int)(tb % B); const int t = (int)(tb / B);
  float v = 0.f;
  if (ck < kTaps) {
    const int c = ck / kLocK, k = ck - c * kLocK, jj = j + k - 15;
    if (jj >= 0 && jj < Te) v = c == 0 ? (t > 0 ? align[((long)b * T + t - 1) * Te + jj] : 0.f) : awc[((long)b * T + t) * Te + jj];
  }
  cols[i] = v;
}
__global__ void reduce_pe_x2_kernel(const float* __restrict__ PE, float* __restrict__ dx2, int B, int nsE) {
  const int b = blockIdx.x, c = threadIdx.x;   // step 0: g_x2 = sum of the partials
  if (b >= B) return;
  float g = 0.f;
  for (int s = 0; s < nsE; ++s) g += PE[((long)s * 64 + b) * kPEld + c];
  dx2[(long)b * 256 + c] = g;
}
__global__ void fill_kernel(float* p, float v, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void sum_rows_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // out[i] = sum_r x[r][i]  (small row counts)
  if (i >= n) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += x[(long)r * n + i];
  out[i] = s;
}
__global__ void prenet_dz_kernel(const float* g, const float* __restrict__ act, float* out, long n) {
  // dropout(p = 0.5) o relu backward: dz = 2 g where the (already masked) activation is positive   model.py:99
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = act[i] > 0.f ? 2.f * g[i] : 0.f;
}

struct BwdWs {
  float *dga, *dgd, *q, *dq, *awc, *pm, *dpm, *dctx, *dy, *gs, *cols, *pb, *pe, *gdc, *gac, *cacc, *gcat, *dv, *ones, *weff,
      *dweff, *tmp, *gproj, *weffT, *inv_scale;
  uint8_t *img_d, *img_a; DecoderCtrl* ctrl;
  void* wg; size_t wg_bytes;
};
size_t carve(char* base, int B, int Te, int T, BwdWs* w) {
  uintptr_t p = (uintptr_t)base;
  auto take = [&](size_t n) { float* r = (float*)p; p += (n * sizeof(float) + 255) & ~(size_t)255; return r; };
  const size_t TB = (size_t)T * B;
  BwdWs d;
  d.dga = take(TB * 4096); d.dgd = take(TB * 4096);
  d.q = take(TB * 128); d.dq = take(TB * 128);
  d.awc = take((size_t)B * T * Te);
  d.pm = take((size_t)B * Te * 128); d.dpm = take((size_t)B * Te * 128);
  d.dctx = take(TB * 512); d.dy = take(TB * 81);
  d.gs = take(TB * Te * 128); d.cols = take(TB * Te * kColsLd);
  d.pb = take((size_t)kSplitMax * 64 * kPBld); d.pe = take((size_t)kSplitMax * 64 * kPEld);
  d.gdc = take((size_t)64 * 1024); d.gac = take((size_t)64 * 1024);
  d.cacc = take((size_t)2 * B * Te); d.gcat = take((size_t)2 * 2 * B * 2 * Te);
  d.dv = take((size_t)B * 128);
  const size_t n_ones = TB > (size_t)B * Te ? TB : (size_t)B * Te;
  d.ones = take(n_ones);
  d.weff = take((size_t)kAtt * kTaps); d.dweff = take((size_t)kAtt * kColsLd);
  d.tmp = take(4096);
  d.gproj = take(TB * 1536); d.weffT = take((size_t)kAtt * kTaps);
  d.inv_scale = take(128);
  d.img_d = (uint8_t*)take(kBwdImgBytes / 4 + 256); d.img_a = (uint8_t*)take(kBwdImgBytes / 4 + 256);
  d.img_d = (uint8_t*)(((uintptr_t)d.img_d + 1023) & ~(uintptr_t)1023);
  d.img_a = (uint8_t*)(((uintptr_t)d.img_a + 1023) & ~(uintptr_t)1023);
  d.ctrl = (DecoderCtrl*)take(sizeof(DecoderCtrl) / 4 + 64);
  d.wg_bytes = wgrad_tc_ws_bytes(B, T);
  d.wg = take(d.wg_bytes / 4 + 64);
  if (w) *w = d;
  return (size_t)(p - (uintptr_t)base);
}

}  // namespace

size_t decoder_backward_ws_bytes(int B, int T_enc, int T_mel) { return carve(nullptr, B, T_enc, T_mel, nullptr) + 256; }

int decoder_backward(T2Model* m, const T2DecoderBwdArgs* a, cudaStream_t s) {
  const int B = a->B, Te = a->T_enc, T = a->T_mel;
  if (B < 1 || B > 64) return fail(T2_ERR_UNSUPPORTED, "decoder backward: 1 <= B <= 64 (got %d)", B);
  if (Te < 1 || T < 1) return fail(T2_ERR_INVALID, "decoder backward: bad sizes");
  const size_t smem = att_bwd_smem(Te);
  if (smem > 220 * 1024) return fail(T2_ERR_UNSUPPORTED, "decoder backward: T_enc = %d too long for the attention kernel", Te);
  if (a->n_grads != W_COUNT) return fail(T2_ERR_INVALID, "decoder backward: expected %d gradient pointers", (int)W_COUNT);
  if (a->ws_bytes < decoder_backward_ws_bytes(B, Te, T)) return fail(T2_ERR_WORKSPACE, "decoder backward workspace too small");
  if (a->stash_bytes < decoder_stash_bytes(B, T)) return fail(T2_ERR_WORKSPACE, "decoder stash too small");
  DecoderStash st;
  decoder_stash_carve(const_cast<void*>(a->stash), B, T, &st);
  BwdWs w;
  carve((char*)(((uintptr_t)a->ws + 255) & ~(uintptr_t)255), B, Te, T, &w);
  const size_t TB = (size_t)T * B;
  const int training = a->training;
  const float p_att = m->cfg.p_attention_dropout, p_dec = m->cfg.p_decoder_dropout;

  // ---- set-up -------------------------------------------------------------------------------------
  T2_CUDA(cudaMemsetAsync(w.gdc, 0, (size_t)64 * 1024 * 4, s));
  T2_CUDA(cudaMemsetAsync(w.gac, 0, (size_t)64 * 1024 * 4, s));
  T2_CUDA(cudaMemsetAsync(w.dv, 0, (size_t)B * 128 * 4, s));
  T2_CUDA(cudaMemsetAsync(w.cacc, 0, (size_t)2 * B * Te * 4, s));
  T2_CUDA(cudaMemsetAsync(w.gcat, 0, (size_t)8 * B * Te * 4, s));
  {
    const long n = (long)TB * 81;
    dy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a->d_mel, a->d_gate, w.dy, B, T);
    T2_LAUNCH_CHECK();
    awc_kernel<<<(B * Te + 127) / 128, 128, 0, s>>>(a->align, w.awc, B, T, Te);
    T2_LAUNCH_CHECK();
    weff_kernel<<<(kAtt * kTaps + 255) / 256, 256, 0, s>>>(m->w[W_ATT_LOC_DENSE], m->w[W_ATT_LOC_CONV], w.weff, w.weffT);
    T2_LAUNCH_CHECK();
    const size_t n_ones = TB > (size_t)B * Te ? TB : (size_t)B * Te;
    fill_kernel<<<(unsigned)((n_ones + 255) / 256), 256, 0, s>>>(w.ones, 1.f, (long)n_ones);
    T2_LAUNCH_CHECK();
  }
  // processed_memory (model.py:288) and the processed queries of all steps (model.py:57)
  T2_TRY(gemm_rm(m, s, false, true, B * Te, kAtt, kEnc, a->memory, kEnc, m->w[W_ATT_MEMORY], kEnc, w.pm, kAtt, 0.f));
  T2_TRY(gemm_rm(m, s, false, true, (int)TB, kAtt, kARnn, st.ha + (size_t)B * kARnn, kARnn, m->w[W_ATT_QUERY], kARnn, w.q, kAtt, 0.f));
  // projection / gate contribution to g_dh and g_ctx of every step: [d_mel_t ; d_gate_t] . [W_proj ; W_gate]  (model.py:373-378)
  T2_TRY(gemm_rm(m, s, false, false, (int)TB, kDRnn + kEnc, 81, w.dy, 81, m->projgate_w, kDRnn + kEnc, w.gproj, kDRnn + kEnc, 0.f));
  T2_CUDA(cudaFuncSetAttribute(att_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

  // T2_BWD_PROFILE=1: CUDA-event time per kernel class over the first 64 steps (stderr), for tuning
  const bool prof = getenv("T2_BWD_PROFILE") != nullptr;
  constexpr int kProfSteps = 64;
  static cudaEvent_t pev[kProfSteps][6];
  static bool pev_init = false;
  if (prof && !pev_init) {
    for (int i = 0; i < kProfSteps; ++i)
      for (int j = 0; j < 6; ++j) cudaEventCreate(&pev[i][j]);
    pev_init = true;
  }
  cudaEvent_t ph[4];
  if (prof) { for (int i = 0; i < 4; ++i) cudaEventCreate(&ph[i]); cudaEventRecord(ph[0], s); }
#define T2_TICK(i)                                                    \
  do {                                                                \
    if (prof && T - 1 - t < kProfSteps) cudaEventRecord(pev[T - 1 - t][i], s); \
  } while (0)
  // skinny GEMMs: tcgen05 split-fp16 engine (default) or the fp32 SIMT kernel (T2_BWD_GEMM=simt, cross-check)
  bool tc_gemm = true;
  {
    const char* e = getenv("T2_BWD_GEMM");
    if (e && e[0] == 's') tc_gemm = false;
  }
  const int nsB = tc_gemm ? kBwdGemmSplit : kSplitB, nsE = tc_gemm ? kBwdGemmSplit : kSplitE;
  if (tc_gemm) {
    T2_TRY(bwd_gemm_prepare(m, s));
    T2_CUDA(cudaMemsetAsync(w.img_d, 0, kBwdImgBytes, s));      // rows >= B stay zero
    T2_CUDA(cudaMemsetAsync(w.img_a, 0, kBwdImgBytes, s));
    T2_CUDA(cudaMemsetAsync(w.ctrl, 0, sizeof(DecoderCtrl), s));
    fill_kernel<<<1, 128, 0, s>>>(w.inv_scale, 1.f, 128);
    T2_LAUNCH_CHECK();
  }
  // ---- phase 1: reverse-time recurrence ---------------------------------------------------------------
  for (int t = T - 1; t >= 0; --t) {
    T2_TICK(0);
    const int carry = t < T - 1;
    {  // KA
      LstmBwdArgs k;
      memset(&k, 0, sizeof(k));
      k.src0 = carry ? w.pb : nullptr; k.ld0 = kPBld; k.off0 = 1536; k.ns0 = nsB;
      k.add = w.gproj + (size_t)t * B * 1536; k.ldadd = 1536;
      k.img = w.img_d; k.inv_scale = w.inv_scale;
      k.keep = a->dec_keep ? a->dec_keep + (size_t)t * B * kDRnn : nullptr;
      k.dropout = training; k.seed = a->seed; k.site = (uint32_t)(t * 4 + 3); k.p = p_dec;
      k.gates = st.gd + (size_t)t * B * 4096; k.c = st.cd + (size_t)(t + 1) * B * kDRnn; k.c_prev = st.cd + (size_t)t * B * kDRnn;
      k.g_c = w.gdc; k.dG = w.dgd + (size_t)t * B * 4096;
      if (tc_gemm) lstm_bwd_row_kernel<<<B, 1024, 0, s>>>(k);
      else lstm_bwd_kernel<<<dim3(4, B), 256, 0, s>>>(k);
      T2_LAUNCH_CHECK();
    }
    T2_TICK(1);
    if (tc_gemm) {
      T2_TRY(bwd_gemm_run(m, 0, w.img_d, w.inv_scale, w.pb, kPBld, w.ctrl, s));
    } else
    {  // KB
      SkinnyArgs k;
      k.A = w.dgd + (size_t)t * B * 4096; k.lda = 4096; k.rows = B; k.nred = 4096;
      k.W0 = m->w[W_DRNN_WIH]; k.ldw0 = kARnn + kEnc; k.cols0 = kARnn + kEnc;
      k.W1 = m->w[W_DRNN_WHH]; k.ldw1 = kDRnn; k.cols1 = kDRnn;
      k.P = w.pb; k.ldp = kPBld; k.nsplit = kSplitB;
      skinny_nn_kernel<<<dim3(kPBld / 128, kSplitB), 256, 0, s>>>(k);
      T2_LAUNCH_CHECK();
    }
    T2_TICK(2);
    {  // KC
      AttBwdArgs k;
      memset(&k, 0, sizeof(k));
      k.t = t; k.T = T; k.B = B; k.Te = Te; k.carry = carry; k.len = a->memory_lengths;
      k.memory = a->memory; k.pm = w.pm; k.q = w.q; k.v = m->w[W_ATT_V]; k.weff = w.weff;
      k.align = a->align; k.awc = w.awc; k.d_align = a->d_align;
      k.PE = w.pe; k.PB = w.pb; k.nsE = nsE; k.nsB = nsB; k.gproj = w.gproj; k.weffT = w.weffT;
      k.dctx = w.dctx; k.dx2 = a->d_prenet; k.dq = w.dq; k.gs = w.gs; k.gcat = w.gcat; k.cacc = w.cacc; k.dv = w.dv;
      att_bwd_kernel<<<dim3(2, B), kAttT, smem, s>>>(k);
      T2_LAUNCH_CHECK();
    }
    T2_TICK(3);
    {  // KD
      LstmBwdArgs k;
      memset(&k, 0, sizeof(k));
      k.src0 = carry ? w.pe : nullptr; k.ld0 = kPEld; k.off0 = 768; k.ns0 = nsE;
      k.src1 = w.pb; k.ld1 = kPBld; k.off1 = 0; k.ns1 = nsB;
      k.img = w.img_a; k.inv_scale = w.inv_scale + 64;
      k.vec = w.dq + (size_t)t * B * 128; k.nvec = 128; k.ldvec = 128; k.Wv = m->w[W_ATT_QUERY]; k.ldwv = kARnn;
      k.keep = a->att_keep ? a->att_keep + (size_t)t * B * kARnn : nullptr;
      k.dropout = training; k.seed = a->seed; k.site = (uint32_t)(t * 4 + 2); k.p = p_att;
      k.gates = st.ga + (size_t)t * B * 4096; k.c = st.ca + (size_t)(t + 1) * B * kARnn; k.c_prev = st.ca + (size_t)t * B * kARnn;
      k.g_c = w.gac; k.dG = w.dga + (size_t)t * B * 4096;
      if (tc_gemm) lstm_bwd_row_kernel<<<B, 1024, 0, s>>>(k);
      else lstm_bwd_kernel<<<dim3(4, B), 256, 0, s>>>(k);
      T2_LAUNCH_CHECK();
    }
    T2_TICK(4);
    if (tc_gemm) {
      T2_TRY(bwd_gemm_run(m, 1, w.img_a, w.inv_scale + 64, w.pe, kPEld, w.ctrl, s));
    } else
    {  // KE
      SkinnyArgs k;
      k.A = w.dga + (size_t)t * B * 4096; k.lda = 4096; k.rows = B; k.nred = 4096;
      k.W0 = m->w[W_ARNN_WIH]; k.ldw0 = kPre + kEnc; k.cols0 = kPre + kEnc;
      k.W1 = m->w[W_ARNN_WHH]; k.ldw1 = kARnn; k.cols1 = kARnn;
      k.P = w.pe; k.ldp = kPEld; k.nsplit = kSplitE;
      skinny_nn_kernel<<<dim3(kPEld / 128, kSplitE), 256, 0, s>>>(k);
      T2_LAUNCH_CHECK();
    }
    T2_TICK(5);
  }
  if (prof) cudaEventRecord(ph[1], s);
  reduce_pe_x2_kernel<<<B, 256, 0, s>>>(w.pe, a->d_prenet, B, nsE);
  T2_LAUNCH_CHECK();

  // ---- phase 2: time-batched gradients --------------------------------------------------------------------
  float* const* G = a->grads;
  const float* x2 = a->teacher_prenet;
  const int TBi = (int)TB;
  // LSTM weight / bias gradients: our tcgen05 split-fp16 engine (default) or plain cuBLAS fp32 GEMMs (T2_WGRAD=cublas)
  bool tc_wgrad = true;
  {
    const char* e = getenv("T2_WGRAD");
    if (e && e[0] == 'c') tc_wgrad = false;
  }
  if (tc_wgrad) {
    T2_TRY(wgrad_tc_run(m, B, T, w.dga, w.dgd, x2, st, G, w.wg, w.wg_bytes, s));
  } else {
  if (G[W_ARNN_WIH]) {   // [x2_t | ctx_{t-1}]                                             model.py:352
      T2_TRY(gemm_rm_wgrad(m, s, true, false, 4096, kPre, TBi, w.dga, 4096, x2, kPre, G[W_ARNN_WIH], kPre + kEnc, 0.f));
      T2_TRY(gemm_rm_wgrad(m, s, true, false, 4096, kEnc, TBi, w.dga, 4096, st.ctx, kEnc, G[W_ARNN_WIH] + kPre, kPre + kEnc, 0.f));
    }
    if (G[W_ARNN_WHH]) T2_TRY(gemm_rm_wgrad(m, s, true, false, 4096, kARnn, TBi, w.dga, 4096, st.ha, kARnn, G[W_ARNN_WHH], kARnn, 0.f));
    if (G[W_ARNN_BIH] || G[W_ARNN_BHH]) {
      T2_TRY(colsum_rm(m, s, w.dga, 4096, TBi, 4096, w.tmp));
      if (G[W_ARNN_BIH]) T2_CUDA(cudaMemcpyAsync(G[W_ARNN_BIH], w.tmp, 4096 * 4, cudaMemcpyDeviceToDevice, s));
      if (G[W_ARNN_BHH]) T2_CUDA(cudaMemcpyAsync(G[W_ARNN_BHH], w.tmp, 4096 * 4, cudaMemcpyDeviceToDevice, s));
    }
    if (G[W_DRNN_WIH]) {   // [ah_t | ctx_t]                                                 model.py:366-367
      T2_TRY(gemm_rm_wgrad(m, s, true, false, 4096, kARnn, TBi, w.dgd, 4096, st.ha + (size_t)B * kARnn, kARnn, G[W_DRNN_WIH], kARnn + kEnc, 0.f));
      T2_TRY(gemm_rm_wgrad(m, s, true, false, 4096, kEnc, TBi, w.dgd, 4096, st.ctx + (size_t)B * kEnc, kEnc, G[W_DRNN_WIH] + kARnn, kARnn + kEnc, 0.f));
    }
    if (G[W_DRNN_WHH]) T2_TRY(gemm_rm_wgrad(m, s, true, false, 4096, kDRnn, TBi, w.dgd, 4096, st.hd, kDRnn, G[W_DRNN_WHH], kDRnn, 0.f));
    if (G[W_DRNN_BIH] || G[W_DRNN_BHH]) {
      T2_TRY(colsum_rm(m, s, w.dgd, 4096, TBi, 4096, w.tmp));
      if (G[W_DRNN_BIH]) T2_CUDA(cudaMemcpyAsync(G[W_DRNN_BIH], w.tmp, 4096 * 4, cudaMemcpyDeviceToDevice, s));
      if (G[W_DRNN_BHH]) T2_CUDA(cudaMemcpyAsync(G[W_DRNN_BHH], w.tmp, 4096 * 4, cudaMemcpyDeviceToDevice, s));
    }
  }
  // projection + gate on [dh_t | ctx_t]                                                   model.py:373-378
  if (G[W_PROJ_W]) {
    T2_TRY(gemm_rm(m, s, true, false, kMel, kDRnn, TBi, w.dy, 81, st.hd + (size_t)B * kDRnn, kDRnn, G[W_PROJ_W], kDRnn + kEnc, 0.f));
    T2_TRY(gemm_rm(m, s, true, false, kMel, kEnc, TBi, w.dy, 81, st.ctx + (size_t)B * kEnc, kEnc, G[W_PROJ_W] + kDRnn, kDRnn + kEnc, 0.f));
  }
  if (G[W_GATE_W]) {
    T2_TRY(gemm_rm(m, s, true, false, 1, kDRnn, TBi, w.dy + 80, 81, st.hd + (size_t)B * kDRnn, kDRnn, G[W_GATE_W], kDRnn + kEnc, 0.f));
    T2_TRY(gemm_rm(m, s, true, false, 1, kEnc, TBi, w.dy + 80, 81, st.ctx + (size_t)B * kEnc, kEnc, G[W_GATE_W] + kDRnn, kDRnn + kEnc, 0.f));
  }
  if (G[W_PROJ_B] || G[W_GATE_B]) {
    T2_TRY(colsum_rm(m, s, w.dy, 81, TBi, 81, w.tmp));
    if (G[W_PROJ_B]) T2_CUDA(cudaMemcpyAsync(G[W_PROJ_B], w.tmp, 80 * 4, cudaMemcpyDeviceToDevice, s));
    if (G[W_GATE_B]) T2_CUDA(cudaMemcpyAsync(G[W_GATE_B], w.tmp + 80, 4, cudaMemcpyDeviceToDevice, s));
  }
  if (G[W_ATT_QUERY])
    T2_TRY(gemm_rm(m, s, true, false, kAtt, kARnn, TBi, w.dq, kAtt, st.ha + (size_t)B * kARnn, kARnn, G[W_ATT_QUERY], kARnn, 0.f));
  if (G[W_ATT_V]) {
    sum_rows_kernel<<<1, 128, 0, s>>>(w.dv, G[W_ATT_V], B, 128);
    T2_LAUNCH_CHECK();
  }
  if (G[W_ATT_LOC_CONV] || G[W_ATT_LOC_DENSE]) {   // through the fused filter Weff = W_ld . W_loc   model.py:23-25
    const long n = (long)TB * Te * kColsLd;
    im2col_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a->align, w.awc, w.cols, B, T, Te);
    T2_LAUNCH_CHECK();
    // K = T*B*Te can exceed what one call accumulates accurately in fp32 order-wise; split per step block
    const long R = (long)TB * Te;
    // cuBLAS cross-check: K = T*B*Te in blocks of 2^20 rows (one fp32 accumulation order per block); our GEMM drains its
    // accumulator every 64 rows and splits K over the SMs, so it takes the whole K in one call
    const long chunk = use_cublas_gemm() ? (1L << 20) : R;
    for (long r0 = 0; r0 < R; r0 += chunk) {
      const int k = (int)(R - r0 < chunk ? R - r0 : chunk);
      T2_TRY(gemm_rm(m, s, true, false, kAtt, kColsLd, k, w.gs + r0 * kAtt, kAtt, w.cols + r0 * kColsLd, kColsLd, w.dweff, kColsLd,
                     r0 ? 1.f : 0.f));
    }
    dweff_split_kernel<<<(kAtt * kLocF + 255) / 256, 256, 0, s>>>(w.dweff, m->w[W_ATT_LOC_DENSE], m->w[W_ATT_LOC_CONV],
                                                                  G[W_ATT_LOC_DENSE], G[W_ATT_LOC_CONV]);
    T2_LAUNCH_CHECK();
  }
  // d_pm = sum_t g_s[t]  -> memory_layer gradient and its share of d_memory              model.py:288
  T2_TRY(colsum_rm(m, s, w.gs, (long)B * Te * kAtt, T, B * Te * kAtt, w.dpm));
  if (G[W_ATT_MEMORY])
    T2_TRY(gemm_rm(m, s, true, false, kAtt, kEnc, B * Te, w.dpm, kAtt, a->memory, kEnc, G[W_ATT_MEMORY], kEnc, 0.f));
  if (a->d_memory) {
    T2_TRY(gemm_rm(m, s, false, false, B * Te, kEnc, kAtt, w.dpm, kAtt, m->w[W_ATT_MEMORY], kEnc, a->d_memory, kEnc, 0.f));
    // + sum_t aw_t[b] (x) g_ctx_t[b]: per row b, (Te x T) . (T x 512)                      model.py:83-84
    if (!use_cublas_gemm()) {
      GemmTc g;        // d_memory[b] (Te x 512) += align[b]^T (T x Te stored) . dctx[:, b, :] (T x 512, rows B * 512 apart)
      g.ta = true; g.tb = false; g.M = Te; g.N = kEnc; g.K = T;
      g.A = a->align; g.lda = Te; g.strideA = (long)T * Te;
      g.B = w.dctx; g.ldb = (long)B * kEnc; g.strideB = kEnc;
      g.C = a->d_memory; g.ldc = kEnc; g.strideC = (long)Te * kEnc; g.beta = 1.f; g.batch = B;
      T2_TRY(gemm_tc(m, s, g));
    } else {
      cublasHandle_t bl;
      T2_TRY(blas_handle(m, s, &bl));
      const float alpha = 1.f, beta = 1.f;
      cublasStatus_t stt = cublasSgemmStridedBatched(bl, CUBLAS_OP_N, CUBLAS_OP_T, kEnc, Te, T, &alpha, w.dctx, B * kEnc, (long long)kEnc,
                                                     a->align, Te, (long long)T * Te, &beta, a->d_memory, kEnc, (long long)Te * kEnc, B);
      if (stt != CUBLAS_STATUS_SUCCESS) return fail(T2_ERR_CUDA, "cublasSgemmStridedBatched failed (%d)", (int)stt);
      g_launch_count++;
    }
  }
  if (prof) {
    cudaEventRecord(ph[2], s);
    cudaStreamSynchronize(s);
    float acc[5] = {0, 0, 0, 0, 0};
    const int n = T < kProfSteps ? T : kProfSteps;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < 5; ++j) { float ms = 0; cudaEventElapsedTime(&ms, pev[i][j], pev[i][j + 1]); acc[j] += ms; }
    float m01 = 0, m12 = 0;
    cudaEventElapsedTime(&m01, ph[0], ph[1]);
    cudaEventElapsedTime(&m12, ph[1], ph[2]);
    fprintf(stderr, "[t2b200] decoder backward: loop %.2f ms (%d steps), batched gradients %.2f ms; per step us: lstm_dec %.1f "
            "gemm_dec %.1f attention %.1f lstm_att %.1f gemm_att %.1f\n", m01, T, m12, acc[0] / n * 1e3f, acc[1] / n * 1e3f,
            acc[2] / n * 1e3f, acc[3] / n * 1e3f, acc[4] / n * 1e3f);
    for (int i = 0; i < 4; ++i) cudaEventDestroy(ph[i]);
  }
  return T2_OK;
}

// ---------------------------------------------------------------------------------------------
// Prenet backward (model.py:97-100): recompute both layers with the same masks, then two plain GEMMs each.
// ---------------------------------------------------------------------------------------------
int prenet_backward(T2Model* m, const T2PrenetBwdArgs* a, cudaStream_t s) {
  const int M = a->M;
  if (a->n_grads != W_COUNT) return fail(T2_ERR_INVALID, "prenet backward: expected %d gradient pointers", (int)W_COUNT);
  if (a->ws_bytes < (size_t)4 * M * kPre * 4 + 1024) return fail(T2_ERR_WORKSPACE, "prenet backward workspace too small");
  float* x1 = (float*)(((uintptr_t)a->ws + 255) & ~(uintptr_t)255);
  float* x2 = x1 + (size_t)M * kPre;
  float* dz2 = x2 + (size_t)M * kPre;
  float* dz1 = dz2 + (size_t)M * kPre;
  GemmArgs g;
  g.seg[0] = {a->frames, kMel, m->w[W_PRENET0], kMel, kMel};
  g.M = M; g.N = kPre; g.C = x1; g.ldc = kPre; g.act = ACT_RELU; g.p_drop = 0.5f;
  if (a->keep) { g.keep = a->keep; g.ldkeep = kPre; } else { g.philox = 1; g.seed = a->seed; g.site = 0xA0; }
  T2_TRY(gemm_f32(g, s));
  GemmArgs h;
  h.seg[0] = {x1, kPre, m->w[W_PRENET1], kPre, kPre};
  h.M = M; h.N = kPre; h.C = x2; h.ldc = kPre; h.act = ACT_RELU; h.p_drop = 0.5f;
  if (a->keep) { h.keep = a->keep + (size_t)M * kPre; h.ldkeep = kPre; } else { h.philox = 1; h.seed = a->seed; h.site = 0xA1; }
  T2_TRY(gemm_f32(h, s));
  const long n = (long)M * kPre;
  prenet_dz_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a->d_out, x2, dz2, n);
  T2_LAUNCH_CHECK();
  if (a->grads[W_PRENET1]) T2_TRY(gemm_rm(m, s, true, false, kPre, kPre, M, dz2, kPre, x1, kPre, a->grads[W_PRENET1], kPre, 0.f));
  // d_x1 = dz2 . W2  -> through dropout o relu of layer 1
  T2_TRY(gemm_rm(m, s, false, false, M, kPre, kPre, dz2, kPre, m->w[W_PRENET1], kPre, dz1, kPre, 0.f));
  prenet_dz_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dz1, x1, dz1, n);
  T2_LAUNCH_CHECK();
  if (a->grads[W_PRENET0]) T2_TRY(gemm_rm(m, s, true, false, kPre, kMel, M, dz1, kPre, a->frames, kMel, a->grads[W_PRENET0], kMel, 0.f));
  return T2_OK;
}

void blas_destroy(T2Model* m) {
  if (m->blas) cublasDestroy((cublasHandle_t)m->blas);
  m->blas = nullptr;
}

}  // namespace t2
