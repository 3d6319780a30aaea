// Fused gradient clipping + Adam update (train.py:229-236: clip_grad_norm_ then torch.optim.Adam.step) as three
// multi-tensor launches over a chunk table: sum of squares -> global norm / clip coefficient -> in-place update of
// gradient (scaled, like clip_grad_norm_), exp_avg, exp_avg_sq and parameter.  fp32.
#include <math.h>

#include <vector>

#include "common.cuh"

namespace t2 {
namespace {

constexpr int kChunk = 32768;
struct Chunk { float* p; float* g; float* m; float* v; int n; };

__global__ void __launch_bounds__(256) sumsq_kernel(const Chunk* __restrict__ chunks, double* __restrict__ partial) {
  __shared__ float red[8];
  const Chunk c = chunks[blockIdx.x];
  float s = 0.f;
  for (int i = threadIdx.x; i < c.n; i += 256) { const float g = c.g[i]; s = fmaf(g, g, s); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int i = 0; i < 8; ++i) a += (double)red[i];
    partial[blockIdx.x] = a;
  }
}
__global__ void norm_kernel(const double* __restrict__ partial, int n, float max_norm, float* __restrict__ out_norm, float* __restrict__ coef) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double a = 0.0;
  for (int i = 0; i < n; ++i) a += partial[i];
  const float norm = (float)sqrt(a);
  *out_norm = norm;
  float c = 1.f;
  if (max_norm > 0.f) { c = max_norm / (norm + 1e-6f); if (c > 1.f) c = 1.f; }   // torch.nn.utils.clip_grad_norm_
  *coef = c;
}
__global__ void __launch_bounds__(256) adam_kernel(const Chunk* __restrict__ chunks, const float* __restrict__ coef, float step_size, float beta1,
                                                   float beta2, float omb1, float omb2, float eps, float wd, float bc2_sqrt) {
  const Chunk c = chunks[blockIdx.x];
  const float k = *coef;
  for (int i = threadIdx.x; i < c.n; i += 256) {
    float g = c.g[i] * k;
    c.g[i] = g;                                   // clip_grad_norm_ scales .grad in place
    const float p = c.p[i];
    g = fmaf(wd, p, g);                           // Adam's L2 weight decay
    const float m = fmaf(beta1, c.m[i], omb1 * g);          // omb = 1 - beta, formed in double on the host
    const float v = fmaf(beta2, c.v[i], omb2 * g * g);
    c.m[i] = m; c.v[i] = v;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    c.p[i] = p - step_size * (m / denom);
  }
}

}  // namespace
}  // namespace t2

extern "C" {

size_t t2_clip_adam_workspace_bytes(int64_t total_elements, int32_t n_tensors) {
  const size_t chunks = (size_t)(total_elements / t2::kChunk) + (size_t)n_tensors + 1;
  return chunks * (sizeof(t2::Chunk) + sizeof(double)) + 1024;
}

int t2_clip_adam_step(const T2AdamArgs* a, void* stream) {
  using namespace t2;
  if (!a || a->n <= 0 || !a->params || !a->grads || !a->exp_avg || !a->exp_avg_sq || !a->numel || !a->grad_norm || !a->ws)
    return fail(T2_ERR_INVALID, "clip_adam: null argument");
  if (a->step < 1) return fail(T2_ERR_INVALID, "clip_adam: step counts from 1");
  cudaStream_t s = (cudaStream_t)stream;
  std::vector<Chunk> chunks;
  int64_t total = 0;
  for (int t = 0; t < a->n; ++t) {
    total += a->numel[t];
    for (int64_t o = 0; o < a->numel[t]; o += kChunk) {
      Chunk c;
      c.p = a->params[t] + o; c.g = a->grads[t] + o; c.m = a->exp_avg[t] + o; c.v = a->exp_avg_sq[t] + o;
      c.n = (int)((a->numel[t] - o) < kChunk ? (a->numel[t] - o) : kChunk);
      chunks.push_back(c);
    }
  }
  if (a->ws_bytes < t2_clip_adam_workspace_bytes(total, a->n)) return fail(T2_ERR_WORKSPACE, "clip_adam workspace too small");
  const size_t nchunk = chunks.size();
  char* p = (char*)(((uintptr_t)a->ws + 255) & ~(uintptr_t)255);
  double* partial = (double*)p; p += ((nchunk * sizeof(double) + 255) & ~(size_t)255);
  float* coef = (float*)p; p += 256;
  Chunk* d_chunks = (Chunk*)p;
  T2_CUDA(cudaMemcpyAsync(d_chunks, chunks.data(), nchunk * sizeof(Chunk), cudaMemcpyHostToDevice, s));   // pageable: staged before return
  sumsq_kernel<<<(unsigned)nchunk, 256, 0, s>>>(d_chunks, partial);
  T2_LAUNCH_CHECK();
  norm_kernel<<<1, 32, 0, s>>>(partial, (int)nchunk, (float)a->max_norm, a->grad_norm, coef);
  T2_LAUNCH_CHECK();
  const double bc1 = 1.0 - pow(a->beta1, (double)a->step), bc2 = 1.0 - pow(a->beta2, (double)a->step);
  adam_kernel<<<(unsigned)nchunk, 256, 0, s>>>(d_chunks, coef, (float)(a->lr / bc1), (float)a->beta1, (float)a->beta2, (float)(1.0 - a->beta1),
                                               (float)(1.0 - a->beta2), (float)a->eps,
                                               (float)a->weight_decay, (float)sqrt(bc2));
  T2_LAUNCH_CHECK();
  return T2_OK;
}

}  // extern "C"
