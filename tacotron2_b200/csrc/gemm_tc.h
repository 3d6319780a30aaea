// General fp32-grade GEMM / column sums of the training path on the tensor cores (gemm_tc.cu).
#pragma once
#include "model.h"

namespace t2 {

// row-major C (M x N, ldc) = op(A) . op(B) + beta C;  ta: A is stored (K x M, lda);  tb: B is stored (N x K, ldb).
// batch > 1: independent problems at element strides strideA / strideB / strideC (one set of operand scales).
struct GemmTc {
  bool ta = false, tb = false;
  int M = 0, N = 0, K = 0;
  const float* A = nullptr; long lda = 0;
  const float* B = nullptr; long ldb = 0;
  float* C = nullptr; long ldc = 0;
  float beta = 0.f;
  int batch = 1; long strideA = 0, strideB = 0, strideC = 0;
};
int gemm_tc(T2Model* m, cudaStream_t s, const GemmTc& g);
int gemm_tc_rm(T2Model* m, cudaStream_t s, bool ta, bool tb, int M, int N, int K, const float* A, long lda, const float* B,
               long ldb, float* C, long ldc, float beta);
// out[c] = sum over `rows` rows of X[r * ld + c]
int colsum_f32(T2Model* m, cudaStream_t s, const float* X, long ld, long rows, int cols, float* out);
void gemm_tc_destroy(T2Model* m);

}  // namespace t2
