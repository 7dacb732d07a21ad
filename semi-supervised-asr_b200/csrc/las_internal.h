// Internal (C++) interfaces shared between the translation units of liblas_b200.so.
// The public C-ABI is include/las_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace las {

int num_sms();

// D[m,n] = sum_k A[m,k] B[n,k] (+bias[n]) (relu) (+= C).  bf16 operands, f32 accumulate.
//   a_mn == false : A is row-major [M, K] with leading dimension lda
//   a_mn == true  : A is row-major [K, M] with leading dimension lda   (same for B / N)
int gemm_bf16(const void* A, int64_t lda, bool a_mn, const void* B, int64_t ldb, bool b_mn, void* C,
              int64_t ldc, bool c_bf16, const float* bias, int M, int N, int K, bool relu,
              bool accumulate, cudaStream_t stream);

}  // namespace las
