/* liblas_b200 — C-ABI of the B200-native LAS (Listen-Attend-Spell) training-step kernels.
 *
 * This is the drop-in boundary below the reference's Python module surface
 * (jjery2243542/semi-supervised-ASR: model.py Encoder/AttLoc/Decoder/E2E/LM, solver.py train steps).
 * The reference has no FFI of its own; the functions below are what the
 * torch.autograd.Function wrappers in semi-supervised-asr_b200/functional.py bind through ctypes.
 *
 * Conventions
 *   - every function returns 0 on success; non-zero means failure and las_last_error() holds a
 *     message (thread-local). Nothing throws, nothing aborts.
 *   - all pointers are DEVICE pointers unless the name ends in _host.
 *   - the library never allocates, frees or synchronises: callers pass workspaces and the CUDA
 *     stream (cudaStream_t as void*) the work is enqueued on.
 *   - "bf16" buffers are uint16 bfloat16; "f32" are float; lengths are int32.
 *   - sm_100a only. There is no CPU path.
 */
#ifndef LAS_B200_H_
#define LAS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LAS_B200_VERSION 1

/* ------------------------------------------------------------------------------------------
 * core
 * ---------------------------------------------------------------------------------------- */
const char* las_last_error(void);
int las_version(void);
int las_num_sms(void);

/* Dense contraction D[m,n] = sum_k A[m,k]*B[n,k] (+bias[n]) (relu) (+=C) on tcgen05 tensor
 * cores, bf16 operands, f32 accumulation, operands fetched by TMA.
 * Replaces every nn.Linear / input-projection matmul on the path: model.py:70,93 (pyramid
 * projection), model.py:67,80 (LSTM input projection inside nn.LSTM), model.py:117,144 (mlp_enc),
 * model.py:263,293 (output layer), model.py:472,522 (LM output layer) and their autograd
 * backward GEMMs.
 *   a_mn_major == 0: A is row-major [M,K] (leading dim lda); == 1: A is row-major [K,M].
 *   b_mn_major likewise for B ([N,K] or [K,N]).
 *   c_is_bf16: output element type (0 = f32, 1 = bf16); bias may be NULL.
 *   lda/ldb must be multiples of 8 elements, bases 16-byte aligned (TMA constraints). */
int las_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb,
                  int b_mn_major, void* C, int64_t ldc, int c_is_bf16, const float* bias, int M,
                  int N, int K, int relu, int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LAS_B200_H_ */
