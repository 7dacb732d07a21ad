// C-ABI plumbing: error reporting, device queries, and the thin extern "C" wrappers around the
// dense GEMM. Entry points for the recurrences / decoder / loss / optimiser live next to their
// kernels; all are declared in include/las_b200.h.
#include "common.cuh"
#include "las_internal.h"
#include "../../include/las_b200.h"
#include <stdarg.h>
#include <string.h>

namespace las {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;
unsigned long long g_path[LAS_PATH_COUNTERS] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error %d (%s) in %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return 1;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace las

extern "C" {

const char* las_last_error(void) { return las::g_err; }

int las_version(void) { return LAS_B200_VERSION; }

int las_num_sms(void) { return las::num_sms(); }

unsigned long long las_launch_count(void) { return las::g_launches; }

int las_path_counters(unsigned long long* out, int reset) {
  for (int i = 0; i < LAS_PATH_COUNTERS; ++i) {
    if (out) out[i] = las::g_path[i];
    if (reset) las::g_path[i] = 0;
  }
  return LAS_PATH_COUNTERS;
}

int las_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb,
                  int b_mn_major, void* C, int64_t ldc, int c_is_bf16, const float* bias, int M,
                  int N, int K, int relu, int accumulate, void* stream) {
  return las::gemm_bf16(A, lda, a_mn_major != 0, B, ldb, b_mn_major != 0, C, ldc, c_is_bf16 != 0,
                        bias, M, N, K, relu != 0, accumulate != 0,
                        static_cast<cudaStream_t>(stream));
}

int las_gemm_bf16_ws(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb,
                     int b_mn_major, void* C, int64_t ldc, int c_is_bf16, const float* bias, int M,
                     int N, int K, int relu, int accumulate, void* ws, int64_t ws_bytes, void* stream) {
  return las::gemm_bf16(A, lda, a_mn_major != 0, B, ldb, b_mn_major != 0, C, ldc, c_is_bf16 != 0,
                        bias, M, N, K, relu != 0, accumulate != 0,
                        static_cast<cudaStream_t>(stream), ws, ws_bytes);
}

}  // extern "C"
