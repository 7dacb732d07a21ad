// Internal (C++) interfaces shared between the translation units of liblas_b200.so.
// The public C-ABI is include/las_b200.h.
#pragma once
#include <stdlib.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace las {

int num_sms();

// D[m,n] = sum_k A[m,k] B[n,k] (+bias[n]) (relu) (+= C).  bf16 operands, f32 accumulate.
//   a_mn == false : A is row-major [M, K] with leading dimension lda
//   a_mn == true  : A is row-major [K, M] with leading dimension lda   (same for B / N)
int gemm_bf16(const void* A, int64_t lda, bool a_mn, const void* B, int64_t ldb, bool b_mn, void* C,
              int64_t ldc, bool c_bf16, const float* bias, int M, int N, int K, bool relu,
              bool accumulate, cudaStream_t stream, void* ws = nullptr, int64_t ws_bytes = 0);

// cluster-persistent recurrence (blstm_persistent.cu)
int persist_supported(int H);
int persist_lstm_fwd(const float* xproj, const void* whh_pk, const int32_t* lens, int B, int T, int H, int ndir,
                     void* y, int64_t y_ld_b, int64_t y_ld_t, int rep_row, void* hprev, int64_t hp_ld_b,
                     int64_t hp_ld_t, void* rec, cudaStream_t stream);
int persist_lstm_bwd(const float* dy, int64_t dy_ld_b, int64_t dy_ld_t, int rep_row, const void* wT_owner_pk,
                     const int32_t* lens, int B, int T, int H, int ndir, const void* rec, void* dG,
                     int64_t dg_ld_b, int64_t dg_ld_t, cudaStream_t stream);

// cluster-persistent attention decoder (decoder_persistent.cu)
}  // namespace las
struct las_dec_args;
namespace las {
extern void* g_dbg_buf_shared;   // las_set_debug_buffer (development aid)
int dec_persist_supported(const las_dec_args* a);
int dec_persist_fwd(const las_dec_args* a, cudaStream_t stream);
int dec_persist_bwd(const las_dec_args* a, cudaStream_t stream);
int64_t dec_persist_pack_bytes(int which, int Hd, int O, int A);
int dec_persist_pack(int which, const float* W, int64_t ld, int Hd, int O, int A, void* out, cudaStream_t stream);

#ifdef __CUDACC__
}  // namespace las
#include <cuda_bf16.h>
#include <cuda_fp16.h>
namespace las {

// One LSTM cell step: gates = xproj + A1 v1 (+ A2 v2); see rnn.cu.
struct CellFwdParams {
  const float* xproj;                  // xproj[n*xp_ld_b + t*xp_ld_t + dir*xp_ld_dir + gate*H + u]
  int64_t xp_ld_b, xp_ld_t, xp_ld_dir;
  const uint32_t* a1;                  // gate-interleaved A fragments [ndir][2*UG][KT1][32][4]
  int64_t a_dir;                       // uint32 elements between directions of a1
  const __nv_bfloat16* v1;             // state operand: v1[dir*v1_dir + n*v1_ld + k]
  int64_t v1_ld, v1_dir;
  int KT1;
  const uint32_t* a2;                  // optional second operand segment (decoder free-run: embedding)
  const __nv_bfloat16* v2;
  int64_t v2_ld;
  int KT2;
  const int32_t* lens;                 // [B] or nullptr (all sequences run T steps)
  __nv_bfloat16* hout;                 // new state: hout[dir*hout_dir + n*hout_ld + u]
  int64_t hout_ld, hout_dir;
  float* c_state;                      // [ndir][B][H]
  __nv_bfloat16* y;                    // y[n*y_ld_b + t*y_ld_t + dir*H + u] or nullptr
  int64_t y_ld_b, y_ld_t;
  __nv_bfloat16* hprev;                // hprev[n*hp_ld_b + t*hp_ld_t + dir*H + u] (state entering step t) or nullptr
  int64_t hp_ld_b, hp_ld_t;
  __half* gates_save;                  // [ndir][B][T][H][4] or nullptr
  float* c_save;                       // [ndir][B][T][H] or nullptr
  int B, T, H, ndir, UG;
  int NT, KS;                          // warp layout, filled by launch_cell_fwd
  int step;                            // 0..T-1 (the reverse direction processes t = T-1-step)
  int rep_row;                         // 1: the value at t == T-1 is also written to row T
};

struct CellBwdParams {
  const float* dy;                     // dy[n*dy_ld_b + t*dy_ld_t + dir*H + u] or nullptr
  int64_t dy_ld_b, dy_ld_t;
  const float* dh_extra;               // optional extra gradient on h_t: dh_extra[n*dhx_ld + u]
  int64_t dhx_ld;
  const uint32_t* a_pk;                // A fragments (rows = hidden units) [ndir][JT][KT][32][4]
  int64_t a_dir;
  const void* v;                       // MMA operand v[dir*v_dir + n*v_ld + t_src*v_ld_t + k]; nullptr = none
  int v_f32;
  int64_t v_ld, v_dir, v_ld_t;
  int v_t_fwd, v_t_rev;
  int KT;
  const int32_t* lens;
  const __half* gates_save;
  const float* c_save;
  __nv_bfloat16* dG;                   // dG[n*dg_ld_b + t*dg_ld_t + dir*4H + gate*H + u]
  int64_t dg_ld_b, dg_ld_t;
  float* dc_state;                     // [ndir][B][H]
  int B, T, H, ndir;
  int NT, KS;                          // warp layout, filled by launch_cell_bwd
  int step;
  int rep_row;
};

struct SmallMMParams {
  const uint32_t* a_pk;                // [MT][KT][32][4]
  const void* v;                       // [N, ldv] bf16 (or f32 when v_f32), finite beyond K up to 16*KT
  int v_f32;
  int64_t ldv;
  const float* bias;                   // [M] or nullptr
  const float* add;                    // [N, ld_add] or nullptr
  int64_t ld_add;
  float* out_f32;                      // [N, ld_out] or nullptr
  int64_t ld_out;
  __nv_bfloat16* out_bf16;             // [N, ld_outb] or nullptr
  int64_t ld_outb;
  int M, N, MT, KT;
  int NT, KS;
};

// LAS_PDL=0 switches programmatic dependent launches off (every chain kernel then launches with full stream order)
inline bool pdl_enabled() {
  static const bool on = getenv("LAS_PDL") == nullptr || atoi(getenv("LAS_PDL")) != 0;
  return on;
}
// <<<grid, block, smem, stream>>> with, optionally, the programmatic-stream-serialization attribute (common.cuh: pdl_*)
template <typename K, typename... Args>
inline cudaError_t launch_k(K kern, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}
int launch_cell_fwd(CellFwdParams& p, cudaStream_t stream, bool pdl = false);
int launch_cell_bwd(CellBwdParams& p, cudaStream_t stream, bool pdl = false);
int pack_afrag(const float* W, int64_t ld, int rows, int cols, int col_offset, int mode, int H,
               bool transposed, int tiles, int KT, uint32_t* out, cudaStream_t stream);
// out[n, m] = sum_k A[m,k] v[n,k] (+bias[m]) (+add[n,m]); A pre-packed by pack_afrag (mode 0).
int smallmm(const uint32_t* a_pk, int M, int K, const void* v, int v_f32, int64_t ldv, int N,
            const float* bias, const float* add, int64_t ld_add, float* out_f32, int64_t ld_out,
            __nv_bfloat16* out_bf16, int64_t ld_outb, cudaStream_t stream, bool pdl = false);
// two products of the same bf16 operand rows in one launch: out0 = A0 v, out1 = A1 v (f32 outputs)
int smallmm_pair(const uint32_t* a0_pk, int M0, float* out0, int64_t ld_out0, const uint32_t* a1_pk, int M1, float* out1,
                 int64_t ld_out1, int K, const void* v, int64_t ldv, int N, cudaStream_t stream, bool pdl = false);
#endif

}  // namespace las
