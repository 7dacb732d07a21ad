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
 *   - "bf16" buffers are uint16 bfloat16; "f16" IEEE half; "f32" float; lengths int32; tokens int64.
 *   - sm_100a only. There is no CPU path.
 */
#ifndef LAS_B200_H_
#define LAS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LAS_B200_VERSION 2

/* ------------------------------------------------------------------------------------------
 * core
 * ---------------------------------------------------------------------------------------- */
const char* las_last_error(void);
int las_version(void);
int las_num_sms(void);
/* kernels launched by this library since load (host counter; bench.py reports the per-step delta) */
unsigned long long las_launch_count(void);
/* Calls per code path since load (or the last reset): which implementation an entry point actually took. A test
 * (or a run) can assert that the cluster-persistent kernels ran instead of discovering a silent per-timestep path
 * as a 2x slower step. out: LAS_PATH_COUNTERS values (may be NULL); reset != 0 zeroes them. Returns the count. */
#define LAS_PATH_LSTM_PERSIST_FWD 0
#define LAS_PATH_LSTM_PERSIST_BWD 1
#define LAS_PATH_LSTM_STEP_FWD 2
#define LAS_PATH_LSTM_STEP_BWD 3
#define LAS_PATH_DEC_PERSIST_FWD 4
#define LAS_PATH_DEC_PERSIST_BWD 5
#define LAS_PATH_DEC_STEP_FWD 6
#define LAS_PATH_DEC_STEP_BWD 7
#define LAS_PATH_COUNTERS 8
int las_path_counters(unsigned long long* out, int reset);

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

/* Same contraction with a caller workspace (f32, ws_bytes): shapes with few output tiles and a long K (the
 * weight-gradient GEMMs: K = all frames of the batch) are split along K over the idle SMs, partial tiles go to
 * the workspace and are summed in a fixed order (deterministic). Falls back to the plain path otherwise. */
int las_gemm_bf16_ws(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb,
                     int b_mn_major, void* C, int64_t ldc, int c_is_bf16, const float* bias, int M,
                     int N, int K, int relu, int accumulate, void* ws, int64_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * element-wise / reduction helpers (HBM-bound)
 * ---------------------------------------------------------------------------------------- */
/* dst[r, c] = bf16(src[r, c]) for c < cols, 0 for cols <= c < ld_dst (utils.py:154 to_gpu + the
 * implicit fp32->compute-dtype cast). */
int las_cvt_pad_bf16(const float* src, int64_t ld_src, int64_t rows, int cols, void* dst,
                     int64_t ld_dst, void* stream);
int las_add2(const float* a, const float* b, float* out, int64_t n, void* stream);
/* Dropout (model.py:82, 95, 285, 512, 520; nn.LSTM inter-layer dropout), in place, on a strided [B, T(+rep_row), W]
 * view: kept elements are scaled by 1/(1-p). The mask is a pure function of (*seed_dev, site, element index)
 * (Philox4x32-10), so the backward pass applies the same call to the gradient instead of storing masks; seed_dev
 * is a device uint64 the caller advances once per training step (CUDA-graph replay safe). rep_row = 1: row T
 * (the replicated row of an odd pyramid extent) shares the mask of row T-1. */
int las_dropout(void* x, int x_is_bf16, int64_t B, int64_t T, int W, int64_t ld_b, int64_t ld_t, int rep_row, float p,
                const void* seed_dev, uint32_t site, void* stream);
/* ReLU backward of model.py:94: dz = dout * (out > 0), dz in bf16 */
int las_relu_bwd(const float* dout, const void* out, int out_is_bf16, void* dz, int64_t n, void* stream);
/* out[c] += sum_r x[r, c] (bias gradients) */
int las_colsum(const void* x, int x_is_bf16, int64_t ld, int64_t rows, int cols, float* out, void* stream);
/* nn.Embedding forward (model.py:261, 310, 465): out[i,:] = bf16(table[idx[i],:]), zero-padded to ld_out */
int las_gather_rows_bf16(const float* table, int dim, const int64_t* idx, int64_t n, void* out,
                         int64_t ld_out, void* stream);
/* nn.Embedding backward with padding_idx */
int las_scatter_add_rows(const float* d, int64_t ld, int dim, const int64_t* idx, int64_t n, int64_t pad,
                         float* dtable, void* stream);

/* ------------------------------------------------------------------------------------------
 * loss: log-softmax + gather + unigram label smoothing (model.py:353-366, 523-530)
 *   row r = (b, t) with b = r / rows_per_b, t = r % rows_per_b lives at logits + b*ld_b + t*ld.
 *   targets == NULL -> gather at the row's own argmax (free-running decode, model.py:360-361).
 *   out_prob / out_pred optional (LM.forward returns probabilities; Decoder returns argmax).
 * backward: dlogits (same addressing) from g_logp[r] (and g_prob[r], may be NULL).
 * ---------------------------------------------------------------------------------------- */
int las_ce_ls_fwd(const float* logits, int64_t ld, int64_t ld_b, int64_t rows_per_b, int64_t rows, int V,
                  const int64_t* targets, const float* dist, float ls, float* out_logp, float* out_prob,
                  int64_t* out_pred, void* stream);
int las_ce_ls_bwd(const float* logits, int64_t ld, int64_t ld_b, int64_t rows_per_b, int64_t rows, int V,
                  const int64_t* targets, const float* dist, float ls, const float* g_logp,
                  const float* g_prob, float* dlogits, void* stream);

/* ------------------------------------------------------------------------------------------
 * optimiser: global-norm clip + Adam / AMSGrad on flat f32 buffers
 * (torch.nn.utils.clip_grad_norm_ + torch.optim.Adam; solver.py:152-153, 171-173, 296-297, 384-385)
 * ---------------------------------------------------------------------------------------- */
int las_grad_norm(const float* g, int64_t n, void* partials_ws /* 256 doubles */, float* out_norm, void* stream);
/* step_dev: device int32 holding the 1-based step count (so a captured graph can be replayed);
 * vmax == NULL -> plain Adam. norm_ptr (device) + max_norm > 0 -> gradients are scaled by
 * min(1, max_norm / (norm + 1e-6)); grad_scale multiplies gradients first (DDP averaging). */
int las_adam_step(float* p, const float* g, float* m, float* v, float* vmax, int64_t n, float lr,
                  float beta1, float beta2, float eps, float weight_decay, const int32_t* step_dev,
                  float max_norm, const float* norm_ptr, float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * recurrences (model.py:67, 79-81 packed BLSTM; model.py:466, 515-517 LM LSTM)
 * ---------------------------------------------------------------------------------------- */
/* f32 weight matrix -> bf16 mma A-fragments. mode 0: rows in order; mode 1: LSTM gate-interleaved
 * (rows = 4H, logical row gate*H + unit; tile = 8 units x 2 gates); mode 2: tile = 4 units x 4 gates; mode 3: mode 2 with the K positions of each k-tile quad-permuted (persistent LSTM / decoder forward); mode 4: mode 0 rows with the same K permutation. transposed: logical A[r][c] = W[c*ld + col_offset + r]. */
int las_pack_afrag(const float* W, int64_t ld, int rows, int cols, int col_offset, int mode, int H,
                   int transposed, void* out, void* stream);
int64_t las_afrag_bytes(int rows, int cols, int mode, int H);
/* out[n, m] = sum_k A[m,k] v[n,k] (+bias[m]) (+add[n,m]) with pre-packed A (mode 0) */
int las_smallmm(const void* a_pk, int M, int K, const void* v, int v_is_f32, int64_t ldv, int N,
                const float* bias, const float* add, int64_t ld_add, float* out_f32, int64_t ld_out,
                void* out_bf16, int64_t ld_outb, void* stream);
/* One stand-alone LSTM cell step with explicit state (nn.LSTMCell; one timestep of nn.LSTM given (h0, c0): the
 * reference's LM.forward_step, model.py:535-542). gates = bias + W_hh h + W_ih x; c_state [B, H] f32 is updated in
 * place, h_out bf16 [B, ld_ho] written (must not alias h_in). whh_pk / wih_pk: las_pack_afrag mode 1; h_in bf16
 * [B, ld_h], x bf16 [B, ld_x] (columns finite up to the next multiple of 16 of H / Kx); bias f32 [4H]. */
int las_lstm_cell_step(const void* whh_pk, const void* wih_pk, const float* bias, const void* h_in, int64_t ld_h,
                       const void* x, int64_t ld_x, int Kx, float* c_state, void* h_out, int64_t ld_ho, int B, int H,
                       void* stream);
int64_t las_lstm_ws_bytes(int B, int H, int ndir);
int las_lstm_seq_fwd(const float* xproj, const void* whh_pk, const int32_t* lens, int B, int T, int H,
                     int ndir, void* y, int64_t y_ld_b, int64_t y_ld_t, int rep_row, void* hprev,
                     int64_t hp_ld_b, int64_t hp_ld_t, void* gates_save, float* c_save, void* ws,
                     void* stream);
int las_lstm_seq_bwd(const float* dy, int64_t dy_ld_b, int64_t dy_ld_t, int rep_row,
                     const void* whhT_pk, int whhT_layout, const int32_t* lens, int B, int T, int H,
                     int ndir, const void* gates_save, const float* c_save, void* dG, int64_t dg_ld_b,
                     int64_t dg_ld_t, void* ws, void* stream);
/* Cluster-persistent recurrence (one launch per layer and pass). las_lstm_persistent_geometry returns 1 (and
 * the backward kernel's cluster size / units per CTA) when hidden size H is served by it. Its operand layouts
 * differ from the per-timestep pair above so that every thread moves 8-16 bytes per access:
 *   xproj, dG  columns are gate-minor: [dir*4H + 4*unit + gate] (permute the rows of W_ih / the bias accordingly)
 *   whh_pk     las_pack_afrag mode 2 per direction; whhT_owner_pk from las_pack_whhT_owner
 *   rec        [ndir, B, T, H] records of 16 bytes: (i,f) f16x2 | (g,o) f16x2 | c f32 | tanh(c) f32
 * y / hprev / dy / lens / rep_row as in las_lstm_seq_{fwd,bwd}, except that the persistent forward writes EVERY row of y
 * (zeros past each length), so y need not be cleared by the caller. */
int las_lstm_persistent_geometry(int H, int* cs, int* upc);
/* development aid: resident-cluster capacity of the persistent LSTM kernels (which: 0 forward with the default 8-CTA
 * clusters, 1 backward, 2 / 3 forward with 7-CTA / 10-CTA clusters) */
int las_lstm_persist_max_clusters(int which, int H);
int las_lstm_persist_fwd(const float* xproj, const void* whh_pk, const int32_t* lens, int B, int T, int H,
                         int ndir, void* y, int64_t y_ld_b, int64_t y_ld_t, int rep_row, void* hprev,
                         int64_t hp_ld_b, int64_t hp_ld_t, void* rec, void* stream);
int las_lstm_persist_bwd(const float* dy, int64_t dy_ld_b, int64_t dy_ld_t, int rep_row,
                         const void* whhT_owner_pk, const int32_t* lens, int B, int T, int H, int ndir,
                         const void* rec, void* dG, int64_t dg_ld_b, int64_t dg_ld_t, void* stream);
/* switch between the cluster-persistent and the per-timestep kernels (returns the previous setting;
 * default on, or off with LAS_DISABLE_PERSISTENT=1 in the environment) */
int las_set_persistent(int on);
/* development aid: device buffer of 128 int64 receiving a clock64() phase trace of the persistent kernels */
int las_set_debug_buffer(void* dev_int64_x128);
int64_t las_whhT_owner_bytes(int H);
int las_pack_whhT_owner(const float* W_hh, int H, void* out, void* stream);
/* lens_out[b] = (lens_in[b] + 1) / sub   (model.py:92) */
int las_pyramid_lens(const int32_t* lens_in, int B, int sub, int32_t* lens_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * attention decoder (model.py:139-173 AttLoc.forward, 283-367 Decoder.forward_step / forward)
 *
 * All per-step buffers are [B, L+1, width]: row r holds what step r-1 produced, row 0 the
 * initial state (zeros; ws row 0 = the initial alignment from las_att_init).
 * ---------------------------------------------------------------------------------------- */
typedef struct las_dec_args {
  int32_t B, L, Te;          /* batch, decoder steps, encoder frames (padded extent)            */
  int32_t Hd, O, A, V, E, H; /* dec hidden, att_odim, att_dim, vocab, embedding, encoder dim    */
  int32_t C, K;              /* conv channels, conv_kernel_size (taps = 2K+1)                   */
  int32_t mode;              /* 0 teacher forcing, 1 greedy free-run, 2 smooth free-run         */
  float att_scaling;         /* AttLoc softmax scaling (2.0, model.py:139)                      */
  float smooth_scaling;      /* softmax temperature of the smooth embedding (model.py:341)      */
  int32_t denc_accumulate;
  int32_t _pad;
  /* inputs */
  const void* enc_h;         /* bf16 [B, Te, H]                                                 */
  const float* P;            /* f32 [B, Te, A] = mlp_enc(enc_h)                                 */
  const float* embx;         /* mode 0: f32 [B, L+1, 4Hd] = W_ih[:, :E] emb(ys_in) + b_ih + b_hh */
  const float* cell_bias;    /* mode 1/2: f32 [4Hd] = b_ih + b_hh                               */
  const void* wr_pk;         /* fragments (mode 1) of [W_hh | W_ih[:, E:]]  ([4Hd, Hd+O])       */
  const void* we_pk;         /* mode 1/2: fragments (mode 1) of W_ih[:, :E]                     */
  const void* mlp_dec_pk;    /* fragments (mode 0) of mlp_dec.weight [A, Hd]                    */
  const void* mlp_o_pk;      /* fragments of mlp_o.weight [O, H]                                */
  const float* mlp_o_b;      /* [O]                                                             */
  const void* out_pk;        /* mode 1/2: fragments of output_layer.weight [V, Hd+O]            */
  const float* out_b;        /* mode 1/2: [V]                                                   */
  const float* emb_w;        /* mode 1/2: embedding.weight [V, E]                               */
  const float* conv_w;       /* loc_conv.weight [C, 2K+1]                                       */
  const float* mlp_att;      /* mlp_att.weight [A, C]                                           */
  const float* gvec;         /* gvec.weight [A]                                                 */
  /* state / outputs (caller-zeroed unless noted) */
  float* ws;                 /* f32 [B, L+1, Te] alignments                                     */
  void* zc;                  /* bf16 [B, L+1, Hd+O] (+64 elements slack): [z_t | c_t]           */
  void* ctx;                 /* bf16 [B, L+1, H] attention context before mlp_o                 */
  float* c_state;            /* f32 [B, Hd] LSTM cell state                                     */
  void* emb_op;              /* mode 1/2: bf16 [B, L+1, Ep] step input embeddings (row 0 = BOS) */
  float* logits;             /* mode 1/2: f32 [B, L+1, V]                                       */
  int64_t* pred;             /* mode 1/2: [B, L]                                                */
  float* e_buf;              /* scratch f32 [B, Te]                                             */
  float* dzf;                /* f32 [B, L, A] mlp_dec(z_t)                                      */
  void* gates_save;          /* f16 [B, L, Hd, 4]                                               */
  float* c_save;             /* f32 [B, L, Hd]                                                  */
  /* backward only */
  const void* wrT_pk;        /* fragments (mode 0, transposed) of [W_hh | W_ih[:, E:]]          */
  const void* mlp_oT_pk;     /* fragments of mlp_o.weight^T                                     */
  const void* mlp_decT_pk;   /* fragments of mlp_dec.weight^T                                   */
  const float* dzc_all;      /* f32 [B, L+1, Hd+O]: dlogits @ output_layer.weight               */
  float* dcz_tot;            /* scratch f32 [B, Hd+O]                                           */
  void* dcz_all;             /* bf16 [B, L+1, Hd+O] total gradient of [z_t | c_t]               */
  float* dctx_all;           /* f32 [B, L, H]                                                   */
  float* dw_buf;             /* scratch f32 [B, Te]                                             */
  float* dattc_all;          /* f32 [L, B, Te, C]                                               */
  float* ddz_all;            /* f32 [B, L+1, A] (+64 slack), zeroed                             */
  float* dP;                 /* f32 [B, Te, A], zeroed                                          */
  float* att_part;           /* scratch f32, las_att_scratch_floats() elements                  */
  float* dc_state;           /* scratch f32 [B, Hd]                                             */
  void* dgates;              /* bf16 [B, L+1, 4Hd], zeroed (row L stays zero)                   */
  float* dmlp_att;           /* += [A, C]                                                       */
  float* dgvec;              /* += [A]                                                          */
  float* dconv_w;            /* += [C, 2K+1]                                                    */
  float* denc;               /* f32 [B, Te, H] gradient through the context (see denc_accumulate) */
  /* cluster-persistent decoder (teacher-forced mode; used when las_dec_persistent_supported() != 0) */
  const void* Q;             /* bf16 [B, Te, O] = enc_h @ mlp_o.weight^T, centred over the frames of each utterance */
  const void* wr2_pk;        /* fragments (mode 2: 4 units x 4 gates per tile) of [W_hh | W_ih[:, E:]] */
  float* cpre;               /* f32 [B, L, O]: c_t - cbias = sum_te w_t[te] Q[te]                */
  float* conv_save;          /* f32 [B, L, Te, 16]: location-conv features of every step        */
  const void* wrT2_pk;       /* backward: las_dec_persistent_pack(0, [W_hh | W_ih[:, E:]])      */
  const void* mlp_decT2_pk;  /* backward: las_dec_persistent_pack(1, mlp_dec.weight)            */
  float* de_all;             /* backward out: f32 [B, L, Te] energy gradients                   */
  float* dc_all;             /* backward out: f32 [B, L, O] total gradient of c_t               */
  const float* cbias;        /* f32 [B, O]: mlp_o.bias + the frame mean removed from Q          */
  /* backward of the free-running smooth mode (mode 2: the gradient of logit_t also arrives through the next
   * input embedding, so the output layer's backward runs inside the loop; dzc_all is then an output) */
  const void* weT_pk;        /* fragments (mode 0, transposed) of W_ih[:, :E]  ([Ep, 4Hd])      */
  const void* outT_pk;       /* fragments (mode 0, transposed) of output_layer.weight ([Hd+O, V]) */
  const float* dlogits;      /* f32 [B, L+1, V] loss gradient of the logits                     */
  float* dl_tot;             /* out: f32 [B, L+1, Vq] (+64 slack), Vq = V rounded up to 4, zeroed: total gradient of the logits */
  float* demb_buf;           /* scratch f32 [B, Ep]                                             */
  /* cell-input dropout (model.py:285): mask element (b*(L+1) + r)*O + o of site drop_site for c_{r-1} as input of
   * step r, element (b*(L+1) + r)*E + j of site drop_site + 1 for the step-r embedding (see las_dropout) */
  float drop_p;              /* 0 = off                                                         */
  uint32_t drop_site;
  const void* seed_dev;      /* device uint64                                                   */
  void* zcd;                 /* per-step path: bf16 [B, L+1, Hd+O] (+64 slack), zeroed: [z | drop(c)] */
  const float* pbar;         /* f32 [B, A]: frame mean removed from P (then P holds P - pbar and dzf = mlp_dec(z_t) + pbar) */
  /* las_dec_fwd, per-step path only: run steps [t_begin, t_end) of the L (t_end == 0: up to L). All state lives in
   * the caller's buffers (c_state, rows of zc / ws / emb_op), so a forward can be issued in chunks -- greedy decoding
   * that stops once every utterance has emitted <EOS> (Solver.test / validation, solver.py:212-286). */
  int32_t t_begin, t_end;
  const void* mlp_dec_pk_p;  /* persistent forward: fragments (mode 4: rows in order, quad-permuted K) of mlp_dec.weight; wr2_pk is mode 3 */
  /* Cluster-persistent GREEDY decoding (mode 1, inference: no dropout, no backward): set Q / wr2_pk / cbias / pbar /
   * mlp_dec_pk_p as for the teacher-forced persistent path, embx = the per-token input table f32 [V, 4Hd]
   * (W_ih[:, :E] emb(v) + b_ih + b_hh, gate-major columns), out_bf = output_layer.weight as bf16 [V, Hd+O], out_b,
   * logits, pred. The argmax of step t selects the table row of step t+1 inside the kernel. stop_token >= 0: a
   * cluster stops once each of its utterances has emitted that token (rows after the stop keep their initial
   * contents); -1: always L steps. */
  const void* out_bf;
  int32_t bos_token, stop_token;
  /* mode 1 on the per-step kernels only. Scheduled sampling (model.py:327-329): tok_teacher = ys_in int64 [B, L+1],
   * tf_mask uint8 [L+1]: step r consumes the teacher token where tf_mask[r] != 0, else the previous step's prediction.
   * sample != 0: the prediction is drawn from softmax(logits) (Categorical sampling, model.py:349-351; needs seed_dev)
   * instead of the argmax. */
  const int64_t* tok_teacher;
  const uint8_t* tf_mask;
  int32_t sample, _pad3;
} las_dec_args;

int las_att_init(const int32_t* enc_lens, int B, int Te, float* w, int64_t w_ld, void* stream);
/* 1 when las_dec_fwd / las_dec_bwd will run this problem (sizes, mode, Q and wr2_pk set) as ONE
 * cluster-persistent launch instead of per-step launches */
int las_dec_persistent_supported(const las_dec_args* args_host);
/* weight fragments of the cluster-persistent backward kernel (which: 0 = [W_hh | W_ih[:, E:]]^T from the
 * f32 [4Hd, Hd+O] matrix, 1 = mlp_dec.weight^T from the f32 [A, Hd] matrix) */
int64_t las_dec_persistent_pack_bytes(int which, int Hd, int O, int A);
int las_dec_persistent_pack(int which, const float* W, int64_t ld, int Hd, int O, int A, void* out, void* stream);
/* Post-loop reductions of the persistent backward (parallel over utterances / frames):
 *   dQ[b,te,o] = sum_t ws[b,t+1,te] dc[b,t,o]                       (las_att_dq)
 *   dP[b,te,a], dmlp_att[a,c] (+=), dgvec[a] (+=) from de, conv, dz, P with tanh recomputed (las_att_param_grads) */
int las_att_dq(const float* ws_alloc, const float* dc_all, int L, int B, int Te, int O, float* dQ, void* stream);
/* dconv_w[c,k] += sum_{t,b,te} dattc[t,b,te,c] * ws[b,t,te+k-K]   (loc_conv.weight gradient) */
int las_att_dconv(const float* dattc_all, const float* ws_alloc, int L, int B, int Te, int C, int K, float* dconv_w,
                  float* scratch, void* stream);
/* floats of scratch (las_dec_args.att_part, las_att_dconv, las_att_param_grads part_ws) for these sizes */
int64_t las_att_scratch_floats(int B, int L, int Te, int A, int C, int K);
/* 1 when the per-timestep decoder backward (las_dec_bwd without the cluster-persistent kernel) can run its lean
 * tensor-core energy backward for these sizes: the caller then passes las_dec_args.conv_save to las_dec_fwd (filled
 * by the per-timestep forward as well) and las_dec_args.de_all to las_dec_bwd, and obtains dP / dmlp_att / dgvec from
 * las_att_param_grads(_part) afterwards, exactly as after the persistent backward (model.py:156-165 backward) */
int las_att_bwd_lean_supported(int A, int C);
int las_att_param_grads(const float* P, const float* dzf, const float* conv_save, const float* de_all,
                        const float* mlp_att, const float* gvec, int B, int L, int Te, int A, int C, float* dP,
                        float* part_ws /* scratch: las_att_scratch_floats() */, float* dmlp_att,
                        float* dgvec, void* stream);
/* the same in two parts, so that only the part the encoder's backward waits for runs on the critical path:
 * what = 1 writes dP only, what = 2 adds the parameter sums into dmlp_att / dgvec only, what = 3 does both */
int las_att_param_grads_part(const float* P, const float* dzf, const float* conv_save, const float* de_all,
                             const float* mlp_att, const float* gvec, int B, int L, int Te, int A, int C, int what,
                             float* dP, float* part_ws, float* dmlp_att, float* dgvec, void* stream);
int las_dec_fwd(const las_dec_args* args_host, void* stream);
/* One stand-alone AttLoc.forward (model.py:139-173) on the per-step kernels, buffer conventions of las_dec_fwd:
 * reads the decoder state z from zc[:, t+1, :Hd] and the previous alignment from ws[:, t]; writes the alignment
 * ws[:, t+1], the context ctx[:, t+1] and c = mlp_o(context) into zc[:, t+1, Hd:]. Needs only the attention
 * operands (enc_h, P, mlp_dec_pk, mlp_o_pk, mlp_o_b, conv_w, mlp_att, gvec, e_buf, dzf). */
int las_att_step(const las_dec_args* args_host, int t, void* stream);
int las_dec_bwd(const las_dec_args* args_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LAS_B200_H_ */
