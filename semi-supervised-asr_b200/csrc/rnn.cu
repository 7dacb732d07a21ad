// LSTM cell steps and small batched mat-vecs on warp-level bf16 MMA. Weights are pre-packed in
// mma.m16n8k16 A-fragment order (one 128-bit load per lane per 16x16 tile), accumulation and the
// cell state are f32.
//
//  * pack_afrag_kernel   : f32 weight matrix -> bf16 A fragments (plain or LSTM gate-interleaved
//                          row order, optionally from the transposed source)
//  * cell_fwd_kernel     : one timestep of an LSTM layer (both directions of a BLSTM in one launch),
//                          per-sequence length predication, zero output past the length.
//                          Replaces nn.LSTM over a PackedSequence (model.py:79-81, 515-517) and
//                          nn.LSTMCell (model.py:286).
//  * cell_bwd_kernel     : the matching BPTT step: dh = A^T-frag MMA + dy (+extra), gate derivatives.
//  * smallmm_kernel      : out[n, m] = sum_k W[m,k] v[n,k] (+bias) (+add) for the per-step
//                          projections of the decoder (mlp_dec, mlp_o, output_layer and transposes).
//
// These are the per-timestep kernels: the host loops in this file (encoder / LM sequences) and in
// decoder.cu enqueue one launch per step on the caller's stream, which the Python side captures
// into a CUDA graph. The cluster-persistent BLSTM kernel lives in blstm_persistent.cu.
#include "common.cuh"
#include "las_internal.h"
#include "../../include/las_b200.h"

namespace las {

// ------------------------------------------------------------------------------------------
// A-fragment packing
// ------------------------------------------------------------------------------------------
// out[((tile*KT + kt)*32 + lane)*4 + j]: j&1 -> row +8, j>>1 -> col +8 (mma.m16n8k16 A layout).
// mode 0: tile rows are 16 consecutive logical rows.
// mode 2: tile = 4 hidden units x 4 gates: row rl -> gate rl>>2 of unit 4*tile + (rl&3).
// mode 1: LSTM gate interleave: tile = 2*ug + half; rows 0-7 -> gate 2*half of units 8ug..8ug+7,
//         rows 8-15 -> gate 2*half+1 of the same units (logical row = gate*H + unit).
// transposed != 0: logical A[row][col] = W[col][row].
__global__ void pack_afrag_kernel(const float* __restrict__ W, int64_t ld, int rows, int cols,
                                  int col_offset, int mode, int H, int transposed, int tiles, int KT,
                                  uint32_t* __restrict__ out) {
  const int64_t total = static_cast<int64_t>(tiles) * KT * 128;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int j = idx & 3;
    const int lane = (idx >> 2) & 31;
    const int64_t tk = idx >> 7;
    const int kt = tk % KT;
    const int tile = tk / KT;
    const int g = lane >> 2, tig = lane & 3;
    const int rl = g + 8 * (j & 1);
    int c0 = 16 * kt + 2 * tig + 8 * (j >> 1);
    int row;
    if (mode == 3 || mode == 4) {   // mode 3: mode 2 rows, mode 4: mode 0 rows; K positions of each k-tile permuted: quad q = units 4q..4q+3 at {2q, 2q+1, 2q+8, 2q+9}
      const int p0 = 2 * tig + 8 * (j >> 1);
      c0 = 16 * kt + 4 * ((p0 & 7) >> 1) + 2 * (p0 >> 3);
    }
    if (mode == 0 || mode == 4) {
      row = 16 * tile + rl;
    } else if (mode == 2 || mode == 3) {
      const int unit = 4 * tile + (rl & 3);
      const int gate = rl >> 2;
      row = (unit < H) ? gate * H + unit : rows;
    } else {
      const int ug = tile >> 1, half = tile & 1;
      const int unit = 8 * ug + (rl & 7);
      const int gate = 2 * half + (rl >> 3);
      row = (unit < H) ? gate * H + unit : rows;  // out of range -> zero
    }
    float v0 = 0.f, v1 = 0.f;
    if (row < rows) {
      if (!transposed) {
        if (c0 < cols) v0 = W[row * ld + col_offset + c0];
        if (c0 + 1 < cols) v1 = W[row * ld + col_offset + c0 + 1];
      } else {
        if (c0 < cols) v0 = W[static_cast<int64_t>(c0) * ld + col_offset + row];
        if (c0 + 1 < cols) v1 = W[static_cast<int64_t>(c0 + 1) * ld + col_offset + row];
      }
    }
    out[idx] = pack_bf16x2(v0, v1);
  }
}

int pack_afrag(const float* W, int64_t ld, int rows, int cols, int col_offset, int mode, int H,
               bool transposed, int tiles, int KT, uint32_t* out, cudaStream_t stream) {
  const int64_t total = static_cast<int64_t>(tiles) * KT * 128;
  if (total == 0) return 0;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
  pack_afrag_kernel<<<blocks, 256, 0, stream>>>(W, ld, rows, cols, col_offset, mode, H,
                                                transposed ? 1 : 0, tiles, KT, out); ++g_launches;
  LAS_LAUNCH_CHECK();
  return 0;
}

// B-operand fragment word: two consecutive K elements of row-major v (bf16, or f32 converted on
// the fly). `off` is an even element offset.
__device__ __forceinline__ uint32_t ldb_frag(const void* v, int is_f32, int64_t off) {
  if (is_f32) {
    const float2 f = *reinterpret_cast<const float2*>(static_cast<const float*>(v) + off);
    return pack_bf16x2(f.x, f.y);
  }
  return *reinterpret_cast<const uint32_t*>(static_cast<const __nv_bfloat16*>(v) + off);
}

// ------------------------------------------------------------------------------------------
// shared MMA machinery of the step kernels
// ------------------------------------------------------------------------------------------
// A CTA is (32 lanes) x NT batch tiles (8 rows each) x KS K-splits: warp = ks * NT + nt. Every
// warp accumulates its K-range with all loads of a chunk issued before the first MMA of the chunk
// (the step kernels are latency-bound: one L2 round trip per chunk instead of one per k-tile),
// K-splits > 0 park their partial sums in shared memory and split 0 finishes.
constexpr int kChunk = 8;
constexpr int kMaxWarps = 16;

// acc[m] += A[m-th tile][k-range] * v^T for NM row tiles sharing one B operand.
template <int NM>
__device__ __forceinline__ void mma_span(const uint4* __restrict__ a, int64_t tile_stride, const void* v,
                                         int v_f32, int64_t voff, int kbeg, int kend, float (&acc)[NM][4]) {
  for (int k0 = kbeg; k0 < kend; k0 += kChunk) {
    uint4 A[NM][kChunk];
    uint32_t B0[kChunk], B1[kChunk];
#pragma unroll
    for (int j = 0; j < kChunk; ++j) {
      if (k0 + j < kend) {
#pragma unroll
        for (int m = 0; m < NM; ++m) A[m][j] = __ldg(a + m * tile_stride + static_cast<int64_t>(k0 + j) * 32);
        B0[j] = ldb_frag(v, v_f32, voff + 16 * (k0 + j));
        B1[j] = ldb_frag(v, v_f32, voff + 16 * (k0 + j) + 8);
      }
    }
#pragma unroll
    for (int j = 0; j < kChunk; ++j) {
      if (k0 + j < kend) {
#pragma unroll
        for (int m = 0; m < NM; ++m) {
          const uint32_t Af[4] = {A[m][j].x, A[m][j].y, A[m][j].z, A[m][j].w};
          mma_bf16_16816(acc[m], Af, B0[j], B1[j]);
        }
      }
    }
  }
}

// Sum the partial accumulators of the K-splits into split 0. Returns false for warps that are done.
template <int NM>
__device__ __forceinline__ bool ksplit_reduce(float (&acc)[NM][4], float* red, int nt, int ks, int NT, int KS,
                                              int lane) {
  if (KS == 1) return true;
  if (ks > 0) {
    float* r = red + ((static_cast<int64_t>(ks - 1) * NT + nt) * 32 + lane) * (NM * 4);
#pragma unroll
    for (int m = 0; m < NM; ++m)
#pragma unroll
      for (int c = 0; c < 4; ++c) r[m * 4 + c] = acc[m][c];
  }
  __syncthreads();
  if (ks > 0) return false;
  for (int k = 1; k < KS; ++k) {
    const float* r = red + ((static_cast<int64_t>(k - 1) * NT + nt) * 32 + lane) * (NM * 4);
#pragma unroll
    for (int m = 0; m < NM; ++m)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[m][c] += r[m * 4 + c];
  }
  return true;
}

struct WarpShape { int NT, KS; };
// batch tiles per CTA and K-splits for an operand of KT k-tiles
static inline WarpShape warp_shape(int N, int KT) {
  WarpShape w;
  w.NT = N >= 32 ? 4 : (N + 7) / 8;
  if (w.NT < 1) w.NT = 1;
  int ks = (KT + 9) / 10;
  const int cap = kMaxWarps / w.NT;
  if (ks > cap) ks = cap;
  if (ks > KT) ks = KT;
  if (ks < 1) ks = 1;
  w.KS = ks;
  return w;
}

// ------------------------------------------------------------------------------------------
// forward LSTM cell step
// ------------------------------------------------------------------------------------------
// grid (UG, ceil(B/(8*NT)), ndir); a CTA owns the 8 hidden units of unit-group blockIdx.x (all
// four gates -> the cell update is thread-local) for 8*NT batch rows.
__global__ void __launch_bounds__(512) cell_fwd_kernel(CellFwdParams p) {
  pdl_enter();
  __shared__ float red[(kMaxWarps - 1) * 32 * 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tig = lane & 3;
  const int nt = warp % p.NT, ks = warp / p.NT;
  const int ug = blockIdx.x;
  const int n0 = (blockIdx.y * p.NT + nt) * 8;
  const int dir = blockIdx.z;
  const int t = (dir == 0) ? p.step : (p.T - 1 - p.step);
  const int nb = min(n0 + g, p.B - 1);  // clamp the operand row; results of rows >= B are dropped

  float acc[2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;

  // the two operand segments form one K range of KT1 + KT2 tiles, split evenly over the KS warps
  const int KT = p.KT1 + p.KT2;
  const int per = (KT + p.KS - 1) / p.KS;
  const int kb = ks * per, ke = min(KT, kb + per);
  if (n0 < p.B) {
    if (kb < p.KT1) {
      const uint4* a = reinterpret_cast<const uint4*>(p.a1 + dir * p.a_dir) + static_cast<int64_t>(2 * ug) * p.KT1 * 32 + lane;
      mma_span<2>(a, static_cast<int64_t>(p.KT1) * 32, p.v1, 0, dir * p.v1_dir + nb * p.v1_ld + 2 * tig, kb,
                  min(ke, p.KT1), acc);
    }
    if (ke > p.KT1) {
      const uint4* a = reinterpret_cast<const uint4*>(p.a2) + static_cast<int64_t>(2 * ug) * p.KT2 * 32 + lane;
      mma_span<2>(a, static_cast<int64_t>(p.KT2) * 32, p.v2, 0, nb * p.v2_ld + 2 * tig, max(kb, p.KT1) - p.KT1,
                  ke - p.KT1, acc);
    }
  }
  if (!ksplit_reduce<2>(acc, red, nt, ks, p.NT, p.KS, lane)) return;
  const int u = 8 * ug + g;
  if (u >= p.H || n0 >= p.B) return;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int n = n0 + 2 * tig + e;
    if (n >= p.B) continue;
    const int len = p.lens ? p.lens[n] : p.T;
    const bool active = t < len;
    const int64_t st = (static_cast<int64_t>(dir) * p.B + n) * p.H + u;  // c_state index
    __nv_bfloat16 h_old = __float2bfloat16(0.f);
    if (p.v1) h_old = p.v1[dir * p.v1_dir + n * p.v1_ld + u];
    __nv_bfloat16 h_new = h_old;
    __nv_bfloat16 y_val = __float2bfloat16(0.f);
    if (active) {
      const float* xp = p.xproj + n * p.xp_ld_b + t * p.xp_ld_t + dir * p.xp_ld_dir + u;
      const float gi = acc[0][e] + xp[0];
      const float gf = acc[0][2 + e] + xp[p.H];
      const float gg = acc[1][e] + xp[2 * p.H];
      const float go = acc[1][2 + e] + xp[3 * p.H];
      const float i = sigmoid_acc(gi), f = sigmoid_acc(gf), gc = tanh_acc(gg), o = sigmoid_acc(go);
      const float c = f * p.c_state[st] + i * gc;
      const float h = o * tanh_acc(c);
      p.c_state[st] = c;
      h_new = __float2bfloat16(h);
      y_val = h_new;
      const int64_t sv = ((static_cast<int64_t>(dir) * p.B + n) * p.T + t) * p.H + u;
      if (p.gates_save) {
        __half2 lo = __floats2half2_rn(i, f), hi = __floats2half2_rn(gc, o);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        reinterpret_cast<uint2*>(p.gates_save)[sv] = pk;
      }
      if (p.c_save) p.c_save[sv] = c;
    }
    p.hout[dir * p.hout_dir + n * p.hout_ld + u] = h_new;
    if (p.y) {
      const int64_t yo = n * p.y_ld_b + t * p.y_ld_t + static_cast<int64_t>(dir) * p.H + u;
      p.y[yo] = y_val;
      if (p.rep_row && t == p.T - 1) p.y[yo + p.y_ld_t] = y_val;
      if (p.hprev) p.hprev[n * p.hp_ld_b + t * p.hp_ld_t + static_cast<int64_t>(dir) * p.H + u] = h_old;
    }
  }
}

int launch_cell_fwd(CellFwdParams& p, cudaStream_t stream, bool pdl) {
  const WarpShape w = warp_shape(p.B, p.KT1 + p.KT2);
  p.NT = w.NT; p.KS = w.KS;
  dim3 grid(p.UG, (p.B + 8 * w.NT - 1) / (8 * w.NT), p.ndir);
  LAS_CUDA(launch_k(cell_fwd_kernel, grid, dim3(32 * w.NT * w.KS), 0, stream, pdl, p)); ++g_launches;
  return 0;
}

// ------------------------------------------------------------------------------------------
// backward LSTM cell step
// ------------------------------------------------------------------------------------------
// grid (ceil(H/16), ceil(B/(8*NT)), ndir). dh[n, j] = sum_k A[j, k] v[n, k] + dy + extra;
// A = W_hh^T (encoder/LM; v = gate gradients of the step that consumed h_t) or mlp_dec^T
// (decoder; v = gradient of the attention's decoder-state projection).
__global__ void __launch_bounds__(512) cell_bwd_kernel(CellBwdParams p) {
  pdl_enter();
  __shared__ float red[(kMaxWarps - 1) * 32 * 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tig = lane & 3;
  const int nt = warp % p.NT, ks = warp / p.NT;
  const int jt = blockIdx.x;
  const int n0 = (blockIdx.y * p.NT + nt) * 8;
  const int dir = blockIdx.z;
  // BPTT order: forward direction walks t = T-1..0, reverse direction walks t = 0..T-1.
  const int t = (dir == 0) ? (p.T - 1 - p.step) : p.step;

  float acc[1][4] = {{0.f, 0.f, 0.f, 0.f}};
  if (p.v != nullptr && p.KT > 0 && n0 < p.B) {
    const uint4* aT = reinterpret_cast<const uint4*>(p.a_pk + dir * p.a_dir) + static_cast<int64_t>(jt) * p.KT * 32 + lane;
    const int nb = min(n0 + g, p.B - 1);
    const int t_src = (dir == 0) ? p.v_t_fwd : p.v_t_rev;  // the step that consumed h_t
    const int64_t off = dir * p.v_dir + nb * p.v_ld + t_src * p.v_ld_t + 2 * tig;
    const int per = (p.KT + p.KS - 1) / p.KS;
    const int kb = ks * per, ke = min(p.KT, kb + per);
    mma_span<1>(aT, 0, p.v, p.v_f32, off, kb, ke, acc);
  }
  if (!ksplit_reduce<1>(acc, red, nt, ks, p.NT, p.KS, lane)) return;
  if (n0 >= p.B) return;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int j = 16 * jt + g + 8 * (e >> 1);
    const int n = n0 + 2 * tig + (e & 1);
    if (j >= p.H || n >= p.B) continue;
    const int len = p.lens ? p.lens[n] : p.T;
    const bool active = t < len;
    __nv_bfloat16* dg = p.dG + n * p.dg_ld_b + t * p.dg_ld_t + static_cast<int64_t>(dir) * 4 * p.H + j;
    if (!active) {
      const __nv_bfloat16 z = __float2bfloat16(0.f);
      dg[0] = z; dg[p.H] = z; dg[2 * p.H] = z; dg[3 * p.H] = z;
      continue;
    }
    const int64_t st = (static_cast<int64_t>(dir) * p.B + n) * p.H + j;
    const int64_t sv = ((static_cast<int64_t>(dir) * p.B + n) * p.T + t) * p.H + j;
    const uint2 pk = reinterpret_cast<const uint2*>(p.gates_save)[sv];
    const float2 if_ = __half22float2(*reinterpret_cast<const __half2*>(&pk.x));
    const float2 go_ = __half22float2(*reinterpret_cast<const __half2*>(&pk.y));
    const float i = if_.x, f = if_.y, gc = go_.x, o = go_.y;
    const float c = p.c_save[sv];
    float c_prev = 0.f;
    if (dir == 0) { if (t > 0) c_prev = p.c_save[sv - p.H]; }
    else          { if (t + 1 < len) c_prev = p.c_save[sv + p.H]; }
    float dh = acc[0][e];
    if (p.dy) {
      const float* dyp = p.dy + n * p.dy_ld_b + t * p.dy_ld_t + static_cast<int64_t>(dir) * p.H + j;
      dh += dyp[0];
      if (p.rep_row && t == p.T - 1) dh += dyp[p.dy_ld_t];
    }
    if (p.dh_extra) dh += p.dh_extra[n * p.dhx_ld + j];
    const float tc = tanh_acc(c);
    const float dc = dh * o * (1.f - tc * tc) + p.dc_state[st];
    p.dc_state[st] = dc * f;
    dg[0] = __float2bfloat16(dc * gc * i * (1.f - i));
    dg[p.H] = __float2bfloat16(dc * c_prev * f * (1.f - f));
    dg[2 * p.H] = __float2bfloat16(dc * i * (1.f - gc * gc));
    dg[3 * p.H] = __float2bfloat16(dh * tc * o * (1.f - o));
  }
}

int launch_cell_bwd(CellBwdParams& p, cudaStream_t stream, bool pdl) {
  const WarpShape w = warp_shape(p.B, p.v ? p.KT : 1);
  p.NT = w.NT; p.KS = w.KS;
  dim3 grid((p.H + 15) / 16, (p.B + 8 * w.NT - 1) / (8 * w.NT), p.ndir);
  LAS_CUDA(launch_k(cell_bwd_kernel, grid, dim3(32 * w.NT * w.KS), 0, stream, pdl, p)); ++g_launches;
  return 0;
}

// ------------------------------------------------------------------------------------------
// small batched mat-vec: out[n, m] = sum_k A[m, k] * v[n, k] (+ bias[m]) (+ add[n, m])
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void smallmm_body(const SmallMMParams& p, int mt, float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tig = lane & 3;
  const int nt = warp % p.NT, ks = warp / p.NT;
  const int n0 = (blockIdx.y * p.NT + nt) * 8;
  float acc[1][4] = {{0.f, 0.f, 0.f, 0.f}};
  if (n0 < p.N) {
    const uint4* a = reinterpret_cast<const uint4*>(p.a_pk) + static_cast<int64_t>(mt) * p.KT * 32 + lane;
    const int64_t off = static_cast<int64_t>(min(n0 + g, p.N - 1)) * p.ldv + 2 * tig;
    const int per = (p.KT + p.KS - 1) / p.KS;
    const int kb = ks * per, ke = min(p.KT, kb + per);
    mma_span<1>(a, 0, p.v, p.v_f32, off, kb, ke, acc);
  }
  if (!ksplit_reduce<1>(acc, red, nt, ks, p.NT, p.KS, lane)) return;
  if (n0 >= p.N) return;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int m = 16 * mt + g + 8 * (e >> 1);
    const int n = n0 + 2 * tig + (e & 1);
    if (m >= p.M || n >= p.N) continue;
    float r = acc[0][e];
    if (p.bias) r += p.bias[m];
    if (p.add) r += p.add[n * p.ld_add + m];
    if (p.out_f32) p.out_f32[n * p.ld_out + m] = r;
    if (p.out_bf16) p.out_bf16[n * p.ld_outb + m] = __float2bfloat16(r);
  }
}

__global__ void __launch_bounds__(512) smallmm_kernel(SmallMMParams p) {
  __shared__ float red[(kMaxWarps - 1) * 32 * 4];
  pdl_enter();
  smallmm_body(p, blockIdx.x, red);
}

// Two problems with the same N and K (hence the same warp shape) in one launch: CTAs [0, p0.MT) serve the first,
// the rest the second. The per-timestep decoder backward multiplies dgates_{t+1} by two weight matrices.
__global__ void __launch_bounds__(512) smallmm_pair_kernel(SmallMMParams p0, SmallMMParams p1) {
  __shared__ float red[(kMaxWarps - 1) * 32 * 4];
  pdl_enter();
  if (static_cast<int>(blockIdx.x) < p0.MT) smallmm_body(p0, blockIdx.x, red);
  else smallmm_body(p1, blockIdx.x - p0.MT, red);
}

static SmallMMParams smallmm_params(const uint32_t* a_pk, int M, int K, const void* v, int v_f32, int64_t ldv, int N,
                                    const float* bias, const float* add, int64_t ld_add, float* out_f32, int64_t ld_out,
                                    __nv_bfloat16* out_bf16, int64_t ld_outb) {
  SmallMMParams p;
  p.a_pk = a_pk; p.v = v; p.v_f32 = v_f32; p.ldv = ldv; p.bias = bias; p.add = add; p.ld_add = ld_add;
  p.out_f32 = out_f32; p.ld_out = ld_out; p.out_bf16 = out_bf16; p.ld_outb = ld_outb;
  p.M = M; p.N = N; p.MT = (M + 15) / 16; p.KT = (K + 15) / 16;
  const WarpShape w = warp_shape(N, p.KT);
  p.NT = w.NT; p.KS = w.KS;
  return p;
}

int smallmm(const uint32_t* a_pk, int M, int K, const void* v, int v_f32, int64_t ldv, int N,
            const float* bias, const float* add, int64_t ld_add, float* out_f32, int64_t ld_out,
            __nv_bfloat16* out_bf16, int64_t ld_outb, cudaStream_t stream, bool pdl) {
  if (M == 0 || N == 0) return 0;
  const SmallMMParams p = smallmm_params(a_pk, M, K, v, v_f32, ldv, N, bias, add, ld_add, out_f32, ld_out, out_bf16, ld_outb);
  dim3 grid(p.MT, (N + 8 * p.NT - 1) / (8 * p.NT));
  LAS_CUDA(launch_k(smallmm_kernel, grid, dim3(32 * p.NT * p.KS), 0, stream, pdl, p)); ++g_launches;
  return 0;
}

// out0 = A0 v, out1 = A1 v (+ nothing else): both f32, the same operand rows v [N, ldv] (bf16) and the same K
int smallmm_pair(const uint32_t* a0_pk, int M0, float* out0, int64_t ld_out0, const uint32_t* a1_pk, int M1, float* out1,
                 int64_t ld_out1, int K, const void* v, int64_t ldv, int N, cudaStream_t stream, bool pdl) {
  if (N == 0) return 0;
  const SmallMMParams p0 = smallmm_params(a0_pk, M0, K, v, 0, ldv, N, nullptr, nullptr, 0, out0, ld_out0, nullptr, 0);
  const SmallMMParams p1 = smallmm_params(a1_pk, M1, K, v, 0, ldv, N, nullptr, nullptr, 0, out1, ld_out1, nullptr, 0);
  dim3 grid(p0.MT + p1.MT, (N + 8 * p0.NT - 1) / (8 * p0.NT));
  LAS_CUDA(launch_k(smallmm_pair_kernel, grid, dim3(32 * p0.NT * p0.KS), 0, stream, pdl, p0, p1)); ++g_launches;
  return 0;
}

}  // namespace las

// ==========================================================================================
// C-ABI
// ==========================================================================================
using namespace las;

extern "C" {

int las_pack_afrag(const float* W, int64_t ld, int rows, int cols, int col_offset, int mode, int H,
                   int transposed, void* out, void* stream) {
  const int tiles = (mode == 1) ? 2 * ((H + 7) / 8) : (mode == 2 || mode == 3) ? (H + 3) / 4 : (rows + 15) / 16;
  const int KT = (cols + 15) / 16;
  return pack_afrag(W, ld, rows, cols, col_offset, mode, H, transposed != 0, tiles, KT,
                    static_cast<uint32_t*>(out), static_cast<cudaStream_t>(stream));
}

int64_t las_afrag_bytes(int rows, int cols, int mode, int H) {
  const int64_t tiles = (mode == 1) ? 2 * ((H + 7) / 8) : (mode == 2 || mode == 3) ? (H + 3) / 4 : (rows + 15) / 16;
  const int64_t KT = (cols + 15) / 16;
  return tiles * KT * 128 * 4;
}

int las_smallmm(const void* a_pk, int M, int K, const void* v, int v_is_f32, int64_t ldv, int N,
                const float* bias, const float* add, int64_t ld_add, float* out_f32, int64_t ld_out,
                void* out_bf16, int64_t ld_outb, void* stream) {
  int rc = smallmm(static_cast<const uint32_t*>(a_pk), M, K, v, v_is_f32, ldv, N, bias, add, ld_add,
                   out_f32, ld_out, static_cast<__nv_bfloat16*>(out_bf16), ld_outb,
                   static_cast<cudaStream_t>(stream));
  if (rc) return rc;
  LAS_LAUNCH_CHECK();
  return 0;
}

int64_t las_lstm_ws_bytes(int B, int H, int ndir) {
  const int64_t Kp = (H + 15) / 16 * 16 + 16;
  // bf16 state ping-pong (2 buffers) + f32 cell state
  return 2 * static_cast<int64_t>(ndir) * B * Kp * 2 + static_cast<int64_t>(ndir) * B * H * 4 + 256;
}

// One LSTM layer (ndir = 1 or 2) over padded [B, T] sequences with per-sequence lengths.
//   xproj      f32 [B, T, ndir*4H]: W_ih x + b_ih + b_hh (gate order i,f,g,o per direction)
//   whh_pk     A fragments of W_hh per direction (las_pack_afrag mode 1), consecutive directions
//   y          bf16, y[b*y_ld_b + t*y_ld_t + dir*H + u]; caller pre-zeroes it (positions >= len[b]
//              stay exact zeros, the pad_packed_sequence semantics of model.py:81)
//   rep_row    1: the value written at t = T-1 is also written at row T (replicate pad of an odd
//              padded extent, model.py:88-89)
//   hprev      bf16, hprev[b*hp_ld_b + t*hp_ld_t + dir*H + u]: the state entering step t (for the W_hh gradient)
//   gates_save f16 [ndir, B, T, H, 4] post-activation (i,f,g,o); c_save f32 [ndir, B, T, H]
int las_lstm_seq_fwd(const float* xproj, const void* whh_pk, const int32_t* lens, int B, int T, int H,
                     int ndir, void* y, int64_t y_ld_b, int64_t y_ld_t, int rep_row, void* hprev,
                     int64_t hp_ld_b, int64_t hp_ld_t, void* gates_save, float* c_save, void* ws,
                     void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LAS_REQUIRE(H % 8 == 0, "lstm: hidden size %d must be a multiple of 8", H);
  LAS_REQUIRE(ndir == 1 || ndir == 2, "lstm: ndir must be 1 or 2");
  if (B == 0 || T == 0) return 0;
  const int KT = (H + 15) / 16;
  const int64_t Kp = KT * 16 + 16;
  CellFwdParams p = {};
  p.B = B; p.T = T; p.H = H; p.ndir = ndir; p.UG = H / 8;
  p.xproj = xproj; p.xp_ld_t = static_cast<int64_t>(ndir) * 4 * H; p.xp_ld_b = p.xp_ld_t * T;
  p.xp_ld_dir = 4 * H;
  p.a1 = static_cast<const uint32_t*>(whh_pk); p.KT1 = KT;
  p.a_dir = static_cast<int64_t>(2) * p.UG * KT * 128;
  p.a2 = nullptr; p.v2 = nullptr; p.KT2 = 0; p.v2_ld = 0;
  p.lens = lens;
  p.y = static_cast<__nv_bfloat16*>(y); p.y_ld_b = y_ld_b; p.y_ld_t = y_ld_t;
  p.hprev = static_cast<__nv_bfloat16*>(hprev); p.hp_ld_b = hp_ld_b; p.hp_ld_t = hp_ld_t;
  p.gates_save = static_cast<__half*>(gates_save); p.c_save = c_save;
  p.rep_row = rep_row;
  const size_t hbytes = static_cast<size_t>(ndir) * B * Kp * sizeof(__nv_bfloat16);
  __nv_bfloat16* hbuf = static_cast<__nv_bfloat16*>(ws);
  p.c_state = reinterpret_cast<float*>(static_cast<char*>(ws) + ((2 * hbytes + 255) / 256) * 256);
  LAS_CUDA(cudaMemsetAsync(ws, 0, static_cast<size_t>(las_lstm_ws_bytes(B, H, ndir)), stream));
  p.v1_ld = Kp; p.v1_dir = static_cast<int64_t>(B) * Kp;
  p.hout_ld = Kp; p.hout_dir = p.v1_dir;
  ++g_path[LAS_PATH_LSTM_STEP_FWD];
  for (int s = 0; s < T; ++s) {
    p.step = s;
    p.v1 = hbuf + static_cast<size_t>(s & 1) * ndir * B * Kp;
    p.hout = hbuf + static_cast<size_t>((s + 1) & 1) * ndir * B * Kp;
    launch_cell_fwd(p, stream, s > 0 && pdl_enabled());     // (the first step follows a memset: full stream order)
  }
  LAS_LAUNCH_CHECK();
  return 0;
}

// One stand-alone LSTM cell step with explicit state (nn.LSTMCell / one timestep of nn.LSTM with (h0, c0); the
// reference's LM.forward_step, model.py:535-542): gates = bias + W_hh h + W_ih x, c_state updated in place,
// h_out written. whh_pk / wih_pk: las_pack_afrag mode 1 (gate-interleaved). h_in bf16 [B, ld_h] (columns up to the
// next multiple of 16 of H finite), x bf16 [B, ld_x] (likewise for the input width Kx), bias f32 [4H] (b_ih + b_hh),
// c_state f32 [B, H], h_out bf16 [B, ld_ho] (must not alias h_in).
int las_lstm_cell_step(const void* whh_pk, const void* wih_pk, const float* bias, const void* h_in, int64_t ld_h,
                       const void* x, int64_t ld_x, int Kx, float* c_state, void* h_out, int64_t ld_ho, int B, int H,
                       void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LAS_REQUIRE(H % 8 == 0, "lstm cell: hidden size %d must be a multiple of 8", H);
  LAS_REQUIRE(h_in != h_out, "lstm cell: h_out must not alias h_in");
  if (B == 0) return 0;
  CellFwdParams p = {};
  p.B = B; p.T = 1; p.H = H; p.ndir = 1; p.UG = H / 8;
  p.xproj = bias; p.xp_ld_b = 0; p.xp_ld_t = 0; p.xp_ld_dir = 0;
  p.a1 = static_cast<const uint32_t*>(whh_pk); p.KT1 = (H + 15) / 16; p.a_dir = 0;
  p.v1 = static_cast<const __nv_bfloat16*>(h_in); p.v1_ld = ld_h; p.v1_dir = 0;
  p.a2 = static_cast<const uint32_t*>(wih_pk); p.KT2 = (Kx + 15) / 16;
  p.v2 = static_cast<const __nv_bfloat16*>(x); p.v2_ld = ld_x;
  p.lens = nullptr; p.c_state = c_state;
  p.hout = static_cast<__nv_bfloat16*>(h_out); p.hout_ld = ld_ho; p.hout_dir = 0;
  p.y = nullptr; p.hprev = nullptr; p.gates_save = nullptr; p.c_save = nullptr; p.rep_row = 0; p.step = 0;
  launch_cell_fwd(p, stream);
  LAS_LAUNCH_CHECK();
  return 0;
}

// BPTT through one LSTM layer. dy f32 indexed like y (may be NULL); dG bf16 out,
// dG[b*dg_ld_b + t*dg_ld_t + dir*4H + gate*H + u] (zeros past the length); whhT_pk = fragments of
// W_hh^T per direction: whhT_layout 0 = las_pack_afrag mode 0, transposed (per-timestep kernels);
// 1 = las_pack_whhT_owner (cluster-persistent kernel). ws: f32 [ndir, B, H] scratch.
int las_lstm_seq_bwd(const float* dy, int64_t dy_ld_b, int64_t dy_ld_t, int rep_row,
                     const void* whhT_pk, int whhT_layout, const int32_t* lens, int B, int T, int H,
                     int ndir, const void* gates_save, const float* c_save, void* dG, int64_t dg_ld_b,
                     int64_t dg_ld_t, void* ws, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LAS_REQUIRE(H % 8 == 0, "lstm: hidden size %d must be a multiple of 8", H);
  if (B == 0 || T == 0) return 0;
  LAS_REQUIRE(whhT_layout == 0, "lstm bwd: owner-ordered weights belong to las_lstm_persist_bwd");
  ++g_path[LAS_PATH_LSTM_STEP_BWD];
  CellBwdParams p = {};
  p.dy = dy; p.dy_ld_b = dy_ld_b; p.dy_ld_t = dy_ld_t; p.rep_row = rep_row;
  p.dh_extra = nullptr; p.dhx_ld = 0;
  p.a_pk = static_cast<const uint32_t*>(whhT_pk);
  p.KT = (4 * H + 15) / 16;
  p.a_dir = static_cast<int64_t>((H + 15) / 16) * p.KT * 128;
  p.lens = lens;
  p.gates_save = static_cast<const __half*>(gates_save); p.c_save = c_save;
  p.dG = static_cast<__nv_bfloat16*>(dG); p.dg_ld_b = dg_ld_b; p.dg_ld_t = dg_ld_t;
  p.dc_state = static_cast<float*>(ws);
  p.B = B; p.T = T; p.H = H; p.ndir = ndir;
  p.v_f32 = 0; p.v_ld = dg_ld_b; p.v_dir = 4 * H;
  LAS_CUDA(cudaMemsetAsync(ws, 0, static_cast<size_t>(ndir) * B * H * sizeof(float), stream));
  for (int s = 0; s < T; ++s) {
    p.step = s;
    // source row of the recurrent term: the step that consumed h_t (t+1 forward, t-1 reverse)
    p.v = (s == 0) ? nullptr : p.dG;
    p.v_t_fwd = (T - 1 - s) + 1;
    p.v_t_rev = s - 1;
    p.v_ld_t = dg_ld_t;
    launch_cell_bwd(p, stream, s > 0 && pdl_enabled());
  }
  LAS_LAUNCH_CHECK();
  return 0;
}

int las_lstm_persist_fwd(const float* xproj, const void* whh_pk, const int32_t* lens, int B, int T, int H,
                         int ndir, void* y, int64_t y_ld_b, int64_t y_ld_t, int rep_row, void* hprev,
                         int64_t hp_ld_b, int64_t hp_ld_t, void* rec, void* stream) {
  LAS_REQUIRE(persist_supported(H), "persistent LSTM: hidden size %d unsupported (or switched off)", H);
  LAS_REQUIRE(ndir == 1 || ndir == 2, "lstm: ndir must be 1 or 2");
  if (B == 0 || T == 0) return 0;
  ++g_path[LAS_PATH_LSTM_PERSIST_FWD];
  int rc = persist_lstm_fwd(xproj, whh_pk, lens, B, T, H, ndir, y, y_ld_b, y_ld_t, rep_row, hprev, hp_ld_b, hp_ld_t,
                            rec, static_cast<cudaStream_t>(stream));
  if (rc) return rc;
  LAS_LAUNCH_CHECK();
  return 0;
}

int las_lstm_persist_bwd(const float* dy, int64_t dy_ld_b, int64_t dy_ld_t, int rep_row,
                         const void* whhT_owner_pk, const int32_t* lens, int B, int T, int H, int ndir,
                         const void* rec, void* dG, int64_t dg_ld_b, int64_t dg_ld_t, void* stream) {
  LAS_REQUIRE(persist_supported(H), "persistent LSTM: hidden size %d unsupported (or switched off)", H);
  if (B == 0 || T == 0) return 0;
  ++g_path[LAS_PATH_LSTM_PERSIST_BWD];
  int rc = persist_lstm_bwd(dy, dy_ld_b, dy_ld_t, rep_row, whhT_owner_pk, lens, B, T, H, ndir, rec, dG, dg_ld_b,
                            dg_ld_t, static_cast<cudaStream_t>(stream));
  if (rc) return rc;
  LAS_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
