// Attention decoder (model.py:114-173 AttLoc, 256-367 Decoder) as a stream of per-timestep
// kernels: LSTM cell step (rnn.cu), decoder-state projection, location-aware attention energy,
// softmax + context, output projections; and the matching backward recursion which recomputes
// tanh/softmax instead of storing the [B, Te, att_dim] attention state of every step.
//
// Row conventions: every per-step buffer is [B, L+1, width]; row r of zc/ctx/ws/logits holds the
// value produced by step r-1 (row 0 = initial state), so "the state entering step t" is row t and
// the shifted weight-gradient GEMMs after the loop need no special case at t = 0.
#include <stdlib.h>
#include "common.cuh"
#include "las_internal.h"
#include "../../include/las_b200.h"

namespace las {

#ifndef LAS_KTT
#define LAS_KTT 8
#endif
// encoder frames per CTA in the per-timestep attention kernels. 8 (was 32): at B = 32, Te = 125 a launch is 512 CTAs of
// 8 frames, four or more resident per SM, instead of 128 CTAs that each walk 32 frames with 10 warps per SM
constexpr int kTT = LAS_KTT;

__device__ __forceinline__ float block_reduce_sum(float v, float* red, int nwarps) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int k = 0; k < nwarps; ++k) t += red[k];
  return t;
}
__device__ __forceinline__ float block_reduce_max(float v, float* red, int nwarps) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = -INFINITY;
  for (int k = 0; k < nwarps; ++k) t = fmaxf(t, red[k]);
  return t;
}

// Sum 16 per-lane values across the warp with 16 shuffles (instead of 16 x 5): each halving step
// keeps half of the values and exchanges the other half with the partner lane. On return, lane l
// holds the warp total of value index ((l>>4)&1)*8 + ((l>>3)&1)*4 + ((l>>2)&1)*2 + ((l>>1)&1).
__device__ __forceinline__ float warp_multi_reduce16(const float (&v)[16], int lane) {
  float a[8], b[4], c[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float mine = h16 ? v[8 + i] : v[i], other = h16 ? v[i] : v[8 + i];
    a[i] = mine + __shfl_xor_sync(0xffffffffu, other, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float mine = h8 ? a[4 + i] : a[i], other = h8 ? a[i] : a[4 + i];
    b[i] = mine + __shfl_xor_sync(0xffffffffu, other, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float mine = h4 ? b[2 + i] : b[i], other = h4 ? b[i] : b[2 + i];
    c[i] = mine + __shfl_xor_sync(0xffffffffu, other, 4);
  }
  const float mine = h2 ? c[1] : c[0], other = h2 ? c[0] : c[1];
  float r = mine + __shfl_xor_sync(0xffffffffu, other, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

struct AttGeom {
  int B, Te, A, C, K;  // K = conv_kernel_size (taps = 2K+1)
  int H;               // encoder feature dim
};

// Shared prologue of the energy kernels: previous alignment (zero padded by K on both sides),
// conv weights, and the location conv for this CTA's frame tile:
//   conv[tl][c] = sum_k wprev[te + k - K] * cw[c][k]            (model.py:120, 156)
template <int CM>
__device__ __forceinline__ void conv_tile(const AttGeom& g, const float* __restrict__ wprev_row,
                                          const float* __restrict__ conv_w, int te0, float* wp,
                                          float* cw, float* conv) {
  const int ksz = 2 * g.K + 1;
  for (int i = threadIdx.x; i < g.Te + 2 * g.K; i += blockDim.x) {
    const int j = i - g.K;
    wp[i] = (j >= 0 && j < g.Te) ? wprev_row[j] : 0.f;
  }
  for (int i = threadIdx.x; i < g.C * ksz; i += blockDim.x) cw[i] = conv_w[i];
  __syncthreads();
  for (int i = threadIdx.x; i < kTT * CM; i += blockDim.x) {
    const int tl = i / CM, c = i % CM;      // padding channels c >= C are written as exact zeros (0 * garbage may be NaN)
    const int te = te0 + tl;
    float s = 0.f;
    if (te < g.Te && c < g.C) {
      const float* wrow = wp + te;  // wp[te + k] = wprev[te + k - K]
      const float* crow = cw + c * ksz;
      for (int k = 0; k < ksz; ++k) s = fmaf(wrow[k], crow[k], s);
    }
    conv[tl * CM + c] = s;
  }
  __syncthreads();
}

struct EnergyFwdParams {
  AttGeom g;
  const float* P;        // [B, Te, A] mlp_enc(enc_h) + bias
  const float* dz;       // [B, A] (row stride dz_ld) mlp_dec(z_t)
  int64_t dz_ld;
  const float* wprev;    // [B] rows of Te (row stride w_ld)
  int64_t w_ld;
  const float* conv_w;   // [C, 2K+1]
  const float* mlp_att;  // [A, C]
  const float* gvec;     // [A]
  float* e;              // [B, Te]
  float* conv_save;      // step slice of the [B, L, Te, 16] location-conv features (row stride cs_ld per utterance) or nullptr
  int64_t cs_ld;
};

// grid (ceil(Te/kTT), B), block = A rounded up to a warp multiple; thread a owns attention dim a.
template <int CM>
__global__ void __launch_bounds__(512) att_energy_fwd_kernel(EnergyFwdParams p) {
  extern __shared__ float sm[];
  pdl_enter();
  const AttGeom& g = p.g;
  const int ksz = 2 * g.K + 1;
  float* wp = sm;
  float* cw = wp + g.Te + 2 * g.K;
  float* conv = cw + g.C * ksz;
  float* red = conv + kTT * CM;  // [nwarps][kTT]
  const int b = blockIdx.y, te0 = blockIdx.x * kTT;
  const int a = threadIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  conv_tile<CM>(g, p.wprev + b * p.w_ld, p.conv_w, te0, wp, cw, conv);
  const int ntl = min(kTT, g.Te - te0);
  if (p.conv_save != nullptr) {
    // the backward (att_energy_bwd_mma_kernel, las_att_param_grads_part) reads the conv features instead of
    // recomputing the 2K+1-tap correlation: rows of 16 floats, channels >= C are zero
    float* cs = p.conv_save + b * p.cs_ld + static_cast<int64_t>(te0) * 16;
    for (int i = threadIdx.x; i < ntl * 16; i += blockDim.x) {
      const int tl = i >> 4, c = i & 15;
      cs[i] = c < CM ? conv[tl * CM + c] : 0.f;
    }
  }

  float matt[CM];
#pragma unroll
  for (int c = 0; c < CM; ++c) matt[c] = (a < g.A && c < g.C) ? p.mlp_att[a * g.C + c] : 0.f;
  const float dza = (a < g.A) ? p.dz[b * p.dz_ld + a] : 0.f;
  const float gv = (a < g.A) ? p.gvec[a] : 0.f;
  const float* Pb = p.P + (static_cast<int64_t>(b) * g.Te + te0) * g.A + a;
  // all P values of the tile first: one global load per frame inside the loop made every iteration wait for L2
  // (long_scoreboard 10.6 warps/issue, 13 % issue utilisation in the first version)
  float pv[kTT];
#pragma unroll
  for (int tl = 0; tl < kTT; ++tl) pv[tl] = (tl < ntl && a < g.A) ? __ldg(Pb + static_cast<int64_t>(tl) * g.A) : 0.f;
#pragma unroll
  for (int tl = 0; tl < kTT; ++tl) {
    if (tl < ntl) {        // (block-uniform)
      float x = dza + pv[tl];
#pragma unroll
      for (int c = 0; c < CM; ++c) x = fmaf(matt[c], conv[tl * CM + c], x);
      float v = gv * tanh_acc(x);
      v = warp_sum(v);
      if (lane == 0) red[warp * kTT + tl] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < ntl) {
    float s = 0.f;
    for (int w = 0; w < nwarps; ++w) s += red[w * kTT + threadIdx.x];
    p.e[b * g.Te + te0 + threadIdx.x] = s;
  }
}

struct CtxFwdParams {
  int B, Te, H;
  float scaling;
  const float* e;               // [B, Te]
  const __nv_bfloat16* enc_h;   // [B, Te, H]
  float* w_out;                 // row b at w_out + b*w_ld (Te floats)
  int64_t w_ld;
  __nv_bfloat16* ctx;           // row b at ctx + b*ctx_ld (H bf16)
  int64_t ctx_ld;
};

// grid (B, ceil(H/64)), block 256: softmax over ALL Te padded frames (no mask, model.py:167),
// then the context for a 64-column slice: thread = (column pair, frame group of 8).
__global__ void __launch_bounds__(256) att_ctx_fwd_kernel(CtxFwdParams p) {
  extern __shared__ float sm[];
  pdl_enter();
  float* w_s = sm;               // [Te]
  float* red = w_s + ((p.Te + 1) & ~1);  // [8]
  float2* part = reinterpret_cast<float2*>(red + 8);  // [8][32]
  const int b = blockIdx.x;
  const float* eb = p.e + static_cast<int64_t>(b) * p.Te;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < p.Te; i += 256) mx = fmaxf(mx, p.scaling * eb[i]);
  mx = block_reduce_max(mx, red, 8);
  float se = 0.f;
  for (int i = threadIdx.x; i < p.Te; i += 256) {
    const float v = __expf(p.scaling * eb[i] - mx);
    w_s[i] = v;
    se += v;
  }
  se = block_reduce_sum(se, red, 8);
  const float inv = 1.f / se;
  __syncthreads();
  for (int i = threadIdx.x; i < p.Te; i += 256) {
    const float v = w_s[i] * inv;
    w_s[i] = v;
    if (blockIdx.y == 0) p.w_out[b * p.w_ld + i] = v;
  }
  __syncthreads();
  const int cp = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int col = blockIdx.y * 64 + 2 * cp;
  float2 acc = make_float2(0.f, 0.f);
  if (col < p.H) {
    const __nv_bfloat16* eh = p.enc_h + static_cast<int64_t>(b) * p.Te * p.H + col;
    for (int te = grp; te < p.Te; te += 8) {
      const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(eh + static_cast<int64_t>(te) * p.H));
      acc.x = fmaf(w_s[te], v.x, acc.x);
      acc.y = fmaf(w_s[te], v.y, acc.y);
    }
  }
  part[grp * 32 + cp] = acc;
  __syncthreads();
  if (grp == 0 && col < p.H) {
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) { s.x += part[k * 32 + cp].x; s.y += part[k * 32 + cp].y; }
    *reinterpret_cast<uint32_t*>(p.ctx + b * p.ctx_ld + col) = pack_bf16x2(s.x, s.y);
  }
}

// Next input embedding in free-running modes (model.py:331-341): argmax token, then either its
// embedding row (greedy) or softmax(scaling * logit) @ E (smooth). One warp per utterance.
struct NextEmbParams {
  float drop_p; const unsigned long long* seed_dev; uint32_t site; int64_t R, row;   // cell-input dropout of the embedding
  int B, V, E;
  const float* logits; int64_t lg_ld;   // row b at logits + b*lg_ld
  const float* emb_w;                   // [V, E]
  float scaling; int smooth;
  int64_t* pred; int64_t pred_ld;       // pred[b*pred_ld]
  __nv_bfloat16* emb_out; int64_t eo_ld;
  // scheduled sampling (model.py:327-329: teacher token with probability tf_rate, else the previous prediction) and
  // Categorical sampling of the prediction (model.py:349-351); both only in the non-smooth branch
  // cell-input dropout of [z_t; c_t] -> zcd row (DropRowParams semantics; zc == nullptr: not done here)
  const __nv_bfloat16* zc; __nv_bfloat16* zcd; int Hd, O; uint32_t zc_site;
  const int64_t* teacher;               // [B, R] teacher tokens (ys_in) or nullptr
  const uint8_t* tf_mask;               // [R]: 1 = step `row` consumes the teacher token
  int sample; uint32_t sample_site;
  const unsigned long long* sample_seed;
};
// One CTA (4 warps) per utterance; dynamic shared memory: V floats (softmax numerators).
__global__ void __launch_bounds__(128) next_emb_kernel(NextEmbParams p) {
  extern __shared__ float ne_sm[];
  pdl_enter();
  __shared__ float s_mx, s_inv;
  __shared__ int s_amax;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x;
  const float* x = p.logits + b * p.lg_ld;
  if (p.zc != nullptr) {       // the other half of the next cell input (was a launch of its own: dec_drop_fwd_kernel)
    const int ZC = p.Hd + p.O;
    const unsigned long long sd = *p.seed_dev;
    const int64_t row_off = (static_cast<int64_t>(b) * p.R + p.row) * ZC;
    for (int j = threadIdx.x; j < ZC; j += 128) {
      __nv_bfloat16 v = p.zc[row_off + j];
      if (j >= p.Hd) {
        const bool keep = dropout_keep(sd, p.zc_site, (static_cast<unsigned long long>(b) * p.R + p.row) * p.O + (j - p.Hd), p.drop_p);
        v = keep ? __float2bfloat16(__bfloat162float(v) / (1.f - p.drop_p)) : __float2bfloat16(0.f);
      }
      p.zcd[row_off + j] = v;
    }
  }
  if (warp == 0) {
    float mx;
    int amax = warp_argmax(x, p.V, lane, &mx);
    if (lane == 0) {
      if (p.sample) {
        // Categorical(logits).sample(): inverse CDF of softmax(logits) at one counter-based uniform per (utterance, step)
        const unsigned long long sd = *p.sample_seed;
        const uint4 r = philox4x32_10(make_uint2(static_cast<uint32_t>(sd), static_cast<uint32_t>(sd >> 32)),
                                      make_uint4(static_cast<uint32_t>(b), static_cast<uint32_t>(p.row), p.sample_site, 0x53414d50u));
        const float u = (static_cast<float>(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
        float tot = 0.f;
        for (int v = 0; v < p.V; ++v) tot += __expf(x[v] - mx);
        float acc = 0.f;
        int pick = p.V - 1;
        for (int v = 0; v < p.V; ++v) {
          acc += __expf(x[v] - mx);
          if (acc >= u * tot) { pick = v; break; }
        }
        amax = pick;
      }
      p.pred[b * p.pred_ld] = amax;
      // the token the NEXT step consumes: the teacher's when scheduled sampling says so
      if (p.teacher && p.tf_mask[p.row]) amax = static_cast<int>(p.teacher[static_cast<int64_t>(b) * p.R + p.row]);
      s_mx = mx; s_amax = amax;
    }
  }
  __syncthreads();
  const float mx = s_mx;
  const int amax = s_amax;
  const unsigned long long seed = p.drop_p > 0.f ? *p.seed_dev : 0ull;
  auto drop = [&](float v, int j) {
    if (p.drop_p <= 0.f) return v;
    return dropout_keep(seed, p.site, (static_cast<unsigned long long>(b) * p.R + p.row) * p.E + j, p.drop_p)
               ? v / (1.f - p.drop_p) : 0.f;
  };
  if (!p.smooth) {
    for (int j = threadIdx.x; j < p.E; j += 128) p.emb_out[b * p.eo_ld + j] = __float2bfloat16(drop(p.emb_w[amax * p.E + j], j));
    return;
  }
  // softmax numerators once per utterance (not once per output dimension), then a V-term dot product per dimension
  for (int v = threadIdx.x; v < p.V; v += 128) ne_sm[v] = __expf(p.scaling * (x[v] - mx));
  __syncthreads();
  if (warp == 0) {
    float se = 0.f;
    for (int v = lane; v < p.V; v += 32) se += ne_sm[v];
    se = warp_sum(se);
    if (lane == 0) s_inv = 1.f / se;
  }
  __syncthreads();
  const float inv = s_inv;
  for (int j = threadIdx.x; j < p.E; j += 128) {
    float s = 0.f;
    for (int v = 0; v < p.V; ++v) s = fmaf(ne_sm[v], p.emb_w[v * p.E + j], s);
    p.emb_out[b * p.eo_ld + j] = __float2bfloat16(drop(s * inv, j));
  }
}

// Backward of the smooth input embedding emb_{t+1} = softmax(s * logit_t) @ E (model.py:341): adds
// s * p * (dp - <p, dp>), dp[v] = <E[v, :], demb_{t+1}>, to the loss gradient of logit_t. One warp
// per utterance; V <= 1024.
struct SmoothBwdParams {
  float drop_p; const unsigned long long* seed_dev; uint32_t site; int64_t R, row;   // dropout mask of emb_{t+1}
  int B, V, E;
  float scaling;
  const float* logits; int64_t lg_ld;     // logits_t rows
  const float* demb; int64_t de_ld;       // [B] rows: W_ih[:, :E]^T dgates_{t+1}
  const float* emb_w;                     // [V, E]
  const float* dlogits; int64_t dl_ld;    // incoming gradient rows
  float* dl_tot; int64_t dt_ld;           // out rows
};
// One CTA (4 warps) per utterance; dynamic shared memory: E floats (demb) + V floats (dp).
__global__ void __launch_bounds__(128) smooth_dlogit_kernel(SmoothBwdParams p) {
  extern __shared__ float sd_sm[];
  pdl_enter();
  float* de_s = sd_sm;            // [E]
  float* dp_s = sd_sm + p.E;      // [V]: dp[v] = <E[v, :], demb>
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x;
  const float* x = p.logits + b * p.lg_ld;
  float* de = const_cast<float*>(p.demb) + b * p.de_ld;
  {
    const unsigned long long seed = p.drop_p > 0.f ? *p.seed_dev : 0ull;
    for (int j = threadIdx.x; j < p.E; j += 128) {
      float v = de[j];
      if (p.drop_p > 0.f) {   // gradient w.r.t. the un-dropped embedding (kept in the scratch row as before)
        v = dropout_keep(seed, p.site, (static_cast<unsigned long long>(b) * p.R + p.row) * p.E + j, p.drop_p)
                ? v / (1.f - p.drop_p) : 0.f;
        de[j] = v;
      }
      de_s[j] = v;
    }
  }
  __syncthreads();
  for (int v = warp; v < p.V; v += 4) {   // a warp per vocabulary row: coalesced reads of E[v, :]
    float dp = 0.f;
    for (int j = lane; j < p.E; j += 32) dp = fmaf(p.emb_w[v * p.E + j], de_s[j], dp);
    dp = warp_sum(dp);
    if (lane == 0) dp_s[v] = dp;
  }
  __syncthreads();
  if (warp != 0) return;
  float mx = -INFINITY;
  for (int v = lane; v < p.V; v += 32) mx = fmaxf(mx, x[v]);
  mx = warp_max(mx);
  float se = 0.f, pd = 0.f;
  for (int v = lane; v < p.V; v += 32) {
    const float pv = __expf(p.scaling * (x[v] - mx));
    se += pv;
    pd = fmaf(pv, dp_s[v], pd);
  }
  se = warp_sum(se);
  pd = warp_sum(pd);
  const float inv = 1.f / se, dot = pd * inv;
  for (int v = lane; v < p.V; v += 32) {
    const float pv = __expf(p.scaling * (x[v] - mx)) * inv;
    p.dl_tot[b * p.dt_ld + v] = p.dlogits[b * p.dl_ld + v] + p.scaling * pv * (dp_s[v] - dot);
  }
}

// Decoder cell-input dropout (model.py:285) for the per-step path. The mask of c_t as an input of step t+1 is
// element (b*R + t+1)*O + o of site `site`; of the step-(t+1) embedding, element (b*R + t+1)*E + j of site + 1.
// forward : zcd row = [z_t | drop(c_t)]  (the undropped [z_t | c_t] still feeds the output layer)
// backward: dcz_tot = [dz | drop'(dc)] + dzc_all row, also stored as bf16
struct DropRowParams {
  int B, Hd, O;
  int64_t R, row;
  float p;
  const unsigned long long* seed_dev;
  uint32_t site;
  const __nv_bfloat16* zc; __nv_bfloat16* zcd;     // forward: rows at b*R*ZC + row*ZC
  float* dcz_tot; const float* dzc_row; __nv_bfloat16* dcz_row; int64_t d_ld;   // backward
};
__global__ void dec_drop_fwd_kernel(DropRowParams p) {
  pdl_enter();
  const int ZC = p.Hd + p.O;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.B * ZC) return;
  const int b = i / ZC, j = i % ZC;
  const int64_t off = (static_cast<int64_t>(b) * p.R + p.row) * ZC + j;
  __nv_bfloat16 v = p.zc[off];
  if (j >= p.Hd) {
    const bool keep = dropout_keep(*p.seed_dev, p.site, (static_cast<unsigned long long>(b) * p.R + p.row) * p.O + (j - p.Hd), p.p);
    v = keep ? __float2bfloat16(__bfloat162float(v) / (1.f - p.p)) : __float2bfloat16(0.f);
  }
  p.zcd[off] = v;
}
__global__ void dec_drop_bwd_kernel(DropRowParams p) {
  pdl_enter();
  const int ZC = p.Hd + p.O;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.B * ZC) return;
  const int b = i / ZC, j = i % ZC;
  float v = p.dcz_tot[b * ZC + j];
  if (j >= p.Hd) {
    const bool keep = dropout_keep(*p.seed_dev, p.site, (static_cast<unsigned long long>(b) * p.R + p.row) * p.O + (j - p.Hd), p.p);
    v = keep ? v / (1.f - p.p) : 0.f;
  }
  v += p.dzc_row[b * p.d_ld + j];
  p.dcz_tot[b * ZC + j] = v;
  p.dcz_row[b * p.d_ld + j] = __float2bfloat16(v);
}

// Initial alignment (model.py:151-153): 1/len over valid frames, exact zeros beyond.
__global__ void att_init_kernel(const int32_t* __restrict__ enc_lens, int B, int Te, float* w, int64_t w_ld) {
  const int b = blockIdx.x;
  const int len = enc_lens[b];
  const float v = 1.0f / static_cast<float>(len);
  for (int i = threadIdx.x; i < Te; i += blockDim.x) w[b * w_ld + i] = (i < len) ? v : 0.f;
}

// ------------------------------------------------------------------------------------------
// backward kernels
// ------------------------------------------------------------------------------------------
struct DwParams {
  AttGeom g;
  const float* dctx; int64_t dctx_ld;   // [B] rows of H
  const __nv_bfloat16* enc_h;           // [B, Te, H]
  const float* dattc_next;              // [B, Te, C] conv-input gradient of step t+1, or nullptr
  const float* conv_w;                  // [C, 2K+1]
  float* dw;                            // [B, Te]
};
// dw[b,te] = <dctx[b], enc_h[b,te]> + sum_c sum_te' dattc_next[b,te',c] * cw[c, te - te' + K]
// grid (ceil(Te/kTT), B), block 256: one warp per frame.
__global__ void __launch_bounds__(256) att_dw_kernel(DwParams p) {
  extern __shared__ float sm[];
  pdl_enter();
  const AttGeom& g = p.g;
  const int ksz = 2 * g.K + 1;
  float* dctx_s = sm;               // [H]
  float* cw = dctx_s + g.H;         // [C*ksz]
  const int b = blockIdx.y, te0 = blockIdx.x * kTT;
  for (int i = threadIdx.x; i < g.H; i += 256) dctx_s[i] = p.dctx[b * p.dctx_ld + i];
  if (p.dattc_next)
    for (int i = threadIdx.x; i < g.C * ksz; i += 256) cw[i] = p.conv_w[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int tl = warp; tl < kTT; tl += 8) {
    const int te = te0 + tl;
    if (te >= g.Te) break;
    const __nv_bfloat16* row = p.enc_h + (static_cast<int64_t>(b) * g.Te + te) * g.H;
    float s = 0.f;
    for (int h = 2 * lane; h < g.H; h += 64) {
      const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(row + h));
      s = fmaf(dctx_s[h], v.x, s);
      s = fmaf(dctx_s[h + 1], v.y, s);
    }
    if (p.dattc_next) {
      const int lo = max(0, te - g.K), hi = min(g.Te - 1, te + g.K);
      const float* dn = p.dattc_next + static_cast<int64_t>(b) * g.Te * g.C;
      for (int tp = lo + lane; tp <= hi; tp += 32) {
        const int k = te - tp + g.K;
        for (int c = 0; c < g.C; ++c) s = fmaf(dn[tp * g.C + c], cw[c * ksz + k], s);
      }
    }
    s = warp_sum(s);
    if (lane == 0) p.dw[b * g.Te + te] = s;
  }
}

struct EnergyBwdParams {
  AttGeom g;
  float scaling;
  const float* P;
  const float* dz; int64_t dz_ld;
  const float* wprev; int64_t w_ld;     // alignment entering the step (rows of Te)
  const float* wcur;                    // alignment produced by the step (same row stride)
  const float* dw;                      // [B, Te]
  const float* conv_w;
  const float* mlp_att;
  const float* gvec;
  float* dP;                            // [B, Te, A]  +=
  float* ddz; int64_t ddz_ld;           // [B] rows of A, atomically accumulated (pre-zeroed)
  float* dattc;                         // [B, Te, C] out
  float* part;                          // [gridDim.y*gridDim.x][CM+1][Ap] per-CTA running sums (dmlp_att, dgvec)
  int Ap;
};

template <int CM>
__global__ void __launch_bounds__(512) att_energy_bwd_kernel(EnergyBwdParams p) {
  extern __shared__ float sm[];
  pdl_enter();
  const AttGeom& g = p.g;
  const int ksz = 2 * g.K + 1;
  float* wp = sm;
  float* cw = wp + g.Te + 2 * g.K;
  float* conv = cw + g.C * ksz;
  float* de_s = conv + kTT * CM;       // [kTT]
  float* red = de_s + kTT;             // [32]
  float* red2 = red + 32;              // [nwarps][kTT][CM]
  const int b = blockIdx.y, te0 = blockIdx.x * kTT;
  const int a = threadIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  conv_tile<CM>(g, p.wprev + b * p.w_ld, p.conv_w, te0, wp, cw, conv);

  // softmax backward: de = scaling * w * (dw - <w, dw>)
  const float* wc = p.wcur + b * p.w_ld;
  const float* dwb = p.dw + static_cast<int64_t>(b) * g.Te;
  float dot = 0.f;
  for (int i = threadIdx.x; i < g.Te; i += blockDim.x) dot = fmaf(wc[i], dwb[i], dot);
  dot = block_reduce_sum(dot, red, nwarps);
  const int ntl = min(kTT, g.Te - te0);
  if (threadIdx.x < kTT) {
    const int te = te0 + threadIdx.x;
    de_s[threadIdx.x] = (threadIdx.x < ntl) ? p.scaling * wc[te] * (dwb[te] - dot) : 0.f;
  }
  __syncthreads();

  float matt[CM], dmatt[CM];
#pragma unroll
  for (int c = 0; c < CM; ++c) {
    matt[c] = (a < g.A && c < g.C) ? p.mlp_att[a * g.C + c] : 0.f;
    dmatt[c] = 0.f;
  }
  const float dza = (a < g.A) ? p.dz[b * p.dz_ld + a] : 0.f;
  const float gv = (a < g.A) ? p.gvec[a] : 0.f;
  float dgv = 0.f, ddz = 0.f;
  const int64_t pb = (static_cast<int64_t>(b) * g.Te + te0) * g.A + a;
  // P and the running dP of the tile are fetched up front (independent loads in flight together) instead of one
  // dependent L2 round trip per frame inside the loop
  float pv[kTT], dpv[kTT];
#pragma unroll
  for (int tl = 0; tl < kTT; ++tl) {
    const bool ok = tl < ntl && a < g.A;
    pv[tl] = ok ? __ldg(p.P + pb + static_cast<int64_t>(tl) * g.A) : 0.f;
    dpv[tl] = ok ? p.dP[pb + static_cast<int64_t>(tl) * g.A] : 0.f;
  }
#pragma unroll 2
  for (int tl = 0; tl < kTT; ++tl) {
    if (tl >= ntl) break;
    float x = dza + pv[tl];
#pragma unroll
    for (int c = 0; c < CM; ++c) x = fmaf(matt[c], conv[tl * CM + c], x);
    const float s = tanh_acc(x);
    const float de = de_s[tl];
    const float ds = de * gv * (1.f - s * s);   // gv == 0 for a >= A
    dgv = fmaf(de, s, dgv);
    ddz += ds;
    if (a < g.A) p.dP[pb + static_cast<int64_t>(tl) * g.A] = dpv[tl] + ds;
    float prod[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) prod[c] = 0.f;
#pragma unroll
    for (int c = 0; c < CM; ++c) {
      dmatt[c] = fmaf(ds, conv[tl * CM + c], dmatt[c]);
      prod[c] = ds * matt[c];
    }
    const float tot = warp_multi_reduce16(prod, lane);
    const int ci = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    if ((lane & 1) == 0 && ci < CM) red2[(warp * kTT + tl) * CM + ci] = tot;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ntl * g.C; i += blockDim.x) {
    const int tl = i / g.C, c = i % g.C;
    float s = 0.f;
    for (int w = 0; w < nwarps; ++w) s += red2[(w * kTT + tl) * CM + c];
    p.dattc[(static_cast<int64_t>(b) * g.Te + te0 + tl) * g.C + c] = s;
  }
  if (a < g.A) {
    atomicAdd(p.ddz + b * p.ddz_ld + a, ddz);
    float* pp = p.part + static_cast<int64_t>(blockIdx.y * gridDim.x + blockIdx.x) * (CM + 1) * p.Ap + a;
#pragma unroll
    for (int c = 0; c < CM; ++c)
      if (c < g.C) pp[c * p.Ap] += dmatt[c];
    pp[CM * p.Ap] += dgv;
  }
}


// ------------------------------------------------------------------------------------------
// "Lean" energy backward of the per-timestep path (tensor cores): what the time loop itself needs from this phase
// is only ddz_t (-> dz_t -> the cell) and the conv-feature gradient (-> dw_{t-1}); dP and the energy-MLP parameter
// sums are plain sums over (b, t) and are produced after the loop by las_att_param_grads_part from the saved conv
// features and the energy gradients de_all written here (exactly what the cluster-persistent backward does).
// The scalar kernel above spends ~230 instructions per (warp, frame), ~100 of them in the 16-value warp reduction of
// ds x mlp_att; here that contraction and the recomputed mlp_att(conv) are mma.m16n8k16 with bf16 hi/lo operands.
//   grid (ceil(Te/16), B), block 32 * min(A/16, 10): warp w owns the attention k-tiles w, w + nw, ..
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16x2(x, y);
  const float2 h = unpack_bf16x2(hi);
  lo = pack_bf16x2(x - h.x, y - h.y);
}

// B fragments of mlp_att for both products: mattB [A/8][32] x uint4 (k = channel, n = attention dim; hi0, hi1, lo0,
// lo1) and mattB2 [A/16][2][32] x uint2 (k = attention dim, n = channel)
__global__ void att_pack_matt_kernel(const float* __restrict__ mlp_att, int A, int C, uint4* __restrict__ mattB,
                                     uint2* __restrict__ mattB2) {
  const int NT = A / 8, KT = A / 16;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < NT * 32 + KT * 64; i += gridDim.x * blockDim.x) {
    if (i < NT * 32) {
      const int nt = i >> 5, l = i & 31, a = 8 * nt + (l >> 2), c0 = 2 * (l & 3);
      float m[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = c0 + (k & 1) + 8 * (k >> 1);
        m[k] = (c < C) ? mlp_att[a * C + c] : 0.f;
      }
      uint4 v;
      split2(m[0], m[1], v.x, v.z);
      split2(m[2], m[3], v.y, v.w);
      mattB[i] = v;
    } else {
      const int j = i - NT * 32;
      const int kt = j >> 6, nc = (j >> 5) & 1, l = j & 31, c = 8 * nc + (l >> 2), a0 = 16 * kt + 2 * (l & 3);
      float m[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int a = a0 + (k & 1) + 8 * (k >> 1);
        m[k] = (c < C) ? mlp_att[a * C + c] : 0.f;
      }
      mattB2[j] = make_uint2(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]));
    }
  }
}

struct EnergyLeanParams {
  int B, L, Te, A, C, t;
  float scaling;
  const float* P;                       // [B*Te, A]
  const float* dz; int64_t dz_ld;       // mlp_dec(z_t) rows
  const float* wcur; int64_t w_ld;      // alignment produced by step t
  const float* dw;                      // [B, Te]
  const float* conv_save;               // [B, L, Te, 16]
  const uint4* mattB; const uint2* mattB2;
  const float* gvec;
  float* ddz; int64_t ddz_ld;           // [B] rows of A, atomically accumulated (pre-zeroed)
  float* dattc;                         // [B, Te, C] out (this step)
  float* de_all;                        // [B, L, Te] out
};

constexpr int kLeanWarps = 10;
__global__ void __launch_bounds__(32 * kLeanWarps) att_energy_bwd_mma_kernel(EnergyLeanParams p) {
  __shared__ float de_s[16];
  pdl_enter();
  __shared__ float red[kLeanWarps];
  __shared__ float4 dcred[kLeanWarps][2][32];
  const int b = blockIdx.y, te0 = blockIdx.x * 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tig = lane & 3;
  const int nw = blockDim.x >> 5;
  const int Te = p.Te, A = p.A;
  const int ntl = min(16, Te - te0);

  // softmax backward: de = scaling * w * (dw - <w, dw>)
  const float* wc = p.wcur + b * p.w_ld;
  const float* dwb = p.dw + static_cast<int64_t>(b) * Te;
  float dot = 0.f;
  for (int i = threadIdx.x; i < Te; i += blockDim.x) dot = fmaf(wc[i], dwb[i], dot);
  dot = block_reduce_sum(dot, red, nw);
  if (threadIdx.x < 16) {
    float de = 0.f;
    if (static_cast<int>(threadIdx.x) < ntl) {
      const int te = te0 + threadIdx.x;
      de = p.scaling * wc[te] * (dwb[te] - dot);
      p.de_all[(static_cast<int64_t>(b) * p.L + p.t) * Te + te] = de;
    }
    de_s[threadIdx.x] = de;
  }
  __syncthreads();

  // conv features of my two rows as A fragments (k = channel)
  const int r0 = gq, r1 = gq + 8;
  uint32_t Ah[4], Al[4];
  {
    const float* cv = p.conv_save + ((static_cast<int64_t>(b) * p.L + p.t) * Te + te0) * 16;
    const float2 z = make_float2(0.f, 0.f);
    const float2 v0 = r0 < ntl ? *reinterpret_cast<const float2*>(cv + r0 * 16 + 2 * tig) : z;
    const float2 v1 = r1 < ntl ? *reinterpret_cast<const float2*>(cv + r1 * 16 + 2 * tig) : z;
    const float2 v2 = r0 < ntl ? *reinterpret_cast<const float2*>(cv + r0 * 16 + 2 * tig + 8) : z;
    const float2 v3 = r1 < ntl ? *reinterpret_cast<const float2*>(cv + r1 * 16 + 2 * tig + 8) : z;
    split2(v0.x, v0.y, Ah[0], Al[0]);
    split2(v1.x, v1.y, Ah[1], Al[1]);
    split2(v2.x, v2.y, Ah[2], Al[2]);
    split2(v3.x, v3.y, Ah[3], Al[3]);
  }
  const float de0 = de_s[r0], de1 = de_s[r1];      // zero for rows past the tile: their P rows only have to be finite
  const float* P0 = p.P + (static_cast<int64_t>(b) * Te + te0 + min(r0, ntl - 1)) * A + 2 * tig;
  const float* P1 = p.P + (static_cast<int64_t>(b) * Te + te0 + min(r1, ntl - 1)) * A + 2 * tig;
  const float* dzr = p.dz + b * p.dz_ld + 2 * tig;
  const float* gvr = p.gvec + 2 * tig;
  float* ddzr = p.ddz + b * p.ddz_ld + 2 * tig;
  float dcv0[4] = {0.f, 0.f, 0.f, 0.f}, dcv1[4] = {0.f, 0.f, 0.f, 0.f};
  const int KT = A >> 4;
  for (int kt = warp; kt < KT; kt += nw) {
    uint32_t Dh[4], Dl[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int nt = 2 * kt + h, a = 8 * nt;
      const uint4 bm = __ldg(p.mattB + nt * 32 + lane);
      const float2 p0 = __ldg(reinterpret_cast<const float2*>(P0 + a));
      const float2 p1 = __ldg(reinterpret_cast<const float2*>(P1 + a));
      const float2 dz2 = *reinterpret_cast<const float2*>(dzr + a);
      const float2 gv2 = __ldg(reinterpret_cast<const float2*>(gvr + a));
      float acc[4] = {p0.x, p0.y, p1.x, p1.y};
      mma_bf16_16816(acc, Ah, bm.x, bm.y);
      mma_bf16_16816(acc, Al, bm.x, bm.y);
      mma_bf16_16816(acc, Ah, bm.z, bm.w);
      const float s00 = tanh_acc(acc[0] + dz2.x), s01 = tanh_acc(acc[1] + dz2.y);
      const float s10 = tanh_acc(acc[2] + dz2.x), s11 = tanh_acc(acc[3] + dz2.y);
      const float d00 = de0 * gv2.x * (1.f - s00 * s00), d01 = de0 * gv2.y * (1.f - s01 * s01);
      const float d10 = de1 * gv2.x * (1.f - s10 * s10), d11 = de1 * gv2.y * (1.f - s11 * s11);
      float v0 = d00 + d10, v1 = d01 + d11;       // column sums over my 2 rows, then over the 8 row groups
      v0 += __shfl_xor_sync(0xffffffffu, v0, 4);  v1 += __shfl_xor_sync(0xffffffffu, v1, 4);
      v0 += __shfl_xor_sync(0xffffffffu, v0, 8);  v1 += __shfl_xor_sync(0xffffffffu, v1, 8);
      v0 += __shfl_xor_sync(0xffffffffu, v0, 16); v1 += __shfl_xor_sync(0xffffffffu, v1, 16);
      if (gq == 0) {
        atomicAdd(ddzr + a, v0);
        atomicAdd(ddzr + a + 1, v1);
      }
      split2(d00, d01, Dh[2 * h], Dl[2 * h]);
      split2(d10, d11, Dh[2 * h + 1], Dl[2 * h + 1]);
    }
    const uint2 b0 = __ldg(p.mattB2 + (kt * 2) * 32 + lane);
    mma_bf16_16816(dcv0, Dh, b0.x, b0.y);
    mma_bf16_16816(dcv0, Dl, b0.x, b0.y);
    if (p.C > 8) {
      const uint2 b1 = __ldg(p.mattB2 + (kt * 2 + 1) * 32 + lane);
      mma_bf16_16816(dcv1, Dh, b1.x, b1.y);
      mma_bf16_16816(dcv1, Dl, b1.x, b1.y);
    }
  }
  dcred[warp][0][lane] = make_float4(dcv0[0], dcv0[1], dcv0[2], dcv0[3]);
  dcred[warp][1][lane] = make_float4(dcv1[0], dcv1[1], dcv1[2], dcv1[3]);
  __syncthreads();
  if (threadIdx.x < 64) {
    const int half = threadIdx.x >> 5;      // channels 0..7 / 8..15
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int w = 0; w < nw; ++w) {
      const float4 v = dcred[w][half][lane];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    const int c0 = 8 * half + 2 * tig;
    float* drow = p.dattc + (static_cast<int64_t>(b) * Te + te0) * p.C;
    if (r0 < ntl) {
      if (c0 < p.C) drow[r0 * p.C + c0] = s.x;
      if (c0 + 1 < p.C) drow[r0 * p.C + c0 + 1] = s.y;
    }
    if (r1 < ntl) {
      if (c0 < p.C) drow[r1 * p.C + c0] = s.z;
      if (c0 + 1 < p.C) drow[r1 * p.C + c0 + 1] = s.w;
    }
  }
}

// Final reduction of the per-CTA running sums: dmlp_att[a, c] (+=), dgvec[a] (+=).
__global__ void att_part_reduce_kernel(const float* __restrict__ part, int ncta, int CM, int Ap, int A, int C,
                                       float* __restrict__ dmlp_att, float* __restrict__ dgvec) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (C + 1) * A) return;
  const int c = idx / A, a = idx % A;
  const int slot = (c < C) ? c : CM;
  float s = 0.f;
  for (int k = 0; k < ncta; ++k) s += part[(static_cast<int64_t>(k) * (CM + 1) + slot) * Ap + a];
  if (c < C) dmlp_att[a * C + c] += s;
  else dgvec[a] += s;
}

// dconv_w[c, k] += sum_{t, b, te} dattc[t, b, te, c] * w_{t-1}[b, te + k - K]   (w_{t-1} = ws_alloc row t).
// Pass 1: grid (ceil(L/kDT), B), block 256 (thread = tap k): partial[cta][c][k] over kDT steps of one
// utterance, operands staged in shared memory. Pass 2: deterministic sum over the CTAs.
constexpr int kDT = 8;
__global__ void __launch_bounds__(256) att_dconv_partial_kernel(const float* __restrict__ dattc_all,
                                                                const float* __restrict__ ws_alloc, int L, int B, int Te,
                                                                int C, int K, float* __restrict__ partial) {
  extern __shared__ float sm[];
  float* wp = sm;                       // [Te + 2K] zero padded alignment
  float* d = wp + Te + 2 * K;           // [Te][C]
  const int ksz = 2 * K + 1;
  const int b = blockIdx.y, t0 = blockIdx.x * kDT;
  float acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = 0.f;
  for (int t = t0; t < min(L, t0 + kDT); ++t) {
    __syncthreads();
    for (int i = threadIdx.x; i < Te + 2 * K; i += 256) {
      const int j = i - K;
      wp[i] = (j >= 0 && j < Te) ? ws_alloc[(static_cast<int64_t>(b) * (L + 1) + t) * Te + j] : 0.f;
    }
    const float* src = dattc_all + (static_cast<int64_t>(t) * B + b) * Te * C;
    for (int i = threadIdx.x; i < Te * C; i += 256) d[i] = src[i];
    __syncthreads();
    for (int k = threadIdx.x; k < ksz; k += 256) {
      const float* wk = wp + k;           // wk[te] = w[te + k - K]
      for (int te = 0; te < Te; ++te) {
        const float w = wk[te];
        const float* dr = d + te * C;
#pragma unroll
        for (int c = 0; c < 16; ++c)
          if (c < C) acc[c] = fmaf(dr[c], w, acc[c]);
      }
    }
  }
  float* out = partial + static_cast<int64_t>(blockIdx.y * gridDim.x + blockIdx.x) * C * ksz;
  for (int k = threadIdx.x; k < ksz; k += 256)
#pragma unroll
    for (int c = 0; c < 16; ++c)
      if (c < C) out[c * ksz + k] = acc[c];
}
__global__ void att_dconv_reduce_kernel(const float* __restrict__ partial, int ncta, int n, float* __restrict__ dconv_w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int k = 0; k < ncta; ++k) s += partial[static_cast<int64_t>(k) * n + i];
  dconv_w[i] += s;
}
// Tensor-core variant: for one (t, b) the conv-weight gradient is D[c][k] = sum_te X[te][c] w[te + k - K], i.e.
// A = X^T (16 channels x Te) times the Toeplitz matrix of the alignment. A CTA walks (t, b) pairs, stages X^T and the
// zero-padded alignment as bf16 in shared memory (the alignment twice, once shifted by one element, so that the
// two consecutive taps of a B-fragment word are one aligned 32-bit load for either parity), and keeps D in the
// accumulators of its 8 warps (warp w owns the 8-tap tiles w, w+8, ...). partial[cta][c][k] is reduced afterwards
// in a fixed order (att_dconv_reduce_kernel). bf16 operands, f32 accumulation.
__global__ void __launch_bounds__(256) att_dconv_mma_kernel(const float* __restrict__ dattc_all,
                                                            const float* __restrict__ ws_alloc, int L, int B, int Te,
                                                            int C, int K, int KT, int NT, float* __restrict__ partial) {
  extern __shared__ __align__(16) uint8_t dsm[];
  const int a_ld = KT * 16 + 8;                                  // bf16 elements per channel row (word stride = 4 mod 32)
  const int wlen = KT * 16 + NT * 8 + 16;                        // padded alignment, wpad[i] = w[i - K]
  const int wwords = (wlen / 2 + 31) / 32 * 32 + 16;             // second array lands 16 banks away from the first
  __nv_bfloat16* A_s = reinterpret_cast<__nv_bfloat16*>(dsm);    // [16][a_ld]
  uint32_t* wE = reinterpret_cast<uint32_t*>(dsm + 16 * a_ld * 2);   // wE[m] = (wpad[2m], wpad[2m+1])
  uint32_t* wO = wE + wwords;                                        // wO[m] = (wpad[2m+1], wpad[2m+2])
  const int ksz = 2 * K + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;
  for (int i = threadIdx.x; i < 16 * a_ld; i += 256) A_s[i] = __float2bfloat16(0.f);
  const uint32_t* wl = ((g & 1) ? wO : wE) + ((2 * tig + g - (g & 1)) >> 1);
  const uint32_t* a0p = reinterpret_cast<const uint32_t*>(A_s + g * a_ld) + tig;
  const uint32_t* a1p = reinterpret_cast<const uint32_t*>(A_s + (g + 8) * a_ld) + tig;
  const int n_items = L * B;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int t = item / B, b = item - t * B;
    __syncthreads();                                               // previous item's fragments are consumed
    const float* src = dattc_all + (static_cast<int64_t>(t) * B + b) * Te * C;
    for (int i = threadIdx.x; i < Te * C; i += 256) {
      const int te = i / C, c = i - te * C;
      A_s[c * a_ld + te] = __float2bfloat16(src[i]);
    }
    const float* wrow = ws_alloc + (static_cast<int64_t>(b) * (L + 1) + t) * Te;
    for (int m = threadIdx.x; m < wlen / 2; m += 256) {
      float v[3];
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        const int j = 2 * m + e - K;
        v[e] = (j >= 0 && j < Te) ? wrow[j] : 0.f;
      }
      wE[m] = pack_bf16x2(v[0], v[1]);
      wO[m] = pack_bf16x2(v[1], v[2]);
    }
    __syncthreads();
    for (int kt = 0; kt < KT; ++kt) {
      const uint32_t Af[4] = {a0p[8 * kt], a1p[8 * kt], a0p[8 * kt + 4], a1p[8 * kt + 4]};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int nt = warp + 8 * j;
        if (nt < NT) mma_bf16_16816(acc[j], Af, wl[8 * kt + 4 * nt], wl[8 * kt + 4 * nt + 4]);
      }
    }
  }
  float* out = partial + static_cast<int64_t>(blockIdx.x) * C * ksz;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int nt = warp + 8 * j;
    if (nt >= NT) continue;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = g + 8 * (e >> 1), k = 8 * nt + 2 * tig + (e & 1);
      if (c < C && k < ksz) out[c * ksz + k] = acc[j][e];
    }
  }
}

static int launch_att_dconv(const float* dattc_all, const float* ws_alloc, int L, int B, int Te, int C, int K,
                            float* dconv_w, float* partial_ws, cudaStream_t stream) {
  const int ksz = 2 * K + 1;
  {
    const int KT = (Te + 15) / 16, NT = (ksz + 7) / 8;
    const int a_ld = KT * 16 + 8, wlen = KT * 16 + NT * 8 + 16, wwords = (wlen / 2 + 31) / 32 * 32 + 16;
    const size_t smem = static_cast<size_t>(16) * a_ld * 2 + static_cast<size_t>(2) * wwords * 4;
    if (NT <= 32 && smem <= 48 * 1024) {
      // as many partial slots as the scratch was sized for (las_att_scratch_floats): ceil(L/kDT) * B
      int ncta = ((L + kDT - 1) / kDT) * B;
      if (ncta > L * B) ncta = L * B;
      if (ncta > 4 * num_sms()) ncta = 4 * num_sms();
      att_dconv_mma_kernel<<<ncta, 256, smem, stream>>>(dattc_all, ws_alloc, L, B, Te, C, K, KT, NT, partial_ws); ++g_launches;
      att_dconv_reduce_kernel<<<(C * ksz + 255) / 256, 256, 0, stream>>>(partial_ws, ncta, C * ksz, dconv_w); ++g_launches;
      return 0;
    }
  }
  LAS_REQUIRE(ksz <= 256, "att_dconv: conv kernel of %d taps is wider than 256", ksz);
  const dim3 grid((L + kDT - 1) / kDT, B);
  const size_t smem = (Te + 2 * K + static_cast<size_t>(Te) * C) * sizeof(float);
  LAS_REQUIRE(smem <= 48 * 1024, "att_dconv: %zu bytes of shared memory needed", smem);
  att_dconv_partial_kernel<<<grid, 256, smem, stream>>>(dattc_all, ws_alloc, L, B, Te, C, K, partial_ws); ++g_launches;
  att_dconv_reduce_kernel<<<(C * ksz + 255) / 256, 256, 0, stream>>>(partial_ws, grid.x * grid.y, C * ksz, dconv_w); ++g_launches;
  return 0;
}

// denc[b, te, h] (+)= sum_t ws[b, t, te] * dctx[b, t, h]     (gradient of the context bmm)
// grid (ceil(Te/kDE), B): a CTA owns kDE frames of one utterance, so dctx[b] is streamed Te/kDE times instead of
// Te times (one frame per CTA made this kernel L2-bound: 123 us at config 2); the alignment columns of the tile
// are staged in shared memory per chunk of kDEC steps.
constexpr int kDE = 8;
constexpr int kDEC = 128;
__global__ void __launch_bounds__(320) att_denc_kernel(const float* __restrict__ ws_alloc, const float* __restrict__ dctx_all,
                                                       int L, int B, int Te, int H, float* __restrict__ denc, int accumulate) {
  __shared__ __align__(16) float wt[kDEC][kDE];
  const int b = blockIdx.y, te0 = blockIdx.x * kDE;
  const int nte = min(kDE, Te - te0);
  const float* w = ws_alloc + (static_cast<int64_t>(b) * (L + 1) + 1) * Te + te0;
  const float* d = dctx_all + static_cast<int64_t>(b) * L * H;
  for (int hb = 0; hb < H; hb += blockDim.x) {
    const int h = hb + threadIdx.x;
    float acc[kDE];
#pragma unroll
    for (int f = 0; f < kDE; ++f) acc[f] = 0.f;
    for (int t0 = 0; t0 < L; t0 += kDEC) {
      const int tn = min(kDEC, L - t0);
      __syncthreads();
      for (int i = threadIdx.x; i < tn * kDE; i += blockDim.x) {
        const int tt = i / kDE, f = i % kDE;
        wt[tt][f] = f < nte ? w[static_cast<int64_t>(t0 + tt) * Te + f] : 0.f;
      }
      __syncthreads();
      if (h < H) {
        const float* dp = d + static_cast<int64_t>(t0) * H + h;
#pragma unroll 4
        for (int tt = 0; tt < tn; ++tt) {
          const float dv = __ldg(dp + static_cast<int64_t>(tt) * H);
          const float4 w0 = *reinterpret_cast<const float4*>(&wt[tt][0]);
          const float4 w1 = *reinterpret_cast<const float4*>(&wt[tt][4]);
          acc[0] = fmaf(w0.x, dv, acc[0]); acc[1] = fmaf(w0.y, dv, acc[1]);
          acc[2] = fmaf(w0.z, dv, acc[2]); acc[3] = fmaf(w0.w, dv, acc[3]);
          acc[4] = fmaf(w1.x, dv, acc[4]); acc[5] = fmaf(w1.y, dv, acc[5]);
          acc[6] = fmaf(w1.z, dv, acc[6]); acc[7] = fmaf(w1.w, dv, acc[7]);
        }
      }
    }
    if (h < H) {
#pragma unroll
      for (int f = 0; f < kDE; ++f) {
        if (f < nte) {
          float* o = denc + (static_cast<int64_t>(b) * Te + te0 + f) * H + h;
          *o = accumulate ? *o + acc[f] : acc[f];
        }
      }
    }
  }
}

// Post-loop parameter gradients of the energy MLP for the persistent backward: every (b, te, a) is
// independent, so nothing of this sits in the serial loop. For its frame tile a CTA walks all steps t
// and recomputes s = tanh(P + dz_t + mlp_att conv_t) from the saved conv features and energy gradients:
//   dP[b,te,a] = sum_t ds,  part[cta][c][a] = sum ds*conv[c],  part[cta][CM][a] = sum de*s,  ds = de gv (1 - s^2)
// grid (ceil(Te/kPG), B), block = A rounded up to a warp multiple.
// WHAT selects the outputs: bit 0 = dP (the critical path: the encoder's backward waits for it), bit 1 = the
// parameter partial sums (nothing waits for them: the trainers run that instance on the weight-gradient stream).
constexpr int kPG = 7;   // frames per CTA (Te = 125: 18 x 32 = 576 CTAs = 1.95 waves of the 296 two-per-SM slots; 8 frames gave 1.73 waves for the price of 2); blocks of <= 320 threads are compiled for two CTAs per SM
constexpr int kPGT = 16;   // decoder steps per shared-memory chunk
template <int CM, int MAXT, int MINB, int WHAT>
__global__ void __launch_bounds__(MAXT, MINB) att_param_grad_kernel(const float* __restrict__ P, const float* __restrict__ dzf,
                                                             const float* __restrict__ conv_save,
                                                             const float* __restrict__ de_all,
                                                             const float* __restrict__ mlp_att,
                                                             const float* __restrict__ gvec, int B, int L, int Te, int A,
                                                             int C, int Ap, float* __restrict__ dP,
                                                             float* __restrict__ part) {
  // step-major loop: dz_t is read once per step, the conv features / energy gradients of the tile's frames are
  // staged per chunk of kPGT steps with cp.async (double buffered), P and the accumulators live in registers
  __shared__ __align__(16) float conv_s[2][kPGT][kPG][16];
  __shared__ float de_s[2][kPGT][kPG];
  const int b = blockIdx.y, te0 = blockIdx.x * kPG, a = threadIdx.x;
  const bool ok = a < A;
  const int ntl = min(kPG, Te - te0);
  float matt[CM], dmatt[CM];
#pragma unroll
  for (int c = 0; c < CM; ++c) {
    matt[c] = (ok && c < C) ? mlp_att[a * C + c] : 0.f;
    dmatt[c] = 0.f;
  }
  const float gv = ok ? gvec[a] : 0.f;
  float dgv = 0.f, pv[kPG], dp[kPG];
#pragma unroll
  for (int f = 0; f < kPG; ++f) {
    pv[f] = (ok && f < ntl) ? P[(static_cast<int64_t>(b) * Te + te0 + f) * A + a] : 0.f;
    dp[f] = 0.f;
  }
  auto stage = [&](int chunk, int buf) {
    const int t0 = chunk * kPGT;
    // conv rows: kPGT x ntl x 16 floats in 16-byte pieces; energy gradients: kPGT x ntl floats
    for (int i = threadIdx.x; i < kPGT * kPG * 4; i += blockDim.x) {
      const int q4 = i & 3, f = (i >> 2) % kPG, tt = i / (4 * kPG);
      const int t = t0 + tt;
      float* dst = &conv_s[buf][tt][f][4 * q4];
      if (t < L && f < ntl) {
        const float* src = conv_save + ((static_cast<int64_t>(b) * L + t) * Te + te0 + f) * 16 + 4 * q4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
      } else {
        *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    for (int i = threadIdx.x; i < kPGT * kPG; i += blockDim.x) {
      const int f = i % kPG, tt = i / kPG, t = t0 + tt;
      float* dst = &de_s[buf][tt][f];
      if (t < L && f < ntl) {
        const float* src = de_all + (static_cast<int64_t>(b) * L + t) * Te + te0 + f;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
      } else {
        *dst = 0.f;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int nchunks = (L + kPGT - 1) / kPGT;
  stage(0, 0);
  for (int ch = 0; ch < nchunks; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < nchunks) stage(ch + 1, buf ^ 1);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    const int tn = min(kPGT, L - ch * kPGT);
    for (int tt = 0; tt < tn; ++tt) {
      const float dz = ok ? __ldg(dzf + (static_cast<int64_t>(b) * L + ch * kPGT + tt) * A + a) : 0.f;
#pragma unroll
      for (int f = 0; f < kPG; ++f) {
        const float4* cv = reinterpret_cast<const float4*>(&conv_s[buf][tt][f][0]);
        float conv[CM];
#pragma unroll
        for (int c4 = 0; c4 < CM / 4; ++c4) {
          const float4 v = cv[c4];
          conv[4 * c4] = v.x; conv[4 * c4 + 1] = v.y; conv[4 * c4 + 2] = v.z; conv[4 * c4 + 3] = v.w;
        }
        const float de = de_s[buf][tt][f];
        float x = pv[f] + dz;
#pragma unroll
        for (int c = 0; c < CM; ++c) x = fmaf(matt[c], conv[c], x);
        const float sx = tanh_fast(x);
        const float ds = de * gv * (1.f - sx * sx);
        if (WHAT & 1) dp[f] += ds;
        if (WHAT & 2) {
          dgv = fmaf(de, sx, dgv);
#pragma unroll
          for (int c = 0; c < CM; ++c) dmatt[c] = fmaf(ds, conv[c], dmatt[c]);
        }
      }
    }
    __syncthreads();
  }
  if (ok && (WHAT & 1)) {
#pragma unroll
    for (int f = 0; f < kPG; ++f)
      if (f < ntl) dP[(static_cast<int64_t>(b) * Te + te0 + f) * A + a] = dp[f];
  }
  if (ok && (WHAT & 2)) {
    float* pp = part + static_cast<int64_t>(blockIdx.y * gridDim.x + blockIdx.x) * (CM + 1) * Ap + a;
#pragma unroll
    for (int c = 0; c < CM; ++c) pp[c * Ap] = dmatt[c];
    pp[CM * Ap] = dgv;
  }
}

// dP alone (the part of the above that the encoder's backward waits for), with the mlp_att contraction on tensor cores:
// for a tile of 16 frames and one step t, x[frame][a] = P + dz_t + conv_t[frame][:] . mlp_att[a][:] is one
// mma.m16n8k16 per 8 attention dims (rows = frames, K = the <= 16 conv channels, bf16 hi/lo split of both operands as
// in the forward kernel: 3 MMAs), initialised from P + dz_t; the element-wise tail (tanh, 1 - s^2, times de gv, sum over
// t) stays in the accumulator layout. Per element ~6 issued instructions instead of ~25 (12 of them FMAs on operands
// every thread fetched from shared memory for itself). grid (ceil(Te/16), B), block A (a warp = 32 attention dims).
constexpr int kDPT = 8;        // decoder steps per shared-memory chunk
constexpr int kDPld = 24;      // row stride (floats) of a staged conv row: conflict-free 8-byte fragment loads
__device__ __forceinline__ void split2_bf16(float x, float y, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16x2(x, y);
  const float2 h = unpack_bf16x2(hi);
  lo = pack_bf16x2(x - h.x, y - h.y);
}
__global__ void __launch_bounds__(320, 2) att_dp_mma_kernel(const float* __restrict__ P, const float* __restrict__ dzf,
                                                            const float* __restrict__ conv_save,
                                                            const float* __restrict__ de_all,
                                                            const float* __restrict__ mlp_att,
                                                            const float* __restrict__ gvec, int B, int L, int Te, int A,
                                                            int C, float* __restrict__ dP) {
  __shared__ __align__(16) float conv_s[2][kDPT][16][kDPld];
  __shared__ float de_s[2][kDPT][16];
  const int b = blockIdx.y, te0 = blockIdx.x * 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tig = lane & 3;
  const int ntl = min(16, Te - te0);
  const int a_w = 32 * warp;                        // this warp's attention dims a_w .. a_w + 31 (4 n-tiles)
  // constant operands: B fragments of mlp_att^T (k = channel, n = attention dim), P tile, gvec
  uint32_t Bh[4][2], Bl[4][2];
  float pv[4][4], gv[4][2], dp[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int an = a_w + 8 * nt + gq;               // B fragment column of this lane
    float m[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = 2 * tig + (k & 1) + 8 * (k >> 1);
      m[k] = (an < A && c < C) ? mlp_att[an * C + c] : 0.f;
    }
    split2_bf16(m[0], m[1], Bh[nt][0], Bl[nt][0]);
    split2_bf16(m[2], m[3], Bh[nt][1], Bl[nt][1]);
    const int ac = a_w + 8 * nt + 2 * tig;          // accumulator columns ac, ac + 1
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int fr = gq + 8 * (e >> 1), a = ac + (e & 1);
      pv[nt][e] = (fr < ntl && a < A) ? P[(static_cast<int64_t>(b) * Te + te0 + fr) * A + a] : 0.f;
      dp[nt][e] = 0.f;
    }
    gv[nt][0] = ac < A ? gvec[ac] : 0.f;
    gv[nt][1] = ac + 1 < A ? gvec[ac + 1] : 0.f;
  }
  auto stage = [&](int chunk, int buf) {
    const int t0 = chunk * kDPT;
    for (int i = threadIdx.x; i < kDPT * 16 * 4; i += blockDim.x) {
      const int q4 = i & 3, f = (i >> 2) & 15, tt = i >> 6, t = t0 + tt;
      float* dst = &conv_s[buf][tt][f][4 * q4];
      if (t < L && f < ntl) {
        const float* src = conv_save + ((static_cast<int64_t>(b) * L + t) * Te + te0 + f) * 16 + 4 * q4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
      } else {
        *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    for (int i = threadIdx.x; i < kDPT * 16; i += blockDim.x) {
      const int f = i & 15, tt = i >> 4, t = t0 + tt;
      float* dst = &de_s[buf][tt][f];
      if (t < L && f < ntl) {
        const float* src = de_all + (static_cast<int64_t>(b) * L + t) * Te + te0 + f;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
      } else {
        *dst = 0.f;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int nchunks = (L + kDPT - 1) / kDPT;
  const float* dz_row = dzf + static_cast<int64_t>(b) * L * A + a_w + 2 * tig;
  stage(0, 0);
  for (int ch = 0; ch < nchunks; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < nchunks) stage(ch + 1, buf ^ 1);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    const int tn = min(kDPT, L - ch * kDPT);
    for (int tt = 0; tt < tn; ++tt) {
      // A fragments of this step's conv tile (rows = frames gq / gq + 8, k = channels 2tig.. / 2tig + 8..)
      const float2 c00 = *reinterpret_cast<const float2*>(&conv_s[buf][tt][gq][2 * tig]);
      const float2 c10 = *reinterpret_cast<const float2*>(&conv_s[buf][tt][gq + 8][2 * tig]);
      const float2 c01 = *reinterpret_cast<const float2*>(&conv_s[buf][tt][gq][2 * tig + 8]);
      const float2 c11 = *reinterpret_cast<const float2*>(&conv_s[buf][tt][gq + 8][2 * tig + 8]);
      uint32_t Ah[4], Al[4];
      split2_bf16(c00.x, c00.y, Ah[0], Al[0]);
      split2_bf16(c10.x, c10.y, Ah[1], Al[1]);
      split2_bf16(c01.x, c01.y, Ah[2], Al[2]);
      split2_bf16(c11.x, c11.y, Ah[3], Al[3]);
      const float de0 = de_s[buf][tt][gq], de1 = de_s[buf][tt][gq + 8];
      const float* dzp = dz_row + static_cast<int64_t>(ch * kDPT + tt) * A;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int ac = a_w + 8 * nt + 2 * tig;
        float2 dz = make_float2(0.f, 0.f);
        if (ac + 1 < A) dz = __ldg(reinterpret_cast<const float2*>(dzp + 8 * nt));
        float x[4] = {pv[nt][0] + dz.x, pv[nt][1] + dz.y, pv[nt][2] + dz.x, pv[nt][3] + dz.y};
        mma_bf16_16816(x, Ah, Bh[nt][0], Bh[nt][1]);
        mma_bf16_16816(x, Al, Bh[nt][0], Bh[nt][1]);
        mma_bf16_16816(x, Ah, Bl[nt][0], Bl[nt][1]);
        const float g0 = gv[nt][0], g1 = gv[nt][1];
        const float s0 = tanh_fast(x[0]), s1 = tanh_fast(x[1]), s2 = tanh_fast(x[2]), s3 = tanh_fast(x[3]);
        dp[nt][0] = fmaf(de0 * g0, 1.f - s0 * s0, dp[nt][0]);
        dp[nt][1] = fmaf(de0 * g1, 1.f - s1 * s1, dp[nt][1]);
        dp[nt][2] = fmaf(de1 * g0, 1.f - s2 * s2, dp[nt][2]);
        dp[nt][3] = fmaf(de1 * g1, 1.f - s3 * s3, dp[nt][3]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int ac = a_w + 8 * nt + 2 * tig;
    if (ac + 1 < A) {
      if (gq < ntl)
        *reinterpret_cast<float2*>(dP + (static_cast<int64_t>(b) * Te + te0 + gq) * A + ac) = make_float2(dp[nt][0], dp[nt][1]);
      if (gq + 8 < ntl)
        *reinterpret_cast<float2*>(dP + (static_cast<int64_t>(b) * Te + te0 + gq + 8) * A + ac) = make_float2(dp[nt][2], dp[nt][3]);
    }
  }
}

static size_t energy_smem(const las_dec_args* a, int CM, int nwarps, bool bwd) {
  const int ksz = 2 * a->K + 1;
  size_t f = a->Te + 2 * a->K + static_cast<size_t>(a->C) * ksz + kTT * CM;
  if (!bwd) f += static_cast<size_t>(nwarps) * kTT;
  else f += kTT + 32 + static_cast<size_t>(nwarps) * kTT * CM;
  return f * sizeof(float);
}

template <typename Kern>
static int ensure_smem(Kern kern, size_t bytes) {
  if (bytes > 48 * 1024) {
    LAS_REQUIRE(bytes <= 227 * 1024, "decoder: attention kernel needs %zu bytes of shared memory", bytes);
    LAS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
  }
  return 0;
}

static int check_args(const las_dec_args* a) {
  LAS_REQUIRE(a->Hd % 8 == 0 && a->O % 8 == 0 && a->H % 8 == 0, "decoder: hidden/att_odim/encoder dims must be multiples of 8");
  LAS_REQUIRE(a->A >= 1 && a->A <= 512, "decoder: att_dim %d out of range [1,512]", a->A);
  LAS_REQUIRE(a->C >= 1 && a->C <= 16, "decoder: conv_channels %d out of range [1,16]", a->C);
  LAS_REQUIRE(a->B >= 1 && a->L >= 1 && a->Te >= 1, "decoder: empty batch / sequence");
  return 0;
}

}  // namespace las

using namespace las;

extern "C" {

int las_att_init(const int32_t* enc_lens, int B, int Te, float* w, int64_t w_ld, void* stream) {
  if (B == 0) return 0;
  att_init_kernel<<<B, 128, 0, static_cast<cudaStream_t>(stream)>>>(enc_lens, B, Te, w, w_ld); ++g_launches;
  LAS_LAUNCH_CHECK();
  return 0;
}

int las_dec_persistent_supported(const las_dec_args* a) { return dec_persist_supported(a); }

int64_t las_dec_persistent_pack_bytes(int which, int Hd, int O, int A) { return dec_persist_pack_bytes(which, Hd, O, A); }

int las_dec_persistent_pack(int which, const float* W, int64_t ld, int Hd, int O, int A, void* out, void* stream) {
  return dec_persist_pack(which, W, ld, Hd, O, A, out, static_cast<cudaStream_t>(stream));
}

int las_att_dq(const float* ws_alloc, const float* dc_all, int L, int B, int Te, int O, float* dQ, void* stream) {
  if (B == 0 || Te == 0) return 0;
  att_denc_kernel<<<dim3((Te + kDE - 1) / kDE, B), (O <= 256 ? 256 : 320), 0, static_cast<cudaStream_t>(stream)>>>(ws_alloc, dc_all, L, B, Te, O, dQ, 0); ++g_launches;
  LAS_LAUNCH_CHECK();
  return 0;
}

static int64_t att_lean_pack_floats(int A) { return (static_cast<int64_t>((A + 7) / 8) * 32 * 4 + static_cast<int64_t>((A + 15) / 16) * 64 * 2 + 63) / 64 * 64; }
static int64_t att_scratch_main_floats(int B, int L, int Te, int A, int C, int K) {
  const int64_t Ap = (A + 31) / 32 * 32;
  const int64_t a = static_cast<int64_t>((Te + kPG - 1) / kPG) * B * 17 * Ap;
  const int64_t b = static_cast<int64_t>((L + kDT - 1) / kDT) * B * C * (2 * K + 1);
  const int64_t c = static_cast<int64_t>((Te + kTT - 1) / kTT) * B * 17 * Ap;
  return ((a > b ? (a > c ? a : c) : (b > c ? b : c)) + 64 + 63) / 64 * 64;
}

int64_t las_att_scratch_floats(int B, int L, int Te, int A, int C, int K) {
  // shared scratch of the attention backward: per-CTA partial sums of (a) the energy-MLP parameter
  // gradients, (b) the conv-weight gradient
  // + the packed mlp_att fragments of the lean per-timestep energy backward, kept BEHIND the shared region
  return att_scratch_main_floats(B, L, Te, A, C, K) + att_lean_pack_floats(A);
}

int las_att_bwd_lean_supported(int A, int C) {
  static const bool on = getenv("LAS_ATT_LEAN") == nullptr || atoi(getenv("LAS_ATT_LEAN")) != 0;
  return (on && A >= 16 && A % 16 == 0 && A <= 1024 && C >= 1 && C <= 16) ? 1 : 0;
}

int las_att_dconv(const float* dattc_all, const float* ws_alloc, int L, int B, int Te, int C, int K, float* dconv_w,
                  float* scratch, void* stream) {
  LAS_REQUIRE(C >= 1 && C <= 16, "att_dconv: conv_channels %d out of range [1,16]", C);
  if (B == 0 || Te == 0 || L == 0) return 0;
  if (int rc = launch_att_dconv(dattc_all, ws_alloc, L, B, Te, C, K, dconv_w, scratch, static_cast<cudaStream_t>(stream))) return rc;
  LAS_LAUNCH_CHECK();
  return 0;
}

int las_att_param_grads_part(const float* P, const float* dzf, const float* conv_save, const float* de_all,
                             const float* mlp_att, const float* gvec, int B, int L, int Te, int A, int C, int what,
                             float* dP, float* part_ws, float* dmlp_att, float* dgvec, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LAS_REQUIRE(A >= 1 && A <= 512 && C >= 1 && C <= 16, "att_param_grads: att_dim / conv_channels out of range");
  LAS_REQUIRE(what >= 1 && what <= 3, "att_param_grads: what = %d (1 = dP, 2 = parameter gradients, 3 = both)", what);
  if (B == 0 || Te == 0 || L == 0) return 0;
  const int CM = (C + 3) / 4 * 4;          // channel count padded to whole float4 pieces: no FMAs on padding beyond that
  const int threads = (A + 31) / 32 * 32;
  static const bool dp_mma = getenv("LAS_DP_MMA") == nullptr || atoi(getenv("LAS_DP_MMA")) != 0;
  if (what == 1 && dp_mma && A % 32 == 0 && A <= 320 && reinterpret_cast<uintptr_t>(dzf) % 8 == 0 &&
      reinterpret_cast<uintptr_t>(dP) % 8 == 0) {
    // the critical-path half (dP alone): tensor-core form
    att_dp_mma_kernel<<<dim3((Te + 15) / 16, B), A, 0, stream>>>(P, dzf, conv_save, de_all, mlp_att, gvec, B, L, Te, A, C, dP);
    ++g_launches;
    LAS_LAUNCH_CHECK();
    return 0;
  }
  const dim3 grid((Te + kPG - 1) / kPG, B);
  // blocks of <= 320 threads are compiled for two CTAs per SM (<= 102 registers): the loop is latency-bound
#define LAS_APG2(CMV, W)                                                                                               \
  do {                                                                                                                 \
    if (threads <= 320)                                                                                                \
      att_param_grad_kernel<CMV, 320, 2, W><<<grid, threads, 0, stream>>>(P, dzf, conv_save, de_all, mlp_att, gvec, B, \
                                                                           L, Te, A, C, threads, dP, part_ws);          \
    else                                                                                                               \
      att_param_grad_kernel<CMV, 512, 1, W><<<grid, threads, 0, stream>>>(P, dzf, conv_save, de_all, mlp_att, gvec, B, \
                                                                           L, Te, A, C, threads, dP, part_ws);          \
  } while (0)
#define LAS_APG(CMV)                                                                                                   \
  do {                                                                                                                 \
    if (what == 1) LAS_APG2(CMV, 1);                                                                                   \
    else if (what == 2) LAS_APG2(CMV, 2);                                                                              \
    else LAS_APG2(CMV, 3);                                                                                             \
  } while (0)
  if (CM == 4) LAS_APG(4);
  else if (CM == 8) LAS_APG(8);
  else if (CM == 12) LAS_APG(12);
  else LAS_APG(16);
#undef LAS_APG
#undef LAS_APG2
  ++g_launches;
  if (what & 2) {
    att_part_reduce_kernel<<<((C + 1) * A + 255) / 256, 256, 0, stream>>>(part_ws, grid.x * grid.y, CM, threads, A, C, dmlp_att, dgvec); ++g_launches;
  }
  LAS_LAUNCH_CHECK();
  return 0;
}

int las_att_param_grads(const float* P, const float* dzf, const float* conv_save, const float* de_all,
                        const float* mlp_att, const float* gvec, int B, int L, int Te, int A, int C, float* dP,
                        float* part_ws, float* dmlp_att, float* dgvec, void* stream_) {
  return las_att_param_grads_part(P, dzf, conv_save, de_all, mlp_att, gvec, B, L, Te, A, C, 3, dP, part_ws, dmlp_att,
                                  dgvec, stream_);
}

// Per-step launches of decoder steps [t_begin, t_end). `with_cell`: LSTMCell update (1); the attention (2)-(5) always
// runs; `with_out`: output layer + next embedding of the free-running modes (6)-(7).
static int dec_fwd_steps(const las_dec_args* a, cudaStream_t stream, int t_begin, int t_end, bool with_cell, bool with_out);

int las_dec_fwd(const las_dec_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = check_args(a)) return rc;
  const int t_begin = a->t_begin, t_end = (a->t_end > 0 && a->t_end < a->L) ? a->t_end : a->L;
  LAS_REQUIRE(t_begin >= 0 && t_begin <= t_end, "decoder: step range [%d, %d) invalid", t_begin, t_end);
  if (dec_persist_supported(a)) {
    LAS_REQUIRE(t_begin == 0 && t_end == a->L, "decoder: the persistent kernel runs all steps in one launch");
    ++g_path[LAS_PATH_DEC_PERSIST_FWD];
    return dec_persist_fwd(a, stream);   // one cluster-persistent launch
  }
  ++g_path[LAS_PATH_DEC_STEP_FWD];
  return dec_fwd_steps(a, stream, t_begin, t_end, true, true);
}

// One stand-alone AttLoc.forward (model.py:139-173) on the per-step kernels: reads the decoder state z from
// zc[:, t+1, :Hd] and the previous alignment from ws[:, t], writes ws[:, t+1], ctx[:, t+1] and c = mlp_o(context)
// into zc[:, t+1, Hd:]. Buffer conventions as las_dec_fwd; no cell / output-layer operands are needed.
int las_att_step(const las_dec_args* a, int t, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = check_args(a)) return rc;
  LAS_REQUIRE(t >= 0 && t < a->L, "attention step %d outside [0, %d)", t, a->L);
  return dec_fwd_steps(a, stream, t, t + 1, false, false);
}

static int dec_fwd_steps(const las_dec_args* a, cudaStream_t stream, int t_begin, int t_end, bool with_cell, bool with_out) {
  const int B = a->B, L = a->L, Te = a->Te, Hd = a->Hd, O = a->O, A = a->A, V = a->V, E = a->E;
  const int ZC = Hd + O;
  const int64_t R = L + 1;  // rows per utterance in the per-step buffers
  const int CM = (a->C + 3) / 4 * 4;       // channels padded to whole float4 pieces
  const int ethreads = (A + 31) / 32 * 32, ewarps = ethreads / 32;
  const size_t esmem = energy_smem(a, CM, ewarps, false);
  if (CM == 4) { if (int rc = ensure_smem(att_energy_fwd_kernel<4>, esmem)) return rc; }
  else if (CM == 8) { if (int rc = ensure_smem(att_energy_fwd_kernel<8>, esmem)) return rc; }
  else if (CM == 12) { if (int rc = ensure_smem(att_energy_fwd_kernel<12>, esmem)) return rc; }
  else { if (int rc = ensure_smem(att_energy_fwd_kernel<16>, esmem)) return rc; }
  const size_t csmem = (Te + 2 + 8 + 8 * 32 * 2) * sizeof(float);
  if (int rc = ensure_smem(att_ctx_fwd_kernel, csmem)) return rc;
  const bool free_run = a->mode != 0 && with_out;
  const bool drop = a->drop_p > 0.f && with_cell;
  if (drop) LAS_REQUIRE(a->seed_dev && a->zcd, "decoder: dropout needs seed_dev and the zcd buffer");
  if (a->mode == 1 && a->sample) LAS_REQUIRE(a->seed_dev, "decoder: sampling needs seed_dev");
  if (a->mode == 1 && a->tok_teacher) LAS_REQUIRE(a->tf_mask, "decoder: scheduled sampling needs tf_mask");
  __nv_bfloat16* zcd = static_cast<__nv_bfloat16*>(a->zcd);
  __nv_bfloat16* zc = static_cast<__nv_bfloat16*>(a->zc);
  __nv_bfloat16* ctx = static_cast<__nv_bfloat16*>(a->ctx);
  __nv_bfloat16* emb_op = static_cast<__nv_bfloat16*>(a->emb_op);
  const int Ep = (E + 15) / 16 * 16;

  CellFwdParams cp = {};
  cp.B = B; cp.T = L; cp.H = Hd; cp.ndir = 1; cp.UG = Hd / 8;
  cp.a1 = static_cast<const uint32_t*>(a->wr_pk); cp.KT1 = (ZC + 15) / 16; cp.a_dir = 0;
  cp.v1_ld = R * ZC; cp.v1_dir = 0; cp.hout_ld = R * ZC; cp.hout_dir = 0;
  cp.lens = nullptr; cp.c_state = a->c_state; cp.y = nullptr; cp.hprev = nullptr;
  cp.gates_save = static_cast<__half*>(a->gates_save); cp.c_save = a->c_save; cp.rep_row = 0;
  if (free_run) {
    cp.xproj = a->cell_bias; cp.xp_ld_b = 0; cp.xp_ld_t = 0; cp.xp_ld_dir = 0;
    cp.a2 = static_cast<const uint32_t*>(a->we_pk); cp.KT2 = Ep / 16; cp.v2_ld = R * Ep;
  } else {
    cp.xproj = a->embx; cp.xp_ld_b = R * 4 * Hd; cp.xp_ld_t = 4 * Hd; cp.xp_ld_dir = 0;
    cp.a2 = nullptr; cp.v2 = nullptr; cp.KT2 = 0;
  }

  EnergyFwdParams ep = {};
  ep.g = {B, Te, A, a->C, a->K, a->H};
  ep.P = a->P; ep.dz_ld = static_cast<int64_t>(L) * A; ep.w_ld = R * Te;
  ep.conv_w = a->conv_w; ep.mlp_att = a->mlp_att; ep.gvec = a->gvec; ep.e = a->e_buf;
  CtxFwdParams xp = {};
  xp.B = B; xp.Te = Te; xp.H = a->H; xp.scaling = a->att_scaling; xp.e = a->e_buf;
  xp.enc_h = static_cast<const __nv_bfloat16*>(a->enc_h); xp.w_ld = R * Te; xp.ctx_ld = R * a->H;

  const dim3 egrid((Te + kTT - 1) / kTT, B), cgrid(B, (a->H + 63) / 64);
  // programmatic dependent launches along the chain (common.cuh: pdl_*); the first launch of the call follows whatever
  // the caller enqueued (other kernels, copies, event waits) and keeps full stream order
  const bool pdl_on = pdl_enabled();
  bool pdl = false;
  for (int t = t_begin; t < t_end; ++t) {
    // (1) LSTMCell: gates = W [emb; c_{t-1}; z_{t-1}] + b   (model.py:284-286)
    cp.step = t;
    cp.v1 = (drop ? zcd : zc) + static_cast<int64_t>(t) * ZC;      // cell input: c_{t-1} after dropout (model.py:285)
    cp.hout = zc + static_cast<int64_t>(t + 1) * ZC;
    if (free_run) cp.v2 = emb_op + static_cast<int64_t>(t) * Ep;
    if (with_cell) { launch_cell_fwd(cp, stream, pdl); pdl = pdl_on; }
    // (2) decoder-state projection mlp_dec(z_t)   (model.py:163)
    float* dz_t = a->dzf + static_cast<int64_t>(t) * A;
    smallmm(static_cast<const uint32_t*>(a->mlp_dec_pk), A, Hd, zc + static_cast<int64_t>(t + 1) * ZC, 0, R * ZC, B,
            nullptr, nullptr, 0, dz_t, static_cast<int64_t>(L) * A, nullptr, 0, stream, pdl);
    pdl = pdl_on;
    // (3) energies e = gvec . tanh(P + dz + mlp_att(conv(w_{t-1})))   (model.py:156-165)
    ep.dz = dz_t;
    ep.wprev = a->ws + static_cast<int64_t>(t) * Te;
    ep.conv_save = a->conv_save ? a->conv_save + static_cast<int64_t>(t) * Te * 16 : nullptr;
    ep.cs_ld = static_cast<int64_t>(L) * Te * 16;
    if (CM == 4) LAS_CUDA(launch_k(att_energy_fwd_kernel<4>, egrid, dim3(ethreads), esmem, stream, pdl, ep));
    else if (CM == 8) LAS_CUDA(launch_k(att_energy_fwd_kernel<8>, egrid, dim3(ethreads), esmem, stream, pdl, ep));
    else if (CM == 12) LAS_CUDA(launch_k(att_energy_fwd_kernel<12>, egrid, dim3(ethreads), esmem, stream, pdl, ep));
    else LAS_CUDA(launch_k(att_energy_fwd_kernel<16>, egrid, dim3(ethreads), esmem, stream, pdl, ep));
    ++g_launches;
    // (4) w = softmax(scaling * e) over all Te, context = w @ enc_h   (model.py:167-171)
    xp.w_out = a->ws + static_cast<int64_t>(t + 1) * Te;
    xp.ctx = ctx + static_cast<int64_t>(t + 1) * a->H;
    LAS_CUDA(launch_k(att_ctx_fwd_kernel, cgrid, dim3(256), csmem, stream, pdl, xp)); ++g_launches;
    // (5) c_t = mlp_o(context)   (model.py:172) -> second half of zc row t+1
    smallmm(static_cast<const uint32_t*>(a->mlp_o_pk), O, a->H, ctx + static_cast<int64_t>(t + 1) * a->H, 0, R * a->H, B,
            a->mlp_o_b, nullptr, 0, nullptr, 0, zc + static_cast<int64_t>(t + 1) * ZC + Hd, R * ZC, stream, pdl);
    if (drop && !free_run) {    // (free-running modes: done inside next_emb_kernel below)
      DropRowParams dp = {};
      dp.B = B; dp.Hd = Hd; dp.O = O; dp.R = R; dp.row = t + 1; dp.p = a->drop_p;
      dp.seed_dev = static_cast<const unsigned long long*>(a->seed_dev); dp.site = a->drop_site;
      dp.zc = zc; dp.zcd = zcd;
      LAS_CUDA(launch_k(dec_drop_fwd_kernel, dim3((B * ZC + 255) / 256), dim3(256), 0, stream, pdl, dp)); ++g_launches;
    }
    if (free_run) {
      // (6) logit_t = output_layer([z_t; c_t]) (model.py:290-293), (7) next input embedding
      float* lg = a->logits + static_cast<int64_t>(t + 1) * V;
      smallmm(static_cast<const uint32_t*>(a->out_pk), V, ZC, zc + static_cast<int64_t>(t + 1) * ZC, 0, R * ZC, B,
              a->out_b, nullptr, 0, lg, R * V, nullptr, 0, stream, pdl);
      NextEmbParams np = {};
      np.drop_p = a->drop_p; np.seed_dev = static_cast<const unsigned long long*>(a->seed_dev);
      np.site = a->drop_site + 1; np.R = R; np.row = t + 1;
      np.B = B; np.V = V; np.E = E; np.logits = lg; np.lg_ld = R * V; np.emb_w = a->emb_w;
      np.scaling = a->smooth_scaling; np.smooth = (a->mode == 2);
      np.pred = a->pred + t; np.pred_ld = L;
      np.teacher = a->mode == 1 ? a->tok_teacher : nullptr; np.tf_mask = a->tf_mask;
      np.sample = a->mode == 1 ? a->sample : 0; np.sample_site = a->drop_site + 7;
      np.sample_seed = static_cast<const unsigned long long*>(a->seed_dev);
      np.emb_out = emb_op + static_cast<int64_t>(t + 1) * Ep; np.eo_ld = R * Ep;
      if (drop) { np.zc = zc; np.zcd = zcd; np.Hd = Hd; np.O = O; np.zc_site = a->drop_site; }
      LAS_CUDA(launch_k(next_emb_kernel, dim3(B), dim3(128), V * sizeof(float), stream, pdl, np)); ++g_launches;
    }
  }
  LAS_LAUNCH_CHECK();
  return 0;
}

int las_dec_bwd(const las_dec_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = check_args(a)) return rc;
  if (dec_persist_supported(a)) {
    ++g_path[LAS_PATH_DEC_PERSIST_BWD];
    return dec_persist_bwd(a, stream);   // one cluster-persistent launch
  }
  ++g_path[LAS_PATH_DEC_STEP_BWD];
  const bool smooth = a->mode == 2;
  const bool drop = a->drop_p > 0.f;
  if (drop) LAS_REQUIRE(a->seed_dev, "decoder backward: dropout needs seed_dev");
  if (smooth)
    LAS_REQUIRE(a->weT_pk && a->outT_pk && a->dlogits && a->dl_tot && a->demb_buf && a->logits && a->emb_w,
                "decoder backward (smooth free-run): missing buffers");
  const int B = a->B, L = a->L, Te = a->Te, Hd = a->Hd, O = a->O, A = a->A;
  const int ZC = Hd + O;
  const int64_t R = L + 1;
  const int CM = (a->C + 3) / 4 * 4;
  const int ethreads = (A + 31) / 32 * 32, ewarps = ethreads / 32;
  const int Ap = ethreads;
  const size_t esmem = energy_smem(a, CM, ewarps, true);
  if (CM == 4) { if (int rc = ensure_smem(att_energy_bwd_kernel<4>, esmem)) return rc; }
  else if (CM == 8) { if (int rc = ensure_smem(att_energy_bwd_kernel<8>, esmem)) return rc; }
  else if (CM == 12) { if (int rc = ensure_smem(att_energy_bwd_kernel<12>, esmem)) return rc; }
  else { if (int rc = ensure_smem(att_energy_bwd_kernel<16>, esmem)) return rc; }
  const int ksz = 2 * a->K + 1;
  const size_t dsmem = (a->H + static_cast<size_t>(a->C) * ksz) * sizeof(float);
  if (int rc = ensure_smem(att_dw_kernel, dsmem)) return rc;
  const dim3 egrid((Te + kTT - 1) / kTT, B);
  const int ncta = egrid.x * egrid.y;
  // lean energy backward (att_energy_bwd_mma_kernel): the caller saved the conv features in the forward, wants the
  // energy gradients in de_all, and produces dP and the energy-MLP parameter sums itself after this call
  // (las_att_param_grads_part), as it does for the cluster-persistent backward
  const bool lean = a->conv_save != nullptr && a->de_all != nullptr && las_att_bwd_lean_supported(A, a->C) != 0;
  EnergyLeanParams lp = {};
  if (lean) {
    float* pk = a->att_part + att_scratch_main_floats(B, L, Te, A, a->C, a->K);
    uint4* mattB = reinterpret_cast<uint4*>(pk);
    uint2* mattB2 = reinterpret_cast<uint2*>(pk + static_cast<int64_t>(A / 8) * 32 * 4);
    att_pack_matt_kernel<<<(A / 8 * 32 + A / 16 * 64 + 255) / 256, 256, 0, stream>>>(a->mlp_att, A, a->C, mattB, mattB2); ++g_launches;
    lp.B = B; lp.L = L; lp.Te = Te; lp.A = A; lp.C = a->C; lp.scaling = a->att_scaling;
    lp.P = a->P; lp.dz_ld = static_cast<int64_t>(L) * A; lp.w_ld = R * Te; lp.dw = a->dw_buf;
    lp.conv_save = a->conv_save; lp.mattB = mattB; lp.mattB2 = mattB2; lp.gvec = a->gvec;
    lp.ddz_ld = R * A; lp.de_all = a->de_all;
  } else {
    LAS_CUDA(cudaMemsetAsync(a->att_part, 0, static_cast<size_t>(ncta) * (CM + 1) * Ap * sizeof(float), stream));
  }
  LAS_CUDA(cudaMemsetAsync(a->dc_state, 0, static_cast<size_t>(B) * Hd * sizeof(float), stream));

  __nv_bfloat16* dgates = static_cast<__nv_bfloat16*>(a->dgates);
  __nv_bfloat16* dcz_all = static_cast<__nv_bfloat16*>(a->dcz_all);

  DwParams wp = {};
  wp.g = {B, Te, A, a->C, a->K, a->H};
  wp.dctx_ld = static_cast<int64_t>(L) * a->H;
  wp.enc_h = static_cast<const __nv_bfloat16*>(a->enc_h); wp.conv_w = a->conv_w; wp.dw = a->dw_buf;
  EnergyBwdParams ep = {};
  ep.g = wp.g; ep.scaling = a->att_scaling; ep.P = a->P; ep.dz_ld = static_cast<int64_t>(L) * A;
  ep.w_ld = R * Te; ep.dw = a->dw_buf; ep.conv_w = a->conv_w; ep.mlp_att = a->mlp_att; ep.gvec = a->gvec;
  ep.dP = a->dP; ep.ddz_ld = R * A; ep.part = a->att_part; ep.Ap = Ap;
  CellBwdParams cb = {};
  cb.dy = nullptr; cb.rep_row = 0;
  cb.dh_extra = a->dcz_tot; cb.dhx_ld = ZC;
  cb.a_pk = static_cast<const uint32_t*>(a->mlp_decT_pk); cb.a_dir = 0; cb.KT = (A + 15) / 16;
  cb.v_f32 = 1; cb.v_ld = R * A; cb.v_dir = 0; cb.v_ld_t = A; cb.v_t_rev = 0;
  cb.lens = nullptr; cb.gates_save = static_cast<const __half*>(a->gates_save); cb.c_save = a->c_save;
  cb.dG = dgates; cb.dg_ld_b = R * 4 * Hd; cb.dg_ld_t = 4 * Hd;
  cb.dc_state = a->dc_state; cb.B = B; cb.T = L; cb.H = Hd; cb.ndir = 1;

  const int V = a->V, E = a->E, Ep = (E + 15) / 16 * 16;
  const int64_t Vq = (V + 3) / 4 * 4;
  // programmatic dependent launches along the chain (see dec_fwd_steps); the launch that follows a memset / memcpy node
  // keeps full stream order
  const bool pdl_on = pdl_enabled();
  bool pdl = false;
  for (int t = L - 1; t >= 0; --t) {
    // In smooth mode with dropout both products of dgates_{t+1} -- W_ih[:, :E]^T (the input-embedding gradient) and
    // Wr^T (the state gradient) -- go out as ONE launch.
    const bool paired_mm = smooth && drop && t + 1 < L;
    if (paired_mm) {
      smallmm_pair(static_cast<const uint32_t*>(a->weT_pk), Ep, a->demb_buf, Ep, static_cast<const uint32_t*>(a->wrT_pk), ZC,
                   a->dcz_tot, ZC, 4 * Hd, dgates + static_cast<int64_t>(t + 1) * 4 * Hd, R * 4 * Hd, B, stream, pdl);
      pdl = pdl_on;
    }
    if (smooth) {
      // (0) free-running smooth mode: the gradient of logit_t also arrives through emb_{t+1}; only then
      //     is the output layer's contribution to d[z_t; c_t] known
      float* dlt = a->dl_tot + static_cast<int64_t>(t + 1) * Vq;   // rows padded to Vq floats: 8-byte aligned operand loads
      if (t + 1 < L) {
        if (!paired_mm) {
          smallmm(static_cast<const uint32_t*>(a->weT_pk), Ep, 4 * Hd, dgates + static_cast<int64_t>(t + 1) * 4 * Hd, 0,
                  R * 4 * Hd, B, nullptr, nullptr, 0, a->demb_buf, Ep, nullptr, 0, stream, pdl);
          pdl = pdl_on;
        }
        SmoothBwdParams sp = {};
        sp.drop_p = a->drop_p; sp.seed_dev = static_cast<const unsigned long long*>(a->seed_dev);
        sp.site = a->drop_site + 1; sp.R = R; sp.row = t + 1;
        sp.B = B; sp.V = V; sp.E = E; sp.scaling = a->smooth_scaling;
        sp.logits = a->logits + static_cast<int64_t>(t + 1) * V; sp.lg_ld = R * V;
        sp.demb = a->demb_buf; sp.de_ld = Ep; sp.emb_w = a->emb_w;
        sp.dlogits = a->dlogits + static_cast<int64_t>(t + 1) * V; sp.dl_ld = R * V;
        sp.dl_tot = dlt; sp.dt_ld = R * Vq;
        LAS_CUDA(launch_k(smooth_dlogit_kernel, dim3(B), dim3(128), (E + V) * sizeof(float), stream, pdl, sp)); ++g_launches;
        pdl = pdl_on;
      } else {
        LAS_CUDA(cudaMemcpy2DAsync(dlt, R * Vq * sizeof(float), a->dlogits + static_cast<int64_t>(t + 1) * V,
                                   R * V * sizeof(float), V * sizeof(float), B, cudaMemcpyDeviceToDevice, stream));
        pdl = false;          // a copy node in front of the next kernel
      }
      smallmm(static_cast<const uint32_t*>(a->outT_pk), ZC, V, dlt, 1, R * Vq, B, nullptr, nullptr, 0,
              const_cast<float*>(a->dzc_all) + static_cast<int64_t>(t + 1) * ZC, R * ZC, nullptr, 0, stream, pdl);
      pdl = pdl_on;
    }
    // (1) d[z_t; c_t] = dzc_all (from the output layer) + Wr^T dgates_{t+1}  (row L of dgates is zero)
    if (!drop) {
      smallmm(static_cast<const uint32_t*>(a->wrT_pk), ZC, 4 * Hd, dgates + static_cast<int64_t>(t + 1) * 4 * Hd, 0, R * 4 * Hd, B,
              nullptr, a->dzc_all + static_cast<int64_t>(t + 1) * ZC, R * ZC, a->dcz_tot, ZC,
              dcz_all + static_cast<int64_t>(t + 1) * ZC, R * ZC, stream, pdl);
      pdl = pdl_on;
    } else {
      // the path through the cell input of step t+1 carries that step's dropout mask on its c part
      if (!paired_mm) {
        smallmm(static_cast<const uint32_t*>(a->wrT_pk), ZC, 4 * Hd, dgates + static_cast<int64_t>(t + 1) * 4 * Hd, 0, R * 4 * Hd, B,
                nullptr, nullptr, 0, a->dcz_tot, ZC, nullptr, 0, stream, pdl);
        pdl = pdl_on;
      }
      DropRowParams dp = {};
      dp.B = B; dp.Hd = Hd; dp.O = O; dp.R = R; dp.row = t + 1; dp.p = a->drop_p;
      dp.seed_dev = static_cast<const unsigned long long*>(a->seed_dev); dp.site = a->drop_site;
      dp.dcz_tot = a->dcz_tot; dp.dzc_row = a->dzc_all + static_cast<int64_t>(t + 1) * ZC;
      dp.dcz_row = dcz_all + static_cast<int64_t>(t + 1) * ZC; dp.d_ld = R * ZC;
      LAS_CUDA(launch_k(dec_drop_bwd_kernel, dim3((B * ZC + 255) / 256), dim3(256), 0, stream, pdl, dp)); ++g_launches;
      pdl = pdl_on;
    }
    // (2) dcontext = mlp_o^T dc_t
    float* dctx_t = a->dctx_all + static_cast<int64_t>(t) * a->H;
    smallmm(static_cast<const uint32_t*>(a->mlp_oT_pk), a->H, O, dcz_all + static_cast<int64_t>(t + 1) * ZC + Hd, 0, R * ZC, B,
            nullptr, nullptr, 0, dctx_t, static_cast<int64_t>(L) * a->H, nullptr, 0, stream, pdl);
    pdl = pdl_on;
    // (3) dw_t = <dcontext, enc_h> + conv-input gradient of step t+1
    wp.dctx = dctx_t;
    wp.dattc_next = (t + 1 < L) ? a->dattc_all + static_cast<int64_t>(t + 1) * B * Te * a->C : nullptr;
    LAS_CUDA(launch_k(att_dw_kernel, egrid, dim3(256), dsmem, stream, pdl, wp)); ++g_launches;
    // (4) softmax + energy backward (tanh recomputed)
    ep.dz = a->dzf + static_cast<int64_t>(t) * A;
    ep.wprev = a->ws + static_cast<int64_t>(t) * Te;
    ep.wcur = a->ws + static_cast<int64_t>(t + 1) * Te;
    ep.ddz = a->ddz_all + static_cast<int64_t>(t + 1) * A;
    ep.dattc = a->dattc_all + static_cast<int64_t>(t) * B * Te * a->C;
    if (lean) {
      lp.t = t; lp.dz = ep.dz; lp.wcur = ep.wcur; lp.ddz = ep.ddz; lp.dattc = ep.dattc;
      const int lw = A / 16 < kLeanWarps ? A / 16 : kLeanWarps;
      LAS_CUDA(launch_k(att_energy_bwd_mma_kernel, dim3((Te + 15) / 16, B), dim3(32 * lw), 0, stream, pdl, lp));
    }
    else if (CM == 4) LAS_CUDA(launch_k(att_energy_bwd_kernel<4>, egrid, dim3(ethreads), esmem, stream, pdl, ep));
    else if (CM == 8) LAS_CUDA(launch_k(att_energy_bwd_kernel<8>, egrid, dim3(ethreads), esmem, stream, pdl, ep));
    else if (CM == 12) LAS_CUDA(launch_k(att_energy_bwd_kernel<12>, egrid, dim3(ethreads), esmem, stream, pdl, ep));
    else LAS_CUDA(launch_k(att_energy_bwd_kernel<16>, egrid, dim3(ethreads), esmem, stream, pdl, ep));
    ++g_launches;
    // (5) dz_t += mlp_dec^T ddz ; LSTMCell backward -> dgates_t
    cb.step = L - 1 - t;
    cb.v = a->ddz_all; cb.v_t_fwd = t + 1;
    launch_cell_bwd(cb, stream, pdl);
  }
  // reductions that were deferred out of the loop
  if (!lean) {
    att_part_reduce_kernel<<<((a->C + 1) * A + 255) / 256, 256, 0, stream>>>(a->att_part, ncta, CM, Ap, A, a->C,
                                                                               a->dmlp_att, a->dgvec); ++g_launches;
  }
  if (int rc = launch_att_dconv(a->dattc_all, a->ws, L, B, Te, a->C, a->K, a->dconv_w, a->att_part, stream)) return rc;
  att_denc_kernel<<<dim3((Te + kDE - 1) / kDE, B), (a->H <= 256 ? 256 : 320), 0, stream>>>(a->ws, a->dctx_all, L, B, Te, a->H, a->denc, a->denc_accumulate); ++g_launches;
  LAS_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
