// HBM-bound helper kernels around the GEMMs and recurrences: dtype conversion with padding,
// bias sums, ReLU backward, column sums (bias gradients), embedding gather / scatter-add,
// fused log-softmax + gather + unigram label smoothing (forward and backward), and the fused
// global-norm clip + Adam/AMSGrad step over a flat parameter buffer.
#include "common.cuh"
#include "las_internal.h"
#include "../../include/las_b200.h"

namespace las {

static inline int grid_for(int64_t n, int block = 256, int per_sm = 8) {
  int64_t b = (n + block - 1) / block;
  int64_t cap = static_cast<int64_t>(num_sms()) * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

// dst[r, c] = bf16(src[r, c]) for c < cols, 0 for cols <= c < ld_dst
__global__ void cvt_pad_bf16_kernel(const float* __restrict__ src, int64_t ld_src, int64_t rows,
                                    int cols, __nv_bfloat16* __restrict__ dst, int64_t ld_dst) {
  const int64_t pairs_per_row = ld_dst / 2;
  const int64_t total = rows * pairs_per_row;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / pairs_per_row;
    const int c = static_cast<int>(i % pairs_per_row) * 2;
    const float* s = src + r * ld_src;
    const float a = (c < cols) ? s[c] : 0.f;
    const float b = (c + 1 < cols) ? s[c + 1] : 0.f;
    reinterpret_cast<uint32_t*>(dst + r * ld_dst)[c / 2] = pack_bf16x2(a, b);
  }
}

__global__ void add2_kernel(const float* __restrict__ a, const float* __restrict__ b,
                            float* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = a[i] + (b ? b[i] : 0.f);
}

// dz = dout * (out > 0): ReLU backward; out_is_bf16 selects the saved activation's type.
__global__ void relu_bwd_kernel(const float* __restrict__ dout, const void* __restrict__ out,
                                int out_is_bf16, __nv_bfloat16* __restrict__ dz, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float o = out_is_bf16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(out)[i])
                                : static_cast<const float*>(out)[i];
    dz[i] = __float2bfloat16(o > 0.f ? dout[i] : 0.f);
  }
}

// the same, 4 elements per thread (n % 4 == 0, 16/8-byte aligned pointers, bf16 activations)
__global__ void relu_bwd4_kernel(const float4* __restrict__ dout, const uint2* __restrict__ out, uint2* __restrict__ dz,
                                 int64_t n4) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 d = __ldg(dout + i);
    const uint2 o = __ldg(out + i);
    const float2 a = unpack_bf16x2(o.x), b = unpack_bf16x2(o.y);
    uint2 r;
    r.x = pack_bf16x2(a.x > 0.f ? d.x : 0.f, a.y > 0.f ? d.y : 0.f);
    r.y = pack_bf16x2(b.x > 0.f ? d.z : 0.f, b.y > 0.f ? d.w : 0.f);
    dz[i] = r;
  }
}

// Column sums of a [rows, cols] matrix (bf16 or f32) into out[cols] (+=): each CTA reduces a
// 32-column x ROWS_PER_CTA-row slab, one atomicAdd per column per CTA.
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, int64_t ld, int64_t rows, int cols,
                              float* __restrict__ out, int rows_per_cta) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;  // 0..7
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_cta;
  const int64_t r1 = min(rows, r0 + rows_per_cta);
  float s = 0.f;
  if (c < cols) {
    for (int64_t r = r0 + ry; r < r1; r += 8) {
      if constexpr (sizeof(T) == 2) s += __bfloat162float(x[r * ld + c]);
      else s += x[r * ld + c];
    }
  }
  red[ry][threadIdx.x & 31] = s;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x & 31];
    atomicAdd(out + c, t);
  }
}

// out[i, :] = bf16(table[idx[i], :]), zero-padded to ld_out
__global__ void gather_rows_bf16_kernel(const float* __restrict__ table, int dim,
                                        const int64_t* __restrict__ idx, int64_t n,
                                        __nv_bfloat16* __restrict__ out, int64_t ld_out) {
  for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
    const float* src = table + idx[i] * dim;
    for (int c = threadIdx.x; c < ld_out; c += blockDim.x)
      out[i * ld_out + c] = __float2bfloat16(c < dim ? src[c] : 0.f);
  }
}

// dtable[idx[i], :] += d[i, :] except for idx == pad (nn.Embedding(padding_idx), model.py:261)
__global__ void scatter_add_rows_kernel(const float* __restrict__ d, int64_t ld, int dim,
                                        const int64_t* __restrict__ idx, int64_t n, int64_t pad,
                                        float* __restrict__ dtable) {
  for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
    const int64_t row = idx[i];
    if (row == pad) continue;
    for (int c = threadIdx.x; c < dim; c += blockDim.x) atomicAdd(dtable + row * dim + c, d[i * ld + c]);
  }
}

// ------------------------------------------------------------------------------------------
// log-softmax + gather + unigram label smoothing (model.py:354-366, 523-530), one warp per row
//   out_logp[r] = (1-ls) * logp[r, y_r] + ls * sum_v dist[v] * logp[r, v]      (ls = 0: plain gather)
//   out_prob[r] = softmax[r, y_r]                                               (optional)
//   out_pred[r] = argmax_v logits[r, v] (first maximal index)                   (optional)
// backward (given g = dL/d out_logp[r], gp = dL/d out_prob[r]):
//   dlogits[r, v] = g * ((1-ls) * (1[v=y] - p_v) + ls * (dist_v - p_v * sum(dist)))
//                 + gp * p_y * (1[v=y] - p_v)
// ------------------------------------------------------------------------------------------
__global__ void ce_ls_fwd_kernel(const float* __restrict__ logits, int64_t ld, int64_t ld_b,
                                 int64_t rows_per_b, int64_t rows, int V,
                                 const int64_t* __restrict__ targets, const float* __restrict__ dist,
                                 float ls, float* __restrict__ out_logp, float* __restrict__ out_prob,
                                 int64_t* __restrict__ out_pred) {
  const int lane = threadIdx.x & 31;
  const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* x = logits + (r / rows_per_b) * ld_b + (r % rows_per_b) * ld;
  float mx;
  const int amax = warp_argmax(x, V, lane, &mx);
  float se = 0.f;
  for (int v = lane; v < V; v += 32) se += expf(x[v] - mx);
  se = warp_sum(se);
  const float lse = mx + logf(se);
  float reg = 0.f;
  if (ls > 0.f) {
    for (int v = lane; v < V; v += 32) reg += dist[v] * (x[v] - lse);
    reg = warp_sum(reg);
  }
  if (lane == 0) {
    int64_t y = targets ? targets[r] : amax;
    const float lp = x[y] - lse;
    out_logp[r] = (ls > 0.f) ? (1.f - ls) * lp + ls * reg : lp;
    if (out_prob) out_prob[r] = expf(lp);
    if (out_pred) out_pred[r] = amax;
  }
}

__global__ void ce_ls_bwd_kernel(const float* __restrict__ logits, int64_t ld, int64_t ld_b,
                                 int64_t rows_per_b, int64_t rows, int V,
                                 const int64_t* __restrict__ targets, const float* __restrict__ dist,
                                 float ls, const float* __restrict__ g_logp,
                                 const float* __restrict__ g_prob, float* __restrict__ dlogits) {
  const int lane = threadIdx.x & 31;
  const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int64_t roff = (r / rows_per_b) * ld_b + (r % rows_per_b) * ld;
  const float* x = logits + roff;
  float mx = -INFINITY;
  for (int v = lane; v < V; v += 32) mx = fmaxf(mx, x[v]);
  mx = warp_max(mx);
  float se = 0.f, sd = 0.f;
  for (int v = lane; v < V; v += 32) {
    se += expf(x[v] - mx);
    if (ls > 0.f) sd += dist[v];
  }
  se = warp_sum(se);
  sd = warp_sum(sd);
  const float inv = 1.f / se;
  int64_t y;
  if (targets) {
    y = targets[r];
  } else {  // free-running decode gathers at the row's own argmax (first maximal index)
    y = warp_argmax(x, V, lane, nullptr);
  }
  const float g = g_logp ? g_logp[r] : 0.f;
  const float py = expf(x[y] - mx) * inv;
  const float gp = g_prob ? g_prob[r] * py : 0.f;
  for (int v = lane; v < V; v += 32) {
    const float pv = expf(x[v] - mx) * inv;
    const float ind = (v == y) ? 1.f : 0.f;
    float d = g * (1.f - ls) * (ind - pv) + gp * (ind - pv);
    if (ls > 0.f) d += g * ls * (dist[v] - pv * sd);
    dlogits[roff + v] = d;
  }
}

// ------------------------------------------------------------------------------------------
// fused global-norm clip + Adam / AMSGrad on a flat f32 buffer
// (torch.nn.utils.clip_grad_norm_ + torch.optim.Adam; solver.py:152-153, 171-173, 296-297, 384-385)
// ------------------------------------------------------------------------------------------
__global__ void sqnorm_partial_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ partials) {
  __shared__ double red[8];
  // 16-byte loads, four independent double accumulators per thread (fixed order: the result is deterministic)
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  const int64_t n4 = (reinterpret_cast<uintptr_t>(g) & 15) == 0 ? n >> 2 : 0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    s0 += static_cast<double>(v.x) * v.x; s1 += static_cast<double>(v.y) * v.y;
    s2 += static_cast<double>(v.z) * v.z; s3 += static_cast<double>(v.w) * v.w;
  }
  for (int64_t i = 4 * n4 + blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double v = g[i];
    s0 += v * v;
  }
  double s = (s0 + s1) + (s2 + s3);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < (blockDim.x >> 5); ++k) t += red[k];
    partials[blockIdx.x] = t;
  }
}

// Deterministic second stage: one warp sums the per-CTA partials in a fixed order.
__global__ void sqnorm_final_kernel(const double* __restrict__ partials, int np, float* __restrict__ out_norm) {
  double s = 0.0;
  for (int i = threadIdx.x; i < np; i += 32) s += partials[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) *out_norm = static_cast<float>(sqrt(s));
}

__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ g,
                                 float* __restrict__ m, float* __restrict__ v,
                                 float* __restrict__ vmax, int64_t n, float lr, float b1, float b2,
                                 float eps, float wd, const int32_t* __restrict__ step_dev, float max_norm,
                                 const float* __restrict__ norm_ptr, float grad_scale) {
  const float stepf = static_cast<float>(*step_dev);
  const float bc1 = 1.f - powf(b1, stepf);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, stepf));
  float coef = grad_scale;
  if (norm_ptr && max_norm > 0.f) {
    const float c = max_norm / (*norm_ptr * grad_scale + 1e-6f);
    coef *= fminf(c, 1.f);
  }
  const float step_size = lr / bc1;
  auto elem = [&](float& pi, float gi, float& mi, float& vi, float& vm) {
    gi *= coef;
    if (wd != 0.f) gi += wd * pi;
    mi = b1 * mi + (1.f - b1) * gi;
    vi = b2 * vi + (1.f - b2) * gi * gi;
    float denom;
    if (vmax) {
      vm = fmaxf(vm, vi);
      denom = sqrtf(vm) / bc2_sqrt + eps;
    } else {
      denom = sqrtf(vi) / bc2_sqrt + eps;
    }
    pi = pi - step_size * (mi / denom);
  };
  // 16-byte accesses over the aligned body (the flat buffers of optim.FusedAdam are 16-byte aligned), scalar tail
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(vmax)) & 15) == 0;
  const int64_t n4 = vec ? n >> 2 : 0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 x4 = vmax ? reinterpret_cast<float4*>(vmax)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    elem(p4.x, g4.x, m4.x, v4.x, x4.x);
    elem(p4.y, g4.y, m4.y, v4.y, x4.y);
    elem(p4.z, g4.z, m4.z, v4.z, x4.z);
    elem(p4.w, g4.w, m4.w, v4.w, x4.w);
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
    if (vmax) reinterpret_cast<float4*>(vmax)[i] = x4;
    reinterpret_cast<float4*>(p)[i] = p4;
  }
  for (int64_t i = 4 * n4 + blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float pi = p[i], mi = m[i], vi = v[i], vm = vmax ? vmax[i] : 0.f;
    elem(pi, g[i], mi, vi, vm);
    m[i] = mi;
    v[i] = vi;
    if (vmax) vmax[i] = vm;
    p[i] = pi;
  }
}

// x[b, t, c] *= keep(seed, site, (b*T + min(t, T-1))*W + c) / (1 - p), t < T + rep_row: the replicated row of
// an odd pyramid extent (model.py:88-89 pads AFTER the dropout of model.py:82) shares the mask of row T-1.
template <typename T>
__global__ void dropout_kernel(T* __restrict__ x, int64_t B, int64_t Tn, int W, int64_t ld_b, int64_t ld_t, int rep_row,
                               float p, const unsigned long long* __restrict__ seed_dev, uint32_t site) {
  const unsigned long long seed = *seed_dev;
  const float scale = 1.0f / (1.0f - p);
  const int64_t rows = Tn + rep_row, total = B * rows * W;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % W);
    const int64_t bt = i / W, t = bt % rows, b = bt / rows;
    const int64_t tm = t < Tn ? t : Tn - 1;
    T* q = x + b * ld_b + t * ld_t + c;
    const bool keep = dropout_keep(seed, site, static_cast<unsigned long long>((b * Tn + tm) * W + c), p);
    if constexpr (sizeof(T) == 2) *q = keep ? __float2bfloat16(__bfloat162float(*q) * scale) : __float2bfloat16(0.f);
    else *q = keep ? *q * scale : 0.f;
  }
}


}  // namespace las

using namespace las;

extern "C" {

int las_cvt_pad_bf16(const float* src, int64_t ld_src, int64_t rows, int cols, void* dst,
                     int64_t ld_dst, void* stream) {
  LAS_REQUIRE(ld_dst % 2 == 0 && ld_dst >= cols, "cvt_pad: bad ld_dst");
  if (rows == 0) return 0;
  cvt_pad_bf16_kernel<<<grid_for(rows * (ld_dst / 2)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, ld_src, rows, cols, static_cast<__nv_bfloat16*>(dst), ld_dst); ++g_launches;
  LAS_LAUNCH_CHECK();
  return 0;
}

int las_add2(const float* a, const float* b, float* out, int64_t n, void* stream) {
  if (n == 0) return 0;
  add2_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, out, n); ++g_launches;
  LAS_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
namespace las {
// W % 4 == 0 and 4-element-aligned rows: one thread = 4 consecutive elements = one Philox call, vector access.
template <typename T>
__global__ void dropout4_kernel(T* __restrict__ x, int64_t B, int64_t Tn, int W, int64_t ld_b, int64_t ld_t, int rep_row,
                                float p, const unsigned long long* __restrict__ seed_dev, uint32_t site) {
  const unsigned long long seed = *seed_dev;
  const float scale = 1.0f / (1.0f - p);
  const int W4 = W >> 2;
  const int64_t rows = Tn + rep_row, total = B * rows * W4;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int c4;
    int64_t t, b;
    if (total < (1ll << 31)) {      // 32-bit divisions: the 64-bit ones cost as much as the Philox rounds
      const uint32_t i32 = static_cast<uint32_t>(i), bt = i32 / static_cast<uint32_t>(W4);
      c4 = static_cast<int>(i32 - bt * static_cast<uint32_t>(W4));
      const uint32_t b32 = bt / static_cast<uint32_t>(rows);
      b = b32;
      t = bt - b32 * static_cast<uint32_t>(rows);
    } else {
      c4 = static_cast<int>(i % W4);
      const int64_t bt = i / W4;
      t = bt % rows;
      b = bt / rows;
    }
    const int64_t tm = t < Tn ? t : Tn - 1;
    const uint32_t k = dropout_keep4(seed, site, static_cast<unsigned long long>(((b * Tn + tm) * W + 4 * c4) >> 2), p);
    T* q = x + b * ld_b + t * ld_t + 4 * c4;
    if constexpr (sizeof(T) == 2) {
      uint2 v = *reinterpret_cast<uint2*>(q);
      const float2 a = unpack_bf16x2(v.x), c = unpack_bf16x2(v.y);
      v.x = pack_bf16x2((k & 1u) ? a.x * scale : 0.f, (k & 2u) ? a.y * scale : 0.f);
      v.y = pack_bf16x2((k & 4u) ? c.x * scale : 0.f, (k & 8u) ? c.y * scale : 0.f);
      *reinterpret_cast<uint2*>(q) = v;
    } else {
      float4 v = *reinterpret_cast<float4*>(q);
      v.x = (k & 1u) ? v.x * scale : 0.f; v.y = (k & 2u) ? v.y * scale : 0.f;
      v.z = (k & 4u) ? v.z * scale : 0.f; v.w = (k & 8u) ? v.w * scale : 0.f;
      *reinterpret_cast<float4*>(q) = v;
    }
  }
}
// W % 8 == 0 and 8-element-aligned rows: one thread = 8 consecutive elements = ONE Philox call, 16-byte accesses
// (the 4-wide kernel above spends two calls on them, and the Philox rounds are what this kernel's time is made of).
template <typename T>
__global__ void dropout8_kernel(T* __restrict__ x, int64_t B, int64_t Tn, int W, int64_t ld_b, int64_t ld_t, int rep_row,
                                float p, const unsigned long long* __restrict__ seed_dev, uint32_t site) {
  const unsigned long long seed = *seed_dev;
  const float scale = 1.0f / (1.0f - p);
  const int W8 = W >> 3;
  const int64_t rows = Tn + rep_row, total = B * rows * W8;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int c8;
    int64_t t, b;
    if (total < (1ll << 31)) {
      const uint32_t i32 = static_cast<uint32_t>(i), bt = i32 / static_cast<uint32_t>(W8);
      c8 = static_cast<int>(i32 - bt * static_cast<uint32_t>(W8));
      const uint32_t b32 = bt / static_cast<uint32_t>(rows);
      b = b32;
      t = bt - b32 * static_cast<uint32_t>(rows);
    } else {
      c8 = static_cast<int>(i % W8);
      const int64_t bt = i / W8;
      t = bt % rows;
      b = bt / rows;
    }
    const int64_t tm = t < Tn ? t : Tn - 1;
    const uint32_t k = dropout_keep8(seed, site, static_cast<unsigned long long>(((b * Tn + tm) * W + 8 * c8) >> 3), p);
    T* q = x + b * ld_b + t * ld_t + 8 * c8;
    if constexpr (sizeof(T) == 2) {
      uint4 v = *reinterpret_cast<uint4*>(q);
      uint32_t* w = &v.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = unpack_bf16x2(w[j]);
        w[j] = pack_bf16x2((k >> (2 * j)) & 1u ? a.x * scale : 0.f, (k >> (2 * j + 1)) & 1u ? a.y * scale : 0.f);
      }
      *reinterpret_cast<uint4*>(q) = v;
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float4 v = *reinterpret_cast<float4*>(q + 4 * h);
        const uint32_t kk = k >> (4 * h);
        v.x = (kk & 1u) ? v.x * scale : 0.f; v.y = (kk & 2u) ? v.y * scale : 0.f;
        v.z = (kk & 4u) ? v.z * scale : 0.f; v.w = (kk & 8u) ? v.w * scale : 0.f;
        *reinterpret_cast<float4*>(q + 4 * h) = v;
      }
    }
  }
}
}  // namespace las
extern "C" {

int las_dropout(void* x, int x_is_bf16, int64_t B, int64_t T, int W, int64_t ld_b, int64_t ld_t, int rep_row, float p,
                const void* seed_dev, uint32_t site, void* stream) {
  LAS_REQUIRE(p >= 0.f && p < 1.f, "dropout: rate %f out of range [0, 1)", p);
  if (p == 0.f || B * T * W == 0) return 0;
  const int64_t n = B * (T + rep_row) * W;
  const int esz = x_is_bf16 ? 2 : 4;
  if (W % 8 == 0 && ld_b % 8 == 0 && ld_t % 8 == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0) {
    if (x_is_bf16)
      dropout8_kernel<__nv_bfloat16><<<grid_for(n / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
          static_cast<__nv_bfloat16*>(x), B, T, W, ld_b, ld_t, rep_row, p, static_cast<const unsigned long long*>(seed_dev), site);
    else
      dropout8_kernel<float><<<grid_for(n / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
          static_cast<float*>(x), B, T, W, ld_b, ld_t, rep_row, p, static_cast<const unsigned long long*>(seed_dev), site);
    ++g_launches;
    LAS_LAUNCH_CHECK();
    return 0;
  }
  if (W % 4 == 0 && ld_b % 4 == 0 && ld_t % 4 == 0 && (reinterpret_cast<uintptr_t>(x) % (4 * esz)) == 0) {
    if (x_is_bf16)
      dropout4_kernel<__nv_bfloat16><<<grid_for(n / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
          static_cast<__nv_bfloat16*>(x), B, T, W, ld_b, ld_t, rep_row, p, static_cast<const unsigned long long*>(seed_dev), site);
    else
      dropout4_kernel<float><<<grid_for(n / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
          static_cast<float*>(x), B, T, W, ld_b, ld_t, rep_row, p, static_cast<const unsigned long long*>(seed_dev), site);
    ++g_launches;
    LAS_LAUNCH_CHECK();
    return 0;
  }
  if (x_is_bf16)
    dropout_kernel<__nv_bfloat16><<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<__nv_bfloat16*>(x), B, T, W, ld_b, ld_t, rep_row, p, static_cast<const unsigned long long*>(seed_dev), site);
  else
    dropout_kernel<float><<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<float*>(x), B, T, W, ld_b, ld_t, rep_row, p, static_cast<const unsigned long long*>(seed_dev), site);
  ++g_launches;
  LAS_LAUNCH_CHECK();
  return 0;
}

int las_relu_bwd(const float* dout, const void* out, int out_is_bf16, void* dz, int64_t n, void* stream) {
  if (n == 0) return 0;
  if (out_is_bf16 && n % 4 == 0 && reinterpret_cast<uintptr_t>(dout) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 8 == 0 &&
      reinterpret_cast<uintptr_t>(dz) % 8 == 0) {
    relu_bwd4_kernel<<<grid_for(n / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(dout), static_cast<const uint2*>(out), static_cast<uint2*>(dz), n / 4); ++g_launches;
    LAS_LAUNCH_CHECK();
    return 0;
  }
  relu_bwd_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dout, out, out_is_bf16, static_cast<__nv_bfloat16*>(dz), n); ++g_launches;
  LAS_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
namespace las {
// Vector path: a thread owns EPT = 16 bytes / sizeof(T) consecutive columns, a warp 32*EPT columns of one row (512
// contiguous bytes), the 8 warps of a CTA walk rows r0+w, r0+w+8, ...; one atomicAdd per column per CTA.
template <typename T>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const T* __restrict__ x, int64_t ld, int64_t rows, int cols,
                                                         float* __restrict__ out, int rows_per_cta) {
  constexpr int EPT = 16 / sizeof(T);
  __shared__ float red[8][32 * EPT + 1];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * 32 + lane) * EPT;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_cta;
  const int64_t r1 = min(rows, r0 + rows_per_cta);
  float s[EPT];
#pragma unroll
  for (int e = 0; e < EPT; ++e) s[e] = 0.f;
  if (c0 < cols) {      // cols % EPT == 0: the whole vector is in range
    for (int64_t r = r0 + w; r < r1; r += 8) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + r * ld + c0));
      if constexpr (sizeof(T) == 2) {
        const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
        s[0] += a.x; s[1] += a.y; s[2] += b.x; s[3] += b.y; s[4] += c.x; s[5] += c.y; s[6] += d.x; s[7] += d.y;
      } else {
        s[0] += __uint_as_float(v.x); s[1] += __uint_as_float(v.y); s[2] += __uint_as_float(v.z); s[3] += __uint_as_float(v.w);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < EPT; ++e) red[w][lane * EPT + e] = s[e];
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * EPT; i += 256) {
    const int c = blockIdx.x * 32 * EPT + i;
    if (c < cols) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += red[k][i];
      atomicAdd(out + c, t);
    }
  }
}
}  // namespace las
extern "C" {

int las_colsum(const void* x, int x_is_bf16, int64_t ld, int64_t rows, int cols, float* out, void* stream) {
  if (rows == 0 || cols == 0) return 0;
  const int ept = x_is_bf16 ? 8 : 4;
  if (cols % ept == 0 && ld % ept == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0) {
    const int xb = (cols + 32 * ept - 1) / (32 * ept);
    // ~4 CTAs per SM in total, at least 64 rows each
    int64_t chunks = (4LL * num_sms() + xb - 1) / xb;
    int64_t rpc = (rows + chunks - 1) / chunks;
    if (rpc < 64) rpc = 64;
    dim3 grid(xb, static_cast<unsigned>((rows + rpc - 1) / rpc));
    if (x_is_bf16)
      colsum_vec_kernel<__nv_bfloat16><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
          static_cast<const __nv_bfloat16*>(x), ld, rows, cols, out, static_cast<int>(rpc));
    else
      colsum_vec_kernel<float><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
          static_cast<const float*>(x), ld, rows, cols, out, static_cast<int>(rpc));
    ++g_launches;
    LAS_LAUNCH_CHECK();
    return 0;
  }
  const int rows_per_cta = 512;
  dim3 grid((cols + 31) / 32, static_cast<unsigned>((rows + rows_per_cta - 1) / rows_per_cta));
  if (x_is_bf16)
    colsum_kernel<__nv_bfloat16><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), ld, rows, cols, out, rows_per_cta);
  else
    colsum_kernel<float><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float*>(x), ld, rows, cols, out, rows_per_cta);
  ++g_launches;
  LAS_LAUNCH_CHECK();
  return 0;
}

int las_gather_rows_bf16(const float* table, int dim, const int64_t* idx, int64_t n, void* out,
                         int64_t ld_out, void* stream) {
  if (n == 0) return 0;
  gather_rows_bf16_kernel<<<static_cast<int>(n < 4096 ? n : 4096), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      table, dim, idx, n, static_cast<__nv_bfloat16*>(out), ld_out); ++g_launches;
  LAS_LAUNCH_CHECK();
  return 0;
}

int las_scatter_add_rows(const float* d, int64_t ld, int dim, const int64_t* idx, int64_t n, int64_t pad,
                         float* dtable, void* stream) {
  if (n == 0) return 0;
  scatter_add_rows_kernel<<<static_cast<int>(n < 4096 ? n : 4096), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      d, ld, dim, idx, n, pad, dtable); ++g_launches;
  LAS_LAUNCH_CHECK();
  return 0;
}

int las_ce_ls_fwd(const float* logits, int64_t ld, int64_t ld_b, int64_t rows_per_b, int64_t rows, int V,
                  const int64_t* targets, const float* dist, float ls, float* out_logp, float* out_prob,
                  int64_t* out_pred, void* stream) {
  if (rows == 0) return 0;
  const int wpb = 8;
  ce_ls_fwd_kernel<<<static_cast<unsigned>((rows + wpb - 1) / wpb), wpb * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, ld, ld_b, rows_per_b, rows, V, targets, dist, ls, out_logp, out_prob, out_pred); ++g_launches;
  LAS_LAUNCH_CHECK();
  return 0;
}

int las_ce_ls_bwd(const float* logits, int64_t ld, int64_t ld_b, int64_t rows_per_b, int64_t rows, int V,
                  const int64_t* targets, const float* dist, float ls, const float* g_logp,
                  const float* g_prob, float* dlogits, void* stream) {
  if (rows == 0) return 0;
  const int wpb = 8;
  ce_ls_bwd_kernel<<<static_cast<unsigned>((rows + wpb - 1) / wpb), wpb * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, ld, ld_b, rows_per_b, rows, V, targets, dist, ls, g_logp, g_prob, dlogits); ++g_launches;
  LAS_LAUNCH_CHECK();
  return 0;
}

__global__ void pyramid_lens_kernel(const int32_t* __restrict__ in, int B, int sub, int32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) out[i] = (in[i] + 1) / sub;
}

int las_pyramid_lens(const int32_t* lens_in, int B, int sub, int32_t* lens_out, void* stream) {
  if (B == 0) return 0;
  LAS_REQUIRE(sub >= 1, "pyramid_lens: subsample factor must be >= 1");
  pyramid_lens_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(lens_in, B, sub, lens_out); ++g_launches;
  LAS_LAUNCH_CHECK();
  return 0;
}

int las_grad_norm(const float* g, int64_t n, void* partials_ws, float* out_norm, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int np = 256;
  sqnorm_partial_kernel<<<np, 256, 0, stream>>>(g, n, static_cast<double*>(partials_ws)); ++g_launches;
  sqnorm_final_kernel<<<1, 32, 0, stream>>>(static_cast<const double*>(partials_ws), np, out_norm); ++g_launches;
  LAS_LAUNCH_CHECK();
  return 0;
}

int las_adam_step(float* p, const float* g, float* m, float* v, float* vmax, int64_t n, float lr,
                  float beta1, float beta2, float eps, float weight_decay, const int32_t* step_dev,
                  float max_norm, const float* norm_ptr, float grad_scale, void* stream) {
  if (n == 0) return 0;
  LAS_REQUIRE(step_dev != nullptr, "adam: step_dev must point to the device step counter");
  adam_step_kernel<<<grid_for((n + 3) / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, vmax, n, lr, beta1, beta2, eps, weight_decay, step_dev, max_norm, norm_ptr,
      grad_scale); ++g_launches;
  LAS_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
