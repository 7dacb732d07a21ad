// Shared device/host helpers for the sm_100a LAS kernels.
// Everything here is inline PTX or plain CUDA; no CUTLASS dependency.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

namespace las {

// ----------------------------------------------------------------------------------------
// error plumbing (C-ABI returns int; message retrievable with las_last_error())
// ----------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern unsigned long long g_launches;  // kernels launched by this library (las_launch_count())
extern unsigned long long g_path[];     // calls per code path (las_path_counters(); indices LAS_PATH_* in las_b200.h)
int check_cuda(cudaError_t e, const char* what);

#define LAS_CUDA(expr)                                   \
  do {                                                   \
    int _rc = ::las::check_cuda((expr), #expr);          \
    if (_rc) return _rc;                                 \
  } while (0)

#define LAS_REQUIRE(cond, ...)                           \
  do {                                                   \
    if (!(cond)) {                                       \
      ::las::set_error(__VA_ARGS__);                     \
      return 2;                                          \
    }                                                    \
  } while (0)

#define LAS_LAUNCH_CHECK() LAS_CUDA(cudaGetLastError())

#ifdef __CUDACC__

// ----------------------------------------------------------------------------------------
// small math
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Accurate variants (per-timestep kernels, decoder cell; the persistent BLSTM forward uses the MUFU.TANH forms, see
// persist_lstm_fwd for the parity measurement).
// ex2.approx + rcp.approx (~2 + 1 ulp): no IEEE-division slow path on the serial critical path.
__device__ __forceinline__ float sigmoid_acc(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_acc(float x) {
  // tanh(x) = 1 - 2/(exp(2x)+1); saturates cleanly (e = inf -> 1, e = 0 -> -1). The clamp keeps
  // e + 1 below 2^126, where __fdividef would flush the quotient to zero anyway.
  float e = __expf(2.0f * fminf(x, 40.0f));
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}

// Bare MUFU.EX2 + MUFU.RCP forms (4 / 5 instructions, ~1e-7 absolute error, clean saturation: ex2 -> inf gives
// rcp -> 0, ex2 -> 0 gives rcp(1) = 1): the gate activations of the persistent BLSTM forward.
__device__ __forceinline__ float sigmoid_er(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
__device__ __forceinline__ float tanh_er(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-2.8853900817779268f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return fmaf(2.0f, r, -1.0f);
}

// Programmatic dependent launch (the per-timestep decoder chain: ~19 small dependent kernels per decoder step). A kernel
// launched with launch_k(.., pdl = true, ..) may be scheduled while its predecessor in the stream is still running; it
// must call pdl_wait() before it touches anything the predecessor writes (the wait returns when the predecessor has
// completed and its memory is visible). Here every such kernel waits as its FIRST statement and then allows its own
// dependent to be scheduled, so at most two kernels of the chain are in flight and no kernel reads or writes memory
// before everything in front of it in the stream has finished; what overlaps is launch latency and CTA ramp-up. In a
// kernel launched without the attribute both calls are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// (Triggering BEFORE the wait lets the whole chain pile up on the SMs, each grid waiting on the one in front of it:
// measured slower, config 4 20.3 against 19.1 ms.)
__device__ __forceinline__ void pdl_enter() { pdl_wait(); pdl_launch_dependents(); }

// argmax with torch.argmax's conventions: first maximal index; NaN counts as the maximum (first NaN wins). Always
// returns an index in [0, V): an all-NaN row (a diverged model) must not turn into an out-of-range token id.
__device__ __forceinline__ bool argmax_better(float av, int ai, float bv, int bi) {
  const bool an = av != av, bn = bv != bv;
  if (an != bn) return an;
  if (an) return ai < bi;
  return av > bv || (av == bv && ai < bi);
}
__device__ __forceinline__ int warp_argmax(const float* __restrict__ x, int V, int lane, float* max_out) {
  float mx = -INFINITY;
  int amax = 0x7fffffff;
  for (int v = lane; v < V; v += 32) {
    const float xv = x[v];
    if (argmax_better(xv, v, mx, amax)) { mx = xv; amax = v; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, amax, o);
    if (argmax_better(om, oa, mx, amax)) { mx = om; amax = oa; }
  }
  if (max_out) *max_out = mx;
  return amax < V ? amax : 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// ----------------------------------------------------------------------------------------
// dropout: counter-based RNG (Philox4x32-10), so the backward pass recomputes the mask instead of
// storing it. Element `idx` of dropout site `site` at the step whose seed is `seed` is kept iff
// its uniform variate is >= p (torch.nn.Dropout semantics: kept values are scaled by 1/(1-p)).
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint2 key, uint4 ctr) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
// One Philox call decides 8 elements: element idx uses the 16-bit lane (idx & 7) of block (idx >> 3) and is kept iff
// that lane, as a fraction of 65536, is >= p (the rate is realised to 2^-16). Every mask on the path -- the vector and
// scalar dropout kernels, the decoder's in-kernel cell-input dropout, forward and backward -- goes through these helpers.
__device__ __forceinline__ uint32_t dropout_thresh16(float p) { return static_cast<uint32_t>(ceilf(p * 65536.0f)); }
// keep decisions of the 8 elements 8*blk .. 8*blk+7: bit i of the result = element 8*blk + i
__device__ __forceinline__ uint32_t dropout_keep8(unsigned long long seed, uint32_t site, unsigned long long blk, float p) {
  const uint4 r = philox4x32_10(make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)),
                                make_uint4(static_cast<uint32_t>(blk), static_cast<uint32_t>(blk >> 32), site, 0x4c41535fu));
  const uint32_t t = dropout_thresh16(p);
  return ((r.x & 0xffffu) >= t ? 1u : 0u) | ((r.x >> 16) >= t ? 2u : 0u) | ((r.y & 0xffffu) >= t ? 4u : 0u) |
         ((r.y >> 16) >= t ? 8u : 0u) | ((r.z & 0xffffu) >= t ? 16u : 0u) | ((r.z >> 16) >= t ? 32u : 0u) |
         ((r.w & 0xffffu) >= t ? 64u : 0u) | ((r.w >> 16) >= t ? 128u : 0u);
}
// keep decisions of the 4 elements 4*blk4 .. 4*blk4+3: bit i of the result = element 4*blk4 + i
__device__ __forceinline__ uint32_t dropout_keep4(unsigned long long seed, uint32_t site, unsigned long long blk4, float p) {
  return (dropout_keep8(seed, site, blk4 >> 1, p) >> (4u * (static_cast<uint32_t>(blk4) & 1u))) & 15u;
}
__device__ __forceinline__ bool dropout_keep(unsigned long long seed, uint32_t site, unsigned long long idx, float p) {
  return ((dropout_keep8(seed, site, idx >> 3, p) >> (static_cast<uint32_t>(idx) & 7u)) & 1u) != 0u;
}

// ----------------------------------------------------------------------------------------
// shared-memory addressing, mbarrier, fences
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap, not hang the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("las: mbarrier wait timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// ----------------------------------------------------------------------------------------
// distributed shared memory: remote stores that carry their own completion (st.async + mbarrier
// complete_tx). The payload and the transaction count land in the destination CTA together, so the
// consumer needs only a CTA-scope mbarrier wait: no cluster barrier, no fence.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_b32(uint32_t addr, uint32_t v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(addr), "r"(v), "r"(mbar) : "memory");
}
__device__ __forceinline__ void st_async_v2(uint32_t addr, uint32_t v0, uint32_t v1, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];"
               ::"r"(addr), "r"(v0), "r"(v1), "r"(mbar) : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint4 v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(mbar) : "memory");
}
// Bounded CTA-scope wait with a tag for the trap message (a protocol bug must trap, not hang the GPU box). The
// report lives in a cold, non-inlined function so that printf's argument buffer does not grow the callers' frames.
static __device__ __noinline__ void mbar_timeout(int tag) {
  printf("las: mbarrier %d wait timeout (block %d,%d thread %d)\n", tag, blockIdx.x, blockIdx.y, threadIdx.x);
  __trap();
}
__device__ __forceinline__ void mbar_wait_tag(uint64_t* bar, uint32_t parity, int tag) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) mbar_timeout(tag);
  }
}

// ----------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) 2D load, global -> shared, completes on an mbarrier
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar,
                                            int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)),
        "r"(c0), "r"(c1)
      : "memory");
}

// ----------------------------------------------------------------------------------------
// tcgen05: TMEM alloc, MMA, commit, load
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_out, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_out)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> f32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 columns of f32: thread i of the warp gets lane (32*(warp%4)+i), 32 consecutive columns.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.
//   start address  bits [0,14)  (>>4)
//   leading offset bits [16,30) (>>4)
//   stride offset  bits [32,46) (>>4)
//   version        bits [46,48) = 1
//   layout type    bits [61,64) = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes,
                                                         uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Instruction descriptor for kind::f16 with bf16 inputs and f32 accumulation.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major,
                                                       int b_mn_major) {
  return (1u << 4)                              // D format = F32
         | (1u << 7)                            // A format = BF16
         | (1u << 10)                           // B format = BF16
         | (static_cast<uint32_t>(a_mn_major) << 15)
         | (static_cast<uint32_t>(b_mn_major) << 16)
         | (static_cast<uint32_t>(N >> 3) << 17)
         | (static_cast<uint32_t>(M >> 4) << 24);
}

// ----------------------------------------------------------------------------------------
// legacy warp MMA (used inside the latency-bound recurrences, where weights live in registers)
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                               uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 "
      "{%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

#endif  // __CUDACC__

}  // namespace las
