// Persistent LSTM recurrence for the pyramidal BLSTM listener (model.py:67, 79-81): ONE launch runs
// all T timesteps of a layer. A thread-block cluster of CS CTAs serves one (direction, group of 8
// utterances); every CTA keeps its slice of the recurrent weights resident in REGISTERS as
// mma.m16n8k16 A fragments for the whole sequence, the cell state lives in registers, and the only
// per-step communication is the exchange of the new hidden state through distributed shared memory:
//
//   forward : all-gather of h_t. Each lane transposes its 8x8 (unit x utterance) tile with movmatrix
//             so that it holds exactly one B-fragment word of the NEXT step's MMA, and pushes it to
//             every CTA of the cluster with st.async (data + mbarrier complete_tx in one message; no
//             fence, no separate flag). A CTA starts step t+1 when its mbarrier has counted all bytes.
//   backward: reduce-scatter of dh_t. A CTA multiplies W_hh^T restricted to ITS OWN gate rows with
//             its own gate gradients (no gather needed), pushes the f32 partial sums to the CTAs that
//             own the units, and each owner adds the CS partials before the gate-derivative math.
//
// Timesteps are therefore synchronised by mbarrier transaction counts fed by remote stores, not by
// kernel boundaries: per step one DSMEM hop instead of a launch + L2 round trips.
#include <cooperative_groups.h>
#include <stdlib.h>
#include "common.cuh"
#include "las_internal.h"
#include "../../include/las_b200.h"

namespace cg = cooperative_groups;

namespace las {

namespace {

constexpr int kMaxUGC = 5;   // unit groups (8 hidden units) per CTA
constexpr int kMaxKTW = 10;  // k-tiles (16) per warp in the backward kernel (= 2 * kMaxUGC)
constexpr int kMaxKT = 20;   // k-tiles (16) of the forward kernel: the whole hidden size
constexpr int kNB = 8;       // utterances per cluster (one MMA n-tile)
constexpr int kDefaultAct = 2;   // gate activations of the forward kernel: 0 = __expf / __fdividef, 1 = MUFU.TANH, 2 = bare ex2 + rcp (LAS_FAST_ACT)
constexpr int kPF = 8;       // timesteps of global-memory prefetch distance (cp.async ring in shared memory)

struct Geom {
  int UG, CS, UGC, KT, KS, KTW, JT, MTW;
};

// forward kernel: Q quads of 4 hidden units, WPC quads (= warps) per CTA, CS CTAs per cluster
struct FGeom {
  int Q, CS, WPC, KT;
};
// Hidden sizes above 320 (the LM judge's 640, model.py:466) do not fit the register file of an 8..10-CTA cluster:
// "wide" geometry = 16-CTA clusters, and the recurrent weights beyond what the registers hold (forward: k-tiles
// kMaxKT.., backward: m-tiles 2..) stay resident in SHARED memory as ready-made A fragments.
constexpr int kWideKTS = 20;
constexpr int kWidePF = 4;    // backward: prefetch ring depth of the wide geometry (shared memory holds weights)   // forward: k-tiles per warp kept in shared memory (wide geometry)
bool fgeom_for(int H, FGeom& g, int max_cs = 8) {
  if (H % 8 != 0 || H < 8) return false;
  g.Q = H / 4;
  if (H > 16 * kMaxKT) max_cs = 16;
  const int cs0 = g.Q < max_cs ? g.Q : max_cs;
  g.WPC = (g.Q + cs0 - 1) / cs0;
  g.CS = (g.Q + g.WPC - 1) / g.WPC;
  g.KT = (H + 15) / 16;
  return g.WPC <= 12 && g.KT <= kMaxKT + kWideKTS;
}

bool geom_for(int H, Geom& g) {
  if (H % 8 != 0 || H < 8 || H > 8 * 16 * kMaxUGC) return false;
  g.UG = H / 8;
  const int csm = H > 8 * 8 * kMaxUGC ? 16 : 8;      // wide geometry above H = 320
  const int cs0 = g.UG < csm ? g.UG : csm;
  g.UGC = (g.UG + cs0 - 1) / cs0;
  g.CS = (g.UG + g.UGC - 1) / g.UGC;
  g.KT = (H + 15) / 16;
  g.KS = g.KT > kMaxKTW ? 2 : 1;
  g.KTW = (g.KT + g.KS - 1) / g.KS;
  g.JT = (H + 15) / 16;
  const int W = 2 * g.UGC;
  g.MTW = (g.JT + W - 1) / W;
  if (csm == 16) return g.MTW <= 4 && g.UGC <= kMaxUGC;   // m-tiles 2, 3 of a warp live in shared memory
  return g.MTW <= 2 && g.KTW <= kMaxKTW && g.UGC <= kMaxUGC;
}

// Asynchronous global->shared copies (LDGSTS): the per-timestep operands are prefetched kPF steps
// ahead without occupying registers or the load scoreboard (a register ring made every branch of
// the step loop wait for DRAM: profiles/r01_persist_v2_stalls.txt).
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
// The payload lands in THIS CTA's shared memory together with its complete_tx (st.async), so the
// default (CTA-scope) wait is enough; a cluster-scope acquire makes ptxas emit CCTL.IVALL (L1
// invalidate) after every wait, which also drains the global prefetch loads: 37 % of all stall
// samples in the first version of this kernel (profiles/r01_persist_v1_stalls.txt).
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap, not hang the GPU box.
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (++spins > (1u << 24)) {
      printf("las: persistent LSTM barrier timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}

// timesteps a cluster has to walk: the longest of its (up to) 8 utterances, T without a length vector
__device__ __forceinline__ int group_steps(const int32_t* lens, int grp, int B, int T, int walk_all) {
  if (lens == nullptr || walk_all) return T;
  int m = 0;
#pragma unroll
  for (int i = 0; i < kNB; ++i) {
    const int nn = grp * kNB + i;
    if (nn < B) m = max(m, min(__ldg(lens + nn), T));
  }
  return m;
}

struct FwdP {
  const float* xproj; int64_t xp_ld_b, xp_ld_t;   // gate-minor columns: [dir*4H + 4*u + gate]
  const uint32_t* whh_pk;      // [ndir][Q][KT][32][4] (las_pack_afrag mode 3: tile = 4 units x 4 gates, quad-permuted K)
  const int32_t* lens;
  __nv_bfloat16* y; int64_t y_ld_b, y_ld_t;
  __nv_bfloat16* hprev; int64_t hp_ld_b, hp_ld_t;
  uint4* rec;                  // [ndir][B][T][H] x 16 B: (i,f) f16x2 | (g,o) f16x2 | c f32 | tanh(c) f32
  int B, T, H, rep_row;
  int Q, WPC, KT;
  int walk_all;                // debugging switch LAS_LSTM_WALK_ALL=1: every cluster walks all T timesteps
  long long* dbg;              // optional clock64() trace of cluster (0,0,0) (las_set_debug_buffer)
};

// grid (CS, ceil(B/8), ndir), cluster (CS,1,1), block 32*WPC. Warp w of CTA `rank` owns the 4 hidden units of
// quad rank*WPC + w: its m16 tile holds their 16 gate rows, K = all of H, so a warp needs no partial-sum
// exchange with other warps and the step loop contains no block-wide barrier: warps are paced only by the
// mbarrier that counts the bytes of h_t arriving from the cluster.
template <int kAct, int KTS>
__global__ void __launch_bounds__(384, 1) lstm_persist_fwd_kernel(FwdP p) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);                    // [2]
  // h_t as B fragments, [2][KT][4 quads][8 utterances][2 words]: the K positions of a k-tile are permuted so that
  // quad q's four units are exactly the (b0, b1) words of the lanes with tig == q. A warp's output (4 units x 8
  // utterances) is then one contiguous 64-byte run, pushed to each peer as 16-byte st.async vectors. The utterance
  // index of quads 2, 3 is XORed with 4 so that a half-warp's 8-byte fragment loads hit 32 distinct banks.
  uint32_t* hs = reinterpret_cast<uint32_t*>(smem + 16);
  float4* xring = reinterpret_cast<float4*>(smem + 16 + 2 * p.KT * 256); // [kPF][threads]
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank(), CS = cluster.num_blocks();
  const int grp = blockIdx.y, dir = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  const int quad = rank * p.WPC + warp;
  const bool quad_ok = quad < p.Q;
  const uint32_t tx_bytes = static_cast<uint32_t>(p.H) * 16u;
  const int T = p.T, H = p.H;
  // A cluster only walks the timesteps its own 8 utterances have (Tg = the longest of them): past that every
  // element is inactive, the state is frozen and only zero rows of y are left to write (tail loop below). Batches
  // are sorted by length, so when there are more clusters than the device holds (B = 64: 16 against 15) the late
  // clusters are the short ones and the second wave ends with the first instead of doubling the launch.
  const int Tg = group_steps(p.lens, blockIdx.y, p.B, T, p.walk_all);

  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 2 * p.KT * 64; i += blockDim.x) hs[i] = 0u;
  __syncthreads();
  if (threadIdx.x == 0) {
    if (Tg > 1) mbar_arrive_expect_tx(&full[1], tx_bytes);   // h_0 -> consumed by step 1
    if (Tg > 2) mbar_arrive_expect_tx(&full[0], tx_bytes);   // h_1 -> consumed by step 2
  }
  cluster.sync();   // every CTA's barriers are armed before any remote store can arrive

  if (quad_ok) {
    // resident recurrent weights: A fragments of this warp's quad over the whole K range
    uint4 A[kMaxKT];
    // wide geometry: k-tiles kMaxKT .. kMaxKT + KTS - 1 of this warp in shared memory, [warp][KTS][32 lanes] x 16 B
    uint4* As = reinterpret_cast<uint4*>(smem + 16 + 2 * p.KT * 256 + kPF * blockDim.x * 16) + warp * KTS * 32 + lane;
    {
      const uint4* src = reinterpret_cast<const uint4*>(p.whh_pk) +
                         (static_cast<int64_t>(dir) * p.Q + quad) * p.KT * 32 + lane;
#pragma unroll
      for (int kt = 0; kt < kMaxKT; ++kt) A[kt] = kt < p.KT ? __ldg(src + kt * 32) : make_uint4(0u, 0u, 0u, 0u);
      if (KTS > 0) {
#pragma unroll 4
        for (int j = 0; j < KTS; ++j)
          As[j * 32] = kMaxKT + j < p.KT ? __ldg(src + (kMaxKT + j) * 32) : make_uint4(0u, 0u, 0u, 0u);
      }
    }
    // element owned by this lane after the in-warp gate exchange: unit u, utterance n
    const int gl = g >> 2, ul = g & 3;
    const int u = 4 * quad + ul;
    const int n_loc = 2 * tig + gl;
    const int n = grp * kNB + n_loc;
    const int len = n < p.B ? (p.lens ? p.lens[n] : T) : 0;
    float c_st = 0.f;
    uint32_t h_bits = 0u;
    uint2 quad_prev = make_uint2(0u, 0u);

    // xproj ring: slot s % kPF holds this lane's 4 gate pre-activations of step s (thread-private, one 16-byte copy)
    float4* xslot = xring + threadIdx.x;
    const int ring_stride = blockDim.x;
    const int64_t xstep = (dir == 0) ? p.xp_ld_t : -p.xp_ld_t;
    const float* xsrc = p.xproj + n * p.xp_ld_b + static_cast<int64_t>(dir == 0 ? 0 : Tg - 1) * p.xp_ld_t +
                        static_cast<int64_t>(dir) * 4 * H + 4 * u;
    int pf_s = 0;   // next step to prefetch
    auto prefetch_xp = [&]() {
      if (pf_s < Tg) {
        const int t = (dir == 0) ? pf_s : (Tg - 1 - pf_s);
        if (t < len) cp_async<16>(xslot + (pf_s % kPF) * ring_stride, xsrc);
        xsrc += xstep;
      }
      ++pf_s;
      cp_async_commit();
    };
    for (int i = 0; i < kPF; ++i) prefetch_xp();

    // running global pointers (start at the first processed timestep)
    const int t0 = (dir == 0) ? 0 : Tg - 1;
    const int64_t rec_step = (dir == 0) ? H : -H;
    const int64_t y_step = (dir == 0) ? p.y_ld_t : -p.y_ld_t;
    const int64_t hp_step = (dir == 0) ? p.hp_ld_t : -p.hp_ld_t;
    uint4* rec_run = p.rec + ((static_cast<int64_t>(dir) * p.B + n) * T + t0) * H + u;
    // the layer output and the entering state are written as 8-byte words (the quad's 4 units of one utterance):
    // lane ul == 0 writes y, lane ul == 1 writes hprev
    __nv_bfloat16* y_run = p.y + n * p.y_ld_b + t0 * p.y_ld_t + static_cast<int64_t>(dir) * H + 4 * quad;
    __nv_bfloat16* hp_run = p.hprev ? p.hprev + n * p.hp_ld_b + t0 * p.hp_ld_t + static_cast<int64_t>(dir) * H + 4 * quad : nullptr;
    // after the in-warp gather the 8 lanes sharing `tig` hold the same 16 bytes (utterances 2tig, 2tig+1 of the
    // quad); lane g pushes them to peer g. Cluster-mapped destination (buffer 0; buffer 1 lies hs_buf_bytes
    // further) and the peer's barrier:
    const uint32_t hs_buf_bytes = static_cast<uint32_t>(p.KT) * 256u;
    const bool peer_ok = static_cast<uint32_t>(g) < CS;
    const uint32_t hs_local = smem_u32(hs + (quad * 8 + ((2 * tig) ^ (((quad & 3) >> 1) << 2))) * 2);
    const uint32_t hs_remote = mapa_u32(hs_local, peer_ok ? g : 0);
    const uint32_t bar_remote = mapa_u32(smem_u32(&full[0]), peer_ok ? g : 0);
    // clusters of more than 8 CTAs (up to 16): lane g also serves peer 8 + g
    const bool peer2_ok = static_cast<uint32_t>(g) + 8u < CS;
    const uint32_t hs_remote2 = mapa_u32(hs_local, peer2_ok ? g + 8 : 0);
    const uint32_t bar_remote2 = mapa_u32(smem_u32(&full[0]), peer2_ok ? g + 8 : 0);

    const bool trace = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0;
// the phase trace costs ~30 instructions per timestep in a loop that is bound by its own instruction stream: it is
// compiled in only with `make TRACE=1` (-DLAS_PHASE_TRACE)
#ifdef LAS_PHASE_TRACE
#define LAS_TRACE(slot) do { if (trace && s >= 64 && s < 72) p.dbg[(s - 64) * 8 + (slot)] = clock64(); } while (0)
#else
#define LAS_TRACE(slot) do { (void)trace; } while (0)
#endif
    for (int s = 0; s < Tg; ++s) {
      const int t = (dir == 0) ? s : (Tg - 1 - s);
      LAS_TRACE(0);
      float acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
      if (s > 0) {
        const int buf = s & 1;
        mbar_wait_cluster(&full[buf], ((s - 1) >> 1) & 1);
        LAS_TRACE(1);
        if (threadIdx.x == 0 && s + 2 < Tg) mbar_arrive_expect_tx(&full[buf], tx_bytes);  // re-arm for step s+2
        const uint2* hb = reinterpret_cast<const uint2*>(hs + buf * p.KT * 64) + tig * 8 + (g ^ ((tig >> 1) << 2));
        if (p.KT == kMaxKT || KTS > 0) {     // H = 320 (or wide): no per-k-tile bound check in the instruction stream
#pragma unroll
          for (int kt = 0; kt < kMaxKT; ++kt) {
            const uint2 b = hb[kt * 32];
            const uint32_t Af[4] = {A[kt].x, A[kt].y, A[kt].z, A[kt].w};
            mma_bf16_16816(acc[kt & 3], Af, b.x, b.y);
          }
          if (KTS > 0) {          // the shared-memory resident part of the weights (zero fragments past KT)
            const int nks = p.KT - kMaxKT;
#pragma unroll 4
            for (int j = 0; j < KTS; ++j) {
              if (j < nks) {
                const uint2 b = hb[(kMaxKT + j) * 32];
                const uint4 a4 = As[j * 32];
                const uint32_t Af[4] = {a4.x, a4.y, a4.z, a4.w};
                mma_bf16_16816(acc[j & 3], Af, b.x, b.y);
              }
            }
          }
        } else {
#pragma unroll
          for (int kt = 0; kt < kMaxKT; ++kt) {
            if (kt < p.KT) {
              const uint2 b = hb[kt * 32];
              const uint32_t Af[4] = {A[kt].x, A[kt].y, A[kt].z, A[kt].w};
              mma_bf16_16816(acc[kt & 3], Af, b.x, b.y);
            }
          }
        }
      }
      LAS_TRACE(2);
      // C fragment: c0/c1 = row g (gate gl of unit ul), c2/c3 = row g+8 (gate 2+gl), utterances 2tig / 2tig+1.
      // Lanes g and g^4 trade halves so that each ends up with all four gates of (unit ul, utterance 2tig+gl).
      float cf[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) cf[c] = (acc[0][c] + acc[1][c]) + (acc[2][c] + acc[3][c]);
      const float r0 = __shfl_xor_sync(0xffffffffu, gl ? cf[0] : cf[1], 16);
      const float r1 = __shfl_xor_sync(0xffffffffu, gl ? cf[2] : cf[3], 16);
      cp_async_wait<kPF - 1>();   // this step's xproj values have landed in the ring
      const bool act = t < len;
      float sv_i = 0.f, sv_f = 0.f, sv_g = 0.f, sv_o = 0.f, tc = 0.f;
      if (act) {
        const float4 xp = xslot[(s % kPF) * ring_stride];
        const float gi = (gl ? r0 : cf[0]) + xp.x;
        const float gf = (gl ? cf[1] : r0) + xp.y;
        const float gg = (gl ? r1 : cf[2]) + xp.z;
        const float go = (gl ? cf[3] : r1) + xp.w;
        if (kAct == 1) {    // MUFU.TANH forms: 3 + 1 instructions (2^-11 relative error; systematic, see persist_lstm_fwd)
          sv_i = fmaf(0.5f, tanh_fast(0.5f * gi), 0.5f); sv_f = fmaf(0.5f, tanh_fast(0.5f * gf), 0.5f);
          sv_g = tanh_fast(gg); sv_o = fmaf(0.5f, tanh_fast(0.5f * go), 0.5f);
        } else if (kAct == 2) {   // bare ex2 + rcp forms: 4 + 5 instructions, ~1e-7 absolute error
          sv_i = sigmoid_er(gi); sv_f = sigmoid_er(gf); sv_g = tanh_er(gg); sv_o = sigmoid_er(go);
        } else {
          sv_i = sigmoid_acc(gi); sv_f = sigmoid_acc(gf); sv_g = tanh_acc(gg); sv_o = sigmoid_acc(go);
        }
        c_st = sv_f * c_st + sv_i * sv_g;
        tc = kAct == 1 ? tanh_fast(c_st) : (kAct == 2 ? tanh_er(c_st) : tanh_acc(c_st));
        const __nv_bfloat16 hb16 = __float2bfloat16(sv_o * tc);
        h_bits = static_cast<uint32_t>(*reinterpret_cast<const unsigned short*>(&hb16));
      }
      LAS_TRACE(3);
      // gather: (units 0,1 | 2,3) words of this lane's utterance, then the neighbour utterance's
      const uint32_t other = __shfl_xor_sync(0xffffffffu, h_bits, 4);
      const uint32_t wp = (ul & 1) ? (other | (h_bits << 16)) : (h_bits | (other << 16));
      const uint32_t wo = __shfl_xor_sync(0xffffffffu, wp, 8);
      const uint2 qw = (ul & 2) ? make_uint2(wo, wp) : make_uint2(wp, wo);
      const uint32_t x0 = __shfl_xor_sync(0xffffffffu, qw.x, 16);
      const uint32_t x1 = __shfl_xor_sync(0xffffffffu, qw.y, 16);
      // critical path first: the new state goes to the peers before anything is written to HBM
      if (s + 1 < Tg) {
        const int nbuf = (s + 1) & 1;
        const uint4 v = gl ? make_uint4(x0, x1, qw.x, qw.y) : make_uint4(qw.x, qw.y, x0, x1);
        if (peer_ok) st_async_v4(hs_remote + nbuf * hs_buf_bytes, v, bar_remote + nbuf * 8u);
        if (peer2_ok) st_async_v4(hs_remote2 + nbuf * hs_buf_bytes, v, bar_remote2 + nbuf * 8u);
      }
      LAS_TRACE(4);
      // saved activations (BPTT operands) and the layer output
      if (act) {
        const __half2 lo = __floats2half2_rn(sv_i, sv_f), hi = __floats2half2_rn(sv_g, sv_o);
        uint4 r;
        r.x = *reinterpret_cast<const uint32_t*>(&lo);
        r.y = *reinterpret_cast<const uint32_t*>(&hi);
        r.z = __float_as_uint(c_st);
        r.w = __float_as_uint(tc);
        *rec_run = r;
      }
      if (ul == 0) {
        // every row of y is written here, zeros past the utterance's length (pad_packed_sequence): the caller does
        // not have to clear the buffer first
        if (n < p.B) {
          const uint2 yv = act ? qw : make_uint2(0u, 0u);
          *reinterpret_cast<uint2*>(y_run) = yv;
          if (p.rep_row && t == T - 1) *reinterpret_cast<uint2*>(y_run + p.y_ld_t) = yv;
        }
      } else if (ul == 1 && hp_run && n < p.B) {
        *reinterpret_cast<uint2*>(hp_run) = quad_prev;
      }
      quad_prev = qw;
      rec_run += rec_step; y_run += y_step;
      if (hp_run) hp_run += hp_step;
      LAS_TRACE(5);
      prefetch_xp();   // refill this ring slot (step s + kPF): DRAM latency hidden behind kPF timesteps
      LAS_TRACE(6);
      LAS_TRACE(7);
    }
#undef LAS_TRACE
    // rows Tg .. T-1: nobody in this cluster is active there. y is zero (pad_packed_sequence); the entering state is
    // the frozen last h in the forward direction and the zero initial state in the reverse one (its walk starts at
    // Tg - 1), exactly what the full-length loop wrote.
    if (Tg < T && n < p.B && ul < 2) {
      __nv_bfloat16* yt = p.y + n * p.y_ld_b + static_cast<int64_t>(Tg) * p.y_ld_t + static_cast<int64_t>(dir) * H + 4 * quad;
      __nv_bfloat16* ht = p.hprev ? p.hprev + n * p.hp_ld_b + static_cast<int64_t>(Tg) * p.hp_ld_t + static_cast<int64_t>(dir) * H + 4 * quad : nullptr;
      const uint2 hv = (dir == 0) ? quad_prev : make_uint2(0u, 0u);
      if (ul == 0) {
        const int rows = T - Tg + (p.rep_row ? 1 : 0);
        for (int i = 0; i < rows; ++i, yt += p.y_ld_t) *reinterpret_cast<uint2*>(yt) = make_uint2(0u, 0u);
      } else if (ht) {
        for (int i = Tg; i < T; ++i, ht += p.hp_ld_t) *reinterpret_cast<uint2*>(ht) = hv;
      }
    }
  }
  cluster.sync();   // nobody exits while remote stores may still target its shared memory
}

struct BwdP {
  const float* dy; int64_t dy_ld_b, dy_ld_t;
  const uint32_t* wT_pk;       // owner-ordered W_hh^T fragments [ndir][JT][CS*KTC][32][4]
  const int32_t* lens;
  const uint4* rec;            // records written by lstm_persist_fwd_kernel
  __nv_bfloat16* dG;           // gate-minor columns: [dir*4H + 4*u + gate]
  int64_t dg_ld_b, dg_ld_t;
  int B, T, H, rep_row;
  int UGC, JT, MTW, CSn;
  int walk_all;
  long long* dbg;
};

// grid (CS, ceil(B/8), ndir), cluster (CS,1,1), block 64*UGC (= 8 utterances x UPC units).
// MTS: m-tiles (16 hidden units) per warp whose W_hh^T fragments live in shared memory, after the 2 held in
// registers (wide geometry: 2; otherwise 0); CSM: largest cluster size served; PF: prefetch ring depth.
template <int MTS, int CSM, int PF>
__global__ void __launch_bounds__(320, 1) lstm_persist_bwd_kernel(BwdP p) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int MT = 2 + MTS;
  const int UPC = 8 * p.UGC, KTC = 2 * p.UGC;
  uint64_t* pfull = reinterpret_cast<uint64_t*>(smem);                   // [2]
  float* part = reinterpret_cast<float*>(smem + 16);                     // [2][CS][UPC*8]
  uint32_t* dgs = reinterpret_cast<uint32_t*>(smem + 16 + 2 * p.CSn * UPC * 8 * 4);   // [KTC][32][2]
  float* pring = reinterpret_cast<float*>(dgs + KTC * 64);                             // [PF][threads][8]
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank(), CS = cluster.num_blocks();
  const int grp = blockIdx.y, dir = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  const int T = p.T, H = p.H;
  const int vu = max(0, min(UPC, H - static_cast<int>(rank) * UPC));     // valid units owned by this CTA
  const uint32_t tx_bytes = CS * static_cast<uint32_t>(vu) * 32u;
  // timesteps this cluster walks (see lstm_persist_fwd_kernel): rows Tg .. T-1 of its utterances only get zero
  // gate gradients, written by the tail loop after the recurrence
  const int Tg = group_steps(p.lens, blockIdx.y, p.B, T, p.walk_all);

  if (threadIdx.x == 0) {
    mbar_init(&pfull[0], 1);
    mbar_init(&pfull[1], 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < KTC * 64; i += blockDim.x) dgs[i] = 0u;
  __syncthreads();
  if (threadIdx.x == 0) {
    if (Tg > 1) mbar_arrive_expect_tx(&pfull[1], tx_bytes);
    if (Tg > 2) mbar_arrive_expect_tx(&pfull[0], tx_bytes);
  }
  cluster.sync();

  // resident W_hh^T fragments: rows = this warp's 16-unit tiles, K = this CTA's own gate rows
  uint4 A[2][kMaxKTW];
  // wide geometry: the fragments of m-tiles 2 .. MT-1 of this warp, [warp][MTS][kMaxKTW][32 lanes] x 16 B
  uint4* As = reinterpret_cast<uint4*>(pring + static_cast<int64_t>(PF) * blockDim.x * 8) + warp * MTS * kMaxKTW * 32 + lane;
  {
    const int KTtot = p.CSn * KTC;
    const int64_t a_dir = static_cast<int64_t>(p.JT) * KTtot * 128;
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int kc = 0; kc < kMaxKTW; ++kc) {
        const int mt = warp * p.MTW + m;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (m < p.MTW && mt < p.JT && kc < KTC)
          v = __ldg(reinterpret_cast<const uint4*>(p.wT_pk + dir * a_dir) +
                    (static_cast<int64_t>(mt) * KTtot + rank * KTC + kc) * 32 + lane);
        if (m < 2) A[m][kc] = v;
        else As[((m - 2) * kMaxKTW + kc) * 32] = v;
      }
  }

  // element owned by this thread in the pointwise phase: local unit ul, utterance row nr
  const int ul = threadIdx.x >> 3, nr = threadIdx.x & 7;
  const int uo = rank * UPC + ul;
  const int nb = grp * kNB + nr;
  const bool own = (uo < H) && (nb < p.B);
  const int len = own ? (p.lens ? p.lens[nb] : T) : 0;
  float dc_st = 0.f;

  // operands of the gate-derivative math: thread-private cp.async ring, fetched kPF timesteps ahead
  //   slot layout (8 floats): [0..3] the forward record (gates 4 x f16, c_t, tanh c_t), [4] dy, [5] dy of the
  //   replicated row. c_prev is the c_t field of the NEXT processed step's slot (already in flight).
  float* pslot = pring + static_cast<int64_t>(threadIdx.x) * 8;
  const int ring_stride = blockDim.x * 8;
  // running source offsets (first processed timestep: t = T-1 for the forward direction, 0 for the reverse)
  const int tb0 = (dir == 0) ? Tg - 1 : 0;
  const int64_t sv_step = (dir == 0) ? -H : H;
  const int64_t dy_step = (dir == 0) ? -p.dy_ld_t : p.dy_ld_t;
  int64_t sv_run = ((static_cast<int64_t>(dir) * p.B + nb) * T + tb0) * H + uo;
  const float* dy_run = p.dy ? p.dy + nb * p.dy_ld_b + tb0 * p.dy_ld_t + static_cast<int64_t>(dir) * H + uo : nullptr;
  const int64_t dg_step = (dir == 0) ? -p.dg_ld_t : p.dg_ld_t;
  __nv_bfloat16* dg_run = p.dG + nb * p.dg_ld_b + tb0 * p.dg_ld_t + static_cast<int64_t>(dir) * 4 * H + 4 * uo;
  int pf_s = 0;
  auto prefetch = [&]() {
    const int t = (dir == 0) ? (Tg - 1 - pf_s) : pf_s;
    if (pf_s < Tg && t < len) {
      float* dst = pslot + (pf_s % PF) * ring_stride;
      cp_async<16>(dst, p.rec + sv_run);
      if (dy_run) {
        cp_async<4>(dst + 4, dy_run);
        if (p.rep_row && t == T - 1) cp_async<4>(dst + 5, dy_run + p.dy_ld_t);
      }
    }
    ++pf_s;
    sv_run += sv_step;
    if (dy_run) dy_run += dy_step;
    cp_async_commit();
  };
  for (int i = 0; i < PF; ++i) prefetch();

  // reduce-scatter destinations of this lane's partial sums (buffer 0; buffer 1 is part_buf_bytes further)
  // Lanes tig and tig^1 trade halves of their C fragments so that each holds 4 consecutive utterances of ONE unit
  // (even tig: row g, odd tig: row g+8) and sends them as one 16-byte st.async.
  uint32_t sc_dst[MT], sc_bar[MT];
  bool sc_ok[MT];
  const uint32_t part_buf_bytes = static_cast<uint32_t>(p.CSn) * UPC * 8 * 4;
  const int hh_own = tig & 1;
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    const int mt = warp * p.MTW + m;
    const int j = 16 * mt + g + 8 * hh_own;
    sc_ok[m] = m < p.MTW && mt < p.JT && j < H;
    const uint32_t owner = sc_ok[m] ? j / UPC : 0;
    const int jl = sc_ok[m] ? j - owner * UPC : 0;
    sc_dst[m] = mapa_u32(smem_u32(part + static_cast<int64_t>(rank) * UPC * 8 + jl * 8 + 2 * (tig & ~1)), owner);
    sc_bar[m] = mapa_u32(smem_u32(&pfull[0]), owner);
  }

  const bool trace = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0;
#ifdef LAS_PHASE_TRACE
#define LAS_TRACE(slot) do { if (trace && s >= 64 && s < 72) p.dbg[64 + (s - 64) * 8 + (slot)] = clock64(); } while (0)
#else
#define LAS_TRACE(slot) do { (void)trace; } while (0)
#endif
  for (int s = 0; s < Tg; ++s) {
    const int t = (dir == 0) ? (Tg - 1 - s) : s;
    const bool active = t < len;
    LAS_TRACE(0);
    // operands of the gate-derivative math into registers, then the ring slot is refilled right away: the loads
    // and the cp.async issue are hidden behind the MMA and the DSMEM round trip below instead of following them
    cp_async_wait<PF - 2>();   // this step's and the next step's operands have landed
    float4 rc = make_float4(0.f, 0.f, 0.f, 0.f);
    float c_prev = 0.f, dy_in = 0.f;
    if (active) {
      const float* q = pslot + (s % PF) * ring_stride;
      rc = *reinterpret_cast<const float4*>(q);
      const float* qn = pslot + ((s + 1) % PF) * ring_stride;
      if (dir == 0) { if (t > 0) c_prev = qn[2]; }
      else          { if (t + 1 < len) c_prev = qn[2]; }
      if (p.dy) {
        dy_in = q[4];
        if (p.rep_row && t == T - 1) dy_in += q[5];
      }
    }
    prefetch();   // refills slot s % kPF (just read) with step s + kPF
    float dh = 0.f;
    if (s > 0) {
      const int buf = s & 1;
      // partial dh for this warp's unit tiles from this CTA's own gate gradients of the previous step
      float acc[MT][2][4];
#pragma unroll
      for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
      const uint2* db = reinterpret_cast<const uint2*>(dgs) + lane;
      if (KTC == kMaxKTW) {       // H = 320 / 640: no per-k-tile bound check in the instruction stream
#pragma unroll
        for (int kc = 0; kc < kMaxKTW; ++kc) {
          const uint2 b = db[kc * 32];
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            const uint32_t Af[4] = {A[m][kc].x, A[m][kc].y, A[m][kc].z, A[m][kc].w};
            mma_bf16_16816(acc[m][kc & 1], Af, b.x, b.y);
          }
#pragma unroll
          for (int m = 2; m < MT; ++m) {          // fragments resident in shared memory (zero past JT / MTW)
            const uint4 a4 = As[((m - 2) * kMaxKTW + kc) * 32];
            const uint32_t Af[4] = {a4.x, a4.y, a4.z, a4.w};
            mma_bf16_16816(acc[m][kc & 1], Af, b.x, b.y);
          }
        }
      } else {
#pragma unroll
        for (int kc = 0; kc < kMaxKTW; ++kc) {
          if (kc < KTC) {
            const uint2 b = db[kc * 32];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
              const uint32_t Af[4] = {A[m][kc].x, A[m][kc].y, A[m][kc].z, A[m][kc].w};
              mma_bf16_16816(acc[m][kc & 1], Af, b.x, b.y);
            }
#pragma unroll
            for (int m = 2; m < MT; ++m) {
              const uint4 a4 = As[((m - 2) * kMaxKTW + kc) * 32];
              const uint32_t Af[4] = {a4.x, a4.y, a4.z, a4.w};
              mma_bf16_16816(acc[m][kc & 1], Af, b.x, b.y);
            }
          }
        }
      }
      LAS_TRACE(1);
      // scatter the partial sums to the CTAs that own the units
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = acc[m][0][c] + acc[m][1][c];
        // give away the row this lane does not send (even tig keeps row g = v[0..1], odd tig keeps row g+8 = v[2..3])
        const float r0 = __shfl_xor_sync(0xffffffffu, hh_own ? v[0] : v[2], 1);
        const float r1 = __shfl_xor_sync(0xffffffffu, hh_own ? v[1] : v[3], 1);
        if (sc_ok[m]) {
          const uint4 q = hh_own ? make_uint4(__float_as_uint(r0), __float_as_uint(r1), __float_as_uint(v[2]), __float_as_uint(v[3]))
                                 : make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(r0), __float_as_uint(r1));
          st_async_v4(sc_dst[m] + buf * part_buf_bytes, q, sc_bar[m] + buf * 8u);
        }
      }
      LAS_TRACE(2);
      mbar_wait_cluster(&pfull[buf], ((s - 1) >> 1) & 1);
      LAS_TRACE(3);
      if (threadIdx.x == 0 && s + 2 < Tg) mbar_arrive_expect_tx(&pfull[buf], tx_bytes);
      const float* pp = part + static_cast<int64_t>(buf) * p.CSn * UPC * 8 + threadIdx.x;
      // the CS <= 8 partial sums: all loads first, then a tree (a serial load-add chain was 15 % of this kernel's
      // stall samples, profiles/r02_ncu_full_summary.txt)
      float pv[CSM];
#pragma unroll
      for (uint32_t src = 0; src < CSM; ++src) pv[src] = src < CS ? pp[src * UPC * 8] : 0.f;
      dh = ((pv[0] + pv[1]) + (pv[2] + pv[3])) + ((pv[4] + pv[5]) + (pv[6] + pv[7]));
      if (CSM > 8) dh += ((pv[CSM - 8] + pv[CSM - 7]) + (pv[CSM - 6] + pv[CSM - 5])) + ((pv[CSM - 4] + pv[CSM - 3]) + (pv[CSM - 2] + pv[CSM - 1]));
    }
    // gate derivatives (same math as cell_bwd_kernel)
    float d4[4] = {0.f, 0.f, 0.f, 0.f};
    LAS_TRACE(4);
    if (active) {
      const uint2 gpk = make_uint2(__float_as_uint(rc.x), __float_as_uint(rc.y));
      const float tc = rc.w;
      dh += dy_in;
      const float2 if_ = __half22float2(*reinterpret_cast<const __half2*>(&gpk.x));
      const float2 go_ = __half22float2(*reinterpret_cast<const __half2*>(&gpk.y));
      const float i = if_.x, f = if_.y, gc = go_.x, o = go_.y;
      const float dc = dh * o * (1.f - tc * tc) + dc_st;
      dc_st = dc * f;
      d4[0] = dc * gc * i * (1.f - i);
      d4[1] = dc * c_prev * f * (1.f - f);
      d4[2] = dc * i * (1.f - gc * gc);
      d4[3] = dh * tc * o * (1.f - o);
    }
    if (own) {
      uint2 pk;
      pk.x = pack_bf16x2(d4[0], d4[1]);
      pk.y = pack_bf16x2(d4[2], d4[3]);
      *reinterpret_cast<uint2*>(dg_run) = pk;
    }
    dg_run += dg_step;
    if (s + 1 < Tg && uo < H) {
      // B fragments of the next step's MMA: K index k = gate*UPC + ul within this CTA
      __nv_bfloat16* d16 = reinterpret_cast<__nv_bfloat16*>(dgs);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = q * UPC + ul;
        const int kt = k >> 4, kk = k & 15;
        d16[((kt * 32 + nr * 4 + ((kk & 7) >> 1)) * 2 + (kk >> 3)) * 2 + (kk & 1)] = __float2bfloat16(d4[q]);
      }
    }
    LAS_TRACE(5);
    LAS_TRACE(6);
    __syncthreads();
    LAS_TRACE(7);
  }
#undef LAS_TRACE
  if (own) {
    __nv_bfloat16* dz = p.dG + nb * p.dg_ld_b + static_cast<int64_t>(Tg) * p.dg_ld_t + static_cast<int64_t>(dir) * 4 * H + 4 * uo;
    for (int i = Tg; i < T; ++i, dz += p.dg_ld_t) *reinterpret_cast<uint2*>(dz) = make_uint2(0u, 0u);
  }
  cluster.sync();
}

// owner-ordered packing of W_hh^T: logical A[j][k], k = r*4*UPC + gate*UPC + ul  <->  W[gate*H + r*UPC + ul][j]
__global__ void pack_whhT_owner_kernel(const float* __restrict__ W, int H, int CS, int UPC, int JT,
                                       uint32_t* __restrict__ out) {
  const int KTtot = CS * UPC / 4;
  const int64_t total = static_cast<int64_t>(JT) * KTtot * 128;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int jj = idx & 3;
    const int lane = (idx >> 2) & 31;
    const int64_t tk = idx >> 7;
    const int kt = tk % KTtot;
    const int tile = tk / KTtot;
    const int g = lane >> 2, tig = lane & 3;
    const int j = 16 * tile + g + 8 * (jj & 1);
    const int k0 = 16 * kt + 2 * tig + 8 * (jj >> 1);
    float v[2] = {0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int k = k0 + q;
      const int r = k / (4 * UPC), rem = k % (4 * UPC);
      const int gate = rem / UPC, ulq = rem % UPC;
      const int uq = r * UPC + ulq;
      if (j < H && uq < H) v[q] = W[(static_cast<int64_t>(gate) * H + uq) * H + j];
    }
    out[idx] = pack_bf16x2(v[0], v[1]);
  }
}

void* g_dbg_buf = nullptr;   // las_set_debug_buffer: >= 128 int64 of device memory, or null
int g_persist = -1;   // -1: take the default from the environment (LAS_DISABLE_PERSISTENT=1 turns it off)
bool persist_enabled() {
  if (g_persist < 0) {
    const char* e = getenv("LAS_DISABLE_PERSISTENT");
    g_persist = (e && e[0] == '1') ? 0 : 1;
  }
  return g_persist == 1;
}

template <typename Kern, typename P>
int launch_cluster(Kern kern, const P& p, int CS, int NG, int ndir, int threads, size_t smem, cudaStream_t stream) {
  // attributes once per KERNEL: instantiations of one kernel template share this function (same pointer type)
  static const void* done[8] = {nullptr};
  bool attr_set = false;
  int slot = -1;
  for (int i = 0; i < 8; ++i) {
    if (done[i] == reinterpret_cast<const void*>(kern)) attr_set = true;
    if (done[i] == nullptr && slot < 0) slot = i;
  }
  if (!attr_set) {
    LAS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    LAS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    if (slot >= 0) done[slot] = reinterpret_cast<const void*>(kern);
  }
  LAS_REQUIRE(smem <= 224 * 1024, "persistent LSTM: %zu bytes of shared memory needed", smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CS, NG, ndir);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  LAS_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  ++g_launches;
  return 0;
}

}  // namespace

void* g_dbg_buf_shared = nullptr;

// resident-cluster capacity of the forward kernel for a geometry (cached per cluster size / block size)
static int fwd_max_clusters(const FGeom& f) {
  static int cache[17][13];   // [CS][WPC], 0 = unknown, -1 = failed
  int& c = cache[f.CS][f.WPC];
  if (c != 0) return c > 0 ? c : 0;
  const int threads = 32 * f.WPC;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = f.CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cfg.gridDim = dim3(f.CS, 1, 1); cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = 16 + static_cast<size_t>(2) * f.KT * 256 + static_cast<size_t>(kPF) * threads * 16;
  int n = 0;
  if (cudaFuncSetAttribute(lstm_persist_fwd_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
      cudaFuncSetAttribute(lstm_persist_fwd_kernel<0, 0>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
      cudaOccupancyMaxActiveClusters(&n, lstm_persist_fwd_kernel<0, 0>, &cfg) != cudaSuccess)
    n = -1;
  (void)cudaGetLastError();
  c = n > 0 ? n : -1;
  return n > 0 ? n : 0;
}

static int walk_all_steps() {
  static const int v = (getenv("LAS_LSTM_WALK_ALL") != nullptr && atoi(getenv("LAS_LSTM_WALK_ALL")) != 0) ? 1 : 0;
  return v;
}

int persist_supported(int H) {
  Geom g;
  FGeom f;
  return persist_enabled() && geom_for(H, g) && fgeom_for(H, f) ? 1 : 0;
}

int persist_lstm_fwd(const float* xproj, const void* whh_pk, const int32_t* lens, int B, int T, int H, int ndir,
                     void* y, int64_t y_ld_b, int64_t y_ld_t, int rep_row, void* hprev, int64_t hp_ld_b,
                     int64_t hp_ld_t, void* rec, cudaStream_t stream) {
  // The forward fragments are packed per quad, so the cluster size is free at launch: fewer warps per CTA (two per
  // scheduler instead of three at H = 320) shorten every phase of the step, as long as all clusters stay resident.
  // (7-CTA clusters of 12 warps were measured as a way to keep B = 64 -- 16 clusters -- in one wave: a B200 holds 15
  // of them, exactly as many as 8-CTA clusters, and the step is slower (1.26 vs 1.14 us), so they are not chosen.)
  FGeom g;
  LAS_REQUIRE(fgeom_for(H, g), "persistent LSTM: hidden size %d unsupported", H);
  {
    const int need = ((B + kNB - 1) / kNB) * ndir;
    FGeom g10;
    if (g.KT <= kMaxKT && fgeom_for(H, g10, 10) && g10.CS > 8 && fwd_max_clusters(g10) >= need) g = g10;
    static const char* force = getenv("LAS_FWD_CS");     // development aid: pin the forward cluster size
    if (force && g.KT <= kMaxKT) {
      FGeom c;
      if (fgeom_for(H, c, atoi(force))) g = c;
    }
  }
  LAS_REQUIRE(y_ld_b % 4 == 0 && y_ld_t % 4 == 0 && hp_ld_b % 4 == 0 && hp_ld_t % 4 == 0 &&
                  reinterpret_cast<uintptr_t>(y) % 8 == 0 && reinterpret_cast<uintptr_t>(hprev) % 8 == 0,
              "persistent LSTM: y / hprev must be 8-byte aligned with strides that are multiples of 4");
  LAS_REQUIRE(reinterpret_cast<uintptr_t>(xproj) % 16 == 0 && reinterpret_cast<uintptr_t>(rec) % 16 == 0,
              "persistent LSTM: xproj / rec must be 16-byte aligned");
  FwdP p;
  p.xproj = xproj; p.xp_ld_t = static_cast<int64_t>(ndir) * 4 * H; p.xp_ld_b = p.xp_ld_t * T;
  p.whh_pk = static_cast<const uint32_t*>(whh_pk); p.lens = lens;
  p.y = static_cast<__nv_bfloat16*>(y); p.y_ld_b = y_ld_b; p.y_ld_t = y_ld_t;
  p.hprev = static_cast<__nv_bfloat16*>(hprev); p.hp_ld_b = hp_ld_b; p.hp_ld_t = hp_ld_t;
  p.rec = static_cast<uint4*>(rec);
  p.B = B; p.T = T; p.H = H; p.rep_row = rep_row;
  p.Q = g.Q; p.WPC = g.WPC; p.KT = g.KT;
  p.walk_all = walk_all_steps();
  p.dbg = static_cast<long long*>(g_dbg_buf);
  const int threads = 32 * g.WPC;
  const bool wide = g.KT > kMaxKT;
  const size_t smem = 16 + static_cast<size_t>(2) * g.KT * 256 + static_cast<size_t>(kPF) * threads * 16 +
                      (wide ? static_cast<size_t>(g.WPC) * kWideKTS * 512 : 0);
  // Gate activations (LAS_FAST_ACT): 2 = bare ex2 + rcp forms (default, ~1e-7 absolute error), 1 = MUFU.TANH forms
  // (2^-11 relative error; 0.9 % shorter config-2 step: 7.88 vs 7.95 ms), 0 = __expf / __fdividef. At config-2 size
  // form 1 measures the same loss error and gradient cosines as form 0 (tools/parity_report.py), but its error is
  // systematic, and on a small ill-conditioned case (tests/test_gpu_lstm.py, seed 8: B = 19, T = 15, H = 32) it took
  // the weight gradients from cosine 0.9997 to 0.9955 against the fp32 oracle; with form 2 the persistent kernel
  // agrees with the per-timestep kernels again (tools/diag_lstm_case.py).
  static const int act = getenv("LAS_FAST_ACT") == nullptr ? kDefaultAct : atoi(getenv("LAS_FAST_ACT"));
  const int NG = (B + kNB - 1) / kNB;
  if (wide) {
    if (act == 1) return launch_cluster(lstm_persist_fwd_kernel<1, kWideKTS>, p, g.CS, NG, ndir, threads, smem, stream);
    if (act == 2) return launch_cluster(lstm_persist_fwd_kernel<2, kWideKTS>, p, g.CS, NG, ndir, threads, smem, stream);
    return launch_cluster(lstm_persist_fwd_kernel<0, kWideKTS>, p, g.CS, NG, ndir, threads, smem, stream);
  }
  if (act == 1) return launch_cluster(lstm_persist_fwd_kernel<1, 0>, p, g.CS, NG, ndir, threads, smem, stream);
  if (act == 2) return launch_cluster(lstm_persist_fwd_kernel<2, 0>, p, g.CS, NG, ndir, threads, smem, stream);
  return launch_cluster(lstm_persist_fwd_kernel<0, 0>, p, g.CS, NG, ndir, threads, smem, stream);
}

int persist_lstm_bwd(const float* dy, int64_t dy_ld_b, int64_t dy_ld_t, int rep_row, const void* wT_owner_pk,
                     const int32_t* lens, int B, int T, int H, int ndir, const void* rec, void* dG, int64_t dg_ld_b,
                     int64_t dg_ld_t, cudaStream_t stream) {
  Geom g;
  LAS_REQUIRE(geom_for(H, g), "persistent LSTM: hidden size %d unsupported", H);
  LAS_REQUIRE(dg_ld_b % 4 == 0 && dg_ld_t % 4 == 0 && reinterpret_cast<uintptr_t>(dG) % 8 == 0 &&
                  reinterpret_cast<uintptr_t>(rec) % 16 == 0,
              "persistent LSTM: dG must be 8-byte aligned with strides that are multiples of 4, rec 16-byte aligned");
  BwdP p;
  p.dy = dy; p.dy_ld_b = dy_ld_b; p.dy_ld_t = dy_ld_t; p.rep_row = rep_row;
  p.wT_pk = static_cast<const uint32_t*>(wT_owner_pk); p.lens = lens;
  p.rec = static_cast<const uint4*>(rec);
  p.dG = static_cast<__nv_bfloat16*>(dG); p.dg_ld_b = dg_ld_b; p.dg_ld_t = dg_ld_t;
  p.B = B; p.T = T; p.H = H;
  p.UGC = g.UGC; p.JT = g.JT; p.MTW = g.MTW; p.CSn = g.CS;
  p.walk_all = walk_all_steps();
  p.dbg = static_cast<long long*>(g_dbg_buf);
  const int UPC = 8 * g.UGC;
  if (g.CS > 8 || g.MTW > 2) {
    // wide geometry: 16-CTA clusters, two more m-tiles per warp from shared memory, a 4-deep prefetch ring
    const size_t smem = 16 + static_cast<size_t>(2) * g.CS * UPC * 8 * 4 + static_cast<size_t>(2 * g.UGC) * 256 +
                        static_cast<size_t>(kWidePF) * 64 * g.UGC * 8 * 4 + static_cast<size_t>(2 * g.UGC) * 2 * kMaxKTW * 512;
    return launch_cluster(lstm_persist_bwd_kernel<2, 16, kWidePF>, p, g.CS, (B + kNB - 1) / kNB, ndir, 64 * g.UGC, smem, stream);
  }
  const size_t smem = 16 + static_cast<size_t>(2) * g.CS * UPC * 8 * 4 + static_cast<size_t>(2 * g.UGC) * 256 +
                      static_cast<size_t>(kPF) * 64 * g.UGC * 8 * 4;
  return launch_cluster(lstm_persist_bwd_kernel<0, 8, kPF>, p, g.CS, (B + kNB - 1) / kNB, ndir, 64 * g.UGC, smem, stream);
}

}  // namespace las

using namespace las;

extern "C" {

/* development aid: clock64() phase trace of timesteps 64..71 of cluster (0,0,0): 8 slots per step,
 * forward kernel in entries [0,64), backward kernel in [64,128) */
int las_set_debug_buffer(void* dev_int64_x128) {
  g_dbg_buf = dev_int64_x128;
  las::g_dbg_buf_shared = dev_int64_x128;
  return 0;
}

int las_set_persistent(int on) {
  const int prev = persist_enabled() ? 1 : 0;
  g_persist = on ? 1 : 0;
  return prev;
}

/* development aid: how many clusters of the persistent LSTM kernels (which = 0 forward, 1 backward) can be resident
 * at once for hidden size H (cudaOccupancyMaxActiveClusters); more clusters than this run as a second wave */
int las_lstm_persist_max_clusters(int which, int H) {
  Geom g;
  FGeom f;
  if (!geom_for(H, g) || !fgeom_for(H, f)) return 0;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (which == 2 || which == 3) {          // forward kernel with 7-CTA / 10-CTA clusters
    FGeom c;
    return fgeom_for(H, c, which == 2 ? 7 : 10) ? fwd_max_clusters(c) : 0;
  }
  if (which == 0) {
    const int threads = 32 * f.WPC;
    cfg.gridDim = dim3(f.CS, 1, 1); cfg.blockDim = dim3(threads); at[0].val.clusterDim.x = f.CS;
    cfg.dynamicSmemBytes = 16 + static_cast<size_t>(2) * f.KT * 256 + static_cast<size_t>(kPF) * threads * 16;
    cudaFuncSetAttribute(lstm_persist_fwd_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (cudaOccupancyMaxActiveClusters(&n, lstm_persist_fwd_kernel<0, 0>, &cfg) != cudaSuccess) n = -1;
  } else {
    const int UPC = 8 * g.UGC;
    cfg.gridDim = dim3(g.CS, 1, 1); cfg.blockDim = dim3(64 * g.UGC); at[0].val.clusterDim.x = g.CS;
    cfg.dynamicSmemBytes = 16 + static_cast<size_t>(2) * g.CS * UPC * 8 * 4 + static_cast<size_t>(2 * g.UGC) * 256 +
                           static_cast<size_t>(kPF) * 64 * g.UGC * 8 * 4;
    cudaFuncSetAttribute(lstm_persist_bwd_kernel<0, 8, kPF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (cudaOccupancyMaxActiveClusters(&n, lstm_persist_bwd_kernel<0, 8, kPF>, &cfg) != cudaSuccess) n = -1;
  }
  (void)cudaGetLastError();
  return n;
}

int las_lstm_persistent_geometry(int H, int* cs, int* upc) {
  Geom g;
  if (!persist_enabled() || !geom_for(H, g)) return 0;
  if (cs) *cs = g.CS;
  if (upc) *upc = 8 * g.UGC;
  return 1;
}

int64_t las_whhT_owner_bytes(int H) {
  Geom g;
  if (!geom_for(H, g)) return 0;
  return static_cast<int64_t>(g.JT) * (g.CS * 2 * g.UGC) * 128 * 4;
}

int las_pack_whhT_owner(const float* W, int H, void* out, void* stream) {
  Geom g;
  LAS_REQUIRE(geom_for(H, g), "pack_whhT_owner: hidden size %d unsupported", H);
  const int64_t total = static_cast<int64_t>(g.JT) * (g.CS * 2 * g.UGC) * 128;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
  pack_whhT_owner_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(W, H, g.CS, 8 * g.UGC, g.JT,
                                                                                  static_cast<uint32_t*>(out));
  ++g_launches;
  LAS_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
