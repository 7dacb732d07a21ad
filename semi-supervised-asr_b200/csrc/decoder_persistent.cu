// Cluster-persistent attention decoder (model.py:139-173 AttLoc.forward, 283-367 Decoder.forward,
// teacher-forced): ONE launch runs all L decoder steps. A cluster of 16 CTAs serves NB utterances
// (NB = 1, 2, 4 or 8); the recurrent weights [W_hh | W_ih[:, E:]] and mlp_dec are sharded over the
// 16 CTAs by output row and stay resident in REGISTERS as mma.m16n8k16 A fragments for the whole
// sequence; per-utterance attention operands (P = mlp_enc(enc_h), Q = mlp_o.weight(enc_h)) are
// sharded by encoder frame over the G = 16/NB "owner" CTAs of each utterance and stay resident in
// SHARED MEMORY. Per step the CTAs exchange only small vectors through distributed shared memory
// (st.shared::cluster) and meet at cluster barriers:
//
//   P1  gates = embx_t + Wr [z_{t-1}; c_{t-1}]   row-sharded MMA -> LSTM cell -> z_t shard
//       -> broadcast z_t (bf16 B-fragment words) to all 16 CTAs                     | barrier
//   P2  dz = mlp_dec z_t  (row-sharded MMA) -> to the owners of each utterance      | barrier
//   P3  owners: e = gvec . tanh(P + dz + mlp_att conv(w_{t-1})) for their frames (the mlp_att
//       contraction on tensor cores, conv split hi/lo), local softmax statistics, partial
//       context sum_te p[te] Q[te] -> reduce-scatter among the owners                | barrier
//   P4  owners: combine statistics -> w_t, c_t = (sum_te w_t[te] Q[te]) + mlp_o.bias
//       -> broadcast c_t to all 16 CTAs                                              | barrier
//
// c_t = mlp_o(sum_te w[te] enc_h[te]) is evaluated as sum_te w[te] (mlp_o.weight enc_h[te]) + bias
// (softmax weights sum to one), which removes mlp_o from the serial loop.
#include <cooperative_groups.h>
#include "common.cuh"
#include "las_internal.h"
#include "../../include/las_b200.h"

namespace las {

namespace {

constexpr int kCS = 16;        // CTAs per cluster
constexpr int kThreads = 512;
constexpr int kWarps = 16;
constexpr int kMaxFG = 14;     // gate fragments per warp
constexpr int kMaxFD = 3;      // mlp_dec fragments per warp

struct DGeom {
  int NB, G;             // utterances per cluster, owner CTAs per utterance
  int UPC, GT;           // hidden units and gate tiles (4 units x 4 gates) per CTA
  int KTg, KTd;          // k-tiles of [z; c] and of z
  int AT, nAT;           // mlp_dec 16-row tiles: total, max per CTA
  int KSg, FG, KTp;      // gate phase: K-splits per tile (warp = tile*KSg + ks), fragments per warp, padded k-tiles of the state buffer
  int KSd, FD;           // mlp_dec phase: likewise
  int TR, TT, WPT, NTW;  // frames per owner, 16-frame tiles, warps per tile, 8-wide att n-tiles per warp
  int AT8, OS, OTs;      // A/8; context dims per owner; 16-row tiles of an owner's context slice
  int KTe, KTc, NC;      // k-tiles over all Te frames; k-tiles over the conv taps; 8-channel n-tiles of the conv
  int Pld, QTld, Tw;     // row strides of the P slice (f32) and of the transposed Q slice (bf16); padded alignment length
  // shared-memory carve-up (byte offsets)
  int o_zB, o_red, o_dzv, o_cred, o_wbuf, o_cwB, o_matt, o_gv, o_P, o_Q, o_epart, o_eall, o_pun, o_wred;
  int smem;
};

inline int rup(int x, int m) { return (x + m - 1) / m * m; }

// Returns false when the problem is not served by the persistent kernel.
bool dec_geom(const las_dec_args* a, int NB, DGeom& g) {
  const int Hd = a->Hd, O = a->O, A = a->A, Te = a->Te;
  if (Hd % 64 != 0 || Hd > 320 || O % 16 != 0 || A % 8 != 0 || A > 512 || a->C > 16 || Te > kThreads) return false;
  g.NB = NB; g.G = kCS / NB;
  g.UPC = Hd / kCS; g.GT = g.UPC / 4;
  g.KTg = (Hd + O) / 16; g.KTd = Hd / 16;
  g.AT = (A + 15) / 16; g.nAT = (g.AT + kCS - 1) / kCS;
  g.KSg = kWarps / g.GT; if (g.KSg > g.KTg) g.KSg = g.KTg;
  g.FG = (g.KTg + g.KSg - 1) / g.KSg;
  g.KTp = g.KSg * g.FG;
  g.KSd = kWarps / g.nAT; if (g.KSd > g.KTd) g.KSd = g.KTd;
  g.FD = (g.KTd + g.KSd - 1) / g.KSd;
  if (g.KSd * g.FD > g.KTp) g.KTp = g.KSd * g.FD;
  if (g.FG > kMaxFG || g.FD > kMaxFD) return false;
  if (O % (2 * g.G) != 0) return false;
  g.OS = O / g.G;
  if (g.OS % 16 != 0 || (Hd / kCS) % 4 != 0) return false;   // whole 16-dim context tiles / whole unit quads per CTA
  g.OTs = (g.OS + 15) / 16;
  g.TR = (Te + g.G - 1) / g.G;
  g.TT = (g.TR + 15) / 16;
  if (g.TT > kWarps) return false;
  g.WPT = kWarps / g.TT;
  g.AT8 = A / 8;
  g.NTW = (g.AT8 + g.WPT - 1) / g.WPT;
  g.KTe = (Te + 15) / 16;
  const int ksz = 2 * a->K + 1;
  g.KTc = (ksz + 15) / 16;
  g.NC = a->C > 8 ? 2 : 1;
  g.Pld = A + ((8 - A % 16) + 16) % 16;   // bf16 row stride, word stride == 4 (mod 8): conflict-free 32-bit fragment loads
  g.QTld = 16 * g.KTe + 8;                // word stride == 4 (mod 8): conflict-free 32-bit fragment loads
  g.Tw = Te + 2 * a->K + 48;              // the Hankel fragments of the last frame tile read up to 45 words past the end
  int off = 0;
  auto take = [&](int bytes) { const int o = off; off += rup(bytes, 16); return o; };
  g.o_zB = take(2 * g.KTp * 256);
  g.o_red = take(kWarps * 128 * 4);
  g.o_dzv = take(A * 4);
  g.o_cred = take(kWarps * 32 * 32);
  g.o_wbuf = take(g.Tw * 4);
  g.o_cwB = take(g.KTc * 2 * 32 * 16);
  g.o_matt = take(g.AT8 * 32 * 16);
  g.o_gv = take(A * 4);
  g.o_P = take(g.TT * 16 * g.Pld * 2);
  g.o_Q = take(g.OTs * 16 * g.QTld * 2);
  g.o_epart = take(g.WPT * g.TT * 16 * 4);
  g.o_eall = take(g.G * g.TR * 4);
  g.o_pun = take(g.KTe * 16 * 4);
  g.o_wred = take(2 * kWarps * 4);
  g.smem = off;
  return g.smem <= 220 * 1024;
}

// ------------------------------------------------------------------------------------------
// cluster primitives
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_remote_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_remote_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_remote_v4_f32(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void st_remote_v4_u32(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// Release/acquire barrier over all threads of the cluster: remote stores issued before it are
// visible to every CTA after it.
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Position (in 32-bit words) of the B-fragment word that holds K elements (k, k+1), k even, of
// column n in a [KT][32][2] fragment buffer.
__device__ __forceinline__ int bfrag_word(int k, int n) {
  const int kt = k >> 4, kk = k & 15;
  return ((kt * 32 + n * 4 + ((kk & 7) >> 1)) * 2) + (kk >> 3);
}

// Forward state buffer: same [KT][32][2] fragment order, but the K positions of a k-tile are quad-permuted (A
// fragments packed with las_pack_afrag modes 3 / 4), so that lane (n, q) holds units 4q..4q+3 of its k-tile: the 16
// units of (k-tile, utterance) are 32 contiguous bytes in natural order and can be pushed as 8/16-byte remote stores.
// Word (two consecutive units k, k+1, k even) of column n:
__device__ __forceinline__ int qfrag_word(int k, int n) {
  const int kt = k >> 4, kk = k & 15;
  return (kt * 32 + n * 4 + (kk >> 2)) * 2 + ((kk >> 1) & 1);
}
__device__ __forceinline__ void st_remote_v2_u32(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

// Element (row, col) of a 16x8 accumulator tile, summed over the KS K-split warps w0 .. w0+KS-1.
__device__ __forceinline__ float red_gather(const float* red, int w0, int KS, int row, int col) {
  const float* r = red + (w0 * 32 + (row & 7) * 4 + (col >> 1)) * 4 + (row >> 3) * 2 + (col & 1);
  float s = 0.f;
  for (int i = 0; i < KS; ++i) s += r[i * 128];
  return s;
}

// bf16 hi/lo split of a float pair: x ~= hi + lo with ~16 mantissa bits in total
__device__ __forceinline__ void split_bf16x2(float x, float y, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16x2(x, y);
  const float2 h = unpack_bf16x2(hi);
  lo = pack_bf16x2(x - h.x, y - h.y);
}

struct DecFwdP {
  int B, L, Te, Hd, O, A, C, K;
  float att_scaling;
  DGeom g;
  const float* P;                // [B, Te, A]
  const __nv_bfloat16* Q;        // [B, Te, O]
  const float* embx;             // [B, L+1, 4Hd]
  const uint32_t* wr_pk;         // [Hd/4][KTg][32][4]  (pack mode 3: quad tiles, quad-permuted K)
  const uint32_t* dec_pk;        // [AT][KTd][32][4]    (pack mode 4: rows in order, quad-permuted K)
  const float* cbias;            // [B, O] mlp_o.bias + frame mean of the uncentred Q
  const float* pbar;             // [B, A] frame mean removed from P: added to dz before it is sent / saved
  const float* conv_w;
  const float* mlp_att;
  const float* gvec;
  float* ws;                     // [B, L+1, Te]; row 0 = initial alignment (input), rows 1.. written
  __nv_bfloat16* zc;             // [B, L+1, Hd+O]; row 0 zeros (input), rows 1.. written
  float* dzf;                    // [B, L, A]
  __half* gates_save;            // [B, L, Hd, 4]
  float* c_save;                 // [B, L, Hd]
  float* cpre;                   // [B, L, O]  context term before the bias
  float* conv_save;              // [B, L, Te, 16] location-conv features (C padded to 16)
  float drop_p; uint32_t drop_site; const unsigned long long* seed_dev;   // cell-input dropout of c_t (model.py:285)
  long long* dbg;                // optional clock64() phase trace (las_set_debug_buffer)
};

// Kernel parameters are copied to shared memory first: every cluster barrier (acquire) invalidates
// the constant/L1 caches, and a constant-bank miss per parameter read was the dominant stall of
// the first version of this kernel (profiles/r01_decfwd_v1_stalls.txt).
__global__ void __launch_bounds__(kThreads, 1) dec_persist_fwd_kernel(const __grid_constant__ DecFwdP p_in) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ DecFwdP p;
  for (int i = threadIdx.x; i < static_cast<int>(sizeof(DecFwdP) / 4); i += kThreads)
    reinterpret_cast<uint32_t*>(&p)[i] = reinterpret_cast<const uint32_t*>(&p_in)[i];
  __syncthreads();
  const DGeom& g = p.g;
  uint32_t* zB = reinterpret_cast<uint32_t*>(smem + g.o_zB);       // [2][KTp][32][2]
  float* red = reinterpret_cast<float*>(smem + g.o_red);           // [16][32][4]
  float* dzv = reinterpret_cast<float*>(smem + g.o_dzv);           // [A]
  float4* cred = reinterpret_cast<float4*>(smem + g.o_cred);       // [16 warps][2][32] partial conv accumulators
  float* wbuf = reinterpret_cast<float*>(smem + g.o_wbuf);         // [Te + 2K + 48], w[j] at wbuf[K + j]
  uint4* cwB = reinterpret_cast<uint4*>(smem + g.o_cwB);           // [KTc][2][32] conv-weight B fragments (hi0, hi1, lo0, lo1)
  uint4* mattB = reinterpret_cast<uint4*>(smem + g.o_matt);        // [AT8][32] mlp_att B fragments (hi0, hi1, lo0, lo1)
  float* gv_s = reinterpret_cast<float*>(smem + g.o_gv);           // [A]
  __nv_bfloat16* P_s = reinterpret_cast<__nv_bfloat16*>(smem + g.o_P);   // [TT*16][Pld]   my frames of P (bf16: below the error of the bf16 GEMM that made it)
  __nv_bfloat16* QT_s = reinterpret_cast<__nv_bfloat16*>(smem + g.o_Q);   // [OTs*16][QTld]  my context dims of Q, transposed
  float* epart = reinterpret_cast<float*>(smem + g.o_epart);       // [WPT][TT*16]
  float* e_all = reinterpret_cast<float*>(smem + g.o_eall);        // [G*TR] scaled energies of all frames (written by the owners)
  float* p_un = reinterpret_cast<float*>(smem + g.o_pun);          // [KTe*16] exp(e - max), zero padded
  float* wred = reinterpret_cast<float*>(smem + g.o_wred);         // [2][16]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tig = lane & 3;
  const uint32_t rank = cluster_rank();
  const int cl = blockIdx.y;
  const int NB = g.NB, G = g.G, UPC = g.UPC, TR = g.TR, OS = g.OS;
  const int Hd = p.Hd, O = p.O, A = p.A, Te = p.Te, L = p.L, K = p.K, C = p.C;
  const int ZC = Hd + O, R = L + 1, ksz = 2 * K + 1;
  const int n_own = rank / G, q = rank % G;
  const int b_own = cl * NB + n_own;
  const bool own_ok = b_own < p.B;
  const int te0 = q * TR;
  const int ntl = own_ok ? max(0, min(TR, Te - te0)) : 0;

  // ---------------- resident weights (registers)
  // gate phase: warp = tile * KSg + ks holds k-tiles [ks*FG, ks*FG + FG) of its tile
  const int g_tile = warp / g.KSg, g_ks = warp % g.KSg;
  const bool g_act = g_tile < g.GT;
  const int g_kt0 = g_ks * g.FG;
  uint4 Ag[kMaxFG];
#pragma unroll
  for (int j = 0; j < kMaxFG; ++j) {
    Ag[j] = make_uint4(0u, 0u, 0u, 0u);
    if (g_act && j < g.FG && g_kt0 + j < g.KTg)
      Ag[j] = __ldg(reinterpret_cast<const uint4*>(p.wr_pk) +
                    (static_cast<int64_t>(rank * g.GT + g_tile) * g.KTg + g_kt0 + j) * 32 + lane);
  }
  const int nATr = (g.AT > static_cast<int>(rank)) ? (g.AT - 1 - static_cast<int>(rank)) / kCS + 1 : 0;   // my mlp_dec tiles
  const int d_tile = warp / g.KSd, d_ks = warp % g.KSd;
  const bool d_act = d_tile < nATr;
  const int d_kt0 = d_ks * g.FD;
  uint4 Ad[kMaxFD];
#pragma unroll
  for (int j = 0; j < kMaxFD; ++j) {
    Ad[j] = make_uint4(0u, 0u, 0u, 0u);
    if (d_act && j < g.FD && d_kt0 + j < g.KTd)
      Ad[j] = __ldg(reinterpret_cast<const uint4*>(p.dec_pk) +
                    (static_cast<int64_t>(rank + kCS * d_tile) * g.KTd + d_kt0 + j) * 32 + lane);
  }
  const int nfg = g.FG, nfd = g.FD;

  // ---------------- resident attention operands (shared memory)
  for (int i = tid; i < 2 * g.KTp * 64; i += kThreads) zB[i] = 0u;
  for (int i = tid; i < g.Tw; i += kThreads) wbuf[i] = 0.f;
  for (int i = tid; i < g.KTe * 16; i += kThreads) p_un[i] = 0.f;
  for (int i = tid; i < kWarps * 64; i += kThreads) cred[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = tid; i < g.KTc * 64; i += kThreads) {
    // B fragment of the conv weights: k = tap, n = channel 8*nc + (l >> 2)
    const int kt = i >> 6, nc = (i >> 5) & 1, l = i & 31, c = 8 * nc + (l >> 2), k0 = 16 * kt + 2 * (l & 3);
    float m[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int kk = k0 + (k & 1) + 8 * (k >> 1);
      m[k] = (c < C && kk < ksz) ? p.conv_w[c * ksz + kk] : 0.f;
    }
    uint4 v;
    split_bf16x2(m[0], m[1], v.x, v.z);
    split_bf16x2(m[2], m[3], v.y, v.w);
    cwB[i] = v;
  }
  for (int i = tid; i < A; i += kThreads) gv_s[i] = p.gvec[i];
  for (int i = tid; i < g.AT8 * 32; i += kThreads) {
    const int nt = i >> 5, l = i & 31, a = 8 * nt + (l >> 2), c0 = 2 * (l & 3);
    float m[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + (k & 1) + 8 * (k >> 1);
      m[k] = (c < C) ? p.mlp_att[a * C + c] : 0.f;
    }
    uint4 v;
    split_bf16x2(m[0], m[1], v.x, v.z);
    split_bf16x2(m[2], m[3], v.y, v.w);
    mattB[i] = v;
  }
  for (int i = tid; i < g.TT * 16 * g.Pld; i += kThreads) {
    const int r = i / g.Pld, a = i % g.Pld;
    P_s[i] = __float2bfloat16((r < ntl && a < A) ? p.P[(static_cast<int64_t>(b_own) * Te + te0 + r) * A + a] : 0.f);
  }
  for (int i = tid; i < g.OTs * 16 * g.QTld; i += kThreads) QT_s[i] = __float2bfloat16(0.f);
  __syncthreads();
  if (own_ok) {
    for (int i = tid; i < Te * OS; i += kThreads) {
      const int te = i / OS, ol = i % OS;
      QT_s[ol * g.QTld + te] = p.Q[(static_cast<int64_t>(b_own) * Te + te) * O + q * OS + ol];
    }
    for (int i = tid; i < Te; i += kThreads) wbuf[K + i] = p.ws[static_cast<int64_t>(b_own) * R * Te + i];
  }

  // ---------------- roles of this thread
  // P1 epilogue: thread -> (utterance n, local unit ulc)
  const int n_e = tid / UPC, ulc = tid % UPC;
  const bool epi = tid < UPC * NB;
  const int b_e = cl * NB + n_e;
  const bool epi_ok = epi && b_e < p.B;
  const int u_e = rank * UPC + ulc;                       // global hidden unit
  const int epi_warps = (UPC * NB + 31) / 32;
  const int e_w0 = (epi ? (ulc >> 2) : 0) * g.KSg;        // first contributing warp of my gate tile
  const int z_word = qfrag_word(u_e & ~3, epi ? n_e : 0);   // first word of my unit's quad
  float cell = 0.f;
  const float* ex_ptr = p.embx + static_cast<int64_t>(epi_ok ? b_e : 0) * R * 4 * Hd + u_e;
  int64_t sv_idx = static_cast<int64_t>(epi_ok ? b_e : 0) * L * Hd + u_e;
  __nv_bfloat16* zc_z_ptr = p.zc + (static_cast<int64_t>(epi_ok ? b_e : 0) * R + 1) * ZC + u_e;
  // P2 epilogue: thread -> (utterance n, local mlp_dec row)
  const int drows = 16 * nATr;
  const int n_d = drows > 0 ? tid / drows : 0, al = drows > 0 ? tid % drows : 0;
  const bool depi = drows > 0 && tid < drows * NB;
  const int a_d = (rank + kCS * (al >> 4)) * 16 + (al & 15);
  const bool depi_ok = depi && a_d < A && (cl * NB + n_d) < p.B;
  const int d_w0 = (al >> 4) * g.KSd;
  float* dzf_ptr = p.dzf + static_cast<int64_t>(depi_ok ? cl * NB + n_d : 0) * L * A + (depi_ok ? a_d : 0);
  const float pbar_d = depi_ok ? p.pbar[static_cast<int64_t>(cl * NB + n_d) * A + a_d] : 0.f;
  // P3: (frame tile, attention-dim slice) of this warp
  const int e_tt = warp / g.WPT, e_wi = warp % g.WPT;
  const bool e_act = warp < g.TT * g.WPT && ntl > 0;
  // conv k-tiles of my frame tile: taps that can touch a valid alignment entry
  const int cm = te0 + 16 * e_tt;
  const int ckt_lo = max(0, K - (cm + 15)) >> 4;
  const int ckt_hi = (cm < Te) ? (min(2 * K, K - cm + Te - 1) >> 4) : -1;
  float* csave_ptr = p.conv_save ? p.conv_save + (static_cast<int64_t>(own_ok ? b_own : 0) * L * Te + cm) * 16 : nullptr;
  // P4
  float* ws_row = p.ws + (static_cast<int64_t>(own_ok ? b_own : 0) * R + 1) * Te;
  const int c_mt = warp;                                   // context m-tile of this warp (if < OTs)
  const int o_l0 = 16 * c_mt + gq;                         // local context dims o_l0, o_l0 + 8 (lanes with tig == 0)
  // cluster-mapped base addresses
  const uint32_t zB_base = smem_u32(zB), dzv_base = smem_u32(dzv), eall_base = smem_u32(e_all);
  const float scal = p.att_scaling;
  const bool drop_on = p.drop_p > 0.f;

  cluster_barrier();   // every CTA's shared memory is initialised before any remote store
  const bool trace = p.dbg != nullptr && blockIdx.y == 0 && rank == 0 && tid == 0;
#define DTRACE(slot) do { if (trace && t >= 8 && t < 12) p.dbg[(t - 8) * 16 + (slot)] = clock64(); } while (0)

  for (int t = 0; t < L; ++t) {
    const int par = t & 1;
    const int zb_nxt_w = (par ^ 1) * g.KTp * 64;     // word offset of the buffer that receives [z_t; c_t]

    // ================= P1: LSTM cell =================
    DTRACE(0);
    float ex[4] = {0.f, 0.f, 0.f, 0.f};
    if (epi_ok) {
#pragma unroll
      for (int k = 0; k < 4; ++k) ex[k] = __ldg(ex_ptr + k * Hd);
    }
    ex_ptr += 4 * Hd;
    if (g_act) {
      float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
      // padded fragments (zero A) read padded, all-zero k-tiles of the state buffer
      const uint2* hb = reinterpret_cast<const uint2*>(zB + par * g.KTp * 64) + lane + g_kt0 * 32;
#pragma unroll
      for (int j = 0; j < kMaxFG; ++j) {
        if (j < nfg) {
          const uint2 b = hb[j * 32];
          const uint32_t Af[4] = {Ag[j].x, Ag[j].y, Ag[j].z, Ag[j].w};
          if (j & 1) mma_bf16_16816(acc1, Af, b.x, b.y);
          else mma_bf16_16816(acc0, Af, b.x, b.y);
        }
      }
      reinterpret_cast<float4*>(red)[warp * 32 + lane] =
          make_float4(acc0[0] + acc1[0], acc0[1] + acc1[1], acc0[2] + acc1[2], acc0[3] + acc1[3]);
    }
    DTRACE(1);
    __syncthreads();
    DTRACE(2);
    // Global stores of a phase's results are issued AFTER the phase's cluster barrier: the barrier's release is a
    // memory fence that waits for every store the thread has in flight, and HBM/L2 write acknowledgements ahead of
    // it were 12 % of this kernel's stall samples (ERRBAR, profiles/r01z_ncu_full_summary.txt). Stored after the
    // barrier, they complete under the next phase.
    uint2 sv_pk = make_uint2(0u, 0u);
    __nv_bfloat16 sv_z = __float2bfloat16(0.f);
    if (warp < epi_warps) {
      uint32_t zbits = 0u;
      if (epi) {
        const int ul4 = ulc & 3;
        const float gi = red_gather(red, e_w0, g.KSg, ul4, n_e) + ex[0];
        const float gf = red_gather(red, e_w0, g.KSg, 4 + ul4, n_e) + ex[1];
        const float gg = red_gather(red, e_w0, g.KSg, 8 + ul4, n_e) + ex[2];
        const float go = red_gather(red, e_w0, g.KSg, 12 + ul4, n_e) + ex[3];
        const float i = sigmoid_acc(gi), f = sigmoid_acc(gf), gc = tanh_acc(gg), o = sigmoid_acc(go);
        cell = f * cell + i * gc;
        const __nv_bfloat16 zb16 = __float2bfloat16(o * tanh_acc(cell));
        zbits = epi_ok ? static_cast<uint32_t>(__bfloat16_as_ushort(zb16)) : 0u;
        __half2 lo = __floats2half2_rn(i, f), hi = __floats2half2_rn(gc, o);
        sv_pk.x = *reinterpret_cast<uint32_t*>(&lo);
        sv_pk.y = *reinterpret_cast<uint32_t*>(&hi);
        sv_z = zb16;
      }
      // the 4 lanes of a quad (UPC % 4 == 0, so they are lanes 4j..4j+3) gather its two words; each then pushes
      // the 8 bytes to a quarter of the cluster
      const uint32_t nb = __shfl_xor_sync(0xffffffffu, zbits, 1);
      const uint32_t wp = (ulc & 1) ? (nb | (zbits << 16)) : (zbits | (nb << 16));
      const uint32_t wo = __shfl_xor_sync(0xffffffffu, wp, 2);
      if (epi) {
        const uint32_t w0 = (ulc & 2) ? wo : wp, w1 = (ulc & 2) ? wp : wo;
        const uint32_t off = zB_base + 4u * static_cast<uint32_t>(zb_nxt_w + z_word);
#pragma unroll
        for (int i = 0; i < kCS / 4; ++i) st_remote_v2_u32(mapa(off, (ulc & 3) * (kCS / 4) + i), w0, w1);
      }
    }
    DTRACE(3);
    cluster_barrier();
    DTRACE(4);
    if (epi_ok) {
      reinterpret_cast<uint2*>(p.gates_save)[sv_idx] = sv_pk;
      p.c_save[sv_idx] = cell;
      zc_z_ptr[0] = sv_z;
    }
    sv_idx += Hd; zc_z_ptr += ZC;

    // ================= P2: dz = mlp_dec z_t; location conv of w_{t-1} on tensor cores =================
    if (d_act) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const uint2* hb = reinterpret_cast<const uint2*>(zB + zb_nxt_w) + lane + d_kt0 * 32;
#pragma unroll
      for (int j = 0; j < kMaxFD; ++j) {
        if (j < nfd) {
          const uint2 b = hb[j * 32];
          const uint32_t Af[4] = {Ad[j].x, Ad[j].y, Ad[j].z, Ad[j].w};
          mma_bf16_16816(acc, Af, b.x, b.y);
        }
      }
      reinterpret_cast<float4*>(red)[warp * 32 + lane] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
    if (e_act) {
      // conv[r][c] = sum_k x[r + k] cw[c][k], x[i] = wbuf[cm + i]: a Hankel matrix times the weights.
      // A fragment of k-tile kt: a0 = y(2kt), a1 = a2 = y(2kt+1), a3 = y(2kt+2), y(j) = x[g + 2tig + 8j .. +1]
      float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
      const float* xb = wbuf + cm + gq + 2 * tig;
      for (int kt = ckt_lo + e_wi; kt <= ckt_hi; kt += g.WPT) {
        const float* xk = xb + 16 * kt;
        uint32_t Ah[4], Al[4];
        split_bf16x2(xk[0], xk[1], Ah[0], Al[0]);
        split_bf16x2(xk[8], xk[9], Ah[1], Al[1]);
        split_bf16x2(xk[16], xk[17], Ah[3], Al[3]);
        Ah[2] = Ah[1]; Al[2] = Al[1];
        const uint4 b0 = cwB[(kt * 2) * 32 + lane];
        mma_bf16_16816(c0, Ah, b0.x, b0.y);
        mma_bf16_16816(c0, Al, b0.x, b0.y);
        mma_bf16_16816(c0, Ah, b0.z, b0.w);
        if (g.NC > 1) {
          const uint4 b1 = cwB[(kt * 2 + 1) * 32 + lane];
          mma_bf16_16816(c1, Ah, b1.x, b1.y);
          mma_bf16_16816(c1, Al, b1.x, b1.y);
          mma_bf16_16816(c1, Ah, b1.z, b1.w);
        }
      }
      cred[(warp * 2) * 32 + lane] = make_float4(c0[0], c0[1], c0[2], c0[3]);
      cred[(warp * 2 + 1) * 32 + lane] = make_float4(c1[0], c1[1], c1[2], c1[3]);
    }
    __syncthreads();
    float dz_v = 0.f;
    if (depi_ok) {
      dz_v = red_gather(red, d_w0, g.KSd, al & 15, n_d) + pbar_d;
      const uint32_t off = dzv_base + 4u * a_d;
      for (int qq = 0; qq < G; ++qq) st_remote_f32(mapa(off, n_d * G + qq), dz_v);
    }
    DTRACE(5);
    cluster_barrier();
    DTRACE(6);
    if (depi_ok) dzf_ptr[0] = dz_v;
    dzf_ptr += A;

    // ================= P3: energies of my frames -> all owners of the utterance =================
    if (e_act) {
      const int r0 = 16 * e_tt + gq, r1 = r0 + 8;
      // conv features in A-fragment position: sum of the K-split partials of my frame tile
      float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int w = 0; w < g.WPT; ++w) {
        const float4 a = cred[((e_tt * g.WPT + w) * 2) * 32 + lane], b = cred[((e_tt * g.WPT + w) * 2 + 1) * 32 + lane];
        s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
        s1.x += b.x; s1.y += b.y; s1.z += b.z; s1.w += b.w;
      }
      if (e_wi == 0 && csave_ptr) {
        if (r0 < ntl) {
          *reinterpret_cast<float2*>(csave_ptr + gq * 16 + 2 * tig) = make_float2(s0.x, s0.y);
          *reinterpret_cast<float2*>(csave_ptr + gq * 16 + 8 + 2 * tig) = make_float2(s1.x, s1.y);
        }
        if (r1 < ntl) {
          *reinterpret_cast<float2*>(csave_ptr + (gq + 8) * 16 + 2 * tig) = make_float2(s0.z, s0.w);
          *reinterpret_cast<float2*>(csave_ptr + (gq + 8) * 16 + 8 + 2 * tig) = make_float2(s1.z, s1.w);
        }
      }
      uint32_t Ah[4], Al[4];
      split_bf16x2(s0.x, s0.y, Ah[0], Al[0]);
      split_bf16x2(s0.z, s0.w, Ah[1], Al[1]);
      split_bf16x2(s1.x, s1.y, Ah[2], Al[2]);
      split_bf16x2(s1.z, s1.w, Ah[3], Al[3]);
      float e0 = 0.f, e1 = 0.f;
      const int nt_end = min((e_wi + 1) * g.NTW, g.AT8);
      const __nv_bfloat16* P0 = P_s + r0 * g.Pld + 2 * tig;
      const __nv_bfloat16* P1 = P_s + r1 * g.Pld + 2 * tig;
#pragma unroll 2
      for (int nt = e_wi * g.NTW; nt < nt_end; ++nt) {
        const uint4 bm = mattB[nt * 32 + lane];
        const int a = 8 * nt;
        const float2 p0 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(P0 + a));
        const float2 p1 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(P1 + a));
        float acc[4] = {p0.x, p0.y, p1.x, p1.y};      // the accumulator starts at P[te][a]
        mma_bf16_16816(acc, Ah, bm.x, bm.y);
        mma_bf16_16816(acc, Al, bm.x, bm.y);
        mma_bf16_16816(acc, Ah, bm.z, bm.w);
        const float2 dz2 = *reinterpret_cast<const float2*>(dzv + a + 2 * tig);
        const float2 gv2 = *reinterpret_cast<const float2*>(gv_s + a + 2 * tig);
        e0 = fmaf(gv2.x, tanh_fast(acc[0] + dz2.x), e0);
        e0 = fmaf(gv2.y, tanh_fast(acc[1] + dz2.y), e0);
        e1 = fmaf(gv2.x, tanh_fast(acc[2] + dz2.x), e1);
        e1 = fmaf(gv2.y, tanh_fast(acc[3] + dz2.y), e1);
      }
      e0 += __shfl_xor_sync(0xffffffffu, e0, 1);
      e0 += __shfl_xor_sync(0xffffffffu, e0, 2);
      e1 += __shfl_xor_sync(0xffffffffu, e1, 1);
      e1 += __shfl_xor_sync(0xffffffffu, e1, 2);
      if (tig == 0) {
        epart[e_wi * g.TT * 16 + r0] = e0;
        epart[e_wi * g.TT * 16 + r1] = e1;
      }
    }
    if (csave_ptr) csave_ptr += Te * 16;
    DTRACE(7);
    __syncthreads();
    DTRACE(8);
    if (tid < ntl) {
      float e = 0.f;
      for (int w = 0; w < g.WPT; ++w) e += epart[w * g.TT * 16 + tid];
      e *= scal;
      const uint32_t off = eall_base + 4u * (te0 + tid);
      for (int qq = 0; qq < G; ++qq) st_remote_f32(mapa(off, n_own * G + qq), e);
    }
    DTRACE(9);
    cluster_barrier();
    DTRACE(10);

    // ================= P4: softmax over all Te frames (every owner), context slice, broadcast =================
    float w_keep = 0.f, ca_keep = 0.f, cb_keep = 0.f;
    __nv_bfloat16 ha_keep = __float2bfloat16(0.f), hb_keep = ha_keep;
    bool va_keep = false, vb_keep = false;
    {
      const bool fr = own_ok && tid < Te;
      const float e = fr ? e_all[tid] : -INFINITY;
      float mx = warp_max(e);
      if (lane == 0) wred[warp] = mx;
      __syncthreads();
      float M = wred[0];
#pragma unroll
      for (int w = 1; w < kWarps; ++w) M = fmaxf(M, wred[w]);
      const float pv = fr ? __expf(e - M) : 0.f;
      if (fr) p_un[tid] = pv;
      const float sm_ = warp_sum(pv);
      if (lane == 0) wred[kWarps + warp] = sm_;
      __syncthreads();
      float S = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) S += wred[kWarps + w];
      const float invS = own_ok ? 1.f / S : 0.f;
      if (fr) {
        w_keep = pv * invS;
        wbuf[K + tid] = w_keep;
      }
      // context slice on tensor cores: c[o] = sum_te QT[o][te] p[te]  (M = my context dims, K = frames, N = column 0)
      if (own_ok && c_mt < g.OTs) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const uint32_t* q0 = reinterpret_cast<const uint32_t*>(QT_s + o_l0 * g.QTld) + tig;
        const uint32_t* q1 = reinterpret_cast<const uint32_t*>(QT_s + (o_l0 + 8) * g.QTld) + tig;
        for (int kt = 0; kt < g.KTe; ++kt) {
          const uint32_t Af[4] = {q0[kt * 8], q1[kt * 8], q0[kt * 8 + 4], q1[kt * 8 + 4]};
          uint32_t bh0 = 0u, bh1 = 0u, bl0 = 0u, bl1 = 0u;
          if (gq == 0) {
            const float2 v0 = *reinterpret_cast<const float2*>(p_un + 16 * kt + 2 * tig);
            const float2 v1 = *reinterpret_cast<const float2*>(p_un + 16 * kt + 2 * tig + 8);
            split_bf16x2(v0.x, v0.y, bh0, bl0);
            split_bf16x2(v1.x, v1.y, bh1, bl1);
          }
          mma_bf16_16816(acc, Af, bh0, bh1);
          mma_bf16_16816(acc, Af, bl0, bl1);
        }
        // lanes with tig == 0 hold column 0: rows o_l0 (acc[0]) and o_l0 + 8 (acc[2])
        const int oa = q * OS + o_l0, ob = oa + 8;
        const bool va = tig == 0 && o_l0 < OS, vb = tig == 0 && o_l0 + 8 < OS;
        const float ca = acc[0] * invS, cb = acc[2] * invS;
        __nv_bfloat16 ha = __float2bfloat16(0.f), hb = ha;
        if (va) ha = __float2bfloat16(ca + p.cbias[static_cast<int64_t>(b_own) * O + oa]);
        if (vb) hb = __float2bfloat16(cb + p.cbias[static_cast<int64_t>(b_own) * O + ob]);
        va_keep = va; vb_keep = vb; ca_keep = ca; cb_keep = cb; ha_keep = ha; hb_keep = hb;
        if (drop_on) {
          // what the NEXT step's gates see is dropout(c_t); the output layer (zc in global memory) sees c_t
          const unsigned long long seed = *p.seed_dev, base = (static_cast<unsigned long long>(b_own) * R + t + 1) * O;
          const float sc = 1.f / (1.f - p.drop_p);
          if (va) ha = dropout_keep(seed, p.drop_site, base + oa, p.drop_p) ? __float2bfloat16(__bfloat162float(ha) * sc) : __float2bfloat16(0.f);
          if (vb) hb = dropout_keep(seed, p.drop_site, base + ob, p.drop_p) ? __float2bfloat16(__bfloat162float(hb) * sc) : __float2bfloat16(0.f);
        }
        // The 8 lanes with tig == 0 hold dims gq (wa) and gq + 8 (wb) of this warp's 16-dim tile = one (k-tile,
        // utterance) row of the state buffer, 32 contiguous bytes. Three xor-shuffle rounds give every one of them
        // the whole row; lane gq pushes it to CTAs 2gq and 2gq+1 as two 16-byte stores each.
        const uint32_t wa = __bfloat16_as_ushort(ha), wb = __bfloat16_as_ushort(hb);
        const uint32_t a1 = __shfl_xor_sync(0xffffffffu, wa, 4), b1 = __shfl_xor_sync(0xffffffffu, wb, 4);
        const uint32_t pa = (gq & 1) ? (a1 | (wa << 16)) : (wa | (a1 << 16));
        const uint32_t pb = (gq & 1) ? (b1 | (wb << 16)) : (wb | (b1 << 16));
        const uint32_t a2 = __shfl_xor_sync(0xffffffffu, pa, 8), b2 = __shfl_xor_sync(0xffffffffu, pb, 8);
        const uint32_t qa0 = (gq & 2) ? a2 : pa, qa1 = (gq & 2) ? pa : a2;
        const uint32_t qb0 = (gq & 2) ? b2 : pb, qb1 = (gq & 2) ? pb : b2;
        const uint32_t xa0 = __shfl_xor_sync(0xffffffffu, qa0, 16), xa1 = __shfl_xor_sync(0xffffffffu, qa1, 16);
        const uint32_t xb0 = __shfl_xor_sync(0xffffffffu, qb0, 16), xb1 = __shfl_xor_sync(0xffffffffu, qb1, 16);
        if (tig == 0 && va) {     // OS % 16 == 0 on this path: every tile is full
          const bool hi = (gq & 4) != 0;
          const uint32_t A0 = hi ? xa0 : qa0, A1 = hi ? xa1 : qa1, A2 = hi ? qa0 : xa0, A3 = hi ? qa1 : xa1;
          const uint32_t B0 = hi ? xb0 : qb0, B1 = hi ? xb1 : qb1, B2 = hi ? qb0 : xb0, B3 = hi ? qb1 : xb1;
          const uint32_t off = zB_base + 4u * static_cast<uint32_t>(zb_nxt_w + qfrag_word(Hd + q * OS + 16 * c_mt, n_own));
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const uint32_t dst = mapa(off, 2 * gq + i);
            st_remote_v4_u32(dst, A0, A1, A2, A3);
            st_remote_v4_u32(dst + 16u, B0, B1, B2, B3);
          }
        }
      }
    }
    DTRACE(11);
    cluster_barrier();
    DTRACE(12);
    if (own_ok) {
      if (tid >= te0 && tid < te0 + ntl) ws_row[tid] = w_keep;      // tid < Te holds: te0 + ntl <= Te
      if (va_keep | vb_keep) {
        const int oa = q * OS + o_l0, ob = oa + 8;
        __nv_bfloat16* zrow = p.zc + (static_cast<int64_t>(b_own) * R + t + 1) * ZC + Hd;
        float* crow = p.cpre ? p.cpre + (static_cast<int64_t>(b_own) * L + t) * O : nullptr;
        if (va_keep) {
          if (crow) crow[oa] = ca_keep;
          zrow[oa] = ha_keep;
        }
        if (vb_keep) {
          if (crow) crow[ob] = cb_keep;
          zrow[ob] = hb_keep;
        }
      }
    }
    ws_row += Te;
  }
#undef DTRACE
}

// ==========================================================================================
// backward (BPTT through the teacher-forced decoder), same cluster organisation, 384 threads
// (more registers per thread: the row-sharded W^T fragments need 96 of them).
//
//   A  d[z_t; c_t] rows of this CTA = dzc_all_t + Wr^T dgates_{t+1}   (dgates all-gathered in the
//      previous iteration) -> dc rows to the owners of each utterance             | barrier
//   B  owners: dw = Q dc + (conv-input gradient from step t+1); softmax backward; energy backward
//      with tanh recomputed on tensor cores (P, dz_t, conv_t saved by the forward); ddz partial
//      -> all CTAs; conv-input gradient for step t-1 -> sibling owners            | barrier
//   C  dz_t += mlp_dec^T ddz; LSTM cell backward -> dgates_t -> all CTAs          | barrier
//
// Parameter gradients that are plain sums over (b, t) are NOT accumulated here: the kernel saves
// de_t, dc_t, ddz_t, dgates_t, dconv_t and the post-loop kernels / GEMMs reduce them in parallel.
// ==========================================================================================
constexpr int kBT = 384;
constexpr int kBW = 12;
constexpr int kMaxFB = 20;     // Wr^T fragments per warp
constexpr int kMaxFD2 = 4;     // mlp_dec^T fragments per warp
constexpr int kCwLd = 20;      // row stride (floats) of the transposed conv weights: conflict-free LDS.128 over 8 consecutive taps

struct BGeom {
  int NB, G, UPC, OPC, RPC;    // utterances per cluster, owners per utterance, z units / c dims / rows per CTA
  int MTb, KTb, KSb, FB;       // Wr^T: m-tiles, k-tiles (4Hd/16), K-splits, fragments per warp
  int MTd, KTa, KSd, FD, KTap; // mlp_dec^T: m-tiles, k-tiles (A/16), K-splits, fragments per warp, padded k-tiles
  int TR, TT, WPT, NTW2;       // frames per owner, 16-frame tiles, warps per tile, 16-wide attention k-tiles per warp
  int KTo, AT8, NC, NT8;       // O/16; A/8; conv channel n-tiles; 8-wide tiles over the conv taps
  int Pld, Qld;
  int o_dgB, o_red, o_dcbuf, o_P, o_Q, o_matt, o_matt2, o_cwB2, o_dwnrx, o_ddzrx, o_ddzB, o_wt, o_cpre, o_dzv,
      o_conv, o_de, o_dwpart, o_dwns, o_dwnout, o_dwnpart, o_gv, o_wred, o_scratch;
  int smem;
};

bool dec_bgeom(const las_dec_args* a, int NB, BGeom& g) {
  const int Hd = a->Hd, O = a->O, A = a->A, Te = a->Te;
  if (Hd % 64 != 0 || Hd > 320 || O % 16 != 0 || A % 16 != 0 || A > 512 || a->C > 16 || Te > kBT) return false;
  g.NB = NB; g.G = kCS / NB;
  g.UPC = Hd / kCS; g.OPC = O / kCS; g.RPC = g.UPC + g.OPC;
  if (g.RPC * NB > kBT) return false;
  g.MTb = (g.RPC + 15) / 16; g.KTb = 4 * Hd / 16;
  g.KSb = kBW / g.MTb; if (g.KSb < 1) return false;
  g.FB = (g.KTb + g.KSb - 1) / g.KSb;
  g.MTd = (g.UPC + 15) / 16; g.KTa = A / 16;
  g.KSd = kBW / g.MTd; if (g.KSd > g.KTa) g.KSd = g.KTa;
  g.FD = (g.KTa + g.KSd - 1) / g.KSd;
  g.KTap = g.KSd * g.FD;
  if (g.FB > kMaxFB || g.FD > kMaxFD2) return false;
  g.TR = (Te + g.G - 1) / g.G;
  g.TT = (g.TR + 15) / 16;
  if (g.TT > kBW) return false;
  g.WPT = kBW / g.TT;
  g.NTW2 = (g.KTa + g.WPT - 1) / g.WPT;
  g.KTo = O / 16; g.AT8 = A / 8; g.NC = a->C > 8 ? 2 : 1;
  const int ksz = 2 * a->K + 1;
  g.NT8 = (ksz + 7) / 8;
  g.Pld = A + ((8 - A % 16) + 16) % 16;
  g.Qld = O + ((8 - O % 16) + 16) % 16;
  int off = 0;
  auto take = [&](int bytes) { const int o = off; off += rup(bytes, 16); return o; };
  g.o_dgB = take(g.KSb * g.FB * 256);
  g.o_red = take(kBW * 128 * 4);
  g.o_dcbuf = take(O * 4);
  g.o_P = take(g.TT * 16 * g.Pld * 2);
  g.o_Q = take(g.TT * 16 * g.Qld * 2);
  g.o_matt = take(g.AT8 * 32 * 16);
  g.o_matt2 = take(g.KTa * 2 * 32 * 8);
  g.o_cwB2 = take(ksz * kCwLd * 4);
  g.o_dwnrx = take(2 * g.G * Te * 4);
  g.o_ddzrx = take(g.G * NB * A * 4);
  g.o_ddzB = take(g.KTap * 256);
  g.o_wt = take(Te * 4);
  g.o_cpre = take(O * 4);
  g.o_dzv = take(A * 4);
  g.o_conv = take(g.TT * 16 * 16 * 4);
  g.o_de = take(g.TT * 16 * 4);
  g.o_dwpart = take(g.WPT * g.TT * 16 * 4);
  g.o_dwns = take(Te * 4);
  g.o_dwnout = take(Te * 4);
  g.o_dwnpart = take((kBT / Te > 0 ? kBT / Te : 1) * Te * 4);
  g.o_gv = take(A * 4);
  g.o_wred = take(kBW * 4);
  // phase-B scratch: ddz partials [TT][A] f32 + dconv partial accumulators [12 warps][2][32] float4
  g.o_scratch = take(g.TT * A * 4 + kBW * 2 * 32 * 16);
  g.smem = off;
  return g.smem <= 220 * 1024;
}

struct DecBwdP {
  int B, L, Te, Hd, O, A, C, K;
  float att_scaling;
  BGeom g;
  const float* P; const __nv_bfloat16* Q;
  const uint32_t* wrT_pk;        // [16][MTb][KTb][32][4]
  const uint32_t* decT_pk;       // [16][MTd][KTa][32][4]
  const float* conv_w; const float* mlp_att; const float* gvec;
  const float* ws; const __half* gates_save; const float* c_save; const float* dzf; const float* cpre;
  const float* conv_save; const float* dzc_all;
  __nv_bfloat16* dgates;         // [B, L+1, 4Hd] (row t)
  __nv_bfloat16* dcz_all;        // [B, L+1, Hd+O] (row t+1, context part)
  float* dc_all;                 // [B, L, O]
  float* ddz_all;                // [B, L+1, A] (row t+1)
  float* de_all;                 // [B, L, Te]
  float* dattc_all;              // [L, B, Te, C]
  float drop_p; uint32_t drop_site; const unsigned long long* seed_dev;
  long long* dbg;
};

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__global__ void __launch_bounds__(kBT, 1) dec_persist_bwd_kernel(const __grid_constant__ DecBwdP p_in) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ DecBwdP p;
  for (int i = threadIdx.x; i < static_cast<int>(sizeof(DecBwdP) / 4); i += kBT)
    reinterpret_cast<uint32_t*>(&p)[i] = reinterpret_cast<const uint32_t*>(&p_in)[i];
  __syncthreads();
  const BGeom& g = p.g;
  uint32_t* dgB = reinterpret_cast<uint32_t*>(smem + g.o_dgB);     // [KSb*FB][32][2] all-gathered dgates_{t+1}
  float* red = reinterpret_cast<float*>(smem + g.o_red);           // [12][32][4]
  float* dcbuf = reinterpret_cast<float*>(smem + g.o_dcbuf);       // [O] dc_t of my utterance
  __nv_bfloat16* P_s = reinterpret_cast<__nv_bfloat16*>(smem + g.o_P);
  __nv_bfloat16* Q_s = reinterpret_cast<__nv_bfloat16*>(smem + g.o_Q);   // [TT*16][Qld] my frames of Q
  uint4* mattB = reinterpret_cast<uint4*>(smem + g.o_matt);        // [AT8][32]      k = channel, n = attention dim
  uint2* mattB2 = reinterpret_cast<uint2*>(smem + g.o_matt2);      // [KTa][2][32]   k = attention dim, n = channel
  float* cw_t = reinterpret_cast<float*>(smem + g.o_cwB2);         // [2K+1][kCwLd] transposed conv weights (channels padded to 16)
  float* dwn_part = reinterpret_cast<float*>(smem + g.o_dwnpart);  // [kBT/Te][Te]
  float* dwn_rx = reinterpret_cast<float*>(smem + g.o_dwnrx);      // [2][G][Te]
  float* ddz_rx = reinterpret_cast<float*>(smem + g.o_ddzrx);      // [G][NB][A] f32 partials (owner partials cancel: bf16 is not enough)
  uint32_t* ddzB = reinterpret_cast<uint32_t*>(smem + g.o_ddzB);   // [KTap][32][2]
  float* wt_s = reinterpret_cast<float*>(smem + g.o_wt);           // [Te] w_t
  float* cpre_s = reinterpret_cast<float*>(smem + g.o_cpre);       // [O]
  float* dzv = reinterpret_cast<float*>(smem + g.o_dzv);           // [A]
  float* conv_s = reinterpret_cast<float*>(smem + g.o_conv);       // [TT*16][16]
  float* de_s = reinterpret_cast<float*>(smem + g.o_de);           // [TT*16]
  float* dwpart = reinterpret_cast<float*>(smem + g.o_dwpart);     // [WPT][TT*16]
  float* dwn_s = reinterpret_cast<float*>(smem + g.o_dwns);        // [Te]
  float* dwn_out = reinterpret_cast<float*>(smem + g.o_dwnout);    // [Te]
  float* gv_s = reinterpret_cast<float*>(smem + g.o_gv);
  float* wred = reinterpret_cast<float*>(smem + g.o_wred);
  float* ddz_part = reinterpret_cast<float*>(smem + g.o_scratch);  // [TT][A]
  float4* dcred = reinterpret_cast<float4*>(smem + g.o_scratch + g.TT * p.A * 4);   // [12][2][32]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tig = lane & 3;
  const uint32_t rank = cluster_rank();
  const int cl = blockIdx.y;
  const int NB = g.NB, G = g.G, UPC = g.UPC, OPC = g.OPC, RPC = g.RPC, TR = g.TR;
  const int Hd = p.Hd, O = p.O, A = p.A, Te = p.Te, L = p.L, K = p.K, C = p.C;
  const int ZC = Hd + O, R = L + 1, ksz = 2 * K + 1, A2 = A >> 1;
  const int n_own = rank / G, q = rank % G;
  const int b_own = cl * NB + n_own;
  const bool own_ok = b_own < p.B;
  const int te0 = q * TR;
  const int ntl = own_ok ? max(0, min(TR, Te - te0)) : 0;
  const float scal = p.att_scaling;

  // ---------------- resident weights (registers)
  const int b_tile = warp / g.KSb, b_ks = warp % g.KSb;
  const bool b_act = b_tile < g.MTb;
  const int b_kt0 = b_ks * g.FB;
  uint4 Ab[kMaxFB];
#pragma unroll
  for (int j = 0; j < kMaxFB; ++j) {
    Ab[j] = make_uint4(0u, 0u, 0u, 0u);
    if (b_act && j < g.FB && b_kt0 + j < g.KTb)
      Ab[j] = __ldg(reinterpret_cast<const uint4*>(p.wrT_pk) +
                    (static_cast<int64_t>(rank * g.MTb + b_tile) * g.KTb + b_kt0 + j) * 32 + lane);
  }
  const int d_tile = warp / g.KSd, d_ks = warp % g.KSd;
  const bool d_act = d_tile < g.MTd;
  const int d_kt0 = d_ks * g.FD;
  uint4 Ad[kMaxFD2];
#pragma unroll
  for (int j = 0; j < kMaxFD2; ++j) {
    Ad[j] = make_uint4(0u, 0u, 0u, 0u);
    if (d_act && j < g.FD && d_kt0 + j < g.KTa)
      Ad[j] = __ldg(reinterpret_cast<const uint4*>(p.decT_pk) +
                    (static_cast<int64_t>(rank * g.MTd + d_tile) * g.KTa + d_kt0 + j) * 32 + lane);
  }
  const int nfb = g.FB, nfd = g.FD;

  // ---------------- resident operands (shared memory)
  for (int i = tid; i < g.KSb * g.FB * 64; i += kBT) dgB[i] = 0u;
  for (int i = tid; i < g.KTap * 64; i += kBT) ddzB[i] = 0u;
  for (int i = tid; i < 2 * G * Te; i += kBT) dwn_rx[i] = 0.f;
  for (int i = tid; i < G * NB * A; i += kBT) ddz_rx[i] = 0.f;
  for (int i = tid; i < g.TT * 16 * 16; i += kBT) conv_s[i] = 0.f;
  for (int i = tid; i < g.TT * 16; i += kBT) de_s[i] = 0.f;
  for (int i = tid; i < A; i += kBT) gv_s[i] = p.gvec[i];
  for (int i = tid; i < g.AT8 * 32; i += kBT) {
    const int nt = i >> 5, l = i & 31, a = 8 * nt + (l >> 2), c0 = 2 * (l & 3);
    float m[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + (k & 1) + 8 * (k >> 1);
      m[k] = (c < C) ? p.mlp_att[a * C + c] : 0.f;
    }
    uint4 v;
    split_bf16x2(m[0], m[1], v.x, v.z);
    split_bf16x2(m[2], m[3], v.y, v.w);
    mattB[i] = v;
  }
  for (int i = tid; i < g.KTa * 64; i += kBT) {
    // k = attention dim 16*kt + .., n = channel 8*nc + (l >> 2)
    const int kt = i >> 6, nc = (i >> 5) & 1, l = i & 31, c = 8 * nc + (l >> 2), a0 = 16 * kt + 2 * (l & 3);
    float m[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int a = a0 + (k & 1) + 8 * (k >> 1);
      m[k] = (c < C && a < A) ? p.mlp_att[a * C + c] : 0.f;
    }
    mattB2[i] = make_uint2(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]));
  }
  for (int i = tid; i < ksz * kCwLd; i += kBT) {
    const int k = i / kCwLd, c = i % kCwLd;
    cw_t[i] = (c < C) ? p.conv_w[c * ksz + k] : 0.f;
  }
  for (int i = tid; i < g.TT * 16 * g.Pld; i += kBT) {
    const int r = i / g.Pld, a = i % g.Pld;
    P_s[i] = __float2bfloat16((r < ntl && a < A) ? p.P[(static_cast<int64_t>(b_own) * Te + te0 + r) * A + a] : 0.f);
  }
  for (int i = tid; i < g.TT * 16 * g.Qld; i += kBT) {
    const int r = i / g.Qld, o = i % g.Qld;
    Q_s[i] = (r < ntl && o < O) ? p.Q[(static_cast<int64_t>(b_own) * Te + te0 + r) * O + o] : __float2bfloat16(0.f);
  }

  // ---------------- roles
  // phase A / C epilogue: thread -> (utterance n, local row); rows < UPC are hidden units, the rest context dims
  const int n_e = tid / RPC, row_l = tid % RPC;
  const bool epi = tid < RPC * NB;
  const int b_e = cl * NB + n_e;
  const bool epi_ok = epi && b_e < p.B;
  const bool is_z = row_l < UPC;
  const int u_e = rank * UPC + row_l;                    // hidden unit (is_z)
  const int o_e = rank * OPC + (row_l - UPC);            // context dim (!is_z)
  const int col_e = is_z ? u_e : Hd + o_e;               // column of [z; c]
  const int a_w0 = (row_l >> 4) * g.KSb;                 // first contributing warp (phase A)
  const int c_w0 = (row_l >> 4) * g.KSd;                 // first contributing warp (phase C, is_z rows)
  float dcell = 0.f, dz_acc = 0.f;
  const int bb = epi_ok ? b_e : 0;
  const float* dzc_ptr = p.dzc_all + (static_cast<int64_t>(bb) * R + L) * ZC + col_e;          // row t+1, t = L-1
  int64_t sv_idx = (static_cast<int64_t>(bb) * L + (L - 1)) * Hd + (is_z ? u_e : 0);           // gates / c_save of step t
  __nv_bfloat16* dg_ptr = p.dgates + (static_cast<int64_t>(bb) * R + (L - 1)) * 4 * Hd + (is_z ? u_e : 0);
  // phase B
  const int e_tt = warp / g.WPT, e_wi = warp % g.WPT;
  const bool e_act = warp < g.TT * g.WPT && ntl > 0;
  const int cm = te0 + 16 * e_tt;                        // first frame of my frame tile
  const int bo = own_ok ? b_own : 0;
  // cluster-mapped bases
  const uint32_t dgB_base = smem_u32(dgB), dcbuf_base = smem_u32(dcbuf), dwnrx_base = smem_u32(dwn_rx),
                 ddzrx_base = smem_u32(ddz_rx);

  __syncthreads();
  cluster_barrier();
  const bool trace = p.dbg != nullptr && blockIdx.y == 0 && rank == 0 && tid == 0;
#define DTRACE(slot) do { if (trace && t <= L - 9 && t > L - 13) p.dbg[64 + (L - 9 - t) * 16 + (slot)] = clock64(); } while (0)

  for (int t = L - 1; t >= 0; --t) {
    const int par = t & 1;
    DTRACE(0);
    // ---------------- prefetch of this step's saved activations (consumed in phases B and C)
    if (own_ok) {
      const float* wrow = p.ws + (static_cast<int64_t>(bo) * R + t + 1) * Te;
      for (int i = tid; i < Te; i += kBT) cp_async4(wt_s + i, wrow + i);
      const float* crow = p.cpre + (static_cast<int64_t>(bo) * L + t) * O;
      for (int i = tid; i < O / 4; i += kBT) cp_async16(cpre_s + 4 * i, crow + 4 * i);
      const float* zrow = p.dzf + (static_cast<int64_t>(bo) * L + t) * A;
      for (int i = tid; i < A / 4; i += kBT) cp_async16(dzv + 4 * i, zrow + 4 * i);
      const float* cvrow = p.conv_save + ((static_cast<int64_t>(bo) * L + t) * Te + te0) * 16;
      for (int i = tid; i < ntl * 4; i += kBT) cp_async16(conv_s + 4 * i, cvrow + 4 * i);
    }
    float dzc_v = 0.f, c_cur = 0.f, c_prev = 0.f;
    uint2 gpk = make_uint2(0u, 0u);
    if (epi_ok) {
      dzc_v = __ldg(dzc_ptr);
      if (is_z) {
        gpk = __ldg(reinterpret_cast<const uint2*>(p.gates_save) + sv_idx);
        c_cur = __ldg(p.c_save + sv_idx);
        if (t > 0) c_prev = __ldg(p.c_save + sv_idx - Hd);
      }
    }
    // ================= phase A: d[z_t; c_t] rows = dzc_all + Wr^T dgates_{t+1} =================
    if (b_act) {
      float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
      const uint2* hb = reinterpret_cast<const uint2*>(dgB) + lane + b_kt0 * 32;
#pragma unroll
      for (int j = 0; j < kMaxFB; ++j) {
        if (j < nfb) {
          const uint2 b = hb[j * 32];
          const uint32_t Af[4] = {Ab[j].x, Ab[j].y, Ab[j].z, Ab[j].w};
          if (j & 1) mma_bf16_16816(acc1, Af, b.x, b.y);
          else mma_bf16_16816(acc0, Af, b.x, b.y);
        }
      }
      reinterpret_cast<float4*>(red)[warp * 32 + lane] =
          make_float4(acc0[0] + acc1[0], acc0[1] + acc1[1], acc0[2] + acc1[2], acc0[3] + acc1[3]);
    }
    __syncthreads();
    float dc_keep = 0.f;
    if (epi_ok) {
      float mm = red_gather(red, a_w0, g.KSb, row_l & 15, n_e);
      if (p.drop_p > 0.f && !is_z)   // the path through the cell input of step t+1 carries that step's dropout mask
        mm = dropout_keep(*p.seed_dev, p.drop_site, (static_cast<unsigned long long>(b_e) * R + t + 1) * O + o_e, p.drop_p)
                 ? mm / (1.f - p.drop_p) : 0.f;
      const float v = mm + dzc_v;
      if (is_z) {
        dz_acc = v;
      } else {
        dc_keep = v;
        const uint32_t off = dcbuf_base + 4u * o_e;
        for (int qq = 0; qq < G; ++qq) st_remote_f32(mapa(off, n_e * G + qq), v);
      }
    }
    dzc_ptr -= ZC;
    cp_async_wait_all();
    DTRACE(1);
    cluster_barrier();
    DTRACE(2);
    // global stores follow the barrier (its release fence would wait for their acknowledgements; see the forward kernel)
    if (epi_ok && !is_z) {
      p.dc_all[(static_cast<int64_t>(b_e) * L + t) * O + o_e] = dc_keep;
      p.dcz_all[(static_cast<int64_t>(b_e) * R + t + 1) * ZC + Hd + o_e] = __float2bfloat16(dc_keep);
    }

    // ================= phase B: attention backward for my frames =================
    // B1: dw = Q dc (tensor cores, K split over the warps of a frame tile); softmax dot product
    {
      float part = 0.f;
      if (own_ok) {
        if (tid < O) part = dcbuf[tid] * cpre_s[tid];
        for (int o2 = tid + kBT; o2 < O; o2 += kBT) part = fmaf(dcbuf[o2], cpre_s[o2], part);
        if (tid < Te) {
          float dn = 0.f;
          for (int qq = 0; qq < G; ++qq) dn += dwn_rx[(par * G + qq) * Te + tid];
          dwn_s[tid] = dn;
          part = fmaf(wt_s[tid], dn, part);
        }
      }
      part = warp_sum(part);
      if (lane == 0) wred[warp] = part;
    }
    if (e_act) {
      float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t* q0 = reinterpret_cast<const uint32_t*>(Q_s + (16 * e_tt + gq) * g.Qld) + tig;
      const uint32_t* q1 = reinterpret_cast<const uint32_t*>(Q_s + (16 * e_tt + gq + 8) * g.Qld) + tig;
      for (int kt = e_wi; kt < g.KTo; kt += g.WPT) {
        const uint32_t Af[4] = {q0[kt * 8], q1[kt * 8], q0[kt * 8 + 4], q1[kt * 8 + 4]};
        uint32_t bh0 = 0u, bh1 = 0u, bl0 = 0u, bl1 = 0u;
        if (gq == 0) {
          const float2 v0 = *reinterpret_cast<const float2*>(dcbuf + 16 * kt + 2 * tig);
          const float2 v1 = *reinterpret_cast<const float2*>(dcbuf + 16 * kt + 2 * tig + 8);
          split_bf16x2(v0.x, v0.y, bh0, bl0);
          split_bf16x2(v1.x, v1.y, bh1, bl1);
        }
        mma_bf16_16816(acc0, Af, bh0, bh1);
        mma_bf16_16816(acc1, Af, bl0, bl1);
      }
      if (tig == 0) {
        dwpart[e_wi * g.TT * 16 + 16 * e_tt + gq] = acc0[0] + acc1[0];
        dwpart[e_wi * g.TT * 16 + 16 * e_tt + gq + 8] = acc0[2] + acc1[2];
      }
    }
    __syncthreads();
    // B2: de = scal * w_t * (dw - <w_t, dw>)
    if (tid < g.TT * 16) {
      float de = 0.f;
      if (tid < ntl) {
        float dw = dwn_s[te0 + tid];
        for (int w = 0; w < g.WPT; ++w) dw += dwpart[w * g.TT * 16 + tid];
        float dot = 0.f;
#pragma unroll
        for (int w = 0; w < kBW; ++w) dot += wred[w];
        de = scal * wt_s[te0 + tid] * (dw - dot);
        p.de_all[(static_cast<int64_t>(b_own) * L + t) * Te + te0 + tid] = de;
      }
      de_s[tid] = de;
    }
    __syncthreads();
    DTRACE(3);
    // B3: energy backward, tanh recomputed: ds = de gv (1 - s^2); ddz[a] = sum_te ds; dconv = ds mlp_att
    if (e_act) {
      const int r0 = 16 * e_tt + gq, r1 = r0 + 8;
      uint32_t Ah[4], Al[4];
      {
        const float2 v0 = *reinterpret_cast<const float2*>(conv_s + r0 * 16 + 2 * tig);
        const float2 v1 = *reinterpret_cast<const float2*>(conv_s + r1 * 16 + 2 * tig);
        const float2 v2 = *reinterpret_cast<const float2*>(conv_s + r0 * 16 + 2 * tig + 8);
        const float2 v3 = *reinterpret_cast<const float2*>(conv_s + r1 * 16 + 2 * tig + 8);
        split_bf16x2(v0.x, v0.y, Ah[0], Al[0]);
        split_bf16x2(v1.x, v1.y, Ah[1], Al[1]);
        split_bf16x2(v2.x, v2.y, Ah[2], Al[2]);
        split_bf16x2(v3.x, v3.y, Ah[3], Al[3]);
      }
      const float de0 = de_s[r0], de1 = de_s[r1];
      float dcv0[4] = {0.f, 0.f, 0.f, 0.f}, dcv1[4] = {0.f, 0.f, 0.f, 0.f};
      const __nv_bfloat16* P0 = P_s + r0 * g.Pld + 2 * tig;
      const __nv_bfloat16* P1 = P_s + r1 * g.Pld + 2 * tig;
      const int kt_end = min((e_wi + 1) * g.NTW2, g.KTa);
      for (int kt2 = e_wi * g.NTW2; kt2 < kt_end; ++kt2) {
        uint32_t Dh[4], Dl[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int nt = 2 * kt2 + h, a = 8 * nt;
          const uint4 bm = mattB[nt * 32 + lane];
          const float2 p0 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(P0 + a));
          const float2 p1 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(P1 + a));
          float acc[4] = {p0.x, p0.y, p1.x, p1.y};
          mma_bf16_16816(acc, Ah, bm.x, bm.y);
          mma_bf16_16816(acc, Al, bm.x, bm.y);
          mma_bf16_16816(acc, Ah, bm.z, bm.w);
          const float2 dz2 = *reinterpret_cast<const float2*>(dzv + a + 2 * tig);
          const float2 gv2 = *reinterpret_cast<const float2*>(gv_s + a + 2 * tig);
          const float s00 = tanh_fast(acc[0] + dz2.x), s01 = tanh_fast(acc[1] + dz2.y);
          const float s10 = tanh_fast(acc[2] + dz2.x), s11 = tanh_fast(acc[3] + dz2.y);
          const float d00 = de0 * gv2.x * (1.f - s00 * s00), d01 = de0 * gv2.y * (1.f - s01 * s01);
          const float d10 = de1 * gv2.x * (1.f - s10 * s10), d11 = de1 * gv2.y * (1.f - s11 * s11);
          float v0 = d00 + d10, v1 = d01 + d11;       // column sums over my 2 rows, then over the 8 row groups
          v0 += __shfl_xor_sync(0xffffffffu, v0, 4);  v1 += __shfl_xor_sync(0xffffffffu, v1, 4);
          v0 += __shfl_xor_sync(0xffffffffu, v0, 8);  v1 += __shfl_xor_sync(0xffffffffu, v1, 8);
          v0 += __shfl_xor_sync(0xffffffffu, v0, 16); v1 += __shfl_xor_sync(0xffffffffu, v1, 16);
          if (gq == 0) *reinterpret_cast<float2*>(ddz_part + e_tt * A + a + 2 * tig) = make_float2(v0, v1);
          split_bf16x2(d00, d01, Dh[2 * h], Dl[2 * h]);
          split_bf16x2(d10, d11, Dh[2 * h + 1], Dl[2 * h + 1]);
        }
        const uint2 b0 = mattB2[(kt2 * 2) * 32 + lane];
        mma_bf16_16816(dcv0, Dh, b0.x, b0.y);
        mma_bf16_16816(dcv0, Dl, b0.x, b0.y);
        if (g.NC > 1) {
          const uint2 b1 = mattB2[(kt2 * 2 + 1) * 32 + lane];
          mma_bf16_16816(dcv1, Dh, b1.x, b1.y);
          mma_bf16_16816(dcv1, Dl, b1.x, b1.y);
        }
      }
      dcred[(warp * 2) * 32 + lane] = make_float4(dcv0[0], dcv0[1], dcv0[2], dcv0[3]);
      dcred[(warp * 2 + 1) * 32 + lane] = make_float4(dcv1[0], dcv1[1], dcv1[2], dcv1[3]);
    }
    __syncthreads();
    DTRACE(4);
    // B4: ddz partial -> all CTAs (f32); conv-input gradient of my frames for step t-1
    if (own_ok) {
      for (int aq = tid; aq < (A >> 2); aq += kBT) {      // 4 attention dims per thread: 16-byte remote stores
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ntl > 0)
          for (int tt = 0; tt < g.TT; ++tt) {
            const float4 x = *reinterpret_cast<const float4*>(ddz_part + tt * A + 4 * aq);
            v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
          }
        const uint32_t off = ddzrx_base + 4u * ((q * NB + n_own) * A + 4 * aq);
#pragma unroll
        for (int r = 0; r < kCS; ++r) st_remote_v4_f32(mapa(off, r), v.x, v.y, v.z, v.w);
      }
    }
    if (e_act && e_wi == 0) {
      // dconv of my frame tile: sum of the partials of the WPT warps -> conv_s (reused) and dattc_all
      float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int w = 0; w < g.WPT; ++w) {
        const float4 a = dcred[((e_tt * g.WPT + w) * 2) * 32 + lane], b = dcred[((e_tt * g.WPT + w) * 2 + 1) * 32 + lane];
        s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
        s1.x += b.x; s1.y += b.y; s1.z += b.z; s1.w += b.w;
      }
      const int r0 = 16 * e_tt + gq, r1 = r0 + 8, c0 = 2 * tig;
      *reinterpret_cast<float2*>(conv_s + r0 * 16 + c0) = make_float2(s0.x, s0.y);
      *reinterpret_cast<float2*>(conv_s + r1 * 16 + c0) = make_float2(s0.z, s0.w);
      *reinterpret_cast<float2*>(conv_s + r0 * 16 + 8 + c0) = make_float2(s1.x, s1.y);
      *reinterpret_cast<float2*>(conv_s + r1 * 16 + 8 + c0) = make_float2(s1.z, s1.w);
      // dattc_all[t][b][te][c] for the conv-weight gradient (post-loop kernel)
      float* drow = p.dattc_all + ((static_cast<int64_t>(t) * p.B + b_own) * Te + cm) * C;
      if (r0 < ntl) {
        if (c0 < C) drow[gq * C + c0] = s0.x;
        if (c0 + 1 < C) drow[gq * C + c0 + 1] = s0.y;
        if (c0 + 8 < C) drow[gq * C + c0 + 8] = s1.x;
        if (c0 + 9 < C) drow[gq * C + c0 + 9] = s1.y;
      }
      if (r1 < ntl) {
        if (c0 < C) drow[(gq + 8) * C + c0] = s0.z;
        if (c0 + 1 < C) drow[(gq + 8) * C + c0 + 1] = s0.w;
        if (c0 + 8 < C) drow[(gq + 8) * C + c0 + 8] = s1.z;
        if (c0 + 9 < C) drow[(gq + 8) * C + c0 + 9] = s1.w;
      }
    }
    __syncthreads();
    // conv-input gradient of my frames for step t-1: dwn[j] = sum_{tl,c} dconv[tl][c] cw[c][j - te + K].
    // thread -> (output frame j, frame residue); deterministic two-stage sum (no shared-memory float atomics)
    const int ngrp = max(1, kBT / Te);
    if (own_ok && t > 0 && tid < ngrp * Te) {
      const int gi = tid / Te, j = tid - gi * Te;
      float acc[12];
#pragma unroll
      for (int c = 0; c < 12; ++c) acc[c] = 0.f;
      float acc2[4] = {0.f, 0.f, 0.f, 0.f};
      const int nc4 = (C + 3) >> 2;
      for (int tl = gi; tl < ntl; tl += ngrp) {
        const int k = j - (te0 + tl) + K;
        if (k < 0 || k > 2 * K) continue;
        const float4* dr = reinterpret_cast<const float4*>(conv_s + tl * 16);
        const float4* cr = reinterpret_cast<const float4*>(cw_t + k * kCwLd);
        const float4 d0 = dr[0], d1 = dr[1], w0 = cr[0], w1 = cr[1];
        acc[0] = fmaf(d0.x, w0.x, acc[0]); acc[1] = fmaf(d0.y, w0.y, acc[1]); acc[2] = fmaf(d0.z, w0.z, acc[2]); acc[3] = fmaf(d0.w, w0.w, acc[3]);
        acc[4] = fmaf(d1.x, w1.x, acc[4]); acc[5] = fmaf(d1.y, w1.y, acc[5]); acc[6] = fmaf(d1.z, w1.z, acc[6]); acc[7] = fmaf(d1.w, w1.w, acc[7]);
        if (nc4 > 2) {
          const float4 d2 = dr[2], w2 = cr[2];
          acc[8] = fmaf(d2.x, w2.x, acc[8]); acc[9] = fmaf(d2.y, w2.y, acc[9]); acc[10] = fmaf(d2.z, w2.z, acc[10]); acc[11] = fmaf(d2.w, w2.w, acc[11]);
        }
        if (nc4 > 3) {
          const float4 d3 = dr[3], w3 = cr[3];
          acc2[0] = fmaf(d3.x, w3.x, acc2[0]); acc2[1] = fmaf(d3.y, w3.y, acc2[1]); acc2[2] = fmaf(d3.z, w3.z, acc2[2]); acc2[3] = fmaf(d3.w, w3.w, acc2[3]);
        }
      }
      float sacc = (acc2[0] + acc2[1]) + (acc2[2] + acc2[3]);
#pragma unroll
      for (int c = 0; c < 12; ++c) sacc += acc[c];
      dwn_part[gi * Te + j] = sacc;
    }
    __syncthreads();
    if (own_ok && t > 0) {
      // partial conv-input gradient -> every owner of this utterance (slot q, parity of step t-1)
      for (int j = tid; j < Te; j += kBT) {
        float v = 0.f;
        for (int gi = 0; gi < ngrp; ++gi) v += dwn_part[gi * Te + j];
        for (int qq = 0; qq < G; ++qq)
          st_remote_f32(mapa(dwnrx_base + 4u * (((par ^ 1) * G + q) * Te + j), n_own * G + qq), v);
      }
    }
    DTRACE(5);
    cluster_barrier();
    DTRACE(6);

    // ================= phase C: dz_t += mlp_dec^T ddz; cell backward; dgates -> all CTAs =================
    for (int idx = tid; idx < NB * A2; idx += kBT) {
      const int n = idx / A2, ap = idx - n * A2;
      float v0 = 0.f, v1 = 0.f;
      for (int qq = 0; qq < G; ++qq) {
        const float2 v = *reinterpret_cast<const float2*>(ddz_rx + (qq * NB + n) * A + 2 * ap);
        v0 += v.x; v1 += v.y;
      }
      ddzB[bfrag_word(2 * ap, n)] = pack_bf16x2(v0, v1);
      const int b = cl * NB + n;
      if (static_cast<int>(rank) == n && b < p.B)
        *reinterpret_cast<float2*>(p.ddz_all + (static_cast<int64_t>(b) * R + t + 1) * A + 2 * ap) = make_float2(v0, v1);
    }
    __syncthreads();
    if (d_act) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const uint2* hb = reinterpret_cast<const uint2*>(ddzB) + lane + d_kt0 * 32;
#pragma unroll
      for (int j = 0; j < kMaxFD2; ++j) {
        if (j < nfd) {
          const uint2 b = hb[j * 32];
          const uint32_t Af[4] = {Ad[j].x, Ad[j].y, Ad[j].z, Ad[j].w};
          mma_bf16_16816(acc, Af, b.x, b.y);
        }
      }
      reinterpret_cast<float4*>(red)[warp * 32 + lane] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
    __syncthreads();
    uint32_t w4[4] = {0u, 0u, 0u, 0u};
    {
      if (epi_ok && is_z) {
        const float dh = dz_acc + red_gather(red, c_w0, g.KSd, row_l & 15, n_e);
        const float2 if_ = __half22float2(*reinterpret_cast<const __half2*>(&gpk.x));
        const float2 go_ = __half22float2(*reinterpret_cast<const __half2*>(&gpk.y));
        const float i = if_.x, f = if_.y, gc = go_.x, o = go_.y;
        const float tc = tanh_acc(c_cur);
        const float dc = dh * o * (1.f - tc * tc) + dcell;
        dcell = dc * f;
        const float d4[4] = {dc * gc * i * (1.f - i), dc * c_prev * f * (1.f - f), dc * i * (1.f - gc * gc),
                             dh * tc * o * (1.f - o)};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          w4[k] = __bfloat16_as_ushort(__float2bfloat16(d4[k]));
        }
      }
      if (warp < (RPC * NB + 31) / 32) {
        // K order of the gathered gate gradients (must match the Wr^T fragments, pack_rowsel_kernel kperm):
        // k-tile = 4 hidden units; the 4 gate values of a unit pair are 4 consecutive fragment words
        uint32_t wd[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) wd[k] = w4[k] | (__shfl_down_sync(0xffffffffu, w4[k], 1) << 16);
        if (epi && is_z && (row_l & 1) == 0) {
          const int kt = u_e >> 2, pp = (u_e >> 1) & 1;
          const uint32_t off = dgB_base + 4u * static_cast<uint32_t>((kt * 32 + n_e * 4 + 2 * pp) * 2);
#pragma unroll
          for (int r = 0; r < kCS; ++r) st_remote_v4_u32(mapa(off, r), wd[0], wd[1], wd[2], wd[3]);
        }
      }
    }
    DTRACE(7);
    cluster_barrier();
    DTRACE(8);
    if (epi_ok && is_z) {
#pragma unroll
      for (int k = 0; k < 4; ++k) dg_ptr[k * Hd] = __ushort_as_bfloat16(static_cast<unsigned short>(w4[k]));
    }
    sv_idx -= Hd; dg_ptr -= 4 * Hd;
  }
#undef DTRACE
}

// A[row i][k] = W[k*ld + col(i)] for the rows a CTA owns in the backward kernel:
// i < UPC -> col = col_z0 + r*UPC + i; UPC <= i < UPC+OPC -> col = col_c0 + r*OPC + (i - UPC); else zero.
// out: [16][MT][KT][32][4] mma A fragments.
// kperm != 0 (Wr^T): fragment column k' of k-tile kt stands for W row gate*Hd + unit with unit = 4*kt + 2*pp + j and
// k' - 16*kt = 4*pp + j + (gate & 1 ? 8 : 0) + (gate & 2 ? 2 : 0)  (so that a unit pair's four gate gradients are four
// consecutive B-fragment words: one 16-byte DSMEM store in the backward kernel).
__global__ void pack_rowsel_kernel(const float* __restrict__ W, int64_t ld, int Ktot, int MT, int KT, int UPC, int OPC,
                                   int col_z0, int col_c0, int kperm, int Hd, uint32_t* __restrict__ out) {
  const int64_t total = static_cast<int64_t>(kCS) * MT * KT * 128;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int j = idx & 3, lane = (idx >> 2) & 31;
    int64_t tk = idx >> 7;
    const int kt = tk % KT; tk /= KT;
    const int mt = tk % MT;
    const int r = tk / MT;
    const int gq = lane >> 2, tig = lane & 3;
    const int i = 16 * mt + gq + 8 * (j & 1);
    const int k0 = 16 * kt + 2 * tig + 8 * (j >> 1);
    int col = -1;
    if (i < UPC) col = col_z0 + r * UPC + i;
    else if (i < UPC + OPC) col = col_c0 + r * OPC + (i - UPC);
    float v0 = 0.f, v1 = 0.f;
    if (col >= 0) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        int k = k0 + e;
        if (kperm) {
          const int kk = k & 15, ktile = k >> 4;
          const int gate = ((kk >> 3) & 1) | (((kk >> 1) & 1) << 1);
          const int unit = 4 * ktile + 2 * ((kk >> 2) & 1) + (kk & 1);
          k = (unit < Hd) ? gate * Hd + unit : Ktot;
        }
        const float v = (k < Ktot) ? W[static_cast<int64_t>(k) * ld + col] : 0.f;
        if (e == 0) v0 = v; else v1 = v;
      }
    }
    out[idx] = pack_bf16x2(v0, v1);
  }
}

bool g_dec_persist_checked = false;
int g_dec_persist_clusters = 0;   // co-resident 16-CTA clusters the device offers (0 = unavailable)

}  // namespace

// Utterances per cluster for a batch of B: the smallest power of two for which all clusters are
// co-resident (a second wave doubles the latency of the whole loop), capped by shared memory.
static int pick_nb(const las_dec_args* a, DGeom& g, BGeom& bg) {
  const int max_cl = g_dec_persist_clusters > 0 ? g_dec_persist_clusters : 7;
  int nb = 1;
  while (nb < 8 && (a->B + nb - 1) / nb > max_cl) nb *= 2;
  for (; nb >= 1; nb /= 2)
    if (dec_geom(a, nb, g) && dec_bgeom(a, nb, bg)) return nb;
  return 0;
}

int dec_persist_supported(const las_dec_args* a) {
  if (a->mode != 0 || a->Q == nullptr || a->wr2_pk == nullptr || a->mlp_dec_pk_p == nullptr || a->cbias == nullptr ||
      a->pbar == nullptr)
    return 0;
  if (a->drop_p > 0.f && a->seed_dev == nullptr) return 0;
  DGeom g;
  BGeom bg;
  if (pick_nb(a, g, bg) == 0) return 0;
  if (!g_dec_persist_checked) {
    // does the device schedule a 16-CTA (non-portable) cluster of this kernel at all?
    g_dec_persist_checked = true;
    g_dec_persist_clusters = 0;
    if (cudaFuncSetAttribute(dec_persist_fwd_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
        cudaFuncSetAttribute(dec_persist_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) == cudaSuccess) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(kCS, 1, 1);
      cfg.blockDim = dim3(kThreads);
      cfg.dynamicSmemBytes = 220 * 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = kCS;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, dec_persist_fwd_kernel, &cfg) == cudaSuccess) g_dec_persist_clusters = n;
    }
    (void)cudaGetLastError();
  }
  return g_dec_persist_clusters > 0 ? 1 : 0;
}

static void cluster_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* at, int nclusters, int threads, int smem,
                        cudaStream_t stream) {
  cfg = {};
  cfg.gridDim = dim3(kCS, nclusters, 1);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kCS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
}

int dec_persist_bwd(const las_dec_args* a, cudaStream_t stream) {
  DGeom g;
  BGeom bg;
  const int nb = pick_nb(a, g, bg);
  LAS_REQUIRE(nb > 0, "persistent decoder: unsupported geometry");
  LAS_REQUIRE(a->wrT2_pk && a->mlp_decT2_pk && a->de_all && a->dc_all && a->cpre && a->conv_save,
              "persistent decoder backward: missing buffers");
  static bool attr_set = false;
  if (!attr_set) {
    LAS_CUDA(cudaFuncSetAttribute(dec_persist_bwd_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    LAS_CUDA(cudaFuncSetAttribute(dec_persist_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set = true;
  }
  DecBwdP p;
  p.B = a->B; p.L = a->L; p.Te = a->Te; p.Hd = a->Hd; p.O = a->O; p.A = a->A; p.C = a->C; p.K = a->K;
  p.att_scaling = a->att_scaling;
  p.g = bg;
  p.P = a->P; p.Q = static_cast<const __nv_bfloat16*>(a->Q);
  p.wrT_pk = static_cast<const uint32_t*>(a->wrT2_pk); p.decT_pk = static_cast<const uint32_t*>(a->mlp_decT2_pk);
  p.conv_w = a->conv_w; p.mlp_att = a->mlp_att; p.gvec = a->gvec;
  p.ws = a->ws; p.gates_save = static_cast<const __half*>(a->gates_save); p.c_save = a->c_save; p.dzf = a->dzf;
  p.cpre = a->cpre; p.conv_save = a->conv_save; p.dzc_all = a->dzc_all;
  p.dgates = static_cast<__nv_bfloat16*>(a->dgates); p.dcz_all = static_cast<__nv_bfloat16*>(a->dcz_all);
  p.dc_all = a->dc_all; p.ddz_all = a->ddz_all; p.de_all = a->de_all; p.dattc_all = a->dattc_all;
  p.drop_p = a->drop_p; p.drop_site = a->drop_site; p.seed_dev = static_cast<const unsigned long long*>(a->seed_dev);
  p.dbg = static_cast<long long*>(g_dbg_buf_shared);
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute at[1];
  cluster_cfg(cfg, at, (a->B + nb - 1) / nb, kBT, bg.smem, stream);
  LAS_CUDA(cudaLaunchKernelEx(&cfg, dec_persist_bwd_kernel, p));
  ++g_launches;
  return 0;
}

// Row-sharded weight fragments of the backward kernel: which = 0 -> Wr^T (W = [W_hh | W_ih[:, E:]], f32
// [4Hd, Hd+O]), rows per CTA = its hidden units then its context dims; which = 1 -> mlp_dec^T
// (W = mlp_dec.weight, f32 [A, Hd]), rows per CTA = its hidden units.
int64_t dec_persist_pack_bytes(int which, int Hd, int O, int A) {
  const int UPC = Hd / kCS, OPC = O / kCS;
  if (which == 0) return static_cast<int64_t>(kCS) * ((UPC + OPC + 15) / 16) * (4 * Hd / 16) * 128 * 4;
  return static_cast<int64_t>(kCS) * ((UPC + 15) / 16) * ((A + 15) / 16) * 128 * 4;
}

int dec_persist_pack(int which, const float* W, int64_t ld, int Hd, int O, int A, void* out, cudaStream_t stream) {
  const int UPC = Hd / kCS, OPC = O / kCS;
  int MT, KT, Ktot, opc, cz, cc;
  if (which == 0) { MT = (UPC + OPC + 15) / 16; KT = 4 * Hd / 16; Ktot = 4 * Hd; opc = OPC; cz = 0; cc = Hd; }
  else            { MT = (UPC + 15) / 16; KT = (A + 15) / 16; Ktot = A; opc = 0; cz = 0; cc = 0; }
  const int64_t total = static_cast<int64_t>(kCS) * MT * KT * 128;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
  pack_rowsel_kernel<<<blocks, 256, 0, stream>>>(W, ld, Ktot, MT, KT, UPC, opc, cz, cc, which == 0 ? 1 : 0, Hd,
                                                 static_cast<uint32_t*>(out));
  ++g_launches;
  LAS_LAUNCH_CHECK();
  return 0;
}

int dec_persist_fwd(const las_dec_args* a, cudaStream_t stream) {
  DGeom g;
  BGeom bg;
  const int nb = pick_nb(a, g, bg);
  LAS_REQUIRE(nb > 0, "persistent decoder: unsupported geometry");
  static bool attr_set = false;
  if (!attr_set) {
    LAS_CUDA(cudaFuncSetAttribute(dec_persist_fwd_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    LAS_CUDA(cudaFuncSetAttribute(dec_persist_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set = true;
  }
  DecFwdP p;
  p.B = a->B; p.L = a->L; p.Te = a->Te; p.Hd = a->Hd; p.O = a->O; p.A = a->A; p.C = a->C; p.K = a->K;
  p.att_scaling = a->att_scaling;
  p.g = g;
  p.P = a->P; p.Q = static_cast<const __nv_bfloat16*>(a->Q); p.embx = a->embx;
  p.wr_pk = static_cast<const uint32_t*>(a->wr2_pk); p.dec_pk = static_cast<const uint32_t*>(a->mlp_dec_pk_p);
  p.cbias = a->cbias; p.pbar = a->pbar; p.conv_w = a->conv_w; p.mlp_att = a->mlp_att; p.gvec = a->gvec;
  p.ws = a->ws; p.zc = static_cast<__nv_bfloat16*>(a->zc); p.dzf = a->dzf;
  p.gates_save = static_cast<__half*>(a->gates_save); p.c_save = a->c_save;
  p.cpre = a->cpre; p.conv_save = a->conv_save;
  p.drop_p = a->drop_p; p.drop_site = a->drop_site; p.seed_dev = static_cast<const unsigned long long*>(a->seed_dev);
  p.dbg = static_cast<long long*>(g_dbg_buf_shared);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kCS, (a->B + nb - 1) / nb, 1);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = g.smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kCS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  LAS_CUDA(cudaLaunchKernelEx(&cfg, dec_persist_fwd_kernel, p));
  ++g_launches;
  return 0;
}

}  // namespace las
