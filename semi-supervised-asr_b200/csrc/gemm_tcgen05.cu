// bf16 x bf16 -> f32 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged by TMA into 128B-swizzled shared memory, persistent over output tiles with a
// double-buffered TMEM accumulator so the epilogue of tile i overlaps the main loop of tile i+1.
//
// This is the dense-contraction workhorse of the LAS step (SURVEY.md §2.1 K1-i, K4, K6 and the
// wgrad/dgrad GEMMs behind them):
//   D[m,n] = sum_k A[m,k] * B[n,k]   (+ bias[n]) (relu) (+= C)
// Each operand can be K-major (row-major [rows, K]) or MN-major (row-major [K, rows]); the
// MN-major form is what lets the weight-gradient GEMMs (dW = dY^T X) read dY and X in place
// instead of through transposed copies.
#include "common.cuh"
#include "las_internal.h"
#include <cuda.h>
#include <cudaTypedefs.h>

namespace las {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 bytes = one swizzle atom row
constexpr int UMMA_K = 16;
constexpr int kStages = 6;
constexpr int kAccStages = 2;
constexpr int kThreads = 384;  // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warps4-11 epilogue (two per TMEM lane quadrant)
constexpr int kEpiWarps = 8;

template <int BN>
struct SmemLayout {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BN * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiOffset = kStages * kStageBytes;          // per epilogue warp: a (32 rows x 32 f32) transpose tile
  static constexpr int kBarOffset = kEpiOffset + kEpiWarps * 4096;
  static constexpr int kTotal = kBarOffset + 256 + 1024;  // barriers + alignment slack
};

struct GemmParams {
  void* C;
  const float* bias;
  int64_t ldc;
  int M, N, K;
  int relu;
  int accumulate;
  int ksplit;        // > 1: work item = (tile, K slice); C is then the f32 workspace [ksplit][M][N] (ldc = N)
};

template <int BN, bool A_MN, bool B_MN, typename OutT>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            GemmParams p) {
  using L = SmemLayout<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + kAccStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_blocks = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int n_blocks = (p.N + BN - 1) / BN;
  const int k_blocks = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int num_tiles = m_blocks * n_blocks;
  const int kb_per = (k_blocks + p.ksplit - 1) / p.ksplit;
  const int num_items = num_tiles * p.ksplit;      // host guarantees that every K slice is non-empty

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiWarps);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kAccStages * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int tile = item % num_tiles, split = item / num_tiles;
      const int m0 = (tile / n_blocks) * BLOCK_M;
      const int n0 = (tile % n_blocks) * BN;
      const int kb0 = split * kb_per, kb1 = min(k_blocks, kb0 + kb_per);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * L::kStageBytes;
        uint8_t* sb = sa + L::kABytes;
        mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
        const int k0 = kb * BLOCK_K;
        if constexpr (!A_MN) {
          tma_load_2d(sa, &tmap_a, &full_bar[stage], k0, m0);           // box {64 k, 128 rows}
        } else {
#pragma unroll
          for (int j = 0; j < BLOCK_M / 64; ++j)                          // boxes {64 mn, 64 k}
            tma_load_2d(sa + j * (BLOCK_K * 128), &tmap_a, &full_bar[stage], m0 + j * 64, k0);
        }
        if constexpr (!B_MN) {
          tma_load_2d(sb, &tmap_b, &full_bar[stage], k0, n0);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_2d(sb + j * (BLOCK_K * 128), &tmap_b, &full_bar[stage], n0 + j * 64, k0);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer (single thread) =====================
    constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int split = item / num_tiles;
      const int kb0 = split * kb_per, kb1 = min(k_blocks, kb0 + kb_per);
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
        const uint32_t sb = sa + L::kABytes;
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          // K-major: advance 16 elements = 32 bytes inside the swizzle row.
          // MN-major: advance 16 k-rows of 128 bytes; atoms along MN are one 64-row box apart.
          const uint64_t adesc = A_MN ? make_smem_desc_sw128(sa + k * (UMMA_K * 128), BLOCK_K * 128, 1024)
                                      : make_smem_desc_sw128(sa + k * (UMMA_K * 2), 16, 1024);
          const uint64_t bdesc = B_MN ? make_smem_desc_sw128(sb + k * (UMMA_K * 128), BLOCK_K * 128, 1024)
                                      : make_smem_desc_sw128(sb + k * (UMMA_K * 2), 16, 1024);
          umma_bf16(d_tmem, adesc, bdesc, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
        if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int q = warp & 3;  // TMEM lane quadrant this warp may touch
    const int half = (warp - 4) >> 2;   // column half of the tile served by this warp
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int tile = item % num_tiles, split = item / num_tiles;
      OutT* Cout = reinterpret_cast<OutT*>(p.C) + static_cast<int64_t>(split) * p.M * p.N;   // split > 0 only with ldc == N
      const int m0 = (tile / n_blocks) * BLOCK_M;
      const int n0 = (tile % n_blocks) * BN;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      // Each lane reads one accumulator ROW from TMEM; the 32x32 chunk is transposed through a swizzled
      // shared-memory tile so that global stores are row-contiguous (a quarter warp writes one full 128-byte
      // line of an f32 row) instead of 32 scattered 16-byte pieces per instruction.
      float4* xt = reinterpret_cast<float4*>(smem + L::kEpiOffset + (warp - 4) * 4096);
      const int row_base = m0 + q * 32;
#pragma unroll 1
      for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c0, v);
        tmem_ld_wait();
        const int col0 = n0 + c0;
        if (col0 >= p.N || row_base >= p.M) continue;      // warp-uniform
        const int ncols = min(32, p.N - col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 o;
          o.x = __uint_as_float(v[4 * j]); o.y = __uint_as_float(v[4 * j + 1]);
          o.z = __uint_as_float(v[4 * j + 2]); o.w = __uint_as_float(v[4 * j + 3]);
          xt[lane * 8 + (j ^ (lane & 7))] = o;
        }
        __syncwarp();
        // lane -> (row sub-index rs = lane >> 3, float4 column jj = lane & 7); 8 passes of 4 rows
        const int jj = lane & 7, rs = lane >> 3;
        const int cbase = col0 + 4 * jj;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias != nullptr) {
          if (cbase + 0 < p.N) bv.x = __ldg(p.bias + cbase + 0);
          if (cbase + 1 < p.N) bv.y = __ldg(p.bias + cbase + 1);
          if (cbase + 2 < p.N) bv.z = __ldg(p.bias + cbase + 2);
          if (cbase + 3 < p.N) bv.w = __ldg(p.bias + cbase + 3);
        }
        const bool vec_ok = (ncols == 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = 4 * i + rs;
          const int row = row_base + rl;
          float4 o = xt[rl * 8 + (jj ^ (rl & 7))];
          o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
          if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          if (row >= p.M) continue;
          OutT* cptr = Cout + static_cast<int64_t>(row) * p.ldc + cbase;
          if constexpr (sizeof(OutT) == 4) {
            if (vec_ok && (reinterpret_cast<uintptr_t>(cptr) & 15) == 0) {
              float4* c4 = reinterpret_cast<float4*>(cptr);
              if (p.accumulate) {
                const float4 old = *c4;
                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
              }
              *c4 = o;
            } else {
              const float f[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (cbase + e < p.N) cptr[e] = p.accumulate ? cptr[e] + f[e] : f[e];
            }
          } else {
            if (vec_ok && !p.accumulate && (reinterpret_cast<uintptr_t>(cptr) & 7) == 0) {
              uint2 pk;
              pk.x = pack_bf16x2(o.x, o.y);
              pk.y = pack_bf16x2(o.z, o.w);
              *reinterpret_cast<uint2*>(cptr) = pk;
            } else {
              const float f[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (cbase + e < p.N) {
                  const float x = p.accumulate ? f[e] + __bfloat162float(cptr[e]) : f[e];
                  cptr[e] = __float2bfloat16(x);
                }
            }
          }
        }
        __syncwarp();      // the tile is rewritten by the next chunk
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kAccStages * BN);
  }
}

// Deterministic split-K reduction: C[m, n] = (accumulate ? C : 0) + sum_s ws[s][m][n] (+ bias[n]) (relu).
template <typename OutT>
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int ksplit, int M, int N, OutT* __restrict__ C,
                                     int64_t ldc, const float* __restrict__ bias, int relu, int accumulate) {
  const int64_t total = static_cast<int64_t>(M) * N;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(i / N), n = static_cast<int>(i % N);
    float v = 0.f;
    for (int s = 0; s < ksplit; ++s) v += ws[s * total + i];
    if (bias != nullptr) v += bias[n];
    if (relu) v = fmaxf(v, 0.f);
    OutT* c = C + m * ldc + n;
    if constexpr (sizeof(OutT) == 4) *c = accumulate ? *c + v : v;
    else *c = __float2bfloat16(accumulate ? __bfloat162float(*c) + v : v);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

int get_encode() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  LAS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  LAS_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess,
              "cuTensorMapEncodeTiled not available from the driver");
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return 0;
}

// 2-D bf16 tensor map over a row-major [rows, cols] matrix with leading dimension ld (elements);
// box = {box_cols (inner), box_rows}; 128B swizzle, zero fill out of bounds.
int make_tmap(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld,
              int box_cols, int box_rows) {
  LAS_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "gemm operand base not 16B aligned");
  LAS_REQUIRE((ld * 2) % 16 == 0, "gemm operand leading dimension %lld not a multiple of 8 elements",
              static_cast<long long>(ld));
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LAS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): rows=%lld cols=%lld ld=%lld",
              static_cast<int>(r), static_cast<long long>(rows), static_cast<long long>(cols),
              static_cast<long long>(ld));
  return 0;
}

template <int BN, bool A_MN, bool B_MN, typename OutT>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  using L = SmemLayout<BN>;
  auto kern = gemm_kernel<BN, A_MN, B_MN, OutT>;
  static bool configured = false;
  if (!configured) {
    LAS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured = true;
  }
  const int tiles = ((p.M + BLOCK_M - 1) / BLOCK_M) * ((p.N + BN - 1) / BN) * p.ksplit;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, kThreads, L::kTotal, stream>>>(ta, tb, p); ++g_launches;
  LAS_LAUNCH_CHECK();
  return 0;
}

template <int BN, typename OutT>
int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb,
                   const GemmParams& p, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch<BN, false, false, OutT>(ta, tb, p, s);
  if (a_mn && b_mn) return launch<BN, true, true, OutT>(ta, tb, p, s);
  if (a_mn && !b_mn) return launch<BN, true, false, OutT>(ta, tb, p, s);
  return launch<BN, false, true, OutT>(ta, tb, p, s);
}

}  // namespace

int gemm_bf16(const void* A, int64_t lda, bool a_mn, const void* B, int64_t ldb, bool b_mn, void* C,
              int64_t ldc, bool c_bf16, const float* bias, int M, int N, int K, bool relu,
              bool accumulate, cudaStream_t stream, void* ws, int64_t ws_bytes) {
  if (M == 0 || N == 0) return 0;
  LAS_REQUIRE(K > 0, "gemm: K must be positive");
  int rc = get_encode();
  if (rc) return rc;
  // Narrow outputs (vocabulary logits, pyramid projections) use the 64-wide tile.
  const int BN = (N <= 64) ? 64 : 128;
  CUtensorMap ta, tb;
  if (!a_mn) rc = make_tmap(&ta, A, M, K, lda, BLOCK_K, BLOCK_M);
  else       rc = make_tmap(&ta, A, K, M, lda, 64, BLOCK_K);
  if (rc) return rc;
  if (!b_mn) rc = make_tmap(&tb, B, N, K, ldb, BLOCK_K, BN);
  else       rc = make_tmap(&tb, B, K, N, ldb, 64, BLOCK_K);
  if (rc) return rc;
  // Split-K for the weight-gradient shapes (few output tiles, K = all frames of the batch): the K slices
  // go to otherwise idle SMs, partial tiles land in the caller's workspace and are summed in a fixed order.
  const int tiles = ((M + BLOCK_M - 1) / BLOCK_M) * ((N + BN - 1) / BN);
  const int k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
  int ksplit = 1;
  if (ws != nullptr && tiles * 2 <= num_sms() && k_blocks >= 32) {
    ksplit = num_sms() / tiles;
    if (ksplit > k_blocks / 8) ksplit = k_blocks / 8;
    if (ksplit > 16) ksplit = 16;
    while (ksplit > 1 && static_cast<int64_t>(ksplit) * M * N * 4 > ws_bytes) --ksplit;
    const int per = (k_blocks + ksplit - 1) / ksplit;
    ksplit = (k_blocks + per - 1) / per;            // no empty slice
  }
  if (ksplit > 1) {
    GemmParams p{ws, nullptr, N, M, N, K, 0, 0, ksplit};
    rc = (BN == 64) ? dispatch_major<64, float>(a_mn, b_mn, ta, tb, p, stream)
                    : dispatch_major<128, float>(a_mn, b_mn, ta, tb, p, stream);
    if (rc) return rc;
    const int64_t total = static_cast<int64_t>(M) * N;
    int blocks = static_cast<int>((total + 255) / 256);
    if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
    if (c_bf16)
      splitk_reduce_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(static_cast<const float*>(ws), ksplit, M, N,
                                                                      static_cast<__nv_bfloat16*>(C), ldc, bias, relu, accumulate);
    else
      splitk_reduce_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const float*>(ws), ksplit, M, N,
                                                              static_cast<float*>(C), ldc, bias, relu, accumulate);
    ++g_launches;
    LAS_LAUNCH_CHECK();
    return 0;
  }
  GemmParams p{C, bias, ldc, M, N, K, relu ? 1 : 0, accumulate ? 1 : 0, 1};
  if (BN == 64) {
    return c_bf16 ? dispatch_major<64, __nv_bfloat16>(a_mn, b_mn, ta, tb, p, stream)
                  : dispatch_major<64, float>(a_mn, b_mn, ta, tb, p, stream);
  }
  return c_bf16 ? dispatch_major<128, __nv_bfloat16>(a_mn, b_mn, ta, tb, p, stream)
                : dispatch_major<128, float>(a_mn, b_mn, ta, tb, p, stream);
}

}  // namespace las
