// Design-time microbenchmarks for the latency-bound recurrences (not part of the product):
//   1. mma.sync m16n8k16 bf16 dependent-chain latency and per-warp issue rate
//   2. which cluster sizes the device will schedule
//   3. per-timestep cost of an all-gather across a thread-block cluster through distributed
//      shared memory (st.shared::cluster + remote mbarrier arrive) vs barrier.cluster
//   4. back-to-back tiny-kernel launch cadence on one stream
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// CHAINS independent accumulators, ITER rounds
template <int CHAINS>
__global__ void hmma_kernel(float* out, long long* cycles, int iters) {
  uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u};
  uint32_t b0 = 0x3c003c00u + threadIdx.x, b1 = 0x3c003c00u;
  float d[CHAINS][4];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) { d[c][0] = d[c][1] = d[c][2] = d[c][3] = 0.f; }
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) mma16816(d[c], a, b0, b1);
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

__global__ void empty_kernel(int* p) { if (p && threadIdx.x == 1000) *p = 1; }

// ---------------- cluster all-gather ----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}"
               : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  return ok != 0;
}

// Every CTA owns `words` u32 per step and must deliver them to all cluster peers; then waits for
// everybody's words. Double-buffered by step parity. mode 0: DSMEM stores + remote mbarrier arrive
// (one arrive per warp per peer); mode 1: DSMEM stores + barrier.cluster.
__global__ void gather_kernel(int steps, int words, int mode, long long* cycles, uint32_t* sink) {
  extern __shared__ __align__(16) uint32_t buf[];   // [2][nc][words]
  __shared__ __align__(8) uint64_t bars[2];
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t nc = cluster.num_blocks();
  const uint32_t rank = cluster.block_rank();
  const int nwarps = blockDim.x / 32;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(nc * nwarps));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster.sync();
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int t = 0; t < steps; ++t) {
    const int b = t & 1;
    const uint32_t par = (t >> 1) & 1;
    // produce: each thread writes its share of `words` to every peer
    for (uint32_t peer = 0; peer < nc; ++peer) {
      uint32_t dst_base = mapa(smem_u32(buf + (b * nc + rank) * words), peer);
      for (int w = threadIdx.x * 4; w < words; w += blockDim.x * 4) {
        uint4 v = make_uint4(acc + w, t, rank, peer);
        st_cluster_v4(dst_base + w * 4, v);
      }
    }
    if (mode == 0) {
      __syncwarp();
      if (lane < nc) mbar_arrive_remote(mapa(smem_u32(&bars[b]), lane));
      uint32_t spins = 0;
      while (!mbar_try_wait_cluster(smem_u32(&bars[b]), par)) { if (++spins > (1u << 24)) { __trap(); } }
    } else {
      asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    // consume: read one word from each peer's slice
    for (uint32_t peer = 0; peer < nc; ++peer) acc += buf[(b * nc + peer) * words + (threadIdx.x * 4) % words];
  }
  long long t1 = clock64();
  cluster.sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s  SMs=%d  clock=%d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
  float* out; long long* cyc; uint32_t* sink;
  CK(cudaMalloc(&out, 1 << 20)); CK(cudaMalloc(&cyc, 8)); CK(cudaMalloc(&sink, 1 << 22));
  long long h;
  // 1. HMMA
  const int iters = 4096;
  for (int warps = 1; warps <= 8; warps *= 2) {
    hmma_kernel<1><<<1, 32 * warps>>>(out, cyc, iters); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("hmma chains=1 warps=%d: %.2f cyc/mma (dependent latency)\n", warps, (double)h / iters);
    hmma_kernel<4><<<1, 32 * warps>>>(out, cyc, iters); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("hmma chains=4 warps=%d: %.2f cyc/mma per warp\n", warps, (double)h / iters / 4);
    hmma_kernel<8><<<1, 32 * warps>>>(out, cyc, iters); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("hmma chains=8 warps=%d: %.2f cyc/mma per warp\n", warps, (double)h / iters / 8);
  }
  // 2. cluster sizes
  CK(cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CK(cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  int sizes[] = {2, 4, 5, 8, 10, 16};
  for (int cs : sizes) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 8); cfg.blockDim = dim3(160); cfg.dynamicSmemBytes = 32 * 1024;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, gather_kernel, &cfg);
    printf("cluster size %2d: maxActiveClusters=%d (%s)\n", cs, nclusters, cudaGetErrorString(e));
    cudaGetLastError();
  }
  // 3. all-gather step cost
  for (int cs : sizes) {
    for (int mode = 0; mode < 2; ++mode) {
      for (int words : {128, 512}) {
        for (int nclusters : {1, 8}) {
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(cs * nclusters); cfg.blockDim = dim3(160);
          cfg.dynamicSmemBytes = 2 * cs * words * 4;
          cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
          cfg.attrs = at; cfg.numAttrs = 1;
          const int steps = 2000;
          cudaError_t e = cudaLaunchKernelEx(&cfg, gather_kernel, steps, words, mode, cyc, sink);
          if (e != cudaSuccess) { printf("gather cs=%d launch failed: %s\n", cs, cudaGetErrorString(e)); cudaGetLastError(); continue; }
          e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("gather cs=%d run failed: %s\n", cs, cudaGetErrorString(e)); return 1; }
          CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
          printf("gather cs=%2d mode=%s bytes/cta=%4d clusters=%d: %.0f cyc/step\n", cs, mode == 0 ? "mbarrier" : "barrier.cluster", words * 4, nclusters, (double)h / steps);
        }
      }
    }
  }
  // 4. launch cadence
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 2000; ++i) empty_kernel<<<148, 128>>>(nullptr);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("launch cadence: %.2f us per tiny kernel (2000 back-to-back)\n", ms * 1000 / 2000);
  }
  // graph cadence
  {
    cudaStream_t s; CK(cudaStreamCreate(&s));
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal));
    for (int i = 0; i < 2000; ++i) empty_kernel<<<148, 128, 0, s>>>(nullptr);
    CK(cudaStreamEndCapture(s, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0, s));
      CK(cudaGraphLaunch(ge, s));
      CK(cudaEventRecord(e1, s)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("graph cadence: %.2f us per tiny kernel (2000-node graph)\n", ms * 1000 / 2000);
    }
  }
  printf("done\n");
  return 0;
}
