// Design-time microbenchmark (not part of the product): cost of one all-gather step inside a 16-CTA cluster
// as a function of HOW the bytes are pushed through distributed shared memory:
//   mode 0  st.async.b32  (4-byte messages + complete_tx)
//   mode 1  st.async.v2   (8-byte)
//   mode 2  st.async.v4   (16-byte)
//   mode 3  cp.async.bulk shared::cta -> shared::cluster (one bulk copy per peer, complete_tx on the peer's mbarrier)
//   mode 4  st.shared::cluster.v4 + barrier.cluster (the round-1 scheme)
// Every CTA owns `bytes` per step, delivers them to all peers, then waits until everybody's bytes have arrived.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dsmem_bench tools/dsmem_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!try_wait(bar, parity)) { if (++spins > (1u << 22)) __trap(); }
}
__device__ __forceinline__ void expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

constexpr int kCS = 16;

// threads_send: number of threads that take part in the sends
__global__ void __launch_bounds__(512, 1) xchg_kernel(int steps, int bytes, int mode, int threads_send, long long* cycles,
                                                      uint32_t* sink) {
  extern __shared__ __align__(128) uint8_t smem[];    // [2][kCS][bytes] receive buffers + [bytes] staging
  __shared__ __align__(8) uint64_t bars[2];
  const uint32_t rank = cluster_rank();
  const int tid = threadIdx.x;
  uint8_t* stage = smem + 2 * kCS * bytes;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t tx = (uint32_t)bytes * kCS;
  if (tid == 0 && mode != 4) { expect_tx(smem_u32(&bars[0]), tx); expect_tx(smem_u32(&bars[1]), tx); }
  cluster_barrier();
  uint32_t acc = 0;
  long long t0 = clock64();
  long long t_send = 0;
  for (int t = 0; t < steps; ++t) {
    const int b = t & 1;
    const uint32_t par = (t >> 1) & 1;
    const uint32_t dst_local = smem_u32(smem + (b * kCS + rank) * bytes);
    const uint32_t bar_local = smem_u32(&bars[b]);
    long long s0 = clock64();
    if (mode == 0) {
      for (int i = tid; i < (bytes / 4) * kCS; i += threads_send) {
        if (tid >= threads_send) break;
        const int peer = i % kCS, w = i / kCS;
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                     ::"r"(mapa(dst_local + 4 * w, peer)), "r"(acc + w), "r"(mapa(bar_local, peer)) : "memory");
      }
    } else if (mode == 1) {
      for (int i = tid; i < (bytes / 8) * kCS; i += threads_send) {
        if (tid >= threads_send) break;
        const int peer = i % kCS, w = i / kCS;
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];"
                     ::"r"(mapa(dst_local + 8 * w, peer)), "r"(acc + w), "r"(t), "r"(mapa(bar_local, peer)) : "memory");
      }
    } else if (mode == 2) {
      for (int i = tid; i < (bytes / 16) * kCS; i += threads_send) {
        if (tid >= threads_send) break;
        const int peer = i % kCS, w = i / kCS;
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                     ::"r"(mapa(dst_local + 16 * w, peer)), "r"(acc + w), "r"(t), "r"(rank), "r"(peer), "r"(mapa(bar_local, peer)) : "memory");
      }
    } else if (mode == 3) {
      // stage locally (generic proxy), make it visible to the async proxy, one bulk copy per peer
      for (int i = tid; i < bytes / 4; i += threads_send) {
        if (tid >= threads_send) break;
        reinterpret_cast<uint32_t*>(stage)[i] = acc + i;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid < kCS) {
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(mapa(dst_local, tid)), "r"(smem_u32(stage)), "r"(bytes), "r"(mapa(bar_local, tid)) : "memory");
      }
    } else {
      for (int i = tid; i < (bytes / 16) * kCS; i += threads_send) {
        if (tid >= threads_send) break;
        const int peer = i % kCS, w = i / kCS;
        asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(mapa(dst_local + 16 * w, peer)), "r"(acc + w), "r"(t), "r"(rank), "r"(peer) : "memory");
      }
    }
    long long s1 = clock64();
    t_send += s1 - s0;
    if (mode == 4) {
      cluster_barrier();
    } else {
      wait(bar_local, par);
      if (mode == 3) __syncthreads();   // the staging buffer is rewritten next step: everybody has passed the wait
      if (tid == 0 && t + 2 < steps) expect_tx(bar_local, tx);
    }
    for (int peer = 0; peer < kCS; ++peer) acc += reinterpret_cast<uint32_t*>(smem + (b * kCS + peer) * bytes)[tid % (bytes / 4)];
  }
  long long t1 = clock64();
  cluster_barrier();
  if (tid == 0 && blockIdx.x == 0) { cycles[0] = t1 - t0; cycles[1] = t_send; }
  sink[blockIdx.x * blockDim.x + tid] = acc;
}

int main() {
  long long* cyc; uint32_t* sink;
  CK(cudaMalloc(&cyc, 16)); CK(cudaMalloc(&sink, 1 << 22));
  CK(cudaFuncSetAttribute(xchg_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CK(cudaFuncSetAttribute(xchg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const char* names[5] = {"st.async.b32", "st.async.v2", "st.async.v4", "cp.async.bulk", "st.cluster.v4+barrier"};
  const int steps = 2000;
  for (int bytes : {64, 256, 320, 640, 1280}) {
    for (int threads_send : {32, 160, 512}) {
      for (int mode = 0; mode < 5; ++mode) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(kCS * 4, 1, 1);
        cfg.blockDim = dim3(512);
        cfg.dynamicSmemBytes = (2 * kCS + 1) * bytes;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = kCS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, xchg_kernel, steps, bytes, mode, threads_send, cyc, sink));
        CK(cudaDeviceSynchronize());
        long long h[2];
        CK(cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost));
        printf("bytes/cta=%4d senders=%3d %-22s: %6.0f cyc/step (send issue %5.0f)\n", bytes, threads_send, names[mode],
               (double)h[0] / steps, (double)h[1] / steps);
      }
    }
  }
  return 0;
}
