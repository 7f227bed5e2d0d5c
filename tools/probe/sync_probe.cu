// Latencies of the synchronisation primitives the tcgen05 pipelines are built from (cycles, one CTA):
//   1. mbarrier.arrive (thread A) -> try_wait returns in thread B (another warp)
//   2. tcgen05.commit (no MMA pending) -> try_wait returns in another warp
//   3. ping-pong round trip producer-warp <-> consumer-thread through two mbarriers (the skeleton of one pipeline stage)
// nvcc -gencode arch=compute_100a,code=sm_100a -I cim_quantization_b200/csrc -o sync_probe tools/probe/sync_probe.cu
#include <cstdio>
#include "tc_ptx.cuh"
using namespace cimq::ptx;

__global__ void __launch_bounds__(256) probe(long long *out, int iters, int mode) {
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full = smem_u32(&bars[0]), empty = smem_u32(&bars[1]);
  if (threadIdx.x == 0) {
    mbar_init(full, mode == 2 ? 128 : 1);
    mbar_init(empty, 1);
    fence_barrier_init();
  }
  if (warp == 7) tmem_alloc(smem_u32(&tmem_slot), 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  long long t0 = clock64();
  if (mode == 0) {
    // ping-pong: warp 0 lane 0 arrives on `full`, warp 4 lane 0 waits then arrives on `empty`, warp 0 waits
    if (warp == 0 && lane == 0) {
      for (int i = 0; i < iters; ++i) { mbar_arrive(full); mbar_wait(empty, i & 1); }
    } else if (warp == 4 && lane == 0) {
      for (int i = 0; i < iters; ++i) { mbar_wait(full, i & 1); mbar_arrive(empty); }
    }
  } else if (mode == 1) {
    // the consumer answers with tcgen05.commit instead of a plain arrive
    if (warp == 0 && lane == 0) {
      for (int i = 0; i < iters; ++i) { mbar_arrive(full); mbar_wait(empty, i & 1); }
    } else if (warp == 4 && lane == 0) {
      for (int i = 0; i < iters; ++i) { mbar_wait(full, i & 1); tc_fence_after(); umma_commit(empty); }
    }
  } else if (mode == 2) {
    // 128 producer threads (fence + arrive each), consumer commits
    if (warp < 4) {
      for (int i = 0; i < iters; ++i) { fence_proxy_async(); mbar_arrive(full); mbar_wait(empty, i & 1); }
    } else if (warp == 4 && lane == 0) {
      for (int i = 0; i < iters; ++i) { mbar_wait(full, i & 1); tc_fence_after(); umma_commit(empty); }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[mode] = (t1 - t0) / iters;
  __syncthreads();
  if (warp == 7) tmem_dealloc(tmem_slot, 32);
}

int main() {
  long long *d, h[4] = {0, 0, 0, 0};
  cudaMalloc(&d, sizeof(h));
  for (int mode = 0; mode < 3; ++mode) {
    probe<<<1, 256>>>(d, 2000, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
  }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("round trip, cycles: arrive<->arrive %lld, arrive<->tcgen05.commit %lld, 128 x (fence+arrive)<->commit %lld\n", h[0], h[1], h[2]);
  return 0;
}
