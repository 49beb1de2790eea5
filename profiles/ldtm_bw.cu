// TMEM read bandwidth probe: W warps of one CTA each read their lane quarter's 32 fp32 columns (tcgen05.ld 32x32b.x32,
// 4 KB per warp instruction) R times; prints clocks per round and bytes/clk/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I stif-continuous-video-representation_b200/csrc profiles/ldtm_bw.cu -o /tmp/ldtm_bw && /tmp/ldtm_bw
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_primitives.cuh"
using namespace stif::tc;
__global__ void probe(int rounds, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 32u;
  uint32_t v[32], acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    tmem_ld32(base, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) acc ^= v[j];
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  sink[threadIdx.x] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}
int main() {
  long long* d; uint32_t* s; cudaMalloc(&d, 8); cudaMalloc(&s, 4096);
  for (int warps : {1, 4, 8, 16}) {
    const int rounds = 2000;
    probe<<<1, warps * 32>>>(rounds, d, s); cudaDeviceSynchronize();
    probe<<<1, warps * 32>>>(rounds, d, s);
    long long clk = 0; cudaMemcpy(&clk, d, 8, cudaMemcpyDeviceToHost);
    printf("%2d warps: %.1f clk per round (4 KB per warp) -> %.1f B/clk/SM  [%s]\n", warps, (double)clk / rounds,
           (double)warps * 4096 * rounds / clk, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
