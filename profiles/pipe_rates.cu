// pipe_rates.cu -- issue-rate probe for the pipes the sine epilogues live on (sm_100a): MUFU.SIN, FMUL, FFMA, FFMA2 (fp32x2),
// F2FP (cvt.bf16x2) and mixes of them, per SM sub-partition, as a function of resident warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/pipe_rates profiles/pipe_rates.cu && ./profiles/pipe_rates
// Every thread runs ILP independent chains, unrolled, so the numbers are throughput, not latency.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float fast_sin(float x) { float r; asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2(float x) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ unsigned pack(float a, float b) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); return r; }

constexpr int ILP = 8, ITERS = 512;

// MODE 0: sin.approx (FMUL + MUFU.SIN)   1: ex2.approx (MUFU.EX2 only)   2: fma.f32   3: fma.f32x2   4: cvt.bf16x2
//      5: per iteration ILP x (sin) + ILP x NF (fma.f32x2)  -- NF = FFMA2 per sine, the mix of the sine epilogues
template <int MODE, int NF>
__global__ void probe(float* out, long long* clk, float seed) {
  float x[ILP];
  unsigned long long y[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { x[i] = seed + i * 0.37f + threadIdx.x * 1e-3f; y[i] = (unsigned long long)__float_as_uint(x[i]) << 32 | __float_as_uint(x[i]); }
  unsigned acc = 0;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0) x[i] = fast_sin(x[i]);
      if (MODE == 1) x[i] = ex2(x[i]);
      if (MODE == 2) x[i] = ffma(x[i], 1.0001f, 0.5f);
      if (MODE == 3) y[i] = fma2(y[i], y[(i + 1) % ILP], y[i]);
      if (MODE == 4) acc += pack(x[i], x[(i + 1) % ILP]);
      if (MODE == 5) {
        x[i] = fast_sin(x[i]);
#pragma unroll
        for (int k = 0; k < NF; ++k) y[i] = fma2(y[i], y[(i + 1) % ILP], y[i]);
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i] + __uint_as_float((unsigned)(y[i] >> 32)) + __uint_as_float((unsigned)y[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE, int NF>
void run(const char* name, int warps) {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&clk, 148 * sizeof(long long));
  probe<MODE, NF><<<148, warps * 32>>>(out, clk, 0.1f);
  probe<MODE, NF><<<148, warps * 32>>>(out, clk, 0.1f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  const double per_smsp_warps = warps / 4.0;
  const double winst = (double)ITERS * ILP * per_smsp_warps;          // "primary" warp-instructions per SMSP
  printf("%-28s warps/SM %2d: %8.0f clk  -> %.2f clk per primary warp-instr per SMSP (%.2f lanes/clk/SMSP)\n", name, warps, c, c / winst, 32.0 * winst / c);
  cudaFree(out); cudaFree(clk);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<0, 0>("sin.approx (FMUL+MUFU.SIN)", w);
    run<1, 0>("ex2.approx (MUFU.EX2)", w);
    run<2, 0>("fma.f32", w);
    run<3, 0>("fma.f32x2", w);
    run<4, 0>("cvt.bf16x2", w);
    run<5, 1>("sin + 1 fma.f32x2", w);
    run<5, 2>("sin + 2 fma.f32x2", w);
    run<5, 3>("sin + 3 fma.f32x2", w);
    run<5, 4>("sin + 4 fma.f32x2", w);
    run<5, 6>("sin + 6 fma.f32x2", w);
  }
  return 0;
}
