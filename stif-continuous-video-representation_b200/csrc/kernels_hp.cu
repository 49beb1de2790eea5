// kernels_hp.cu -- the dense layers of the high-precision mode (STIF_MODE_FP32) on the tensor cores.
//
// STIF_MODE_FP32 evaluates the hoisted formulation layer by layer with fp32 activations in the workspace (kernels_fp32.cu).
// Its GEMMs -- C[M,N] = act(A[M,K] W[N,K]^T + b), M = a chunk of queries, K in {64, 256}, N in {64, 128, 256} -- used to run
// on a SIMT SGEMM.  Here they are tcgen05 MMAs on a 2-term bf16 split of BOTH operands,
//     a = a_hi + a_lo,  w = w_hi + w_lo   (hi = bf16(x), lo = bf16(x - hi)),     a.w ~= a_hi w_hi + a_lo w_hi + a_hi w_lo,
// three MMAs per K step into the same fp32 TMEM accumulator (the dropped a_lo w_lo term is 2^-16 relative).  SURVEY.md
// section 7.3-5 measured this split at 9e-6 max-abs RGB error with the stress weights, where single-pass TF32 gives 5.5e-4
// and misses the 1e-4 bound.  Activations keep their accurate sinf; tables stay fp32.
//
// One CTA = 128 rows of A: the fp32 rows are read once (coalesced), split into hi / lo SW128 tiles in shared memory, and
// reused for every 64-column chunk of N; the chunk's weight slices (hi and lo, pre-split and pre-swizzled on the host)
// arrive by bulk TMA while the previous chunk's epilogue (bias, sinf, coalesced fp32 stores) runs from the other TMEM slot.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "stif_internal.h"
#include "tc_pack.h"
#include "tc_primitives.cuh"

namespace stif {
namespace {

using namespace tc;

struct HpGemmParams {
  const float* A;          // [M, K] fp32, row stride K
  const uint8_t* w_hi;     // SW128 image of the whole layer: K/64 K-blocks x n_total rows x 128 B (bf16 hi parts)
  const uint8_t* w_lo;     // ... lo parts
  const float* bias;       // [N] (already offset to the first output of this call), may be null
  float* C;                // [M, ldc]
  long ldc, M;
  int n_total, n_off, N, K;
  int act;                 // 0 identity, 1 sine
  // version 2 only: the activations are not stored but contracted with an output layer proj_w [proj_n, 256] (+ proj_b) in the
  // epilogue (the 256 -> 4 flow and 256 -> 3 RGB layers: saves writing and re-reading 1 KB per query); needs N == 256
  // version 2 only: output columns >= n_split go to C2[m * ldc2 + (col - n_split)] (the composed 256 -> 64 | 128 layer writes F
  // and Q1|Q2 to different tables from ONE pass over its 256-wide input); n_split is a multiple of 32, 0 = no split
  float* C2 = nullptr;
  long ldc2 = 0;
  int n_split = 0;
  const float* proj_w = nullptr;
  const float* proj_b = nullptr;
  float* proj_out = nullptr;   // out[m * proj_scm + j * proj_scn]
  long proj_scm = 0, proj_scn = 0;
  int proj_n = 0;
};

extern __shared__ __align__(1024) uint8_t hp_smem[];

// STIF_HP_SINF=1 at build time keeps libdevice's sinf in the epilogues (the accuracy anchor of reduced_sin)
__device__ __forceinline__ float hp_sin(float x) {
#ifdef STIF_HP_SINF
  return sinf(x);
#else
  return reduced_sin(x);
#endif
}

__device__ __forceinline__ void wait_or_trap(uint64_t* bar, uint32_t parity) {
  for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it)
    if (it > (1u << 24)) __trap();
}
// the roles of the persistent GEMM wait for microseconds at a time: back off so that the spinning warps leave the issue slots to
// the epilogue (the spin loops were 18 % of the executed instructions)
__device__ __forceinline__ void wait_backoff_or_trap(uint64_t* bar, uint32_t parity) {
  for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it) {
    if (it > 4) __nanosleep(it < 64 ? 40 : 200);
    if (it > (1u << 22)) __trap();
  }
}

template <int K>
__global__ void __launch_bounds__(256, K == 64 ? 2 : 1) hp_gemm_kernel(const __grid_constant__ HpGemmParams p) {
  constexpr int KB = K / 64;                                     // 64-wide K blocks
  constexpr uint32_t kA = KB * 16384u, kW = KB * 8192u;          // bytes of one A tile (hi or lo) / one weight slice
  uint8_t* sA_hi = hp_smem;
  uint8_t* sA_lo = sA_hi + kA;
  uint8_t* sW_hi = sA_lo + kA;
  uint8_t* sW_lo = sW_hi + kW;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW_lo + kW);      // [0] weights landed, [1] weights read, [2,3] accumulator slot ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long m0 = (long)blockIdx.x * 128;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  // ---- A tile: fp32 rows -> (hi, lo) bf16 SW128 tiles; 4 consecutive k per thread and step (one float4)
  for (int idx = tid; idx < 128 * (K / 4); idx += 256) {
    const int row = idx / (K / 4), k = (idx - row * (K / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + row < p.M) v = __ldg(reinterpret_cast<const float4*>(p.A + (m0 + row) * K + k));
    const uint32_t h0 = pack_bf16x2(v.x, v.y), h1 = pack_bf16x2(v.z, v.w);
    const float r0 = v.x - __uint_as_float(h0 << 16), r1 = v.y - __uint_as_float(h0 & 0xFFFF0000u);
    const float r2 = v.z - __uint_as_float(h1 << 16), r3 = v.w - __uint_as_float(h1 & 0xFFFF0000u);
    const uint32_t off = (uint32_t)(k >> 6) * 16384u + sw128_offset(row, k & 63);
    *reinterpret_cast<uint2*>(sA_hi + off) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(sA_lo + off) = make_uint2(pack_bf16x2(r0, r1), pack_bf16x2(r2, r3));
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int nchunks = p.N / 64;
  const int quarter = warp & 3, colhalf = warp >> 2;
  auto epilogue = [&](int c) {
    wait_or_trap(&bars[2 + (c & 1)], (c >> 1) & 1);
    tc_fence_after();
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c & 1) * 64u + (uint32_t)colhalf * 32u, v);
    tmem_ld_wait();
    const long row = m0 + quarter * 32 + lane;
    const int col0 = c * 64 + colhalf * 32;
    float o[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float x = __uint_as_float(v[j]) + (p.bias ? __ldg(p.bias + col0 + j) : 0.f);
      o[j] = p.act ? hp_sin(x) : x;
    }
    if (row < p.M) {
      float4* dst = reinterpret_cast<float4*>(p.C + row * p.ldc + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    }
    tc_fence_before();
  };
  for (int c = 0; c < nchunks; ++c) {
    if (tid == 0) {
      if (c > 0) wait_or_trap(&bars[1], (c - 1) & 1);            // the previous chunk's MMAs have read the weight slices
      mbar_arrive_expect_tx(&bars[0], 2 * kW);
      const size_t kb_stride = (size_t)p.n_total * 128, row_off = (size_t)(p.n_off + c * 64) * 128;
      for (int kb = 0; kb < KB; ++kb) {
        bulk_copy_g2s(sW_hi + kb * 8192, p.w_hi + kb * kb_stride + row_off, 8192, &bars[0]);
        bulk_copy_g2s(sW_lo + kb * 8192, p.w_lo + kb * kb_stride + row_off, 8192, &bars[0]);
      }
      wait_or_trap(&bars[0], c & 1);
      tc_fence_after();
      const uint32_t idesc = make_idesc_bf16(128, 64), d = tmem + (uint32_t)(c & 1) * 64u;
      const uint32_t ah = smem_u32(sA_hi), al = smem_u32(sA_lo), wh = smem_u32(sW_hi), wl = smem_u32(sW_lo);
      bool first = true;
#pragma unroll
      for (int term = 0; term < 3; ++term) {                     // a_hi w_hi, a_lo w_hi, a_hi w_lo
        const uint32_t a = term == 1 ? al : ah, w = term == 2 ? wl : wh;
#pragma unroll
        for (int j = 0; j < K / 16; ++j) {
          umma_ss(d, make_desc_sw128(a + (j >> 2) * 16384) + 2 * (j & 3), make_desc_sw128(w + (j >> 2) * 8192) + 2 * (j & 3), idesc, !first);
          first = false;
        }
      }
      umma_commit(&bars[2 + (c & 1)]);
      umma_commit(&bars[1]);
    }
    __syncwarp();
    if (c > 0) epilogue(c - 1);                                  // overlaps this chunk's weight TMA + MMAs
    __syncthreads();                                             // slot (c - 1) & 1 is drained before chunk c + 1 is issued into it
  }
  epilogue(nchunks - 1);
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}


// ---- version 2: persistent, pipelined over K blocks and tiles ----------------------------------------------------------------
// Version 1 above is one 128-row tile per CTA with everything in sequence (A load -> split -> per chunk: weight TMA -> MMAs ->
// epilogue): 1.4 TB/s of HBM traffic where the layers need ~5.  Here a CTA is persistent and has three roles:
//   warps 8-15  producers (one group of four warps per stage of a 2-deep ring): the next 64-wide K block of A (fp32 -> hi / lo
//               SW128 tiles) and, by bulk TMA, the layer's weight rows for that K block (hi and lo, all N columns);
//   warp 16     issuer: 3 x 4 x N/64 MMAs per stage into the tile's N accumulator columns, tcgen05.commit frees the stage;
//   warps 0-7   epilogue of the PREVIOUS tile from the other half of TMEM (2 x 256 columns): bias, sine, fp32 rows.
// so loads, tensor work and the sine / store epilogue of neighbouring tiles overlap.  Same arithmetic as version 1 (same three
// terms, same order within an accumulator), so the two agree bit for bit.
constexpr uint32_t kHp2StageA = 2 * 16384;   // A hi | A lo of one K block
__host__ __device__ constexpr uint32_t hp2_stage_bytes(int N) { return kHp2StageA + 2u * (uint32_t)N * 128u; }
__host__ __device__ constexpr uint32_t hp2_smem_bytes(int N) { return 2u * hp2_stage_bytes(N) + 128u + 8u * 4096u + 1024u; }   // + one 32 x 32 fp32 transpose tile per epilogue warp + the layer's bias

__global__ void __launch_bounds__(544, 1) hp_gemm2_kernel(const __grid_constant__ HpGemmParams p) {
  const int KB = p.K / 64, NC = p.N / 64;
  const uint32_t stage_bytes = hp2_stage_bytes(p.N), w_bytes = (uint32_t)p.N * 128u;
  uint64_t* bars = reinterpret_cast<uint64_t*>(hp_smem + 2 * stage_bytes);   // [0,1] stage full, [2,3] stage empty, [4,5] accumulator full, [6,7] accumulator empty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars[s], 129);       // 128 producer threads + the expect_tx arrival of the weight TMA
      mbar_init(&bars[2 + s], 1);     // tcgen05.commit
      mbar_init(&bars[4 + s], 1);     // tcgen05.commit
      mbar_init(&bars[6 + s], 256);   // the epilogue threads
    }
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const long ntiles = (p.M + 127) / 128;
  if (warp < 8) {
    // ---------------- epilogue: thread = one row x half of the N columns
    const int quarter = warp & 3, ncol = p.N / 2, cbeg = (warp >> 2) * ncol;
    float4* xpose = reinterpret_cast<float4*>(hp_smem + 2 * stage_bytes + 128) + warp * 256;
    float4* pw = reinterpret_cast<float4*>(hp_smem + 2 * stage_bytes + 128);            // proj mode: proj_w [proj_n][256] ...
    float4* pex = pw + 256;                                                                // ... and the partial-sum exchange [2][4][32]
    // the bias in shared memory: a per-element __ldg sat on every accumulator's dependency chain (16 % of the stall samples)
    float* sbias = reinterpret_cast<float*>(hp_smem + 2 * stage_bytes + 128 + 8 * 4096);
    for (int i = tid; i < p.N; i += 256) sbias[i] = p.bias ? __ldg(p.bias + i) : 0.f;
    if (p.proj_n)
      for (int i = tid; i < p.proj_n * 64; i += 256) pw[i] = __ldg(reinterpret_cast<const float4*>(p.proj_w) + i);
    asm volatile("bar.sync 9, 256;" ::: "memory");   // the 8 epilogue warps
    long it = 0;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const uint32_t b = (uint32_t)(it & 1);
      wait_backoff_or_trap(&bars[4 + b], (uint32_t)(it >> 1) & 1);
      tc_fence_after();
      const long row0 = tile * 128 + quarter * 32;
      if (p.proj_n) {
        // thread = row x 128 of the 256 activations: partial dot products with the output layer, the two column halves (warps w
        // and w + 4) meet in shared memory; nothing of the 256-wide activation ever reaches HBM
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c0 = cbeg; c0 < cbeg + ncol; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + b * 256u + (uint32_t)c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float o[4];
            const float4 b4 = *reinterpret_cast<const float4*>(sbias + c0 + 4 * j);
            o[0] = hp_sin(__uint_as_float(v[4 * j]) + b4.x); o[1] = hp_sin(__uint_as_float(v[4 * j + 1]) + b4.y);
            o[2] = hp_sin(__uint_as_float(v[4 * j + 2]) + b4.z); o[3] = hp_sin(__uint_as_float(v[4 * j + 3]) + b4.w);
#pragma unroll
            for (int n = 0; n < 4; ++n)
              if (n < p.proj_n) {
                const float4 w4 = pw[n * 64 + (c0 >> 2) + j];
                acc[n] = fmaf(o[3], w4.w, fmaf(o[2], w4.z, fmaf(o[1], w4.y, fmaf(o[0], w4.x, acc[n]))));
              }
          }
        }
        tc_fence_before();
        const int slot = ((int)b * 4 + quarter) * 32 + lane, barid = 1 + (int)b * 4 + quarter;
        if (warp >= 4) {
          pex[slot] = make_float4(acc[0], acc[1], acc[2], acc[3]);
          __threadfence_block();
          asm volatile("bar.arrive %0, 64;" ::"r"(barid) : "memory");
        } else {
          asm volatile("bar.sync %0, 64;" ::"r"(barid) : "memory");
          const float4 o = pex[slot];
          const long grow = row0 + lane;
          if (grow < p.M) {
            const float r[4] = {acc[0] + o.x, acc[1] + o.y, acc[2] + o.z, acc[3] + o.w};
            if (p.proj_n == 4 && p.proj_scn == 1 && p.proj_scm == 4) {
              *reinterpret_cast<float4*>(p.proj_out + grow * 4) =
                  make_float4(r[0] + __ldg(p.proj_b), r[1] + __ldg(p.proj_b + 1), r[2] + __ldg(p.proj_b + 2), r[3] + __ldg(p.proj_b + 3));
            } else {
              for (int n = 0; n < p.proj_n; ++n) p.proj_out[grow * p.proj_scm + n * p.proj_scn] = r[n] + __ldg(p.proj_b + n);
            }
          }
        }
        mbar_arrive(&bars[6 + b]);
        continue;
      }
      for (int c0 = cbeg; c0 < cbeg + ncol; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + b * 256u + (uint32_t)c0, v);
        tmem_ld_wait();
        // thread = row holds 32 consecutive outputs: transposed through the warp's shared tile (16-byte slots XOR-swizzled by the row,
        // 4 wavefronts per access both ways) so that every store instruction of the warp writes 4 rows x 128 contiguous bytes
        // instead of 32 rows x 16 bytes (the row-per-thread stores of version 1 cost one L2 request per 16 bytes: 8.6 us per tile).
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = *reinterpret_cast<const float4*>(sbias + c0 + 4 * j);
          float4 o = make_float4(__uint_as_float(v[4 * j]) + b4.x, __uint_as_float(v[4 * j + 1]) + b4.y,
                                 __uint_as_float(v[4 * j + 2]) + b4.z, __uint_as_float(v[4 * j + 3]) + b4.w);
          if (p.act) o = make_float4(hp_sin(o.x), hp_sin(o.y), hp_sin(o.z), hp_sin(o.w));
          xpose[lane * 8 + (j ^ (lane & 7))] = o;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 4 + (lane >> 3), slot = lane & 7;
          const float4 o = xpose[r * 8 + (slot ^ (r & 7))];
          const long grow = row0 + r;
          if (grow < p.M) {
            if (p.n_split && c0 >= p.n_split) *reinterpret_cast<float4*>(p.C2 + grow * p.ldc2 + (c0 - p.n_split) + slot * 4) = o;
            else *reinterpret_cast<float4*>(p.C + grow * p.ldc + c0 + slot * 4) = o;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&bars[6 + b]);
    }
  } else if (warp < 16) {
    // ---------------- producers: group 0 (warps 8-11) fills stage 0 with the even items, group 1 (warps 12-15) stage 1 with the
    // odd ones, so that two stages' worth of loads (64 KB per SM) are in flight
    const int pt = (tid - 256) & 127;
    const uint32_t group = (uint32_t)(tid - 256) >> 7;
    long item = 0;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
      for (int kb = 0; kb < KB; ++kb, ++item) {
        const uint32_t s = (uint32_t)(item & 1);
        if (s != group) continue;
        wait_backoff_or_trap(&bars[2 + s], ((uint32_t)(item >> 1) & 1) ^ 1);   // (passes at once the first time round)
        uint8_t* st = hp_smem + s * stage_bytes;
        if (pt == 0) {
          mbar_arrive_expect_tx(&bars[s], 2 * w_bytes);
          const size_t off = ((size_t)kb * p.n_total + p.n_off) * 128;
          bulk_copy_g2s(st + kHp2StageA, p.w_hi + off, w_bytes, &bars[s]);
          bulk_copy_g2s(st + kHp2StageA + w_bytes, p.w_lo + off, w_bytes, &bars[s]);
        }
        float4 v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int idx = pt + 128 * i, row = idx >> 4, k = (idx & 15) * 4;
          const long m = tile * 128 + row;
          v[i] = m < p.M ? __ldg(reinterpret_cast<const float4*>(p.A + m * p.K + kb * 64 + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int idx = pt + 128 * i, row = idx >> 4, k = (idx & 15) * 4;
          const uint32_t h0 = pack_bf16x2(v[i].x, v[i].y), h1 = pack_bf16x2(v[i].z, v[i].w);
          const float r0 = v[i].x - __uint_as_float(h0 << 16), r1 = v[i].y - __uint_as_float(h0 & 0xFFFF0000u);
          const float r2 = v[i].z - __uint_as_float(h1 << 16), r3 = v[i].w - __uint_as_float(h1 & 0xFFFF0000u);
          const uint32_t off = sw128_offset(row, k);
          *reinterpret_cast<uint2*>(st + off) = make_uint2(h0, h1);
          *reinterpret_cast<uint2*>(st + 16384 + off) = make_uint2(pack_bf16x2(r0, r1), pack_bf16x2(r2, r3));
        }
        fence_proxy_async_smem();
        mbar_arrive(&bars[s]);
      }
  } else if (lane == 0) {
    // ---------------- issuer
    const uint32_t idesc = make_idesc_bf16(128, 64);
    long item = 0, it = 0;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const uint32_t b = (uint32_t)(it & 1);
      wait_or_trap(&bars[6 + b], ((uint32_t)(it >> 1) & 1) ^ 1);       // the epilogue has drained this half of TMEM
      tc_fence_after();
      for (int kb = 0; kb < KB; ++kb, ++item) {
        const uint32_t s = (uint32_t)(item & 1);
        wait_or_trap(&bars[s], (uint32_t)(item >> 1) & 1);
        tc_fence_after();
        const uint32_t ah = smem_u32(hp_smem + s * stage_bytes), al = ah + 16384, wh = ah + kHp2StageA, wl = wh + w_bytes;
        for (int c = 0; c < NC; ++c) {
          const uint32_t d = tmem + b * 256u + (uint32_t)c * 64u;
#pragma unroll
          for (int term = 0; term < 3; ++term) {                       // a_hi w_hi, a_lo w_hi, a_hi w_lo
            const uint32_t a = term == 1 ? al : ah, w = (term == 2 ? wl : wh) + (uint32_t)c * 8192u;
#pragma unroll
            for (int j = 0; j < 4; ++j) umma_ss(d, make_desc_sw128(a) + 2 * j, make_desc_sw128(w) + 2 * j, idesc, kb > 0 || term > 0 || j > 0);
          }
        }
        umma_commit(&bars[2 + s]);
      }
      umma_commit(&bars[4 + b]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int K>
constexpr size_t hp_smem_bytes() { return 2 * (size_t)(K / 64) * 16384 + 2 * (size_t)(K / 64) * 8192 + 64; }

void split_images(const std::vector<float>& w, int N, int K, std::vector<uint8_t>& hi, std::vector<uint8_t>& lo) {
  std::vector<float> h((size_t)N * K), l((size_t)N * K);
  for (size_t i = 0; i < h.size(); ++i) {
    h[i] = bf16_round_host(w[i]);
    l[i] = w[i] - h[i];
  }
  append_sw128_image(hi, h.data(), N, K);
  append_sw128_image(lo, l.data(), N, K);
}

}  // namespace

struct HpWeights {
  HpLayer layer[HP_NUM_LAYERS];
};

const HpLayer* hp_layer(const HpWeights* w, int id) { return w ? &w->layer[id] : nullptr; }

HpWeights* hp_weights_create(const FoldedWeights& hw, std::string& err) {
  auto* t = new HpWeights();
  // the projection's [256, 198] matrices padded to K = 256 with zero columns (the packed input rows carry zeros there too)
  std::vector<float> k0[2] = {std::vector<float>((size_t)256 * 256, 0.f), std::vector<float>((size_t)256 * 256, 0.f)};
  for (int r = 0; r < 256; ++r)
    for (int c = 0; c < 198; ++c) {
      k0[0][(size_t)r * 256 + c] = hw.w_tab[(size_t)r * 198 + c];
      k0[1][(size_t)r * 256 + c] = hw.w_tab_lat[(size_t)r * 198 + c];
    }
  const std::vector<float>* src[HP_NUM_LAYERS] = {&hw.f1_w, &hw.f2_w, &hw.f3_w, &hw.l1_w, &hw.l2_w, &hw.e1_w, &hw.e2_w, &hw.e3_w, &k0[0], &k0[1]};
  const int N[HP_NUM_LAYERS] = {64, 256, 192, 64, 256, 64, 256, 256, 256, 256}, K[HP_NUM_LAYERS] = {64, 64, 256, 64, 64, 64, 64, 256, 256, 256};
  for (int i = 0; i < HP_NUM_LAYERS; ++i) t->layer[i] = HpLayer{nullptr, nullptr, N[i], K[i]};
  for (int i = 0; i < HP_NUM_LAYERS; ++i) {
    std::vector<uint8_t> hi, lo;
    split_images(*src[i], N[i], K[i], hi, lo);
    cudaError_t e = cudaMalloc(&t->layer[i].hi, hi.size());
    if (e == cudaSuccess) e = cudaMalloc(&t->layer[i].lo, lo.size());
    if (e == cudaSuccess) e = cudaMemcpy(t->layer[i].hi, hi.data(), hi.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(t->layer[i].lo, lo.data(), lo.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      err = cudaGetErrorString(e);
      hp_weights_destroy(t);
      return nullptr;
    }
  }
  cudaError_t e = cudaFuncSetAttribute(hp_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hp_smem_bytes<64>());
  if (e == cudaSuccess) e = cudaFuncSetAttribute(hp_gemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hp_smem_bytes<256>());
  if (e == cudaSuccess) e = cudaFuncSetAttribute(hp_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hp2_smem_bytes(256));
  if (e != cudaSuccess) {
    err = cudaGetErrorString(e);
    hp_weights_destroy(t);
    return nullptr;
  }
  return t;
}

void hp_weights_destroy(HpWeights* t) {
  if (!t) return;
  for (auto& l : t->layer) {
    if (l.hi) cudaFree(l.hi);
    if (l.lo) cudaFree(l.lo);
  }
  delete t;
}

cudaError_t hp_gemm(const LaunchCtx& cx, const HpLayer& L, int n_off, int N, const float* A, const float* bias, float* C, long ldc,
                    long M, int act) {
  if (M <= 0) return cudaSuccess;
  if (N % 64 != 0 || n_off % 8 != 0 || n_off + N > L.N || (L.K != 64 && L.K != 256) || ldc % 4 != 0) return cudaErrorInvalidValue;
  HpGemmParams p{A, L.hi, L.lo, bias, C, ldc, M, L.N, n_off, N, L.K, act};
  const unsigned grid = (unsigned)((M + 127) / 128);
  static const bool v1 = getenv("STIF_HP_V1") && atoi(getenv("STIF_HP_V1")) != 0;   // A/B switch: the one-tile-per-CTA kernel
  if (!v1 && N <= 256) {
    hp_gemm2_kernel<<<std::min<unsigned>(grid, (unsigned)cx.num_sms), 544, hp2_smem_bytes(N), cx.stream>>>(p);
    ++*cx.launch_counter;
    return cudaGetLastError();
  }
  if (L.K == 64) hp_gemm_kernel<64><<<grid, 256, hp_smem_bytes<64>(), cx.stream>>>(p);
  else hp_gemm_kernel<256><<<grid, 256, hp_smem_bytes<256>(), cx.stream>>>(p);
  ++*cx.launch_counter;
  return cudaGetLastError();
}

// One pass over A, two destinations: columns [0, n_split) of the layer -> C, [n_split, L.N) -> C2 (version 2 kernel)
cudaError_t hp_gemm_split(const LaunchCtx& cx, const HpLayer& L, const float* A, const float* bias, long M, int act, int n_split, float* C,
                          long ldc, float* C2, long ldc2) {
  if (M <= 0) return cudaSuccess;
  if (L.N % 64 != 0 || L.N > 256 || n_split % 32 != 0 || n_split <= 0 || n_split >= L.N || ldc % 4 != 0 || ldc2 % 4 != 0) return cudaErrorInvalidValue;
  HpGemmParams p{A, L.hi, L.lo, bias, C, ldc, M, L.N, 0, L.N, L.K, act};
  p.C2 = C2; p.ldc2 = ldc2; p.n_split = n_split;
  const unsigned grid = (unsigned)((M + 127) / 128);
  hp_gemm2_kernel<<<std::min<unsigned>(grid, (unsigned)cx.num_sms), 544, hp2_smem_bytes(L.N), cx.stream>>>(p);
  ++*cx.launch_counter;
  return cudaGetLastError();
}

// act(A W^T + b) contracted with the NOUT x 256 output layer in the epilogue (version 2 kernel): out[m * scm + j * scn]
cudaError_t hp_gemm_proj(const LaunchCtx& cx, const HpLayer& L, const float* A, const float* bias, long M, const float* proj_w,
                         const float* proj_b, int proj_n, float* out, long scm, long scn) {
  if (M <= 0) return cudaSuccess;
  if (L.N != 256 || (L.K != 64 && L.K != 256) || proj_n < 1 || proj_n > 4 || !bias) return cudaErrorInvalidValue;
  HpGemmParams p{A, L.hi, L.lo, bias, nullptr, 0, M, L.N, 0, 256, L.K, 1};
  p.proj_w = proj_w; p.proj_b = proj_b; p.proj_out = out; p.proj_scm = scm; p.proj_scn = scn; p.proj_n = proj_n;
  const unsigned grid = (unsigned)((M + 127) / 128);
  hp_gemm2_kernel<<<std::min<unsigned>(grid, (unsigned)cx.num_sms), 544, hp2_smem_bytes(256), cx.stream>>>(p);
  ++*cx.launch_counter;
  return cudaGetLastError();
}

// On-device check of the split GEMM against a double-precision host product (stif_selftest).
int hp_selftest(std::string& report) {
  const int M = 300, Ks[2] = {64, 256}, N = 128;
  int fails = 0;
  for (int t = 0; t < 2; ++t) {
    const int K = Ks[t];
    std::vector<float> A((size_t)M * K), W((size_t)192 * K), b(N);
    uint32_t s = 12345u + t;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 32768.0f - 1.0f; };
    for (auto& x : A) x = rnd();
    for (auto& x : W) x = 3.0f * rnd();
    for (auto& x : b) x = rnd();
    std::vector<uint8_t> hi, lo;
    split_images(W, 192, K, hi, lo);
    HpLayer L{nullptr, nullptr, 192, K};
    float *dA = nullptr, *dB = nullptr, *dC = nullptr;
    cudaMalloc(&L.hi, hi.size()); cudaMalloc(&L.lo, lo.size());
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, b.size() * 4); cudaMalloc(&dC, (size_t)M * N * 4);
    cudaMemcpy(L.hi, hi.data(), hi.size(), cudaMemcpyHostToDevice); cudaMemcpy(L.lo, lo.data(), lo.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(hp_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hp_smem_bytes<64>());
    cudaFuncSetAttribute(hp_gemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hp_smem_bytes<256>());
    cudaFuncSetAttribute(hp_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hp2_smem_bytes(256));
    int64_t launches = 0;
    LaunchCtx cx{nullptr, &launches, 148};
    cudaError_t e = hp_gemm(cx, L, 64, N, dA, dB, dC, N, M, 0);       // outputs 64..191 of the layer
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    std::vector<float> C((size_t)M * N);
    cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0.0, ref_max = 0.0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double r = b[n];
        for (int k = 0; k < K; ++k) r += (double)A[(size_t)m * K + k] * W[(size_t)(64 + n) * K + k];
        worst = std::max(worst, std::fabs(r - C[(size_t)m * N + n]));
        ref_max = std::max(ref_max, std::fabs(r));
      }
    char line[256];
    snprintf(line, sizeof line, "T%d split-bf16 GEMM k%d n%d: max_abs_err %.3e (ref max %.3e) %s\n", 4 + t, K, N, worst, ref_max,
             e == cudaSuccess ? "" : cudaGetErrorString(e));
    report += line;
    if (e != cudaSuccess || !(worst <= 2e-4 * ref_max)) ++fails;
    cudaFree(L.hi); cudaFree(L.lo); cudaFree(dA); cudaFree(dB); cudaFree(dC);
  }
  return fails;
}

}  // namespace stif
