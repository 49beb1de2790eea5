// kernels_tc.cu -- the fused bf16 tensor-core path of the STIF query decoder (STIF_MODE_BF16).
//
// Two persistent kernels per (t, b) slab, one CTA per SM, 256 threads = two independent
// "workgroups" (WG, 4 warps each) that each own a 128-query tile and half of tensor memory:
//
//   K1  stage A + B  (reference Sakuya_arch_test.py:382-422): nearest gather of the projected latent
//       -> feat_imnet trunk -> composed last layer writes the projected HR table (Q1|Q2, fp16) ->
//       + bilinear gather of TB -> flow_imnet trunk -> flow (fp32).
//   K2  stage C + D + E (warplayer.py:25-39, :424-458): warp positions from the flow, four bilinear
//       gathers (Q1@g1, Q2@g2, TE1@g1, TE2@g2) -> encode_imnet trunk -> RGB (fp32 planar).
//
// Per tile every MLP layer is a chain of tcgen05.mma (M=128 queries, N=64 output chunk, K=16 per
// instruction, bf16 x bf16 -> fp32 in TMEM).  Weights stay resident in shared memory for the
// whole kernel (bulk-TMA loaded once); activations never leave the SM: the epilogue warps read
// an accumulator chunk (tcgen05.ld), apply bias + sine (MUFU), round to bf16 and write it back
// to TMEM as the next layer's A operand (tcgen05.st, "TS" MMA form).  The 256->4 (flow) and
// 256->3 (RGB) output layers run on the FMA pipe in fp32 straight from the sine outputs.
//
// TMEM map of one WG (256 columns): A-area [0,128) = up to 256 bf16 activations per query;
// D-area [128,256) = two 64-column fp32 accumulator slots used as a ring (MMA of chunk c+1
// overlaps the sine epilogue of chunk c; the other WG's tile fills the remaining bubbles).
#include <algorithm>
#include <cstdio>
#include <string>
#include <vector>

#include "stif_internal.h"
#include "tc_pack.h"
#include "tc_primitives.cuh"

namespace stif {
namespace {

using namespace tc;

constexpr int kTile = 128;
constexpr uint32_t kColA = 0;     // A-area: 256-wide activations, channel k at column k/2
constexpr uint32_t kColAin = 96;  // 64-wide activations live in the last 32 columns of the A-area
constexpr uint32_t kColD = 128;   // two accumulator slots: [128,192), [192,256)

// ---- shared-memory images (bytes) ----------------------------------------------------------
constexpr uint32_t kW64x64 = 64 * 64 * 2, kW256x64 = 256 * 64 * 2, kW192x256 = 192 * 256 * 2, kW256x256 = 256 * 256 * 2;
// K1: F1 | F2 | F3(composed 192x256) | L1 | L2
constexpr uint32_t k1F1 = 0, k1F2 = k1F1 + kW64x64, k1F3 = k1F2 + kW256x64, k1L1 = k1F3 + kW192x256,
                   k1L2 = k1L1 + kW64x64, k1WBytes = k1L2 + kW256x64;
constexpr uint32_t k1Bars = k1WBytes, k1Smem = k1Bars + 128 + 1024;
// K2: E1 | E2 | E3 | A0[2] | tap staging[8 warps x 2 KB]
constexpr uint32_t k2E1 = 0, k2E2 = k2E1 + kW64x64, k2E3 = k2E2 + kW256x64, k2WBytes = k2E3 + kW256x256;
constexpr uint32_t k2A0 = k2WBytes, k2Taps = k2A0 + 2 * 16384, k2Bars = k2Taps + 8 * 2048, k2Smem = k2Bars + 128 + 1024;
static_assert(k2A0 % 1024 == 0 && k1F3 % 1024 == 0 && k1L1 % 1024 == 0 && k2E3 % 1024 == 0, "SW128 tiles need 1024 B alignment");
static_assert(k2Smem <= 232448 && k1Smem <= 232448, "exceeds 227 KB of shared memory");

struct K1Consts {
  float cA[64];       // feat_imnet L0: w_t * t + b        (per launch)
  float a_rel[128];   // feat_imnet L0: (rel_y, rel_x) columns, [c][2]
  float f1_b[64], f2_b[256], f3_b[192];
  float cB[64];       // flow_imnet L0: w_t * t + b        (per launch)
  float l1_b[64], l2_b[256];
  float l3_w[4 * 256], l3_b[4];
};
struct K2Consts {
  float cE[64];       // encode_imnet L0: w_t * t + b      (per launch)
  float e1_b[64], e2_b[256], e3_b[256];
  float e4_w[3 * 256], e4_b[4];
};
struct K1Params {
  K1Consts c;
  Geometry g;
  const __half* tab;   // [H*W,256]   TA | TB | TE1 | TE2
  __half* qtab;        // [HH*WW,128] Q1 | Q2
  float* flow;         // [HH*WW,4]
  const uint8_t* wimg;
  long q_begin, q_end;
};
struct K2Params {
  K2Consts c;
  Geometry g;
  const __half* tab;
  const __half* qtab;
  const float* flow;
  float* out;          // [3, plane]
  long plane;
  const uint8_t* wimg;
  long q_begin, q_end;
  int band_lo, band_hi, band_mode;
  int* flag;
};

// ---- workgroup context -----------------------------------------------------------------------
struct WgCtx {
  uint32_t tmem;       // TMEM address of this WG's column 0, lane 0
  uint32_t lane_addr;  // same + this warp's lane quarter (for tcgen05.ld/st)
  uint64_t* full;      // two mbarriers: accumulator slot s is complete
  uint32_t n_issued, n_waited;
  int wg, tid_wg;
};

__device__ __forceinline__ void wg_barrier(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }

__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, uint32_t parity) {
  long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    if (mbar_try_wait(bar, parity)) return;
    if ((it & 0xFFF) == 0xFFF) {
      long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) __trap();   // ~2 s: a lost arrival must not hang the GPU
    }
  }
}

// Issue the MMAs of one 64-wide output chunk (thread 0 of the WG only) and commit to the slot's barrier.
//   a_smem != 0: A operand is an SW128 tile in shared memory (K = 64); else A is in TMEM at column a_col.
__device__ __forceinline__ void issue_chunk(WgCtx& cx, uint32_t a_smem, uint32_t a_col, uint32_t w_smem, int n_rows,
                                            int chunk, int ksteps) {
  const uint32_t slot = cx.n_issued & 1;
  if (cx.tid_wg == 0) {
    const uint32_t idesc = make_idesc_bf16(128, 64);
    const uint32_t d = cx.tmem + kColD + slot * 64;
    for (int j = 0; j < ksteps; ++j) {
      const uint64_t bdesc = make_desc_sw128(w_smem + (uint32_t)(j >> 2) * (uint32_t)n_rows * 128u + (uint32_t)chunk * 8192u) + 2 * (j & 3);
      if (a_smem) umma_ss(d, make_desc_sw128(a_smem) + 2 * j, bdesc, idesc, j > 0);
      else umma_ts(d, cx.tmem + a_col + 8 * j, bdesc, idesc, j > 0);
    }
    umma_commit(&cx.full[slot]);
  }
  ++cx.n_issued;
}

// Wait for the oldest outstanding chunk; returns the lane-adjusted TMEM address of its accumulator slot.
__device__ __forceinline__ uint32_t wait_chunk(WgCtx& cx) {
  const uint32_t slot = cx.n_waited & 1, parity = (cx.n_waited >> 1) & 1;
  mbar_wait_or_trap(&cx.full[slot], parity);
  ++cx.n_waited;
  tc_fence_after();
  return cx.lane_addr + kColD + slot * 64;
}

// One MLP layer: NC output chunks of 64.  Preconditions: the A operand is complete and a WG barrier
// has been passed since it was written and since both accumulator slots were last read.
template <int NC, class ChunkOf, class Epi>
__device__ __forceinline__ void run_layer(WgCtx& cx, uint32_t a_smem, uint32_t a_col, uint32_t w_smem, int n_rows, int ksteps,
                                          ChunkOf chunk_of, Epi epi) {
  if (cx.tid_wg == 0) tc_fence_after();
  issue_chunk(cx, a_smem, a_col, w_smem, n_rows, chunk_of(0), ksteps);
  if (NC > 1) issue_chunk(cx, a_smem, a_col, w_smem, n_rows, chunk_of(1), ksteps);
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const uint32_t d = wait_chunk(cx);
    epi(i, d);
    tc_fence_before();
    wg_barrier(cx.wg);
    if (i + 2 < NC) {
      if (cx.tid_wg == 0) tc_fence_after();
      issue_chunk(cx, a_smem, a_col, w_smem, n_rows, chunk_of(i + 2), ksteps);
    }
  }
}

// ---- epilogues (thread = query row; 64 accumulator columns per chunk) ---------------------------
// act = sin(acc + bias) -> bf16 -> TMEM A operand at column dst (32 columns)
__device__ __forceinline__ void epi_sin_to_tmem(uint32_t src, uint32_t dst, const float* __restrict__ bias) {
  uint32_t v0[32], v1[32], pk[16];
  tmem_ld32(src, v0);
  tmem_ld32(src + 32, v1);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 16; ++j)
    pk[j] = pack_bf16x2(fast_sin(__uint_as_float(v0[2 * j]) + bias[2 * j]), fast_sin(__uint_as_float(v0[2 * j + 1]) + bias[2 * j + 1]));
  tmem_st16(dst, pk);
#pragma unroll
  for (int j = 0; j < 16; ++j)
    pk[j] = pack_bf16x2(fast_sin(__uint_as_float(v1[2 * j]) + bias[32 + 2 * j]),
                        fast_sin(__uint_as_float(v1[2 * j + 1]) + bias[32 + 2 * j + 1]));
  tmem_st16(dst + 16, pk);
  tmem_st_wait();
}

// act = sin(acc + bias) kept in fp32 and contracted with the NOUT x 256 output layer on the FMA pipe
template <int NOUT>
__device__ __forceinline__ void epi_sin_fma(uint32_t src, const float* __restrict__ bias, const float* __restrict__ w,
                                            float (&acc)[NOUT]) {
  uint32_t v0[32], v1[32];
  tmem_ld32(src, v0);
  tmem_ld32(src + 32, v1);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float s = fast_sin(__uint_as_float(v0[j]) + bias[j]);
#pragma unroll
    for (int k = 0; k < NOUT; ++k) acc[k] = fmaf(w[k * 256 + j], s, acc[k]);
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float s = fast_sin(__uint_as_float(v1[j]) + bias[32 + j]);
#pragma unroll
    for (int k = 0; k < NOUT; ++k) acc[k] = fmaf(w[k * 256 + 32 + j], s, acc[k]);
  }
}

__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_half2(uint32_t v) {
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}

// acc + bias -> fp16 -> 128 bytes of the projected HR table
__device__ __forceinline__ void epi_store_qtab(uint32_t src, const float* __restrict__ bias, __half* dst, bool valid) {
  uint32_t v0[32], v1[32];
  tmem_ld32(src, v0);
  tmem_ld32(src + 32, v1);
  tmem_ld_wait();
  if (!valid) return;
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 o;
    o.x = pack_half2(__uint_as_float(v0[8 * j + 0]) + bias[8 * j + 0], __uint_as_float(v0[8 * j + 1]) + bias[8 * j + 1]);
    o.y = pack_half2(__uint_as_float(v0[8 * j + 2]) + bias[8 * j + 2], __uint_as_float(v0[8 * j + 3]) + bias[8 * j + 3]);
    o.z = pack_half2(__uint_as_float(v0[8 * j + 4]) + bias[8 * j + 4], __uint_as_float(v0[8 * j + 5]) + bias[8 * j + 5]);
    o.w = pack_half2(__uint_as_float(v0[8 * j + 6]) + bias[8 * j + 6], __uint_as_float(v0[8 * j + 7]) + bias[8 * j + 7]);
    d4[j] = o;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 o;
    o.x = pack_half2(__uint_as_float(v1[8 * j + 0]) + bias[32 + 8 * j + 0], __uint_as_float(v1[8 * j + 1]) + bias[32 + 8 * j + 1]);
    o.y = pack_half2(__uint_as_float(v1[8 * j + 2]) + bias[32 + 8 * j + 2], __uint_as_float(v1[8 * j + 3]) + bias[32 + 8 * j + 3]);
    o.z = pack_half2(__uint_as_float(v1[8 * j + 4]) + bias[32 + 8 * j + 4], __uint_as_float(v1[8 * j + 5]) + bias[32 + 8 * j + 5]);
    o.w = pack_half2(__uint_as_float(v1[8 * j + 6]) + bias[32 + 8 * j + 6], __uint_as_float(v1[8 * j + 7]) + bias[32 + 8 * j + 7]);
    d4[4 + j] = o;
  }
}

// f0 = sin(F + g) -> bf16 -> TMEM (g already holds bilinear(TB) + time constant + composed bias)
__device__ __forceinline__ void epi_flow_first_layer(uint32_t src, uint32_t dst, const float (&g)[64]) {
  uint32_t v0[32], v1[32], pk[16];
  tmem_ld32(src, v0);
  tmem_ld32(src + 32, v1);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 16; ++j)
    pk[j] = pack_bf16x2(fast_sin(__uint_as_float(v0[2 * j]) + g[2 * j]), fast_sin(__uint_as_float(v0[2 * j + 1]) + g[2 * j + 1]));
  tmem_st16(dst, pk);
#pragma unroll
  for (int j = 0; j < 16; ++j)
    pk[j] = pack_bf16x2(fast_sin(__uint_as_float(v1[2 * j]) + g[32 + 2 * j]), fast_sin(__uint_as_float(v1[2 * j + 1]) + g[32 + 2 * j + 1]));
  tmem_st16(dst + 16, pk);
  tmem_st_wait();
}

// ---- common prologue / epilogue of both kernels ---------------------------------------------------
struct CtaSetup {
  uint8_t* smem;
  uint64_t* bars;  // [0] weights landed, [1,2] WG0 slots, [3,4] WG1 slots
  uint32_t tmem_base;
};

__device__ __forceinline__ CtaSetup cta_prologue(uint8_t* smem_raw, uint32_t bars_off, const uint8_t* wimg, uint32_t wbytes) {
  CtaSetup s;
  s.smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  s.bars = reinterpret_cast<uint64_t*>(s.smem + bars_off);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s.bars + 8);
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&s.bars[0], 1);
    for (int i = 1; i <= 4; ++i) mbar_init(&s.bars[i], 1);
    fence_mbar_init();
  }
  if (tid < 32) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    mbar_arrive_expect_tx(&s.bars[0], wbytes);
    for (uint32_t off = 0; off < wbytes; off += 32768) {
      const uint32_t n = min(32768u, wbytes - off);
      bulk_copy_g2s(s.smem + off, wimg + off, n, &s.bars[0]);
    }
  }
  s.tmem_base = *tmem_slot;
  mbar_wait_or_trap(&s.bars[0], 0);
  return s;
}

__device__ __forceinline__ void cta_epilogue(uint32_t tmem_base) {
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem_base, 512);
}

__device__ __forceinline__ WgCtx make_wg(const CtaSetup& s) {
  WgCtx cx;
  const int tid = threadIdx.x;
  cx.wg = tid >> 7;
  cx.tid_wg = tid & 127;
  cx.tmem = s.tmem_base + (uint32_t)cx.wg * 256u;
  cx.lane_addr = cx.tmem + ((uint32_t)(cx.tid_wg & ~31) << 16);
  cx.full = s.bars + 1 + 2 * cx.wg;
  cx.n_issued = cx.n_waited = 0;
  return cx;
}

// =================================================================================================
// K1: stage A + B
// =================================================================================================
__global__ void __launch_bounds__(256, 1) k1_stage_ab_kernel(const __grid_constant__ K1Params p) {
  extern __shared__ uint8_t smem_raw[];
  const CtaSetup s = cta_prologue(smem_raw, k1Bars, p.wimg, k1WBytes);
  WgCtx cx = make_wg(s);
  const uint32_t wsm = smem_u32(s.smem);
  const Geometry& g = p.g;
  const long ntiles = (p.q_end - p.q_begin + kTile - 1) / kTile;
  const uint4* __restrict__ tab4 = reinterpret_cast<const uint4*>(p.tab);  // 32 uint4 per texel

  for (long tile = (long)blockIdx.x * 2 + cx.wg; tile < ntiles; tile += (long)gridDim.x * 2) {
    const long q = p.q_begin + tile * kTile + cx.tid_wg;
    const bool valid = q < p.q_end;
    const long qc = valid ? q : p.q_end - 1;
    const int jy = (int)(qc / g.WW), jx = (int)(qc - (long)jy * g.WW);

    // ---- stage A, first layer (hoisted): h0 = sin(TA[iy,ix] + rel . w_rel + cA)      (:382-400)
    {
      const int iy = g.y.idx[jy], ix = g.x.idx[jx];
      const float rely = g.y.rel[jy], relx = g.x.rel[jx];
      const bool inb = (iy >= 0) & (iy < g.H) & (ix >= 0) & (ix < g.W);
      const uint4* ta = tab4 + (inb ? ((long)iy * g.W + ix) : 0) * 32;
      const float mask = inb ? 1.f : 0.f;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 v = __ldg(ta + half * 4 + j);
          const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = half * 32 + j * 8 + e * 2;
            const float2 f = unpack_half2(w4[e]);
            const float a0 = fmaf(f.x, mask, fmaf(rely, p.c.a_rel[2 * c], fmaf(relx, p.c.a_rel[2 * c + 1], p.c.cA[c])));
            const float a1 = fmaf(f.y, mask, fmaf(rely, p.c.a_rel[2 * c + 2], fmaf(relx, p.c.a_rel[2 * c + 3], p.c.cA[c + 1])));
            pk[j * 4 + e] = pack_bf16x2(fast_sin(a0), fast_sin(a1));
          }
        }
        tmem_st16(cx.lane_addr + kColAin + half * 16, pk);
      }
      tmem_st_wait();
    }
    tc_fence_before();
    wg_barrier(cx.wg);

    // ---- feat_imnet hidden layers
    run_layer<1>(cx, 0, kColAin, wsm + k1F1, 64, 4, [](int) { return 0; },
                 [&](int, uint32_t d) { epi_sin_to_tmem(d, cx.lane_addr + kColAin, p.c.f1_b); });
    run_layer<4>(cx, 0, kColAin, wsm + k1F2, 256, 4, [](int i) { return i; },
                 [&](int i, uint32_t d) { epi_sin_to_tmem(d, cx.lane_addr + kColA + 32 * i, p.c.f2_b + 64 * i); });

    // ---- stage B gather, issued before the composed layer so its latency hides behind the MMAs:
    //      gB = bilinear(TB; query position) + cB + composed bias of F                  (:410-418)
    float gB[64];
    {
      const Taps tp = make_taps_tables(g, jy, jx);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = p.c.cB[8 * j + e] + p.c.f3_b[8 * j + e];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint4 v = __ldg(tab4 + (long)tp.off[k] * 32 + 8 + j);
          const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = unpack_half2(w4[e]);
            acc[2 * e] = fmaf(tp.w[k], f.x, acc[2 * e]);
            acc[2 * e + 1] = fmaf(tp.w[k], f.y, acc[2 * e + 1]);
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) gB[8 * j + e] = acc[e];
      }
    }

    // ---- composed last layer of feat_imnet: chunk order Q1, Q2, F (F last: its epilogue overwrites h2)
    run_layer<3>(cx, 0, kColA, wsm + k1F3, 192, 16, [](int i) { return i == 2 ? 0 : i + 1; },
                 [&](int i, uint32_t d) {
                   if (i < 2) epi_store_qtab(d, p.c.f3_b + 64 * (i + 1), p.qtab + qc * 128 + 64 * i, valid);
                   else epi_flow_first_layer(d, cx.lane_addr + kColAin, gB);
                 });

    // ---- flow_imnet hidden layers; the 256->4 output layer rides the FMA pipe          (:419-422)
    run_layer<1>(cx, 0, kColAin, wsm + k1L1, 64, 4, [](int) { return 0; },
                 [&](int, uint32_t d) { epi_sin_to_tmem(d, cx.lane_addr + kColAin, p.c.l1_b); });
    float fl[4] = {p.c.l3_b[0], p.c.l3_b[1], p.c.l3_b[2], p.c.l3_b[3]};
    run_layer<4>(cx, 0, kColAin, wsm + k1L2, 256, 4, [](int i) { return i; },
                 [&](int i, uint32_t d) { epi_sin_fma<4>(d, p.c.l2_b + 64 * i, p.c.l3_w + 64 * i, fl); });
    if (valid) reinterpret_cast<float4*>(p.flow)[q] = make_float4(fl[0], fl[1], fl[2], fl[3]);
  }
  cta_epilogue(s.tmem_base);
}

// =================================================================================================
// K2: stage C + D + E
// =================================================================================================
// Warp-cooperative gather of encode_imnet's (hoisted) first layer for the warp's 32 queries:
// lanes 0..15 / 16..31 first compute the two warps' bilinear footprints of 16 queries (thread per
// query), stage them in the warp's private 2 KB of shared memory, then the whole warp walks the
// queries one at a time with lane = channel pair, so every tap is one coalesced 128-byte load.
__device__ __forceinline__ void k2_gather(const K2Params& p, uint8_t* a0, uint4* stg, long tile_q0, int warp_in_wg, int lane,
                                          float cE0, float cE1) {
  const Geometry& g = p.g;
  const char* __restrict__ qtab_b = reinterpret_cast<const char*>(p.qtab);
  const char* __restrict__ tab_b = reinterpret_cast<const char*>(p.tab);
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    {
      const int qi = lane & 15, which = lane >> 4;
      const long q = min(tile_q0 + warp_in_wg * 32 + pass * 16 + qi, p.q_end - 1);
      const int jy = (int)(q / g.WW), jx = (int)(q - (long)jy * g.WW);
      const float4 fl = __ldg(reinterpret_cast<const float4*>(p.flow) + q);
      float gy, gx;
      warp_position(g, jy, jx, which ? fl.z : fl.x, which ? fl.w : fl.y, gy, gx);   // (warplayer.py:25-39)
      const Taps hr = make_taps(gy, gx, g.HH, g.WW);
      const Taps lr = make_taps(gy, gx, g.H, g.W);
      if (p.band_mode) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (hr.w[k] != 0.f) {
            const int row = hr.off[k] / g.WW;
            if (row < p.band_lo || row >= p.band_hi) atomicOr(p.flag, 1);
          }
      }
      uint4 oh, ol, wh, wl;
      oh.x = (uint32_t)hr.off[0] * 256u + which * 128u; oh.y = (uint32_t)hr.off[1] * 256u + which * 128u;
      oh.z = (uint32_t)hr.off[2] * 256u + which * 128u; oh.w = (uint32_t)hr.off[3] * 256u + which * 128u;
      ol.x = (uint32_t)lr.off[0] * 512u + 256u + which * 128u; ol.y = (uint32_t)lr.off[1] * 512u + 256u + which * 128u;
      ol.z = (uint32_t)lr.off[2] * 512u + 256u + which * 128u; ol.w = (uint32_t)lr.off[3] * 512u + 256u + which * 128u;
      wh = make_uint4(__float_as_uint(hr.w[0]), __float_as_uint(hr.w[1]), __float_as_uint(hr.w[2]), __float_as_uint(hr.w[3]));
      wl = make_uint4(__float_as_uint(lr.w[0]), __float_as_uint(lr.w[1]), __float_as_uint(lr.w[2]), __float_as_uint(lr.w[3]));
      uint4* dst = stg + qi * 8 + which * 4;
      dst[0] = oh; dst[1] = ol; dst[2] = wh; dst[3] = wl;
    }
    __syncwarp();
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
      const uint4* sq = stg + i * 8;
      float acc0 = cE0, acc1 = cE1;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        const uint4 oh = sq[which * 4 + 0], ol = sq[which * 4 + 1], wh = sq[which * 4 + 2], wl = sq[which * 4 + 3];
        const uint32_t o_h[4] = {oh.x, oh.y, oh.z, oh.w}, o_l[4] = {ol.x, ol.y, ol.z, ol.w};
        const uint32_t w_h[4] = {wh.x, wh.y, wh.z, wh.w}, w_l[4] = {wl.x, wl.y, wl.z, wl.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_half2(__ldg(reinterpret_cast<const uint32_t*>(qtab_b + o_h[k]) + lane));
          const float w = __uint_as_float(w_h[k]);
          acc0 = fmaf(w, f.x, acc0);
          acc1 = fmaf(w, f.y, acc1);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_half2(__ldg(reinterpret_cast<const uint32_t*>(tab_b + o_l[k]) + lane));
          const float w = __uint_as_float(w_l[k]);
          acc0 = fmaf(w, f.x, acc0);
          acc1 = fmaf(w, f.y, acc1);
        }
      }
      const int r = warp_in_wg * 32 + pass * 16 + i;
      *reinterpret_cast<uint32_t*>(a0 + sw128_offset(r, 2 * lane)) = pack_bf16x2(fast_sin(acc0), fast_sin(acc1));
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256, 1) k2_stage_cde_kernel(const __grid_constant__ K2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const CtaSetup s = cta_prologue(smem_raw, k2Bars, p.wimg, k2WBytes);
  WgCtx cx = make_wg(s);
  const uint32_t wsm = smem_u32(s.smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warp_in_wg = warp & 3;
  uint8_t* a0 = s.smem + k2A0 + cx.wg * 16384;
  uint4* stg = reinterpret_cast<uint4*>(s.smem + k2Taps + warp * 2048);
  const float cE0 = p.c.cE[2 * lane], cE1 = p.c.cE[2 * lane + 1];
  const long ntiles = (p.q_end - p.q_begin + kTile - 1) / kTile;

  for (long tile = (long)blockIdx.x * 2 + cx.wg; tile < ntiles; tile += (long)gridDim.x * 2) {
    const long tile_q0 = p.q_begin + tile * kTile;
    const long q = tile_q0 + cx.tid_wg;
    const bool valid = q < p.q_end;

    // ---- stage C + D + first layer of encode_imnet (hoisted)                         (:424-456)
    k2_gather(p, a0, stg, tile_q0, warp_in_wg, lane, cE0, cE1);
    fence_proxy_async_smem();
    tc_fence_before();
    wg_barrier(cx.wg);

    // ---- encode_imnet hidden layers; the 256->3 output layer rides the FMA pipe       (:456-457)
    run_layer<1>(cx, smem_u32(a0), 0, wsm + k2E1, 64, 4, [](int) { return 0; },
                 [&](int, uint32_t d) { epi_sin_to_tmem(d, cx.lane_addr + kColAin, p.c.e1_b); });
    run_layer<4>(cx, 0, kColAin, wsm + k2E2, 256, 4, [](int i) { return i; },
                 [&](int i, uint32_t d) { epi_sin_to_tmem(d, cx.lane_addr + kColA + 32 * i, p.c.e2_b + 64 * i); });
    float rgb[3] = {p.c.e4_b[0], p.c.e4_b[1], p.c.e4_b[2]};
    run_layer<4>(cx, 0, kColA, wsm + k2E3, 256, 16, [](int i) { return i; },
                 [&](int i, uint32_t d) { epi_sin_fma<3>(d, p.c.e3_b + 64 * i, p.c.e4_w + 64 * i, rgb); });
    if (valid) {
      p.out[q] = rgb[0];
      p.out[p.plane + q] = rgb[1];
      p.out[2 * p.plane + q] = rgb[2];
    }
  }
  cta_epilogue(s.tmem_base);
}

void fill_from(float* dst, const std::vector<float>& src, size_t n) { std::copy(src.begin(), src.begin() + n, dst); }

}  // namespace

// =================================================================================================
// host side
// =================================================================================================
struct TcWeights {
  uint8_t* d_k1 = nullptr;
  uint8_t* d_k2 = nullptr;
  K1Consts c1{};
  K2Consts c2{};
  std::vector<float> a_t, a_b, b_t, b_b, e_t, e_b;
  bool attrs_set = false;
};

TcWeights* tc_weights_create(const FoldedWeights& hw, std::string& err) {
  auto* t = new TcWeights();
  std::vector<uint8_t> i1, i2;
  append_sw128_image(i1, hw.f1_w.data(), 64, 64);
  append_sw128_image(i1, hw.f2_w.data(), 256, 64);
  append_sw128_image(i1, hw.f3_w.data(), 192, 256);
  append_sw128_image(i1, hw.l1_w.data(), 64, 64);
  append_sw128_image(i1, hw.l2_w.data(), 256, 64);
  append_sw128_image(i2, hw.e1_w.data(), 64, 64);
  append_sw128_image(i2, hw.e2_w.data(), 256, 64);
  append_sw128_image(i2, hw.e3_w.data(), 256, 256);
  if (i1.size() != k1WBytes || i2.size() != k2WBytes) {
    err = "internal: weight image size mismatch";
    delete t;
    return nullptr;
  }
  cudaError_t e = cudaMalloc(&t->d_k1, i1.size());
  if (e == cudaSuccess) e = cudaMalloc(&t->d_k2, i2.size());
  if (e == cudaSuccess) e = cudaMemcpy(t->d_k1, i1.data(), i1.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(t->d_k2, i2.data(), i2.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k1_stage_ab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1Smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k2_stage_cde_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2Smem);
  if (e != cudaSuccess) {
    err = cudaGetErrorString(e);
    tc_weights_destroy(t);
    return nullptr;
  }
  fill_from(t->c1.a_rel, hw.a_rel, 128);
  fill_from(t->c1.f1_b, hw.f1_b, 64);
  fill_from(t->c1.f2_b, hw.f2_b, 256);
  fill_from(t->c1.f3_b, hw.f3_b, 192);
  fill_from(t->c1.l1_b, hw.l1_b, 64);
  fill_from(t->c1.l2_b, hw.l2_b, 256);
  fill_from(t->c1.l3_w, hw.l3_w, 1024);
  fill_from(t->c1.l3_b, hw.l3_b, 4);
  fill_from(t->c2.e1_b, hw.e1_b, 64);
  fill_from(t->c2.e2_b, hw.e2_b, 256);
  fill_from(t->c2.e3_b, hw.e3_b, 256);
  fill_from(t->c2.e4_w, hw.e4_w, 768);
  fill_from(t->c2.e4_b, hw.e4_b, 3);
  t->c2.e4_b[3] = 0.f;
  t->a_t = hw.a_t; t->a_b = hw.a_b; t->b_t = hw.b_t; t->b_b = hw.b_b; t->e_t = hw.e_t; t->e_b = hw.e_b;
  return t;
}

void tc_weights_destroy(TcWeights* t) {
  if (!t) return;
  if (t->d_k1) cudaFree(t->d_k1);
  if (t->d_k2) cudaFree(t->d_k2);
  delete t;
}

cudaError_t decode_slab_tc(const LaunchCtx& cx, const TcWeights* tw, const Geometry& geo, const Workspace& ws, float t,
                           int row_begin, int row_end, int k1_row_begin, int k1_row_end, float* out_rgb, int stage) {
  const long WW = geo.WW;
  if (stage == 1) {
    K1Params p;
    p.c = tw->c1;
    for (int c = 0; c < 64; ++c) {
      p.c.cA[c] = tw->a_t[c] * t + tw->a_b[c];
      p.c.cB[c] = tw->b_t[c] * t + tw->b_b[c];
    }
    p.g = geo;
    p.tab = reinterpret_cast<const __half*>(ws.tab);
    p.qtab = reinterpret_cast<__half*>(ws.qtab);
    p.flow = ws.flow;
    p.wimg = tw->d_k1;
    p.q_begin = k1_row_begin * WW;
    p.q_end = k1_row_end * WW;
    const long ntiles = (p.q_end - p.q_begin + kTile - 1) / kTile;
    const int grid = (int)std::min<long>(cx.num_sms, (ntiles + 1) / 2);
    k1_stage_ab_kernel<<<grid, 256, k1Smem, cx.stream>>>(p);
    ++*cx.launch_counter;
    return cudaGetLastError();
  }
  K2Params p;
  p.c = tw->c2;
  for (int c = 0; c < 64; ++c) p.c.cE[c] = tw->e_t[c] * t + tw->e_b[c];
  p.g = geo;
  p.tab = reinterpret_cast<const __half*>(ws.tab);
  p.qtab = reinterpret_cast<const __half*>(ws.qtab);
  p.flow = ws.flow;
  p.out = out_rgb;
  p.plane = (long)geo.HH * geo.WW;
  p.wimg = tw->d_k2;
  p.q_begin = row_begin * WW;
  p.q_end = row_end * WW;
  p.band_lo = k1_row_begin;
  p.band_hi = k1_row_end;
  p.band_mode = (k1_row_begin > 0 || k1_row_end < geo.HH) ? 1 : 0;
  p.flag = ws.flag;
  const long ntiles = (p.q_end - p.q_begin + kTile - 1) / kTile;
  const int grid = (int)std::min<long>(cx.num_sms, (ntiles + 1) / 2);
  k2_stage_cde_kernel<<<grid, 256, k2Smem, cx.stream>>>(p);
  ++*cx.launch_counter;
  return cudaGetLastError();
}

}  // namespace stif
