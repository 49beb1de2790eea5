// kernels_tc.cu -- the fused bf16 tensor-core path of the STIF query decoder (STIF_MODE_BF16).
//
// Three kernels, all persistent (one CTA per SM), all tcgen05 / TMEM:
//
//   K0  latent projection (per frame pair, t-independent): tab[texel, 256] = W_tab . [latent; frames]
//       -- the hoisted first layers of the three SIRENs.
//   K1  stage A + B  (reference Sakuya_arch_test.py:382-422): nearest gather of the projected latent
//       -> feat_imnet trunk -> composed last layer writes the projected HR table (Q1|Q2, fp16) ->
//       + bilinear gather of TB -> flow_imnet trunk -> flow (fp32).
//   K2  stage C + D + E (warplayer.py:25-39, :424-458): warp positions from the flow, four bilinear
//       gathers (Q1@g1, Q2@g2, TE1@g1, TE2@g2) -> encode_imnet trunk -> RGB (fp32 planar).
//
// K1/K2: 512 threads = two "workgroups" (WG, 8 warps) that each own a 128-query tile and half of
// tensor memory.  Per tile every MLP layer is a chain of tcgen05.mma (M=128 queries, N=64 output
// chunk, K=16 per instruction, bf16 x bf16 -> fp32 in TMEM).  Weights stay resident in shared
// memory for the whole kernel (bulk-TMA loaded once); activations never leave the SM: the
// epilogue warps read an accumulator chunk (tcgen05.ld), apply bias + sine (MUFU), round to bf16
// and write it back to TMEM as the next layer's A operand (tcgen05.st, "TS" MMA form).  The
// 256->4 (flow) and 256->3 (RGB) output layers run on the FMA pipe in fp32 (packed FFMA2) straight
// from the sine outputs.  Two warps share each TMEM lane quarter and split a chunk's 64 columns.
//
// TMEM map of one WG (256 columns): A-area [0,128) = up to 256 bf16 activations per query;
// D-area [128,256) = two 64-column fp32 accumulator slots used as a ring (MMA of chunk c+1
// overlaps the sine epilogue of chunk c; the other WG's tile fills the remaining bubbles).
#include <algorithm>
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

constexpr int kTile = 128;
#ifndef KPOLY
#define KPOLY 0
#endif
constexpr int kPolyPairs = KPOLY;   // of the 16 sine pairs per thread per chunk, this many run on the FMA pipe (poly_sin2)
constexpr uint32_t kColA = 0;     // A-area: 256-wide activations, channel k at column k/2
constexpr uint32_t kColAin = 96;  // 64-wide activations live in the last 32 columns of the A-area
constexpr uint32_t kColD = 128;   // two accumulator slots: [128,192), [192,256)

// ---- shared-memory images (bytes) ----------------------------------------------------------
constexpr uint32_t kW64x64 = 64 * 64 * 2, kW256x64 = 256 * 64 * 2, kW192x256 = 192 * 256 * 2, kW256x256 = 256 * 256 * 2;
// K1: F1 | F2 | F3(composed 192x256) | L1 | L2 | partial-sum exchange
constexpr uint32_t k1F1 = 0, k1F2 = k1F1 + kW64x64, k1F3 = k1F2 + kW256x64, k1L1 = k1F3 + kW192x256,
                   k1L2 = k1L1 + kW64x64, k1WBytes = k1L2 + kW256x64;
// The fp32 output layers (256 -> 4 flow, 256 -> 3 RGB) are read once per (thread, column) on the FMA pipe.  From shared
// memory that is one broadcast LDS.128 per four weights; from the constant bank it is a uniform load (LDCU.64) per two
// (measured: K1 -3 %, K2 -1 %).  The biases stay in the constant bank: an LDS in the middle of the MUFU-bound sine
// epilogues competes with the MUFU for the MIO queue and made K1 6 % slower.
constexpr uint32_t kc1L3W = 0, kc1Floats = 1024;
constexpr uint32_t kc2E4W = 0, kc2Floats = 768;
constexpr uint32_t k1Part = k1WBytes, k1Const = k1Part + 2 * 128 * 16, k1Bars = k1Const + kc1Floats * 4, k1Smem = k1Bars + 128;
// K2: E1 | E2 | E3 | A0[2 WGs] (tap staging of the NEXT tile aliases the A tile once the first MMA has read it) |
//     partial-sum exchange | fp32 constants
constexpr uint32_t k2E1 = 0, k2E2 = k2E1 + kW64x64, k2E3 = k2E2 + kW256x64, k2WBytes = k2E3 + kW256x256;
constexpr uint32_t k2A0 = k2WBytes, k2Part = k2A0 + 2 * 16384, k2Const = k2Part + 2 * 128 * 16, k2Bars = k2Const + kc2Floats * 4,
                   k2Smem = k2Bars + 128;
// K2, three-workgroup rotation: the same with three A tiles and three partial-sum exchanges
constexpr uint32_t k2rPart = k2A0 + 3 * 16384, k2rConst = k2rPart + 3 * 128 * 16, k2rBars = k2rConst + kc2Floats * 4, k2rSmem = k2rBars + 128;
static_assert(k2rSmem <= 232448, "exceeds 227 KB of shared memory");
// K0: A tile (4 K-blocks x 128 rows) | W_tab (4 K-blocks x 256 rows)
constexpr uint32_t k0A = 0, k0B = 4 * 128 * 128, k0WBytes = 4 * 256 * 128, k0Bars = k0B + k0WBytes, k0Smem = k0Bars + 128;
static_assert(k2A0 % 1024 == 0 && k1F3 % 1024 == 0 && k1L1 % 1024 == 0 && k2E3 % 1024 == 0 && k0B % 1024 == 0,
              "SW128 tiles need 1024 B alignment");
static_assert(k2Smem <= 232448 && k1Smem <= 232448 && k0Smem <= 232448, "exceeds 227 KB of shared memory");

struct alignas(16) K1Consts {
  float cA[64];       // feat_imnet L0: w_t * t + b        (per launch)
  float a_rel[128];   // feat_imnet L0: (rel_y, rel_x) columns, [c][2]
  float f1_b[64], f2_b[256], f3_b[192];
  float cB[64];       // flow_imnet L0: w_t * t + b        (per launch)
  float l1_b[64], l2_b[256];
  float l3_w[4 * 256], l3_b[4];
};
struct alignas(16) K2Consts {
  float cE[64];       // encode_imnet L0: w_t * t + b      (per launch)
  float e1_b[64], e2_b[256], e3_b[256];
  float e4_w[3 * 256], e4_b[4];
};
// MULTI launches (band-major host pipeline): one launch covers the same rows of up to four timesteps ("slabs"), each with its
// own time constants and tables; tile t of the launch is tile t % tiles_per_slab of slab t / tiles_per_slab.  Halves the
// launch count of a band, and with it the pipeline fill / drain each persistent launch pays (~25-45 us).
constexpr int kMaxSlabs = kMaxSlabsHost;
struct K1Slab { float cA[64]; float cB[64]; __half* qtab; float* flow; };
struct K2Slab { float cE[64]; const __half* qtab; const float* flow; float* out; uint8_t* out_u8; };
struct K1Params {
  K1Consts c;
  K1Slab slab[kMaxSlabs];     // MULTI launches only
  long tiles_per_slab, ntiles_total;
  Geometry g;
  const __half* tab;   // [H*W,256]   TA | TB | TE1 | TE2
  __half* qtab;        // [HH*WW,128] Q1 | Q2
  float* flow;         // [HH*WW,4]
  const __half* utab;  // decoding_test at x4 only: [HH*WW,192] fp16 UB | UE1 | UE2 (frame terms on the query grid), else null
  float* ftab;         // local-ensemble passes only: F + composed bias, [HH*WW,64] fp32 (stage B reads it at OTHER pixels)
  const uint8_t* wimg;
  long q_begin, q_end;
  long long* trace;    // debug: clock64 timestamps of block 0 (STIF_TRACE=<file>), else null
};
struct K2Params {
  K2Consts c;
  K2Slab slab[kMaxSlabs];     // MULTI launches only
  long tiles_per_slab, ntiles_total;
  Geometry g;
  const __half* tab;
  const __half* qtab;
  const float* flow;
  float* out;          // [3, plane] fp32 planar, or ...
  uint8_t* out_u8;     // ... [plane, 3] uint8 HWC (STIF_FLAG_OUT_U8: custom_video_test.py:102's conversion in the output stage), else null
  long plane;
  const uint8_t* wimg;
  int row_begin, row_end;   // output rows this launch decodes
  int col_begin, col_end;   // ... and columns (the whole width except for decoding_memory's zoom window)
  const __half* uadd;       // decoding_test away from x4: [HH*WW, 64] upsampled-frame terms at the warped positions (null otherwise)
  int tiles_x;              // K2 tiles are 8 rows x 16 columns (2 x 4 warp patches of 4 x 4 queries): neighbouring
                            // queries of a warp share bilinear taps in x AND y, so the L1 / in-flight-miss merge removes
                            // most of the duplicate line requests of the gather
  int band_lo_off, band_hi_off;   // Q-table pixels [lo, hi) exist (stage A+B rows of this launch); taps outside raise *flag
  int* flag;
  long long* trace;
};
struct K0Params {
  const float* latent;  // [192, HW] fp32, or ...
  const uint16_t* latent16;   // ... bf16 bit patterns (stif_decode_host_bf16: the projection rounds the latent to bf16 anyway), else null
  const float* frames;  // [6, HW]
  __half* tab;          // [HW, 256]
  const uint8_t* wimg;  // W_tab, bf16, [256 x 256 (K padded)] SW128 image
  long HW;
  long m_begin, m_end;  // texel range of this launch (row bands let the host pipeline H2D copies with K0)
};

extern __shared__ __align__(1024) uint8_t smem[];

// ---- workgroup context -----------------------------------------------------------------------
struct WgCtx {
  uint32_t tmem;       // TMEM address of this WG's column 0, lane 0
  uint32_t lane_addr;  // same + this warp's lane quarter (for tcgen05.ld/st)
  uint64_t* full;      // two mbarriers: accumulator slot s is complete
  uint32_t n_issued, n_waited, n_steps;
  int wg, tid_wg, warp_in_wg, row, colhalf;
  int bar_base;        // first of the two alternating step-barrier ids of the TMEM slot this warp works on (1 + 2 * slot)
  int quarter;         // TMEM lane quarter of this warp (warp id % 4)
  int slot;            // logical warp slot: 0..15 epilogue warps (WG0 then WG1), 16/17 the issuers (trace rows, per-warp smem)
  bool issuer;         // this warp only issues the WG's MMAs (warps 0, 1); the other 8 warps of the WG only run epilogues
  uint64_t* extra_commit;   // issuer only: a second mbarrier every chunk's tcgen05.commit also arrives on while it is set
  uint32_t bias_smem;       // issuer only: shared address of the current layer's bias blocks (one K = 16 block per chunk), 0 = none
  uint32_t ones_smem;       // issuer only: the constant A tile of the bias K step
  long long* trace;    // this thread's trace cursor (null unless tracing)
};

// debug timeline: one (tag, clock) pair per call, only for the traced threads of block 0
// (compiled in only with -DSTIF_ENABLE_TRACE: even predicated-off marks cost ~8% of K1's issue slots)
__device__ __forceinline__ void trace_mark(WgCtx& cx, int tag) {
#ifdef STIF_ENABLE_TRACE
  if (cx.trace) {
    cx.trace[0] = tag;
    cx.trace[1] = clock64();
    cx.trace += 2;
  }
#else
  (void)cx; (void)tag;
#endif
}

// Named barriers.  Per WG: two alternating "step" barriers (ids 1..4, 288 threads = 8 epilogue warps that only ARRIVE
// + the issuer warp that SYNCs: the epilogue warps never wait for each other, only for accumulators), and one
// epilogue-only barrier (ids 5, 6, 256 threads) for the rare smem exchanges.  A warp can run at most one step ahead of
// the slowest warp of its WG (step s+2's accumulator is issued only after everyone arrived for step s), hence two ids.
__device__ __forceinline__ void wg_barrier(int wg) { asm volatile("bar.sync %0, 256;" ::"r"(wg + 5) : "memory"); }
template <bool ISSUER>
__device__ __forceinline__ void step_done(WgCtx& cx) {
  const int id = cx.bar_base + (int)(cx.n_steps & 1);
  if (ISSUER) asm volatile("bar.sync %0, 288;" ::"r"(id) : "memory");
  else asm volatile("bar.arrive %0, 288;" ::"r"(id) : "memory");
  ++cx.n_steps;
}

// K2: the two workgroups take turns at the gather.  Left alone they drift into lockstep -- the trailing WG finds its
// neighbour tile's table lines in L1 and catches up -- and then both sit in their gather (memory latency, MUFU idle)
// and both run their sine epilogues at the same time.  WG1 may start the gather of its j-th tile only after WG0 has
// finished the gather of ITS j-th tile (named barrier 7: WG0's 256 threads arrive, WG1's 256 sync).  Measured: K2 -6 %.
// WG0 in turn may not signal tile j+1 before WG1 has taken the signal for tile j (barrier 8, roles swapped).  That second
// barrier costs WG0 nothing in steady state (WG1 picked the previous signal up a whole tile ago) but it is REQUIRED:
// without it WG0 can get two tiles ahead, and two arrivals of the same 256 threads complete a 512-thread phase of
// barrier 7 on their own -- the pairing is lost and WG1's last sync waits forever (seen as a hang at 4K).  (A handshake
// in both directions at the half-tile points, by contrast, serialised the WGs and was 25 % slower.)
#ifndef STIF_EARLY_RELEASE
#define STIF_EARLY_RELEASE 1
#endif
// -DSTIF_CHECK_BOUNDS: every table index the gathers / stores of K1 and K2 form is range-checked on the device and a
// violation traps (the launch fails, the C ABI returns STIF_ECUDA).  compute-sanitizer is closed on the GPU pool this was
// developed on, so the parity suite is run once per round against a build with these checks instead (profiles/).
#ifdef STIF_CHECK_BOUNDS
#define STIF_BOUND(idx, n) do { if ((long)(idx) < 0 || (long)(idx) >= (long)(n)) __trap(); } while (0)
#else
#define STIF_BOUND(idx, n) do { } while (0)
#endif
// Diagnostic builds only (results are WRONG): bit 0 drops the Q-table stores, bit 1 the TA loads, bit 2 the stage-B (TB) loads,
// bit 3 K2's Q-table tap loads, bit 4 K2's TE tap loads -- what each memory stream costs end to end (profiles/diag_streams.sh).
#ifndef STIF_DIAG
#define STIF_DIAG 0
#endif
// WG0 hands the gather turn over once the loads of half-step STIF_TURN_EARLY (1..8) are in flight; 9 = after the whole gather and
// its step barrier.  Measured (K2 ms per launch at config 2): 9: 0.594, 8: 0.580, 7: 0.576-0.580, 6: 0.600, 4: 0.594 -- with all of
// WG0's loads issued, WG1's first load batch overlaps WG0's last blend + sines instead of waiting behind them.
#ifndef STIF_TURN_EARLY
#define STIF_TURN_EARLY 8
#endif
#ifndef STIF_GATHER_TURNS
#define STIF_GATHER_TURNS 1
#endif

// Both sides are unconditional (no tile counts): WG0 has as many tiles as WG1 or one more, so every sync finds its
// partner; the at most one unmatched arrival per barrier is the very last one and nobody waits behind it.
// WG1, top of every tile: wait for WG0's signal, acknowledge it.
__device__ __forceinline__ void gather_turn_wait(const WgCtx& cx) {
  if (STIF_GATHER_TURNS && cx.wg == 1) {
    asm volatile("bar.sync 7, 512;" ::: "memory");
    asm volatile("bar.arrive 8, 512;" ::: "memory");
  }
}
// WG0, after the gather of every tile: take WG1's acknowledgement of the previous signal (none before the first), signal.
__device__ __forceinline__ void gather_turn_done(const WgCtx& cx, bool first_tile) {
  if (STIF_GATHER_TURNS && cx.wg == 0) {
    if (!first_tile) asm volatile("bar.sync 8, 512;" ::: "memory");
    asm volatile("bar.arrive 7, 512;" ::: "memory");
  }
}

// ---- sine turns ------------------------------------------------------------------------------------------------------
// The two workgroups share one MUFU pipe per SM sub-partition.  Left alone they run in LOCKSTEP (same phase of their tiles
// at the same time): both burn sines together, each at half rate, and both then sit in the same MMA wait / gather / store
// phase with the MUFU idle -- a tile costs 2 S + N clocks (S = its sine time alone, N = everything else), measured.  The
// sine-heavy phases ("BIG": the four-chunk 64->256 layers, 2 x 256 of K1's 768 sines per query; K2: 64->256 + 256->256,
// 512 of 640; STIF_K2_SINE_TURNS = 1: those two layers as one phase, 2: as two phases, 3: all three hidden layers as one)
// are therefore taken in turns: a token alternates WG0 -> WG1 -> WG0 ..., so the holder's sines run at the
// full MUFU rate while the other WG does its MMA waits, gathers, Q-table stores and small layers underneath: S + N.
// Named barriers A (WG0's 256 epilogue threads arrive, WG1's sync) and B (roles swapped); every arrival is consumed by
// exactly one sync before the next arrival on the same barrier can happen (the arriving WG needs the other barrier's
// hand-back first), and nobody waits for an arrival that is never made: WG0 owns tile 2b + 2Gj, WG1 tile 2b + 1 + 2Gj of
// round j, so WG1's rounds are a prefix of WG0's and "does my partner have this / the next round" is a tile-index test.
#ifndef STIF_K1_SINE_TURNS
#define STIF_K1_SINE_TURNS 1
#endif
#ifndef STIF_K1_ROT
#define STIF_K1_ROT 1
#endif
#ifndef STIF_K2_ROT
#define STIF_K2_ROT 1
#endif
#ifndef STIF_K2_SINE_TURNS
#define STIF_K2_SINE_TURNS 0
#endif
template <int BAR_A, int BAR_B>
struct SineTurn {
  int wg;
  bool first_round;     // WG0: no hand-back to take before the very first phase
  bool partner_now;     // WG0: WG1 has a tile in this round
  bool partner_next;    // WG1: WG0 has a tile in the next round
  // phase s of NPH phases per round
  template <int NPH>
  __device__ __forceinline__ void acquire(int s) const {
    if (wg == 0) {
      const bool take = s > 0 ? partner_now : !first_round;
      if (take) asm volatile("bar.sync %0, 512;" ::"n"(BAR_B) : "memory");
    } else {
      asm volatile("bar.sync %0, 512;" ::"n"(BAR_A) : "memory");
    }
  }
  template <int NPH>
  __device__ __forceinline__ void release(int s) const {
    if (wg == 0) {
      if (partner_now) asm volatile("bar.arrive %0, 512;" ::"n"(BAR_A) : "memory");
    } else {
      if (s + 1 < NPH || partner_next) asm volatile("bar.arrive %0, 512;" ::"n"(BAR_B) : "memory");
    }
  }
};

// STIF_WAIT_MODE: 0 = plain try_wait loop; 1 = try_wait with a suspend-time hint (the warp sleeps in hardware instead of
// re-issuing the probe: spinning warps were taking ~15 % of the rotation kernel's issue slots); 2 = nanosleep back-off.
#ifndef STIF_WAIT_MODE
#define STIF_WAIT_MODE 0
#endif
#if STIF_WAIT_MODE == 1
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
#endif
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, uint32_t parity) {
  // try_wait suspends the warp in hardware for a bounded time, so this loop turns only a few times per wait; the
  // iteration cap converts a lost arrival into a launch failure instead of a hung GPU.
#if STIF_WAIT_MODE == 1
  for (uint32_t it = 0; !mbar_try_wait_hint(bar, parity, 20000u); ++it)
    if (it > (1u << 20)) __trap();
#elif STIF_WAIT_MODE == 2
  if (mbar_try_wait(bar, parity)) return;
  for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it) {
    __nanosleep(40);
    if (it > (1u << 22)) __trap();
  }
#else
  for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it)
    if (it > (1u << 24)) __trap();
#endif
}

// Long waits (a workgroup of the rotation kernels waiting for a TMEM slot: thousands of clocks): back off between probes so
// that eight spinning warps do not take issue slots from the sixteen that hold the slots (the spin loop was 25 % of K1's
// executed instructions).  STIF_LONG_WAIT_NS = 0 falls back to the plain loop.
#ifndef STIF_LONG_WAIT_NS
#define STIF_LONG_WAIT_NS 0
#endif
__device__ __forceinline__ void mbar_wait_long(uint64_t* bar, uint32_t parity) {
#if STIF_LONG_WAIT_NS == 0
  mbar_wait_or_trap(bar, parity);
#else
  for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it) {
    __nanosleep(STIF_LONG_WAIT_NS);
    if (it > (1u << 22)) __trap();
  }
#endif
}

__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
  return p != 0;
}

constexpr uint32_t kBiasBlock = 64 * 32, kOnesTile = 128 * 32;   // interleaved K = 16 operands (tc_primitives.cuh: nosw_offset)
// Issue the MMAs of one 64-wide output chunk and commit to the slot's barrier.  Executed by warp 0 of
// the WG with warp-uniform operands (one elected lane issues), so the descriptors live in uniform
// registers and each tcgen05.mma costs a couple of instructions.
//   A_SMEM: A operand is an SW128 tile in shared memory at a_base (K = 64); else A is in TMEM at address a_base.
template <int KSTEPS, bool A_SMEM, bool ISSUER>
__device__ __forceinline__ void issue_chunk(WgCtx& cx, uint32_t a_base, uint32_t w_smem, int n_rows, int chunk) {
  const uint32_t slot = cx.n_issued & 1;
  if (ISSUER) {
    trace_mark(cx, 40);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, 64);
    const uint32_t d = cx.tmem + kColD + slot * 64;
    const uint64_t b0 = make_desc_sw128(w_smem + (uint32_t)chunk * 8192u);
    const uint64_t kb = (uint64_t)(((uint32_t)n_rows * 128u) >> 4);   // K-block stride in descriptor units
    const uint64_t a0 = A_SMEM ? make_desc_sw128(a_base) : 0;
    if (elect_one()) {
#pragma unroll
      for (int j = 0; j < KSTEPS; ++j) {
        const uint64_t bdesc = b0 + (uint64_t)(j >> 2) * kb + 2 * (j & 3);
        if (A_SMEM) umma_ss(d, a0 + 2 * j, bdesc, idesc, j > 0);
        else umma_ts(d, a_base + 8 * j, bdesc, idesc, j > 0);
      }
      // bias as one more K step: [1 1 1 0 ..] x [hi lo lo2 0 ..]^T accumulates bias[n] (exact: three bf16 terms) into every row
      if (cx.bias_smem) umma_ss(d, make_desc_nosw(cx.ones_smem), make_desc_nosw(cx.bias_smem + (uint32_t)chunk * kBiasBlock), idesc, true);
      umma_commit(&cx.full[slot]);
      if (cx.extra_commit) umma_commit(cx.extra_commit);
    }
    __syncwarp();
    trace_mark(cx, 41);
  }
  ++cx.n_issued;
}

// Wait for the oldest outstanding chunk; returns the lane-adjusted TMEM address of this thread's
// 32-column half of its accumulator slot.
__device__ __forceinline__ uint32_t wait_chunk(WgCtx& cx) {
  const uint32_t slot = cx.n_waited & 1, parity = (cx.n_waited >> 1) & 1;
  mbar_wait_or_trap(&cx.full[slot], parity);
  ++cx.n_waited;
  tc_fence_after();
  return cx.lane_addr + kColD + slot * 64 + cx.colhalf * 32;
}

// One MLP layer = NC output chunks of 64, in two halves so that independent work (a gather) can be
// placed between the first MMA issue and the first wait.
// Preconditions of layer_begin: the A operand is complete and a WG barrier has been passed since it
// was written and since both accumulator slots were last read.
template <int NC, int KSTEPS, bool A_SMEM, bool ISSUER, class ChunkOf>
__device__ __forceinline__ void layer_begin(WgCtx& cx, uint32_t a_base, uint32_t w_smem, int n_rows, ChunkOf chunk_of) {
  issue_chunk<KSTEPS, A_SMEM, ISSUER>(cx, a_base, w_smem, n_rows, chunk_of(0));
  if (NC > 1) issue_chunk<KSTEPS, A_SMEM, ISSUER>(cx, a_base, w_smem, n_rows, chunk_of(1));
}
template <int NC, int KSTEPS, bool A_SMEM, bool ISSUER, class ChunkOf, class Epi>
__device__ __forceinline__ void layer_finish(WgCtx& cx, uint32_t a_base, uint32_t w_smem, int n_rows, ChunkOf chunk_of, Epi epi) {
  // The accumulator of chunk i+1 is fetched (tcgen05.ld, asynchronous) as soon as chunk i's registers are dead, so its
  // TMEM round trip and mbarrier wake-up hide behind chunk i's store / fence / barrier / MMA issue.
  if constexpr (ISSUER) {
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      step_done<true>(cx);   // wait until every epilogue warp has drained the slot and written its part of A'
      if (i + 2 < NC) issue_chunk<KSTEPS, A_SMEM, true>(cx, a_base, w_smem, n_rows, chunk_of(i + 2));
    }
  } else {
    uint32_t v[32];
    tmem_ld32(wait_chunk(cx), v);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      tmem_ld_wait();
      trace_mark(cx, 10 + i);
      // Chunk i+2 reuses this chunk's accumulator slot and reads the SAME A operand, so all the issuer needs from us is
      // "slot drained": arrive as soon as the accumulator is in registers, not after the epilogue.  (The epilogue's own
      // prefetch waits for chunk i+1 -- arriving after it meant chunk i+2 could not even be issued before chunk i+1 had
      // completed, a tensor-pipe bubble per chunk wherever the epilogue is short: the Q-table stores of the composed
      // layer.)  The last two chunks arrive after their epilogue: those arrivals tell the issuer that the NEXT layer's A
      // operand is complete (every warp's earlier tcgen05.st precedes them in program order).
      const bool early = STIF_EARLY_RELEASE && i + 2 < NC;
      if (early) {
        tc_fence_before();
        step_done<false>(cx);
      }
      epi(i, v, [&]() { if (i + 1 < NC) tmem_ld32(wait_chunk(cx), v); });
      trace_mark(cx, 20 + i);
      if (!early) {
        tc_fence_before();
        step_done<false>(cx);   // arrive and move on: epilogue warps never wait for each other
      }
      if (i + 2 < NC) issue_chunk<KSTEPS, A_SMEM, false>(cx, a_base, w_smem, n_rows, chunk_of(i + 2));   // (counter only)
    }
  }
}
template <int NC, int KSTEPS, bool A_SMEM, bool ISSUER, class ChunkOf, class Epi>
__device__ __forceinline__ void run_layer(WgCtx& cx, uint32_t a_base, uint32_t w_smem, int n_rows, ChunkOf chunk_of, Epi epi) {
  layer_begin<NC, KSTEPS, A_SMEM, ISSUER>(cx, a_base, w_smem, n_rows, chunk_of);
  layer_finish<NC, KSTEPS, A_SMEM, ISSUER>(cx, a_base, w_smem, n_rows, chunk_of, epi);
}

// (No software prefetch: an L2 prefetch of the next tile's Q lines paid off while both workgroups gathered at once; with the
// gather turns it costs 2 % of K2, and exact per-tap prefetches cost 3 %.)

// which of a thread's 16 sine pairs go to the FMA-pipe polynomial (evenly interleaved with the MUFU ones)
__host__ __device__ constexpr bool use_poly(int j) { return ((j * kPolyPairs) % 16) < kPolyPairs; }

// rotation kernels: of the 8 sine pairs of a 16-column half-chunk, KPOLY16 run on the FMA pipe (poly_sin2), spread evenly
#ifndef KPOLY16
#define KPOLY16 0
#endif
__host__ __device__ constexpr bool use_poly16(int jp) { return ((jp * KPOLY16) % 8) < KPOLY16; }
__device__ __forceinline__ float2 sin2_mixed(float2 a, int jp) {
  if (use_poly16(jp)) return poly_sin2(a);
  return make_float2(fast_sin(a.x), fast_sin(a.y));
}
// 4 consecutive constants in one 128-bit constant-bank load (all bias / weight arrays are 16-byte aligned at 4-element steps)
__device__ __forceinline__ float4 ldc4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// ---- epilogues (thread = query row x 32 of the chunk's 64 accumulator columns) --------------------
// act = sin(acc + bias) -> bf16 -> TMEM A operand (16 columns at dst)
template <class Pf>
__device__ __forceinline__ void epi_sin_to_tmem(uint32_t (&v)[32], uint32_t dst, const float* __restrict__ bias, Pf&& next_ld) {
  uint32_t pk[16];
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 b4 = ldc4(bias + 4 * j4);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = 2 * j4 + h;
      const float2 a = add2(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])),
                            h ? make_float2(b4.z, b4.w) : make_float2(b4.x, b4.y));
      if (use_poly(j)) {
        const float2 sn = poly_sin2(a);
        pk[j] = pack_bf16x2(sn.x, sn.y);
      } else {
        pk[j] = pack_bf16x2(fast_sin(a.x), fast_sin(a.y));
      }
    }
  }
  next_ld();
  tmem_st16(dst, pk);
  tmem_st_wait();
}

// act = sin(acc + bias) kept in fp32 and contracted with the NOUT x 256 output layer on the FMA pipe.
// acc[k] holds (even-column, odd-column) partial sums.
template <int NOUT, class Pf>
__device__ __forceinline__ void epi_sin_fma(uint32_t (&v)[32], const float* __restrict__ bias, const float* __restrict__ w,
                                            float2 (&acc)[NOUT], Pf&& next_ld) {
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 b4 = ldc4(bias + 4 * j4);
    float4 w4[NOUT];
#pragma unroll
    for (int k = 0; k < NOUT; ++k) w4[k] = ldc4(w + k * 256 + 4 * j4);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = 2 * j4 + h;
      float2 s = add2(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])),
                      h ? make_float2(b4.z, b4.w) : make_float2(b4.x, b4.y));
      if (use_poly(j)) {
        s = poly_sin2(s);
      } else {
        s.x = fast_sin(s.x);
        s.y = fast_sin(s.y);
      }
#pragma unroll
      for (int k = 0; k < NOUT; ++k) acc[k] = fma2(s, h ? make_float2(w4[k].z, w4[k].w) : make_float2(w4[k].x, w4[k].y), acc[k]);
    }
  }
  next_ld();
}

// two fp32 -> packed fp16x2, round-to-nearest-even, SATURATING to +-65504 (and NaN -> 0x7FFF stays NaN only for NaN
// inputs): a projected-table entry beyond fp16's range -- first-layer pre-activations of ~6e4 rad, far outside anything a
// SIREN produces -- must not become an Inf whose sine is a NaN for every query that gathers it
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// acc + bias -> fp16 -> 64 bytes of the projected HR table
// (ADD: `add` points at 32 fp16 values added to the row -- the upsampled-frame term of decoding_test at x4, which lives on
// the same grid as the Q table and is sampled with the same taps, so it can ride inside it)
template <bool ADD = false, class Pf>
__device__ __forceinline__ void epi_store_qtab(uint32_t (&v)[32], const float* __restrict__ bias, __half* dst, bool valid, Pf&& next_ld,
                                               const __half* add = nullptr) {
  uint32_t o[16];
  U8x32 u0, u1;
  if constexpr (ADD) { u0 = ldg256(add); u1 = ldg256(add + 16); }
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 b4 = ldc4(bias + 4 * j4);
    float2 a0 = add2(make_float2(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1])), make_float2(b4.x, b4.y));
    float2 a1 = add2(make_float2(__uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3])), make_float2(b4.z, b4.w));
    if constexpr (ADD) {
      const uint32_t h0 = j4 < 4 ? u0.r[2 * j4] : u1.r[2 * j4 - 8], h1 = j4 < 4 ? u0.r[2 * j4 + 1] : u1.r[2 * j4 - 7];
      a0.x = add_f16((uint16_t)(h0 & 0xFFFF), a0.x); a0.y = add_f16((uint16_t)(h0 >> 16), a0.y);
      a1.x = add_f16((uint16_t)(h1 & 0xFFFF), a1.x); a1.y = add_f16((uint16_t)(h1 >> 16), a1.y);
    }
    o[2 * j4] = pack_half2(a0.x, a0.y);
    o[2 * j4 + 1] = pack_half2(a1.x, a1.y);
  }
  next_ld();
  if (!valid || (STIF_DIAG & 1)) return;
  stg256(dst, o);            // two full 32-byte sectors per thread
  stg256(dst + 16, o + 8);
}

// local-ensemble stage A: F + bias -> fp32 table row (this thread's 32 channels = one 128-byte line)
template <class Pf>
__device__ __forceinline__ void epi_store_ftab(uint32_t (&v)[32], const float* __restrict__ bias, float* dst, bool valid, Pf&& next_ld) {
  uint32_t o[32];
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 b4 = ldc4(bias + 4 * j4);
    const float2 a0 = add2(make_float2(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1])), make_float2(b4.x, b4.y));
    const float2 a1 = add2(make_float2(__uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3])), make_float2(b4.z, b4.w));
    o[4 * j4] = __float_as_uint(a0.x); o[4 * j4 + 1] = __float_as_uint(a0.y);
    o[4 * j4 + 2] = __float_as_uint(a1.x); o[4 * j4 + 3] = __float_as_uint(a1.y);
  }
  next_ld();
  if (!valid) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) stg256(dst + 8 * j, o + 8 * j);
}

// f0 = sin(F + g) -> bf16 -> TMEM (g already holds bilinear(TB) + time constant + composed bias)
template <class Pf>
__device__ __forceinline__ void epi_flow_first_layer(uint32_t (&v)[32], uint32_t dst, const float (&g)[32], Pf&& next_ld) {
  uint32_t pk[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float2 a = add2(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), make_float2(g[2 * j], g[2 * j + 1]));
    if (use_poly(j)) {
      const float2 sn = poly_sin2(a);
      pk[j] = pack_bf16x2(sn.x, sn.y);
    } else {
      pk[j] = pack_bf16x2(fast_sin(a.x), fast_sin(a.y));
    }
  }
  next_ld();
  tmem_st16(dst, pk);
  tmem_st_wait();
}


// ---- half-chunk pipelined epilogues (rotation kernels) -----------------------------------------------------------------
// A thread's 32 accumulator columns are fetched and processed as two 16-column halves A | B: while the sines of A run,
// tcgen05.ld of B is in flight; while the sines of B run, the NEXT chunk's A half is in flight and A's tcgen05.st drains.
// A workgroup that has the MUFU pipe to itself (the other one waiting for MMAs, gathering or storing) then loses ~80
// instead of ~300 clocks per chunk to TMEM round trips.  Each tcgen05.wait::ld has exactly one load outstanding.
//   epi(i, h, v16): process columns [16 h, 16 h + 16) of this thread's half of chunk i (v16 = the 16 accumulators).
template <int NC, int KSTEPS, bool A_SMEM, bool ISSUER, class ChunkOf, class Epi>
__device__ __forceinline__ void layer_finish2(WgCtx& cx, uint32_t a_base, uint32_t w_smem, int n_rows, ChunkOf chunk_of, Epi epi) {
  if constexpr (ISSUER) {
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      step_done<true>(cx);
      if (i + 2 < NC) issue_chunk<KSTEPS, A_SMEM, true>(cx, a_base, w_smem, n_rows, chunk_of(i + 2));
    }
  } else {
    uint32_t va[16], vb[16];
    uint32_t addr = wait_chunk(cx);
    tmem_ld16(addr, va);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      tmem_ld_wait();                       // A(i)
      tmem_ld16(addr + 16, vb);             // B(i) in flight under A's sines
      trace_mark(cx, 10 + i);
      epi(i, 0, va);
      tmem_ld_wait();                       // B(i): the accumulator slot is drained
      const bool early = STIF_EARLY_RELEASE && i + 2 < NC;
      if (early) {
        tc_fence_before();
        step_done<false>(cx);
      }
      if (i + 1 < NC) {                     // A(i + 1) in flight under B's sines
        addr = wait_chunk(cx);
        tmem_ld16(addr, va);
      }
      epi(i, 1, vb);
      trace_mark(cx, 20 + i);
      if (!early) {
        tmem_st_wait();                     // covers every tcgen05.st of this layer so far (the next layer's A operand)
        tc_fence_before();
        step_done<false>(cx);
      }
      if (i + 2 < NC) issue_chunk<KSTEPS, A_SMEM, false>(cx, a_base, w_smem, n_rows, chunk_of(i + 2));   // (counter only)
    }
  }
}
template <int NC, int KSTEPS, bool A_SMEM, bool ISSUER, class ChunkOf, class Epi>
__device__ __forceinline__ void run_layer2(WgCtx& cx, uint32_t a_base, uint32_t w_smem, int n_rows, ChunkOf chunk_of, Epi epi) {
  layer_begin<NC, KSTEPS, A_SMEM, ISSUER>(cx, a_base, w_smem, n_rows, chunk_of);
  layer_finish2<NC, KSTEPS, A_SMEM, ISSUER>(cx, a_base, w_smem, n_rows, chunk_of, epi);
}

// act = sin(acc + bias) -> bf16 -> 8 TMEM columns at dst (no wait: layer_finish2 waits once per late chunk)
template <bool BIAS = true>
__device__ __forceinline__ void epi16_sin_to_tmem(const uint32_t (&v)[16], uint32_t dst, const float* __restrict__ bias) {
  uint32_t pk[8];
#ifdef STIF_DIAG_SKELETON   // diagnostic (WRONG results): the MMA / TMEM / barrier protocol alone
#pragma unroll
  for (int j = 0; j < 8; ++j) pk[j] = v[2 * j] & 0x3f803f80u;
  tmem_st8(dst, pk);
  return;
#endif
  if constexpr (!BIAS) {   // the bias arrived through the MMAs (issue_chunk, bias_smem)
#pragma unroll
    for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(fast_sin(__uint_as_float(v[2 * j])), fast_sin(__uint_as_float(v[2 * j + 1])));
    tmem_st8(dst, pk);
    return;
  }
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) {
    const float4 b4 = ldc4(bias + 4 * j4);
    const float2 a0 = add2(make_float2(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1])), make_float2(b4.x, b4.y));
    const float2 a1 = add2(make_float2(__uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3])), make_float2(b4.z, b4.w));
    const float2 s0 = sin2_mixed(a0, 2 * j4), s1 = sin2_mixed(a1, 2 * j4 + 1);
    pk[2 * j4] = pack_bf16x2(s0.x, s0.y);
    pk[2 * j4 + 1] = pack_bf16x2(s1.x, s1.y);
  }
  tmem_st8(dst, pk);
}
// act = sin(acc + bias) in fp32, contracted with the NOUT x 256 output layer on the FMA pipe
template <int NOUT, bool BIAS = true>
__device__ __forceinline__ void epi16_sin_fma(const uint32_t (&v)[16], const float* __restrict__ bias, const float* __restrict__ w,
                                              float2 (&acc)[NOUT]) {
#ifdef STIF_DIAG_SKELETON
  acc[0].x += __uint_as_float(v[0] & 0x3f800000u);
  return;
#endif
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) {
    const float4 b4 = BIAS ? ldc4(bias + 4 * j4) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 w4[NOUT];
#pragma unroll
    for (int k = 0; k < NOUT; ++k) w4[k] = ldc4(w + k * 256 + 4 * j4);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float2 s = make_float2(__uint_as_float(v[4 * j4 + 2 * h]), __uint_as_float(v[4 * j4 + 2 * h + 1]));
      if constexpr (BIAS) s = add2(s, h ? make_float2(b4.z, b4.w) : make_float2(b4.x, b4.y));
      s = sin2_mixed(s, 2 * j4 + h);
#pragma unroll
      for (int k = 0; k < NOUT; ++k) acc[k] = fma2(s, h ? make_float2(w4[k].z, w4[k].w) : make_float2(w4[k].x, w4[k].y), acc[k]);
    }
  }
}
// acc + bias -> fp16 -> 32 bytes of the projected HR table (one full sector)
template <bool BIAS = true>
__device__ __forceinline__ void epi16_store_qtab(const uint32_t (&v)[16], const float* __restrict__ bias, __half* dst, bool valid) {
  uint32_t o[8];
#ifdef STIF_DIAG_SKELETON
  if (v[0] == 0x12345678u) stg256(dst, v);
  return;
#endif
  if constexpr (!BIAS) {
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = pack_half2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
    if (valid) stg256(dst, o);
    return;
  }
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) {
    const float4 b4 = ldc4(bias + 4 * j4);
    const float2 a0 = add2(make_float2(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1])), make_float2(b4.x, b4.y));
    const float2 a1 = add2(make_float2(__uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3])), make_float2(b4.z, b4.w));
    o[2 * j4] = pack_half2(a0.x, a0.y);
    o[2 * j4 + 1] = pack_half2(a1.x, a1.y);
  }
  if (valid) stg256(dst, o);
}

// ---- common prologue / epilogue of the kernels -----------------------------------------------------
struct CtaSetup {
  uint64_t* bars;  // [0] weights landed, [1,2] accumulator ring of TMEM slot 0, [3,4] of slot 1, [5,6] slot free, [7,8] h2 read (rotation kernels)
  uint32_t tmem_base;
};

// Programmatic dependent launch: every kernel lets its successor in the stream be scheduled at once (its CTAs take SMs
// as ours retire and run their prologue -- barrier init, TMEM allocation, the weight image's TMA load -- early), and
// itself waits for its predecessor's results only after its own prologue.  All grids are persistent (<= one CTA per
// SM, all resident from the start), so an early successor can never starve its predecessor of SMs.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_for_predecessor() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ CtaSetup cta_prologue(uint32_t bars_off, uint32_t w_off, const uint8_t* wimg, uint32_t wbytes,
                                                 uint32_t tmem_cols) {
  CtaSetup s;
  pdl_launch_dependents();
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // SW128 operands assume a 1024-byte aligned window
  s.bars = reinterpret_cast<uint64_t*>(smem + bars_off);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + bars_off + 120);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i <= 4; ++i) mbar_init(&s.bars[i], 1);
    for (int i = 5; i <= 6; ++i) mbar_init(&s.bars[i], 1);   // rotation kernels: "TMEM slot 0 / 1 is free again" (its issuer arrives)
    for (int i = 7; i <= 8; ++i) mbar_init(&s.bars[i], 3);   // K1 rotation: "slot's composed layer has read h2" (3 tcgen05.commit)
    fence_mbar_init();
  }
  if (tid < 32) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    mbar_arrive_expect_tx(&s.bars[0], wbytes);
    for (uint32_t off = 0; off < wbytes; off += 32768) {
      const uint32_t n = min(32768u, wbytes - off);
      bulk_copy_g2s(smem + w_off + off, wimg + off, n, &s.bars[0]);
    }
  }
  s.tmem_base = *tmem_slot;
  pdl_wait_for_predecessor();   // nothing above reads or writes data another kernel produces
  return s;
}

__device__ __forceinline__ void cta_epilogue(uint32_t tmem_base, uint32_t tmem_cols) {
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem_base, tmem_cols);
}

// EW0 = index of the first epilogue warp: 2 in the 18-warp kernels (warps 0, 1 issue), 4 in the producer kernels (warps 0, 1
// issue, 2, 3 idle so that the epilogue warps start on a warpgroup boundary, 4..19 epilogue, 20..23 produce)
template <int EW0 = 2>
__device__ __forceinline__ WgCtx make_wg(const CtaSetup& s) {
  WgCtx cx;
  const int tid = threadIdx.x;
  // warp-uniform quantities are broadcast from lane 0 so that ptxas keeps them in uniform registers
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  // Warps 0 and 1 are the MMA issuers, warps 2..17 the epilogue warps.  The issuers must be the OLDEST warps of their
  // scheduler: with the issuers last (warps 16, 17) the greedy-then-oldest pick starved them behind four always-ready
  // epilogue warps -- ~2400 clocks between "accumulator slot free" and the next chunk's first tcgen05.mma.
  cx.issuer = warp < 2;
  const int ew = warp - EW0;                 // epilogue warp 0..15
  cx.wg = cx.issuer ? warp : ew >> 3;
  const int warp_in_wg = ew & 7;
  cx.slot = cx.issuer ? 16 + warp : ew;
  cx.tid_wg = warp_in_wg * 32 + (tid & 31);
  cx.warp_in_wg = warp_in_wg;
  const int quarter = warp & 3;              // the TMEM lane quarter this warp may access = (warp id) % 4
  cx.colhalf = warp_in_wg >> 2;              // epilogue warps w and w+4 of a WG share a quarter and split the columns
  cx.row = quarter * 32 + (tid & 31);
  cx.tmem = __shfl_sync(0xffffffffu, s.tmem_base, 0) + (uint32_t)cx.wg * 256u;
  cx.lane_addr = cx.tmem + ((uint32_t)(quarter * 32) << 16);
  cx.full = s.bars + 1 + 2 * cx.wg;
  cx.bar_base = 1 + 2 * cx.wg;
  cx.quarter = quarter;
  cx.n_issued = cx.n_waited = cx.n_steps = 0;
  cx.trace = nullptr;
  cx.extra_commit = nullptr;
  cx.bias_smem = 0;
  cx.ones_smem = 0;
  return cx;
}

// =================================================================================================
// K0: latent projection  tab[HW,256] (fp16) = [latent(192) ; frames(6)]^T  W_tab^T      (bf16 MMA)
// =================================================================================================
// 512 threads: thread = (texel row, quarter of the channel groups).  The fp32 planes of the NEXT tile are fetched into
// registers (up to 54 loads per thread = the whole 101 KB tile in flight per SM) before the current tile's MMA is
// awaited and its accumulator is written out, so HBM latency hides behind the epilogue instead of serialising with it.
__global__ void __launch_bounds__(512, 1) k0_project_kernel(const __grid_constant__ K0Params p) {
  const CtaSetup s = cta_prologue(k0Bars, k0B, p.wimg, k0WBytes, 256);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row = tid & 127, part = tid >> 7;
  const long ntiles = (p.m_end - p.m_begin + kTile - 1) / kTile;
  const uint32_t a_sm = smem_u32(smem + k0A), b_sm = smem_u32(smem + k0B);
  // K index 200..207 (group 25) is padding: zero it once (B is zero there too, but 0 * garbage could be NaN)
  if (part == 1) *reinterpret_cast<uint4*>(smem + k0A + 3 * 16384 + sw128_offset(row, 8)) = make_uint4(0, 0, 0, 0);
  // channel groups of 8: this thread owns g8 = part + 4 i, i = 0..5 (latent) and, for part 0, group 24 (the 6 frame planes)
  float v[48], f[6];
  auto fetch = [&](long tile) {
    const long m = min(p.m_begin + tile * kTile + row, p.m_end - 1);
    if (p.latent16) {   // (uniform: one branch per tile, not per element)
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) v[8 * i + e] = __uint_as_float((uint32_t)__ldg(p.latent16 + (long)((part + 4 * i) * 8 + e) * p.HW + m) << 16);
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) v[8 * i + e] = __ldg(p.latent + (long)((part + 4 * i) * 8 + e) * p.HW + m);
    }
    if (part == 0) {
#pragma unroll
      for (int e = 0; e < 6; ++e) f[e] = __ldg(p.frames + (long)e * p.HW + m);
    }
  };
  if ((long)blockIdx.x < ntiles) fetch(blockIdx.x);
  mbar_wait_or_trap(&s.bars[0], 0);
  uint32_t phase = 0;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    // ---- A tile: 8 channels -> one 16-byte chunk of the SW128 image
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int g8 = part + 4 * i;
      *reinterpret_cast<uint4*>(smem + k0A + (g8 >> 3) * 16384 + sw128_offset(row, (g8 & 7) * 8)) =
          make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]), pack_bf16x2(v[8 * i + 4], v[8 * i + 5]),
                     pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
    }
    if (part == 0)
      *reinterpret_cast<uint4*>(smem + k0A + 3 * 16384 + sw128_offset(row, 0)) =
          make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), 0u);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        const uint32_t idesc = make_idesc_bf16(128, 256);
#pragma unroll
        for (int j = 0; j < 13; ++j)    // K = 208 = 13 x 16
          umma_ss(s.tmem_base, make_desc_sw128(a_sm + (j >> 2) * 16384) + 2 * (j & 3), make_desc_sw128(b_sm + (j >> 2) * 32768) + 2 * (j & 3),
                  idesc, j > 0);
        umma_commit(&s.bars[1]);
      }
      __syncwarp();
    }
    if (tile + gridDim.x < ntiles) fetch(tile + gridDim.x);   // in flight across the MMA wait and the epilogue
    mbar_wait_or_trap(&s.bars[1], phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue: warp = (lane quarter, column quarter); thread = one texel row, 64 channels = one 128-byte line
    {
      const int quarter = warp & 3, cq = warp >> 2;
      const long texel = p.m_begin + tile * kTile + quarter * 32 + lane;
      const uint32_t src = s.tmem_base + ((uint32_t)(quarter * 32) << 16) + cq * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t a[32], o[16];
        tmem_ld32(src + c * 32, a);
        tmem_ld_wait();
#pragma unroll
        // clamped to +-16000 rad: the sum of the four tables K2 blends in fp16 then stays finite whatever the latent holds
        // (pre-activations of that size carry no information anyway: a bf16 operand's step there is 64 rad)
        for (int j = 0; j < 16; ++j)
          o[j] = pack_half2(fminf(fmaxf(__uint_as_float(a[2 * j]), -16000.f), 16000.f), fminf(fmaxf(__uint_as_float(a[2 * j + 1]), -16000.f), 16000.f));
        if (texel < p.m_end) {
          __half* dst = p.tab + texel * 256 + cq * 64 + c * 32;
          stg256(dst, o);
          stg256(dst + 16, o + 8);
        }
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  cta_epilogue(s.tmem_base, 256);
}

// =================================================================================================
// K1: stage A + B
// =================================================================================================
// CH = this thread's column half (compile-time so that every bias / weight index is an immediate
// constant-bank operand instead of a per-thread LDC)
// UPF = decoding_test at x4 ("upsampled frames"): the bilinear frame gathers of stage B / stage D read the x4-upsampled
// pair, whose grid IS the query grid at x4 -- stage B's term is the query's own texel (added to gB) and stage D's terms
// are folded into the Q planes (bilinear(Q;g) + bilinear(UE;g) = bilinear(Q + UE;g)), so K2 runs unchanged.
template <bool ISSUER, bool UPF = false>
__device__ __forceinline__ void k1_tile_loop(const K1Params& p, const CtaSetup& s, WgCtx& cx) {
  const int CH = cx.colhalf;   // warp-uniform (broadcast from lane 0): constant-bank indices stay uniform-register loads
  const uint32_t wsm = smem_u32(smem);
  const Geometry& g = p.g;
  const long ntiles = (p.q_end - p.q_begin + kTile - 1) / kTile;
  const uint4* __restrict__ tab4 = reinterpret_cast<const uint4*>(p.tab);  // 32 uint4 per texel
  float4* part = reinterpret_cast<float4*>(smem + k1Part) + cx.wg * 128;
  const float* cs = reinterpret_cast<const float*>(smem + k1Const);   // shared-memory copy of the fp32 output layer
  const int ch0 = CH * 32;   // this thread's 32 channels of every 64-wide vector

  const long tile_first = (long)blockIdx.x * 2 + cx.wg;
  // (jy, jx) of this thread's query advance by a constant step from tile to tile: one 64-bit division up front
  // instead of one per tile (and one more in the last, partial tile, whose missing rows are clamped onto the last query)
  const long q_step = (long)gridDim.x * 2 * kTile;
  const int jy_step = (int)(q_step / g.WW), jx_step = (int)(q_step - (long)jy_step * g.WW);
  long q = p.q_begin + tile_first * kTile + cx.row;
  int jy_run = (int)(q / g.WW), jx_run = (int)(q - (long)jy_run * g.WW);
  auto next_tile = [&]() {
    q += q_step;
    jy_run += jy_step;
    jx_run += jx_step;
    if (jx_run >= g.WW) { jx_run -= g.WW; ++jy_run; }
  };
  for (long tile = tile_first; tile < ntiles; tile += (long)gridDim.x * 2, next_tile()) {
    if (cx.trace && tile >= (long)gridDim.x * 2 * 16) cx.trace = nullptr;   // trace the first 16 tiles only
    const bool valid = q < p.q_end;
    long qc = q;
    int jy = jy_run, jx = jx_run;
    if (!valid) {
      qc = p.q_end - 1;
      jy = (int)(qc / g.WW);
      jx = (int)(qc - (long)jy * g.WW);
    }
    SineTurn<7, 8> turn{cx.wg, tile == tile_first, tile + 1 < ntiles, tile - 1 + (long)gridDim.x * 2 < ntiles};
    trace_mark(cx, 1);
    // ---- stage A, first layer (hoisted): h0 = sin(TA[iy,ix] + rel . w_rel + cA)      (:382-400)
    if constexpr (!ISSUER) {
      const int iy = g.y.idx[jy], ix = g.x.idx[jx];
      const float rely = g.y.rel[jy], relx = g.x.rel[jx];
      const bool inb = (iy >= 0) & (iy < g.H) & (ix >= 0) & (ix < g.W);
      const uint4* ta = tab4 + (inb ? ((long)iy * g.W + ix) : 0) * 32 + CH * 4;
      STIF_BOUND(jy, g.HH); STIF_BOUND(jx, g.WW); STIF_BOUND(qc, (long)g.HH * g.WW);
      uint32_t pk[16];
      U8x32 ta2;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if ((j & 1) == 0) { if (STIF_DIAG & 2) ta2 = U8x32{}; else ta2 = ldg256(ta + j); }
        uint32_t w4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) w4[e] = inb ? ta2.r[(j & 1) * 4 + e] : 0u;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = ch0 + j * 8 + e * 2;
          const float b0 = fmaf(rely, p.c.a_rel[2 * c], fmaf(relx, p.c.a_rel[2 * c + 1], p.c.cA[c]));
          const float b1 = fmaf(rely, p.c.a_rel[2 * c + 2], fmaf(relx, p.c.a_rel[2 * c + 3], p.c.cA[c + 1]));
          pk[j * 4 + e] = pack_bf16x2(fast_sin(add_f16((uint16_t)(w4[e] & 0xFFFF), b0)), fast_sin(add_f16((uint16_t)(w4[e] >> 16), b1)));
        }
      }
      tmem_st16(cx.lane_addr + kColAin + CH * 16, pk);
      tmem_st_wait();
      tc_fence_before();
    }
    trace_mark(cx, 2);
    step_done<ISSUER>(cx);
    trace_mark(cx, 3);

    // ---- feat_imnet hidden layers
    run_layer<1, 4, false, ISSUER>(cx, cx.tmem + kColAin, wsm + k1F1, 64, [](int) { return 0; },
                 [&](int, uint32_t(&v)[32], auto&& pf) { epi_sin_to_tmem(v, cx.lane_addr + kColAin + CH * 16, p.c.f1_b + ch0, pf); });
    run_layer<4, 4, false, ISSUER>(cx, cx.tmem + kColAin, wsm + k1F2, 256, [](int i) { return i; }, [&](int i, uint32_t(&v)[32], auto&& pf) {
      if (STIF_K1_SINE_TURNS && i == 0) turn.template acquire<2>(0);
      if (i == 0) trace_mark(cx, 50);
      epi_sin_to_tmem(v, cx.lane_addr + kColA + 32 * i + CH * 16, p.c.f2_b + 64 * i + ch0, pf);
      if (STIF_K1_SINE_TURNS && i == 3) turn.template release<2>(0);
      if (i == 3) trace_mark(cx, 51);
    });

    // ---- composed last layer of feat_imnet: chunk order Q1, Q2, F (F last: its epilogue overwrites h2)
    auto f3_order = [](int i) { return i == 2 ? 0 : i + 1; };
    layer_begin<3, 16, false, ISSUER>(cx, cx.tmem + kColA, wsm + k1F3, 192, f3_order);
    // ---- stage B gather, placed after the composed layer's first MMAs are issued so its latency hides behind them:
    //      gB = bilinear(TB; query position) + cB + composed bias of F                  (:410-418)
    float gB[32];
    if constexpr (!ISSUER) {
      const Taps tp = make_taps_tables(g, jy, jx);
      uint16_t wq[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { wq[k] = __half_as_ushort(__float2half_rn(tp.w[k])); STIF_BOUND(tp.off[k], (long)g.H * g.W); }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
#pragma unroll
        for (int e = 0; e < 16; ++e) gB[16 * j + e] = p.c.cB[ch0 + 16 * j + e] + p.c.f3_b[ch0 + 16 * j + e];
        if constexpr (UPF) {   // + UB at the query's own texel of the upsampled frames
          const U8x32 ub = ldg256(p.utab + qc * 192 + ch0 + 16 * j);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            gB[16 * j + 2 * e] = add_f16((uint16_t)(ub.r[e] & 0xFFFF), gB[16 * j + 2 * e]);
            gB[16 * j + 2 * e + 1] = add_f16((uint16_t)(ub.r[e] >> 16), gB[16 * j + 2 * e + 1]);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const U8x32 v = (STIF_DIAG & 4) ? U8x32{} : ldg256(tab4 + (long)tp.off[k] * 32 + 8 + CH * 4 + 2 * j);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            gB[16 * j + 2 * e] = fma_f16((uint16_t)(v.r[e] & 0xFFFF), wq[k], gB[16 * j + 2 * e]);
            gB[16 * j + 2 * e + 1] = fma_f16((uint16_t)(v.r[e] >> 16), wq[k], gB[16 * j + 2 * e + 1]);
          }
        }
      }
    }

    trace_mark(cx, 4);
    layer_finish<3, 16, false, ISSUER>(cx, cx.tmem + kColA, wsm + k1F3, 192, f3_order, [&](int i, uint32_t(&v)[32], auto&& pf) {
      if (i < 2) {
        if constexpr (UPF) epi_store_qtab<true>(v, p.c.f3_b + 64 * (i + 1) + ch0, p.qtab + qc * 128 + 64 * i + ch0, valid, pf,
                                                p.utab + qc * 192 + 64 * (i + 1) + ch0);
        else epi_store_qtab(v, p.c.f3_b + 64 * (i + 1) + ch0, p.qtab + qc * 128 + 64 * i + ch0, valid, pf);
      } else {
        epi_flow_first_layer(v, cx.lane_addr + kColAin + CH * 16, gB, pf);
      }
    });

    // ---- flow_imnet hidden layers; the 256->4 output layer rides the FMA pipe          (:419-422)
    run_layer<1, 4, false, ISSUER>(cx, cx.tmem + kColAin, wsm + k1L1, 64, [](int) { return 0; },
                 [&](int, uint32_t(&v)[32], auto&& pf) { epi_sin_to_tmem(v, cx.lane_addr + kColAin + CH * 16, p.c.l1_b + ch0, pf); });
    float2 fl[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
    run_layer<4, 4, false, ISSUER>(cx, cx.tmem + kColAin, wsm + k1L2, 256, [](int i) { return i; }, [&](int i, uint32_t(&v)[32], auto&& pf) {
      if (STIF_K1_SINE_TURNS && i == 0) turn.template acquire<2>(1);
      if (i == 0) trace_mark(cx, 52);
      epi_sin_fma<4>(v, p.c.l2_b + 64 * i + ch0, cs + kc1L3W + 64 * i + ch0, fl, pf);
      if (STIF_K1_SINE_TURNS && i == 3) turn.template release<2>(1);
      if (i == 3) trace_mark(cx, 53);
    });
    // combine the two column halves and store
    if constexpr (ISSUER) continue;
    const float4 mine = make_float4(fl[0].x + fl[0].y, fl[1].x + fl[1].y, fl[2].x + fl[2].y, fl[3].x + fl[3].y);
    if (CH == 1) part[cx.row] = mine;
    wg_barrier(cx.wg);
    if (CH == 0 && valid) {
      const float4 o = part[cx.row];
      STIF_BOUND(q, (long)g.HH * g.WW);
      reinterpret_cast<float4*>(p.flow)[q] =
          make_float4(mine.x + o.x + p.c.l3_b[0], mine.y + o.y + p.c.l3_b[1], mine.z + o.z + p.c.l3_b[2], mine.w + o.w + p.c.l3_b[3]);
    }
  }
}

// decoding_test at x4 on the tensor-core path: the fused loop with the upsampled-frame terms (see k1_tile_loop, UPF)
__global__ void __launch_bounds__(576, 1) k1_stage_ab_upf_kernel(const __grid_constant__ K1Params p) {
  const CtaSetup s = cta_prologue(k1Bars, 0, p.wimg, k1WBytes, 512);
  {
    float* cs = reinterpret_cast<float*>(smem + k1Const);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) cs[kc1L3W + i] = p.c.l3_w[i];
    __syncthreads();
  }
  WgCtx cx = make_wg(s);
  mbar_wait_or_trap(&s.bars[0], 0);
  if (cx.issuer) k1_tile_loop<true, true>(p, s, cx);
  else k1_tile_loop<false, true>(p, s, cx);
  cta_epilogue(s.tmem_base, 512);
}

// fp16 table of the upsampled-frame terms for that kernel: utab[4H*4W, 192] = w_up . bilinear_upsample_x4(frames)
// (ATen's formula, as project_frames_up4_kernel in kernels_fp32.cu; fp32 arithmetic, fp16 storage like the other tables)
__global__ void project_frames_up4_half_kernel(const float* __restrict__ frames, int H, int W, const float* __restrict__ w_up,
                                               __half* __restrict__ utab) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)16 * H * W * 192) return;
  const int c = (int)(i % 192);
  const long texel = i / 192;
  const int y = (int)(texel / (4 * W)), x = (int)(texel % (4 * W));
  const float sy = fmaxf(0.f, __fadd_rn(__fmul_rn(0.25f, (float)y + 0.5f), -0.5f)), sx = fmaxf(0.f, __fadd_rn(__fmul_rn(0.25f, (float)x + 0.5f), -0.5f));
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
  const float ly = __fadd_rn(sy, -(float)y0), lx = __fadd_rn(sx, -(float)x0);
  const float wy0 = __fadd_rn(1.f, -ly), wx0 = __fadd_rn(1.f, -lx);
  float s = 0.f;
#pragma unroll
  for (int ch = 0; ch < 6; ++ch) {
    const float* f = frames + (long)ch * H * W;
    const float top = __fadd_rn(__fmul_rn(wx0, f[(long)y0 * W + x0]), __fmul_rn(lx, f[(long)y0 * W + x1]));
    const float bot = __fadd_rn(__fmul_rn(wx0, f[(long)y1 * W + x0]), __fmul_rn(lx, f[(long)y1 * W + x1]));
    s = fmaf(w_up[c * 6 + ch], __fadd_rn(__fmul_rn(wy0, top), __fmul_rn(ly, bot)), s);
  }
  utab[i] = __float2half_rn(s);
}

__global__ void __launch_bounds__(576, 1) k1_stage_ab_kernel(const __grid_constant__ K1Params p) {
  const CtaSetup s = cta_prologue(k1Bars, 0, p.wimg, k1WBytes, 512);
  {
    float* cs = reinterpret_cast<float*>(smem + k1Const);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) cs[kc1L3W + i] = p.c.l3_w[i];
    __syncthreads();
  }
  WgCtx cx = make_wg(s);
  if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0) cx.trace = p.trace + cx.slot * 4096;
  mbar_wait_or_trap(&s.bars[0], 0);
  if (cx.issuer) k1_tile_loop<true>(p, s, cx);
  else k1_tile_loop<false>(p, s, cx);
  cta_epilogue(s.tmem_base, 512);
}


// ---- K1 as a three-workgroup rotation over two TMEM slots (STIF_K1_ROT) -----------------------------------------------
// Same idea as K2's rotation (k2_rot_loop): tile n of the CTA belongs to workgroup n % 3 and runs its MMA phase on TMEM
// slot n % 2.  What K1 can do without its slot is the tile-opening first layer (index tables -> TA row -> 64 sines); its
// result h0 is parked in TMEM columns [0, 32) of the TARGET slot while the slot's previous tile is still in its flow
// layers: those columns (part of h2) are dead once the composed layer's MMAs have completed, which the slot's issuer
// reports on bars[7 + slot] (three commits, one per chunk of that layer).  The stage-B gather stays inside the MMA phase
// (its sum has to live somewhere until the composed layer's F chunk arrives: as fp16 pairs in 16 registers here, the
// kernel runs at 72 registers per thread).  13 accumulator chunks and 14 step barriers per tile.
constexpr uint32_t kColH0 = 0;
#ifndef STIF_K1_BIAS_MMA
// 1: the biases of the 13 MMA chunks ride the tensor pipe as an extra K = 16 step (constant ones tile x [hi lo lo2] bias block,
// interleaved no-swizzle operands; stif_selftest T6) instead of LDCU + MOV + FADD2 per accumulator pair.  Parity-green, 18 % fewer
// epilogue instructions -- and no faster (K1 0.562 vs 0.560 ms): each extra MMA costs its 32-clock floor on the serial
// issue -> MMA -> epilogue chain (+0.021 ms) and the shorter epilogues give back only 0.015 ms.  Off by default.
#define STIF_K1_BIAS_MMA 0
#endif
constexpr bool kK1BiasMma = STIF_K1_BIAS_MMA != 0;
// ... | ones tile | 13 bias blocks (F1, F2 x4, F3 x3, L1, L2 x4 in weight-row order)
constexpr uint32_t k1rOnes = k1WBytes, k1rBias = k1rOnes + kOnesTile, k1rPart = kK1BiasMma ? k1rBias + 13 * kBiasBlock : k1WBytes;
constexpr uint32_t k1rConst = k1rPart + 3 * 128 * 16, k1rBars = k1rConst + kc1Floats * 4, k1rSmem = k1rBars + 128;
constexpr uint32_t k1rBiasF1 = k1rBias, k1rBiasF2 = k1rBias + kBiasBlock, k1rBiasF3 = k1rBias + 5 * kBiasBlock, k1rBiasL1 = k1rBias + 8 * kBiasBlock,
                   k1rBiasL2 = k1rBias + 9 * kBiasBlock;
static_assert(k1rSmem <= 232448, "exceeds 227 KB of shared memory");

template <bool ISSUER, bool MULTI>
__device__ __forceinline__ void k1_rot_loop(const K1Params& p, const CtaSetup& s, WgCtx& cx, int me) {
  const int CH = cx.colhalf;
  const uint32_t wsm = smem_u32(smem);
  const Geometry& g = p.g;
  const long ntiles = MULTI ? p.ntiles_total : (p.q_end - p.q_begin + kTile - 1) / kTile;
  const long stride = gridDim.x;
  const uint4* __restrict__ tab4 = reinterpret_cast<const uint4*>(p.tab);
  float4* part = reinterpret_cast<float4*>(smem + k1rPart) + me * 128;
  const float* cs = reinterpret_cast<const float*>(smem + k1rConst);
  const int ch0 = CH * 32;
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, s.tmem_base, 0);
  uint64_t* slot_free = s.bars + 5;
  uint64_t* h2_read = s.bars + 7;
  constexpr int kStep = ISSUER ? 2 : 3;
#ifdef STIF_DIAG_BIAS   // diagnostics (wrong results): 1 = bias MMAs AND epilogue adds, 2 = neither
  constexpr bool NB = STIF_DIAG_BIAS == 1;
#else
  constexpr bool NB = !kK1BiasMma;   // epilogues add the bias themselves
#endif
  if constexpr (ISSUER) cx.ones_smem = wsm + k1rOnes;
#if defined(STIF_DIAG_BIAS) && STIF_DIAG_BIAS == 2
  auto bias_blocks = [&](uint32_t) {};
#else
  auto bias_blocks = [&](uint32_t off) { if constexpr (ISSUER && kK1BiasMma) cx.bias_smem = wsm + off; };
#endif
  for (long n = me;; n += kStep) {
    const long tile = (long)blockIdx.x + n * stride;
    if (tile >= ntiles) break;
    const int slot = (int)(n & 1);
    const uint32_t seq = (uint32_t)(n >> 1);
    cx.tmem = tmem_base + (uint32_t)slot * 256u;
    cx.lane_addr = cx.tmem + ((uint32_t)(cx.quarter * 32) << 16);
    cx.full = s.bars + 1 + 2 * slot;
    cx.bar_base = 1 + 2 * slot;
    cx.n_steps = 0;
    cx.n_issued = cx.n_waited = 13u * seq;
    long ltile = tile;
    int sidx = 0;
    if constexpr (MULTI) {   // <= 4 slabs: compares instead of a 64-bit division
      const long tps = p.tiles_per_slab;
      sidx = (tile >= tps) + (tile >= 2 * tps) + (tile >= 3 * tps);
      ltile = tile - (long)sidx * tps;
    }
    const long q = p.q_begin + ltile * kTile + cx.row;
    const bool valid = q < p.q_end;
    const long qc = valid ? q : p.q_end - 1;
    // (rasters of this mode have at most 2^24 pixels, stif_api.cu: 32-bit division)
    const int jy = (int)((uint32_t)qc / (uint32_t)g.WW), jx = (int)((uint32_t)qc - (uint32_t)jy * (uint32_t)g.WW);
    trace_mark(cx, 1);
    // ---- stage A, first layer (hoisted), WITHOUT the slot: h0 = sin(TA[iy,ix] + rel . w_rel + cA)      (:382-400)
    if constexpr (!ISSUER) {
      const int iy = g.y.idx[jy], ix = g.x.idx[jx];
      const float rely = g.y.rel[jy], relx = g.x.rel[jx];
      const bool inb = (iy >= 0) & (iy < g.H) & (ix >= 0) & (ix < g.W);
      const uint4* ta = tab4 + (inb ? ((long)iy * g.W + ix) : 0) * 32 + CH * 4;
      STIF_BOUND(jy, g.HH); STIF_BOUND(jx, g.WW); STIF_BOUND(qc, (long)g.HH * g.WW);
      uint32_t pk[16];
      U8x32 ta2;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if ((j & 1) == 0) ta2 = ldg256(ta + j);
        uint32_t w4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) w4[e] = inb ? ta2.r[(j & 1) * 4 + e] : 0u;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = ch0 + j * 8 + e * 2;
          const float b0 = fmaf(rely, p.c.a_rel[2 * c], fmaf(relx, p.c.a_rel[2 * c + 1], MULTI ? p.slab[sidx].cA[c] : p.c.cA[c]));
          const float b1 = fmaf(rely, p.c.a_rel[2 * c + 2], fmaf(relx, p.c.a_rel[2 * c + 3], MULTI ? p.slab[sidx].cA[c + 1] : p.c.cA[c + 1]));
          pk[j * 4 + e] = pack_bf16x2(fast_sin(add_f16((uint16_t)(w4[e] & 0xFFFF), b0)), fast_sin(add_f16((uint16_t)(w4[e] >> 16), b1)));
        }
      }
      trace_mark(cx, 2);
      if (seq > 0) {
        mbar_wait_long(&h2_read[slot], (seq - 1) & 1);   // the slot's previous tile no longer reads columns [0, 96)
        tc_fence_after();
      }
      tmem_st16(cx.lane_addr + kColH0 + CH * 16, pk);
      tmem_st_wait();
      tc_fence_before();
      if (seq > 0) mbar_wait_long(&slot_free[slot], (seq - 1) & 1);   // ... and has left the slot and its named barriers
    }
    step_done<ISSUER>(cx);
    trace_mark(cx, 3);

    // ---- feat_imnet hidden layers (MMA phase on the slot from here on)
    bias_blocks(k1rBiasF1);
    run_layer2<1, 4, false, ISSUER>(cx, cx.tmem + kColH0, wsm + k1F1, 64, [](int) { return 0; }, [&](int, int h, const uint32_t(&v)[16]) {
      epi16_sin_to_tmem<NB>(v, cx.lane_addr + kColAin + CH * 16 + h * 8, p.c.f1_b + ch0 + 16 * h);
    });
    bias_blocks(k1rBiasF2);
    run_layer2<4, 4, false, ISSUER>(cx, cx.tmem + kColAin, wsm + k1F2, 256, [](int i) { return i; }, [&](int i, int h, const uint32_t(&v)[16]) {
      epi16_sin_to_tmem<NB>(v, cx.lane_addr + kColA + 32 * i + CH * 16 + h * 8, p.c.f2_b + 64 * i + ch0 + 16 * h);
    });

    // ---- composed last layer of feat_imnet: chunk order Q1, Q2, F
    auto f3_order = [](int i) { return i == 2 ? 0 : i + 1; };
    if constexpr (ISSUER) cx.extra_commit = &h2_read[slot];
    bias_blocks(k1rBiasF3);
    layer_begin<3, 16, false, ISSUER>(cx, cx.tmem + kColA, wsm + k1F3, 192, f3_order);
    // ---- stage B gather: gB = bilinear(TB; query position) + cB + composed bias of F, kept as 32 fp16 values  (:410-418)
    uint32_t gBh[16];
    if constexpr (!ISSUER) {
      const Taps tp = make_taps_tables(g, jy, jx);
      uint16_t wq[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { wq[k] = __half_as_ushort(__float2half_rn(tp.w[k])); STIF_BOUND(tp.off[k], (long)g.H * g.W); }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float gB[16];
#pragma unroll
        for (int e = 0; e < 16; ++e)
          gB[e] = (MULTI ? p.slab[sidx].cB[ch0 + 16 * j + e] : p.c.cB[ch0 + 16 * j + e]) + (NB ? p.c.f3_b[ch0 + 16 * j + e] : 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const U8x32 v = ldg256(tab4 + (long)tp.off[k] * 32 + 8 + CH * 4 + 2 * j);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            gB[2 * e] = fma_f16((uint16_t)(v.r[e] & 0xFFFF), wq[k], gB[2 * e]);
            gB[2 * e + 1] = fma_f16((uint16_t)(v.r[e] >> 16), wq[k], gB[2 * e + 1]);
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) gBh[8 * j + e] = pack_half2(gB[2 * e], gB[2 * e + 1]);
      }
    }
    trace_mark(cx, 4);
    layer_finish2<3, 16, false, ISSUER>(cx, cx.tmem + kColA, wsm + k1F3, 192, f3_order, [&](int i, int h, const uint32_t(&v)[16]) {
      if (i < 2) {
        epi16_store_qtab<NB>(v, p.c.f3_b + 64 * (i + 1) + ch0 + 16 * h, (MULTI ? p.slab[sidx].qtab : p.qtab) + qc * 128 + 64 * i + ch0 + 16 * h, valid);
      } else {   // f0 = sin(F + gB) -> bf16 -> TMEM
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          pk[j] = pack_bf16x2(fast_sin(add_f16((uint16_t)(gBh[8 * h + j] & 0xFFFF), __uint_as_float(v[2 * j]))),
                              fast_sin(add_f16((uint16_t)(gBh[8 * h + j] >> 16), __uint_as_float(v[2 * j + 1]))));
        tmem_st8(cx.lane_addr + kColAin + CH * 16 + h * 8, pk);
      }
    });
    if constexpr (ISSUER) cx.extra_commit = nullptr;

    // ---- flow_imnet hidden layers; the 256->4 output layer rides the FMA pipe          (:419-422)
    bias_blocks(k1rBiasL1);
    run_layer2<1, 4, false, ISSUER>(cx, cx.tmem + kColAin, wsm + k1L1, 64, [](int) { return 0; }, [&](int, int h, const uint32_t(&v)[16]) {
      epi16_sin_to_tmem<NB>(v, cx.lane_addr + kColAin + CH * 16 + h * 8, p.c.l1_b + ch0 + 16 * h);
    });
    float2 fl[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
    bias_blocks(k1rBiasL2);
    run_layer2<4, 4, false, ISSUER>(cx, cx.tmem + kColAin, wsm + k1L2, 256, [](int i) { return i; }, [&](int i, int h, const uint32_t(&v)[16]) {
      epi16_sin_fma<4, NB>(v, p.c.l2_b + 64 * i + ch0 + 16 * h, cs + kc1L3W + 64 * i + ch0 + 16 * h, fl);
    });
    if constexpr (ISSUER) {
      if ((threadIdx.x & 31) == 0) mbar_arrive(&slot_free[slot]);
      __syncwarp();
      continue;
    }
    const float4 mine = make_float4(fl[0].x + fl[0].y, fl[1].x + fl[1].y, fl[2].x + fl[2].y, fl[3].x + fl[3].y);
    if (CH == 1) part[cx.row] = mine;
    asm volatile("bar.sync %0, 256;" ::"r"(5 + me) : "memory");
    if (CH == 0 && valid) {
      const float4 o = part[cx.row];
      STIF_BOUND(q, (long)g.HH * g.WW);
      reinterpret_cast<float4*>(MULTI ? p.slab[sidx].flow : p.flow)[q] =
          make_float4(mine.x + o.x + p.c.l3_b[0], mine.y + o.y + p.c.l3_b[1], mine.z + o.z + p.c.l3_b[2], mine.w + o.w + p.c.l3_b[3]);
    }
  }
}

template <bool MULTI = false>
__global__ void __launch_bounds__(832, 1) k1_stage_ab_rot_kernel(const __grid_constant__ K1Params p) {
  const CtaSetup s = cta_prologue(k1rBars, 0, p.wimg, k1WBytes, 512);
  {
    float* cs = reinterpret_cast<float*>(smem + k1rConst);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) cs[kc1L3W + i] = p.c.l3_w[i];
    if constexpr (kK1BiasMma) {   // 832 threads = 13 chunks x 64 bias rows; f1_b | f2_b | f3_b and l1_b | l2_b are contiguous in K1Consts
      const int t = threadIdx.x;
      const float b = t < 512 ? p.c.f1_b[t] : p.c.l1_b[t - 512];
      const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
      uint8_t* row = smem + k1rBias + (t >> 6) * kBiasBlock;
      *reinterpret_cast<uint4*>(row + nosw_offset(t & 63, 0)) = pack_bias_3term(b);
      *reinterpret_cast<uint4*>(row + nosw_offset(t & 63, 8)) = zero;
      if (t < 128) {
        *reinterpret_cast<uint4*>(smem + k1rOnes + nosw_offset(t, 0)) = make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u);
        *reinterpret_cast<uint4*>(smem + k1rOnes + nosw_offset(t, 8)) = zero;
      }
      fence_proxy_async_smem();
    }
    __syncthreads();
  }
  WgCtx cx = make_wg(s);
  if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0) cx.trace = p.trace + (cx.issuer ? 24 + cx.wg : cx.slot) * 4096;
  mbar_wait_or_trap(&s.bars[0], 0);
  if (cx.issuer) k1_rot_loop<true, MULTI>(p, s, cx, cx.wg);
  else k1_rot_loop<false, MULTI>(p, s, cx, cx.wg);
  cta_epilogue(s.tmem_base, 512);
}

// ---- decoding_localensemble passes (Sakuya_arch_test.py:962-1085) ------------------------------------------------
// Every gather of a pass uses coordinates shifted by half an LR texel (the per-pass axis tables in p.g), including the
// "identity" nearest gather of HRfeat in stage B: F is needed at ANOTHER pixel (g.y.hidx / g.x.hidx), so stage A and
// stage B cannot share a tile.  PASS 1 = stage A for the raster: Q1 | Q2 (fp16) and F + bias (fp32) tables.
// PASS 2 = stage B: f0 = sin(F[hidx] + bilinear(TB) + cB) -> flow_imnet -> flow.  Same building blocks as the fused loop.
template <bool ISSUER, int PASS>
__device__ __forceinline__ void k1_ens_tile_loop(const K1Params& p, const CtaSetup& s, WgCtx& cx) {
  const int CH = cx.colhalf;
  const uint32_t wsm = smem_u32(smem);
  const Geometry& g = p.g;
  const long ntiles = (p.q_end - p.q_begin + kTile - 1) / kTile;
  const uint4* __restrict__ tab4 = reinterpret_cast<const uint4*>(p.tab);
  float4* part = reinterpret_cast<float4*>(smem + k1Part) + cx.wg * 128;
  const float* cs = reinterpret_cast<const float*>(smem + k1Const);
  const int ch0 = CH * 32;
  for (long tile = (long)blockIdx.x * 2 + cx.wg; tile < ntiles; tile += (long)gridDim.x * 2) {
    const long q = p.q_begin + tile * kTile + cx.row;
    const bool valid = q < p.q_end;
    const long qc = valid ? q : p.q_end - 1;
    const int jy = (int)(qc / g.WW), jx = (int)(qc - (long)jy * g.WW);
    if constexpr (PASS == 1) {
      if constexpr (!ISSUER) {   // stage A first layer, exactly as in the fused loop
        const int iy = g.y.idx[jy], ix = g.x.idx[jx];
        const float rely = g.y.rel[jy], relx = g.x.rel[jx];
        const bool inb = (iy >= 0) & (iy < g.H) & (ix >= 0) & (ix < g.W);
        const uint4* ta = tab4 + (inb ? ((long)iy * g.W + ix) : 0) * 32 + CH * 4;
        uint32_t pk[16];
        U8x32 ta2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if ((j & 1) == 0) ta2 = ldg256(ta + j);
          uint32_t w4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) w4[e] = inb ? ta2.r[(j & 1) * 4 + e] : 0u;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = ch0 + j * 8 + e * 2;
            const float b0 = fmaf(rely, p.c.a_rel[2 * c], fmaf(relx, p.c.a_rel[2 * c + 1], p.c.cA[c]));
            const float b1 = fmaf(rely, p.c.a_rel[2 * c + 2], fmaf(relx, p.c.a_rel[2 * c + 3], p.c.cA[c + 1]));
            pk[j * 4 + e] = pack_bf16x2(fast_sin(add_f16((uint16_t)(w4[e] & 0xFFFF), b0)), fast_sin(add_f16((uint16_t)(w4[e] >> 16), b1)));
          }
        }
        tmem_st16(cx.lane_addr + kColAin + CH * 16, pk);
        tmem_st_wait();
        tc_fence_before();
      }
      step_done<ISSUER>(cx);
      run_layer<1, 4, false, ISSUER>(cx, cx.tmem + kColAin, wsm + k1F1, 64, [](int) { return 0; },
                   [&](int, uint32_t(&v)[32], auto&& pf) { epi_sin_to_tmem(v, cx.lane_addr + kColAin + CH * 16, p.c.f1_b + ch0, pf); });
      run_layer<4, 4, false, ISSUER>(cx, cx.tmem + kColAin, wsm + k1F2, 256, [](int i) { return i; }, [&](int i, uint32_t(&v)[32], auto&& pf) {
        epi_sin_to_tmem(v, cx.lane_addr + kColA + 32 * i + CH * 16, p.c.f2_b + 64 * i + ch0, pf);
      });
      auto f3_order = [](int i) { return i == 2 ? 0 : i + 1; };
      run_layer<3, 16, false, ISSUER>(cx, cx.tmem + kColA, wsm + k1F3, 192, f3_order, [&](int i, uint32_t(&v)[32], auto&& pf) {
        if (i < 2) epi_store_qtab(v, p.c.f3_b + 64 * (i + 1) + ch0, p.qtab + qc * 128 + 64 * i + ch0, valid, pf);
        else epi_store_ftab(v, p.c.f3_b + ch0, p.ftab + qc * 64 + ch0, valid, pf);
      });
    } else {
      if constexpr (!ISSUER) {   // f0 = sin(F[shifted nearest HR pixel] + bilinear(TB; shifted position) + cB)    (:1003-1030)
        float gB[32];
        const Taps tp = make_taps_tables(g, jy, jx);
        uint16_t wq[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) wq[k] = __half_as_ushort(__float2half_rn(tp.w[k]));
        const int hy = g.y.hidx[jy], hx = g.x.hidx[jx];
        const bool hin = (hy >= 0) & (hy < g.HH) & (hx >= 0) & (hx < g.WW);   // zero padding of the nearest gather
        const float* frow = p.ftab + ((long)(hin ? hy : 0) * g.WW + (hin ? hx : 0)) * 64 + ch0;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const U8x32 f0 = ldg256(frow + 16 * j), f1 = ldg256(frow + 16 * j + 8);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            gB[16 * j + e] = p.c.cB[ch0 + 16 * j + e] + (hin ? __uint_as_float(f0.r[e]) : 0.f);
            gB[16 * j + 8 + e] = p.c.cB[ch0 + 16 * j + 8 + e] + (hin ? __uint_as_float(f1.r[e]) : 0.f);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const U8x32 v = ldg256(tab4 + (long)tp.off[k] * 32 + 8 + CH * 4 + 2 * j);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              gB[16 * j + 2 * e] = fma_f16((uint16_t)(v.r[e] & 0xFFFF), wq[k], gB[16 * j + 2 * e]);
              gB[16 * j + 2 * e + 1] = fma_f16((uint16_t)(v.r[e] >> 16), wq[k], gB[16 * j + 2 * e + 1]);
            }
          }
        }
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(fast_sin(gB[2 * j]), fast_sin(gB[2 * j + 1]));
        tmem_st16(cx.lane_addr + kColAin + CH * 16, pk);
        tmem_st_wait();
        tc_fence_before();
      }
      step_done<ISSUER>(cx);
      run_layer<1, 4, false, ISSUER>(cx, cx.tmem + kColAin, wsm + k1L1, 64, [](int) { return 0; },
                   [&](int, uint32_t(&v)[32], auto&& pf) { epi_sin_to_tmem(v, cx.lane_addr + kColAin + CH * 16, p.c.l1_b + ch0, pf); });
      float2 fl[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
      run_layer<4, 4, false, ISSUER>(cx, cx.tmem + kColAin, wsm + k1L2, 256, [](int i) { return i; },
                   [&](int i, uint32_t(&v)[32], auto&& pf) { epi_sin_fma<4>(v, p.c.l2_b + 64 * i + ch0, cs + kc1L3W + 64 * i + ch0, fl, pf); });
      if constexpr (ISSUER) continue;
      const float4 mine = make_float4(fl[0].x + fl[0].y, fl[1].x + fl[1].y, fl[2].x + fl[2].y, fl[3].x + fl[3].y);
      if (CH == 1) part[cx.row] = mine;
      wg_barrier(cx.wg);
      if (CH == 0 && valid) {
        const float4 o = part[cx.row];
        reinterpret_cast<float4*>(p.flow)[q] =
            make_float4(mine.x + o.x + p.c.l3_b[0], mine.y + o.y + p.c.l3_b[1], mine.z + o.z + p.c.l3_b[2], mine.w + o.w + p.c.l3_b[3]);
      }
      wg_barrier(cx.wg);   // `part` is rewritten one short tile later
    }
  }
}

template <int PASS>
__global__ void __launch_bounds__(576, 1) k1_ensemble_kernel(const __grid_constant__ K1Params p) {
  const CtaSetup s = cta_prologue(k1Bars, 0, p.wimg, k1WBytes, 512);
  {
    float* cs = reinterpret_cast<float*>(smem + k1Const);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) cs[kc1L3W + i] = p.c.l3_w[i];
    __syncthreads();
  }
  WgCtx cx = make_wg(s);
  mbar_wait_or_trap(&s.bars[0], 0);
  if (cx.issuer) k1_ens_tile_loop<true, PASS>(p, s, cx);
  else k1_ens_tile_loop<false, PASS>(p, s, cx);
  cta_epilogue(s.tmem_base, 512);
}

// =================================================================================================
// K2: stage C + D + E
// =================================================================================================
// Warp-cooperative gather of encode_imnet's (hoisted) first layer for the warp's 16 queries.
// Phase 1: lane = (query, which warp) computes one warp position and its two bilinear footprints
// (HR grid for Q, LR grid for TE) and stages byte offsets + fp16 weights in the warp's private
// 1.5 KB of shared memory.  Phase 2: 8 lanes per query, 8 channels per lane: every tap is a
// 16-byte load (one 128-byte line per query), blended with mixed-precision FMAs (fp16 table value
// x fp16 weight + fp32 accumulator), then sine -> bf16 -> the SW128 A tile of the first MMA.
// K2 tile geometry: row r (0..127) of tile `tile` -> linear query index (clamped into the launch's rows / the raster so
// that loads stay in range; `valid` says whether the query really exists).
__device__ __forceinline__ long k2_query(const K2Params& p, long tile, int r, bool& valid, int& jy, int& jx) {
  const int t32 = (int)tile;   // tiles per launch < 2^31 (a 16 K x 16 K raster has 2^21)
  const int ty = t32 / p.tiles_x, tx = t32 - ty * p.tiles_x;
  const int patch = r >> 4, i = r & 15;
  const int y = p.row_begin + ty * 8 + (patch >> 2) * 4 + (i >> 2);
  const int x = p.col_begin + tx * 16 + (patch & 3) * 4 + (i & 3);
  valid = (y < p.row_end) & (x < p.col_end);
  jy = min(y, p.row_end - 1);
  jx = min(x, p.col_end - 1);
  return (long)jy * p.g.WW + jx;
}

// Tap staging of a warp's 16 queries lives INSIDE the warp's own 16 rows (2 KB) of the WG's A tile: query group it
// (4 queries, 384 B of taps) sits at the start of rows [4 it, 4 it + 4) (512 B; the SW128 swizzle stays inside a row),
// which phase 2 overwrites only after it has read that group's taps.  uint4 index of query qi's 6-entry slot:
__device__ __forceinline__ int tap_slot(int qi) { return (qi >> 2) * 32 + (qi & 3) * 6; }

// phase 1 (bilinear footprints -> per-warp staging); issued one tile ahead, under the 256->256 layer's first MMA wait
template <bool BAND, bool MULTI = false>
__device__ __forceinline__ void k2_gather_taps(const K2Params& p, uint4* stg, long tile, int warp_in_wg, int lane, int sidx = 0) {
  const Geometry& g = p.g;
  {
    const int qi = lane & 15, which = lane >> 4;
    bool valid_;
    int jy, jx;
    const long q = k2_query(p, tile, warp_in_wg * 16 + qi, valid_, jy, jx);
    const float4 fl = __ldg(reinterpret_cast<const float4*>(MULTI ? p.slab[sidx].flow : p.flow) + q);
    float gy, gx;
    warp_position(g, jy, jx, which ? fl.z : fl.x, which ? fl.w : fl.y, gy, gx);   // (warplayer.py:25-39)
    Taps hr = make_taps(gy, gx, g.HH, g.WW);
    Taps lr = make_taps(gy, gx, g.H, g.W);
#pragma unroll
    for (int k = 0; k < 4; ++k) { STIF_BOUND(hr.off[k], (long)g.HH * g.WW); STIF_BOUND(lr.off[k], (long)g.H * g.W); }
    STIF_BOUND(q, (long)g.HH * g.WW);
    // Row-band launches: rows outside [band_lo, band_hi) of the Q table (and the LR rows behind them) may not be
    // written yet.  A tap that carries weight there is a halo violation (flagged; the host repeats the launch);
    // a zero-weight tap is redirected to data that certainly exists (the query's own pixel / texel 0), because
    // 0 x stale bits is not 0 when the bits are a NaN.  (BAND = false: full-raster launches, where every table row exists and the check is compiled out.)
    if constexpr (BAND) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (hr.w[k] == 0.f) hr.off[k] = (int)q;
        else if (hr.off[k] < p.band_lo_off || hr.off[k] >= p.band_hi_off) atomicOr(p.flag, 1);
        if (lr.w[k] == 0.f) lr.off[k] = 0;
      }
    }
    const uint32_t cb = which * 128u;
    uint4* dst = stg + tap_slot(qi);
    dst[which * 2 + 0] = make_uint4((uint32_t)hr.off[0] * 256u + cb, (uint32_t)hr.off[1] * 256u + cb,
                                    (uint32_t)hr.off[2] * 256u + cb, (uint32_t)hr.off[3] * 256u + cb);
    dst[which * 2 + 1] = make_uint4((uint32_t)lr.off[0] * 512u + 256u + cb, (uint32_t)lr.off[1] * 512u + 256u + cb,
                                    (uint32_t)lr.off[2] * 512u + 256u + cb, (uint32_t)lr.off[3] * 512u + 256u + cb);
    dst[4 + which] = make_uint4(pack_half2(hr.w[0], hr.w[1]), pack_half2(hr.w[2], hr.w[3]), pack_half2(lr.w[0], lr.w[1]),
                                pack_half2(lr.w[2], lr.w[3]));
  }
  __syncwarp();
}

// 16-byte read-only load with an L1 eviction-priority hint (tuning knob of the K2 gather: 0 default, 1 evict_last,
// 2 no_allocate, 3 evict_first)
#ifndef STIF_QTAP_HINT
#define STIF_QTAP_HINT 0
#endif
#ifndef STIF_TTAP_HINT
#define STIF_TTAP_HINT 0
#endif
template <int HINT>
__device__ __forceinline__ uint4 ldg128_hint(const uint4* p) {
  uint4 v;
  if (HINT == 1) asm volatile("ld.global.nc.L1::evict_last.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else if (HINT == 2) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else if (HINT == 3) asm volatile("ld.global.nc.L1::evict_first.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else v = __ldg(p);
  return v;
}

// phase 2 (loads + blend + sine -> A tile)
// UADD (compile time: the hook costs registers the default kernel does not have): + p.uadd[q], decoding_test away from x4
template <bool UADD = false, bool MULTI = false, class Sig>
__device__ __forceinline__ void k2_gather_blend(const K2Params& p, uint8_t* a0, const uint4* stg, int warp_in_wg, int lane, Sig&& loads_issued,
                                                long tile = 0, int sidx = 0) {
  const char* __restrict__ qtab_b = reinterpret_cast<const char*>(MULTI ? p.slab[sidx].qtab : p.qtab);
  const char* __restrict__ tab_b = reinterpret_cast<const char*>(p.tab);
  const int sub = lane & 7;
  float cE[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) cE[e] = MULTI ? p.slab[sidx].cE[sub * 8 + e] : p.c.cE[sub * 8 + e];
  // Software pipeline over 8 half-steps (4 queries x one warp position = 8 taps = 8 x 16 B per lane each): the loads
  // of half-step s+1 are in flight while half-step s is blended, so a warp exposes one memory round trip per tile
  // instead of four.  Step s: query group it = s/2, warp position which = s%2 (which 0 -> Q1/TE1 @ g1, 1 -> Q2/TE2 @ g2).
  uint4 va[8], vb[8];
  uint32_t wa[4], wb[4];
  auto load_step = [&](int s_, uint4 (&v)[8], uint32_t (&w)[4]) {
    const int qloc = (s_ >> 1) * 4 + (lane >> 3), which = s_ & 1;
    const uint4* sq = stg + tap_slot(qloc);
    const uint4 oh = sq[which * 2 + 0], ol = sq[which * 2 + 1], wq = sq[4 + which];
    const uint32_t off[8] = {oh.x, oh.y, oh.z, oh.w, ol.x, ol.y, ol.z, ol.w};
    w[0] = wq.x; w[1] = wq.y; w[2] = wq.z; w[3] = wq.w;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      v[k] = (STIF_DIAG & (k < 4 ? 8 : 16)) ? make_uint4(0, 0, 0, 0)
                                             : (k < 4 ? ldg128_hint<STIF_QTAP_HINT>(reinterpret_cast<const uint4*>(qtab_b + off[k]) + sub)
                                                      : ldg128_hint<STIF_TTAP_HINT>(reinterpret_cast<const uint4*>(tab_b + off[k]) + sub));
  };
  __half2 acc2[4];
  // packed fp16 FMAs (2 channels per instruction, fp16 accumulate: the emulator shows the RGB error is unchanged)
  auto blend_step = [&](const uint4 (&v)[8], const uint32_t (&w)[4]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const __half2 wpair = *reinterpret_cast<const __half2*>(&w[k >> 1]);
      const __half2 w2 = (k & 1) ? __high2half2(wpair) : __low2half2(wpair);
      const uint32_t w4[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) acc2[e] = __hfma2(*reinterpret_cast<const __half2*>(&w4[e]), w2, acc2[e]);
    }
  };
  load_step(0, va, wa);
#pragma unroll
  for (int s_ = 0; s_ < 8; ++s_) {
    if (s_ + 1 < 8) {
      if (s_ & 1) load_step(s_ + 1, va, wa);
      else load_step(s_ + 1, vb, wb);
    }
    if (s_ + 1 == STIF_TURN_EARLY) loads_issued();
    if ((s_ & 1) == 0) {
#pragma unroll
      for (int e = 0; e < 4; ++e) acc2[e] = __float2half2_rn(0.f);
      blend_step(va, wa);
    } else {
      blend_step(vb, wb);
      float acc[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t a = *reinterpret_cast<const uint32_t*>(&acc2[e]);
        acc[2 * e] = add_f16((uint16_t)(a & 0xFFFF), cE[2 * e]);       // + fp32 time constant
        acc[2 * e + 1] = add_f16((uint16_t)(a >> 16), cE[2 * e + 1]);
      }
      const int r = warp_in_wg * 16 + (s_ >> 1) * 4 + (lane >> 3);
      if constexpr (UADD) {   // decoding_test away from x4: + bilinear(UE1; g1) + bilinear(UE2; g2), precomputed per query (warp_u_terms)
        bool valid_;
        int jy_, jx_;
        const long q_ = k2_query(p, tile, r, valid_, jy_, jx_);
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(p.uadd + q_ * 64) + sub);
        const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[2 * e] = add_f16((uint16_t)(uu[e] & 0xFFFF), acc[2 * e]);
          acc[2 * e + 1] = add_f16((uint16_t)(uu[e] >> 16), acc[2 * e + 1]);
        }
      }
      *reinterpret_cast<uint4*>(a0 + sw128_offset(r, sub * 8)) =
          make_uint4(pack_bf16x2(fast_sin(acc[0]), fast_sin(acc[1])), pack_bf16x2(fast_sin(acc[2]), fast_sin(acc[3])),
                     pack_bf16x2(fast_sin(acc[4]), fast_sin(acc[5])), pack_bf16x2(fast_sin(acc[6]), fast_sin(acc[7])));
    }
  }
  __syncwarp();
}

// output stage: RGB of query q, either fp32 planar (the reference's tensor) or what its caller makes of it
// (custom_video_test.py:102: `(img.clamp(0, 1) * 255).astype(uint8)` on the HWC frame -- fp32 clamp, fp32 multiply, truncation)
template <bool MULTI = false>
__device__ __forceinline__ void k2_store_rgb(const K2Params& p, long q, float r, float g, float b, int sidx = 0) {
  STIF_BOUND(q, p.plane);
  uint8_t* out_u8 = MULTI ? p.slab[sidx].out_u8 : p.out_u8;
  float* out = MULTI ? p.slab[sidx].out : p.out;
  if (out_u8) {
    uint8_t* o = out_u8 + q * 3;
    o[0] = (uint8_t)(fminf(fmaxf(r, 0.f), 1.f) * 255.f);
    o[1] = (uint8_t)(fminf(fmaxf(g, 0.f), 1.f) * 255.f);
    o[2] = (uint8_t)(fminf(fmaxf(b, 0.f), 1.f) * 255.f);
  } else {
    out[q] = r;
    out[p.plane + q] = g;
    out[2 * p.plane + q] = b;
  }
}

template <bool ISSUER, bool BAND>
__device__ __forceinline__ void k2_tile_loop(const K2Params& p, const CtaSetup& s, WgCtx& cx) {
  const int CH = cx.colhalf;
  const uint32_t wsm = smem_u32(smem);
  const int lane = threadIdx.x & 31, warp_in_wg = cx.warp_in_wg;
  uint8_t* a0 = smem + k2A0 + cx.wg * 16384;
  uint4* stg = reinterpret_cast<uint4*>(a0 + warp_in_wg * 2048);
  float4* part = reinterpret_cast<float4*>(smem + k2Part) + cx.wg * 128;
  const float* cs = reinterpret_cast<const float*>(smem + k2Const);   // shared-memory copy of the fp32 output layer
  const long ntiles = (long)p.tiles_x * ((p.row_end - p.row_begin + 7) / 8);
  const int ch0 = CH * 32;

  const long tile_first = (long)blockIdx.x * 2 + cx.wg;
  if (tile_first < ntiles && !ISSUER) k2_gather_taps<BAND>(p, stg, tile_first, warp_in_wg, lane);
  for (long tile = tile_first; tile < ntiles; tile += (long)gridDim.x * 2) {
    if (cx.trace && tile >= (long)gridDim.x * 2 * 16) cx.trace = nullptr;
    bool valid;
    int jy_, jx_;
    const long q = k2_query(p, tile, cx.row, valid, jy_, jx_);
    if constexpr (!ISSUER) gather_turn_wait(cx);

    // ---- stage C + D + first layer of encode_imnet (hoisted)                         (:424-456)
    trace_mark(cx, 1);
    if constexpr (!ISSUER) {
      k2_gather_blend(p, a0, stg, warp_in_wg, lane, [&]() { gather_turn_done(cx, tile == tile_first); });
      fence_proxy_async_smem();
      tc_fence_before();
    }
    trace_mark(cx, 2);
    step_done<ISSUER>(cx);
    trace_mark(cx, 3);

    if constexpr (!ISSUER) { if (STIF_TURN_EARLY > 8) gather_turn_done(cx, tile == tile_first); }
    SineTurn<9, 10> turn{cx.wg, tile == tile_first, tile + 1 < ntiles, tile - 1 + (long)gridDim.x * 2 < ntiles};
    // ---- encode_imnet hidden layers; the 256->3 output layer rides the FMA pipe       (:456-457)
    run_layer<1, 4, true, ISSUER>(cx, smem_u32(a0), wsm + k2E1, 64, [](int) { return 0; }, [&](int, uint32_t(&v)[32], auto&& pf) {
      if (STIF_K2_SINE_TURNS == 3) turn.template acquire<1>(0);
      epi_sin_to_tmem(v, cx.lane_addr + kColAin + CH * 16, p.c.e1_b + ch0, pf);
    });
    run_layer<4, 4, false, ISSUER>(cx, cx.tmem + kColAin, wsm + k2E2, 256, [](int i) { return i; }, [&](int i, uint32_t(&v)[32], auto&& pf) {
      if ((STIF_K2_SINE_TURNS == 1 || STIF_K2_SINE_TURNS == 2) && i == 0) turn.template acquire<STIF_K2_SINE_TURNS>(0);
      epi_sin_to_tmem(v, cx.lane_addr + kColA + 32 * i + CH * 16, p.c.e2_b + 64 * i + ch0, pf);
      if (STIF_K2_SINE_TURNS == 2 && i == 3) turn.template release<2>(0);
    });
    float2 rgb[3] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
    layer_begin<4, 16, false, ISSUER>(cx, cx.tmem + kColA, wsm + k2E3, 256, [](int i) { return i; });
    {  // footprints of this WG's next tile, computed while the first 256->256 chunk is on the tensor pipe
      const long tile_next = tile + (long)gridDim.x * 2;
      if (tile_next < ntiles && !ISSUER) k2_gather_taps<BAND>(p, stg, tile_next, warp_in_wg, lane);
    }
    layer_finish<4, 16, false, ISSUER>(cx, cx.tmem + kColA, wsm + k2E3, 256, [](int i) { return i; }, [&](int i, uint32_t(&v)[32], auto&& pf) {
      if (STIF_K2_SINE_TURNS == 2 && i == 0) turn.template acquire<2>(1);
      epi_sin_fma<3>(v, p.c.e3_b + 64 * i + ch0, cs + kc2E4W + 64 * i + ch0, rgb, pf);
      if (STIF_K2_SINE_TURNS && i == 3) turn.template release<(STIF_K2_SINE_TURNS == 2 ? 2 : 1)>(STIF_K2_SINE_TURNS == 2 ? 1 : 0);
    });
    if constexpr (ISSUER) continue;
    const float4 mine = make_float4(rgb[0].x + rgb[0].y, rgb[1].x + rgb[1].y, rgb[2].x + rgb[2].y, 0.f);
    if (CH == 1) part[cx.row] = mine;
    wg_barrier(cx.wg);
    if (CH == 0 && valid) {
      const float4 o = part[cx.row];
      k2_store_rgb(p, q, mine.x + o.x + p.c.e4_b[0], mine.y + o.y + p.c.e4_b[1], mine.z + o.z + p.c.e4_b[2]);
    }
    // (`part` is rewritten one whole tile later; no warp can run that far ahead of a sibling: every chunk's MMA waits
    // for all eight warps' arrival on the step barrier)
  }
}

template <bool BAND>
__global__ void __launch_bounds__(576, 1) k2_stage_cde_kernel(const __grid_constant__ K2Params p) {
  const CtaSetup s = cta_prologue(k2Bars, 0, p.wimg, k2WBytes, 512);
  {
    float* cs = reinterpret_cast<float*>(smem + k2Const);
    for (int i = threadIdx.x; i < 768; i += blockDim.x) cs[kc2E4W + i] = p.c.e4_w[i];
    __syncthreads();
  }
  WgCtx cx = make_wg(s);
  if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0) cx.trace = p.trace + cx.slot * 4096;
  mbar_wait_or_trap(&s.bars[0], 0);
  if (cx.issuer) k2_tile_loop<true, BAND>(p, s, cx);
  else k2_tile_loop<false, BAND>(p, s, cx);
  cta_epilogue(s.tmem_base, 512);
}



// ---- K2 as a three-workgroup rotation over two TMEM slots (STIF_K2_ROT) ---------------------------------------------
// A tile's life has a phase that needs no tensor memory -- the gather (warp positions, 16 tap loads per query, blend, first
// sine -> the bf16 A tile in SHARED memory; 36 % of a workgroup's time per tile, MUFU almost idle) -- and a phase that
// does (the three hidden layers: MMA round trips and 512 of the 640 sines).  With two workgroups the MUFU pipe idles
// whenever both are outside their sine epilogues, which at ~60 % non-sine time each is ~35 % of the time (ncu: XU 58 %).
// TMEM (512 columns = two tiles) forbids a third tile in the MMA phase, but not a third workgroup in the GATHER phase:
//   832 threads = 2 issuer warps (one per TMEM slot) + 3 workgroups x 8 warps.  Tile n of the CTA is gathered and decoded by
//   workgroup n % 3 on TMEM slot n % 2; while two workgroups run their MMA phases on the two slots the third gathers its
//   next tile, then takes over the slot that frees up first in the static order.
// Hand-over of a slot: its issuer warp runs the slot's tiles in order and arrives on bars[5 + slot] after the last step
// barrier of a tile (every epilogue warp has drained the slot and made its last arrival on the slot's named barriers);
// the next tile's workgroup waits for that before its first arrival on the same named barriers.  Counters restart per
// tile: 10 step barriers (even, so the id alternation restarts at 0) and 9 accumulator chunks per tile (the ring position
// of tile k of a slot is 9 k).
template <bool ISSUER, bool BAND, bool UADD, bool MULTI>
__device__ __forceinline__ void k2_rot_loop(const K2Params& p, const CtaSetup& s, WgCtx& cx, int me) {
  const int CH = cx.colhalf;
  const uint32_t wsm = smem_u32(smem);
  const int lane = threadIdx.x & 31, warp_in_wg = cx.warp_in_wg;
  const float* cs = reinterpret_cast<const float*>(smem + k2rConst);
  const long ntiles = MULTI ? p.ntiles_total : (long)p.tiles_x * ((p.row_end - p.row_begin + 7) / 8);
  const long stride = gridDim.x;
  const int ch0 = CH * 32;
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, s.tmem_base, 0);
  uint64_t* slot_free = s.bars + 5;
  constexpr int kStep = ISSUER ? 2 : 3;   // an issuer serves every second tile (its slot), a workgroup every third
  uint8_t* a0_mine = smem + k2A0 + me * 16384;                                    // (workgroups only)
  uint4* stg = reinterpret_cast<uint4*>(a0_mine + warp_in_wg * 2048);
  float4* part = reinterpret_cast<float4*>(smem + k2rPart) + me * 128;
  auto slab_of = [&](long t, long& lt) {   // launch tile -> (slab, tile within the slab)
    if constexpr (MULTI) {   // <= 4 slabs: compares instead of a 64-bit division
      const long tps = p.tiles_per_slab;
      const int sx = (t >= tps) + (t >= 2 * tps) + (t >= 3 * tps);
      lt = t - (long)sx * tps;
      return sx;
    }
    lt = t;
    return 0;
  };
  if constexpr (!ISSUER) {
    const long t0 = (long)blockIdx.x + (long)me * stride;
    long lt0;
    const int s0 = slab_of(t0 < ntiles ? t0 : 0, lt0);
    if (t0 < ntiles) k2_gather_taps<BAND, MULTI>(p, stg, lt0, warp_in_wg, lane, s0);
  }
  for (long n = me;; n += kStep) {
    const long tile = (long)blockIdx.x + n * stride;
    if (tile >= ntiles) break;
    const int wg = (int)(n % 3), slot = (int)(n & 1);
    const uint32_t seq = (uint32_t)(n >> 1);   // this tile is the slot's seq-th
    cx.tmem = tmem_base + (uint32_t)slot * 256u;
    cx.lane_addr = cx.tmem + ((uint32_t)(cx.quarter * 32) << 16);
    cx.full = s.bars + 1 + 2 * slot;
    cx.bar_base = 1 + 2 * slot;
    cx.n_steps = 0;
    cx.n_issued = cx.n_waited = 9u * seq;
    uint8_t* a0 = smem + k2A0 + wg * 16384;
    bool valid;
    int jy_, jx_;
    long ltile;
    const int sidx = slab_of(tile, ltile);
    const long q = k2_query(p, ltile, cx.row, valid, jy_, jx_);
    trace_mark(cx, 1);
    if constexpr (!ISSUER) {
      // ---- gather phase: no tensor memory, no named barrier of the slot                       (:424-456)
      k2_gather_blend<UADD, MULTI>(p, a0, stg, warp_in_wg, lane, []() {}, ltile, sidx);
      fence_proxy_async_smem();
      tc_fence_before();
      trace_mark(cx, 2);
      if (seq > 0) mbar_wait_long(&slot_free[slot], (seq - 1) & 1);   // the slot's previous tile has left it
      tc_fence_after();
    }
    step_done<ISSUER>(cx);
    trace_mark(cx, 3);
    // ---- MMA phase on the slot: encode_imnet hidden layers; the 256->3 output layer rides the FMA pipe   (:456-457)
    run_layer2<1, 4, true, ISSUER>(cx, smem_u32(a0), wsm + k2E1, 64, [](int) { return 0; }, [&](int, int h, const uint32_t(&v)[16]) {
      epi16_sin_to_tmem(v, cx.lane_addr + kColAin + CH * 16 + h * 8, p.c.e1_b + ch0 + 16 * h);
    });
    run_layer2<4, 4, false, ISSUER>(cx, cx.tmem + kColAin, wsm + k2E2, 256, [](int i) { return i; }, [&](int i, int h, const uint32_t(&v)[16]) {
      epi16_sin_to_tmem(v, cx.lane_addr + kColA + 32 * i + CH * 16 + h * 8, p.c.e2_b + 64 * i + ch0 + 16 * h);
    });
    float2 rgb[3] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
    layer_begin<4, 16, false, ISSUER>(cx, cx.tmem + kColA, wsm + k2E3, 256, [](int i) { return i; });
    if constexpr (!ISSUER) {   // footprints of this workgroup's next tile, under the first 256->256 chunk's MMA time
      const long tile_next = tile + 3 * stride;
      if (tile_next < ntiles) {
        long ltn;
        const int sn = slab_of(tile_next, ltn);
        k2_gather_taps<BAND, MULTI>(p, stg, ltn, warp_in_wg, lane, sn);
      }
    }
    layer_finish2<4, 16, false, ISSUER>(cx, cx.tmem + kColA, wsm + k2E3, 256, [](int i) { return i; }, [&](int i, int h, const uint32_t(&v)[16]) {
      epi16_sin_fma<3>(v, p.c.e3_b + 64 * i + ch0 + 16 * h, cs + kc2E4W + 64 * i + ch0 + 16 * h, rgb);
    });
    if constexpr (ISSUER) {
      if (lane == 0) mbar_arrive(&slot_free[slot]);   // after the tile's last step barrier: the slot and its barrier ids are free
      __syncwarp();
      continue;
    }
    const float4 mine = make_float4(rgb[0].x + rgb[0].y, rgb[1].x + rgb[1].y, rgb[2].x + rgb[2].y, 0.f);
    if (CH == 1) part[cx.row] = mine;
    asm volatile("bar.sync %0, 256;" ::"r"(5 + me) : "memory");
    if (CH == 0 && valid) {
      const float4 o = part[cx.row];
      k2_store_rgb<MULTI>(p, q, mine.x + o.x + p.c.e4_b[0], mine.y + o.y + p.c.e4_b[1], mine.z + o.z + p.c.e4_b[2], sidx);
    }
    // (`part` is rewritten one whole tile later, after the workgroup has passed two of its own barriers)
  }
}

template <bool BAND, bool UADD = false, bool MULTI = false>
__global__ void __launch_bounds__(832, 1) k2_stage_cde_rot_kernel(const __grid_constant__ K2Params p) {
  const CtaSetup s = cta_prologue(k2rBars, 0, p.wimg, k2WBytes, 512);
  {
    float* cs = reinterpret_cast<float*>(smem + k2rConst);
    for (int i = threadIdx.x; i < 768; i += blockDim.x) cs[kc2E4W + i] = p.c.e4_w[i];
    __syncthreads();
  }
  WgCtx cx = make_wg(s);   // warps 0, 1: issuers (cx.wg = TMEM slot); warps 2..25: workgroups 0..2
  if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0) cx.trace = p.trace + (cx.issuer ? 24 + cx.wg : cx.slot) * 4096;
  mbar_wait_or_trap(&s.bars[0], 0);
  if (cx.issuer) k2_rot_loop<true, BAND, UADD, MULTI>(p, s, cx, cx.wg);
  else k2_rot_loop<false, BAND, UADD, MULTI>(p, s, cx, cx.wg);
  cta_epilogue(s.tmem_base, 512);
}

void fill_from(float* dst, const std::vector<float>& src, size_t n) { std::copy(src.begin(), src.begin() + n, dst); }

}  // namespace

// =================================================================================================
// host side
// =================================================================================================
template <class P>
cudaError_t launch_pdl(void (*kern)(const P), int grid, int block, size_t smem_bytes, cudaStream_t stream, const P& p) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

struct TcWeights {
  uint8_t* d_k0 = nullptr;
  uint8_t* d_k0_lat = nullptr;   // decoding_test variant: W_tab without the frame columns of TB | TE1 | TE2
  float* d_w_up = nullptr;       // ... which become this [192,6] map on the upsampled frames
  uint8_t* d_k1 = nullptr;
  uint8_t* d_k2 = nullptr;
  K1Consts c1{};
  K2Consts c2{};
  std::vector<float> a_t, a_b, b_t, b_b, e_t, e_b;
};

TcWeights* tc_weights_create(const FoldedWeights& hw, std::string& err) {
  auto* t = new TcWeights();
  std::vector<uint8_t> i0, i0lat, i1, i2;
  {
    std::vector<float> wpad((size_t)256 * 256, 0.f);   // K padded 198 -> 256 with zeros
    for (int r = 0; r < 256; ++r)
      for (int k = 0; k < 198; ++k) wpad[(size_t)r * 256 + k] = hw.w_tab[(size_t)r * 198 + k];
    append_sw128_image(i0, wpad.data(), 256, 256);
    for (int r = 0; r < 256; ++r)
      for (int k = 0; k < 198; ++k) wpad[(size_t)r * 256 + k] = hw.w_tab_lat[(size_t)r * 198 + k];
    append_sw128_image(i0lat, wpad.data(), 256, 256);
  }
  append_sw128_image(i1, hw.f1_w.data(), 64, 64);
  append_sw128_image(i1, hw.f2_w.data(), 256, 64);
  append_sw128_image(i1, hw.f3_w.data(), 192, 256);
  append_sw128_image(i1, hw.l1_w.data(), 64, 64);
  append_sw128_image(i1, hw.l2_w.data(), 256, 64);
  append_sw128_image(i2, hw.e1_w.data(), 64, 64);
  append_sw128_image(i2, hw.e2_w.data(), 256, 64);
  append_sw128_image(i2, hw.e3_w.data(), 256, 256);
  if (i0.size() != k0WBytes || i1.size() != k1WBytes || i2.size() != k2WBytes) {
    err = "internal: weight image size mismatch";
    delete t;
    return nullptr;
  }
  cudaError_t e = cudaMalloc(&t->d_k0, i0.size());
  if (e == cudaSuccess) e = cudaMalloc(&t->d_k1, i1.size());
  if (e == cudaSuccess) e = cudaMalloc(&t->d_k2, i2.size());
  if (e == cudaSuccess) e = cudaMemcpy(t->d_k0, i0.data(), i0.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&t->d_k0_lat, i0lat.size());
  if (e == cudaSuccess) e = cudaMemcpy(t->d_k0_lat, i0lat.data(), i0lat.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&t->d_w_up, hw.w_up.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(t->d_w_up, hw.w_up.data(), hw.w_up.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k1_stage_ab_upf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1Smem);
  if (e == cudaSuccess) e = cudaMemcpy(t->d_k1, i1.data(), i1.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(t->d_k2, i2.data(), i2.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k0_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k0Smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k1_stage_ab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1Smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k2_stage_cde_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2Smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k2_stage_cde_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2Smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k1_stage_ab_rot_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1rSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k1_stage_ab_rot_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1rSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k2_stage_cde_rot_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2rSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k2_stage_cde_rot_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2rSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k2_stage_cde_rot_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2rSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k2_stage_cde_rot_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2rSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k2_stage_cde_rot_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2rSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k2_stage_cde_rot_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2rSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k1_ensemble_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1Smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k1_ensemble_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1Smem);
  if (getenv("STIF_DEBUG_ATTRS")) {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, k1_stage_ab_kernel);
    fprintf(stderr, "k1: regs %d maxThreads %d smem static %zu local %zu\n", a.numRegs, a.maxThreadsPerBlock, a.sharedSizeBytes, a.localSizeBytes);
    cudaFuncGetAttributes(&a, k2_stage_cde_kernel<false>);
    fprintf(stderr, "k2: regs %d maxThreads %d smem static %zu local %zu\n", a.numRegs, a.maxThreadsPerBlock, a.sharedSizeBytes, a.localSizeBytes);
  }
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
  if (t->d_k0) cudaFree(t->d_k0);
  if (t->d_k0_lat) cudaFree(t->d_k0_lat);
  if (t->d_w_up) cudaFree(t->d_w_up);
  if (t->d_k1) cudaFree(t->d_k1);
  if (t->d_k2) cudaFree(t->d_k2);
  delete t;
}

cudaError_t project_frames_up4_tc(const LaunchCtx& cx, const TcWeights* tw, const float* frames6, int H, int W, void* utab) {
  const long n = (long)16 * H * W * 192;
  project_frames_up4_half_kernel<<<(unsigned)((n + 255) / 256), 256, 0, cx.stream>>>(frames6, H, W, tw->d_w_up, reinterpret_cast<__half*>(utab));
  ++*cx.launch_counter;
  return cudaGetLastError();
}

cudaError_t project_latent_tc(const LaunchCtx& cx, const TcWeights* tw, const float* latent192, const float* frames6, int H,
                              int W, void* tab, int row_begin, int row_end, bool test_variant, bool latent_is_bf16) {
  K0Params p;
  p.latent = latent_is_bf16 ? nullptr : latent192;
  p.latent16 = latent_is_bf16 ? reinterpret_cast<const uint16_t*>(latent192) : nullptr;
  p.frames = frames6;
  p.tab = reinterpret_cast<__half*>(tab);
  p.wimg = test_variant ? tw->d_k0_lat : tw->d_k0;
  p.HW = (long)H * W;
  p.m_begin = (long)row_begin * W;
  p.m_end = (long)row_end * W;
  const long ntiles = (p.m_end - p.m_begin + kTile - 1) / kTile;
  if (cudaError_t e = launch_pdl(k0_project_kernel, (int)std::min<long>(cx.num_sms, ntiles), 512, k0Smem, cx.stream, p)) return e;
  ++*cx.launch_counter;
  return cudaGetLastError();
}

namespace {
// STIF_TRACE=<path>: dump clock64 timelines of block 0 (16 warps x first 16 tiles) for K1 and K2.
long long* trace_buffer() {
  static long long* buf = nullptr;
  static bool init = false;
  if (!init) {
    init = true;
    if (getenv("STIF_TRACE")) {
      cudaMalloc(&buf, 26 * 4096 * sizeof(long long));
    }
  }
  return buf;
}
void trace_dump(const char* kernel, cudaStream_t stream) {
  long long* buf = trace_buffer();
  if (!buf) return;
  cudaStreamSynchronize(stream);
  std::vector<long long> h(26 * 4096);
  cudaMemcpy(h.data(), buf, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  FILE* f = fopen(getenv("STIF_TRACE"), "a");
  if (!f) return;
  for (int w = 0; w < 26; ++w) {
    fprintf(f, "%s warp %d:", kernel, w);
    for (int i = 0; i + 1 < 4096 && h[w * 4096 + i] != 0; i += 2) fprintf(f, " %lld:%lld", h[w * 4096 + i], h[w * 4096 + i + 1]);
    fprintf(f, "\n");
  }
  fclose(f);
}
}  // namespace

cudaError_t decode_slab_tc(const LaunchCtx& cx, const TcWeights* tw, const Geometry& geo, const Workspace& ws, float t,
                           int row_begin, int row_end, int k1_row_begin, int k1_row_end, float* out_rgb, int stage, uint8_t* out_u8,
                           int col_begin, int col_end, const void* uadd) {
  const long WW = geo.WW;
  // Plain stage C-E launches go through the multi-slab instantiation with one slab: same arithmetic, and ptxas happens to allocate
  // it better (20-36 B of spills instead of 28-40; K2 0.512 -> 0.503 ms at config 2).
  if (stage == 2 && !uadd && col_begin == 0 && col_end == geo.WW) {
    float* o1 = out_rgb;
    uint8_t* o8 = out_u8;
    return decode_multi_tc(cx, tw, geo, &ws, &t, 1, row_begin, row_end, k1_row_begin, k1_row_end, &o1, out_u8 ? &o8 : nullptr, 2);
  }
  if (stage == 1 || stage == 3 || stage == 4 || stage == 5) {   // 1: fused stage A+B; 3 / 4: stage A / B of a local-ensemble pass; 5: fused, decoding_test at x4
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
    p.ftab = ws.ftab;
    p.utab = reinterpret_cast<const __half*>(ws.utab);
    p.wimg = tw->d_k1;
    p.q_begin = k1_row_begin * WW;
    p.q_end = k1_row_end * WW;
    const long ntiles = (p.q_end - p.q_begin + kTile - 1) / kTile;
    const int grid = (int)std::min<long>(cx.num_sms, (ntiles + 1) / 2);
    p.trace = trace_buffer();
    if (p.trace) cudaMemsetAsync(p.trace, 0, 26 * 4096 * sizeof(long long), cx.stream);
    if ((stage == 3 || stage == 4) && !ws.ftab) return cudaErrorInvalidValue;
    if (stage == 5 && !ws.utab) return cudaErrorInvalidValue;
#if STIF_K1_ROT
    if (stage == 1) {
      if (cudaError_t e = launch_pdl(k1_stage_ab_rot_kernel<false>, (int)std::min<long>(cx.num_sms, ntiles), 832, k1rSmem, cx.stream, p)) return e;
      ++*cx.launch_counter;
      trace_dump("K1", cx.stream);
      return cudaGetLastError();
    }
#endif
    if (cudaError_t e = stage == 1   ? launch_pdl(k1_stage_ab_kernel, grid, 576, k1Smem, cx.stream, p)
                        : stage == 3 ? launch_pdl(k1_ensemble_kernel<1>, grid, 576, k1Smem, cx.stream, p)
                        : stage == 4 ? launch_pdl(k1_ensemble_kernel<2>, grid, 576, k1Smem, cx.stream, p)
                                     : launch_pdl(k1_stage_ab_upf_kernel, grid, 576, k1Smem, cx.stream, p))
      return e;
    ++*cx.launch_counter;
    trace_dump("K1", cx.stream);
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
  p.out_u8 = out_u8;
  p.plane = (long)geo.HH * geo.WW;
  p.wimg = tw->d_k2;
  p.row_begin = row_begin;
  p.row_end = row_end;
  p.col_begin = col_begin < 0 ? 0 : col_begin;
  p.col_end = col_end < 0 ? geo.WW : col_end;
  p.uadd = reinterpret_cast<const __half*>(uadd);
  p.tiles_x = (p.col_end - p.col_begin + 15) / 16;
  p.band_lo_off = k1_row_begin * geo.WW;
  p.band_hi_off = k1_row_end * geo.WW;
  p.flag = ws.flag;
  const long ntiles = (long)p.tiles_x * ((row_end - row_begin + 7) / 8);
  const int grid = (int)std::min<long>(cx.num_sms, (ntiles + 1) / 2);
  p.trace = trace_buffer();
  if (p.trace) cudaMemsetAsync(p.trace, 0, 26 * 4096 * sizeof(long long), cx.stream);
  const bool band = k1_row_begin > 0 || k1_row_end < geo.HH;   // stage A+B rows are incomplete: check every weighted tap
#if STIF_K2_ROT
  const int grid_rot = (int)std::min<long>(cx.num_sms, ntiles);
  if (cudaError_t e = p.uadd ? (band ? launch_pdl(k2_stage_cde_rot_kernel<true, true>, grid_rot, 832, k2rSmem, cx.stream, p)
                                     : launch_pdl(k2_stage_cde_rot_kernel<false, true>, grid_rot, 832, k2rSmem, cx.stream, p))
                      : band   ? launch_pdl(k2_stage_cde_rot_kernel<true>, grid_rot, 832, k2rSmem, cx.stream, p)
                               : launch_pdl(k2_stage_cde_rot_kernel<false>, grid_rot, 832, k2rSmem, cx.stream, p))
    return e;
#else
  if (cudaError_t e = band ? launch_pdl(k2_stage_cde_kernel<true>, grid, 576, k2Smem, cx.stream, p)
                           : launch_pdl(k2_stage_cde_kernel<false>, grid, 576, k2Smem, cx.stream, p))
    return e;
#endif
  ++*cx.launch_counter;
  trace_dump("K2", cx.stream);
  return cudaGetLastError();
}


// The same stage of up to kMaxSlabs timesteps in ONE launch (band-major host pipeline): slab g has its own time t[g], tables
// ws[g] and output; rows / band limits are shared.  stage 1 = K1 (stage A+B), stage 2 = K2 (stage C-E).
cudaError_t decode_multi_tc(const LaunchCtx& cx, const TcWeights* tw, const Geometry& geo, const Workspace* ws, const float* t, int nslab,
                            int row_begin, int row_end, int k1_row_begin, int k1_row_end, float* const* out_rgb, uint8_t* const* out_u8,
                            int stage) {
  if (nslab < 1 || nslab > kMaxSlabs) return cudaErrorInvalidValue;
  const long WW = geo.WW;
  if (stage == 1) {
    K1Params p;
    p.c = tw->c1;
    p.g = geo;
    p.tab = reinterpret_cast<const __half*>(ws[0].tab);
    p.qtab = nullptr; p.flow = nullptr; p.ftab = nullptr; p.utab = nullptr;
    p.wimg = tw->d_k1;
    p.q_begin = k1_row_begin * WW;
    p.q_end = k1_row_end * WW;
    p.trace = nullptr;
    p.tiles_per_slab = (p.q_end - p.q_begin + kTile - 1) / kTile;
    p.ntiles_total = p.tiles_per_slab * nslab;
    for (int g = 0; g < nslab; ++g) {
      for (int c = 0; c < 64; ++c) {
        p.slab[g].cA[c] = tw->a_t[c] * t[g] + tw->a_b[c];
        p.slab[g].cB[c] = tw->b_t[c] * t[g] + tw->b_b[c];
      }
      p.slab[g].qtab = reinterpret_cast<__half*>(ws[g].qtab);
      p.slab[g].flow = ws[g].flow;
    }
    if (cudaError_t e = launch_pdl(k1_stage_ab_rot_kernel<true>, (int)std::min<long>(cx.num_sms, p.ntiles_total), 832, k1rSmem, cx.stream, p)) return e;
    ++*cx.launch_counter;
    return cudaGetLastError();
  }
  K2Params p;
  p.c = tw->c2;
  p.g = geo;
  p.tab = reinterpret_cast<const __half*>(ws[0].tab);
  p.qtab = nullptr; p.flow = nullptr; p.out = nullptr; p.out_u8 = nullptr; p.uadd = nullptr;
  p.plane = (long)geo.HH * geo.WW;
  p.wimg = tw->d_k2;
  p.row_begin = row_begin;
  p.row_end = row_end;
  p.col_begin = 0;
  p.col_end = geo.WW;
  p.tiles_x = (geo.WW + 15) / 16;
  p.band_lo_off = k1_row_begin * geo.WW;
  p.band_hi_off = k1_row_end * geo.WW;
  p.flag = ws[0].flag;
  p.trace = nullptr;
  p.tiles_per_slab = (long)p.tiles_x * ((row_end - row_begin + 7) / 8);
  p.ntiles_total = p.tiles_per_slab * nslab;
  for (int g = 0; g < nslab; ++g) {
    for (int c = 0; c < 64; ++c) p.slab[g].cE[c] = tw->e_t[c] * t[g] + tw->e_b[c];
    p.slab[g].qtab = reinterpret_cast<const __half*>(ws[g].qtab);
    p.slab[g].flow = ws[g].flow;
    p.slab[g].out = out_rgb[g];
    p.slab[g].out_u8 = out_u8 ? out_u8[g] : nullptr;
  }
  const int grid = (int)std::min<long>(cx.num_sms, p.ntiles_total);
  const bool band = k1_row_begin > 0 || k1_row_end < geo.HH;
  if (cudaError_t e = band ? launch_pdl(k2_stage_cde_rot_kernel<true, false, true>, grid, 832, k2rSmem, cx.stream, p)
                           : launch_pdl(k2_stage_cde_rot_kernel<false, false, true>, grid, 832, k2rSmem, cx.stream, p))
    return e;
  ++*cx.launch_counter;
  return cudaGetLastError();
}

// decoding_localensemble on the tensor-core kernels: four passes (shifted axis tables), each = stage A for the raster
// (Q + F tables), stage B (F gathered at the shifted nearest HR pixel), stage C-E into the pass prediction, then the
// area-weighted accumulation (bit-exact weights, kernels_fp32.cu).
cudaError_t decode_slab_tc_ensemble(const LaunchCtx& cx, const TcWeights* tw, const Geometry geo_pass[4], const AxisTables ens_y[2],
                                    const AxisTables ens_x[2], const Workspace& ws, float t, float* out_rgb) {
  for (int k = 0; k < 4; ++k) {
    const Geometry& geo = geo_pass[k];
    for (int stage : {3, 4})
      if (cudaError_t e = decode_slab_tc(cx, tw, geo, ws, t, 0, geo.HH, 0, geo.HH, ws.pred, stage)) return e;
    if (cudaError_t e = decode_slab_tc(cx, tw, geo, ws, t, 0, geo.HH, 0, geo.HH, ws.pred, 2)) return e;
    if (cudaError_t e = ensemble_blend_launch(cx, ws.pred, out_rgb, geo.HH, geo.WW, ens_y, ens_x, k)) return e;
  }
  return cudaSuccess;
}

}  // namespace stif
