// tc_primitives.cuh -- thin inline-PTX wrappers for the Blackwell (sm_100a) features the fused
// decoder kernels use: mbarrier, 1-D bulk TMA copies, tensor memory (TMEM) allocation and
// tcgen05.{mma,commit,ld,st,fence,wait}.  No CUTLASS dependency.
//
// Layout conventions used throughout kernels_tc.cu
//   * smem operands are K-major, 128-byte swizzled ("SW128"): a tile of R rows x 64 bf16 (128 B per row);
//     rows are grouped by 8 (1024 B, the swizzle atom), 16-byte chunk j of row r is stored at chunk
//     position j ^ (r & 7).  The tile base must be 1024-byte aligned.
//   * accumulators: TMEM lane = tile row (query), 32-bit column = output channel (fp32).
//   * TMEM-resident A operands (tcgen05.mma "TS" form): lane = row, 32-bit column c holds the bf16
//     pair (k = 2c in bits [0,16), k = 2c+1 in bits [16,32)).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace stif {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ------------------------------------------------------------------ bulk async copy (TMA, 1-D)
// global -> shared::cta, completion signalled on an mbarrier (bytes must be a multiple of 16).
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// 256-bit global accesses (sm_100+: LDG.256 / STG.256): one instruction per full 32-byte sector
struct alignas(32) U8x32 { uint32_t r[8]; };
__device__ __forceinline__ U8x32 ldg256(const void* p) {
  U8x32 a;
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.r[0]), "=r"(a.r[1]), "=r"(a.r[2]), "=r"(a.r[3]), "=r"(a.r[4]), "=r"(a.r[5]), "=r"(a.r[6]), "=r"(a.r[7])
               : "l"(p));
  return a;
}
__device__ __forceinline__ void stg256(void* p, const uint32_t* a) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]),
               "r"(a[5]), "r"(a[6]), "r"(a[7])
               : "memory");
}
// shared::cta -> global bulk store (TMA); completion tracked with bulk groups of the issuing thread
__device__ __forceinline__ void bulk_store_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// the issuing thread's bulk stores have finished READING shared memory (the buffer may be rewritten)
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... and have fully completed (global writes performed)
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make generic-proxy smem writes visible to the async proxy (tcgen05.mma reads smem through it)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------ tensor memory
// One full warp executes alloc/dealloc.  ncols: power of two in [32, 512].
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ descriptors
// Shared-memory matrix descriptor, K-major, SWIZZLE_128B, 8-row groups 1024 B apart.
//   bits [0,14)  start address >> 4          bits [16,30) leading byte offset >> 4 (unused for swizzled K-major; 1)
//   bits [32,46) stride byte offset >> 4     bits [46,48) descriptor version = 1 (sm_100)
//   bits [61,64) layout type: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// K-major INTERLEAVED (no swizzle) operand of K = 16 bf16: 8 x 16-byte core matrices, the second K half LBO bytes on, the next
// 8 rows SBO bytes on (cute UMMA canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units).  Used for the bias K step.
constexpr uint32_t kNoswLBO = 128, kNoswSBO = 256;
__host__ __device__ constexpr uint32_t nosw_offset(int r, int k) {
  return (uint32_t)((r >> 3) * kNoswSBO + (k >> 3) * kNoswLBO + (r & 7) * 16 + (k & 7) * 2);
}
__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(kNoswLBO >> 4) << 16;
  d |= (uint64_t)(kNoswSBO >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;   // layout type 0 = SWIZZLE_NONE
}
// Instruction descriptor for kind::f16 with bf16 A/B (K-major both), fp32 accumulate, M x N tile.
//   bit 4: D fp32 | bits [7,10): A fmt (1 = bf16) | bits [10,13): B fmt | bits [17,23): N>>3 | bits [24,29): M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of bf16 element (row r, k) inside an SW128 K-major tile of 64 k per row
__host__ __device__ constexpr uint32_t sw128_offset(int r, int k) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + (k & 7) * 2);
}

// ------------------------------------------------------------------ MMA issue (one thread)
// D[tmem] (+)= A[smem] * B[smem]^T
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------ TMEM <-> registers
// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp gets lane (taddr.lane + i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// 16-column / 8-column variants (half-chunk software pipelining of the epilogues)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

// two fp32 -> packed bf16x2 (lo in bits [0,16)), round-to-nearest-even
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// packed fp32x2 arithmetic (sm_100: FADD2 / FFMA2) and mixed-precision FMA (FHFMA: f16 * f16 + f32)
// bias as two bf16 K entries (hi in the low half-word = k 0, lo = bias - hi in k 1): exact to ~2^-17 relative
__device__ __forceinline__ uint32_t pack_bias_hi_lo(float b) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(b);
  const __nv_bfloat16 lo = __float2bfloat16_rn(b - __bfloat162float(hi));
  return (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
}
// ... and with a third term the fp32 value exactly (3 x 8 mantissa bits): k 0..2 = hi, lo, lo2, k 3..7 = 0
__device__ __forceinline__ uint4 pack_bias_3term(float b) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(b);
  const float r1 = b - __bfloat162float(hi);
  const __nv_bfloat16 lo = __float2bfloat16_rn(r1);
  const __nv_bfloat16 lo2 = __float2bfloat16_rn(r1 - __bfloat162float(lo));
  return make_uint4((uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16), (uint32_t)__bfloat16_as_ushort(lo2), 0u, 0u);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long r, x, y;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b.x), "f"(b.y));
  asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y));
  float2 o;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r));
  return o;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long r, x, y, z;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(z) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(y), "l"(z));
  float2 o;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r));
  return o;
}
// acc + h * w with h, w fp16 (16-bit operands) and fp32 accumulate
__device__ __forceinline__ float fma_f16(uint16_t h, uint16_t w, float acc) {
  float r;
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r) : "h"(h), "h"(w), "f"(acc));
  return r;
}
// acc + h with h fp16
__device__ __forceinline__ float add_f16(uint16_t h, float acc) {
  float r;
  asm("add.f32.f16 %0, %1, %2;" : "=f"(r) : "h"(h), "f"(acc));
  return r;
}

__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  unsigned long long r, x, y;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b.x), "f"(b.y));
  asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y));
  float2 o;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r));
  return o;
}

// sin of two radian arguments evaluated entirely on the FMA pipe (packed fp32x2): range reduction to
// r = x/2pi - rint(x/2pi) in [-0.5, 0.5] with the 1.5*2^23 rounding trick, then the odd degree-9 minimax
// polynomial of sin(2 pi r) (max abs error 1.3e-5).  Used for a fraction of every epilogue's sines so
// that the MUFU pipe (16 sines/clk/SM) and the FMA pipe share the transcendental load.
__device__ __forceinline__ float2 poly_sin2(float2 x) {
  const float2 inv2pi = make_float2(0.15915494309189535f, 0.15915494309189535f);
  const float2 magic = make_float2(12582912.0f, 12582912.0f), nmagic = make_float2(-12582912.0f, -12582912.0f);
  const float2 none = make_float2(-1.0f, -1.0f);
  const float2 u = mul2(x, inv2pi);
  const float2 k = add2(add2(u, magic), nmagic);
  const float2 r = fma2(k, none, u);
  const float2 s = mul2(r, r);
  float2 p = fma2(s, make_float2(32.1492805f, 32.1492805f), make_float2(-74.1115570f, -74.1115570f));
  p = fma2(p, s, make_float2(81.2946777f, 81.2946777f));
  p = fma2(p, s, make_float2(-41.3257561f, -41.3257561f));
  p = fma2(p, s, make_float2(6.28293371f, 6.28293371f));
  return mul2(p, r);
}

// fast sine on the MUFU pipe (abs error ~5e-7 for |x| <~ 100; the result is rounded to bf16 anyway)
__device__ __forceinline__ float fast_sin(float x) {
#ifdef STIF_DIAG_NOSIN   // diagnostic builds only (results are WRONG): what the kernels cost without the MUFU work
  return x * 0.5f;
#endif
  float r;
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// High-precision mode: sine with an explicit two-constant Cody-Waite reduction to [-pi, pi] ahead of the MUFU (whose own range
// scaling is only accurate for small arguments): ~6e-7 absolute error for |x| up to a few hundred radians, 7 instructions instead
// of sinf's ~40 (the epilogues of kernels_hp.cu were bound by sinf, not by memory).
__device__ __forceinline__ float reduced_sin(float x) {
  const float k = rintf(x * 0.15915494309189535f);
  float r = fmaf(k, -6.2831854820251465f, x);      // fl32(2 pi)
  r = fmaf(k, 1.7484555314695172e-7f, r);          // fl32(2 pi) - 2 pi
  float y;
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(r));
  return y;
}

}  // namespace tc
}  // namespace stif
