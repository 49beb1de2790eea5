// tc_selftest.cu -- on-device unit checks of the tcgen05 / TMEM / bulk-TMA building blocks.
//
// Three small GEMMs exercise exactly the operand forms the fused kernels rely on:
//   T1  D1[128x64]  = A1[128x64]  * B1[64x64]^T     A, B in SW128 K-major smem ("SS" form)
//   T2  D2[128x256] = A2[128x64]  * B2[256x64]^T    A in TMEM as packed bf16 ("TS" form), N = 256
//   T3  D3[128x128] = A3[128x256] * B3[128x256]^T   TS form, K = 256 from 4 smem K-blocks, A written
//                                                    in place over the accumulator it was derived from
// Results are compared on the host against fp32 references computed from the same bf16-rounded
// operands.  Exposed through stif_selftest() (include/stif_b200.h).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "stif_internal.h"
#include "tc_primitives.cuh"

namespace stif {
namespace {

using namespace tc;

struct SelfTestParams {
  const uint8_t* b1;  // SW128 image, 64 rows x 64 k          (8 KB)
  const uint8_t* b2;  // SW128 image, 256 rows x 64 k         (32 KB)
  const uint8_t* b3;  // 4 K-blocks of (128 rows x 64 k)      (64 KB)
  const float* a1;    // [128,64] fp32
  float* d1;          // [128,64]
  float* d2;          // [128,256]
  float* d3;          // [128,128]
  const float* bias6; // [64] fp32
  float* d6;          // [128,64]
  int* status;        // 0 = ok, else the step that timed out
};

constexpr uint32_t kOnesBytes = 128 * 32, kBiasBytes = 64 * 32;
constexpr uint32_t kB1 = 64 * 128, kB2 = 256 * 128, kB3 = 4 * 128 * 128, kA1 = 128 * 128;

__device__ bool wait_bounded(uint64_t* bar, uint32_t parity) {
  for (int i = 0; i < (1 << 22); ++i)
    if (mbar_try_wait(bar, parity)) return true;
  return false;
}

__global__ void __launch_bounds__(128) tc_selftest_kernel(SelfTestParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA1 = smem;
  uint8_t* sB1 = sA1 + kA1;
  uint8_t* sB2 = sB1 + kB1;
  uint8_t* sB3 = sB2 + kB2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB3 + kB3);  // [0] load, [1] mma
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  uint8_t* sOnes = reinterpret_cast<uint8_t*>(bars) + 64;   // 128 x 16 bf16, interleaved (no swizzle): 4 KB
  uint8_t* sBias = sOnes + kOnesBytes;                        // 64 x 16 bf16, interleaved: 2 KB
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], kB1 + kB2 + kB3);
    bulk_copy_g2s(sB1, p.b1, kB1, &bars[0]);
    bulk_copy_g2s(sB2, p.b2, kB2, &bars[0]);
    bulk_copy_g2s(sB3, p.b3, kB3, &bars[0]);
  }
  // A1: thread = row, 8 chunks of 8 bf16
  {
    const float* src = p.a1 + tid * 64;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint4 v;
      v.x = pack_bf16x2(src[j * 8 + 0], src[j * 8 + 1]);
      v.y = pack_bf16x2(src[j * 8 + 2], src[j * 8 + 3]);
      v.z = pack_bf16x2(src[j * 8 + 4], src[j * 8 + 5]);
      v.w = pack_bf16x2(src[j * 8 + 6], src[j * 8 + 7]);
      *reinterpret_cast<uint4*>(sA1 + sw128_offset(tid, j * 8)) = v;
    }
  }
  // T6 operands: A = [1, 1, 0 ...] per row, B row n = [bf16 hi(bias[n]), bf16 lo(bias[n]), 0 ...]
  {
    uint4 ones = make_uint4(0x3F803F80u, 0u, 0u, 0u), zero = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(sOnes + nosw_offset(tid, 0)) = ones;
    *reinterpret_cast<uint4*>(sOnes + nosw_offset(tid, 8)) = zero;
    if (tid < 64) {
      *reinterpret_cast<uint4*>(sBias + nosw_offset(tid, 0)) = make_uint4(pack_bias_hi_lo(p.bias6[tid]), 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(sBias + nosw_offset(tid, 8)) = zero;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  if (!wait_bounded(&bars[0], 0)) { if (tid == 0) *p.status = 1; goto done; }

  // ---------------- T1: SS, K = 64, N = 64
  if (tid == 0) {
    uint64_t da = make_desc_sw128(smem_u32(sA1)), db = make_desc_sw128(smem_u32(sB1));
    for (int k = 0; k < 4; ++k) umma_ss(tmem, da + 2 * k, db + 2 * k, make_idesc_bf16(128, 64), k > 0);
    umma_commit(&bars[1]);
  }
  if (!wait_bounded(&bars[1], 0)) { if (tid == 0) *p.status = 2; goto done; }
  tc_fence_after();
  {
    uint32_t v[32];
    uint32_t a2[16];
    for (int c = 0; c < 2; ++c) {
      tmem_ld32(lane_base + c * 32, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) p.d1[tid * 64 + c * 32 + j] = __uint_as_float(v[j]);
      // A2 = bf16(0.25 * D1), TMEM columns [256, 288)
      for (int j = 0; j < 16; ++j)
        a2[j] = pack_bf16x2(0.25f * __uint_as_float(v[2 * j]), 0.25f * __uint_as_float(v[2 * j + 1]));
      tmem_st16(lane_base + 256 + c * 16, a2);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  // ---------------- T2: TS, K = 64, N = 256
  if (tid == 0) {
    tc_fence_after();
    uint64_t db = make_desc_sw128(smem_u32(sB2));
    for (int k = 0; k < 4; ++k) umma_ts(tmem, tmem + 256 + 8 * k, db + 2 * k, make_idesc_bf16(128, 256), k > 0);
    umma_commit(&bars[1]);
  }
  if (!wait_bounded(&bars[1], 1)) { if (tid == 0) *p.status = 3; goto done; }
  tc_fence_after();
  {
    uint32_t v[32];
    uint32_t a3[16];
    for (int c = 0; c < 8; ++c) {
      tmem_ld32(lane_base + c * 32, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) p.d2[tid * 256 + c * 32 + j] = __uint_as_float(v[j]);
      // A3 = bf16(0.125 * D2) written IN PLACE over D2: columns [16c, 16c+16)
      for (int j = 0; j < 16; ++j)
        a3[j] = pack_bf16x2(0.125f * __uint_as_float(v[2 * j]), 0.125f * __uint_as_float(v[2 * j + 1]));
      tmem_st16(lane_base + c * 16, a3);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  // ---------------- T3: TS, K = 256 (4 K-blocks of B), N = 128, D at columns [256, 384)
  if (tid == 0) {
    tc_fence_after();
    for (int j = 0; j < 16; ++j) {
      uint64_t db = make_desc_sw128(smem_u32(sB3 + (j >> 2) * (128 * 128))) + 2 * (j & 3);
      umma_ts(tmem + 256, tmem + 8 * j, db, make_idesc_bf16(128, 128), j > 0);
    }
    umma_commit(&bars[1]);
  }
  if (!wait_bounded(&bars[1], 0)) { if (tid == 0) *p.status = 4; goto done; }
  tc_fence_after();
  {
    uint32_t v[32];
    for (int c = 0; c < 4; ++c) {
      tmem_ld32(lane_base + 256 + c * 32, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) p.d3[tid * 128 + c * 32 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  // ---------------- T6: TS K = 64 from A3's first 64 channels x B1, plus the bias as one extra K = 16 SS MMA (interleaved
  // no-swizzle operands: a constant ones tile x a [hi, lo] bias block), D at columns [384, 448)
  if (tid == 0) {
    tc_fence_after();
    uint64_t db = make_desc_sw128(smem_u32(sB1));
    for (int k = 0; k < 4; ++k) umma_ts(tmem + 384, tmem + 8 * k, db + 2 * k, make_idesc_bf16(128, 64), k > 0);
    umma_ss(tmem + 384, make_desc_nosw(smem_u32(sOnes)), make_desc_nosw(smem_u32(sBias)), make_idesc_bf16(128, 64), true);
    umma_commit(&bars[1]);
  }
  if (!wait_bounded(&bars[1], 1)) { if (tid == 0) *p.status = 6; goto done; }
  tc_fence_after();
  {
    uint32_t v[32];
    for (int c = 0; c < 2; ++c) {
      tmem_ld32(lane_base + 384 + c * 32, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) p.d6[tid * 64 + c * 32 + j] = __uint_as_float(v[j]);
    }
  }
done:
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

float bf16_round(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  u = (u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u;
  float r;
  std::memcpy(&r, &u, 4);
  return r;
}
uint16_t bf16_bits(float x) {
  float r = bf16_round(x);
  uint32_t u;
  std::memcpy(&u, &r, 4);
  return (uint16_t)(u >> 16);
}

// rows x K (K multiple of 64) -> K/64 consecutive SW128 K-blocks of (rows x 64)
std::vector<uint8_t> pack_sw128(const std::vector<float>& w, int rows, int K) {
  std::vector<uint8_t> img((size_t)rows * K * 2, 0);
  for (int kb = 0; kb < K / 64; ++kb)
    for (int r = 0; r < rows; ++r)
      for (int k = 0; k < 64; ++k) {
        uint16_t b = bf16_bits(w[(size_t)r * K + kb * 64 + k]);
        std::memcpy(&img[(size_t)kb * rows * 128 + sw128_offset(r, k)], &b, 2);
      }
  return img;
}

double compare(const std::vector<float>& got, const std::vector<float>& ref, double* refmax) {
  double e = 0, m = 0;
  for (size_t i = 0; i < got.size(); ++i) {
    e = std::max(e, (double)std::fabs(got[i] - ref[i]));
    m = std::max(m, (double)std::fabs(ref[i]));
  }
  *refmax = m;
  return e;
}

}  // namespace

int tc_selftest(int device, std::string& report) {
  char line[512];
  auto fail = [&](const char* what, cudaError_t e) {
    snprintf(line, sizeof line, "selftest: %s: %s\n", what, cudaGetErrorString(e));
    report += line;
    return -1;
  };
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail("cudaSetDevice", e);
  uint32_t seed = 12345u;
  auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return ((seed >> 8) & 0xFFFF) / 32768.0f - 1.0f; };
  std::vector<float> a1(128 * 64), b1(64 * 64), b2(256 * 64), b3(128 * 256);
  for (auto& v : a1) v = bf16_round(rnd());
  for (auto& v : b1) v = bf16_round(rnd());
  for (auto& v : b2) v = bf16_round(rnd());
  for (auto& v : b3) v = bf16_round(rnd());
  std::vector<float> bias6(64);
  for (auto& v : bias6) v = 7.0f * rnd();
  auto i1 = pack_sw128(b1, 64, 64), i2 = pack_sw128(b2, 256, 64), i3 = pack_sw128(b3, 128, 256);
  uint8_t *db1, *db2, *db3;
  float *da1, *dd1, *dd2, *dd3;
  int* dstatus;
  float *dbias6, *dd6;
  cudaMalloc(&dbias6, 64 * 4); cudaMalloc(&dd6, 128 * 64 * 4);
  cudaMemcpy(dbias6, bias6.data(), 64 * 4, cudaMemcpyHostToDevice);
  cudaMemset(dd6, 0, 128 * 64 * 4);
  cudaMalloc(&db1, i1.size()); cudaMalloc(&db2, i2.size()); cudaMalloc(&db3, i3.size());
  cudaMalloc(&da1, a1.size() * 4); cudaMalloc(&dd1, 128 * 64 * 4); cudaMalloc(&dd2, 128 * 256 * 4);
  cudaMalloc(&dd3, 128 * 128 * 4); cudaMalloc(&dstatus, 4);
  cudaMemcpy(db1, i1.data(), i1.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(db2, i2.data(), i2.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(db3, i3.data(), i3.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(da1, a1.data(), a1.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dstatus, 0, 4);
  cudaMemset(dd1, 0, 128 * 64 * 4); cudaMemset(dd2, 0, 128 * 256 * 4); cudaMemset(dd3, 0, 128 * 128 * 4);
  SelfTestParams p{db1, db2, db3, da1, dd1, dd2, dd3, dbias6, dd6, dstatus};
  size_t smem = kA1 + kB1 + kB2 + kB3 + 64 + kOnesBytes + kBiasBytes + 1024;
  e = cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail("cudaFuncSetAttribute", e);
  tc_selftest_kernel<<<1, 128, smem>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail("launch", e);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return fail("kernel", e);
  int status = 0;
  std::vector<float> d1(128 * 64), d2(128 * 256), d3(128 * 128), d6(128 * 64);
  cudaMemcpy(d6.data(), dd6, d6.size() * 4, cudaMemcpyDeviceToHost);
  cudaFree(dbias6); cudaFree(dd6);
  cudaMemcpy(&status, dstatus, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(d1.data(), dd1, d1.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(d2.data(), dd2, d2.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(d3.data(), dd3, d3.size() * 4, cudaMemcpyDeviceToHost);
  cudaFree(db1); cudaFree(db2); cudaFree(db3); cudaFree(da1); cudaFree(dd1); cudaFree(dd2); cudaFree(dd3); cudaFree(dstatus);
  int rc = 0;
  if (status != 0) {
    snprintf(line, sizeof line, "selftest: mbarrier wait timed out at step %d\n", status);
    report += line;
    rc = -1;
  }
  // host references (operands exactly as the device rounded them; T2/T3 start from the DEVICE's D1/D2)
  std::vector<float> r1(128 * 64), r2(128 * 256), r2swap(128 * 256), r3(128 * 128);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 64; ++n) {
      double acc = 0;
      for (int k = 0; k < 64; ++k) acc += (double)a1[m * 64 + k] * b1[n * 64 + k];
      r1[m * 64 + n] = (float)acc;
    }
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 256; ++n) {
      double acc = 0, accs = 0;
      for (int k = 0; k < 64; ++k) {
        acc += (double)bf16_round(0.25f * d1[m * 64 + k]) * b2[n * 64 + k];
        accs += (double)bf16_round(0.25f * d1[m * 64 + (k ^ 1)]) * b2[n * 64 + k];  // hypothesis: halves swapped
      }
      r2[m * 256 + n] = (float)acc;
      r2swap[m * 256 + n] = (float)accs;
    }
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 128; ++n) {
      double acc = 0;
      for (int k = 0; k < 256; ++k) acc += (double)bf16_round(0.125f * d2[m * 256 + k]) * b3[n * 256 + k];
      r3[m * 128 + n] = (float)acc;
    }
  std::vector<float> r6(128 * 64);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 64; ++n) {
      double acc = bias6[n];
      for (int k = 0; k < 64; ++k) acc += (double)bf16_round(0.125f * d2[m * 256 + k]) * b1[n * 64 + k];
      r6[m * 64 + n] = (float)acc;
    }
  double m6;
  const double e6 = compare(d6, r6, &m6);
  double m1, m2, m2s, m3;
  double e1 = compare(d1, r1, &m1), e2 = compare(d2, r2, &m2), e2s = compare(d2, r2swap, &m2s), e3 = compare(d3, r3, &m3);
  snprintf(line, sizeof line, "T1 SS  k64 n64 : max_abs_err %.3e (ref max %.3e)\n", e1, m1); report += line;
  snprintf(line, sizeof line, "T2 TS  k64 n256: max_abs_err %.3e (ref max %.3e) [swapped-halves hypothesis err %.3e]\n", e2, m2, e2s);
  report += line;
  snprintf(line, sizeof line, "T3 TS k256 n128: max_abs_err %.3e (ref max %.3e)\n", e3, m3); report += line;
  snprintf(line, sizeof line, "T6 TS k64 n64 + bias as a K=16 no-swizzle SS MMA: max_abs_err %.3e (ref max %.3e)\n", e6, m6); report += line;
  if (!(e6 < 2e-4 * std::max(1.0, m6))) rc = -1;
  snprintf(line, sizeof line, "samples D1[0][0..3] got %.5f %.5f %.5f %.5f ref %.5f %.5f %.5f %.5f\n", d1[0], d1[1], d1[2],
           d1[3], r1[0], r1[1], r1[2], r1[3]);
  report += line;
  snprintf(line, sizeof line, "samples D1[77][60..63] got %.5f %.5f %.5f %.5f ref %.5f %.5f %.5f %.5f\n", d1[77 * 64 + 60],
           d1[77 * 64 + 61], d1[77 * 64 + 62], d1[77 * 64 + 63], r1[77 * 64 + 60], r1[77 * 64 + 61], r1[77 * 64 + 62],
           r1[77 * 64 + 63]);
  report += line;
  snprintf(line, sizeof line, "samples D2[5][0..3] got %.5f %.5f %.5f %.5f ref %.5f %.5f %.5f %.5f\n", d2[5 * 256], d2[5 * 256 + 1],
           d2[5 * 256 + 2], d2[5 * 256 + 3], r2[5 * 256], r2[5 * 256 + 1], r2[5 * 256 + 2], r2[5 * 256 + 3]);
  report += line;
  snprintf(line, sizeof line, "samples D3[9][0..3] got %.5f %.5f %.5f %.5f ref %.5f %.5f %.5f %.5f\n", d3[9 * 128], d3[9 * 128 + 1],
           d3[9 * 128 + 2], d3[9 * 128 + 3], r3[9 * 128], r3[9 * 128 + 1], r3[9 * 128 + 2], r3[9 * 128 + 3]);
  report += line;
  if (!(e1 < 1e-3 * std::max(1.0, m1)) || !(e2 < 1e-3 * std::max(1.0, m2)) || !(e3 < 1e-3 * std::max(1.0, m3))) rc = -1;
  if (hp_selftest(report) != 0) rc = -1;   // T4 / T5: the split-bf16 GEMM of the high-precision mode (kernels_hp.cu)
  report += rc == 0 ? "selftest: PASS\n" : "selftest: FAIL\n";
  return rc;
}

}  // namespace stif
