// kernels_fp32.cu -- fp32 FMA-pipe path of the STIF query decoder (STIF_MODE_FP32) and the
// latent-projection kernel shared with the bf16 path.
//
// This is the high-precision mode (RGB within 1e-4 of the reference): the hoisted formulation of
// DESIGN.md section 3 evaluated layer by layer in fp32 with accurate sinf.  Activations round-trip
// through the caller's workspace in query chunks; it is the parity anchor for the fused tcgen05
// kernels, not the throughput path.
//
// Reference semantics: LunaTokis.decoding, codes/models/modules/Sakuya_arch_test.py:364-459.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "stif_internal.h"
#include "tc_primitives.cuh"

namespace stif {
namespace {

struct GemmArgs {
  const float* A;    // A(m,k) = A[m*sam + k*sak]               for k <  ksplit
  const float* A2;   // A(m,k) = A2[m*sam + (k-ksplit)*sak]     for k >= ksplit (may be null)
  long sam, sak;
  int ksplit;
  const float* Wt;   // [N,K] row-major
  const float* bias; // [N] or null
  void* C;           // C(m,n) at C[m*scm + n*scn]
  long scm, scn;
  long M;
  int N, K;
  int act;           // 0 = identity, 1 = sin
  int out_half;      // store __half instead of float
};

constexpr int BM = 128, BN = 64, BK = 16;

// C = act(A W^T + b).  256 threads, 128x64 tile, each thread 8 rows x 4 cols.
template <bool A_MCONTIG>
__global__ void __launch_bounds__(256) sgemm_bias_act_kernel(GemmArgs g) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long m0 = (long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < g.K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int idx = tid + i * 256, m, k;
      if (A_MCONTIG) { m = idx % BM; k = idx / BM; } else { k = idx % BK; m = idx / BK; }
      long gm = m0 + m;
      int gk = k0 + k;
      float v = 0.f;
      if (gm < g.M && gk < g.K)
        v = (gk < g.ksplit) ? g.A[gm * g.sam + (long)gk * g.sak] : g.A2[gm * g.sam + (long)(gk - g.ksplit) * g.sak];
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + i * 256, k = idx % BK, n = idx / BK;
      int gn = n0 + n, gk = k0 + k;
      Ws[k][n] = (gn < g.N && gk < g.K) ? g.Wt[(long)gn * g.K + gk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], w[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[kk][ty * 8 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long gm = m0 + ty * 8 + i;
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= g.N) continue;
      float v = acc[i][j] + (g.bias ? g.bias[gn] : 0.f);
      if (g.act == 1) v = sinf(v);
      long o = gm * g.scm + (long)gn * g.scn;
      if (g.out_half) reinterpret_cast<__half*>(g.C)[o] = __float2half_rn(v);
      else reinterpret_cast<float*>(g.C)[o] = v;
    }
  }
}

cudaError_t launch_gemm(const LaunchCtx& cx, const GemmArgs& g, bool a_mcontig) {
  if (g.M <= 0) return cudaSuccess;
  dim3 grid((unsigned)((g.M + BM - 1) / BM), (unsigned)((g.N + BN - 1) / BN));
  if (a_mcontig) sgemm_bias_act_kernel<true><<<grid, 256, 0, cx.stream>>>(g);
  else sgemm_bias_act_kernel<false><<<grid, 256, 0, cx.stream>>>(g);
  ++*cx.launch_counter;
  return cudaGetLastError();
}

GemmArgs dense(const float* A, int K, const float* Wt, const float* bias, void* C, long ldc, long M, int N, int act) {
  GemmArgs g{};
  g.A = A; g.A2 = nullptr; g.sam = K; g.sak = 1; g.ksplit = K;
  g.Wt = Wt; g.bias = bias; g.C = C; g.scm = ldc; g.scn = 1; g.M = M; g.N = N; g.K = K; g.act = act; g.out_half = 0;
  return g;
}

// The 256 -> 4 (flow) and 256 -> 3 (RGB) output layers: C(m, n) = A[m, :] . W[n, :] + b[n], one warp per row of A (a 1 KB
// coalesced read; the SGEMM tile wasted 60 of its 64 columns on them and ran at a tenth of the memory rate).
template <int NOUT>
__global__ void __launch_bounds__(256) out_layer_kernel(const float* __restrict__ A, const float* __restrict__ Wt, const float* __restrict__ bias,
                                                        float* __restrict__ C, long scm, long scn, long M) {
  const int lane = threadIdx.x & 31;
  float w[NOUT][8];
#pragma unroll
  for (int n = 0; n < NOUT; ++n) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(Wt + n * 256) + lane), b = __ldg(reinterpret_cast<const float4*>(Wt + n * 256) + 32 + lane);
    w[n][0] = a.x; w[n][1] = a.y; w[n][2] = a.z; w[n][3] = a.w; w[n][4] = b.x; w[n][5] = b.y; w[n][6] = b.z; w[n][7] = b.w;
  }
  for (long m = (long)blockIdx.x * 8 + (threadIdx.x >> 5); m < M; m += (long)gridDim.x * 8) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(A + m * 256) + lane), b = __ldg(reinterpret_cast<const float4*>(A + m * 256) + 32 + lane);
    float s[NOUT];
#pragma unroll
    for (int n = 0; n < NOUT; ++n) {
      float t = a.x * w[n][0];
      t = fmaf(a.y, w[n][1], t); t = fmaf(a.z, w[n][2], t); t = fmaf(a.w, w[n][3], t);
      t = fmaf(b.x, w[n][4], t); t = fmaf(b.y, w[n][5], t); t = fmaf(b.z, w[n][6], t); t = fmaf(b.w, w[n][7], t);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      s[n] = t;
    }
    if (lane < NOUT) {
      float v = s[0];
#pragma unroll
      for (int n = 1; n < NOUT; ++n) v = lane == n ? s[n] : v;
      C[m * scm + (long)lane * scn] = v + bias[lane];
    }
  }
}
template <int NOUT>
cudaError_t launch_out_layer(const LaunchCtx& cx, const float* A, const float* Wt, const float* bias, float* C, long scm, long scn, long M) {
  if (M <= 0) return cudaSuccess;
  const unsigned grid = (unsigned)std::min<long>((M + 7) / 8, (long)cx.num_sms * 16);
  out_layer_kernel<NOUT><<<grid, 256, 0, cx.stream>>>(A, Wt, bias, C, scm, scn, M);
  ++*cx.launch_counter;
  return cudaGetLastError();
}

struct Vec64 { float v[64]; };

// Stage A first layer (hoisted): h0 = sin(TA[iy,ix] + rel_y*w_ry + rel_x*w_rx + (w_t t + b))
// reference: Sakuya_arch_test.py:382-400 (nearest gathers, rel_coord, pe_coord, first SineLayer).
// (first-layer kernels: thread = (query, 4 consecutive channels) -- 16 threads share a query's index / tap arithmetic instead of
// 64, every table access is one 16-byte load; the sine is tc::reduced_sin as in the dense layers' epilogues)
// raster index -> (row, column): rasters have at most 2^30 pixels (check_shape), so one 32-bit division instead of two 64-bit ones
// (which were most of these kernels' instructions)
__device__ __forceinline__ void raster_pos(long q, int WW, int& jy, int& jx) {
  const uint32_t qq = (uint32_t)q, w = (uint32_t)WW, y = qq / w;
  jy = (int)y;
  jx = (int)(qq - y * w);
}
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void fma4(float w, const float4 v, float4& s) {
  s.x = fmaf(w, v.x, s.x); s.y = fmaf(w, v.y, s.y); s.z = fmaf(w, v.z, s.z); s.w = fmaf(w, v.w, s.w);
}
__device__ __forceinline__ float4 sin4(const float4 v) {
  return make_float4(tc::reduced_sin(v.x), tc::reduced_sin(v.y), tc::reduced_sin(v.z), tc::reduced_sin(v.w));
}
__global__ void stage_a_first_layer(const float* __restrict__ tab, Geometry g, const float* __restrict__ a_rel, Vec64 cst,
                                    long q0, long n, float* __restrict__ out) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 16) return;
  int c = (int)(i & 15) * 4;
  long q = q0 + (i >> 4);
  int jy, jx;
  raster_pos(q, g.WW, jy, jx);
  int iy = g.y.idx[jy], ix = g.x.idx[jx];
  float4 ta = make_float4(0.f, 0.f, 0.f, 0.f);
  if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) ta = ld4(tab + ((long)iy * g.W + ix) * 256 + c);
  const float ry = g.y.rel[jy], rx = g.x.rel[jx];
  const float4 r0 = ld4(a_rel + c * 2), r1 = ld4(a_rel + c * 2 + 4);   // [c][2]: (w_ry, w_rx) pairs
  float4 v;
  v.x = ta.x + ry * r0.x + rx * r0.y + cst.v[c];
  v.y = ta.y + ry * r0.z + rx * r0.w + cst.v[c + 1];
  v.z = ta.z + ry * r1.x + rx * r1.y + cst.v[c + 2];
  v.w = ta.w + ry * r1.z + rx * r1.w + cst.v[c + 3];
  reinterpret_cast<float4*>(out)[i] = sin4(v);
}

// Stage B first layer (hoisted): f0 = sin(F + bilinear(TB; query position) + (w_t t + b))
// reference: Sakuya_arch_test.py:406-419.
// f_table != null (local-ensemble pass): F is gathered at the nearest HR pixel of the SHIFTED coordinate
// (Sakuya_arch_test.py:1026-1029) from the whole-slab table instead of being the query's own value.
// pixel-centre query coordinate as make_coord builds it (Sakuya_arch_test.py:1233-1248 + clamp :373): two separately
// rounded fp32 operations on double-derived constants
__device__ __forceinline__ float query_coord(int j, int n) {
  const float r0 = (float)(-1.0 + 1.0 / (double)n), step = (float)(2.0 / (double)n);
  return fminf(fmaxf(__fadd_rn(r0, __fmul_rn(step, (float)j)), kClampLo), kClampHi);
}

__global__ void stage_b_first_layer(const float* __restrict__ tab, Geometry g, Vec64 cst, long q0, long n,
                                    float* __restrict__ f_inout, const float* __restrict__ f_table,
                                    const float* __restrict__ utab) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 16) return;
  int c = (int)(i & 15) * 4;
  long q = q0 + (i >> 4);
  int jy, jx;
  raster_pos(q, g.WW, jy, jx);
  Taps tp = make_taps_tables(g, jy, jx);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) fma4(tp.w[k], ld4(tab + (long)tp.off[k] * 256 + 64 + c), s);
  if (utab) {   // decoding_test: frames sampled bilinearly from the x4-upsampled pair at the query position (:520-523)
    const Taps up = make_taps(query_coord(jy, g.HH), query_coord(jx, g.WW), 4 * g.H, 4 * g.W);
#pragma unroll
    for (int k = 0; k < 4; ++k) fma4(up.w[k], ld4(utab + (long)up.off[k] * 192 + c), s);
  }
  float4 f = reinterpret_cast<float4*>(f_inout)[i];
  if (f_table) {
    const int hy = g.y.hidx[jy], hx = g.x.hidx[jx];
    f = (hy >= 0 && hy < g.HH && hx >= 0 && hx < g.WW) ? ld4(f_table + ((long)hy * g.WW + hx) * 64 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  reinterpret_cast<float4*>(f_inout)[i] =
      sin4(make_float4(f.x + s.x + cst.v[c], f.y + s.y + cst.v[c + 1], f.z + s.z + cst.v[c + 2], f.w + s.w + cst.v[c + 3]));
}

// ret = ret + pred_k * (area_{3-k} / tot_area), every operation separately rounded as the reference's ATen kernels do
// (Sakuya_arch_test.py:1011-1012, 1077-1084).  The weights are recomputed here from the per-axis rel tables.
__global__ void ensemble_blend(const float* __restrict__ pred, float* __restrict__ out, int WW, long Q, AxisTables ym, AxisTables yp,
                               AxisTables xm, AxisTables xp, int k) {
  long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  const int jy = (int)(q / WW), jx = (int)(q % WW);
  const float ry[2] = {ym.rel[jy], yp.rel[jy]}, rx[2] = {xm.rel[jx], xp.rel[jx]};
  float a[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = __fadd_rn(fabsf(__fmul_rn(ry[i >> 1], rx[i & 1])), 1e-9f);
  const float tot = __fadd_rn(__fadd_rn(__fadd_rn(a[0], a[1]), a[2]), a[3]);
  const float w = __fdiv_rn(a[3 - k], tot);
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float prev = k == 0 ? 0.f : out[ch * Q + q];
    out[ch * Q + q] = __fadd_rn(prev, __fmul_rn(pred[ch * Q + q], w));
  }
}

// Stage C+D + encode_imnet first layer (hoisted):
// e0 = sin(bilin(Q1;g1) + bilin(Q2;g2) + bilin(TE1;g1) + bilin(TE2;g2) + (w_t t + b))
// reference: warplayer.py:25-39, Sakuya_arch_test.py:424-456.
__global__ void stage_e_first_layer(const float* __restrict__ tab, const float* __restrict__ qtab,
                                    const float* __restrict__ flow, Geometry g, Vec64 cst, long q0, long n,
                                    int band_lo, int band_hi, int* __restrict__ flag, float* __restrict__ out,
                                    const float* __restrict__ utab) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 16) return;
  int c = (int)(i & 15) * 4;
  long q = q0 + (i >> 4);
  int jy, jx;
  raster_pos(q, g.WW, jy, jx);
  float4 fl = reinterpret_cast<const float4*>(flow)[q];
  float4 s = make_float4(cst.v[c], cst.v[c + 1], cst.v[c + 2], cst.v[c + 3]);
#pragma unroll
  for (int wv = 0; wv < 2; ++wv) {
    float gy, gx;
    warp_position(g, jy, jx, wv == 0 ? fl.x : fl.z, wv == 0 ? fl.y : fl.w, gy, gx);
    Taps hr = make_taps(gy, gx, g.HH, g.WW);
    Taps lr = make_taps(gy, gx, g.H, g.W);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // A zero-weight tap (out-of-grid taps are redirected to row 0; exact-integer positions give in-grid taps with w == 0)
      // may point at a Q-table row this row-band launch never wrote: 0 x stale NaN bits is NaN, so the load is skipped,
      // as the tensor-core path does (kernels_tc.cu, k2_gather_taps).
      if (hr.w[k] == 0.f) continue;
      const int row = hr.off[k] / g.WW;
      if (row < band_lo || row >= band_hi) { if (c == 0) atomicOr(flag, 1); continue; }
      fma4(hr.w[k], ld4(qtab + (long)hr.off[k] * 128 + wv * 64 + c), s);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) fma4(lr.w[k], ld4(tab + (long)lr.off[k] * 256 + 128 + wv * 64 + c), s);
    if (utab) {   // decoding_test: frames at the warped position come from the x4-upsampled pair (:548-551, :562-565)
      const Taps up = make_taps(gy, gx, 4 * g.H, 4 * g.W);
#pragma unroll
      for (int k = 0; k < 4; ++k) fma4(up.w[k], ld4(utab + (long)up.off[k] * 192 + 64 + wv * 64 + c), s);
    }
  }
  reinterpret_cast<float4*>(out)[i] = sin4(s);
}

// The same without the upsampled-frame terms (every decode except decoding_test), with the tap arithmetic shared: the 16 lanes of a
// query each derive ONE of its 16 taps (2 warps x {HR, LR} x 4 corners) and the (offset, weight) pairs travel by shuffle, in the
// accumulation order of the kernel above (bit-identical result, ~3x fewer instructions).
__global__ void stage_e_first_layer_shfl(const float* __restrict__ tab, const float* __restrict__ qtab,
                                         const float* __restrict__ flow, Geometry g, Vec64 cst, long q0, long n,
                                         int band_lo, int band_hi, int* __restrict__ flag, float* __restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n * 16;                     // (whole 16-lane groups are live or not: n * 16 is a multiple of 16)
  const int l16 = threadIdx.x & 15, c = l16 * 4;
  const long q = q0 + (live ? (i >> 4) : 0);
  int jy, jx;
  raster_pos(q, g.WW, jy, jx);
  const float4 fl = reinterpret_cast<const float4*>(flow)[q];
  const int wv = l16 >> 3, kind = (l16 >> 2) & 1, k = l16 & 3;
  float gy, gx;
  warp_position(g, jy, jx, wv == 0 ? fl.x : fl.z, wv == 0 ? fl.y : fl.w, gy, gx);
  const Taps t = kind ? make_taps(gy, gx, g.H, g.W) : make_taps(gy, gx, g.HH, g.WW);
  int my_off = k == 0 ? t.off[0] : k == 1 ? t.off[1] : k == 2 ? t.off[2] : t.off[3];
  float my_w = k == 0 ? t.w[0] : k == 1 ? t.w[1] : k == 2 ? t.w[2] : t.w[3];
  if (kind == 0 && my_w != 0.f) {                   // HR tap: must lie in the rows stage A+B has produced (see the kernel above)
    const int row = my_off / g.WW;
    if (row < band_lo || row >= band_hi) { if (live) atomicOr(flag, 1); my_w = 0.f; }
  }
  float4 s = make_float4(cst.v[c], cst.v[c + 1], cst.v[c + 2], cst.v[c + 3]);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int off = __shfl_sync(0xffffffffu, my_off, j, 16);
    const float w = __shfl_sync(0xffffffffu, my_w, j, 16);
    const int jwv = j >> 3, jkind = (j >> 2) & 1;
    if (jkind == 0) {
      if (w != 0.f) fma4(w, ld4(qtab + (long)off * 128 + jwv * 64 + c), s);
    } else {
      fma4(w, ld4(tab + (long)off * 256 + 128 + jwv * 64 + c), s);
    }
  }
  if (live) reinterpret_cast<float4*>(out)[i] = sin4(s);
}

Vec64 time_constant(const std::vector<float>& wt, const std::vector<float>& b, float t) {
  Vec64 v;
  for (int c = 0; c < 64; ++c) v.v[c] = wt[c] * t + b[c];
  return v;
}

}  // namespace

// tab[texel, 0:256] = w_tab [256,198] . [latent(192); frames(6)][texel]      (t-independent)
// Output conversion of the reference's caller (custom_video_test.py:102): `(img.clamp(0,1) * 255).astype(np.uint8)` on the
// HWC-permuted frame, i.e. fp32 clamp, fp32 multiply, truncation.  planar [3, Q] fp32 -> interleaved [Q, 3] uint8, rows
// [row_begin, row_end) of the raster.  Four pixels per thread: 3 coalesced plane reads, one 12-byte run written as words.
__global__ void rgb_to_u8_hwc_kernel(const float* __restrict__ rgb, uint8_t* __restrict__ out, long Q, long q_begin, long q_end) {
  const long q0 = (q_begin & ~3L) + ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (q0 >= q_end) return;
  uint8_t v[12];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long q = min(max(q0 + i, q_begin), q_end - 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) v[3 * i + c] = (uint8_t)(fminf(fmaxf(__ldg(rgb + c * Q + q), 0.f), 1.f) * 255.f);
  }
  // word stores need 4-byte alignment of the absolute address: the 12-byte run starts at a multiple of 12 bytes inside
  // the slab, but the slab itself starts at slab_index * 3 * HH * WW bytes, which is odd-sized for odd rasters
  if (q0 >= q_begin && q0 + 3 < q_end && (reinterpret_cast<uintptr_t>(out + q0 * 3) & 3) == 0) {
    uint32_t* o = reinterpret_cast<uint32_t*>(out + q0 * 3);
#pragma unroll
    for (int w = 0; w < 3; ++w) o[w] = v[4 * w] | (v[4 * w + 1] << 8) | (v[4 * w + 2] << 16) | ((uint32_t)v[4 * w + 3] << 24);
  } else {
    for (int i = 0; i < 4; ++i)
      if (q0 + i >= q_begin && q0 + i < q_end)
        for (int c = 0; c < 3; ++c) out[(q0 + i) * 3 + c] = v[3 * i + c];
  }
}

cudaError_t rgb_to_u8_hwc(const LaunchCtx& cx, const float* rgb_planar, uint8_t* out_hwc, int HH, int WW, int row_begin, int row_end) {
  const long Q = (long)HH * WW, q_begin = (long)row_begin * WW, q_end = (long)row_end * WW;
  if (q_end <= q_begin) return cudaSuccess;
  const long groups = (q_end - (q_begin & ~3L) + 3) / 4;
  rgb_to_u8_hwc_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, cx.stream>>>(rgb_planar, out_hwc, Q, q_begin, q_end);
  ++*cx.launch_counter;
  return cudaGetLastError();
}

// decoding_test variant: one thread per (texel of the 4H x 4W grid, output channel).  The six upsampled frame values are
// ATen's upsample_bilinear2d (align_corners=False, scale 1/4): src = max(0, 0.25 (dst + 0.5) - 0.5), i1 = min(i0 + 1, n - 1),
// value = wy0 (wx0 a + wx1 b) + wy1 (wx0 c + wx1 d), all in fp32 without contraction.
__global__ void project_frames_up4_kernel(const float* __restrict__ frames, int H, int W, const float* __restrict__ w_up,
                                          float* __restrict__ utab) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long n = (long)16 * H * W;
  if (i >= n * 192) return;
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
    const float v = __fadd_rn(__fmul_rn(wy0, top), __fmul_rn(ly, bot));
    s = fmaf(w_up[c * 6 + ch], v, s);
  }
  utab[i] = s;
}

// ---- decoding_test on the tensor-core kernels at sizes other than x4 (Sakuya_arch_test.py:461-598) --------------------
// utab [4H*4W, 192] fp16 = UB | UE1 | UE2, the frame terms of the three hoisted first layers on the x4-upsampled grid.
// Thread = (query, 16-byte chunk of 8 channels); loads are 16 bytes, arithmetic fp32.
__device__ __forceinline__ void blend8(const __half* __restrict__ row0, int chunk, const Taps& t, float (&acc)[8]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (t.w[k] == 0.f) continue;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(row0 + (long)t.off[k] * 192) + chunk);
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __half22float2(h[e]);
      acc[2 * e] = fmaf(t.w[k], f.x, acc[2 * e]);
      acc[2 * e + 1] = fmaf(t.w[k], f.y, acc[2 * e + 1]);
    }
  }
}
__device__ __forceinline__ uint4 pack8(const float (&a)[8]) {
  __half2 h[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(a[2 * e], a[2 * e + 1]);
  return *reinterpret_cast<uint4*>(h);
}
// stage B's term (:520-523): uq[q, 0:64] = bilinear(UB; query position on the 4H x 4W grid); uq[q, 64:192] = 0 (the x4 kernel
// that consumes this table folds columns 64..191 into the Q planes, which is only valid when the two grids coincide)
__global__ void resample_ub_half_kernel(const __half* __restrict__ utab, Geometry g, __half* __restrict__ uq) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long Q = (long)g.HH * g.WW;
  if (i >= Q * 24) return;
  const long q = i / 24;
  const int chunk = (int)(i - q * 24);
  uint4 o = make_uint4(0, 0, 0, 0);
  if (chunk < 8) {
    const int jy = (int)(q / g.WW), jx = (int)(q - (long)jy * g.WW);
    const Taps up = make_taps(query_coord(jy, g.HH), query_coord(jx, g.WW), 4 * g.H, 4 * g.W);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    blend8(utab, chunk, up, acc);
    o = pack8(acc);
  }
  reinterpret_cast<uint4*>(uq + q * 192)[chunk] = o;
}
// stage D's terms (:548-551, :562-565): uadd[q] = bilinear(UE1; g1) + bilinear(UE2; g2) at the flow-warped positions
__global__ void warp_u_terms_half_kernel(const __half* __restrict__ utab, const float* __restrict__ flow, Geometry g, long q_begin,
                                         long q_end, __half* __restrict__ uadd) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long q = q_begin + (i >> 3);
  if (q >= q_end) return;
  const int chunk = (int)(i & 7);
  const int jy = (int)(q / g.WW), jx = (int)(q - (long)jy * g.WW);
  const float4 fl = __ldg(reinterpret_cast<const float4*>(flow) + q);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int wv = 0; wv < 2; ++wv) {
    float gy, gx;
    warp_position(g, jy, jx, wv == 0 ? fl.x : fl.z, wv == 0 ? fl.y : fl.w, gy, gx);
    const Taps up = make_taps(gy, gx, 4 * g.H, 4 * g.W);
    blend8(utab, 8 + 8 * wv + chunk, up, acc);
  }
  reinterpret_cast<uint4*>(uadd + q * 64)[chunk] = pack8(acc);
}

cudaError_t resample_ub_tc(const LaunchCtx& cx, const void* utab, const Geometry& geo, void* uq) {
  const long n = (long)geo.HH * geo.WW * 24;
  resample_ub_half_kernel<<<(unsigned)((n + 255) / 256), 256, 0, cx.stream>>>(reinterpret_cast<const __half*>(utab), geo,
                                                                               reinterpret_cast<__half*>(uq));
  ++*cx.launch_counter;
  return cudaGetLastError();
}
cudaError_t warp_u_terms_tc(const LaunchCtx& cx, const void* utab, const float* flow, const Geometry& geo, int row_begin, int row_end,
                            void* uadd) {
  const long q0 = (long)row_begin * geo.WW, q1 = (long)row_end * geo.WW, n = (q1 - q0) * 8;
  if (n <= 0) return cudaSuccess;
  warp_u_terms_half_kernel<<<(unsigned)((n + 255) / 256), 256, 0, cx.stream>>>(reinterpret_cast<const __half*>(utab), flow, geo, q0, q1,
                                                                                reinterpret_cast<__half*>(uadd));
  ++*cx.launch_counter;
  return cudaGetLastError();
}

cudaError_t project_frames_up4(const LaunchCtx& cx, const DeviceWeights32& w, const float* frames6, int H, int W, float* utab) {
  const long n = (long)16 * H * W * 192;
  project_frames_up4_kernel<<<(unsigned)((n + 255) / 256), 256, 0, cx.stream>>>(frames6, H, W, w.w_up, utab);
  ++*cx.launch_counter;
  return cudaGetLastError();
}

// ---- the projection K0 of the fp32 mode --------------------------------------------------------------------------------------
// tab[m, n] = sum_k X[k, m] W[n, k] with X = [latent(192); frames(6)] in the reference's channel-major layout (row stride HW) and
// W [256, 198]: exact fp32 FMAs in ascending k (the same rounding sequence as sgemm_bias_act_kernel, so the tables are bit-identical
// to the anchor's), as a register-tiled SGEMM: 128 x 128 tile per 256-thread block, 8 x 8 outputs per thread in two 4-wide strips per
// axis (conflict-free LDS.128), K in steps of 8 with the next step's global loads in flight under the FMAs.  ~3x the generic kernel.
template <bool VEC>
__global__ void __launch_bounds__(256) k0_sgemm_kernel(const float* __restrict__ lat, const float* __restrict__ frames,
                                                       const float* __restrict__ W, float* __restrict__ C, long HW) {
  constexpr int K = 198, KT = 8, NK = (K + KT - 1) / KT;
  __shared__ __align__(16) float As[2][KT][128];
  __shared__ __align__(16) float Bs[2][KT][128];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const long m0 = (long)blockIdx.x * 128;
  const int n0 = blockIdx.y * 128;
  const int ak = t >> 5, am = (t & 31) * 4;          // A: one float4 (4 texels of channel k0 + ak) per thread
  const int bn = t >> 1, bk = (t & 1) * 4;           // B: W[n0 + bn][k0 + bk .. +3]
  auto load_a = [&](int k0) {
    const int k = k0 + ak;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < K) {
      const float* src = (k < 192 ? lat + (long)k * HW : frames + (long)(k - 192) * HW) + m0 + am;
      if (VEC) {
        if (m0 + am + 3 < HW) v = __ldg(reinterpret_cast<const float4*>(src));
        else {
          if (m0 + am < HW) v.x = __ldg(src);
          if (m0 + am + 1 < HW) v.y = __ldg(src + 1);
          if (m0 + am + 2 < HW) v.z = __ldg(src + 2);
        }
      } else {
        if (m0 + am < HW) v.x = __ldg(src);
        if (m0 + am + 1 < HW) v.y = __ldg(src + 1);
        if (m0 + am + 2 < HW) v.z = __ldg(src + 2);
        if (m0 + am + 3 < HW) v.w = __ldg(src + 3);
      }
    }
    return v;
  };
  auto load_b = [&](int k0, float (&w)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = (k0 + bk + j < K) ? __ldg(W + (long)(n0 + bn) * K + k0 + bk + j) : 0.f;
  };
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float4 a_next = load_a(0);
  float b_next[4];
  load_b(0, b_next);
  for (int step = 0; step < NK; ++step) {
    const int buf = step & 1;
    *reinterpret_cast<float4*>(&As[buf][ak][am]) = a_next;
#pragma unroll
    for (int j = 0; j < 4; ++j) Bs[buf][bk + j][bn] = b_next[j];
    __syncthreads();                                   // (two buffers: the stores of step s + 1 cannot overtake the reads of step s - 1's ... see below)
    if (step + 1 < NK) {
      a_next = load_a((step + 1) * KT);
      load_b((step + 1) * KT, b_next);
    }
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]), a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]), b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    // a block-wide barrier per step is enough with two buffers: buffer `buf` is next written at step + 2, after the barrier of step + 1,
    // which every thread reaches only after finishing the reads above
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= HW) continue;
    float* dst = C + m * 256 + n0;
    *reinterpret_cast<float4*>(dst + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    *reinterpret_cast<float4*>(dst + 64 + tx * 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
  }
}

// [latent(192); frames(6)] for texels [m0, m0 + n), channel-major in the reference's layout, as row-major [n, 256] rows (columns 198..255
// zero): what the tensor-core GEMM reads.  32 x 32 tiles through shared memory, coalesced both ways.
__global__ void pack_latent_rows_kernel(const float* __restrict__ latent, const float* __restrict__ frames, long HW, long m0, long n,
                                        float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const long m = (long)blockIdx.x * 32 + tx;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ch = blockIdx.y * 32 + ty + 8 * i;
    float v = 0.f;
    if (m < n) {
      if (ch < 192) v = __ldg(latent + (long)ch * HW + m0 + m);
      else if (ch < 198) v = __ldg(frames + (long)(ch - 192) * HW + m0 + m);
    }
    tile[ty + 8 * i][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long row = (long)blockIdx.x * 32 + ty + 8 * i;
    if (row < n) out[row * 256 + blockIdx.y * 32 + tx] = tile[tx][ty + 8 * i];
  }
}

cudaError_t project_latent(const LaunchCtx& cx, const DeviceWeights32& w, const float* latent192, const float* frames6,
                           int H, int W, void* tab, bool tab_half, bool test_variant, float* scratch, size_t scratch_rows) {
  const long HW = (long)H * W;
  const HpLayer* L = hp_layer(w.hp, test_variant ? HP_K0T : HP_K0);
  // STIF_HP_K0=1 (opt-in): the projection on the split-bf16 tensor-core GEMM, 1.0 -> 0.15 ms per 270 x 480 pair.  Off by default: the
  // projected tables ARE the large part of every first-layer sine argument, and the split's ~2^-17 relative error per product shows
  // there first (RGB 1.05e-3 instead of 1.9e-4 in the 50-radian stress case of test_bf16_error_envelope_vs_sine_argument_scale, with
  // no measurable difference inside the envelope); the exact fp32 SGEMM keeps this mode's margin where it is needed.
  static const bool hp_k0 = getenv("STIF_HP_K0") && atoi(getenv("STIF_HP_K0")) != 0;
  if (hp_k0 && L && scratch && scratch_rows >= 128 && !tab_half) {
    for (long m0 = 0; m0 < HW; m0 += (long)scratch_rows) {
      const long n = std::min<long>((long)scratch_rows, HW - m0);
      pack_latent_rows_kernel<<<dim3((unsigned)((n + 31) / 32), 8), dim3(32, 8), 0, cx.stream>>>(latent192, frames6, HW, m0, n, scratch);
      ++*cx.launch_counter;
      if (cudaError_t e = cudaGetLastError()) return e;
      if (cudaError_t e = hp_gemm(cx, *L, 0, 256, scratch, nullptr, (float*)tab + m0 * 256, 256, n, 0)) return e;
    }
    return cudaSuccess;
  }
  if (!tab_half && w.hp) {   // (the SIMT anchor, STIF_FP32_SIMT=1, keeps the generic kernel below: same bits, a third of the speed)
    const dim3 grid((unsigned)((HW + 127) / 128), 2);
    const float* Wt = test_variant ? w.w_tab_lat : w.w_tab;
    const bool vec = HW % 4 == 0 && (reinterpret_cast<uintptr_t>(latent192) & 15) == 0 && (reinterpret_cast<uintptr_t>(frames6) & 15) == 0;
    if (vec) k0_sgemm_kernel<true><<<grid, 256, 0, cx.stream>>>(latent192, frames6, Wt, (float*)tab, HW);
    else k0_sgemm_kernel<false><<<grid, 256, 0, cx.stream>>>(latent192, frames6, Wt, (float*)tab, HW);
    ++*cx.launch_counter;
    return cudaGetLastError();
  }
  GemmArgs g{};
  g.A = latent192; g.A2 = frames6; g.sam = 1; g.sak = HW; g.ksplit = 192;
  g.Wt = test_variant ? w.w_tab_lat : w.w_tab; g.bias = nullptr; g.C = tab; g.scm = 256; g.scn = 1;
  g.M = HW; g.N = 256; g.K = 198; g.act = 0; g.out_half = tab_half ? 1 : 0;
  return launch_gemm(cx, g, true);
}

#define STIF_TRY(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return e__; } while (0)

// One dense layer C[M, N] = act(A[M, K] . W[n_off : n_off + N]^T + b): on the tensor cores (split-bf16, kernels_hp.cu)
// when the handle carries the hi / lo weight images, else the SIMT SGEMM (STIF_FP32_SIMT=1: the test-only anchor).
static cudaError_t dense_layer(const LaunchCtx& cx, const DeviceWeights32& w, int id, int n_off, int N, const float* A, int K,
                               const float* Wt, const float* bias, float* C, long ldc, long M, int act) {
  if (const HpLayer* L = hp_layer(w.hp, id)) return hp_gemm(cx, *L, n_off, N, A, bias + n_off, C, ldc, M, act);
  return launch_gemm(cx, dense(A, K, Wt + (size_t)n_off * K, bias + n_off, C, ldc, M, N, act), false);
}

// The composed last layer of feat_imnet (256 -> F 64 | Q1 64 | Q2 64, no activation): F and Q1|Q2 live in different tables; one pass
// over the 256-wide input on the tensor-core path (hp_gemm_split), two on the SIMT anchor.
static cudaError_t composed_layer(const LaunchCtx& cx, const DeviceWeights32& w, const float* A, float* f_out, float* q_out, long M) {
  if (const HpLayer* L = hp_layer(w.hp, HP_F3)) return hp_gemm_split(cx, *L, A, w.f3_b, M, 0, 64, f_out, 64, q_out, 128);
  STIF_TRY(dense_layer(cx, w, HP_F3, 0, 64, A, 256, w.f3_w, w.f3_b, f_out, 64, M, 0));
  return dense_layer(cx, w, HP_F3, 64, 128, A, 256, w.f3_w, w.f3_b, q_out, 128, M, 0);
}

// A 256-wide sine layer followed by its NOUT-wide output layer: one launch on the tensor-core path (the activations stay in the
// epilogue's registers, hp_gemm_proj), two on the SIMT anchor (through the scratch activations `act`).
template <int NOUT>
static cudaError_t dense_out_layer(const LaunchCtx& cx, const DeviceWeights32& w, int id, const float* A, int K, const float* Wt,
                                   const float* bias, float* act, const float* out_w, const float* out_b, float* out, long scm, long scn,
                                   long M) {
  if (const HpLayer* L = hp_layer(w.hp, id)) return hp_gemm_proj(cx, *L, A, bias, M, out_w, out_b, NOUT, out, scm, scn);
  STIF_TRY(dense_layer(cx, w, id, 0, 256, A, K, Wt, bias, act, 256, M, 1));
  return launch_out_layer<NOUT>(cx, act, out_w, out_b, out, scm, scn, M);
}

cudaError_t decode_slab_fp32(const LaunchCtx& cx, const DeviceWeights32& w, const FoldedWeights& hw, const Geometry& geo,
                             const Workspace& ws, float t, int row_begin, int row_end, int k1_row_begin,
                             int k1_row_end, float* out_rgb, int stage) {
  const long WW = geo.WW;
  const long Qall = (long)geo.HH * geo.WW;
  const float* tab = reinterpret_cast<const float*>(ws.tab);
  float* qtab = reinterpret_cast<float*>(ws.qtab);
  const Vec64 cA = time_constant(hw.a_t, hw.a_b, t), cB = time_constant(hw.b_t, hw.b_b, t),
              cE = time_constant(hw.e_t, hw.e_b, t);
  const long chunk = (long)ws.chunk;
  // ---- K1: stage A + B over rows [k1_row_begin, k1_row_end)
  for (long q0 = k1_row_begin * WW; stage == 1 && q0 < k1_row_end * WW; q0 += chunk) {
    long n = std::min(chunk, k1_row_end * WW - q0);
    unsigned blocks = (unsigned)((n * 16 + 255) / 256);
    stage_a_first_layer<<<blocks, 256, 0, cx.stream>>>(tab, geo, w.a_rel, cA, q0, n, ws.act_c);
    ++*cx.launch_counter;
    STIF_TRY(cudaGetLastError());
    STIF_TRY(dense_layer(cx, w, HP_F1, 0, 64, ws.act_c, 64, w.f1_w, w.f1_b, ws.act_a, 64, n, 1));
    STIF_TRY(dense_layer(cx, w, HP_F2, 0, 256, ws.act_a, 64, w.f2_w, w.f2_b, ws.act_b, 256, n, 1));
    STIF_TRY(composed_layer(cx, w, ws.act_b, ws.act_c, qtab + q0 * 128, n));
    stage_b_first_layer<<<blocks, 256, 0, cx.stream>>>(tab, geo, cB, q0, n, ws.act_c, nullptr, reinterpret_cast<const float*>(ws.utab));
    ++*cx.launch_counter;
    STIF_TRY(cudaGetLastError());
    STIF_TRY(dense_layer(cx, w, HP_L1, 0, 64, ws.act_c, 64, w.l1_w, w.l1_b, ws.act_a, 64, n, 1));
    STIF_TRY((dense_out_layer<4>(cx, w, HP_L2, ws.act_a, 64, w.l2_w, w.l2_b, ws.act_b, w.l3_w, w.l3_b, ws.flow + q0 * 4, 4, 1, n)));
  }
  // ---- K2: stage C + D + E over rows [row_begin, row_end)
  for (long q0 = row_begin * WW; stage == 2 && q0 < row_end * WW; q0 += chunk) {
    long n = std::min(chunk, row_end * WW - q0);
    unsigned blocks = (unsigned)((n * 16 + 255) / 256);
    if (ws.utab)
      stage_e_first_layer<<<blocks, 256, 0, cx.stream>>>(tab, qtab, ws.flow, geo, cE, q0, n, k1_row_begin, k1_row_end,
                                                        ws.flag, ws.act_c, reinterpret_cast<const float*>(ws.utab));
    else
      stage_e_first_layer_shfl<<<blocks, 256, 0, cx.stream>>>(tab, qtab, ws.flow, geo, cE, q0, n, k1_row_begin, k1_row_end, ws.flag, ws.act_c);
    ++*cx.launch_counter;
    STIF_TRY(cudaGetLastError());
    STIF_TRY(dense_layer(cx, w, HP_E1, 0, 64, ws.act_c, 64, w.e1_w, w.e1_b, ws.act_a, 64, n, 1));
    STIF_TRY(dense_layer(cx, w, HP_E2, 0, 256, ws.act_a, 64, w.e2_w, w.e2_b, ws.act_b, 256, n, 1));
    STIF_TRY((dense_out_layer<3>(cx, w, HP_E3, ws.act_b, 256, w.e3_w, w.e3_b, ws.act_a, w.e4_w, w.e4_b, out_rgb + q0, 1, Qall, n)));   // planar [3,HH,WW] (Sakuya_arch_test.py:457)
  }
  return cudaSuccess;
}

cudaError_t ensemble_blend_launch(const LaunchCtx& cx, const float* pred, float* out_rgb, int HH, int WW, const AxisTables ens_y[2],
                                  const AxisTables ens_x[2], int k) {
  const long Q = (long)HH * WW;
  ensemble_blend<<<(unsigned)((Q + 255) / 256), 256, 0, cx.stream>>>(pred, out_rgb, WW, Q, ens_y[0], ens_y[1], ens_x[0], ens_x[1], k);
  ++*cx.launch_counter;
  return cudaGetLastError();
}

cudaError_t decode_slab_fp32_ensemble(const LaunchCtx& cx, const DeviceWeights32& w, const FoldedWeights& hw,
                                      const Geometry geo_pass[4], const AxisTables ens_y[2], const AxisTables ens_x[2],
                                      const Workspace& ws, float t, float* out_rgb) {
  const Geometry& g0 = geo_pass[0];
  const long Q = (long)g0.HH * g0.WW;
  const float* tab = reinterpret_cast<const float*>(ws.tab);
  float* qtab = reinterpret_cast<float*>(ws.qtab);
  const Vec64 cA = time_constant(hw.a_t, hw.a_b, t), cB = time_constant(hw.b_t, hw.b_b, t);   // (stage E's constant: decode_slab_fp32)
  const long chunk = (long)ws.chunk;
  for (int k = 0; k < 4; ++k) {
    const Geometry& geo = geo_pass[k];
    // stage A for the whole slab: F and Q tables (stage B gathers F at OTHER pixels in this mode)
    for (long q0 = 0; q0 < Q; q0 += chunk) {
      long n = std::min(chunk, Q - q0);
      unsigned blocks = (unsigned)((n * 16 + 255) / 256);
      stage_a_first_layer<<<blocks, 256, 0, cx.stream>>>(tab, geo, w.a_rel, cA, q0, n, ws.act_c);
      ++*cx.launch_counter;
      STIF_TRY(cudaGetLastError());
      STIF_TRY(dense_layer(cx, w, HP_F1, 0, 64, ws.act_c, 64, w.f1_w, w.f1_b, ws.act_a, 64, n, 1));
      STIF_TRY(dense_layer(cx, w, HP_F2, 0, 256, ws.act_a, 64, w.f2_w, w.f2_b, ws.act_b, 256, n, 1));
      STIF_TRY(composed_layer(cx, w, ws.act_b, ws.ftab + q0 * 64, qtab + q0 * 128, n));
    }
    // stage B
    for (long q0 = 0; q0 < Q; q0 += chunk) {
      long n = std::min(chunk, Q - q0);
      unsigned blocks = (unsigned)((n * 16 + 255) / 256);
      stage_b_first_layer<<<blocks, 256, 0, cx.stream>>>(tab, geo, cB, q0, n, ws.act_c, ws.ftab, nullptr);
      ++*cx.launch_counter;
      STIF_TRY(cudaGetLastError());
      STIF_TRY(dense_layer(cx, w, HP_L1, 0, 64, ws.act_c, 64, w.l1_w, w.l1_b, ws.act_a, 64, n, 1));
      STIF_TRY((dense_out_layer<4>(cx, w, HP_L2, ws.act_a, 64, w.l2_w, w.l2_b, ws.act_b, w.l3_w, w.l3_b, ws.flow + q0 * 4, 4, 1, n)));
    }
    // stage C + D + E into the per-pass prediction, then the area-weighted accumulation
    STIF_TRY(decode_slab_fp32(cx, w, hw, geo, ws, t, 0, geo.HH, 0, geo.HH, ws.pred, 2));
    ensemble_blend<<<(unsigned)((Q + 255) / 256), 256, 0, cx.stream>>>(ws.pred, out_rgb, geo.WW, Q, ens_y[0], ens_y[1], ens_x[0],
                                                                     ens_x[1], k);
    ++*cx.launch_counter;
    STIF_TRY(cudaGetLastError());
  }
  return cudaSuccess;
}

}  // namespace stif
