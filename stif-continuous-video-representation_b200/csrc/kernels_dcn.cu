// kernels_dcn.cu -- modulated deformable convolution v2, forward, for sm_100a.
//
// Replaces, behind the same call signature, the reference's DCNv2 extension entry `_ext.dcn_v2_forward`
// (codes/models/modules/DCNv2/dcn_v2.py:24-27 -> src/cuda/dcn_v2_cuda.cu:42-172): bias broadcast, modulated deformable
// im2col (src/cuda/dcn_v2_im2col_cuda.cu:125-195, bilinear taps :24-52) and a batched SGEMM columns x weight.  That
// extension is THC-era code that does not build against torch >= 1.11; the encoder of the reference (`gen_feat`,
// Sakuya_arch_test.py:313-362) calls it 78 times per frame pair, always with the same geometry:
//     64 -> 64 channels, 3 x 3, stride 1, padding 1, dilation 1, 8 deformable groups      (PCD_Align / Easy_PCD, :38-66, :135-160)
// which is the one configuration this kernel implements (anything else returns STIF_EINVAL and the caller keeps its
// fallback).  Here the column buffer never exists: per 128-pixel tile and per kernel tap the sampled, mask-modulated values
// [128 pixels x 64 channels] are built straight into a K-major SW128 shared-memory tile and multiplied with that tap's
// [64 x 64] weight block by tcgen05.mma, nine taps accumulating into one fp32 TMEM accumulator.  fp32-class accuracy comes
// from the same 2-term bf16 split as the high-precision decoder mode (kernels_hp.cu): a_hi w_hi + a_lo w_hi + a_hi w_lo.
#include <algorithm>
#include <mutex>
#include <cstdio>
#include <string>

#include "../../include/stif_b200.h"
#include "stif_internal.h"
#include "tc_primitives.cuh"

namespace stif {
namespace {

using namespace tc;

struct DcnParams {
  const float* input;    // [B, 8 groups, H * W, 8 channels]: the group-major channels-last copy of the NCHW input (dcn_pack_input_kernel)
  const float* weight;   // [64, 64, 3, 3]
  const float* bias;     // [64]
  const float* offset;   // [B, 8 * 2 * 9, H, W]   channel (g * 18 + 2 k) = dy, (+ 1) = dx of tap k = i * 3 + j   (im2col :162-167)
  const float* mask;     // [B, 8 * 9, H, W]
  float* out;            // [B, 64, H, W]
  int B, H, W;
};

constexpr uint32_t kWBlk = 8192, kATile = 16384;
constexpr uint32_t dW_hi = 0, dW_lo = 9 * kWBlk, dA = 18 * kWBlk, dBars = dA + 4 * kATile, dSmem = dBars + 64;
static_assert(dSmem <= 232448, "exceeds 227 KB of shared memory");

extern __shared__ __align__(1024) uint8_t dcn_smem[];

__device__ __forceinline__ void wait_or_trap(uint64_t* bar, uint32_t parity) {
  for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it)
    if (it > (1u << 24)) __trap();
}

__global__ void __launch_bounds__(512, 1) dcn_v2_forward_kernel(const __grid_constant__ DcnParams p) {
  uint8_t* sW_hi = dcn_smem + dW_hi;
  uint8_t* sW_lo = dcn_smem + dW_lo;
  uint8_t* sA = dcn_smem + dA;                                          // [buffer 0: hi | lo][buffer 1: hi | lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(dcn_smem + dBars);       // [0,1] MMAs that read A buffer 0 / 1 are done, [2] accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long HW = (long)p.H * p.W, total = (long)p.B * HW;
  const long ntiles = (total + 127) / 128;
  if (tid == 0) {
    for (int i = 0; i < 3; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  // weights: [co][c][tap] fp32 -> per tap a [64 co x 64 c] K-major SW128 block, split into bf16 hi / lo
  for (int idx = tid; idx < 64 * 64 * 9; idx += 512) {
    const int co = idx / 576, rem = idx - co * 576, c = rem / 9, kb = rem - c * 9;
    const float w = __ldg(p.weight + idx);
    const uint32_t h = pack_bf16x2(w, 0.f) & 0xFFFFu;
    const float r = w - __uint_as_float(h << 16);
    const uint32_t o = (uint32_t)kb * kWBlk + sw128_offset(co, c);
    *reinterpret_cast<uint16_t*>(sW_hi + o) = (uint16_t)h;
    *reinterpret_cast<uint16_t*>(sW_lo + o) = (uint16_t)(pack_bf16x2(r, 0.f) & 0xFFFFu);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int row = tid & 127, gq = tid >> 7;                             // this thread builds pixel `row`, deformable groups 2 gq, 2 gq + 1
  uint32_t it = 0;                                                       // taps built so far (A buffer = it & 1)
  uint32_t ntile_done = 0;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ntile_done) {
    const long pflat = tile * 128 + row;
    const bool live = pflat < total;
    const int b = live ? (int)(pflat / HW) : 0;
    const long pix = live ? pflat - (long)b * HW : 0;
    const int y = (int)(pix / p.W), x = (int)(pix - (long)y * p.W);
#pragma unroll 1
    for (int kb = 0; kb < 9; ++kb, ++it) {
      const uint32_t buf = it & 1;
      if (it >= 2) wait_or_trap(&bars[buf], ((it >> 1) - 1) & 1);       // the tap that used this buffer two taps ago has been multiplied
      uint8_t* a_hi = sA + buf * 2 * kATile;
      uint8_t* a_lo = a_hi + kATile;
      const int i = kb / 3, j = kb - i * 3;
#pragma unroll
      for (int gi = 0; gi < 2; ++gi) {
        const int g = gq * 2 + gi;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (live) {
          const float* off = p.offset + ((long)(b * 8 + g) * 18 + 2 * kb) * HW + pix;
          const float off_h = __ldg(off), off_w = __ldg(off + HW);
          const float m = __ldg(p.mask + ((long)(b * 8 + g) * 9 + kb) * HW + pix);
          const float h_im = (float)(y - 1 + i) + off_h, w_im = (float)(x - 1 + j) + off_w;      // (:176-177)
          if (h_im > -1.f && w_im > -1.f && h_im < (float)p.H && w_im < (float)p.W) {              // (:179)
            const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
            const float lh = h_im - (float)h_low, lw = w_im - (float)w_low, hh = 1.f - lh, hw = 1.f - lw;
            const bool t1 = h_low >= 0 && w_low >= 0, t2 = h_low >= 0 && w_low + 1 <= p.W - 1;
            const bool t3 = h_low + 1 <= p.H - 1 && w_low >= 0, t4 = h_low + 1 <= p.H - 1 && w_low + 1 <= p.W - 1;
            const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
            const long o1 = (long)h_low * p.W + w_low;
            // one 32-byte load per corner: the 8 channels of this deformable group are contiguous in the packed copy (the
            // NCHW original needs 8 scalar loads per corner, each up to 32 L1 wavefronts when the offsets scatter the lanes)
            const float* src = p.input + ((long)(b * 8 + g) * HW) * 8;
            const U8x32 z{};
            const U8x32 c1 = t1 ? ldg256(src + o1 * 8) : z, c2 = t2 ? ldg256(src + (o1 + 1) * 8) : z;
            const U8x32 c3 = t3 ? ldg256(src + (o1 + p.W) * 8) : z, c4 = t4 ? ldg256(src + (o1 + p.W + 1) * 8) : z;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              v[c] = (w1 * __uint_as_float(c1.r[c]) + w2 * __uint_as_float(c2.r[c]) + w3 * __uint_as_float(c3.r[c]) +
                      w4 * __uint_as_float(c4.r[c])) * m;                                             // (:44-50, :190)
          }
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          hi[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
          lo[e] = pack_bf16x2(v[2 * e] - __uint_as_float(hi[e] << 16), v[2 * e + 1] - __uint_as_float(hi[e] & 0xFFFF0000u));
        }
        const uint32_t o = sw128_offset(row, g * 8);
        *reinterpret_cast<uint4*>(a_hi + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(a_lo + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(128, 64);
        const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), wh = smem_u32(sW_hi + kb * kWBlk), wl = smem_u32(sW_lo + kb * kWBlk);
        bool first = kb == 0;
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint32_t a = term == 1 ? al : ah, w = term == 2 ? wl : wh;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_ss(tmem, make_desc_sw128(a) + 2 * k, make_desc_sw128(w) + 2 * k, idesc, !first);
            first = false;
          }
        }
        umma_commit(&bars[buf]);
        if (kb == 8) umma_commit(&bars[2]);
      }
    }
    // ---- epilogue: accumulator + bias -> out[b, co, y, x]; warp = (lane quarter, 16 output channels)
    wait_or_trap(&bars[2], ntile_done & 1);
    tc_fence_after();
    {
      const int quarter = warp & 3, cg = warp >> 2;
      uint32_t acc[16];
      tmem_ld16(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)cg * 16u, acc);
      tmem_ld_wait();
      const long pe = tile * 128 + quarter * 32 + lane;
      if (pe < total) {
        const int be = (int)(pe / HW);
        float* dst = p.out + ((long)be * 64 + cg * 16) * HW + (pe - (long)be * HW);
#pragma unroll
        for (int c = 0; c < 16; ++c) dst[(long)c * HW] = __uint_as_float(acc[c]) + __ldg(p.bias + cg * 16 + c);
      }
    }
    tc_fence_before();
    __syncthreads();                                                     // the accumulator is drained before the next tile's first MMA
  }
  if (warp == 0) tmem_dealloc(tmem, 64);
}

// NCHW [B, 64, H, W] -> [B, 8 groups, H * W, 8 channels]: thread = (b, g, pixel); reads coalesced along the pixel axis per
// channel, writes one 32-byte sector.
__global__ void dcn_pack_input_kernel(const float* __restrict__ in, float* __restrict__ out, long HW, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long bg = i / HW, pix = i - bg * HW;
  uint32_t v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = __float_as_uint(__ldg(in + (bg * 8 + c) * HW + pix));
  stg256(out + i * 8, v);
}

// cudaFuncSetAttribute is per device: remember where the kernel's shared-memory limit has been raised
// ... and keep one stream-ordered memory pool per device for the packed input (release threshold = never: after the first call an
// allocation is a pool hit, a few microseconds; the device's default pool would hand its memory back at every synchronisation)
std::mutex g_dcn_mutex;
bool g_dcn_configured[64] = {};
cudaMemPool_t g_dcn_pool[64] = {};

}  // namespace
}  // namespace stif

extern "C" int stif_dcn_v2_forward(const float* input, const float* weight, const float* bias, const float* offset, const float* mask,
                                   int B, int C, int H, int W, int Cout, int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw,
                                   int dg, float* out, void* stream) {
  using namespace stif;
  if (!input || !weight || !bias || !offset || !mask || !out || B < 1 || H < 1 || W < 1) return STIF_EINVAL;
  if (C != 64 || Cout != 64 || kh != 3 || kw != 3 || sh != 1 || sw != 1 || ph != 1 || pw != 1 || dh != 1 || dw != 1 || dg != 8)
    return STIF_EINVAL;   // not the encoder's configuration: the caller keeps its own path
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return STIF_ECUDA;
  if (dev < 0 || dev >= 64) return STIF_EINVAL;
  {
    std::lock_guard<std::mutex> lock(g_dcn_mutex);
    if (!g_dcn_configured[dev]) {
      if (cudaFuncSetAttribute(dcn_v2_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dSmem) != cudaSuccess) return STIF_ECUDA;
      cudaMemPoolProps props = {};
      props.allocType = cudaMemAllocationTypePinned;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = dev;
      if (cudaMemPoolCreate(&g_dcn_pool[dev], &props) != cudaSuccess) return STIF_ECUDA;
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(g_dcn_pool[dev], cudaMemPoolAttrReleaseThreshold, &keep);
      g_dcn_configured[dev] = true;
    }
  }
  const long ntiles = ((long)B * H * W + 127) / 128;
  // packed copy of the input: stream-ordered allocation (the pool recycles it), so the entry point keeps no buffer of its own and
  // concurrent calls on different streams do not share one
  float* packed = nullptr;
  if (cudaMallocFromPoolAsync(&packed, (size_t)B * 64 * H * W * sizeof(float), g_dcn_pool[dev], (cudaStream_t)stream) != cudaSuccess) return STIF_ECUDA;
  const long npack = (long)B * 8 * H * W;
  dcn_pack_input_kernel<<<(unsigned)((npack + 255) / 256), 256, 0, (cudaStream_t)stream>>>(input, packed, (long)H * W, npack);
  DcnParams p{packed, weight, bias, offset, mask, out, B, H, W};
  dcn_v2_forward_kernel<<<(unsigned)std::min<long>(sms, ntiles), 512, dSmem, (cudaStream_t)stream>>>(p);
  const cudaError_t e = cudaGetLastError();
  const cudaError_t ef = cudaFreeAsync(packed, (cudaStream_t)stream);
  return e == cudaSuccess && ef == cudaSuccess ? STIF_OK : STIF_ECUDA;
}
