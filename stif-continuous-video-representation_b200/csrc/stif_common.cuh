// stif_common.cuh -- shared definitions for the STIF query-decoder kernels (sm_100a).
//
// Vocabulary (follows the reference, codes/models/modules/Sakuya_arch_test.py:364-459):
//   latent  [192,H,W]   = cat(self.feat[:,0..2])            frames [6,H,W] = self.inp.view(6,H,W)
//   query   one output pixel (jy,jx) of the HH x WW raster at one time t
//   slab    all HH*WW queries of one (t, b)
//   tab     "projected latent table"  [H*W, 256] : the four first-layer products that depend only
//           on the LR texel (TA | TB | TE1 | TE2, 64 channels each), pre-multiplied by omega_0
//   qtab    "projected HR table" [HH*WW, 128] : Q1 | Q2, encode_imnet's first-layer products of
//           HRfeat, folded into feat_imnet's last layer
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace stif {

constexpr float kClampLo = -0.99999898672103881836f;  // fl32(-1 + 1e-6)  (Sakuya_arch_test.py:373)
constexpr float kClampHi = 0.99999898672103881836f;   // fl32( 1 - 1e-6)

// Per-axis query tables, built on the HOST with separately rounded fp32 operations
// (axis_tables.cpp) so that the nearest index / rel chain is bit-exact by construction.
struct AxisTables {
  const int32_t* idx;    // [n_hr] nearest LR texel                       (:382-393)
  const float* rel;      // [n_hr] (c - lr_c[idx]) * n_lr                 (:394-396)
  const int32_t* b0;     // [n_hr] floor of the unnormalised query coord  (stage-B bilinear, :410-417)
  const float* bw;       // [n_hr] its fractional part
  const float* base;     // [n_hr] linspace(-1,1,n_hr)                    (warplayer.py:28-31)
  const int32_t* hidx;   // [n_hr] local-ensemble passes only: nearest HR index of the shifted coordinate (:1026-1029), else null
};

struct Geometry {
  int H, W, HH, WW;
  AxisTables y, x;
  float half_h, half_w;  // (HH-1)/2, (WW-1)/2 : warpgrid's flow normalisers (warplayer.py:35-36)
};

// One zero-padded bilinear footprint on an (n_y x n_x) grid: 4 texel offsets (in texels, -1 = outside)
// and 4 weights.  grid_sampler_unnormalize(align_corners=False) = ((c+1)*n-1)/2
// (ATen/native/cuda/GridSampler.cuh:23-31); taps outside the grid contribute zero.
struct Taps {
  int off[4];
  float w[4];
};

#ifdef __CUDACC__
__device__ __forceinline__ float unnormalize(float c, int n) {
  return ((c + 1.0f) * (float)n - 1.0f) * 0.5f;
}

__device__ __forceinline__ Taps make_taps(float gy, float gx, int ny, int nx) {
  float v = unnormalize(gy, ny), u = unnormalize(gx, nx);
  float fy = floorf(v), fx = floorf(u);
  float wy1 = v - fy, wx1 = u - fx;
  float wy0 = 1.0f - wy1, wx0 = 1.0f - wx1;
  int y0 = (int)fy, x0 = (int)fx;
  bool yv0 = (y0 >= 0) & (y0 < ny), yv1 = (y0 + 1 >= 0) & (y0 + 1 < ny);
  bool xv0 = (x0 >= 0) & (x0 < nx), xv1 = (x0 + 1 >= 0) & (x0 + 1 < nx);
  Taps t;
  t.off[0] = (yv0 & xv0) ? y0 * nx + x0 : -1;           t.w[0] = wy0 * wx0;
  t.off[1] = (yv0 & xv1) ? y0 * nx + x0 + 1 : -1;       t.w[1] = wy0 * wx1;
  t.off[2] = (yv1 & xv0) ? (y0 + 1) * nx + x0 : -1;     t.w[2] = wy1 * wx0;
  t.off[3] = (yv1 & xv1) ? (y0 + 1) * nx + x0 + 1 : -1; t.w[3] = wy1 * wx1;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (t.off[k] < 0) { t.off[k] = 0; t.w[k] = 0.0f; }
  return t;
}

// Stage-B footprint from the host-built per-axis tables (query's own position on the LR grid).
__device__ __forceinline__ Taps make_taps_tables(const Geometry& g, int jy, int jx) {
  int y0 = g.y.b0[jy], x0 = g.x.b0[jx];
  float wy1 = g.y.bw[jy], wx1 = g.x.bw[jx];
  float wy0 = 1.0f - wy1, wx0 = 1.0f - wx1;
  bool yv0 = (y0 >= 0) & (y0 < g.H), yv1 = (y0 + 1 >= 0) & (y0 + 1 < g.H);
  bool xv0 = (x0 >= 0) & (x0 < g.W), xv1 = (x0 + 1 >= 0) & (x0 + 1 < g.W);
  Taps t;
  t.off[0] = (yv0 & xv0) ? y0 * g.W + x0 : -1;           t.w[0] = wy0 * wx0;
  t.off[1] = (yv0 & xv1) ? y0 * g.W + x0 + 1 : -1;       t.w[1] = wy0 * wx1;
  t.off[2] = (yv1 & xv0) ? (y0 + 1) * g.W + x0 : -1;     t.w[2] = wy1 * wx0;
  t.off[3] = (yv1 & xv1) ? (y0 + 1) * g.W + x0 + 1 : -1; t.w[3] = wy1 * wx1;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (t.off[k] < 0) { t.off[k] = 0; t.w[k] = 0.0f; }
  return t;
}

// Stage C (warpgrid, warplayer.py:25-39 + clamp at Sakuya_arch_test.py:428,441):
// normalised sampling position of one warp from the flow in HR-pixel units.
__device__ __forceinline__ void warp_position(const Geometry& g, int jy, int jx, float dx, float dy,
                                              float& gy, float& gx) {
  gx = g.x.base[jx] + __fdiv_rn(dx, g.half_w);
  gy = g.y.base[jy] + __fdiv_rn(dy, g.half_h);
  gx = fminf(fmaxf(gx, kClampLo), kClampHi);
  gy = fminf(fmaxf(gy, kClampLo), kClampHi);
}

#endif  // __CUDACC__

}  // namespace stif
