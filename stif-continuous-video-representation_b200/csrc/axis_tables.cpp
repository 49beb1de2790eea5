// axis_tables.cpp -- per-axis query tables, built on the host in strict (uncontracted) fp32.
//
// The reference builds its query coordinates on the CPU with two separately rounded ATen kernels
// (make_coord, Sakuya_arch_test.py:1233-1248: `v0 + r + (2*r) * arange(n).float()`), clamps them
// (:373) and lets grid_sample(mode='nearest', align_corners=False) round the unnormalised
// coordinate with nearbyint (half-to-even).  At non-integer scales this chain hits exact .5 ties
// (SURVEY.md section 7.3-3), so it is replayed here operation by operation instead of being
// re-derived in a kernel where nvcc may contract to FMA.  Compile with -ffp-contract=off.
#include <cmath>

#include "stif_internal.h"

namespace stif {

static inline float axis_coord(int n, int j) {
  const double r = 1.0 / (double)n;           // (v1 - v0) / (2 n), python double
  const float c0 = (float)(-1.0 + r);         // python double -> fp32 at the tensor op
  const float step = (float)(2.0 * r);
  volatile float prod = step * (float)j;      // rounded product (mul kernel)
  volatile float sum = c0 + prod;             // rounded sum     (add kernel)
  return sum;
}

void build_axis(int n_lr, int n_hr, HostAxis& o, int shift_sign) {
  o.coord.resize(n_hr);
  o.rel.resize(n_hr);
  o.bw.resize(n_hr);
  o.base.resize(n_hr);
  o.idx.resize(n_hr);
  o.b0.resize(n_hr);
  o.hidx.assign(shift_sign ? n_hr : 0, 0);
  // local-ensemble shift: python double (v * (2/n_lr/2) + 1e-6) -> fp32 at the in-place tensor add (:993-994)
  const float shift = (float)((double)shift_sign * (2.0 / (double)n_lr / 2.0) + 1e-6);
  const float lo = kClampLo, hi = kClampHi;
  for (int j = 0; j < n_hr; ++j) {
    float c = axis_coord(n_hr, j);
    c = c < lo ? lo : (c > hi ? hi : c);                           // clamp (:373)
    float cs = c;                                                  // coordinate the gathers use
    if (shift_sign) {
      volatile float moved = c + shift;
      cs = moved;
      cs = cs < lo ? lo : (cs > hi ? hi : cs);                     // clamp_ (:995)
      volatile float ha = cs + 1.0f;
      volatile float hb = ha * (float)n_hr;
      volatile float hd = hb - 1.0f;
      volatile float hu = hd / 2.0f;
      o.hidx[j] = (int)std::nearbyintf((float)hu);                 // nearest HR pixel of the shifted coordinate (:1026-1029)
    }
    // grid_sampler_unnormalize, align_corners=False: ((c + 1) * n - 1) / 2, each op rounded
    volatile float a = cs + 1.0f;
    volatile float b = a * (float)n_lr;
    volatile float d = b - 1.0f;
    volatile float u = d / 2.0f;
    const float uu = u;
    const int i = (int)std::nearbyintf(uu);                        // round-half-even (default FE mode)
    const bool inb = i >= 0 && i < n_lr;
    const float q = inb ? axis_coord(n_lr, i) : 0.0f;              // feat_coord is not clamped (:375-377)
    volatile float diff = c - q;
    volatile float rel = diff * (float)n_lr;                       // (:394-396)
    const float f = std::floor(uu);
    volatile float frac = uu - f;
    o.coord[j] = c;
    o.idx[j] = i;
    o.rel[j] = rel;
    o.b0[j] = (int)f;
    o.bw[j] = frac;
    // warp base grid: linspace(-1, 1, n_hr) (warplayer.py:28-31), correctly rounded from float64.
    double basev = (n_hr == 1) ? -1.0 : (-1.0 + (double)j * (2.0 / (double)(n_hr - 1)));
    if (j == n_hr - 1 && n_hr > 1) basev = 1.0;
    o.base[j] = (float)basev;
  }
}

void ensemble_weights_host(int H, int W, int HH, int WW, float* w) {
  HostAxis ym, yp, xm, xp;
  build_axis(H, HH, ym, -1);
  build_axis(H, HH, yp, +1);
  build_axis(W, WW, xm, -1);
  build_axis(W, WW, xp, +1);
  const size_t Q = (size_t)HH * WW;
  for (int jy = 0; jy < HH; ++jy)
    for (int jx = 0; jx < WW; ++jx) {
      const float ry[2] = {ym.rel[jy], yp.rel[jy]}, rx[2] = {xm.rel[jx], xp.rel[jx]};
      float a[4];
      for (int k = 0; k < 4; ++k) {                                 // (vx, vy) = (-1,-1), (-1,1), (1,-1), (1,1)
        volatile float prod = ry[k >> 1] * rx[k & 1];
        volatile float ar = std::fabs((float)prod) + 1e-9f;         // area + 1e-9 (:1011-1012)
        a[k] = ar;
      }
      volatile float s01 = a[0] + a[1];
      volatile float s012 = s01 + a[2];
      volatile float tot = s012 + a[3];                             // torch.stack(areas).sum(0)
      for (int k = 0; k < 4; ++k) {
        volatile float wk = a[3 - k] / tot;
        w[(size_t)k * Q + (size_t)jy * WW + jx] = wk;
      }
    }
}

}  // namespace stif
