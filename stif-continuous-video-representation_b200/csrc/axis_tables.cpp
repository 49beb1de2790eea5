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

void build_axis(int n_lr, int n_hr, HostAxis& o) {
  o.coord.resize(n_hr);
  o.rel.resize(n_hr);
  o.bw.resize(n_hr);
  o.base.resize(n_hr);
  o.idx.resize(n_hr);
  o.b0.resize(n_hr);
  const float lo = kClampLo, hi = kClampHi;
  for (int j = 0; j < n_hr; ++j) {
    float c = axis_coord(n_hr, j);
    c = c < lo ? lo : (c > hi ? hi : c);                           // clamp (:373)
    // grid_sampler_unnormalize, align_corners=False: ((c + 1) * n - 1) / 2, each op rounded
    volatile float a = c + 1.0f;
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

}  // namespace stif
