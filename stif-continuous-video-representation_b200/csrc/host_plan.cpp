// host_plan.cpp -- band plan of stif_decode_host's band-major pipeline (pure host logic, unit-tested on the CPU through
// stif_debug_band_plan).  See decode_host_banded in stif_api.cu for how the plan is executed.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "stif_internal.h"

namespace stif {

// The kernels are persistent with static tile assignment (tile = blockIdx + n * gridDim, three workgroups per CTA taking
// turns on two TMEM slots), so a launch of n tiles costs about fill + ceil(n / num_sms) tile times: band heights are
// chosen so that each launch fills just under a whole number of such waves.
// he[k] / ge[k] = end of band k in HR rows for stage A+B / stage C-E, lr_end[k] = LR rows that must have landed.
// Invariants (tests/test_abi.py): all three are non-decreasing and end at HH / HH / H; ge[k] <= he[k]; every HR row
// below he[k] has its nearest and bilinear LR footprint inside rows [0, lr_end[k]); ge[k] <= he[k] - halo except where
// he[k] == HH (the speculative lag of stage C-E behind stage A+B).
HostBandPlan plan_host_bands(int H, int W, int HH, int WW, int G, int bands_hint, bool bands_forced, int halo, int num_sms) {
  HostAxis ay;
  build_axis(H, HH, ay);
  const long slots = num_sms;
  auto k1_tiles = [&](int rows) { return ((long)rows * WW + 127) / 128; };
  auto k2_tiles = [&](int rows) { return (long)((rows + 7) / 8) * ((WW + 15) / 16); };
  auto wave_aligned = [&](int target, int step, auto tiles_of) {
    int best = std::max(step, target / step * step);
    double best_eff = 0.0;
    for (int r = std::max(step, (int)(0.8 * target) / step * step); r <= (int)(1.05 * target); r += step) {
      const double w = (double)tiles_of(r) / (double)slots, eff = w / std::ceil(w);
      if (eff > best_eff + 1e-9 || (eff > best_eff - 1e-9 && std::abs(r - target) < std::abs(best - target))) { best = r; best_eff = eff; }
    }
    return best;
  };
  // The first band is short (the pipeline starts after a ~0.1 ms upload; the GPU would idle otherwise), the rest are
  // equal (heights are multiples of 8 rows so that stage C-E, whose tiles are 8 rows high, keeps pace with stage A+B),
  // and stage C-E of the last rows is cut once more so that only a short final download is exposed.
  auto make_plan = [&](int nb) {
    const int n_entries = nb > 1 ? nb + 1 : 1;
    HostBandPlan pl{std::vector<int>(n_entries, HH), std::vector<int>(n_entries, HH), std::vector<int>(n_entries, H), 0.0};
    // Band heights: a short first band, nb - 2 equal ones, and a LAST band of `taper` times their height -- the upload is
    // what bounds the call (PCIe), so everything downstream of the last band's arrival is exposed and should be small.
    // (A launch costs whole waves of ONE tile per SM since the three-workgroup rotation: no wave alignment of the heights.)
    static const double taper = getenv("STIF_HOST_TAPER") ? atof(getenv("STIF_HOST_TAPER")) : 0.5;
    const int first = std::max(8, (HH / (4 * std::max(nb, 1))) & ~7);
    const int r1 = nb > 2 ? std::max(8, ((int)((HH - first) / (nb - 2 + taper)) + 7) & ~7) : (nb > 1 ? HH - first : HH);
    const int tail = std::min(HH / 2, wave_aligned(std::max(16, HH / 16), 8, k2_tiles));
    for (int k = 0; k + 1 < nb; ++k) {
      pl.he[k] = std::min(HH, first + k * r1);
      pl.ge[k] = std::max(k ? pl.ge[k - 1] : 0, (pl.he[k] - halo) & ~7);
      if (pl.ge[k] - (k ? pl.ge[k - 1] : 0) < 32) pl.ge[k] = k ? pl.ge[k - 1] : 0;   // too few rows for a launch: next band
      const int last = pl.he[k] - 1;   // LR rows the footprints of HR rows < he[k] touch (tables are monotone)
      pl.lr_end[k] = std::min(H, std::max(ay.idx[last], ay.b0[last] + 1) + 1);
      if (k && pl.lr_end[k] < pl.lr_end[k - 1]) pl.lr_end[k] = pl.lr_end[k - 1];
    }
    if (nb > 1) pl.ge[nb - 1] = std::max(pl.ge[nb - 2], (HH - tail) & ~7);   // band nb-1: last upload / stage A+B; band nb: only the tail of C-E
    nb = n_entries;
    // cost model (microseconds; constants measured on B200, profiles/launch_overhead.py + the banded timeline): the
    // compute stream starts band k when its upload has landed and band k-1 is done; a launch costs ~4.3 us per wave of
    // one tile per SM + ~19 us fixed (launch + filling the three-workgroup rotation); PCIe moves ~55 GB/s each way; the
    // last band's download is exposed
    double t = 0.0;
    for (int k = 0; k < nb; ++k) {
      const int h0 = k ? pl.he[k - 1] : 0, g0 = k ? pl.ge[k - 1] : 0, l0 = k ? pl.lr_end[k - 1] : 0;
      t = std::max(t, (double)pl.lr_end[k] * W * 198 * 4 / 55e3);
      if (pl.lr_end[k] > l0) t += 11.0 + 0.3 * (pl.lr_end[k] - l0) * W / 1000.0;   // K0: fixed + ~0.3 ns per texel
      // (the G resident timesteps of a band share one launch per stage: decode_multi_tc)
      if (pl.he[k] > h0) t += 4.3 * std::ceil((double)G * k1_tiles(pl.he[k] - h0) / slots) + 19.0;
      if (pl.ge[k] > g0) t += 3.9 * std::ceil((double)G * k2_tiles(pl.ge[k] - g0) / slots) + 19.0;
    }
    t += (double)(HH - (nb > 1 ? pl.ge[nb - 2] : 0)) * WW * 12 * G / 55e3;
    pl.cost_us = t;
    return pl;
  };
  int nbands = std::max(1, std::min(bands_hint, H));
  if (!bands_forced) {   // the hint is a starting point: take the cheapest plan nearby (1 band for rasters too small to split)
    int best = 1;
    double best_cost = make_plan(1).cost_us;
    for (int nb = std::max(2, nbands - 2); nb <= std::min(H, nbands + 4); ++nb) {
      if (k1_tiles(HH) < 4 * slots * nb) break;   // a band would not even fill four waves
      const double c = make_plan(nb).cost_us;
      if (c < best_cost) { best = nb; best_cost = c; }
    }
    nbands = best;
  }
  return make_plan(nbands);
}

}  // namespace stif
