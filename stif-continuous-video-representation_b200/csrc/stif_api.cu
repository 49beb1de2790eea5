// stif_api.cu -- the C ABI of libstif_b200.so (include/stif_b200.h): handle, weight upload,
// geometry cache, workspace carve-up and the per-slab orchestration of the decode kernels.
//
// Orchestration mirrors LunaTokis.decoding (codes/models/modules/Sakuya_arch_test.py:364-459):
//   for each batch item b:   project the latent once (t-independent work the reference repeats per t)
//     for each time c:       K1 = stage A+B (feat_imnet, flow_imnet), K2 = stage C+D+E (warp, encode_imnet)
#include <algorithm>
#include <array>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>

#include "../../include/stif_b200.h"
#include "stif_internal.h"

using namespace stif;

namespace {

thread_local std::string g_last_error;

int set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CUDA_OR_RETURN(expr)                                                                        \
  do {                                                                                              \
    cudaError_t e__ = (expr);                                                                       \
    if (e__ != cudaSuccess) return set_error(STIF_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

struct DeviceGeometry {
  Geometry geo{};
  void* blob = nullptr;       // one allocation holding all axis tables
  bool has_ensemble = false;  // local-ensemble tables are built lazily
  Geometry geo_pass[4];       // pass k of decoding_localensemble: (vx,vy) = (-1,-1), (-1,1), (1,-1), (1,1)
  AxisTables ens_y[2], ens_x[2];
  void* ens_blob[4] = {nullptr, nullptr, nullptr, nullptr};
};

// upload one axis (6 arrays of n entries) and return its device-side view
int upload_axis(const HostAxis& a, int n, AxisTables* out, void** blob) {
  const size_t stride = align256((size_t)n * 4);
  std::vector<char> host(6 * stride, 0);
  memcpy(host.data() + 0 * stride, a.idx.data(), (size_t)n * 4);
  memcpy(host.data() + 1 * stride, a.rel.data(), (size_t)n * 4);
  memcpy(host.data() + 2 * stride, a.b0.data(), (size_t)n * 4);
  memcpy(host.data() + 3 * stride, a.bw.data(), (size_t)n * 4);
  memcpy(host.data() + 4 * stride, a.base.data(), (size_t)n * 4);
  if (!a.hidx.empty()) memcpy(host.data() + 5 * stride, a.hidx.data(), (size_t)n * 4);
  CUDA_OR_RETURN(cudaMalloc(blob, host.size()));
  CUDA_OR_RETURN(cudaMemcpy(*blob, host.data(), host.size(), cudaMemcpyHostToDevice));
  char* b = (char*)*blob;
  *out = AxisTables{(const int32_t*)(b + 0 * stride), (const float*)(b + 1 * stride), (const int32_t*)(b + 2 * stride),
                    (const float*)(b + 3 * stride), (const float*)(b + 4 * stride),
                    a.hidx.empty() ? nullptr : (const int32_t*)(b + 5 * stride)};
  return STIF_OK;
}

}  // namespace

struct stif_decoder {
  int device = 0;
  int num_sms = 0;
  bool weights_loaded = false;
  FoldedWeights hw;
  float* d_w32 = nullptr;  // fp32 folded weights (one allocation)
  DeviceWeights32 w32{};
  TcWeights* tcw = nullptr;
  HpWeights* hpw = nullptr;  // split-bf16 images of the dense layers for STIF_MODE_FP32 (kernels_hp.cu)
  std::map<std::array<int, 5>, DeviceGeometry> geos;   // key: H, W, HH, WW, warp-base variant
  std::vector<std::array<int, 5>> geo_order;           // insertion order of `geos` (oldest first): single-entry eviction
  int64_t launches = 0;
  // last decoded slab (debug introspection)
  const float* last_flow = nullptr;
  size_t last_flow_floats = 0;
  // per-kernel-group event timing (stif_profile_*)
  bool profiling = false;
  struct Span { cudaEvent_t a, b; int kind; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> event_pool;
  // stif_decode_host scratch
  void* host_scratch = nullptr;
  size_t host_scratch_bytes = 0;
  cudaStream_t host_stream = nullptr, h2d_stream = nullptr, d2h_stream = nullptr;
  int host_bands = 6;        // LR row bands of the host pipeline
  bool host_bands_forced = false;   // set through stif_debug_host_pipeline: keep the band count even for tiny rasters
  int host_halo = 32;        // HR rows by which K2 trails K1 in the banded host pipeline (doubles after a miss)
  int64_t host_respins = 0;  // how often the speculation missed and K2 was repeated
};

namespace stif {

Workspace carve_workspace(void* base, int H, int W, int HH, int WW, int mode) {
  Workspace ws{};
  const bool fp32 = (mode & 0xFF) == STIF_MODE_FP32;
  const size_t esz = fp32 ? 4 : 2;
  const size_t Q = (size_t)HH * WW;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? (void*)((char*)base + off) : nullptr;
    off += align256(bytes);
    return p;
  };
  ws.tab = take((size_t)H * W * 256 * esz);
  ws.qtab = take(Q * 128 * esz);
  ws.flow = (float*)take(Q * 4 * sizeof(float));
  ws.flag = (int*)take(256);
  if (mode & STIF_FLAG_OUT_U8) ws.rgb32 = (float*)take(Q * 3 * sizeof(float));
  if (mode & STIF_FLAG_TEST_VARIANT) {
    ws.utab = take((size_t)16 * H * W * 192 * esz);
    if (!fp32 && (HH != 4 * H || WW != 4 * W || (mode & STIF_FLAG_WARP_FROM_COORD))) {   // tensor-core path away from the x4 fast path
      ws.uq = take(Q * 192 * 2);
      ws.uadd = take(Q * 64 * 2);
    }
  }
  if (fp32) {
    ws.chunk = std::min<size_t>(Q, (size_t)1 << 20);   // queries per launch of the layer-by-layer pipeline (2.3 KB of activations each)
    ws.act_a = (float*)take(ws.chunk * 256 * sizeof(float));
    ws.act_b = (float*)take(ws.chunk * 256 * sizeof(float));
    ws.act_c = (float*)take(ws.chunk * 64 * sizeof(float));
    if (mode & STIF_FLAG_LOCAL_ENSEMBLE) {
      ws.ftab = (float*)take(Q * 64 * sizeof(float));
      ws.pred = (float*)take(Q * 3 * sizeof(float));
    }
  } else {
    ws.chunk = Q;
    if (mode & STIF_FLAG_LOCAL_ENSEMBLE) {
      ws.ftab = (float*)take(Q * 64 * sizeof(float));
      ws.pred = (float*)take(Q * 3 * sizeof(float));
    }
  }
  ws.total_bytes = off;
  return ws;
}

}  // namespace stif

namespace {

int check_shape(int B, int H, int W, int HH, int WW, int T) {
  if (B < 1 || H < 1 || W < 1 || HH < 1 || WW < 1 || T < 1)
    return set_error(STIF_EINVAL, "invalid shape B=%d H=%d W=%d HH=%d WW=%d T=%d", B, H, W, HH, WW, T);
  if ((long long)HH * WW > (1ll << 30) || (long long)H * W > (1ll << 28))
    return set_error(STIF_EINVAL, "raster too large (HH*WW=%lld, H*W=%lld)", (long long)HH * WW, (long long)H * W);
  return STIF_OK;
}

// warp_from_coord: the warp base grid is the query's pixel-centre coordinate (warpgrid2, warplayer.py:41-47) instead of
// torch.linspace(-1, 1, n) (warpgrid, :28-31) -- the only thing decoding_memory's stage C does differently
int get_geometry(stif_decoder* d, int H, int W, int HH, int WW, cudaStream_t stream, const Geometry** out, bool warp_from_coord = false) {
  std::array<int, 5> key{H, W, HH, WW, warp_from_coord ? 1 : 0};
  auto it = d->geos.find(key);
  if (it == d->geos.end()) {
    HostAxis ay, ax;
    build_axis(H, HH, ay);
    build_axis(W, WW, ax);
    if (warp_from_coord) { ay.base = ay.coord; ax.base = ax.coord; }
    // blob layout: y{idx,rel,b0,bw,base} x{idx,rel,b0,bw,base}, each 256-byte aligned
    size_t ny = align256((size_t)HH * 4), nx = align256((size_t)WW * 4);
    size_t total = 5 * ny + 5 * nx;
    std::vector<char> host(total, 0);
    auto put = [&](size_t off, const void* src, size_t n) { memcpy(host.data() + off, src, n); };
    put(0 * ny, ay.idx.data(), HH * 4); put(1 * ny, ay.rel.data(), HH * 4); put(2 * ny, ay.b0.data(), HH * 4);
    put(3 * ny, ay.bw.data(), HH * 4);  put(4 * ny, ay.base.data(), HH * 4);
    size_t xo = 5 * ny;
    put(xo + 0 * nx, ax.idx.data(), WW * 4); put(xo + 1 * nx, ax.rel.data(), WW * 4); put(xo + 2 * nx, ax.b0.data(), WW * 4);
    put(xo + 3 * nx, ax.bw.data(), WW * 4);  put(xo + 4 * nx, ax.base.data(), WW * 4);
    DeviceGeometry dg;
    CUDA_OR_RETURN(cudaMalloc(&dg.blob, total));
    // synchronous copy: `host` dies at scope exit
    CUDA_OR_RETURN(cudaMemcpy(dg.blob, host.data(), total, cudaMemcpyHostToDevice));
    char* b = (char*)dg.blob;
    Geometry& g = dg.geo;
    g.H = H; g.W = W; g.HH = HH; g.WW = WW;
    g.y = AxisTables{(const int32_t*)(b + 0 * ny), (const float*)(b + 1 * ny), (const int32_t*)(b + 2 * ny),
                     (const float*)(b + 3 * ny), (const float*)(b + 4 * ny), nullptr};
    g.x = AxisTables{(const int32_t*)(b + xo + 0 * nx), (const float*)(b + xo + 1 * nx), (const int32_t*)(b + xo + 2 * nx),
                     (const float*)(b + xo + 3 * nx), (const float*)(b + xo + 4 * nx), nullptr};
    g.half_h = (float)((HH - 1.0) / 2.0);  // python double -> fp32 at the tensor division (warplayer.py:35-36)
    g.half_w = (float)((WW - 1.0) / 2.0);
    // bounded cache (the reference's warp-grid cache is unbounded, warplayer.py:6): drop the OLDEST geometry only.
    // cudaFree synchronises the device, so no kernel still reads the evicted tables.
    while (d->geos.size() >= 64 && !d->geo_order.empty()) {
      auto old = d->geos.find(d->geo_order.front());
      d->geo_order.erase(d->geo_order.begin());
      if (old == d->geos.end()) continue;
      cudaFree(old->second.blob);
      for (void* eb : old->second.ens_blob) if (eb) cudaFree(eb);
      d->geos.erase(old);
    }
    it = d->geos.emplace(key, dg).first;
    d->geo_order.push_back(key);
  }
  *out = &it->second.geo;
  (void)stream;
  return STIF_OK;
}

// Shifted axis tables of decoding_localensemble, built on first use for this geometry.
int get_ensemble_geometry(stif_decoder* d, int H, int W, int HH, int WW, DeviceGeometry** out) {
  const Geometry* base = nullptr;
  if (int rc = get_geometry(d, H, W, HH, WW, nullptr, &base)) return rc;
  DeviceGeometry& dg = d->geos.find(std::array<int, 5>{H, W, HH, WW, 0})->second;
  if (!dg.has_ensemble) {
    for (int s = 0; s < 2; ++s) {
      HostAxis ay, ax;
      build_axis(H, HH, ay, s == 0 ? -1 : +1);
      build_axis(W, WW, ax, s == 0 ? -1 : +1);
      if (int rc = upload_axis(ay, HH, &dg.ens_y[s], &dg.ens_blob[s])) return rc;
      if (int rc = upload_axis(ax, WW, &dg.ens_x[s], &dg.ens_blob[2 + s])) return rc;
    }
    for (int k = 0; k < 4; ++k) {
      dg.geo_pass[k] = dg.geo;
      dg.geo_pass[k].y = dg.ens_y[k >> 1];
      dg.geo_pass[k].x = dg.ens_x[k & 1];
    }
    dg.has_ensemble = true;
  }
  *out = &dg;
  return STIF_OK;
}

cudaEvent_t take_event(stif_decoder* d) {
  if (!d->event_pool.empty()) { cudaEvent_t e = d->event_pool.back(); d->event_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

struct ScopedSpan {  // brackets one kernel group with events when profiling is on
  stif_decoder* d; cudaStream_t s; cudaEvent_t b = nullptr;
  ScopedSpan(stif_decoder* d_, cudaStream_t s_, int kind) : d(d_), s(s_) {
    if (!d->profiling) return;
    if (d->spans.size() >= 16384) {   // nobody is reading (stif_profile_read): recycle the oldest half instead of growing
      for (size_t i = 0; i < 8192; ++i) { d->event_pool.push_back(d->spans[i].a); d->event_pool.push_back(d->spans[i].b); }
      d->spans.erase(d->spans.begin(), d->spans.begin() + 8192);
    }
    cudaEvent_t a = take_event(d);
    b = take_event(d);
    cudaEventRecord(a, s);
    d->spans.push_back({a, b, kind});
  }
  ~ScopedSpan() { if (b) cudaEventRecord(b, s); }
};

// Error exits of the host pipelines: async copies to / from the CALLER's host buffers may still be in flight on the
// copy streams when a later call fails.  The guard drains all three streams before the error code is returned (the caller
// may free its buffers right after) and hands every event of the call back to the pool on every exit.
struct PipeGuard {
  stif_decoder* d;
  cudaStream_t compute, h2d, d2h;
  std::vector<cudaEvent_t>* used;
  bool ok = false;
  ~PipeGuard() {
    if (!ok) {
      if (h2d) cudaStreamSynchronize(h2d);
      cudaStreamSynchronize(compute);
      if (d2h) cudaStreamSynchronize(d2h);
    }
    for (auto e : *used) d->event_pool.push_back(e);
    used->clear();
  }
};

// stif_decode_host: host<->device copies pipelined with the kernels.  The latent is uploaded in row
// bands on `h2d` (K0 projects each band as soon as it lands) and every finished (t,b) slab is
// downloaded on `d2h` while the next slab is being decoded.
struct HostPipe {
  int latent_elem;           // bytes per latent element on the host AND in the device staging copy: 4 (fp32) or 2 (bf16)
  const float* latent_host;
  const float* frames_host;
  void* out_host;            // fp32 [T,B,3,HH,WW], or uint8 [T,B,HH,WW,3] with STIF_FLAG_OUT_U8
  uint8_t* out_u8_dev;       // device copy of the uint8 result (STIF_FLAG_OUT_U8), else null
  cudaStream_t h2d, d2h;
  int bands;
};

// Number of timesteps whose Q table + flow stay resident at once in the banded host pipeline.
constexpr int kHostGroup = kMaxHostGroup;
size_t host_group_extra_bytes(int HH, int WW, int T) {
  const size_t Q = (size_t)HH * WW;
  return (size_t)(std::min(T, kHostGroup) - 1) * (align256(Q * 128 * 2) + align256(Q * 4 * sizeof(float)));
}

// Band-major host pipeline (bf16 mode).  The latent arrives in LR row bands; everything downstream is stream-ordered
// behind the band that makes it computable, so uploads, kernels and downloads of different bands overlap:
//   band k:  H2D(k) -> K0(k) -> K1(t, HR rows [he[k-1], he[k]))  for every t of the group
//                            -> K2(t, HR rows [ge[k-1], ge[k]))  -> D2H of those RGB rows
// he[k] = first HR row whose nearest / bilinear LR footprint is not yet covered by bands 0..k (exact, from the axis
// tables).  ge[k] = he[k] - halo is SPECULATIVE: stage D reads stage-A rows at flow-displaced positions, so K2 of a
// band trails K1 by `halo` HR rows and raises the workspace flag when a warp reaches a row that is not there yet
// (the same check stif_decode_rows uses).  If that ever happens the group's K2 launches are repeated on the complete
// tables after the loop, and the handle doubles its halo for the following calls (flows of a video are consistent).
// `workspace` must hold stif_workspace_bytes() + host_group_extra_bytes().
int decode_host_banded(stif_decoder* d, const float* latent, const float* frames, int B, int H, int W, int HH, int WW,
                       const float* times, int T, int mode, void* workspace, float* out, cudaStream_t stream, const HostPipe& hp) {
  const Geometry* geo = nullptr;
  if (int rc = get_geometry(d, H, W, HH, WW, stream, &geo)) return rc;
  const Workspace ws = carve_workspace(workspace, H, W, HH, WW, mode);
  LaunchCtx cx{stream, &d->launches, d->num_sms};
  const size_t Q = (size_t)HH * WW, plane = (size_t)H * W;
  const int G = std::min(T, kHostGroup);
  const size_t qtab_b = align256(Q * 128 * 2), flow_b = align256(Q * 4 * sizeof(float));
  auto slab_ws = [&](int g) {   // timestep g of the group: its own Q table and flow
    Workspace w = ws;
    if (g > 0) {
      char* extra = (char*)workspace + ws.total_bytes + (size_t)(g - 1) * (qtab_b + flow_b);
      w.qtab = extra;
      w.flow = (float*)(extra + qtab_b);
    }
    return w;
  };
  const int halo = std::min(HH, std::max(1, d->host_halo));
  const HostBandPlan plan = plan_host_bands(H, W, HH, WW, G, hp.bands, d->host_bands_forced, halo, d->num_sms);
  const std::vector<int>&he = plan.he, &ge = plan.ge, &lr_end = plan.lr_end;
  const int nbands = (int)he.size();
  std::vector<cudaEvent_t> used_events;
  PipeGuard guard{d, stream, hp.h2d, hp.d2h, &used_events};
  auto chain = [&](cudaStream_t from, cudaStream_t to) -> cudaError_t {   // `to` waits for what `from` holds now
    cudaEvent_t ev = take_event(d);
    used_events.push_back(ev);
    cudaError_t e = cudaEventRecord(ev, from);
    return e != cudaSuccess ? e : cudaStreamWaitEvent(to, ev, 0);
  };
  // STIF_HOST_MULTI=0: one launch per (band, timestep) as in round 1; default: the resident timesteps of a band share one K1 and
  // one K2 launch (decode_multi_tc) -- half the launches of a T=2 call, and each launch's rotation fill / drain is paid once
  static const bool multi = !(getenv("STIF_HOST_MULTI") && atoi(getenv("STIF_HOST_MULTI")) == 0);
  static_assert(kHostGroup <= kMaxSlabsHost, "decode_multi_tc takes at most kMaxSlabsHost timesteps");
  // K2 of RGB rows [g0,g1) for timesteps [c0,c1) back to back (nothing between the launches, so each one's prologue
  // overlaps its predecessor's tail), then one event and the downloads of those rows
  auto k2_rows = [&](int b, int c0, int c1, bool resident, int g0, int g1, int k1_hi) -> int {
    if (g1 <= g0) return STIF_OK;
    if (multi && resident && c1 - c0 > 1) {   // all resident timesteps of these rows in one launch
      ScopedSpan sp(d, stream, 2);
      Workspace wsv[kHostGroup];
      float tv[kHostGroup];
      float* ov[kHostGroup];
      uint8_t* o8[kHostGroup];
      for (int c = c0; c < c1; ++c) {
        const size_t slab = ((size_t)c * B + b) * 3 * Q;
        wsv[c - c0] = slab_ws(c);
        tv[c - c0] = times[(size_t)c * B + b];
        ov[c - c0] = out + slab;
        o8[c - c0] = hp.out_u8_dev ? hp.out_u8_dev + slab : nullptr;
      }
      cudaError_t e = decode_multi_tc(cx, d->tcw, *geo, wsv, tv, c1 - c0, g0, g1, 0, k1_hi, ov, hp.out_u8_dev ? o8 : nullptr, 2);
      if (e != cudaSuccess) return set_error(STIF_ECUDA, "decode kernels failed: %s", cudaGetErrorString(e));
    } else
    for (int c = c0; c < c1; ++c) {
      ScopedSpan sp(d, stream, 2);
      const size_t slab = ((size_t)c * B + b) * 3 * Q;
      cudaError_t e = decode_slab_tc(cx, d->tcw, *geo, resident ? slab_ws(c) : ws, times[(size_t)c * B + b], g0, g1, 0, k1_hi,
                                     out + slab, 2, hp.out_u8_dev ? hp.out_u8_dev + slab : nullptr);
      if (e != cudaSuccess) return set_error(STIF_ECUDA, "decode kernels failed: %s", cudaGetErrorString(e));
    }
    CUDA_OR_RETURN(chain(stream, hp.d2h));
    for (int c = c0; c < c1; ++c) {
      const size_t slab = ((size_t)c * B + b) * 3 * Q, o = slab + (size_t)g0 * WW;
      if (hp.out_u8_dev)   // HWC rows are contiguous
        CUDA_OR_RETURN(cudaMemcpyAsync((uint8_t*)hp.out_host + slab + (size_t)g0 * WW * 3, hp.out_u8_dev + slab + (size_t)g0 * WW * 3,
                                       (size_t)(g1 - g0) * WW * 3, cudaMemcpyDeviceToHost, hp.d2h));
      else
        CUDA_OR_RETURN(cudaMemcpy2DAsync((float*)hp.out_host + o, Q * 4, out + o, Q * 4, (size_t)(g1 - g0) * WW * 4, 3, cudaMemcpyDeviceToHost,
                                         hp.d2h));
    }
    return STIF_OK;
  };
  // STIF_HOST_TIMELINE=1: print when each band's upload / stage A+B / stage C-E / download finished (ms since the
  // first copy was queued).  Debug aid; the extra event records sit between kernels.
  static const bool timeline = getenv("STIF_HOST_TIMELINE") != nullptr;
  struct Mark { cudaEvent_t ev; const char* what; int b, k; };
  std::vector<Mark> marks;
  auto mark = [&](cudaStream_t st, const char* what, int b, int k) {
    if (!timeline) return;
    cudaEvent_t ev = take_event(d);
    used_events.push_back(ev);
    cudaEventRecord(ev, st);
    marks.push_back({ev, what, b, k});
  };
  mark(hp.h2d, "start", 0, 0);
  CUDA_OR_RETURN(cudaMemsetAsync(ws.flag, 0, sizeof(int), stream));
  // all uploads are queued up front (they depend on nothing); one event per (item, band)
  std::vector<cudaEvent_t> landed((size_t)B * nbands);
  for (int b = 0; b < B; ++b)
    for (int k = 0; k < nbands; ++k) {
      const int r0 = k ? lr_end[k - 1] : 0, r1 = lr_end[k];
      const size_t off = (size_t)b * 192 * plane + (size_t)r0 * W, offf = (size_t)b * 6 * plane + (size_t)r0 * W;
      const size_t width = (size_t)(r1 - r0) * W * sizeof(float);
      const size_t le = (size_t)hp.latent_elem;
      if (r1 > r0) {
        CUDA_OR_RETURN(cudaMemcpy2DAsync((char*)latent + off * le, plane * le, (const char*)hp.latent_host + off * le, plane * le,
                                         (size_t)(r1 - r0) * W * le, 192, cudaMemcpyHostToDevice, hp.h2d));
        CUDA_OR_RETURN(cudaMemcpy2DAsync((float*)frames + offf, plane * 4, hp.frames_host + offf, plane * 4, width, 6,
                                         cudaMemcpyHostToDevice, hp.h2d));
      }
      cudaEvent_t ev = take_event(d);
      used_events.push_back(ev);
      CUDA_OR_RETURN(cudaEventRecord(ev, hp.h2d));
      landed[(size_t)b * nbands + k] = ev;
    }
  for (int b = 0; b < B; ++b) {
    const float* lat_b = (const float*)((const char*)latent + (size_t)b * 192 * plane * hp.latent_elem);
    const float* fr_b = frames + (size_t)b * 6 * plane;
    for (int k = 0; k < nbands; ++k) {
      const int r0 = k ? lr_end[k - 1] : 0, r1 = lr_end[k];
      CUDA_OR_RETURN(cudaStreamWaitEvent(stream, landed[(size_t)b * nbands + k], 0));
      if (r1 > r0) {
        ScopedSpan sp(d, stream, 0);
        CUDA_OR_RETURN(project_latent_tc(cx, d->tcw, lat_b, fr_b, H, W, ws.tab, r0, r1, false, hp.latent_elem == 2));
      }
      mark(stream, "K0", b, k);
      const int h0 = k ? he[k - 1] : 0, h1 = he[k];
      if (multi && G > 1 && h1 > h0) {
        ScopedSpan sp(d, stream, 1);
        Workspace wsv[kHostGroup];
        float tv[kHostGroup];
        for (int g = 0; g < G; ++g) { wsv[g] = slab_ws(g); tv[g] = times[(size_t)g * B + b]; }
        cudaError_t e = decode_multi_tc(cx, d->tcw, *geo, wsv, tv, G, 0, HH, h0, h1, nullptr, nullptr, 1);
        if (e != cudaSuccess) return set_error(STIF_ECUDA, "decode kernels failed: %s", cudaGetErrorString(e));
      } else
      for (int g = 0; g < G && h1 > h0; ++g) {
        ScopedSpan sp(d, stream, 1);
        cudaError_t e = decode_slab_tc(cx, d->tcw, *geo, slab_ws(g), times[(size_t)g * B + b], 0, HH, h0, h1, out, 1);
        if (e != cudaSuccess) return set_error(STIF_ECUDA, "decode kernels failed: %s", cudaGetErrorString(e));
      }
      mark(stream, "K1", b, k);
      if (int rc = k2_rows(b, 0, G, true, k ? ge[k - 1] : 0, ge[k], h1)) return rc;
      mark(stream, "K2", b, k);
      mark(hp.d2h, "D2H", b, k);
    }
    // speculation check for this item (the flag is only read here, after the tables are complete)
    int flag = 0;
    if (nbands > 1) {
      CUDA_OR_RETURN(cudaMemcpyAsync(&flag, ws.flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
      CUDA_OR_RETURN(cudaStreamSynchronize(stream));
    }
    if (flag) {
      d->host_halo = std::min(HH, 2 * halo);
      ++d->host_respins;
      CUDA_OR_RETURN(cudaMemsetAsync(ws.flag, 0, sizeof(int), stream));
      if (int rc = k2_rows(b, 0, G, true, 0, HH, HH)) return rc;
    }
    // timesteps beyond the resident group: one slab at a time on complete tables; the download of slab c overlaps
    // slab c+1, and the last slab is decoded in row bands so that only its last band's download is exposed
    for (int c = G; c < T; ++c) {
      {
        ScopedSpan sp(d, stream, 1);
        cudaError_t e = decode_slab_tc(cx, d->tcw, *geo, ws, times[(size_t)c * B + b], 0, HH, 0, HH, out, 1);
        if (e != cudaSuccess) return set_error(STIF_ECUDA, "decode kernels failed: %s", cudaGetErrorString(e));
      }
      const int parts = (c == T - 1 && b == B - 1) ? nbands : 1;
      for (int k = 0; k < parts; ++k)
        if (int rc = k2_rows(b, c, c + 1, false, (int)((long)HH * k / parts) & ~7,
                             k + 1 == parts ? HH : (int)((long)HH * (k + 1) / parts) & ~7, HH))
          return rc;
    }
  }
  CUDA_OR_RETURN(cudaStreamSynchronize(hp.d2h));
  CUDA_OR_RETURN(cudaStreamSynchronize(stream));
  if (timeline) {
    for (int b = 0; b < B; ++b)
      for (int k = 0; k < nbands; ++k) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, marks[0].ev, landed[(size_t)b * nbands + k]);
        fprintf(stderr, "[stif host] b%d band %d: H2D %.3f", b, k, ms);
        for (size_t i = 1; i < marks.size(); ++i)
          if (marks[i].b == b && marks[i].k == k) {
            cudaEventElapsedTime(&ms, marks[0].ev, marks[i].ev);
            fprintf(stderr, "  %s %.3f", marks[i].what, ms);
          }
        fprintf(stderr, "   (HR rows A+B < %d, C-E < %d)\n", he[k], ge[k]);
      }
  }
  guard.ok = true;
  const Workspace last = slab_ws(T <= G ? T - 1 : 0);
  d->last_flow = last.flow;
  d->last_flow_floats = Q * 4;
  return STIF_OK;
}

int decode_impl(stif_decoder* d, const float* latent, const float* frames, int B, int H, int W, int HH, int WW,
                const float* times, int T, int mode, int row_begin, int row_end, int halo, void* workspace,
                size_t workspace_bytes, void* out_any, cudaStream_t stream, bool check_band, const HostPipe* hp = nullptr,
                int col_begin = 0, int col_end = -1) {
  float* out = (float*)out_any;   // fp32 [T,B,3,HH,WW], or with STIF_FLAG_OUT_U8 uint8 [T,B,HH,WW,3] (see out_u8 below)
  if (!d) return set_error(STIF_EINVAL, "null decoder");
  if (!d->weights_loaded) return set_error(STIF_ESTATE, "stif_load_weights has not been called");
  if (!latent || !frames || !times || !out || !workspace) return set_error(STIF_EINVAL, "null buffer");
  if (int rc = check_shape(B, H, W, HH, WW, T)) return rc;
  const int prec = mode & 0xFF;
  if (prec != STIF_MODE_BF16 && prec != STIF_MODE_FP32) return set_error(STIF_EINVAL, "unknown mode 0x%x", mode);
  const bool ensemble = (mode & STIF_FLAG_LOCAL_ENSEMBLE) != 0, u8 = (mode & STIF_FLAG_OUT_U8) != 0;
  const bool test_variant = (mode & STIF_FLAG_TEST_VARIANT) != 0, warp_from_coord = (mode & STIF_FLAG_WARP_FROM_COORD) != 0;
  // tensor-core kernels, decoding_test: at x4 the upsampled-frame grid IS the query grid and the frame terms ride inside the Q
  // planes (k1_tile_loop, UPF; "fast"); at any other size -- and for decoding_memory's coordinate warp base -- stage B's term is
  // resampled onto the query grid once per pair and stage D's terms are gathered per slab at the warped positions (tc_general)
  const bool tc_variant = test_variant && prec == STIF_MODE_BF16;
  const bool tc_general = tc_variant && (HH != 4 * H || WW != 4 * W || warp_from_coord);
  if (col_end < 0) col_end = WW;
  if (col_begin < 0 || col_end > WW || col_begin >= col_end)
    return set_error(STIF_EINVAL, "invalid column window [%d,%d) for WW=%d", col_begin, col_end, WW);
  if ((test_variant || warp_from_coord) && ensemble)
    return set_error(STIF_EINVAL, "STIF_FLAG_TEST_VARIANT / STIF_FLAG_WARP_FROM_COORD cannot be combined with STIF_FLAG_LOCAL_ENSEMBLE");
  if (warp_from_coord && prec == STIF_MODE_BF16 && !test_variant)
    return set_error(STIF_EINVAL, "STIF_FLAG_WARP_FROM_COORD with STIF_MODE_BF16 needs STIF_FLAG_TEST_VARIANT (decoding_memory, Sakuya_arch_test.py:600-861)");
  if ((col_begin != 0 || col_end != WW) && (ensemble || hp))
    return set_error(STIF_EINVAL, "a column window needs device buffers and no STIF_FLAG_LOCAL_ENSEMBLE");
  if (ensemble && (B != 1 || row_begin != 0 || row_end != HH || hp))
    return set_error(STIF_EINVAL, "STIF_FLAG_LOCAL_ENSEMBLE needs B == 1 (Sakuya_arch_test.py:989) and a full raster on device buffers");
  // The tensor-core K2 gather stages tap addresses as 32-bit BYTE offsets (256 B per HR pixel, 512 B per LR texel,
  // kernels_tc.cu k2_gather_taps): they wrap above 2^24 HR pixels / 2^23 LR texels.  Refuse instead of gathering wrong rows.
  if (prec == STIF_MODE_BF16 && ((long long)HH * WW > (1ll << 24) || (long long)H * W >= (1ll << 23)))
    return set_error(STIF_EINVAL, "raster too large for STIF_MODE_BF16 (HH*WW=%lld > 2^24 or H*W=%lld >= 2^23); decode it in row-band "
                     "sized pieces of a smaller raster or use STIF_MODE_FP32", (long long)HH * WW, (long long)H * W);
  if (row_begin < 0 || row_end > HH || row_begin >= row_end || halo < 0)
    return set_error(STIF_EINVAL, "invalid row band [%d,%d) halo %d for HH=%d", row_begin, row_end, halo, HH);
  const size_t need = stif_workspace_bytes(B, H, W, HH, WW, T, mode);
  if (workspace_bytes < need)
    return set_error(STIF_ENOMEM, "workspace too small: %zu bytes given, %zu needed", workspace_bytes, need);
  CUDA_OR_RETURN(cudaSetDevice(d->device));
  if (hp && prec == STIF_MODE_BF16 && !ensemble && !test_variant) return decode_host_banded(d, latent, frames, B, H, W, HH, WW, times, T, mode, workspace, out, stream, *hp);
  if (hp && hp->latent_elem != 4)
    return set_error(STIF_EINVAL, "bf16 host latents (stif_decode_host_bf16) are supported by plain STIF_MODE_BF16 decodes only");
  const Geometry* geo = nullptr;
  if (int rc = get_geometry(d, H, W, HH, WW, stream, &geo, warp_from_coord)) return rc;
  Workspace ws = carve_workspace(workspace, H, W, HH, WW, mode);
  LaunchCtx cx{stream, &d->launches, d->num_sms};
  const int k1_lo = std::max(0, row_begin - halo), k1_hi = std::min(HH, row_end + halo);
  const size_t Q = (size_t)HH * WW;
  CUDA_OR_RETURN(cudaMemsetAsync(ws.flag, 0, sizeof(int), stream));
  std::vector<cudaEvent_t> used_events;
  PipeGuard guard{d, stream, hp ? hp->h2d : nullptr, hp ? hp->d2h : nullptr, &used_events};
  if (!hp) guard.ok = true;   // device-buffer calls are purely stream-ordered: nothing of the caller's to drain
  const int nbands = hp ? std::max(1, std::min(hp->bands, H)) : 1;
  const size_t plane = (size_t)H * W;
  for (int b = 0; b < B; ++b) {
    const float* lat_b = latent + (size_t)b * 192 * plane;
    const float* fr_b = frames + (size_t)b * 6 * plane;
    for (int k = 0; k < nbands; ++k) {
      const int r0 = (int)((long)H * k / nbands), r1 = (int)((long)H * (k + 1) / nbands);
      if (hp) {   // upload this row band of all 198 channels, then let the compute stream wait for it
        const size_t off = (size_t)r0 * W, width = (size_t)(r1 - r0) * W * sizeof(float);
        CUDA_OR_RETURN(cudaMemcpy2DAsync((float*)lat_b + off, plane * 4, hp->latent_host + (size_t)b * 192 * plane + off, plane * 4,
                                         width, 192, cudaMemcpyHostToDevice, hp->h2d));
        CUDA_OR_RETURN(cudaMemcpy2DAsync((float*)fr_b + off, plane * 4, hp->frames_host + (size_t)b * 6 * plane + off, plane * 4,
                                         width, 6, cudaMemcpyHostToDevice, hp->h2d));
        cudaEvent_t ev = take_event(d);
        used_events.push_back(ev);
        CUDA_OR_RETURN(cudaEventRecord(ev, hp->h2d));
        CUDA_OR_RETURN(cudaStreamWaitEvent(stream, ev, 0));
      }
      ScopedSpan sp(d, stream, 0);
      if (prec == STIF_MODE_BF16) {
        CUDA_OR_RETURN(project_latent_tc(cx, d->tcw, lat_b, fr_b, H, W, ws.tab, r0, r1, tc_variant));
        if (tc_variant && k == nbands - 1) {
          CUDA_OR_RETURN(project_frames_up4_tc(cx, d->tcw, fr_b, H, W, ws.utab));
          if (tc_general) CUDA_OR_RETURN(resample_ub_tc(cx, ws.utab, *geo, ws.uq));
        }
      }
      else if (k == nbands - 1) {
        CUDA_OR_RETURN(project_latent(cx, d->w32, lat_b, fr_b, H, W, ws.tab, false, test_variant, ws.act_a, ws.chunk));
        if (test_variant) CUDA_OR_RETURN(project_frames_up4(cx, d->w32, fr_b, H, W, (float*)ws.utab));
      }
    }
    for (int c = 0; c < T; ++c) {
      const float t = times[(size_t)c * B + b];
      uint8_t* out_u8 = u8 ? (uint8_t*)out_any + ((size_t)c * B + b) * 3 * Q : nullptr;
      float* out_slab = u8 ? ws.rgb32 : out + ((size_t)c * B + b) * 3 * Q;   // uint8 output: decode into the staging slab
      if (ensemble) {
        DeviceGeometry* dg = nullptr;
        if (int rc = get_ensemble_geometry(d, H, W, HH, WW, &dg)) return rc;
        ScopedSpan sp(d, stream, 1);
        cudaError_t e = prec == STIF_MODE_FP32
                            ? decode_slab_fp32_ensemble(cx, d->w32, d->hw, dg->geo_pass, dg->ens_y, dg->ens_x, ws, t, out_slab)
                            : decode_slab_tc_ensemble(cx, d->tcw, dg->geo_pass, dg->ens_y, dg->ens_x, ws, t, out_slab);
        if (e != cudaSuccess) return set_error(STIF_ECUDA, "decode kernels failed: %s", cudaGetErrorString(e));
      }
      const bool u8_fused = u8 && prec == STIF_MODE_BF16 && !ensemble;   // K2's output stage converts (no fp32 staging slab)
      for (int stage = 1; stage <= 2 && !ensemble; ++stage) {
        ScopedSpan sp(d, stream, stage);
        cudaError_t e;
        if (prec == STIF_MODE_FP32) {
          e = decode_slab_fp32(cx, d->w32, d->hw, *geo, ws, t, row_begin, row_end, k1_lo, k1_hi, out_slab, stage);
        } else if (stage == 1) {
          Workspace w1 = ws;
          if (tc_general) w1.utab = ws.uq;   // the x4 kernel with "UB at the query | zeros": nothing is folded into the Q planes
          e = decode_slab_tc(cx, d->tcw, *geo, w1, t, row_begin, row_end, k1_lo, k1_hi, out_slab, tc_variant ? 5 : 1);
        } else {
          e = tc_general ? warp_u_terms_tc(cx, ws.utab, ws.flow, *geo, row_begin, row_end, ws.uadd) : cudaSuccess;
          if (e == cudaSuccess)
            e = decode_slab_tc(cx, d->tcw, *geo, ws, t, row_begin, row_end, k1_lo, k1_hi, out_slab, 2, u8_fused ? out_u8 : nullptr,
                               col_begin, col_end, tc_general ? ws.uadd : nullptr);
        }
        if (e != cudaSuccess) return set_error(STIF_ECUDA, "decode kernels failed: %s", cudaGetErrorString(e));
      }
      if (u8 && !u8_fused) CUDA_OR_RETURN(rgb_to_u8_hwc(cx, ws.rgb32, out_u8, HH, WW, row_begin, row_end));
      if (hp) {   // download this slab while the next one is decoded
        cudaEvent_t ev = take_event(d);
        used_events.push_back(ev);
        CUDA_OR_RETURN(cudaEventRecord(ev, stream));
        CUDA_OR_RETURN(cudaStreamWaitEvent(hp->d2h, ev, 0));
        if (u8)
          CUDA_OR_RETURN(cudaMemcpyAsync((uint8_t*)hp->out_host + ((size_t)c * B + b) * 3 * Q, out_u8, 3 * Q, cudaMemcpyDeviceToHost, hp->d2h));
        else
          CUDA_OR_RETURN(cudaMemcpyAsync((float*)hp->out_host + ((size_t)c * B + b) * 3 * Q, out_slab, 3 * Q * sizeof(float),
                                         cudaMemcpyDeviceToHost, hp->d2h));
      }
    }
  }
  if (hp) {
    CUDA_OR_RETURN(cudaStreamSynchronize(hp->d2h));
    CUDA_OR_RETURN(cudaStreamSynchronize(stream));
    guard.ok = true;
  }
  d->last_flow = ws.flow;
  d->last_flow_floats = Q * 4;
  if (check_band) {
    int flag = 0;
    CUDA_OR_RETURN(cudaMemcpyAsync(&flag, ws.flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CUDA_OR_RETURN(cudaStreamSynchronize(stream));
    if (flag)
      return set_error(STIF_EINVAL, "row band [%d,%d): a warp reached outside the %d-row halo; increase halo", row_begin,
                       row_end, halo);
  }
  return STIF_OK;
}

}  // namespace

extern "C" {

int stif_abi_version(void) { return STIF_ABI_VERSION; }

const char* stif_last_error(void) { return g_last_error.c_str(); }

int stif_create(stif_decoder_t** out, int device) {
  if (!out) return set_error(STIF_EINVAL, "null out pointer");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return set_error(STIF_ENODEV, "no CUDA device available (%s); this library has no CPU fallback",
                     e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return set_error(STIF_ENODEV, "device %d out of range (count %d)", device, n);
  cudaDeviceProp prop;
  CUDA_OR_RETURN(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return set_error(STIF_ENODEV, "device %d is sm_%d%d; libstif_b200 is built for sm_100a only", device, prop.major,
                     prop.minor);
  CUDA_OR_RETURN(cudaSetDevice(device));
  auto* d = new stif_decoder();
  d->device = device;
  d->num_sms = prop.multiProcessorCount;
  *out = d;
  return STIF_OK;
}

int stif_destroy(stif_decoder_t* d) {
  if (!d) return STIF_OK;
  cudaSetDevice(d->device);
  for (auto& kv : d->geos) {
    cudaFree(kv.second.blob);
    for (void* eb : kv.second.ens_blob) if (eb) cudaFree(eb);
  }
  if (d->d_w32) cudaFree(d->d_w32);
  if (d->tcw) tc_weights_destroy(d->tcw);
  if (d->hpw) hp_weights_destroy(d->hpw);
  if (d->host_scratch) cudaFree(d->host_scratch);
  if (d->host_stream) cudaStreamDestroy(d->host_stream);
  if (d->h2d_stream) cudaStreamDestroy(d->h2d_stream);
  if (d->d2h_stream) cudaStreamDestroy(d->d2h_stream);
  for (auto& sp : d->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
  for (auto e : d->event_pool) cudaEventDestroy(e);
  delete d;
  return STIF_OK;
}

int stif_load_weights(stif_decoder_t* d, const float* const* tensors, int num_tensors) {
  if (!d || !tensors) return set_error(STIF_EINVAL, "null argument");
  if (num_tensors != STIF_NUM_WEIGHT_TENSORS)
    return set_error(STIF_EINVAL, "expected %d weight tensors, got %d", STIF_NUM_WEIGHT_TENSORS, num_tensors);
  for (int i = 0; i < num_tensors; ++i)
    if (!tensors[i]) return set_error(STIF_EINVAL, "weight tensor %d is null", i);
  CUDA_OR_RETURN(cudaSetDevice(d->device));
  fold_weights(tensors, d->hw);
  const FoldedWeights& h = d->hw;
  const std::vector<float>* parts[] = {&h.w_tab, &h.w_tab_lat, &h.w_up, &h.a_rel, &h.a_t, &h.a_b, &h.f1_w, &h.f1_b, &h.f2_w, &h.f2_b, &h.f3_w, &h.f3_b,
                                       &h.b_t, &h.b_b, &h.l1_w, &h.l1_b, &h.l2_w, &h.l2_b, &h.l3_w, &h.l3_b,
                                       &h.e_t, &h.e_b, &h.e1_w, &h.e1_b, &h.e2_w, &h.e2_b, &h.e3_w, &h.e3_b, &h.e4_w, &h.e4_b};
  constexpr int NP = sizeof(parts) / sizeof(parts[0]);
  size_t offs[NP], total = 0;
  for (int i = 0; i < NP; ++i) { offs[i] = total; total += (parts[i]->size() + 63) & ~size_t(63); }
  std::vector<float> host(total, 0.f);
  for (int i = 0; i < NP; ++i) memcpy(host.data() + offs[i], parts[i]->data(), parts[i]->size() * sizeof(float));
  if (d->d_w32) { cudaFree(d->d_w32); d->d_w32 = nullptr; }
  CUDA_OR_RETURN(cudaMalloc(&d->d_w32, total * sizeof(float)));
  CUDA_OR_RETURN(cudaMemcpy(d->d_w32, host.data(), total * sizeof(float), cudaMemcpyHostToDevice));
  const float** fields[] = {&d->w32.w_tab, &d->w32.w_tab_lat, &d->w32.w_up, &d->w32.a_rel, &d->w32.a_t, &d->w32.a_b, &d->w32.f1_w, &d->w32.f1_b, &d->w32.f2_w,
                            &d->w32.f2_b, &d->w32.f3_w, &d->w32.f3_b, &d->w32.b_t, &d->w32.b_b, &d->w32.l1_w, &d->w32.l1_b,
                            &d->w32.l2_w, &d->w32.l2_b, &d->w32.l3_w, &d->w32.l3_b, &d->w32.e_t, &d->w32.e_b, &d->w32.e1_w,
                            &d->w32.e1_b, &d->w32.e2_w, &d->w32.e2_b, &d->w32.e3_w, &d->w32.e3_b, &d->w32.e4_w, &d->w32.e4_b};
  for (int i = 0; i < NP; ++i) *fields[i] = d->d_w32 + offs[i];
  if (d->tcw) { tc_weights_destroy(d->tcw); d->tcw = nullptr; }
  std::string err;
  d->tcw = tc_weights_create(d->hw, err);
  if (!d->tcw) return set_error(STIF_ECUDA, "packing tensor-core weights failed: %s", err.c_str());
  if (d->hpw) { hp_weights_destroy(d->hpw); d->hpw = nullptr; }
  d->w32.hp = nullptr;
  if (!getenv("STIF_FP32_SIMT")) {   // STIF_FP32_SIMT=1 keeps the SIMT SGEMM (test anchor for the tensor-core split GEMM)
    d->hpw = hp_weights_create(d->hw, err);
    if (!d->hpw) return set_error(STIF_ECUDA, "packing split-bf16 weights failed: %s", err.c_str());
    d->w32.hp = d->hpw;
  }
  CUDA_OR_RETURN(cudaDeviceSynchronize());
  d->weights_loaded = true;
  return STIF_OK;
}

size_t stif_workspace_bytes(int B, int H, int W, int HH, int WW, int T, int mode) {
  if (B < 1 || H < 1 || W < 1 || HH < 1 || WW < 1 || T < 1) return 0;
  return carve_workspace(nullptr, H, W, HH, WW, mode).total_bytes;
}

int stif_prepare(stif_decoder_t* d, int H, int W, int HH, int WW, int mode) {
  if (!d) return set_error(STIF_EINVAL, "null decoder");
  if (int rc = check_shape(1, H, W, HH, WW, 1)) return rc;
  CUDA_OR_RETURN(cudaSetDevice(d->device));
  const Geometry* geo = nullptr;
  if (int rc = get_geometry(d, H, W, HH, WW, nullptr, &geo, (mode & STIF_FLAG_WARP_FROM_COORD) != 0)) return rc;
  if (mode & STIF_FLAG_LOCAL_ENSEMBLE) {
    DeviceGeometry* dg = nullptr;
    if (int rc = get_ensemble_geometry(d, H, W, HH, WW, &dg)) return rc;
  }
  return STIF_OK;
}

int stif_decode(stif_decoder_t* d, const float* latent, const float* frames, int B, int H, int W, int HH, int WW,
                const float* times, int T, int mode, void* workspace, size_t workspace_bytes, void* out, void* stream) {
  return decode_impl(d, latent, frames, B, H, W, HH, WW, times, T, mode, 0, HH, 0, workspace, workspace_bytes, out,
                     (cudaStream_t)stream, false);
}

int stif_decode_rows(stif_decoder_t* d, const float* latent, const float* frames, int B, int H, int W, int HH, int WW,
                     const float* times, int T, int mode, int row_begin, int row_end, int halo, void* workspace,
                     size_t workspace_bytes, void* out, void* stream) {
  return decode_impl(d, latent, frames, B, H, W, HH, WW, times, T, mode, row_begin, row_end, halo, workspace, workspace_bytes,
                     out, (cudaStream_t)stream, true);
}

namespace {
int decode_host_entry(stif_decoder_t* d, const void* latent_host, int latent_elem, const float* frames_host, int B, int H, int W, int HH,
                      int WW, const float* times, int T, int mode, void* out_host) {
  if (!d) return set_error(STIF_EINVAL, "null decoder");
  if (!latent_host || !frames_host || !out_host) return set_error(STIF_EINVAL, "null buffer");
  if (int rc = check_shape(B, H, W, HH, WW, T)) return rc;
  CUDA_OR_RETURN(cudaSetDevice(d->device));
  if (!d->host_stream) {
    CUDA_OR_RETURN(cudaStreamCreateWithFlags(&d->host_stream, cudaStreamNonBlocking));
    CUDA_OR_RETURN(cudaStreamCreateWithFlags(&d->h2d_stream, cudaStreamNonBlocking));
    CUDA_OR_RETURN(cudaStreamCreateWithFlags(&d->d2h_stream, cudaStreamNonBlocking));
  }
  const size_t lat_b = align256((size_t)B * 192 * H * W * latent_elem), fr_b = align256((size_t)B * 6 * H * W * 4);
  const size_t out_b = align256((size_t)T * B * 3 * HH * WW * 4);
  const size_t ws_b = stif_workspace_bytes(B, H, W, HH, WW, T, mode) + host_group_extra_bytes(HH, WW, T);
  const size_t out8_b = (mode & STIF_FLAG_OUT_U8) ? align256((size_t)T * B * 3 * HH * WW) : 0;
  const size_t total = lat_b + fr_b + out_b + ws_b + out8_b;
  if (d->host_scratch_bytes < total) {
    if (d->host_scratch) cudaFree(d->host_scratch);
    d->host_scratch = nullptr;
    d->host_scratch_bytes = 0;
    CUDA_OR_RETURN(cudaMalloc(&d->host_scratch, total));
    d->host_scratch_bytes = total;
  }
  char* base = (char*)d->host_scratch;
  float* lat = (float*)base;
  float* fr = (float*)(base + lat_b);
  float* out = (float*)(base + lat_b + fr_b);
  void* ws = base + lat_b + fr_b + out_b;
  cudaStream_t s = d->host_stream;
  HostPipe hp{latent_elem, (const float*)latent_host, frames_host, out_host,
              out8_b ? (uint8_t*)base + lat_b + fr_b + out_b + ws_b : nullptr, d->h2d_stream, d->d2h_stream, d->host_bands};
  return decode_impl(d, lat, fr, B, H, W, HH, WW, times, T, mode, 0, HH, 0, ws, ws_b, out, s, false, &hp);
}
}  // namespace

int stif_decode_window(stif_decoder_t* d, const float* latent, const float* frames, int B, int H, int W, int HH, int WW,
                       const float* times, int T, int mode, int row_begin, int row_end, int col_begin, int col_end, int halo,
                       void* workspace, size_t workspace_bytes, void* out, void* stream) {
  return decode_impl(d, latent, frames, B, H, W, HH, WW, times, T, mode, row_begin, row_end, halo, workspace, workspace_bytes,
                     out, (cudaStream_t)stream, true, nullptr, col_begin, col_end);
}

int stif_decode_host(stif_decoder_t* d, const float* latent_host, const float* frames_host, int B, int H, int W, int HH,
                     int WW, const float* times, int T, int mode, void* out_host) {
  return decode_host_entry(d, latent_host, 4, frames_host, B, H, W, HH, WW, times, T, mode, out_host);
}

int stif_decode_host_bf16(stif_decoder_t* d, const uint16_t* latent_bf16_host, const float* frames_host, int B, int H, int W,
                          int HH, int WW, const float* times, int T, int mode, void* out_host) {
  if ((mode & 0xFF) != STIF_MODE_BF16 || (mode & (STIF_FLAG_LOCAL_ENSEMBLE | STIF_FLAG_TEST_VARIANT | STIF_FLAG_WARP_FROM_COORD)))
    return set_error(STIF_EINVAL, "stif_decode_host_bf16: plain STIF_MODE_BF16 decodes only (STIF_FLAG_OUT_U8 allowed)");
  return decode_host_entry(d, latent_bf16_host, 2, frames_host, B, H, W, HH, WW, times, T, mode, out_host);
}

int stif_debug_host_pipeline(stif_decoder_t* d, int bands, int halo, int64_t* respins) {
  if (!d) return set_error(STIF_EINVAL, "null decoder");
  if (bands > 0) { d->host_bands = bands; d->host_bands_forced = true; }
  if (halo > 0) d->host_halo = halo;
  if (respins) *respins = d->host_respins;
  return STIF_OK;
}

int stif_debug_band_plan(int H, int W, int HH, int WW, int T, int bands, int bands_forced, int halo, int num_sms, int max_entries,
                         int* lr_end, int* ab_end, int* ce_end, double* cost_us) {
  if (H < 1 || W < 1 || HH < 1 || WW < 1 || T < 1 || bands < 1 || halo < 1 || num_sms < 1 || max_entries < 1 || !lr_end || !ab_end || !ce_end)
    return set_error(STIF_EINVAL, "invalid argument");
  const HostBandPlan pl = plan_host_bands(H, W, HH, WW, std::min(T, kMaxHostGroup), bands, bands_forced != 0, std::min(HH, halo), num_sms);
  const int n = (int)pl.he.size();
  if (n > max_entries) return set_error(STIF_ENOMEM, "plan has %d entries, room for %d", n, max_entries);
  for (int k = 0; k < n; ++k) { lr_end[k] = pl.lr_end[k]; ab_end[k] = pl.he[k]; ce_end[k] = pl.ge[k]; }
  if (cost_us) *cost_us = pl.cost_us;
  return n;
}

int stif_axis_tables(int n_lr, int n_hr, float* coord, int32_t* index, float* rel, float* base) {
  if (n_lr < 1 || n_hr < 1) return set_error(STIF_EINVAL, "invalid axis sizes %d -> %d", n_lr, n_hr);
  HostAxis a;
  build_axis(n_lr, n_hr, a);
  if (coord) memcpy(coord, a.coord.data(), (size_t)n_hr * 4);
  if (index) memcpy(index, a.idx.data(), (size_t)n_hr * 4);
  if (rel) memcpy(rel, a.rel.data(), (size_t)n_hr * 4);
  if (base) memcpy(base, a.base.data(), (size_t)n_hr * 4);
  return STIF_OK;
}

int stif_ensemble_weights(int H, int W, int HH, int WW, float* weights_host, size_t num_floats) {
  if (H < 1 || W < 1 || HH < 1 || WW < 1 || !weights_host) return set_error(STIF_EINVAL, "invalid argument");
  if (num_floats < (size_t)4 * HH * WW) return set_error(STIF_EINVAL, "weights buffer too small");
  ensemble_weights_host(H, W, HH, WW, weights_host);
  return STIF_OK;
}

int stif_debug_last_flow(stif_decoder_t* d, float* flow_host, size_t num_floats) {
  if (!d || !flow_host) return set_error(STIF_EINVAL, "null argument");
  if (!d->last_flow) return set_error(STIF_ESTATE, "no decode has run on this handle");
  if (num_floats < d->last_flow_floats)
    return set_error(STIF_EINVAL, "flow buffer too small: %zu floats given, %zu needed", num_floats, d->last_flow_floats);
  CUDA_OR_RETURN(cudaSetDevice(d->device));
  CUDA_OR_RETURN(cudaDeviceSynchronize());
  CUDA_OR_RETURN(cudaMemcpy(flow_host, d->last_flow, d->last_flow_floats * 4, cudaMemcpyDeviceToHost));
  return STIF_OK;
}

int64_t stif_launch_count(const stif_decoder_t* d) { return d ? d->launches : 0; }

int stif_profile_enable(stif_decoder_t* d, int enable) {
  if (!d) return set_error(STIF_EINVAL, "null decoder");
  d->profiling = enable != 0;
  return STIF_OK;
}

int stif_profile_read(stif_decoder_t* d, double* ms, int64_t* count) {
  if (!d || !ms || !count) return set_error(STIF_EINVAL, "null argument");
  CUDA_OR_RETURN(cudaSetDevice(d->device));
  CUDA_OR_RETURN(cudaDeviceSynchronize());
  for (int k = 0; k < 3; ++k) { ms[k] = 0.0; count[k] = 0; }
  for (auto& sp : d->spans) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, sp.a, sp.b) == cudaSuccess) { ms[sp.kind] += t; count[sp.kind] += 1; }
    d->event_pool.push_back(sp.a);
    d->event_pool.push_back(sp.b);
  }
  d->spans.clear();
  return STIF_OK;
}

int stif_selftest(int device, char* report, size_t cap) {
  std::string rep;
  int rc = tc_selftest(device, rep);
  if (report && cap) {
    size_t n = std::min(cap - 1, rep.size());
    memcpy(report, rep.data(), n);
    report[n] = 0;
  }
  if (rc != 0) set_error(STIF_ECUDA, "tcgen05 selftest failed");
  return rc;
}

}  // extern "C"
