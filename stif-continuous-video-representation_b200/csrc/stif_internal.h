// stif_internal.h -- host-side structures shared by the translation units of libstif_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "stif_common.cuh"

namespace stif {

// ---------------------------------------------------------------------------------------------
// Folded weights (pack_weights.cpp).  omega_0 = 30 is folded into every sine layer
// (sin(30(Wx+b)) = sin((30W)x + 30b), reference SIREN.py:45) and the three consumers of HRfeat
// are composed with feat_imnet's last linear layer.  All fp32, row-major [out,in].
struct FoldedWeights {
  std::vector<float> w_tab;          // [256,198]  rows: TA | TB | TE1 | TE2 ; cols: latent(192), frames(6)
  std::vector<float> w_tab_lat;      // decoding_test variant: same, frame columns of TB | TE1 | TE2 zeroed (they move to w_up)
  std::vector<float> w_up;           // decoding_test variant: [192,6] frame columns of TB | TE1 | TE2, applied to the x4-upsampled frames
  std::vector<float> a_rel;          // [64,2]     (rely, relx) columns of feat_imnet layer 0
  std::vector<float> a_t, a_b;       // [64]       t column / bias of feat_imnet layer 0
  std::vector<float> f1_w, f1_b;     // [64,64]
  std::vector<float> f2_w, f2_b;     // [256,64]
  std::vector<float> f3_w, f3_b;     // [192,256]  rows: F | Q1 | Q2   (composed)
  std::vector<float> b_t, b_b;       // [64]       flow_imnet layer 0
  std::vector<float> l1_w, l1_b;     // [64,64]
  std::vector<float> l2_w, l2_b;     // [256,64]
  std::vector<float> l3_w, l3_b;     // [4,256]    plain linear (not scaled)
  std::vector<float> e_t, e_b;       // [64]       encode_imnet layer 0
  std::vector<float> e1_w, e1_b;     // [64,64]
  std::vector<float> e2_w, e2_b;     // [256,64]
  std::vector<float> e3_w, e3_b;     // [256,256]
  std::vector<float> e4_w, e4_b;     // [3,256]    plain linear
};

// tensors: the 26 host pointers of stif_load_weights, in ABI order.
void fold_weights(const float* const* tensors, FoldedWeights& out);

// Split-bf16 tensor-core GEMMs of the high-precision mode (kernels_hp.cu): per dense layer the bf16 hi / lo SW128 images.
struct HpLayer { uint8_t* hi; uint8_t* lo; int N, K; };
struct HpWeights;
enum HpLayerId { HP_F1 = 0, HP_F2, HP_F3, HP_L1, HP_L2, HP_E1, HP_E2, HP_E3, HP_K0, HP_K0T, HP_NUM_LAYERS };   // K0 / K0T: the projection (K = 198 padded to 256), plain / decoding_test

// Device-side view of the fp32 copy (kernels_fp32.cu).
struct DeviceWeights32 {
  const HpWeights* hp = nullptr;   // non-null: the dense layers run on the tensor cores (hp_gemm) instead of the SIMT SGEMM
  const float *w_tab, *w_tab_lat, *w_up, *a_rel, *a_t, *a_b, *f1_w, *f1_b, *f2_w, *f2_b, *f3_w, *f3_b;
  const float *b_t, *b_b, *l1_w, *l1_b, *l2_w, *l2_b, *l3_w, *l3_b;
  const float *e_t, *e_b, *e1_w, *e1_b, *e2_w, *e2_b, *e3_w, *e3_b, *e4_w, *e4_b;
};

// ---------------------------------------------------------------------------------------------
// Axis tables (axis_tables.cpp), host side.
struct HostAxis {
  std::vector<float> coord, rel, bw, base;
  std::vector<int32_t> idx, b0, hidx;
};
// shift_sign = 0: the plain decode.  +-1: a local-ensemble pass (Sakuya_arch_test.py:981-995): every gather uses the
// coordinate shifted by sign/n_lr + 1e-6 and re-clamped, `rel` keeps the un-shifted coordinate, `hidx` is filled.
void build_axis(int n_lr, int n_hr, HostAxis& out, int shift_sign = 0);
// area_k / tot_area after the reference's 0<->3, 1<->2 swap (:1078-1084): w[k*Q + q] multiplies pass k's prediction.
void ensemble_weights_host(int H, int W, int HH, int WW, float* w /* [4, HH*WW] */);

// ---------------------------------------------------------------------------------------------
// Launch plumbing shared by the kernel translation units.
struct LaunchCtx {
  cudaStream_t stream;
  int64_t* launch_counter;   // incremented per kernel launch
  int num_sms;
};

// Workspace carve-up for one decode call (all offsets 256-byte aligned).
struct Workspace {
  void* tab;      // [H*W,256]  fp32 (FP32 mode) or fp16 (BF16 mode)
  void* qtab;     // [HH*WW,128] same element type as tab
  float* flow;    // [HH*WW,4]
  float* act_a;   // FP32 mode only: ping-pong activation buffers [chunk,256]
  float* act_b;
  float* act_c;   // [chunk,64]
  float* ftab;    // local-ensemble mode: F = 30 Wl0[:, :64] HRfeat for the whole slab [HH*WW,64]
  float* pred;    // local-ensemble mode: one pass's prediction [3,HH*WW]
  int* flag;      // device int: row-band halo violation flag
  void* utab;     // STIF_FLAG_TEST_VARIANT: [4H*4W,192] UB | UE1 | UE2 = frame columns applied to the x4-upsampled frames
                  // (fp32 in FP32 mode, fp16 in BF16 mode)
  void* uq;       // STIF_FLAG_TEST_VARIANT, tensor-core path away from x4: [HH*WW,192] fp16 (UB at the query positions | zeros)
  void* uadd;     // ... and [HH*WW,64] fp16 upsampled-frame terms of stage D for the slab being decoded
  float* rgb32;   // STIF_FLAG_OUT_U8: fp32 staging of one slab [3,HH*WW] ahead of the uint8 conversion
  size_t chunk;   // queries per activation chunk (FP32 mode)
  size_t total_bytes;
};
Workspace carve_workspace(void* base, int H, int W, int HH, int WW, int mode);

// fp32 FMA-pipe path (kernels_fp32.cu)
// decoding_test variant (Sakuya_arch_test.py:513-514): utab[4H*4W,192] = w_up . bilinear_upsample_x4(frames)
cudaError_t project_frames_up4(const LaunchCtx& cx, const DeviceWeights32& w, const float* frames6, int H, int W, float* utab);
// scratch (optional): [scratch_rows, 256] fp32 for the row-major copy of [latent; frames] the tensor-core projection reads
cudaError_t project_latent(const LaunchCtx& cx, const DeviceWeights32& w, const float* latent192, const float* frames6,
                           int H, int W, void* tab, bool tab_half, bool test_variant = false, float* scratch = nullptr,
                           size_t scratch_rows = 0);
cudaError_t decode_slab_fp32(const LaunchCtx& cx, const DeviceWeights32& w, const FoldedWeights& hw, const Geometry& geo,
                             const Workspace& ws, float t, int row_begin, int row_end, int k1_row_begin,
                             int k1_row_end, float* out_rgb /* [3,HH,WW] */, int stage /* 1 = K1 (A+B), 2 = K2 (C+D+E) */);
// decoding_localensemble (Sakuya_arch_test.py:962-1085): 4 shifted passes blended by swapped areas.  geo_pass[k] carries the
// shifted axis tables of pass k (loop order (vx,vy) = (-1,-1),(-1,1),(1,-1),(1,1)); ens_y/ens_x = tables for sign -1, +1.
struct TcWeights;
// pass k of decoding_localensemble: out (+)= pred * area_{3-k} / tot_area (bit-exact weights)
cudaError_t ensemble_blend_launch(const LaunchCtx& cx, const float* pred, float* out_rgb, int HH, int WW, const AxisTables ens_y[2],
                                  const AxisTables ens_x[2], int k);
cudaError_t decode_slab_tc_ensemble(const LaunchCtx& cx, const TcWeights* tw, const Geometry geo_pass[4], const AxisTables ens_y[2],
                                    const AxisTables ens_x[2], const Workspace& ws, float t, float* out_rgb);
// custom_video_test.py:102 output conversion: planar fp32 [3,HH*WW] -> uint8 HWC, rows [row_begin,row_end)
cudaError_t rgb_to_u8_hwc(const LaunchCtx& cx, const float* rgb_planar, uint8_t* out_hwc, int HH, int WW, int row_begin, int row_end);
cudaError_t decode_slab_fp32_ensemble(const LaunchCtx& cx, const DeviceWeights32& w, const FoldedWeights& hw,
                                      const Geometry geo_pass[4], const AxisTables ens_y[2], const AxisTables ens_x[2],
                                      const Workspace& ws, float t, float* out_rgb);

// bf16 tcgen05 path (kernels_tc.cu)
struct TcWeights;  // opaque: device smem images + host constant blocks
TcWeights* tc_weights_create(const FoldedWeights& hw, std::string& err);
void tc_weights_destroy(TcWeights*);
// Band plan of the band-major host pipeline (host_plan.cpp).
constexpr int kMaxHostGroup = 4;   // timesteps whose Q table + flow stay resident at once
struct HostBandPlan {
  std::vector<int> he, ge, lr_end;   // per band: end (HR rows) of stage A+B, of stage C-E; LR rows that must have landed
  double cost_us;                    // the cost model's estimate of the call
};
HostBandPlan plan_host_bands(int H, int W, int HH, int WW, int G, int bands_hint, bool bands_forced, int halo, int num_sms);

// out_u8 != null (stage 2 only): the slab's uint8 HWC frame is written by K2's output stage instead of fp32 planar RGB
cudaError_t decode_slab_tc(const LaunchCtx& cx, const TcWeights* tw, const Geometry& geo, const Workspace& ws, float t,
                           int row_begin, int row_end, int k1_row_begin, int k1_row_end, float* out_rgb, int stage,
                           uint8_t* out_u8 = nullptr, int col_begin = -1, int col_end = -1 /* stage 2: column window, default whole width */,
                           const void* uadd = nullptr /* stage 2: [HH*WW,64] fp16 term added to the first layer (decoding_test away from x4) */);
// decoding_test on the tensor-core kernels at sizes other than x4 (the upsampled-frame grid is then not the query grid):
//   uq[HH*WW,192]  = UB bilinearly sampled at every query position | zeros  (per frame pair; k1_stage_ab_upf_kernel adds it)
//   uadd[HH*WW,64] = bilinear(UE1; g1) + bilinear(UE2; g2) at the flow-warped positions of rows [row_begin,row_end) (per slab)
constexpr int kMaxSlabsHost = 4;   // == kMaxSlabs of kernels_tc.cu
// stage 1 (K1) or 2 (K2) of up to four timesteps in one launch (same rows, per-timestep tables ws[g] / outputs): band-major host pipeline
cudaError_t decode_multi_tc(const LaunchCtx& cx, const TcWeights* tw, const Geometry& geo, const Workspace* ws, const float* t, int nslab,
                            int row_begin, int row_end, int k1_row_begin, int k1_row_end, float* const* out_rgb, uint8_t* const* out_u8,
                            int stage);
cudaError_t resample_ub_tc(const LaunchCtx& cx, const void* utab, const Geometry& geo, void* uq);
cudaError_t warp_u_terms_tc(const LaunchCtx& cx, const void* utab, const float* flow, const Geometry& geo, int row_begin, int row_end, void* uadd);
cudaError_t project_frames_up4_tc(const LaunchCtx& cx, const TcWeights* tw, const float* frames6, int H, int W, void* utab);
// latent_is_bf16: `latent192` points at bf16 bit patterns [192,H,W] instead of fp32 (host entry stif_decode_host_bf16)
cudaError_t project_latent_tc(const LaunchCtx& cx, const TcWeights* tw, const float* latent192, const float* frames6, int H,
                              int W, void* tab /* fp16 [H*W,256] */, int row_begin, int row_end, bool test_variant = false,
                              bool latent_is_bf16 = false);
int tc_selftest(int device, std::string& report);
// C[M, N] (row stride ldc) = act(A[M, K] . W[n_off : n_off + N, :]^T + bias): three tcgen05 MMAs per K step on a 2-term bf16
// split of both operands, fp32 accumulation (<= ~1e-5 relative).  N a multiple of 64, K = the layer's (64 or 256).
struct FoldedWeights;
HpWeights* hp_weights_create(const FoldedWeights& hw, std::string& err);
void hp_weights_destroy(HpWeights*);
const HpLayer* hp_layer(const HpWeights* w, int id);
cudaError_t hp_gemm(const LaunchCtx& cx, const HpLayer& L, int n_off, int N, const float* A, const float* bias, float* C, long ldc,
                    long M, int act);
// ... with two destinations: columns [0, n_split) -> C, the rest -> C2
cudaError_t hp_gemm_split(const LaunchCtx& cx, const HpLayer& L, const float* A, const float* bias, long M, int act, int n_split, float* C,
                          long ldc, float* C2, long ldc2);
// ... with the 256-wide sine activations contracted with a proj_n x 256 output layer in the epilogue instead of being stored
cudaError_t hp_gemm_proj(const LaunchCtx& cx, const HpLayer& L, const float* A, const float* bias, long M, const float* proj_w,
                         const float* proj_b, int proj_n, float* out, long scm, long scn);
int hp_selftest(std::string& report);

}  // namespace stif
