/* stif_b200.h -- C ABI of the B200-native STIF space-time query decoder.
 *
 * This is the drop-in boundary for ONE path of paperwave/STIF-continuous-video-representation:
 * `LunaTokis.decoding(times, scale)` and its siblings
 * (reference: codes/models/modules/Sakuya_arch_test.py:364-459, :863-960, :962-1085), i.e. the
 * three-SIREN continuous space-time decoder that turns the encoder's latent volume
 * (`self.feat` [B,3,64,H,W], set at :361) and the LR frame pair (`self.inp` [B,2,3,H,W], :1224)
 * into RGB at an arbitrary output raster (HH,WW) and arbitrary times t.
 *
 * The reference has no FFI for this path: the boundary there is a Python
 * method on the model.  The only FFI precedent in the reference is the DCNv2 extension
 * (codes/models/modules/DCNv2/src/vision.cpp:4-9, src/dcn_v2.h:9-39): free functions taking device
 * buffers, asserting CUDA residency, erroring on unsupported devices, running on the
 * caller's current stream.  This header follows that precedent with plain C types:
 * pointers + sizes in, int status out, no torch types.  INTEGRATION.md shows the ctypes
 * binding and the class-level patch a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative STIF_E* code on failure;
 *     stif_last_error() returns a thread-local human-readable message for the last failure.
 *   - "dev" pointers are CUDA device pointers on the decoder's device; "host" pointers are
 *     ordinary host memory.  There is NO CPU fallback: without a usable sm_100 device
 *     stif_create fails with STIF_ENODEV.
 *   - all work is stream-ordered on the cudaStream_t passed as `void* stream`
 *     (NULL = legacy default stream), except the one-off table upload on first sight of a geometry
 *     (see stif_prepare).  A handle is not re-entrant.
 *   - the caller owns every buffer, including the workspace; the library owns only its
 *     packed copy of the weights and a few KB of per-geometry axis tables.
 */
#ifndef STIF_B200_H_
#define STIF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STIF_ABI_VERSION 2

/* status codes */
#define STIF_OK        0
#define STIF_EINVAL   -1   /* bad argument (shape, null pointer, unknown mode ...)          */
#define STIF_ENODEV   -2   /* no CUDA device / not an sm_100 device / kernels not loadable  */
#define STIF_ECUDA    -3   /* a CUDA runtime call or kernel launch failed                   */
#define STIF_ENOMEM   -4   /* workspace too small (see stif_workspace_bytes)                */
#define STIF_ESTATE   -5   /* weights not loaded yet                                        */

/* precision / algorithm modes (the `mode` argument) */
#define STIF_MODE_BF16      0  /* tcgen05 bf16 tensor-core kernels, fp32 accumulate; RGB within 2e-2 of the reference */
#define STIF_MODE_FP32      1  /* fp32 FMA-pipe kernels; RGB within 1e-4 of the reference                           */
/* flags OR-ed into `mode` */
#define STIF_FLAG_LOCAL_ENSEMBLE 0x100  /* decoding_localensemble semantics (Sakuya_arch_test.py:962-1085), either precision mode */
#define STIF_FLAG_TEST_VARIANT   0x400  /* decoding_test semantics (Sakuya_arch_test.py:461-598, what VideoSRBaseModel.test runs):
                                         * the frame pair is bilinearly upsampled x4 (:513-514) before every bilinear frame
                                         * gather.  Both precision modes, any size (the tensor-core kernels have a fast path at the
                                         * method's own x4, where the upsampled-frame grid is the query grid). */
#define STIF_FLAG_WARP_FROM_COORD 0x800 /* warpgrid2 semantics (warplayer.py:41-47): the warp starts from the query's own
                                         * pixel-centre coordinate instead of the linspace base grid.  Together with
                                         * STIF_FLAG_TEST_VARIANT and stif_decode_rows this is decoding_memory
                                         * (Sakuya_arch_test.py:600-861) without its file-system side effects (stif_decode_window). */
#define STIF_FLAG_OUT_U8         0x200  /* write what the reference's caller makes of the result (custom_video_test.py:102):
                                         * `(img.clamp(0,1).permute(1,2,0) * 255).astype(uint8)` -- uint8 [T,B,HH,WW,3],
                                         * fp32 clamp / multiply, truncation.  The `out` pointer is then a uint8_t buffer and
                                         * device<->host output traffic drops 4x. */

/* number of weight tensors the decoder consumes (state-dict order, see stif_load_weights) */
#define STIF_NUM_WEIGHT_TENSORS 26

typedef struct stif_decoder stif_decoder_t;

/* ABI version of the loaded library (compare with STIF_ABI_VERSION). */
int stif_abi_version(void);

/* Message for the most recent failure on the calling thread ("" if none). */
const char* stif_last_error(void);

/* Create a decoder bound to CUDA device `device`.
 * Replaces: model construction + `.to('cuda')` for the decoder part
 * (codes/custom_video_test.py:35-39). */
int stif_create(stif_decoder_t** out, int device);

int stif_destroy(stif_decoder_t* dec);

/* Load the 26 decoder tensors, fp32, HOST pointers, in this order (shapes [out,in] / [out]):
 *   feat_imnet.net.{0,1,2}.linear.{weight,bias}, feat_imnet.net.3.{weight,bias},
 *   flow_imnet.net.{0,1,2}.linear.{weight,bias}, flow_imnet.net.3.{weight,bias},
 *   encode_imnet.net.{0,1,2,3}.linear.{weight,bias}, encode_imnet.net.4.{weight,bias}
 * i.e. [64,201],[64] [64,64],[64] [256,64],[256] [64,256],[64] | [64,263].. [4,256],[4] |
 * [64,525].. [256,256],[256] [3,256],[3].
 * Replaces: `model.load_state_dict(torch.load('latest_G.pth'), strict=True)` for the decoder
 * sub-modules (codes/custom_video_test.py:36; layer shapes Sakuya_arch_test.py:306-311).
 * The library folds omega_0 = 30 (SIREN.py:45) into its packed copy and synchronises the
 * device before returning, so the host buffers may be freed immediately. */
int stif_load_weights(stif_decoder_t* dec, const float* const* tensors_host, int num_tensors);

/* Build and cache the per-geometry axis tables of an [H,W] -> [HH,WW] decode (plus the shifted tables of
 * STIF_FLAG_LOCAL_ENSEMBLE / the coordinate warp base of STIF_FLAG_WARP_FROM_COORD if `mode` carries the flag) ahead of
 * time.  Optional: the first stif_decode of a geometry does this itself, but that first call then contains a cudaMalloc
 * and a synchronous upload (a device-wide sync; illegal inside a stream capture) before its stream-ordered work.  After
 * stif_prepare -- or after any earlier decode of the same geometry -- stif_decode / stif_decode_rows are purely
 * stream-ordered.  The cache holds 64 geometries; the oldest one is evicted (the reference's own warp-grid cache,
 * warplayer.py:6,26-33, grows without bound). */
int stif_prepare(stif_decoder_t* dec, int H, int W, int HH, int WW, int mode);

/* Bytes of device workspace stif_decode needs for this problem (0 on invalid arguments). */
size_t stif_workspace_bytes(int B, int H, int W, int HH, int WW, int T, int mode);

/* Decode.  Replaces `LunaTokis.decoding(times, scale)` (Sakuya_arch_test.py:364-459).
 *   latent_dev  [B,3,64,H,W] fp32  (== self.feat; channel c of the 192 = (c/64, c%64))
 *   frames_dev  [B,2,3,H,W]  fp32  (== self.inp)
 *   times_host  [T,B] fp32         (times[c][b]; the reference passes [1,1] or [B,1] tensors per c)
 *   (HH,WW)     output raster size; the reference's `scale` argument IS this size
 *               (Sakuya_arch_test.py:368-371); the x4 default is (4H,4W)
 *   out_rgb_dev [T,B,3,HH,WW] fp32, unclamped (== torch.stack(preds));
 *               with STIF_FLAG_OUT_U8: uint8 [T,B,HH,WW,3] (see the flag)
 * With STIF_FLAG_LOCAL_ENSEMBLE the result is
 * decoding_localensemble's: B must be 1 as in the reference, out is [T,1,3,HH,WW]. */
int stif_decode(stif_decoder_t* dec,
                const float* latent_dev, const float* frames_dev,
                int B, int H, int W, int HH, int WW,
                const float* times_host, int T, int mode,
                void* workspace_dev, size_t workspace_bytes,
                void* out_rgb_dev, void* stream);

/* Same as stif_decode but the query raster is restricted to rows [row_begin,row_end) of every
 * (t,b) slab -- the unit the multi-GPU launcher shards when slabs < GPUs.  Stage A/B of the
 * reference are point-wise (Sakuya_arch_test.py:382-422), so the halo rows the warp of
 * stage D may reach (|flow_y| pixels, :424-453) are recomputed locally: rows
 * [row_begin-halo, row_end+halo) are decoded through stage A/B and the call FAILS with
 * STIF_EINVAL (message names the needed halo) if a flow reaches outside that band.
 * out_rgb_dev is still addressed as the full [T,B,3,HH,WW] tensor; only the band is written. */
int stif_decode_rows(stif_decoder_t* dec,
                     const float* latent_dev, const float* frames_dev,
                     int B, int H, int W, int HH, int WW,
                     const float* times_host, int T, int mode,
                     int row_begin, int row_end, int halo,
                     void* workspace_dev, size_t workspace_bytes,
                     void* out_rgb_dev, void* stream);

/* stif_decode_rows restricted further to columns [col_begin,col_end): the zoom window of decoding_memory
 * (Sakuya_arch_test.py:600-861: stage A on the whole raster, stages B-E on a 4H x 4W window around `center`; use halo = HH,
 * STIF_FLAG_TEST_VARIANT | STIF_FLAG_WARP_FROM_COORD).  Only the window is guaranteed to be written; the tensor-core kernels
 * compute nothing else, the fp32 kernels decode the window's rows at full width. */
int stif_decode_window(stif_decoder_t* dec,
                       const float* latent_dev, const float* frames_dev,
                       int B, int H, int W, int HH, int WW,
                       const float* times_host, int T, int mode,
                       int row_begin, int row_end, int col_begin, int col_end, int halo,
                       void* workspace_dev, size_t workspace_bytes,
                       void* out_rgb_dev, void* stream);

/* End-to-end convenience for FFI callers that hold HOST buffers: allocates device memory
 * internally (cached on the handle), copies latent/frames host->device, decodes, copies RGB
 * device->host, synchronises.  This is the call bench.py's `e2e` number goes through. */
int stif_decode_host(stif_decoder_t* dec,
                     const float* latent_host, const float* frames_host,
                     int B, int H, int W, int HH, int WW,
                     const float* times_host, int T, int mode,
                     void* out_rgb_host);

/* stif_decode_host for callers that hold the latent in bf16 (uint16_t bit patterns, [B,3,64,H,W], round-to-nearest-even of
 * the encoder's fp32 output): half the host->device bytes, which are what bounds this entry point (PCIe).  The tensor-core
 * path rounds the latent to bf16 before its first MMA anyway, so the result is bit-identical to stif_decode_host on the
 * fp32 latent the bf16 values were rounded from.  Plain STIF_MODE_BF16 decodes only (STIF_FLAG_OUT_U8 may be added);
 * frames stay fp32 (3 % of the bytes). */
int stif_decode_host_bf16(stif_decoder_t* dec,
                          const uint16_t* latent_bf16_host, const float* frames_host,
                          int B, int H, int W, int HH, int WW,
                          const float* times_host, int T, int mode,
                          void* out_rgb_host);

/* ---- the step before the path: the encoder's modulated deformable convolution ----
 * Forward of DCNv2 behind the call signature of the reference's extension entry `_ext.dcn_v2_forward(input, weight, bias,
 * offset, mask, kh, kw, sh, sw, ph, pw, dh, dw, dg)` (codes/models/modules/DCNv2/dcn_v2.py:24-27 -> src/cuda/dcn_v2_cuda.cu:42-172,
 * im2col src/cuda/dcn_v2_im2col_cuda.cu:125-195).  All pointers are DEVICE pointers, fp32, contiguous NCHW:
 *   input [B,C,H,W], weight [Cout,C,kh,kw], bias [Cout], offset [B, dg*2*kh*kw, H, W] (channel g*2*kh*kw + 2k = dy, +1 = dx of
 *   tap k), mask [B, dg*kh*kw, H, W], out [B,Cout,H,W]; stream-ordered on `stream`.
 * Implemented: the one geometry the reference's encoder uses (C = Cout = 64, 3x3, stride 1, padding 1, dilation 1, dg = 8;
 * Sakuya_arch_test.py:38-66,135-160) as a fused deformable-im2col + tcgen05 kernel with a 2-term bf16 operand split
 * (fp32-class accuracy).  Any other geometry returns STIF_EINVAL without touching `out`: the caller keeps its fallback.
 * No decoder handle and no state of the caller's: the channels-last copy of the input the kernel gathers from is a stream-ordered
 * allocation from a per-device pool the library owns, so calls on different streams / threads are independent. */
int stif_dcn_v2_forward(const float* input_dev, const float* weight_dev, const float* bias_dev,
                        const float* offset_dev, const float* mask_dev,
                        int B, int C, int H, int W, int Cout, int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw,
                        int dg, float* out_dev, void* stream);

/* ---- introspection used by the parity tests (same device functions as the decode path) ---- */

/* Per-axis query tables for an (n_lr -> n_hr) axis, HOST outputs of length n_hr (any may be NULL):
 *   coord  clamped pixel-centre query coordinate   (make_coord :1233-1248 + clamp :373)
 *   index  nearest LR texel (grid_sample nearest, :382-393)            -- bit-exact contract
 *   rel    (coord - lr_coord[index]) * n_lr         (:394-396)         -- bit-exact contract
 *   base   warp base grid, linspace(-1,1,n_hr)      (warplayer.py:28-31) */
int stif_axis_tables(int n_lr, int n_hr, float* coord, int32_t* index, float* rel, float* base);

/* Blend weights of decoding_localensemble (Sakuya_arch_test.py:1011-1012, 1077-1084), HOST output
 * [4, HH*WW]: weights[k*HH*WW + q] = area_{3-k}[q] / tot_area[q] multiplies pass k's prediction
 * (pass order (vx,vy) = (-1,-1), (-1,1), (1,-1), (1,1)).  Bit-exact contract; the decode kernels
 * recompute the same value with IEEE-rounded intrinsics. */
int stif_ensemble_weights(int H, int W, int HH, int WW, float* weights_host, size_t num_floats);

/* Copy stage intermediates of the LAST slab decoded by `dec` to HOST buffers (any may be NULL):
 *   flow [HH*WW,4] fp32 -- flow_imnet output (dx1,dy1,dx2,dy2), HR-pixel units (:419-422).
 * Returns STIF_ESTATE if no decode has run. */
int stif_debug_last_flow(stif_decoder_t* dec, float* flow_host, size_t num_floats);

/* Tuning / introspection of stif_decode_host's band-major pipeline (bf16 mode): the latent is uploaded in `bands` LR row
 * bands (default 6) and stage C-E of a band trails stage A-B by `halo` HR rows (default 32, doubled by the library after a
 * call in which a warp reached further; that call repeats stage C-E on the complete tables, so results never depend on
 * these knobs).  Values <= 0 leave a knob unchanged; *respins (may be NULL) receives how many times a repeat was needed. */
int stif_debug_host_pipeline(stif_decoder_t* dec, int bands, int halo, int64_t* respins);

/* The band plan stif_decode_host would use (pure host arithmetic, no device needed): for a [H,W] -> [HH,WW] decode of T
 * timesteps with the given band-count hint (`bands_forced` != 0 takes it literally), halo and SM count, writes per band
 * the LR rows that must have been uploaded, the end of stage A+B and the end of stage C-E in HR rows (arrays of
 * `max_entries`), the cost model's estimate in microseconds (may be NULL) and returns the number of bands (< 0: error). */
int stif_debug_band_plan(int H, int W, int HH, int WW, int T, int bands, int bands_forced, int halo, int num_sms,
                         int max_entries, int* lr_end, int* ab_end, int* ce_end, double* cost_us);

/* Number of kernel launches issued by this handle since creation (bench.py's gpu_launches). */
int64_t stif_launch_count(const stif_decoder_t* dec);

/* Per-kernel device timing.  While enabled, stif_decode brackets each kernel group with CUDA
 * events on the caller's stream: group 0 = latent projection (K0), 1 = stage A+B (K1),
 * 2 = stage C+D+E (K2).  stif_profile_read synchronises the device, adds the elapsed
 * milliseconds and group launch counts accumulated since the last read into ms[3] / count[3]
 * (overwriting them) and resets the accumulators.  bench.py's roofline uses this. */
int stif_profile_enable(stif_decoder_t* dec, int enable);
int stif_profile_read(stif_decoder_t* dec, double* ms, int64_t* count);

/* Run the built-in tcgen05/TMEM unit checks on the device (small GEMMs against a host fp32
 * reference).  Writes a report into `report` (NUL-terminated, truncated to cap).  Returns 0 if
 * every check passed. */
int stif_selftest(int device, char* report, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* STIF_B200_H_ */
