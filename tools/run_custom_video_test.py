#!/usr/bin/env python
"""Run the reference's UNMODIFIED `codes/custom_video_test.py` with the B200 decoder swapped in (BASELINE.json config 5).

Needs, on the same machine: a B200, the reference tree (`oracle/_ref` staged by `oracle/stage_ref.py`, or a checkout),
`latest_G.pth` (or `--synthetic-weights`) and `video_sequences/train/*.png` in the working directory (the script only
processes a folder named `train`, custom_video_test.py:66; `--make-video N` writes a synthetic one).  What this wrapper
does, all outside the reference's files:

  1. installs a torchvision-backed `_ext` module so that the reference's DCNv2 encoder (`DCNv2/dcn_v2.py:11,24`) runs on
     torch >= 1.11 (the THC-era extension does not build): `dcn_v2_forward` -> `torchvision.ops.deform_conv2d`
     (same offset/mask layout, `DCNv2/src/cuda/dcn_v2_im2col_cuda.cu:162-189`);
  2. patches `Sakuya_arch_test.LunaTokis.decoding*` at CLASS level (the script builds the model itself, :35) so that
     `model(imgs, times)` (:52) runs the reference encoder and then libstif_b200's kernels;
  3. `runpy`-executes the script from the current directory.

    python tools/run_custom_video_test.py [--reference DIR] --mode bf16|fp32|reference [--synthetic-weights]
                                          [--make-video N] [--report out.json] [--dcn torchvision|b200]

`--mode reference` leaves the decoder untouched (the unpatched A/B arm).  `--report` writes encoder / decoder seconds
per frame pair (CUDA-synchronised around `gen_feat` and `decoding`) and the script's wall time.
"""
from __future__ import annotations

import argparse
import json
import os
import runpy
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stif-continuous-video-representation_b200"))
sys.path.insert(0, ROOT)


def default_reference() -> str:
    staged = os.path.join(ROOT, "oracle", "_ref")
    return staged if os.path.isdir(os.path.join(staged, "codes")) else "/root/reference"


def make_ext_shim(dcn: str = "torchvision"):
    """`_ext` for the reference's encoder.  dcn='torchvision': `torchvision.ops.deform_conv2d`; dcn='b200': this repo's sm_100a
    kernel (`stif_dcn_v2_forward`, the encoder's 64->64 3x3 dg=8 geometry; anything else still goes to torchvision)."""
    from torchvision.ops import deform_conv2d

    ext = types.ModuleType("_ext")
    ext.calls = {"b200": 0, "torchvision": 0}

    def dcn_v2_forward(input, weight, bias, offset, mask, kh, kw, sh, sw, ph, pw, dh, dw, dg):
        if dcn == "b200":
            import stif_b200
            out = stif_b200.dcn_v2_forward(input, weight, bias, offset, mask, kh, kw, sh, sw, ph, pw, dh, dw, dg)
            if out is not None:
                ext.calls["b200"] += 1
                return out
        ext.calls["torchvision"] += 1
        return deform_conv2d(input, offset, weight, bias, stride=(sh, sw), padding=(ph, pw), dilation=(dh, dw), mask=mask)

    def unsupported(*a, **k):
        raise NotImplementedError("inference-only shim")

    ext.dcn_v2_forward = dcn_v2_forward
    ext.dcn_v2_backward = unsupported
    ext.dcn_v2_psroi_pooling_forward = unsupported
    ext.dcn_v2_psroi_pooling_backward = unsupported
    return ext


def make_synthetic_video(n_frames: int, width: int = 960, height: int = 540, folder: str = "video_sequences/train") -> None:
    """`im01.png .. imNN.png`: a smooth colour field translating at constant velocity with two moving discs on top --
    consistent motion between consecutive frames, BGR uint8 as `cv2.imread` expects (custom_video_test.py:85)."""
    import cv2
    import numpy as np

    os.makedirs(folder, exist_ok=True)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    for i in range(n_frames):
        ph = 3.0 * i
        r = 0.5 + 0.5 * np.sin((xx + ph) / 37.0) * np.cos((yy - 0.5 * ph) / 53.0)
        g = 0.5 + 0.5 * np.sin((xx - 2.0 * ph) / 91.0 + (yy + ph) / 29.0)
        b = 0.5 + 0.5 * np.cos((xx + yy + ph) / 67.0)
        img = np.stack([b, g, r], -1)
        for k, (cx0, cy0, vx, vy, rad) in enumerate(((200.0, 150.0, 5.0, 2.0, 60.0), (700.0, 400.0, -4.0, -1.5, 90.0))):
            d2 = (xx - (cx0 + vx * i)) ** 2 + (yy - (cy0 + vy * i)) ** 2
            img[d2 < rad * rad] = (0.9, 0.2, 0.1) if k == 0 else (0.1, 0.8, 0.9)
        cv2.imwrite(os.path.join(folder, f"im{i + 1:02d}.png"), (np.clip(img, 0, 1) * 255).astype(np.uint8))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=default_reference(), help="root of the reference tree (contains codes/)")
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32", "reference"],
                    help="'reference' leaves the decoder untouched (for A/B comparisons)")
    ap.add_argument("--synthetic-weights", action="store_true", help="write a SIREN-init latest_G.pth if none exists")
    ap.add_argument("--make-video", type=int, default=0, metavar="N", help="write N synthetic 960x540 frames if video_sequences/train is missing")
    ap.add_argument("--time-decoder", action="store_true", help="print decoder seconds per frame pair")
    ap.add_argument("--report", default=None, help="write encoder/decoder seconds per pair + wall time as JSON")
    ap.add_argument("--dcn", default="torchvision", choices=["torchvision", "b200"],
                    help="what serves the encoder's `_ext.dcn_v2_forward`: torchvision.ops.deform_conv2d or this repo's sm_100a kernel")
    args = ap.parse_args()

    import torch

    sys.modules["_ext"] = make_ext_shim(args.dcn)
    codes = os.path.join(os.path.abspath(args.reference), "codes")
    if not os.path.isfile(os.path.join(codes, "custom_video_test.py")):
        raise SystemExit(f"no reference under {args.reference} (run `python -m oracle.stage_ref` in the build container)")
    sys.path.insert(0, codes)
    import models.modules.Sakuya_arch_test as sat  # noqa: E402

    if args.make_video and not os.path.isdir("video_sequences/train"):
        make_synthetic_video(args.make_video)
    if args.synthetic_weights and not os.path.exists("latest_G.pth"):
        torch.manual_seed(0)
        torch.save(sat.LunaTokis(64, 6, 8, 5, 40).state_dict(), "latest_G.pth")
    if args.mode != "reference":
        import stif_b200
        stif_b200.install_class_patch(sat.LunaTokis, mode=args.mode)
    report = {"mode": args.mode, "decoder_s": [], "encoder_s": [], "timesteps": [], "out_shape": None,
              "reference": os.path.abspath(args.reference)}
    if args.time_decoder or args.report:
        inner_dec, inner_enc = sat.LunaTokis.decoding, sat.LunaTokis.gen_feat

        def timed_dec(self, times=None, scale=None):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = inner_dec(self, times, scale)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            report["decoder_s"].append(dt)
            report["timesteps"].append(len(times))
            report["out_shape"] = list(out[0].shape)
            if args.time_decoder:
                print(f"[decoder] {dt:.4f} s for {len(times)} timesteps")
            return out

        def timed_enc(self, x):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = inner_enc(self, x)
            torch.cuda.synchronize()
            report["encoder_s"].append(time.perf_counter() - t0)
            return out
        sat.LunaTokis.decoding = timed_dec
        sat.LunaTokis.gen_feat = timed_enc
    t0 = time.perf_counter()
    runpy.run_path(os.path.join(codes, "custom_video_test.py"), run_name="__main__")
    report["wall_s"] = time.perf_counter() - t0
    report["pairs"] = len(report["decoder_s"])
    report["dcn"] = args.dcn
    report["dcn_calls"] = dict(sys.modules["_ext"].calls)
    if args.mode != "reference":
        import stif_b200
        report["native_lib"] = stif_b200.LIB_PATH
    if args.report:
        with open(args.report, "w") as f:
            json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
