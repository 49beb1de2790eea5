#!/usr/bin/env python
"""Run the reference's UNMODIFIED `codes/custom_video_test.py` with the B200 decoder swapped in (BASELINE.json config 5).

Needs, on the same machine: a B200, the reference checkout, `latest_G.pth` (or `--synthetic-weights`) and
`video_sequences/train/*.png` in the working directory (the script only processes a folder named `train`,
custom_video_test.py:66).  What this wrapper does, all outside the reference's files:

  1. installs a torchvision-backed `_ext` module so that the reference's DCNv2 encoder (`DCNv2/dcn_v2.py:11,24`) runs on
     torch >= 1.11 (the THC-era extension does not build): `dcn_v2_forward` -> `torchvision.ops.deform_conv2d`
     (same offset/mask layout, `DCNv2/src/cuda/dcn_v2_im2col_cuda.cu:162-189`);
  2. patches `Sakuya_arch_test.LunaTokis.decoding*` at CLASS level (the script builds the model itself, :35) so that
     `model(imgs, times)` (:52) runs the reference encoder and then libstif_b200's kernels;
  3. `runpy`-executes the script.

    python tools/run_custom_video_test.py --reference /path/to/STIF --mode bf16 [--synthetic-weights] [--time-decoder]
"""
from __future__ import annotations

import argparse
import os
import runpy
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stif-continuous-video-representation_b200"))
sys.path.insert(0, ROOT)


def make_ext_shim():
    import torch
    from torchvision.ops import deform_conv2d

    ext = types.ModuleType("_ext")

    def dcn_v2_forward(input, weight, bias, offset, mask, kh, kw, sh, sw, ph, pw, dh, dw, dg):
        return deform_conv2d(input, offset, weight, bias, stride=(sh, sw), padding=(ph, pw), dilation=(dh, dw), mask=mask)

    def unsupported(*a, **k):
        raise NotImplementedError("inference-only shim")

    ext.dcn_v2_forward = dcn_v2_forward
    ext.dcn_v2_backward = unsupported
    ext.dcn_v2_psroi_pooling_forward = unsupported
    ext.dcn_v2_psroi_pooling_backward = unsupported
    return ext


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="root of the reference checkout (contains codes/)")
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32", "reference"],
                    help="'reference' leaves the decoder untouched (for A/B comparisons)")
    ap.add_argument("--synthetic-weights", action="store_true", help="write a SIREN-init latest_G.pth if none exists")
    ap.add_argument("--time-decoder", action="store_true", help="print decoder seconds per frame pair")
    args = ap.parse_args()

    import torch

    sys.modules["_ext"] = make_ext_shim()
    codes = os.path.join(args.reference, "codes")
    sys.path.insert(0, codes)
    import models.modules.Sakuya_arch_test as sat  # noqa: E402

    if args.synthetic_weights and not os.path.exists("latest_G.pth"):
        torch.manual_seed(0)
        torch.save(sat.LunaTokis(64, 6, 8, 5, 40).state_dict(), "latest_G.pth")
    if args.mode != "reference":
        import stif_b200
        stif_b200.install_class_patch(sat.LunaTokis, mode=args.mode)
    if args.time_decoder:
        inner = sat.LunaTokis.decoding

        def timed(self, times=None, scale=None):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = inner(self, times, scale)
            torch.cuda.synchronize()
            print(f"[decoder] {time.perf_counter() - t0:.4f} s for {len(times)} timesteps")
            return out
        sat.LunaTokis.decoding = timed
    runpy.run_path(os.path.join(codes, "custom_video_test.py"), run_name="__main__")


if __name__ == "__main__":
    main()
