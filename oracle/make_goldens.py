"""Generate ``tests/golden/*.npz`` by EXECUTING THE REFERENCE (build container only).

    python -m oracle.make_goldens            # from the repo root; needs /root/reference

The reference ships no golden vectors, so these fixtures are the pin for
``oracle/restate_np.py`` and ``oracle/port_torch.py`` and, through them, for the CUDA path.
Inputs and weights are NOT stored: they are regenerated from seeds by ``oracle/synth.py``
(numpy PCG64, platform-stable); a checksum of each is stored to detect drift.

What is captured, per case, from the unmodified ``LunaTokis.decoding``
(``Sakuya_arch_test.py:364-459``) via forward hooks on its three SIREN sub-modules:
  * the full 201/263/525-wide MLP inputs at a strided subset of queries -- these contain the
    nearest-gathered latents, ``rel_coord``, the bilinear gathers and the warped gathers,
    i.e. every intermediate of stages A-D;
  * HRfeat / flow at the same subset, the full RGB output;
and separately the per-axis nearest-index tables of ``F.grid_sample(mode='nearest')`` and the
``make_coord`` axes for every (n_lr, n_hr) pair the configs use.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import synth  # noqa: E402
from oracle.ref_loader import build_reference_model, clear_warp_cache, load_reference_module  # noqa: E402

ROWS = 200  # about this many sampled queries per case in the stage dumps

CASES = {
    # name: dict(B,H,W,scale,times,wseed,stress,iseed,latent_std)
    "x4_init": dict(B=1, H=16, W=16, scale=None, times=[0.0, 0.375], wseed=0, stress=False, iseed=0, latent_std=0.05),
    "x6p5_stress": dict(B=1, H=16, W=16, scale=(104, 104), times=[1 / 9, 8 / 9], wseed=1, stress=True, iseed=1,
                        latent_std=0.05),
    "odd_b2_stress": dict(B=2, H=12, W=10, scale=(37, 53), times=[[0.25, 0.75]], wseed=2, stress=True, iseed=2,
                          latent_std=1.0),
    "down_stress": dict(B=1, H=20, W=24, scale=(13, 17), times=[0.5], wseed=3, stress=True, iseed=3, latent_std=0.3),
}

AXIS_PAIRS = [(64, 256), (64, 416), (270, 1080), (270, 1755), (480, 1920), (480, 3120), (540, 2160),
              (960, 3840), (272, 1088), (10, 37), (12, 53), (16, 64), (16, 104), (20, 13), (24, 17), (7, 7)]


def checksum(*arrays) -> float:
    return float(sum(np.abs(np.asarray(a, dtype=np.float64)).sum() for a in arrays))


def run_case(name: str, cfg: dict) -> dict:
    import torch

    weights = synth.make_weights(cfg["wseed"], cfg["stress"])
    latent, frames = synth.make_inputs(cfg["iseed"], cfg["B"], cfg["H"], cfg["W"], cfg["latent_std"])
    model = build_reference_model(weights)
    clear_warp_cache()
    model.feat = torch.from_numpy(latent)
    model.inp = torch.from_numpy(frames)
    B = cfg["B"]
    tm = np.asarray(cfg["times"], dtype=np.float32)
    if tm.ndim == 1:
        times = [torch.tensor([[float(t)]], dtype=torch.float32) for t in tm]       # [1,1] (custom_video_test.py:50)
    else:
        times = [torch.tensor(row, dtype=torch.float32).view(B, 1) for row in tm]   # [B,1] (VideoSR_base_model.py:93)
    cap = {"feat_imnet": [], "flow_imnet": [], "encode_imnet": []}
    hooks = []
    for net in cap:
        def hook(mod, inp, out, net=net):
            cap[net].append((inp[0].detach().numpy().copy(), out.detach().numpy().copy()))
        hooks.append(getattr(model, net).register_forward_hook(hook))
    with torch.no_grad():
        preds = model.decoding(times, cfg["scale"])
    for h in hooks:
        h.remove()
    rgb = np.stack([p.numpy() for p in preds], 0)                                   # [T,B,3,HH,WW]
    T = len(times)
    HH, WW = rgb.shape[-2:]
    Q = HH * WW
    sel = np.arange(0, B * Q, max(1, (B * Q // ROWS)) | 1)
    out = {"rgb": rgb.astype(np.float32), "sel": sel.astype(np.int64),
           "input_checksum": np.float64(checksum(latent, frames)),
           "weight_checksum": np.float64(checksum(*weights.values()))}
    for c in range(T):
        out[f"feat_in_{c}"] = cap["feat_imnet"][c][0][sel]
        out[f"hr_{c}"] = cap["feat_imnet"][c][1][sel]
        out[f"flow_in_{c}"] = cap["flow_imnet"][c][0][sel]
        out[f"flow_{c}"] = cap["flow_imnet"][c][1]                                  # full [B*Q,4]
        out[f"enc_in_{c}"] = cap["encode_imnet"][c][0][sel]
    # sibling methods (same building blocks) where applicable
    if B == 1 and tm.ndim == 1:
        tl = [float(t) for t in tm]
        with torch.no_grad():
            clear_warp_cache()
            fast = model.decoding_fasttest(tl, cfg["scale"])
            clear_warp_cache()
            # the blend weights are locals of the reference method: capture them at its return
            # (post-swap `areas` and `tot_area`, Sakuya_arch_test.py:1078-1084)
            grabbed = {}

            def tracer(frame, event, arg):
                if frame.f_code.co_name != "decoding_localensemble":
                    return None

                def local(frame, event, arg):
                    if event == "return":
                        grabbed["areas"] = [a.numpy().copy() for a in frame.f_locals["areas"]]
                        grabbed["tot"] = frame.f_locals["tot_area"].numpy().copy()
                    return local
                return local
            sys.settrace(tracer)
            try:
                ens = model.decoding_localensemble(tl, cfg["scale"])
            finally:
                sys.settrace(None)
            out["ensemble_weights"] = np.stack([(a / grabbed["tot"])[0] for a in grabbed["areas"]], 0).astype(np.float32)
        out["rgb_fasttest"] = fast.numpy().astype(np.float32)
        out["rgb_localensemble"] = ens.numpy().astype(np.float32)
    return out


TEST_VARIANT_CASES = {
    # decoding_test (Sakuya_arch_test.py:461-598): frames x4-upsampled before the bilinear taps; int scale only
    "testvar_x4_init": dict(H=16, W=16, scale=None, times=[0.0, 0.375], wseed=0, stress=False, iseed=0, latent_std=0.05),
    "testvar_x3_stress": dict(H=12, W=10, scale=3, times=[0.25, 0.8], wseed=2, stress=True, iseed=7, latent_std=0.3),
    "testvar_x5_stress": dict(H=9, W=14, scale=5, times=[0.6], wseed=3, stress=True, iseed=8, latent_std=1.0),
}


def run_test_variant(cfg: dict) -> dict:
    """`LunaTokis.decoding_test(times, scale)` on seeded inputs: full RGB + the flow of every timestep."""
    import torch

    weights = synth.make_weights(cfg["wseed"], cfg["stress"])
    latent, frames = synth.make_inputs(cfg["iseed"], 1, cfg["H"], cfg["W"], cfg["latent_std"])
    model = build_reference_model(weights)
    clear_warp_cache()
    model.feat = torch.from_numpy(latent)
    model.inp = torch.from_numpy(frames)
    times = [torch.tensor([[float(t)]], dtype=torch.float32) for t in cfg["times"]]
    flows = []
    hook = model.flow_imnet.register_forward_hook(lambda mod, inp, out: flows.append(out.detach().numpy().copy()))
    with torch.no_grad():
        preds = model.decoding_test(times, cfg["scale"])
    hook.remove()
    Q = preds[0].shape[-2] * preds[0].shape[-1]
    # the method runs flow_imnet on three query chunks per timestep (:519-527): stitch them back
    per_t = [np.concatenate(flows[3 * c:3 * c + 3], 0) for c in range(len(times))]
    assert all(f.shape == (Q, 4) for f in per_t)
    return {"rgb": np.stack([p.numpy() for p in preds], 0).astype(np.float32), "flow": np.stack(per_t, 0).astype(np.float32),
            "input_checksum": np.float64(checksum(latent, frames)), "weight_checksum": np.float64(checksum(*weights.values()))}


MEMORY_VARIANT_CASES = {
    # decoding_memory (Sakuya_arch_test.py:600-861): stage A on the full (HH,WW) raster, stages B-E (decoding_test's upsampled
    # frames, warpgrid2) on a 4H x 4W window around `center`; the window is clamped into the raster
    "memvar_mid": dict(H=12, W=10, scale=(100, 90), center=(0.1, -0.2), times=[0.3, 0.9], wseed=1, stress=True, iseed=3, latent_std=0.3),
    "memvar_corner": dict(H=8, W=9, scale=(40, 50), center=(-0.9, 0.8), times=[0.5], wseed=2, stress=True, iseed=4, latent_std=1.0),
}


def window_of(H, W, HH, WW, center):
    """Rows [x0,x1) and columns [y0,y1) of the window, exactly as the method computes them (:636-650)."""
    H0, W0 = 4 * H, 4 * W
    cc = ((np.asarray(center, dtype=np.float64) + 1) / 2) * np.array((HH, WW))
    x0, x1, y0, y1 = int(cc[0]) - H0 // 2, int(cc[0]) + H0 - H0 // 2, int(cc[1]) - W0 // 2, int(cc[1]) + W0 - W0 // 2
    if x0 < 0:
        x1 -= x0
        x0 = 0
    elif x1 > HH:
        x0 -= (x1 - HH)
        x1 = HH
    if y0 < 0:
        y1 -= y0
        y0 = 0
    elif y1 > WW:
        y0 -= (y1 - WW)
        y1 = WW
    return x0, x1, y0, y1


def run_memory_variant(cfg: dict) -> dict:
    """`LunaTokis.decoding_memory(times, scale, center, input_img)` with its file-system side effects (hard-coded
    `/home/users/...` directories, JPEG saves, :609-651) neutralised for the duration of the call -- nothing outside the
    numerical path is touched."""
    import os as _os
    import torch
    from unittest import mock
    from PIL import Image

    weights = synth.make_weights(cfg["wseed"], cfg["stress"])
    latent, frames = synth.make_inputs(cfg["iseed"], 1, cfg["H"], cfg["W"], cfg["latent_std"])
    model = build_reference_model(weights)
    clear_warp_cache()
    model.feat = torch.from_numpy(latent)
    model.inp = torch.from_numpy(frames)
    times = [torch.tensor([[float(t)]], dtype=torch.float32) for t in cfg["times"]]
    with mock.patch.object(_os, "makedirs", lambda *a, **k: None), mock.patch.object(_os.path, "exists", lambda p: True), \
            mock.patch.object(Image.Image, "save", lambda self, *a, **k: None), torch.no_grad():
        preds = model.decoding_memory(times, cfg["scale"], np.asarray(cfg["center"], dtype=np.float64), torch.from_numpy(frames))
    x0, x1, y0, y1 = window_of(cfg["H"], cfg["W"], cfg["scale"][0], cfg["scale"][1], cfg["center"])
    return {"rgb": np.stack([p.numpy() for p in preds], 0).astype(np.float32), "window": np.asarray([x0, x1, y0, y1], dtype=np.int64),
            "input_checksum": np.float64(checksum(latent, frames)), "weight_checksum": np.float64(checksum(*weights.values()))}


def run_config1(stress: bool) -> dict:
    """Config 1 of BASELINE.json (64x64 latent -> 256x256, 8 timesteps): store a strided sample."""
    import torch

    weights = synth.make_weights(0, stress)
    latent, frames = synth.make_inputs(0, 1, 64, 64, 0.05)
    model = build_reference_model(weights)
    clear_warp_cache()
    model.feat = torch.from_numpy(latent)
    model.inp = torch.from_numpy(frames)
    times = [torch.tensor([[i / 8.0]], dtype=torch.float32) for i in range(8)]
    with torch.no_grad():
        preds = model.decoding(times, None)
    rgb = np.stack([p.numpy() for p in preds], 0)                                   # [8,1,3,256,256]
    return {"rgb_sub": rgb[:, :, :, 1::5, 2::5].astype(np.float32),
            "mean": rgb.mean(axis=(1, 2, 3, 4)).astype(np.float64),
            "absmax": np.abs(rgb).max(axis=(1, 2, 3, 4)).astype(np.float64),
            "input_checksum": np.float64(checksum(latent, frames)),
            "weight_checksum": np.float64(checksum(*weights.values()))}


def run_e2e_small() -> dict:
    """The caller's path of BASELINE.json config 5 in miniature: `custom_video_test.single_forward` (:41-54) pads the
    frame pair to a multiple of 4, builds `time_Tensors = [tensor([i/8])[None] for i in range(8)]` and calls
    `model(imgs, times)` = `gen_feat` (reference encoder, here through a torchvision-backed `_ext`) + `decoding`.
    Stored: the encoder's latent (what the drop-in decoder receives), the padded frames and the reference's RGB."""
    import importlib.util
    import torch

    spec = importlib.util.spec_from_file_location("rcvt", os.path.join(ROOT, "tools", "run_custom_video_test.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    sys.modules["_ext"] = tool.make_ext_shim()
    weights = synth.make_weights(4, True)
    model = build_reference_model(weights)                      # encoder: torch.manual_seed(0) random init
    clear_warp_cache()
    rng = np.random.default_rng(77)
    h, w = 26, 34                                               # not multiples of 4 -> padded to 28 x 36
    frames = rng.random((1, 2, 3, h, w), dtype=np.float32)
    imgs = torch.from_numpy(frames)
    H, W = (h + 3) // 4 * 4, (w + 3) // 4 * 4
    padded = torch.zeros(1, 2, 3, H, W)
    padded[:, :, :, :h, :w] = imgs                              # custom_video_test.py:44-48
    times = [torch.tensor([i / 8.0])[None] for i in range(8)]   # :50
    with torch.no_grad():
        out = model(padded, times)                              # LunaTokis.forward -> gen_feat + decoding (:1222-1231)
    rgb = np.stack([o.numpy() for o in out], 0)                 # [8,1,3,4H,4W]
    return {"latent": model.feat.numpy().astype(np.float32), "frames": padded.numpy().astype(np.float32),
            "rgb": rgb.astype(np.float32), "weight_checksum": np.float64(checksum(*weights.values()))}


def axis_goldens() -> dict:
    """Per-axis nearest indices straight from ``F.grid_sample(mode='nearest')`` on an index ramp,
    the ``make_coord`` axes and ``torch.linspace`` for every (n_lr, n_hr) pair."""
    import torch
    import torch.nn.functional as F

    sat = load_reference_module()
    out = {"pairs": np.asarray(AXIS_PAIRS, dtype=np.int64)}
    for n_lr, n_hr in AXIS_PAIRS:
        c = sat.make_coord((n_hr, 1)).clamp(-1 + 1e-6, 1 - 1e-6)                    # [n_hr,2] (y, x=0)
        ramp = torch.arange(n_lr, dtype=torch.float32).view(1, 1, n_lr, 1)           # value == row index
        g = c.flip(-1).view(1, 1, n_hr, 2)
        idx = F.grid_sample(ramp, g, mode="nearest", align_corners=False).view(-1)
        lr_c = sat.make_coord((n_lr, 1), flatten=False)[:, 0, 0]
        out[f"idx_{n_lr}_{n_hr}"] = idx.numpy().astype(np.int32)
        out[f"coord_{n_hr}"] = c[:, 0].numpy().astype(np.float32)
        out[f"coord_{n_lr}"] = sat.make_coord((n_lr, 1)).clamp(-1 + 1e-6, 1 - 1e-6)[:, 0].numpy().astype(np.float32)
        out[f"lrcoord_{n_lr}"] = lr_c.numpy().astype(np.float32)
        out[f"linspace_{n_hr}"] = torch.linspace(-1.0, 1.0, n_hr).numpy().astype(np.float32)
    return out


def main():
    os.makedirs(GOLD, exist_ok=True)
    for name, cfg in CASES.items():
        res = run_case(name, cfg)
        np.savez_compressed(os.path.join(GOLD, f"case_{name}.npz"), **res)
        print(name, {k: getattr(v, "shape", None) for k, v in res.items() if k.startswith("rgb")})
    for stress in (False, True):
        res = run_config1(stress)
        np.savez_compressed(os.path.join(GOLD, f"config1_{'stress' if stress else 'init'}.npz"), **res)
        print("config1", stress, res["absmax"])
    for name, cfg in TEST_VARIANT_CASES.items():
        res = run_test_variant(cfg)
        np.savez_compressed(os.path.join(GOLD, f"case_{name}.npz"), **res)
        print(name, res["rgb"].shape)
    for name, cfg in MEMORY_VARIANT_CASES.items():
        res = run_memory_variant(cfg)
        np.savez_compressed(os.path.join(GOLD, f"case_{name}.npz"), **res)
        print(name, res["rgb"].shape, res["window"])
    np.savez_compressed(os.path.join(GOLD, "e2e_small.npz"), **run_e2e_small())
    np.savez_compressed(os.path.join(GOLD, "axis_tables.npz"), **axis_goldens())
    print("done")


if __name__ == "__main__":
    main()
