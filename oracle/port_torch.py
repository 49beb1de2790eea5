"""The reference decoder's algorithm on torch CPU ops -- the TIMED CPU BASELINE (test-only).

``bench.py``'s ``cpu_baseline`` leg and ``bench.py --impl reference`` execute this on the GPU box's
host cores (``/root/reference`` is Python and cannot travel to the box, so the "reference arm"
is this port; ``kind: "port"``).  It issues the same ATen work per timestep as
``LunaTokis.decoding`` (``Sakuya_arch_test.py:364-459``): 4 nearest + 8 zero-padded bilinear
``F.grid_sample`` calls, three wide ``torch.cat``s, 13 fp32 ``addmm`` + 10 ``sin``, using every host
thread torch has -- including the per-timestep recomputation of the t-independent gathers.
The one thing it omits is the reference's discarded border-padded warp inside ``warpgrid``
(``warplayer.py:39``), whose result the reference throws away.

Pinned against the reference-generated fixtures by ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

_EPS = 1e-6


def _axis(n: int) -> torch.Tensor:
    r = 1.0 / n
    return (-1.0 + r) + (2.0 * r) * torch.arange(n).float()          # make_coord (:1233-1248)


def _grid(hh: int, ww: int) -> torch.Tensor:
    yy, xx = torch.meshgrid(_axis(hh), _axis(ww), indexing="ij")
    return torch.stack([yy, xx], -1).view(-1, 2)                        # [Q,2] (y,x)


def _sample(img: torch.Tensor, yx: torch.Tensor, mode: str) -> torch.Tensor:
    """img [B,C,h,w], yx [B,Q,2] (y,x) -> [B,Q,C]"""
    g = yx.flip(-1).unsqueeze(1)
    return F.grid_sample(img, g, mode=mode, align_corners=False)[:, :, 0, :].permute(0, 2, 1)


def _siren(x: torch.Tensor, w: dict, net: str) -> torch.Tensor:
    li = 0
    while f"{net}.net.{li}.linear.weight" in w:
        x = torch.sin(30.0 * F.linear(x, w[f"{net}.net.{li}.linear.weight"], w[f"{net}.net.{li}.linear.bias"]))
        li += 1
    return F.linear(x, w[f"{net}.net.{li}.weight"], w[f"{net}.net.{li}.bias"])


@torch.no_grad()
def decode(latent, frames, weights, times, scale=None, device: str = "cpu") -> torch.Tensor:
    """latent [B,3,64,H,W], frames [B,2,3,H,W], times [T] or [T,B] -> rgb [T,B,3,HH,WW] (torch fp32).  ``device`` = "cpu"
    for the timed CPU baseline; "cuda" runs the SAME eager op sequence on the GPU (``profiles/eager_gpu_baseline.py``:
    the reference's own execution model on the same box)."""
    latent = torch.as_tensor(latent, dtype=torch.float32).to(device)
    frames = torch.as_tensor(frames, dtype=torch.float32).to(device)
    w = {k: torch.as_tensor(v, dtype=torch.float32).to(device) for k, v in weights.items()}
    B, _, _, H, W = latent.shape
    feat = latent.reshape(B, 192, H, W)
    inp6 = frames.reshape(B, 6, H, W)
    HH, WW = (4 * H, 4 * W) if scale is None else (int(scale[0]), int(scale[1]))
    Q = HH * WW
    tm = np.asarray(times, dtype=np.float32)
    if tm.ndim == 1:
        tm = np.repeat(tm[:, None], B, 1)
    coord = _grid(HH, WW).unsqueeze(0).repeat(B, 1, 1).clamp(-1 + _EPS, 1 - _EPS).to(device)      # (:373)
    lr = _grid(H, W).view(H, W, 2).permute(2, 0, 1).unsqueeze(0).expand(B, 2, H, W).to(device)    # (:375-377)
    bx = torch.linspace(-1.0, 1.0, WW).view(1, 1, WW).expand(B, HH, WW).to(device)                # warplayer.py:28-31
    by = torch.linspace(-1.0, 1.0, HH).view(1, HH, 1).expand(B, HH, WW).to(device)
    outs = []
    for c in range(tm.shape[0]):
        t = torch.from_numpy(tm[c]).view(B, 1, 1).to(device)
        pe = torch.ones(B, Q, 1, device=device) * t
        # stage A (:382-401)
        q_feat = _sample(feat, coord, "nearest")
        q_inp = _sample(inp6, coord, "nearest")
        q_coord = _sample(lr, coord, "nearest")
        rel = coord - q_coord
        rel = torch.stack([rel[..., 0] * H, rel[..., 1] * W], -1)
        hr = _siren(torch.cat([q_feat, q_inp, rel, pe], -1).view(B * Q, -1), w, "feat_imnet").view(B, Q, 64)
        hr_map = hr.permute(0, 2, 1).reshape(B, 64, HH, WW)
        # stage B (:406-422)
        b_in = torch.cat([_sample(hr_map, coord, "nearest"), _sample(feat, coord, "bilinear"),
                          _sample(inp6, coord, "bilinear"), pe], -1)
        flow = _siren(b_in.view(B * Q, -1), w, "flow_imnet").view(B, Q, 4).permute(0, 2, 1).reshape(B, 4, HH, WW)
        # stage C (warplayer.py:25-39, :428, :441)
        grids = []
        for k in (0, 2):
            gx = bx + flow[:, k] / ((WW - 1.0) / 2.0)
            gy = by + flow[:, k + 1] / ((HH - 1.0) / 2.0)
            grids.append(torch.stack([gy, gx], -1).view(B, Q, 2).clamp(-1 + _EPS, 1 - _EPS))
        g1, g2 = grids
        # stage D + E (:429-458)
        e_in = torch.cat([_sample(hr_map, g1, "bilinear"), _sample(hr_map, g2, "bilinear"),
                          _sample(feat, g1, "bilinear"), _sample(feat, g2, "bilinear"),
                          _sample(inp6, g1, "bilinear"), _sample(inp6, g2, "bilinear"), pe], -1)
        rgb = _siren(e_in.view(B * Q, -1), w, "encode_imnet").view(B, Q, 3)
        outs.append(rgb.permute(0, 2, 1).reshape(B, 3, HH, WW))
    return torch.stack(outs, 0)
