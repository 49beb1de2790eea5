"""numpy restatement of STIF's space-time query decoder -- THE PARITY ORACLE (test-only).

Follows, stage by stage, ``LunaTokis.decoding`` (``/root/reference/codes/models/modules/
Sakuya_arch_test.py:364-459``), ``make_coord`` (``:1233-1248``), ``Siren``/``SineLayer``
(``SIREN.py:44-45,76-79``), ``warpgrid`` (``warplayer.py:25-39``) and -- for the arithmetic
that lives in the un-vendored dependency ``torch==2.11.0`` -- ATen's ``grid_sampler``
(``ATen/native/cuda/GridSampler.cuh:23-31`` ``grid_sampler_unnormalize``; nearest uses
``nearbyint``, bilinear uses 4 zero-padded taps).  All arithmetic is fp32 with explicit,
separately rounded operations (numpy never contracts to FMA).

Parity status: PINNED against reference-generated fixtures (``tests/golden``, made by
``oracle/make_goldens.py``; checked by ``tests/test_oracle_golden.py``): nearest indices and
``rel`` bit-exact, every stage to <=2e-6.

One knowingly inexact item: the warp base grid.  ``torch.linspace`` (CPU: vector-lane
``arange`` from per-vector bases; CUDA: FMA-contracted two-sided formula) differs from the
correctly rounded ``-1 + 2 i/(n-1)`` used here in the last ulp for about half the entries.
It only feeds continuous bilinear taps (1 ulp ~ 1e-4 HR px at 4K): RGB moves by < 1e-6.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
OMEGA0 = F32(30.0)
CLAMP_LO = F32(-1 + 1e-6)   # python double -> fp32, as torch.clamp does (Sakuya_arch_test.py:373)
CLAMP_HI = F32(1 - 1e-6)


# ----------------------------------------------------------------------------- axes
def make_axis(n: int) -> np.ndarray:
    """``make_coord`` for one axis (``:1233-1248``): ``fl32(-1+1/n) + fl32(2/n) * fl32(j)`` with
    the product and the sum rounded separately (two ATen kernels, no FMA)."""
    r = 1.0 / n                      # (v1 - v0) / (2 n) in python double
    prod = (F32(2.0 * r) * np.arange(n, dtype=F32)).astype(F32)
    return (F32(-1.0 + r) + prod).astype(F32)


def clamp_axis(c: np.ndarray) -> np.ndarray:
    return np.clip(c, CLAMP_LO, CLAMP_HI).astype(F32)


def unnormalize(c: np.ndarray, n: int) -> np.ndarray:
    """``grid_sampler_unnormalize(align_corners=False)``: ``((c + 1) * n - 1) / 2`` in fp32."""
    c = np.asarray(c, dtype=F32)
    return (((c + F32(1.0)) * F32(n) - F32(1.0)) / F32(2.0)).astype(F32)


def nearest_index(c: np.ndarray, n: int) -> np.ndarray:
    """nearest-mode texel index: ``nearbyint`` (round-half-even) of the unnormalised coordinate."""
    return np.rint(unnormalize(c, n)).astype(np.int64)


def linspace_axis(n: int) -> np.ndarray:
    """warp base grid, ``torch.linspace(-1, 1, n)`` (``warplayer.py:28-31``); see module note."""
    if n == 1:
        return np.array([-1.0], dtype=F32)
    return np.linspace(-1.0, 1.0, n, dtype=np.float64).astype(F32)


def query_axis_tables(n_lr: int, n_hr: int) -> dict:
    """Everything the decoder needs along one axis (separable until the MLPs):
    ``c``   clamped HR query coordinate (``:373``)
    ``i``   nearest LR texel index (``:382-393``) -- must be bit-exact
    ``rel`` ``(c - lr_c[i]) * n_lr`` (``:394-396``)
    ``b0``/``bw`` floor index / fractional weight of the stage-B bilinear tap (``:410-417``)
    ``base`` linspace warp base (``warplayer.py:28-31``)."""
    c = clamp_axis(make_axis(n_hr))
    lr_c = make_axis(n_lr)           # feat_coord is NOT clamped (``:375-377``)
    i = nearest_index(c, n_lr)
    inb = (i >= 0) & (i < n_lr)
    q = np.where(inb, lr_c[np.clip(i, 0, n_lr - 1)], F32(0.0)).astype(F32)
    rel = ((c - q) * F32(n_lr)).astype(F32)
    u = unnormalize(c, n_lr)
    b0 = np.floor(u)
    bw = (u - b0).astype(F32)
    return {"c": c, "i": i, "rel": rel, "lr_c": lr_c, "b0": b0.astype(np.int64), "bw": bw,
            "base": linspace_axis(n_hr)}


# ----------------------------------------------------------------------------- gathers
def gather_nearest(img: np.ndarray, y: np.ndarray, x: np.ndarray) -> np.ndarray:
    """``F.grid_sample(mode='nearest', align_corners=False)`` at points (y, x); img ``[C,Hi,Wi]`` -> ``[Q,C]``."""
    C, Hi, Wi = img.shape
    iy = nearest_index(y, Hi)
    ix = nearest_index(x, Wi)
    ok = (iy >= 0) & (iy < Hi) & (ix >= 0) & (ix < Wi)
    v = img[:, np.clip(iy, 0, Hi - 1), np.clip(ix, 0, Wi - 1)]
    return (v * ok.astype(F32)).T.astype(F32)


def gather_bilinear(img: np.ndarray, y: np.ndarray, x: np.ndarray) -> np.ndarray:
    """``F.grid_sample(mode='bilinear', padding_mode='zeros', align_corners=False)``;
    img ``[C,Hi,Wi]`` -> ``[Q,C]``.  Taps outside the image contribute 0."""
    C, Hi, Wi = img.shape
    u = unnormalize(x, Wi)
    v = unnormalize(y, Hi)
    x0 = np.floor(u)
    y0 = np.floor(v)
    wx1 = (u - x0).astype(F32)
    wy1 = (v - y0).astype(F32)
    wx0 = (F32(1.0) - wx1).astype(F32)
    wy0 = (F32(1.0) - wy1).astype(F32)
    x0 = x0.astype(np.int64)
    y0 = y0.astype(np.int64)
    out = np.zeros((y.shape[0], C), dtype=F32)
    for dy, wy in ((0, wy0), (1, wy1)):
        for dx, wx in ((0, wx0), (1, wx1)):
            xi = x0 + dx
            yi = y0 + dy
            ok = (xi >= 0) & (xi < Wi) & (yi >= 0) & (yi < Hi)
            vals = img[:, np.clip(yi, 0, Hi - 1), np.clip(xi, 0, Wi - 1)].T
            w = ((wy * wx).astype(F32) * ok.astype(F32)).astype(F32)
            out += (vals * w[:, None]).astype(F32)
    return out


# ----------------------------------------------------------------------------- SIREN
def siren(x: np.ndarray, weights: dict, net: str) -> np.ndarray:
    """``Siren.forward`` (``SIREN.py:76-79``): sine layers ``sin(30 * (x W^T + b))`` (``:44-45``)
    followed by the outermost plain linear layer (``:63-69``)."""
    li = 0
    while f"{net}.net.{li}.linear.weight" in weights:
        w = weights[f"{net}.net.{li}.linear.weight"]
        b = weights[f"{net}.net.{li}.linear.bias"]
        x = np.sin(OMEGA0 * ((x @ w.T).astype(F32) + b).astype(F32)).astype(F32)
        li += 1
    w = weights[f"{net}.net.{li}.weight"]
    b = weights[f"{net}.net.{li}.bias"]
    return ((x @ w.T).astype(F32) + b).astype(F32)


# ----------------------------------------------------------------------------- decode
def _times_matrix(times, B: int) -> np.ndarray:
    t = np.asarray(times, dtype=F32)
    if t.ndim == 1:
        t = np.repeat(t[:, None], B, axis=1)
    assert t.ndim == 2 and t.shape[1] == B, "times must be [T] or [T,B]"
    return t


def upsample4_bilinear(img: np.ndarray) -> np.ndarray:
    """``F.upsample(x, scale_factor=4, mode='bilinear')`` (``Sakuya_arch_test.py:514``; align_corners=False) for ``[C,H,W]``:
    ATen ``area_pixel_compute_source_index``: ``src = max(0, 0.25*(dst+0.5) - 0.5)`` in fp32, taps ``i0 = floor(src)``,
    ``i1 = min(i0+1, n-1)``, weights ``(1-l, l)`` with ``l = src - i0``; rows first, then columns, fp32 throughout."""
    img = np.asarray(img, dtype=F32)
    C, H, W = img.shape

    def axis(n):
        dst = np.arange(4 * n, dtype=F32)
        src = np.maximum(F32(0.0), (F32(0.25) * (dst + F32(0.5)) - F32(0.5)).astype(F32)).astype(F32)
        i0 = np.floor(src).astype(np.int64)
        i1 = np.minimum(i0 + 1, n - 1)
        l1 = (src - i0.astype(F32)).astype(F32)
        return i0, i1, (F32(1.0) - l1).astype(F32), l1

    y0, y1, wy0, wy1 = axis(H)
    x0, x1, wx0, wx1 = axis(W)
    # ATen's upsample_bilinear2d: w_y0*(w_x0*a + w_x1*b) + w_y1*(w_x0*c + w_x1*d)
    top = (img[:, y0][:, :, x0] * wx0[None, None, :] + img[:, y0][:, :, x1] * wx1[None, None, :]).astype(F32)
    bot = (img[:, y1][:, :, x0] * wx0[None, None, :] + img[:, y1][:, :, x1] * wx1[None, None, :]).astype(F32)
    return (wy0[None, :, None] * top + wy1[None, :, None] * bot).astype(F32)


def decode(latent, frames, weights, times, scale=None, return_stages: bool = False,
           chunk: int = 1 << 16, upsampled_frames: bool = False, window=None):
    """Restatement of ``LunaTokis.decoding(times, scale)``; with ``upsampled_frames`` of ``decoding_test`` (``:461-598``):
    identical except that the frame pair is bilinearly upsampled x4 (``:513-514``) before every BILINEAR frame gather
    (stage B ``:520-523``, stage D ``:548-551, :562-565``); the nearest gather of stage A still reads the LR frames
    (``:486-489``), and ``scale`` is an integer factor there (``:467``).
    ``window = (x0, x1, y0, y1)`` (rows, columns) gives ``decoding_memory`` (``:600-861``): stage A on the whole raster,
    stages B-E of ``decoding_test`` on the window only, with ``warpgrid2`` (``warplayer.py:41-47``: the warp starts from the
    query's pixel-centre coordinate instead of the linspace base grid); the result is ``[T,B,3,x1-x0,y1-y0]``.

    latent ``[B,3,64,H,W]`` (``self.feat``), frames ``[B,2,3,H,W]`` (``self.inp``), ``times`` ``[T]`` or
    ``[T,B]``, ``scale`` = None (x4) or the OUTPUT SIZE ``(HH, WW)`` (``:368-371``).
    Returns rgb ``[T,B,3,HH,WW]`` fp32, unclamped; with ``return_stages`` also a dict of the
    stage tensors of the LAST (t, b) slab processed plus the axis tables."""
    latent = np.asarray(latent, dtype=F32)
    frames = np.asarray(frames, dtype=F32)
    B, _, _, H, W = latent.shape
    if upsampled_frames and scale is not None and np.ndim(scale) == 0:
        scale = (H * int(scale), W * int(scale))
    HH, WW = (4 * H, 4 * W) if scale is None else (int(scale[0]), int(scale[1]))
    T = len(times)
    tm = _times_matrix(times, B)
    ay = query_axis_tables(H, HH)
    ax = query_axis_tables(W, WW)
    Q = HH * WW
    jy, jx = np.divmod(np.arange(Q), WW)
    cy, cx = ay["c"][jy], ax["c"][jx]
    if window is not None:
        x0, x1, y0, y1 = (int(v) for v in window)
        sel = (np.arange(x0, x1)[:, None] * WW + np.arange(y0, y1)[None, :]).ravel()     # stages B-E see these queries only
        out_h, out_w = x1 - x0, y1 - y0
    else:
        sel, out_h, out_w = np.arange(Q), HH, WW
    QS = sel.size
    sjy, sjx, scy, scx = jy[sel], jx[sel], cy[sel], cx[sel]
    out = np.zeros((T, B, 3, out_h, out_w), dtype=F32)
    stages = {}
    for b in range(B):
        feat = latent[b].reshape(192, H, W)          # cat of self.feat[:,0..2] on channels (:365)
        fr = frames[b].reshape(6, H, W)              # self.inp.view(bs,-1,H,W)      (:387)
        fr_bil = upsample4_bilinear(fr) if upsampled_frames else fr   # what the bilinear frame gathers sample
        # t-independent part of stage A / B (recomputed per timestep by the reference)
        iy, ix = ay["i"][jy], ax["i"][jx]
        ok = ((iy >= 0) & (iy < H) & (ix >= 0) & (ix < W)).astype(F32)[:, None]
        iyc, ixc = np.clip(iy, 0, H - 1), np.clip(ix, 0, W - 1)
        for c in range(T):
            t = tm[c, b]
            hr = np.zeros((Q, 64), dtype=F32)
            a_in_keep = None
            for s in range(0, Q, chunk):
                e = min(Q, s + chunk)
                a_in = np.concatenate([
                    feat[:, iyc[s:e], ixc[s:e]].T * ok[s:e],
                    fr[:, iyc[s:e], ixc[s:e]].T * ok[s:e],
                    ay["rel"][jy[s:e]][:, None], ax["rel"][jx[s:e]][:, None],
                    np.full((e - s, 1), t, dtype=F32)], axis=1).astype(F32)      # 201 (:399)
                hr[s:e] = siren(a_in, weights, "feat_imnet")                      # (:400)
                if s == 0:
                    a_in_keep = a_in
            hr_map = np.ascontiguousarray(hr.T).reshape(64, HH, WW)               # (:401)
            flow = np.zeros((QS, 4), dtype=F32)
            b_in_keep = None
            for s in range(0, QS, chunk):
                e = min(QS, s + chunk)
                b_in = np.concatenate([
                    gather_nearest(hr_map, scy[s:e], scx[s:e]),                   # identity gather (:406-409)
                    gather_bilinear(feat, scy[s:e], scx[s:e]),                    # (:414-417)
                    gather_bilinear(fr_bil, scy[s:e], scx[s:e]),                  # (:410-413)
                    np.full((e - s, 1), t, dtype=F32)], axis=1).astype(F32)      # 263 (:418)
                flow[s:e] = siren(b_in, weights, "flow_imnet")                    # (:419)
                if s == 0:
                    b_in_keep = b_in
            # stage C: warpgrid (warplayer.py:25-39); flow ch0/1 = (dx,dy) to frame 0, ch2/3 to frame 1
            # (decoding_memory: warpgrid2 starts from the query coordinate itself, warplayer.py:41-47)
            bx_, by_ = (scx, scy) if window is not None else (ax["base"][sjx], ay["base"][sjy])
            gx1 = (bx_ + flow[:, 0] / F32((WW - 1.0) / 2.0)).astype(F32)
            gy1 = (by_ + flow[:, 1] / F32((HH - 1.0) / 2.0)).astype(F32)
            gx2 = (bx_ + flow[:, 2] / F32((WW - 1.0) / 2.0)).astype(F32)
            gy2 = (by_ + flow[:, 3] / F32((HH - 1.0) / 2.0)).astype(F32)
            gx1, gy1, gx2, gy2 = (clamp_axis(g) for g in (gx1, gy1, gx2, gy2))   # (:428,441)
            rgb = np.zeros((QS, 3), dtype=F32)
            c_in_keep = None
            for s in range(0, QS, chunk):
                e = min(QS, s + chunk)
                y1, x1, y2, x2 = gy1[s:e], gx1[s:e], gy2[s:e], gx2[s:e]
                c_in = np.concatenate([
                    gather_bilinear(hr_map, y1, x1), gather_bilinear(hr_map, y2, x2),   # (:429-432,442-445)
                    gather_bilinear(feat, y1, x1), gather_bilinear(feat, y2, x2),       # (:437-440,450-453)
                    gather_bilinear(fr_bil, y1, x1), gather_bilinear(fr_bil, y2, x2),   # (:433-436,446-449)
                    np.full((e - s, 1), t, dtype=F32)], axis=1).astype(F32)            # 525 (:455)
                rgb[s:e] = siren(c_in, weights, "encode_imnet")                         # (:456)
                if s == 0:
                    c_in_keep = c_in
            out[c, b] = rgb.T.reshape(3, out_h, out_w)                                   # (:457)
            if return_stages:
                stages = {"feat_in": a_in_keep, "hr": hr, "flow_in": b_in_keep, "flow": flow,
                          "grid1": np.stack([gy1, gx1], 1), "grid2": np.stack([gy2, gx2], 1),
                          "enc_in": c_in_keep, "rgb": rgb, "iy": ay["i"], "ix": ax["i"],
                          "rely": ay["rel"], "relx": ax["rel"], "cy": ay["c"], "cx": ax["c"]}
    return (out, stages) if return_stages else out


# ----------------------------------------------------------------------------- local ensemble
ENSEMBLE_SHIFTS = [(-1, -1), (-1, 1), (1, -1), (1, 1)]      # (vx, vy) loop order, Sakuya_arch_test.py:978-988
EPS_SHIFT = 1e-6


def shifted_axis_tables(n_lr: int, n_hr: int, v: int) -> dict:
    """Per-axis tables of ``decoding_localensemble`` for shift sign ``v`` (``:981-995``): the query coordinate is
    moved by ``v * (1/n_lr) + 1e-6`` (python double -> fp32 at the in-place add) and re-clamped; EVERY gather uses
    the shifted coordinate, but ``rel`` is taken from the un-shifted one (``:1008``)."""
    c = clamp_axis(make_axis(n_hr))
    cs = clamp_axis((c + F32(v * (2.0 / n_lr / 2.0) + EPS_SHIFT)).astype(F32))
    lr_c = make_axis(n_lr)
    i = nearest_index(cs, n_lr)
    inb = (i >= 0) & (i < n_lr)
    q = np.where(inb, lr_c[np.clip(i, 0, n_lr - 1)], F32(0.0)).astype(F32)
    rel = ((c - q) * F32(n_lr)).astype(F32)
    return {"c": c, "cs": cs, "i": i, "rel": rel, "hi": nearest_index(cs, n_hr), "base": linspace_axis(n_hr)}


def ensemble_weights(H: int, W: int, HH: int, WW: int) -> np.ndarray:
    """``area_k / tot_area`` AFTER the 0<->3, 1<->2 swap (``:1079-1084``), shape ``[4, HH*WW]`` -- the weight that
    multiplies pass k's prediction.  fp32, separately rounded: area = abs(rel_y * rel_x) + fl32(1e-9)."""
    jy, jx = np.divmod(np.arange(HH * WW), WW)
    areas = []
    for vx, vy in ENSEMBLE_SHIFTS:
        ry = shifted_axis_tables(H, HH, vx)["rel"][jy]
        rx = shifted_axis_tables(W, WW, vy)["rel"][jx]
        areas.append((np.abs((ry * rx).astype(F32)) + F32(1e-9)).astype(F32))
    tot = (((areas[0] + areas[1]).astype(F32) + areas[2]).astype(F32) + areas[3]).astype(F32)
    return np.stack([(areas[3 - k] / tot).astype(F32) for k in range(4)], 0)


def decode_localensemble(latent, frames, weights, times, scale=None):
    """Restatement of ``LunaTokis.decoding_localensemble`` (``:962-1085``; batch size 1 as in the reference).
    Returns ``[T,3,HH,WW]``."""
    latent = np.asarray(latent, dtype=F32)
    frames = np.asarray(frames, dtype=F32)
    assert latent.shape[0] == 1
    _, _, _, H, W = latent.shape
    HH, WW = (4 * H, 4 * W) if scale is None else (int(scale[0]), int(scale[1]))
    feat, fr = latent[0].reshape(192, H, W), frames[0].reshape(6, H, W)
    Q = HH * WW
    jy, jx = np.divmod(np.arange(Q), WW)
    wts = ensemble_weights(H, W, HH, WW)
    out = np.zeros((len(times), 3, HH, WW), dtype=F32)
    for c, t in enumerate(times):
        t = F32(t)
        ret = np.zeros((Q, 3), dtype=F32)
        for k, (vx, vy) in enumerate(ENSEMBLE_SHIFTS):
            ay, ax = shifted_axis_tables(H, HH, vx), shifted_axis_tables(W, WW, vy)
            iy, ix = ay["i"][jy], ax["i"][jx]
            ok = ((iy >= 0) & (iy < H) & (ix >= 0) & (ix < W)).astype(F32)[:, None]
            iyc, ixc = np.clip(iy, 0, H - 1), np.clip(ix, 0, W - 1)
            a_in = np.concatenate([feat[:, iyc, ixc].T * ok, fr[:, iyc, ixc].T * ok, ay["rel"][jy][:, None],
                                   ax["rel"][jx][:, None], np.full((Q, 1), t, dtype=F32)], 1).astype(F32)
            hr = siren(a_in, weights, "feat_imnet")
            hr_map = np.ascontiguousarray(hr.T).reshape(64, HH, WW)
            ys, xs = ay["cs"][jy], ax["cs"][jx]
            b_in = np.concatenate([gather_nearest(hr_map, ys, xs), gather_bilinear(feat, ys, xs),
                                   gather_bilinear(fr, ys, xs), np.full((Q, 1), t, dtype=F32)], 1).astype(F32)
            flow = siren(b_in, weights, "flow_imnet")
            g = []
            for f0 in (0, 2):
                gx = clamp_axis((ax["base"][jx] + flow[:, f0] / F32((WW - 1.0) / 2.0)).astype(F32))
                gy = clamp_axis((ay["base"][jy] + flow[:, f0 + 1] / F32((HH - 1.0) / 2.0)).astype(F32))
                g.append((gy, gx))
            (y1, x1), (y2, x2) = g
            c_in = np.concatenate([gather_bilinear(hr_map, y1, x1), gather_bilinear(hr_map, y2, x2),
                                   gather_bilinear(feat, y1, x1), gather_bilinear(feat, y2, x2),
                                   gather_bilinear(fr, y1, x1), gather_bilinear(fr, y2, x2),
                                   np.full((Q, 1), t, dtype=F32)], 1).astype(F32)
            pred = siren(c_in, weights, "encode_imnet")
            ret = (ret + (pred * wts[k][:, None]).astype(F32)).astype(F32)
        out[c] = ret.T.reshape(3, HH, WW)
    return out


def psnr255(a: np.ndarray, b: np.ndarray) -> float:
    """``utils/util.py:140-151`` ``calculate_psnr`` on clamp(0,1)*255 images."""
    a = np.clip(a, 0, 1).astype(np.float64) * 255.0
    b = np.clip(b, 0, 1).astype(np.float64) * 255.0
    mse = float(np.mean((a - b) ** 2))
    if mse == 0:
        return float("inf")
    return 20.0 * float(np.log10(255.0 / np.sqrt(mse)))
