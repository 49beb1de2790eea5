"""Host-side mirror of the reference's decode interface, calling the CUDA kernels through the C ABI.

Reference surface being mirrored (``codes/models/modules/Sakuya_arch_test.py``):

* ``LunaTokis.decoding(times, scale)`` (``:364-459``) -> ``STIFQueryDecoder.decode`` /
  the bound method installed by ``patch_reference_model``: reads ``model.feat`` ``[B,3,64,H,W]``
  and ``model.inp`` ``[B,2,3,H,W]``, ``times`` = list of ``[1,1]`` / ``[B,1]`` tensors (or floats),
  ``scale`` = ``None`` (x4) or the OUTPUT SIZE ``(HH, WW)``; returns a python list of ``T`` tensors
  ``[B,3,HH,WW]`` fp32, unclamped, on the input device.
* ``LunaTokis.decoding_fasttest(times, scale)`` (``:863-960``) -> same numbers, one ``[T,3,HH,WW]`` tensor.
* the north-star ``forward(feat, coord, cell)`` surface does not exist in the reference
  (SURVEY.md section 0); it is provided here as an adapter over full query rasters.

PyTorch is plumbing only (device memory, streams).  There is no eager fallback: every decode
goes through ``libstif_b200.so``.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import (STIF_FLAG_LOCAL_ENSEMBLE, STIF_FLAG_OUT_U8, STIF_FLAG_TEST_VARIANT, STIF_FLAG_WARP_FROM_COORD,
                   STIF_MODE_BF16, STIF_MODE_FP32, StifError, check, lib)

_MODES = {"bf16": STIF_MODE_BF16, "fp32": STIF_MODE_FP32}

NET_SHAPES = {
    "feat_imnet": [201, 64, 64, 256, 64],
    "flow_imnet": [263, 64, 64, 256, 4],
    "encode_imnet": [525, 64, 64, 256, 256, 3],
}


def weight_keys() -> list[str]:
    """The 26 state-dict keys in the C ABI's order (``stif_load_weights``)."""
    keys = []
    for net, dims in NET_SHAPES.items():
        n = len(dims) - 1
        for li in range(n):
            stem = f"{net}.net.{li}" if li == n - 1 else f"{net}.net.{li}.linear"
            keys += [f"{stem}.weight", f"{stem}.bias"]
    return keys


def _expected_shape(key: str) -> tuple[int, ...]:
    net, _, li = key.split(".")[:3]
    dims = NET_SHAPES[net]
    li = int(li)
    return (dims[li + 1], dims[li]) if key.endswith("weight") else (dims[li + 1],)


def _times_matrix(times, B: int) -> np.ndarray:
    """list of [1,1]/[B,1] tensors (``decoding``) or floats (``decoding_fasttest``) -> float32 [T,B]."""
    rows = []
    for t in times:
        if isinstance(t, torch.Tensor):
            v = t.detach().to("cpu", torch.float32).reshape(-1).numpy()
        else:
            v = np.asarray(t, dtype=np.float32).reshape(-1)
        if v.size == 1:
            v = np.repeat(v, B)
        if v.size != B:
            raise ValueError(f"times entry has {v.size} values for batch size {B}")
        rows.append(v.astype(np.float32))
    if not rows:
        raise ValueError("times is empty")
    return np.ascontiguousarray(np.stack(rows, 0))


class STIFQueryDecoder(torch.nn.Module):
    """B200 decoder for STIF's continuous space-time queries.

    ``mode='bf16'``: fused tcgen05 tensor-core kernels (RGB within 2e-2 of the reference);
    ``mode='fp32'``: fp32 FMA-pipe kernels (RGB within 1e-4)."""

    def __init__(self, device: int | str | torch.device | None = None, mode: str = "bf16"):
        super().__init__()
        if not torch.cuda.is_available():
            raise StifError("STIFQueryDecoder needs a CUDA device (sm_100a); there is no CPU fallback")
        if mode not in _MODES:
            raise ValueError(f"mode must be one of {sorted(_MODES)}")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise StifError("STIFQueryDecoder only runs on CUDA devices")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self.mode = mode
        self._handle = C.c_void_p()
        check(lib.stif_create(C.byref(self._handle), self.device.index))
        self._workspace: torch.Tensor | None = None
        self._loaded = False

    # ------------------------------------------------------------------ weights
    def load_weights(self, state_dict) -> "STIFQueryDecoder":
        """Ingest the 26 decoder tensors from a (possibly full-model, possibly ``module.``-prefixed)
        state dict -- what ``load_state_dict(torch.load('latest_G.pth'))`` feeds the reference
        (``custom_video_test.py:36``)."""
        sd = {}
        for k, v in state_dict.items():
            k = k[len("module."):] if k.startswith("module.") else k
            sd[k] = v
        arrays = []
        for key in weight_keys():
            if key not in sd:
                raise KeyError(f"state dict is missing decoder tensor '{key}'")
            v = sd[key]
            a = v.detach().to("cpu", torch.float32).contiguous().numpy() if isinstance(v, torch.Tensor) \
                else np.ascontiguousarray(v, dtype=np.float32)
            if tuple(a.shape) != _expected_shape(key):
                raise ValueError(f"'{key}' has shape {tuple(a.shape)}, expected {_expected_shape(key)}")
            arrays.append(a)
        ptrs = (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
        check(lib.stif_load_weights(self._handle, ptrs, len(arrays)))
        self._loaded = True
        return self

    # ------------------------------------------------------------------ decode
    def _workspace_for(self, B, H, W, HH, WW, T, mode) -> torch.Tensor:
        need = int(lib.stif_workspace_bytes(B, H, W, HH, WW, T, mode))
        if need == 0:
            raise ValueError("invalid decode shape")
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = None
            self._workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._workspace

    def _prep(self, latent: torch.Tensor, frames: torch.Tensor, scale):
        if latent.dim() == 4:                                   # [B,192,H,W] accepted as well
            latent = latent.reshape(latent.shape[0], 3, 64, *latent.shape[-2:])
        if latent.dim() != 5 or latent.shape[1] != 3 or latent.shape[2] != 64:
            raise ValueError(f"latent must be [B,3,64,H,W] or [B,192,H,W], got {tuple(latent.shape)}")
        B, _, _, H, W = latent.shape
        if frames.dim() == 4:
            frames = frames.reshape(B, 2, 3, H, W)
        if tuple(frames.shape) != (B, 2, 3, H, W):
            raise ValueError(f"frames must be [B,2,3,H,W] = {(B, 2, 3, H, W)}, got {tuple(frames.shape)}")
        if latent.device.type != "cuda" or frames.device.type != "cuda":
            raise StifError("latent/frames must be CUDA tensors (no CPU fallback)")
        latent = latent.to(self.device, torch.float32).contiguous()
        frames = frames.to(self.device, torch.float32).contiguous()
        if scale is None:
            HH, WW = 4 * H, 4 * W                                # Sakuya_arch_test.py:368-369
        else:
            HH, WW = int(scale[0]), int(scale[1])                # `scale` IS the output size (:370-371)
        return latent, frames, B, H, W, HH, WW

    def decode_stacked(self, latent, frames, times, scale=None, mode: str | None = None,
                       rows: tuple[int, int] | None = None, halo: int = 0,
                       out: torch.Tensor | None = None, local_ensemble: bool = False, uint8: bool = False,
                       test_variant: bool = False, warp_from_coord: bool = False,
                       cols: tuple[int, int] | None = None) -> torch.Tensor:
        """Decode to one ``[T,B,3,HH,WW]`` fp32 tensor.  ``rows=(r0,r1)`` restricts the call to a row band
        (``stif_decode_rows``), used by the sharding launcher.  ``uint8=True`` returns what the reference's caller
        saves (``custom_video_test.py:102``): ``(clamp(0,1) * 255).astype(uint8)`` as ``[T,B,HH,WW,3]``."""
        if not self._loaded:
            raise StifError("load_weights() has not been called")
        latent, frames, B, H, W, HH, WW = self._prep(latent, frames, scale)
        tm = _times_matrix(times, B)
        T = tm.shape[0]
        m = (_MODES[mode or self.mode] | (STIF_FLAG_LOCAL_ENSEMBLE if local_ensemble else 0) | (STIF_FLAG_OUT_U8 if uint8 else 0)
             | (STIF_FLAG_TEST_VARIANT if test_variant else 0) | (STIF_FLAG_WARP_FROM_COORD if warp_from_coord else 0))
        ws = self._workspace_for(B, H, W, HH, WW, T, m)
        shape, dtype = ((T, B, HH, WW, 3), torch.uint8) if uint8 else ((T, B, 3, HH, WW), torch.float32)
        if out is None:
            # a row-band call writes only its band: the other rows are defined (zero), not allocator garbage
            out = (torch.zeros if rows is not None else torch.empty)(shape, dtype=dtype, device=self.device)
        elif tuple(out.shape) != shape or out.dtype != dtype or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous {dtype} {list(shape)} tensor")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        fp = tm.ctypes.data_as(C.POINTER(C.c_float))
        with torch.cuda.device(self.device):
            if rows is not None and cols is not None:
                check(lib.stif_decode_window(self._handle, latent.data_ptr(), frames.data_ptr(), B, H, W, HH, WW, fp, T, m,
                                             int(rows[0]), int(rows[1]), int(cols[0]), int(cols[1]), int(halo), ws.data_ptr(),
                                             ws.numel(), out.data_ptr(), stream))
            elif rows is None:
                check(lib.stif_decode(self._handle, latent.data_ptr(), frames.data_ptr(), B, H, W, HH, WW, fp, T, m,
                                      ws.data_ptr(), ws.numel(), out.data_ptr(), stream))
            else:
                check(lib.stif_decode_rows(self._handle, latent.data_ptr(), frames.data_ptr(), B, H, W, HH, WW, fp, T, m,
                                           int(rows[0]), int(rows[1]), int(halo), ws.data_ptr(), ws.numel(),
                                           out.data_ptr(), stream))
        return out

    def decode(self, latent, frames, times, scale=None, mode: str | None = None) -> list[torch.Tensor]:
        """``LunaTokis.decoding`` return convention: list of ``T`` tensors ``[B,3,HH,WW]``."""
        return list(self.decode_stacked(latent, frames, times, scale, mode).unbind(0))

    def decode_test(self, latent, frames, times, scale=None, mode: str | None = None) -> list[torch.Tensor]:
        """``LunaTokis.decoding_test`` (``Sakuya_arch_test.py:461-598``, what ``VideoSRBaseModel.test`` runs): as ``decode`` but
        the bilinear frame gathers read the x4-upsampled frame pair.  ``scale`` is the reference's integer factor
        (``:467``) or, as the shipped evaluation loops pass it, an output size tuple.  At x4 (the default) the decoder's own
        mode is used -- on the tensor-core path the upsampled-frame terms share the query grid and ride inside the Q
        table; at any other size stage B's term is resampled onto the query grid once and stage D's terms are gathered
        per timestep at the warped positions."""
        H, W = int(latent.shape[-2]), int(latent.shape[-1])
        if scale is not None and not isinstance(scale, (tuple, list)):
            scale = (H * int(scale), W * int(scale))
        return list(self.decode_stacked(latent, frames, times, scale, mode=mode or self.mode, test_variant=True).unbind(0))

    @staticmethod
    def memory_window(H: int, W: int, HH: int, WW: int, center) -> tuple[int, int, int, int]:
        """Rows ``[x0,x1)`` and columns ``[y0,y1)`` of ``decoding_memory``'s 4H x 4W window around ``center`` (normalised
        (y, x) in [-1,1]), clamped into the raster exactly as the method does (``Sakuya_arch_test.py:636-650``)."""
        H0, W0 = 4 * H, 4 * W
        c0 = ((float(center[0]) + 1) / 2) * HH
        c1 = ((float(center[1]) + 1) / 2) * WW
        x0, x1, y0, y1 = int(c0) - H0 // 2, int(c0) + H0 - H0 // 2, int(c1) - W0 // 2, int(c1) + W0 - W0 // 2
        if x0 < 0:
            x0, x1 = 0, x1 - x0
        elif x1 > HH:
            x0, x1 = x0 - (x1 - HH), HH
        if y0 < 0:
            y0, y1 = 0, y1 - y0
        elif y1 > WW:
            y0, y1 = y0 - (y1 - WW), WW
        return x0, x1, y0, y1

    def decode_memory(self, latent, frames, times, scale, center, mode: str | None = None) -> list[torch.Tensor]:
        """``LunaTokis.decoding_memory`` (``Sakuya_arch_test.py:600-861``) WITHOUT its side effects (the hard-coded
        ``/home/users/...`` directories and JPEG saves, ``:609-651``): stage A on the whole ``scale = (HH, WW)`` raster,
        stages B-E -- ``decoding_test``'s upsampled frames, ``warpgrid2`` -- on the 4H x 4W window around ``center``.
        Returns ``T`` tensors ``[B,3,4H,4W]``.  Decoded as a row + column window (``stif_decode_window``) in the decoder's
        precision mode; the window is cropped out of the full-size output tensor."""
        H, W = int(latent.shape[-2]), int(latent.shape[-1])
        HH, WW = int(scale[0]), int(scale[1])
        if HH < 4 * H or WW < 4 * W:
            raise ValueError("decoding_memory needs an output raster at least as large as its 4H x 4W window")
        x0, x1, y0, y1 = self.memory_window(H, W, HH, WW, center)
        full = self.decode_stacked(latent, frames, times, (HH, WW), mode=mode or self.mode, rows=(x0, x1), cols=(y0, y1), halo=HH,
                                   test_variant=True, warp_from_coord=True)
        return list(full[:, :, :, x0:x1, y0:y1].contiguous().unbind(0))

    def decode_localensemble(self, latent, frames, times, scale=None, mode: str | None = None) -> torch.Tensor:
        """``LunaTokis.decoding_localensemble`` (``Sakuya_arch_test.py:962-1085``): four shifted passes blended by
        swapped areas (blend weights bit-exact in both modes); batch size 1, returns ``[T,3,HH,WW]``.  ``mode`` defaults
        to the decoder's: "bf16" = tensor-core kernels (RGB within 2e-2), "fp32" = fp32 kernels (within 1e-4)."""
        if latent.shape[0] != 1:
            raise ValueError("decoding_localensemble requires batch size 1 (Sakuya_arch_test.py:989)")
        return self.decode_stacked(latent, frames, list(times), scale, mode=mode or self.mode, local_ensemble=True)[:, 0]

    def decode_host(self, latent: np.ndarray | torch.Tensor, frames, times, scale=None, mode: str | None = None,
                    out: torch.Tensor | None = None, uint8: bool = False) -> torch.Tensor:
        """End-to-end call on HOST buffers (``stif_decode_host``): H2D copy, decode, D2H copy, sync.
        ``uint8=True``: ``[T,B,HH,WW,3]`` uint8 frames as in ``decode_stacked`` (a quarter of the download).
        A ``torch.bfloat16`` latent goes through ``stif_decode_host_bf16`` (half the upload; bit-identical to the call on
        the fp32 latent it was rounded from, because the projection rounds to bf16 anyway)."""
        if not self._loaded:
            raise StifError("load_weights() has not been called")
        bf16_in = isinstance(latent, torch.Tensor) and latent.dtype == torch.bfloat16
        lat = latent.contiguous() if bf16_in else torch.as_tensor(latent, dtype=torch.float32).contiguous()
        fr = torch.as_tensor(frames, dtype=torch.float32).contiguous()
        if lat.device.type != "cpu" or fr.device.type != "cpu":
            raise ValueError("decode_host takes host tensors")
        B, _, _, H, W = lat.shape
        HH, WW = (4 * H, 4 * W) if scale is None else (int(scale[0]), int(scale[1]))
        tm = _times_matrix(times, B)
        T = tm.shape[0]
        shape, dtype = ((T, B, HH, WW, 3), torch.uint8) if uint8 else ((T, B, 3, HH, WW), torch.float32)
        if out is None:
            out = torch.empty(shape, dtype=dtype)
        elif tuple(out.shape) != shape or out.dtype != dtype or not out.is_contiguous() or out.device.type != "cpu":
            raise ValueError(f"out must be a contiguous host {dtype} {list(shape)} tensor")
        m = _MODES[mode or self.mode] | (STIF_FLAG_OUT_U8 if uint8 else 0)
        fn = lib.stif_decode_host_bf16 if bf16_in else lib.stif_decode_host
        check(fn(self._handle, lat.data_ptr(), fr.data_ptr(), B, H, W, HH, WW,
                 tm.ctypes.data_as(C.POINTER(C.c_float)), T, m, out.data_ptr()))
        return out

    # ------------------------------------------------------------------ north-star adapter
    def forward(self, feat, coord: torch.Tensor, cell: torch.Tensor) -> torch.Tensor:
        """``forward(feat, coord, cell)`` adapter.

        ``feat`` = ``(latent [B,3,64,H,W] | [B,192,H,W], frames [B,2,3,H,W])``;
        ``coord`` ``[B,Q,3]`` = (y, x, t) of one or more FULL query rasters in raster order
        (the (y,x,t)+cell convention of ``codes/myutils.py:291-308``);
        ``cell`` ``[B,Q,3]``: ``HH = round(2/cell_y)``, ``WW = round(2/cell_x)``; the third component
        is ignored (the reference has no cell arithmetic).  Returns RGB ``[B,Q,3]``.
        Non-raster ``coord`` is rejected: stage D samples the HR feature map of the whole raster
        (``Sakuya_arch_test.py:429-453``)."""
        latent, frames = feat
        B = latent.shape[0]
        if coord.dim() != 3 or coord.shape[-1] != 3 or coord.shape[0] != B or cell.shape != coord.shape:
            raise ValueError("coord and cell must both be [B,Q,3]")
        HH = int(round(2.0 / float(cell[0, 0, 0])))
        WW = int(round(2.0 / float(cell[0, 0, 1])))
        Q = coord.shape[1]
        if HH < 1 or WW < 1 or Q % (HH * WW) != 0:
            raise ValueError(f"coord holds {Q} queries, not a multiple of the {HH}x{WW} raster implied by cell")
        S = Q // (HH * WW)
        c = coord.detach().to("cpu", torch.float32).reshape(B, S, HH, WW, 3)
        ax_y = torch.from_numpy(_lib.axis_tables(max(1, latent.shape[-2]), HH)["coord"])
        ax_x = torch.from_numpy(_lib.axis_tables(max(1, latent.shape[-1]), WW)["coord"])
        yy = ax_y.view(1, 1, HH, 1).expand(B, S, HH, WW)
        xx = ax_x.view(1, 1, 1, WW).expand(B, S, HH, WW)
        if not (torch.allclose(c[..., 0], yy, atol=2e-6, rtol=0) and torch.allclose(c[..., 1], xx, atol=2e-6, rtol=0)):
            raise ValueError("coord is not a full pixel-centre raster in raster order; arbitrary scattered queries are "
                             "not decodable because stage D warps over the whole HR feature map")
        tt = c[..., 2].reshape(B, S, -1)
        if not torch.equal(tt, tt[..., :1].expand_as(tt)):
            raise ValueError("every raster in coord must carry a single time t")
        times = [tt[:, s, 0].reshape(B, 1) for s in range(S)]
        out = self.decode_stacked(latent, frames, times, (HH, WW))          # [S,B,3,HH,WW]
        return out.permute(1, 0, 3, 4, 2).reshape(B, Q, 3)

    # ------------------------------------------------------------------ introspection
    def last_flow(self, HH: int, WW: int) -> np.ndarray:
        buf = np.empty((HH * WW, 4), np.float32)
        check(lib.stif_debug_last_flow(self._handle, buf.ctypes.data_as(C.POINTER(C.c_float)), buf.size))
        return buf

    def profile(self, enable: bool) -> None:
        """Bracket each kernel group (K0 projection, K1 stage A+B, K2 stage C+D+E) with CUDA events."""
        check(lib.stif_profile_enable(self._handle, int(enable)))

    def profile_read(self) -> dict:
        """{'ms': [K0,K1,K2], 'count': [..]} accumulated since the last read (synchronises)."""
        ms = (C.c_double * 3)()
        cnt = (C.c_int64 * 3)()
        check(lib.stif_profile_read(self._handle, ms, cnt))
        return {"ms": list(ms), "count": list(cnt)}

    def host_pipeline(self, bands: int = 0, halo: int = 0) -> int:
        """Tune ``decode_host``'s band-major pipeline (``stif_debug_host_pipeline``; values <= 0 keep the current setting)
        and return how many calls so far had to repeat stage C-E because a warp out-ran the speculative halo."""
        n = C.c_int64(0)
        check(lib.stif_debug_host_pipeline(self._handle, int(bands), int(halo), C.byref(n)))
        return int(n.value)

    @property
    def launch_count(self) -> int:
        return int(lib.stif_launch_count(self._handle))

    def close(self) -> None:
        if getattr(self, "_handle", None) is not None and self._handle:
            lib.stif_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------- drop-in patching
def _decoder_state(model) -> dict:
    sd = {}
    for net in NET_SHAPES:
        for k, v in getattr(model, net).state_dict().items():
            sd[f"{net}.{k}"] = v
    return sd


def patch_reference_model(model, mode: str = "bf16"):
    """Rebind the reference model's decode methods to the B200 decoder (instance-level).

    After ``patch_reference_model(model)``, ``model(imgs, times)`` (``LunaTokis.forward``,
    ``Sakuya_arch_test.py:1222-1231``) runs the reference encoder (``gen_feat``) and then THIS decoder.
    Weights are snapshotted from ``feat_imnet / flow_imnet / encode_imnet`` now; call
    ``model.stif_refresh_weights()`` after changing them."""
    dev = next(model.parameters()).device
    dec = STIFQueryDecoder(dev, mode=mode)
    dec.load_weights(_decoder_state(model))

    def decoding(times=None, scale=None):
        return dec.decode(model.feat, model.inp, times, scale)

    def decoding_fasttest(times=None, scale=None):
        if model.feat.shape[0] != 1:
            raise ValueError("decoding_fasttest requires batch size 1 (Sakuya_arch_test.py:877)")
        return dec.decode_stacked(model.feat, model.inp, list(times), scale)[:, 0]

    model.decoding = decoding
    model.decoding_fasttest = decoding_fasttest
    model.decoding_localensemble = lambda times=None, scale=None: dec.decode_localensemble(model.feat, model.inp, times, scale)
    model.decoding_test = lambda times=None, scale=None: dec.decode_test(model.feat, model.inp, times, scale)
    model.decoding_memory = lambda times=None, scale=None, center=None, input_img=None, index=0, save=0: dec.decode_memory(
        model.feat, model.inp, times, scale, center)   # (input_img / index / save only drive the reference's JPEG side effects)
    model.decoding_fasttest_memory = decoding_fasttest
    model.stif_decoder = dec
    model.stif_refresh_weights = lambda: dec.load_weights(_decoder_state(model))
    return model


def install_class_patch(luna_tokis_cls, mode: str = "bf16"):
    """Class-level patch for callers that construct the model themselves, e.g. the unmodified
    ``custom_video_test.py`` (``:35``): the decoder is created lazily on first decode."""

    def _dec(self):
        d = self.__dict__.get("_stif_decoder")
        if d is None:
            d = STIFQueryDecoder(next(self.parameters()).device, mode=mode)
            d.load_weights(_decoder_state(self))
            self.__dict__["_stif_decoder"] = d
        return d

    def decoding(self, times=None, scale=None):
        return _dec(self).decode(self.feat, self.inp, times, scale)

    def decoding_fasttest(self, times=None, scale=None):
        if self.feat.shape[0] != 1:
            raise ValueError("decoding_fasttest requires batch size 1 (Sakuya_arch_test.py:877)")
        return _dec(self).decode_stacked(self.feat, self.inp, list(times), scale)[:, 0]

    def decoding_localensemble(self, times=None, scale=None):
        return _dec(self).decode_localensemble(self.feat, self.inp, times, scale)

    def decoding_test(self, times=None, scale=None):
        return _dec(self).decode_test(self.feat, self.inp, times, scale)

    def decoding_memory(self, times=None, scale=None, center=None, input_img=None, index=0, save=0):
        return _dec(self).decode_memory(self.feat, self.inp, times, scale, center)   # (no JPEG side effects)

    luna_tokis_cls.decoding = decoding
    luna_tokis_cls.decoding_fasttest = decoding_fasttest
    luna_tokis_cls.decoding_fasttest_memory = decoding_fasttest
    luna_tokis_cls.decoding_localensemble = decoding_localensemble
    luna_tokis_cls.decoding_test = decoding_test
    luna_tokis_cls.decoding_memory = decoding_memory
    return luna_tokis_cls
