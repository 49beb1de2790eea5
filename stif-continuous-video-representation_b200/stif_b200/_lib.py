"""ctypes binding of ``libstif_b200.so`` (the C ABI declared in ``include/stif_b200.h``).

There is deliberately no fallback: if the shared library has not been built
(``python -c 'import __graft_entry__ as g; g.build()'``) importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# STIF_LIB selects another in-tree build of the same library (tuning variants under build_variants/); there is still no
# fallback -- the named file must exist.
LIB_PATH = os.path.abspath(os.environ["STIF_LIB"]) if os.environ.get("STIF_LIB") else \
    os.path.normpath(os.path.join(_HERE, "..", "lib", "libstif_b200.so"))

STIF_OK = 0
STIF_MODE_BF16 = 0
STIF_MODE_FP32 = 1
STIF_FLAG_LOCAL_ENSEMBLE = 0x100
STIF_FLAG_OUT_U8 = 0x200
STIF_FLAG_TEST_VARIANT = 0x400
STIF_FLAG_WARP_FROM_COORD = 0x800
STIF_NUM_WEIGHT_TENSORS = 26
STIF_ABI_VERSION = 2

# every symbol include/stif_b200.h declares (tests/test_abi.py checks the .so exports them all)
EXPORTS = [
    "stif_abi_version", "stif_last_error", "stif_create", "stif_destroy", "stif_load_weights",
    "stif_prepare", "stif_workspace_bytes", "stif_decode", "stif_decode_rows", "stif_decode_window", "stif_decode_host", "stif_decode_host_bf16", "stif_dcn_v2_forward", "stif_axis_tables",
    "stif_ensemble_weights", "stif_debug_last_flow", "stif_debug_host_pipeline", "stif_debug_band_plan", "stif_launch_count", "stif_profile_enable", "stif_profile_read", "stif_selftest",
]


class StifError(RuntimeError):
    """A libstif_b200 call failed; the message is ``stif_last_error()``."""


def _load():
    if not os.path.isfile(LIB_PATH):
        raise StifError(
            f"{LIB_PATH} not found: build the CUDA library first (__graft_entry__.build()); "
            "stif_b200 has no CPU or PyTorch fallback")
    lib = C.CDLL(LIB_PATH)
    fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_void_p
    lib.stif_abi_version.restype = C.c_int
    lib.stif_last_error.restype = C.c_char_p
    lib.stif_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.stif_destroy.argtypes = [vp]
    lib.stif_load_weights.argtypes = [vp, C.POINTER(vp), C.c_int]
    lib.stif_prepare.argtypes = [vp] + [C.c_int] * 5
    lib.stif_workspace_bytes.argtypes = [C.c_int] * 7
    lib.stif_workspace_bytes.restype = C.c_size_t
    lib.stif_decode.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fp, C.c_int, C.c_int,
                                vp, C.c_size_t, vp, vp]
    lib.stif_decode_rows.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fp, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_int, vp, C.c_size_t, vp, vp]
    lib.stif_decode_window.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fp, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_size_t, vp, vp]
    lib.stif_decode_host.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fp, C.c_int, C.c_int, vp]
    lib.stif_decode_host_bf16.argtypes = lib.stif_decode_host.argtypes
    lib.stif_dcn_v2_forward.argtypes = [vp] * 5 + [C.c_int] * 14 + [vp, vp]
    lib.stif_axis_tables.argtypes = [C.c_int, C.c_int, fp, ip, fp, fp]
    lib.stif_ensemble_weights.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, fp, C.c_size_t]
    lib.stif_debug_last_flow.argtypes = [vp, fp, C.c_size_t]
    lib.stif_debug_host_pipeline.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_int64)]
    lib.stif_debug_band_plan.argtypes = [C.c_int] * 10 + [ip, ip, ip, C.POINTER(C.c_double)]
    lib.stif_launch_count.argtypes = [vp]
    lib.stif_launch_count.restype = C.c_int64
    lib.stif_profile_enable.argtypes = [vp, C.c_int]
    lib.stif_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.stif_selftest.argtypes = [C.c_int, C.c_char_p, C.c_size_t]
    for name in EXPORTS:  # fail at import, not at first use, if a symbol is missing
        getattr(lib, name)
    if lib.stif_abi_version() != STIF_ABI_VERSION:
        raise StifError(f"ABI mismatch: library {lib.stif_abi_version()} != binding {STIF_ABI_VERSION}")
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != STIF_OK:
        raise StifError(f"libstif_b200 error {rc}: {lib.stif_last_error().decode(errors='replace')}")


def axis_tables(n_lr: int, n_hr: int):
    """Host-side per-axis query tables (numpy): coord, index, rel, base."""
    import numpy as np

    coord = np.empty(n_hr, np.float32)
    index = np.empty(n_hr, np.int32)
    rel = np.empty(n_hr, np.float32)
    base = np.empty(n_hr, np.float32)
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
    check(lib.stif_axis_tables(n_lr, n_hr, coord.ctypes.data_as(fp), index.ctypes.data_as(ip),
                               rel.ctypes.data_as(fp), base.ctypes.data_as(fp)))
    return {"coord": coord, "index": index, "rel": rel, "base": base}


def ensemble_weights(H: int, W: int, HH: int, WW: int):
    """decoding_localensemble's blend weights [4, HH*WW] (host computation, bit-exact contract)."""
    import numpy as np

    w = np.empty((4, HH * WW), np.float32)
    check(lib.stif_ensemble_weights(H, W, HH, WW, w.ctypes.data_as(C.POINTER(C.c_float)), w.size))
    return w


def selftest(device: int = 0) -> tuple[int, str]:
    buf = C.create_string_buffer(8192)
    rc = lib.stif_selftest(device, buf, len(buf))
    return rc, buf.value.decode(errors="replace")


def band_plan(H: int, W: int, HH: int, WW: int, T: int = 2, bands: int = 6, forced: bool = False, halo: int = 32,
              num_sms: int = 148):
    """Band plan of the host pipeline (``stif_debug_band_plan``): ``(lr_end, ab_end, ce_end, cost_us)`` per band."""
    import numpy as np
    n_max = 64
    a, b, c = (np.zeros(n_max, np.int32) for _ in range(3))
    cost = C.c_double(0.0)
    ip = C.POINTER(C.c_int32)
    n = lib.stif_debug_band_plan(H, W, HH, WW, T, bands, int(forced), halo, num_sms, n_max, a.ctypes.data_as(ip),
                                 b.ctypes.data_as(ip), c.ctypes.data_as(ip), C.byref(cost))
    if n < 0:
        check(n)
    return a[:n].copy(), b[:n].copy(), c[:n].copy(), float(cost.value)


def dcn_v2_forward(input, weight, bias, offset, mask, kh, kw, sh, sw, ph, pw, dh, dw, dg):
    """``_ext.dcn_v2_forward`` (``DCNv2/dcn_v2.py:24-27``) on the B200 kernel; torch CUDA tensors in, a new tensor out.
    Returns ``None`` when the geometry is not the reference encoder's (the caller keeps its fallback)."""
    import torch

    if not (input.is_cuda and input.dtype == torch.float32):
        return None
    B, Cin, H, W = input.shape
    Cout = weight.shape[0]
    if (Cin, Cout, kh, kw, sh, sw, ph, pw, dh, dw, dg) != (64, 64, 3, 3, 1, 1, 1, 1, 1, 1, 8):
        return None
    t = [x.contiguous().float() for x in (input, weight, bias, offset, mask)]
    out = torch.empty((B, Cout, H, W), dtype=torch.float32, device=input.device)
    with torch.cuda.device(input.device):
        rc = lib.stif_dcn_v2_forward(*[x.data_ptr() for x in t], B, Cin, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, dg,
                                     out.data_ptr(), torch.cuda.current_stream(input.device).cuda_stream)
    check(rc)
    return out
