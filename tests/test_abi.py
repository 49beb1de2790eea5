"""C-ABI checks that need no GPU: the library loads, exports every symbol include/stif_b200.h declares,
its host-side axis tables are bit-exact against the torch-generated fixtures, and it fails loudly
(no CPU fallback) when there is no device."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLD, ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "stif_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(stif_[a-z0-9_]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported(built_lib):
    lib = C.CDLL(built_lib)
    names = _declared_symbols()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/stif_b200.h but not exported"


def test_binding_lists_the_same_symbols(stif):
    from stif_b200 import _lib
    assert sorted(_lib.EXPORTS) == _declared_symbols()
    assert _lib.lib.stif_abi_version() == _lib.STIF_ABI_VERSION == 2


def test_axis_tables_bit_exact_vs_torch_fixtures(stif):
    """nearest indices (F.grid_sample) and make_coord axes: the bit-exact contract of the north star."""
    a = np.load(os.path.join(GOLD, "axis_tables.npz"))
    from oracle import restate_np as R
    for n_lr, n_hr in a["pairs"]:
        t = stif.axis_tables(int(n_lr), int(n_hr))
        assert np.array_equal(t["index"], a[f"idx_{n_lr}_{n_hr}"]), (n_lr, n_hr)
        assert np.array_equal(t["coord"], a[f"coord_{n_hr}"]), (n_lr, n_hr)
        o = R.query_axis_tables(int(n_lr), int(n_hr))
        assert np.array_equal(t["rel"], o["rel"])
        assert np.array_equal(t["base"], o["base"])


def test_ensemble_weights_bit_exact_vs_reference(stif):
    """area/tot_area of decoding_localensemble, captured from the reference run (oracle/make_goldens.py): bit-exact."""
    from oracle.make_goldens import CASES
    for name, cfg in CASES.items():
        if cfg["B"] != 1:
            continue
        g = np.load(os.path.join(GOLD, f"case_{name}.npz"))
        HH, WW = g["rgb"].shape[-2:]
        w = stif.ensemble_weights(cfg["H"], cfg["W"], HH, WW)
        assert np.array_equal(w, g["ensemble_weights"]), name


def test_axis_tables_reject_bad_sizes(stif):
    with pytest.raises(stif.StifError):
        stif.axis_tables(0, 16)


def test_workspace_bytes(stif):
    from stif_b200._lib import lib
    assert lib.stif_workspace_bytes(1, 0, 4, 16, 16, 1, 0) == 0
    bf = lib.stif_workspace_bytes(1, 16, 16, 64, 64, 2, 0)
    fp = lib.stif_workspace_bytes(1, 16, 16, 64, 64, 2, 1)
    assert 0 < bf < fp


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device error path")
def test_no_cpu_fallback(stif):
    from stif_b200._lib import lib
    h = C.c_void_p()
    rc = lib.stif_create(C.byref(h), 0)
    assert rc == -2 and not h
    assert b"no CPU fallback" in lib.stif_last_error()
    with pytest.raises(stif.StifError):
        stif.STIFQueryDecoder()
    rc, rep = stif.selftest(0)
    assert rc != 0


def test_weight_key_order_matches_fixture_generator(stif):
    from oracle import synth
    assert stif.weight_keys() == synth.weight_keys()


@pytest.mark.parametrize("shape", [(270, 480, 1080, 1920), (540, 960, 2160, 3840), (270, 480, 1755, 3120), (64, 64, 416, 416),
                                   (96, 40, 384, 163), (20, 24, 13, 17), (7, 9, 1000, 33)])
@pytest.mark.parametrize("bands,forced", [(6, False), (6, True), (3, True), (11, True)])
@pytest.mark.parametrize("halo", [1, 32, 500])
def test_host_band_plan_invariants(stif, shape, bands, forced, halo):
    """The band plan of stif_decode_host's pipeline (pure host arithmetic, csrc/host_plan.cpp).  Whatever the knobs:
    the three boundaries are monotone and complete; stage C-E never runs ahead of stage A+B, and trails it by the halo
    wherever stage A+B is not complete yet (the speculation the kernels then verify); every HR row handed to stage A+B
    has its nearest AND bilinear LR footprint inside the rows uploaded so far (exact, from the library's own tables)."""
    from stif_b200 import _lib
    H, W, HH, WW = shape
    lr_end, ab_end, ce_end, cost = _lib.band_plan(H, W, HH, WW, T=2, bands=bands, forced=forced, halo=halo)
    n = len(lr_end)
    assert n >= 1 and lr_end[-1] == H and ab_end[-1] == HH and ce_end[-1] == HH and cost > 0
    for a in (lr_end, ab_end, ce_end):
        assert (np.diff(a) >= 0).all() and a[0] >= 0
    assert (ce_end <= ab_end).all()
    h = min(halo, HH)
    for k in range(n):
        if ab_end[k] < HH:
            assert ce_end[k] <= max(0, ab_end[k] - h)
    idx = stif.axis_tables(H, HH)["index"]
    u = ((np.arange(HH, dtype=np.float64) + 0.5) / HH) * H - 0.5          # bilinear source row of each HR row centre
    b1 = np.clip(np.floor(u).astype(np.int64) + 1, 0, H - 1)               # lower tap of the bilinear footprint
    for k in range(n):
        rows = ab_end[k]
        if rows > 0:
            assert idx[:rows].max() < lr_end[k] and b1[:rows].max() < lr_end[k], (k, rows, lr_end[k])
    if not forced and HH * WW < 128 * 296 * 4:
        assert n == 1                                                       # small rasters are not split
