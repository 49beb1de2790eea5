"""The oracle against the reference-generated fixtures (CPU; no GPU, no /root/reference needed).

Fixtures: tests/golden/*.npz, written by oracle/make_goldens.py from the UNMODIFIED reference
(LunaTokis.decoding, Sakuya_arch_test.py:364-459).  Tolerances: indices / coordinates / rel
bit-exact; fp32 stages <= 5e-6 (BLAS summation order only)."""
import os

import numpy as np
import pytest

from conftest import GOLD
from oracle import emulate_hoisted as E
from oracle import port_torch, synth
from oracle import restate_np as R
from oracle.make_goldens import CASES, MEMORY_VARIANT_CASES, TEST_VARIANT_CASES, checksum, window_of

FP32_TOL = 5e-6


def _inputs(cfg):
    w = synth.make_weights(cfg["wseed"], cfg["stress"])
    lat, fr = synth.make_inputs(cfg["iseed"], cfg["B"], cfg["H"], cfg["W"], cfg["latent_std"])
    return w, lat, fr


@pytest.mark.parametrize("name", list(CASES))
def test_synth_matches_fixture_checksums(name):
    cfg = CASES[name]
    g = np.load(os.path.join(GOLD, f"case_{name}.npz"))
    w, lat, fr = _inputs(cfg)
    assert checksum(lat, fr) == pytest.approx(float(g["input_checksum"]), rel=1e-12)
    assert checksum(*w.values()) == pytest.approx(float(g["weight_checksum"]), rel=1e-12)


@pytest.mark.parametrize("name", list(CASES))
def test_restatement_rgb_and_stages(name):
    cfg = CASES[name]
    g = np.load(os.path.join(GOLD, f"case_{name}.npz"))
    w, lat, fr = _inputs(cfg)
    rgb, st = R.decode(lat, fr, w, cfg["times"], cfg["scale"], return_stages=True)
    assert rgb.shape == g["rgb"].shape
    assert np.abs(rgb - g["rgb"]).max() <= FP32_TOL
    # stage tensors of the last (t, b) slab against the hooked MLP inputs/outputs of the reference
    T = len(cfg["times"])
    B = cfg["B"]
    Q = rgb.shape[-1] * rgb.shape[-2]
    sel = g["sel"]
    mine = sel[(sel >= (B - 1) * Q)] - (B - 1) * Q          # sampled queries that fall in the last batch item
    rows = np.nonzero(sel >= (B - 1) * Q)[0]
    if mine.size and mine.max() < min(Q, 1 << 16):
        fin = g[f"feat_in_{T - 1}"][rows]
        assert np.array_equal(st["feat_in"][mine][:, :198], fin[:, :198])        # nearest gathers: exact copies
        assert np.array_equal(st["feat_in"][mine][:, 198:200], fin[:, 198:200])  # rel_coord: bit-exact
        assert np.array_equal(st["feat_in"][mine][:, 200], fin[:, 200])          # t
        assert np.abs(st["hr"][mine] - g[f"hr_{T - 1}"][rows]).max() <= FP32_TOL
        assert np.abs(st["flow_in"][mine] - g[f"flow_in_{T - 1}"][rows]).max() <= FP32_TOL
        assert np.abs(st["enc_in"][mine] - g[f"enc_in_{T - 1}"][rows]).max() <= 2e-5   # warped gathers (flow up to 27 px)
    assert np.abs(st["flow"] - g[f"flow_{T - 1}"][(B - 1) * Q:]).max() <= 5e-5


@pytest.mark.parametrize("name", [n for n in CASES if CASES[n]["B"] == 1])
def test_fasttest_is_the_same_function(name):
    """decoding_fasttest (Sakuya_arch_test.py:863-960) is bit-identical to decoding in the reference."""
    g = np.load(os.path.join(GOLD, f"case_{name}.npz"))
    assert np.array_equal(g["rgb_fasttest"], g["rgb"][:, 0])


@pytest.mark.parametrize("name", [n for n in CASES if CASES[n]["B"] == 1])
def test_localensemble_restatement(name):
    """decoding_localensemble (Sakuya_arch_test.py:962-1085): RGB and the bit-exact blend weights."""
    cfg = CASES[name]
    g = np.load(os.path.join(GOLD, f"case_{name}.npz"))
    w, lat, fr = _inputs(cfg)
    out = R.decode_localensemble(lat, fr, w, cfg["times"], cfg["scale"])
    assert np.abs(out - g["rgb_localensemble"]).max() <= FP32_TOL
    HH, WW = out.shape[-2:]
    assert np.array_equal(R.ensemble_weights(cfg["H"], cfg["W"], HH, WW), g["ensemble_weights"])


@pytest.mark.parametrize("name", list(CASES))
def test_torch_port(name):
    cfg = CASES[name]
    g = np.load(os.path.join(GOLD, f"case_{name}.npz"))
    w, lat, fr = _inputs(cfg)
    rgb = port_torch.decode(lat, fr, w, cfg["times"], cfg["scale"]).numpy()
    assert np.abs(rgb - g["rgb"]).max() <= FP32_TOL


def test_e2e_small_encoder_latents():
    """Config 5 in miniature: latents produced by the reference ENCODER (gen_feat through a torchvision `_ext` shim) for
    a padded frame pair, decoded at t = i/8 as custom_video_test.py:44-52 does."""
    g = np.load(os.path.join(GOLD, "e2e_small.npz"))
    w = synth.make_weights(4, True)
    assert checksum(*w.values()) == pytest.approx(float(g["weight_checksum"]), rel=1e-12)
    out = R.decode(g["latent"], g["frames"], w, [i / 8.0 for i in range(8)], None)
    assert np.abs(out - g["rgb"]).max() <= FP32_TOL


def test_axis_tables_bit_exact():
    a = np.load(os.path.join(GOLD, "axis_tables.npz"))
    for n_lr, n_hr in a["pairs"]:
        t = R.query_axis_tables(int(n_lr), int(n_hr))
        assert np.array_equal(t["i"], a[f"idx_{n_lr}_{n_hr}"]), (n_lr, n_hr)     # F.grid_sample nearest
        assert np.array_equal(t["c"], a[f"coord_{n_hr}"])                          # make_coord + clamp
        assert np.array_equal(t["lr_c"], a[f"lrcoord_{n_lr}"])
        assert np.abs(t["base"] - a[f"linspace_{n_hr}"]).max() <= 6e-8             # <= 1 ulp (see restate_np docstring)


def test_integer_shortcut_is_wrong_at_non_integer_scale():
    """Guards the design decision to replay the fp32 chain."""
    a = np.load(os.path.join(GOLD, "axis_tables.npz"))
    idx = a["idx_270_1755"]
    shortcut = (np.arange(1755) * 270) // 1755
    assert (idx != shortcut).sum() > 0
    assert np.array_equal(a["idx_270_1080"], (np.arange(1080) * 270) // 1080)


@pytest.mark.parametrize("stress", [False, True])
def test_config1_sample(stress):
    """BASELINE.json config 1 (64x64 latent -> 256x256, 8 timesteps), strided sample of the reference's output."""
    g = np.load(os.path.join(GOLD, f"config1_{'stress' if stress else 'init'}.npz"))
    w = synth.make_weights(0, stress)
    lat, fr = synth.make_inputs(0, 1, 64, 64, 0.05)
    rgb = port_torch.decode(lat, fr, w, [i / 8.0 for i in range(8)], None).numpy()
    assert np.abs(rgb[:, :, :, 1::5, 2::5] - g["rgb_sub"]).max() <= FP32_TOL
    assert np.abs(rgb.mean(axis=(1, 2, 3, 4)) - g["mean"]).max() <= 1e-6


@pytest.mark.parametrize("name", list(CASES))
def test_hoisted_algebra_model(name):
    """The product's hoisted formulation is the same function in fp32, and its
    bf16 numerics model stays inside the 2e-2 bound the north star sets for bf16 mode."""
    cfg = CASES[name]
    g = np.load(os.path.join(GOLD, f"case_{name}.npz"))
    w, lat, fr = _inputs(cfg)
    rgb = E.decode(lat, fr, w, cfg["times"], cfg["scale"], mode="fp32")
    assert np.abs(rgb - g["rgb"]).max() <= 1e-5
    rgb16 = E.decode(lat, fr, w, cfg["times"], cfg["scale"], mode="bf16")
    assert np.abs(rgb16 - g["rgb"]).max() <= 2e-2


@pytest.mark.parametrize("name", list(TEST_VARIANT_CASES))
def test_decoding_test_variant_restatement(name):
    """`decoding_test` (Sakuya_arch_test.py:461-598): the x4-bilinear-upsampled frame pair feeds the bilinear frame gathers.
    The restatement (incl. ATen's upsample source-index formula) against the reference's own run; the variant is
    numerically distinct from `decoding` (2e-3 ... 5e-2 here), so this is its own golden set."""
    cfg = TEST_VARIANT_CASES[name]
    g = np.load(os.path.join(GOLD, f"case_{name}.npz"))
    w = synth.make_weights(cfg["wseed"], cfg["stress"])
    lat, fr = synth.make_inputs(cfg["iseed"], 1, cfg["H"], cfg["W"], cfg["latent_std"])
    assert abs(checksum(lat, fr) - float(g["input_checksum"])) < 1e-6 * float(g["input_checksum"])
    out, st = R.decode(lat, fr, w, cfg["times"], cfg["scale"], return_stages=True, upsampled_frames=True)
    assert out.shape == g["rgb"].shape
    assert np.abs(out - g["rgb"]).max() <= FP32_TOL
    assert np.abs(st["flow"] - g["flow"][-1]).max() <= 5e-5          # flows reach +-26 px here
    size = None if cfg["scale"] is None else (cfg["H"] * cfg["scale"], cfg["W"] * cfg["scale"])
    assert np.abs(R.decode(lat, fr, w, cfg["times"], size) - g["rgb"]).max() > 1e-3   # not the same function as `decoding`


@pytest.mark.parametrize("name", list(MEMORY_VARIANT_CASES))
def test_decoding_memory_variant_restatement(name):
    """`decoding_memory` (Sakuya_arch_test.py:600-861; fixtures generated with its file-system side effects neutralised):
    stage A on the full raster, `decoding_test`'s stages B-E with `warpgrid2` on the clamped 4H x 4W window."""
    cfg = MEMORY_VARIANT_CASES[name]
    g = np.load(os.path.join(GOLD, f"case_{name}.npz"))
    w = synth.make_weights(cfg["wseed"], cfg["stress"])
    lat, fr = synth.make_inputs(cfg["iseed"], 1, cfg["H"], cfg["W"], cfg["latent_std"])
    win = window_of(cfg["H"], cfg["W"], cfg["scale"][0], cfg["scale"][1], cfg["center"])
    assert tuple(g["window"]) == win
    out = R.decode(lat, fr, w, cfg["times"], cfg["scale"], upsampled_frames=True, window=win)
    assert out.shape == g["rgb"].shape and np.abs(out - g["rgb"]).max() <= FP32_TOL
    crop = R.decode(lat, fr, w, cfg["times"], cfg["scale"], upsampled_frames=True)[:, :, :, win[0]:win[1], win[2]:win[3]]
    assert np.abs(crop - g["rgb"]).max() > 1e-3          # warpgrid2 makes it more than a crop of decoding_test
