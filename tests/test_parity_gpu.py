"""Parity of the CUDA path (through the C ABI) against the oracle and the reference-generated
fixtures.  GPU box only (-m gpu); nothing here reads /root/reference.

Bars (BASELINE.json north_star): RGB max-abs <= 1e-4 in fp32 mode, <= 2e-2 in bf16 mode,
PSNR delta <= 0.05 dB; nearest indices / rel bit-exact (host tables: tests/test_abi.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLD
from oracle import port_torch, synth
from oracle import restate_np as R
from oracle.make_goldens import CASES, MEMORY_VARIANT_CASES, TEST_VARIANT_CASES

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "bf16": 2e-2}


@pytest.fixture(scope="module")
def decoders(stif):
    cache = {}

    def get(wseed, stress, mode):
        key = (wseed, stress, mode)
        if key not in cache:
            d = stif.STIFQueryDecoder(0, mode=mode)
            d.load_weights(synth.make_weights(wseed, stress))
            cache[key] = d
        return cache[key]
    return get


def _run(dec, lat, fr, times, scale):
    out = dec.decode_stacked(torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda(), _times(times), scale)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _times(times):
    t = np.asarray(times, dtype=np.float32)
    return [torch.tensor(r, dtype=torch.float32).view(-1, 1) for r in t] if t.ndim == 2 else [float(x) for x in t]


def test_tcgen05_selftest(stif):
    rc, report = stif.selftest(0)
    print(report)
    assert rc == 0, report


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name", list(CASES))
def test_golden_cases(name, mode, decoders):
    cfg = CASES[name]
    g = np.load(os.path.join(GOLD, f"case_{name}.npz"))
    lat, fr = synth.make_inputs(cfg["iseed"], cfg["B"], cfg["H"], cfg["W"], cfg["latent_std"])
    dec = decoders(cfg["wseed"], cfg["stress"], mode)
    rgb = _run(dec, lat, fr, cfg["times"], cfg["scale"])
    assert rgb.shape == g["rgb"].shape
    err = np.abs(rgb - g["rgb"]).max()
    print(f"{name} {mode}: rgb max-abs {err:.3e}")
    assert err <= TOL[mode]
    # flow of the last slab (HR-pixel units) against the reference's flow_imnet output
    T, B = len(cfg["times"]), cfg["B"]
    HH, WW = rgb.shape[-2:]
    flow = dec.last_flow(HH, WW)
    ref = g[f"flow_{T - 1}"][(B - 1) * HH * WW:]
    ferr = np.abs(flow - ref).max()
    print(f"{name} {mode}: flow max-abs {ferr:.3e} px (|flow| max {np.abs(ref).max():.1f})")
    assert ferr <= (2e-3 if mode == "fp32" else 0.5)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("stress", [False, True])
def test_config1(stress, mode, decoders):
    """BASELINE.json config 1: 64x64 latent -> 256x256, t = i/8, against the reference's own output sample."""
    g = np.load(os.path.join(GOLD, f"config1_{'stress' if stress else 'init'}.npz"))
    lat, fr = synth.make_inputs(0, 1, 64, 64, 0.05)
    rgb = _run(decoders(0, stress, mode), lat, fr, [i / 8.0 for i in range(8)], None)
    err = np.abs(rgb[:, :, :, 1::5, 2::5] - g["rgb_sub"]).max()
    print(f"config1 stress={stress} {mode}: max-abs {err:.3e}")
    assert err <= TOL[mode]
    assert np.abs(rgb.mean(axis=(1, 2, 3, 4)) - g["mean"]).max() <= TOL[mode] / 10


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_config3_noninteger_scale_vs_oracle(mode, decoders):
    """x6.5 (exact .5 ties in the nearest index), 8 intermediate timesteps k/9, 64x64 latent -> 416x416."""
    lat, fr = synth.smooth_inputs(5, 1, 64, 64, 0.05)
    w = synth.make_weights(1, True)
    times = [k / 9.0 for k in range(1, 9)]
    scale = (int(6.5 * 64), int(6.5 * 64))
    ref = port_torch.decode(lat, fr, w, times, scale).numpy()
    rgb = _run(decoders(1, True, mode), lat, fr, times, scale)
    err = np.abs(rgb - ref).max()
    psnr = R.psnr255(rgb, ref)
    print(f"config3 {mode}: max-abs {err:.3e} PSNR(new,ref) {psnr:.1f} dB")
    assert err <= TOL[mode]
    assert psnr >= 50.0


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_psnr_delta_against_pseudo_ground_truth(mode, decoders):
    """|PSNR(new,GT) - PSNR(ref,GT)| <= 0.05 dB with GT = the reference output + noise-free offset image.
    (utils/util.py:140-151 PSNR on clamp(0,1)*255.)"""
    lat, fr = synth.smooth_inputs(7, 1, 48, 40, 0.05)
    w = synth.make_weights(2, True)
    ref = port_torch.decode(lat, fr, w, [0.3], None).numpy()
    rng = np.random.default_rng(0)
    gt = ref + rng.normal(0, 0.02, ref.shape).astype(np.float32)      # a 34 dB "ground truth"
    rgb = _run(decoders(2, True, mode), lat, fr, [0.3], None)
    d = abs(R.psnr255(rgb, gt) - R.psnr255(ref, gt))
    print(f"{mode}: PSNR delta {d:.4f} dB")
    assert d <= 0.05


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_properties_at_config2_size(mode, decoders):
    """Size-independent properties at BASELINE.json's config-2 size (270x480 latent -> 1080x1920, t in {0, 0.5}):
    determinism, independence of timesteps, row-band decode == full decode, fp32/bf16 agreement."""
    lat, fr = synth.smooth_inputs(11, 1, 270, 480, 0.05)
    dec = decoders(0, True, mode)
    a = _run(dec, lat, fr, [0.0, 0.5], None)
    assert a.shape == (2, 1, 3, 1080, 1920) and np.isfinite(a).all()
    b = _run(dec, lat, fr, [0.0, 0.5], None)
    assert np.array_equal(a, b)                                        # deterministic
    c = _run(dec, lat, fr, [0.5], None)
    assert np.array_equal(a[1], c[0])                                  # slabs are independent (Sakuya_arch_test.py:380)
    latc, frc = torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda()
    band = torch.zeros((1, 1, 3, 1080, 1920), device="cuda")
    dec.decode_stacked(latc, frc, [0.5], None, rows=(300, 420), halo=64, out=band)
    torch.cuda.synchronize()
    assert np.array_equal(band.cpu().numpy()[0, 0, :, 300:420], a[1, 0, :, 300:420])
    assert float(band[0, 0, :, :300].abs().max()) == 0.0
    other = _run(decoders(0, True, "fp32" if mode == "bf16" else "bf16"), lat, fr, [0.0, 0.5], None)
    err = np.abs(a - other).max()
    print(f"config2 size: fp32 vs bf16 max-abs {err:.3e}, PSNR {R.psnr255(a, other):.1f} dB")
    assert err <= 2e-2


def test_properties_at_config4_size(decoders):
    """BASELINE.json config 4 (540x960 latent -> 2160x3840, 8 timesteps i/8, one timestep per GPU when sharded): the
    oracle cannot reach 66 M queries, so parity at this size rests on size-independent properties -- the bf16 tensor-core
    path against the fp32 path (itself <= 1e-4 from the reference wherever the oracle reaches) on two of the timesteps,
    a slab decoded alone == the same slab decoded in the batch (what the sharding launcher relies on), the row-band
    decode (the unit when slabs < GPUs) == the full decode, and the host entry point == the device path."""
    lat, fr = synth.smooth_inputs(12, 1, 540, 960, 0.05)
    times = [i / 8 for i in range(8)]
    bf, fp = decoders(0, True, "bf16"), decoders(0, True, "fp32")
    a = _run(bf, lat, fr, times, None)
    assert a.shape == (8, 1, 3, 2160, 3840) and np.isfinite(a).all()
    for c in (3, 7):
        alone = _run(bf, lat, fr, [times[c]], None)
        assert np.array_equal(alone[0], a[c])
        ref32 = _run(fp, lat, fr, [times[c]], None)
        err = np.abs(ref32[0] - a[c]).max()
        print(f"config4 size t={times[c]}: bf16 vs fp32 max-abs {err:.3e}")
        assert err <= 2e-2
    latc, frc = torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda()
    band = torch.zeros((1, 1, 3, 2160, 3840), device="cuda")
    bf.decode_stacked(latc, frc, [times[5]], None, rows=(1080, 1350), halo=48, out=band)
    torch.cuda.synchronize()
    assert np.array_equal(band.cpu().numpy()[0, 0, :, 1080:1350], a[5, 0, :, 1080:1350])
    del band
    host = bf.decode_host(lat, fr, times[:5], None).numpy()               # 5 > the resident group of 4
    assert np.array_equal(host, a[:5])


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name", [n for n in CASES if CASES[n]["B"] == 1])
def test_localensemble_mode(name, mode, decoders):
    """decoding_localensemble through STIF_FLAG_LOCAL_ENSEMBLE against the reference's own output: fp32 kernels and the
    tensor-core kernels (four passes of stage A tables -> stage B with F gathered at the shifted nearest HR pixel ->
    stage C-E, blended with the bit-exact area weights)."""
    cfg = CASES[name]
    g = np.load(os.path.join(GOLD, f"case_{name}.npz"))
    lat, fr = synth.make_inputs(cfg["iseed"], 1, cfg["H"], cfg["W"], cfg["latent_std"])
    dec = decoders(cfg["wseed"], cfg["stress"], mode)
    out = dec.decode_localensemble(torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda(), cfg["times"], cfg["scale"])
    torch.cuda.synchronize()
    err = np.abs(out.cpu().numpy() - g["rgb_localensemble"]).max()
    print(f"{name} local-ensemble {mode}: max-abs {err:.3e} (differs from plain decode by {np.abs(g['rgb_localensemble'] - g['rgb'][:, 0]).max():.2e})")
    assert err <= TOL[mode]


def test_localensemble_bf16_at_config2_size(decoders):
    """Tensor-core local ensemble at 270x480 -> 1080x1920 against the fp32 kernels (size-independent check), plus timing."""
    lat, fr = synth.smooth_inputs(13, 1, 270, 480, 0.05)
    L, F = torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda()
    bf, fp = decoders(0, True, "bf16"), decoders(0, True, "fp32")
    a = bf.decode_localensemble(L, F, [0.5], None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    a2 = bf.decode_localensemble(L, F, [0.5], None)
    e1.record()
    torch.cuda.synchronize()
    assert torch.equal(a, a2)
    b = fp.decode_localensemble(L, F, [0.5], None)
    err = float((a - b).abs().max())
    print(f"local ensemble 1080p: bf16 vs fp32 max-abs {err:.3e}; bf16 {e0.elapsed_time(e1):.2f} ms per timestep (4 passes)")
    assert err <= 2e-2


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_config5_caller_path_small(mode, decoders):
    """Encoder-produced latents + padded frames + the 8 time tensors of custom_video_test.py:44-52, against the
    reference's own model(imgs, times) output (tests/golden/e2e_small.npz)."""
    g = np.load(os.path.join(GOLD, "e2e_small.npz"))
    dec = decoders(4, True, mode)
    times = [torch.tensor([i / 8.0])[None] for i in range(8)]                 # exactly the caller's time_Tensors (:50)
    out = dec.decode(torch.from_numpy(g["latent"]).cuda(), torch.from_numpy(g["frames"]).cuda(), times)
    assert isinstance(out, list) and len(out) == 8 and tuple(out[0].shape) == (1, 3, 112, 144)
    rgb = torch.stack(out, 0).cpu().numpy()
    err = np.abs(rgb - g["rgb"]).max()
    d = abs(R.psnr255(rgb, g["rgb"] + 0.02) - R.psnr255(g["rgb"], g["rgb"] + 0.02))
    print(f"config5-small {mode}: max-abs {err:.3e}, PSNR(new,ref) {R.psnr255(rgb, g['rgb']):.1f} dB")
    assert err <= TOL[mode]


def test_config5_shape_properties(decoders):
    """Config 5's decode shape (480x272 padded latent -> 1088x1920, 8 timesteps): determinism and bf16/fp32 agreement on
    one frame pair; sizes where the oracle cannot run."""
    lat, fr = synth.smooth_inputs(21, 1, 272, 480, 0.05)
    times = [i / 8.0 for i in range(8)]
    a = _run(decoders(0, True, "bf16"), lat, fr, times, None)
    assert a.shape == (8, 1, 3, 1088, 1920) and np.isfinite(a).all()
    b = _run(decoders(0, True, "bf16"), lat, fr, times[3:5], None)
    assert np.array_equal(a[3:5], b)
    c = _run(decoders(0, True, "fp32"), lat, fr, times[3:4], None)
    assert np.abs(a[3] - c[0]).max() <= 2e-2


def test_band_halo_violation_is_reported(decoders, stif):
    lat, fr = synth.make_inputs(1, 1, 16, 16, 0.05)
    dec = decoders(1, True, "fp32")                                   # stress weights: flows of ~13 px
    out = torch.zeros((1, 1, 3, 104, 104), device="cuda")
    with pytest.raises(stif.StifError, match="halo"):
        dec.decode_stacked(torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda(), [0.5], (104, 104),
                           rows=(40, 60), halo=1, out=out)


def test_forward_feat_coord_cell_adapter(decoders):
    """north-star surface forward(feat, coord, cell): (y,x,t) rasters + cell=(2/HH, 2/WW, .)."""
    lat, fr = synth.make_inputs(0, 1, 16, 16, 0.05)
    dec = decoders(0, False, "fp32")
    HH = WW = 64
    ay, ax = R.clamp_axis(R.make_axis(HH)), R.clamp_axis(R.make_axis(WW))
    yy, xx = np.meshgrid(ay, ax, indexing="ij")
    slabs = [np.stack([yy, xx, np.full_like(yy, t)], -1).reshape(-1, 3) for t in (0.0, 0.375)]
    coord = torch.from_numpy(np.concatenate(slabs, 0)[None].astype(np.float32))
    cell = torch.tensor([2.0 / HH, 2.0 / WW, 0.5]).expand_as(coord).contiguous()
    feat = (torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda())
    out = dec(feat, coord, cell).cpu().numpy()                        # [1, 2*Q, 3]
    g = np.load(os.path.join(GOLD, "case_x4_init.npz"))["rgb"]       # [2,1,3,64,64]
    ref = np.concatenate([g[t, 0].reshape(3, -1).T for t in range(2)], 0)[None]
    assert np.abs(out - ref).max() <= 1e-4
    bad = coord.clone()
    bad[0, 5, 1] += 0.01
    with pytest.raises(ValueError, match="raster"):
        dec(feat, bad, cell)


def test_host_entry_point(decoders):
    lat, fr = synth.make_inputs(0, 1, 16, 16, 0.05)
    dec = decoders(0, False, "fp32")
    out = dec.decode_host(lat, fr, [0.0, 0.375], None).numpy()
    g = np.load(os.path.join(GOLD, "case_x4_init.npz"))["rgb"]
    assert np.abs(out - g).max() <= 1e-4


@pytest.mark.parametrize("stress,halo,expect_respin", [(False, 32, False), (True, 1, True)])
def test_host_pipeline_matches_device_path(stif, stress, halo, expect_respin):
    """stif_decode_host's band-major pipeline (bf16): uploads, kernels and downloads of different row bands overlap and
    stage C-E trails stage A-B by a speculative halo.  Whatever the knobs, the result must be BIT-identical to the
    device-buffer path (same kernels, same per-query arithmetic); with +-20 px stress flows and a 1-row halo the
    speculation must miss, be detected and be repaired.  T = 6 also covers the timesteps beyond the resident group."""
    dec = stif.STIFQueryDecoder(0, mode="bf16")
    dec.load_weights(synth.make_weights(1, stress))
    lat, fr = synth.make_inputs(5, 1, 96, 40, 1.0 if stress else 0.05)
    times = [0.0, 0.2, 0.4, 0.6, 0.8, 1.0]
    ref = _run(dec, lat, fr, times, (384, 163))
    before = dec.host_pipeline(bands=6, halo=halo)
    out = dec.decode_host(lat, fr, times, (384, 163)).numpy()
    respins = dec.host_pipeline() - before
    assert np.isfinite(out).all()
    assert np.array_equal(out, ref)
    assert (respins > 0) == expect_respin, respins
    # a second call after a miss runs with the doubled halo and still agrees
    out2 = dec.decode_host(lat, fr, times, (384, 163)).numpy()
    assert np.array_equal(out2, ref)


def test_host_pipeline_batch_of_two(stif):
    dec = stif.STIFQueryDecoder(0, mode="bf16")
    dec.load_weights(synth.make_weights(2, True))
    lat, fr = synth.make_inputs(6, 2, 24, 20, 0.3)
    times = [[0.25, 0.75], [0.5, 0.1]]
    ref = _run(dec, lat, fr, times, (61, 77))
    out = dec.decode_host(lat, fr, _times(times), (61, 77)).numpy()
    assert np.array_equal(out, ref)


def _to_u8_like_reference(rgb):
    """custom_video_test.py:102: (img.clamp(0,1).permute(1,2,0) * 255).numpy().astype(np.uint8), per (t, b) frame."""
    x = torch.from_numpy(rgb).clamp(0.0, 1.0).permute(0, 1, 3, 4, 2) * 255
    return np.ascontiguousarray(x.numpy().astype(np.uint8))


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_uint8_output_flag(mode, decoders):
    """STIF_FLAG_OUT_U8 = the conversion the reference's caller applies before saving a frame.  Bit-exact against that
    conversion of the library's own fp32 result (device and host entry points, odd sizes, row bands), and within
    tolerance*255 + 1 code values of the converted reference fixture."""
    cfg = CASES["down_stress"]
    lat, fr = synth.make_inputs(cfg["iseed"], cfg["B"], cfg["H"], cfg["W"], cfg["latent_std"])
    dec = decoders(cfg["wseed"], cfg["stress"], mode)
    # offset + gain so that the clamp bites on both sides
    rgb = _run(dec, lat, fr, cfg["times"], cfg["scale"])
    want = _to_u8_like_reference(rgb)
    L, F = torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda()
    got = dec.decode_stacked(L, F, _times(cfg["times"]), cfg["scale"], uint8=True).cpu().numpy()
    assert got.dtype == np.uint8 and got.shape == want.shape
    assert np.array_equal(got, want)
    assert np.array_equal(dec.decode_host(lat, fr, _times(cfg["times"]), cfg["scale"], uint8=True).numpy(), want)
    band = torch.zeros_like(torch.from_numpy(want)).cuda()
    dec.decode_stacked(L, F, _times(cfg["times"]), cfg["scale"], rows=(3, 9), halo=13, out=band, uint8=True)
    assert np.array_equal(band.cpu().numpy()[:, :, 3:9], want[:, :, 3:9]) and not band.cpu().numpy()[:, :, 9:].any()
    ref = _to_u8_like_reference(np.load(os.path.join(GOLD, "case_down_stress.npz"))["rgb"])
    assert np.abs(got.astype(np.int32) - ref.astype(np.int32)).max() <= int(TOL[mode] * 255) + 1


def test_uint8_host_pipeline_banded(stif):
    dec = stif.STIFQueryDecoder(0, mode="bf16")
    dec.load_weights(synth.make_weights(1, True))
    lat, fr = synth.make_inputs(5, 1, 96, 40, 0.05)
    times = [0.0, 0.3, 0.6, 0.9, 1.0]
    want = _to_u8_like_reference(_run(dec, lat, fr, times, (384, 163)))
    dec.host_pipeline(bands=5, halo=16)
    assert np.array_equal(dec.decode_host(lat, fr, times, (384, 163), uint8=True).numpy(), want)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name", list(TEST_VARIANT_CASES))
def test_decoding_test_variant(name, mode, decoders):
    """STIF_FLAG_TEST_VARIANT = `LunaTokis.decoding_test` (Sakuya_arch_test.py:461-598): the bilinear frame gathers read
    the x4-upsampled frame pair.  Both precision modes against the reference's own run (int scale as the method takes it,
    and the equivalent output-size tuple the shipped eval loops pass), flow included.  On the tensor-core kernels x4 is
    the fast path (frame terms inside the Q planes); x3 / x5 go through the resampled / warped frame-term tables."""
    cfg = TEST_VARIANT_CASES[name]
    g = np.load(os.path.join(GOLD, f"case_{name}.npz"))
    lat, fr = synth.make_inputs(cfg["iseed"], 1, cfg["H"], cfg["W"], cfg["latent_std"])
    dec = decoders(cfg["wseed"], cfg["stress"], mode)
    L, F = torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda()
    out = torch.stack(dec.decode_test(L, F, cfg["times"], cfg["scale"]), 0)
    torch.cuda.synchronize()
    err = np.abs(out.cpu().numpy() - g["rgb"]).max()
    flow_err = np.abs(dec.last_flow(out.shape[-2], out.shape[-1]) - g["flow"][-1]).max()
    print(f"{name} {mode}: decoding_test rgb max-abs {err:.3e}, flow max-abs {flow_err:.3e}")
    assert err <= TOL[mode] and flow_err <= (1e-3 if mode == "fp32" else 0.5)
    if cfg["scale"] is not None:
        size = (cfg["H"] * cfg["scale"], cfg["W"] * cfg["scale"])
        again = torch.stack(dec.decode_test(L, F, cfg["times"], size), 0)
        assert torch.equal(again, out)
    plain = torch.stack(dec.decode(L, F, cfg["times"], None if cfg["scale"] is None else size), 0)
    assert float((plain - out).abs().max()) > 1e-3        # a different function from `decoding`


def test_random_shapes_fuzz(stif):
    """Seeded sweep over awkward geometries (rasters smaller than one 128-query tile, widths that are not multiples of the
    8x16 K2 tile, down-scaling, B up to 3, T up to 9): tensor-core path vs fp32 path (itself pinned to the reference
    wherever fixtures exist), row band == full, host entry == device entry, uint8 flag == conversion of the fp32 output."""
    rng = np.random.default_rng(2024)
    bf = stif.STIFQueryDecoder(0, mode="bf16")
    fp = stif.STIFQueryDecoder(0, mode="fp32")
    w = synth.make_weights(9, True)
    bf.load_weights(w)
    fp.load_weights(w)
    worst = 0.0
    for it in range(24):
        H, W = int(rng.integers(3, 40)), int(rng.integers(3, 40))
        HH, WW = int(rng.integers(1, 6 * H)), int(rng.integers(1, 6 * W))
        B, T = int(rng.integers(1, 4)), int(rng.integers(1, 10))
        lat, fr = synth.make_inputs(1000 + it, B, H, W, float(rng.choice([0.05, 0.3])))
        times = [list(rng.random(B).astype(np.float32)) for _ in range(T)]
        a = _run(bf, lat, fr, times, (HH, WW))
        b = _run(fp, lat, fr, times, (HH, WW))
        assert a.shape == (T, B, 3, HH, WW) and np.isfinite(a).all(), (H, W, HH, WW, B, T)
        err = float(np.abs(a - b).max())
        if min(HH, WW) >= 8:
            # (rasters a few pixels across are kept for the exactness checks below but not for the tolerance: with +-10 px
            # stress flows most of their warped taps sit in the zero-padded border ramp, where a 0.03 px bf16 flow
            # error is a 3 % feature error -- 3e-2 ... 1e-1 RGB there, 3e-3 ... 1e-2 everywhere else)
            worst = max(worst, err)
            assert err <= 2e-2, (H, W, HH, WW, B, T, err)
        host = bf.decode_host(lat, fr, _times(times), (HH, WW)).numpy()
        assert np.array_equal(host, a), (H, W, HH, WW, B, T)
        if HH >= 4:
            r0 = int(rng.integers(0, HH - 1)); r1 = int(rng.integers(r0 + 1, HH + 1))
            band = torch.zeros((T, B, 3, HH, WW), device="cuda")
            bf.decode_stacked(torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda(), _times(times), (HH, WW), rows=(r0, r1), halo=HH,
                              out=band)
            assert np.array_equal(band.cpu().numpy()[:, :, :, r0:r1], a[:, :, :, r0:r1]), (H, W, HH, WW, r0, r1)
        u8 = bf.decode_stacked(torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda(), _times(times), (HH, WW), uint8=True)
        assert np.array_equal(u8.cpu().numpy(), _to_u8_like_reference(a)), (H, W, HH, WW)
    print(f"fuzz: worst bf16-vs-fp32 max-abs {worst:.3e} over 24 geometries")


def test_random_shapes_fuzz_modes(stif):
    """Second seeded sweep: the local-ensemble mode (tensor-core vs fp32 kernels) and the band-major host pipeline with
    forced band counts / tiny halos on mid-size rasters (speculation misses included) -- results must not depend on them."""
    rng = np.random.default_rng(77)
    bf = stif.STIFQueryDecoder(0, mode="bf16")
    fp = stif.STIFQueryDecoder(0, mode="fp32")
    w = synth.make_weights(4, True)
    bf.load_weights(w)
    fp.load_weights(w)
    worst = 0.0
    for it in range(8):
        H, W = int(rng.integers(8, 48)), int(rng.integers(8, 48))
        HH, WW = int(rng.integers(2 * H, 5 * H)), int(rng.integers(2 * W, 5 * W))
        lat, fr = synth.make_inputs(2000 + it, 1, H, W, 0.3)
        L, F = torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda()
        times = [float(t) for t in rng.random(int(rng.integers(1, 4)))]
        a = bf.decode_localensemble(L, F, times, (HH, WW))
        b = fp.decode_localensemble(L, F, times, (HH, WW))
        err = float((a - b).abs().max())
        worst = max(worst, err)
        assert err <= 2e-2, (H, W, HH, WW, err)
    misses = 0
    for it in range(8):
        H, W = int(rng.integers(40, 120)), int(rng.integers(20, 90))
        HH, WW = int(rng.integers(2 * H, 5 * H)), int(rng.integers(2 * W, 5 * W))
        T = int(rng.integers(1, 7))
        lat, fr = synth.make_inputs(3000 + it, 1, H, W, float(rng.choice([0.05, 1.0])))
        times = [float(t) for t in rng.random(T)]
        ref = _run(bf, lat, fr, times, (HH, WW))
        before = bf.host_pipeline(bands=int(rng.integers(2, 9)), halo=int(rng.choice([1, 4, 16, 64])))
        host = bf.decode_host(lat, fr, times, (HH, WW)).numpy()
        misses += bf.host_pipeline() - before
        assert np.array_equal(host, ref), (H, W, HH, WW, T)
    print(f"fuzz modes: worst ensemble bf16-vs-fp32 {worst:.3e}; host pipeline speculation misses repaired: {misses}")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name", list(MEMORY_VARIANT_CASES))
def test_decoding_memory_variant(name, mode, decoders, stif):
    """`LunaTokis.decoding_memory` (windowed zoom queries, Sakuya_arch_test.py:600-861) = STIF_FLAG_TEST_VARIANT |
    STIF_FLAG_WARP_FROM_COORD on a row + column window (stif_decode_window); both precision modes against the reference's
    own run (the tensor-core K2 computes the window's columns only)."""
    cfg = MEMORY_VARIANT_CASES[name]
    g = np.load(os.path.join(GOLD, f"case_{name}.npz"))
    lat, fr = synth.make_inputs(cfg["iseed"], 1, cfg["H"], cfg["W"], cfg["latent_std"])
    dec = decoders(cfg["wseed"], cfg["stress"], mode)
    assert dec.memory_window(cfg["H"], cfg["W"], cfg["scale"][0], cfg["scale"][1], cfg["center"]) == tuple(int(v) for v in g["window"])
    out = torch.stack(dec.decode_memory(torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda(), cfg["times"], cfg["scale"],
                                        cfg["center"]), 0)
    torch.cuda.synchronize()
    err = np.abs(out.cpu().numpy() - g["rgb"]).max()
    print(f"{name} {mode}: decoding_memory window {tuple(g['window'])} rgb max-abs {err:.3e}")
    assert out.shape == g["rgb"].shape and err <= TOL[mode]


def test_decoding_test_variant_tensor_core_x4(decoders):
    """decoding_test at x4 on the tensor-core kernels: the upsampled-frame grid is the query grid there, so stage B's term is
    the query's own texel and stage D's terms are folded into the Q planes.  Against the reference fixture (bf16
    tolerance) and, at 270x480 -> 1080x1920, against the fp32 kernels; any other size falls back to fp32."""
    cfg = TEST_VARIANT_CASES["testvar_x4_init"]
    g = np.load(os.path.join(GOLD, "case_testvar_x4_init.npz"))
    lat, fr = synth.make_inputs(cfg["iseed"], 1, cfg["H"], cfg["W"], cfg["latent_std"])
    bf = decoders(cfg["wseed"], cfg["stress"], "bf16")
    out = torch.stack(bf.decode_test(torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda(), cfg["times"], None), 0)
    torch.cuda.synchronize()
    err = np.abs(out.cpu().numpy() - g["rgb"]).max()
    lat, fr = synth.smooth_inputs(14, 1, 270, 480, 0.05)
    L, F = torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda()
    bfs, fps = decoders(0, True, "bf16"), decoders(0, True, "fp32")
    a = torch.stack(bfs.decode_test(L, F, [0.0, 0.5], 4), 0)
    b = torch.stack(fps.decode_test(L, F, [0.0, 0.5], 4), 0)
    plain = torch.stack(bfs.decode(L, F, [0.0, 0.5], None), 0)
    err2 = float((a - b).abs().max())
    print(f"decoding_test x4 tensor-core: vs reference fixture {err:.3e}; 1080p vs fp32 kernels {err2:.3e} "
          f"(differs from plain decoding by {float((a - plain).abs().max()):.2e})")
    assert err <= 2e-2 and err2 <= 2e-2 and float((a - plain).abs().max()) > 1e-3
    # away from x4 the tensor-core kernels take the general path (resampled / warped frame-term tables); at 1080p-ish size:
    Ls, Fs = L[:, :, :, :135, :240].contiguous(), F[:, :, :, :135, :240].contiguous()
    c = torch.stack(bfs.decode_test(Ls, Fs, [0.3], (877, 1560)), 0)                   # x6.5
    d = torch.stack(fps.decode_test(Ls, Fs, [0.3], (877, 1560)), 0)
    err3 = float((c - d).abs().max())
    print(f"decoding_test x6.5 tensor-core (135x240 -> 877x1560) vs fp32 kernels {err3:.3e}")
    assert err3 <= 2e-2


class _SineLayer(torch.nn.Module):          # SIREN.py:14-45 shape: a Linear called `linear`
    def __init__(self, i, o):
        super().__init__()
        self.linear = torch.nn.Linear(i, o)


class _Siren(torch.nn.Module):              # SIREN.py:48-79 shape: `net` = sine layers + a final plain Linear
    def __init__(self, dims):
        super().__init__()
        self.net = torch.nn.Sequential(*[_SineLayer(dims[i], dims[i + 1]) for i in range(len(dims) - 2)],
                                       torch.nn.Linear(dims[-2], dims[-1]))


class _StandInLuna(torch.nn.Module):
    """The attributes of LunaTokis the patch touches (Sakuya_arch_test.py:306-311, :361, :1224): three SIREN sub-modules with
    the reference's state-dict key layout, `feat`, `inp`.  (The reference itself cannot travel to the GPU box.)"""
    def __init__(self, weights):
        super().__init__()
        from stif_b200.decoder import NET_SHAPES
        for net, dims in NET_SHAPES.items():
            setattr(self, net, _Siren(dims))
        self.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()}, strict=True)


@pytest.mark.parametrize("how", ["instance", "class"])
def test_drop_in_patch_rebinds_every_decode_method(how, stif):
    """patch_reference_model / install_class_patch: weights snapshotted from the three sub-modules' state dicts, and
    decoding / decoding_fasttest / decoding_localensemble / decoding_test / decoding_memory answer with the reference's
    calling conventions and return types; checked against the reference fixtures of the same seeded case."""
    from stif_b200.decoder import install_class_patch, patch_reference_model
    cfg = CASES["x4_init"]
    g = np.load(os.path.join(GOLD, "case_x4_init.npz"))
    gt = np.load(os.path.join(GOLD, "case_testvar_x4_init.npz"))
    lat, fr = synth.make_inputs(cfg["iseed"], 1, cfg["H"], cfg["W"], cfg["latent_std"])
    if how == "class":
        cls = type("PatchedLuna", (_StandInLuna,), {})
        install_class_patch(cls, mode="fp32")
        model = cls(synth.make_weights(cfg["wseed"], cfg["stress"])).cuda()
    else:
        model = patch_reference_model(_StandInLuna(synth.make_weights(cfg["wseed"], cfg["stress"])).cuda(), mode="fp32")
    model.feat, model.inp = torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda()
    times = [torch.tensor([[float(t)]]).cuda() for t in cfg["times"]]           # custom_video_test.py:50
    preds = model.decoding(times, cfg["scale"])
    assert isinstance(preds, list) and len(preds) == len(times) and preds[0].shape == (1, 3, 64, 64)
    assert np.abs(torch.stack(preds, 0).cpu().numpy() - g["rgb"]).max() <= 1e-4
    fast = model.decoding_fasttest([float(t) for t in cfg["times"]], cfg["scale"])
    assert fast.shape == (2, 3, 64, 64) and np.abs(fast.cpu().numpy() - g["rgb_fasttest"]).max() <= 1e-4
    ens = model.decoding_localensemble([float(t) for t in cfg["times"]], cfg["scale"])
    assert np.abs(ens.cpu().numpy() - g["rgb_localensemble"]).max() <= 1e-4
    tst = model.decoding_test(times, None)
    assert np.abs(torch.stack(tst, 0).cpu().numpy() - gt["rgb"]).max() <= 1e-4
    win = model.decoding_memory(times, (100, 90), np.array([0.1, -0.2]), input_img=None)
    assert isinstance(win, list) and win[0].shape == (1, 3, 64, 64)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_row_band_with_nan_prefilled_workspace(mode, stif):
    """A row-band decode must not read Q-table rows it never wrote, even through zero-weight taps (0 x NaN = NaN): the
    workspace is pre-filled with 0xFF bytes (NaN in fp32 and fp16) and the band must still equal the full decode.  Exact
    integer warp positions (zero flow at x4 hits them for the linspace base's end points) and out-of-grid taps both occur."""
    lat, fr = synth.smooth_inputs(21, 1, 24, 20, 0.05)
    w = synth.make_weights(0, False)                     # init weights: flows ~ +-0.1 px, many taps with tiny / zero weights
    dec = stif.STIFQueryDecoder(0, mode=mode)
    dec.load_weights(w)
    latc, frc = torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda()
    full = dec.decode_stacked(latc, frc, [0.5], None)
    torch.cuda.synchronize()
    for rows in ((32, 56), (0, 16), (80, 96)):
        dec._workspace.fill_(0xFF)
        band = dec.decode_stacked(latc, frc, [0.5], None, rows=rows, halo=8)
        torch.cuda.synchronize()
        got = band[0, 0, :, rows[0]:rows[1]]
        assert torch.isfinite(got).all(), rows
        assert torch.equal(got, full[0, 0, :, rows[0]:rows[1]]), rows
        assert float(band[0, 0, :, :rows[0]].abs().max() if rows[0] else 0.0) == 0.0     # rows outside the band: zeros
    dec.close()


def test_bf16_rejects_rasters_beyond_32bit_tap_offsets(stif):
    """The tensor-core K2 stages tap addresses as 32-bit byte offsets: HH*WW > 2^24 (or H*W >= 2^23) must be refused with
    STIF_EINVAL, not decoded wrongly.  The check precedes every buffer access, so dummy pointers suffice."""
    import ctypes as C
    from stif_b200 import _lib
    dec = stif.STIFQueryDecoder(0, mode="bf16")
    dec.load_weights(synth.make_weights(0, False))
    dummy = torch.zeros(64, device="cuda")
    t = (C.c_float * 1)(0.5)
    args = lambda HH, WW, mode: (dec._handle, dummy.data_ptr(), dummy.data_ptr(), 1, 8, 8, HH, WW, t, 1, mode, dummy.data_ptr(), 256,
                                 dummy.data_ptr(), None)
    rc = _lib.lib.stif_decode(*args(4097, 4096, _lib.STIF_MODE_BF16))
    assert rc == -1 and b"too large for STIF_MODE_BF16" in _lib.lib.stif_last_error()
    rc = _lib.lib.stif_decode(*args(4096, 4096, _lib.STIF_MODE_BF16))          # exactly 2^24: allowed (fails later: workspace)
    assert rc == -4, _lib.lib.stif_last_error()
    rc = _lib.lib.stif_decode(*args(4097, 4096, _lib.STIF_MODE_FP32))          # the fp32 path indexes with 64 bits
    assert rc == -4, _lib.lib.stif_last_error()
    dec.close()


def test_prepare_makes_decode_stream_ordered(stif):
    """After stif_prepare the decode of a new geometry can be captured in a CUDA graph (no allocation / sync inside)."""
    from stif_b200 import _lib
    dec = stif.STIFQueryDecoder(0, mode="bf16")
    w = synth.make_weights(0, True)
    dec.load_weights(w)
    lat, fr = synth.smooth_inputs(3, 1, 20, 28, 0.05)
    latc, frc = torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda()
    _lib.check(_lib.lib.stif_prepare(dec._handle, 20, 28, 80, 112, _lib.STIF_MODE_BF16))
    out = torch.empty((1, 1, 3, 80, 112), device="cuda")
    dec._workspace_for(1, 20, 28, 80, 112, 1, _lib.STIF_MODE_BF16)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            dec.decode_stacked(latc, frc, [0.25], (80, 112), out=out)
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    ref = dec.decode_stacked(latc, frc, [0.25], (80, 112))
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    dec.close()


def test_host_entry_bf16_latent_is_bit_identical(stif):
    """stif_decode_host_bf16 on the RN-rounded latent == stif_decode_host on the fp32 latent it was rounded from (the
    projection rounds to bf16 anyway), fp32 and uint8 outputs, banded pipeline included."""
    lat, fr = synth.smooth_inputs(13, 1, 96, 128, 0.05)
    dec = stif.STIFQueryDecoder(0, mode="bf16")
    dec.load_weights(synth.make_weights(0, True))
    dec.host_pipeline(bands=4)
    lat32 = torch.from_numpy(lat)
    lat16 = lat32.to(torch.bfloat16)
    for u8 in (False, True):
        a = dec.decode_host(lat32, fr, [0.25, 0.75], None, uint8=u8)
        b = dec.decode_host(lat16, fr, [0.25, 0.75], None, uint8=u8)
        assert torch.equal(a, b), u8
    dev = dec.decode_stacked(lat32.cuda(), torch.from_numpy(fr).cuda(), [0.25, 0.75], None, uint8=True)
    torch.cuda.synchronize()
    assert torch.equal(dev.cpu(), b)                     # fused uint8 output stage == host pipeline's
    dec.close()


def _scaled_first_layers(w, s):
    w = {k: v.copy() for k, v in w.items()}
    for net in ("feat_imnet", "flow_imnet", "encode_imnet"):
        w[f"{net}.net.0.linear.weight"] *= s
    return w


@pytest.mark.parametrize("scale,std,bound", [(1.0, 0.05, 5e-3), (1.0, 1.0, 2e-2), (2.0, 1.0, None), (3.0, 3.0, None)])
def test_bf16_error_envelope_vs_sine_argument_scale(scale, std, bound, stif):
    """How the tensor-core mode's RGB error grows with the size of the first-layer sine arguments (stress weights: RGB
    gain x10, flows of +-20 px, white-noise latents).  The numpy model of the kernels' rounding points
    (oracle/emulate_hoisted.py) predicts the GPU error; the dominant term is the bf16 rounding of the MMA operands
    (error ~ output gain x |argument| x 2^-9), NOT the fp16 tables or the fp16 blend accumulation: the model with fp32
    tables predicts the same error.  Inside the documented envelope (arguments up to ~6 rad) the 2e-2
    bound holds with margin; beyond it the bound is out of reach of bf16 operands and the fp32 mode is the answer."""
    from oracle import emulate_hoisted as E
    w = _scaled_first_layers(synth.make_weights(1, True), scale)
    lat, fr = synth.make_inputs(3, 1, 16, 16, std)
    ref = R.decode(lat, fr, w, [0.3], None)
    model = E.decode(lat, fr, w, [0.3], None, mode="bf16")
    model32 = E.decode(lat, fr, w, [0.3], None, mode="bf16", table_round=lambda v: v)
    dec = stif.STIFQueryDecoder(0, mode="bf16")
    dec.load_weights(w)
    rgb = _run(dec, lat, fr, [0.3], None)
    dec.close()
    err, pred, pred32 = np.abs(rgb - ref).max(), np.abs(model - ref).max(), np.abs(model32 - ref).max()
    print(f"w0 x{scale} latent std {std}: GPU bf16 err {err:.3e}, model {pred:.3e}, model with fp32 tables {pred32:.3e}")
    assert np.isfinite(rgb).all()
    if bound is not None:
        assert err <= 1.5 * pred + 1e-3              # inside the envelope the GPU tracks the model of its own rounding points
        assert err <= bound
    # (beyond the envelope -- measured 3.3e-2 at 11 rad, 0.46 at 50 rad -- the fp16 accumulation of K2's sixteen-tap
    #  blend adds to the operand-rounding term the model predicts, 2.4e-2 / 9.7e-2: both are past the bound already)
    fp = stif.STIFQueryDecoder(0, mode="fp32")
    fp.load_weights(w)
    e32 = np.abs(_run(fp, lat, fr, [0.3], None) - ref).max()
    fp.close()
    print(f"   fp32 mode (split-bf16 tensor-core GEMMs) err {e32:.3e}")
    assert e32 <= (1e-4 if bound is not None else 1e-3)   # 1e-4 inside the envelope; 1.9e-4 measured at 50 rad arguments


def test_projected_tables_saturate_instead_of_overflowing(stif):
    """Latents large enough to push first-layer pre-activations past fp16's range give finite RGB (saturating table
    stores), not NaN."""
    w = synth.make_weights(0, False)
    lat, fr = synth.make_inputs(5, 1, 12, 12, 0.05)
    lat = (lat * 1e5).astype(np.float32)             # |tab| ~ 1e5 rad: meaningless sines, but finite
    dec = stif.STIFQueryDecoder(0, mode="bf16")
    dec.load_weights(w)
    rgb = _run(dec, lat, fr, [0.5], None)
    dec.close()
    assert np.isfinite(rgb).all()


def test_fp32_mode_gemm_backends_agree(tmp_path):
    """STIF_MODE_FP32's dense layers have three back-ends: the persistent split-bf16 tcgen05 GEMM (default), the first one-tile-per-CTA
    tensor-core kernel (STIF_HP_V1=1) and the SIMT SGEMM anchor (STIF_FP32_SIMT=1; plain fp32 FMAs, libdevice sinf).  The switches are
    read when the library loads weights, so each runs in its own process on the same seeded inputs (odd sizes, two timesteps, stress
    weights): the tensor-core paths must sit within 3e-5 of the anchor and within 2e-5 of each other."""
    import subprocess
    import sys
    script = (
        "import sys, numpy as np, torch\n"
        f"sys.path.insert(0, {os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'stif-continuous-video-representation_b200')!r})\n"
        "import stif_b200\n"
        "from stif_b200 import synthetic as synth\n"
        "dec = stif_b200.STIFQueryDecoder(0, mode='fp32'); dec.load_weights(synth.make_weights(4, True))\n"
        "lat, fr = synth.make_inputs(21, 1, 37, 53, 0.3)\n"
        "out = dec.decode_stacked(torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda(), [0.25, 0.8], (301, 197))\n"
        "np.save(sys.argv[1], out.cpu().numpy())\n")
    outs = {}
    for name, env in (("v2", {}), ("v1", {"STIF_HP_V1": "1"}), ("simt", {"STIF_FP32_SIMT": "1"})):
        path = str(tmp_path / f"{name}.npy")
        subprocess.run([sys.executable, "-c", script, path], check=True, env=dict(os.environ, **env), timeout=600)
        outs[name] = np.load(path)
    assert np.isfinite(outs["simt"]).all() and np.abs(outs["simt"]).max() > 0.05
    d_anchor = max(np.abs(outs["v2"] - outs["simt"]).max(), np.abs(outs["v1"] - outs["simt"]).max())
    d_tc = np.abs(outs["v2"] - outs["v1"]).max()
    print(f"fp32 back-ends: tensor-core vs SIMT anchor {d_anchor:.3e}, v2 vs v1 {d_tc:.3e}")
    assert d_anchor <= 3e-5 and d_tc <= 2e-5
