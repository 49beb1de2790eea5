"""Seeded synthetic weights / latents / frames for the STIF query decoder (used by bench.py, the tests and the oracle).

Everything is drawn from ``numpy.random.default_rng`` (PCG64: the stream is stable across
numpy versions and machines), so the build container and the GPU box generate identical
inputs without shipping them.

The weight distributions follow the reference's SIREN initialisation
(``codes/models/modules/SIREN.py:35-42`` for sine layers, ``:63-67`` for the outermost
linear layer; biases use ``nn.Linear``'s default ``U(-1/sqrt(fan_in), 1/sqrt(fan_in))``).
Layer shapes are those of ``LunaTokis.__init__`` (``Sakuya_arch_test.py:306-311``), i.e. the
26 decoder tensors of ``latest_G.pth``.
"""
from __future__ import annotations

import numpy as np

OMEGA0 = 30.0

# (network, [fan_in, widths...]) -- the last width is the outermost linear layer.
NET_SHAPES = {
    "feat_imnet": [201, 64, 64, 256, 64],
    "flow_imnet": [263, 64, 64, 256, 4],
    "encode_imnet": [525, 64, 64, 256, 256, 3],
}


def weight_keys() -> list[str]:
    """The 26 state-dict keys in canonical (C-ABI) order."""
    keys = []
    for net, dims in NET_SHAPES.items():
        n_layers = len(dims) - 1
        for li in range(n_layers):
            last = li == n_layers - 1
            stem = f"{net}.net.{li}" if last else f"{net}.net.{li}.linear"
            keys += [f"{stem}.weight", f"{stem}.bias"]
    return keys


def weight_shapes() -> dict[str, tuple[int, ...]]:
    shapes = {}
    for net, dims in NET_SHAPES.items():
        n_layers = len(dims) - 1
        for li in range(n_layers):
            last = li == n_layers - 1
            stem = f"{net}.net.{li}" if last else f"{net}.net.{li}.linear"
            shapes[f"{stem}.weight"] = (dims[li + 1], dims[li])
            shapes[f"{stem}.bias"] = (dims[li + 1],)
    return shapes


def make_weights(seed: int = 0, stress: bool = False) -> dict[str, np.ndarray]:
    """SIREN-initialised decoder weights; ``stress`` applies SURVEY.md section 8c's variant
    (flows of roughly +-20 HR pixels, RGB of order 0..0.4) so that warps and tolerances
    are actually exercised."""
    rng = np.random.default_rng(1000 + seed)
    out: dict[str, np.ndarray] = {}
    for net, dims in NET_SHAPES.items():
        n_layers = len(dims) - 1
        for li in range(n_layers):
            fan_in, fan_out = dims[li], dims[li + 1]
            last = li == n_layers - 1
            stem = f"{net}.net.{li}" if last else f"{net}.net.{li}.linear"
            if li == 0:
                bound = 1.0 / fan_in
            else:
                bound = float(np.sqrt(6.0 / fan_in) / OMEGA0)
            w = rng.uniform(-bound, bound, size=(fan_out, fan_in)).astype(np.float32)
            bb = 1.0 / float(np.sqrt(fan_in))
            b = rng.uniform(-bb, bb, size=(fan_out,)).astype(np.float32)
            out[f"{stem}.weight"] = w
            out[f"{stem}.bias"] = b
    if stress:
        out["flow_imnet.net.3.weight"] = (out["flow_imnet.net.3.weight"] * np.float32(200.0)).astype(np.float32)
        out["flow_imnet.net.3.bias"] = (out["flow_imnet.net.3.bias"] + np.float32(3.0)).astype(np.float32)
        out["encode_imnet.net.4.weight"] = (out["encode_imnet.net.4.weight"] * np.float32(10.0)).astype(np.float32)
    return out


def make_inputs(seed: int, B: int, H: int, W: int, latent_std: float = 0.05):
    """latent ``[B,3,64,H,W]`` (what ``gen_feat`` leaves in ``self.feat``,
    ``Sakuya_arch_test.py:361``) and the LR frame pair ``[B,2,3,H,W]`` in [0,1)
    (``self.inp``, ``:1224``)."""
    rng = np.random.default_rng(2000 + seed)
    latent = (latent_std * rng.standard_normal((B, 3, 64, H, W))).astype(np.float32)
    frames = rng.random((B, 2, 3, H, W), dtype=np.float32)
    return latent, frames


def smooth_inputs(seed: int, B: int, H: int, W: int, latent_std: float = 0.05):
    """Spatially smooth variant (low-pass filtered noise): closer to what a real encoder
    emits, used by PSNR-style checks where white noise would make every pixel an outlier."""
    latent, frames = make_inputs(seed, B, H, W, latent_std)

    def blur(x):
        for ax in (-1, -2):
            x = (np.roll(x, 1, ax) + 2.0 * x + np.roll(x, -1, ax)) / 4.0
        return x

    for _ in range(3):
        latent = blur(latent)
        frames = blur(frames)
    latent = (latent * (latent_std / max(float(latent.std()), 1e-12))).astype(np.float32)
    return latent, frames.astype(np.float32)
