"""Import the UNMODIFIED reference decoder (test infrastructure only).

Where the reference comes from, in this order: ``$STIF_REFERENCE_ROOT``; ``/root/reference`` (the build
container -- used by ``make_goldens.py`` to produce the golden fixtures that pin ``restate_np.py`` /
``port_torch.py``); ``oracle/_ref`` -- the byte-for-byte staged copy made by ``oracle/stage_ref.py`` (git-ignored,
travels to the GPU box with the snapshot), which is what the ``-m gpu`` tests and ``bench.py --impl reference`` run:
the reference itself, eager PyTorch, on the B200 or on the box's host cores.  Nothing under
``stif-continuous-video-representation_b200/`` imports this module.

Two harness shims, both outside the reference's arithmetic:

1. ``sys.modules['_ext']`` stub -- ``Sakuya_arch_test.py:9-12`` imports ``DCNv2/dcn_v2.py`` which does
   ``import _ext`` (``dcn_v2.py:11``); the THC-era extension cannot be built against torch 2.11 and
   the decoder never calls it.
2. on a CUDA-less host ``torch.Tensor.cuda`` becomes the identity, because ``decoding``
   hard-codes ``.cuda()`` (``Sakuya_arch_test.py:372-375``).
"""
from __future__ import annotations

import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ROOT = os.path.join(_HERE, "_ref")


def _has_decoder(root: str) -> bool:
    return os.path.isfile(os.path.join(root, "codes", "models", "modules", "Sakuya_arch_test.py"))


def _resolve_root() -> str:
    env = os.environ.get("STIF_REFERENCE_ROOT")
    if env:
        return env
    if _has_decoder("/root/reference"):
        return "/root/reference"
    return STAGED_ROOT


REFERENCE_ROOT = _resolve_root()


def reference_available() -> bool:
    if not _has_decoder(REFERENCE_ROOT):
        return False
    if os.path.abspath(REFERENCE_ROOT) == os.path.abspath(STAGED_ROOT):
        from oracle import stage_ref
        return stage_ref.verify(STAGED_ROOT)      # a staged tree that was edited is not the reference any more
    return True


def reference_kind() -> str:
    """'checkout' (/root/reference or $STIF_REFERENCE_ROOT) or 'staged' (oracle/_ref)."""
    return "staged" if os.path.abspath(REFERENCE_ROOT) == os.path.abspath(STAGED_ROOT) else "checkout"


def load_reference_module():
    """Return the reference's ``models.modules.Sakuya_arch_test`` module object."""
    import torch

    if not reference_available():
        raise RuntimeError(f"reference not found (or staged copy modified) under {REFERENCE_ROOT}; "
                           "run `python -m oracle.stage_ref` in the build container")
    sys.dont_write_bytecode = True
    if "_ext" not in sys.modules:
        sys.modules["_ext"] = types.ModuleType("_ext")
    codes = os.path.join(REFERENCE_ROOT, "codes")
    if codes not in sys.path:
        sys.path.insert(0, codes)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self  # noqa: E731  (harness shim 2)
    import models.modules.Sakuya_arch_test as sat  # type: ignore

    return sat


def build_reference_model(weights: dict):
    """``LunaTokis(64, 6, 8, 5, 40)`` as ``custom_video_test.py:35`` builds it, with the 26 decoder
    tensors overwritten by ``weights`` (numpy arrays keyed by state-dict name)."""
    import torch

    sat = load_reference_module()
    torch.manual_seed(0)
    model = sat.LunaTokis(64, 6, 8, 5, 40)
    sd = {k: torch.from_numpy(v.copy()) for k, v in weights.items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert not any(k.split(".")[0] in ("feat_imnet", "flow_imnet", "encode_imnet") for k in missing), missing
    model.eval()
    return model


def clear_warp_cache():
    """``warplayer.backwarp_tenGrid`` memoises one base grid per flow size forever
    (``warplayer.py:6,26-33``); clear it between shapes."""
    mod = sys.modules.get("models.modules.warplayer")
    if mod is not None:
        mod.backwarp_tenGrid.clear()


def reference_decode(latent, frames, weights: dict, times, scale=None, device: str = "cpu", method: str = "decoding",
                     model=None):
    """Run the reference's own ``LunaTokis.<method>(times, scale)`` (default ``decoding``,
    ``Sakuya_arch_test.py:364-459``) on numpy inputs; returns ``[T,B,3,HH,WW]`` float32 (numpy).

    ``device='cuda'``: the eager fp32 reference on the GPU (TF32 off) -- BASELINE.md section 3's "second baseline", the
    oracle for RGB tolerances at sizes the CPU cannot finish in seconds.  ``times``: floats, or rows of per-item times."""
    import numpy as np
    import torch

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    if model is None:
        model = build_reference_model(weights)
    model = model.to(device)
    clear_warp_cache()
    lat = torch.as_tensor(latent, dtype=torch.float32, device=device)
    fr = torch.as_tensor(frames, dtype=torch.float32, device=device)
    B = lat.shape[0]
    model.feat, model.inp = lat, fr
    tm = np.asarray(times, dtype=np.float32)
    with torch.no_grad():
        if method == "decoding":
            if tm.ndim == 1:
                tl = [torch.tensor([[float(t)]], dtype=torch.float32, device=device) for t in tm]
            else:
                tl = [torch.tensor(row, dtype=torch.float32, device=device).view(B, 1) for row in tm]
            out = torch.stack(model.decoding(tl, scale), 0)
        else:
            out = getattr(model, method)([float(t) for t in tm], scale)[:, None]
    if device != "cpu":
        torch.cuda.synchronize()
    res = out.float().cpu().numpy()
    model.feat = model.inp = None
    return res
