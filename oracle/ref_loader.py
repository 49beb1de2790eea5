"""Import the UNMODIFIED reference decoder from ``/root/reference`` (test infrastructure only).

Works only where the reference checkout is mounted (the build container).  Nothing under
``tests -m gpu``, ``smoke()`` or ``bench.py`` may call this: ``/root/reference`` does not
exist on the GPU box.  Its one job is to produce the golden fixtures (``make_goldens.py``)
that pin ``restate_np.py`` / ``port_torch.py``.

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

REFERENCE_ROOT = os.environ.get("STIF_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "codes", "models", "modules", "Sakuya_arch_test.py"))


def load_reference_module():
    """Return the reference's ``models.modules.Sakuya_arch_test`` module object."""
    import torch

    if not reference_available():
        raise RuntimeError(f"reference checkout not found under {REFERENCE_ROOT}")
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
