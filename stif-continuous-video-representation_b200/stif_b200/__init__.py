"""stif_b200 -- B200-native (sm_100a) space-time query decoder for STIF.

Drop-in for ONE path of paperwave/STIF-continuous-video-representation: ``LunaTokis.decoding*``
(``codes/models/modules/Sakuya_arch_test.py:364-459``).  The compute lives in
``lib/libstif_b200.so`` (hand-written CUDA behind the C ABI of ``include/stif_b200.h``);
this package is the thin host side.  Importing it without the built library raises.
"""
from ._lib import (LIB_PATH, STIF_MODE_BF16, STIF_MODE_FP32, StifError, axis_tables, dcn_v2_forward,  # noqa: F401
                   ensemble_weights, selftest)
from .decoder import (STIFQueryDecoder, install_class_patch, patch_reference_model,  # noqa: F401
                      weight_keys)

__all__ = ["STIFQueryDecoder", "patch_reference_model", "install_class_patch", "weight_keys", "axis_tables",
           "selftest", "dcn_v2_forward", "StifError", "LIB_PATH", "STIF_MODE_BF16", "STIF_MODE_FP32"]
from .launcher import QueryShardLauncher, WorkUnit, plan_units, units_for_rank  # noqa: E402,F401
