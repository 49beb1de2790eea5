"""Seeded synthetic fixtures (test infrastructure): thin re-export of ``stif_b200.synthetic`` so that the fixture
generator, the tests and ``bench.py`` all draw the SAME numpy PCG64 streams.  The generators themselves live in the
product package because ``bench.py``'s GPU arm must not import anything from ``oracle/``."""
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "stif-continuous-video-representation_b200",
                    "stif_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
from synthetic import (NET_SHAPES, OMEGA0, make_inputs, make_weights, smooth_inputs, weight_keys,  # noqa: E402,F401
                       weight_shapes)
