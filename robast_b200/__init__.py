"""robast_b200 — B200-native non-sequential ray tracer behind ROBAST's API.

`import robast_b200 as ROOT` gives the class names a PyROOT + ROBAST script uses on the
TraceNonSequential path (AOpticsManager, ARayShooter, ALens, TGeoBBox, ...).  All tracing runs in
the CUDA library `librobast_b200.so` through the C ABI in include/robast_b200.h; there is no CPU
fallback: importing fails loudly if the native library has not been built (run
`python -c "import __graft_entry__ as g; g.build()"`), and tracing fails loudly without a GPU.
"""
import ctypes
import os

_here = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_here, "librobast_b200.so")
if not os.path.exists(LIB_PATH):
    raise ImportError(
        "robast_b200: native library %s is missing — build it with __graft_entry__.build() "
        "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)

# the C ABI, for callers that want plain pointers (bench.py, tests, other FFIs)
lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)

from ._robast import *  # noqa: F401,F403  (host-side mirror classes)
from . import _robast as _ext
from .cabi import *  # noqa: F401,F403

gRandom = _ext.gRandom
