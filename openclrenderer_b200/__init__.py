"""openclrenderer_b200 — B200-native (sm_100a) drop-in for OpenCLRenderer's per-frame draw path.

The product is the C-ABI shared library `librr_b200.so` (include/rr.h) built from csrc/*.cu; this package is the
Python-side binding used by the tests and the bench, plus the host-side scene helpers (OBJ loader, atlas planner,
synthetic scenes) that mirror the reference's host classes. There is no CPU path: importing works anywhere,
creating a `Renderer` needs the CUDA library and a B200.
"""
from ._abi import (Config, Timings, RRError, TRIANGLE, VERTEX, OBJ_DESC, LIGHT, FEATURE_TWO_SIDED, FEATURE_IS_STATIC,
                   FEATURE_NO_DYNAMIC_SHADOWS, FEATURE_SS_REFLECTIVE, FEATURE_OUTLINE)
from .rr import Renderer, load_library, library_path

__all__ = ["Config", "Timings", "RRError", "Renderer", "load_library", "library_path", "TRIANGLE", "VERTEX", "OBJ_DESC", "LIGHT",
           "FEATURE_TWO_SIDED", "FEATURE_IS_STATIC", "FEATURE_NO_DYNAMIC_SHADOWS", "FEATURE_SS_REFLECTIVE", "FEATURE_OUTLINE"]
