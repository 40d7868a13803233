"""ctypes binding of oracle/liboracle.so (CPU restatement of cl2.cl). TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

from openclrenderer_b200._abi import CApi, _P, _ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def load_oracle():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            from oracle import build as obuild
            obuild.build_oracle()
        _LIB = C.CDLL(path)
        L = _LIB
        L.orc_fov_const_from_hfov.restype = C.c_float
        L.orc_fov_const_from_hfov.argtypes = [C.c_float, C.c_float]
        for n in ("orc_depth_samples", "orc_shadow_samples", "orc_saturation_events"):
            getattr(L, n).restype = C.c_uint64
            getattr(L, n).argtypes = [_P]
        L.orc_threads.restype = C.c_int
        L.orc_threads.argtypes = [_P]
        L.orc_read_colour_f32.argtypes = [_P, _P]
        L.orc_write_shadow.argtypes = [_P, C.c_int, C.c_uint32, _P]
        L.orc_stage_setup.argtypes = [_P, C.c_float * 4, C.c_float * 4]
        L.orc_stage_depth.argtypes = [_P]
        L.orc_stage_ids.argtypes = [_P]
        L.orc_unit_wang_hash.restype = C.c_uint32
        L.orc_unit_wang_hash.argtypes = [C.c_uint32]
        L.orc_unit_xorshift.restype = C.c_uint32
        L.orc_unit_xorshift.argtypes = [C.c_uint32]
        L.orc_unit_log2_approx.restype = C.c_float
        L.orc_unit_log2_approx.argtypes = [C.c_float]
        L.orc_unit_rational_acos.restype = C.c_float
        L.orc_unit_rational_acos.argtypes = [C.c_float]
        L.orc_unit_point_in_tri.restype = C.c_int
        L.orc_unit_point_in_tri.argtypes = [C.c_float, C.c_float, _P]
        L.orc_unit_cubeface.restype = C.c_int
        L.orc_unit_cubeface.argtypes = [_P, _P]
        L.orc_unit_scan.restype = C.c_int
        L.orc_unit_scan.argtypes = [_P, C.c_int, C.c_uint32, _P, C.c_int]
        L.orc_unit_clip_project.restype = C.c_int
        L.orc_unit_clip_project.argtypes = [_P, C.c_int, C.c_float, C.c_float, C.c_float, _P]
    return _LIB


class Oracle(CApi):
    """Same interface as openclrenderer_b200.Renderer, computed on the CPU by the restatement."""

    def __init__(self, cfg, threads=1):
        super().__init__(load_oracle(), "orc_", cfg, create_args=(threads,))

    @property
    def depth_samples(self):
        return int(self._lib.orc_depth_samples(self._ctx))

    @property
    def shadow_samples(self):
        return int(self._lib.orc_shadow_samples(self._ctx))

    @property
    def saturation_events(self):
        return int(self._lib.orc_saturation_events(self._ctx))

    @property
    def threads(self):
        return int(self._lib.orc_threads(self._ctx))

    def read_colour_f32(self):
        out = np.empty((self.H, self.W, 4), dtype=np.float32)
        self._lib.orc_read_colour_f32(self._ctx, _ptr(out))
        return out

    def write_shadow(self, is_static, slab, data):
        data = np.ascontiguousarray(data, dtype=np.uint32)
        self._lib.orc_write_shadow(self._ctx, int(is_static), slab, _ptr(data))


def _v(a, n):
    return np.ascontiguousarray(a, dtype=np.float32).reshape(n)


def unit_rot(p, cpos, crot, back=False):
    L = load_oracle()
    out = np.zeros(3, np.float32)
    (L.orc_unit_back_rot if back else L.orc_unit_rot)(_ptr(_v(p, 3)), _ptr(_v(cpos, 3)), _ptr(_v(crot, 3)), _ptr(out))
    return out


def unit_rot_quat(p, q, back=False):
    L = load_oracle()
    out = np.zeros(3, np.float32)
    (L.orc_unit_back_rot_quat if back else L.orc_unit_rot_quat)(_ptr(_v(p, 3)), _ptr(_v(q, 4)), _ptr(out))
    return out


def unit_scan(mm, op_size, distance, max_out=1024):
    L = load_oracle()
    out = np.zeros((max_out, 2), np.int32)
    n = L.orc_unit_scan(_ptr(_v(mm, 4)), op_size, distance, _ptr(out), max_out)
    return out[:min(n, max_out)]


def unit_clip_project(pr, icut, w, h, fov):
    L = load_oracle()
    out = np.zeros(18, np.float32)
    n = L.orc_unit_clip_project(_ptr(_v(pr, 9)), icut, w, h, fov, _ptr(out))
    return out.reshape(2, 3, 3)[:n]
