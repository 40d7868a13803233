"""ctypes binding of the product library (include/rr.h). No fallback: a missing library is an error."""
import ctypes as C
import os

import numpy as np

from ._abi import CApi, Config, MgpuHandle, RRError, RR_OK, _F4, _P, _ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def library_path():
    # RR_LIB: an alternative build of the same library (A/B runs of build-time knobs); the default is the in-tree product
    return os.environ.get("RR_LIB") or os.path.join(_HERE, "librr_b200.so")


def load_library():
    """dlopen librr_b200.so. Raises if it has not been built (python -m openclrenderer_b200._build / __graft_entry__.build())."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise RRError(-2, f"{path} is missing: build it with `python -m openclrenderer_b200._build` (nvcc, sm_100a). "
                              "There is no CPU implementation to fall back to.")
        _LIB = C.CDLL(path)
        _LIB.rr_fov_const_from_hfov.restype = C.c_float
        _LIB.rr_fov_const_from_hfov.argtypes = [C.c_float, C.c_float]
        _LIB.rr_version.restype = C.c_char_p
        _LIB.rr_default_config.argtypes = [C.POINTER(Config)]
        _LIB.rr_bind_external.restype = C.c_int
        _LIB.rr_bind_external.argtypes = [_P, C.c_int, _P, C.c_size_t]
        _LIB.rr_device_ptr.restype = _P
        _LIB.rr_device_ptr.argtypes = [_P, C.c_int]
        _LIB.rr_stream.restype = _P
        _LIB.rr_stream.argtypes = [_P]
        _LIB.rr_shadow_stream.restype = _P
        _LIB.rr_shadow_stream.argtypes = [_P]
        _LIB.rr_shadows_done.restype = C.c_int
        _LIB.rr_shadows_done.argtypes = [_P]
        _LIB.rr_frame_e2e.restype = C.c_int
        _LIB.rr_frame_e2e.argtypes = [_P, _F4, _F4, _F4, C.c_int, _P]
        _LIB.rr_set_profiling.restype = C.c_int
        _LIB.rr_set_profiling.argtypes = [_P, C.c_int]
        _LIB.rr_set_pipeline_depth.restype = C.c_int
        _LIB.rr_set_pipeline_depth.argtypes = [_P, C.c_int]
        _LIB.rr_host_alloc.restype = _P
        _LIB.rr_host_alloc.argtypes = [C.c_size_t]
        _LIB.rr_host_free.argtypes = [_P]
        _LIB.rr_microbench_atomic_min.restype = C.c_int
        _LIB.rr_microbench_atomic_min.argtypes = [_P, C.c_size_t, C.c_uint64, C.POINTER(C.c_float)]
        _LIB.rr_mgpu_export.restype = C.c_int
        _LIB.rr_mgpu_export.argtypes = [_P, C.POINTER(MgpuHandle)]
        _LIB.rr_mgpu_connect.restype = C.c_int
        _LIB.rr_mgpu_connect.argtypes = [_P, C.c_int, C.c_int, C.POINTER(MgpuHandle)]
        _LIB.rr_mgpu_connect_local.restype = C.c_int
        _LIB.rr_mgpu_connect_local.argtypes = [C.POINTER(_P), C.c_int]
        _LIB.rr_mgpu_disconnect.restype = C.c_int
        _LIB.rr_mgpu_disconnect.argtypes = [_P]
        _LIB.rr_mgpu_pushed_bytes.restype = C.c_int
        _LIB.rr_mgpu_pushed_bytes.argtypes = [_P, C.POINTER(C.c_uint64)]
        _LIB.rr_mgpu_set_readback.restype = C.c_int
        _LIB.rr_mgpu_set_readback.argtypes = [_P, C.c_int]
        _LIB.rr_host_register.restype = C.c_int
        _LIB.rr_host_register.argtypes = [_P, C.c_size_t]
        _LIB.rr_host_unregister.restype = C.c_int
        _LIB.rr_host_unregister.argtypes = [_P]
        _LIB.rr_set_readback_tiles.restype = C.c_int
        _LIB.rr_set_readback_tiles.argtypes = [_P, C.c_int]
        _LIB.rr_readback_tile_bytes.restype = C.c_int
        _LIB.rr_readback_tile_bytes.argtypes = [_P, C.POINTER(C.c_uint64)]
        _LIB.rr_microbench_copy.restype = C.c_int
        _LIB.rr_microbench_copy.argtypes = [_P, C.c_size_t, C.POINTER(C.c_float)]
    return _LIB


class _PinnedBlock:
    """owner of one rr_host_alloc block; numpy keeps it alive as the base of every view and it frees the block when collected"""

    def __init__(self, lib, ptr, nbytes):
        self._lib, self._ptr = lib, ptr
        self.__array_interface__ = {"data": (ptr, False), "shape": (nbytes,), "typestr": "|u1", "version": 3}

    def __del__(self):
        try:
            if self._ptr:
                self._lib.rr_host_free(self._ptr)
                self._ptr = None
        except Exception:
            pass


def host_alloc(shape, dtype=np.uint8):
    """numpy array over page-locked memory from rr_host_alloc (rr_host_free runs when the last view of it is collected)."""
    lib = load_library()
    n = max(1, int(np.prod(shape)) * np.dtype(dtype).itemsize)
    p = lib.rr_host_alloc(n)
    if not p:
        raise RRError(-3, "rr_host_alloc failed")
    return np.asarray(_PinnedBlock(lib, p, n)).view(dtype)[:int(np.prod(shape))].reshape(shape)


def fov_const_from_hfov(hfov_deg, width):
    return float(load_library().rr_fov_const_from_hfov(hfov_deg, width))


class Renderer(CApi):
    """One rr_ctx: the CUDA rasteriser on one B200."""

    def __init__(self, cfg):
        super().__init__(load_library(), "rr_", cfg)

    def bind_external(self, which, device_ptr, nbytes):
        r = self._lib.rr_bind_external(self._ctx, which, device_ptr, nbytes)
        if r != RR_OK:
            raise RRError(r, self.last_error())

    def device_ptr(self, which):
        return self._lib.rr_device_ptr(self._ctx, which)

    def stream(self):
        return self._lib.rr_stream(self._ctx)

    def shadow_stream(self):
        return self._lib.rr_shadow_stream(self._ctx)

    def shadows_done(self):
        r = self._lib.rr_shadows_done(self._ctx)
        if r != RR_OK:
            raise RRError(r, self.last_error())

    def set_profiling(self, on=True):
        """per-stage CUDA events for timings() (off by default: they cost frame time, see rr.h)"""
        self._lib.rr_set_profiling(self._ctx, int(bool(on)))

    def set_pipeline_depth(self, depth):
        r = self._lib.rr_set_pipeline_depth(self._ctx, int(depth))
        if r != RR_OK:
            raise RRError(r, self.last_error())

    def set_readback_tiles(self, on):
        """dirty-tile read-back of frame_e2e (include/rr.h rr_set_readback_tiles)"""
        r = self._lib.rr_set_readback_tiles(self._ctx, int(bool(on)))
        if r != RR_OK:
            raise RRError(r, self.last_error())

    def readback_tile_bytes(self):
        """bytes the tile path has stored into host buffers since the last call"""
        n = C.c_uint64(0)
        r = self._lib.rr_readback_tile_bytes(self._ctx, C.byref(n))
        if r != RR_OK:
            raise RRError(r, self.last_error())
        return int(n.value)

    def frame_e2e(self, c_pos, c_rot, clear, with_shadows, host_rgba8):
        def f4(v):
            v = list(v) + [0.0] * (4 - len(v))
            return _F4(*[float(x) for x in v[:4]])
        r = self._lib.rr_frame_e2e(self._ctx, f4(c_pos), f4(c_rot), f4(clear), int(with_shadows), _ptr(host_rgba8))
        if r != RR_OK:
            raise RRError(r, self.last_error())

    # -- multi-GPU exchange over peer memory (include/rr.h rr_mgpu_*)
    def mgpu_export(self):
        """-> bytes of this context's rr_mgpu_handle (to be exchanged between the processes, any transport)"""
        h = MgpuHandle()
        r = self._lib.rr_mgpu_export(self._ctx, C.byref(h))
        if r != RR_OK:
            raise RRError(r, self.last_error())
        return bytes(h)

    def mgpu_connect(self, rank, world, handles):
        """handles: list of `world` byte strings from mgpu_export, in rank order"""
        arr = (MgpuHandle * world)()
        for k, b in enumerate(handles):
            C.memmove(C.byref(arr[k]), bytes(b), C.sizeof(MgpuHandle))
        r = self._lib.rr_mgpu_connect(self._ctx, rank, world, arr)
        if r != RR_OK:
            raise RRError(r, self.last_error())

    def mgpu_set_readback(self, distributed):
        r = self._lib.rr_mgpu_set_readback(self._ctx, int(distributed))
        if r != RR_OK:
            raise RRError(r, self.last_error())

    def mgpu_pushed_bytes(self):
        """bytes of cubemap faces this context has stored into its peers' memory since the last call"""
        n = C.c_uint64(0)
        r = self._lib.rr_mgpu_pushed_bytes(self._ctx, C.byref(n))
        if r != RR_OK:
            raise RRError(r, self.last_error())
        return int(n.value)

    def mgpu_disconnect(self):
        self._lib.rr_mgpu_disconnect(self._ctx)

    def microbench_atomic_min(self, footprint_bytes, n_ops):
        ms = C.c_float(0)
        r = self._lib.rr_microbench_atomic_min(self._ctx, footprint_bytes, n_ops, C.byref(ms))
        if r != RR_OK:
            raise RRError(r, self.last_error())
        return ms.value

    def microbench_copy(self, nbytes):
        ms = C.c_float(0)
        r = self._lib.rr_microbench_copy(self._ctx, nbytes, C.byref(ms))
        if r != RR_OK:
            raise RRError(r, self.last_error())
        return ms.value


def mgpu_connect_local(renderers):
    """Wire contexts that live in this process (one per GPU, or several on one GPU in the tests): renderers[k] becomes rank k."""
    lib = load_library()
    arr = (_P * len(renderers))(*[r._ctx for r in renderers])
    r = lib.rr_mgpu_connect_local(arr, len(renderers))
    if r != RR_OK:
        raise RRError(r, lib.rr_last_error().decode())


def host_register(arr):
    """page-lock caller-owned memory (e.g. a numpy view of a multiprocessing.shared_memory block) for direct DMA"""
    lib = load_library()
    r = lib.rr_host_register(_ptr(arr), arr.nbytes)
    if r != RR_OK:
        raise RRError(r, lib.rr_last_error().decode())


def host_unregister(arr):
    load_library().rr_host_unregister(_ptr(arr))
