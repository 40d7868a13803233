"""ctypes / numpy mirrors of the PODs in include/rr.h (which mirror cl2.cl:77-151 and the host structs).

`CApi` drives any shared library that exports the rr.h entry points under a prefix: the product
(`librr_b200.so`, prefix ``rr_``) and — from tests/ and bench.py only — the CPU oracle (prefix ``orc_``).
"""
import ctypes as C

import numpy as np

# struct vertex cl2.cl:131-138 (48 B), struct triangle cl2.cl:148-151 (144 B)
VERTEX = np.dtype([("pos", "<f4", 4), ("normal", "<f4", 4), ("vt", "<f4", 2), ("object_id", "<u4"), ("vertex_col", "<u4")])
TRIANGLE = np.dtype([("vertices", VERTEX, 3)])
# struct obj_g_descriptor cl2.cl:99-124 (136 B -> 144 B)
OBJ_DESC = np.dtype([
    ("world_pos", "<f4", 4), ("world_rot_quat", "<f4", 4),
    ("old_world_pos_1", "<f4", 4), ("old_world_pos_2", "<f4", 4),
    ("old_world_rot_quat_1", "<f4", 4), ("old_world_rot_quat_2", "<f4", 4),
    ("scale", "<f4"), ("tid", "<u4"), ("rid", "<u4"), ("ssid", "<u4"), ("has_bump", "<u4"),
    ("specular", "<f4"), ("spec_mult", "<f4"), ("diffuse", "<f4"),
    ("buffer_offset", "<i4"), ("feature_flag", "<i4"), ("_pad", "<u4", 2)])
# struct light cl2.cl:77-87 (56 B -> 64 B)
LIGHT = np.dtype([("pos", "<f4", 4), ("col", "<f4", 4), ("shadow", "<u4"), ("brightness", "<f4"), ("radius", "<f4"),
                  ("diffuse", "<f4"), ("godray_intensity", "<f4"), ("is_static", "<i4"), ("_pad", "<u4", 2)])
assert VERTEX.itemsize == 48 and TRIANGLE.itemsize == 144 and OBJ_DESC.itemsize == 144 and LIGHT.itemsize == 64

FEATURE_SS_REFLECTIVE, FEATURE_TWO_SIDED, FEATURE_OUTLINE, FEATURE_IS_STATIC, FEATURE_NO_DYNAMIC_SHADOWS = 1, 2, 4, 8, 16

RR_OK, RR_ERR_INVALID, RR_ERR_CUDA, RR_ERR_OOM, RR_ERR_OVERFLOW, RR_ERR_PEER = 0, -1, -2, -3, -4, -5
RR_BUF_RGBA8, RR_BUF_SHADOW_DYNAMIC, RR_BUF_SHADOW_STATIC, RR_BUF_DEPTH, RR_BUF_IDS = range(5)


class Config(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("light_dim", C.c_int32), ("fov_const", C.c_float), ("hfov_deg", C.c_float),
                ("depth_icutoff", C.c_int32), ("ambient", C.c_float), ("ssao_rad", C.c_float), ("ssao_div", C.c_float), ("mip_bias", C.c_float),
                ("shadow_bias", C.c_float), ("shadow_exp", C.c_float), ("test_linear", C.c_int32), ("use_linear_rendering", C.c_int32),
                ("no_ssao", C.c_int32), ("device", C.c_int32), ("band_y0", C.c_int32), ("band_y1", C.c_int32), ("band_halo", C.c_int32),
                ("face_rank", C.c_int32), ("face_world", C.c_int32), ("max_fragments", C.c_uint32), ("max_cutdown", C.c_uint32),
                ("band_tile", C.c_int32), ("band_rank", C.c_int32), ("band_world", C.c_int32), ("face_interleave", C.c_int32),
                ("cluster_cull", C.c_int32)]

    @staticmethod
    def default(width=800, height=600, **kw):
        """Reference defaults: SSAO_RAD 5, no TEST_LINEAR ("profile B")."""
        c = Config(width=width, height=height, light_dim=1024, fov_const=0.0, hfov_deg=120.0, depth_icutoff=20, ambient=0.2, ssao_rad=5.0,
                   ssao_div=2.5, mip_bias=1.1, shadow_bias=50.0, shadow_exp=1.0, test_linear=0, use_linear_rendering=1, no_ssao=0, device=0,
                   band_y0=0, band_y1=0, band_halo=-1, face_rank=0, face_world=0, max_fragments=0, max_cutdown=0,
                   band_tile=0, band_rank=0, band_world=0, face_interleave=0, cluster_cull=0)
        for k, v in kw.items():
            if not hasattr(c, k):
                raise AttributeError(k)
            setattr(c, k, v)
        return c

    @staticmethod
    def profile_a(width, height, **kw):
        """main.cpp:80-85: -D depth_icutoff=20 -D AMBIENT=0.2f -D SSAO_RAD=2.f -D TEST_LINEAR, use_linear_rendering = 1."""
        return Config.default(width, height, ssao_rad=2.0, test_linear=1, use_linear_rendering=1, **kw)

    def copy(self, **kw):
        c = Config()
        C.memmove(C.byref(c), C.byref(self), C.sizeof(Config))
        for k, v in kw.items():
            setattr(c, k, v)
        return c


class MgpuHandle(C.Structure):
    """rr_mgpu_handle (include/rr.h): the IPC handles one context exports for the peer-memory exchange."""
    _fields_ = [("shadow", C.c_uint8 * 64), ("fb", C.c_uint8 * 64), ("ctrl", C.c_uint8 * 64), ("shadow_bytes", C.c_uint64), ("fb_bytes", C.c_uint64),
                ("device", C.c_int32), ("n_shadow", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("light_dim", C.c_int32), ("_pad", C.c_int32)]


class Timings(C.Structure):
    _fields_ = [("shadow_clear_ms", C.c_float), ("shadow_setup_ms", C.c_float), ("shadow_depth_ms", C.c_float), ("setup_ms", C.c_float),
                ("depth_ms", C.c_float), ("id_ms", C.c_float), ("shade_ms", C.c_float), ("frame_ms", C.c_float), ("n_cutdown", C.c_uint32),
                ("n_fragments", C.c_uint32), ("n_shadow_fragments", C.c_uint32), ("overflow", C.c_uint32), ("launches", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class RRError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


_P = C.c_void_p
_F4 = C.c_float * 4

# name -> (restype, argtypes); ctx is always the first argument
_SIGS = {
    "scene_alloc": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "scene_write_tris": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P]),
    "scene_write_objs": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P]),
    "scene_patch_obj": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32, _P]),
    "atlas_alloc": (C.c_int, [_P, C.c_uint32, _P, C.c_uint32, _P, C.c_uint32, C.c_uint32]),
    "scene_build_begin": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "scene_build_write_objs": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P]),
    "scene_build_write_tris": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P]),
    "scene_build_ready": (C.c_int, [_P]),
    "scene_build_commit": (C.c_int, [_P]),
    "atlas_upload": (C.c_int, [_P, C.c_uint32, _P, C.c_uint32, C.c_uint32, C.c_int]),
    "atlas_upload_batch": (C.c_int, [_P, C.c_uint32, _P, _P, _P, _P, C.c_int]),
    "atlas_fill_colour": (C.c_int, [_P, C.c_uint32, _F4, C.c_uint32, C.c_uint32]),
    "atlas_upload_mono": (C.c_int, [_P, C.c_uint32, _P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]),
    "atlas_write_raw": (C.c_int, [_P, _P, C.c_size_t]),
    "atlas_read_raw": (C.c_int, [_P, _P, C.c_size_t]),
    "lights_write": (C.c_int, [_P, _P, C.c_uint32]),
    "frame_shadows": (C.c_int, [_P, C.c_int]),
    "frame_draw": (C.c_int, [_P, _F4, _F4, _F4]),
    "swap_buffers": (C.c_int, [_P]),
    "post_pseudo_aa": (C.c_int, [_P]),
    "post_motion_blur": (C.c_int, [_P, C.c_float, C.c_float]),
    "post_godrays": (C.c_int, [_P]),
    "scene_read_objs": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P]),
    "sync": (C.c_int, [_P]),
    "read_depth": (C.c_int, [_P, _P]),
    "read_ids": (C.c_int, [_P, _P]),
    "read_rgba8": (C.c_int, [_P, _P]),
    "read_normals": (C.c_int, [_P, _P]),
    "read_shadow": (C.c_int, [_P, C.c_int, C.c_uint32, _P]),
    "read_fragments": (C.c_int, [_P, _P, C.c_uint32, C.POINTER(C.c_uint32)]),
    "read_cutdown": (C.c_int, [_P, _P, C.c_uint32, C.POINTER(C.c_uint32)]),
    "get_timings": (C.c_int, [_P, C.POINTER(Timings)]),
}


def _ptr(a):
    return a.ctypes.data_as(_P)


class CApi:
    """Thin object wrapper over one context of an rr.h-shaped library."""

    def __init__(self, lib, prefix, cfg, create_args=()):
        self._lib, self._prefix = lib, prefix
        self.cfg = cfg.copy()
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, prefix + name)
            fn.restype, fn.argtypes = res, args
        create = getattr(lib, prefix + "create")
        create.restype = _P
        create.argtypes = [C.POINTER(Config)] + [C.c_int] * len(create_args)
        getattr(lib, prefix + "destroy").argtypes = [_P]
        getattr(lib, prefix + "last_error").restype = C.c_char_p
        self._ctx = create(C.byref(self.cfg), *create_args)
        if not self._ctx:
            raise RRError(RR_ERR_CUDA, self.last_error())
        self.W, self.H, self.L = cfg.width, cfg.height, cfg.light_dim
        self.n_shadow = self.n_static = 0

    # -- plumbing
    def last_error(self):
        return getattr(self._lib, self._prefix + "last_error")().decode()

    def _call(self, name, *args):
        r = getattr(self._lib, self._prefix + name)(self._ctx, *args)
        if r != RR_OK:
            raise RRError(r, f"{self._prefix}{name}: {self.last_error()}")

    def close(self):
        if getattr(self, "_ctx", None):
            getattr(self._lib, self._prefix + "destroy")(self._ctx)
            self._ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- scene
    def scene_alloc(self, n_tris, n_objs):
        self._call("scene_alloc", n_tris, n_objs)

    def scene_write_tris(self, tris, first=0):
        tris = np.ascontiguousarray(tris, dtype=TRIANGLE)
        self._call("scene_write_tris", first, len(tris), _ptr(tris))

    def scene_write_objs(self, objs, first=0):
        objs = np.ascontiguousarray(objs, dtype=OBJ_DESC)
        self._call("scene_write_objs", first, len(objs), _ptr(objs))

    def scene_patch_obj(self, obj_id, byte_off, data):
        b = np.frombuffer(bytes(data), dtype=np.uint8)
        self._call("scene_patch_obj", obj_id, byte_off, len(b), _ptr(b))

    # -- asynchronous rebuild (object_context::build(async) + flip_buffers)
    def scene_build(self, tris, objs, commit=True):
        """upload a whole new scene into the back buffers while the current one keeps rendering; flip on commit"""
        tris = np.ascontiguousarray(tris, dtype=TRIANGLE)
        objs = np.ascontiguousarray(objs, dtype=OBJ_DESC)
        self._call("scene_build_begin", len(tris), len(objs))
        self._call("scene_build_write_objs", 0, len(objs), _ptr(objs))
        self._call("scene_build_write_tris", 0, len(tris), _ptr(tris))
        if commit:
            self._call("scene_build_commit")

    def scene_build_ready(self):
        return bool(getattr(self._lib, self._prefix + "scene_build_ready")(self._ctx))

    def scene_build_commit(self):
        self._call("scene_build_commit")

    # -- atlas
    def atlas_alloc(self, n_slices, nums, sizes, mipmap_start):
        nums = np.ascontiguousarray(nums, dtype=np.uint32)
        sizes = np.ascontiguousarray(sizes, dtype=np.uint32)
        self.atlas_slices = max(int(n_slices), 2)
        self._call("atlas_alloc", n_slices, _ptr(nums), len(nums), _ptr(sizes), len(sizes), mipmap_start)

    def atlas_upload(self, gpu_id, rgba, flip=1):
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        h, w = rgba.shape[:2]
        self._call("atlas_upload", gpu_id, _ptr(rgba), w, h, flip)

    def atlas_upload_batch(self, gpu_ids, images, flip=1):
        """all textures of an atlas build in one call (texture_context::alloc_gpu's upload loop)"""
        images = [np.ascontiguousarray(im, dtype=np.uint8) for im in images]
        n = len(images)
        ids = np.ascontiguousarray(gpu_ids, dtype=np.uint32)
        ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in images])
        w = np.array([im.shape[1] for im in images], dtype=np.uint32)
        h = np.array([im.shape[0] for im in images], dtype=np.uint32)
        self._call("atlas_upload_batch", n, _ptr(ids), C.cast(ptrs, _P), _ptr(w), _ptr(h), flip)

    def atlas_fill_colour(self, gpu_id, col_0_255, w, h):
        """texture::update_gpu_texture_col: a flat colour (0..255 units) into texture gpu_id and its mips"""
        self._call("atlas_fill_colour", gpu_id, _F4(*[float(x) for x in col_0_255]), w, h)

    def atlas_upload_mono(self, gpu_id, raw, w, h, flip=1):
        """texture::update_gpu_texture_mono: one byte per texel"""
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        self._call("atlas_upload_mono", gpu_id, _ptr(raw), raw.nbytes, w, h, flip)

    def atlas_read_raw(self):
        out = np.empty((self.atlas_slices, 2048, 2048, 4), dtype=np.uint8)
        self._call("atlas_read_raw", _ptr(out), out.nbytes)
        return out

    def atlas_write_raw(self, atlas):
        atlas = np.ascontiguousarray(atlas, dtype=np.uint8)
        self._call("atlas_write_raw", _ptr(atlas), atlas.nbytes)

    # -- lights
    def lights_write(self, lights):
        lights = np.ascontiguousarray(lights, dtype=LIGHT)
        self.n_shadow = int((lights["shadow"] == 1).sum())
        self.n_static = int(((lights["shadow"] != 0) & (lights["is_static"] != 0)).sum())
        self._call("lights_write", _ptr(lights), len(lights))

    # -- frame
    def frame_shadows(self, static_dirty=0):
        self._call("frame_shadows", int(static_dirty))

    def frame_draw(self, c_pos, c_rot, clear=(0, 0, 0, 0)):
        def f4(v):
            v = list(v) + [0.0] * (4 - len(v))
            return _F4(*[float(x) for x in v[:4]])
        self._call("frame_draw", f4(c_pos), f4(c_rot), f4(clear))

    def post_pseudo_aa(self):
        """engine::do_pseudo_aa: edge smoothing of the frame just drawn (before swap_buffers)"""
        self._call("post_pseudo_aa")

    def post_motion_blur(self, strength=1.0, camera_contribution=1.0):
        """engine::do_motion_blur: blur along each pixel's screen-space motion since the previous frame (before swap_buffers)"""
        self._call("post_motion_blur", float(strength), float(camera_contribution))

    def post_godrays(self):
        """engine::draw_godrays: screen-space light shafts of the lights with godray_intensity > 0 (before swap_buffers)"""
        self._call("post_godrays")

    def scene_read_objs(self, first, count):
        """the device copy of the descriptors (do_motion_blur advances their motion history)"""
        out = np.zeros(count, dtype=OBJ_DESC)
        if count:
            self._call("scene_read_objs", first, count, _ptr(out))
        return out

    def swap_buffers(self):
        self._call("swap_buffers")

    def sync(self):
        self._call("sync")

    # -- read-back
    def read_depth(self):
        out = np.empty((self.H, self.W), dtype=np.uint32)
        self._call("read_depth", _ptr(out))
        return out

    def read_ids(self):
        out = np.empty((self.H, self.W), dtype=np.uint32)
        self._call("read_ids", _ptr(out))
        return out

    def read_rgba8(self):
        out = np.empty((self.H, self.W, 4), dtype=np.uint8)
        self._call("read_rgba8", _ptr(out))
        return out

    def read_normals(self):
        out = np.empty((self.H, self.W, 2), dtype=np.uint16)
        self._call("read_normals", _ptr(out))
        return out

    def read_shadow(self, is_static, slab):
        out = np.empty((6, self.L, self.L), dtype=np.uint32)
        self._call("read_shadow", int(is_static), slab, _ptr(out))
        return out

    def read_fragments(self):
        n = C.c_uint32(0)
        self._call("read_fragments", None, 0, C.byref(n))
        out = np.empty((n.value, 5), dtype=np.uint32)
        if n.value:
            self._call("read_fragments", _ptr(out), n.value, C.byref(n))
        return out

    def read_cutdown(self):
        n = C.c_uint32(0)
        self._call("read_cutdown", None, 0, C.byref(n))
        out = np.empty((n.value, 3, 4), dtype=np.float32)
        if n.value:
            self._call("read_cutdown", _ptr(out), n.value, C.byref(n))
        return out

    def timings(self):
        t = Timings()
        self._call("get_timings", C.byref(t))
        return t.as_dict()
