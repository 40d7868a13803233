"""TEST INFRASTRUCTURE ONLY — runs the reference's own, unmodified cl2.cl through the NVIDIA OpenCL ICD that ships
with the driver on the GPU box (libnvidia-opencl.so.1, reached through the CUDA toolkit's libOpenCL.so.1 loader with
OCL_ICD_FILENAMES; no OpenCL headers, no pyopencl: plain ctypes).

It is the strongest pin the oracle has: the reference kernels themselves (`prearrange`, `kernel1`, `kernel2`,
`kernel3`, `prearrange_realtime_shadowing`, `kernel1_realtime_shadowing`, `update_gpu_tex`, `generate_mips`,
`generate_mip_mips`), launched the way engine.cpp / texture.cpp launch them (argument order, global/local sizes,
build options), on the same inputs as the oracle and the CUDA product.

  RefCL(cfg, mode="shipped")  build options exactly as ocl.h:227-236 + main.cpp:80-85 (-cl-fast-relaxed-math …)
  RefCL(cfg, mode="pinned")   same source, but native_divide/native_recip/native_sin/native_cos/fast_* are mapped to
                              their exact counterparts with -D and fast-relaxed-math is dropped: the arithmetic pinned
                              in SURVEY.md §8c as far as a real OpenCL compiler can be asked to provide it
                              (FP_CONTRACT stays ON in the source, so products may still be fused).

The source travels to the GPU box as oracle/_ref/cl2.cl.gz (git-ignored; written by oracle/build_ref.py from
/root/reference/cl2.cl). Same Python interface as openclrenderer_b200.Renderer / oracle.binding.Oracle.
"""
import ctypes as C
import gzip
import os
import time

import numpy as np

from openclrenderer_b200._abi import TRIANGLE, OBJ_DESC, LIGHT

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC_GZ = os.path.join(_HERE, "_ref", "cl2.cl.gz")

P = C.c_void_p
CL_MEM_READ_WRITE, CL_MEM_READ_ONLY, CL_MEM_COPY_HOST_PTR = 1, 4, 32
CL_R, CL_RGBA = 0x10B0, 0x10B5
CL_UNORM_INT8, CL_UNSIGNED_INT32, CL_FLOAT = 0x10D2, 0x10DC, 0x10DE
CL_DEVICE_TYPE_GPU = 4
CL_PROGRAM_BUILD_LOG = 0x1183
CL_QUEUE_PROFILING_ENABLE = 2
CL_PROFILING_COMMAND_START, CL_PROFILING_COMMAND_END = 0x1282, 0x1283


PINNED_PREFIX = b"""
#pragma OPENCL FP_CONTRACT OFF
#define RRO __attribute__((overloadable))
float RRO rr_pin_dot(float2 a, float2 b) { return a.x*b.x + a.y*b.y; }
float RRO rr_pin_dot(float3 a, float3 b) { return a.x*b.x + a.y*b.y + a.z*b.z; }
float RRO rr_pin_dot(float4 a, float4 b) { return a.x*b.x + a.y*b.y + a.z*b.z + a.w*b.w; }
float3 RRO rr_pin_cross(float3 a, float3 b) { return (float3)(a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x); }
float4 RRO rr_pin_cross(float4 a, float4 b) { return (float4)(a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x, 0.0f); }
float RRO rr_pin_dot(float a, float b) { return a*b; }
float RRO rr_pin_length(float a) { return fabs(a); }
float RRO rr_pin_normalize(float a) { return a / fabs(a); }
float RRO rr_pin_length(float2 a) { return sqrt(rr_pin_dot(a, a)); }
float RRO rr_pin_length(float3 a) { return sqrt(rr_pin_dot(a, a)); }
float RRO rr_pin_length(float4 a) { return sqrt(rr_pin_dot(a, a)); }
float2 RRO rr_pin_normalize(float2 a) { return a / sqrt(rr_pin_dot(a, a)); }
float3 RRO rr_pin_normalize(float3 a) { return a / sqrt(rr_pin_dot(a, a)); }
float4 RRO rr_pin_normalize(float4 a) { return a / sqrt(rr_pin_dot(a, a)); }
#line 1
"""


class ImageFormat(C.Structure):
    _fields_ = [("order", C.c_uint), ("type", C.c_uint)]


class CLError(RuntimeError):
    pass


_cl = None


def available():
    try:
        _load()
        n = C.c_uint(0)
        return _cl.clGetPlatformIDs(0, None, C.byref(n)) == 0 and n.value > 0 and os.path.exists(SRC_GZ)
    except Exception:
        return False


def _load():
    global _cl
    if _cl is None:
        os.environ.setdefault("OCL_ICD_FILENAMES", "libnvidia-opencl.so.1")
        _cl = C.CDLL("libOpenCL.so.1")
        for name in ("clCreateContext", "clCreateCommandQueue", "clCreateProgramWithSource", "clCreateKernel", "clCreateBuffer", "clCreateImage2D"):
            getattr(_cl, name).restype = P
        S, U = C.c_size_t, C.c_uint
        _cl.clEnqueueWriteImage.argtypes = [P, P, U, C.POINTER(S), C.POINTER(S), S, S, P, U, P, P]
        _cl.clEnqueueReadImage.argtypes = [P, P, U, C.POINTER(S), C.POINTER(S), S, S, P, U, P, P]
        _cl.clEnqueueWriteBuffer.argtypes = [P, P, U, S, S, P, U, P, P]
        _cl.clEnqueueReadBuffer.argtypes = [P, P, U, S, S, P, U, P, P]
        _cl.clEnqueueCopyBuffer.argtypes = [P, P, P, S, S, S, U, P, P]
        _cl.clEnqueueFillBuffer.argtypes = [P, P, P, S, S, S, U, P, P]
        _cl.clEnqueueNDRangeKernel.argtypes = [P, P, U, C.POINTER(S), C.POINTER(S), C.POINTER(S), U, P, P]
        _cl.clSetKernelArg.argtypes = [P, U, S, P]
        _cl.clCreateImage2D.argtypes = [P, C.c_uint64, C.POINTER(ImageFormat), S, S, S, P, C.POINTER(C.c_int)]
        _cl.clCreateBuffer.argtypes = [P, C.c_uint64, S, P, C.POINTER(C.c_int)]
        _cl.clGetProgramBuildInfo.argtypes = [P, P, U, S, P, C.POINTER(S)]
        _cl.clGetEventProfilingInfo.argtypes = [P, U, S, P, C.POINTER(S)]
        _cl.clFinish.argtypes = [P]
        _cl.clReleaseMemObject.argtypes = [P]
        _cl.clReleaseEvent.argtypes = [P]
    return _cl


def _chk(r, what):
    if r != 0:
        raise CLError(f"{what} failed: {r}")


class RefCL:
    def __init__(self, cfg, mode="shipped", fov=None, profile=False):
        cl = _load()
        self.cl, self.cfg, self.mode = cl, cfg, mode
        self.W, self.H, self.L = cfg.width, cfg.height, cfg.light_dim
        n = C.c_uint(0)
        plats = (P * 4)()
        _chk(cl.clGetPlatformIDs(4, plats, C.byref(n)), "clGetPlatformIDs")
        devs = (P * 8)()
        _chk(cl.clGetDeviceIDs(P(plats[0]), C.c_uint64(CL_DEVICE_TYPE_GPU), 8, devs, C.byref(n)), "clGetDeviceIDs")
        self.dev = P(devs[cfg.device if cfg.device < n.value else 0])
        err = C.c_int(0)
        self.ctx = P(cl.clCreateContext(None, 1, C.byref(self.dev), None, None, C.byref(err)))
        _chk(err.value, "clCreateContext")
        self.q = P(cl.clCreateCommandQueue(self.ctx, self.dev, C.c_uint64(CL_QUEUE_PROFILING_ENABLE if profile else 0), C.byref(err)))
        _chk(err.value, "clCreateCommandQueue")
        self.profile = profile
        src = gzip.open(SRC_GZ, "rb").read()
        prefix = b""
        if mode.startswith("pinned"):
            # The pinned arithmetic of SURVEY.md §8c, asked of a real OpenCL compiler. Two deviations from "as shipped":
            #  (1) cl2.cl:3 `#pragma OPENCL FP_CONTRACT ON` is flipped to OFF in memory (products are not fused),
            #  (2) a prefix string supplies the geometric builtins in their pinned form (left-to-right sums, v / sqrt(dot))
            #      and -D remaps route the reference's calls to them. The reference's own code is otherwise untouched.
            assert src.count(b"#pragma OPENCL FP_CONTRACT ON") == 1
            src = src.replace(b"#pragma OPENCL FP_CONTRACT ON", b"#pragma OPENCL FP_CONTRACT OFF")
            prefix = PINNED_PREFIX
        src = prefix + src
        sp, ln = C.c_char_p(src), C.c_size_t(len(src))
        self.prog = P(cl.clCreateProgramWithSource(self.ctx, 1, C.byref(sp), C.byref(ln), C.byref(err)))
        _chk(err.value, "clCreateProgramWithSource")
        from openclrenderer_b200.scene import fov_for
        self.fov = fov if fov is not None else fov_for(cfg)
        # ocl.h:227-236 (+ engine.cpp:474-477 FOV literal) + main.cpp:80-85 extras as carried by cfg
        opts = ["-cl-single-precision-constant", f"-D SCREENWIDTH={self.W}", f"-D SCREENHEIGHT={self.H}", f"-D LIGHTBUFFERDIM={self.L}",
                f"-D SHADOWBIAS={int(cfg.shadow_bias)}", "-D FOV_CONST=%ff" % self.fov, f"-D depth_icutoff={cfg.depth_icutoff}",
                "-D AMBIENT=%rf" % float(np.float32(cfg.ambient)), "-D SSAO_RAD=%rf" % float(np.float32(cfg.ssao_rad))]
        if cfg.test_linear:
            opts.append("-D TEST_LINEAR")
        if cfg.no_ssao:
            opts.append("-D NO_SSAO")
        if mode == "shipped":
            opts = ["-cl-fast-relaxed-math", "-cl-no-signed-zeros", "-cl-denorms-are-zero"] + opts
        else:
            opts += ["-cl-fp32-correctly-rounded-divide-sqrt", "-Dnative_divide(a,b)=((a)/(b))", "-Dnative_recip(a)=(1.0f/(a))", "-Dnative_sin=sin",
                     "-Dnative_cos=cos", "-Dnative_exp=exp", "-Dnative_powr=powr", "-Ddot=rr_pin_dot", "-Dcross=rr_pin_cross",
                     "-Dlength=rr_pin_length", "-Dfast_length=rr_pin_length", "-Dnormalize=rr_pin_normalize", "-Dfast_normalize=rr_pin_normalize"]
        if mode == "pinned_noopt":
            opts.append("-cl-opt-disable")
        self.options = " ".join(opts)
        t0 = time.time()
        r = cl.clBuildProgram(self.prog, 1, C.byref(self.dev), self.options.encode(), None, None)
        self.build_s = time.time() - t0
        if r != 0:
            sz = C.c_size_t(0)
            cl.clGetProgramBuildInfo(self.prog, self.dev, CL_PROGRAM_BUILD_LOG, 0, None, C.byref(sz))
            log = C.create_string_buffer(sz.value + 1)
            cl.clGetProgramBuildInfo(self.prog, self.dev, CL_PROGRAM_BUILD_LOG, sz.value, C.cast(log, P), None)
            raise CLError("clBuildProgram failed %d:\n%s" % (r, log.value.decode(errors="replace")[-4000:]))
        self.k = {}
        for name in ("prearrange", "kernel1", "kernel2", "kernel3", "prearrange_realtime_shadowing", "kernel1_realtime_shadowing",
                     "update_gpu_tex", "generate_mips", "generate_mip_mips"):
            self.k[name] = P(cl.clCreateKernel(self.prog, name.encode(), C.byref(err)))
            _chk(err.value, "clCreateKernel " + name)
        Pn = self.W * self.H
        ones = np.full(Pn, 0xFFFFFFFF, np.uint32)
        self.depth = [self._buf(ones), self._buf(ones)]                     # object_context.cpp:43-52
        self.cur = 0
        self.id_img = self._image(CL_R, CL_UNSIGNED_INT32, self.W, self.H)  # object_context.cpp:36-38
        self.screen = [self._image(CL_RGBA, CL_FLOAT, self.W, self.H), self._image(CL_RGBA, CL_FLOAT, self.W, self.H)]
        self.normals = self._buf(np.zeros(Pn * 2, np.uint16))
        self.g_tid_buf = self._buf(nbytes=40 << 20 if Pn < (1 << 22) else 400 << 20)   # engine.cpp:601 is 10 Mi uint; larger for 4K+ scenes (q8)
        self.g_tid_buf_max_len = self._buf(np.array([(40 << 20) // 4], np.uint32))
        self.cnt_main = self._buf(np.zeros(1, np.uint32))
        self.cnt_light = self._buf(np.zeros(1, np.uint32))
        self.cut_num = self._buf(np.zeros(1, np.uint32))
        self.dummy_buf = self._buf(np.zeros(4, np.uint32))
        self.dummy_img = self._image(CL_RGBA, CL_FLOAT, 4, 4)
        self.scratch_img = None                                             # third full-size colour image (post passes: see post_pseudo_aa)
        self.c_pos = self.c_rot = self.c_pos_old = self.c_rot_old = self._f4((0, 0, 0, 0))
        self.n_tris = 0
        self.lights = np.zeros(0, LIGHT)
        self.n_shadow = self.n_static = 0
        self.frame_id = 0
        self.kernel_ms = {}
        self.shadow_counts = []
        self.steady_count = None
        self.steady = False
        self.steady_shadow_counts = []

    # ---- helpers
    def _buf(self, arr=None, nbytes=None, flags=CL_MEM_READ_WRITE):
        err = C.c_int(0)
        if arr is not None:
            arr = np.ascontiguousarray(arr)
            b = self.cl.clCreateBuffer(self.ctx, flags | CL_MEM_COPY_HOST_PTR, max(arr.nbytes, 4), arr.ctypes.data_as(P), C.byref(err))
        else:
            b = self.cl.clCreateBuffer(self.ctx, flags, max(nbytes, 4), None, C.byref(err))
        _chk(err.value, "clCreateBuffer")
        return P(b)

    def _image(self, order, typ, w, h, flags=CL_MEM_READ_WRITE):
        err = C.c_int(0)
        fmt = ImageFormat(order, typ)
        im = self.cl.clCreateImage2D(self.ctx, flags, C.byref(fmt), w, h, 0, None, C.byref(err))
        _chk(err.value, "clCreateImage2D")
        return P(im)

    def _write(self, buf, arr, offset=0):
        arr = np.ascontiguousarray(arr)
        _chk(self.cl.clEnqueueWriteBuffer(self.q, buf, 1, offset, arr.nbytes, arr.ctypes.data_as(P), 0, None, None), "write")

    def _read(self, buf, arr, offset=0):
        _chk(self.cl.clEnqueueReadBuffer(self.q, buf, 1, offset, arr.nbytes, arr.ctypes.data_as(P), 0, None, None), "read")
        return arr

    def _fill(self, buf, nbytes, pattern=0xFFFFFFFF):
        pat = C.c_uint(pattern)
        _chk(self.cl.clEnqueueFillBuffer(self.q, buf, C.cast(C.byref(pat), P), 4, 0, nbytes, 0, None, None), "fill")

    def _run(self, name, args, gws, lws):
        """run_kernel_with_string, engine.hpp:646-676: every arg set every call, gws rounded up to a multiple of lws."""
        k = self.k[name]
        keep = []
        for i, a in enumerate(args):
            if isinstance(a, P):
                v = P(a.value)
                r = self.cl.clSetKernelArg(k, i, 8, C.cast(C.byref(v), P))
                keep.append(v)
            else:
                a = np.ascontiguousarray(a)
                r = self.cl.clSetKernelArg(k, i, a.nbytes, a.ctypes.data_as(P))
                keep.append(a)
            _chk(r, f"clSetKernelArg {name}[{i}]")
        g = [((int(x) + l - 1) // l) * l for x, l in zip(gws, lws)]
        if min(g) <= 0:
            return
        G = (C.c_size_t * len(g))(*g)
        Lw = (C.c_size_t * len(lws))(*lws)
        ev = P()
        _chk(self.cl.clEnqueueNDRangeKernel(self.q, k, len(g), None, G, Lw, 0, None, C.cast(C.byref(ev), P) if self.profile else None), "launch " + name)
        if self.profile:
            self.cl.clFinish(self.q)
            t0, t1 = C.c_uint64(0), C.c_uint64(0)
            self.cl.clGetEventProfilingInfo(ev, CL_PROFILING_COMMAND_START, 8, C.cast(C.byref(t0), P), None)
            self.cl.clGetEventProfilingInfo(ev, CL_PROFILING_COMMAND_END, 8, C.cast(C.byref(t1), P), None)
            self.kernel_ms.setdefault(name, []).append((t1.value - t0.value) * 1e-6)
            self.cl.clReleaseEvent(ev)

    @staticmethod
    def _f4(v):
        v = list(v) + [0.0] * (4 - len(v))
        return np.array(v[:4], np.float32)

    # ---- the rr.h-shaped interface
    def scene_alloc(self, n_tris, n_objs):
        self.n_tris, self.n_objs = n_tris, n_objs
        self.g_tri_mem = self._buf(nbytes=max(n_tris, 1) * 144)
        self.g_obj_desc = self._buf(nbytes=max(n_objs, 1) * 144)
        self.g_cut_tri_mem = self._buf(nbytes=max(n_tris, 1) * 16 * 3 * 2 * 6)         # object_context.cpp:352-354
        self.g_tri_num = self._buf(np.array([n_tris], np.uint32))

    def scene_write_tris(self, tris, first=0):
        self._write(self.g_tri_mem, np.ascontiguousarray(tris, TRIANGLE), first * 144)

    def scene_write_objs(self, objs, first=0):
        self._write(self.g_obj_desc, np.ascontiguousarray(objs, OBJ_DESC), first * 144)

    def atlas_alloc(self, n_slices, nums, sizes, mipmap_start):
        self.atlas_slices = max(int(n_slices), 2)
        self.g_texture_array = self._buf(np.zeros(self.atlas_slices * 2048 * 2048 * 4, np.uint8))
        self.g_nums = self._buf(np.ascontiguousarray(nums, np.uint32))
        self.g_sizes = self._buf(np.ascontiguousarray(sizes, np.uint32))
        self.mipmap_start = int(mipmap_start)

    def atlas_upload(self, gpu_id, rgba, flip=1):
        """texture::update_me_to_gpu (texture.cpp:323-358) + update_gpu_mipmaps (465-493)."""
        rgba = np.ascontiguousarray(rgba, np.uint8)
        h, w = rgba.shape[:2]
        img = self._image(CL_RGBA, CL_UNORM_INT8, w, h, CL_MEM_READ_ONLY)
        origin, region = (C.c_size_t * 3)(0, 0, 0), (C.c_size_t * 3)(w, h, 1)
        _chk(self.cl.clEnqueueWriteImage(self.q, img, 1, origin, region, 0, 0, rgba.ctypes.data_as(P), 0, None, None), "clEnqueueWriteImage")
        u = lambda v: np.array([v], np.uint32)
        i32 = lambda v: np.array([v], np.int32)
        self._run("update_gpu_tex", [img, u(gpu_id), u(self.mipmap_start), self.g_nums, self.g_sizes, self.g_texture_array, i32(flip)], (w, h), (16, 16))
        self._run("generate_mips", [u(gpu_id), u(self.mipmap_start), self.g_nums, self.g_sizes, self.g_texture_array, self.g_texture_array], (w, h), (16, 16))
        for i in range(3):
            self._run("generate_mip_mips", [u(gpu_id), u(i), u(self.mipmap_start), self.g_nums, self.g_sizes, self.g_texture_array, self.g_texture_array], (w, h), (16, 16))
        self.cl.clFinish(self.q)
        self.cl.clReleaseMemObject(img)

    def atlas_read_raw(self):
        out = np.empty((self.atlas_slices, 2048, 2048, 4), np.uint8)
        return self._read(self.g_texture_array, out)

    def atlas_write_raw(self, atlas):
        self._write(self.g_texture_array, np.ascontiguousarray(atlas, np.uint8))

    def lights_write(self, lights):
        """light::build, light.cpp:145-276."""
        self.lights = np.ascontiguousarray(lights, LIGHT)
        self.n_shadow = int((self.lights["shadow"] == 1).sum())
        self.n_static = int(((self.lights["shadow"] != 0) & (self.lights["is_static"] != 0)).sum())
        self.g_light_mem = self._buf(self.lights if len(self.lights) else np.zeros(1, LIGHT))
        self.g_light_num = self._buf(np.array([len(self.lights)], np.uint32))
        slab = 4 * 6 * self.L * self.L
        self.g_shadow = self._buf(nbytes=max(slab * self.n_shadow, 4))
        self.g_static_shadow = self._buf(nbytes=max(slab * self.n_static, 4))
        self._fill(self.g_shadow, max(slab * self.n_shadow, 4))
        self._fill(self.g_static_shadow, max(slab * self.n_static, 4))

    def _shadow_pair(self, light, only_static, slab_buf, slab_index, exact_count=True):
        slab_bytes = 4 * 6 * self.L * self.L
        # the reference uses clCreateSubBuffer(origin = nn*slab) (engine.cpp:1637-1646); here the slab is rendered into a
        # scratch buffer of one slab and copied into place, which needs no sub-buffer API
        if not hasattr(self, "_slab_scratch"):
            self._slab_scratch = self._buf(nbytes=slab_bytes)
        self._fill(self._slab_scratch, slab_bytes)
        no_rot = np.zeros(4, np.float32)
        zero = np.zeros(1, np.uint32)
        self._write(self.cnt_light, zero)
        self._write(self.cut_num, zero)
        self._run("prearrange_realtime_shadowing", [self.g_tri_mem, self.g_tri_num, self._f4(light["pos"]), no_rot, self.g_tid_buf, self.g_tid_buf_max_len,
                                                    self.cnt_light, self.cut_num, self.g_cut_tri_mem, self.g_obj_desc, np.array([only_static], np.int32)],
                  (self.n_tris,), (256,))
        idx = len(self.shadow_counts)
        if self.steady and idx < len(self.steady_shadow_counts):
            c0 = self.steady_shadow_counts[idx]          # engine.cpp:1680: non-blocking read, i.e. the previous frame's count
        else:
            c0 = int(self._read(self.cnt_light, np.zeros(1, np.uint32))[0])
        self.shadow_counts.append(c0)
        fragments_number = int(c0 * 1.1) + 256                                    # engine.cpp:1682
        self._run("kernel1_realtime_shadowing", [self.g_tri_mem, self.g_tid_buf, self._slab_scratch, self.cnt_light, self.g_cut_tri_mem], (fragments_number,), (256,))
        _chk(self.cl.clEnqueueCopyBuffer(self.q, self._slab_scratch, slab_buf, 0, slab_index * slab_bytes, slab_bytes, 0, None, None), "copy slab")

    def frame_shadows(self, static_dirty=0):
        """engine::generate_realtime_shadowing, engine.cpp:1601-1790."""
        slab = 4 * 6 * self.L * self.L
        self.shadow_counts = []
        if len(self.lights):
            if self.n_shadow:
                self._fill(self.g_shadow, slab * self.n_shadow)
            if static_dirty and self.n_static:
                self._fill(self.g_static_shadow, slab * self.n_static)
        nn = kk = 0
        for l in self.lights:
            if l["shadow"] == 1:
                self._shadow_pair(l, 0, self.g_shadow, nn)
                nn += 1
            if l["shadow"] and l["is_static"] and static_dirty:
                self._shadow_pair(l, 1, self.g_static_shadow, kk)
                kk += 1

    def frame_draw(self, c_pos, c_rot, clear=(0, 0, 0, 0)):
        """render_tris, engine.cpp:1794-2025."""
        if self.n_tris <= 0:
            return
        pos, rot = self._f4(c_pos), self._f4(c_rot)
        zero = np.zeros(1, np.uint32)
        self._write(self.cnt_main, zero)
        self._write(self.cut_num, zero)
        self._run("prearrange", [self.g_tri_mem, self.g_tri_num, pos, rot, self.g_tid_buf, self.g_tid_buf_max_len, self.cnt_main, self.cut_num,
                                 self.g_cut_tri_mem, self.g_obj_desc], (self.n_tris,), (256,))
        if self.steady_count is None:
            cnt = int(self._read(self.cnt_main, np.zeros(1, np.uint32))[0])     # converged current_cpu_id_num (q6)
        else:
            cnt = self.steady_count
        self.last_count = cnt
        gws = int(cnt * 1.2 + 1000)                                             # engine.cpp:1899
        d0, d1 = self.depth[self.cur], self.depth[self.cur ^ 1]
        self._run("kernel1", [self.g_tri_mem, self.g_tid_buf, d0, self.cnt_main, self.g_cut_tri_mem, self.id_img], (gws,), (256,))
        self._run("kernel2", [self.g_tri_mem, self.g_tid_buf, d0, self.id_img, self.cnt_main, self.g_cut_tri_mem], (gws,), (256,))
        u = lambda v: np.array([v], np.uint32)
        args = [self.g_tri_mem, pos, rot, d0, self.id_img, self.g_texture_array, self.screen[0], self.screen[1], self.g_nums, self.g_sizes, self.g_obj_desc,
                self.g_light_num, self.g_light_mem, self.g_shadow, self.g_static_shadow, d1, self.g_tid_buf, self.g_cut_tri_mem, self._f4(clear),
                u(self.frame_id), self.normals, np.array([0], np.int32), self.dummy_buf, self.dummy_img, u(self.mipmap_start),
                u(1 if self.cfg.use_linear_rendering else 0)]
        self._run("kernel3", args, (self.W, self.H), (16, 16))
        self.frame_id += 1
        self.c_pos, self.c_rot = pos, rot

    # ---- post passes (engine.cpp:1463-1538). The reference's own kernels, but never in place: kernel3 has written the frame to both
    # screen images (screen / backup_screen), so a pass reads screen[1] and writes screen[0] — what read_rgba8 returns.
    def _post_kernel(self, name):
        if name not in self.k:
            err = C.c_int(0)
            self.k[name] = P(self.cl.clCreateKernel(self.prog, name.encode(), C.byref(err)))
            _chk(err.value, "clCreateKernel " + name)

    def post_pseudo_aa(self):
        """do_pseudo_aa, cl2.cl:6437-6657. The engine binds `screen` and `in_screen` to the same image (engine.cpp:1854-1856); here
        `screen` is a scratch image, `front_screen` (which gets the same values, cl2.cl:6628-6629) is screen[0]."""
        self._post_kernel("do_pseudo_aa")
        if self.scratch_img is None:
            self.scratch_img = self._image(CL_RGBA, CL_FLOAT, self.W, self.H)
        self._run("do_pseudo_aa", [self.id_img, self.g_tid_buf, self.screen[1], self.scratch_img, self.screen[0], self.g_cut_tri_mem, self.depth[self.cur],
                                   self.normals], (self.W, self.H), (8, 8))

    def post_motion_blur(self, strength=1.0, camera_contribution=1.0):
        """engine::do_motion_blur, engine.cpp:1518-1538 (in = gl_screen[1], out = gl_screen[0] there too)."""
        self._post_kernel("do_motion_blur")
        f = lambda v: np.array([v], np.float32)
        self._run("do_motion_blur", [self.id_img, self.g_tid_buf, self.screen[1], self.screen[0], self.g_cut_tri_mem, self.depth[self.cur], self.g_obj_desc,
                                     np.array([self.frame_id], np.uint32), self.c_pos, self.c_rot, self.c_pos_old, self.c_rot_old, f(strength),
                                     f(camera_contribution)], (self.W, self.H), (16, 16))

    def post_godrays(self):
        """engine::draw_godrays, engine.cpp:1463-1482 (in place there; screen[1] -> screen[0] here)."""
        if not (self.lights["godray_intensity"] > 0).any():
            return
        self._post_kernel("screenspace_godrays")
        self._run("screenspace_godrays", [self.depth[self.cur], self.screen[1], self.screen[0], self.g_light_num, self.g_light_mem, self.c_pos, self.c_rot],
                  (self.W, self.H), (16, 16))

    def scene_patch_obj(self, obj_id, byte_off, data):
        b = np.frombuffer(bytes(data), dtype=np.uint8)
        self._write(self.g_obj_desc, b, obj_id * 144 + byte_off)

    def scene_read_objs(self, first, count):
        from openclrenderer_b200._abi import OBJ_DESC
        return self._read(self.g_obj_desc, np.zeros(count, OBJ_DESC), first * 144)

    def enter_steady_state(self):
        """after a converged frame: launch sizes come from that frame's counts with no blocking reads, as the engine's
        stale asynchronous reads do in steady state (engine.cpp:1680, 1836; object_context.cpp:8-15)."""
        self.steady = True
        self.steady_shadow_counts = list(self.shadow_counts)
        self.steady_count = self.last_count

    def swap_buffers(self):
        self.cur ^= 1
        self.c_pos_old, self.c_rot_old = self.c_pos, self.c_rot                 # object_context.cpp:23-24

    def sync(self):
        self.cl.clFinish(self.q)

    def read_depth(self):
        return self._read(self.depth[self.cur], np.empty((self.H, self.W), np.uint32))

    def _read_image(self, img, arr):
        origin, region = (C.c_size_t * 3)(0, 0, 0), (C.c_size_t * 3)(self.W, self.H, 1)
        _chk(self.cl.clEnqueueReadImage(self.q, img, 1, origin, region, 0, 0, arr.ctypes.data_as(P), 0, None, None), "clEnqueueReadImage")
        return arr

    def read_ids(self):
        return self._read_image(self.id_img, np.empty((self.H, self.W), np.uint32))

    def read_colour_f32(self):
        return self._read_image(self.screen[0], np.empty((self.H, self.W, 4), np.float32))

    def read_rgba8(self):
        c = self.read_colour_f32()
        return (np.clip(np.nan_to_num(c, nan=0.0), 0.0, 1.0) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)   # same quantiser as q15

    def read_shadow(self, is_static, slab):
        out = np.empty((6, self.L, self.L), np.uint32)
        return self._read(self.g_static_shadow if is_static else self.g_shadow, out, slab * out.nbytes)

    def read_fragments(self):
        n = int(self._read(self.cnt_main, np.zeros(1, np.uint32))[0])
        return self._read(self.g_tid_buf, np.empty((n, 5), np.uint32))

    def read_cutdown(self):
        n = int(self._read(self.cut_num, np.zeros(1, np.uint32))[0])
        return self._read(self.g_cut_tri_mem, np.empty((n, 3, 4), np.float32))
