"""Host-side scene assembly that mirrors the reference's host classes for the draw path.

  load_obj                 <- obj_load.cpp:181-568 (triangulated OBJ + MTL map_Kd)
  plan_atlas               <- texture_context.cpp:94-261 (page planner, `nums` / `sizes` descriptors)
  make_obj_desc            <- object_context.cpp:228-339 (generate_gpu_object_descriptor) + object.cpp:54-83 defaults
  make_light               <- light.cpp (light::light defaults) / light.hpp:25-34
  Scene.upload             <- object_context::build (object_context.cpp:646-797) + texture_context::alloc_gpu (350-517) + light::build (145-276)
  scene_c1 / c2 / spheres  <- the configurations of BASELINE.json

Everything here is float32 numpy producing the byte layouts of include/rr.h; it drives the CUDA product and (from
tests / bench only) the CPU oracle through the same `upload`.
"""
import math
import os
from dataclasses import dataclass, field

import numpy as np

from ._abi import (Config, TRIANGLE, OBJ_DESC, LIGHT, FEATURE_IS_STATIC, FEATURE_TWO_SIDED)

MIP_LEVELS = 4          # texture_context.hpp:17
ATLAS_DIM = 2048        # texture_context.hpp:16
f32 = np.float32


# ---------------------------------------------------------------------------------------------------------------------
# OBJ loading (obj_load.cpp:181-568): `f a/b/c d/e/f g/h/i` triangulated faces, 1-based, one `usemtl` per object
# ---------------------------------------------------------------------------------------------------------------------
def load_obj(path, requested_scale=1.0):
    """Returns a list of (material_name, triangles[TRIANGLE]) in file order, one entry per `usemtl` block."""
    vl, vtl, vnl, faces, usemtl_pos, usemtl_name = [], [], [], [], [], []
    with open(path, "r") as fh:
        for ln in fh:
            ln = ln.rstrip("\r\n")
            if len(ln) < 2:
                continue
            if ln[0] == "f" and ln[1] == " ":
                idx = []
                for tok in ln.split()[1:4]:
                    a = tok.split("/")
                    idx.append((int(a[0]) - 1, int(a[1]) - 1, int(a[2]) - 1))       # decompose_face, obj_load.cpp:142-158
                faces.append(idx)
            elif ln[0] == "v" and ln[1] == " ":
                vl.append([float(x) for x in ln.split()[1:4]])
            elif ln.startswith("vt "):
                vtl.append([float(x) for x in ln.split()[1:3]])
            elif ln.startswith("vn "):
                vnl.append([float(x) for x in ln.split()[1:4]])
            elif ln.startswith("use"):
                usemtl_pos.append(len(faces))
                usemtl_name.append(ln[ln.rfind(" ") + 1:])
    vl = np.asarray(vl, dtype=f32)
    vtl = np.asarray(vtl, dtype=f32)
    vnl = np.asarray(vnl, dtype=f32)
    tris = np.zeros(len(faces), dtype=TRIANGLE)
    fa = np.asarray(faces, dtype=np.int64)                                             # [F,3,(v,vt,vn)]
    tris["vertices"]["pos"][:, :, :3] = vl[fa[:, :, 0]] * f32(requested_scale)        # vert[j].set_pos(mult(v, requested_scale)) obj_load.cpp:407
    tris["vertices"]["vt"] = vtl[fa[:, :, 1]]
    tris["vertices"]["normal"][:, :, :3] = vnl[fa[:, :, 2]]
    usemtl_pos.append(len(faces))
    return [(usemtl_name[i], tris[usemtl_pos[i]:usemtl_pos[i + 1]].copy()) for i in range(len(usemtl_name))]


def mtl_diffuse_map(mtl_path, material):
    """retrieve_diffuse_new, obj_load.cpp:21-47: the map_Kd of `newmtl material`."""
    found = False
    with open(mtl_path, "r") as fh:
        for ln in fh:
            ln = ln.strip()
            if ln.startswith("newmtl "):
                found = ln.split()[-1] == material
            elif found and ln.startswith("map_Kd "):
                return ln.split()[-1]
    return None


# ---------------------------------------------------------------------------------------------------------------------
# texture atlas planner (texture_context.cpp:94-261)
# ---------------------------------------------------------------------------------------------------------------------
def plan_atlas(tex_sizes):
    """tex_sizes: largest dimension of each texture in gpu_id (= ascending texture id) order.

    Returns (n_slices, nums[uint32], sizes[uint32], mipmap_start) exactly as texture_context::alloc_gpu computes them:
    nums[i] = slice << 16 | index for the base levels, then nums[mipmap_start + 4*i + level] for the four mips.
    """
    size_to_numbers = {}
    for s in tex_sizes:                                        # calculate_texture_pages, 107-129
        size_to_numbers[s] = size_to_numbers.get(s, 0) + 1
        for j in range(MIP_LEVELS):
            ms = s // (2 ** (j + 1))
            size_to_numbers[ms] = size_to_numbers.get(ms, 0) + 1
    pages = []                                                 # calculate_fitted_texture_pages, 131-165 (std::map: ascending size)
    for size in sorted(size_to_numbers):
        if size <= 0:
            raise ValueError("texture too small for 4 mip levels (reference divides by zero here)")
        remaining = size_to_numbers[size]
        per_page = (ATLAS_DIM // size) * (ATLAS_DIM // size)
        while remaining >= per_page:
            pages.append([size, per_page])
            remaining -= per_page
        if remaining > 0:
            pages.append([size, remaining])
    free = [p[:] for p in pages]

    def take(size):                                            # calculate_texture_slice_descriptor, 167-249
        for sl, p in enumerate(free):
            if p[0] == size and p[1] > 0:
                p[1] -= 1
                return (sl << 16) | p[1]
        raise RuntimeError("could not find a free texture page")

    nums = [take(s) for s in tex_sizes]
    for s in tex_sizes:
        for j in range(MIP_LEVELS):
            nums.append(take(s // (2 ** (j + 1))))
    sizes = [p[0] for p in pages]
    return len(pages), np.asarray(nums, dtype=np.uint32), np.asarray(sizes, dtype=np.uint32), len(tex_sizes)


# ---------------------------------------------------------------------------------------------------------------------
# descriptors / lights
# ---------------------------------------------------------------------------------------------------------------------
def make_obj_desc(pos=(0, 0, 0), quat=(0, 0, 0, 1), scale=1.0, tid=0, specular=0.9, spec_mult=1.0, diffuse=1.0, feature_flag=0):
    d = np.zeros((), dtype=OBJ_DESC)
    d["world_pos"][:3] = pos
    d["world_rot_quat"] = quat
    d["old_world_pos_1"] = d["world_pos"]
    d["old_world_pos_2"] = d["world_pos"]
    d["old_world_rot_quat_1"] = d["world_rot_quat"]
    d["old_world_rot_quat_2"] = d["world_rot_quat"]
    d["scale"] = scale
    d["tid"] = tid
    d["rid"] = 0xFFFFFFFF        # get_gpu_position_id(-1) == -1
    d["ssid"] = 0xFFFFFFFF
    d["specular"], d["spec_mult"], d["diffuse"] = specular, spec_mult, diffuse     # object.cpp:76-78
    d["feature_flag"] = feature_flag
    return d


def make_light(pos, col=(1, 1, 1), shadow=0, brightness=1.0, radius=20000.0, diffuse=1.0, is_static=0):
    l = np.zeros((), dtype=LIGHT)
    l["pos"][:3] = pos
    l["col"][:3] = col
    l["shadow"], l["brightness"], l["radius"], l["diffuse"], l["godray_intensity"], l["is_static"] = shadow, brightness, radius, diffuse, 0.0, is_static
    return l


@dataclass
class Scene:
    cfg: Config
    tris: np.ndarray                      # TRIANGLE[T], vertices[0].object_id already stamped (fill_ids, cl2.cl:4231)
    objs: np.ndarray                      # OBJ_DESC[n]
    lights: np.ndarray                    # LIGHT[n]
    textures: list                        # RGBA8 arrays in gpu_id order
    c_pos: tuple = (0.0, 0.0, 0.0)
    c_rot: tuple = (0.0, 0.0, 0.0)
    clear: tuple = (0.0, 0.0, 0.0, 0.0)
    name: str = ""
    meta: dict = field(default_factory=dict)

    def upload(self, r, atlas_raw=None):
        """object_context::build(true): textures -> descriptors -> triangles -> lights."""
        n_slices, nums, sizes, mip_start = plan_atlas([max(t.shape[0], t.shape[1]) for t in self.textures])
        r.atlas_alloc(n_slices, nums, sizes, mip_start)
        if atlas_raw is not None:
            r.atlas_write_raw(atlas_raw)
        else:
            for gid, t in enumerate(self.textures):
                r.atlas_upload(gid, t, flip=1)              # texture::update_me_to_gpu passes flip = true (texture.cpp:355)
        r.scene_alloc(len(self.tris), len(self.objs))
        r.scene_write_objs(self.objs)
        r.scene_write_tris(self.tris)
        r.lights_write(self.lights)
        return r

    def render(self, r, frames=1, shadows=True):
        """The frame loop of main.cpp:262-291, `frames` times with a static camera; leaves the last frame readable."""
        for i in range(frames):
            if i:
                r.swap_buffers()
            if shadows and len(self.lights):
                r.frame_shadows(1 if i == 0 else 0)
            r.frame_draw(self.c_pos, self.c_rot, self.clear)
        r.sync()
        return r


def stamp_object_ids(tri_lists):
    """object_context::alloc_gpu writes `pad = object_g_id` into vertex 0 of every triangle (object_context.cpp:427 / fill_ids)."""
    out = []
    for oid, t in enumerate(tri_lists):
        t = t.copy()
        t["vertices"]["object_id"][:, 0] = oid
        out.append(t)
    return np.concatenate(out) if out else np.zeros(0, dtype=TRIANGLE)


# ---------------------------------------------------------------------------------------------------------------------
# assets of configs 1-2: read from the reference tree when present, else from the committed fixture
# ---------------------------------------------------------------------------------------------------------------------
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSET_FIXTURE = os.path.join(_ROOT, "tests", "golden", "assets.npz")


def load_assets():
    """cube.obj, high_cylinder_forward.obj, red.png, test_reflection_map.png as arrays (tests/golden/make_assets.py wrote them)."""
    z = np.load(ASSET_FIXTURE)
    return {k: z[k] for k in z.files}


def _front_facing(tri_pos, obj_pos, scale, c_pos, c_rot, w, h, fov):
    """float64 sanity helper used only to choose a winding for procedurally built quads."""
    def rot(p):
        c, s = np.cos(c_rot), np.sin(c_rot)
        rel = p - np.asarray(c_pos, dtype=np.float64)
        t = s[2] * rel[1] + c[2] * rel[0]
        u = c[1] * rel[2] + s[1] * t
        v = c[2] * rel[1] - s[2] * rel[0]
        return np.array([c[1] * t - s[1] * rel[2], s[0] * u + c[0] * v, c[0] * u - s[0] * v])
    pr = [rot(np.asarray(p, dtype=np.float64) * scale + np.asarray(obj_pos, dtype=np.float64)) for p in tri_pos]
    sp = [np.array([p[0] * fov / p[2] + w / 2, p[1] * fov / p[2] + h / 2]) for p in pr]
    a, b = sp[1] - sp[0], sp[2] - sp[0]
    return a[0] * b[1] - a[1] * b[0] < 0                      # backface_cull_expanded, cl2.cl:491-494


def fov_for(cfg):
    if cfg.fov_const > 0:
        return cfg.fov_const
    fr = f32(float(f32(cfg.hfov_deg) / f32(360.0)) * 2 * math.pi)
    v = f32((cfg.width / 2) / math.tan(float(f32(fr / f32(2)))))
    return float(f32(float("%f" % v)))


def scene_c1(profile="A"):
    """config 1: objects/cube.obj, dynamic_scale 100, one non-shadow light, 800x600."""
    a = load_assets()
    cfg = Config.profile_a(800, 600) if profile == "A" else Config.default(800, 600)
    tris = stamp_object_ids([a["cube_tris"].view(TRIANGLE).reshape(-1)])
    objs = np.array([make_obj_desc(pos=(0, 0, 0), scale=100.0, tid=0)], dtype=OBJ_DESC)
    lights = np.array([make_light((-200, 300, -300), shadow=0, brightness=1.0, radius=20000.0)], dtype=LIGHT)
    return Scene(cfg, tris, objs, lights, [a["red_png"]], c_pos=(0, 150, -400), c_rot=(0.3, 0, 0), name="c1_cube_800x600")


def ground_quad(y, half, cam, cfg, scale=1.0):
    """two triangles spanning x,z in +-half at height y, normal +y, wound to be front-facing for `cam`."""
    p = [(-half, y, -half), (half, y, -half), (half, y, half), (-half, y, half)]
    uv = [(0, 0), (1, 0), (1, 1), (0, 1)]
    order = [(0, 1, 2), (0, 2, 3)]
    t = np.zeros(2, dtype=TRIANGLE)
    fov = fov_for(cfg)
    for k, (i0, i1, i2) in enumerate(order):
        idx = (i0, i1, i2)
        if not _front_facing([p[i] for i in idx], (0, 0, 0), scale, cam[0], np.asarray(cam[1], dtype=np.float64), cfg.width, cfg.height, fov):
            idx = (i0, i2, i1)
        for j, i in enumerate(idx):
            t["vertices"]["pos"][k, j, :3] = p[i]
            t["vertices"]["normal"][k, j, :3] = (0, 1, 0)
            t["vertices"]["vt"][k, j] = uv[i]
    return t


def scene_c2(width=1920, height=1080, light_dim=1024):
    """config 2: high_cylinder_forward.obj (scale 200) + a ground quad, test_reflection_map.png, one shadow-casting light."""
    a = load_assets()
    cfg = Config.profile_a(width, height, light_dim=light_dim)
    cam = ((0, 200, -700), (0.25, 0, 0))
    cyl = a["cylinder_tris"].view(TRIANGLE).reshape(-1)
    quad = ground_quad(-250.0, 1000.0, cam, cfg)
    tris = stamp_object_ids([cyl, quad])
    objs = np.array([make_obj_desc(scale=200.0, tid=0), make_obj_desc(scale=1.0, tid=0)], dtype=OBJ_DESC)
    lights = np.array([make_light((-300, 600, -300), shadow=1, is_static=0)], dtype=LIGHT)
    return Scene(cfg, tris, objs, lights, [a["reflection_png"]], c_pos=cam[0], c_rot=cam[1], name=f"c2_cylinder_{width}x{height}")


# ---------------------------------------------------------------------------------------------------------------------
# synthetic tessellated-sphere field (configs 3-5)
# ---------------------------------------------------------------------------------------------------------------------
def uv_sphere(slices=40, stacks=26):
    """Unit UV sphere: `stacks` latitude bands, the two polar bands as single fans -> 2*slices*(stacks-1) triangles
    (40 x 26 -> 2000), outward winding, smooth normals (= positions), spherical UVs."""
    bands = stacks

    def vert(i, j):
        th = math.pi * i / bands
        ph = 2.0 * math.pi * j / slices
        p = (math.sin(th) * math.cos(ph), math.cos(th), math.sin(th) * math.sin(ph))
        return p, (j / slices, i / bands)

    out = []
    for i in range(bands):
        for j in range(slices):
            a, b, c, d = vert(i, j), vert(i + 1, j), vert(i + 1, j + 1), vert(i, j + 1)
            if i == 0:
                out.append((a, b, c))
            elif i == bands - 1:
                out.append((a, b, d))
            else:
                out.append((a, b, c))
                out.append((a, c, d))
    t = np.zeros(len(out), dtype=TRIANGLE)
    for k, tri in enumerate(out):
        P = np.array([v[0] for v in tri], dtype=np.float64)
        n = np.cross(P[1] - P[0], P[2] - P[0])
        order = (0, 1, 2) if np.dot(n, P.mean(axis=0)) > 0 else (0, 2, 1)          # outward CCW (right-handed sense)
        for j, o in enumerate(order):
            t["vertices"]["pos"][k, j, :3] = tri[o][0]
            t["vertices"]["normal"][k, j, :3] = tri[o][0]
            t["vertices"]["vt"][k, j] = tri[o][1]
    return t


def procedural_texture(size, seed):
    """Deterministic RGBA8 texture: smooth gradients + checker + hash noise, alpha 255."""
    y, x = np.mgrid[0:size, 0:size].astype(np.uint32)
    h = (x * np.uint32(0x9E3779B1)) ^ (y * np.uint32(0x85EBCA77)) ^ np.uint32((seed * 0xC2B2AE3D) & 0xFFFFFFFF)
    h ^= h >> np.uint32(15)
    h *= np.uint32(0x2C1B3C6D)
    h ^= h >> np.uint32(12)
    noise = (h & np.uint32(0xFF)).astype(np.float32)
    fx, fy = x.astype(np.float32) / size, y.astype(np.float32) / size
    cell = max(size // 8, 1)
    checker = (((x // cell) + (y // cell)) & 1).astype(np.float32)
    rng = np.random.Generator(np.random.PCG64(seed))
    base = rng.uniform(40, 215, size=3).astype(np.float32)
    img = np.zeros((size, size, 4), dtype=np.uint8)
    img[..., 0] = np.clip(base[0] * (0.6 + 0.4 * fx) + 30 * checker + 0.15 * noise, 0, 255)
    img[..., 1] = np.clip(base[1] * (0.6 + 0.4 * fy) + 20 * (1 - checker) + 0.15 * noise, 0, 255)
    img[..., 2] = np.clip(base[2] * (0.5 + 0.5 * fx * fy) + 25 * checker + 0.15 * noise, 0, 255)
    img[..., 3] = 255
    return img


def scene_spheres(width=3840, height=2160, n_spheres=500, grid=(25, 20), seed=20260, n_lights=4, light_dim=1024, shadows=True,
                  tex_sizes=(1024, 1024, 512, 512, 256, 256, 128, 128), region_scale=1.0, slices=40, stacks=26, profile="A", name=None):
    """configs 3-5: n_spheres UV-spheres of slices x stacks on a jittered grid, one object per sphere,
    textures round-robin, `n_lights` lights on a ring. Defaults = config 3 (1,000,000 triangles at 3840x2160)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    cfg = (Config.profile_a if profile == "A" else Config.default)(width, height, light_dim=light_dim)
    unit = uv_sphere(slices, stacks)
    gx, gz = grid
    assert gx * gz >= n_spheres
    xs = np.linspace(-6000.0, 6000.0, gx) * region_scale
    zs = np.linspace(500.0, 12000.0, gz) * region_scale
    cell_x = (xs[1] - xs[0]) if gx > 1 else 0.0
    cell_z = (zs[1] - zs[0]) if gz > 1 else 0.0
    objs = np.zeros(n_spheres, dtype=OBJ_DESC)
    tri_lists = []
    for i in range(n_spheres):
        ix, iz = i % gx, i // gx
        pos = (xs[ix] + rng.uniform(-0.3, 0.3) * cell_x, rng.uniform(-500.0, 500.0), zs[iz] + rng.uniform(-0.3, 0.3) * cell_z)
        radius = rng.uniform(40.0, 160.0)
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        objs[i] = make_obj_desc(pos=pos, quat=q, scale=radius, tid=i % len(tex_sizes))
        tri_lists.append(unit)
    tris = np.tile(unit, n_spheres)
    tris["vertices"]["object_id"][:, 0] = np.repeat(np.arange(n_spheres, dtype=np.uint32), len(unit))
    zc = float(zs.mean())
    lights = np.zeros(n_lights, dtype=LIGHT)
    for k in range(n_lights):
        ang = 2.0 * math.pi * k / max(n_lights, 1) + 0.3
        lights[k] = make_light((4000.0 * region_scale * math.cos(ang), 2500.0, zc + 4000.0 * region_scale * math.sin(ang)),
                               col=(1.0, 0.95 - 0.05 * (k % 3), 0.9 - 0.1 * (k % 2)), shadow=1 if shadows else 0, brightness=1.0, radius=20000.0)
    textures = [procedural_texture(s, seed * 31 + k) for k, s in enumerate(tex_sizes)]
    nm = name or f"spheres_{n_spheres}x{len(unit)}_{width}x{height}_L{n_lights}"
    return Scene(cfg, tris, objs, lights, textures, c_pos=(0.0, 800.0, -1500.0), c_rot=(0.2, 0.0, 0.0), name=nm,
                 meta={"n_spheres": n_spheres, "tris_per_sphere": len(unit), "seed": seed})


def scene_c3():
    return scene_spheres(3840, 2160, 500, (25, 20), 20260, 4, 1024, name="c3_spheres_1Mtri_3840x2160_4lights")


def scene_c4():
    return scene_spheres(7680, 4320, 8000, (100, 80), 20261, 4, 1024, region_scale=4.0, name="c4_spheres_16Mtri_7680x4320_4lights")


def scene_c5():
    return scene_spheres(3840, 2160, 500, (25, 20), 20262, 8, 2048, name="c5_spheres_1Mtri_3840x2160_8lights_L2048")
