"""not-gpu: host-side logic that mirrors the reference's host classes (scene.py) and the C ABI surface."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from openclrenderer_b200 import scene, _abi, _build, distributed as rrd
from openclrenderer_b200._abi import Config, TRIANGLE, OBJ_DESC, LIGHT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fov_literals():
    """engine.cpp:119-133 + std::to_string (SURVEY.md §8 config sizes)."""
    for w, lit in ((800, 230.940094), (1920, 554.256226), (3840, 1108.512451), (7680, 2217.024902)):
        assert scene.fov_for(Config.default(w, 600)) == float(np.float32(lit))


def test_plan_atlas_matches_reference_planner_by_hand():
    # one 1024 texture: sizes ascending 64,128,256,512,1024 -> one page each; index counts down from count-1 = 0
    n, nums, sizes, mip = scene.plan_atlas([1024])
    assert n == 5 and list(sizes) == [64, 128, 256, 512, 1024] and mip == 1
    assert list(nums) == [(4 << 16) | 0, (3 << 16) | 0, (2 << 16) | 0, (1 << 16) | 0, (0 << 16) | 0]
    # two 512s and one 256: 256-page holds {tex2, mip0 of tex0, mip0 of tex1} -> indices handed out downwards 2,1,0
    n, nums, sizes, mip = scene.plan_atlas([512, 512, 256])
    assert list(sizes) == [16, 32, 64, 128, 256, 512] and mip == 3
    sl256, sl512 = 4, 5
    assert nums[0] == (sl512 << 16 | 1) and nums[1] == (sl512 << 16 | 0) and nums[2] == (sl256 << 16 | 2)
    assert nums[3] == (sl256 << 16 | 1)               # tex0 mip0 (256)
    assert nums[3 + 4] == (sl256 << 16 | 0)           # tex1 mip0
    assert nums[3 + 1] == (3 << 16 | 2)               # tex0 mip1 (128): the 128-page holds 3 tiles, handed out 2,1,0
    assert nums[3 + 5] == (3 << 16 | 1) and nums[3 + 8] == (3 << 16 | 0)
    # a page overflows into a second slice of the same size: 2048/1024 = 2 -> 4 tiles per page
    n, nums, sizes, mip = scene.plan_atlas([1024] * 5)
    assert list(sizes).count(1024) == 2
    with pytest.raises(ValueError):
        scene.plan_atlas([8])


def test_load_obj_semantics(tmp_path):
    p = tmp_path / "t.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nvt 0 0\nvt 1 0\nvt 0 1\nvn 0 0 1\nvn 1 0 0\n"
                 "usemtl A\nf 1/1/1 2/2/1 3/3/1\nusemtl B\nf 1/1/2 3/3/2 4/2/2\nf 2/2/2 3/3/2 4/1/2\n")
    objs = scene.load_obj(str(p), requested_scale=2.0)
    assert [m for m, _ in objs] == ["A", "B"] and [len(t) for _, t in objs] == [1, 2]
    t = objs[0][1]
    assert np.array_equal(t["vertices"]["pos"][0, 1], [2, 0, 0, 0]) and np.array_equal(t["vertices"]["vt"][0, 2], [0, 1])
    assert np.array_equal(objs[1][1]["vertices"]["normal"][1, 0], [1, 0, 0, 0])


def test_asset_fixture_and_scenes():
    a = scene.load_assets()
    assert a["cube_tris"].size == 12 * 144 and a["cylinder_tris"].size == 4506 * 144
    assert a["red_png"].shape == (1024, 1024, 4) and a["reflection_png"].shape == (512, 512, 4)
    s = scene.scene_c1()
    assert len(s.tris) == 12 and s.cfg.width == 800 and s.cfg.test_linear == 1 and s.cfg.ssao_rad == 2.0
    assert (s.tris["vertices"]["object_id"][:, 0] == 0).all()
    s2 = scene.scene_c2()
    assert len(s2.tris) == 4508 and (s2.tris["vertices"]["object_id"][-2:, 0] == 1).all() and s2.lights["shadow"][0] == 1
    assert len(scene.uv_sphere(40, 26)) == 2000
    s3 = scene.scene_spheres(320, 200, n_spheres=3, grid=(3, 1), seed=1, n_lights=2, light_dim=64, tex_sizes=(64,))
    assert len(s3.tris) == 6000 and len(s3.objs) == 3 and len(s3.lights) == 2
    s3b = scene.scene_spheres(320, 200, n_spheres=3, grid=(3, 1), seed=1, n_lights=2, light_dim=64, tex_sizes=(64,))
    assert np.array_equal(s3.tris, s3b.tris) and np.array_equal(s3.objs, s3b.objs)         # seeded


def test_upload_call_order_mirrors_object_context_build():
    calls = []

    class Rec:
        def __getattr__(self, n):
            return lambda *a, **k: calls.append(n)
    s = scene.scene_c1()
    s.upload(Rec())
    assert calls == ["atlas_alloc", "atlas_upload", "scene_alloc", "scene_write_objs", "scene_write_tris", "lights_write"]
    calls.clear()
    s.render(Rec(), frames=2)
    assert calls == ["frame_shadows", "frame_draw", "swap_buffers", "frame_shadows", "frame_draw", "sync"]


def test_band_and_face_partition():
    assert rrd.band_rows(2160, 8, 3) == (810, 1080)
    with pytest.raises(ValueError):
        rrd.band_rows(2160, 7, 0)
    assert rrd.face_chunk(4, 8) == 3 and rrd.face_chunk(4, 5) == 5 and rrd.face_chunk(1, 4) == 2
    c = rrd.band_config(Config.default(640, 480), 4, 2)
    assert (c.band_y0, c.band_y1, c.face_rank, c.face_world) == (240, 360, 2, 4)


def test_interleaved_tiles_and_round_robin_faces():
    """the p2p split: tile t of band_tile rows belongs to rank t % world; the row sets partition the screen; the config
    carries round-robin face ownership; the tile chooser returns something the screen divides into."""
    H, world, tile = 2160, 8, 24
    masks = [rrd.owned_rows(H, tile, world, k) for k in range(world)]
    assert np.array_equal(np.sum(masks, axis=0), np.ones(H, int))
    assert masks[3][3 * tile] and masks[3][(3 + world) * tile + 5] and not masks[3][0]
    c = rrd.tile_config(Config.default(3840, H), world, 5, tile, halo=12)
    assert (c.band_tile, c.band_rank, c.band_world, c.band_halo, c.face_rank, c.face_world, c.face_interleave) == (tile, 5, world, 12, 5, world, 1)
    assert (c.band_y0, c.band_y1) == (0, 0)
    for hh, ww, halo in [(2160, 8, 12), (2160, 2, 12), (4320, 8, 24), (384, 4, 24)]:
        t = rrd.choose_tile(hh, ww, halo)
        assert 8 <= t <= 64


def test_struct_layout_matches_header():
    """compile a tiny C program against include/rr.h and compare sizeof/offsetof with the ctypes / numpy mirrors."""
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "rr.h"
int main(void){
 printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(rr_vertex), sizeof(rr_triangle), sizeof(rr_obj_desc), sizeof(rr_light), sizeof(rr_config),
   sizeof(rr_timings), offsetof(rr_obj_desc, scale), offsetof(rr_obj_desc, feature_flag), offsetof(rr_light, shadow), offsetof(rr_config, max_fragments));
 printf("%zu %zu %zu\n", sizeof(rr_mgpu_handle), offsetof(rr_config, band_tile), offsetof(rr_mgpu_handle, shadow_bytes));
 return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    want = [48, 144, 144, 64, ctypes.sizeof(Config), ctypes.sizeof(_abi.Timings), OBJ_DESC.fields["scale"][1], OBJ_DESC.fields["feature_flag"][1],
            LIGHT.fields["shadow"][1], Config.max_fragments.offset,
            ctypes.sizeof(_abi.MgpuHandle), Config.band_tile.offset, _abi.MgpuHandle.shadow_bytes.offset]
    assert got == want


def test_library_exports_every_declared_symbol():
    """the C-ABI library loads on a machine without a GPU and exports every entry point include/rr.h declares."""
    lib = _build.build_product()
    dll = ctypes.CDLL(lib)
    hdr = open(os.path.join(ROOT, "include", "rr.h")).read()
    names = set(re.findall(r"\b(rr_[a-z0-9_]+)\s*\(", hdr)) - {"rr_ctx", "rr_config", "rr_timings"}
    assert len(names) >= 30
    for n in sorted(names):
        assert hasattr(dll, n), n
    dll.rr_fov_const_from_hfov.restype = ctypes.c_float
    dll.rr_fov_const_from_hfov.argtypes = [ctypes.c_float, ctypes.c_float]
    assert dll.rr_fov_const_from_hfov(120.0, 3840.0) == float(np.float32(1108.512451))
    cfg = Config()
    dll.rr_default_config(ctypes.byref(cfg))
    assert (cfg.width, cfg.height, cfg.light_dim, cfg.depth_icutoff) == (800, 600, 1024, 20) and abs(cfg.ssao_rad - 5.0) < 1e-6


def test_product_has_no_cpu_fallback():
    """on a box without a GPU rr_create must fail loudly (and nothing in the product imports the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from openclrenderer_b200 import Renderer, RRError
    with pytest.raises(RRError) as e:
        Renderer(Config.default(64, 64))
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)
    pkg = os.path.join(ROOT, "openclrenderer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "from oracle" not in txt and "import oracle" not in txt and "liboracle" not in txt, f


def test_dirty_tile_readback_rule_keeps_the_host_buffers_exact():
    """the rule of rr_set_readback_tiles (k_tile_copy, include/rr.h), modelled on the CPU: a tile is sent when it holds a shaded pixel
    in this frame or held one in the frame last written into the same host buffer; everything else in that buffer is already the clear
    colour. With a ring of D buffers, random coverage per frame, a clear-colour change and a foreign buffer, every buffer equals its
    frame after every copy — and strictly fewer tiles travel than with whole-frame copies."""
    rng = np.random.Generator(np.random.PCG64(3))
    n_tiles, D, n_frames = 200, 3, 40
    host = [np.full(n_tiles, -1, np.int64) for _ in range(D + 1)]             # what each host buffer holds per tile (-1: never written)
    prev = [np.zeros(n_tiles, bool) for _ in range(D)]                        # per ring slot: tiles shaded in the frame last written through it
    slot_buf, slot_clear, valid = [None] * D, [None] * D, [False] * D
    sent = 0
    for f in range(n_frames):
        clear = 1000 if f < 25 else 2000                                      # "colour" of unshaded tiles; shaded tiles get a per-frame value
        shaded = rng.uniform(size=n_tiles) < 0.15
        frame = np.where(shaded, 10 * f + 7, clear)
        k = f % D
        buf = D if f == 17 else k                                             # frame 17 goes into a buffer the slot has never seen
        everything = (not valid[k]) or slot_buf[k] != buf or slot_clear[k] != clear
        need = np.ones(n_tiles, bool) if everything else (shaded | prev[k])
        host[buf][need] = frame[need]
        sent += int(need.sum())
        prev[k], slot_buf[k], slot_clear[k], valid[k] = shaded, buf, clear, True
        assert np.array_equal(host[buf], frame), f"frame {f}"
    assert sent < 0.5 * n_frames * n_tiles
