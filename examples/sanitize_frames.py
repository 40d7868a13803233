#!/usr/bin/env python
"""A small workload that touches every kernel of a frame, for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
  compute-sanitizer --tool racecheck python examples/sanitize_frames.py
c1 (cube, no shadows), a small sphere scene with shadows, clipping (camera inside the field), post passes, the atlas paths, and a
two-context peer exchange (rr_mgpu_connect_local) on one device: the spin-wait protocols (look-back scan, k_wait_flags) run too."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np  # noqa: E402

from openclrenderer_b200 import Renderer, distributed as rrd, rr, scene  # noqa: E402

which = set(sys.argv[1:]) or {"single", "mgpu"}
if "single" in which:
    s = scene.scene_c1("A")
    r = Renderer(s.cfg)
    s.upload(r)
    s.render(r, frames=2)
    r.sync()
    s = scene.scene_spheres(320, 192, n_spheres=6, grid=(3, 2), seed=3, n_lights=2, light_dim=64, tex_sizes=(64, 32))
    r = Renderer(s.cfg)
    s.upload(r)
    lights = s.lights.copy()
    lights["godray_intensity"][0] = 0.5
    r.lights_write(lights)
    for i in range(3):
        r.frame_shadows(1 if i == 0 else 0)
        # the last camera sits inside the field: near-plane clipping, big fragments, the work-list raster
        c_pos = (s.c_pos[0] + 40.0 * i, s.c_pos[1] - 300.0 * i, s.c_pos[2] + 2500.0 * i)
        r.frame_draw(c_pos, s.c_rot, s.clear)
        r.post_godrays()
        r.post_motion_blur(1.0, 1.0)
        r.post_pseudo_aa()
        r.sync()
        r.swap_buffers()
    # rr_frame_e2e: ring of host buffers, plain copies and the dirty-tile read-back (stores into mapped host memory)
    r.set_pipeline_depth(2)
    bufs = [rr.host_alloc((s.cfg.height, s.cfg.width, 4)) for _ in range(2)]
    for tiles in (0, 1):
        r.set_readback_tiles(tiles)
        for i in range(5):
            r.frame_e2e((s.c_pos[0] + 60.0 * i, s.c_pos[1], s.c_pos[2]), s.c_rot, s.clear, 1, bufs[i % 2])
        r.sync()
    r.set_readback_tiles(0)
    r.atlas_fill_colour(0, (255, 0, 0, 255), 64, 64)
    r.atlas_upload_mono(1, np.arange(32 * 32, dtype=np.uint8).reshape(32, 32), 32, 32)
    r.sync()
    print("single ok", r.timings()["launches"], "launches")
if "mgpu" in which:
    s = scene.scene_spheres(320, 192, n_spheres=6, grid=(3, 2), seed=3, n_lights=2, light_dim=64, tex_sizes=(64, 32))
    world, tile = 2, 16
    rs = [Renderer(rrd.tile_config(s.cfg, world, k, tile, 24)) for k in range(world)]
    for r in rs:
        s.upload(r)
    rr.mgpu_connect_local(rs)
    for i in range(3):
        for r in rs:
            r.frame_shadows(1 if i == 0 else 0)
        for r in reversed(rs):
            r.frame_draw((s.c_pos[0] + 30.0 * i, s.c_pos[1], s.c_pos[2]), s.c_rot, s.clear)
        for r in rs:
            r.sync()
        for r in rs:
            r.swap_buffers()
    print("mgpu ok")
