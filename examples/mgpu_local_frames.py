#!/usr/bin/env python
"""The peer-memory exchange at full size on ONE device: `world` contexts of config 3 in this process, wired with rr_mgpu_connect_local
(the same kernels — k_fill_faces, k_push_faces, k_wait_flags, k_signal_flag(s) — as the cross-process / cross-GPU path). For ncu
captures of the exchange kernels, which must not be taken under a multi-rank launch:
  ncu --set full -k regex:k_push_faces -c 2 python examples/mgpu_local_frames.py --world 2"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from bench import camera, make_scene  # noqa: E402
from openclrenderer_b200 import Renderer, distributed as rrd, rr  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3")
ap.add_argument("--world", type=int, default=2)
ap.add_argument("--tile", type=int, default=64)
ap.add_argument("--frames", type=int, default=3)
a = ap.parse_args()
s = make_scene(a.workload)
halo = rrd.ssao_halo(s, [camera(s, i) for i in range(7)])
rs = [Renderer(rrd.tile_config(s.cfg, a.world, k, a.tile, halo)) for k in range(a.world)]
for r in rs:
    s.upload(r)
rr.mgpu_connect_local(rs)
for i in range(a.frames):
    c_pos, c_rot = camera(s, i)
    for r in rs:
        r.frame_shadows(1 if i == 0 else 0)
    for r in reversed(rs):
        r.frame_draw(c_pos, c_rot, s.clear)
    for r in rs:
        r.sync()
    for r in rs:
        r.swap_buffers()
print("pushed bytes per frame:", [r.mgpu_pushed_bytes() // a.frames for r in rs])
