#!/usr/bin/env python
"""Render a few frames of ONE rank's share of the sort-first split, alone (no exchange) — for ncu launch lists of a rank.
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python examples/rank_frame.py --world 8 --rank 4 --tile 32"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import camera, make_scene  # noqa: E402
from openclrenderer_b200 import Renderer, distributed as rrd  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3")
ap.add_argument("--world", type=int, default=1)
ap.add_argument("--rank", type=int, default=0)
ap.add_argument("--tile", type=int, default=32)
ap.add_argument("--frames", type=int, default=4)
a = ap.parse_args()
s = make_scene(a.workload)
cfg = s.cfg
if a.world > 1:
    halo = rrd.ssao_halo(s, [camera(s, i) for i in range(7)])
    cfg = rrd.tile_config(s.cfg, a.world, a.rank, a.tile, halo)
r = Renderer(cfg)
s.upload(r)
r.set_profiling(True)
for i in range(a.frames):
    c_pos, c_rot = camera(s, i)
    r.frame_shadows(0)
    r.frame_draw(c_pos, c_rot, s.clear)
    r.swap_buffers()
r.sync()
print(r.timings())
if os.environ.get("RANK_FRAME_LOOP"):                      # steady-state time of this rank's frames alone (no exchange, no profiling events)
    import time
    r.set_profiling(False)
    n = int(os.environ["RANK_FRAME_LOOP"])
    for rep in range(2):
        t0 = time.perf_counter()
        for i in range(n):
            c_pos, c_rot = camera(s, i)
            r.frame_shadows(0)
            r.frame_draw(c_pos, c_rot, s.clear)
            r.swap_buffers()
        r.sync()
        ms = 1e3 * (time.perf_counter() - t0) / n
    print(f"rank {a.rank} of {a.world}: {ms:.4f} ms/frame alone")
