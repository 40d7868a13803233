#!/usr/bin/env python
"""End-to-end loop with the dirty-tile read-back: ms per frame, bytes per frame (RR_TILE_GRID, DEPTH from the environment)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import camera, make_scene  # noqa: E402
from openclrenderer_b200 import Renderer, rr  # noqa: E402

s = make_scene("c3")
H, W = s.cfg.height, s.cfg.width
r = Renderer(s.cfg)
s.upload(r)
D = int(os.environ.get("DEPTH", "3"))
r.set_pipeline_depth(D)
ring = [rr.host_alloc((H, W, 4), np.uint8) for _ in range(D)]
r.frame_shadows(1)
N = 200
for tiles in (0, 1):
    r.set_readback_tiles(tiles)
    k = 0
    for i in range(3 * D):
        r.frame_e2e(*camera(s, i), s.clear, 1, ring[k % D]); k += 1
    r.sync()
    r.readback_tile_bytes()
    t0 = time.perf_counter()
    for i in range(N):
        r.frame_e2e(*camera(s, 50 + i), s.clear, 1, ring[k % D]); k += 1
    r.sync()
    ms = 1e3 * (time.perf_counter() - t0) / N
    print(f"tiles={tiles} grid={os.environ.get('RR_TILE_GRID', '16')} depth={D}: {ms:.4f} ms/frame, {r.readback_tile_bytes() / N / 1e6:.2f} MB/frame through the tile path")
