#!/usr/bin/env python
"""How much would pipelining consecutive frames buy? Two independent contexts render the same scene on one GPU, their frames
submitted alternately from one host thread (no synchronisation in between): the GPU is free to overlap one context's shading with
the other's setup / shadow kernels. Aggregate ms per frame against one context alone = the head-room of cross-frame overlap."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import camera, make_scene  # noqa: E402
from openclrenderer_b200 import Renderer  # noqa: E402

s = make_scene(sys.argv[1] if len(sys.argv) > 1 else "c3")
N = 200


def loop(rs, n):
    for i in range(n):
        c_pos, c_rot = camera(s, i)
        for r in rs:
            r.frame_shadows(0)
            r.frame_draw(c_pos, c_rot, s.clear)
            r.swap_buffers()


a = Renderer(s.cfg)
s.upload(a)
loop([a], 10)
a.sync()
t0 = time.perf_counter()
loop([a], N)
a.sync()
one = 1e3 * (time.perf_counter() - t0) / N
b = Renderer(s.cfg)
s.upload(b)
loop([a, b], 10)
a.sync(), b.sync()
t0 = time.perf_counter()
loop([a, b], N)
a.sync(), b.sync()
two = 1e3 * (time.perf_counter() - t0) / (2 * N)
print(f"one context: {one:.4f} ms/frame; two contexts interleaved: {two:.4f} ms/frame aggregate ({100 * (one / two - 1):.1f} % more frames/s)")
