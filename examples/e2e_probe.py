#!/usr/bin/env python
"""Where does the end-to-end frame time go? GPU stage times inside the e2e loop, host time per call, enqueue-only time."""
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
r.set_profiling(os.environ.get("PROFILING", "0") == "1")   # stage events cost frame time: off unless asked for
D = int(os.environ.get("DEPTH", "3"))
r.set_pipeline_depth(D)
ring = [rr.host_alloc((H, W, 4), np.uint8) for _ in range(D)]
r.frame_shadows(1)
for i in range(5):
    r.frame_e2e(*camera(s, i), s.clear, 1, ring[i % D])
r.sync()
N = 60
t_call = []
t0 = time.perf_counter()
for i in range(N):
    a = time.perf_counter()
    r.frame_e2e(*camera(s, 5 + i), s.clear, 1, ring[(5 + i) % D])
    t_call.append(time.perf_counter() - a)
r.sync()
t1 = time.perf_counter()
print(f"depth {D}: e2e {1e3 * (t1 - t0) / N:.4f} ms/frame; host time in call: median {1e3 * np.median(t_call):.4f} ms, min {1e3 * min(t_call):.4f}, max {1e3 * max(t_call):.4f}")
print("stage times of the last e2e frame:", {k: round(v, 4) for k, v in r.timings().items() if k.endswith("_ms")})
# enqueue-only cost: the same calls without read-back, never syncing -> host time per frame when the GPU queue is never empty
t0 = time.perf_counter()
for i in range(N):
    c_pos, c_rot = camera(s, 100 + i)
    r.frame_shadows(0)
    r.frame_draw(c_pos, c_rot, s.clear)
    r.swap_buffers()
t_enq = time.perf_counter() - t0
r.sync()
t_all = time.perf_counter() - t0
print(f"plain loop: host enqueue {1e3 * t_enq / N:.4f} ms/frame, with final sync {1e3 * t_all / N:.4f} ms/frame")
# plain loop with an unrelated 33 MB D2H copy per frame on another stream: does the DMA itself slow the kernels?
import torch
src = torch.zeros(H * W * 4, dtype=torch.uint8, device="cuda")
dst = torch.empty(H * W * 4, dtype=torch.uint8).pin_memory()
cs = torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(N):
    c_pos, c_rot = camera(s, 200 + i)
    r.frame_shadows(0)
    r.frame_draw(c_pos, c_rot, s.clear)
    r.swap_buffers()
    with torch.cuda.stream(cs):
        dst.copy_(src, non_blocking=True)
r.sync()
t_r = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"plain loop + background D2H: render done after {1e3 * t_r / N:.4f} ms/frame, copies done after {1e3 * t_all / N:.4f} ms/frame")
print("stage times of the last frame:", {k: round(v, 4) for k, v in r.timings().items() if k.endswith("_ms")})
