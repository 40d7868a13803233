#!/usr/bin/env python
"""Diagnostic: split a workload over `world` contexts on ONE GPU (rr_mgpu_connect_local) and report where the composite
differs from the frame a single context renders (expected: nowhere)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from bench import camera, make_scene  # noqa: E402
from openclrenderer_b200 import Renderer, distributed as rrd, rr  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c4")
ap.add_argument("--world", type=int, default=4)
ap.add_argument("--tile", type=int, default=32)
ap.add_argument("--halo", type=int, default=None)
ap.add_argument("--no-cull", action="store_true")
a = ap.parse_args()
s = make_scene(a.workload)
H, W = s.cfg.height, s.cfg.width
halo = a.halo if a.halo is not None else rrd.ssao_halo(s, [camera(s, i) for i in range(7)])
print("halo", halo, flush=True)
c_pos, c_rot = camera(s, 12345)
solo = Renderer(s.cfg)
s.upload(solo)
solo.frame_shadows(0)
solo.frame_draw(c_pos, c_rot, s.clear)
solo.sync()
want, wdepth = solo.read_rgba8(), solo.read_depth()
wfr, wid = solo.read_fragments(), solo.read_ids()
solo.close()
rs = [Renderer(rrd.tile_config(s.cfg, a.world, k, a.tile, halo).copy(cluster_cull=-1 if a.no_cull else 0)) for k in range(a.world)]
for r in rs:
    s.upload(r)
rr.mgpu_connect_local(rs)
for r in rs:
    r.frame_shadows(0)
for r in reversed(rs):
    r.frame_draw(c_pos, c_rot, s.clear)
for r in rs:
    r.sync()
got = rs[0].read_rgba8()
bad = (got != want).any(axis=-1)
ys, xs = np.nonzero(bad)
print("differing pixels", int(bad.sum()), "of", H * W)
if len(ys):
    print("rows mod tile histogram", np.bincount(ys % a.tile, minlength=a.tile))
    print("max channel diff", int(np.abs(got.astype(int) - want.astype(int)).max()))
    cov = wdepth != 0xFFFFFFFF
    print("differing pixels covered in the solo frame:", int(cov[bad].sum()))
    for k, r in enumerate(rs):
        own = rrd.owned_rows(H, a.tile, a.world, k)
        d = r.read_depth()
        m = bad & own[:, None]
        dd = (d != wdepth) & own[:, None]
        fr, ids = r.read_fragments(), r.read_ids()
        okc = cov & own[:, None]
        t_mine = np.full((H, W), -1, np.int64)
        t_mine[okc] = fr[np.minimum(ids[okc], len(fr) - 1), 0]
        t_want = np.full((H, W), -1, np.int64)
        t_want[okc] = wfr[wid[okc], 0]
        print(f"rank {k}: colour diffs in owned rows {int(m.sum())}, depth diffs in owned rows {int(dd.sum())}, triangle-id diffs {int((t_mine != t_want).sum())}", flush=True)
        if m.sum():
            yy, xx = np.nonzero(m)
            for j in range(min(5, len(yy))):
                y, x = yy[j], xx[j]
                print("   px", x, y, "got", got[y, x], "want", want[y, x], "depth", d[y, x], wdepth[y, x], "tri", t_mine[y, x], t_want[y, x])
