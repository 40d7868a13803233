#!/usr/bin/env python
"""Load-balance probe for the sort-first split, on ONE GPU: every rank's configuration (interleaved row tiles + its share of
the cubemap faces) is rendered alone, without the exchange, and its stage times recorded. max over ranks = what the slowest
GPU of an N-GPU run computes per frame (the exchange comes on top). Used to pick the tile size.

  python examples/scaling_probe.py --workload c3 --worlds 2,4,8 --tiles 16,24,32,48,64
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import camera, make_scene  # noqa: E402
from openclrenderer_b200 import Renderer, distributed as rrd  # noqa: E402


def run(s, cfg, frames=6):
    r = Renderer(cfg)
    s.upload(r)
    r.set_profiling(True)
    acc = {}
    for i in range(frames):
        c_pos, c_rot = camera(s, i)
        r.frame_shadows(0)
        r.frame_draw(c_pos, c_rot, s.clear)
        r.swap_buffers()
        t = r.timings()
        if i >= 2:
            for k in ("shadow_depth_ms", "setup_ms", "depth_ms", "id_ms", "shade_ms", "frame_ms"):
                acc[k] = acc.get(k, 0.0) + t[k] / (frames - 2)
    acc["n_cutdown"], acc["n_fragments"] = t["n_cutdown"], t["n_fragments"]
    r.close()
    return acc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--worlds", default="2,4,8")
    ap.add_argument("--tiles", default="16,32,64")
    ap.add_argument("--contiguous", action="store_true", help="also probe the contiguous-band split")
    a = ap.parse_args()
    s = make_scene(a.workload)
    halo = rrd.ssao_halo(s, [camera(s, i) for i in range(7)])
    base = run(s, s.cfg)
    print(json.dumps({"world": 1, "halo": halo, **{k: round(v, 4) for k, v in base.items()}}), flush=True)
    for world in [int(x) for x in a.worlds.split(",")]:
        variants = [("tile", int(t)) for t in a.tiles.split(",")] + ([("band", 0)] if a.contiguous else [])
        for kind, tile in variants:
            rows = []
            for rank in range(world):
                cfg = rrd.tile_config(s.cfg, world, rank, tile, halo) if kind == "tile" else \
                    rrd.band_config(s.cfg, world, rank, halo).copy(face_interleave=0)
                rows.append(run(s, cfg))
            worst = {k: round(max(r[k] for r in rows), 4) for k in rows[0]}
            mean = round(sum(r["frame_ms"] for r in rows) / world, 4)
            print(json.dumps({"world": world, "split": kind, "tile": tile, "max": worst, "mean_frame_ms": mean,
                              "frame_ms_per_rank": [round(r["frame_ms"], 4) for r in rows],
                              "shadow_ms_per_rank": [round(r["shadow_depth_ms"], 4) for r in rows]}), flush=True)


if __name__ == "__main__":
    main()
