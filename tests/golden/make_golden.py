"""Generates the golden fixtures of tests/golden/ from THE REFERENCE ITSELF: the unmodified kernels of cl2.cl run through
the NVIDIA OpenCL ICD on the GPU box (oracle/ref_opencl.py, mode "pinned" = the arithmetic of SURVEY.md §8c; the two
deviations from "as shipped" are listed there). Run on the GPU box:

    gpurun -- python tests/golden/make_golden.py          # writes gpurun_out/golden_ref.npz
    cp gpurun_out/golden_ref.npz tests/golden/golden_ref.npz

Per scene it stores the reference's depth buffer, the triangle id seen through its id buffer (fragment ids are
allocation-order dependent in the reference, triangle ids are not), its RGBA8 frame (float output quantised with q15),
its fragment count, and for shadow-casting scenes the cubemap of light 0. The CPU tests check the oracle against these
(-m "not gpu"); the GPU tests check the CUDA product against them."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from openclrenderer_b200 import scene  # noqa: E402

def _sph_close():
    """the sphere field seen from inside it: a quarter of the screen covered, triangles from sub-pixel to hundreds of pixels, some cut
    by the near plane (multi-chunk fragments, the clip path) — the `sph` scene alone covers under 1 % of its frame."""
    s = scene.scene_spheres(960, 540, n_spheres=24, grid=(6, 4), seed=7, n_lights=4, light_dim=256, tex_sizes=(256, 128, 64, 64))
    s.c_pos = (511.0, 25.0, 257.0)          # ten units off the surface of sphere 3: 27 % of the frame covered, triangles of up to 83 chunks
    s.c_rot = (0.0, 0.0, 0.0)
    return s


SCENES = {
    "c1A": lambda: scene.scene_c1("A"),
    "c1B": lambda: scene.scene_c1("B"),
    "c2": lambda: scene.scene_c2(),
    "c2_small": lambda: scene.scene_c2(640, 360, 256),
    "sph": lambda: scene.scene_spheres(960, 540, n_spheres=24, grid=(6, 4), seed=7, n_lights=4, light_dim=256, tex_sizes=(256, 128, 64, 64)),
    "sph_close": _sph_close,
}


def tri_ids(r):
    ids, fr, d = r.read_ids(), r.read_fragments(), r.read_depth()
    cov = d != 0xFFFFFFFF
    t = np.full(ids.shape, -1, np.int32)
    ok = cov & (ids < len(fr))
    t[ok] = fr[ids[ok], 0]
    return t


def main():
    from oracle.ref_opencl import RefCL, available
    assert available(), "needs the NVIDIA OpenCL ICD and oracle/_ref/cl2.cl.gz (GPU box)"
    out, meta = {}, {}
    for name, mk in SCENES.items():
        s = mk()
        r = RefCL(s.cfg, mode="pinned")
        s.upload(r)
        s.render(r, frames=2)
        out[name + "_depth"] = r.read_depth()
        out[name + "_tri"] = tri_ids(r)
        out[name + "_rgba"] = r.read_rgba8()
        fr = r.read_fragments()
        # allocation-order independent view of the fragment records: sorted (triangle, chunk, bits(rconst), object)
        key = np.lexsort((fr[:, 1], fr[:, 0]))
        out[name + "_frags"] = fr[key][:, [0, 1, 3, 4]]
        if name in ("c2_small", "sph", "sph_close"):
            out[name + "_shadow0"] = r.read_shadow(0, 0)
        meta[name] = {"options": r.options, "fragments": int(len(fr)), "covered": int((out[name + "_depth"] != 0xFFFFFFFF).sum())}
        print(name, meta[name]["fragments"], meta[name]["covered"])
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    dst = os.path.join(ROOT, "gpurun_out", "golden_ref.npz")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst))


if __name__ == "__main__":
    main()
