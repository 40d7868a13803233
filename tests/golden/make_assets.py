"""Generates tests/golden/assets.npz from the reference's own assets (run in the build container only;
/root/reference does not exist on the GPU box).  python tests/golden/make_assets.py

  cube_tris / cylinder_tris : TRIANGLE bytes produced by openclrenderer_b200.scene.load_obj (obj_load.cpp semantics)
  red_png / reflection_png  : RGBA8 pixels as sf::Image::loadFromFile would return them (lossless PNG decode)
"""
import os
import sys

import numpy as np
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from openclrenderer_b200.scene import load_obj, mtl_diffuse_map  # noqa: E402

REF = "/root/reference/objects"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets.npz")


def png(name):
    im = Image.open(os.path.join(REF, name)).convert("RGBA")
    return np.asarray(im, dtype=np.uint8)


def main():
    cube = load_obj(os.path.join(REF, "cube.obj"))
    cyl = load_obj(os.path.join(REF, "high_cylinder_forward.obj"))
    assert len(cube) == 1 and len(cyl) == 1
    assert mtl_diffuse_map(os.path.join(REF, "cube.mtl"), cube[0][0]) == "red.png"
    print("cube tris", len(cube[0][1]), "cylinder tris", len(cyl[0][1]))
    np.savez_compressed(OUT, cube_tris=cube[0][1].view(np.uint8), cylinder_tris=cyl[0][1].view(np.uint8),
                        red_png=png("red.png"), reflection_png=png("test_reflection_map.png"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
