"""The C++ host layer that keeps the reference's class names (openclrenderer_b200/host/rr_host.hpp) and the main.cpp-shaped
example: compiled against include/rr.h here (no GPU), and on the GPU box run against the Python-driven renderer."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_example(tmp):
    from openclrenderer_b200 import _build
    lib = _build.build_product()
    exe = os.path.join(tmp, "main_headless")
    libdir = os.path.dirname(lib)
    cuda_lib = "/usr/local/cuda/lib64"
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-o", exe, os.path.join(ROOT, "examples", "main_headless.cpp"), "-L" + libdir, "-lrr_b200",
                    "-L" + cuda_lib, "-lcudart", "-lz", "-Wl,-rpath," + libdir, "-Wl,-rpath," + cuda_lib], check=True)
    return exe


def write_obj_assets(tmp):
    """cube.obj + cube.mtl + red.png rebuilt from the committed asset fixture (the reference tree is not on the GPU box)."""
    from PIL import Image
    from openclrenderer_b200 import scene
    from openclrenderer_b200._abi import TRIANGLE
    a = scene.load_assets()
    tris = a["cube_tris"].view(TRIANGLE).reshape(-1)
    with open(os.path.join(tmp, "cube.obj"), "w") as f:
        f.write("mtllib cube.mtl\no Cube\n")
        n = 0
        for t in tris:
            for v in t["vertices"]:
                f.write("v %.9g %.9g %.9g\nvt %.9g %.9g\nvn %.9g %.9g %.9g\n" % (*v["pos"][:3], *v["vt"], *v["normal"][:3]))
        f.write("usemtl CubeMat\ns off\n")
        for t in tris:
            f.write("f %d/%d/%d %d/%d/%d %d/%d/%d\n" % tuple(x for k in range(3) for x in (n + k + 1,) * 3))
            n += 3
    open(os.path.join(tmp, "cube.mtl"), "w").write("newmtl CubeMat\nKd 0.8 0.8 0.8\nmap_Kd red.png\n")
    Image.fromarray(a["red_png"], "RGBA").save(os.path.join(tmp, "red.png"))
    return os.path.join(tmp, "cube.obj")


def test_host_layer_compiles_and_links(tmp_path):
    exe = build_example(str(tmp_path))
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


@pytest.mark.parametrize("n_tex,seed", [(1, 0), (9, 1), (300, 2), (4000, 3)])
def test_page_planner_cpp_equals_python_mirror(tmp_path, n_tex, seed):
    """texture_context::alloc_gpu's page planner (texture_context.cpp:94-261) exists twice, in the C++ host layer and in the
    Python harness: both must produce the same nums[] / sizes[] for anything from one texture to thousands (SURVEY.md §8f
    rank 3: atlas build at scale). 4000 textures of 16..2048 px fill dozens of 2048^2 slices."""
    from openclrenderer_b200 import _build, scene
    lib = _build.build_product()
    libdir, cuda_lib = os.path.dirname(lib), "/usr/local/cuda/lib64"
    exe = str(tmp_path / "plan_pages")
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "examples", "plan_pages.cpp"), "-L" + libdir, "-lrr_b200",
                    "-L" + cuda_lib, "-lcudart", "-lz", "-Wl,-rpath," + libdir, "-Wl,-rpath," + cuda_lib], check=True)
    rng = np.random.Generator(np.random.PCG64(seed))
    dims = [int(2 ** rng.integers(4, 12)) for _ in range(n_tex)]           # 16 .. 2048, all have four mip levels >= 1
    if n_tex >= 9:
        dims[3] = 48                                                       # a non-power-of-two size: 48, 24, 12, 6, 3
    r = subprocess.run([exe], input=" ".join(map(str, dims)), capture_output=True, text=True, check=True)
    out = r.stdout.split("\n")
    n_slices, n_nums = (int(x) for x in out[0].split())
    nums = np.array(out[1].split(), dtype=np.uint64).astype(np.uint32)
    sizes = np.array(out[2].split(), dtype=np.uint64).astype(np.uint32)
    p_slices, p_nums, p_sizes, p_start = scene.plan_atlas(dims)
    assert n_slices == p_slices and n_nums == len(p_nums) == 5 * n_tex and p_start == n_tex
    assert np.array_equal(nums, p_nums) and np.array_equal(sizes, p_sizes)
    # every (slice, index) is used once and fits its page
    assert len(set(nums.tolist())) == len(nums)
    for v in nums[:: max(1, len(nums) // 500)]:
        sl, idx = int(v) >> 16, int(v) & 0xFFFF
        assert idx < (2048 // int(sizes[sl])) ** 2


def test_png_loader_roundtrip(tmp_path):
    """the host layer's zlib PNG decoder against Pillow on the reference's own texture"""
    write_obj_assets(str(tmp_path))
    src = r'''
#include "%s/openclrenderer_b200/host/rr_host.hpp"
int main(int, char** a){ std::vector<uint8_t> px; int w,h; if(!rrhost::load_png_rgba(a[1],px,w,h)) return 1;
 FILE* f=fopen(a[2],"wb"); fwrite(px.data(),1,px.size(),f); fclose(f); printf("%%d %%d\n",w,h); return 0; }''' % ROOT
    c = tmp_path / "png.cpp"
    c.write_text(src)
    exe = str(tmp_path / "png")
    from openclrenderer_b200 import _build
    libdir = os.path.dirname(_build.build_product())
    subprocess.run(["g++", "-std=c++17", "-o", exe, str(c), "-L" + libdir, "-lrr_b200", "-L/usr/local/cuda/lib64", "-lcudart", "-lz",
                    "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)
    out = str(tmp_path / "px.raw")
    r = subprocess.run([exe, str(tmp_path / "red.png"), out], capture_output=True, text=True, check=True)
    assert r.stdout.split() == ["1024", "1024"]
    from openclrenderer_b200 import scene
    assert np.array_equal(np.fromfile(out, np.uint8).reshape(1024, 1024, 4), scene.load_assets()["red_png"])


@pytest.mark.gpu
def test_main_headless_equals_python_driven_frame(tmp_path):
    """config 1 through the C++ host classes (obj_load, texture_context::alloc_gpu, object_context::build, light::build,
    engine::draw_bulk_objs_n) == the same frame driven from Python, bit for bit."""
    from openclrenderer_b200 import Renderer, scene
    exe = build_example(str(tmp_path))
    obj = write_obj_assets(str(tmp_path))
    out = str(tmp_path / "frame")
    r = subprocess.run([exe, obj, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "tris 12 objs 1" in r.stdout
    s = scene.scene_c1("A")
    g = Renderer(s.cfg)
    s.upload(g)
    s.render(g, frames=2)
    assert np.array_equal(np.fromfile(out + ".depth", np.uint32).reshape(600, 800), g.read_depth())
    assert np.array_equal(np.fromfile(out + ".ids", np.uint32).reshape(600, 800), g.read_ids())
    assert np.array_equal(np.fromfile(out + ".rgba", np.uint8).reshape(600, 800, 4), g.read_rgba8())
