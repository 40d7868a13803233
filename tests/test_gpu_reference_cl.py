"""-m gpu: the reference's own kernels (unmodified cl2.cl through the NVIDIA OpenCL ICD on the same B200) against the
CPU oracle and against the CUDA product, plus the committed golden vectors against the CUDA product."""
import os

import numpy as np
import pytest

from openclrenderer_b200 import Renderer
from oracle.binding import Oracle
from tests.golden.make_golden import SCENES, tri_ids
from tests.test_oracle_golden import GOLD, check_against_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    z = np.load(GOLD)
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name,tri_floor", [("c1A", 0.999), ("c1B", 0.999), ("c2", 0.99), ("c2_small", 0.99), ("sph", 0.95), ("sph_close", 0.95)])
def test_cuda_matches_reference_golden(gold, name, tri_floor):
    s = SCENES[name]()
    g = Renderer(s.cfg)
    s.upload(g)
    s.render(g, frames=2)
    check_against_golden(g, gold, name, tri_floor)


def _refcl():
    from oracle import ref_opencl
    if not ref_opencl.available():
        pytest.skip("NVIDIA OpenCL ICD or oracle/_ref/cl2.cl.gz not available on this box")
    return ref_opencl


@pytest.mark.parametrize("name", ["c1A", "c2_small", "sph", "c2", "sph_close"])
def test_live_reference_pinned_equals_oracle_and_cuda(name):
    """pinned arithmetic: reference == oracle == CUDA, bit for bit, on depth, shadow cubemaps, atlas and fragment multiset."""
    cl = _refcl()
    s = SCENES[name]()
    r, o, g = cl.RefCL(s.cfg, mode="pinned"), Oracle(s.cfg, threads=0), Renderer(s.cfg)
    for x in (r, o, g):
        s.upload(x)
    ra = r.atlas_read_raw()
    assert np.array_equal(ra, o.atlas_read_raw()) and np.array_equal(ra, g.atlas_read_raw())
    for x in (r, o, g):
        s.render(x, frames=2)
    rd = r.read_depth()
    assert np.array_equal(rd, o.read_depth()) and np.array_equal(rd, g.read_depth())
    for k in range(r.n_shadow):
        rs = r.read_shadow(0, k)
        assert np.array_equal(rs, o.read_shadow(0, k)) and np.array_equal(rs, g.read_shadow(0, k))
    for k in range(r.n_static):
        rs = r.read_shadow(1, k)
        assert np.array_equal(rs, o.read_shadow(1, k)) and np.array_equal(rs, g.read_shadow(1, k))

    def canon(fr):
        return fr[np.lexsort((fr[:, 1], fr[:, 0]))][:, [0, 1, 3, 4]]
    assert np.array_equal(canon(r.read_fragments()), canon(g.read_fragments()))
    cov = rd != 0xFFFFFFFF
    same = (tri_ids(r) == tri_ids(g)) & cov
    assert same.sum() / cov.sum() > 0.95
    diff = np.abs(r.read_rgba8().astype(np.int16) - g.read_rgba8().astype(np.int16)).max(axis=-1)
    assert (diff <= 1).mean() >= 0.999
    assert diff[same | ~cov].max() <= 2


@pytest.mark.parametrize("name", ["c1A", "c2_small", "c2", "sph", "sph_close"])
def test_live_reference_as_shipped_is_close(name):
    """as shipped (-cl-fast-relaxed-math, FP_CONTRACT ON, native_* intrinsics): same coverage and fragments up to a handful
    of boundary pixels, depth within the reference's own +-20 tolerance almost everywhere, colours within +-2 LSB on >= 99.9 %."""
    cl = _refcl()
    s = SCENES[name]()
    r, g = cl.RefCL(s.cfg, mode="shipped"), Renderer(s.cfg)
    for x in (r, g):
        s.upload(x)
        s.render(x, frames=2)
    rd, gd = r.read_depth(), g.read_depth()
    rc, gc = rd != 0xFFFFFFFF, gd != 0xFFFFFFFF
    assert (rc != gc).sum() <= 1e-4 * gc.sum() + 8
    assert abs(len(r.read_fragments()) - len(g.read_fragments())) <= 8
    diff = np.abs(r.read_rgba8().astype(np.int16) - g.read_rgba8().astype(np.int16)).max(axis=-1)
    assert (diff <= 2).mean() >= 0.999


@pytest.mark.parametrize("which", ["pseudo_aa", "motion_blur", "godrays"])
def test_live_reference_post_passes(which):
    """The post passes pinned against the reference's own kernels (do_pseudo_aa cl2.cl:6437, do_motion_blur 6714,
    screenspace_godrays 1792) run with separate input and output images. The reference filters a float image with the GPU's
    fixed-point CLK_FILTER_LINEAR weights, the product the quantised RGBA8 target with the specification's formula: +-2 LSB."""
    from openclrenderer_b200 import scene
    from tests.test_post_passes import _moving_frames
    cl = _refcl()
    s = scene.scene_c2(640, 360, light_dim=256)
    out = {}
    for key, x in (("ref", cl.RefCL(s.cfg, mode="pinned")), ("cuda", Renderer(s.cfg))):
        if which == "pseudo_aa":
            s.upload(x)
            x.frame_shadows(1)
            x.frame_draw(s.c_pos, s.c_rot, s.clear)
            x.post_pseudo_aa()
            x.sync()
            out[key] = ([x.read_rgba8()], None)
        else:
            out[key] = _moving_frames(x, s, 3, blur=(which == "motion_blur"), godray=(which == "godrays"))
    (rf, robjs), (gf, gobjs) = out["ref"], out["cuda"]
    for i, (a, b) in enumerate(zip(rf, gf)):
        d = np.abs(a.astype(np.int16) - b.astype(np.int16))[..., :3].max(axis=-1)
        assert (d <= 2).mean() >= 0.995, f"{which} frame {i}: only {(d <= 2).mean() * 100:.3f}% within +-2 LSB (max {d.max()})"
        assert (d <= 1).mean() >= 0.97, f"{which} frame {i}: only {(d <= 1).mean() * 100:.3f}% within +-1 LSB"
    if which == "motion_blur":
        for f in ("old_world_pos_1", "old_world_pos_2", "old_world_rot_quat_1", "old_world_rot_quat_2"):
            assert np.array_equal(robjs[f][:, :3], gobjs[f][:, :3]), f
