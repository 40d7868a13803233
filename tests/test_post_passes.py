"""Post passes on the G-buffer: do_pseudo_aa (cl2.cl:6437-6657).
not-gpu: properties of the CPU restatement; gpu: the CUDA kernel against it through the C ABI.
Parity for this pass is UNPINNED against the reference itself: the reference runs it in place on one image
(engine.cpp:1854-1856), so its own output depends on scheduling; both sides here read kernel3's frame (canonical)."""
import numpy as np
import pytest

from openclrenderer_b200 import scene
from oracle.binding import Oracle


def _render(x, s, aa):
    s.upload(x)
    x.frame_shadows(1)
    x.frame_draw(s.c_pos, s.c_rot, s.clear)
    if aa:
        x.post_pseudo_aa()
    x.sync()
    return x.read_rgba8(), x.read_depth()


def test_oracle_pseudo_aa_touches_only_edges():
    s = scene.scene_c1("A")
    plain, depth = _render(Oracle(s.cfg, threads=0), s, False)
    aa, depth2 = _render(Oracle(s.cfg, threads=0), s, True)
    assert np.array_equal(depth, depth2)
    changed = (plain != aa).any(axis=-1)
    cov = depth != 0xFFFFFFFF
    assert 0 < changed.sum() < 0.05 * cov.sum(), "AA should touch a thin set of edge pixels"
    assert not changed[~cov].any(), "uncovered pixels are never written (cl2.cl:6513)"
    assert not changed[0].any() and not changed[-1].any() and not changed[:, 0].any() and not changed[:, -1].any()
    assert (aa[changed][:, 3] == 255).all()
    # every changed pixel has a depth step or a crease in its 3x3 neighbourhood: at least its colour neighbourhood is not flat
    ys, xs = np.nonzero(changed)
    for y, x in list(zip(ys, xs))[:200]:
        nb = plain[y - 1:y + 2, x - 1:x + 2, :3].reshape(-1, 3).astype(int)
        lo, hi = nb.min(axis=0), nb.max(axis=0)
        assert ((aa[y, x, :3].astype(int) >= lo - 1) & (aa[y, x, :3].astype(int) <= hi + 1)).all(), "result is a convex mix of the neighbours"


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["c1", "spheres"])
def test_pseudo_aa_matches_oracle(which):
    from openclrenderer_b200 import Renderer
    if which == "c1":
        s = scene.scene_c1("A")
    else:
        s = scene.scene_spheres(640, 384, n_spheres=12, grid=(4, 3), seed=11, n_lights=2, light_dim=128, tex_sizes=(128, 64))
    g_plain, _ = _render(Renderer(s.cfg), s, False)
    o_plain, _ = _render(Oracle(s.cfg, threads=0), s, False)
    g, gd = _render(Renderer(s.cfg), s, True)
    o, od = _render(Oracle(s.cfg, threads=0), s, True)
    assert np.array_equal(gd, od)
    assert ((g != g_plain).any(axis=-1)).sum() > 50, "the pass changed nothing"
    same_in = (g_plain == o_plain).all(axis=-1)
    # where the 3x3 input neighbourhoods are identical the outputs must be identical (the arithmetic is pinned)
    nb_same = same_in.copy()
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            nb_same &= np.roll(np.roll(same_in, dy, axis=0), dx, axis=1)
    assert np.array_equal(g[nb_same], o[nb_same])
    d = np.abs(g.astype(int) - o.astype(int)).max(axis=-1)
    assert (d <= 1).mean() >= 0.999 and d.max() <= 2


@pytest.mark.gpu
def test_pseudo_aa_rejected_on_a_split_frame():
    from openclrenderer_b200 import Renderer, RRError
    s = scene.scene_c1("A")
    r = Renderer(s.cfg.copy(band_y0=0, band_y1=300))
    s.upload(r)
    r.frame_draw(s.c_pos, s.c_rot, s.clear)
    with pytest.raises(RRError):
        r.post_pseudo_aa()
