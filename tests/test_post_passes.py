"""Post passes on the G-buffer: do_pseudo_aa (cl2.cl:6437-6657), do_motion_blur (6714-6860),
screenspace_godrays (1792-1917).
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


# ---- do_motion_blur (cl2.cl:6714-6860) and screenspace_godrays (cl2.cl:1792-1917) -----------------------------------------
def _moving_frames(x, s, n_frames, blur=True, godray=False):
    """n frames with a moving camera and two moving objects (object::g_flush patches of world_pos, byte offset 0); returns the
    colour of every frame after the post pass and the descriptors at the end (the motion history lives in them)."""
    s.upload(x)
    if godray:
        lights = s.lights.copy()
        lights["godray_intensity"][0] = 1.0
        x.lights_write(lights)
    out = []
    for i in range(n_frames):
        for oid in (0, len(s.objs) - 1):
            pos = np.array(s.objs[oid]["world_pos"], dtype=np.float32).copy()
            pos[0] += 40.0 * i * (1 if oid == 0 else -1)
            x.scene_patch_obj(oid, 0, pos.tobytes())
        c_pos = (s.c_pos[0] + 60.0 * i, s.c_pos[1], s.c_pos[2] + 25.0 * i)
        c_rot = (s.c_rot[0], s.c_rot[1] + 0.015 * i, s.c_rot[2])
        x.frame_shadows(1 if i == 0 else 0)
        x.frame_draw(c_pos, c_rot, s.clear)
        if blur:
            x.post_motion_blur(1.0, 1.0)
        if godray:
            x.post_godrays()
        x.sync()
        out.append(x.read_rgba8())
        x.swap_buffers()
    return out, x.scene_read_objs(0, len(s.objs))


def _post_scene():
    return scene.scene_c2(640, 360, light_dim=256)         # cylinder + ground quad: two objects, most of the screen covered


def test_oracle_motion_blur_properties():
    s = _post_scene()
    plain, _ = _moving_frames(Oracle(s.cfg, threads=0), s, 3, blur=False)
    blurred, objs = _moving_frames(Oracle(s.cfg, threads=0), s, 3, blur=True)
    # frame 0: the previous camera is the zero camera of a fresh context and the history equals the placement -> large vectors, but
    # every output is an average of in-image samples: inside the input's range; uncovered pixels never change
    for a, b in zip(plain, blurred):
        assert b.min() >= a.min() and b.max() <= a.max()
    changed = (plain[2] != blurred[2]).any(axis=-1)
    assert changed.sum() > 100, "moving camera + moving objects must blur something"
    clear = (plain[2] == plain[2][0, 0]).all(axis=-1)
    # the history slots alternate with frame parity (cl2.cl:6768-6784): after 3 frames (ids 1, 2, 3) slot 1 holds frame 3's placement
    # of every object that was visible, slot 2 frame 2's
    moved = objs[0]
    assert np.allclose(moved["old_world_pos_1"][:3], moved["world_pos"][:3])
    assert abs(moved["old_world_pos_2"][0] - (moved["world_pos"][0] - 40.0)) < 1e-3
    assert clear.sum() > 0


def test_oracle_godrays_properties():
    s = _post_scene()
    plain, _ = _moving_frames(Oracle(s.cfg, threads=0), s, 1, blur=False)
    rays, _ = _moving_frames(Oracle(s.cfg, threads=0), s, 1, blur=False, godray=True)
    a, b = plain[0].astype(int), rays[0].astype(int)
    assert (b[..., 3] == 252).all(), "alpha = 1 * exposure 0.99 -> 252"
    # rays only add light on top of the (0.99-scaled, quarter-pixel resampled) frame
    assert (b[..., :3] - a[..., :3]).max() > 3, "a light with godray_intensity 1 must brighten some pixels"


@pytest.mark.gpu
def test_motion_blur_matches_oracle():
    from openclrenderer_b200 import Renderer
    s = _post_scene()
    g, gobjs = _moving_frames(Renderer(s.cfg), s, 4, blur=True)
    o, oobjs = _moving_frames(Oracle(s.cfg, threads=0), s, 4, blur=True)
    for i, (a, b) in enumerate(zip(g, o)):
        d = np.abs(a.astype(int) - b.astype(int)).max(axis=-1)
        assert (d <= 1).mean() >= 0.999 and d.max() <= 2, f"frame {i}: {(d <= 1).mean()} within 1 LSB, max {d.max()}"
    assert gobjs.tobytes() == oobjs.tobytes(), "motion history in the descriptors differs"
    plain, _ = _moving_frames(Renderer(s.cfg), s, 4, blur=False)
    assert ((plain[3] != g[3]).any(axis=-1)).sum() > 100


@pytest.mark.gpu
def test_godrays_match_oracle():
    from openclrenderer_b200 import Renderer
    s = _post_scene()
    g, _ = _moving_frames(Renderer(s.cfg), s, 2, blur=False, godray=True)
    o, _ = _moving_frames(Oracle(s.cfg, threads=0), s, 2, blur=False, godray=True)
    for i, (a, b) in enumerate(zip(g, o)):
        d = np.abs(a.astype(int) - b.astype(int)).max(axis=-1)
        assert (d <= 1).mean() >= 0.999 and d.max() <= 2, f"frame {i}: {(d <= 1).mean()} within 1 LSB, max {d.max()}"
    plain, _ = _moving_frames(Renderer(s.cfg), s, 2, blur=False)
    assert (g[1].astype(int)[..., :3] - plain[1].astype(int)[..., :3]).max() > 3


@pytest.mark.gpu
def test_blur_then_aa_chain_matches_oracle():
    """the passes compose: each reads the previous one's output (engine.cpp's call order: godrays, motion blur, pseudo AA)"""
    from openclrenderer_b200 import Renderer
    s = _post_scene()
    res = []
    for x in (Renderer(s.cfg), Oracle(s.cfg, threads=0)):
        s.upload(x)
        lights = s.lights.copy()
        lights["godray_intensity"][0] = 0.7
        x.lights_write(lights)
        for i in range(2):
            x.frame_shadows(1 if i == 0 else 0)
            x.frame_draw((s.c_pos[0] + 20.0 * i, s.c_pos[1], s.c_pos[2]), s.c_rot, s.clear)
            x.post_godrays()
            x.post_motion_blur(0.8, 0.5)
            x.post_pseudo_aa()
            x.sync()
            img = x.read_rgba8()
            x.swap_buffers()
        res.append(img)
    d = np.abs(res[0].astype(int) - res[1].astype(int)).max(axis=-1)
    assert (d <= 1).mean() >= 0.998 and d.max() <= 3, f"{(d <= 1).mean()} within 1 LSB, max {d.max()}"
