"""-m gpu: CUDA product vs CPU oracle through the C ABI, on the configurations of BASELINE.json at sizes the oracle
finishes in seconds, plus size-independent properties at full size."""
import numpy as np
import pytest

from openclrenderer_b200 import Renderer, scene
from openclrenderer_b200._abi import Config, TRIANGLE, OBJ_DESC, LIGHT, FEATURE_TWO_SIDED, FEATURE_IS_STATIC, FEATURE_NO_DYNAMIC_SHADOWS
from oracle.binding import Oracle
from tests.parity import render_both, assert_frame_parity, colour_stats

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("profile", ["A", "B"])
def test_c1_cube(profile):
    s = scene.scene_c1(profile)
    g, o = render_both(s, frames=2)
    st = assert_frame_parity(g, o, label=f"c1/{profile}")
    assert st["covered"] > 10000


def test_c1_atlas_bit_exact():
    s = scene.scene_c1()
    g, o = Renderer(s.cfg), Oracle(s.cfg)
    s.upload(g), s.upload(o)
    assert np.array_equal(g.atlas_read_raw(), o.atlas_read_raw())


def test_c2_cylinder_shadowed():
    s = scene.scene_c2()
    g, o = render_both(s, frames=2, threads=0)
    assert np.array_equal(g.read_shadow(0, 0), o.read_shadow(0, 0)), "shadow cubemap differs"
    st = assert_frame_parity(g, o, label="c2")
    assert st["covered"] > 100000


def test_c2_atlas_bit_exact():
    s = scene.scene_c2(640, 360, 256)
    g, o = Renderer(s.cfg), Oracle(s.cfg)
    s.upload(g), s.upload(o)
    assert np.array_equal(g.atlas_read_raw(), o.atlas_read_raw())


def test_spheres_small_all_features():
    """scaled-down config 3: 24 spheres (48k triangles), multi-slice atlas with mips, 4 shadow lights, 960x540."""
    s = scene.scene_spheres(960, 540, n_spheres=24, grid=(6, 4), seed=7, n_lights=4, light_dim=256, tex_sizes=(256, 128, 64, 64))
    # make the scene exercise the feature flags: two-sided, static (skipped by dynamic shadow passes), no-dynamic-shadow receivers
    s.objs["feature_flag"][::5] |= FEATURE_TWO_SIDED
    s.objs["feature_flag"][1::7] |= FEATURE_IS_STATIC
    s.objs["feature_flag"][2::9] |= FEATURE_NO_DYNAMIC_SHADOWS
    s.lights["is_static"][1] = 1
    g, o = render_both(s, frames=2, threads=0)
    for k in range(4):
        assert np.array_equal(g.read_shadow(0, k), o.read_shadow(0, k)), f"dynamic cubemap {k} differs"
    assert np.array_equal(g.read_shadow(1, 0), o.read_shadow(1, 0)), "static cubemap differs"
    assert_frame_parity(g, o, label="spheres24")


def _soup(seed, n, w, h, big=False):
    """random triangle soup in front of (and through) the camera: near-plane clipping, screen-edge clamping, huge bboxes."""
    rng = np.random.Generator(np.random.PCG64(seed))
    tris = np.zeros(n, dtype=TRIANGLE)
    c = rng.uniform([-600, -400, -100], [600, 400, 1500], size=(n, 1, 3))
    ext = rng.uniform(5, 900 if big else 120, size=(n, 1, 1))
    tris["vertices"]["pos"][:, :, :3] = (c + rng.normal(size=(n, 3, 3)) * ext).astype(np.float32)
    nrm = rng.normal(size=(n, 3, 3))
    tris["vertices"]["normal"][:, :, :3] = (nrm / np.linalg.norm(nrm, axis=-1, keepdims=True)).astype(np.float32)
    tris["vertices"]["vt"] = rng.uniform(-1.5, 2.5, size=(n, 3, 2)).astype(np.float32)
    col = rng.integers(1, 2 ** 32, size=(n, 1), dtype=np.uint64).astype(np.uint32)
    tris["vertices"]["vertex_col"] = np.where(rng.uniform(size=(n, 1)) < 0.3, col, 0)          # 30 % vertex-coloured
    tris["vertices"]["object_id"][:, 0] = rng.integers(0, 3, size=n)
    objs = np.array([scene.make_obj_desc(pos=(0, 0, 200), scale=1.0, tid=0, feature_flag=FEATURE_TWO_SIDED),
                     scene.make_obj_desc(pos=(50, -20, 300), quat=(0.1, 0.7, 0.2, 0.6), scale=0.8, tid=1),
                     scene.make_obj_desc(pos=(-80, 60, 500), quat=(-0.3, 0.2, 0.9, 0.1), scale=1.7, tid=0, specular=0.5)], dtype=OBJ_DESC)
    lights = np.array([scene.make_light((300, -500, -200), shadow=1), scene.make_light((-400, 300, 100), col=(0.9, 0.8, 1.0), shadow=0)], dtype=LIGHT)
    cfg = Config.default(w, h, light_dim=128)
    tex = [scene.procedural_texture(64, 5), scene.procedural_texture(128, 6)]
    return scene.Scene(cfg, tris, objs, lights, tex, c_pos=(10, -20, -150), c_rot=(0.1, -0.2, 0.05), clear=(0.1, 0.2, 0.3, 1.0), name=f"soup{seed}")


@pytest.mark.parametrize("seed,big", [(1, False), (2, True), (3, True)])
def test_triangle_soup(seed, big):
    s = _soup(seed, 600 if big else 3000, 640, 360, big)
    g, o = render_both(s, frames=2, threads=0)
    assert np.array_equal(g.read_shadow(0, 0), o.read_shadow(0, 0))
    assert_frame_parity(g, o, label=s.name)


def test_odd_resolution_and_empty_scene():
    s = _soup(4, 500, 333, 217, True)
    g, o = render_both(s, frames=1, threads=0)
    assert_frame_parity(g, o, label="odd")
    # empty scene: draw is a no-op (engine.cpp:1806), buffers stay cleared
    cfg = Config.default(64, 48)
    g2 = Renderer(cfg)
    g2.scene_alloc(0, 0)
    g2.lights_write(np.zeros(0, dtype=LIGHT))
    g2.frame_draw((0, 0, 0), (0, 0, 0))
    g2.sync()
    assert (g2.read_depth() == 0xFFFFFFFF).all()


def _tri_of(r):
    ids, fr, d = r.read_ids(), r.read_fragments(), r.read_depth()
    t = np.full(ids.shape, -1, np.int64)
    ok = (d != 0xFFFFFFFF) & (ids < len(fr))
    t[ok] = fr[ids[ok], 0]
    return t


@pytest.mark.parametrize("halo", [-1, 24])
def test_band_split_equals_full_frame(halo):
    """sort-first bands: compositing the bands of 4 contexts == the single-context frame, bit for bit in
    depth and colour; ids compared as triangle ids (each context numbers only the fragments of the objects it set up).
    halo -1 rasterises every row on every context; halo 24 (>= this scene's SSAO reach) culls objects outside band+halo."""
    s = scene.scene_spheres(640, 384, n_spheres=12, grid=(4, 3), seed=11, n_lights=2, light_dim=128, tex_sizes=(128, 64))
    full = Renderer(s.cfg)
    s.upload(full)
    s.render(full, frames=2)
    fd, ft, fc = full.read_depth(), _tri_of(full), full.read_rgba8()
    n_full = len(full.read_fragments())
    n = 4
    culled = 0
    for k in range(n):
        y0, y1 = k * 96, (k + 1) * 96
        cfg = s.cfg.copy(band_y0=y0, band_y1=y1, band_halo=halo, face_rank=0, face_world=0)
        b = Renderer(cfg)
        s.upload(b)
        s.render(b, frames=2)
        assert np.array_equal(b.read_depth()[y0:y1], fd[y0:y1])
        assert np.array_equal(_tri_of(b)[y0:y1], ft[y0:y1])
        assert np.array_equal(b.read_rgba8()[y0:y1], fc[y0:y1])
        culled += int(len(b.read_fragments()) < n_full)
    if halo >= 0:
        assert culled >= 2, "object culling never kicked in"


def test_face_sharding_union_equals_full_cubemap():
    """shadow (light, face) pairs split over 3 contexts: the element-wise min of their slabs == the full cubemap."""
    s = scene.scene_spheres(320, 200, n_spheres=12, grid=(4, 3), seed=13, n_lights=2, light_dim=128, tex_sizes=(64,))
    full = Renderer(s.cfg)
    s.upload(full)
    full.frame_shadows(1)
    full.sync()
    ref = [full.read_shadow(0, k) for k in range(2)]
    acc = [np.full_like(ref[0], 0xFFFFFFFF) for _ in range(2)]
    for rank in range(3):
        b = Renderer(s.cfg.copy(face_rank=rank, face_world=3))
        s.upload(b)
        b.frame_shadows(1)
        b.sync()
        for k in range(2):
            part = b.read_shadow(0, k)
            owned = [(k * 6 + f) // 4 == rank for f in range(6)]      # 12 pairs over 3 contexts -> chunks of 4
            for f in range(6):
                if not owned[f]:
                    assert (part[f] == 0xFFFFFFFF).all()
            acc[k] = np.minimum(acc[k], part)
    for k in range(2):
        assert np.array_equal(acc[k], ref[k])


@pytest.mark.parametrize("rot_y,rot_x", [(1.3, 0.1), (-2.4, 0.3), (0.4, -0.6)])
def test_cluster_culling_keeps_records_exact(rot_y, rot_x):
    """camera inside the sphere field looking sideways / backwards / up: most 128-triangle clusters are off screen and are
    skipped by k_cluster_vis; depth, ids, fragment records (incl. the c_id numbering with its holes, cl2.cl:4342) and colours
    must still equal the oracle's, which tests every triangle."""
    s = scene.scene_spheres(640, 384, n_spheres=24, grid=(6, 4), seed=17, n_lights=2, light_dim=128, tex_sizes=(128, 64))
    s.c_pos = (s.c_pos[0] + 150.0, s.c_pos[1] - 300.0, s.c_pos[2] + 2500.0)
    s.c_rot = (rot_x, rot_y, 0.0)
    s.cfg = s.cfg.copy(cluster_cull=1)              # single context: culling is opt-in (it is on by default only when the frame is split)
    g, o = render_both(s, frames=2)
    assert_frame_parity(g, o, label=f"culling rot_y={rot_y}")


def test_async_rebuild_flips_between_scenes():
    """object_context::build(async) + flip_buffers (object_context.cpp:520-797): a new scene is uploaded
    into the back buffers while frames keep coming from the old one; after the commit the frame equals the oracle's (and a
    fresh context's) render of the new scene. Covers a smaller scene (buffers reused), a larger one (buffers grow) and the
    way back."""
    kw = dict(n_lights=2, light_dim=128, tex_sizes=(128, 64))
    sA = scene.scene_spheres(640, 360, n_spheres=8, grid=(4, 2), seed=5, **kw)
    sB = scene.scene_spheres(640, 360, n_spheres=5, grid=(3, 2), seed=6, **kw)
    sC = scene.scene_spheres(640, 360, n_spheres=14, grid=(5, 3), seed=7, **kw)
    g, o = Renderer(sA.cfg), Oracle(sA.cfg, threads=0)
    for x in (g, o):
        sA.upload(x)

    def frame(label, dirty=0):
        for x in (g, o):
            x.frame_shadows(dirty)
            x.frame_draw(sA.c_pos, sA.c_rot, sA.clear)
            x.sync()
        st = assert_frame_parity(g, o, label=label)
        col = g.read_rgba8()
        for x in (g, o):
            x.swap_buffers()
        return col, st

    frame("A", 1)
    cur = sA
    for nxt, name in ((sB, "B"), (sC, "C"), (sA, "A again")):
        for x in (g, o):
            x.scene_build(nxt.tris, nxt.objs, commit=False)
        before, _ = frame(f"still {cur.name} while {name} uploads")      # the old scene keeps rendering during the upload
        for x in (g, o):
            x.scene_build_commit()
        after, st = frame(name)
        assert st["covered"] > 50
        assert not np.array_equal(before, after)
        fresh = Renderer(sA.cfg)                                         # same atlas / lights, the new geometry
        sA.upload(fresh)
        fresh.scene_alloc(len(nxt.tris), len(nxt.objs))
        fresh.scene_write_objs(nxt.objs)
        fresh.scene_write_tris(nxt.tris)
        fresh.frame_shadows(0)
        fresh.frame_draw(sA.c_pos, sA.c_rot, sA.clear)
        fresh.sync()
        assert np.array_equal(fresh.read_rgba8(), after), f"{name}: rebuilt context differs from a fresh one"
        fresh.close()
        cur = nxt
    assert g.scene_build_ready() is False                                # nothing in flight


@pytest.mark.parametrize("cap", [1, 5000])
def test_sample_list_overflow_falls_back_to_walking(cap, monkeypatch):
    """kernel2 normally streams the samples the depth kernels recorded; fragments whose samples did not fit the list (depth complexity
    above 4 on the whole screen — forced here with a tiny list) are walked again by the tail of k_ids_list: same ids either way."""
    monkeypatch.setenv("RR_SAMPLE_CAP", str(cap))
    for s in (scene.scene_c2(640, 360, 256), _soup(11, 300, 320, 200, big=True)):
        g, o = render_both(s, frames=2, threads=0)
        assert_frame_parity(g, o, label=f"{s.name} sample cap {cap}")


@pytest.mark.parametrize("knobs", [{"RR_SPLIT_CLEAR": "0"}, {"RR_LIST_FROM_IDS": "0"}, {"RR_CLEAR_AT": "1", "RR_CLEAR_GRID": "1"}])
def test_alternative_frame_paths_render_the_same_frame(knobs, monkeypatch):
    """the A/B knobs of INTEGRATION.md §6 select other arrangements of the same work — kernel3's streaming stores inside
    k_shade_pre4 instead of on the side stream, the covered-pixel list from a pass over the screen instead of from the id resolve,
    the side stream forked behind the setup kernel: same frame, two frames in a row (the clears of frame 1 are frame 2's start)."""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    for s in (scene.scene_c2(640, 360, 256), _soup(5, 400, 322, 200)):          # (322: a width that is not a multiple of 4 takes k_shade_pre)
        g, o = render_both(s, frames=3, threads=0)
        assert_frame_parity(g, o, label=f"{s.name} {knobs}")


def test_second_draw_without_swap_still_shades_every_covered_pixel():
    """rr_frame_draw twice without rr_swap_buffers (the id image of the second draw is not empty): the covered-pixel list must not
    rely on 'first sample to resolve a pixel' then — every pixel with a depth is shaded, as in a frame drawn once."""
    s = scene.scene_c1("A")
    g = Renderer(s.cfg)
    s.upload(g)
    g.frame_shadows(1)
    g.frame_draw(s.c_pos, s.c_rot, s.clear)
    g.sync()
    once = g.read_rgba8()
    g.frame_draw(s.c_pos, s.c_rot, s.clear)                              # same camera, no swap: same depth, same winners
    g.sync()
    assert np.array_equal(g.read_rgba8(), once)


def test_overflow_is_reported():
    s = _soup(5, 400, 320, 200, True)
    cfg = s.cfg.copy(max_fragments=64)
    g = Renderer(cfg)
    s.upload(g)
    g.frame_draw(s.c_pos, s.c_rot)
    from openclrenderer_b200 import RRError
    with pytest.raises(RRError) as e:
        g.sync()
    assert e.value.code == -4


def test_full_size_properties_c3():
    """config 3 at full size (1 M triangles, 3840x2160, 4 shadow lights): properties that need no oracle run.
    determinism (two contexts, identical buffers), idempotence (re-rendering the same frame gives the same frame),
    id/depth consistency (every covered pixel's id names a fragment whose triangle covers that pixel at that depth +-20)."""
    s = scene.scene_c3()
    a = Renderer(s.cfg)
    s.upload(a)
    s.render(a, frames=2)
    d1, i1, c1 = a.read_depth(), a.read_ids(), a.read_rgba8()
    a.swap_buffers()
    a.frame_shadows(0)
    a.frame_draw(s.c_pos, s.c_rot, s.clear)
    a.sync()
    assert np.array_equal(a.read_depth(), d1) and np.array_equal(a.read_ids(), i1) and np.array_equal(a.read_rgba8(), c1)
    b = Renderer(s.cfg)
    s.upload(b)
    s.render(b, frames=1)
    assert np.array_equal(b.read_depth(), d1) and np.array_equal(b.read_ids(), i1) and np.array_equal(b.read_rgba8(), c1)
    cov = d1 != 0xFFFFFFFF
    assert cov.sum() > 100000
    frags = a.read_fragments()
    assert i1[cov].max() < len(frags)
    # fragment records are prefix sums in triangle order: triangle ids non-decreasing, chunk numbers restart at 0
    assert (np.diff(frags[:, 0].astype(np.int64)) >= 0).all()
    assert (np.diff(frags[:, 2].astype(np.int64)) >= 0).all()
    t = a.timings()
    assert t["overflow"] == 0 and t["n_fragments"] == len(frags)


@pytest.mark.parametrize("depth", [2, 3, 4])
def test_frame_e2e_pipelined_readback_matches_plain_frames(depth):
    """rr_frame_e2e (descriptor upload + shadows + draw + pipelined read-back into a ring of pinned buffers) returns the
    same frames as frame_draw + read_rgba8, for a moving camera; ring depth D: the buffer passed D-1 calls ago is complete."""
    from openclrenderer_b200 import rr
    s = scene.scene_spheres(640, 360, n_spheres=8, grid=(4, 2), seed=21, n_lights=2, light_dim=128, tex_sizes=(128, 64))
    cams = [((s.c_pos[0] + 40.0 * i, s.c_pos[1], s.c_pos[2]), (s.c_rot[0] + 0.01 * i, 0.0, 0.0)) for i in range(7)]
    a = Renderer(s.cfg)
    s.upload(a)
    want = []
    for i, (p, r) in enumerate(cams):
        a.frame_shadows(1 if i == 0 else 0)
        a.frame_draw(p, r, s.clear)
        a.sync()
        want.append(a.read_rgba8())
        a.swap_buffers()
    b = Renderer(s.cfg)
    s.upload(b)
    b.frame_shadows(1)
    b.set_pipeline_depth(depth)
    bufs = [rr.host_alloc((s.cfg.height, s.cfg.width, 4)) for _ in range(depth)]
    got = []
    for i, (p, r) in enumerate(cams):
        b.frame_e2e(p, r, s.clear, 1, bufs[i % depth])
        if i >= depth - 1:
            got.append(bufs[(i - depth + 1) % depth].copy())          # the buffer passed depth-1 calls ago is complete on return
    b.sync()
    for i in range(len(cams) - depth + 1, len(cams)):
        got.append(bufs[i % depth].copy())
    for i in range(len(cams)):
        assert np.array_equal(got[i], want[i]), f"frame {i}"


@pytest.mark.parametrize("depth,size", [(2, (640, 360)), (3, (644, 363)), (4, (640, 360))])
def test_frame_e2e_dirty_tile_readback_is_bit_identical(depth, size):
    """rr_set_readback_tiles: only the 32x4-pixel tiles that can differ from what the host buffer holds cross the PCIe link, and the
    host buffers are still the frames bit for bit — camera moving by whole spheres (tiles covered D frames ago must be rewritten with
    the clear colour), an odd frame size (partial tiles at the right and bottom edges), a changed clear colour and a foreign host
    buffer in the middle of the run (both force a full copy), and fewer bytes than full copies on the way."""
    from openclrenderer_b200 import rr
    W, H = size
    s = scene.scene_spheres(W, H, n_spheres=8, grid=(4, 2), seed=21, n_lights=2, light_dim=128, tex_sizes=(128, 64))
    n = 11
    cams = [((s.c_pos[0] + 300.0 * (i % 5), s.c_pos[1] + 50.0 * i, s.c_pos[2]), (s.c_rot[0] + 0.02 * i, 0.0, 0.0)) for i in range(n)]
    clears = [s.clear if i < 6 else (0.3, 0.1, 0.2, 1.0) for i in range(n)]
    a = Renderer(s.cfg)
    s.upload(a)
    want = []
    for i, (p, r) in enumerate(cams):
        a.frame_shadows(1 if i == 0 else 0)
        a.frame_draw(p, r, clears[i])
        a.sync()
        want.append(a.read_rgba8())
        a.swap_buffers()
    b = Renderer(s.cfg)
    s.upload(b)
    b.frame_shadows(1)
    b.set_pipeline_depth(depth)
    b.set_readback_tiles(True)
    bufs = [rr.host_alloc((H, W, 4)) for _ in range(depth)]
    foreign = rr.host_alloc((H, W, 4))
    foreign[:] = 77                                                     # a buffer the library has never written: must come back complete
    used = []
    for i, (p, r) in enumerate(cams):
        buf = foreign if i == 8 else bufs[i % depth]
        b.frame_e2e(p, r, clears[i], 1, buf)
        used.append(buf)
        if i >= depth - 1:
            j = i - depth + 1
            assert np.array_equal(used[j], want[j]), f"frame {j} (complete on return of call {i})"
    b.sync()
    for j in range(n - depth + 1, n):
        assert np.array_equal(used[j], want[j]), f"frame {j}"
    sent = b.readback_tile_bytes()
    assert 0 < sent < n * W * H * 4, f"{sent} bytes for {n} frames of {W * H * 4}"      # (first uses, the new clear colour and the foreign buffer are whole frames)
    # the same host buffer for every call (against the contract, but it must not produce a stale frame): whole frames again
    b.readback_tile_bytes()
    for i in range(4):
        b.frame_e2e(cams[i][0], cams[i][1], clears[i], 1, bufs[0])
        b.sync()
        assert np.array_equal(bufs[0], want[i]), f"single buffer, frame {i}"
    # switching it off goes back to plain copies of whole frames
    b.set_readback_tiles(False)
    b.readback_tile_bytes()
    b.frame_e2e(cams[0][0], cams[0][1], clears[0], 1, bufs[0])
    b.sync()
    assert np.array_equal(bufs[0], want[0]) and b.readback_tile_bytes() == 0


def test_object_transform_patches_match_oracle():
    """object::g_flush (object.cpp:652-857): 12/16-byte writes of world_pos / world_rot_quat / scale into the descriptor
    array between frames; both sides re-render the moved scene identically."""
    s = scene.scene_spheres(640, 360, n_spheres=6, grid=(3, 2), seed=5, n_lights=2, light_dim=128, tex_sizes=(64, 32))
    g, o = Renderer(s.cfg), Oracle(s.cfg, threads=0)
    for x in (g, o):
        s.upload(x)
        s.render(x, frames=1)
    rng = np.random.Generator(np.random.PCG64(9))
    for step in range(3):
        for x in (g, o):
            x.swap_buffers()
        for oid in range(len(s.objs)):
            pos = (s.objs["world_pos"][oid] + np.array([rng.uniform(-300, 300), rng.uniform(-100, 100), rng.uniform(-200, 400), 0], np.float32)).astype(np.float32)
            q = rng.normal(size=4).astype(np.float32)
            sc = np.float32(s.objs["scale"][oid] * rng.uniform(0.7, 1.4))
            for x in (g, o):
                x.scene_patch_obj(oid, 0, pos.tobytes())
                x.scene_patch_obj(oid, 16, q.tobytes())
                x.scene_patch_obj(oid, OBJ_DESC.fields["scale"][1], sc.tobytes())
        for x in (g, o):
            x.frame_shadows(0)
            x.frame_draw(s.c_pos, s.c_rot, s.clear)
            x.sync()
        for k in range(2):
            assert np.array_equal(g.read_shadow(0, k), o.read_shadow(0, k))
        assert_frame_parity(g, o, label=f"patched step {step}")


def _micro_soup(seed, n, w, h, light_dim):
    """thousands of tiny triangles (0.2 .. 8 px) placed in SCREEN space over the whole frame, its borders and corners, a band of
    them in the top rows (where the walk's float row counter can lag, cl2.cl:5076) and some sub-pixel ones that round to
    collinear vertices: the cases the exact integer rasteriser of the setup kernels special-cases, against the literal walk."""
    rng = np.random.Generator(np.random.PCG64(seed))
    cfg = Config.default(w, h, light_dim=light_dim)
    fov = scene.fov_for(cfg)
    cx = rng.uniform(-12, w + 12, size=n)
    cy = rng.uniform(-12, h + 12, size=n)
    top = rng.uniform(size=n) < 0.25
    cy[top] = rng.uniform(-4, 60, size=top.sum())
    left = rng.uniform(size=n) < 0.1
    cx[left] = rng.uniform(-4, 30, size=left.sum())
    size = np.exp(rng.uniform(np.log(0.2), np.log(8.0), size=n))
    z = np.exp(rng.uniform(np.log(60.0), np.log(4000.0), size=n))
    px = cx[:, None] + rng.normal(size=(n, 3)) * size[:, None]
    py = cy[:, None] + rng.normal(size=(n, 3)) * size[:, None]
    pz = z[:, None] * (1.0 + rng.normal(size=(n, 3)) * 0.01)
    tris = np.zeros(n, dtype=TRIANGLE)
    pos = np.stack([(px - w / 2) * pz / fov, (py - h / 2) * pz / fov, pz], axis=-1)
    tris["vertices"]["pos"][:, :, :3] = pos.astype(np.float32)
    tris["vertices"]["normal"][:, :, 2] = -1.0
    tris["vertices"]["vt"] = rng.uniform(0, 1, size=(n, 3, 2)).astype(np.float32)
    tris["vertices"]["object_id"][:, 0] = rng.integers(0, 2, size=n)       # half two-sided, half culled by winding (early shadow back-face cull)
    objs = np.array([scene.make_obj_desc(pos=(0, 0, 0), scale=1.0, tid=0, feature_flag=FEATURE_TWO_SIDED),
                     scene.make_obj_desc(pos=(0, 0, 0), scale=1.0, tid=0)], dtype=OBJ_DESC)
    lights = np.array([scene.make_light((40, -60, 30), shadow=1), scene.make_light((-900, 500, 2500), shadow=1)], dtype=LIGHT)
    tex = [scene.procedural_texture(64, 9)]
    return scene.Scene(cfg, tris, objs, lights, tex, c_pos=(0, 0, 0), c_rot=(0, 0, 0), clear=(0, 0, 0, 1.0), name=f"micro{seed}")


@pytest.mark.parametrize("seed,w,h,L", [(11, 640, 360, 64), (12, 333, 217, 128), (13, 41 * 8, 200, 256)])
def test_micro_triangles_on_borders_and_top_rows(seed, w, h, L):
    s = _micro_soup(seed, 20000, w, h, L)
    g, o = render_both(s, frames=2, threads=0)
    for k in range(2):
        assert np.array_equal(g.read_shadow(0, k), o.read_shadow(0, k)), f"cubemap {k} differs"
    st = assert_frame_parity(g, o, label=s.name)
    assert st["covered"] > 2000
    d = o.read_depth()
    assert (d[:8] != 0xFFFFFFFF).any() and (d[:, :4] != 0xFFFFFFFF).any(), "the scene must reach the top rows and the left columns"
