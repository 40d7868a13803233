"""not-gpu: pieces of the oracle against independent float64 restatements of the reference's HOST mirrors
(engine::rot_about engine.cpp:778-848, depth_project_singular 914-932, back_project_about_camera 855-886) and
hand-computed cases (SURVEY.md §4 "cross-check material")."""
import math

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import binding as ob

f32 = np.float32


def rot_about64(p, c_pos, c_rot):
    """engine::rot_about (engine.cpp:778-848) in float64."""
    c, s = np.cos(np.asarray(c_rot, np.float64)), np.sin(np.asarray(c_rot, np.float64))
    rel = np.asarray(p, np.float64) - np.asarray(c_pos, np.float64)
    x = c[1] * (s[2] * rel[1] + c[2] * rel[0]) - s[1] * rel[2]
    y = s[0] * (c[1] * rel[2] + s[1] * (s[2] * rel[1] + c[2] * rel[0])) + c[0] * (c[2] * rel[1] - s[2] * rel[0])
    z = c[0] * (c[1] * rel[2] + s[1] * (s[2] * rel[1] + c[2] * rel[0])) - s[0] * (c[2] * rel[1] - s[2] * rel[0])
    return np.array([x, y, z])


@settings(max_examples=200, deadline=None)
@given(st.lists(st.floats(-3000, 3000, width=32), min_size=3, max_size=3), st.lists(st.floats(-500, 500, width=32), min_size=3, max_size=3),
       st.lists(st.floats(-3.25, 3.25, width=32), min_size=3, max_size=3))
def test_rot_matches_host_mirror(p, cpos, crot):
    ob.load_oracle()
    got = ob.unit_rot(p, cpos, crot)
    want = rot_about64(np.asarray(p, f32), np.asarray(cpos, f32), np.asarray(crot, f32))
    assert np.allclose(got, want, rtol=2e-5, atol=2e-2)


@settings(max_examples=100, deadline=None)
@given(st.lists(st.floats(-2000, 2000, width=32), min_size=3, max_size=3), st.lists(st.floats(-3.25, 3.25, width=32), min_size=3, max_size=3))
def test_back_rot_inverts_rot(p, crot):
    """engine::back_rotate (engine.cpp:888-912) is the inverse of rotate."""
    fwd = ob.unit_rot(p, (0, 0, 0), crot)
    back = ob.unit_rot(fwd, (0, 0, 0), crot, back=True)
    assert np.allclose(back, np.asarray(p, f32), rtol=1e-4, atol=5e-2)


def test_rot_quat_known_values():
    # identity quaternion (object.cpp:54 default) leaves points untouched, bit for bit
    p = np.array([1.25, -7.5, 300.0], f32)
    assert np.array_equal(ob.unit_rot_quat(p, (0, 0, 0, 1)), p)
    # 90 degrees about +y: (1,0,0) -> (0,0,-1)
    h = math.sqrt(0.5)
    assert np.allclose(ob.unit_rot_quat((1, 0, 0), (0, h, 0, h)), (0, 0, -1), atol=1e-6)
    # non-unit quaternions are normalised first (cl2.cl:352)
    assert np.allclose(ob.unit_rot_quat((1, 0, 0), (0, 2, 0, 2)), (0, 0, -1), atol=1e-6)
    # back_rot_quat inverts
    q = (0.3, -0.2, 0.9, 0.1)
    r = ob.unit_rot_quat(ob.unit_rot_quat((3, 4, 5), q), q, back=True)
    assert np.allclose(r, (3, 4, 5), atol=1e-4)


def test_wang_hash_and_xorshift_known_values():
    L = ob.load_oracle()

    def wang(seed):
        seed &= 0xFFFFFFFF
        seed = (seed ^ 61) ^ (seed >> 16)
        seed = (seed * 9) & 0xFFFFFFFF
        seed = seed ^ (seed >> 4)
        seed = (seed * 0x27d4eb2d) & 0xFFFFFFFF
        return seed ^ (seed >> 15)

    def xs(s):
        s ^= (s << 13) & 0xFFFFFFFF
        s ^= s >> 17
        s ^= (s << 5) & 0xFFFFFFFF
        return s & 0xFFFFFFFF
    for v in (0, 1, 61, 12345, 0xFFFFFFFF, 3840 * 2160 * 259 + 17):
        assert L.orc_unit_wang_hash(v & 0xFFFFFFFF) == wang(v)
        assert L.orc_unit_xorshift(v & 0xFFFFFFFF) == xs(v & 0xFFFFFFFF)
    # q12: x + y*W*H overflows int32 from y ~ 259 at 4K; the oracle wraps mod 2^32
    x, y, W, H = 100, 2000, 3840, 2160
    assert wang((x + y * W * H) & 0xFFFFFFFF) == L.orc_unit_wang_hash((x + y * W * H) & 0xFFFFFFFF)


def test_point_in_tri_edges_inclusive():
    L = ob.load_oracle()
    tri = np.array([10, 10, 20, 10, 10, 20], f32)                # q2: tolerances make every edge inclusive
    inside = lambda x, y: bool(L.orc_unit_point_in_tri(f32(x), f32(y), tri.ctypes.data))
    assert inside(12, 12) and inside(10, 10) and inside(20, 10) and inside(10, 20) and inside(15, 15) and inside(15, 10)
    assert not inside(16, 15) and not inside(9, 12) and not inside(12, 9) and not inside(21, 10)
    tri2 = np.array([10, 10, 10, 20, 20, 10], f32)               # opposite winding: same set (sign handling)
    assert bool(L.orc_unit_point_in_tri(f32(12), f32(12), tri2.ctypes.data))


def test_cubeface_rules():
    L = ob.load_oracle()
    zero = np.zeros(3, f32)

    def face(p):
        a = np.asarray(p, f32)                   # keep the array alive across the call
        return L.orc_unit_cubeface(a.ctypes.data, zero.ctypes.data)
    assert face((-5, 1, 1)) == 4 and face((5, 1, 1)) == 5          # |x| dominant (ties to x)
    assert face((1, -5, 1)) == 1 and face((1, 5, 1)) == 3
    assert face((1, 1, -5)) == 2 and face((1, 1, 5)) == 0
    assert face((3, 3, 3)) == 5 and face((0, 0, 0)) == 5           # ties go to x; the zero vector too (>= comparisons)
    assert face((1, 2, 2)) == 3                                    # y ties with z -> y


def test_log2_approx_and_texture_mod_and_acos():
    L = ob.load_oracle()
    for v in (1.0, 2.0, 3.7, 16.0, 1000.0):
        assert abs(L.orc_unit_log2_approx(f32(v)) - math.log2(v)) < 0.01
    assert L.orc_unit_log2_approx(f32(0.0)) < -100                # worst == 0 -> clamped to mip 0 by the caller
    out = np.zeros(2, f32)
    for vin, want in (((0.25, 0.75), (0.25, 0.75)), ((1.25, 2.5), (0.75, 0.5)), ((-0.25, -1.5), (0.25, 0.5)), ((1.0, 0.0), (1.0, 0.0))):
        a = np.asarray(vin, f32)
        L.orc_unit_texture_mod(a.ctypes.data, out.ctypes.data)       # q14 asymmetric mirror repeat
        assert np.allclose(out, want, atol=1e-6), (vin, out)
    for x in (0.0, 0.05, 0.5, 0.95, 1.0):
        assert abs(L.orc_unit_rational_acos(f32(x)) - math.acos(x)) < 0.02


def test_clip_project_cases():
    W, H, fov = 800.0, 600.0, 230.940094
    # all in front: one triangle, pinhole projection xy*fov/z + (W/2,H/2) (depth_project_singular host mirror, engine.cpp:914-932)
    t = ob.unit_clip_project([0, 0, 100, 50, 0, 100, 0, 50, 200], 20, W, H, fov)
    assert len(t) == 1
    assert np.allclose(t[0][1], [50 * fov / 100 + 400, 300, 100], rtol=1e-6)
    assert np.allclose(t[0][2], [400, 50 * fov / 200 + 300, 200], rtol=1e-6)
    # all behind (z <= 20) -> nothing; beyond depth_far counts as behind
    assert len(ob.unit_clip_project([0, 0, 10, 5, 0, 20, 0, 5, -3], 20, W, H, fov)) == 0
    assert len(ob.unit_clip_project([0, 0, 400000, 5, 0, 400000, 0, 5, 400000], 20, W, H, fov)) == 0
    # one behind -> two triangles sharing the clipped edge at z == 20
    t = ob.unit_clip_project([0, 0, 10, 100, 0, 120, 0, 100, 120], 20, W, H, fov)
    assert len(t) == 2 and abs(t[0][0][2] - 20) < 1e-4 and abs(t[1][2][2] - 20) < 1e-4
    # two behind -> one triangle, valid vertex kept in its slot
    t = ob.unit_clip_project([0, 0, 10, 100, 0, 10, 0, 100, 120], 20, W, H, fov)
    assert len(t) == 1 and abs(t[0][2][2] - 120) < 1e-6 and abs(t[0][0][2] - 20) < 1e-4 and abs(t[0][1][2] - 20) < 1e-4


def _fmaf(a, b, c):
    return f32(np.float64(f32(a)) * np.float64(f32(b)) + np.float64(f32(c)))     # exact in double for these magnitudes


def scan_closed_form(mm, op, distance):
    """Closed form of the reference's pixel walk (rule q1 generalised): pixel k of the row-major box gets
    y_k = floor(fma(k, 1/width, mm2)) and x_k = mm0 + (j mod width) + (k - j), j = last index <= k at which the float row
    counter changed (or the chunk start). Used by the warp-cooperative CUDA rasteriser; property-tested here."""
    mm0, mm1, mm2, mm3 = [f32(v) for v in mm]
    width = int(mm1 - mm0)
    if width <= 0:
        return np.zeros((0, 2), np.int32)
    iw = f32(1.0) / f32(width)
    k0 = op * distance
    yk = lambda k: np.floor(_fmaf(k, iw, mm2))
    out = []
    for k in range(k0, k0 + op + 1):
        y = yk(k)
        if y >= mm3:
            break
        # first index with this row value
        c = int(y - mm2) * width
        while c > k0 and yk(c - 1) == y:
            c -= 1
        while yk(c) < y:
            c += 1
        j = max(c, k0)
        x = mm0 + f32(j % width) + f32(k - j)
        if x >= mm1:
            continue
        out.append((int(x), int(y)))
    return np.asarray(out, np.int32).reshape(-1, 2)


@settings(max_examples=400, deadline=None)
@given(st.integers(0, 7000), st.integers(1, 7679), st.integers(0, 4000), st.integers(1, 4319), st.integers(0, 40))
def test_scan_closed_form_equals_literal_walk(mm0, w, mm2, h, chunk):
    mm = (mm0, mm0 + w, mm2, mm2 + h)
    area = w * h
    for op in (500, 300):
        n_chunks = -(-area // op)
        d = min(chunk * 37 % max(n_chunks, 1), n_chunks - 1) if n_chunks > 0 else 0
        lit = ob.unit_scan(mm, op, d)
        cf = scan_closed_form(mm, op, d)
        assert np.array_equal(lit, cf), (mm, op, d)


def test_scan_covers_half_open_box_with_one_pixel_overlap():
    mm = (5, 12, 3, 9)                       # 7 x 6 box -> 42 pixels in one chunk
    px = ob.unit_scan(mm, 500, 0)
    want = [(x, y) for y in range(3, 9) for x in range(5, 12)]
    assert [tuple(p) for p in px] == want
    mm = (0, 100, 0, 100)                    # 10,000 pixels -> 20 chunks of 500 (+1 overlap each)
    seen = set()
    for d in range(20):
        p = ob.unit_scan(mm, 500, d)
        assert len(p) == (501 if d < 19 else 500)
        seen |= {tuple(q) for q in p}
    assert len(seen) == 10000
    # documented quirk (q1): a lagging float row counter skips the first column of some rows, e.g. width 41
    mm = (0, 41, 0, 200)
    cols0 = {tuple(q) for d in range(17) for q in ob.unit_scan(mm, 500, d) if q[0] == 0}
    assert 0 < len(cols0) < 200


# ---- the invariants the exact integer rasteriser of the setup kernels relies on (rr_kernels.cuh: shadow_raster_small, InlineRaster) ----
@settings(max_examples=300, deadline=None)
@given(st.integers(0, 4000), st.integers(1, 48), st.integers(0, 4200), st.integers(1, 48))
def test_small_box_walk_has_no_lagging_row_when_rows_fit_under_min_y(mm0, w, mm2, h):
    """a single small chunk (<= 48 slots) whose box starts at row min_y >= rows - 1 is walked exactly row-major: the float row
    counter floor(fma(k, 1/width, min_y)) cannot lag at a row start r * width when r <= min_y (the fast path's condition)"""
    if w * h > 48 or h - 1 > mm2:
        return
    px = ob.unit_scan((mm0, mm0 + w, mm2, mm2 + h), 300, 0)
    want = [(x, y) for y in range(mm2, mm2 + h) for x in range(mm0, mm0 + w)]
    assert [tuple(p) for p in px] == want


@settings(max_examples=300, deadline=None)
@given(st.lists(st.integers(-2047, 2047), min_size=6, max_size=6), st.integers(0, 2 ** 31))
def test_point_in_tri_on_integer_vertices_is_closed_triangle_membership(v, seed):
    """rounded vertices with small extents: every partial sum of point_in_tri (cl2.cl:4798-4807) is an integer below 2^24, so the
    fp32 result equals exact integer arithmetic — inside <=> s >= 0, t >= 0, s + t <= 2|A| — and a covered pixel lies inside the
    vertices' own bounding box (the fast path visits nothing else)"""
    x0, y0 = v[0], v[1]
    x1, y1 = x0 + v[2] % 65 - 32, y0 + v[3] % 65 - 32
    x2, y2 = x0 + v[4] % 65 - 32, y0 + v[5] % 65 - 32
    rng = np.random.default_rng(seed)
    A2 = -y1 * x2 + y0 * (-x1 + x2) + x0 * (y1 - y2) + x1 * y2          # 2A, exact
    sg = -1 if A2 < 0 else 1
    lo_x, hi_x, lo_y, hi_y = min(x0, x1, x2), max(x0, x1, x2), min(y0, y1, y2), max(y0, y1, y2)
    for _ in range(40):
        px, py = int(rng.integers(lo_x - 2, hi_x + 3)), int(rng.integers(lo_y - 2, hi_y + 3))
        s = (y0 * x2 - x0 * y2 + (y2 - y0) * px + (x0 - x2) * py) * sg
        t = (x0 * y1 - y0 * x1 + (y0 - y1) * px + (x1 - x0) * py) * sg
        want = s >= 0 and t >= 0 and s + t <= abs(A2) and A2 != 0
        got = bool(ob.load_oracle().orc_unit_point_in_tri(float(px), float(py), (np.array([x0, y0, x1, y1, x2, y2], np.float32)).ctypes.data_as(ob._P)))
        assert got == want, (v, px, py, s, t, A2)
        if got:
            assert lo_x <= px <= hi_x and lo_y <= py <= hi_y
