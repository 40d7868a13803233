"""Atlas build at scale and dynamic atlas writes: the batched upload must leave exactly the bytes the
per-texture path leaves (texture_context::alloc_gpu, texture_context.cpp:478-517), update_gpu_tex_colour (cl2.cl:955-984) and
generate_from_raw (cl2.cl:1006-1031) exactly the oracle's."""
import numpy as np
import pytest

from openclrenderer_b200 import scene
from openclrenderer_b200._abi import Config
from oracle.binding import Oracle


def _textures(rng, sizes):
    out = []
    for k, s in enumerate(sizes):
        t = rng.integers(0, 256, size=(s, s, 4), dtype=np.uint8)
        t[..., 3] = 255 if k % 3 else rng.integers(0, 256, size=(s, s), dtype=np.uint8)     # some with real alpha (premultiplied mips)
        out.append(t)
    return out


def _alloc(x, texs):
    n_slices, nums, sizes, mip_start = scene.plan_atlas([t.shape[0] for t in texs])
    x.atlas_alloc(n_slices, nums, sizes, mip_start)
    return x


def test_oracle_batch_equals_serial_and_dynamic_writes():
    rng = np.random.default_rng(3)
    texs = _textures(rng, [64, 32, 64, 16, 128, 32])
    cfg = Config.default(64, 64, light_dim=16)
    a, b = _alloc(Oracle(cfg), texs), _alloc(Oracle(cfg), texs)
    for i, t in enumerate(texs):
        a.atlas_upload(i, t, 1)
    b.atlas_upload_batch(range(len(texs)), texs, 1)
    assert np.array_equal(a.atlas_read_raw(), b.atlas_read_raw())
    before = a.atlas_read_raw().copy()
    a.atlas_fill_colour(2, (255, 10.7, 300, 128), 64, 64)
    after = a.atlas_read_raw()
    ch = (before != after).any(axis=-1)
    # the 64x64 tile plus its 32, 16, 8, 4 mips, minus texels that already held the colour
    assert 0.95 * (64 * 64 + 32 * 32 + 16 * 16 + 8 * 8 + 4 * 4) < ch.sum() <= 64 * 64 + 32 * 32 + 16 * 16 + 8 * 8 + 4 * 4
    assert set(map(tuple, after[ch])) == {(255, 10, 44, 128)}, "convert_uint4 truncates (300 -> 300), convert_uchar4 keeps the low byte (300 -> 44)"
    raw = rng.integers(0, 256, size=(32, 40), dtype=np.uint8)                       # stride 40, image 32 wide
    a.atlas_upload_mono(1, raw, 32, 32)
    mono = a.atlas_read_raw()
    ch2 = (mono != after).any(axis=-1)
    assert ch2.sum() > 900 and (mono[ch2][:, 0] == mono[ch2][:, 3]).all()


@pytest.mark.gpu
def test_batched_atlas_build_equals_serial_and_oracle():
    from openclrenderer_b200 import Renderer
    rng = np.random.default_rng(7)
    sizes = [256, 128, 128, 64, 64, 64, 32, 32, 16, 512, 16, 16] + [32] * 40 + [64] * 20
    texs = _textures(rng, sizes)
    cfg = Config.default(64, 64, light_dim=16)
    serial, batch, orc = _alloc(Renderer(cfg), texs), _alloc(Renderer(cfg), texs), _alloc(Oracle(cfg, threads=0), texs)
    for i, t in enumerate(texs):
        serial.atlas_upload(i, t, 1)
    batch.atlas_upload_batch(range(len(texs)), texs, 1)
    orc.atlas_upload_batch(range(len(texs)), texs, 1)
    s_raw = serial.atlas_read_raw()
    assert np.array_equal(s_raw, batch.atlas_read_raw()), "batched upload differs from the per-texture path"
    assert np.array_equal(s_raw, orc.atlas_read_raw()), "atlas differs from the oracle's"
    assert batch.timings()["launches"] <= 5 + 4, "one launch per phase, whatever the number of textures"


@pytest.mark.gpu
def test_dynamic_atlas_writes_match_oracle():
    from openclrenderer_b200 import Renderer
    rng = np.random.default_rng(11)
    texs = _textures(rng, [128, 64, 64, 32, 256])
    cfg = Config.default(64, 64, light_dim=16)
    g, o = _alloc(Renderer(cfg), texs), _alloc(Oracle(cfg, threads=0), texs)
    for x in (g, o):
        x.atlas_upload_batch(range(len(texs)), texs, 1)
        x.atlas_fill_colour(1, (12.9, 255, 0, 999), 64, 64)
        x.atlas_fill_colour(4, (1, 2, 3, 4), 200, 256)                              # launch smaller than the tile: partial fill
    mono_img = rng.integers(0, 256, size=(64, 80), dtype=np.uint8)
    for x in (g, o):
        x.atlas_upload_mono(2, mono_img, 64, 64)
        x.atlas_upload_mono(3, mono_img[:20], 32, 20)
    g.sync()
    assert np.array_equal(g.atlas_read_raw(), o.atlas_read_raw())
