"""gpu: the multi-GPU exchange over peer memory (include/rr.h rr_mgpu_*) exercised on ONE device — several
contexts in this process wired with rr_mgpu_connect_local, so the whole protocol (owned-face clear, k_push_faces, frame
flags, k_wait_flags, shading straight into rank 0's frame buffer, interleaved row tiles, per-object face reach) runs and
is compared bit for bit with a single-context frame. The cross-process / cross-GPU variant (cudaIpc handles exchanged over
torch.distributed) is what `bench.py --gpus N` runs; `bench.py --gpus N --verify` checks its composite against the frame one
GPU renders alone (profiles/r1m_bench_*gpu_p2p.json: 0 pixels differ at 2, 4 and 8 GPUs) and tests/test_gpu_mgpu_ipc.py runs
exactly that under pytest when two GPUs are visible."""
import numpy as np
import pytest

from openclrenderer_b200 import Renderer, distributed as rrd, rr, scene

pytestmark = pytest.mark.gpu


def _cams(s, n):
    return [((s.c_pos[0] + 35.0 * i, s.c_pos[1] + 5.0 * i, s.c_pos[2]), (s.c_rot[0] + 0.01 * i, 0.0, 0.0)) for i in range(n)]


def _frames_single(s, cams):
    a = Renderer(s.cfg)
    s.upload(a)
    out = []
    for i, (p, r) in enumerate(cams):
        a.frame_shadows(1 if i == 0 else 0)
        a.frame_draw(p, r, s.clear)
        a.sync()
        out.append((a.read_rgba8(), a.read_depth(), [a.read_shadow(0, k) for k in range(a.n_shadow)]))
        a.swap_buffers()
    return out


@pytest.mark.parametrize("world,tile,halo", [(2, 16, 24), (4, 8, 24), (3, 32, -1), (8, 16, 24)])
def test_peer_exchange_equals_single_context(world, tile, halo):
    s = scene.scene_spheres(640, 384, n_spheres=12, grid=(4, 3), seed=11, n_lights=3, light_dim=128, tex_sizes=(128, 64))
    cams = _cams(s, 4)
    want = _frames_single(s, cams)
    rs = [Renderer(rrd.tile_config(s.cfg, world, k, tile, halo)) for k in range(world)]
    for r in rs:
        s.upload(r)
    rr.mgpu_connect_local(rs)
    for i, (p, rot) in enumerate(cams):
        for r in rs:                       # producers first: every wait only depends on work already enqueued
            r.frame_shadows(1 if i == 0 else 0)
        for r in reversed(rs):             # rank 0 last: its final wait needs every peer's shading kernel
            r.frame_draw(p, rot, s.clear)
        for r in rs:
            r.sync()
        col, depth, shadows = want[i]
        assert np.array_equal(rs[0].read_rgba8(), col), f"frame {i}: composite differs"
        for k, r in enumerate(rs):
            own = rrd.owned_rows(s.cfg.height, tile, world, k)
            assert np.array_equal(r.read_depth()[own], depth[own]), f"frame {i} rank {k}: depth differs"
            for li in range(r.n_shadow):
                assert np.array_equal(r.read_shadow(0, li), shadows[li]), f"frame {i} rank {k}: cubemap {li} differs"
        for r in rs:
            r.swap_buffers()

def test_peer_exchange_e2e_pipelined():
    """rr_frame_e2e in multi-GPU mode: alternating colour targets on rank 0, pipelined read-back, same frames."""
    s = scene.scene_spheres(640, 360, n_spheres=8, grid=(4, 2), seed=21, n_lights=2, light_dim=128, tex_sizes=(128, 64))
    cams = _cams(s, 5)
    want = _frames_single(s, cams)
    world, tile = 3, 24
    rs = [Renderer(rrd.tile_config(s.cfg, world, k, tile, 24)) for k in range(world)]
    for r in rs:
        s.upload(r)
    rr.mgpu_connect_local(rs)
    for r in rs:
        r.frame_shadows(1)
    bufs = [rr.host_alloc((s.cfg.height, s.cfg.width, 4)), rr.host_alloc((s.cfg.height, s.cfg.width, 4))]
    dummy = rr.host_alloc((s.cfg.height, s.cfg.width, 4))
    got = []
    for i, (p, rot) in enumerate(cams):
        for r in reversed(rs):
            r.frame_e2e(p, rot, s.clear, 1, bufs[i % 2] if r is rs[0] else dummy)
        if i >= 1:
            got.append(bufs[(i - 1) % 2].copy())
    for r in rs:
        r.sync()
    got.append(bufs[(len(cams) - 1) % 2].copy())
    for i in range(len(cams)):
        assert np.array_equal(got[i], want[i][0]), f"frame {i}"


@pytest.mark.parametrize("depth,tiles", [(2, 0), (3, 0), (2, 1), (3, 1)])
def test_distributed_readback_shared_host_frame(depth, tiles):
    """rr_mgpu_set_readback(1): every context keeps its rows and copies exactly those into ONE shared host frame — as whole rows,
    or (tiles = 1, rr_set_readback_tiles) as the 32x4-pixel tiles of its rows that can differ from what the frame already holds."""
    s = scene.scene_spheres(640, 360, n_spheres=8, grid=(4, 2), seed=21, n_lights=2, light_dim=128, tex_sizes=(128, 64))
    cams = _cams(s, 9 if tiles else 5)
    want = _frames_single(s, cams)
    world, tile = 4, 32                                  # 360 rows: 11 full tiles and a partial one (rank 3's)
    rs = [Renderer(rrd.tile_config(s.cfg, world, k, tile, 24)) for k in range(world)]
    for r in rs:
        s.upload(r)
    rr.mgpu_connect_local(rs)
    for r in rs:
        r.mgpu_set_readback(1)
        r.set_pipeline_depth(depth)
        r.set_readback_tiles(tiles)
        r.frame_shadows(1)
    bufs = [rr.host_alloc((s.cfg.height, s.cfg.width, 4)) for _ in range(depth)]
    for b in bufs:
        b[:] = 7
    got = []
    for i, (p, rot) in enumerate(cams):
        for r in rs:
            r.frame_e2e(p, rot, s.clear, 1, bufs[i % depth])
        if i >= depth - 1:
            got.append(bufs[(i - depth + 1) % depth].copy())   # every context's call for frame i has returned: frame i-depth+1 is whole
    for r in rs:
        r.sync()
    for i in range(len(cams) - depth + 1, len(cams)):
        got.append(bufs[i % depth].copy())
    for i in range(len(cams)):
        assert np.array_equal(got[i], want[i][0]), f"frame {i}"
    if tiles:
        sent = sum(r.readback_tile_bytes() for r in rs)
        assert 0 < sent < len(cams) * s.cfg.height * s.cfg.width * 4


def test_interleaved_rows_without_exchange():
    """band_tile ownership alone (no peer wiring): each context's owned rows equal the full frame's."""
    s = scene.scene_spheres(640, 384, n_spheres=12, grid=(4, 3), seed=11, n_lights=2, light_dim=128, tex_sizes=(128, 64))
    full = Renderer(s.cfg)
    s.upload(full)
    s.render(full, frames=2)
    fd, fc = full.read_depth(), full.read_rgba8()
    world, tile = 4, 24
    for k in range(world):
        cfg = s.cfg.copy(band_tile=tile, band_rank=k, band_world=world, band_halo=24)
        b = Renderer(cfg)
        s.upload(b)
        s.render(b, frames=2)
        own = rrd.owned_rows(s.cfg.height, tile, world, k)
        assert np.array_equal(b.read_depth()[own], fd[own])
        assert np.array_equal(b.read_rgba8()[own], fc[own])


def test_peer_that_never_arrives_times_out_with_rr_err_peer(monkeypatch):
    """the bounded wait of the exchange (k_wait_flags): a context whose peer never renders its share must not hang — the wait
    gives up after RR_MGPU_TIMEOUT_MS, raises the error bit in the control block, and rr_sync reports RR_ERR_PEER."""
    from openclrenderer_b200 import RRError
    monkeypatch.setenv("RR_MGPU_TIMEOUT_MS", "100")
    s = scene.scene_spheres(320, 192, n_spheres=6, grid=(3, 2), seed=3, n_lights=2, light_dim=64, tex_sizes=(64, 32))
    rs = [Renderer(rrd.tile_config(s.cfg, 2, k, 16, 24)) for k in range(2)]
    for r in rs:
        s.upload(r)
    rr.mgpu_connect_local(rs)
    rs[1].frame_shadows(1)                    # rank 1 alone: rank 0's faces and its "target is free" flag never come
    rs[1].frame_draw(s.c_pos, s.c_rot, s.clear)
    with pytest.raises(RRError) as e:
        rs[1].sync()
    assert e.value.code == -5, e.value
