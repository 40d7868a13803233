"""not-gpu: the N > 1 path with world_size 2 over gloo. Each process renders its screen band and its share of the
shadow-cubemap faces with the CPU oracle standing in for the device renderer (same rr.h-shaped interface, same
band/face configuration fields), exchanges faces with the in-place all-gather and gathers the bands to rank 0 through
openclrenderer_b200.distributed — the same calls bench.py makes over NCCL. Rank 0 checks the composited frame and the
gathered cubemaps against a single-process render."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from openclrenderer_b200 import scene, distributed as rrd
    from oracle.binding import Oracle
    s = scene.scene_spheres(320, 192, n_spheres=6, grid=(3, 2), seed=3, n_lights=2, light_dim=64, tex_sizes=(64, 32))
    W, H, L = s.cfg.width, s.cfg.height, s.cfg.light_dim
    cfg = rrd.band_config(s.cfg, world, rank)
    o = Oracle(cfg, threads=1)
    s.upload(o)
    n_shadow = o.n_shadow
    chunk = rrd.face_chunk(n_shadow, world)
    # shadows: render the owned faces, all-gather in place, hand the complete cubemaps back to the renderer
    o.frame_shadows(1)
    flat = torch.full((chunk * world * L * L,), -1, dtype=torch.int32)
    slabs = np.concatenate([o.read_shadow(0, k).reshape(-1) for k in range(n_shadow)])
    flat[:slabs.size] = torch.from_numpy(slabs.view(np.int32))
    rrd.all_gather_faces(flat, chunk * L * L, rank)
    full = flat.numpy().view(np.uint32)[:n_shadow * 6 * L * L].reshape(n_shadow, 6, L, L)
    for k in range(n_shadow):
        o.write_shadow(0, k, full[k])
    o.frame_draw(s.c_pos, s.c_rot, s.clear)
    fb = torch.from_numpy(o.read_rgba8().copy())
    rows = H // world
    rrd.gather_bands(fb, rows, rank, world, dst=0)
    # interleaved split (the peer-memory path's row ownership): rows of foreign tiles blanked, then gather_tiles to rank 0
    tile = 16
    own = rrd.owned_rows(H, tile, world, rank)
    whole = Oracle(s.cfg, threads=1)
    s.upload(whole)
    whole.frame_shadows(1)
    whole.frame_draw(s.c_pos, s.c_rot, s.clear)
    fbt = torch.from_numpy(whole.read_rgba8().copy())
    fbt[torch.from_numpy(~own)] = 0
    rrd.gather_tiles(fbt, tile, rank, world, dst=0)
    if rank == 0:
        ref = Oracle(s.cfg, threads=1)
        s.upload(ref)
        ref.frame_shadows(1)
        ref.frame_draw(s.c_pos, s.c_rot, s.clear)
        ok_shadow = all(np.array_equal(ref.read_shadow(0, k), full[k]) for k in range(n_shadow))
        ok_frame = np.array_equal(ref.read_rgba8(), fb.numpy())
        y0, y1 = rrd.band_rows(H, world, 0)
        ok_depth = np.array_equal(ref.read_depth(), o.read_depth())
        ok_tiles = np.array_equal(ref.read_rgba8(), fbt.numpy())
        q.put((ok_shadow, ok_frame, ok_depth, ok_tiles))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_band_and_face_exchange_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res == (True, True, True, True), res
