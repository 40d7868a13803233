"""Sort-first multi-GPU plumbing: one process per GPU, torch.distributed for the exchange.

  bands       : the screen is cut into `world` equal horizontal bands; rank k resolves ids and shades rows [k*rows, (k+1)*rows)
                and the finished bands are gathered to rank 0 (RGBA8, 4*W*H*(world-1)/world bytes over NVLink).
  shadow faces: the 6*S (light, face) cubemap faces are owned in contiguous chunks of ceil(6S/world); every rank renders
                its chunk, then one in-place all-gather makes all faces visible everywhere (4*L*L bytes per face).

Works on any backend: NCCL on the B200s, gloo in the CPU tests (tests/test_multiproc_gloo.py).

Two exchange paths:
  "p2p"  (product) : interleaved row tiles + round-robin faces; the contexts map each other's buffers (rr_mgpu_export /
                     rr_mgpu_connect, handles exchanged here with one all_gather) and the kernels themselves push cubemap
                     faces into the peers and store shaded rows into rank 0's frame buffer over NVLink. No collective
                     runs per frame.
  "nccl" (baseline): contiguous bands + contiguous face chunks, one all_gather_into_tensor and one gather per frame.
"""
import torch
import torch.distributed as dist


def band_rows(height, world, rank):
    if height % world:
        raise ValueError(f"height {height} is not divisible into {world} equal bands")
    rows = height // world
    return rank * rows, (rank + 1) * rows


def face_chunk(n_shadow_lights, world):
    """faces per rank; the cubemap buffer is padded to chunk*world faces so the all-gather is uniform."""
    pairs = 6 * n_shadow_lights
    return (pairs + world - 1) // world


def band_config(cfg, world, rank, halo=-1):
    y0, y1 = band_rows(cfg.height, world, rank)
    return cfg.copy(band_y0=y0, band_y1=y1, band_halo=halo, face_rank=rank, face_world=world)


def all_gather_faces(shadow_flat, chunk_words, rank, group=None):
    """shadow_flat: 1-D int32/uint32 tensor of chunk_words*world words holding this rank's faces at [rank*chunk_words, ...)."""
    mine = shadow_flat[rank * chunk_words:(rank + 1) * chunk_words]
    dist.all_gather_into_tensor(shadow_flat, mine, group=group)


def gather_bands(fb, rows, rank, world, dst=0, group=None):
    """fb: (H, W, 4) uint8 tensor; rank k's band is fb[k*rows:(k+1)*rows]; after the call rank `dst` holds the whole frame."""
    band = fb[rank * rows:(rank + 1) * rows]
    if dist.get_backend(group) == "gloo":
        lst = [torch.empty_like(band) for _ in range(world)] if rank == dst else None
        dist.gather(band.contiguous(), lst, dst=dst, group=group)
        if rank == dst:
            for k in range(world):
                fb[k * rows:(k + 1) * rows].copy_(lst[k])
    else:
        dist.gather(band, [fb[k * rows:(k + 1) * rows] for k in range(world)] if rank == dst else None, dst=dst, group=group)


def ssao_halo(scene, cameras=None, margin=8):
    """Rows of depth a band needs beyond its own for exact SSAO (cl2.cl:2194-2260): the taps reach round(2 * world_rad)
    pixels with world_rad = (SSAO_RAD + 0.5) * FOV_CONST / depth, so the nearest geometry bounds it. Computed from the
    objects' bounding spheres for the given cameras; -1 (rasterise every row) when geometry can reach the near plane."""
    import numpy as np
    from .scene import fov_for
    if scene.cfg.no_ssao:
        return margin
    tris, objs = scene.tris, scene.objs
    oid = tris["vertices"]["object_id"][:, 0]
    r_obj = np.zeros(len(objs))
    np.maximum.at(r_obj, oid, np.linalg.norm(tris["vertices"]["pos"][:, :, :3].astype(np.float64), axis=-1).max(axis=1))
    R = r_obj * np.abs(objs["scale"].astype(np.float64)) * 1.001
    z_near = np.inf
    for c_pos, c_rot in (cameras or [(scene.c_pos, scene.c_rot)]):
        cr, sr = np.cos(np.asarray(c_rot, np.float64)), np.sin(np.asarray(c_rot, np.float64))
        rel = objs["world_pos"][:, :3].astype(np.float64) - np.asarray(c_pos, np.float64)[:3]
        t = sr[2] * rel[:, 1] + cr[2] * rel[:, 0]
        u = cr[1] * rel[:, 2] + sr[1] * t
        v = cr[2] * rel[:, 1] - sr[2] * rel[:, 0]
        z = cr[0] * u - sr[0] * v                                   # rot(), cl2.cl:236
        z_near = min(z_near, float(np.where(z + R > scene.cfg.depth_icutoff, np.maximum(z - R, 0.0), np.inf).min()))
    if not np.isfinite(z_near) or z_near <= scene.cfg.depth_icutoff + 1:
        return -1
    reach = 2.0 * (scene.cfg.ssao_rad + 0.5) * fov_for(scene.cfg) / z_near
    return int(np.ceil(reach)) + margin


# ---- interleaved split + peer-memory exchange ("p2p") ------------------------------------------------------------------
def choose_tile(height, world, halo, min_tile=8, max_tile=64):
    """Rows per tile of the interleaved split: tile t belongs to rank t % world. Small tiles balance the load when the
    geometry sits in a few hundred rows, but every tile drags 2*halo extra depth rows along (SSAO reach) and whole
    clusters / triangles are set up wherever they touch a needed row, so small tiles also replicate more setup work.
    Measured with examples/scaling_probe.py on config 3 (profiles/r1m_scaling_probe.jsonl): 64 rows is best at 2 ranks,
    32 rows at 4 and 8 (max-over-ranks frame time 0.44 / 0.33 / 0.26 ms)."""
    tile = 64 if world <= 2 else 32
    while tile > min_tile and -(-height // tile) < 2 * world:      # tiny frames: keep at least two tiles per rank
        tile //= 2
    return max(min_tile, min(max_tile, tile))


def owned_rows(height, tile, world, rank):
    import numpy as np
    y = np.arange(height)
    return (y // tile) % world == rank


def tile_config(cfg, world, rank, tile, halo=-1):
    """rr_config of rank `rank` for the p2p path: interleaved row tiles, round-robin (light, face) pairs."""
    return cfg.copy(band_y0=0, band_y1=0, band_tile=tile, band_rank=rank, band_world=world, band_halo=halo,
                    face_rank=rank, face_world=world, face_interleave=1)


def connect_peers(renderer, rank, world, group=None, device=None, barrier=True):
    """rr_mgpu_export on every rank, one all_gather of the handle bytes, rr_mgpu_connect. After this the contexts exchange
    cubemap faces and frame rows themselves; call frame_shadows / frame_draw in lock-step on all ranks."""
    mine = renderer.mgpu_export()
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    handles = [bytes(o.cpu().numpy().tobytes()) for o in out]
    renderer.mgpu_connect(rank, world, handles)
    if barrier:
        dist.barrier(group=group)
    return handles


def gather_tiles(fb, tile, rank, world, dst=0, group=None):
    """baseline / CPU-test counterpart of the in-kernel composite for the interleaved split: rank k's rows are the tiles
    t = k, k + world, ...; after the call rank `dst` holds the whole frame. fb: (H, W, 4) uint8."""
    H = fb.shape[0]
    n_tiles = -(-H // tile)
    rows = lambda k: [y for t in range(k, n_tiles, world) for y in range(t * tile, min(H, (t + 1) * tile))]
    mine = fb[rows(rank)].contiguous()
    if rank == dst:
        lst = [torch.empty((len(rows(k)),) + tuple(fb.shape[1:]), dtype=fb.dtype, device=fb.device) for k in range(world)]
        dist.gather(mine, lst, dst=dst, group=group)
        for k in range(world):
            if k != rank:
                fb[rows(k)] = lst[k]
    else:
        dist.gather(mine, None, dst=dst, group=group)


class SharedFrames:
    """Host side of the distributed read-back (rr_mgpu_set_readback(1)): a ring of `depth` full frames plus one completion
    counter per rank in ONE block of shared memory (/dev/shm), mapped by every process and page-locked for its GPU
    (rr_host_register). Every rank's rr_frame_e2e DMAs its own rows of frame i into frames[i % depth]; when the call for
    frame i + depth - 1 returns those rows are complete (rr_set_pipeline_depth(depth)) and the rank publishes i + 1 in its
    counter; the consumer (rank 0) owns frame i once every counter has reached i + 1."""
    HEADER = 4096

    def __init__(self, height, width, rank, world, group=None, depth=2):
        import os
        import numpy as np
        from . import rr
        self.rank, self.world = rank, world
        self.depth = depth
        self.nbytes = self.HEADER + depth * height * width * 4
        name = [f"/dev/shm/rr_frames_{os.getpid()}"] if rank == 0 else [None]
        if rank == 0:
            with open(name[0], "wb") as f:
                f.truncate(self.nbytes)
        if world > 1:
            dist.broadcast_object_list(name, src=0, group=group)
        self.path = name[0]
        self.map = np.memmap(self.path, dtype=np.uint8, mode="r+", shape=(self.nbytes,))
        self.ok, self.error = True, None                # page-locking can fail; every rank still reaches the barrier below
        try:
            rr.host_register(self.map)
        except Exception as e:
            self.ok, self.error = False, e
        self.counters = self.map[:8 * world].view(np.int64)
        fsz = height * width * 4
        self.frames = [self.map[self.HEADER + i * fsz:self.HEADER + (i + 1) * fsz].reshape(height, width, 4) for i in range(depth)]
        if world > 1:
            dist.barrier(group=group)

    def publish(self, n_complete):
        self.counters[self.rank] = n_complete

    def wait_complete(self, n_complete, timeout_s=30.0):
        """block until every rank has published >= n_complete (frames 0 .. n_complete-1 are whole in host memory)"""
        import time
        t0 = time.perf_counter()
        while int(self.counters.min()) < n_complete:
            if time.perf_counter() - t0 > timeout_s:
                raise TimeoutError(f"frame {n_complete - 1}: ranks at {list(self.counters)}")

    def close(self, group=None):
        import os
        from . import rr
        if self.world > 1:
            dist.barrier(group=group)
        if self.ok:
            rr.host_unregister(self.map)
        del self.frames, self.counters
        self.map._mmap.close()
        if self.rank == 0:
            os.unlink(self.path)
