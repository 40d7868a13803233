#!/usr/bin/env python
"""bench.py — headline metric of BASELINE.json: Mtri/s (and frames/s, Mfrag/s) of the per-frame raster path on
config 3 (synthetic 1 M-triangle sphere field, texture atlas with mips, 4 shadow-casting lights, 3840x2160).

  python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference [...]                         the reference algorithm on the host cores (CPU oracle)

A step = one whole frame: 4 shadow-cubemap passes (24 faces) + prearrange + depth + id resolve + deferred shading.
N > 1: the same frame split sort-first over N contexts (interleaved row tiles), the 24 cubemap faces sharded round-robin;
faces are pushed and rows composited by the kernels themselves over peer memory (NVLink), `--exchange nccl` keeps the
collective baseline ("scaling": "strong"). Every line carries `stages_ms` (max over ranks) and `roofline`; N > 1 lines also
carry `verify` (composite == the frame one GPU renders alone, checked outside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from openclrenderer_b200 import scene as scn  # noqa: E402

METRIC = "Mtri_per_s_4K_1Mtri_frame"
UNIT = "Mtri/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c3_small", "c4", "c5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-opencl-reference", action="store_true")
    ap.add_argument("--halo", type=int, default=None, help="band halo rows at N>1 (default: the SSAO reach bound of the scene, exact; -1 = every row)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: p2p = interleaved row tiles, faces pushed and rows composited by the kernels over peer memory (product); "
                         "nccl = contiguous bands, all_gather + gather collectives per frame (baseline)")
    ap.add_argument("--tile", type=int, default=0, help="rows per tile of the interleaved split (p2p; 0 = choose)")
    ap.add_argument("--readback", default="distributed", choices=["distributed", "rank0"],
                    help="N>1 p2p end-to-end leg: distributed = every GPU DMAs the rows it shaded into one shared host frame over its own PCIe "
                         "link; rank0 = rows composited on GPU 0 over NVLink, GPU 0 reads the whole frame back")
    ap.add_argument("--depth", type=int, default=3, help="colour-target ring of the end-to-end leg (rr_set_pipeline_depth): frames in flight + 1")
    ap.add_argument("--verify", action="store_true", help="(kept for compatibility) N>1 lines always carry `verify`: rank 0 also renders the frame "
                                                          "alone, outside the timed region, and counts the pixels in which the composite differs")
    return ap.parse_args()


def make_scene(name):
    if name == "c3":
        return scn.scene_c3()
    if name == "c5":
        return scn.scene_c5()
    if name == "c4":
        return scn.scene_c4()
    return scn.scene_spheres(1920, 1080, 60, (10, 6), 20260, 4, 512, name="c3_small_spheres_120ktri_1920x1080_4lights")


def config_dict(s, extra=None):
    d = {"workload": s.name, "triangles": int(len(s.tris)), "objects": int(len(s.objs)), "resolution": f"{s.cfg.width}x{s.cfg.height}",
         "lights": int(len(s.lights)), "shadow_lights": int((s.lights["shadow"] == 1).sum()), "light_dim": int(s.cfg.light_dim),
         "textures": [int(t.shape[0]) for t in s.textures], "macro_profile": "A (main.cpp:80-85: SSAO_RAD=2, TEST_LINEAR)",
         "l2": "inputs larger than L2 (per-frame working set ~0.6 GB vs 126 MB L2); camera jittered every step"}
    if extra:
        d.update(extra)
    return d


def camera(s, i):
    """static scene, camera jittered deterministically per step (SURVEY.md §8d timing method)."""
    j = (i % 7) - 3
    return (s.c_pos[0] + 3.0 * j, s.c_pos[1] + 1.0 * j, s.c_pos[2]), (s.c_rot[0] + 0.001 * j, s.c_rot[1], s.c_rot[2])


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py touches oracle/)
# ---------------------------------------------------------------------------------------------------------------------
def host_cores():
    """cores this process may run on. torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which would make
    omp_get_max_threads() 1: the CPU legs pass the count explicitly (the oracle's pragmas carry num_threads)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_frame_loop(s, steps, warmup, threads=None):
    from oracle.binding import Oracle
    o = Oracle(s.cfg, threads=threads or host_cores())
    s.upload(o)
    times = []
    for i in range(warmup + steps):
        c_pos, c_rot = camera(s, i)
        t0 = time.perf_counter()
        o.frame_shadows(1 if i == 0 else 0)
        o.frame_draw(c_pos, c_rot, s.clear)
        o.swap_buffers()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    stats = {"depth_samples": o.depth_samples, "shadow_samples": o.shadow_samples, "threads": o.threads, "timings": o.timings()}
    return times, stats


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    s = make_scene(args.workload)
    times, st = cpu_frame_loop(s, args.steps, args.warmup)
    ms = 1e3 * sum(times) / len(times)
    T = len(s.tris)
    val = T / (ms * 1e-3) / 1e6
    sample = f"{args.steps} full frames of {s.name} (shadows + draw), mean after {args.warmup} warm-up, OpenMP over work-groups"
    line = {"impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+u32", "data": "synthetic",
            "config": config_dict(s), "fps": round(1e3 / ms, 3),
            "cpu_baseline": {"value": round(val, 4), "unit": UNIT, "cores": st["threads"], "kind": "port", "sample": sample,
                             "note": "CPU restatement of cl2.cl (PoCL / OpenCL unavailable in the image, SURVEY.md §8c)"},
            "e2e": {"value": round(val, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def opencl_reference(s, steps):
    """The reference's own, unmodified cl2.cl kernels on the same B200 through the NVIDIA OpenCL ICD, launched as
    engine.cpp launches them in steady state. Reported beside the CUDA path; a checker/baseline, never the product."""
    from oracle import ref_opencl
    if not ref_opencl.available():
        raise RuntimeError("NVIDIA OpenCL ICD or oracle/_ref/cl2.cl.gz not present")
    T = len(s.tris)

    def run(profile):
        r = ref_opencl.RefCL(s.cfg, mode="shipped", profile=profile)
        s.upload(r)
        for i in range(2):
            c_pos, c_rot = camera(s, i)
            r.frame_shadows(1 if i == 0 else 0)
            r.frame_draw(c_pos, c_rot, s.clear)
            r.swap_buffers()
        r.sync()
        r.enter_steady_state()
        return r
    r = run(False)
    t0 = time.perf_counter()
    for i in range(steps):
        c_pos, c_rot = camera(s, 2 + i)
        r.frame_shadows(0)
        r.frame_draw(c_pos, c_rot, s.clear)
        r.swap_buffers()
    r.sync()
    ms = 1e3 * (time.perf_counter() - t0) / steps
    rp = run(True)
    rp.kernel_ms.clear()
    for i in range(3):
        c_pos, c_rot = camera(s, 2 + i)
        rp.frame_shadows(0)
        rp.frame_draw(c_pos, c_rot, s.clear)
        rp.swap_buffers()
    kernels = {k: round(sum(v) / 3, 4) for k, v in rp.kernel_ms.items()}
    return {"ms_per_step": round(ms, 4), "value": round(T / (ms * 1e-3) / 1e6, 3), "unit": UNIT, "fps": round(1e3 / ms, 2), "steps": steps,
            "kernels_ms_per_frame": kernels, "kernel_sum_ms": round(sum(kernels.values()), 4), "build_options": r.options,
            "what": "unmodified cl2.cl (prearrange, kernel1, kernel2, kernel3, *_realtime_shadowing) via NVIDIA OpenCL on the same GPU, "
                    "launch sizes from the previous frame's counts as engine.cpp does; wall clock over clFinish"}


# ---------------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[1])), mx.append(float(f[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from openclrenderer_b200 import Renderer, rr
    from openclrenderer_b200._abi import RR_BUF_RGBA8, RR_BUF_SHADOW_DYNAMIC

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the one JSON line: some boxes print NCCL's version banner there unless the level is set explicitly
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # whatever NCCL logs (its version banner at WARN) goes to stderr
        dist.init_process_group("nccl", device_id=dev)
    s = make_scene(args.workload)
    W, H, L = s.cfg.width, s.cfg.height, s.cfg.light_dim
    n_shadow = int((s.lights["shadow"] == 1).sum())
    from openclrenderer_b200 import distributed as rrd
    rows = H // world
    cfg = s.cfg.copy(device=local)
    p2p = world > 1 and args.exchange == "p2p"
    if world > 1:
        if args.halo is None:
            args.halo = rrd.ssao_halo(s, [camera(s, i) for i in range(7)])
        if p2p:
            if not args.tile:
                args.tile = rrd.choose_tile(H, world, args.halo)
            cfg = rrd.tile_config(cfg, world, rank, args.tile, args.halo)
        else:
            cfg = rrd.band_config(cfg, world, rank, args.halo)
    r = Renderer(cfg)
    # nccl baseline: colour target and cubemap slab live in torch tensors so torch.distributed (NCCL) can move them
    fb = torch.zeros((H, W, 4), dtype=torch.uint8, device=dev) if (world > 1 and not p2p) else None
    if fb is not None:
        r.bind_external(RR_BUF_RGBA8, fb.data_ptr(), fb.numel())
    chunk = rrd.face_chunk(n_shadow, world)
    if world > 1 and not p2p:
        shadow = torch.full((chunk * world * L * L,), -1, dtype=torch.int32, device=dev)
        r.bind_external(RR_BUF_SHADOW_DYNAMIC, shadow.data_ptr(), shadow.numel() * 4)
    s.upload(r)

    def all_ok(ok):
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(int(t.item()))

    if p2p:
        try:
            rrd.connect_peers(r, rank, world, device=dev, barrier=False)     # after this the kernels exchange faces and rows themselves
            ok = True
        except Exception as e:                                               # no peer access between these GPUs
            sys.stderr.write(f"rank {rank}: peer-memory exchange unavailable ({e}); using the NCCL exchange\n")
            ok = False
        if not all_ok(ok):
            # same renderer, other exchange: contiguous bands + collectives (every rank takes this branch together)
            r.close()
            p2p = False
            args.exchange = "nccl (peer access unavailable)"
            cfg = rrd.band_config(s.cfg.copy(device=local), world, rank, args.halo)
            r = Renderer(cfg)
            fb = torch.zeros((H, W, 4), dtype=torch.uint8, device=dev)
            r.bind_external(RR_BUF_RGBA8, fb.data_ptr(), fb.numel())
            shadow = torch.full((chunk * world * L * L,), -1, dtype=torch.int32, device=dev)
            r.bind_external(RR_BUF_SHADOW_DYNAMIC, shadow.data_ptr(), shadow.numel() * 4)
            s.upload(r)
        dist.barrier()
    stream = torch.cuda.ExternalStream(r.stream(), device=dev)
    sh_stream = torch.cuda.ExternalStream(r.shadow_stream(), device=dev)

    def frame(i):
        c_pos, c_rot = camera(s, i)
        r.frame_shadows(0)
        if p2p:
            r.frame_draw(c_pos, c_rot, s.clear)
            r.swap_buffers()
            return
        if world > 1:
            with torch.cuda.stream(sh_stream):                                   # on the shadow stream: overlaps the main view's setup/depth/ids
                rrd.all_gather_faces(shadow, chunk * L * L, rank)                # faces rendered elsewhere arrive in place
            r.shadows_done()
        r.frame_draw(c_pos, c_rot, s.clear)
        if world > 1:
            with torch.cuda.stream(stream):
                rrd.gather_bands(fb, rows, rank, world, dst=0)
        r.swap_buffers()

    def barrier():
        r.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    r.frame_shadows(1)                     # static-light pass of the first frame (none in this scene) outside the loop
    for i in range(args.warmup):
        frame(i)
    barrier()
    l0 = r.timings()["launches"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    tw0 = time.perf_counter()
    e0.record(stream)
    for i in range(args.steps):
        frame(args.warmup + i)
    e1.record(stream)
    barrier()
    tw1 = time.perf_counter()
    ms_total = e0.elapsed_time(e1)
    launches = r.timings()["launches"] - l0
    clocks = sampler.stop(tw0, tw1) if rank == 0 else None
    t = torch.tensor([ms_total, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, launches = float(tmax[0]), int(tsum[1])
    push_bytes = None
    if p2p and world > 1:                  # cubemap bytes the ranks stored into each other's memory during the timed frames (NVLink)
        pb = torch.tensor([float(r.mgpu_pushed_bytes())], dtype=torch.float64, device=dev)
        dist.all_reduce(pb, op=dist.ReduceOp.SUM)
        push_bytes = float(pb[0]) / (args.steps + args.warmup)
    if os.environ.get("RR_BENCH_RANK_TIMINGS"):
        r.set_profiling(True)
        frame(10_000)
        tm = r.timings()
        r.set_profiling(False)
        sys.stderr.write("rank %d: %s\n" % (rank, json.dumps({k: round(v, 4) if isinstance(v, float) else v for k, v in tm.items()})))
    ms = ms_total / args.steps
    T = len(s.tris)
    val = T / (ms * 1e-3) / 1e6

    # ---- end to end through the public API with host buffers: H2D of the per-frame inputs (object descriptors from
    # pinned memory, as object_context::flush_locations does) and D2H of the finished frame, inside the timed region
    dist_rb = p2p and args.readback == "distributed"
    shared = None
    D = max(2, min(4, args.depth))
    if dist_rb:
        shared = rrd.SharedFrames(H, W, rank, world, depth=D)      # one ring of host frames mapped and page-locked by every rank
        if not shared.ok:
            sys.stderr.write(f"rank {rank}: shared host frame unavailable ({shared.error}); GPU 0 reads the frame back\n")
        if not all_ok(shared.ok):
            shared.close()
            dist_rb, shared = False, None
    if dist_rb:
        r.mgpu_set_readback(1)
        host_ring = shared.frames
    else:
        host_ring = [rr.host_alloc((H, W, 4), np.uint8) for _ in range(D if (world == 1 or p2p) else 1)] if rank == 0 else None
    if world == 1 or p2p:
        r.set_pipeline_depth(D)
    pinned_t = torch.from_numpy(host_ring[0]) if (rank == 0 and not dist_rb) else None
    dummy = np.zeros(4, np.uint8)
    e2e_i = [0]

    last_cam = [0]

    def frame_e2e(i):
        last_cam[0] = i
        c_pos, c_rot = camera(s, i)
        if world == 1:
            r.frame_e2e(c_pos, c_rot, s.clear, 1, host_ring[e2e_i[0] % D])   # pipelined read-back over a ring of D targets; swaps buffers itself
            e2e_i[0] += 1
        elif dist_rb:
            # every rank uploads its descriptors, renders its rows and DMAs them into the shared host frame; when the call
            # returns this rank's rows of the frame issued D-1 calls ago are in host memory, which it publishes; rank 0 (the
            # consumer) takes a frame only when every rank has published it
            n = e2e_i[0]
            r.frame_e2e(c_pos, c_rot, s.clear, 1, host_ring[n % D])
            done = max(0, n - D + 2)
            shared.publish(done)
            if rank == 0:
                shared.wait_complete(done)
            e2e_i[0] = n + 1
        elif p2p:
            # every rank uploads its descriptors and renders its rows into rank 0's target; rank 0 reads the composite back
            r.frame_e2e(c_pos, c_rot, s.clear, 1, host_ring[e2e_i[0] % D] if rank == 0 else dummy)
            e2e_i[0] += 1
        else:
            r.scene_write_objs(s.objs)
            frame(i)
            if rank == 0:
                with torch.cuda.stream(stream):
                    pinned_t.copy_(fb, non_blocking=True)
            r.sync()

    for i in range(3):
        frame_e2e(i)
    barrier()
    te0 = time.perf_counter()
    for i in range(args.steps):
        frame_e2e(3 + i)
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - te0) / args.steps
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te[0])
    e2e = {"value": round(T / (e2e_ms * 1e-3) / 1e6, 3), "unit": UNIT, "ms_per_step": round(e2e_ms, 4),
           "h2d_bytes_per_step": int(len(s.objs) * 144 * world + 32 * world), "d2h_bytes_per_step": int(W * H * 4),
           "readback": ("distributed: every GPU DMAs the rows it shaded into one shared, page-locked host frame" if dist_rb else
                        "GPU 0 reads the whole frame back") if world > 1 else "pipelined",
           "ring_depth": D if (world == 1 or p2p) else 1}
    if world == 1 or dist_rb:
        # the same loop with the dirty-tile read-back (rr_set_readback_tiles): the host buffers hold the same frames bit for bit
        # (tests/test_gpu_parity.py, test_gpu_mgpu.py), but only the tiles that can differ from what a buffer already holds cross
        # the PCIe link(s). Reported beside `e2e` (which stays the plain copy of the whole frame every step), with the bytes moved.
        r.sync()
        r.set_readback_tiles(True)
        for i in range(2 * D):
            frame_e2e(1000 + i)
        r.sync()
        r.readback_tile_bytes()
        barrier()
        td0 = time.perf_counter()
        for i in range(args.steps):
            frame_e2e(1000 + 2 * D + i)
        r.sync()
        barrier()
        dt_ms = 1e3 * (time.perf_counter() - td0) / args.steps
        tb = torch.tensor([dt_ms, float(r.readback_tile_bytes()) / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            tmx = tb.clone()
            dist.all_reduce(tmx, op=dist.ReduceOp.MAX)
            dist.all_reduce(tb, op=dist.ReduceOp.SUM)
            dt_ms, tile_bytes = float(tmx[0]), float(tb[1])
        else:
            dt_ms, tile_bytes = float(tb[0]), float(tb[1])
        e2e["dirty_tiles"] = {"value": round(T / (dt_ms * 1e-3) / 1e6, 3), "unit": UNIT, "ms_per_step": round(dt_ms, 4),
                              "d2h_bytes_per_step": int(tile_bytes),
                              "what": "rr_set_readback_tiles(1): 32x4-pixel tiles that hold a shaded pixel now or held one in the frame last written "
                                      "into the same host buffer are stored into the page-locked host frame; the rest of it already holds the clear colour"}
        if world == 1:
            last = (e2e_i[0] - 1) % D
            c_pos, c_rot = camera(s, last_cam[0])
            r.set_readback_tiles(False)
            r.frame_shadows(0)
            r.frame_draw(c_pos, c_rot, s.clear)
            r.sync()
            e2e["dirty_tiles"]["host_frame_equals_device_frame"] = bool(np.array_equal(r.read_rgba8(), host_ring[last]))
            r.swap_buffers()
    e2e_check = None
    if dist_rb:
        # the last frame of the e2e loop, as the consumer sees it in host memory, against the device composite of the same camera
        r.sync()
        dist.barrier()
        last = e2e_i[0] - 1
        r.mgpu_set_readback(0)
        r.set_readback_tiles(False)
        c_pos, c_rot = camera(s, last_cam[0])
        r.frame_shadows(0)
        r.frame_draw(c_pos, c_rot, s.clear)
        barrier()
        if rank == 0:
            dev_img = r.read_rgba8()
            e2e_check = int((dev_img != shared.frames[last % D]).any(axis=-1).sum())
            e2e["host_frame_pixels_differing_from_device_composite"] = e2e_check
        r.swap_buffers()
        barrier()
        shared.close()

    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": round(val, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+u32",
                "data": "synthetic", "config": config_dict(s, {"parallelism": (f"tiles{world}x{args.tile}rows+faces{world} peer-memory (in-kernel push/composite over NVLink)" if p2p
                                                                                else f"bands{world}+faces{world} nccl") if world > 1 else "single",
                                                                "band_halo": args.halo if world > 1 else None}),
                "fps": round(1e3 / ms, 2), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches)}
        if push_bytes is not None:
            slab_bytes = 4 * 6 * L * L * n_shadow
            line["nvlink"] = {"face_push_bytes_per_frame": int(push_bytes), "cubemap_bytes": int(slab_bytes),
                              "what": "bytes k_push_faces stored into peer memory per frame, all ranks (only 512-byte chunks that hold a sample or held one "
                                      "last time are sent; a full exchange would move cubemap_bytes x (N-1)); colour rows go straight from k_shade into "
                                      "rank 0's target or, with distributed read-back, to the host over each GPU's own PCIe link",
                              "GBps_if_serial": round(push_bytes / (ms * 1e-3) / 1e9, 1)}

    # ---- stage profile (every N: max over ranks), verification of the split frame (N > 1, always), roofline, CPU baseline (N = 1)
    stage_keys = ("shadow_depth_ms", "setup_ms", "depth_ms", "id_ms", "shade_ms", "frame_ms")
    stage = {k: 0.0 for k in stage_keys}
    n_prof = 8
    tm = None
    r.set_profiling(True)                  # per-stage events only for these frames: they are not free (rr.h rr_set_profiling)
    for i in range(n_prof):
        frame(1000 + i)
        tm = r.timings()
        for k in stage:
            stage[k] += tm[k] / n_prof
    r.set_profiling(False)
    barrier()
    if world > 1:
        tst = torch.tensor([stage[k] for k in stage_keys], dtype=torch.float64, device=dev)
        dist.all_reduce(tst, op=dist.ReduceOp.MAX)
        stage = {k: float(tst[i]) for i, k in enumerate(stage_keys)}

    solo = None
    if world > 1:
        # the composite on rank 0 must equal the frame one context renders alone, bit for bit (outside the timed region)
        c_pos, c_rot = camera(s, 12345)
        r.frame_shadows(0)
        if p2p:
            r.frame_draw(c_pos, c_rot, s.clear)
        else:
            with torch.cuda.stream(sh_stream):
                rrd.all_gather_faces(shadow, chunk * L * L, rank)
            r.shadows_done()
            r.frame_draw(c_pos, c_rot, s.clear)
            with torch.cuda.stream(stream):
                rrd.gather_bands(fb, rows, rank, world, dst=0)
        barrier()
        if rank == 0:
            got = r.read_rgba8()
            solo = Renderer(s.cfg.copy(device=local))
            s.upload(solo)
            solo.frame_shadows(0)
            solo.frame_draw(c_pos, c_rot, s.clear)
            solo.sync()
            want = solo.read_rgba8()
            bad = int((got != want).any(axis=-1).sum())
            line["verify"] = {"pixels_differing_from_single_gpu_frame": bad, "pixels": int(W * H),
                              "what": "composite of the N contexts vs the same frame rendered by one context on GPU 0, all four channels"}
        r.swap_buffers()
        barrier()

    if rank == 0:
        # counts of the whole frame: from this context at N = 1, from the solo context that verified the composite at N > 1
        q = r if world == 1 else solo
        if world == 1:
            r.swap_buffers()               # the frame just drawn is in the previous buffer after swap: swap back to read it
        ids, depth, frags = q.read_ids(), q.read_depth(), q.read_fragments()
        qt = q.timings()
        if world == 1:
            r.swap_buffers()
        cov = depth != 0xFFFFFFFF
        P = W * H
        V = int(len(np.unique(frags[ids[cov], 0]))) if cov.any() else 0
        C, F = qt["n_cutdown"], qt["n_fragments"]
        S = n_shadow
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        atom_ms = q.microbench_atomic_min(32 << 20, 1 << 30)
        r_atomic = (1 << 30) / (atom_ms * 1e-3)
        copy_ms = q.microbench_copy(1 << 30)
        if solo is not None:
            solo.close()
        # algorithmic bytes per stage of the WHOLE frame
        B = {"setup": 40 * T + 48 * C + 20 * F,
             "shadow": S * (40 * T + 2 * 4 * 6 * L * L),
             "depth": 4 * P,
             "id": 8 * P,
             "shade": 8 * P + 144 * V + 144 * len(s.objs) + 5 * P + 4 * P + 4 * P + 4 * P}
        B_lookup = S * min(P, 6 * L * L) * 4
        ms_of = {"setup": stage["setup_ms"], "shadow": stage["shadow_depth_ms"], "depth": stage["depth_ms"], "id": stage["id_ms"], "shade": stage["shade_ms"]}
        dom = max(ms_of, key=lambda k: ms_of[k])
        achieved = B[dom] / (ms_of[dom] * 1e-3) / 1e9
        peak = hbm * world                     # N contexts: the frame's bytes over the max-over-ranks stage time against N x the measured copy bandwidth
        line["roofline"] = {"bound": "hbm", "limiter": "instruction issue (ncu: DRAM 3-7 % of peak, issue slots 50-68 % busy; profiles/r2b_ncu_top_kernels.txt)",
                            "kernel": {"setup": "k_setup_main (+ inline depth of small triangles)",
                                       "shadow": "k_shadow_setup (all %d lights, inline raster) + k_raster_shadow_warp" % S,
                                       "depth": "k_raster_warp_depth (work list of the fragments not rasterised inline)",
                                       "id": "k_ids_list (stream over the recorded samples; builds kernel3's pixel list)", "shade": "k_shade over the covered-pixel list (built by k_ids_list); k_clear_next on a side stream"}[dom],
                            "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": None,
                            "peak_source": peak_src + (f" x {world} GPUs" if world > 1 else ""), "algorithmic_bytes_per_launch": int(B[dom]), "avg_ms": round(ms_of[dom], 4)}
        if world == 1:
            try:                               # DRAM bytes of the stage's dominant kernel from the committed ncu capture (per launch)
                tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
                kname = {"setup": "k_setup_main", "shadow": "k_shadow_setup", "shade": "k_shade"}.get(dom)
                if kname in tr:
                    line["roofline"]["traffic"] = int(tr[kname])
                    line["roofline"]["traffic_kernel"] = kname
                    line["roofline"]["traffic_source"] = "profiles/ncu_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of that kernel)"
            except Exception:
                pass
        line["roofline"]["note"] = ("the HBM roofline is the hardware bound of this byte/integer path, but the stages are currently limited by instruction issue, "
                                    "not by memory: the pinned IEEE arithmetic of the reference's per-pixel / per-triangle math dominates. The shadow stage runs "
                                    "concurrently with setup/depth/id on a second stream, so stage times overlap and do not add up to ms_per_step"
                                    + ("; stage times are the max over ranks" if world > 1 else ""))
        line["stages_ms"] = {k: round(v, 4) for k, v in stage.items()}
        line["counts"] = {"cutdown": int(C), "fragments": int(F), "visible_tris": V, "covered_px": int(cov.sum()), "shadow_fragments": int(tm["n_shadow_fragments"])}
        line["microbench"] = {"atomic_min_Gops": round(r_atomic / 1e9, 2), "copy_GBs": round(2 * (1 << 30) / (copy_ms * 1e-3) / 1e9, 1)}
        if world == 1 and not args.no_cpu_baseline:
            n_cpu = 5
            times, st = cpu_frame_loop(s, n_cpu, 1)
            cms = 1e3 * sum(times) / len(times)
            line["cpu_baseline"] = {"value": round(T / (cms * 1e-3) / 1e6, 4), "unit": UNIT, "cores": st["threads"], "kind": "port",
                                    "sample": f"{n_cpu} full frames of {s.name} after 1 warm-up ({cms:.0f} ms/frame), OpenMP over reference work-groups",
                                    "ms_per_step": round(cms, 2)}
            A_depth, A_shadow = st["depth_samples"], st["shadow_samples"]
            B_frame = B["setup"] + B["depth"] + B["id"] + B["shade"] + B["shadow"] + B_lookup
            t_roof_ms = 1e3 * (B_frame / (hbm * 1e9) + (A_depth + A_shadow) / r_atomic)
            line["roofline_frame"] = {"B_frame_bytes": int(B_frame), "atomics": int(A_depth + A_shadow), "t_roof_ms": round(t_roof_ms, 4),
                                      "frac": round(t_roof_ms / ms, 4), "formula": "B_frame/BW_hbm + (A_depth+A_shadow)/R_atomic (SURVEY.md §8d)"}
            line["Mfrag_per_s"] = round((A_depth + A_shadow) / (ms * 1e-3) / 1e6, 1)
    if world == 1 and not args.no_opencl_reference:
        try:
            line["reference_opencl_b200"] = opencl_reference(s, min(args.steps, 20))
        except Exception as e:                         # informative only; never fails the bench
            line["reference_opencl_b200"] = {"unavailable": str(e)[:200]}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
