"""gpu, >= 2 devices: the peer-memory exchange across PROCESSES — one process per GPU under torchrun, cudaIpc handles exchanged
with torch.distributed, faces pushed and rows composited by the kernels over NVLink, distributed read-back into a shared host
frame. `bench.py --verify` renders the same frame on one GPU and counts differing pixels; both counts must be zero.
(tests/test_gpu_mgpu.py runs the same protocol inside one process on one device.)"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_two_processes_two_gpus_equal_one_gpu(exchange):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "4", "--warmup", "3",
           "--workload", "c3_small", "--exchange", exchange, "--verify"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert lines, r.stdout[-500:]
    d = json.loads(lines[-1])
    assert d["n_gpus"] == 2 and d["value"] > 0 and d["e2e"]["value"] > 0
    assert d["verify"]["pixels_differing_from_single_gpu_frame"] == 0
    if exchange == "p2p":
        assert "peer-memory" in d["config"]["parallelism"]
        assert d["e2e"]["host_frame_pixels_differing_from_device_composite"] == 0
