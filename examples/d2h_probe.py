#!/usr/bin/env python
"""PCIe probe for the e2e leg: D2H bandwidth of one 4K RGBA8 frame from page-locked memory, alone and while kernels run."""
import time
import torch

n = 3840 * 2160 * 4
dev = torch.device("cuda", 0)
src = torch.zeros(n, dtype=torch.uint8, device=dev)
dst = torch.empty(n, dtype=torch.uint8).pin_memory()
a = torch.randn(8192, 8192, device=dev)
s_copy, s_k = torch.cuda.Stream(), torch.cuda.Stream()


def run(busy, reps=40):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if busy:
            with torch.cuda.stream(s_k):
                torch.sin_(a)
        with torch.cuda.stream(s_copy):
            dst.copy_(src, non_blocking=True)
    s_copy.synchronize()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / reps


for busy in (False, True, False, True):
    t = run(busy)
    print(f"busy={busy}: {t * 1e3:.3f} ms per 33 MB frame = {n / t / 1e9:.1f} GB/s")
