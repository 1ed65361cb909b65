#!/usr/bin/env python
"""Decode / encode throughput per tree size on one GPU (CUDA events), for kernel tuning.

    python profiles/decode_timing.py [n ...]          # default: a spread of sizes
    P2V_DECODE_WARP_KMAX=2016 python profiles/decode_timing.py 4096 16384   # merge kernel instead
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phylo2vec_b200 as p2v  # noqa: E402

dev = torch.device("cuda:0")
sizes = [int(a) for a in sys.argv[1:]] or [16, 32, 64, 128, 256, 512, 1000, 2000, 3000, 4096, 8192, 16384, 65536]


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


for n in sizes:
    k = n - 1
    B = max(1, min(4_000_000, 250_000_000 // k))
    g = torch.Generator(device=dev).manual_seed(n)
    hi = (2 * torch.arange(k, device=dev) + 1).to(torch.float64)
    V = torch.empty((B, k), dtype=torch.int64, device=dev)
    step = max(1, (1 << 26) // k)
    for b0 in range(0, B, step):
        b1 = min(B, b0 + step)
        V[b0:b1] = (torch.rand((b1 - b0, k), generator=g, device=dev, dtype=torch.float64) * hi).to(torch.int64)
    out = torch.empty((B, k, 3), dtype=torch.int64, device=dev)
    p2v.to_ancestry_batch(V, out=out, check=False)
    t = timeit(lambda: p2v.to_ancestry_batch(V, out=out, check=False))
    v2 = torch.empty_like(V)
    t2 = timeit(lambda: p2v.from_ancestry_batch(out, out=v2, check=False))
    ok = bool(torch.equal(v2, V))
    print(f"n={n:6d} B={B:8d} to_ancestry {B / t:12.4e} trees/s {B * k * 32 / t / 1e9:7.1f} GB/s | "
          f"from_ancestry {B / t2:12.4e} trees/s {B * k * 32 / t2 / 1e9:7.1f} GB/s round_trip={ok}", flush=True)
    del V, out, v2
