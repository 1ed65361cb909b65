#!/usr/bin/env python
"""One-shot cophenetic_distances_batch throughput for small trees (CUDA events).

    python profiles/coph_small_timing.py [n ...]
    P2V_COPH_SMALL_NMAX=0 python profiles/coph_small_timing.py 100     # the general kernels instead
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phylo2vec_b200 as p2v  # noqa: E402
from phylo2vec_b200 import batch as p2v_batch  # noqa: E402

dev = torch.device("cuda:0")
sizes = [int(a) for a in sys.argv[1:]] or [16, 32, 64, 100, 128, 200, 256, 400, 512]
for n in sizes:
    k = n - 1
    B = max(1, (1 << 31) // (n * n * 8))
    g = torch.Generator(device=dev).manual_seed(n)
    hi = (2 * torch.arange(k, device=dev) + 1).to(torch.float64)
    V = (torch.rand((B, k), generator=g, device=dev, dtype=torch.float64) * hi).to(torch.int64)
    bls = torch.rand((B, k, 2), generator=g, device=dev, dtype=torch.float64)
    out = torch.empty((B, n, n), dtype=torch.float64, device=dev)
    ws = torch.empty(p2v_batch.cophenetic_workspace_bytes(B, k) + 16, dtype=torch.uint8, device=dev)
    res = []
    for name, fn in (("topo", lambda: p2v.cophenetic_distances_batch(V, out=out, check=False, workspace=ws)),
                     ("fp64", lambda: p2v.cophenetic_distances_batch(V, bls=bls, out=out, check=False, workspace=ws)),
                     ("vcv64", lambda: p2v.vcv_batch(V, bls=bls, out=out, check=False, workspace=ws))):
        fn(); fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = a.elapsed_time(b) / 3 * 1e-3
        res.append(f"{name} {B / t:10.3e} trees/s {B * n * n * 8 / t / 1e9:7.0f} GB/s")
    print(f"n={n:4d} B={B:8d} " + " | ".join(res), flush=True)
