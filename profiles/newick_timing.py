#!/usr/bin/env python
"""to_newick / from_newick throughput per tree size on one GPU (CUDA events), random trees and the
ladder (v[i] = i: every row hangs off the row before it, the deepest in-batch dependency chains).

    python profiles/newick_timing.py [n ...]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phylo2vec_b200 as p2v  # noqa: E402

dev = torch.device("cuda:0")
sizes = [int(a) for a in sys.argv[1:]] or [16, 100, 256, 512, 1000, 1024, 2000, 5000]


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
    B = max(1, min(2_000_000, 100_000_000 // k))
    g = torch.Generator(device=dev).manual_seed(n)
    hi = (2 * torch.arange(k, device=dev) + 1).to(torch.float64)
    V = (torch.rand((B, k), generator=g, device=dev, dtype=torch.float64) * hi).to(torch.int64)
    for name, VV in (("random", V), ("ladder", torch.arange(k, device=dev).repeat(B, 1).contiguous())):
        buf, ln = p2v.to_newick_batch(VV, raw=True)
        nbytes = int(ln.sum()) + B * k * 8
        t = timeit(lambda: p2v.to_newick_batch(VV, raw=True))
        t2 = timeit(lambda: p2v.from_newick_batch(buf, lengths=ln, n_leaves=n))
        ok = bool(torch.equal(p2v.from_newick_batch(buf, lengths=ln, n_leaves=n), VV))
        print(f"n={n:5d} {name:6s} B={B:8d} to_newick {B / t:10.3e} trees/s {nbytes / t / 1e9:7.1f} GB/s | "
              f"from_newick {B / t2:10.3e} trees/s {nbytes / t2 / 1e9:7.1f} GB/s round_trip={ok}", flush=True)
    del V
