#!/usr/bin/env python
"""Event timings of the cophenetic phases (not a bench line; used to steer optimisation)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phylo2vec_b200 as p2v  # noqa: E402
from phylo2vec_b200 import _capi  # noqa: E402

lib = _capi.load()
dev = torch.device("cuda:0")
s = torch.cuda.current_stream().cuda_stream


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def trees(B, n, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    k = n - 1
    hi = (2 * torch.arange(k, device=dev) + 1).to(torch.float64)
    v = (torch.rand((B, k), generator=g, device=dev, dtype=torch.float64) * hi).to(torch.int64)
    bls = torch.rand((B, k, 2), generator=g, device=dev, dtype=torch.float64)
    return v, bls


for n, B, row_sets in ((65536, 1, [(0, 65536), (0, 8192), (8192, 16384), (0, 32768)]), (16384, 1, [(0, 16384), (0, 2048)]),
                       (2000, 256, [(0, 2000)]), (1000, 1024, [(0, 1000)]), (200, 8192, [(0, 200)])):
    k = n - 1
    v, bls = trees(B, n)
    need = lib.p2v_cophenetic_workspace_bytes(B, k)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    t_dec = timeit(lambda: p2v.to_ancestry_batch(v.to(torch.int32), check=False))
    for weighted, blp, name in ((1, bls.data_ptr(), "fp64"), (0, 0, "topo")):
        t_prep = timeit(lambda: lib.p2v_cophenetic_prepare_batch(v.data_ptr(), 8, blp, B, k, 0, None, ws.data_ptr(), need, s))
        for (r0, r1) in row_sets:
            out = torch.empty((B, r1 - r0, n), dtype=torch.float64, device=dev)
            t_rows = timeit(lambda: lib.p2v_cophenetic_rows_batch(ws.data_ptr(), need, B, k, weighted, r0, r1, out.data_ptr(), s))
            gb = out.numel() * 8 / 1e9
            print(f"n={n} B={B} {name} rows=[{r0},{r1}) decode(int32)={t_dec:.3f}ms prepare={t_prep:.3f}ms rows={t_rows:.3f}ms "
                  f"{gb / t_rows * 1e3:.0f} GB/s  ws={need / 1e6:.1f}MB", flush=True)
            del out
