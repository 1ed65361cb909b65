#!/usr/bin/env python
"""BASELINE config C4 (one tree, n = 65536) per phase: prepare (narrow + decode + tables), the rows
call (blocks + minima + column tables + rows) and the one-shot call, for the whole matrix and for one
eighth of the rows (what a rank of an 8-GPU job writes).  CUDA-event timings used to steer the work;
the contract numbers are bench.py's.

    python profiles/c4_timing.py [reps] [n]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phylo2vec_b200 import _capi  # noqa: E402

lib = _capi.load()
dev = torch.device("cuda:0")
s = torch.cuda.current_stream().cuda_stream
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
PEAK = 6551.4


def timeit(fn, reps=reps, warm=3):
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


k = n - 1
g = torch.Generator(device=dev).manual_seed(7)
hi = (2 * torch.arange(k, device=dev) + 1).to(torch.float64)
v = (torch.rand((1, k), generator=g, device=dev, dtype=torch.float64) * hi).to(torch.int64)
bls = torch.rand((1, k, 2), generator=g, device=dev, dtype=torch.float64)
need = lib.p2v_cophenetic_workspace_bytes(1, k)
ws = torch.empty(need, dtype=torch.uint8, device=dev)
v32 = v.to(torch.int32)
anc = torch.empty((1, k, 3), dtype=torch.int32, device=dev)
st = torch.empty(1, dtype=torch.int32, device=dev)
dneed = lib.p2v_decode_workspace_bytes(1, k)
dws = torch.empty(max(dneed, 16), dtype=torch.uint8, device=dev)
t_dec = timeit(lambda: lib.p2v_to_ancestry_batch(v32.data_ptr(), 4, 1, k, anc.data_ptr(), st.data_ptr(), dws.data_ptr(), dneed, s))
print(f"n={n} decode(int32, to_ancestry) {t_dec * 1e3:.1f} us", flush=True)


def phases(which, names, first=0, start=15):
    import ctypes
    import numpy as np
    ts = np.zeros(256, dtype=np.uint64)
    lib.p2v_debug_phase_ns(which, ts.ctypes.data_as(ctypes.c_void_p))
    t0 = int(ts[start])
    prev = t0
    parts = []
    for i, nm in enumerate(names):
        parts.append(f"{nm}={(int(ts[first + i]) - prev) / 1e3:.1f}")
        prev = int(ts[first + i])
    if which == 1 and first == 38:
        print("    per-CTA warp0 loop us:", " ".join(f"{int(ts[64 + c]) / 1e3:.0f}" for c in range(64)))
        print("    block sizes:", " ".join(f"{(int(ts[128 + c]) & 0xFFFFFFFF) - (int(ts[128 + c]) >> 32)}" for c in range(64)))
    if which == 1 and first == 38:
        parts.append(f"maxloop={int(ts[40]) / 1e3:.1f}us at cta={int(ts[41]) >> 32} warp={(int(ts[41]) >> 16) & 0xFFFF} blocks={int(ts[41]) & 0xFFFF} nid={int(ts[42])}")
    return " ".join(parts) + f" | total={(prev - t0) / 1e3:.1f} us" + (f" rounds={int(ts[16])},{int(ts[17])}" if which else "")


print("  decode phases (us):", phases(0, ["load", "level0", "merge", "labels-scan", "labels"]),
      "then", phases(0, ["record+publish", "combine", "sweep"], first=6, start=4), flush=True)
for name, weighted, blp in (("fp64", 1, bls.data_ptr()), ("topo", 0, 0)):
    t_prep = timeit(lambda: lib.p2v_cophenetic_prepare_batch(v.data_ptr(), 8, blp, 1, k, 0, None, ws.data_ptr(), need, s))
    print(f"  tables phases {name} (us):", phases(1, ["init", "jump", "rank", "positions"]), flush=True)
    for (r0, r1) in ((0, n // 8), (3 * n // 8, n // 2), (0, n)):
        out = torch.empty((1, r1 - r0, n), dtype=torch.float64, device=dev)
        t_rows = timeit(lambda: lib.p2v_cophenetic_rows_batch(ws.data_ptr(), need, 1, k, weighted, r0, r1, out.data_ptr(), s))
        print("  blocks phases (us):", phases(1, ["count", "ids", "tiles", "minima", "columns"], first=33, start=32).replace(" rounds=", " tables-rounds="), flush=True)
        t_all = timeit(lambda: lib.p2v_cophenetic_batch(v.data_ptr(), 8, blp, 1, k, 0, r0, r1, out.data_ptr(), None,
                                                        ws.data_ptr(), need, s))
        gb = out.numel() * 8 / 1e9
        print(f"n={n} {name} rows=[{r0},{r1}) prepare={t_prep * 1e3:.1f}us rows={t_rows * 1e3:.1f}us "
              f"({gb / t_rows * 1e3:.0f} GB/s, {gb / t_rows * 1e3 / PEAK:.3f}) one-shot={t_all * 1e3:.1f}us "
              f"({gb / t_all * 1e3:.0f} GB/s, {gb / t_all * 1e3 / PEAK:.3f} of measured HBM)", flush=True)
        del out
