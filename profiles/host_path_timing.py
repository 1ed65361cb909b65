#!/usr/bin/env python
"""p2v_to_ancestry_host with pageable (numpy) and pinned buffers, trees/s (wall clock)."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phylo2vec_b200 import _capi  # noqa: E402

lib = _capi.load()
B, n = 400_000, 1000
k = n - 1
rng = np.random.default_rng(0)
V = (rng.random((B, k)) * (2 * np.arange(k) + 1)).astype(np.int64)
vp = ctypes.c_void_p
for name, pin in (("pageable", False), ("pinned", True)):
    if pin:
        hV = torch.from_numpy(V).pin_memory()
        hO = torch.empty((B, k, 3), dtype=torch.int64).pin_memory()
        hS = torch.empty(B, dtype=torch.int32).pin_memory()
        a, o, s = hV.data_ptr(), hO.data_ptr(), hS.data_ptr()
    else:
        O = np.empty((B, k, 3), dtype=np.int64)
        S = np.empty(B, dtype=np.int32)
        O[:] = 0
        a, o, s = V.ctypes.data, O.ctypes.data, S.ctypes.data
    for mode in ("1", "0"):
        os.environ["P2V_HOST_COMPACT"] = mode
        lib.p2v_to_ancestry_host(vp(a), 8, B, k, vp(o), vp(s))
        t0 = time.perf_counter()
        for _ in range(2):
            rc = lib.p2v_to_ancestry_host(vp(a), 8, B, k, vp(o), vp(s))
        t = (time.perf_counter() - t0) / 2
        print(f"{name:9s} compact={mode} rc={rc} {B / t:10.3e} trees/s", flush=True)
