#!/usr/bin/env python
"""Small fixed workloads for ncu (kept short: ncu replays every kernel ~40 times).

    python profiles/profile_cmd.py decode    # to_ancestry, 131072 trees of n=1000, int64
    python profiles/profile_cmd.py rows      # cophenetic rows, n=65536, rows [0,8192), fp64 + topological
    python profiles/profile_cmd.py rowsfull  # the same with all 65536 rows (34 GB)
    python profiles/profile_cmd.py encode    # from_ancestry, 131072 trees
    python profiles/profile_cmd.py newick    # to_newick + from_newick, 32768 trees of n=1000
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phylo2vec_b200 as p2v  # noqa: E402
from phylo2vec_b200 import _capi  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "decode"
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
if what in ("decode", "encode"):
    B, k = 131072, 999
    hi = (2 * torch.arange(k, device=dev) + 1).to(torch.float64)
    V = (torch.rand((B, k), generator=g, device=dev, dtype=torch.float64) * hi).to(torch.int64)
    for _ in range(3):
        anc = p2v.to_ancestry_batch(V, check=False)
    if what == "encode":
        for _ in range(3):
            p2v.from_ancestry_batch(anc, check=False)
elif what == "newick":
    # to_newick and from_newick, 32768 trees of n=1000
    B, k = 32768, 999
    hi = (2 * torch.arange(k, device=dev) + 1).to(torch.float64)
    V = (torch.rand((B, k), generator=g, device=dev, dtype=torch.float64) * hi).to(torch.int64)
    for _ in range(2):
        buf, ln = p2v.to_newick_batch(V, raw=True)
    for _ in range(2):
        p2v.from_newick_batch(buf, lengths=ln, n_leaves=k + 1)
elif what.startswith("cophsmall"):
    # small trees (BASELINE config C1 shape n=100): which kernel of the distance path costs what
    n = int(what[9:] or 100)
    B = max(1, (1 << 31) // (n * n * 8))
    k = n - 1
    lib = _capi.load()
    hi = (2 * torch.arange(k, device=dev) + 1).to(torch.float64)
    v = (torch.rand((B, k), generator=g, device=dev, dtype=torch.float64) * hi).to(torch.int64)
    bls = torch.rand((B, k, 2), generator=g, device=dev, dtype=torch.float64)
    out = torch.empty((B, n, n), dtype=torch.float64, device=dev)
    need = lib.p2v_cophenetic_workspace_bytes(B, k)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    for blp in (bls.data_ptr(), 0):
        for _ in range(2):
            lib.p2v_cophenetic_batch(v.data_ptr(), 8, blp, B, k, 0, 0, n, out.data_ptr(), None, ws.data_ptr(), need, s)
elif what == "rows2000":
    # BASELINE config C3 shape: batch of n=2000 trees, fp64 branch lengths
    n, B = 2000, 256
    k = n - 1
    lib = _capi.load()
    hi = (2 * torch.arange(k, device=dev) + 1).to(torch.float64)
    v = (torch.rand((B, k), generator=g, device=dev, dtype=torch.float64) * hi).to(torch.int64)
    bls = torch.rand((B, k, 2), generator=g, device=dev, dtype=torch.float64)
    out = torch.empty((B, n, n), dtype=torch.float64, device=dev)
    need = lib.p2v_cophenetic_workspace_bytes(B, k)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        lib.p2v_cophenetic_batch(v.data_ptr(), 8, bls.data_ptr(), B, k, 0, 0, n, out.data_ptr(), None, ws.data_ptr(), need, s)
else:
    n = 65536
    k = n - 1
    lib = _capi.load()
    hi = (2 * torch.arange(k, device=dev) + 1).to(torch.float64)
    v = (torch.rand((1, k), generator=g, device=dev, dtype=torch.float64) * hi).to(torch.int64)
    bls = torch.rand((1, k, 2), generator=g, device=dev, dtype=torch.float64)
    nrows = n if what == "rowsfull" else 8192
    out = torch.empty((nrows, n), dtype=torch.float64, device=dev)
    need = lib.p2v_cophenetic_workspace_bytes(1, k)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    for weighted, blp in ((1, bls.data_ptr()), (0, 0)):
        lib.p2v_cophenetic_prepare_batch(v.data_ptr(), 8, blp, 1, k, 0, None, ws.data_ptr(), need, s)
        for _ in range(3):
            lib.p2v_cophenetic_rows_batch(ws.data_ptr(), need, 1, k, weighted, 0, nrows, out.data_ptr(), s)
torch.cuda.synchronize()
print("done", what, _capi.launch_count())
