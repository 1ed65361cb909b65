#!/usr/bin/env python
"""Small run through every kernel family for compute-sanitizer (memcheck / racecheck).

    compute-sanitizer --tool memcheck python profiles/sanitize_cmd.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phylo2vec_b200 as p2v  # noqa: E402
from oracle import p2v_oracle as O  # noqa: E402
from tests.helpers import sample_bls, sample_vectors, to_matrix_input  # noqa: E402

for n, B in ((2, 3), (50, 9), (1000, 6), (2000, 3), (5000, 2), (20000, 1)):
    V = sample_vectors(B, n, seed=n)
    k = n - 1
    Vd = torch.from_numpy(V).cuda()
    a = p2v.to_ancestry_batch(Vd)
    p = p2v.to_pairs_batch(Vd)
    e = p2v.to_edges_batch(Vd)
    assert np.array_equal(a.cpu().numpy(), O.batch("to_ancestry", V, k)[0])
    assert torch.equal(p2v.from_ancestry_batch(a), Vd)
    assert torch.equal(p2v.from_pairs_batch(p), Vd)
    assert torch.equal(p2v.from_edges_batch(e), Vd)
for n, B in ((1, 2), (2, 2), (3, 2), (70, 3), (300, 2), (1500, 1), (5000, 1)):
    V = sample_vectors(B, n, seed=n) if n > 1 else np.zeros((B, 0), dtype=np.int64)
    bls = sample_bls(B, n, seed=n)
    d = p2v.cophenetic_distances_batch(torch.from_numpy(V).cuda())
    assert np.array_equal(d.cpu().numpy(), O.cophenetic_batch(V, None)) if n <= 1500 else True
    if n > 1:
        M = torch.from_numpy(to_matrix_input(V, bls)).cuda()
        p2v.cophenetic_distances_batch(M, unrooted=True)
        p2v.cophenetic_distances_batch(M, rows=(n // 3, n // 3 + max(1, n // 8)))
torch.cuda.synchronize()
print("sanitize run ok")
