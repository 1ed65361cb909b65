#!/usr/bin/env python
"""Small fixed workload for ncu: to_ancestry (and from_ancestry) of a batch of mid-size trees.
    python profiles/decode_mid_cmd.py [n] [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phylo2vec_b200 as p2v  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
k = n - 1
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
hi = (2 * torch.arange(k, device=dev) + 1).to(torch.float64)
V = (torch.rand((B, k), generator=g, device=dev, dtype=torch.float64) * hi).to(torch.int64)
for _ in range(3):
    anc = p2v.to_ancestry_batch(V, check=False)
for _ in range(2):
    p2v.from_ancestry_batch(anc, check=False)
torch.cuda.synchronize()
