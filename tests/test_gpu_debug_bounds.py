"""The decode / encode / distance kernels once more through the -DP2V_DEBUG_BOUNDS build (libphylo2vec_b200_dbg.so): every index
into a shared-memory table is asserted on the device, a bad one traps and fails the launch.  compute-sanitizer is
closed on the GPU pool; this is the memory check that runs instead, over the sizes where the kernels change shape."""

import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_DBG = os.path.join(ROOT, "phylo2vec_b200", "lib", "libphylo2vec_b200_dbg.so")

SCRIPT = r"""
import sys
import numpy as np, torch
sys.path.insert(0, %(root)r)
import phylo2vec_b200 as p2v
from phylo2vec_b200 import _capi
assert _capi.load()._name.endswith("libphylo2vec_b200_dbg.so")
from tests.helpers import adversarial_vectors, sample_vectors
sizes = [2, 3, 33, 64, 65, 1000, 1024, 1025, 1026, 2016, 2017, 2018, 2048, 2049, 4096, 4097, 8193, 12288, 12289,
         16384, 16385, 16386, 20000, 32768, 32769, 65535, 65536, 65537, 70000]
for n in sizes:
    k = n - 1
    B = 160 if n <= 20000 else 40
    V = sample_vectors(B, n, seed=n)
    adv = list(adversarial_vectors(n).values())
    V[:len(adv)] = np.stack(adv)
    for dtype in (torch.int64, torch.int32):
        Vd = torch.from_numpy(V).cuda().to(dtype)
        anc = p2v.to_ancestry_batch(Vd)
        assert torch.equal(p2v.from_ancestry_batch(anc), Vd), n
        assert torch.equal(p2v.from_pairs_batch(p2v.to_pairs_batch(Vd)), Vd), n
        assert torch.equal(p2v.from_edges_batch(p2v.to_edges_batch(Vd)), Vd), n
    one = torch.from_numpy(V[:2]).cuda()                       # a few trees: the CTA / cluster / whole-grid kernels
    assert torch.equal(p2v.from_ancestry_batch(p2v.to_ancestry_batch(one)), one), n
    torch.cuda.synchronize()
# the distance kernels with shared-memory tables: the fused small-tree kernel in all its shapes, the shared-memory
# tables kernel of mid-size batches
from oracle import p2v_oracle as O
from tests.helpers import sample_bls, to_matrix_input
for n in [3, 4, 16, 32, 33, 40, 41, 64, 100, 128, 129, 161, 200, 320, 321, 700, 2000]:
    B = 5
    V = sample_vectors(B, n, seed=7 * n)
    M = np.stack([to_matrix_input(V[b], sample_bls(1, n, b)[0]) for b in range(B)])
    Dt = p2v.cophenetic_distances_batch(torch.from_numpy(V).cuda()).cpu().numpy()
    Df = p2v.cophenetic_distances_batch(torch.from_numpy(M).cuda()).cpu().numpy()
    Cv = p2v.vcv_batch(torch.from_numpy(M).cuda()).cpu().numpy()
    assert np.isfinite(Cv).all()
    if n <= 700:
        ref = O.cophenetic_distances(V[0]); assert np.array_equal(Dt[0], ref), n
        ref = O.cophenetic_distances(M[1]); assert np.max(np.abs(Df[1] - ref) / np.maximum(ref, 1e-300)) <= 1e-12, n
    torch.cuda.synchronize()
print("bounds ok", len(sizes))
"""


def test_boundary_sizes_with_asserted_indices():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    if not os.path.exists(LIB_DBG):
        pytest.fail("libphylo2vec_b200_dbg.so is missing: python -m phylo2vec_b200.build --debug-bounds")
    env = dict(os.environ, P2V_B200_LIB=LIB_DBG, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "bounds ok" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
