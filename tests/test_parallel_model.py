"""Checks the data-parallel formulations used by the CUDA kernels (modelled in
tests/parallel_model.py) against the oracle, on the CPU."""

import numpy as np
import pytest

from oracle import p2v_oracle as O
from tests import parallel_model as M
from tests.helpers import adversarial_vectors, sample_bls, sample_vectors, to_matrix_input


@pytest.mark.parametrize("n_leaves", [2, 3, 4, 5, 9, 32, 33, 34, 64, 65, 100, 257])
def test_decode_model(n_leaves):
    vs = list(sample_vectors(6, n_leaves, seed=n_leaves)) + \
        list(sample_vectors(2, n_leaves, seed=n_leaves, ordered=True)) + \
        list(adversarial_vectors(n_leaves).values())
    for v in vs:
        pairs, anc, edges = M.decode_model(v)
        assert np.array_equal(pairs, O.to_pairs(v))
        assert np.array_equal(anc, O.to_ancestry(v))
        assert np.array_equal(edges, O.to_edges(v))


@pytest.mark.parametrize("n_leaves", [2, 3, 5, 33, 34, 100, 200])
def test_encode_model(n_leaves):
    for v in list(sample_vectors(4, n_leaves, seed=n_leaves + 1)) + list(adversarial_vectors(n_leaves).values()):
        pairs = O.to_pairs(v)
        ch = np.concatenate([pairs, pairs.max(axis=1, keepdims=True)], axis=1)
        assert np.array_equal(M.build_vector_model(ch), v)
        assert np.array_equal(M.build_vector_model(M.min_desc_cherries_model(O.to_ancestry(v))), v)
        assert np.array_equal(M.build_vector_rows_model(ch), v)                     # the one-pass form of round 2
        assert np.array_equal(M.build_vector_rows_model(M.min_desc_cherries_model(O.to_ancestry(v))), v)


@pytest.mark.parametrize("n_leaves", [3, 4, 5, 8, 9, 17, 40, 70])
@pytest.mark.parametrize("unrooted", [False, True])
@pytest.mark.parametrize("R", [4, 8])
def test_distance_rows_model(n_leaves, unrooted, R):
    for s, v in enumerate(list(sample_vectors(3, n_leaves, seed=n_leaves)) + list(adversarial_vectors(n_leaves).values())):
        d = M.distance_rows_model(v, None, unrooted, R=R)
        assert np.array_equal(d, O.cophenetic_distances(v, unrooted))
        bls = sample_bls(1, n_leaves, seed=s)[0]
        d = M.distance_rows_model(v, bls, unrooted, R=R)
        ref = O.cophenetic_distances(to_matrix_input(v, bls), unrooted)
        assert np.allclose(d, ref, rtol=1e-13, atol=0)


@pytest.mark.parametrize("n_leaves", [3, 4, 5, 9, 17, 40, 70])
@pytest.mark.parametrize("unrooted", [False, True])
def test_distance_rows_model_v2(n_leaves, unrooted):
    """Blocks sized by the requested rows + the three-term form used by rows_kernel."""
    for s, v in enumerate(list(sample_vectors(2, n_leaves, seed=n_leaves)) + list(adversarial_vectors(n_leaves).values())):
        ref_t = O.cophenetic_distances(v, unrooted)
        bls = sample_bls(1, n_leaves, seed=s)[0]
        ref_b = O.cophenetic_distances(to_matrix_input(v, bls), unrooted)
        for (r0, r1) in ((0, n_leaves), (n_leaves // 3, n_leaves // 3 + max(1, n_leaves // 4)), (n_leaves - 1, n_leaves)):
            for R, PMAX in ((4, 16), (2, 3), (64, 512)):
                d = M.distance_rows_model_v2(v, None, unrooted, r0, r1, R=R, PMAX=PMAX)
                assert np.array_equal(d, ref_t[r0:r1])
                d = M.distance_rows_model_v2(v, bls, unrooted, r0, r1, R=R, PMAX=PMAX)
                assert np.allclose(d, ref_b[r0:r1], rtol=1e-13, atol=0)


@pytest.mark.parametrize("n_leaves", [2, 3, 5, 33, 34, 64, 65, 100, 257, 1000, 1500])
def test_merge_slot_assignment_and_forward_sweep(n_leaves):
    for v in list(sample_vectors(3, n_leaves, seed=n_leaves)) + list(adversarial_vectors(n_leaves).values()):
        ref, _ = M.slot_assignment(v)
        assert np.array_equal(M.slot_assignment_merge(v), ref)
        assert np.array_equal(M.ancestry_forward_model(v), O.to_ancestry(v))


@pytest.mark.parametrize("n_leaves", [2, 3, 4, 5, 8, 33, 64, 100])
def test_robinson_foulds_day_model(n_leaves):
    vs = list(sample_vectors(5, n_leaves, seed=3 * n_leaves)) + list(adversarial_vectors(n_leaves).values())
    for a in range(len(vs)):
        for b in (a, (a + 1) % len(vs), (a + 3) % len(vs)):
            assert M.robinson_foulds_day_model(vs[a], vs[b]) == O.robinson_foulds(vs[a], vs[b])
