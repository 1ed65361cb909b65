"""Property tests of the CPU oracle, modelled on the reference's randomised
tests (phylo2vec/src/vector/convert.rs:581-638; py-phylo2vec/tests/test_base.py
:182-226; py-phylo2vec/tests/test_stats.py:28-126) with harness-owned seeds."""

import numpy as np
import pytest

from oracle import p2v_oracle as O
from tests.helpers import (adversarial_vectors, sample_bls, sample_vectors, to_matrix_input,
                           tree_distances_by_definition)


@pytest.mark.parametrize("n_leaves", [2, 3, 4, 5, 10, 33, 100, 257, 1000])
def test_loop_vs_avl_vs_list(n_leaves):
    # convert.rs:594-612
    for ordered in (False, True):
        for v in sample_vectors(4, n_leaves, seed=n_leaves, ordered=ordered):
            a = O.to_pairs_avl(v)
            assert np.array_equal(a, O.to_pairs_loop(v))
            assert np.array_equal(a, O.to_pairs_list(v))
            assert np.array_equal(a, O.to_pairs(v))


@pytest.mark.parametrize("n_leaves", list(range(2, 40)) + [100, 200, 1000])
def test_round_trips(n_leaves):
    for v in sample_vectors(3, n_leaves, seed=100 + n_leaves):
        assert np.array_equal(O.from_pairs(O.to_pairs(v)), v)
        assert np.array_equal(O.from_ancestry(O.to_ancestry(v)), v)
        assert np.array_equal(O.from_edges(O.to_edges(v)), v)


@pytest.mark.parametrize("n_leaves", [2, 3, 7, 64, 500])
def test_adversarial_shapes(n_leaves):
    for name, v in adversarial_vectors(n_leaves).items():
        O.check_v(v)
        a = O.to_pairs_avl(v)
        assert np.array_equal(a, O.to_pairs_list(v)), name
        assert np.array_equal(O.from_ancestry(O.to_ancestry(v)), v), name
        assert np.array_equal(O.from_edges(O.to_edges(v)), v), name


def test_ancestry_structure():
    v = sample_vectors(1, 50, seed=3)[0]
    k = v.size
    anc = O.to_ancestry(v)
    assert np.array_equal(anc[:, 2], np.arange(k + 1, 2 * k + 1))  # parent of row i is k+1+i
    kids = np.sort(anc[:, :2].ravel())
    assert np.array_equal(kids, np.arange(2 * k))                   # every non-root node once
    e = O.to_edges(v)
    assert np.array_equal(e[0::2, 0], anc[:, 0]) and np.array_equal(e[1::2, 0], anc[:, 1])
    assert np.array_equal(e[0::2, 1], anc[:, 2]) and np.array_equal(e[1::2, 1], anc[:, 2])


def test_invalid_vector_rejected():
    for fn in (O.to_pairs, O.to_ancestry, O.to_edges):
        with pytest.raises(AssertionError, match=r"v\[3\] = 10"):
            fn([0, 0, 1, 10, 2])
    with pytest.raises(AssertionError):
        O.cophenetic_distances(np.array([0, 5, 1]))


def test_from_edges_odd_length():
    # convert.rs:202
    with pytest.raises(ValueError):
        O.from_edges([[2, 4], [3, 4], [0, 5]])


@pytest.mark.parametrize("n_leaves", [3, 4, 5, 8, 17, 50])
@pytest.mark.parametrize("unrooted", [False, True])
def test_cophenetic_vs_definition(n_leaves, unrooted):
    """The literal restatement agrees with distances read off the tree itself
    (what the reference checks against ete4)."""
    for s in range(3):
        v = sample_vectors(1, n_leaves, seed=10 * n_leaves + s)[0]
        bls = sample_bls(1, n_leaves, seed=s)[0]
        anc = O.to_ancestry(v)
        d_top = O.cophenetic_distances(v, unrooted)
        assert np.array_equal(d_top, tree_distances_by_definition(anc, None, unrooted))
        d_bl = O.cophenetic_distances(to_matrix_input(v, bls), unrooted)
        assert np.allclose(d_bl, tree_distances_by_definition(anc, bls, unrooted), rtol=1e-13, atol=0)
        assert np.array_equal(d_bl, d_bl.T) and np.all(np.diag(d_bl) == 0)


@pytest.mark.parametrize("n_leaves", [2, 3, 20, 300])
@pytest.mark.parametrize("unrooted", [False, True])
def test_closed_form_rows_vs_literal(n_leaves, unrooted):
    v = sample_vectors(1, n_leaves, seed=n_leaves)[0]
    bls = sample_bls(1, n_leaves, seed=n_leaves)[0]
    lit = O.cophenetic_distances(v, unrooted)
    assert np.array_equal(O.cophenetic_rows(v, unrooted, 0, n_leaves), lit)
    m = to_matrix_input(v, bls)
    lit = O.cophenetic_distances(m, unrooted)
    rows = O.cophenetic_rows(m, unrooted, 0, n_leaves)
    rel = np.abs(rows - lit) / np.maximum(np.abs(lit), 1e-300)
    assert rel.max() <= 1e-13
    lo, hi = n_leaves // 3, min(n_leaves, n_leaves // 3 + 2)
    assert np.array_equal(O.cophenetic_rows(m, unrooted, lo, hi), rows[lo:hi])


def test_matrix_truncating_cast():
    # matrix/base.rs:108-121: v = m[:,0] as usize (truncation)
    m = np.array([[0.0, 0.4, 0.5], [2.9, 0.1, 0.2], [2.2, 0.3, 0.6]])
    m2 = m.copy()
    m2[:, 0] = [0, 2, 2]
    assert np.array_equal(O.cophenetic_distances(m), O.cophenetic_distances(m2))


def test_batch_drivers_match_single():
    V = sample_vectors(16, 40, seed=5)
    k = V.shape[1]
    for op, single in (("to_pairs", O.to_pairs), ("to_ancestry", O.to_ancestry), ("to_edges", O.to_edges)):
        out, st = O.batch(op, V, k, nthreads=2)
        assert not st.any()
        for b in range(V.shape[0]):
            assert np.array_equal(out[b], single(V[b]))
    anc, _ = O.batch("to_ancestry", V, k)
    back, st = O.batch("from_ancestry", anc, k)
    assert not st.any() and np.array_equal(back, V)
    pairs, _ = O.batch("to_pairs", V, k)
    assert np.array_equal(O.batch("from_pairs", pairs, k)[0], V)
    edges, _ = O.batch("to_edges", V, k)
    assert np.array_equal(O.batch("from_edges", edges, k)[0], V)
    bad = V.copy()
    bad[3, 5] = 99
    _, st = O.batch("to_ancestry", bad, k)
    assert st[3] == 6 and st.sum() == 6
    D = O.cophenetic_batch(V[:4], None)
    for b in range(4):
        assert np.array_equal(D[b], O.cophenetic_distances(V[b]))
