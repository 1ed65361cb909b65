"""Pins the CPU oracle against the reference's own golden vectors
(tests/golden/reference_goldens.json, transcribed from the reference's rstest
tables, doctests and docs notebook -- see make_reference_goldens.py)."""

import numpy as np
import pytest

from oracle import p2v_oracle as O


def test_to_pairs(goldens):
    for g in goldens["to_pairs"]:
        exp = np.array(g["pairs"])
        for fn in (O.to_pairs, O.to_pairs_avl, O.to_pairs_list):
            assert np.array_equal(fn(g["v"]), exp), (fn.__name__, g["src"])


def test_to_ancestry(goldens):
    for g in goldens["to_ancestry"]:
        assert np.array_equal(O.to_ancestry(g["v"]), np.array(g["ancestry"])), g["src"]


def test_to_edges(goldens):
    for g in goldens["to_edges"]:
        assert np.array_equal(O.to_edges(g["v"]), np.array(g["edges"])), g["src"]


def test_to_newick(goldens):
    for g in goldens["to_newick"]:
        assert O.to_newick(g["v"]) == g["newick"], g["src"]


def test_to_newick_single_leaf():
    # build_newick with no pairs returns the cache of leaf 0 (vector/convert.rs:361-381)
    assert O.to_newick(np.zeros(0, dtype=np.int64)) == "0;"
    assert O.to_newick([0]) == "(0,1)2;"


def test_encode(goldens):
    for g in goldens["from_pairs"]:
        assert np.array_equal(O.from_pairs(g["pairs"]), g["v"]), g["src"]
    for g in goldens["from_ancestry"]:
        assert np.array_equal(O.from_ancestry(g["ancestry"]), g["v"]), g["src"]
    for g in goldens["from_edges"]:
        assert np.array_equal(O.from_edges(g["edges"]), g["v"]), g["src"]
    for g in goldens["build_vector"]:
        assert np.array_equal(O.build_vector(g["cherries"]), g["v"]), g["src"]
    for g in goldens["order_cherries"]:
        assert np.array_equal(O.order_cherries(g["in"]), np.array(g["out"])), g["src"]


def test_validation(goldens):
    for v in goldens["check_v"]["ok"]:
        O.check_v(v)
    for v in goldens["check_v"]["panic"]:
        with pytest.raises(AssertionError, match="out of bounds"):
            O.check_v(v)
    for g in goldens["is_unordered"]:
        if g["expected"] == "panic":
            with pytest.raises(AssertionError):
                O.is_unordered(g["v"])
        else:
            assert O.is_unordered(g["v"]) == g["expected"], g["src"]


CHECK_M_MSG = {"columns": "Matrix must have at least 3 columns", "v": "out of bounds", "negative": "Branch lengths must be positive"}


def test_check_m(goldens):
    # phylo2vec/src/matrix/base.rs:76-95, test table :131-170
    for m in goldens["check_m"]["ok"]:
        O.check_m(m)
    for g in goldens["check_m"]["panic"]:
        with pytest.raises(AssertionError, match=CHECK_M_MSG[g["why"]]):
            O.check_m(g["m"])
    with pytest.raises(AssertionError, match="Matrix should not be empty"):
        O.check_m(np.zeros((1, 0)))


def check_rf_golden(fn, g):
    got = fn(g["v1"], g["v2"], g["normalize"])
    exp = g["expected"]
    if exp == "symmetric":
        assert got == fn(g["v2"], g["v1"], g["normalize"]), g["src"]
    elif exp == "unit_interval":
        assert 0.0 <= got <= 1.0, g["src"]
    elif exp == "range_0_2":
        assert 0.0 <= got <= 2.0, g["src"]
    else:
        assert got == exp, g["src"]


def test_robinson_foulds(goldens):
    # phylo2vec/src/vector/distance.rs:41-162 and its test module :164-239
    for g in goldens["robinson_foulds"]:
        check_rf_golden(O.robinson_foulds, g)
    with pytest.raises(AssertionError, match="same number of leaves"):      # distance.rs:220-225
        O.robinson_foulds([0, 1, 2], [0, 1])
    # the definition itself (what the reference checks against ete4): splits read off the tree by brute force
    for n in (4, 5, 6, 9, 17):
        rng = np.random.Generator(np.random.Philox(n))
        for _ in range(5):
            v1 = (rng.random(n - 1) * (2 * np.arange(n - 1) + 1)).astype(np.int64)
            v2 = (rng.random(n - 1) * (2 * np.arange(n - 1) + 1)).astype(np.int64)

            def splits(v):
                anc = O.to_ancestry(v)
                below = {i: frozenset([i]) for i in range(n)}
                out = set()
                for c1, c2, p in anc.tolist():
                    below[p] = below[c1] | below[c2]
                for p, s in below.items():
                    if 2 <= len(s) <= n - 2:
                        out.add(frozenset([s, frozenset(range(n)) - s]))
                return out
            s1, s2 = splits(v1), splits(v2)
            assert O.robinson_foulds(v1, v2) == len(s1 ^ s2)
            assert len(s1) == n - 3 and len(s2) == n - 3


def test_matrix_newick(goldens):
    # phylo2vec/src/matrix/convert.rs:26-122 with its doctests and tables (:18-24, :73-78, :131-205)
    for g in goldens["to_newick_matrix"]:
        assert O.to_newick_matrix(g["m"]) == g["newick"], g["src"]
    for g in goldens["from_newick_matrix"]:
        assert np.array_equal(O.from_newick_matrix(g["newick"]), np.array(g["m"])), g["src"]
    assert O.from_newick_matrix("").shape == (0, 3)                       # :171-178
    # test_newick (:207-222): matrix -> Newick -> matrix is the identity, bit for bit
    from tests.helpers import sample_bls, sample_vectors, to_matrix_input
    for n in (5, 10, 50, 100):
        m = to_matrix_input(sample_vectors(1, n, seed=n)[0], sample_bls(1, n, seed=n)[0])
        assert np.array_equal(O.from_newick_matrix(O.to_newick_matrix(m)), m)
    # ryu's layout (matrix/convert.rs:37 `Buffer::format`): the strings the reference's tables hold
    for x, sx in ((0.5, "0.5"), (3.0, "3.0"), (3.12, "3.12"), (0.059750594, "0.059750594"), (0.0017238181, "0.0017238181"),
                  (1e-7, "1e-7"), (1e16, "1e16"), (1e15, "1000000000000000.0"), (1.5e-5, "0.000015"), (0.0, "0.0")):
        assert O.ryu_format(x) == sx


def test_check_v_message():
    # phylo2vec/src/vector/base.rs:60-67
    with pytest.raises(AssertionError) as e:
        O.check_v([0, 0, 9, 1])
    assert str(e.value) == "Validation failed: v[2] = 9 is out of bounds (max = 4)"


def test_cophenetic_vector(goldens):
    for g in goldens["cophenetic_vector"]:
        got = O.cophenetic_distances(np.array(g["v"], dtype=np.int64), g["unrooted"])
        exp = np.array(g["d"], dtype=np.float64)
        assert got.dtype == np.float64 and got.shape == exp.shape
        assert np.array_equal(got, exp), g["src"]      # exact: reference uses assert_eq
        n = exp.shape[0]
        rows = O.cophenetic_rows(np.array(g["v"], dtype=np.int64), g["unrooted"], 0, n)
        assert np.array_equal(rows, exp), g["src"]


def test_cophenetic_matrix(goldens):
    for g in goldens["cophenetic_matrix"]:
        m = np.array(g["m"], dtype=np.float64)
        got = O.cophenetic_distances(m, g["unrooted"])
        exp = np.array(g["d"], dtype=np.float64)
        if "round" in g:
            assert np.array_equal(got.round(g["round"]), exp), g["src"]
        else:
            # the reference's own tolerance: matrix/graph.rs:99-109
            assert np.allclose(got, exp, rtol=1e-5, atol=1e-8), g["src"]
        rows = O.cophenetic_rows(m, g["unrooted"], 0, exp.shape[0])
        assert np.allclose(rows, got, rtol=1e-14, atol=0), g["src"]


def test_vcv_vector(goldens):
    for g in goldens["vcv_vector"]:
        v = np.array(g["v"], dtype=np.int64)
        exp = np.array(g["c"], dtype=np.float64)
        got = O.vcv(v)
        assert got.dtype == np.float64 and np.array_equal(got, exp), g["src"]   # reference: assert_eq
        assert np.array_equal(O.vcv_rows(v, 0, exp.shape[0]), exp), g["src"]
    with pytest.raises(AssertionError):      # vector/graph.rs:636-641 (should_panic)
        O.vcv(np.zeros(0, dtype=np.int64))


def test_vcv_matrix(goldens):
    for g in goldens["vcv_matrix"]:
        m = np.array(g["m"], dtype=np.float64)
        exp = np.array(g["c"], dtype=np.float64)
        got = O.vcv(m)
        assert np.allclose(got, exp, rtol=1e-5, atol=1e-8), g["src"]   # matrix/graph.rs:72-88
        assert np.array_equal(O.vcv_rows(m, 0, exp.shape[0]), got), g["src"]


def test_node_depths(goldens):
    for g in goldens["node_depths"]:
        t = np.array(g["tree"])
        t = t.astype(np.float64) if t.ndim == 2 else t.astype(np.int64)
        got = O.get_node_depths(t)
        k = t.shape[0]
        assert got.shape == (2 * k + 1,) and got[2 * k] == 0.0
        for node, d in g["depths"].items():
            assert abs(got[int(node)] - d) < 1e-10, (g["src"], node)   # vector/ops.rs:609-612


def test_from_newick(goldens):
    # with parent labels and without (convert.rs:259-266, 662-664, 674-676)
    for g in goldens["from_newick"]:
        assert np.array_equal(O.from_newick(g["newick"]), np.array(g["v"])), g["src"]
    for g in goldens["newick_parse"]:
        assert np.array_equal(O.newick_parse(g["newick"]), np.array(g["ancestry"])), g["src"]
    assert O.from_newick("").size == 0


def test_newick_round_trip_oracle():
    # test_newick (convert.rs:640-656): v -> to_newick -> from_newick == v, n = 10, 100, 1000;
    # without parent labels (py tests/test_base.py:56-72 does the same for matrices) as well
    import re
    from .helpers import sample_vectors
    for n in (2, 3, 10, 100, 1000):
        for v in sample_vectors(4, n, seed=n):
            nw = O.to_newick(v)
            assert np.array_equal(O.from_newick(nw), v), n
            assert np.array_equal(O.from_newick(re.sub(r"\)\d+", ")", nw)), v), n


def test_balance(goldens):
    # vector/balance.rs tests: exact Sackin values, rtol 1e-5 / atol 1e-8 for the floats (their `close`)
    for g in goldens["balance"]:
        got = O.balance(g["v"])
        if "sackin" in g:
            assert got[0] == g["sackin"], g["src"]
        if "variance" in g:
            assert abs(got[1] - g["variance"]) <= 1e-8 + 1e-5 * abs(g["variance"]), g["src"]
        assert abs(got[2] - g["b2"]) <= 1e-8 + 1e-5 * abs(g["b2"]), g["src"]
    assert np.array_equal(O.balance(np.zeros(0, dtype=np.int64)), [0.0, 0.0, 0.0])
