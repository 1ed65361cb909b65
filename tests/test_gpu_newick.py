"""GPU parity tests for to_newick / from_newick (run with -m gpu on a B200): byte-exact against the
restatements of phylo2vec/src/vector/convert.rs:341-397 and newick/mod.rs:73-127 +
convert.rs:268-278, which tests/test_oracle_goldens.py pins to the reference's own Newick goldens."""

import ctypes

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import p2v_oracle as O  # noqa: E402
from tests.helpers import adversarial_vectors, sample_vectors  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def p2v():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import phylo2vec_b200 as m
    return m


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def test_goldens_single_tree(p2v, goldens):
    from phylo2vec_b200.base.newick import to_newick
    for g in goldens["to_newick"]:
        assert to_newick(np.array(g["v"], dtype=np.int64)) == g["newick"], g["src"]
    assert to_newick(np.zeros(0, dtype=np.int64)) == "0;"          # a single leaf
    assert to_newick([0]) == "(0,1)2;"
    with pytest.raises(AssertionError, match="out of bounds"):
        to_newick(np.array([0, 5, 1]))
    assert to_newick(np.array([[0.0, 0.5, 0.8], [1.0, 0.7, 0.6]])) == "(0:0.7,(1:0.5,2:0.8)3:0.6)4;"     # matrices: test_gpu_newick_matrix.py


SIZES = [2, 3, 4, 5, 9, 10, 11, 33, 64, 65, 99, 100, 101, 129, 500, 999, 1000, 1001, 1023, 1024]
# warp-per-tree kernels up to k = 4095, then uint32 tables in the workspace
BIG_SIZES = [1025, 2000, 2049, 3000, 4095, 4096, 4097, 5000, 9999, 10001, 20000]


@pytest.mark.parametrize("n_leaves", SIZES)
@pytest.mark.parametrize("dtype", [torch.int64, torch.int32])
def test_batch_parity(p2v, n_leaves, dtype):
    B = 40
    V = sample_vectors(B, n_leaves, seed=11 * n_leaves)
    V[1] = sample_vectors(1, n_leaves, seed=1, ordered=True)[0]
    adv = list(adversarial_vectors(n_leaves).values())
    V[2:2 + len(adv)] = np.stack(adv)
    got = p2v.to_newick_batch(dev(V, dtype))
    assert len(got) == B
    for b in range(B):
        assert got[b] == O.to_newick(V[b]), (n_leaves, b)
    buf, lengths = p2v.to_newick_batch(dev(V, dtype), raw=True)
    assert buf.dtype == torch.uint8 and lengths.dtype == torch.int64
    host = buf.cpu().numpy()
    for b in range(B):
        ln = int(lengths[b])
        assert host[b, ln] == 0 and host[b, ln - 1] == ord(";")
    assert int(lengths.max()) + 1 <= buf.shape[1]


def test_invalid_and_limits(p2v):
    from phylo2vec_b200 import _capi
    V = sample_vectors(6, 50, seed=4)
    V[4, 7] = 15
    status = torch.zeros(6, dtype=torch.int32, device="cuda")
    buf, lengths = p2v.to_newick_batch(dev(V), status=status, check=False, raw=True)
    assert list(status.cpu().numpy()) == [0, 0, 0, 0, 8, 0] and int(lengths[4]) == 0
    with pytest.raises(AssertionError, match=r"v\[7\] = 15"):
        p2v.to_newick_batch(dev(V))
    lib = _capi.load()
    assert lib.p2v_newick_max_bytes(3) == len("((0,1)5,(2,3)4)6;") + 1


def test_large_batch_and_host_entry(p2v):
    """BASELINE.json configs[0] shape (n = 100, 10k trees) and the host entry point."""
    from phylo2vec_b200 import _capi
    lib = _capi.load()
    n, B = 100, 10000
    V = sample_vectors(B, n, seed=9)
    got = p2v.to_newick_batch(dev(V))
    for b in np.linspace(0, B - 1, 50).astype(int):
        assert got[b] == O.to_newick(V[b])
    stride = int(lib.p2v_newick_max_bytes(n - 1))
    out = np.zeros((B, stride), dtype=np.uint8)
    lengths = np.zeros(B, dtype=np.int64)
    st = np.ones(B, dtype=np.int32)
    vp = ctypes.c_void_p
    rc = lib.p2v_to_newick_host(V.ctypes.data_as(vp), 8, B, n - 1, out.ctypes.data_as(vp), stride,
                                lengths.ctypes.data_as(vp), st.ctypes.data_as(vp))
    assert rc == 0 and not st.any()
    for b in (0, 1, B // 2, B - 1):
        assert out[b, :lengths[b]].tobytes().decode() == got[b]


@pytest.mark.parametrize("n_leaves", BIG_SIZES)
def test_big_trees_parity(p2v, n_leaves):
    B = 12 if n_leaves < 10000 else 5
    V = sample_vectors(B, n_leaves, seed=5 * n_leaves)
    adv = list(adversarial_vectors(n_leaves).values())[:B - 1]
    V[1:1 + len(adv)] = np.stack(adv)
    for dtype in (torch.int64, torch.int32):
        got = p2v.to_newick_batch(dev(V, dtype))
        for b in range(B):
            assert got[b] == O.to_newick(V[b]), (n_leaves, b)
    # and back: string -> vector
    back = p2v.from_newick_batch(got)
    assert np.array_equal(back.cpu().numpy(), V)
    buf, lengths = p2v.to_newick_batch(dev(V), raw=True)
    back32 = p2v.from_newick_batch(buf, n_leaves=n_leaves, dtype=torch.int32)       # NUL-terminated rows
    assert back32.dtype == torch.int32 and np.array_equal(back32.cpu().numpy(), V)
    import re
    bare = [re.sub(r"\)\d+", ")", s) for s in got]                                    # without parent labels
    assert np.array_equal(p2v.from_newick_batch(bare).cpu().numpy(), V)


def test_from_newick_goldens(p2v, goldens):
    from phylo2vec_b200.base.newick import from_newick
    for g in goldens["from_newick"]:                                     # with and without parent labels
        got = from_newick(g["newick"])
        assert got.dtype == np.int64 and np.array_equal(got, np.array(g["v"])), g["src"]
    assert from_newick("0;").size == 0
    assert np.array_equal(from_newick("(0,1)2;"), [0])
    assert from_newick("((0:0.1,1:0.2)3:0.3,2:0.1)4;").shape == (2, 3)              # branch lengths: the matrix form
    assert np.array_equal(from_newick("(0,1)5,(2,3)4)6;"), [0, 2, 2])      # like the reference, "(" is not checked
    for bad in ("((0,1)5,(2,3)4)6", "((0,1)5,(2,3)4)9;", "((0,1)5,(2,x)4)6;", "((0,1)5,2,3)6;",
                "((0,1)5,(2,2)4)6;", "((0,1)5,(2,3)5)6;", "((0,1)5;(2,3)4)6;"):
        with pytest.raises(ValueError):
            from_newick(bad)


def test_from_newick_batch_goldens_and_malformed(p2v, goldens):
    """The same known answers and malformed strings through the batched entry point (16-byte aligned
    rows: the warp-per-tree tokenizer), one string per call and all of one size in one call."""
    for g in goldens["from_newick"]:
        got = p2v.from_newick_batch([g["newick"]])
        assert np.array_equal(got.cpu().numpy()[0], np.array(g["v"])), g["src"]
    bad = ["((0,1)5,(2,3)4)6", "((0,1)5,(2,3)4)9;", "((0,1)5,(2,x)4)6;", "((0,1)5,2,3)6;", "((0,1)5,(2,2)4)6;",
           "((0,1)5,(2,3)5)6;", "((0,1)5;(2,3)4)6;", "((0,1)5,(2,3)4)6;;", "0,1)5,(2,3)4)6;", "((0,1)5,(2,3)4)123456789;",
           "((0,1)5,(2,3)),6;", "((0 ,1)5,(2,3)4)6;", "((0,1)5,(7,3)4)6;"]
    for b in bad:
        with pytest.raises(ValueError):
            p2v.from_newick_batch([b], n_leaves=4)
    # leading zeros make a string longer than the canonical length of its size: rejected (the reference parses them)
    bad.append("((00,1)5,(2,3)004)6;")
    good = ["((0,1)5,(2,3)4)6;", "((0,1),(2,3));", "((0,2)5,(1,3)4)6;", "(0,1)5,(2,3)4)6;"]
    status = torch.zeros(len(bad) + len(good), dtype=torch.int32, device="cuda")
    got = p2v.from_newick_batch(bad + good, n_leaves=4, status=status, check=False).cpu().numpy()
    st = status.cpu().numpy()
    assert (st[:len(bad)] == -2).all() and not st[len(bad):].any()
    assert np.array_equal(got[len(bad):], [[0, 2, 2], [0, 2, 2], [0, 0, 1], [0, 2, 2]])


@pytest.mark.parametrize("n_leaves", SIZES)
@pytest.mark.parametrize("dtype", [torch.int64, torch.int32])
def test_from_newick_batch_parity(p2v, n_leaves, dtype):
    """to_newick strings, and the same trees with subtrees swapped / in another label order, against
    the restatement of parse + order_cherries + build_vector."""
    B = 40
    V = sample_vectors(B, n_leaves, seed=17 * n_leaves)
    V[1] = sample_vectors(1, n_leaves, seed=1, ordered=True)[0]
    adv = list(adversarial_vectors(n_leaves).values())
    V[2:2 + len(adv)] = np.stack(adv)
    strings = [O.to_newick(v) for v in V]
    got = p2v.from_newick_batch(strings, dtype=dtype)
    assert got.dtype == dtype and np.array_equal(got.cpu().numpy().astype(np.int64), V)      # round trip
    # children swapped at every internal node (py tests/test_base.py:75-104 permutes cherries)
    rng = np.random.default_rng(n_leaves)
    swapped = [_swap_children(s, rng) for s in strings]
    ref = np.stack([O.from_newick(s) for s in swapped])
    got = p2v.from_newick_batch(swapped, dtype=dtype)
    assert np.array_equal(got.cpu().numpy().astype(np.int64), ref)
    # without parent labels (order_cherries_no_parents, convert.rs:534-573): the same vectors back,
    # and the swapped trees like the restatement
    import re
    bare = [re.sub(r"\)\d+", ")", s) for s in strings]
    got = p2v.from_newick_batch(bare, dtype=dtype)
    assert np.array_equal(got.cpu().numpy().astype(np.int64), V)
    bare_swapped = [re.sub(r"\)\d+", ")", s) for s in swapped]
    ref = np.stack([O.from_newick(s) for s in bare_swapped])
    got = p2v.from_newick_batch(bare_swapped, dtype=dtype)
    assert np.array_equal(got.cpu().numpy().astype(np.int64), ref)
    # labelled, unlabelled, malformed and half-labelled strings inside one batch
    mixed = list(strings)
    mixed[3] = bare[3]
    mixed[5] = strings[5].replace(",", ";", 1)
    if n_leaves > 2:
        mixed[7] = re.sub(r"\)\d+", ")", strings[7], count=1)
    status = torch.zeros(B, dtype=torch.int32, device="cuda")
    got = p2v.from_newick_batch(mixed, n_leaves=n_leaves, dtype=dtype, status=status, check=False)
    st = status.cpu().numpy()
    bad = [5, 7] if n_leaves > 2 else [5]
    assert all(st[b] == -2 for b in bad) and np.count_nonzero(st) == len(bad)
    keep = [b for b in range(B) if b not in bad]
    assert np.array_equal(got.cpu().numpy().astype(np.int64)[keep], V[keep])
    with pytest.raises(ValueError):
        p2v.from_newick_batch(mixed[:8], n_leaves=n_leaves)


def _swap_children(newick, rng):
    """The same tree with the two children of random internal nodes exchanged."""
    def parse(s, i):
        if s[i] == "(":
            a, i = parse(s, i + 1)
            assert s[i] == ","
            b, i = parse(s, i + 1)
            assert s[i] == ")"
            j = i + 1
            while s[j].isdigit():
                j += 1
            return (a, b, s[i + 1:j]), j
        j = i
        while s[j].isdigit():
            j += 1
        return s[i:j], j

    import sys
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 20000))
    tree, _ = parse(newick, 0)

    def emit(t):
        if isinstance(t, str):
            return t
        a, b, lab = t
        if rng.random() < 0.5:
            a, b = b, a
        return "(" + emit(a) + "," + emit(b) + ")" + lab

    return emit(tree) + ";"


def test_from_newick_host_entry(p2v):
    from phylo2vec_b200 import _capi
    lib = _capi.load()
    n, B = 100, 5000
    V = sample_vectors(B, n, seed=21)
    buf, lengths = O.to_newick_batch(V, int(lib.p2v_newick_max_bytes(n - 1)))
    out = np.zeros((B, n - 1), dtype=np.int64)
    st = np.ones(B, dtype=np.int32)
    vp = ctypes.c_void_p
    rc = lib.p2v_from_newick_host(buf.ctypes.data_as(vp), buf.shape[1], B, n - 1, out.ctypes.data_as(vp), 8,
                                  st.ctypes.data_as(vp))
    assert rc == 0 and not st.any() and np.array_equal(out, V)
    ref, rst = O.from_newick_batch(buf, lengths, n - 1)
    assert not rst.any() and np.array_equal(ref, V)


@pytest.mark.parametrize("n_leaves", [2, 17, 500, 1500, 3000, 5000, 7000])
def test_single_tree_round_trip_unaligned_rows(p2v, n_leaves):
    """The reference-named single-tree functions go through the host entry points with rows that are
    not 16-byte aligned: the CTA-per-tree tokenizer at every size (test_newick, convert.rs:640-656)."""
    import re
    from phylo2vec_b200.base.newick import from_newick, to_newick
    for v in sample_vectors(2, n_leaves, seed=23 * n_leaves):
        s = to_newick(v)
        assert s == O.to_newick(v)
        assert np.array_equal(from_newick(s), v)
        assert np.array_equal(from_newick(re.sub(r"\)\d+", ")", s)), v)
