"""GPU parity tests for the matrix Newick forms (run with -m gpu on a B200): to_newick of matrices byte-exact, and
to_matrix / from_newick with branch lengths bit-exact, against the literal ports of
phylo2vec/src/matrix/convert.rs:26-122 and newick/mod.rs:161-255 (oracle/p2v_oracle.py), which are pinned to
the reference's own tables."""

import ctypes
import re

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import p2v_oracle as O  # noqa: E402
from tests.helpers import adversarial_vectors, sample_bls, sample_vectors, to_matrix_input  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def p2v():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import phylo2vec_b200 as m
    return m


def matrices(B, n, seed):
    V = sample_vectors(B, n, seed=seed)
    adv = list(adversarial_vectors(n).values())
    V[1:1 + min(len(adv), B - 1)] = np.stack(adv)[:B - 1]
    bls = sample_bls(B, n, seed=seed + 1)
    # lengths of every kind of layout: whole numbers, tiny, huge, few digits, zero
    special = np.array([0.0, 1.0, 3.0, 0.5, 1e-7, 1.5e-5, 1e16, 123456.789, 1e300, 5e-324, 0.1, 2.5e-10, 12.0, 1e15, 0.30118185])
    rng = np.random.Generator(np.random.Philox(seed + 2))
    mask = rng.random(bls.shape) < 0.15
    bls[mask] = rng.choice(special, size=int(mask.sum()))
    return to_matrix_input(V, bls)


def swap_children(newick, rng):
    """The same tree with the two children (and their lengths) of random internal nodes exchanged."""
    import sys
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 50000))
    tok = re.compile(r"[^(),;]+")

    def parse(s, i):
        if s[i] == "(":
            a, i = parse(s, i + 1)
            assert s[i] == ","
            b, i = parse(s, i + 1)
            assert s[i] == ")"
            m = tok.match(s, i + 1)
            ann = m.group(0) if m else ""
            return (a, b, ann), i + 1 + len(ann)
        m = tok.match(s, i)
        return m.group(0), m.end()

    tree, _ = parse(newick, 0)

    def emit(t):
        if isinstance(t, str):
            return t
        a, b, ann = t
        if rng.random() < 0.5:
            a, b = b, a
        return "(" + emit(a) + "," + emit(b) + ")" + ann

    return emit(tree) + ";"


def test_goldens_single_tree(p2v, goldens):
    for g in goldens["to_newick_matrix"]:
        assert p2v.to_newick(np.array(g["m"])) == g["newick"], g["src"]
    for g in goldens["from_newick_matrix"]:
        got = p2v.from_newick(g["newick"])
        assert got.dtype == np.float64 and np.array_equal(got, np.array(g["m"])), g["src"]
        assert np.array_equal(p2v.to_matrix(g["newick"]), np.array(g["m"]))
    assert p2v.to_matrix("").shape == (0, 3)
    # vectors still take the vector path; the dispatch is has_branch_lengths (base/newick.py:21-24)
    assert np.array_equal(p2v.from_newick("((0,1)5,(2,3)4)6;"), [0, 2, 2])
    with pytest.raises(AssertionError, match="Branch lengths must be positive"):
        p2v.to_newick(np.array([[0.0, 0.5, -0.8], [1.0, 0.7, 0.6]]))
    with pytest.raises(AssertionError, match="out of bounds"):
        p2v.to_newick(np.array([[0.0, 0.5, 0.8], [3.0, 0.7, 0.6]]))
    for bad in ("(0:0.1,1)2;", "(0:0.1,1:0.2,2:0.3)3;", "(0:0.1,1:abc)2;", "((0:0.1,1:0.2)3,2:0.5)4;", "(0:0.1,1:0.2)2",
                "(0:0.1,5:0.2)2;", "(0:0.1,,1:0.2)2;"):
        with pytest.raises(ValueError):
            p2v.to_matrix(bad)


@pytest.mark.parametrize("n_leaves", [2, 3, 4, 5, 8, 33, 64, 65, 100, 257, 1000, 3000])
def test_to_newick_matrix_batch(p2v, n_leaves):
    B = 24 if n_leaves <= 1000 else 6
    M = matrices(B, n_leaves, seed=n_leaves)
    got = p2v.to_newick_batch(torch.from_numpy(M).cuda())
    want = [O.to_newick_matrix(M[b]) for b in range(B)]
    assert got == want
    # and back: bit-exact matrices, from the strings as written and through the raw device buffers
    back = p2v.to_matrix_batch(got).cpu().numpy()
    assert np.array_equal(back, M)
    buf, ln = p2v.to_newick_batch(torch.from_numpy(M).cuda(), raw=True)
    assert np.array_equal(p2v.to_matrix_batch(buf, lengths=ln, n_leaves=n_leaves).cpu().numpy(), M)
    assert np.array_equal(p2v.to_matrix_batch(buf, n_leaves=n_leaves).cpu().numpy(), M)      # NUL-terminated rows


@pytest.mark.parametrize("n_leaves", [2, 3, 5, 9, 40, 100, 500, 2000])
def test_to_matrix_batch_against_oracle(p2v, n_leaves):
    B = 16 if n_leaves <= 500 else 4
    M = matrices(B, n_leaves, seed=1000 + n_leaves)
    rng = np.random.Generator(np.random.Philox(n_leaves))
    strings = [O.to_newick_matrix(M[b]) for b in range(B)]
    swapped = [swap_children(s, rng) for s in strings]                       # bl_rows_to_swap at work
    bare = [re.sub(r"\)\d+", ")", s) for s in swapped]                        # no parent labels
    # other spellings of the same lengths: exponents, leading / trailing zeros, "%.18e" (numpy.savetxt)
    respelt = [re.sub(r":([0-9.e+-]+)", lambda mm: ":%.18e" % float(mm.group(1)), s) for s in strings]
    for batch in (strings, swapped, bare, respelt):
        want = np.stack([O.from_newick_matrix(s) for s in batch])
        got = p2v.to_matrix_batch(batch).cpu().numpy()
        assert np.array_equal(got, want)
    # the matrix does not depend on the order of the children, with or without labels (matrix/convert.rs:98-113)
    assert np.array_equal(p2v.to_matrix_batch(swapped).cpu().numpy(), M)
    assert np.array_equal(np.stack([O.from_newick_matrix(s) for s in bare])[:, :, 0], M[:, :, 0]) or n_leaves > 2
    # a malformed string inside a batch is flagged; its neighbours are not affected
    if n_leaves >= 3:
        mixed = list(strings)
        mixed[1] = mixed[1].replace(":", "", 1)
        mixed[2] = mixed[2][:-2] + ";"
        st = torch.zeros(B, dtype=torch.int32, device="cuda")
        got = p2v.to_matrix_batch(mixed, status=st, check=False).cpu().numpy()
        s = st.cpu().numpy()
        assert s[1] == -2 and s[0] == 0 and np.array_equal(got[0], M[0]) and np.array_equal(got[3], M[3])
        with pytest.raises(ValueError):
            p2v.to_matrix_batch(mixed)


@pytest.mark.parametrize("n_leaves", [5, 40, 300])
def test_long_literals_are_read_exactly(p2v, n_leaves, tmp_path):
    """Branch lengths of more than 19 significant digits, many of them exactly half way between two doubles or a
    hair beside it: the fast parser leaves those undecided and parse_f64_exact settles them -- the matrix is the one
    Python's (and Rust's) correctly rounded float() gives, bit for bit.  The same through the CSV reader."""
    from tests.test_float_text import hard_long_literals
    B = 6
    M = matrices(B, n_leaves, seed=77 + n_leaves)
    rng = np.random.Generator(np.random.Philox(n_leaves))
    lits = [x for x in hard_long_literals(rng, 400, 300) if not x.startswith("-") and len(x) < 900]
    strings = []
    for b in range(B):
        it = iter(rng.permutation(len(lits)).tolist())
        strings.append(re.sub(r":([0-9.e+-]+)", lambda mm: ":" + lits[next(it)], O.to_newick_matrix(M[b])))
    want = np.stack([O.from_newick_matrix(s) for s in strings])              # Python float(): correctly rounded
    got = p2v.to_matrix_batch(strings).cpu().numpy()
    assert np.array_equal(got, want)
    # CSV: one flattened matrix per line, lengths written as the same long literals
    k = n_leaves - 1
    lines = []
    lits = [x for x in lits if len(x) < 500]                                # the CSV reader's fields are shorter than 512 bytes
    for b in range(B):
        it = iter(rng.permutation(len(lits)).tolist())
        lines.append(",".join(f"{int(want[b, r, 0])},{lits[next(it)]},{lits[next(it)]}" for r in range(k)))
    path = tmp_path / "long.csv"
    path.write_text("\n".join(lines) + "\n")
    from phylo2vec_b200.io import reader
    got_csv = reader.load_batch(str(path), kind="matrix", check=False).cpu().numpy()
    want_csv = np.array([[float(x) for x in ln.split(",")] for ln in lines]).reshape(B, k, 3)
    assert np.array_equal(got_csv, want_csv)


def test_host_entry_points(p2v):
    from phylo2vec_b200 import _capi
    lib = _capi.load()
    n, B = 300, 500
    k = n - 1
    M = matrices(B, n, seed=5)
    stride = int(lib.p2v_newick_matrix_max_bytes(k))
    out = np.zeros((B, stride), dtype=np.uint8)
    ln = np.zeros(B, dtype=np.int64)
    st = np.zeros(B, dtype=np.int32)
    cp = lambda a: a.ctypes.data_as(ctypes.c_void_p)       # noqa: E731
    assert lib.p2v_to_newick_matrix_host(cp(M), B, k, cp(out), stride, cp(ln), cp(st)) == 0 and not st.any()
    for b in (0, 7, B - 1):
        assert out[b, :ln[b]].tobytes().decode() == O.to_newick_matrix(M[b])
    back = np.zeros_like(M)
    assert lib.p2v_from_newick_matrix_host(cp(out), stride, B, k, cp(back), cp(st)) == 0 and not st.any()
    assert np.array_equal(back, M)
