"""GPU parity tests for decode/encode (run with -m gpu on a B200).

The CUDA path (through the C ABI) is compared bit-exactly with the CPU oracle on
the same seeded inputs, against the reference's golden vectors, and -- at
BASELINE.json's full size -- through round-trip properties."""

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


# ---------------------------------------------------------------- goldens, single-tree drop-ins
def test_goldens_single_tree(p2v, goldens):
    for g in goldens["to_pairs"]:
        got = p2v.to_pairs(np.array(g["v"]))
        assert isinstance(got, list) and isinstance(got[0], tuple)
        assert got == [tuple(x) for x in g["pairs"]], g["src"]
    for g in goldens["to_ancestry"]:
        got = p2v.to_ancestry(np.array(g["v"]))
        assert got.dtype == np.int64 and got.shape == (len(g["v"]), 3)
        assert np.array_equal(got, np.array(g["ancestry"])), g["src"]
    for g in goldens["to_edges"]:
        assert p2v.to_edges(g["v"]) == [tuple(x) for x in g["edges"]], g["src"]
    for g in goldens["from_pairs"]:
        got = p2v.from_pairs([tuple(x) for x in g["pairs"]])
        assert got.dtype == np.int64 and np.array_equal(got, g["v"]), g["src"]
    for g in goldens["from_ancestry"]:
        assert np.array_equal(p2v.from_ancestry(np.array(g["ancestry"])), g["v"]), g["src"]
    for g in goldens["from_edges"]:
        assert np.array_equal(p2v.from_edges([tuple(x) for x in g["edges"]]), g["v"]), g["src"]
    for g in goldens["build_vector"]:
        pairs = np.array(g["cherries"])[:, :2]
        if (pairs.max(axis=1) == np.array(g["cherries"])[:, 2]).all():
            assert np.array_equal(p2v.from_pairs(pairs), g["v"]), g["src"]


def test_check_v(p2v, goldens):
    for v in goldens["check_v"]["ok"]:
        p2v.check_v(v)
    for v in goldens["check_v"]["panic"]:
        with pytest.raises(AssertionError, match="out of bounds"):
            p2v.check_v(v)
    with pytest.raises(AssertionError) as e:
        p2v.to_ancestry([0, 0, 9, 1])
    assert str(e.value) == "Validation failed: v[2] = 9 is out of bounds (max = 4)"
    with pytest.raises(OverflowError):
        p2v.to_pairs([0, -1])
    with pytest.raises(TypeError):
        p2v.to_pairs(np.array([0.0, 1.0]))
    with pytest.raises(AssertionError, match="even"):
        p2v.from_edges([(2, 4), (3, 4), (0, 5)])


def test_empty_and_tiny(p2v):
    assert p2v.to_pairs(np.zeros(0, dtype=np.int64)) == []
    assert p2v.to_ancestry(np.zeros(0, dtype=np.int64)).shape == (0, 3)
    assert p2v.to_pairs([0]) == [(0, 1)]
    assert np.array_equal(p2v.to_ancestry([0]), [[0, 1, 2]])
    assert p2v.to_edges([0]) == [(0, 2), (1, 2)]
    assert np.array_equal(p2v.from_pairs([(0, 1)]), [0])
    assert np.array_equal(p2v.from_ancestry([[0, 1, 2]]), [0])
    assert np.array_equal(p2v.from_edges([(0, 2), (1, 2)]), [0])
    B0 = torch.zeros((0, 7), dtype=torch.int64, device="cuda")
    assert tuple(p2v.to_ancestry_batch(B0).shape) == (0, 7, 3)


# ---------------------------------------------------------------- batched parity vs oracle
SIZES = [2, 3, 4, 5, 17, 32, 33, 34, 64, 65, 100, 257, 1000, 1024, 1025, 1026, 1500, 2000, 2016, 2017, 2018, 2048, 2049, 5000, 12288, 12289, 20000]


@pytest.mark.parametrize("n_leaves", SIZES)
@pytest.mark.parametrize("dtype", [torch.int64, torch.int32])
def test_decode_parity(p2v, n_leaves, dtype):
    B = 37 if n_leaves <= 1100 else 9
    V = sample_vectors(B, n_leaves, seed=n_leaves)
    V[1] = sample_vectors(1, n_leaves, seed=1, ordered=True)[0]
    adv = list(adversarial_vectors(n_leaves).values())
    V[2:2 + len(adv)] = np.stack(adv)
    k = n_leaves - 1
    Vd = dev(V, dtype)
    for name, fn in (("to_pairs", p2v.to_pairs_batch), ("to_ancestry", p2v.to_ancestry_batch),
                     ("to_edges", p2v.to_edges_batch)):
        ref, st = O.batch(name, V, k)
        assert not st.any()
        got = fn(Vd)
        assert got.dtype == dtype and got.is_cuda
        assert np.array_equal(got.cpu().numpy().astype(np.int64), ref), (name, n_leaves)


# One to four huge trees take the whole-grid (cooperative) form of the merge kernel: every SM works on the
# same tree; slots that fit 16 bits (k <= 65535) use the per-CTA shared-memory label tables, larger
# trees the per-worker tables in the workspace.  Sizes straddle both limits; adversarial shapes
# (one label for every row, one row per label) stress the cross-worker prefix.
@pytest.mark.parametrize("n_leaves", [8193, 8200, 12289, 20000, 40000, 65535, 65536, 65537, 70000, 100000])
def test_decode_parity_whole_grid(p2v, n_leaves):
    k = n_leaves - 1
    adv = adversarial_vectors(n_leaves)
    cases = [sample_vectors(1, n_leaves, seed=n_leaves)[0], sample_vectors(1, n_leaves, seed=3, ordered=True)[0],
             adv["ladder_zero"], adv["two_i"], adv["identity"], adv["alt"]]
    for B in (1, 2, 4):
        for first in range(0, len(cases), B):
            V = np.stack([cases[(first + j) % len(cases)] for j in range(B)])
            for dtype in ((torch.int64, torch.int32) if B == 1 else (torch.int32,)):
                Vd = dev(V, dtype)
                for name, fn in (("to_ancestry", p2v.to_ancestry_batch), ("to_pairs", p2v.to_pairs_batch),
                                 ("to_edges", p2v.to_edges_batch)):
                    if name != "to_ancestry" and (first or B == 2):
                        continue
                    ref, st = O.batch(name, V, k)
                    assert not st.any()
                    got = fn(Vd)
                    assert np.array_equal(got.cpu().numpy().astype(np.int64), ref), (name, n_leaves, B, first)
    # an invalid entry anywhere in a huge tree is reported with its index
    bad = cases[0].copy()
    bad[k // 2] = k + 7 if k // 2 < (k + 7) // 2 else 2 * (k // 2) + 1
    bad[k // 2] = 2 * (k // 2) + 1
    st = torch.zeros(1, dtype=torch.int32, device="cuda")
    p2v.to_ancestry_batch(dev(bad[None], torch.int64), status=st, check=False)
    assert int(st[0]) == k // 2 + 1


# Batches of big trees (16384 < k <= 262143, at least 38 of them) take the warp-per-tree kernel with the mask in
# shared memory and the label table in the workspace (p2v_decode_warp_big.cuh; uint32 table entries above 65535
# slots); beyond that the merge kernel.
@pytest.mark.parametrize("n_leaves", [16386, 16400, 20000, 33000, 65535, 65536, 65537, 65570, 100000, 262144, 262145])
def test_decode_parity_big_batches(p2v, n_leaves):
    k = n_leaves - 1
    B = 40
    V = sample_vectors(B, n_leaves, seed=n_leaves)
    V[1] = sample_vectors(1, n_leaves, seed=1, ordered=True)[0]
    adv = list(adversarial_vectors(n_leaves).values())
    V[2:2 + len(adv)] = np.stack(adv)
    for dtype in (torch.int64, torch.int32):
        Vd = dev(V, dtype)
        for name, fn in (("to_ancestry", p2v.to_ancestry_batch), ("to_pairs", p2v.to_pairs_batch),
                         ("to_edges", p2v.to_edges_batch)):
            if dtype == torch.int32 and name == "to_edges":
                continue
            if n_leaves > 70000 and (dtype, name) not in ((torch.int64, "to_ancestry"), (torch.int32, "to_pairs")):
                continue                                   # the oracle takes seconds per call at these sizes
            ref, st = O.batch(name, V, k)
            assert not st.any()
            got = fn(Vd)
            assert np.array_equal(got.cpu().numpy().astype(np.int64), ref), (name, n_leaves, dtype)
    bad = V.copy()
    bad[7, k - 3] = 2 * (k - 3) + 1
    bad[9, 5] = 11
    bad[9, 100] = 999
    st = torch.zeros(B, dtype=torch.int32, device="cuda")
    p2v.to_ancestry_batch(dev(bad, torch.int64), status=st, check=False)
    want = np.zeros(B, dtype=np.int32)
    want[7] = k - 3 + 1
    want[9] = 5 + 1
    assert np.array_equal(st.cpu().numpy(), want)


# Batches of mid-size trees stay on the warp-per-tree kernel (4 / 8 / 16 mask words per lane,
# integer rank scan); a handful of such trees goes to the merge kernel (test_decode_parity above).
MID_SIZES = [2017, 2049, 3000, 4096, 4097, 4098, 5000, 8192, 8193, 8194, 12289, 16384, 16385, 16386]


@pytest.mark.parametrize("n_leaves", MID_SIZES)
@pytest.mark.parametrize("dtype", [torch.int64, torch.int32])
def test_decode_parity_mid_batches(p2v, n_leaves, dtype):
    B = 48
    V = sample_vectors(B, n_leaves, seed=7 * n_leaves)
    V[1] = sample_vectors(1, n_leaves, seed=1, ordered=True)[0]
    adv = list(adversarial_vectors(n_leaves).values())
    V[2:2 + len(adv)] = np.stack(adv)
    k = n_leaves - 1
    Vd = dev(V, dtype)
    for name, fn in (("to_pairs", p2v.to_pairs_batch), ("to_ancestry", p2v.to_ancestry_batch),
                     ("to_edges", p2v.to_edges_batch)):
        ref, st = O.batch(name, V, k)
        assert not st.any()
        got = fn(Vd)
        assert np.array_equal(got.cpu().numpy().astype(np.int64), ref), (name, n_leaves)
    # the same trees through the few-trees path give the same bits
    assert torch.equal(p2v.to_ancestry_batch(Vd[:5]), p2v.to_ancestry_batch(Vd)[:5])
    V[3, k // 2] = 2 * (k // 2) + 1
    status = torch.zeros(B, dtype=torch.int32, device="cuda")
    p2v.to_ancestry_batch(dev(V, dtype), status=status, check=False)
    assert status[3].item() == k // 2 + 1 and int(status.count_nonzero()) == 1


@pytest.mark.parametrize("n_leaves", [2048, 2049, 2050, 3000, 4096, 4097, 4098, 5000, 8192, 8193, 8194, 9000])
@pytest.mark.parametrize("dtype", [torch.int64, torch.int32])
def test_encode_parity_mid_batches(p2v, n_leaves, dtype):
    """Batches of mid-size trees on the warp-per-tree encode kernel (4 / 8 mask words per lane)."""
    B = 40
    V = sample_vectors(B, n_leaves, seed=13 * n_leaves)
    adv = list(adversarial_vectors(n_leaves).values())
    V[:len(adv)] = np.stack(adv)
    k = n_leaves - 1
    pairs, _ = O.batch("to_pairs", V, k)
    anc, _ = O.batch("to_ancestry", V, k)
    edges, _ = O.batch("to_edges", V, k)
    rng = np.random.default_rng(n_leaves)
    perm = rng.permutation(k)
    shuffled = edges.reshape(B, k, 2, 2)[:, perm].reshape(B, 2 * k, 2)     # from_edges orders by parent itself
    for name, fn, x in (("from_pairs", p2v.from_pairs_batch, pairs), ("from_ancestry", p2v.from_ancestry_batch, anc),
                        ("from_edges", p2v.from_edges_batch, edges), ("from_edges", p2v.from_edges_batch, shuffled)):
        got = fn(dev(x, dtype))
        assert got.dtype == dtype
        assert np.array_equal(got.cpu().numpy().astype(np.int64), V), (name, n_leaves)
    bad = anc.copy()
    bad[5, 3, 2] = 1                      # a leaf label as parent
    status = torch.zeros(B, dtype=torch.int32, device="cuda")
    p2v.from_ancestry_batch(dev(bad, dtype), status=status, check=False)
    assert status[5].item() == -2 and int(status.count_nonzero()) == 1


@pytest.mark.parametrize("n_leaves", [16385, 16386, 20000, 32768, 32769, 65535, 65536, 65537, 100000, 262144, 262145])
def test_encode_parity_big_batches(p2v, n_leaves):
    """Batches of big trees on the warp-per-tree encoder with its label table in the workspace (above 65536 leaves:
    64-bit parked words for int64 rows, the CTA kernel for int32)."""
    B = 150 if n_leaves <= 20000 else 40                  # more trees than SMs / 4 for the smaller sizes
    V = sample_vectors(B, n_leaves, seed=17 * n_leaves)
    adv = list(adversarial_vectors(n_leaves).values())
    V[:len(adv)] = np.stack(adv)
    k = n_leaves - 1
    rng = np.random.default_rng(n_leaves)
    perm = torch.from_numpy(rng.permutation(k)).cuda()
    for dtype in (torch.int64, torch.int32):
        Vd = dev(V, dtype)
        anc = p2v.to_ancestry_batch(Vd)                    # checked against the oracle by the decode tests
        assert torch.equal(p2v.from_ancestry_batch(anc), Vd), (n_leaves, dtype)
        assert torch.equal(p2v.from_pairs_batch(p2v.to_pairs_batch(Vd)), Vd), (n_leaves, dtype)
        edges = p2v.to_edges_batch(Vd)
        assert torch.equal(p2v.from_edges_batch(edges), Vd), (n_leaves, dtype)
        shuffled = edges.reshape(B, k, 2, 2)[:, perm].reshape(B, 2 * k, 2).contiguous()
        assert torch.equal(p2v.from_edges_batch(shuffled), Vd), (n_leaves, dtype)
    ref, st = O.batch("from_ancestry", anc[:3].cpu().numpy().astype(np.int64), k)
    assert not st.any() and np.array_equal(ref, V[:3])
    bad = anc.clone()
    bad[2, 7, 2] = 1                                      # a leaf label as parent
    bad[4, 9] = bad[4, 8]                                 # one parent label, two rows
    status = torch.zeros(B, dtype=torch.int32, device="cuda")
    p2v.from_ancestry_batch(bad, status=status, check=False)
    assert status[2].item() == -2 and status[4].item() == -2 and int(status.count_nonzero()) == 2


@pytest.mark.parametrize("n_leaves", SIZES)
@pytest.mark.parametrize("dtype", [torch.int64, torch.int32])
def test_encode_parity_and_round_trip(p2v, n_leaves, dtype):
    B = 21 if n_leaves <= 1100 else 5
    V = sample_vectors(B, n_leaves, seed=1000 + n_leaves)
    adv = list(adversarial_vectors(n_leaves).values())
    V[:len(adv)] = np.stack(adv)
    k = n_leaves - 1
    pairs, _ = O.batch("to_pairs", V, k)
    anc, _ = O.batch("to_ancestry", V, k)
    edges, _ = O.batch("to_edges", V, k)
    for name, fn, x in (("from_pairs", p2v.from_pairs_batch, pairs), ("from_ancestry", p2v.from_ancestry_batch, anc),
                        ("from_edges", p2v.from_edges_batch, edges)):
        ref, st = O.batch(name, x, k)
        assert not st.any() and np.array_equal(ref, V)
        got = fn(dev(x, dtype))
        assert got.dtype == dtype
        assert np.array_equal(got.cpu().numpy().astype(np.int64), V), (name, n_leaves)
    # GPU -> GPU round trips
    Vd = dev(V, dtype)
    assert torch.equal(p2v.from_pairs_batch(p2v.to_pairs_batch(Vd)), Vd)
    assert torch.equal(p2v.from_ancestry_batch(p2v.to_ancestry_batch(Vd)), Vd)
    assert torch.equal(p2v.from_edges_batch(p2v.to_edges_batch(Vd)), Vd)


@pytest.mark.parametrize("n_leaves", [50, 1000, 3000])
def test_encode_accepts_permuted_rows(p2v, n_leaves):
    """from_edges reorders cherries by parent label (order_cherries, convert.rs:449-477)."""
    V = sample_vectors(6, n_leaves, seed=7)
    k = n_leaves - 1
    edges, _ = O.batch("to_edges", V, k)
    rng = np.random.default_rng(0)
    shuffled = np.empty_like(edges)
    for b in range(V.shape[0]):
        perm = rng.permutation(k)
        e = edges[b].reshape(k, 2, 2)[perm].reshape(2 * k, 2)
        shuffled[b] = e
    ref, st = O.batch("from_edges", shuffled, k)
    assert not st.any()
    got = p2v.from_edges_batch(dev(shuffled)).cpu().numpy()
    assert np.array_equal(got, ref) and np.array_equal(got, V)


@pytest.mark.parametrize("n_leaves", [10, 1000, 2500])
def test_invalid_vectors_flagged(p2v, n_leaves):
    V = sample_vectors(8, n_leaves, seed=3)
    k = n_leaves - 1
    V[2, k // 2] = 2 * (k // 2) + 1
    V[5, 0] = 1
    V[6, k - 1] = 10 * k
    Vd = dev(V)
    status = torch.zeros(8, dtype=torch.int32, device="cuda")
    out = p2v.to_ancestry_batch(Vd, status=status, check=False)
    st = status.cpu().numpy()
    assert list(st) == [0, 0, k // 2 + 1, 0, 0, 1, k, 0]
    _, ost = O.batch("to_ancestry", V, k)
    assert np.array_equal(st, ost)
    good = [0, 1, 3, 4, 7]
    ref, _ = O.batch("to_ancestry", V[good], k)
    assert np.array_equal(out.cpu().numpy()[good], ref)
    assert np.array_equal(p2v.check_v_batch(Vd).cpu().numpy(), st)
    with pytest.raises(AssertionError, match=r"v\[%d\]" % (k // 2)):
        p2v.to_pairs_batch(Vd)


def test_malformed_encode_input_flagged(p2v):
    V = sample_vectors(4, 40, seed=9)
    k = 39
    anc, _ = O.batch("to_ancestry", V, k)
    bad = anc.copy()
    bad[1, 5, 0] = 10 * k          # label out of range
    bad[2, 3] = bad[2, 4]          # duplicated cherry -> c_max not a permutation
    status = torch.zeros(4, dtype=torch.int32, device="cuda")
    out = p2v.from_ancestry_batch(dev(bad), status=status, check=False)
    st = status.cpu().numpy()
    assert st[0] == 0 and st[3] == 0 and st[1] == -2 and st[2] == -2
    assert np.array_equal(out.cpu().numpy()[[0, 3]], V[[0, 3]])
    with pytest.raises(ValueError):
        p2v.from_ancestry_batch(dev(bad))


def test_dlpack_handoff(p2v):
    V = dev(sample_vectors(5, 30, seed=2))

    class Capsule:
        def __init__(self, t):
            self.t = t

        def __dlpack__(self, stream=None):
            return self.t.__dlpack__()

        def __dlpack_device__(self):
            return self.t.__dlpack_device__()

    a = p2v.to_ancestry_batch(Capsule(V))
    b = p2v.to_ancestry_batch(V)
    assert torch.equal(a, b)
    back = torch.from_dlpack(a)          # consumer side: zero copy
    assert back.data_ptr() == a.data_ptr()


def test_non_default_stream_and_out_buffers(p2v):
    V = dev(sample_vectors(64, 500, seed=11))
    out = torch.empty((64, 499, 3), dtype=torch.int64, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        got = p2v.to_ancestry_batch(V, out=out, check=False)
    s.synchronize()
    assert got.data_ptr() == out.data_ptr()
    assert torch.equal(out, p2v.to_ancestry_batch(V))


# ---------------------------------------------------------------- BASELINE config C2 at full size
def test_full_size_round_trip_properties(p2v):
    """1M trees, n=1000 (BASELINE.json configs[1]): decode then encode returns v;
    structural invariants of the ancestry hold for every tree; a sample is compared
    with the oracle."""
    B, n = 1_000_000, 1000
    k = n - 1
    g = torch.Generator(device="cuda").manual_seed(0)
    hi = (2 * torch.arange(k, device="cuda") + 1).to(torch.float64)
    V = (torch.rand((B, k), generator=g, device="cuda", dtype=torch.float64) * hi).to(torch.int64)
    anc = p2v.to_ancestry_batch(V, check=False)
    # parent column is k+1+row for every tree
    assert bool((anc[:, :, 2] == (torch.arange(k, device="cuda") + k + 1)).all())
    # every non-root node is a child exactly once: sum and sum of squares of 0..2k-1
    kids = anc[:, :, :2].reshape(B, -1)
    assert bool((kids.sum(1) == (2 * k - 1) * (2 * k) // 2).all())
    assert bool(((kids * kids).sum(1) == (2 * k - 1) * (2 * k) * (4 * k - 1) // 6).all())
    back = p2v.from_ancestry_batch(anc, check=False)
    assert torch.equal(back, V)
    del back
    idx = torch.arange(0, B, B // 64, device="cuda")[:64]
    ref, _ = O.batch("to_ancestry", V[idx].cpu().numpy(), k)
    assert np.array_equal(anc[idx].cpu().numpy(), ref)
    del anc
    pairs = p2v.to_pairs_batch(V[:200_000], check=False)
    assert torch.equal(p2v.from_pairs_batch(pairs, check=False), V[:200_000])
    del pairs
    edges = p2v.to_edges_batch(V[:200_000], check=False)
    assert torch.equal(p2v.from_edges_batch(edges, check=False), V[:200_000])


def test_host_entry_points_batched(p2v):
    """C-ABI *_host functions with host buffers (what a Rust/ctypes caller binds)."""
    import ctypes
    from phylo2vec_b200 import _capi
    lib = _capi.load()
    V = sample_vectors(300, 128, seed=4)
    k = 127
    out = np.zeros((300, k, 3), dtype=np.int64)
    st = np.ones(300, dtype=np.int32)
    rc = lib.p2v_to_ancestry_host(V.ctypes.data_as(ctypes.c_void_p), 8, 300, k,
                                  out.ctypes.data_as(ctypes.c_void_p), st.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0 and not st.any()
    ref, _ = O.batch("to_ancestry", V, k)
    assert np.array_equal(out, ref)
    back = np.zeros((300, k), dtype=np.int64)
    rc = lib.p2v_from_ancestry_host(out.ctypes.data_as(ctypes.c_void_p), 8, 300, k,
                                    back.ctypes.data_as(ctypes.c_void_p), None)
    assert rc == 0 and np.array_equal(back, V)
    V32 = V.astype(np.int32)
    out32 = np.zeros((300, 2 * k, 2), dtype=np.int32)
    rc = lib.p2v_to_edges_host(V32.ctypes.data_as(ctypes.c_void_p), 4, 300, k,
                               out32.ctypes.data_as(ctypes.c_void_p), None)
    assert rc == 0 and np.array_equal(out32.astype(np.int64), O.batch("to_edges", V, k)[0])


def test_host_to_ancestry_compact_transport(p2v):
    """Big int64 batches through p2v_to_ancestry_host travel as two int32 columns and are widened by
    host threads of the library (several chunks, pageable buffers, an invalid tree in the middle);
    the result must be what the device path and the oracle give."""
    import ctypes
    from phylo2vec_b200 import _capi
    lib = _capi.load()
    vp = ctypes.c_void_p
    # labels below 65536 travel as uint16 pairs, larger trees (n = 40000) as int32 pairs
    for n_leaves, B in ((1000, 120_000), (100, 60_000), (3000, 3_000), (20000, 300), (40000, 120)):
        k = n_leaves - 1
        V = sample_vectors(B, n_leaves, seed=n_leaves + 5)
        bad = B // 2 + 1
        V[bad, k // 3] = 2 * (k // 3) + 7
        out = np.full((B, k, 3), -1, dtype=np.int64)
        st = np.full(B, 77, dtype=np.int32)
        rc = lib.p2v_to_ancestry_host(V.ctypes.data_as(vp), 8, B, k, out.ctypes.data_as(vp), st.ctypes.data_as(vp))
        assert rc == 0, lib.p2v_last_error()
        assert st[bad] == k // 3 + 1 and np.count_nonzero(st) == 1
        idx = np.concatenate([np.arange(0, 40), np.linspace(40, B - 1, 200).astype(int)])
        idx = idx[idx != bad]
        ref, _ = O.batch("to_ancestry", V[idx], k)
        assert np.array_equal(out[idx], ref), n_leaves
        dev_out = p2v.to_ancestry_batch(dev(V), check=False).cpu().numpy()
        good = np.ones(B, dtype=bool)
        good[bad] = False
        assert np.array_equal(out[good], dev_out[good]), n_leaves
    lib.p2v_host_release()


def test_numpy_batches_go_through_the_host_entry_points(p2v):
    """numpy in, numpy out for the batched functions (no torch tensors involved)."""
    n, B = 700, 9000                      # 6.3 M elements: the compact int64 transport for to_ancestry
    k = n - 1
    V = sample_vectors(B, n, seed=21)
    anc = p2v.to_ancestry_batch(V)
    assert isinstance(anc, np.ndarray) and anc.dtype == np.int64 and anc.shape == (B, k, 3)
    idx = np.linspace(0, B - 1, 64).astype(int)
    assert np.array_equal(anc[idx], O.batch("to_ancestry", V[idx], k)[0])
    assert np.array_equal(p2v.to_pairs_batch(V[:50]), O.batch("to_pairs", V[:50], k)[0])
    assert np.array_equal(p2v.to_edges_batch(V[:50].astype(np.int32)).astype(np.int64), O.batch("to_edges", V[:50], k)[0])
    assert np.array_equal(p2v.from_ancestry_batch(anc), V)
    # big pageable buffers: threaded staging through pinned memory, several chunks
    edges = p2v.to_edges_batch(V)
    assert np.array_equal(edges[idx], O.batch("to_edges", V[idx], k)[0])
    assert np.array_equal(p2v.from_edges_batch(edges), V)
    pairs = p2v.to_pairs_batch(V)
    assert np.array_equal(pairs[idx], O.batch("to_pairs", V[idx], k)[0])
    assert np.array_equal(p2v.from_pairs_batch(pairs), V)
    assert np.array_equal(p2v.from_pairs_batch(p2v.to_pairs_batch(V[:50])), V[:50])
    assert np.array_equal(p2v.from_edges_batch(p2v.to_edges_batch(V[:50])), V[:50])
    bad = V[:8].copy()
    bad[3, 5] = 11
    with pytest.raises(AssertionError, match=r"v\[5\] = 11"):
        p2v.to_ancestry_batch(bad)
    with pytest.raises(TypeError):
        p2v.to_ancestry_batch(V.astype(np.float64))
