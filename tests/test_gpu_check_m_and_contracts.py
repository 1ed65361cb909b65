"""GPU tests: check_m (phylo2vec/src/matrix/base.rs:76-95) against the oracle and the reference's
test table, the per-tree status for negative branch lengths on the distance path, the alignment /
status-tensor contract of the C ABI (include/phylo2vec_b200.h "Alignment"), the host-cache limit,
and re-entrancy: the ABI called from two host threads (and two devices when present) at once."""

import ctypes
import threading

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import p2v_oracle as O  # noqa: E402
from tests.helpers import sample_bls, sample_vectors, to_matrix_input  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def p2v():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import phylo2vec_b200 as m
    return m


MSG = {"columns": "Matrix must have at least 3 columns", "v": "out of bounds", "negative": "Branch lengths must be positive"}


def test_check_matrix_reference_table(p2v, goldens):
    for m in goldens["check_m"]["ok"]:
        p2v.check_matrix(np.array(m, dtype=np.float64))
    for g in goldens["check_m"]["panic"]:
        with pytest.raises(AssertionError, match=MSG[g["why"]]) as e:
            p2v.check_matrix(np.array(g["m"], dtype=np.float64))
        with pytest.raises(AssertionError) as eo:          # same message as the oracle's restatement
            O.check_m(g["m"])
        assert str(e.value) == str(eo.value)
    with pytest.raises(AssertionError, match="Matrix should not be empty"):
        p2v.check_matrix(np.zeros((1, 0)))
    with pytest.raises(TypeError):
        p2v.check_matrix(np.zeros((3, 3), dtype=np.float32))
    with pytest.raises(TypeError):
        p2v.check_matrix(np.zeros(3))


def test_check_m_batch_against_oracle(p2v):
    from phylo2vec_b200 import _capi
    rng = np.random.Generator(np.random.Philox(11))
    for n in (2, 3, 33, 100, 1000, 5000):
        k = n - 1
        B = 96
        V = sample_vectors(B, n, seed=n)
        bls = sample_bls(B, n, seed=n + 1)
        M = to_matrix_input(V, bls)
        want = np.zeros(B, dtype=np.int32)
        for b in range(0, B, 3):               # every third tree is broken in one of four ways
            kind = (b // 3) % 4
            i = int(rng.integers(0, k))
            if kind == 0 and i > 0:
                M[b, i, 0] = 2 * i + 1 + int(rng.integers(0, 5))
                want[b] = i + 1
            elif kind == 1:
                M[b, i, 1 + int(rng.integers(0, 2))] = -1e-300
                want[b] = _capi.P2V_TREE_NEGATIVE_BRANCH
            elif kind == 2:
                M[b, i, 2] = np.nan
                want[b] = _capi.P2V_TREE_NEGATIVE_BRANCH
            elif kind == 3:
                M[b, i, 0] = -3.0                 # `as usize` saturates to 0: valid
                M[b, i, 1] = 0.0                  # zero is allowed
        got = p2v.check_m_batch(torch.from_numpy(M).cuda()).cpu().numpy()
        assert np.array_equal(got, want), n
        for b in range(B):                        # and the oracle agrees tree by tree
            try:
                O.check_m(M[b])
                ok = True
            except AssertionError:
                ok = False
            assert ok == (want[b] == 0)
        # host entry point
        st = np.zeros(B, dtype=np.int32)
        rc = _capi.load().p2v_check_m_host(M.ctypes.data_as(ctypes.c_void_p), B, k, st.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0 and np.array_equal(st, want)


@pytest.mark.parametrize("n_leaves", [2, 40, 200, 1000, 20000])
def test_negative_lengths_flagged_on_the_distance_path(p2v, n_leaves):
    """The reference's cophenetic_distances does not call check_m; here a tree with a negative or NaN
    length gets status P2V_TREE_NEGATIVE_BRANCH (the three-term distance assumes depth grows
    downwards) and the other trees of the batch are unaffected."""
    from phylo2vec_b200 import _capi
    k = n_leaves - 1
    B = 4 if n_leaves > 5000 else 12
    V = sample_vectors(B, n_leaves, seed=5)
    bls = sample_bls(B, n_leaves, seed=6)
    M = to_matrix_input(V, bls)
    M[1, k // 2, 1] = -0.25
    M[B - 1, 0, 2] = np.nan
    Md = torch.from_numpy(M).cuda()
    st = torch.zeros(B, dtype=torch.int32, device="cuda")
    D = p2v.cophenetic_distances_batch(Md, status=st, check=False)
    got = st.cpu().numpy()
    want = np.zeros(B, dtype=np.int32)
    want[1] = want[B - 1] = _capi.P2V_TREE_NEGATIVE_BRANCH
    assert np.array_equal(got, want)
    if n_leaves <= 1000:
        ref = O.cophenetic_distances(M[0])
        assert np.max(np.abs(D[0].cpu().numpy() - ref) / np.maximum(ref, 1e-300)) <= 1e-12
    with pytest.raises(AssertionError, match="Branch lengths must be positive"):
        p2v.cophenetic_distances_batch(Md)
    with pytest.raises(AssertionError, match="Branch lengths must be positive"):
        p2v.cophenetic_distances(M[1])
    # bls given next to an integer vector take the same guard
    stv = torch.zeros(B, dtype=torch.int32, device="cuda")
    p2v.cophenetic_distances_batch(torch.from_numpy(V).cuda(), bls=torch.from_numpy(np.ascontiguousarray(M[:, :, 1:])).cuda(),
                                   status=stv, check=False)
    assert np.array_equal(stv.cpu().numpy(), want)


def test_alignment_contract(p2v):
    """pairs / edges outputs are written as two-element vector stores: a buffer at an odd element
    offset is rejected with an error instead of faulting the context."""
    from phylo2vec_b200 import _capi
    lib = _capi.load()
    B, n = 8, 50
    k = n - 1
    V = torch.from_numpy(sample_vectors(B, n, seed=1)).cuda()
    for dtype in (torch.int64, torch.int32):
        Vt = V.to(dtype)
        es = Vt.element_size()
        big = torch.zeros(B * k * 4 + 1, dtype=dtype, device="cuda")
        odd = big[1:1 + B * k * 2]
        assert odd.data_ptr() % (2 * es) != 0
        st = torch.zeros(B, dtype=torch.int32, device="cuda")
        rc = lib.p2v_to_pairs_batch(Vt.data_ptr(), es, B, k, odd.data_ptr(), st.data_ptr(), None, 0,
                                    torch.cuda.current_stream().cuda_stream)
        assert rc == -1 and b"aligned" in lib.p2v_last_error()
        rc = lib.p2v_to_edges_batch(Vt.data_ptr(), es, B, k, big[1:].data_ptr(), st.data_ptr(), None, 0,
                                    torch.cuda.current_stream().cuda_stream)
        assert rc == -1
        with pytest.raises(ValueError, match="aligned"):
            p2v.to_pairs_batch(Vt, out=odd.view(B, k, 2))
        # ancestry rows are scalar stores: an odd element offset is fine and the result is right
        anc_odd = big[1:1 + B * k * 3].view(B, k, 3)
        got = p2v.to_ancestry_batch(Vt, out=anc_odd)
        ref, _ = O.batch("to_ancestry", V.cpu().numpy(), k)
        assert np.array_equal(got.cpu().numpy(), ref)
    torch.cuda.synchronize()          # the context is still healthy
    assert np.array_equal(p2v.to_pairs_batch(V).cpu().numpy(), O.batch("to_pairs", V.cpu().numpy(), k)[0])


@pytest.mark.parametrize("n_leaves", [130, 131, 200, 257, 320])
def test_staged_branch_lengths_at_any_alignment(p2v, n_leaves):
    """For 128 < n <= 320 the branch lengths of a tree reach the small-tree kernel by a TMA bulk copy of the 16-byte
    pieces that cover its matrix rows.  Matrices that start 8 bytes off a 16-byte boundary (an odd k, a view into a
    bigger buffer) shift every second tree's block: same distances, bit for bit, as from an aligned copy, including
    the first and last tree, whose pieces would reach outside the array."""
    B, k = 7, n_leaves - 1
    V = sample_vectors(B, n_leaves, seed=3 * n_leaves)
    M = np.stack([to_matrix_input(V[b], sample_bls(1, n_leaves, b)[0]) for b in range(B)])
    want = p2v.cophenetic_distances_batch(torch.from_numpy(M).cuda())
    for lead in (1, 2, 3):                                  # 8, 16, 24 bytes into a buffer
        big = torch.zeros(B * k * 3 + 8, dtype=torch.float64, device="cuda")
        view = big[lead:lead + B * k * 3].view(B, k, 3)
        view.copy_(torch.from_numpy(M))
        assert view.data_ptr() % 16 == (8 * lead) % 16
        assert torch.equal(p2v.cophenetic_distances_batch(view), want), (n_leaves, lead)
        assert torch.equal(p2v.vcv_batch(view), p2v.vcv_batch(torch.from_numpy(M).cuda())), (n_leaves, lead)
    ref = O.cophenetic_distances(M[B - 1])
    got = want[B - 1].cpu().numpy()
    assert np.max(np.abs(got - ref) / np.maximum(ref, 1e-300)) <= 1e-12


@pytest.mark.parametrize("n_leaves", [130, 200, 320])
def test_invalid_trees_between_staged_ones(p2v, n_leaves):
    """A batch for the small-tree kernel's staged (TMA) path with invalid vectors at the start, in the middle, twice in a
    row and at the end: the skipped trees keep the copy / barrier sequence going, their neighbours get the distances
    of a clean batch, and the statuses name the bad entries."""
    B, k = 11, n_leaves - 1
    V = sample_vectors(B, n_leaves, seed=5 * n_leaves)
    M = np.stack([to_matrix_input(V[b], sample_bls(1, n_leaves, b)[0]) for b in range(B)])
    clean = p2v.cophenetic_distances_batch(torch.from_numpy(M).cuda()).cpu().numpy()
    bad = M.copy()
    for b in (0, 4, 5, 10):
        bad[b, 7, 0] = 2 * 7 + 1                      # v[7] out of range
    st = torch.zeros(B, dtype=torch.int32, device="cuda")
    got = p2v.cophenetic_distances_batch(torch.from_numpy(bad).cuda(), status=st, check=False).cpu().numpy()
    torch.cuda.synchronize()
    want_st = np.zeros(B, dtype=np.int32)
    want_st[[0, 4, 5, 10]] = 8
    assert np.array_equal(st.cpu().numpy(), want_st)
    for b in range(B):
        if want_st[b] == 0:
            assert np.array_equal(got[b], clean[b]), b


def test_status_tensor_is_validated(p2v):
    V = torch.from_numpy(sample_vectors(4, 30, seed=2)).cuda()
    for bad in (torch.zeros(4, dtype=torch.int64, device="cuda"), torch.zeros(3, dtype=torch.int32, device="cuda"),
                torch.zeros(4, dtype=torch.int32)):
        with pytest.raises(ValueError, match="status"):
            p2v.to_ancestry_batch(V, status=bad)
        with pytest.raises(ValueError, match="status"):
            p2v.cophenetic_distances_batch(V, status=bad)


def test_host_cache_limit_and_release(p2v):
    """The *_host staging is freed when a call leaves more than the limit behind."""
    from phylo2vec_b200 import _capi
    lib = _capi.load()
    V = sample_vectors(20000, 1000, seed=3)               # 160 MB in, 480 MB out
    try:
        _capi.set_host_cache_limit(64 << 20)
        a = p2v.to_ancestry_batch(V)
        assert _capi.host_cache_bytes() == 0
        _capi.set_host_cache_limit(-1)
        b = p2v.to_ancestry_batch(V)
        assert _capi.host_cache_bytes() > (64 << 20)
        assert np.array_equal(a, b)
        _capi.release_host_cache()
        assert _capi.host_cache_bytes() == 0
    finally:
        lib.p2v_host_cache_limit(512 << 20)
    ref, _ = O.batch("to_ancestry", V[:64], 999)
    assert np.array_equal(a[:64], ref)


def test_two_host_threads_share_the_abi(p2v):
    """Row b4 of the scope table: re-entrant C ABI, no global mutable state.  Two host threads call
    the *_host and *_batch entry points at the same time (on two devices when there are two), each
    on its own stream; every result must equal the oracle's."""
    from phylo2vec_b200 import _capi
    lib = _capi.load()
    ndev = torch.cuda.device_count()
    n, B = 700, 3000
    k = n - 1
    inputs = [sample_vectors(B, n, seed=40 + t) for t in range(2)]
    results, errors = [None, None], []

    def work(t):
        try:
            d = t % ndev
            torch.cuda.set_device(d)
            V = inputs[t]
            out_host = np.empty((B, k, 3), dtype=np.int64)
            st = np.zeros(B, dtype=np.int32)
            stream = torch.cuda.Stream(device=d)
            for _ in range(4):
                rc = lib.p2v_to_ancestry_host(V.ctypes.data_as(ctypes.c_void_p), 8, B, k, out_host.ctypes.data_as(ctypes.c_void_p),
                                              st.ctypes.data_as(ctypes.c_void_p))
                assert rc == 0, lib.p2v_last_error()
                with torch.cuda.stream(stream):
                    Vd = torch.from_numpy(V).cuda(d, non_blocking=True)
                    anc = p2v.to_ancestry_batch(Vd)
                    back = p2v.from_ancestry_batch(anc)
                    D = p2v.cophenetic_distances_batch(Vd[:8])
                stream.synchronize()
                assert torch.equal(back, Vd)
                assert np.array_equal(anc.cpu().numpy(), out_host)
            # an error on this thread stays on this thread
            assert lib.p2v_to_ancestry_batch(None, 3, 1, 4, None, None, None, 0, None) == -1
            assert b"idx_bytes" in lib.p2v_last_error()
            results[t] = (out_host, D.cpu().numpy())
        except BaseException as exc:      # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for t in range(2):
        ref, _ = O.batch("to_ancestry", inputs[t], k)
        assert np.array_equal(results[t][0], ref)
        for b in range(8):
            assert np.array_equal(results[t][1][b], O.cophenetic_distances(inputs[t][b]))
