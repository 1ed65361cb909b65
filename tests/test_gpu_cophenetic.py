"""GPU parity tests for cophenetic distances (run with -m gpu on a B200).

Topological distances must be bit-exact; fp64 branch-length distances must be
within 1e-12 relative error of the literal restatement of
phylo2vec/src/vector/graph.rs:20-90 (BASELINE.json north_star)."""

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import p2v_oracle as O  # noqa: E402
from tests.helpers import adversarial_vectors, sample_bls, sample_vectors, to_matrix_input  # noqa: E402

pytestmark = pytest.mark.gpu
RTOL = 1e-12  # north_star: <= 1e-12 relative error for fp64 branch-length distances


@pytest.fixture(scope="module")
def p2v():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import phylo2vec_b200 as m
    return m


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def rel_err(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


def assert_symmetric(d, exact):
    """Topological matrices are bit-symmetric.  fp64 matrices are symmetric to the
    last bits only: D[i][j] and D[j][i] may combine the same three double-double
    depths in a different order (DESIGN.md "fp64 parity")."""
    dt = np.swapaxes(d, -1, -2)
    if exact:
        assert np.array_equal(d, dt)
    else:
        assert rel_err(d, dt) <= 1e-15


def test_goldens_single_tree(p2v, goldens):
    for g in goldens["cophenetic_vector"]:
        got = p2v.cophenetic_distances(np.array(g["v"], dtype=np.int64), unrooted=g["unrooted"])
        exp = np.array(g["d"], dtype=np.float64)
        assert got.dtype == np.float64 and got.shape == exp.shape
        assert np.array_equal(got, exp), g["src"]
    for g in goldens["cophenetic_matrix"]:
        got = p2v.cophenetic_distances(np.array(g["m"], dtype=np.float64), unrooted=g["unrooted"])
        exp = np.array(g["d"], dtype=np.float64)
        if "round" in g:
            assert np.array_equal(got.round(g["round"]), exp), g["src"]
        else:
            assert np.allclose(got, exp, rtol=1e-5, atol=1e-8), g["src"]   # the reference's tolerance
        lit = O.cophenetic_distances(np.array(g["m"], dtype=np.float64), g["unrooted"])
        assert rel_err(got, lit) <= RTOL, g["src"]
    v = np.array([0, 2, 2, 5, 4, 1])
    assert np.array_equal(p2v.pairwise_distances(v), p2v.cophenetic_distances(v))


def test_type_dispatch(p2v):
    # PyTree: py-phylo2vec/src/lib.rs:20-24; tests/test_base.py:37-48 style errors
    with pytest.raises(TypeError):
        p2v.cophenetic_distances(np.zeros((5, 3, 1)))
    with pytest.raises(TypeError):
        p2v.cophenetic_distances(np.array(0))
    with pytest.raises(TypeError):
        p2v.cophenetic_distances(np.zeros((4, 3), dtype=np.int64))
    with pytest.raises(AssertionError, match="out of bounds"):
        p2v.cophenetic_distances(np.array([0, 5, 1]))


# 40 / 41: warp-per-tree small kernel below, CTA-per-tree above; 128 / 129 and 160 / 161 / 320 / 321: the small kernel's other shapes
SIZES = [1, 2, 3, 4, 5, 8, 9, 16, 31, 32, 33, 40, 41, 48, 63, 64, 65, 66, 100, 128, 129, 160, 161, 320, 321, 500, 1000, 1001, 1026, 1100]


@pytest.mark.parametrize("n_leaves", SIZES)
@pytest.mark.parametrize("unrooted", [False, True])
def test_batch_parity(p2v, n_leaves, unrooted):
    k = n_leaves - 1
    B = 6
    if k == 0:
        V = np.zeros((B, 0), dtype=np.int64)
    else:
        V = sample_vectors(B, n_leaves, seed=n_leaves)
        adv = list(adversarial_vectors(n_leaves).values())[:4]
        V[:len(adv)] = np.stack(adv)
    bls = sample_bls(B, n_leaves, seed=n_leaves)
    # topological, int64 and int32 vectors
    ref = O.cophenetic_batch(V, None, unrooted)
    for dt in (torch.int64, torch.int32):
        got = p2v.cophenetic_distances_batch(dev(V, dt), unrooted=unrooted)
        assert got.dtype == torch.float64 and tuple(got.shape) == (B, n_leaves, n_leaves)
        assert np.array_equal(got.cpu().numpy(), ref), (n_leaves, unrooted, dt)
    # branch lengths: (v, bls) and the reference's (k,3) matrix form
    ref = O.cophenetic_batch(V, bls, unrooted)
    got = p2v.cophenetic_distances_batch(dev(V), unrooted=unrooted, bls=dev(bls)).cpu().numpy()
    assert rel_err(got, ref) <= RTOL
    M = to_matrix_input(V, bls)
    got_m = p2v.cophenetic_distances_batch(dev(M), unrooted=unrooted).cpu().numpy()
    assert np.array_equal(got_m, got)
    assert_symmetric(got, exact=False)
    assert_symmetric(p2v.cophenetic_distances_batch(dev(V), unrooted=unrooted).cpu().numpy(), exact=True)
    assert not np.diagonal(got, axis1=1, axis2=2).any()


@pytest.mark.parametrize("n_leaves", [70, 1000, 3000])
def test_row_blocks(p2v, n_leaves):
    """Row blocks (how one tree is sharded over GPUs).  Topological rows are bit-identical
    to the full matrix; fp64 rows agree to the last bits only, because the DFS blocks -- and
    with them the order in which the three path pieces are summed -- follow the requested
    rows (DESIGN.md "fp64 parity")."""
    V = sample_vectors(3, n_leaves, seed=5)
    bls = sample_bls(3, n_leaves, seed=5)
    full_t = p2v.cophenetic_distances_batch(dev(V))
    full_b = p2v.cophenetic_distances_batch(dev(V), bls=dev(bls))
    ref_b = O.cophenetic_batch(V[:1], bls[:1])
    for r0, r1 in ((0, 1), (n_leaves // 3, n_leaves // 3 + 17), (n_leaves - 5, n_leaves), (7, 7),
                   (0, n_leaves // 8), (n_leaves // 2, n_leaves)):
        part = p2v.cophenetic_distances_batch(dev(V), rows=(r0, r1))
        assert tuple(part.shape) == (3, r1 - r0, n_leaves)
        assert torch.equal(part, full_t[:, r0:r1])
        part = p2v.cophenetic_distances_batch(dev(V), rows=(r0, r1), bls=dev(bls))
        assert rel_err(part.cpu().numpy(), full_b[:, r0:r1].cpu().numpy()) <= 1e-15
        assert rel_err(part[0].cpu().numpy(), ref_b[0, r0:r1]) <= RTOL
    # shards of a row partition reassemble the matrix (what the multi-GPU path does)
    G = 4
    cuts = [n_leaves * g // G for g in range(G + 1)]
    parts = [p2v.cophenetic_distances_batch(dev(V), rows=(cuts[g], cuts[g + 1])) for g in range(G)]
    assert torch.equal(torch.cat(parts, dim=1), full_t)
    parts = [p2v.cophenetic_distances_batch(dev(V), rows=(cuts[g], cuts[g + 1]), bls=dev(bls)) for g in range(G)]
    assert rel_err(torch.cat(parts, dim=1).cpu().numpy(), full_b.cpu().numpy()) <= 1e-15


def test_invalid_tree_in_batch(p2v):
    V = sample_vectors(4, 60, seed=1)
    V[2, 10] = 99
    status = torch.zeros(4, dtype=torch.int32, device="cuda")
    out = p2v.cophenetic_distances_batch(dev(V), status=status, check=False).cpu().numpy()
    assert list(status.cpu().numpy()) == [0, 0, 11, 0]
    for b in (0, 1, 3):
        assert np.array_equal(out[b], O.cophenetic_distances(V[b]))
    with pytest.raises(AssertionError, match=r"v\[10\] = 99"):
        p2v.cophenetic_distances_batch(dev(V))


def test_config_c3_sample(p2v):
    """BASELINE.json configs[2] shape: n=2000 trees, topological and fp64."""
    n = 2000
    V = sample_vectors(24, n, seed=42)
    bls = sample_bls(24, n, seed=42)
    got = p2v.cophenetic_distances_batch(dev(V)).cpu().numpy()
    ref = O.cophenetic_batch(V[:6], None)
    assert np.array_equal(got[:6], ref)
    for unrooted in (False, True):
        got = p2v.cophenetic_distances_batch(dev(V), unrooted=unrooted, bls=dev(bls)).cpu().numpy()
        ref = O.cophenetic_batch(V[:4], bls[:4], unrooted)
        assert rel_err(got[:4], ref) <= RTOL
        assert_symmetric(got, exact=False)


@pytest.mark.parametrize("n_leaves", [16384, 65536])
def test_single_large_tree(p2v, n_leaves):
    """BASELINE.json configs[3]: one tree, n = 65536, row blocks.  The literal
    algorithm cannot run at this size (137 GB scratch), so sampled
    rows are compared with the closed-form oracle and the whole block with
    size-independent properties."""
    V = sample_vectors(1, n_leaves, seed=n_leaves)
    bls = sample_bls(1, n_leaves, seed=n_leaves)
    M = to_matrix_input(V, bls)
    rows = [(0, 8), (n_leaves // 2 - 3, n_leaves // 2 + 5), (n_leaves - 8, n_leaves)]
    for unrooted in (False, True):
        for r0, r1 in rows:
            got = p2v.cophenetic_distances_batch(dev(V), unrooted=unrooted, rows=(r0, r1))[0].cpu().numpy()
            assert np.array_equal(got, O.cophenetic_rows(V[0], unrooted, r0, r1))
            got = p2v.cophenetic_distances_batch(dev(M), unrooted=unrooted, rows=(r0, r1))[0].cpu().numpy()
            assert rel_err(got, O.cophenetic_rows(M[0], unrooted, r0, r1)) <= RTOL
    # a full row block of the kind one GPU of 8 owns: symmetry against the transposed block
    G = 8
    blk = n_leaves // G
    a = p2v.cophenetic_distances_batch(dev(M), rows=(0, blk))[0]            # rows [0,blk)
    b = p2v.cophenetic_distances_batch(dev(M), rows=(blk, 2 * blk))[0]      # rows [blk,2blk)
    assert rel_err(a[:, blk:2 * blk].cpu().numpy(), b[:, :blk].T.cpu().numpy()) <= 1e-15
    assert_symmetric(a[:, :blk].cpu().numpy(), exact=False)
    assert not bool(torch.diagonal(a[:, :blk]).any())
    assert bool((a >= 0).all())
    t = p2v.cophenetic_distances_batch(dev(V), rows=(0, blk))[0]
    t2 = p2v.cophenetic_distances_batch(dev(V), rows=(blk, 2 * blk))[0]
    assert torch.equal(t[:, blk:2 * blk], t2[:, :blk].T) and torch.equal(t[:, :blk], t[:, :blk].T)
    assert torch.equal(t, t.round()) and float(t.max()) < 2 * n_leaves


def test_host_entry_point_matrix(p2v):
    import ctypes
    from phylo2vec_b200 import _capi
    lib = _capi.load()
    n = 300
    V = sample_vectors(5, n, seed=8)
    bls = sample_bls(5, n, seed=8)
    M = np.ascontiguousarray(to_matrix_input(V, bls))
    out = np.zeros((5, n, n))
    st = np.ones(5, dtype=np.int32)
    rc = lib.p2v_cophenetic_matrix_host(M.ctypes.data_as(ctypes.c_void_p), 5, n - 1, 0, 0, n,
                                        out.ctypes.data_as(ctypes.c_void_p), st.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0 and not st.any()
    assert rel_err(out, O.cophenetic_batch(V, bls)) <= RTOL
    out2 = np.zeros((5, n, n))
    rc = lib.p2v_cophenetic_host(V.ctypes.data_as(ctypes.c_void_p), 8, bls.ctypes.data_as(ctypes.c_void_p), 5,
                                 n - 1, 1, 0, n, out2.ctypes.data_as(ctypes.c_void_p), None)
    assert rc == 0 and rel_err(out2, O.cophenetic_batch(V, bls, True)) <= RTOL


@pytest.mark.parametrize("n_leaves", [3, 5, 33, 100, 257, 512, 513])
def test_two_phase_api_matches_one_shot(p2v, n_leaves):
    """p2v_cophenetic_prepare_batch + p2v_cophenetic_rows_batch (the general kernels at every size)
    against the one-shot call (the fused small-tree kernel up to n = 512)."""
    from phylo2vec_b200 import _capi
    lib = _capi.load()
    B, k = 7, n_leaves - 1
    V = sample_vectors(B, n_leaves, seed=3 * n_leaves)
    bls = sample_bls(B, n_leaves, seed=3 * n_leaves)
    Vd, bd = dev(V), dev(bls)
    need = lib.p2v_cophenetic_workspace_bytes(B, k)
    ws = torch.empty(max(need, 16), dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for unrooted in (0, 1):
        for weighted, blp in ((0, 0), (1, bd.data_ptr())):
            out = torch.empty((B, n_leaves, n_leaves), dtype=torch.float64, device="cuda")
            assert lib.p2v_cophenetic_prepare_batch(Vd.data_ptr(), 8, blp, B, k, unrooted, None, ws.data_ptr(), need, s) == 0
            assert lib.p2v_cophenetic_rows_batch(ws.data_ptr(), need, B, k, weighted, 0, n_leaves, out.data_ptr(), s) == 0
            one = p2v.cophenetic_distances_batch(Vd, unrooted=bool(unrooted), bls=bd if weighted else None)
            ref = O.cophenetic_batch(V, bls if weighted else None, bool(unrooted))
            if weighted:
                assert rel_err(out.cpu().numpy(), ref) <= RTOL and rel_err(one.cpu().numpy(), ref) <= RTOL
            else:
                assert np.array_equal(out.cpu().numpy(), ref) and torch.equal(one, out)
        # variance-covariance through both paths
        out = torch.empty((B, n_leaves, n_leaves), dtype=torch.float64, device="cuda")
        assert lib.p2v_cophenetic_prepare_batch(Vd.data_ptr(), 8, bd.data_ptr(), B, k, 0, None, ws.data_ptr(), need, s) == 0
        assert lib.p2v_vcv_rows_batch(ws.data_ptr(), need, B, k, 1, 0, n_leaves, out.data_ptr(), s) == 0
        assert torch.equal(out, p2v.vcv_batch(Vd, bls=bd))


def test_small_tree_batches(p2v):
    """Batches large enough to fill the GPU with the fused small-tree kernel (BASELINE.json
    configs[0] shape n = 100), row blocks and odd sizes included."""
    for n_leaves, B in ((100, 3000), (16, 5000), (101, 700), (300, 300), (512, 160)):
        V = sample_vectors(B, n_leaves, seed=n_leaves + 1)
        bls = sample_bls(B, n_leaves, seed=n_leaves + 1)
        idx = np.linspace(0, B - 1, 12).astype(int)
        got = p2v.cophenetic_distances_batch(dev(V)).cpu().numpy()
        assert np.array_equal(got[idx], O.cophenetic_batch(V[idx], None))
        assert np.array_equal(got, np.swapaxes(got, 1, 2))
        got = p2v.cophenetic_distances_batch(dev(V), bls=dev(bls), unrooted=True).cpu().numpy()
        assert rel_err(got[idx], O.cophenetic_batch(V[idx], bls[idx], True)) <= RTOL
        r0, r1 = n_leaves // 3, n_leaves // 3 + 9
        part = p2v.cophenetic_distances_batch(dev(V), bls=dev(bls), unrooted=True, rows=(r0, r1)).cpu().numpy()
        # the fused kernel computes every entry on its own; above its size limit the general kernels
        # agree to the last bits only (DESIGN.md "fp64 parity")
        assert rel_err(part, got[:, r0:r1]) <= 1e-15
        if n_leaves <= 256:
            assert np.array_equal(part, got[:, r0:r1])


def test_host_entry_point_large_pageable(p2v):
    """Big pageable output (numpy): the host entry point stages it through pinned memory with threads."""
    import ctypes
    from phylo2vec_b200 import _capi
    lib = _capi.load()
    n, B = 300, 400                       # 288 MB of distances
    V = sample_vectors(B, n, seed=12)
    bls = sample_bls(B, n, seed=12)
    M = np.ascontiguousarray(to_matrix_input(V, bls))
    out = np.zeros((B, n, n))
    st = np.ones(B, dtype=np.int32)
    vp = ctypes.c_void_p
    rc = lib.p2v_cophenetic_matrix_host(M.ctypes.data_as(vp), B, n - 1, 0, 0, n, out.ctypes.data_as(vp), st.ctypes.data_as(vp))
    assert rc == 0 and not st.any()
    idx = np.linspace(0, B - 1, 16).astype(int)
    assert rel_err(out[idx], O.cophenetic_batch(V[idx], bls[idx])) <= RTOL
    dev_out = p2v.cophenetic_distances_batch(dev(M)).cpu().numpy()
    assert np.array_equal(out, dev_out)
