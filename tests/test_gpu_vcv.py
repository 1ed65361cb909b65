"""GPU parity tests for the variance-covariance matrix and node depths (run with -m gpu on a B200).

Topological results must be bit-exact; fp64 branch-length results within 1e-12 relative error of
the literal restatements of phylo2vec/src/vector/graph.rs:211-258 (_vcv) and
phylo2vec/src/vector/ops.rs:201-233 (_get_node_depths)."""

import ctypes

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import p2v_oracle as O  # noqa: E402
from tests.helpers import adversarial_vectors, sample_bls, sample_vectors, to_matrix_input  # noqa: E402

pytestmark = pytest.mark.gpu
RTOL = 1e-12


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


def test_goldens_single_tree(p2v, goldens):
    from phylo2vec_b200.stats import cov
    from phylo2vec_b200.utils.vector import get_node_depth, get_node_depths
    for g in goldens["vcv_vector"]:
        got = cov(np.array(g["v"], dtype=np.int64))
        exp = np.array(g["c"], dtype=np.float64)
        assert got.dtype == np.float64 and got.shape == exp.shape
        assert np.array_equal(got, exp), g["src"]
    for g in goldens["vcv_matrix"]:
        m = np.array(g["m"], dtype=np.float64)
        got = cov(m)
        assert np.allclose(got, np.array(g["c"]), rtol=1e-5, atol=1e-8), g["src"]   # the reference's tolerance
        assert rel_err(got, O.vcv(m)) <= RTOL, g["src"]
    for g in goldens["node_depths"]:
        t = np.array(g["tree"])
        t = t.astype(np.float64) if t.ndim == 2 else t.astype(np.int64)
        got = get_node_depths(t)
        assert got.shape == (2 * t.shape[0] + 1,) and got[-1] == 0.0
        for node, d in g["depths"].items():
            assert abs(got[int(node)] - d) < 1e-10, (g["src"], node)
            assert abs(get_node_depth(t, int(node)) - d) < 1e-10
    with pytest.raises(AssertionError, match="not supported for trees with < 2 leaves"):
        cov(np.zeros(0, dtype=np.int64))                    # vector/graph.rs:636-641
    with pytest.raises(ValueError, match="Node index out of bounds. Max node = 6, got node = 7"):
        get_node_depth(np.array([0, 1, 2]), 7)             # vector/ops.rs:663-668
    with pytest.raises(ValueError, match="out of bounds"):
        get_node_depth(np.array([0]), 3)
    with pytest.raises(TypeError):
        cov(np.zeros((4, 3), dtype=np.int64))
    with pytest.raises(AssertionError, match="out of bounds"):
        cov(np.array([0, 5, 1]))
    assert np.array_equal(get_node_depths(np.zeros(0, dtype=np.int64)), [0.0])   # ops.rs:206-209


SIZES = [2, 3, 4, 5, 8, 9, 33, 63, 64, 65, 66, 100, 129, 500, 513, 1000, 1001, 1026, 1100]


@pytest.mark.parametrize("n_leaves", SIZES)
def test_batch_parity(p2v, n_leaves):
    k = n_leaves - 1
    B = 6
    V = sample_vectors(B, n_leaves, seed=n_leaves)
    adv = list(adversarial_vectors(n_leaves).values())[:4]
    V[:len(adv)] = np.stack(adv)
    bls = sample_bls(B, n_leaves, seed=n_leaves)
    M = to_matrix_input(V, bls)
    ref_t = np.stack([O.vcv(V[b]) for b in range(B)])
    ref_b = np.stack([O.vcv(M[b]) for b in range(B)])
    for dt in (torch.int64, torch.int32):
        got = p2v.vcv_batch(dev(V, dt))
        assert got.dtype == torch.float64 and tuple(got.shape) == (B, n_leaves, n_leaves)
        assert np.array_equal(got.cpu().numpy(), ref_t), (n_leaves, dt)
    got = p2v.vcv_batch(dev(V), bls=dev(bls)).cpu().numpy()
    assert rel_err(got, ref_b) <= RTOL
    got_m = p2v.vcv_batch(dev(M)).cpu().numpy()
    assert np.array_equal(got_m, got)
    assert np.array_equal(got, np.swapaxes(got, 1, 2))      # min of the same two candidates: bit-symmetric
    # node depths
    dep_t = np.stack([O.get_node_depths(V[b]) for b in range(B)])
    dep_b = np.stack([O.get_node_depths(M[b]) for b in range(B)])
    for dt in (torch.int64, torch.int32):
        assert np.array_equal(p2v.node_depths_batch(dev(V, dt)).cpu().numpy(), dep_t)
    gd = p2v.node_depths_batch(dev(M)).cpu().numpy()
    assert gd.shape == (B, 2 * k + 1) and rel_err(gd, dep_b) <= RTOL
    assert np.array_equal(p2v.node_depths_batch(dev(V), bls=dev(bls)).cpu().numpy(), gd)
    # diagonal of the vcv = depth of the leaves; D = d_i + d_j - 2 C (graph.rs:20-90 vs :211-258)
    assert np.array_equal(np.diagonal(got, axis1=1, axis2=2), gd[:, :n_leaves])
    D = p2v.cophenetic_distances_batch(dev(M)).cpu().numpy()
    d = gd[:, :n_leaves]
    assert np.allclose(D, d[:, :, None] + d[:, None, :] - 2 * got, rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("n_leaves", [70, 1000, 3000])
def test_row_blocks(p2v, n_leaves):
    V = sample_vectors(3, n_leaves, seed=5)
    bls = sample_bls(3, n_leaves, seed=5)
    full_t = p2v.vcv_batch(dev(V))
    full_b = p2v.vcv_batch(dev(V), bls=dev(bls))
    assert rel_err(full_b[0].cpu().numpy(), O.vcv(to_matrix_input(V, bls)[0])) <= RTOL
    for r0, r1 in ((0, 1), (n_leaves // 3, n_leaves // 3 + 17), (n_leaves - 5, n_leaves), (7, 7),
                   (0, n_leaves // 8), (n_leaves // 2, n_leaves)):
        part = p2v.vcv_batch(dev(V), rows=(r0, r1))
        assert tuple(part.shape) == (3, r1 - r0, n_leaves)
        assert torch.equal(part, full_t[:, r0:r1])
        part = p2v.vcv_batch(dev(V), rows=(r0, r1), bls=dev(bls))
        assert torch.equal(part, full_b[:, r0:r1])          # a selection, not a sum: bit-identical rows


def test_invalid_tree_in_batch(p2v):
    V = sample_vectors(4, 60, seed=1)
    V[2, 10] = 99
    status = torch.zeros(4, dtype=torch.int32, device="cuda")
    out = p2v.vcv_batch(dev(V), status=status, check=False).cpu().numpy()
    assert list(status.cpu().numpy()) == [0, 0, 11, 0]
    for b in (0, 1, 3):
        assert np.array_equal(out[b], O.vcv(V[b]))
    with pytest.raises(AssertionError, match=r"v\[10\] = 99"):
        p2v.vcv_batch(dev(V))
    with pytest.raises(AssertionError, match=r"v\[10\] = 99"):
        p2v.node_depths_batch(dev(V))
    with pytest.raises(AssertionError, match="< 2 leaves"):
        p2v.vcv_batch(torch.zeros((3, 0), dtype=torch.int64, device="cuda"))


@pytest.mark.parametrize("n_leaves", [2000, 16384, 65536])
def test_large_trees(p2v, n_leaves):
    """Sizes of BASELINE.json configs[2] and [3]: sampled rows against the oracle."""
    V = sample_vectors(1, n_leaves, seed=n_leaves)
    bls = sample_bls(1, n_leaves, seed=n_leaves)
    M = to_matrix_input(V, bls)
    for r0, r1 in [(0, 8), (n_leaves // 2 - 3, n_leaves // 2 + 5), (n_leaves - 8, n_leaves)]:
        got = p2v.vcv_batch(dev(V), rows=(r0, r1))[0].cpu().numpy()
        assert np.array_equal(got, O.vcv_rows(V[0], r0, r1))
        got = p2v.vcv_batch(dev(M), rows=(r0, r1))[0].cpu().numpy()
        assert rel_err(got, O.vcv_rows(M[0], r0, r1)) <= RTOL
    blk = n_leaves // 8
    a = p2v.vcv_batch(dev(M), rows=(0, blk))[0]
    b = p2v.vcv_batch(dev(M), rows=(blk, 2 * blk))[0]
    assert torch.equal(a[:, blk:2 * blk], b[:, :blk].T) and torch.equal(a[:, :blk], a[:, :blk].T)
    dep = p2v.node_depths_batch(dev(M))[0]
    assert rel_err(dep.cpu().numpy(), O.get_node_depths(M[0])) <= RTOL
    assert torch.equal(torch.diagonal(a[:, :blk]), dep[:blk])
    assert bool((a <= torch.minimum(dep[:blk, None], dep[None, :n_leaves])).all()) and bool((a >= 0).all())


def test_host_entry_points(p2v):
    from phylo2vec_b200 import _capi
    lib = _capi.load()
    n = 300
    B = 5
    V = sample_vectors(B, n, seed=8)
    bls = sample_bls(B, n, seed=8)
    M = np.ascontiguousarray(to_matrix_input(V, bls))
    vp = ctypes.c_void_p
    out = np.zeros((B, n, n))
    st = np.ones(B, dtype=np.int32)
    rc = lib.p2v_vcv_matrix_host(M.ctypes.data_as(vp), B, n - 1, 0, n, out.ctypes.data_as(vp), st.ctypes.data_as(vp))
    assert rc == 0 and not st.any()
    assert rel_err(out, np.stack([O.vcv(M[b]) for b in range(B)])) <= RTOL
    part = np.zeros((B, 10, n))
    rc = lib.p2v_vcv_host(V.ctypes.data_as(vp), 8, bls.ctypes.data_as(vp), B, n - 1, 20, 30,
                          part.ctypes.data_as(vp), None)
    assert rc == 0 and np.array_equal(part, out[:, 20:30])
    dep = np.zeros((B, 2 * n - 1))
    rc = lib.p2v_node_depths_host(V.ctypes.data_as(vp), 8, None, B, n - 1, dep.ctypes.data_as(vp), None)
    assert rc == 0 and np.array_equal(dep, np.stack([O.get_node_depths(V[b]) for b in range(B)]))
    rc = lib.p2v_vcv_host(V.ctypes.data_as(vp), 8, None, B, 0, 0, 1, out.ctypes.data_as(vp), None)
    assert rc != 0 and b"< 2 leaves" in lib.p2v_last_error()
