"""GPU parity tests for to_newick (run with -m gpu on a B200): byte-exact against the restatement
of phylo2vec/src/vector/convert.rs:341-397, which tests/test_oracle_goldens.py pins to the
reference's own Newick goldens."""

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
    with pytest.raises(NotImplementedError):
        to_newick(np.zeros((3, 3)))


SIZES = [2, 3, 4, 5, 9, 10, 11, 33, 64, 65, 99, 100, 101, 129, 500, 999, 1000, 1001, 1023, 1024]


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
    with pytest.raises(RuntimeError, match="1023"):                 # larger trees: not built yet
        p2v.to_newick_batch(dev(sample_vectors(2, 1026, seed=1)))
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
