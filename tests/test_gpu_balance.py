"""GPU parity tests for the balance indices (run with -m gpu on a B200) against the restatement of
phylo2vec/src/vector/balance.rs:14-58, which tests/test_oracle_goldens.py pins to the reference's own
test tables.  Sackin and the two sums behind the variance are integers: bit-exact; B2 is a sum of
positive terms added in another order: <= 1e-12 relative."""

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


def test_goldens_single_tree(p2v, goldens):
    from phylo2vec_b200.stats import b2, leaf_depth_variance, sackin
    for g in goldens["balance"]:
        v = np.array(g["v"], dtype=np.int64)
        if "sackin" in g:
            s = sackin(v)
            assert isinstance(s, int) and s == g["sackin"], g["src"]
        if "variance" in g:
            assert abs(leaf_depth_variance(v) - g["variance"]) <= 1e-8 + 1e-5 * abs(g["variance"]), g["src"]
        assert abs(b2(v) - g["b2"]) <= 1e-8 + 1e-5 * abs(g["b2"]), g["src"]
    # a matrix: only the topology column counts (stats/balance.py:30-31)
    m = np.array([[0, 0.1, 0.2], [2, 0.3, 0.4], [2, 0.5, 0.6]])
    assert sackin(m) == 8 and b2(m) == 2.0 and leaf_depth_variance(m) == 0.0
    assert sackin(np.zeros(0, dtype=np.int64)) == 0
    with pytest.raises(ValueError):
        sackin(np.zeros((2, 2, 2)))
    with pytest.raises(AssertionError, match="out of bounds"):
        sackin(np.array([0, 5, 1]))


@pytest.mark.parametrize("n_leaves", [2, 3, 4, 17, 100, 101, 1000, 2000, 5000, 20000])
@pytest.mark.parametrize("dtype", [torch.int64, torch.int32])
def test_batch_parity(p2v, n_leaves, dtype):
    B = 24 if n_leaves <= 5000 else 6
    V = sample_vectors(B, n_leaves, seed=19 * n_leaves)
    adv = list(adversarial_vectors(n_leaves).values())[:B - 1]
    V[1:1 + len(adv)] = np.stack(adv)
    ref = np.stack([O.balance(v) for v in V])
    got = p2v.balance_indices_batch(torch.from_numpy(V).cuda().to(dtype))
    assert got.dtype == torch.float64 and got.shape == (B, 3)
    got = got.cpu().numpy()
    assert np.array_equal(got[:, 0], ref[:, 0])                       # Sackin: exact
    assert np.array_equal(got[:, 1], ref[:, 1])                       # the same two integer sums, the same formula
    assert np.all(np.abs(got[:, 2] - ref[:, 2]) <= 1e-12 * np.abs(ref[:, 2]))


def test_invalid_vector_in_batch(p2v):
    V = sample_vectors(5, 40, seed=2)
    V[3, 9] = 19
    status = torch.zeros(5, dtype=torch.int32, device="cuda")
    out = p2v.balance_indices_batch(torch.from_numpy(V).cuda(), status=status, check=False)
    assert list(status.cpu().numpy()) == [0, 0, 0, 10, 0] and float(out[3].abs().sum()) == 0.0
    with pytest.raises(AssertionError, match=r"v\[9\] = 19"):
        p2v.balance_indices_batch(torch.from_numpy(V).cuda())
