"""GPU parity tests for the batched Robinson-Foulds distance (run with -m gpu on a B200): bit-exact
against the literal port of phylo2vec/src/vector/distance.rs:41-162 (integers stored as float64; the
normalised form is one correctly rounded division of two small integers, also exact)."""

import ctypes

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import p2v_oracle as O  # noqa: E402
from tests.helpers import adversarial_vectors, sample_bls, sample_vectors, to_matrix_input  # noqa: E402
from tests.test_oracle_goldens import check_rf_golden  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def p2v():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import phylo2vec_b200 as m
    return m


def test_reference_table(p2v, goldens):
    from phylo2vec_b200.stats import robinson_foulds
    for g in goldens["robinson_foulds"]:
        check_rf_golden(lambda a, b, nz: robinson_foulds(np.array(a), np.array(b), normalize=nz), g)
    with pytest.raises(AssertionError, match="same number of leaves"):
        robinson_foulds(np.array([0, 1, 2]), np.array([0, 1]))
    with pytest.raises(AssertionError, match="out of bounds"):
        robinson_foulds(np.array([0, 1, 2, 9]), np.array([0, 1, 2, 3]))
    # matrices are reduced to their topology (py-phylo2vec/src/lib.rs:295-306), in any combination
    v1, v2 = sample_vectors(2, 40, seed=9)
    m1, m2 = to_matrix_input(v1, sample_bls(1, 40, 1)[0]), to_matrix_input(v2, sample_bls(1, 40, 2)[0])
    want = O.robinson_foulds(v1, v2)
    assert robinson_foulds(m1, m2) == want and robinson_foulds(m1, v2) == want and robinson_foulds(v1, m2) == want
    assert isinstance(robinson_foulds(v1, v2), float)


@pytest.mark.parametrize("n_leaves", [2, 3, 4, 5, 6, 8, 31, 32, 33, 64, 65, 100, 257, 1000, 2000, 4096, 4097, 6000])
def test_pairs_against_oracle(p2v, n_leaves):
    k = n_leaves - 1
    B = 24 if n_leaves <= 1000 else 8
    V1 = sample_vectors(B, n_leaves, seed=n_leaves)
    V2 = sample_vectors(B, n_leaves, seed=n_leaves + 1)
    adv = list(adversarial_vectors(n_leaves).values())
    V1[1:1 + len(adv)] = np.stack(adv)                   # ladders and combs against random trees ...
    V2[3] = V1[3]                                        # ... an identical pair ...
    V2[2] = adv[0]                                       # ... and two adversarial shapes against each other
    V2[B - 1] = sample_vectors(1, n_leaves, seed=5, ordered=True)[0]
    for dtype in (torch.int64, torch.int32):
        d1, d2 = torch.from_numpy(V1).cuda().to(dtype), torch.from_numpy(V2).cuda().to(dtype)
        for nz in (False, True):
            got = p2v.robinson_foulds_batch(d1, d2, normalize=nz).cpu().numpy()
            want = np.array([O.robinson_foulds(V1[b], V2[b], nz) for b in range(B)])
            assert np.array_equal(got, want), (n_leaves, dtype, nz)
            back = p2v.robinson_foulds_batch(d2, d1, normalize=nz).cpu().numpy()
            assert np.array_equal(got, back)             # symmetric, bit for bit
        assert float(p2v.robinson_foulds_batch(d1, d1)[0:].abs().max()) == 0.0
    # one tree against many, both ways round
    one = torch.from_numpy(V1[:1]).cuda()
    many = torch.from_numpy(V2).cuda()
    want = np.array([O.robinson_foulds(V1[0], V2[b]) for b in range(B)])
    assert np.array_equal(p2v.robinson_foulds_batch(one, many).cpu().numpy(), want)
    assert np.array_equal(p2v.robinson_foulds_batch(many, one).cpu().numpy(), want)
    if n_leaves >= 4:                                    # a binary tree has n - 3 splits: the bound is reached or not
        assert want.max() <= 2 * (n_leaves - 3)


def test_invalid_vectors_and_shapes(p2v):
    from phylo2vec_b200 import _capi
    n, B = 50, 6
    V1 = sample_vectors(B, n, seed=1)
    V2 = sample_vectors(B, n, seed=2)
    V2[4, 7] = 99
    st = torch.zeros(B, dtype=torch.int32, device="cuda")
    got = p2v.robinson_foulds_batch(torch.from_numpy(V1).cuda(), torch.from_numpy(V2).cuda(), status=st, check=False)
    assert st.cpu().numpy().tolist() == [0, 0, 0, 0, 8, 0]
    assert float(got[0]) == O.robinson_foulds(V1[0], V2[0])
    with pytest.raises(AssertionError, match=r"v\[7\] = 99 is out of bounds"):
        p2v.robinson_foulds_batch(torch.from_numpy(V1).cuda(), torch.from_numpy(V2).cuda())
    with pytest.raises(AssertionError, match="same number of leaves"):
        p2v.robinson_foulds_batch(torch.from_numpy(V1).cuda(), torch.from_numpy(V2[:, :-1].copy()).cuda())
    with pytest.raises(ValueError):
        p2v.robinson_foulds_batch(torch.from_numpy(V1).cuda(), torch.from_numpy(V2[:4]).cuda())
    # the host entry point, through ctypes with host buffers
    lib = _capi.load()
    V2[4, 7] = 1
    out = np.zeros(B)
    sth = np.zeros(B, dtype=np.int32)
    cp = lambda a: a.ctypes.data_as(ctypes.c_void_p)     # noqa: E731
    assert lib.p2v_robinson_foulds_host(cp(V1), cp(V2), 8, B, B, n - 1, 0, cp(out), cp(sth)) == 0
    assert np.array_equal(out, [O.robinson_foulds(V1[b], V2[b]) for b in range(B)]) and not sth.any()
    assert lib.p2v_robinson_foulds_host(cp(V1), cp(V2), 8, 3, 2, n - 1, 0, cp(out), cp(sth)) == -1


def test_all_pairs_of_a_batch_through_one_vs_many(p2v):
    """The use the scope table names: every tree of a decoded batch against every other."""
    n, B = 300, 40
    V = sample_vectors(B, n, seed=77)
    Vd = torch.from_numpy(V).cuda()
    D = torch.stack([p2v.robinson_foulds_batch(Vd[b:b + 1], Vd) for b in range(B)]).cpu().numpy()
    assert np.array_equal(D, D.T) and not D.diagonal().any()
    for (a, b) in ((0, 1), (5, 17), (39, 2)):
        assert D[a, b] == O.robinson_foulds(V[a], V[b])
