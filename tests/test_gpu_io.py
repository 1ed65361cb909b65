"""GPU tests of the I/O row: files in the reference's format (py-phylo2vec/phylo2vec/io/
writer.py:13-50: numpy.savetxt with "%d" / "%.18e") are parsed on the GPU and come back bit for bit; the cases
follow py-phylo2vec/tests/test_io.py:17-170."""

import io as _io

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import p2v_oracle as O  # noqa: E402
from tests.helpers import sample_bls, sample_vectors, to_matrix_input  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pio():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from phylo2vec_b200 import io
    return io


@pytest.mark.parametrize("n_leaves", [5, 6, 50, 200, 3000])
def test_save_and_load_vector_and_matrix(pio, tmp_path, n_leaves):
    v = sample_vectors(1, n_leaves, seed=n_leaves)[0]
    m = to_matrix_input(v, sample_bls(1, n_leaves, seed=n_leaves)[0])
    for arr in (v, m):
        path = tmp_path / "test.csv"
        pio.save(arr, path)
        # byte for byte what the reference writes (io/writer.py:45-50)
        ref = _io.StringIO()
        np.savetxt(ref, arr, fmt="%d" if arr.ndim == 1 else "%.18e", delimiter=",")
        assert path.read_text() == ref.getvalue()
        got = pio.load(path)
        assert got.dtype == arr.dtype and np.array_equal(got, arr)
        assert np.array_equal(pio.read(path.read_text()), arr)
        bad = tmp_path / "test.random"
        np.savetxt(bad, arr, delimiter=",")
        with pytest.raises(ValueError, match="Unsupported file extension"):
            pio.save(arr, bad)
        with pytest.raises(ValueError, match="Unsupported file extension"):
            pio.load(bad)
    # what numpy writes by default (floats for an integer vector are NOT a vector: reader.py:47-55)
    path = tmp_path / "floats.txt"
    np.savetxt(path, v.astype(float), delimiter=",")
    with pytest.raises(ValueError, match="vector .* or matrix"):
        pio.load(path)
    with pytest.raises(FileNotFoundError):
        pio.load(tmp_path / "missing.csv")


def test_load_validates_like_the_reference(pio, tmp_path):
    path = tmp_path / "bad_vector.csv"
    path.write_text("0\n0\n9\n1\n")
    with pytest.raises(AssertionError, match=r"v\[2\] = 9 is out of bounds"):
        pio.load(path)
    path = tmp_path / "bad_matrix.csv"
    path.write_text("0.0,0.1,0.2\n0.0,-0.3,0.4\n")
    with pytest.raises(AssertionError, match="Branch lengths must be positive"):
        pio.load(path)
    # the edge case the reference documents (reader.py:43-46): a one-row matrix reads as a 1-D float array
    path = tmp_path / "one_row.csv"
    path.write_text("0.0,0.1,0.2\n")
    with pytest.raises(ValueError):
        pio.load(path)
    path = tmp_path / "ragged.csv"
    path.write_text("0.0,0.1,0.2\n0.0,0.3\n")
    with pytest.raises(ValueError):
        pio.load(path)
    path = tmp_path / "words.csv"
    path.write_text("0\nzero\n")
    with pytest.raises(ValueError):
        pio.load(path)
    # other delimiters, CRLF line ends, blanks, no newline at the end
    v = np.array([0, 1, 0, 5, 2])
    assert np.array_equal(pio.read("0\r\n1\r\n0\r\n5\r\n2"), v)
    assert np.array_equal(pio.read(" 0.5; 0.25 ;1e-3\n0;1.5;2\n", delimiter=";"), [[0.5, 0.25, 1e-3], [0.0, 1.5, 2.0]])
    assert pio.write(v) == "0,1,0,5,2"
    with pytest.raises(ValueError):
        pio.save(np.array(0), "test.csv")
    with pytest.raises(ValueError):
        pio.save(np.zeros((5, 3, 1)), "test.csv")


@pytest.mark.parametrize("n_leaves,B", [(2, 7), (100, 5000), (1000, 2000)])
def test_batches_round_trip_through_files(pio, tmp_path, n_leaves, B):
    import phylo2vec_b200 as p2v
    V = sample_vectors(B, n_leaves, seed=3)
    M = to_matrix_input(V, sample_bls(B, n_leaves, seed=4))
    path = tmp_path / "vectors.csv"
    pio.save_batch(torch.from_numpy(V).cuda(), path)
    Vd = pio.load_batch(path)
    assert Vd.is_cuda and Vd.dtype == torch.int64 and np.array_equal(Vd.cpu().numpy(), V)
    # the loaded batch feeds the device API directly
    ref, _ = O.batch("to_ancestry", V[:16], n_leaves - 1)
    assert np.array_equal(p2v.to_ancestry_batch(Vd[:16]).cpu().numpy(), ref)
    path = tmp_path / "matrices.txt"
    pio.save_batch(torch.from_numpy(M).cuda(), path)
    Md = pio.load_batch(path, kind="matrix")
    assert Md.shape == (B, n_leaves - 1, 3) and np.array_equal(Md.cpu().numpy(), M)       # "%.18e" reads back exactly
    # a broken tree in the file is reported with the reference's message
    V2 = V.copy()
    if n_leaves > 2:
        V2[B // 2, 1] = 7
        pio.save_batch(V2, path)
        with pytest.raises(AssertionError, match=r"v\[1\] = 7 is out of bounds"):
            pio.load_batch(path)
        assert np.array_equal(pio.load_batch(path, check=False).cpu().numpy(), V2)


def test_newick_files(pio, tmp_path):
    v = sample_vectors(1, 40, seed=8)[0]
    m = to_matrix_input(v, sample_bls(1, 40, seed=9)[0])
    for arr in (v, m):
        path = tmp_path / "tree.newick"
        pio.save_newick(arr, path)
        assert np.array_equal(pio.load_newick(str(path)), arr)
        assert np.array_equal(pio.load_newick(path.read_text()), arr)
        with pytest.raises(ValueError, match="Unsupported file extension"):
            pio.save_newick(arr, tmp_path / "tree.random")
    pio.save_newick(np.array([0, 2, 2]), tmp_path / "named.nwk", labels={"0": "a", "1": "b", "2": "c", "3": "d"})
    assert (tmp_path / "named.nwk").read_text() == "((a,b)5,(c,d)4)6;"
