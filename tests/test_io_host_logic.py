"""CPU checks of the file I/O layer's host logic (no GPU): extension validation and the text formats of the writers,
which must stay byte for byte those of py-phylo2vec/phylo2vec/io/writer.py:13-50 (numpy's "%d" / "%.18e")."""

from io import StringIO

import numpy as np
import pytest

from phylo2vec_b200.io import _validation, writer


def test_extensions_like_the_reference():
    for ok in ("a.csv", "b.txt", "/tmp/x/y.csv"):
        _validation.check_path(ok, "array")
    for ok in ("t.nwk", "t.newick", "t.tree", "t.treefile", "t.txt"):
        _validation.check_path(ok, "newick")
    with pytest.raises(ValueError, match=r"Unsupported file extension: \.json\. Accepted extensions for array files"):
        _validation.check_path("v.json", "array")
    with pytest.raises(ValueError, match=r"Accepted extensions for newick files"):
        _validation.check_path("v.csv", "newick")
    with pytest.raises(ValueError, match=r"Unsupported file extension: \."):
        _validation.check_path("noext", "array")


def test_writer_text_formats(tmp_path):
    v = np.array([0, 2, 2, 5, 2], dtype=np.int64)
    m = np.array([[0.0, 0.25, 1e-7], [1.0, 3.0, 0.1 + 0.2]])
    # strings (io/writer.py:13-24)
    assert writer.write(v) == np.array2string(v, separator=",")[1:-1] == "0,2,2,5,2"
    buf = StringIO()
    np.savetxt(buf, m, delimiter=",")
    assert writer.write(m) == buf.getvalue()
    with pytest.raises(ValueError):
        writer.write(np.zeros((2, 2, 2)))
    # files (io/writer.py:27-50): "%d" for vectors, "%.18e" for matrices
    pv, pm = tmp_path / "v.csv", tmp_path / "m.csv"
    writer.save(v, str(pv))
    writer.save(m, str(pm))
    assert pv.read_text() == "0\n2\n2\n5\n2\n"
    assert pm.read_text().splitlines()[1] == ",".join("%.18e" % x for x in m[1])
    assert np.array_equal(np.loadtxt(str(pm), delimiter=","), m)            # %.18e reads back bit for bit
    with pytest.raises(ValueError, match="Unsupported file extension"):
        writer.save(v, str(tmp_path / "v.bin"))
    # batches: one tree per line, the same number formats
    V = np.array([[0, 1, 2], [0, 0, 4]], dtype=np.int64)
    pb = tmp_path / "batch.csv"
    writer.save_batch(V, str(pb))
    assert pb.read_text() == "0,1,2\n0,0,4\n"
    M = np.stack([m, 2 * m])
    writer.save_batch(M, str(pb))
    back = np.loadtxt(str(pb), delimiter=",").reshape(2, 2, 3)
    assert np.array_equal(back, M)
    with pytest.raises(ValueError):
        writer.save_batch(np.zeros((2, 2), dtype=np.float64), str(pb))
