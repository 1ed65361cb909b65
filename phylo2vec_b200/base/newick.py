"""Newick conversion; drop-in for py-phylo2vec/phylo2vec/base/newick.py (to_newick and from_newick,
vectors).

The matrix forms (branch lengths) are not built."""

import ctypes

import numpy as np

from .. import _capi
from .._host import as_index_array, ptr, raise_status


def to_newick(vector_or_matrix) -> str:
    """Newick string of a Phylo2Vec vector, like core.to_newick (py-phylo2vec/src/lib.rs:90-96;
    phylo2vec/src/vector/convert.rs:341-397): every node labelled, e.g. ``((0,1)5,(2,3)4)6;``."""
    t = vector_or_matrix if isinstance(vector_or_matrix, np.ndarray) else np.asarray(vector_or_matrix)
    if t.ndim == 2:
        raise NotImplementedError("to_newick of a matrix (branch lengths) is not built in phylo2vec_b200")
    if t.ndim != 1:
        raise TypeError("tree must be a 1-D integer vector")
    v = as_index_array(t, 1, "to_newick")
    k = v.shape[0]
    lib = _capi.load()
    stride = int(lib.p2v_newick_max_bytes(k))
    buf = ctypes.create_string_buffer(stride)
    length = np.zeros(1, dtype=np.int64)
    status = np.zeros(1, dtype=np.int32)
    rc = lib.p2v_to_newick_host(ptr(v), 8, 1, k, ctypes.cast(buf, ctypes.c_void_p), stride, ptr(length), ptr(status))
    _capi.check(rc, "to_newick")
    raise_status(status, v)
    return buf.raw[:int(length[0])].decode("ascii")


def from_newick(newick: str) -> np.ndarray:
    """Phylo2Vec vector of a Newick string with or without parent labels, like core.to_vector
    (py-phylo2vec/src/lib.rs:98-101; phylo2vec/src/vector/convert.rs:268-278):
    ``from_newick("((0,1)5,(2,3)4)6;") -> [0, 2, 2]``, ``from_newick("((0,1),(2,3));") -> [0, 2, 2]``.
    Branch lengths (the matrix form) raise NotImplementedError."""
    if not isinstance(newick, str):
        raise TypeError("newick must be a str")
    if ":" in newick:
        raise NotImplementedError("from_newick with branch lengths (matrix form) is not built in phylo2vec_b200")
    data = newick.encode("ascii")
    n_leaves = data.count(b",") + 1
    k = n_leaves - 1
    v = np.zeros(k, dtype=np.int64)
    if k == 0:
        return v
    buf = ctypes.create_string_buffer(data, len(data) + 1)
    status = np.zeros(1, dtype=np.int32)
    rc = _capi.load().p2v_from_newick_host(ctypes.cast(buf, ctypes.c_void_p), len(data) + 1, 1, k, ptr(v), 8, ptr(status))
    _capi.check(rc, "from_newick")
    if int(status[0]) != 0:
        raise ValueError("malformed Newick string: not a binary tree with leaves 0..n-1 and parents n..2n-2")
    return v
