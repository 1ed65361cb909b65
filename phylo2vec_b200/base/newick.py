"""Newick conversion; drop-in for py-phylo2vec/phylo2vec/base/newick.py: to_newick and from_newick for
vectors and for matrices (branch lengths), plus the two core names behind from_newick, to_vector and to_matrix
(py-phylo2vec/src/lib.rs:98-106)."""

import ctypes

import numpy as np

from .. import _capi
from .._host import as_index_array, ptr, raise_status


def to_newick(vector_or_matrix) -> str:
    """Newick string of a Phylo2Vec vector or matrix, like core.to_newick (py-phylo2vec/src/lib.rs:90-96;
    phylo2vec/src/vector/convert.rs:341-397, matrix/convert.rs:26-57): every node labelled, e.g.
    ``((0,1)5,(2,3)4)6;`` -- with a matrix every node but the root is followed by ``:length``, e.g.
    ``(0:0.7,(1:0.5,2:0.8)3:0.6)4;``."""
    t = vector_or_matrix if isinstance(vector_or_matrix, np.ndarray) else np.asarray(vector_or_matrix)
    lib = _capi.load()
    length = np.zeros(1, dtype=np.int64)
    status = np.zeros(1, dtype=np.int32)
    if t.ndim == 2:
        if t.dtype != np.float64:
            raise TypeError("a matrix must be float64")
        from ..utils.matrix import check_matrix
        check_matrix(t)                      # the reference validates first (matrix/convert.rs:28)
        m = np.ascontiguousarray(t)
        k = m.shape[0]
        stride = int(lib.p2v_newick_matrix_max_bytes(k))
        buf = ctypes.create_string_buffer(stride)
        rc = lib.p2v_to_newick_matrix_host(ptr(m), 1, k, ctypes.cast(buf, ctypes.c_void_p), stride, ptr(length), ptr(status))
        _capi.check(rc, "to_newick")
        raise_status(status, m[:, 0].astype(np.int64))
        return buf.raw[:int(length[0])].decode("ascii")
    if t.ndim != 1:
        raise TypeError("tree must be a 1-D integer vector or a 2-D float64 matrix")
    v = as_index_array(t, 1, "to_newick")
    k = v.shape[0]
    stride = int(lib.p2v_newick_max_bytes(k))
    buf = ctypes.create_string_buffer(stride)
    rc = lib.p2v_to_newick_host(ptr(v), 8, 1, k, ctypes.cast(buf, ctypes.c_void_p), stride, ptr(length), ptr(status))
    _capi.check(rc, "to_newick")
    raise_status(status, v)
    return buf.raw[:int(length[0])].decode("ascii")


def has_branch_lengths(newick: str) -> bool:
    """newick/mod.rs:322-328 (pattern ``:\\d+(\\.\\d+)?...``: a colon followed by a digit)."""
    import re
    return re.search(r":\d", newick) is not None


def to_vector(newick: str) -> np.ndarray:
    """Phylo2Vec vector of a Newick string with or without parent labels, like core.to_vector
    (py-phylo2vec/src/lib.rs:98-101; phylo2vec/src/vector/convert.rs:268-278):
    ``to_vector("((0,1)5,(2,3)4)6;") -> [0, 2, 2]``, ``to_vector("((0,1),(2,3));") -> [0, 2, 2]``."""
    if not isinstance(newick, str):
        raise TypeError("newick must be a str")
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


def to_matrix(newick: str) -> np.ndarray:
    """Phylo2Vec matrix ``(k,3)`` float64 ``[v, bl1, bl2]`` of a Newick string with branch lengths, with or without
    parent labels, like core.to_matrix (py-phylo2vec/src/lib.rs:103-106; phylo2vec/src/matrix/convert.rs:84-122):
    ``to_matrix("(0:0.7,(1:0.5,2:0.8)3:0.6)4;") -> [[0, 0.5, 0.8], [1, 0.7, 0.6]]``."""
    if not isinstance(newick, str):
        raise TypeError("newick must be a str")
    if newick == "":
        return np.zeros((0, 3), dtype=np.float64)
    data = newick.encode("ascii")
    k = data.count(b",")
    m = np.zeros((k, 3), dtype=np.float64)
    if k == 0:
        return m
    buf = ctypes.create_string_buffer(data, len(data) + 1)
    status = np.zeros(1, dtype=np.int32)
    rc = _capi.load().p2v_from_newick_matrix_host(ctypes.cast(buf, ctypes.c_void_p), len(data) + 1, 1, k, ptr(m), ptr(status))
    _capi.check(rc, "from_newick")
    if int(status[0]) != 0:
        raise ValueError("malformed Newick string: not a binary tree with leaves 0..n-1 and a length after every node")
    return m


def from_newick(newick: str) -> np.ndarray:
    """Vector, or matrix if the string has branch lengths (py-phylo2vec/phylo2vec/base/newick.py:8-26)."""
    if not isinstance(newick, str):
        raise TypeError("newick must be a str")
    return to_matrix(newick) if has_branch_lengths(newick) else to_vector(newick)
