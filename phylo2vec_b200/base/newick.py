"""Newick output; drop-in for py-phylo2vec/phylo2vec/base/newick.py:29-43 (to_newick, vectors).

from_newick and the matrix form of to_newick (branch lengths) are not built."""

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
