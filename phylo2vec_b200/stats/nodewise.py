"""Node-wise distances and covariance; drop-in for py-phylo2vec/phylo2vec/stats/nodewise.py:10-71."""

import numpy as np

from .. import _capi
from .._host import as_index_array, ptr, raise_status


def cophenetic_distances(tree, unrooted=False) -> np.ndarray:
    """Cophenetic distance matrix of a Phylo2Vec vector (topological) or matrix
    (branch lengths): float64 (n,n), like core.cophenetic_distances
    (py-phylo2vec/src/lib.rs:144-156; phylo2vec/src/vector/graph.rs:20-90;
    phylo2vec/src/matrix/graph.rs:23-26).

    Dispatch follows PyTree (py-phylo2vec/src/lib.rs:20-24): a 2-D float64 array is
    a matrix [v, bl1, bl2]; otherwise the input must be a 1-D integer vector."""
    t = tree if isinstance(tree, np.ndarray) else np.asarray(tree)
    lib = _capi.load()
    status = np.zeros(1, dtype=np.int32)
    if t.ndim == 2 and t.dtype == np.float64:
        m = np.ascontiguousarray(t)
        if m.shape[1] != 3:
            raise TypeError("matrix input must have 3 columns [v, bl1, bl2]")
        k = m.shape[0]
        out = np.zeros((k + 1, k + 1), dtype=np.float64)
        rc = lib.p2v_cophenetic_matrix_host(ptr(m), 1, k, int(bool(unrooted)), 0, k + 1, ptr(out), ptr(status))
        _capi.check(rc, "cophenetic_distances")
        raise_status(status, m[:, 0].astype(np.int64))
        return out
    if t.ndim != 1:
        raise TypeError("tree must be a 1-D integer vector or a 2-D float64 matrix")
    v = as_index_array(t, 1, "cophenetic_distances")
    k = v.shape[0]
    out = np.zeros((k + 1, k + 1), dtype=np.float64)
    rc = lib.p2v_cophenetic_host(ptr(v), 8, None, 1, k, int(bool(unrooted)), 0, k + 1, ptr(out), ptr(status))
    _capi.check(rc, "cophenetic_distances")
    raise_status(status, v)
    return out


NODEWISE_DISTANCES = {"cophenetic": cophenetic_distances}


def pairwise_distances(tree, metric="cophenetic") -> np.ndarray:
    """py-phylo2vec/phylo2vec/stats/nodewise.py:31-52 (rooted only)."""
    return NODEWISE_DISTANCES[metric](tree)


def cov(tree) -> np.ndarray:
    """Variance-covariance matrix of a Phylo2Vec vector (unit lengths) or matrix: float64 (n,n),
    like core.vcv (py-phylo2vec/src/lib.rs:166-172; phylo2vec/src/vector/graph.rs:211-282;
    phylo2vec/src/matrix/graph.rs:61-64; py-phylo2vec/phylo2vec/stats/nodewise.py:55-71)."""
    t = tree if isinstance(tree, np.ndarray) else np.asarray(tree)
    lib = _capi.load()
    status = np.zeros(1, dtype=np.int32)
    if t.ndim == 2 and t.dtype == np.float64:
        m = np.ascontiguousarray(t)
        if m.shape[1] != 3:
            raise TypeError("matrix input must have 3 columns [v, bl1, bl2]")
        k, v = m.shape[0], m[:, 0].astype(np.int64)
    elif t.ndim == 1:
        v = as_index_array(t, 1, "vcv")
        k, m = v.shape[0], None
    else:
        raise TypeError("tree must be a 1-D integer vector or a 2-D float64 matrix")
    if k == 0:      # the reference's assert, vector/graph.rs:214-217
        raise AssertionError("Variance-covariance matrix not supported for trees with < 2 leaves.")
    out = np.zeros((k + 1, k + 1), dtype=np.float64)
    if m is not None:
        rc = lib.p2v_vcv_matrix_host(ptr(m), 1, k, 0, k + 1, ptr(out), ptr(status))
    else:
        rc = lib.p2v_vcv_host(ptr(v), 8, None, 1, k, 0, k + 1, ptr(out), ptr(status))
    _capi.check(rc, "vcv")
    raise_status(status, v)
    return out


vcv = cov
