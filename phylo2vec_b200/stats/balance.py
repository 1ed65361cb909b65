"""Tree balance indices; drop-in for py-phylo2vec/phylo2vec/stats/balance.py:8-101 (Sackin, B2, leaf
depth variance of the topology of a vector or matrix)."""

import numpy as np

from .. import _capi
from .._host import as_index_array, ptr, raise_status


def _topology(vector_or_matrix) -> np.ndarray:
    t = vector_or_matrix if isinstance(vector_or_matrix, np.ndarray) else np.asarray(vector_or_matrix)
    if t.ndim == 2:
        t = t[:, 0].astype(int)          # branch lengths are ignored (stats/balance.py:30-31)
    elif t.ndim != 1:
        raise ValueError("vector_or_matrix should either be a vector (ndim == 1) or matrix (ndim == 2)")
    return as_index_array(t, 1, "balance")


def _indices(vector_or_matrix) -> np.ndarray:
    v = _topology(vector_or_matrix)
    out = np.zeros(3, dtype=np.float64)
    status = np.zeros(1, dtype=np.int32)
    rc = _capi.load().p2v_balance_host(ptr(v), 8, 1, v.shape[0], ptr(out), ptr(status))
    _capi.check(rc, "balance")
    raise_status(status, v)
    return out


def sackin(vector_or_matrix) -> int:
    """Sum of the depths of all leaves, like core.sackin (phylo2vec/src/vector/balance.rs:14-18)."""
    return int(_indices(vector_or_matrix)[0])


def leaf_depth_variance(vector_or_matrix) -> float:
    """Variance of the leaf depths, like core.leaf_depth_variance (vector/balance.rs:20-33)."""
    return float(_indices(vector_or_matrix)[1])


def b2(vector_or_matrix) -> float:
    """B2 = sum d_i 2^-d_i over the leaves, like core.b2 (vector/balance.rs:49-58)."""
    return float(_indices(vector_or_matrix)[2])
