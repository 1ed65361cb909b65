"""Distances between trees; drop-in for py-phylo2vec/phylo2vec/stats/treewise.py:10-58."""

import numpy as np

from .. import _capi
from .._host import as_index_array, ptr, raise_status


def _topology(tree, what) -> np.ndarray:
    """PyTree (py-phylo2vec/src/lib.rs:20-24, 295-306): a 2-D float64 matrix is reduced to its vector."""
    t = tree if isinstance(tree, np.ndarray) else np.asarray(tree)
    if t.ndim == 2 and t.dtype == np.float64:
        return as_index_array(np.nan_to_num(t[:, 0]).clip(0).astype(np.int64), 1, what)
    if t.ndim != 1:
        raise TypeError("tree must be a 1-D integer vector or a 2-D float64 matrix")
    return as_index_array(t, 1, what)


def robinson_foulds(tree1, tree2, normalize: bool = False) -> float:
    """Robinson-Foulds distance of two trees (vectors or matrices; only the topology counts), like
    core.robinson_foulds (phylo2vec/src/vector/distance.rs:41-74): the number of bipartitions found in one
    tree only; ``normalize`` divides by the total number of bipartitions."""
    v1 = _topology(tree1, "robinson_foulds")
    v2 = _topology(tree2, "robinson_foulds")
    if v1.shape[0] != v2.shape[0]:           # the reference's assert_eq!, distance.rs:42-46
        raise AssertionError("Trees must have the same number of leaves")
    out = np.zeros(1, dtype=np.float64)
    status = np.zeros(1, dtype=np.int32)
    rc = _capi.load().p2v_robinson_foulds_host(ptr(v1), ptr(v2), 8, 1, 1, v1.shape[0], int(bool(normalize)), ptr(out),
                                               ptr(status))
    _capi.check(rc, "robinson_foulds")
    if status[0] != 0:                        # which of the two vectors is invalid
        for v in (v1, v2):
            st = np.zeros(1, dtype=np.int32)
            _capi.check(_capi.load().p2v_check_v_host(ptr(v), 8, 1, v.shape[0], ptr(st)), "check_v")
            raise_status(st, v)
    return float(out[0])
