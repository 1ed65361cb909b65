"""Hot-path helpers named like py-phylo2vec/phylo2vec/utils/vector.py: check_vector
(:16-26 -> core.check_v, py-phylo2vec/src/lib.rs:48-66), get_node_depths (:297-331) and
get_node_depth (:260-294)."""

import numpy as np

from .. import _capi
from .._host import as_index_array, ptr, raise_status


def check_v(v) -> None:
    """Raise AssertionError with the reference's message if any v[i] > 2i
    (phylo2vec/src/vector/base.rs:50-67)."""
    v = as_index_array(v, 1, "check_v")
    status = np.zeros(1, dtype=np.int32)
    rc = _capi.load().p2v_check_v_host(ptr(v), 8, 1, v.shape[0], ptr(status))
    _capi.check(rc, "check_v")
    raise_status(status, v)


check_vector = check_v


def _split(tree, what):
    """PyTree dispatch (py-phylo2vec/src/lib.rs:20-24): 2-D float64 = matrix, else 1-D integer vector."""
    t = tree if isinstance(tree, np.ndarray) else np.asarray(tree)
    if t.ndim == 2 and t.dtype == np.float64:
        m = np.ascontiguousarray(t)
        if m.shape[1] != 3:
            raise TypeError("matrix input must have 3 columns [v, bl1, bl2]")
        return m, 0, m[:, 0].astype(np.int64)
    if t.ndim != 1:
        raise TypeError("tree must be a 1-D integer vector or a 2-D float64 matrix")
    v = as_index_array(t, 1, what)
    return v, 8, v


def get_node_depths(vector_or_matrix) -> np.ndarray:
    """Depth (distance from the root) of every node, float64 (2*n_leaves - 1,); unit branch
    lengths for a vector, the matrix's lengths otherwise (phylo2vec/src/vector/ops.rs:201-233,
    matrix/ops.rs:42-45; py-phylo2vec/src/lib.rs:281-287)."""
    a, idx_bytes, v = _split(vector_or_matrix, "get_node_depths")
    k = a.shape[0]
    out = np.zeros(2 * k + 1, dtype=np.float64)
    status = np.zeros(1, dtype=np.int32)
    rc = _capi.load().p2v_node_depths_host(ptr(a), idx_bytes, None, 1, k, ptr(out), ptr(status))
    _capi.check(rc, "get_node_depths")
    raise_status(status, v)
    return out


def get_node_depth(vector_or_matrix, node: int) -> float:
    """Depth of one node; ValueError with the reference's text when the node is out of range
    (phylo2vec/src/vector/ops.rs:243-261, py-phylo2vec/src/lib.rs:267-278)."""
    t = np.asarray(vector_or_matrix)
    n_nodes = 2 * t.shape[0] + 1
    if node < 0:
        raise ValueError(f"Node index must be non-negative, got {node}")
    if node >= n_nodes:
        raise ValueError(f"Node index out of bounds. Max node = {n_nodes - 1}, got node = {node}")
    return float(get_node_depths(vector_or_matrix)[node])
