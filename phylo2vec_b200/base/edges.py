"""Vector <-> edge list; drop-in for py-phylo2vec/phylo2vec/base/edges.py."""

from typing import List, Tuple

import numpy as np

from .._host import as_index_array, run_host


def from_edges(edges) -> np.ndarray:
    """Convert a list of (child, parent) edges to a vector, int64 (k,).

    Same contract as py-phylo2vec/phylo2vec/base/edges.py:12-28
    (core.from_edges, phylo2vec/src/vector/convert.rs:201-213; an odd number of
    edges trips the assertion at :202)."""
    e = as_index_array(edges, 2, "from_edges") if len(edges) else np.zeros((0, 2), dtype=np.int64)
    if e.shape[1] != 2:
        raise TypeError("from_edges: expected (child, parent) pairs")
    if e.shape[0] % 2 != 0:
        raise AssertionError("The number of edges must be even")
    k = e.shape[0] // 2
    return run_host("from_edges", e, 1, k, np.zeros(k, dtype=np.int64))


def to_edges(v) -> List[Tuple[int, int]]:
    """Convert a vector to its list of (child, parent) edges (a list of tuples,
    like core.get_edges: py-phylo2vec/phylo2vec/base/edges.py:31-47,
    phylo2vec/src/vector/convert.rs:227-248)."""
    v = as_index_array(v, 1, "to_edges")
    k = v.shape[0]
    out = run_host("to_edges", v, 1, k, np.zeros((2 * k, 2), dtype=np.int64), v_for_msg=v)
    return [(int(a), int(b)) for a, b in out]
