"""Vector <-> list of pairs; drop-in for py-phylo2vec/phylo2vec/base/pairs.py."""

from typing import List, Tuple

import numpy as np

from .._host import as_index_array, run_host


def from_pairs(pairs) -> np.ndarray:
    """Convert a list of (branch, leaf) pairs to a vector, int64 (k,).

    Same contract as py-phylo2vec/phylo2vec/base/pairs.py:11-28
    (core.from_pairs, phylo2vec/src/vector/convert.rs:23-30)."""
    p = as_index_array(pairs, 2, "from_pairs") if len(pairs) else np.zeros((0, 2), dtype=np.int64)
    if p.shape[1] != 2:
        raise TypeError("from_pairs: expected pairs of integers")
    k = p.shape[0]
    return run_host("from_pairs", p, 1, k, np.zeros(k, dtype=np.int64))


def to_pairs(v) -> List[Tuple[int, int]]:
    """Convert a vector to its list of (branch, leaf) pairs (a list of tuples,
    like core.get_pairs: py-phylo2vec/phylo2vec/base/pairs.py:31-48,
    phylo2vec/src/vector/convert.rs:103-109)."""
    v = as_index_array(v, 1, "to_pairs")
    k = v.shape[0]
    out = run_host("to_pairs", v, 1, k, np.zeros((k, 2), dtype=np.int64), v_for_msg=v)
    return [(int(a), int(b)) for a, b in out]
