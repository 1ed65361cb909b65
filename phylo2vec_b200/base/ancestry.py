"""Vector <-> ancestry matrix; drop-in for py-phylo2vec/phylo2vec/base/ancestry.py."""

import numpy as np

from .._host import as_index_array, run_host


def from_ancestry(ancestry) -> np.ndarray:
    """Convert an ancestry matrix (k,3) [child1, child2, parent] to a vector.

    Same contract as py-phylo2vec/phylo2vec/base/ancestry.py:10-27
    (core.from_ancestry, phylo2vec/src/vector/convert.rs:129-136); computed on the
    GPU.  Returns an int64 array of shape (k,)."""
    a = as_index_array(ancestry, 2, "from_ancestry")
    if a.shape[1] != 3:
        raise TypeError("from_ancestry: expected rows of 3 integers")
    k = a.shape[0]
    return run_host("from_ancestry", a, 1, k, np.zeros(k, dtype=np.int64))


def to_ancestry(v) -> np.ndarray:
    """Convert a Phylo2Vec vector to its ancestry matrix, int64 (k,3).

    Same contract as py-phylo2vec/phylo2vec/base/ancestry.py:30-66
    (core.get_ancestry, phylo2vec/src/vector/convert.rs:167-188)."""
    v = as_index_array(v, 1, "to_ancestry")
    k = v.shape[0]
    out = np.zeros((k, 3), dtype=np.int64)
    return run_host("to_ancestry", v, 1, k, out, v_for_msg=v)
