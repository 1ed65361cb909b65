"""phylo2vec_b200 -- B200 (sm_100a) implementation of phylo2vec's batched
decode / encode / cophenetic-distance path.

The single-tree functions keep the reference's names, dtypes and layouts
(py-phylo2vec/phylo2vec/base/*.py, stats/nodewise.py); the ``*_batch`` functions
are the device-resident additions (torch tensors via DLPack).  Everything runs
as hand-written CUDA behind the C ABI in include/phylo2vec_b200.h; there is no
CPU fallback.
"""

from . import _capi
from .base.ancestry import from_ancestry, to_ancestry
from .base.edges import from_edges, to_edges
from .base.newick import from_newick, to_matrix, to_newick, to_vector
from .base.pairs import from_pairs, to_pairs
from .stats.nodewise import cophenetic_distances, cov, pairwise_distances
from .utils.matrix import check_m, check_matrix
from .utils.vector import check_v, get_node_depth, get_node_depths

__all__ = [
    "to_ancestry", "to_pairs", "to_edges", "from_ancestry", "from_pairs", "from_edges",
    "cophenetic_distances", "pairwise_distances", "check_v", "check_m", "check_matrix", "cov", "to_newick", "from_newick", "to_matrix", "to_vector", "get_node_depths", "get_node_depth",
    "to_ancestry_batch", "to_pairs_batch", "to_edges_batch", "from_ancestry_batch",
    "from_pairs_batch", "from_edges_batch", "cophenetic_distances_batch", "check_v_batch", "check_m_batch",
    "vcv_batch", "node_depths_batch", "to_newick_batch", "from_newick_batch", "balance_indices_batch",
    "robinson_foulds_batch", "to_matrix_batch",
]

_BATCH = {"to_ancestry_batch", "to_pairs_batch", "to_edges_batch", "from_ancestry_batch",
          "from_pairs_batch", "from_edges_batch", "cophenetic_distances_batch", "check_v_batch", "check_m_batch",
          "vcv_batch", "node_depths_batch", "to_newick_batch", "from_newick_batch", "balance_indices_batch",
          "robinson_foulds_batch", "to_matrix_batch"}


def __getattr__(name):
    # torch is only imported when the device API is used
    if name in _BATCH:
        from . import batch
        return getattr(batch, name)
    raise AttributeError(name)
