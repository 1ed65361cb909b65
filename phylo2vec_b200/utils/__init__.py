"""Same module layout as py-phylo2vec/phylo2vec/utils (hot-path part only)."""
from .matrix import check_m, check_matrix
from .vector import check_v, check_vector, get_node_depth, get_node_depths

__all__ = ["check_v", "check_vector", "check_m", "check_matrix", "get_node_depth", "get_node_depths"]
