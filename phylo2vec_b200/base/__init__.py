"""Same module layout as py-phylo2vec/phylo2vec/base (ancestry, pairs, edges)."""
from .ancestry import from_ancestry, to_ancestry
from .edges import from_edges, to_edges
from .pairs import from_pairs, to_pairs

__all__ = ["from_ancestry", "to_ancestry", "from_edges", "to_edges", "from_pairs", "to_pairs"]
