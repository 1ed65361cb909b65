"""Same module layout as py-phylo2vec/phylo2vec/base (ancestry, pairs, edges, newick output)."""
from .ancestry import from_ancestry, to_ancestry
from .edges import from_edges, to_edges
from .newick import from_newick, to_newick
from .pairs import from_pairs, to_pairs

__all__ = ["from_ancestry", "to_ancestry", "from_edges", "to_edges", "from_pairs", "to_pairs", "to_newick", "from_newick"]
