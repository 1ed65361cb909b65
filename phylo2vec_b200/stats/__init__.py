"""Same module layout as py-phylo2vec/phylo2vec/stats (nodewise part only)."""
from .nodewise import cophenetic_distances, cov, pairwise_distances, vcv

__all__ = ["cophenetic_distances", "pairwise_distances", "cov", "vcv"]
