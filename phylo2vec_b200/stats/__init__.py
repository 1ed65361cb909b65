"""Same module layout as py-phylo2vec/phylo2vec/stats (node-wise distances / covariance, balance indices,
tree-wise distances)."""
from .balance import b2, leaf_depth_variance, sackin
from .nodewise import cophenetic_distances, cov, pairwise_distances, vcv
from .treewise import robinson_foulds

__all__ = ["cophenetic_distances", "pairwise_distances", "cov", "vcv", "sackin", "b2", "leaf_depth_variance",
           "robinson_foulds"]
