"""Phylo2Vec vectors / matrices -> files and strings; drop-in for py-phylo2vec/phylo2vec/io/writer.py:13-79,
plus the batched form (one tree per line) for tensors that live on the GPU.

Formatting is numpy's, exactly as in the reference ("%d" for vectors, "%.18e" for matrices), so that the bytes
on disk are the reference's; a device batch comes back through pinned host memory first."""

from io import StringIO
from typing import Dict, Optional

import numpy as np

from ..base.newick import to_newick
from ._validation import check_path


def write(vector_or_matrix: np.ndarray, delimiter: str = ",") -> str:
    """io/writer.py:13-24"""
    if vector_or_matrix.ndim == 2:
        buffer = StringIO()
        np.savetxt(buffer, vector_or_matrix, delimiter=delimiter)
        return buffer.getvalue()
    if vector_or_matrix.ndim == 1:
        return np.array2string(vector_or_matrix, separator=delimiter)[1:-1]
    raise ValueError("vector_or_matrix should either be a vector (ndim == 1) or matrix (ndim == 2)")


def save(vector_or_matrix: np.ndarray, filepath, delimiter: str = ",") -> None:
    """Save a Phylo2Vec vector or matrix to a file: io/writer.py:27-50."""
    check_path(filepath, "array")
    if vector_or_matrix.ndim == 2:
        fmt = "%.18e"
    elif vector_or_matrix.ndim == 1:
        fmt = "%d"
    else:
        raise ValueError("vector_or_matrix should either be a vector (ndim == 1) or matrix (ndim == 2)")
    np.savetxt(filepath, vector_or_matrix, fmt=fmt, delimiter=delimiter)


def save_batch(trees, filepath, delimiter: str = ",") -> None:
    """A batch -- int ``(B,k)`` vectors or float64 ``(B,k,3)`` matrices, a torch tensor (any device) or a numpy
    array -- to a file with one tree per line, in the formats of :func:`save`; :func:`reader.load_batch` reads it
    back bit for bit."""
    check_path(filepath, "array")
    if not isinstance(trees, np.ndarray):
        import torch
        t = trees.detach()
        if t.is_cuda:
            host = torch.empty(t.shape, dtype=t.dtype).pin_memory()
            host.copy_(t)
            t = host
        trees = t.numpy()
    if trees.ndim == 2 and np.issubdtype(trees.dtype, np.integer):
        np.savetxt(filepath, trees, fmt="%d", delimiter=delimiter)
    elif trees.ndim == 3 and trees.shape[2] == 3 and np.issubdtype(trees.dtype, np.floating):
        np.savetxt(filepath, trees.reshape(trees.shape[0], -1), fmt="%.18e", delimiter=delimiter)
    else:
        raise ValueError("a batch is an integer (B,k) array of vectors or a float (B,k,3) array of matrices")


def save_newick(vector_or_matrix: np.ndarray, filepath, labels: Optional[Dict[str, str]] = None) -> None:
    """Save a Phylo2Vec vector or matrix in Newick format: io/writer.py:53-79.  ``labels`` maps integer labels
    (as strings) to taxon names (utils/newick.py apply_label_mapping: whole leaf tokens are replaced)."""
    import re
    check_path(filepath, "newick")
    newick = to_newick(vector_or_matrix)
    if labels:
        newick = re.sub(r"(?<=[(,])(\d+)(?=[:,)])", lambda m: str(labels.get(m.group(1), labels.get(int(m.group(1)), m.group(1)))), newick)
    with open(filepath, "w", encoding="utf-8") as f:
        f.write(newick)
