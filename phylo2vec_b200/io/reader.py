"""Files and strings -> Phylo2Vec vectors / matrices; drop-in for py-phylo2vec/phylo2vec/io/reader.py:16-79,
plus the batched form feeding device tensors.

The text goes from the file into pinned host memory, across to the GPU, and is parsed THERE
(include/phylo2vec_b200.h, p2v_csv_parse_batch: integers digit by digit, float64 correctly rounded); the result
is validated by the library's own check_v / check_m kernels.  There is no host parser."""

import os
from io import TextIOBase
from typing import Union

import numpy as np

from .. import _capi
from ..base.newick import from_newick
from ..utils.matrix import check_matrix
from ..utils.vector import check_vector
from ._validation import check_path


def _text_bytes(filepath_or_buffer) -> bytes:
    if hasattr(filepath_or_buffer, "read"):
        data = filepath_or_buffer.read()
        return data.encode("ascii") if isinstance(data, str) else bytes(data)
    with open(filepath_or_buffer, "rb") as f:
        return f.read()


def _parse_on_gpu(data: bytes, delimiter: str, kind: int, device=None):
    """(flat tensor on the GPU, fields, rows, error bits).  kind 0: int64, 1: float64."""
    import torch
    if len(delimiter) != 1:
        raise ValueError("delimiter must be a single character")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    lib = _capi.load()
    n = len(data)
    with torch.cuda.device(dev):
        pinned = torch.empty(max(n, 1), dtype=torch.uint8).pin_memory()        # the staging the file is read into
        pinned[:n] = torch.frombuffer(bytearray(data), dtype=torch.uint8) if n else pinned[:0]
        text = pinned.to(dev, non_blocking=True)
        capacity = n // 2 + 1                                                   # a field and its separator take two bytes
        out = torch.empty(capacity, dtype=torch.int64 if kind == 0 else torch.float64, device=dev)
        counts = torch.zeros(3, dtype=torch.int64, device=dev)
        need = int(lib.p2v_csv_workspace_bytes(n))
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        rc = lib.p2v_csv_parse_batch(text.data_ptr(), n, ord(delimiter), kind, out.data_ptr(), 8, capacity,
                                     counts.data_ptr(), ws.data_ptr(), need, torch.cuda.current_stream(dev).cuda_stream)
        _capi.check(rc, "csv parse")
        fields, rows, err = (int(x) for x in counts.cpu())
    return out[:fields], fields, rows, err


def read(buffer: str, delimiter: str = ",") -> np.ndarray:
    """io/reader.py:16-17"""
    from io import StringIO
    return load(StringIO(buffer), delimiter=delimiter)


def load(filepath_or_buffer: Union[str, TextIOBase], delimiter: str = ",") -> np.ndarray:
    """Read a text / csv file into a Phylo2Vec vector (one integer per line) or matrix (k lines ``v,bl1,bl2``):
    io/reader.py:20-57.  Like numpy.genfromtxt(dtype=None) there, integers everywhere give an int64 array and
    anything else a float64 one; a single line or a single column gives a 1-D array."""
    if not hasattr(filepath_or_buffer, "read"):
        if not os.path.isfile(filepath_or_buffer):
            raise FileNotFoundError(f"File not found: {filepath_or_buffer}")
        check_path(filepath_or_buffer, "array")
    data = _text_bytes(filepath_or_buffer)
    flat, fields, rows, err = _parse_on_gpu(data, delimiter, 0)
    if err & 1:                                  # not all integers: floats
        flat, fields, rows, err = _parse_on_gpu(data, delimiter, 1)
        if err & 1:
            raise ValueError("could not parse the file: a field is neither an integer nor a float")
    if fields == 0 or rows == 0 or fields % rows != 0:
        raise ValueError("Input file should either be a vector (ndim == 1) or matrix (ndim == 2)")
    arr = flat.cpu().numpy()
    cols = fields // rows
    if rows > 1 and cols > 1:
        arr = arr.reshape(rows, cols)
    if arr.ndim == 1 and np.issubdtype(arr.dtype, np.integer):
        check_vector(arr)
    elif arr.ndim == 2 and np.issubdtype(arr.dtype, np.floating):
        check_matrix(arr)
    else:
        raise ValueError("Input file should either be a vector (ndim == 1) or matrix (ndim == 2)")
    return arr


def load_batch(filepath_or_buffer, kind: str = "vector", delimiter: str = ",", device=None, check: bool = True):
    """A file with ONE TREE PER LINE -> a tensor on the GPU: ``kind="vector"``: k integers per line -> int64
    ``(B,k)``; ``kind="matrix"``: 3k numbers per line (the rows ``v,bl1,bl2`` of the tree one after the other) ->
    float64 ``(B,k,3)``.  Validated on the device (check_v / check_m) unless ``check=False``."""
    from .. import batch
    if kind not in ("vector", "matrix"):
        raise ValueError("kind must be 'vector' or 'matrix'")
    if not hasattr(filepath_or_buffer, "read"):
        if not os.path.isfile(filepath_or_buffer):
            raise FileNotFoundError(f"File not found: {filepath_or_buffer}")
        check_path(filepath_or_buffer, "array")
    data = _text_bytes(filepath_or_buffer)
    flat, fields, rows, err = _parse_on_gpu(data, delimiter, 0 if kind == "vector" else 1, device)
    if err & 1:
        raise ValueError(f"could not parse the file as a batch of {kind} rows")
    if fields == 0 or rows == 0 or fields % rows != 0:
        raise ValueError("every line of a batch file must hold the same number of fields")
    cols = fields // rows
    if kind == "vector":
        V = flat.reshape(rows, cols).contiguous()
        if check:
            batch.raise_for_status(batch.check_v_batch(V), V)
        return V
    if cols % 3 != 0:
        raise ValueError("a matrix line holds 3 numbers per row of the tree")
    M = flat.reshape(rows, cols // 3, 3).contiguous()
    if check:
        batch.raise_for_status(batch.check_m_batch(M), M[..., 0].long())
    return M


def load_newick(filepath_or_buffer: str) -> np.ndarray:
    """Read a Newick string / file into a Phylo2Vec vector or matrix: io/reader.py:60-79."""
    if os.path.isfile(filepath_or_buffer):
        check_path(filepath_or_buffer, "newick")
        with open(filepath_or_buffer, "r", encoding="utf-8") as f:
            newick = f.read().strip()
    else:
        newick = filepath_or_buffer
    return from_newick(newick)
