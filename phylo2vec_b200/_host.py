"""Host-side (numpy) plumbing behind the reference-named single-tree functions.

Every call goes through the C ABI's *_host entry points (H2D copy, CUDA
kernels, D2H copy).  No CPU implementation exists here."""

from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _capi

_NEG_MSG = "can't convert negative int to unsigned"


def as_index_array(x, ndim: int, what: str) -> np.ndarray:
    """int64 C-contiguous array; mirrors PyO3's Vec<usize> extraction
    (py-phylo2vec/src/lib.rs:108-136): non-integers are a TypeError, negatives an
    OverflowError."""
    a = np.asarray(x)
    if a.dtype == object or not (np.issubdtype(a.dtype, np.integer) or a.size == 0):
        raise TypeError(f"{what}: expected a sequence of non-negative integers, got dtype {a.dtype}")
    if a.ndim != ndim:
        if a.size == 0 and ndim >= 1:
            a = a.reshape((0,) + (1,) * 0) if ndim == 1 else a
        if a.ndim != ndim:
            raise TypeError(f"{what}: expected {ndim} dimensions, got {a.ndim}")
    a = np.ascontiguousarray(a, dtype=np.int64)
    if a.size and a.min() < 0:
        raise OverflowError(_NEG_MSG)
    return a


def ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def raise_status(status: np.ndarray, v: Optional[np.ndarray] = None) -> None:
    bad = np.flatnonzero(status)
    if bad.size == 0:
        return
    b = int(bad[0])
    code = int(status[b])
    if code > 0:
        i = code - 1
        val = int(v.reshape(status.size, -1)[b, i]) if v is not None else "?"
        # message of check_max, phylo2vec/src/vector/base.rs:60-67
        raise AssertionError(f"Validation failed: v[{i}] = {val} is out of bounds (max = {2 * i})")
    if code == _capi.P2V_TREE_NEGATIVE_BRANCH:      # check_m, phylo2vec/src/matrix/base.rs:88-94
        raise AssertionError("Branch lengths must be positive")
    raise ValueError("malformed tree: the input is not a phylo2vec ancestry / pair / edge list")


def run_host(name: str, x: np.ndarray, B: int, k: int, out: np.ndarray, v_for_msg=None) -> np.ndarray:
    status = np.zeros(max(B, 1), dtype=np.int32)
    rc = getattr(_capi.load(), f"p2v_{name}_host")(ptr(x), 8, B, k, ptr(out), ptr(status))
    _capi.check(rc, name)
    raise_status(status[:B], v_for_msg)
    return out
