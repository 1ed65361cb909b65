"""Hot-path helpers named like py-phylo2vec/phylo2vec/utils/matrix.py: check_matrix (:8-18 ->
core.check_m, py-phylo2vec/src/lib.rs:68-88; phylo2vec/src/matrix/base.rs:76-95)."""

import numpy as np

from .. import _capi
from .._host import ptr


def _as_usize(col: np.ndarray) -> np.ndarray:
    """Rust's saturating ``f64 as usize`` (matrix/base.rs:80): NaN and negatives -> 0, truncation."""
    c = np.nan_to_num(col, nan=0.0, posinf=9.0e18, neginf=0.0)
    return np.clip(np.trunc(c), 0.0, 9.0e18).astype(np.int64)


def check_matrix(m) -> None:
    """Raise AssertionError with the reference's message if ``m`` is not a Phylo2Vec matrix:
    "Matrix should not be empty", check_v's message for column 0, "Matrix must have at least 3
    columns (vector and two branch lengths)", "Branch lengths must be positive" -- in the order
    the reference asserts them (phylo2vec/src/matrix/base.rs:76-95)."""
    a = m if isinstance(m, np.ndarray) else np.asarray(m)
    if a.ndim != 2 or a.dtype != np.float64:          # PyReadonlyArray2<f64> extraction
        raise TypeError("check_matrix: expected a 2-D float64 array")
    if a.size == 0:
        raise AssertionError("Matrix should not be empty")
    lib = _capi.load()
    k = a.shape[0]
    status = np.zeros(1, dtype=np.int32)
    if a.shape[1] != 3:
        # the reference validates column 0 before it looks at the column count
        v = np.ascontiguousarray(_as_usize(a[:, 0]))
        rc = lib.p2v_check_v_host(ptr(v), 8, 1, k, ptr(status))
        _capi.check(rc, "check_m")
        _raise(status, a)
        raise AssertionError("Matrix must have at least 3 columns (vector and two branch lengths)")
    a = np.ascontiguousarray(a)
    rc = lib.p2v_check_m_host(ptr(a), 1, k, ptr(status))
    _capi.check(rc, "check_m")
    _raise(status, a)


def _raise(status: np.ndarray, a: np.ndarray) -> None:
    code = int(status[0])
    if code > 0:
        i = code - 1
        val = int(_as_usize(a[i:i + 1, 0])[0])
        raise AssertionError(f"Validation failed: v[{i}] = {val} is out of bounds (max = {2 * i})")
    if code == _capi.P2V_TREE_NEGATIVE_BRANCH:
        raise AssertionError("Branch lengths must be positive")


check_m = check_matrix
