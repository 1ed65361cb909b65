"""Validation helper named like the reference's (phylo2vec.utils.vector.check_vector ->
core.check_v, py-phylo2vec/src/lib.rs:48-66)."""

import numpy as np

from . import _capi
from ._host import as_index_array, ptr, raise_status


def check_v(v) -> None:
    """Raise AssertionError with the reference's message if any v[i] > 2i
    (phylo2vec/src/vector/base.rs:50-67)."""
    v = as_index_array(v, 1, "check_v")
    status = np.zeros(1, dtype=np.int32)
    rc = _capi.load().p2v_check_v_host(ptr(v), 8, 1, v.shape[0], ptr(status))
    _capi.check(rc, "check_v")
    raise_status(status, v)
