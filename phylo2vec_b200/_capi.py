"""ctypes binding of the C ABI declared in include/phylo2vec_b200.h.

There is no CPU fallback: if the CUDA library is missing or cannot be loaded the
import fails loudly, and every entry point needs a CUDA device.
"""

from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libphylo2vec_b200.so")

P2V_OK = 0
P2V_TREE_MALFORMED = -2
P2V_TREE_NEGATIVE_BRANCH = -3

_i64 = ctypes.c_int64
_int = ctypes.c_int
_vp = ctypes.c_void_p
_sz = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/phylo2vec_b200.h one to one
SIGNATURES = {
    "p2v_version": (_int, []),
    "p2v_last_error": (ctypes.c_char_p, []),
    "p2v_launch_count": (_i64, []),
    "p2v_decode_workspace_bytes": (_sz, [_i64, _i64]),
    "p2v_encode_workspace_bytes": (_sz, [_i64, _i64]),
    "p2v_cophenetic_workspace_bytes": (_sz, [_i64, _i64]),
    "p2v_check_v_batch": (_int, [_vp, _int, _i64, _i64, _vp, _vp]),
    "p2v_cophenetic_batch": (_int, [_vp, _int, _vp, _i64, _i64, _int, _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "p2v_cophenetic_matrix_batch": (_int, [_vp, _i64, _i64, _int, _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "p2v_cophenetic_prepare_batch": (_int, [_vp, _int, _vp, _i64, _i64, _int, _vp, _vp, _sz, _vp]),
    "p2v_cophenetic_rows_batch": (_int, [_vp, _sz, _i64, _i64, _int, _i64, _i64, _vp, _vp]),
    "p2v_check_v_host": (_int, [_vp, _int, _i64, _i64, _vp]),
    "p2v_vcv_batch": (_int, [_vp, _int, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "p2v_vcv_matrix_batch": (_int, [_vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "p2v_vcv_rows_batch": (_int, [_vp, _sz, _i64, _i64, _int, _i64, _i64, _vp, _vp]),
    "p2v_node_depths_batch": (_int, [_vp, _int, _vp, _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "p2v_vcv_host": (_int, [_vp, _int, _vp, _i64, _i64, _i64, _i64, _vp, _vp]),
    "p2v_vcv_matrix_host": (_int, [_vp, _i64, _i64, _i64, _i64, _vp, _vp]),
    "p2v_node_depths_host": (_int, [_vp, _int, _vp, _i64, _i64, _vp, _vp]),
    "p2v_newick_max_bytes": (_sz, [_i64]),
    "p2v_newick_workspace_bytes": (_sz, [_i64, _i64]),
    "p2v_to_newick_batch": (_int, [_vp, _int, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "p2v_to_newick_host": (_int, [_vp, _int, _i64, _i64, _vp, _i64, _vp, _vp]),
    "p2v_robinson_foulds_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "p2v_robinson_foulds_batch": (_int, [_vp, _vp, _int, _i64, _i64, _i64, _int, _vp, _vp, _vp, _sz, _vp]),
    "p2v_robinson_foulds_host": (_int, [_vp, _vp, _int, _i64, _i64, _i64, _int, _vp, _vp]),
    "p2v_newick_matrix_max_bytes": (_sz, [_i64]),
    "p2v_newick_matrix_workspace_bytes": (_sz, [_i64, _i64]),
    "p2v_to_newick_matrix_batch": (_int, [_vp, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "p2v_to_newick_matrix_host": (_int, [_vp, _i64, _i64, _vp, _i64, _vp, _vp]),
    "p2v_from_newick_matrix_workspace_bytes": (_sz, [_i64, _i64]),
    "p2v_from_newick_matrix_batch": (_int, [_vp, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "p2v_from_newick_matrix_host": (_int, [_vp, _i64, _i64, _i64, _vp, _vp]),
    "p2v_csv_workspace_bytes": (_sz, [_i64]),
    "p2v_csv_parse_batch": (_int, [_vp, _i64, _int, _int, _vp, _int, _i64, _vp, _vp, _sz, _vp]),
    "p2v_balance_workspace_bytes": (_sz, [_i64, _i64]),
    "p2v_balance_batch": (_int, [_vp, _int, _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "p2v_balance_host": (_int, [_vp, _int, _i64, _i64, _vp, _vp]),
    "p2v_from_newick_workspace_bytes": (_sz, [_i64, _i64, _int]),
    "p2v_from_newick_batch": (_int, [_vp, _i64, _vp, _i64, _i64, _vp, _int, _vp, _vp, _sz, _vp]),
    "p2v_from_newick_host": (_int, [_vp, _i64, _i64, _i64, _vp, _int, _vp]),
    "p2v_cophenetic_host": (_int, [_vp, _int, _vp, _i64, _i64, _int, _i64, _i64, _vp, _vp]),
    "p2v_cophenetic_matrix_host": (_int, [_vp, _i64, _i64, _int, _i64, _i64, _vp, _vp]),
    "p2v_host_release": (None, []),
    "p2v_host_cache_bytes": (_i64, []),
    "p2v_host_bind_thread": (_int, []),
    "p2v_host_cache_limit": (None, [_i64]),
    "p2v_debug_phase_ns": (_int, [_int, _vp]),
    "p2v_check_m_batch": (_int, [_vp, _i64, _i64, _vp, _vp]),
    "p2v_check_m_host": (_int, [_vp, _i64, _i64, _vp]),
    "p2v_fill_f64": (_int, [_vp, _i64, ctypes.c_double, _vp]),
}
for _name in ("to_pairs", "to_ancestry", "to_edges", "from_pairs", "from_ancestry", "from_edges"):
    SIGNATURES[f"p2v_{_name}_batch"] = (_int, [_vp, _int, _i64, _i64, _vp, _vp, _vp, _sz, _vp])
    SIGNATURES[f"p2v_{_name}_host"] = (_int, [_vp, _int, _i64, _i64, _vp, _vp])


class P2VError(RuntimeError):
    """A C-ABI call returned a negative code."""


_lib = None


def load() -> ctypes.CDLL:
    """Load libphylo2vec_b200.so (built by ``python -m phylo2vec_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("P2V_B200_LIB") or LIB_PATH      # e.g. the -DP2V_DEBUG_BOUNDS build, libphylo2vec_b200_dbg.so
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -m phylo2vec_b200.build` "
            "(needs nvcc; sm_100a). There is no CPU fallback.")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    import atexit
    atexit.register(release_host_cache)
    return lib


def release_host_cache() -> None:
    """Free the staging (device + pinned memory) the *_host entry points keep for the calling thread."""
    if _lib is not None:
        _lib.p2v_host_release()


def host_cache_bytes() -> int:
    return int(load().p2v_host_cache_bytes())


def set_host_cache_limit(nbytes: int) -> None:
    """Bytes of staging a host thread may keep between *_host calls (-1: everything)."""
    load().p2v_host_cache_limit(int(nbytes))


def check(rc: int, what: str) -> None:
    if rc != P2V_OK:
        msg = load().p2v_last_error()
        raise P2VError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(load().p2v_launch_count())
