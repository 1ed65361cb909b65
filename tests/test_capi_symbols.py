"""CPU-side checks of the drop-in boundary: the shared library loads and exports
every symbol that include/phylo2vec_b200.h declares (no compute calls here)."""

import ctypes
import os
import re

from phylo2vec_b200 import _capi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "phylo2vec_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(p2v_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    build.build()
    lib = ctypes.CDLL(_capi.LIB_PATH)
    names = _header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_ctypes_signatures_cover_header():
    assert sorted(_capi.SIGNATURES) == _header_symbols()
    lib = _capi.load()
    assert lib.p2v_version() >= 100
    assert lib.p2v_launch_count() >= 0


def test_workspace_queries_need_no_gpu():
    lib = _capi.load()
    assert lib.p2v_decode_workspace_bytes(1000, 999) == 0          # warp-per-tree path: no scratch
    assert lib.p2v_decode_workspace_bytes(4, 5000) == 0           # mid-size trees: shared memory only
    assert lib.p2v_decode_workspace_bytes(4, 20000) > 0
    assert lib.p2v_encode_workspace_bytes(4, 5000) > 0
    assert lib.p2v_cophenetic_workspace_bytes(2, 1999) > 2 * 1999 * 100


def test_argument_validation_without_gpu():
    lib = _capi.load()
    assert lib.p2v_to_ancestry_batch(None, 3, 1, 4, None, None, None, 0, None) == -1   # bad idx_bytes
    assert b"idx_bytes" in lib.p2v_last_error()
    assert lib.p2v_to_ancestry_batch(None, 8, 1, 4, None, None, None, 0, None) == -1   # null buffers
    assert lib.p2v_cophenetic_batch(None, 8, None, 1, 4, 0, 3, 9, None, None, None, 0, None) == -1  # rows > n
    assert lib.p2v_to_pairs_batch(None, 8, 0, 10, None, None, None, 0, None) == 0      # empty batch is fine


def test_python_api_mirrors_reference_names():
    import phylo2vec_b200 as p2v
    for name in ("to_ancestry", "to_pairs", "to_edges", "from_ancestry", "from_pairs", "from_edges",
                 "cophenetic_distances", "pairwise_distances", "check_v"):
        assert callable(getattr(p2v, name))
    from phylo2vec_b200.base.ancestry import from_ancestry, to_ancestry  # noqa: F401
    from phylo2vec_b200.base.edges import from_edges, to_edges  # noqa: F401
    from phylo2vec_b200.base.pairs import from_pairs, to_pairs  # noqa: F401
    from phylo2vec_b200.stats.nodewise import cophenetic_distances  # noqa: F401
