"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE -- see p2v_oracle.c).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this module.  Nothing under ``phylo2vec_b200/`` does.

Every function restates one reference function (paths relative to
/root/reference); the citation is in the docstring and in p2v_oracle.c.
"""

from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libp2v_oracle.so")
_SRC = os.path.join(_HERE, "p2v_oracle.c")

EBADINPUT = -2


def build(force: bool = False) -> str:
    """Compile oracle/p2v_oracle.c with oracle/Makefile (gcc, OpenMP)."""
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(_SRC)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        i64 = ctypes.c_int64
        p = ctypes.c_void_p
        _lib.p2vo_check_v.restype = i64
        _lib.p2vo_check_v.argtypes = [p, i64]
        _lib.p2vo_is_unordered.restype = ctypes.c_int
        _lib.p2vo_is_unordered.argtypes = [p, i64, p]
        for name in ("p2vo_to_pairs", "p2vo_to_pairs_avl", "p2vo_to_pairs_loop", "p2vo_to_pairs_list",
                     "p2vo_to_ancestry", "p2vo_to_edges", "p2vo_build_vector", "p2vo_from_pairs",
                     "p2vo_from_ancestry", "p2vo_from_edges"):
            f = getattr(_lib, name)
            f.restype = ctypes.c_int
            f.argtypes = [p, i64, p]
        _lib.p2vo_order_cherries.restype = ctypes.c_int
        _lib.p2vo_order_cherries.argtypes = [p, i64]
        _lib.p2vo_cophenetic_literal.restype = ctypes.c_int
        _lib.p2vo_cophenetic_literal.argtypes = [p, p, i64, ctypes.c_int, p]
        _lib.p2vo_cophenetic_rows.restype = ctypes.c_int
        _lib.p2vo_cophenetic_rows.argtypes = [p, p, i64, ctypes.c_int, i64, i64, p]
        _lib.p2vo_cophenetic_matrix_literal.restype = ctypes.c_int
        _lib.p2vo_cophenetic_matrix_literal.argtypes = [p, i64, ctypes.c_int, p]
        _lib.p2vo_vcv_literal.restype = ctypes.c_int
        _lib.p2vo_vcv_literal.argtypes = [p, p, i64, p]
        _lib.p2vo_vcv_rows.restype = ctypes.c_int
        _lib.p2vo_vcv_rows.argtypes = [p, p, i64, i64, i64, p]
        _lib.p2vo_node_depths.restype = ctypes.c_int
        _lib.p2vo_node_depths.argtypes = [p, p, i64, p]
        _lib.p2vo_to_newick.restype = i64
        _lib.p2vo_to_newick.argtypes = [p, i64, ctypes.c_char_p, i64]
        _lib.p2vo_balance.restype = ctypes.c_int
        _lib.p2vo_balance.argtypes = [p, i64, p]
        _lib.p2vo_newick_parse.restype = i64
        _lib.p2vo_newick_parse.argtypes = [ctypes.c_char_p, i64, p, i64]
        _lib.p2vo_from_newick.restype = i64
        _lib.p2vo_from_newick.argtypes = [ctypes.c_char_p, i64, p, i64]
        _lib.p2vo_from_newick_batch.restype = ctypes.c_int
        _lib.p2vo_from_newick_batch.argtypes = [p, i64, p, i64, i64, p, p, ctypes.c_int]
        _lib.p2vo_to_newick_batch.restype = ctypes.c_int
        _lib.p2vo_to_newick_batch.argtypes = [p, i64, i64, p, i64, p, ctypes.c_int]
        _lib.p2vo_num_threads.restype = ctypes.c_int
        _lib.p2vo_batch.restype = ctypes.c_int
        _lib.p2vo_batch.argtypes = [ctypes.c_int, p, i64, i64, p, p, ctypes.c_int]
        _lib.p2vo_cophenetic_batch.restype = ctypes.c_int
        _lib.p2vo_cophenetic_batch.argtypes = [p, p, i64, i64, ctypes.c_int, p, ctypes.c_int]
    return _lib


def _i64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64))


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class InvalidVector(AssertionError):
    """Raised where the reference panics in check_max (vector/base.rs:60-67)."""


def _raise(rc: int, v: np.ndarray):
    if rc > 0:
        i = rc - 1
        raise InvalidVector(
            f"Validation failed: v[{i}] = {int(v[i])} is out of bounds (max = {2 * i})")
    if rc < 0:
        raise ValueError(f"oracle: malformed input (code {rc})")


def check_v(v) -> None:
    """vector/base.rs:50-54"""
    v = _i64(v)
    _raise(int(lib().p2vo_check_v(_ptr(v), v.size)), v)


def check_m(m) -> None:
    """matrix/base.rs:76-95, in the reference's order: non-empty; check_v on column 0 cast like
    ``as usize`` (saturating: NaN and negatives give 0); exactly three columns; lengths >= 0."""
    m = np.asarray(m, dtype=np.float64)
    if m.size == 0:
        raise AssertionError("Matrix should not be empty")
    col = np.nan_to_num(m[:, 0], nan=0.0, posinf=9.0e18, neginf=0.0)
    check_v(np.clip(np.trunc(col), 0.0, 9.0e18).astype(np.int64))
    if m.shape[1] != 3:
        raise AssertionError("Matrix must have at least 3 columns (vector and two branch lengths)")
    for row in m:
        if not (row[1] >= 0.0 and row[2] >= 0.0):
            raise AssertionError("Branch lengths must be positive")


def is_unordered(v) -> bool:
    """vector/base.rs:92-101"""
    v = _i64(v)
    bad = ctypes.c_int64(0)
    r = lib().p2vo_is_unordered(_ptr(v), v.size, ctypes.byref(bad))
    _raise(int(bad.value), v)
    return bool(r)


def _decode(fn: str, v, cols: int, rows_per_k: int = 1) -> np.ndarray:
    v = _i64(v)
    k = v.size
    out = np.zeros((rows_per_k * k, cols), dtype=np.int64)
    _raise(int(getattr(lib(), fn)(_ptr(v), k, _ptr(out))), v)
    return out


def to_pairs(v) -> np.ndarray:
    """vector/convert.rs:103-109 -> (k,2)"""
    return _decode("p2vo_to_pairs", v, 2)


def to_pairs_avl(v) -> np.ndarray:
    """vector/convert.rs:83-87 + vector/avl.rs:35-52,193-211"""
    return _decode("p2vo_to_pairs_avl", v, 2)


def to_pairs_loop(v) -> np.ndarray:
    """vector/convert.rs:39-76"""
    return _decode("p2vo_to_pairs_loop", v, 2)


def to_pairs_list(v) -> np.ndarray:
    """naive list-insert form of vector/avl.rs:35-52"""
    return _decode("p2vo_to_pairs_list", v, 2)


def to_ancestry(v) -> np.ndarray:
    """vector/convert.rs:167-188 -> (k,3)"""
    return _decode("p2vo_to_ancestry", v, 3)


def to_edges(v) -> np.ndarray:
    """vector/convert.rs:227-248 -> (2k,2)"""
    return _decode("p2vo_to_edges", v, 2, rows_per_k=2)


def _encode(fn: str, a, cols: int) -> np.ndarray:
    a = _i64(a).reshape(-1, cols)
    k = a.shape[0]
    out = np.zeros(k, dtype=np.int64)
    rc = int(getattr(lib(), fn)(_ptr(a), k, _ptr(out)))
    if rc:
        raise ValueError(f"oracle: malformed input (code {rc})")
    return out


def build_vector(cherries) -> np.ndarray:
    """vector/convert.rs:320-337"""
    return _encode("p2vo_build_vector", cherries, 3)


def from_pairs(pairs) -> np.ndarray:
    """vector/convert.rs:23-30"""
    return _encode("p2vo_from_pairs", pairs, 2)


def from_ancestry(anc) -> np.ndarray:
    """vector/convert.rs:129-136"""
    return _encode("p2vo_from_ancestry", anc, 3)


def from_edges(edges) -> np.ndarray:
    """vector/convert.rs:201-213"""
    e = _i64(edges).reshape(-1, 2)
    out = np.zeros(e.shape[0] // 2, dtype=np.int64)
    rc = int(lib().p2vo_from_edges(_ptr(e), e.shape[0], _ptr(out)))
    if rc:
        raise ValueError(f"oracle: malformed input (code {rc})")
    return out


def order_cherries(anc) -> np.ndarray:
    """vector/convert.rs:449-477 (returns the reordered, normalised cherries)"""
    a = _i64(anc).reshape(-1, 3).copy()
    rc = int(lib().p2vo_order_cherries(_ptr(a), a.shape[0]))
    if rc:
        raise ValueError(f"oracle: malformed input (code {rc})")
    return a


def _split_tree(tree):
    """PyTree dispatch of py-phylo2vec/src/lib.rs:20-24,144-156."""
    t = np.asarray(tree)
    if t.ndim == 2:
        m = np.ascontiguousarray(t, dtype=np.float64)
        v = m[:, 0].astype(np.int64)  # matrix/base.rs:108-121 (`as usize`)
        bls = np.ascontiguousarray(m[:, 1:3])
        return v, bls
    if t.ndim == 1:
        return _i64(t), None
    raise TypeError("tree must be a 1-D vector or a 2-D matrix")


def cophenetic_distances(tree, unrooted: bool = False) -> np.ndarray:
    """Literal vector/graph.rs:20-90 (+ matrix/graph.rs:23-26). (n,n) float64."""
    v, bls = _split_tree(tree)
    k = v.size
    out = np.zeros((k + 1, k + 1), dtype=np.float64)
    _raise(int(lib().p2vo_cophenetic_literal(_ptr(v), _ptr(bls), k, int(unrooted), _ptr(out))), v)
    return out


def cophenetic_rows(tree, unrooted: bool, row_begin: int, row_end: int) -> np.ndarray:
    """Closed-form rows (long-double path sums) for sizes where the literal
    (2k+1)^2 scratch does not fit; pinned against the literal form in tests."""
    v, bls = _split_tree(tree)
    k = v.size
    out = np.zeros((row_end - row_begin, k + 1), dtype=np.float64)
    _raise(int(lib().p2vo_cophenetic_rows(_ptr(v), _ptr(bls), k, int(unrooted), row_begin, row_end,
                                          _ptr(out))), v)
    return out


def vcv(tree) -> np.ndarray:
    """Literal vector/graph.rs:211-258 (+ matrix/graph.rs:61-64). (n,n) float64; k == 0 is the
    reference's assert."""
    v, bls = _split_tree(tree)
    k = v.size
    if k == 0:
        raise AssertionError("Variance-covariance matrix not supported for trees with < 2 leaves.")
    out = np.zeros((k + 1, k + 1), dtype=np.float64)
    _raise(int(lib().p2vo_vcv_literal(_ptr(v), _ptr(bls), k, _ptr(out))), v)
    return out


def vcv_rows(tree, row_begin: int, row_end: int) -> np.ndarray:
    """Rows of the vcv matrix as node depth of the LCA (for sizes where (n,n) is too large)."""
    v, bls = _split_tree(tree)
    k = v.size
    out = np.zeros((row_end - row_begin, k + 1), dtype=np.float64)
    _raise(int(lib().p2vo_vcv_rows(_ptr(v), _ptr(bls), k, row_begin, row_end, _ptr(out))), v)
    return out


def get_node_depths(tree) -> np.ndarray:
    """vector/ops.rs:201-233 (+ matrix/ops.rs:42-45). (2k+1,) float64."""
    v, bls = _split_tree(tree)
    k = v.size
    out = np.zeros(2 * k + 1, dtype=np.float64)
    _raise(int(lib().p2vo_node_depths(_ptr(v), _ptr(bls), k, _ptr(out))), v)
    return out


def to_newick(v) -> str:
    """vector/convert.rs:394-397 (host only)"""
    v = _i64(v)
    n = int(lib().p2vo_to_newick(_ptr(v), v.size, None, 0))
    if n < 0:
        raise ValueError(f"oracle: to_newick failed ({n})")
    buf = ctypes.create_string_buffer(n + 2)
    lib().p2vo_to_newick(_ptr(v), v.size, buf, n + 2)
    return buf.value.decode()


# ---- batch drivers (CPU baseline) -----------------------------------------
OPS = {"to_pairs": 0, "to_ancestry": 1, "to_edges": 2, "from_pairs": 3, "from_ancestry": 4,
       "from_edges": 5}


def num_threads() -> int:
    return int(lib().p2vo_num_threads())


def batch(op: str, x: np.ndarray, k: int, nthreads: int = 0, out: Optional[np.ndarray] = None):
    """Run one reference call per tree, OpenMP parallel-for over trees."""
    x = _i64(x)
    B = x.shape[0]
    shape = {0: (B, k, 2), 1: (B, k, 3), 2: (B, 2 * k, 2), 3: (B, k), 4: (B, k), 5: (B, k)}[OPS[op]]
    if out is None:
        out = np.empty(shape, dtype=np.int64)
    status = np.zeros(B, dtype=np.int32)
    lib().p2vo_batch(OPS[op], _ptr(x), B, k, _ptr(out), _ptr(status), nthreads)
    return out, status


def cophenetic_batch(v: np.ndarray, bls: Optional[np.ndarray], unrooted: bool = False,
                     nthreads: int = 0) -> np.ndarray:
    v = _i64(v)
    B, k = v.shape
    if bls is not None:
        bls = np.ascontiguousarray(bls, dtype=np.float64)
    out = np.empty((B, k + 1, k + 1), dtype=np.float64)
    lib().p2vo_cophenetic_batch(_ptr(v), _ptr(bls), B, k, int(unrooted), _ptr(out), nthreads)
    return out


def to_newick_batch(v: np.ndarray, stride: int, nthreads: int = 0):
    """One reference call per tree (OpenMP over trees): uint8 (B, stride) buffer and lengths."""
    v = _i64(v)
    B, k = v.shape
    out = np.zeros((B, stride), dtype=np.uint8)
    lengths = np.zeros(B, dtype=np.int64)
    rc = lib().p2vo_to_newick_batch(_ptr(v), B, k, _ptr(out), stride, _ptr(lengths), nthreads)
    if rc:
        raise ValueError(f"oracle: to_newick_batch failed ({rc})")
    return out, lengths


def newick_parse(newick: str) -> np.ndarray:
    """newick::parse (newick/mod.rs:73-127): rows [c1, c2, p] in string order."""
    data = newick.encode("ascii")
    anc = np.zeros((len(data) + 1, 3), dtype=np.int64)
    rows = int(lib().p2vo_newick_parse(data, len(data), _ptr(anc), anc.shape[0]))
    if rows < 0:
        raise ValueError(f"oracle: newick parse failed ({rows})")
    return anc[:rows].copy()


def from_newick(newick: str) -> np.ndarray:
    """vector::convert::from_newick (convert.rs:268-278), with or without parent labels."""
    data = newick.encode("ascii")
    v = np.zeros(len(data) + 1, dtype=np.int64)
    k = int(lib().p2vo_from_newick(data, len(data), _ptr(v), v.size))
    if k < 0:
        raise ValueError(f"oracle: from_newick failed ({k})")
    return v[:k].copy()


def from_newick_batch(buf: np.ndarray, lengths: np.ndarray, k: int, nthreads: int = 0):
    """buf uint8 (B, stride); returns (v (B,k) int64, status int32)."""
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    lengths = np.ascontiguousarray(lengths, dtype=np.int64)
    B, stride = buf.shape
    v = np.zeros((B, k), dtype=np.int64)
    status = np.zeros(B, dtype=np.int32)
    lib().p2vo_from_newick_batch(_ptr(buf), stride, _ptr(lengths), B, k, _ptr(v), _ptr(status), nthreads)
    return v, status


def balance(v) -> np.ndarray:
    """[sackin, leaf_depth_variance, b2] (vector/balance.rs:14-58)."""
    v = _i64(v)
    out = np.zeros(3, dtype=np.float64)
    rc = lib().p2vo_balance(_ptr(v), v.size, _ptr(out))
    if rc:
        raise ValueError(f"oracle: balance failed ({rc})")
    return out


def _bipartitions(v) -> set:
    """vector/distance.rs:82-137 get_bipartitions: descendant leaf sets bottom-up through the ancestry (Python
    ints as bit sets), trivial splits skipped, each split canonicalised (:142-153): the smaller side, the side
    holding leaf 0 when the sides are equal."""
    anc = to_ancestry(v)
    k = anc.shape[0]
    n = k + 1
    desc = [1 << i for i in range(n)] + [0] * k
    for c1, c2, p in anc.tolist():
        desc[p] = desc[c1] | desc[c2]
    full = (1 << n) - 1
    splits = set()
    for i in range(k):
        d = desc[n + i]
        size = bin(d).count("1")
        if size <= 1 or size >= n - 1:
            continue
        comp = full ^ d
        if size < n - size:
            splits.add(d)
        elif size > n - size:
            splits.add(comp)
        else:
            splits.add(d if d & 1 else comp)
    return splits


def robinson_foulds(v1, v2, normalize: bool = False) -> float:
    """vector/distance.rs:41-74 (PyO3 robinson_foulds, py-phylo2vec/src/lib.rs:295-309: a matrix is
    reduced to its first column)."""
    a, b = np.asarray(v1), np.asarray(v2)
    if a.ndim == 2:
        a = a[:, 0].astype(np.int64)
    if b.ndim == 2:
        b = b[:, 0].astype(np.int64)
    if a.shape[0] != b.shape[0]:
        raise AssertionError("Trees must have the same number of leaves")
    if a.shape[0] + 1 <= 3:
        return 0.0
    s1, s2 = _bipartitions(a), _bipartitions(b)
    diff = len(s1 - s2) + len(s2 - s1)
    if normalize:
        m = len(s1) + len(s2)
        return 0.0 if m == 0 else diff / m
    return float(diff)


# ---- matrices <-> Newick with branch lengths (matrix/convert.rs, newick/mod.rs:161-255) ---------------
def ryu_format(x: float) -> str:
    """The Rust `ryu` crate's Buffer::format (what matrix/convert.rs:37-57 writes): the shortest round-trip
    digits -- the same digits Python's repr finds -- laid out by ryu's rules: plain decimal for
    1e-5 <= |x| < 1e16, otherwise d[.ddd]e[-]x without '+' or padding; a whole number ends in ".0"."""
    import math
    from decimal import Decimal
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "-inf" if x < 0 else "inf"
    sign = "-" if math.copysign(1.0, x) < 0 else ""
    if x == 0:
        return sign + "0.0"
    t = Decimal(repr(abs(x))).as_tuple()
    digits = "".join(map(str, t.digits)).lstrip("0")
    k = t.exponent
    stripped = digits.rstrip("0")
    k += len(digits) - len(stripped)
    digits = stripped
    length = len(digits)
    kk = length + k
    if 0 <= k and kk <= 16:
        return sign + digits + "0" * k + ".0"
    if 0 < kk <= 16:
        return sign + digits[:kk] + "." + digits[kk:]
    if -5 < kk <= 0:
        return sign + "0." + "0" * (-kk) + digits
    e = kk - 1
    return sign + digits[0] + ("." + digits[1:] if length > 1 else "") + "e" + str(e)


def to_newick_matrix(m) -> str:
    """matrix/convert.rs:26-57: check_m, parse_matrix, to_pairs, build_newick_with_bls."""
    m = np.asarray(m, dtype=np.float64)
    check_m(m)
    v = m[:, 0].astype(np.int64)
    pairs = to_pairs(v)
    k = pairs.shape[0]
    n = k + 1
    cache = [str(i) for i in range(n)]
    for c1, _c2 in pairs.tolist():               # prepare_cache, vector/convert.rs:341-359
        cache[c1] = "(" + cache[c1]
    for i, ((c1, c2), (bl1, bl2)) in enumerate(zip(pairs.tolist(), m[:, 1:3].tolist())):
        s2 = cache[c2]
        cache[c2] = ""
        cache[c1] = cache[c1] + ":" + ryu_format(bl1) + "," + s2 + ":" + ryu_format(bl2) + ")" + str(n + i)
    return cache[0] + ";"


def _node_substr(s: str, start: int):
    """newick/mod.rs:23-38"""
    end = start
    for i in range(start, len(s)):
        if s[i] in ",);":
            end = i
            break
    return s[start:end], end


def parse_with_bls(newick: str):
    """newick/mod.rs:161-255 (floats by Python's float(): correctly rounded, like Rust's parse::<f64>)."""
    anc, bls, stack, bl_stack = [], [], [], []
    i = 0
    while i < len(newick):
        c = newick[i]
        if c == ")":
            c2 = stack.pop()
            c1 = stack.pop()
            bl2 = bl_stack.pop()
            bl1 = bl_stack.pop()
            bls.append([bl1, bl2])
            annotation, end = _node_substr(newick, i + 1)
            i = end - 1
            if ":" in annotation:
                p, blp = annotation.split(":", 1)
                if p == "":
                    anc.append([c1, c2, max(c1, c2)])
                    stack.append(min(c1, c2))
                else:
                    anc.append([c1, c2, int(p)])
                    stack.append(int(p))
                bl_stack.append(float(blp))
            else:
                if end == len(newick) - 1:
                    anc.append([c1, c2, max(c1, c2)] if annotation == "" else [c1, c2, int(annotation)])
                    break
                raise ValueError("Missing annotation in the Newick string")
        elif c.isdigit():
            annotated, end = _node_substr(newick, i)
            i = end - 1
            node, bln = annotated.split(":", 1)
            stack.append(int(node))
            bl_stack.append(float(bln))
        i += 1
    return anc, bls


def from_newick_matrix(newick: str) -> np.ndarray:
    """matrix/convert.rs:84-122 with order_cherries / build_cherries (vector/convert.rs:398-477) and
    order_cherries_no_parents (:534-573) including the rows whose two lengths swap."""
    import re
    if newick == "":
        return np.zeros((0, 3))
    cherries, bls = parse_with_bls(newick)
    k = len(cherries)
    has_parents = re.search(r"\)(\d+)", newick) is not None          # newick/mod.rs NewickPatterns.parents
    swaps = []
    if has_parents:
        offset = max(c[2] for c in cherries) - 2 * k + 1
        row_idxs = list(range(k))
        new = [None] * k
        for i, ch in enumerate(cherries):
            idx = ch[2] - k - offset
            new[idx] = list(ch)
            row_idxs[idx] = i
        min_desc = {}
        for i, (c1, c2, p) in enumerate(new):
            m1 = min_desc.get(c1, c1)
            m2 = min_desc.get(c2, c2)
            if m1 < m2:
                lo, hi = m1, m2
            else:
                swaps.append(row_idxs[i])
                lo, hi = m2, m1
            min_desc[p] = lo
            new[i] = [m1, m2, hi]
        cherries = new
    else:
        to_sort, visited = [], {}
        for c1, c2, cmax in cherries:
            cmin = min(c1, c2)
            sister = visited[cmin] if (cmin in visited and visited[cmin] < cmax) else cmax
            to_sort.append(sister)
            visited[cmin] = sister
        row_idxs = sorted(range(k), key=lambda i: -to_sort[i])        # stable, descending
        for i in row_idxs:
            if cherries[i][0] > cherries[i][1]:
                swaps.append(i)
        cherries = [cherries[i] for i in row_idxs]
    v = build_vector(np.array(cherries, dtype=np.int64))
    for i in swaps:
        bls[i] = [bls[i][1], bls[i][0]]
    out = np.zeros((k, 3))
    for i in range(k):
        out[i, 0] = float(v[i])
        out[i, 1], out[i, 2] = bls[row_idxs[i]]
    return out
