"""Batched API: torch CUDA tensors (or any DLPack producer) in, torch tensors out -- or numpy
arrays in, numpy arrays out through the host entry points of the C ABI.

These are additions next to the reference's single-tree functions
(SURVEY section 8b): ``(B,k)`` vectors -> ``(B,k,3)`` ancestry etc.  Tensors
stay on the device; hand-off is zero-copy (DLPack), the kernels run on torch's
current stream through the C ABI in include/phylo2vec_b200.h.  torch is plumbing
here (device memory, streams); there is no eager/PyTorch implementation of the
path and no CPU fallback.
"""

from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _capi
from ._host import ptr as _np_ptr, raise_status as _np_raise_status

_DECODE = {"to_pairs": 2, "to_ancestry": 3, "to_edges": 4}
_ENCODE = {"from_pairs": 2, "from_ancestry": 3, "from_edges": 4}


def as_tensor(x) -> torch.Tensor:
    """torch tensor from a tensor or any object exposing ``__dlpack__`` (zero copy)."""
    if isinstance(x, torch.Tensor):
        return x
    if hasattr(x, "__dlpack__"):
        return torch.from_dlpack(x)
    raise TypeError("expected a torch.Tensor or a DLPack-capable object on a CUDA device")


def _idx_tensor(x, ndim: int, what: str) -> torch.Tensor:
    t = as_tensor(x)
    if not t.is_cuda:
        raise ValueError(f"{what}: tensor must live on a CUDA device (use the *_host functions for host data)")
    if t.dtype not in (torch.int64, torch.int32):
        raise TypeError(f"{what}: dtype must be int64 or int32, got {t.dtype}")
    if t.dim() != ndim:
        raise ValueError(f"{what}: expected {ndim} dimensions, got {t.dim()}")
    return t.contiguous()


def _check_status(status, B: int, device, what: str) -> torch.Tensor:
    """A caller-supplied status tensor is written by the kernels: int32, B entries, same device."""
    if status is None:
        return torch.empty(B, dtype=torch.int32, device=device)
    status = as_tensor(status)
    if status.dtype != torch.int32 or status.numel() != B or status.device != device or not status.is_contiguous():
        raise ValueError(f"{what}: status must be a contiguous int32 tensor of {B} entries on {device}")
    return status


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _ws(nbytes: int, device) -> Tuple[Optional[torch.Tensor], int]:
    if nbytes == 0:
        return None, 0
    t = torch.empty(nbytes, dtype=torch.uint8, device=device)
    return t, t.data_ptr()


def raise_for_status(status: torch.Tensor, v: Optional[torch.Tensor] = None) -> None:
    """Raise like the reference for the first failed tree of a batch.

    Invalid vectors: AssertionError with check_max's message
    (phylo2vec/src/vector/base.rs:60-67)."""
    if not bool(status.any()):          # the usual case: one reduction and one word back, nothing allocated
        return
    bad = torch.nonzero(status)
    b = int(bad[0, 0])
    code = int(status[b])
    if code > 0:
        i = code - 1
        val = int(v[b, i]) if v is not None else "?"
        raise AssertionError(
            f"Validation failed: v[{i}] = {val} is out of bounds (max = {2 * i})" +
            (f" [tree {b} of the batch]" if status.numel() > 1 else ""))
    if code == _capi.P2V_TREE_NEGATIVE_BRANCH:      # check_m, phylo2vec/src/matrix/base.rs:88-94
        raise AssertionError("Branch lengths must be positive" +
                             (f" [tree {b} of the batch]" if status.numel() > 1 else ""))
    raise ValueError(f"malformed tree input (tree {b} of the batch, code {code})")


def _host_call(name: str, X: np.ndarray, out_shape, k: int, check: bool, v_for_msg=None) -> np.ndarray:
    """numpy in, numpy out: the *_host entry point (H2D, kernels, D2H inside the call)."""
    if X.dtype not in (np.int64, np.int32):
        raise TypeError(f"{name}: dtype must be int64 or int32, got {X.dtype}")
    X = np.ascontiguousarray(X)
    B = X.shape[0]
    out = np.empty(out_shape, dtype=X.dtype)
    status = np.zeros(max(B, 1), dtype=np.int32)
    rc = getattr(_capi.load(), f"p2v_{name}_host")(_np_ptr(X), X.itemsize, B, k, _np_ptr(out), _np_ptr(status))
    _capi.check(rc, name)
    if check:
        _np_raise_status(status[:B], v_for_msg)
    return out


def _decode(name: str, V, out=None, status=None, check: bool = True):
    cols = _DECODE[name]
    if isinstance(V, np.ndarray):           # host data: numpy in, numpy out
        if V.ndim != 2:
            raise ValueError(f"{name}: expected 2 dimensions, got {V.ndim}")
        B, k = V.shape
        shape = {"to_pairs": (B, k, 2), "to_ancestry": (B, k, 3), "to_edges": (B, 2 * k, 2)}[name]
        return _host_call(name, V, shape, k, check, V)
    V = _idx_tensor(V, 2, name)
    B, k = V.shape
    shape = {"to_pairs": (B, k, 2), "to_ancestry": (B, k, 3), "to_edges": (B, 2 * k, 2)}[name]
    if out is None:
        out = torch.empty(shape, dtype=V.dtype, device=V.device)
    else:
        out = as_tensor(out)
        if tuple(out.shape) != shape or out.dtype != V.dtype or not out.is_contiguous() or out.device != V.device:
            raise ValueError(f"{name}: out must be a contiguous {shape} {V.dtype} tensor on {V.device}")
        if name != "to_ancestry" and out.data_ptr() % (2 * out.element_size()) != 0:
            # pairs / edges are written as two-element vector stores (include/phylo2vec_b200.h, "Alignment")
            raise ValueError(f"{name}: out must be aligned to {2 * out.element_size()} bytes "
                             "(a view at an odd element offset is not)")
    status = _check_status(status, B, V.device, name)
    lib = _capi.load()
    with torch.cuda.device(V.device):
        ws_t, ws_p = _ws(lib.p2v_decode_workspace_bytes(B, k), V.device)
        rc = getattr(lib, f"p2v_{name}_batch")(V.data_ptr(), V.element_size(), B, k, out.data_ptr(),
                                              status.data_ptr(), ws_p, 0 if ws_t is None else ws_t.numel(),
                                              _stream_ptr(V.device))
    _capi.check(rc, name)
    del cols
    if check:
        raise_for_status(status, V)
    return out


def to_pairs_batch(V, out=None, status=None, check=True):
    """(B,k) int vectors -> (B,k,2) pairs (branch, leaf).  Batched to_pairs
    (phylo2vec/src/vector/convert.rs:103-109)."""
    return _decode("to_pairs", V, out, status, check)


def to_ancestry_batch(V, out=None, status=None, check=True):
    """(B,k) -> (B,k,3) rows [child1, child2, parent].  Batched to_ancestry
    (phylo2vec/src/vector/convert.rs:167-188)."""
    return _decode("to_ancestry", V, out, status, check)


def to_edges_batch(V, out=None, status=None, check=True):
    """(B,k) -> (B,2k,2) edges (child, parent).  Batched to_edges
    (phylo2vec/src/vector/convert.rs:227-248)."""
    return _decode("to_edges", V, out, status, check)


def _encode(name: str, X, out=None, status=None, check: bool = True):
    cols = _ENCODE[name]
    if isinstance(X, np.ndarray):           # host data: numpy in, numpy out
        if X.ndim != 3:
            raise ValueError(f"{name}: expected 3 dimensions, got {X.ndim}")
        if name == "from_edges" and X.shape[1] % 2 != 0:
            raise AssertionError("The number of edges must be even")
        k = X.shape[1] // 2 if name == "from_edges" else X.shape[1]
        if X.shape[2] != (2 if name == "from_edges" else cols):
            raise ValueError(f"{name}: bad trailing dimension {tuple(X.shape)}")
        return _host_call(name, X, (X.shape[0], k), k, check)
    X = _idx_tensor(X, 3, name)
    B = X.shape[0]
    if name == "from_edges":
        if X.shape[1] % 2 != 0:  # phylo2vec/src/vector/convert.rs:202
            raise AssertionError("The number of edges must be even")
        k = X.shape[1] // 2
        ok = X.shape[2] == 2
    else:
        k = X.shape[1]
        ok = X.shape[2] == cols
    if not ok:
        raise ValueError(f"{name}: bad trailing dimension {tuple(X.shape)}")
    if out is None:
        out = torch.empty((B, k), dtype=X.dtype, device=X.device)
    else:
        out = as_tensor(out)
        if tuple(out.shape) != (B, k) or out.dtype != X.dtype or not out.is_contiguous() or out.device != X.device:
            raise ValueError(f"{name}: out must be a contiguous {(B, k)} {X.dtype} tensor on {X.device}")
    status = _check_status(status, B, X.device, name)
    lib = _capi.load()
    with torch.cuda.device(X.device):
        ws_t, ws_p = _ws(lib.p2v_encode_workspace_bytes(B, k), X.device)
        rc = getattr(lib, f"p2v_{name}_batch")(X.data_ptr(), X.element_size(), B, k, out.data_ptr(),
                                              status.data_ptr(), ws_p, 0 if ws_t is None else ws_t.numel(),
                                              _stream_ptr(X.device))
    _capi.check(rc, name)
    if check:
        raise_for_status(status)
    return out


def from_pairs_batch(P, out=None, status=None, check=True):
    """(B,k,2) -> (B,k).  Batched from_pairs (phylo2vec/src/vector/convert.rs:23-30)."""
    return _encode("from_pairs", P, out, status, check)


def from_ancestry_batch(A, out=None, status=None, check=True):
    """(B,k,3) -> (B,k).  Batched from_ancestry (phylo2vec/src/vector/convert.rs:129-136)."""
    return _encode("from_ancestry", A, out, status, check)


def from_edges_batch(E, out=None, status=None, check=True):
    """(B,2k,2) -> (B,k).  Batched from_edges (phylo2vec/src/vector/convert.rs:201-213)."""
    return _encode("from_edges", E, out, status, check)


def check_v_batch(V) -> torch.Tensor:
    """Per-tree status of check_v (phylo2vec/src/vector/base.rs:50-67): 0 or bad index + 1."""
    V = _idx_tensor(V, 2, "check_v")
    B, k = V.shape
    status = torch.empty(B, dtype=torch.int32, device=V.device)
    with torch.cuda.device(V.device):
        rc = _capi.load().p2v_check_v_batch(V.data_ptr(), V.element_size(), B, k, status.data_ptr(),
                                            _stream_ptr(V.device))
    _capi.check(rc, "check_v")
    return status


def check_m_batch(M) -> torch.Tensor:
    """Per-tree status of check_m (phylo2vec/src/matrix/base.rs:76-95) for float64 ``(B,k,3)`` matrices:
    0, bad index + 1 (column 0 fails check_v) or ``_capi.P2V_TREE_NEGATIVE_BRANCH``."""
    M = as_tensor(M)
    if not M.is_cuda or M.dtype != torch.float64 or M.dim() != 3 or M.shape[2] != 3:
        raise TypeError("check_m: expected a float64 CUDA tensor (B,k,3)")
    M = M.contiguous()
    B, k = M.shape[0], M.shape[1]
    status = torch.empty(B, dtype=torch.int32, device=M.device)
    with torch.cuda.device(M.device):
        rc = _capi.load().p2v_check_m_batch(M.data_ptr(), B, k, status.data_ptr(), _stream_ptr(M.device))
    _capi.check(rc, "check_m")
    return status


def cophenetic_workspace_bytes(B: int, k: int) -> int:
    return int(_capi.load().p2v_cophenetic_workspace_bytes(B, k))


def _tree_batch(trees, bls, what: str):
    """PyTree dispatch of py-phylo2vec/src/lib.rs:20-24 lifted to batches."""
    T = as_tensor(trees)
    if not T.is_cuda:
        raise ValueError(f"{what}: tensor must live on a CUDA device")
    if T.dim() == 3:
        if T.dtype != torch.float64 or T.shape[2] != 3:
            raise TypeError("matrix input must be float64 with shape (B,k,3)")
        if bls is not None:
            raise ValueError("bls given together with matrix input")
        return T.contiguous(), None, True
    if T.dim() == 2:
        T = _idx_tensor(T, 2, what)
        B, k = T.shape
        if bls is not None:
            bls = as_tensor(bls)
            if bls.dtype != torch.float64 or tuple(bls.shape) != (B, k, 2) or bls.device != T.device:
                raise TypeError("bls must be float64 (B,k,2) on the same device")
            bls = bls.contiguous()
        return T, bls, False
    raise TypeError("trees must be (B,k) integer vectors or (B,k,3) float64 matrices")


def _out_f64(out, shape, device):
    if out is None:
        return torch.empty(shape, dtype=torch.float64, device=device)
    out = as_tensor(out)
    if tuple(out.shape) != shape or out.dtype != torch.float64 or not out.is_contiguous() or out.device != device:
        raise ValueError(f"out must be a contiguous float64 {shape} tensor on {device}")
    return out


def _rows_call(what, trees, unrooted, rows, out, bls, status, check, workspace):
    T, bls, is_matrix = _tree_batch(trees, bls, what)
    lib = _capi.load()
    B, k = T.shape[0], T.shape[1]
    n = k + 1
    if what == "vcv" and k == 0:    # vector/graph.rs:214-217
        raise AssertionError("Variance-covariance matrix not supported for trees with < 2 leaves.")
    r0, r1 = (0, n) if rows is None else (int(rows[0]), int(rows[1]))
    if not (0 <= r0 <= r1 <= n):
        raise ValueError(f"rows {rows} outside [0, {n}]")
    out = _out_f64(out, (B, r1 - r0, n), T.device)
    status = _check_status(status, B, T.device, what)
    need = lib.p2v_cophenetic_workspace_bytes(B, k)
    with torch.cuda.device(T.device):
        if workspace is None or workspace.numel() * workspace.element_size() < need:
            workspace = torch.empty(max(need, 16), dtype=torch.uint8, device=T.device)
        wsb = workspace.numel() * workspace.element_size()
        tail = (r0, r1, out.data_ptr(), status.data_ptr(), workspace.data_ptr(), wsb, _stream_ptr(T.device))
        blp = 0 if bls is None else bls.data_ptr()
        if what == "vcv":
            if is_matrix:
                rc = lib.p2v_vcv_matrix_batch(T.data_ptr(), B, k, *tail)
            else:
                rc = lib.p2v_vcv_batch(T.data_ptr(), T.element_size(), blp, B, k, *tail)
        else:
            if is_matrix:
                rc = lib.p2v_cophenetic_matrix_batch(T.data_ptr(), B, k, int(bool(unrooted)), *tail)
            else:
                rc = lib.p2v_cophenetic_batch(T.data_ptr(), T.element_size(), blp, B, k, int(bool(unrooted)), *tail)
    _capi.check(rc, what)
    if check and bool(status.any()):
        raise_for_status(status, T[..., 0].to(torch.int64) if is_matrix else T)
    return out


def cophenetic_distances_batch(trees, unrooted: bool = False, rows: Optional[Tuple[int, int]] = None,
                               out=None, bls=None, status=None, check: bool = True,
                               workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Distance matrices of a batch of trees, float64 ``(B, rows, n)``.

    trees: int ``(B,k)`` vectors (topological unless ``bls`` ``(B,k,2)`` float64 is
    given) or float64 ``(B,k,3)`` matrices ``[v, bl1, bl2]`` -- the PyTree dispatch of
    py-phylo2vec/src/lib.rs:20-24,144-156 lifted to batches.  ``rows=(r0,r1)`` selects a
    row block of every matrix (how one huge tree is sharded over GPUs).
    Replaces vector::graph::_cophenetic_distances (phylo2vec/src/vector/graph.rs:20-90)."""
    return _rows_call("cophenetic_distances", trees, unrooted, rows, out, bls, status, check, workspace)


def vcv_batch(trees, rows: Optional[Tuple[int, int]] = None, out=None, bls=None, status=None,
              check: bool = True, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Variance-covariance matrices of a batch of trees, float64 ``(B, rows, n)``: entry (l, r) is the
    depth of the node where leaves l and r meet, the diagonal the depth of the leaf.  Same input
    conventions as :func:`cophenetic_distances_batch`.
    Replaces vector::graph::_vcv (phylo2vec/src/vector/graph.rs:211-282, matrix/graph.rs:61-64)."""
    return _rows_call("vcv", trees, False, rows, out, bls, status, check, workspace)


def node_depths_batch(trees, out=None, bls=None, status=None, check: bool = True,
                      workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Depth (distance from the root) of every node of every tree, float64 ``(B, 2k+1)``.
    Replaces vector::ops::_get_node_depths (phylo2vec/src/vector/ops.rs:201-233, matrix/ops.rs:42-45)."""
    T, bls, is_matrix = _tree_batch(trees, bls, "get_node_depths")
    lib = _capi.load()
    B, k = T.shape[0], T.shape[1]
    out = _out_f64(out, (B, 2 * k + 1), T.device)
    status = _check_status(status, B, T.device, "get_node_depths")
    need = lib.p2v_cophenetic_workspace_bytes(B, k)
    with torch.cuda.device(T.device):
        if workspace is None or workspace.numel() * workspace.element_size() < need:
            workspace = torch.empty(max(need, 16), dtype=torch.uint8, device=T.device)
        wsb = workspace.numel() * workspace.element_size()
        rc = lib.p2v_node_depths_batch(T.data_ptr(), 0 if is_matrix else T.element_size(),
                                       0 if bls is None else bls.data_ptr(), B, k, out.data_ptr(),
                                       status.data_ptr(), workspace.data_ptr(), wsb, _stream_ptr(T.device))
    _capi.check(rc, "get_node_depths")
    if check and bool(status.any()):
        raise_for_status(status, T[..., 0].to(torch.int64) if is_matrix else T)
    return out


def to_newick_batch(V, status=None, check: bool = True, raw: bool = False):
    """Newick strings of a batch of vectors ``(B,k)`` or of matrices ``(B,k,3)`` float64 (branch lengths).  Returns
    a list of ``str``; with ``raw=True`` the device buffers instead: a uint8 tensor ``(B, stride)`` of
    NUL-terminated strings and an int64 tensor of their lengths.
    Replaces vector::convert::to_newick (phylo2vec/src/vector/convert.rs:341-397) and matrix::convert::to_newick
    (phylo2vec/src/matrix/convert.rs:26-57)."""
    T = as_tensor(V)
    lib = _capi.load()
    if T.dim() == 3:
        if not T.is_cuda or T.dtype != torch.float64 or T.shape[2] != 3:
            raise TypeError("to_newick: matrices must be a float64 CUDA tensor (B,k,3)")
        M = T.contiguous()
        B, k = M.shape[0], M.shape[1]
        if k == 0:
            raise AssertionError("Matrix should not be empty")
        stride = (int(lib.p2v_newick_matrix_max_bytes(k)) + 15) & ~15
        out = torch.empty((B, stride), dtype=torch.uint8, device=M.device)
        lengths = torch.empty(B, dtype=torch.int64, device=M.device)
        status = _check_status(status, B, M.device, "to_newick")
        need = int(lib.p2v_newick_matrix_workspace_bytes(B, k))
        with torch.cuda.device(M.device):
            ws = torch.empty(max(need, 16), dtype=torch.uint8, device=M.device)
            rc = lib.p2v_to_newick_matrix_batch(M.data_ptr(), B, k, out.data_ptr(), stride, lengths.data_ptr(),
                                                status.data_ptr(), ws.data_ptr(), need, _stream_ptr(M.device))
        _capi.check(rc, "to_newick")
        if check and bool(status.any()):
            raise_for_status(status, M[..., 0].to(torch.int64))
        V = M
    else:
        V = _idx_tensor(V, 2, "to_newick")
        B, k = V.shape
        stride = (int(lib.p2v_newick_max_bytes(k)) + 15) & ~15
        out = torch.empty((B, stride), dtype=torch.uint8, device=V.device)
        lengths = torch.empty(B, dtype=torch.int64, device=V.device)
        status = _check_status(status, B, V.device, "to_newick")
        need = int(lib.p2v_newick_workspace_bytes(B, k))
        with torch.cuda.device(V.device):
            ws = torch.empty(max(need, 16), dtype=torch.uint8, device=V.device)
            rc = lib.p2v_to_newick_batch(V.data_ptr(), V.element_size(), B, k, out.data_ptr(), stride,
                                         lengths.data_ptr(), status.data_ptr(), ws.data_ptr(), need,
                                         _stream_ptr(V.device))
        _capi.check(rc, "to_newick")
        if check and bool(status.any()):
            raise_for_status(status, V)
    if raw:
        return out, lengths
    ln = lengths.cpu().numpy()
    width = int(ln.max()) if B else 0
    host = out[:, :max(width, 1)].cpu().numpy()
    return [host[b, :ln[b]].tobytes().decode("ascii") for b in range(B)]


def to_matrix_batch(newick, lengths=None, n_leaves: Optional[int] = None, status=None, check: bool = True,
                    device=None) -> torch.Tensor:
    """Matrices ``(B,k,3)`` float64 ``[v, bl1, bl2]`` of a batch of Newick strings with branch lengths (with or
    without parent labels); every tree has the same number of leaves.  ``newick`` is a list of ``str`` or the raw
    form :func:`to_newick_batch` returns.  Replaces matrix::convert::from_newick
    (phylo2vec/src/matrix/convert.rs:84-122; PyO3 to_matrix, py-phylo2vec/src/lib.rs:103-106)."""
    newick, lengths, n_leaves = _newick_buffers(newick, lengths, n_leaves, device, "to_matrix")
    B, stride = newick.shape
    k = int(n_leaves) - 1
    lib = _capi.load()
    m = torch.empty((B, k, 3), dtype=torch.float64, device=newick.device)
    status = _check_status(status, B, newick.device, "to_matrix")
    need = int(lib.p2v_from_newick_matrix_workspace_bytes(B, k))
    with torch.cuda.device(newick.device):
        ws = torch.empty(max(need, 16), dtype=torch.uint8, device=newick.device)
        rc = lib.p2v_from_newick_matrix_batch(newick.data_ptr(), stride, 0 if lengths is None else lengths.data_ptr(), B, k,
                                              m.data_ptr(), status.data_ptr(), ws.data_ptr(), need,
                                              _stream_ptr(newick.device))
    _capi.check(rc, "to_matrix")
    if check:
        st = status.cpu().numpy()
        if st.any():
            raise ValueError(f"malformed Newick string (tree {int(np.flatnonzero(st)[0])})")
    return m


def _newick_buffers(newick, lengths, n_leaves, device, what):
    """list of str -> (uint8 (B, stride) CUDA tensor, int64 lengths, n_leaves); raw buffers pass through."""
    if isinstance(newick, (list, tuple)):
        if len(newick) == 0:
            raise ValueError(f"{what}: empty batch")
        enc = [s.encode("ascii") for s in newick]
        if n_leaves is None:
            n_leaves = enc[0].count(b",") + 1
        stride = (max(len(e) for e in enc) + 1 + 15) & ~15
        host = np.zeros((len(enc), stride), dtype=np.uint8)
        for b, e in enumerate(enc):
            host[b, :len(e)] = np.frombuffer(e, dtype=np.uint8)
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        newick = torch.from_numpy(host).to(dev)
        lengths = torch.tensor([len(e) for e in enc], dtype=torch.int64, device=dev)
    if not (isinstance(newick, torch.Tensor) and newick.dtype == torch.uint8 and newick.dim() == 2 and newick.is_cuda):
        raise TypeError(f"{what}: expected a list of str or a uint8 CUDA tensor (B, stride)")
    if n_leaves is None:
        raise ValueError(f"{what}: n_leaves is required with a raw buffer")
    newick = newick.contiguous()
    if lengths is not None:
        lengths = lengths.to(device=newick.device, dtype=torch.int64).contiguous()
    return newick, lengths, n_leaves


def from_newick_batch(newick, lengths=None, n_leaves: Optional[int] = None, dtype=torch.int64, status=None,
                      check: bool = True, device=None) -> torch.Tensor:
    """Vectors ``(B,k)`` of a batch of Newick strings (with or without parent labels); every tree has the same
    number of leaves.  ``newick`` is a list of ``str`` or the raw form :func:`to_newick_batch` returns
    (uint8 tensor ``(B, stride)`` on the GPU, with ``lengths`` or NUL-terminated rows).
    Replaces newick::parse + vector::convert::from_newick (phylo2vec/src/newick/mod.rs:73-127,
    vector/convert.rs:268-278, 534-573)."""
    if isinstance(newick, (list, tuple)):
        if len(newick) == 0:
            raise ValueError("from_newick: empty batch")
        enc = [s.encode("ascii") for s in newick]
        if n_leaves is None:
            n_leaves = enc[0].count(b",") + 1
        stride = (max(len(e) for e in enc) + 1 + 15) & ~15
        host = np.zeros((len(enc), stride), dtype=np.uint8)
        for b, e in enumerate(enc):
            host[b, :len(e)] = np.frombuffer(e, dtype=np.uint8)
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        newick = torch.from_numpy(host).to(dev)
        lengths = torch.tensor([len(e) for e in enc], dtype=torch.int64, device=dev)
    if not (isinstance(newick, torch.Tensor) and newick.dtype == torch.uint8 and newick.dim() == 2 and newick.is_cuda):
        raise TypeError("from_newick: expected a list of str or a uint8 CUDA tensor (B, stride)")
    if n_leaves is None:
        raise ValueError("from_newick: n_leaves is required with a raw buffer")
    newick = newick.contiguous()
    B, stride = newick.shape
    k = int(n_leaves) - 1
    lib = _capi.load()
    idx_bytes = 8 if dtype == torch.int64 else 4
    if dtype not in (torch.int64, torch.int32):
        raise TypeError("from_newick: dtype must be torch.int64 or torch.int32")
    v = torch.empty((B, k), dtype=dtype, device=newick.device)
    if status is None:
        status = torch.empty(B, dtype=torch.int32, device=newick.device)
    if lengths is not None:
        lengths = lengths.to(device=newick.device, dtype=torch.int64).contiguous()
    need = int(lib.p2v_from_newick_workspace_bytes(B, k, idx_bytes))
    with torch.cuda.device(newick.device):
        ws = torch.empty(max(need, 16), dtype=torch.uint8, device=newick.device)
        rc = lib.p2v_from_newick_batch(newick.data_ptr(), stride, 0 if lengths is None else lengths.data_ptr(), B, k,
                                       v.data_ptr(), idx_bytes, status.data_ptr(), ws.data_ptr(), need,
                                       _stream_ptr(newick.device))
    _capi.check(rc, "from_newick")
    if check:
        st = status.cpu().numpy()
        if st.any():
            raise ValueError(f"malformed Newick string (tree {int(np.flatnonzero(st)[0])})")
    return v


def robinson_foulds_batch(V1, V2, normalize: bool = False, out=None, status=None, check: bool = True) -> torch.Tensor:
    """Robinson-Foulds distances of tree pairs, float64 ``(P,)``: ``V1`` ``(B1,k)`` against ``V2`` ``(B2,k)`` with
    ``B1 == B2`` (pair b = tree b of each batch) or one of them a single tree (against every tree of the other
    batch).  Replaces vector::distance::robinson_foulds (phylo2vec/src/vector/distance.rs:41-162)."""
    V1 = _idx_tensor(V1, 2, "robinson_foulds")
    V2 = _idx_tensor(V2, 2, "robinson_foulds")
    if V1.shape[1] != V2.shape[1]:
        raise AssertionError("Trees must have the same number of leaves")
    if V1.dtype != V2.dtype or V1.device != V2.device:
        raise TypeError("robinson_foulds: both batches must share dtype and device")
    B1, B2, k = V1.shape[0], V2.shape[0], V1.shape[1]
    if B1 != B2 and B1 != 1 and B2 != 1:
        raise ValueError("robinson_foulds: batches must have equal sizes, or one of them a single tree")
    P = max(B1, B2)
    lib = _capi.load()
    out = _out_f64(out, (P,), V1.device)
    status = _check_status(status, P, V1.device, "robinson_foulds")
    need = int(lib.p2v_robinson_foulds_workspace_bytes(B1, B2, k))
    with torch.cuda.device(V1.device):
        ws = torch.empty(max(need, 16), dtype=torch.uint8, device=V1.device)
        rc = lib.p2v_robinson_foulds_batch(V1.data_ptr(), V2.data_ptr(), V1.element_size(), B1, B2, k,
                                           int(bool(normalize)), out.data_ptr(), status.data_ptr(), ws.data_ptr(), need,
                                           _stream_ptr(V1.device))
    _capi.check(rc, "robinson_foulds")
    if check and bool(status.any()):
        bad = int(torch.nonzero(status)[0, 0])
        which = V1 if int(check_v_batch(V1[bad:bad + 1] if B1 > 1 else V1)[0]) != 0 else V2
        raise_for_status(status[bad:bad + 1], which[bad:bad + 1] if which.shape[0] > 1 else which)
    return out


def balance_indices_batch(V, out=None, status=None, check: bool = True) -> torch.Tensor:
    """Sackin index, leaf-depth variance and B2 of every vector of a batch ``(B,k)``: float64 ``(B,3)``.
    Replaces vector::balance::{sackin, leaf_depth_variance, b2} (phylo2vec/src/vector/balance.rs:14-58)."""
    V = _idx_tensor(V, 2, "balance")
    B, k = V.shape
    lib = _capi.load()
    out = _out_f64(out, (B, 3), V.device)
    if status is None:
        status = torch.empty(B, dtype=torch.int32, device=V.device)
    need = int(lib.p2v_balance_workspace_bytes(B, k))
    with torch.cuda.device(V.device):
        ws = torch.empty(max(need, 16), dtype=torch.uint8, device=V.device)
        rc = lib.p2v_balance_batch(V.data_ptr(), V.element_size(), B, k, out.data_ptr(), status.data_ptr(),
                                   ws.data_ptr(), need, _stream_ptr(V.device))
    _capi.check(rc, "balance")
    if check and bool(status.any()):
        raise_for_status(status, V)
    return out
