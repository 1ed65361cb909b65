"""Multi-GPU partitioning of the path (one process per GPU, torch.distributed).

Two shardings, both without any data-path collective:

* batches of trees: trees are independent, so rank r owns the contiguous batch
  slice ``shard_batch(B, r, W)`` and nothing is exchanged;
* one very large tree: rank r owns the row block ``shard_rows(n, r, W)`` of the
  n x n distance matrix.  Every rank decodes the (small) tree itself and builds
  its own tables -- cheaper than broadcasting them -- and writes its rows.

``all_gather_rows`` is the one optional collective: it assembles the full matrix
on every rank (NCCL over NVLink on GPUs; 34 GB at n = 65536, so it is off by
default and timed separately).
"""

from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of rank's share of `total` items: floor(total*r/W) cuts, so the
    shares are contiguous, disjoint, cover everything and differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    return total * rank // world, total * (rank + 1) // world


def shard_batch(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Trees [begin, end) of a batch of B owned by `rank`."""
    return shard_bounds(B, rank, world)


def shard_rows(n_leaves: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [begin, end) (leaf labels) of one tree's n x n matrix owned by `rank`."""
    return shard_bounds(n_leaves, rank, world)


def all_shards(total: int, world: int) -> List[Tuple[int, int]]:
    return [shard_bounds(total, r, world) for r in range(world)]


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def to_ancestry_sharded(V_global_or_local, local: bool = True, group=None, **kw) -> torch.Tensor:
    """Decode this rank's slice of a batch.  With ``local=True`` the tensor already is
    the rank's shard (the normal case: inputs are produced or loaded per rank); otherwise
    the rank slices its share out of the replicated batch.  No communication."""
    from . import batch
    rank, world = _world(group)
    V = V_global_or_local
    if not local:
        b0, b1 = shard_batch(V.shape[0], rank, world)
        V = V[b0:b1]
    return batch.to_ancestry_batch(V, **kw)


def cophenetic_rows_sharded(tree, unrooted: bool = False, bls=None, group=None,
                            out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Tuple[int, int]]:
    """This rank's row block of ONE tree's distance matrix.

    tree: (1,k) int vector (optionally with bls (1,k,2)) or (1,k,3) float64 matrix,
    replicated on every rank.  Returns ((rows, n) float64 tensor, (row_begin, row_end))."""
    from . import batch
    rank, world = _world(group)
    k = tree.shape[1]
    r0, r1 = shard_rows(k + 1, rank, world)
    D = batch.cophenetic_distances_batch(tree, unrooted=unrooted, rows=(r0, r1), bls=bls,
                                         out=None if out is None else out.view(1, r1 - r0, k + 1))
    return D[0], (r0, r1)


def all_gather_rows_(full: torch.Tensor, n_leaves: int, group=None) -> torch.Tensor:
    """Optional collective, IN PLACE: ``full`` is the preallocated (n,n) matrix of which this rank has
    written its own row block ``shard_rows(n, rank, world)`` (e.g. by passing ``out=full[r0:r1]`` to
    :func:`cophenetic_rows_sharded`); afterwards every rank holds every block.  Equal blocks
    (``n % world == 0``) are one ``all_gather_into_tensor`` whose input is the rank's own slice of the
    output -- ncclAllGather in place, ``sendcount = rows * n`` float64, no staging buffer (SURVEY
    section 8e).  Unequal blocks are broadcast block by block, also in place."""
    rank, world = _world(group)
    if world == 1:
        return full
    if tuple(full.shape) != (n_leaves, n_leaves) or not full.is_contiguous():
        raise ValueError(f"all_gather_rows_: expected a contiguous ({n_leaves},{n_leaves}) tensor")
    shards = all_shards(n_leaves, world)
    r0, r1 = shards[rank]
    if n_leaves % world == 0:
        dist.all_gather_into_tensor(full, full[r0:r1], group=group)
    else:
        for src, (a, b) in enumerate(shards):
            if b > a:
                dist.broadcast(full[a:b], src=dist.get_global_rank(group, src) if group is not None else src, group=group)
    return full


def all_gather_rows(local_rows: torch.Tensor, n_leaves: int, group=None,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Optional: assemble the full (n,n) matrix on every rank from the row blocks.  ``out`` is the
    (n,n) destination (allocated here if absent); the rank's block is copied into its slice unless
    ``local_rows`` already is that slice, then :func:`all_gather_rows_` gathers in place."""
    rank, world = _world(group)
    if world == 1:
        return local_rows
    r0, r1 = shard_rows(n_leaves, rank, world)
    if tuple(local_rows.shape) != (r1 - r0, n_leaves):
        raise ValueError(f"all_gather_rows: rank {rank} holds {tuple(local_rows.shape)}, expected {(r1 - r0, n_leaves)}")
    if out is None:
        out = torch.empty((n_leaves, n_leaves), dtype=local_rows.dtype, device=local_rows.device)
    mine = out[r0:r1]
    if mine.data_ptr() != local_rows.data_ptr():
        mine.copy_(local_rows)
    return all_gather_rows_(out, n_leaves, group)
