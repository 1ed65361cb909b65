"""Host-side logic of the multi-GPU path on the CPU: world_size-2 gloo processes.

The kernels need a GPU, so the row blocks here come from the oracle; what is tested is
the partitioning, that shards reassemble the reference matrix, and the optional
all-gather of row blocks (phylo2vec_b200/distributed.py)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from phylo2vec_b200 import distributed as D
from tests.helpers import sample_bls, sample_vectors, to_matrix_input


def test_shard_bounds_partition():
    for total in (0, 1, 7, 64, 1000, 65536, 1_000_000):
        for world in (1, 2, 3, 4, 8):
            cuts = D.all_shards(total, world)
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_leaves, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import p2v_oracle as O
        v = sample_vectors(1, n_leaves, seed=11)[0]
        bls = sample_bls(1, n_leaves, seed=11)[0]
        m = to_matrix_input(v, bls)
        full = O.cophenetic_distances(m)
        # row-block sharding of one tree + optional all-gather
        r0, r1 = D.shard_rows(n_leaves, rank, world)
        local = torch.from_numpy(O.cophenetic_rows(m, False, r0, r1))
        gathered = D.all_gather_rows(local, n_leaves)
        ok_rows = np.allclose(gathered.numpy(), full, rtol=1e-13, atol=0)
        # the in-place form: the rank writes its block into its slice of the (n,n) buffer, nothing is staged
        buf = torch.full((n_leaves, n_leaves), -1.0, dtype=torch.float64)
        buf[r0:r1] = local
        same = D.all_gather_rows_(buf, n_leaves)
        ok_rows = ok_rows and same.data_ptr() == buf.data_ptr() and np.array_equal(buf.numpy(), gathered.numpy())
        ok_rows = ok_rows and D.all_gather_rows(buf[r0:r1], n_leaves, out=buf).data_ptr() == buf.data_ptr()
        # batch sharding: shards are disjoint and cover the batch
        B = 37
        V = sample_vectors(B, 50, seed=3)
        b0, b1 = D.shard_batch(B, rank, world)
        anc_local, _ = O.batch("to_ancestry", V[b0:b1], 49)
        counts = torch.tensor([b1 - b0], dtype=torch.int64)
        dist.all_reduce(counts)
        chk = torch.tensor([int(anc_local.sum())], dtype=torch.int64)
        dist.all_reduce(chk)
        ref_all, _ = O.batch("to_ancestry", V, 49)
        ok_batch = int(counts) == B and int(chk) == int(ref_all.sum())
        q.put((rank, bool(ok_rows), bool(ok_batch), (r0, r1)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_leaves", [33, 200])
def test_gloo_world2_row_blocks_and_batch_shards(n_leaves):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_leaves, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort()
    assert [r[0] for r in results] == [0, 1]
    assert all(r[1] and r[2] for r in results)
    assert results[0][3][1] == results[1][3][0] and results[1][3][1] == n_leaves
