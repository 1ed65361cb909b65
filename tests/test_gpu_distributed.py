"""Multi-GPU path on real devices (needs >= 2 GPUs; skipped on a 1-GPU box)."""

import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import phylo2vec_b200 as p2v
        from oracle import p2v_oracle as O
        from phylo2vec_b200 import distributed as D
        from tests.helpers import sample_bls, sample_vectors, to_matrix_input
        n = 3000
        v = sample_vectors(1, n, seed=5)
        bls = sample_bls(1, n, seed=5)
        M = torch.from_numpy(to_matrix_input(v, bls)).cuda()
        rows, (r0, r1) = D.cophenetic_rows_sharded(M)
        ref = O.cophenetic_rows(to_matrix_input(v, bls)[0], False, r0, r1)
        err = np.max(np.abs(rows.cpu().numpy() - ref) / np.maximum(ref, 1e-300))
        full = D.all_gather_rows(rows, n)
        sym = float((full - full.T).abs().max() / full.abs().max())
        V = torch.from_numpy(sample_vectors(64, 500, seed=rank)).cuda()
        anc = D.to_ancestry_sharded(V)
        ok = np.array_equal(anc.cpu().numpy(), O.batch("to_ancestry", V.cpu().numpy(), 499)[0])
        q.put((rank, float(err), sym, bool(ok), tuple(full.shape)))
    finally:
        dist.destroy_process_group()


def test_two_gpu_row_blocks_and_gather():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, err, sym, ok, shape in res:
        assert err <= 1e-12 and sym <= 1e-15 and ok and shape == (3000, 3000)
