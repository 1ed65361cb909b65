#!/usr/bin/env python
"""Mixed-size sweep (BASELINE.json configs[0], [2], [4]): decode and distance throughput per
tree size on one GPU, next to the CPU oracle port on the host cores.

    python benchmarks/sweep.py [--out profiles/r01_sweep.json] [--quick]

Not the bench contract (that is bench.py); this fills the per-size table in profiles/.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import phylo2vec_b200 as p2v  # noqa: E402
from oracle import p2v_oracle as O  # noqa: E402
from phylo2vec_b200 import _capi  # noqa: E402

lib = _capi.load()
dev = torch.device("cuda:0")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def timeit(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


def trees(B, n, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    k = n - 1
    hi = (2 * torch.arange(k, device=dev) + 1).to(torch.float64)
    v = torch.empty((B, k), dtype=torch.int64, device=dev)
    step = max(1, (1 << 26) // max(k, 1))
    for b0 in range(0, B, step):
        b1 = min(B, b0 + step)
        v[b0:b1] = (torch.rand((b1 - b0, k), generator=g, device=dev, dtype=torch.float64) * hi).to(torch.int64)
    return v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_sweep.json"))
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    sizes = [16, 32, 64, 100, 128, 256, 512, 1000, 2000, 4096, 16384, 65536, 100000]
    if args.quick:
        sizes = [100, 1000, 2000, 16384]
    nthreads = host_threads()
    rows = []
    for n in sizes:
        k = n - 1
        rec = {"n_leaves": n}
        # ---- decode / encode: B*k ~ 2.5e8 elements, at most 4M trees -------------------------
        B = int(min(4_000_000, max(8, 250_000_000 // k)))
        V = trees(B, n)
        need = lib.p2v_decode_workspace_bytes(B, k)
        ws = torch.empty(max(need, 16), dtype=torch.uint8, device=dev)
        st = torch.empty(B, dtype=torch.int32, device=dev)
        anc = torch.empty((B, k, 3), dtype=torch.int64, device=dev)
        s = torch.cuda.current_stream().cuda_stream
        t = timeit(lambda: lib.p2v_to_ancestry_batch(V.data_ptr(), 8, B, k, anc.data_ptr(), st.data_ptr(), ws.data_ptr(), need, s))
        assert int(st.count_nonzero()) == 0
        rec["to_ancestry"] = {"batch": B, "trees_per_s": B / t, "gbs": 4 * k * 8 * B / t / 1e9,
                              "frac_of_measured_hbm": 4 * k * 8 * B / t / 1e9 / PEAK}
        need_e = lib.p2v_encode_workspace_bytes(B, k)
        wse = torch.empty(max(need_e, 16), dtype=torch.uint8, device=dev)
        back = torch.empty((B, k), dtype=torch.int64, device=dev)
        t = timeit(lambda: lib.p2v_from_ancestry_batch(anc.data_ptr(), 8, B, k, back.data_ptr(), st.data_ptr(), wse.data_ptr(), need_e, s))
        rec["from_ancestry"] = {"batch": B, "trees_per_s": B / t, "round_trip_equal": bool(torch.equal(back, V))}
        # CPU oracle port on a bounded sample (about a second of wall time)
        Bc = int(max(4, min(B, 30_000_000 // max(k * 12, 1))))
        Vc = V[:Bc].cpu().numpy()
        t0 = time.perf_counter()
        ref, stc = O.batch("to_ancestry", Vc, k, nthreads=nthreads)
        tc = time.perf_counter() - t0
        rec["to_ancestry_cpu"] = {"sample": Bc, "trees_per_s": Bc / tc, "cores": nthreads,
                                  "matches_gpu": bool(np.array_equal(ref, anc[:Bc].cpu().numpy()))}
        del anc, back, ws, wse
        # ---- distances: B * n^2 * 8 <= 8 GB -----------------------------------------------------
        Bd = int(max(1, min(B, (8 << 30) // (n * n * 8))))
        if n * n * 8 > (40 << 30):
            Bd = 0
        if Bd:
            Vd = V[:Bd].contiguous()
            g = torch.Generator(device=dev).manual_seed(1)
            bls = torch.rand((Bd, k, 2), generator=g, device=dev, dtype=torch.float64)
            needc = lib.p2v_cophenetic_workspace_bytes(Bd, k)
            wsc = torch.empty(needc, dtype=torch.uint8, device=dev)
            out = torch.empty((Bd, n, n), dtype=torch.float64, device=dev)
            for name, blp in (("topological", 0), ("fp64", bls.data_ptr())):
                t = timeit(lambda: lib.p2v_cophenetic_batch(Vd.data_ptr(), 8, blp, Bd, k, 0, 0, n, out.data_ptr(), None,
                                                            wsc.data_ptr(), needc, s), reps=2)
                rec[f"cophenetic_{name}"] = {"batch": Bd, "trees_per_s": Bd / t, "gbs": Bd * n * n * 8 / t / 1e9,
                                             "frac_of_measured_hbm": Bd * n * n * 8 / t / 1e9 / PEAK,
                                             "includes": "decode + tables + blocks + rows"}
            if n <= 4096:   # literal reference algorithm on the CPU, bounded sample
                Bcc = int(max(1, min(Bd, (1 << 31) // ((2 * n) ** 2 * 8), 4 * nthreads)))
                t0 = time.perf_counter()
                refd = O.cophenetic_batch(Vd[:Bcc].cpu().numpy(), None, nthreads=nthreads)
                tcd = time.perf_counter() - t0
                lib.p2v_cophenetic_batch(Vd.data_ptr(), 8, 0, Bd, k, 0, 0, n, out.data_ptr(), None, wsc.data_ptr(), needc, s)
                rec["cophenetic_cpu"] = {"sample": Bcc, "trees_per_s": Bcc / tcd, "cores": nthreads,
                                         "matches_gpu": bool(np.array_equal(refd, out[:Bcc].cpu().numpy()))}
            del out, wsc, bls
        del V
        torch.cuda.empty_cache()
        rows.append(rec)
        print(json.dumps(rec), flush=True)
    with open(args.out, "w") as f:
        json.dump({"peak_hbm_gbs": PEAK, "host_threads": nthreads, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
