#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json): batched to_ancestry at n=1000, batch 1M,
plus the cophenetic row kernel at n=65536.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A step = one pass of to_ancestry over one batch of synthetic vectors (1M trees per
GPU, int64, inputs resident in HBM and larger than L2).  Prints ONE JSON line (rank 0).
"""

from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_LEAVES = 1000
K = N_LEAVES - 1
BATCH = 1_000_000
METRIC = "trees/s to_ancestry @n=1k batch 1M; cophenetic matrix HBM GB/s @n=64k"
WORKLOAD = "to_ancestry: batch of 1M random trees n=1000, int64 (BASELINE.json configs[1]); v ~ UniformInt{0..2i}"
COPH_N = 65536
C3_TREES, C3_N = 10_000, 2000        # BASELINE.json configs[2]
E2E_BATCH = 500_000                  # trees per GPU per end-to-end step, the same at every N
CPU_SAMPLE = 262_144                 # trees of the cpu_baseline leg (about 2 s on 16 threads)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def synth_vectors_device(torch, B, k, seed, device):
    g = torch.Generator(device=device).manual_seed(seed)
    hi = (2 * torch.arange(k, device=device) + 1).to(torch.float64)
    out = torch.empty((B, k), dtype=torch.int64, device=device)
    step = 1 << 17
    for b0 in range(0, B, step):
        b1 = min(B, b0 + step)
        out[b0:b1] = (torch.rand((b1 - b0, k), generator=g, device=device, dtype=torch.float64) * hi).to(torch.int64)
    return out


def synth_vectors_host(np, B, k, seed):
    rng = np.random.Generator(np.random.Philox(seed))
    hi = 2 * np.arange(k, dtype=np.int64) + 1
    return (rng.random((B, k)) * hi).astype(np.int64)


# --------------------------------------------------------------------------- reference arm
def host_threads():
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1; the baseline
    passes its thread count to OpenMP explicitly)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def host_threads_per_rank(world):
    """Worker threads the host entry points use per process (p2v_capi.cu host_thread_count)."""
    e = os.environ.get("P2V_HOST_THREADS")
    if e and int(e) > 0:
        return int(e)
    n = host_threads()
    lw = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
    if lw > 1:
        n //= lw
    return max(2, min(16, n))


def cpu_baseline(np, sample_trees, nthreads=0, repeats=1):
    """The reference's CPU algorithm (oracle port: AVL decode + relabel pass,
    phylo2vec/src/vector/avl.rs, convert.rs:167-188), one reference call per tree,
    OpenMP parallel-for over trees on the host cores."""
    from oracle import p2v_oracle as O
    if nthreads <= 0:
        nthreads = host_threads()
    V = synth_vectors_host(np, sample_trees, K, seed=1)
    out = np.empty((sample_trees, K, 3), dtype=np.int64)
    O.batch("to_ancestry", V[:256], K, nthreads=nthreads)      # warm up threads / pages
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        O.batch("to_ancestry", V, K, nthreads=nthreads, out=out)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    cores = nthreads
    return sample_trees / best, cores, best


def run_reference(args):
    """The reference's CPU algorithm on the accelerated arm's own configuration: every step decodes
    --batch trees (1 M by default) of n = 1000 with all host threads."""
    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.batch
    from oracle import p2v_oracle as O
    nthreads = host_threads()
    V = synth_vectors_host(np, sample, K, seed=1)
    out = np.empty((sample, K, 3), dtype=np.int64)
    for _ in range(args.warmup):
        O.batch("to_ancestry", V[:min(sample, 65536)], K, nthreads=nthreads, out=out[:min(sample, 65536)])
    t_all = 0.0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        O.batch("to_ancestry", V, K, nthreads=nthreads, out=out)
        t_all += time.perf_counter() - t0
    value = sample * args.steps / t_all
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "trees/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": sample, "n_leaves": N_LEAVES, "idx_bytes": 8,
                   "note": "host arm: one process, all host threads, the same trees per step as the accelerated arm's "
                           "per-GPU batch; warm-up steps use 65536 trees"},
        "cpu_baseline": {"value": value, "unit": "trees/s", "cores": nthreads, "kind": "port",
                         "sample": f"{sample} trees of n=1000 per step, all host threads (OpenMP over trees)"},
        "e2e": {"value": value, "unit": "trees/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from phylo2vec_b200 import _capi, batch as pb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = _capi.load()
    peaks, peak_src = measured_peaks()
    # this rank's host thread (and the pinned buffers it allocates) next to its GPU's NUMA node
    numa_cpus = int(lib.p2v_host_bind_thread())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B = args.batch
    V = synth_vectors_device(torch, B, K, seed=1234 + rank, device=device)
    out = torch.empty((B, K, 3), dtype=torch.int64, device=device)
    status = torch.empty(B, dtype=torch.int32, device=device)
    stream = torch.cuda.current_stream(device)

    def step():
        rc = lib.p2v_to_ancestry_batch(V.data_ptr(), 8, B, K, out.data_ptr(), status.data_ptr(), None, 0,
                                       stream.cuda_stream)
        if rc != 0:
            raise RuntimeError(f"p2v_to_ancestry_batch failed: {lib.p2v_last_error()}")

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _capi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = _capi.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if int(status.count_nonzero()) != 0:
        raise RuntimeError("decode reported invalid vectors on synthetic input")
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * B / (ms_step * 1e-3)
    alg_bytes = 4 * K * 8 * B                       # 4*k*w per tree, w = 8
    achieved = alg_bytes / (ms_total / args.steps * 1e-3) / 1e9
    peak = float(peaks["hbm_gbs"])

    # ---- parity spot check on the timed buffers (never timed) -----------------------------
    parity = None
    if rank == 0:
        from oracle import p2v_oracle as O
        idx = torch.arange(0, B, max(1, B // 32), device=device)[:32]
        ref, _ = O.batch("to_ancestry", V[idx].cpu().numpy(), K)
        parity = bool(np.array_equal(out[idx].cpu().numpy(), ref))

    # ---- the other functions of BASELINE config C2, same batch (device resident, event timed) ----
    c2 = {}
    if rank == 0:
        def timed(fn, reps=3):
            fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                fn()
            b.record(stream)
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        Bs = min(B, 250_000)
        Vs = V[:Bs]
        for name, cols in (("to_pairs", 2), ("to_edges", 4)):
            o = torch.empty((Bs, K * cols), dtype=torch.int64, device=device)
            fnc = getattr(lib, f"p2v_{name}_batch")
            ms = timed(lambda: fnc(Vs.data_ptr(), 8, Bs, K, o.data_ptr(), status.data_ptr(), None, 0, stream.cuda_stream))
            c2[name] = {"trees_per_s": Bs / (ms * 1e-3), "gbs": (1 + cols) * K * 8 * Bs / (ms * 1e-3) / 1e9}
            del o
        anc_s = out[:Bs]
        vb = torch.empty((Bs, K), dtype=torch.int64, device=device)
        need_enc = int(lib.p2v_encode_workspace_bytes(Bs, K))
        ws_enc = torch.empty(max(need_enc, 16), dtype=torch.uint8, device=device)

        def enc_call():
            rc = lib.p2v_from_ancestry_batch(anc_s.data_ptr(), 8, Bs, K, vb.data_ptr(), status.data_ptr(), ws_enc.data_ptr(),
                                             need_enc, stream.cuda_stream)
            assert rc == 0, lib.p2v_last_error()
        ms = timed(enc_call)
        c2["from_ancestry"] = {"trees_per_s": Bs / (ms * 1e-3), "gbs": 4 * K * 8 * Bs / (ms * 1e-3) / 1e9,
                               "round_trip_equal": bool(torch.equal(vb, Vs))}
        c2["batch"] = Bs
        del vb, ws_enc

    # ---- end to end: host buffers through the C ABI (H2D + kernels + D2H in the timed region) ----
    # The same per-GPU batch at every N, --steps timed steps; the staging of the host entry points is
    # kept between the steps (a streaming caller's setting, include/phylo2vec_b200.h).
    Be = args.e2e_batch if args.e2e_batch > 0 else min(B, E2E_BATCH)
    lib.p2v_host_cache_limit(-1)
    hV = torch.empty((Be, K), dtype=torch.int64).pin_memory()
    hV.copy_(V[:Be])
    hOut = torch.empty((Be, K, 3), dtype=torch.int64).pin_memory()
    hSt = torch.empty(Be, dtype=torch.int32).pin_memory()

    def e2e_step(hv=hV, ho=hOut, hs=hSt, nb=Be):
        rc = lib.p2v_to_ancestry_host(hv.data_ptr(), 8, nb, K, ho.data_ptr(), hs.data_ptr())
        if rc != 0:
            raise RuntimeError(f"p2v_to_ancestry_host failed: {lib.p2v_last_error()}")

    e2e_step()
    barrier()
    e2e_steps = max(1, args.steps)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * Be * e2e_steps / float(te.item())
    # every 997th tree of the whole output against the device path (a strided sample over all chunks)
    sel = torch.arange(0, Be, 997)
    e2e_ok = bool(torch.equal(hOut[sel], out[sel.to(device)].cpu())) and int(hSt.count_nonzero()) == 0
    # the same call on pageable buffers (what a plain numpy caller passes), rank 0, bounded
    e2e_pageable = None
    if rank == 0:
        Bp = min(Be, 200_000)
        pV = hV[:Bp].numpy().copy()
        pOut = np.empty((Bp, K, 3), dtype=np.int64)
        pSt = np.zeros(Bp, dtype=np.int32)
        cp = lambda a: a.ctypes.data_as(ctypes.c_void_p)     # noqa: E731
        lib.p2v_to_ancestry_host(cp(pV), 8, Bp, K, cp(pOut), cp(pSt))
        t0 = time.perf_counter()
        rc = lib.p2v_to_ancestry_host(cp(pV), 8, Bp, K, cp(pOut), cp(pSt))
        tp = time.perf_counter() - t0
        e2e_pageable = {"value": Bp / tp, "unit": "trees/s", "batch": Bp, "ok": rc == 0 and bool(
            np.array_equal(pOut[::997], hOut[:Bp][::997].numpy()))}
        del pV, pOut
    del hV, hOut, hSt
    lib.p2v_host_release()

    del V, out
    torch.cuda.empty_cache()
    # ---- second half of the metric: the n = 65536 matrix, row blocks over the ranks (config C4) ----
    coph = None
    try:
        coph = bench_c4(torch, dist, lib, device, rank, world, peak, args)
    except Exception as exc:  # keep the headline line even if this part fails
        coph = {"error": repr(exc)}
    torch.cuda.empty_cache()
    # ---- config C3: 10 k trees of n = 2000, topological and branch-length matrices, batch-sharded ----
    c3 = None
    try:
        c3 = bench_c3(torch, dist, lib, device, rank, world, peak, args)
    except Exception as exc:
        c3 = {"error": repr(exc)}
    torch.cuda.empty_cache()
    newick = (None, None, None)
    if rank == 0:
        try:
            newick = bench_newick(torch, lib, device, peak)
        except Exception as exc:
            newick = ({"error": repr(exc)},) * 3
    round2 = {}
    if rank == 0:
        try:
            round2 = bench_round2_rows(torch, device)
        except Exception as exc:
            round2 = {"round2_rows_error": repr(exc)}

    if rank == 0:
        cb_rate, cb_cores, cb_dt = cpu_baseline(np, CPU_SAMPLE)
        cb1_rate, _, cb1_dt = cpu_baseline(np, 4096, nthreads=1)      # what one reference call does: one core
        try:
            import phylo2vec as _ref_pkg  # noqa: F401  (a driver-provided install of the real reference)
            ref_python = "importable (not timed here)"
        except Exception:
            ref_python = "unavailable (no Rust toolchain / wheel in this image)"
        line = {
            "metric": METRIC, "value": value, "unit": "trees/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "n_leaves": N_LEAVES, "idx_bytes": 8,
                       "sharding": "tree-batch index, no collective",
                       "l2": "inputs (8 GB) and outputs (24 GB) per step exceed the 126 MB L2"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "frac_of_8tbs_spec": achieved / 8000.0,
                         "traffic": NCU_DRAM_BYTES_PER_TREE * B,
                         "traffic_source": f"ncu --set full capture {NCU_DRAM_SOURCE}: "
                                           "dram read+write per launch of 131072 trees = "
                                           f"{NCU_DRAM_BYTES_PER_TREE} B/tree, scaled to this launch's {B} trees",
                         "peak_source": peak_src,
                         "kernel": "decode_small_kernel<int64, ancestry>",
                         "algorithmic_bytes_per_tree": 4 * K * 8,
                         "note": "integer kernel bound by the L1/shared-memory data pipe (ncu: 90 % of peak), not by HBM; HBM fraction reported as required"},
            "cpu_baseline": {"value": cb_rate, "unit": "trees/s", "cores": cb_cores, "kind": "port",
                             "sample": f"{CPU_SAMPLE} trees of n=1000, {cb_dt:.2f} s, OpenMP over trees",
                             "single_core_value": cb1_rate, "single_core_sample": f"4096 trees, {cb1_dt:.2f} s",
                             "reference_python_package": ref_python},
            "e2e": {"value": e2e_value, "unit": "trees/s", "h2d_bytes_per_step": Be * K * 8,
                    "d2h_bytes_per_step": Be * K * 4 + Be * 4, "batch_per_gpu": Be, "steps": e2e_steps,
                    "host_gbs": world * Be * e2e_steps * K * (8 + 4 + 24) / float(te.item()) / 1e9,
                    "host_threads_per_rank": host_threads_per_rank(world), "numa_bound_cpus": numa_cpus,
                    "pageable_buffers": e2e_pageable,
                    "result_bytes_per_step": Be * K * 24 + Be * 4,
                    "transport": "the int64 (k,3) result crosses PCIe as its two informative columns in 16 bits each (the "
                                 "third is k+1+row); host threads of the library widen them into the caller's int64 rows",
                    "api": "p2v_to_ancestry_host (C ABI, pinned host buffers)", "matches_device_path": e2e_ok},
            "gpu_launches": launches, "clocks": clocks, "parity_spot_check": parity,
            "config_c2_other_functions": c2, "cophenetic_c4": coph, "c3": c3, "next_rows": {"to_newick": newick[0], "from_newick": newick[1], "balance_indices": newick[2], **round2},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_newick(torch, lib, device, peak):
    """The second row beyond the scope table: batched to_newick, GPU next to the CPU
    restatement on a bounded sample.  Bytes = v read + string bytes written."""
    import numpy as np
    from oracle import p2v_oracle as O
    res, res2, res3 = {}, {}, {}
    for n, B in ((100, 1_000_000), (1000, 100_000)):
        k = n - 1
        g = torch.Generator(device=device).manual_seed(n)
        hi = (2 * torch.arange(k, device=device) + 1).to(torch.float64)
        V = (torch.rand((B, k), generator=g, device=device, dtype=torch.float64) * hi).to(torch.int64)
        stride = (int(lib.p2v_newick_max_bytes(k)) + 15) & ~15
        out = torch.empty((B, stride), dtype=torch.uint8, device=device)
        lengths = torch.empty(B, dtype=torch.int64, device=device)
        st = torch.empty(B, dtype=torch.int32, device=device)
        need = int(lib.p2v_newick_workspace_bytes(B, k))
        ws = torch.empty(max(need, 16), dtype=torch.uint8, device=device)
        stream = torch.cuda.current_stream(device)

        def call():
            rc = lib.p2v_to_newick_batch(V.data_ptr(), 8, B, k, out.data_ptr(), stride, lengths.data_ptr(),
                                         st.data_ptr(), ws.data_ptr(), need, stream.cuda_stream)
            assert rc == 0, lib.p2v_last_error()
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5):
            call()
        e1.record(stream)
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 5 * 1e-3
        nbytes = int(lengths.sum()) + B + B * k * 8
        Bc = min(B, 20000 if n == 100 else 4000)
        Vc = V[:Bc].cpu().numpy()
        t0 = time.perf_counter()
        ref, reflen = O.to_newick_batch(Vc, stride, nthreads=host_threads())
        tc = time.perf_counter() - t0
        same = bool(np.array_equal(reflen, lengths[:Bc].cpu().numpy())) and bool(
            np.array_equal(ref[:, :int(reflen.max())], out[:Bc, :int(reflen.max())].cpu().numpy()))
        res[f"n{n}"] = {"batch": B, "trees_per_s": B / t, "gbs": nbytes / t / 1e9, "frac_of_peak": nbytes / t / 1e9 / peak,
                        "cpu": {"sample": Bc, "trees_per_s": Bc / tc, "cores": host_threads(), "kind": "port"},
                        "matches_cpu": same}
        # and back: from_newick on the strings just written (bytes = string bytes read + v written)
        back = torch.empty((B, k), dtype=torch.int64, device=device)
        need2 = int(lib.p2v_from_newick_workspace_bytes(B, k, 8))
        del ws
        ws2 = torch.empty(max(need2, 16), dtype=torch.uint8, device=device)

        def call2():
            rc = lib.p2v_from_newick_batch(out.data_ptr(), stride, lengths.data_ptr(), B, k, back.data_ptr(), 8,
                                           st.data_ptr(), ws2.data_ptr(), need2, stream.cuda_stream)
            assert rc == 0, lib.p2v_last_error()
        for _ in range(3):
            call2()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(5):
            call2()
        e1.record(stream)
        torch.cuda.synchronize()
        t2 = e0.elapsed_time(e1) / 5 * 1e-3
        t0 = time.perf_counter()
        refv, refst = O.from_newick_batch(ref, reflen, k, nthreads=host_threads())
        tc2 = time.perf_counter() - t0
        res2[f"n{n}"] = {"batch": B, "trees_per_s": B / t2, "gbs": nbytes / t2 / 1e9, "frac_of_peak": nbytes / t2 / 1e9 / peak,
                         "cpu": {"sample": Bc, "trees_per_s": Bc / tc2, "cores": host_threads(), "kind": "port"},
                         "round_trip_equal": bool(torch.equal(back, V)) and not bool(st.any()),
                         "matches_cpu": bool(np.array_equal(refv, Vc)) and not bool(refst.any())}
        del out, ws2, back
        # third row beyond the scope table, first half: Sackin / leaf-depth variance / B2 (bytes = v read + 24 written)
        bal = torch.empty((B, 3), dtype=torch.float64, device=device)
        need3 = int(lib.p2v_balance_workspace_bytes(B, k))
        ws3 = torch.empty(max(need3, 16), dtype=torch.uint8, device=device)

        def call3():
            rc = lib.p2v_balance_batch(V.data_ptr(), 8, B, k, bal.data_ptr(), st.data_ptr(), ws3.data_ptr(), need3,
                                       stream.cuda_stream)
            assert rc == 0, lib.p2v_last_error()
        for _ in range(3):
            call3()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(5):
            call3()
        e1.record(stream)
        torch.cuda.synchronize()
        t3 = e0.elapsed_time(e1) / 5 * 1e-3
        Bc3 = min(Bc, 2000)
        t0 = time.perf_counter()
        refb = np.stack([O.balance(v) for v in Vc[:Bc3]])
        tc3 = time.perf_counter() - t0
        gotb = bal[:Bc3].cpu().numpy()
        nb3 = B * k * 8 + B * 24
        res3[f"n{n}"] = {"batch": B, "trees_per_s": B / t3, "gbs": nb3 / t3 / 1e9, "frac_of_peak": nb3 / t3 / 1e9 / peak,
                         "cpu": {"sample": Bc3, "trees_per_s": Bc3 / tc3, "cores": 1, "kind": "port"},
                         "matches_cpu": bool(np.array_equal(gotb[:, :2], refb[:, :2])) and bool(
                             np.all(np.abs(gotb[:, 2] - refb[:, 2]) <= 1e-12 * np.abs(refb[:, 2])))}
        del V, ws3, bal
    return res, res2, res3


def bench_round2_rows(torch, device):
    """The rows added in round 2 (SURVEY 8f-2 matrices, 8f-3 Robinson-Foulds, 8f-4 CSV): throughput through the
    package's batched API (CUDA events) with the oracle on a small sample beside it.  First versions: these
    numbers say where they stand, not what bounds them."""
    import io as _io
    import numpy as np
    import phylo2vec_b200 as p2v
    from oracle import p2v_oracle as O
    from phylo2vec_b200.io import reader
    out = {}
    stream = torch.cuda.current_stream(device)

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            r = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3, r

    for n, B in ((100, 200_000), (1000, 20_000)):
        k = n - 1
        g = torch.Generator(device=device).manual_seed(7 * n)
        hi = (2 * torch.arange(k, device=device) + 1).to(torch.float64)
        V1 = (torch.rand((B, k), generator=g, device=device, dtype=torch.float64) * hi).to(torch.int64)
        V2 = (torch.rand((B, k), generator=g, device=device, dtype=torch.float64) * hi).to(torch.int64)
        t, rf = timed(lambda: p2v.robinson_foulds_batch(V1, V2, check=False))
        ns = 200 if n == 100 else 20
        v1, v2 = V1[:ns].cpu().numpy(), V2[:ns].cpu().numpy()
        t0 = time.perf_counter()
        ref = np.array([O.robinson_foulds(v1[i], v2[i]) for i in range(ns)])
        tc = time.perf_counter() - t0
        out.setdefault("robinson_foulds", {})[f"n{n}"] = {
            "pairs": B, "pairs_per_s": B / t, "cpu": {"sample": ns, "pairs_per_s": ns / tc, "cores": 1, "kind": "port (numpy / Python sets)"},
            "matches_cpu": bool(np.array_equal(rf[:ns].cpu().numpy(), ref))}
        # matrices with branch lengths -> Newick strings -> matrices
        Bm = B // 4
        M = torch.empty((Bm, k, 3), dtype=torch.float64, device=device)
        M[..., 0] = V1[:Bm].to(torch.float64)
        M[..., 1:] = torch.rand((Bm, k, 2), generator=g, device=device, dtype=torch.float64)
        t1, (buf, ln) = timed(lambda: p2v.to_newick_batch(M, check=False, raw=True))
        t2, M2 = timed(lambda: p2v.to_matrix_batch(buf, lengths=ln, n_leaves=n, check=False))
        s0 = bytes(buf[0, :int(ln[0])].cpu().numpy()).decode()
        out.setdefault("matrix_newick", {})[f"n{n}"] = {
            "trees": Bm, "to_newick_trees_per_s": Bm / t1, "from_newick_trees_per_s": Bm / t2,
            "string_bytes_per_tree": float(ln.double().mean()), "round_trip_bit_equal": bool(torch.equal(M2, M)),
            "first_string_matches_cpu": s0 == O.to_newick_matrix(M[0].cpu().numpy())}
    # CSV: a file of vectors, one per line, parsed on the GPU
    n, B = 1000, 20_000
    k = n - 1
    rng = np.random.default_rng(3)
    Vh = (rng.random((B, k)) * (2 * np.arange(k) + 1)).astype(np.int64)
    sio = _io.StringIO()
    np.savetxt(sio, Vh, fmt="%d", delimiter=",")
    text = sio.getvalue().encode()
    t0 = time.perf_counter()
    for _ in range(3):
        Vd = reader.load_batch(_io.BytesIO(text), kind="vector", device=device, check=False)
    torch.cuda.synchronize()
    tw = (time.perf_counter() - t0) / 3
    t0 = time.perf_counter()
    head = b"\n".join(text.split(b"\n", B // 20)[: B // 20])           # the first twentieth of the lines
    np.loadtxt(_io.BytesIO(head), delimiter=",", dtype=np.int64)
    tnp = (time.perf_counter() - t0) * 20
    out["csv_load_batch"] = {"trees": B, "n_leaves": n, "file_bytes": len(text), "seconds_wall_incl_h2d": tw,
                             "file_gbs": len(text) / tw / 1e9, "numpy_loadtxt_seconds_extrapolated": tnp,
                             "equal": bool(np.array_equal(Vd.cpu().numpy(), Vh))}
    return out


def _evt_ms(torch, stream, fn, reps):
    """Mean device time of fn() over reps back-to-back calls, CUDA events on the launching stream."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _max_over_ranks(torch, dist, device, world, x):
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _all_true(torch, dist, device, world, ok):
    t = torch.tensor([0.0 if ok else 1.0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) == 0.0


def bench_c4(torch, dist, lib, device, rank, world, peak, args):
    """BASELINE config C4: ONE tree of n = 65536, float64 matrix, strong-scaled: rank r writes rows
    [n*r/W, n*(r+1)/W) of the matrix into its slice of a preallocated (n,n) buffer; every rank decodes
    the tree and builds its tables itself (prepare), then writes its rows.  Timed per rank with CUDA
    events, max over ranks: prepare, the rows call, and the one-shot call (prepare + rows, the number
    the 70 % target is about).  The optional NCCL all-gather of the row blocks is in place
    (all_gather_into_tensor on the same buffer) and timed separately.  Sampled rows of every rank are
    checked against the closed-form CPU restatement (never timed)."""
    import numpy as np
    from phylo2vec_b200 import distributed as pd
    n = COPH_N
    k = n - 1
    g = torch.Generator(device=device).manual_seed(7)            # same tree on every rank
    hi = (2 * torch.arange(k, device=device) + 1).to(torch.float64)
    v = (torch.rand((1, k), generator=g, device=device, dtype=torch.float64) * hi).to(torch.int64)
    bls = torch.rand((1, k, 2), generator=g, device=device, dtype=torch.float64)
    r0, r1 = pd.shard_rows(n, rank, world)
    rows = r1 - r0
    full = torch.empty((n, n), dtype=torch.float64, device=device)
    mine = full[r0:r1]
    need = lib.p2v_cophenetic_workspace_bytes(1, k)
    ws = torch.empty(need, dtype=torch.uint8, device=device)
    stream = torch.cuda.current_stream(device)
    sp = stream.cuda_stream
    reps = max(3, min(args.steps, 10))
    res = {"workload": "one tree n=65536, float64 matrix, row blocks over the ranks (BASELINE.json configs[3])",
           "n_leaves": n, "n_gpus": world, "scaling": "strong", "rows_this_rank": [r0, r1],
           "bytes_per_rank": rows * n * 8, "reps": reps}
    nbytes = rows * n * 8
    hv = v[0].cpu().numpy()
    hb = bls[0].cpu().numpy()
    for label, weighted, blp in (("fp64_branch_lengths", 1, bls.data_ptr()), ("topological", 0, 0)):
        def prep():
            rc = lib.p2v_cophenetic_prepare_batch(v.data_ptr(), 8, blp, 1, k, 0, None, ws.data_ptr(), need, sp)
            assert rc == 0, lib.p2v_last_error()

        def rws():
            rc = lib.p2v_cophenetic_rows_batch(ws.data_ptr(), need, 1, k, weighted, r0, r1, mine.data_ptr(), sp)
            assert rc == 0, lib.p2v_last_error()

        def oneshot():
            rc = lib.p2v_cophenetic_batch(v.data_ptr(), 8, blp, 1, k, 0, r0, r1, mine.data_ptr(), None, ws.data_ptr(), need, sp)
            assert rc == 0, lib.p2v_last_error()
        for _ in range(3):
            oneshot()
        if world > 1:
            dist.barrier()
        prep_ms = _max_over_ranks(torch, dist, device, world, _evt_ms(torch, stream, prep, reps))
        rows_ms = _max_over_ranks(torch, dist, device, world, _evt_ms(torch, stream, rws, reps))
        launches0 = int(lib.p2v_launch_count())
        total_ms = _max_over_ranks(torch, dist, device, world, _evt_ms(torch, stream, oneshot, reps))
        launches = (int(lib.p2v_launch_count()) - launches0) // reps
        # parity of this rank's block: four sampled rows against the closed form (long-double sums)
        ok, worst = True, 0.0
        tree = np.concatenate([hv[:, None].astype(np.float64), hb], axis=1) if weighted else hv
        from oracle import p2v_oracle as O
        for rr in sorted({r0, r0 + rows // 3, r0 + (2 * rows) // 3, r1 - 1}):
            ref = O.cophenetic_rows(tree, False, rr, rr + 1)[0]
            got = mine[rr - r0].cpu().numpy()
            if weighted:
                err = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)))
                worst = max(worst, err)
                ok = ok and err <= 1e-12
            else:
                ok = ok and bool(np.array_equal(got, ref))
        res[label] = {
            "prepare_ms": prep_ms, "rows_ms": rows_ms, "total_ms": total_ms, "launches_per_call": launches,
            "rows_gbs_per_gpu": nbytes / (rows_ms * 1e-3) / 1e9, "rows_frac_of_measured_hbm": nbytes / (rows_ms * 1e-3) / 1e9 / peak,
            "total_gbs_per_gpu": nbytes / (total_ms * 1e-3) / 1e9,
            "total_frac_of_measured_hbm": nbytes / (total_ms * 1e-3) / 1e9 / peak,
            "total_frac_of_8tbs_spec": nbytes / (total_ms * 1e-3) / 1e9 / 8000.0,
            "aggregate_gbs": n * n * 8 / (total_ms * 1e-3) / 1e9,
            "parity": _all_true(torch, dist, device, world, ok),
            "parity_bar": "<= 1e-12 relative, 4 rows per rank vs closed form" if weighted else "bit-exact, 4 rows per rank",
        }
        if weighted:
            res[label]["max_rel_err_this_rank"] = worst
    # ---- optional collective: in-place all-gather of the row blocks (NCCL over NVLink / NVSwitch) ----
    if world > 1:
        try:
            lib.p2v_cophenetic_batch(v.data_ptr(), 8, bls.data_ptr(), 1, k, 0, r0, r1, mine.data_ptr(), None, ws.data_ptr(), need, sp)
            torch.cuda.synchronize()
            pd.all_gather_rows_(full, n)           # warm-up (communicator set-up)
            ag_ms = _max_over_ranks(torch, dist, device, world, _evt_ms(torch, stream, lambda: pd.all_gather_rows_(full, n), 2))
            # one row of the next rank's block must now be here, equal to the closed form
            o0, _o1 = pd.shard_rows(n, (rank + 1) % world, world)
            from oracle import p2v_oracle as O
            tree = np.concatenate([hv[:, None].astype(np.float64), hb], axis=1)
            ref = O.cophenetic_rows(tree, False, o0 + 5, o0 + 6)[0]
            got = full[o0 + 5].cpu().numpy()
            err = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)))
            res["all_gather"] = {
                "ms": ag_ms, "bytes_received_per_gpu": (n - rows) * n * 8,
                "algbw_gbs": n * n * 8 / (ag_ms * 1e-3) / 1e9, "busbw_gbs": (n - rows) * n * 8 / (ag_ms * 1e-3) / 1e9,
                "in_place": True, "parity": _all_true(torch, dist, device, world, err <= 1e-12),
                "note": "dist.all_gather_into_tensor on the preallocated (n,n) buffer, sendcount = rows*n float64"}
        except Exception as exc:
            res["all_gather"] = {"error": repr(exc)}
    fp = res["fp64_branch_lengths"]
    res["roofline"] = {"bound": "hbm", "achieved": fp["total_gbs_per_gpu"], "peak": peak, "unit": "GB/s",
                       "frac": fp["total_frac_of_measured_hbm"], "kernel": "prepare (decode + tables) + rows_kernel<fp64>, per GPU",
                       "algorithmic_bytes": nbytes, "traffic": int(34.332404e9 * rows / n),
                       "traffic_source": "ncu --set full, profiles/r01_rows_full_fp64_v7.summary.txt: 34.303 GB written "
                                         "+ 0.030 GB read for all 65536 rows, scaled to this rank's rows"}
    return res


def bench_c3(torch, dist, lib, device, rank, world, peak, args):
    """BASELINE config C3: 10 000 trees of n = 2000; topological matrices from the vectors and
    branch-length matrices from the (k,3) float64 matrices, float64 (n,n) each (32 MB per tree, 320 GB in
    all).  Strong-scaled: the batch is sharded by tree index over the ranks, no collective; every rank
    walks its shard in chunks that reuse one output buffer.  Max over ranks; one full matrix per rank and
    mode is checked against the literal CPU restatement (never timed)."""
    import numpy as np
    from oracle import p2v_oracle as O
    from phylo2vec_b200 import distributed as pd
    n, k = C3_N, C3_N - 1
    total = args.c3_trees
    b0, b1 = pd.shard_batch(total, rank, world)
    Bl = b1 - b0
    rng = np.random.Generator(np.random.Philox(2000))
    hV = (rng.random((total, k)) * (2 * np.arange(k) + 1)).astype(np.int64)[b0:b1]
    hB = np.random.Generator(np.random.Philox(2001)).random((total, k, 2))[b0:b1]
    V = torch.from_numpy(np.ascontiguousarray(hV)).to(device)
    M = torch.from_numpy(np.concatenate([hV[..., None].astype(np.float64), hB], axis=2)).to(device)
    chunk = max(1, min(Bl, 512))
    out = torch.empty((chunk, n, n), dtype=torch.float64, device=device)
    need = lib.p2v_cophenetic_workspace_bytes(chunk, k)
    ws = torch.empty(need, dtype=torch.uint8, device=device)
    st = torch.empty(chunk, dtype=torch.int32, device=device)
    stream = torch.cuda.current_stream(device)
    sp = stream.cuda_stream
    res = {"workload": "10k trees n=2000: cophenetic_distances, topological (vectors) and branch lengths "
                       "((k,3) float64 matrices), float64 (n,n) out (BASELINE.json configs[2])",
           "trees": total, "n_leaves": n, "n_gpus": world, "scaling": "strong", "sharding": "tree-batch index, no collective",
           "trees_this_rank": Bl, "chunk_trees": chunk, "bytes_total": total * n * n * 8}

    def sweep(weighted):
        for c0 in range(0, Bl, chunk):
            nb = min(chunk, Bl - c0)
            if weighted:
                rc = lib.p2v_cophenetic_matrix_batch(M[c0:c0 + nb].data_ptr(), nb, k, 0, 0, n, out.data_ptr(), st.data_ptr(),
                                                     ws.data_ptr(), need, sp)
            else:
                rc = lib.p2v_cophenetic_batch(V[c0:c0 + nb].data_ptr(), 8, 0, nb, k, 0, 0, n, out.data_ptr(), st.data_ptr(),
                                              ws.data_ptr(), need, sp)
            assert rc == 0, lib.p2v_last_error()

    for label, weighted in (("topological", 0), ("fp64_branch_lengths", 1)):
        sweep(weighted)                                      # warm-up pass over the shard
        if world > 1:
            dist.barrier()
        launches0 = int(lib.p2v_launch_count())
        ms_local = _evt_ms(torch, stream, lambda: sweep(weighted), 1)
        launches = int(lib.p2v_launch_count()) - launches0
        ms = _max_over_ranks(torch, dist, device, world, ms_local)
        # parity: the last chunk is still in `out`; its first tree against the literal reference loop
        last0 = ((Bl - 1) // chunk) * chunk
        got = out[0].cpu().numpy()
        tree = M[last0].cpu().numpy() if weighted else hV[last0]
        ref = O.cophenetic_distances(tree)
        if weighted:
            err = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)))
            ok = err <= 1e-12
        else:
            err = 0.0
            ok = bool(np.array_equal(got, ref))
        ok = ok and int(st.count_nonzero()) == 0
        res[label] = {"ms": ms, "trees_per_s": total / (ms * 1e-3), "aggregate_gbs": total * n * n * 8 / (ms * 1e-3) / 1e9,
                      "gbs_per_gpu": Bl * n * n * 8 / (ms_local * 1e-3) / 1e9,
                      "frac_of_measured_hbm_per_gpu": Bl * n * n * 8 / (ms_local * 1e-3) / 1e9 / peak,
                      "launches": launches, "includes": "input adapters + decode + tables + blocks + rows, every chunk",
                      "parity": _all_true(torch, dist, device, world, ok),
                      "parity_bar": "<= 1e-12 relative, one full matrix per rank vs the literal reference loop" if weighted
                      else "bit-exact, one full matrix per rank vs the literal reference loop",
                      "max_rel_err_this_rank": err}
    del out, ws
    torch.cuda.empty_cache()
    # ---- the same call end to end: host (pinned) buffers through p2v_cophenetic_matrix_host, rank 0's shard ----
    try:
        Be = min(Bl, 64)
        hM = torch.empty((Be, k, 3), dtype=torch.float64).pin_memory()
        hM.copy_(M[:Be])
        hD = torch.empty((Be, n, n), dtype=torch.float64).pin_memory()
        hS = torch.empty(Be, dtype=torch.int32).pin_memory()

        def host_call():
            rc = lib.p2v_cophenetic_matrix_host(hM.data_ptr(), Be, k, 0, 0, n, hD.data_ptr(), hS.data_ptr())
            assert rc == 0, lib.p2v_last_error()
        lib.p2v_host_cache_limit(-1)
        host_call()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        steps = 3
        for _ in range(steps):
            host_call()
        dt = _max_over_ranks(torch, dist, device, world, time.perf_counter() - t0)
        ref = O.cophenetic_distances(hM[Be - 1].numpy())
        got = hD[Be - 1].numpy()
        err = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)))
        res["e2e"] = {"value": world * Be * steps / dt, "unit": "trees/s", "api": "p2v_cophenetic_matrix_host (pinned host buffers)",
                      "batch_per_gpu": Be, "steps": steps, "h2d_bytes_per_step": Be * k * 24, "d2h_bytes_per_step": Be * n * n * 8 + Be * 4,
                      "d2h_gbs_per_gpu": Be * steps * n * n * 8 / dt / 1e9,
                      "parity": _all_true(torch, dist, device, world, err <= 1e-12 and int(hS.count_nonzero()) == 0)}
        del hM, hD, hS
        lib.p2v_host_release()
        lib.p2v_host_cache_limit(512 << 20)
    except Exception as exc:
        res["e2e"] = {"error": repr(exc)}
    return res


# dram__bytes_read.sum + dram__bytes_write.sum of decode_small_kernel<int64, ancestry> per tree
# (profiles/r01_decode_small_v7.summary.txt: (1.049430 + 3.083787) GB / 131072 trees)
NCU_DRAM_BYTES_PER_TREE = 31534
NCU_DRAM_SOURCE = "profiles/r01_decode_small_v7.summary.txt"



# ---------------------------------------------------------------------------------------------
# Mixed-size sweep (BASELINE.json configs[0], [2], [4]): `python bench.py --sweep [--sweep-out F]`.
# Not the bench contract line; it fills the per-size table in profiles/.  It lives here because
# bench.py's CPU-baseline leg is the one place outside tests/ that may run the oracle.
def run_sweep(args):
    import numpy as np
    import torch
    from oracle import p2v_oracle as O
    from phylo2vec_b200 import _capi
    lib = _capi.load()
    dev = torch.device("cuda:0")
    PEAK = measured_peaks()[0]["hbm_gbs"]

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

    sizes = [16, 32, 64, 100, 128, 256, 512, 1000, 2000, 4096, 16384, 65536, 100000]
    if args.sweep_quick:
        sizes = [100, 1000, 2000, 16384]
    nthreads = host_threads()
    rows = []
    for n in sizes:
        k = n - 1
        rec = {"n_leaves": n}
        # ---- decode / encode: B*k ~ 1e9 elements (BASELINE.json configs[4]), at most 10M trees ------
        B = int(min(10_000_000, max(8, 1_000_000_000 // k)))
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
        # ---- distances: B * n^2 * 8 <= 64 GB per GPU (one tree at n = 100 000: 80 GB) ----------------
        Bd = int(max(1, min(B, (64 << 30) // (n * n * 8))))
        if n * n * 8 > (100 << 30):
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
    with open(args.sweep_out, "w") as f:
        json.dump({"peak_hbm_gbs": PEAK, "host_threads": nthreads, "rows": rows}, f, indent=1)




def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="trees per GPU per step")
    ap.add_argument("--e2e-batch", type=int, default=0)
    ap.add_argument("--c3-trees", type=int, default=C3_TREES, help="trees of the C3 leg (whole job)")
    ap.add_argument("--sweep", action="store_true", help="per-size sweep instead of the bench line")
    ap.add_argument("--sweep-out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_sweep.json"))
    ap.add_argument("--sweep-quick", action="store_true")
    args = ap.parse_args()
    if args.sweep:
        run_sweep(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    # stdout carries the JSON line(s) only: libraries that write to file descriptor 1 on their own
    # (NCCL's version banner and NCCL_DEBUG output) are sent to stderr for the whole run
    sys.stdout.flush()
    _json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = _json_out
    try:
        main()
    finally:
        _json_out.flush()
