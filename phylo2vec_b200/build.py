"""Builds libphylo2vec_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m phylo2vec_b200.build [--force] [--verbose] [--debug-bounds]

--debug-bounds also builds lib/libphylo2vec_b200_dbg.so: the decode / encode kernels compiled with
-DP2V_DEBUG_BOUNDS (every shared-memory table index asserted, see p2v_common.cuh), linked with the ordinary
objects of the other sources; tests/test_gpu_debug_bounds.py runs the boundary sizes through it.

The shared object stays inside the package (phylo2vec_b200/lib/), so it travels
with the source tree to the GPU box; there is no JIT cache and no fallback.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libphylo2vec_b200.so")
LIB_DBG = os.path.join(LIBDIR, "libphylo2vec_b200_dbg.so")
DBG_SOURCES = ["p2v_decode.cu", "p2v_encode.cu", "p2v_cophenetic.cu"]      # the kernels with scattered shared-memory tables
SOURCES = ["p2v_capi.cu", "p2v_decode.cu", "p2v_encode.cu", "p2v_cophenetic.cu", "p2v_newick.cu", "p2v_treewise.cu", "p2v_newick_matrix.cu", "p2v_csv.cu"]
HEADERS = ["p2v_common.cuh", "p2v_internal.h", "p2v_decode_warp.cuh", "p2v_decode_warp_big.cuh", "p2v_encode_light.cuh", "p2v_tree_pass.cuh", "p2v_float_text.cuh", "p2v_float_tables.h", os.path.join("..", "..", "include", "phylo2vec_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def _stale_dbg() -> bool:
    if not os.path.exists(LIB_DBG):
        return True
    t = os.path.getmtime(LIB_DBG)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, debug_bounds: bool = False) -> str:
    """Every source is compiled to its own object file (in parallel, only the stale ones) and the
    objects are linked into the shared library.  debug_bounds: also libphylo2vec_b200_dbg.so -- DBG_SOURCES
    compiled once more with -DP2V_DEBUG_BOUNDS (in the same pool), linked with the ordinary objects of the rest."""
    need_main = force or _stale()
    need_dbg = debug_bounds and (force or need_main or _stale_dbg())
    if not need_main and not need_dbg:
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    extra = os.environ.get("P2V_NVCC_EXTRA", "").split()
    tag = "_".join(extra).replace("/", "_").replace("=", "_")
    hdr_t = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)
    hdr_t = max(hdr_t, os.path.getmtime(os.path.abspath(__file__)))
    nvcc = _nvcc()
    compile_flags = [f for f in NVCC_FLAGS if f != "--shared"]

    def one(job):
        src, dbg = job
        suffix = ".dbg" if dbg else ((("." + tag) if tag else ""))
        obj = os.path.join(objdir, os.path.splitext(src)[0] + suffix + ".o")
        path = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(hdr_t, os.path.getmtime(path)):
            return obj, None
        flags = compile_flags + (["-DP2V_DEBUG_BOUNDS"] if dbg else extra)
        cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, path]
        if verbose:
            print(" ".join(cmd))
        return obj, subprocess.run(cmd, capture_output=True, text=True)

    jobs = [(s, False) for s in SOURCES] + ([(s, True) for s in DBG_SOURCES] if need_dbg else [])
    with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
        results = list(ex.map(one, jobs))
    for obj, r in results:
        if r is None:
            continue
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed building libphylo2vec_b200.so")
    main_objs = [o for (o, _), (_, dbg) in zip(results, jobs) if not dbg]
    if need_main:
        r = subprocess.run([nvcc, "--shared", "-o", LIB] + main_objs + ["-lpthread"], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed linking libphylo2vec_b200.so")
    if need_dbg:
        dbg_objs = [o for (o, _), (_, dbg) in zip(results, jobs) if dbg]
        rest = [o for o, (s, dbg) in zip(main_objs, [j for j in jobs if not j[1]]) if s not in DBG_SOURCES]
        r = subprocess.run([nvcc, "--shared", "-o", LIB_DBG] + dbg_objs + rest + ["-lpthread"], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed linking libphylo2vec_b200_dbg.so")
    return LIB


def build_debug_bounds(force: bool = False) -> str:
    build(force=force, debug_bounds=True)
    return LIB_DBG


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, debug_bounds="--debug-bounds" in sys.argv))
