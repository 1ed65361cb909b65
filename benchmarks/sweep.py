#!/usr/bin/env python
"""Mixed-size sweep (BASELINE.json configs[0], [2], [4]); the code lives in bench.py (`--sweep`),
the one place outside tests/ that may run the CPU oracle.

    python benchmarks/sweep.py [--out profiles/r01_sweep.json] [--quick]
"""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_sweep.json"))
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--sweep", "--sweep-out", a.out]
    if a.quick:
        cmd.append("--sweep-quick")
    sys.exit(subprocess.call(cmd))
