#!/usr/bin/env python
"""Per-source-line stall samples of one kernel of an .ncu-rep (needs -lineinfo and --import-source on).

    python profiles/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [launch_index] [top]
"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--launch-skip",
                      str(skip), "--launch-count", "1", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fname = None
hdr = None
cur = None
agg = {}
for row in csv.reader(io.StringIO(raw)):
    if not row:
        continue
    if row[0] == "File Name":
        fname = row[1].split("/")[-1]
        continue
    if row[0] == "Line No":
        hdr = row
        continue
    if hdr is None or len(row) != len(hdr):
        continue
    if row[0]:
        cur = (fname, int(row[0]), row[1].strip())
        agg.setdefault(cur, [0, 0, {}])
    if cur is None or row[2] in ("", "..."):
        continue
    rec = dict(zip(hdr[4:], row[4:]))
    try:
        ns = int(rec["# Samples"])
    except (KeyError, ValueError):
        continue
    a = agg[cur]
    a[0] += ns
    a[1] += int(rec.get("Instructions Executed", "0") or 0)
    for k, v in rec.items():
        if k.startswith("stall_") and "(Not Issued)" not in k and v not in ("", "0", "-"):
            a[2][k] = a[2].get(k, 0) + int(v)
tot = sum(a[0] for a in agg.values()) or 1
print(f"total samples {tot}")
for (f, ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = sorted(a[2].items(), key=lambda kv: -kv[1])[:2]
    print(f"{100.0 * a[0] / tot:5.1f}%  {f}:{ln:<5d} inst={a[1]:<9d} {' '.join(f'{k[6:]}={v}' for k, v in st):32s} | {src[:100]}")
