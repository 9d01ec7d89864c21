#!/usr/bin/env python
"""Hottest SASS lines of a kernel in an `ncu --set full --import-source on` report (stall samples per instruction):
    python tools/ncu_hot.py gpurun_out/prof.ncu-rep [min_pct]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = None
    data = []
    for r in rows:
        if r and r[0] == "Kernel Name":
            print("====", r[1][:100])
            continue
        if r and r[0] == "Address":
            hdr = {h: i for i, h in enumerate(r)}
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        try:
            data.append((int(r[hdr["# Samples"]]), int(r[hdr["Instructions Executed"]]), r[hdr["Source"]].strip()))
        except ValueError:
            pass
    tot = sum(d[0] for d in data) or 1
    print("total samples", tot, "instructions", sum(d[1] for d in data))
    for i, (n, ex, s) in enumerate(data):
        if 100.0 * n / tot >= min_pct:
            print("%5d %6d %5.1f%% x%-8d %s" % (i, n, 100.0 * n / tot, ex, s[:120]))


if __name__ == "__main__":
    main()
