#!/usr/bin/env python
"""Per-source-line totals from an .ncu-rep captured with --import-source on (kernels compiled with -lineinfo).
usage: python tools/ncu_lines.py rep.ncu-rep [top_n]   -> line, samples, instructions executed, smem wavefronts"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = None
    cur = None
    agg = {}
    fname = ""
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) - 2:
            continue
        if r[0] != "":
            cur = (fname, int(r[0]), r[1].strip())
            agg.setdefault(cur, [0, 0, 0])
            continue
        if cur is None or r[2] in ("...", "-"):
            continue
        def col(name):
            try:
                return float(r[hdr.index(name)] or 0)
            except (ValueError, IndexError):
                return 0.0
        a = agg[cur]
        a[0] += col("# Samples")
        a[1] += col("Instructions Executed")
        a[2] += col("L1 Wavefronts Shared")
    tot = [sum(v[i] for v in agg.values()) for i in range(3)]
    print(f"total samples {tot[0]:.0f}  warp-instructions {tot[1]:.0f}  smem wavefronts {tot[2]:.0f}")
    for (f, ln, src), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{f}:{ln:4d} samp {v[0]:7.0f} ({100*v[0]/max(tot[0],1):4.1f}%) inst {v[1]:10.0f} ({100*v[1]/max(tot[1],1):4.1f}%) wf {v[2]:10.0f} | {src[:90]}")


if __name__ == "__main__":
    main()
