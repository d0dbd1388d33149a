#!/usr/bin/env python
"""Warp-stall summary of ONE kernel launch from an `ncu --set full --import-source on` report (compile with -lineinfo):

    python tools/ncu_stalls.py gpurun_out/prof_X.ncu-rep > profiles/..._stalls.txt

Prints the share of every stall reason over all sampled warps, the executed code regions with their sample counts, and
the hottest instructions with their dominant reasons: where the warps of a latency-bound kernel actually wait."""
import csv
import io
import subprocess
import sys


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print("# %s" % rows[0][1] if len(rows[0]) > 1 else rows[0])
    hdr, data = rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    smp = [int(r[idx["# Samples"]]) for r in data]
    exe = [int(r[idx["Instructions Executed"]]) for r in data]
    tot = sum(smp)
    print("# %d SASS instructions, %d warp samples, %d warp-instructions executed" % (len(data), tot, sum(exe)))
    agg = {c: sum(int(r[idx[c]]) for r in data) for c in stall}
    print("## stall reasons (all sampled warps, idle producer / MMA / exited warps included)")
    for c, v in sorted(agg.items(), key=lambda x: -x[1]):
        if v:
            print("%6.1f%%  %5d  %s" % (100.0 * v / tot, v, c))
    print("## executed code regions (SASS index range: samples, warp-instructions)")
    cur = None
    for i, (e, s) in enumerate(zip(exe, smp)):
        if e > 0:
            if cur is None:
                cur = [i, i, 0, 0]
            cur[1], cur[2], cur[3] = i, cur[2] + s, cur[3] + e
        elif cur is not None and i - cur[1] > 40:
            if cur[2] >= 10:
                print("  [%5d, %5d]  samples %5d  warp-instructions %9d" % tuple(cur))
            cur = None
    if cur and cur[2] >= 10:
        print("  [%5d, %5d]  samples %5d  warp-instructions %9d" % tuple(cur))
    print("## hottest instructions (index, SASS, samples, times executed, dominant reasons)")
    for i in sorted(sorted(range(len(data)), key=lambda i: -smp[i])[:top]):
        r = data[i]
        why = sorted(((c, int(r[idx[c]])) for c in stall if int(r[idx[c]]) > 0), key=lambda x: -x[1])[:3]
        print("%6d  %-58s %5d %8d  %s" % (i, r[idx["Source"]].strip()[:58], smp[i], exe[i], ", ".join("%s %d" % w for w in why)))


if __name__ == "__main__":
    main(sys.argv[1])
