#!/usr/bin/env python
"""DRAM traffic per launch of tfk_gemm2_kernel from an `ncu --set full` capture of `bench.py --steps 3 --warmup 3`
(one whole step = 14 consecutive GEMM launches), in the JSON format bench.py's `roofline.traffic` reads:

    python tools/ncu_traffic.py gpurun_out/prof_c2_TAG.ncu-rep c2:bf16:8192 "<source note>" > profiles/r2_ncu_traffic.json
"""
import csv
import io
import json
import subprocess
import sys


def main(path, key, source):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = []
    for r in data:
        if "tfk_gemm2_kernel" not in r[idx["Kernel Name"]]:
            continue
        b = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(r[idx[m]].replace(",", "")) * scale.get(units[idx[m]], 1)
        per.append(b)
    # the capture may start mid-step: take the first 14 launches from the first layer-0 forward (the smallest launch
    # that is followed by a hidden forward); with only 14+ launches captured any window of 14 is one step's worth
    n = 14
    if len(per) < n:
        raise SystemExit("only %d GEMM launches captured" % len(per))
    window = per[:n]
    res = {key: {"kernel": "tfk_gemm2_kernel", "launches_per_step": n, "dram_mb_per_launch": [round(b / 1e6, 1) for b in window],
                 "dram_bytes_per_step": int(sum(window)), "dram_bytes_per_launch": int(sum(window) / n), "source": source}}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
