#!/usr/bin/env python
"""Turn the two ncu outputs of tools/gpu_profile.sh into the text summaries kept under profiles/.

  python tools/ncu_summarize.py launches gpurun_out/launches_TAG.csv   > profiles/..._launch_list_summary.txt
  python tools/ncu_summarize.py full     gpurun_out/prof_TAG.ncu-rep   > profiles/..._full_summary.txt

`launches`: per-kernel share of the serialised, cold-cache launch list (compare SHARES with bench.py's timers).
`full`: one row per captured launch from `ncu --set full` (DRAM traffic, tensor-pipe activity, occupancy, clocks).
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("unnamed>::", "").replace("tfk::", "").replace("(anonymous namespace)::", "")
    return name.split("::")[-1].strip()


def launches(path):
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    agg = OrderedDict()
    total = 0.0
    n = 0
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        us = float(r["Metric Value"].replace(",", "")) / (1000.0 if r["Metric Unit"] == "ns" else 1.0)
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
        n += 1
    print("# ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --steps 3 --warmup 3 (C2, bf16), first 400 launches")
    print("# cold-cache, serialised per-launch times: compare SHARES, not absolutes")
    print("launches %d total_us %.1f" % (n, total))
    for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%5.1f%%  n=%4d  avg=%8.1fus  %s" % (100.0 * us / total, c, us / c, k))


METRICS = [
    ("time_us", "gpu__time_duration.sum", 1e-3),
    ("dram_rd_MB", "dram__bytes_read.sum", None),
    ("dram_wr_MB", "dram__bytes_write.sum", None),
    ("dram_%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("tensor_%act", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", 1),
    ("warps_%", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    ("regs", "launch__registers_per_thread", 1),
    ("grid", "launch__grid_size", 1),
    ("L2hit_%", "lts__t_sector_hit_rate.pct", 1),
    ("L2_%", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("sm_%", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("GHz", "sm__cycles_elapsed.avg.per_second", None),
]


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * m.get(unit, 1)


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    tensor_alts = [h for h in hdr if "pipe_tensor" in h and "pct_of_peak_sustained_active" in h]
    print("# ncu --set full --clock-control none --import-source on, bench.py --steps 3 --warmup 3 (C2 bf16); -k tfk_gemm2|adam|softmax_ce")
    print("# one row per captured launch; replayed (cold-cache) numbers: use for traffic / pipe utilisation, not for timing")
    print("\t".join(["kernel"] + [m[0] for m in METRICS]))
    for r in data:
        cells = [short(r[idx["Kernel Name"]])]
        for label, name, scale in METRICS:
            if name not in idx and label == "tensor_%act" and tensor_alts:
                name = tensor_alts[0]
            if name not in idx:
                cells.append("-")
                continue
            v = r[idx[name]].replace(",", "")
            u = units[idx[name]]
            try:
                if label.endswith("_MB"):
                    cells.append("%.1f" % (to_bytes(v, u) / 1e6))
                elif label == "GHz":
                    f = float(v) * {"hz": 1e-9, "Khz": 1e-6, "Mhz": 1e-3, "Ghz": 1.0}.get(u, 1e-9)
                    cells.append("%.2f" % f)
                elif label == "time_us":
                    f = float(v) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
                    cells.append("%.1f" % f)
                elif label in ("regs", "grid"):
                    cells.append("%d" % float(v))
                else:
                    cells.append("%.1f" % float(v))
            except ValueError:
                cells.append(v)
        print("\t".join(cells))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
