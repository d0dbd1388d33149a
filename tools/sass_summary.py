#!/usr/bin/env python
"""Per-kernel SASS opcode summary of the built library (no GPU needed):

    python tools/sass_summary.py tfkaldi_b200/libtfkaldi_b200.so > profiles/r2_sass_opcodes.txt

For every kernel: instruction count and the counts of the opcodes that identify the hardware paths used — tcgen05
(UTC*MMA, LDTM/STTM, UTCBAR), TMA (UTMALDG / UTMASTG / UTMAREDG / UTMAPF / UTMACCTL), mbarrier (SYNCS), programmatic dependent
launch (ACQBULK / PREEXIT-style griddepcontrol show as these), legacy tensor path (HMMA: must be absent), vector width of
global accesses."""
import collections
import re
import subprocess
import sys

KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UTMACCTL", "UTMACMDFLUSH", "SYNCS",
        "HMMA", "LDG.E.128", "LDG.E.64", "LDG.E", "STG.E.128", "STG.E.64", "STG.E", "LDS", "STS", "SHFL", "MUFU", "REDG", "ATOMG", "RED.", "ATOM",
        "CCTL", "MEMBAR", "ERRBAR", "FENCE", "BAR.SYNC", "ACQBULK", "PREEXIT", "UCGABAR", "IMAD.WIDE", "LDC"]


def main(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::|tfk::", "", name)
            name = re.sub(r"\(.*", "", name)
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            cur[op] += 1
    print("# cuobjdump -sass %s : opcode counts per kernel (sm_100a)" % path)
    for name, c in kernels.items():
        picks = []
        for k in KEYS:
            n = sum(v for op, v in c.items() if op == k or op.startswith(k + ".")) if not k.endswith(".") else sum(v for op, v in c.items() if op.startswith(k))
            if k == "UTCHMMA":
                n = sum(v for op, v in c.items() if op.startswith("UTCHMMA") and ".2CTA" not in op)
            if k == "UTCHMMA.2CTA":
                n = sum(v for op, v in c.items() if op.startswith("UTCHMMA") and ".2CTA" in op)
            if k in ("LDG.E", "STG.E"):
                n = sum(v for op, v in c.items() if op.startswith(k) and ".128" not in op and ".64" not in op)
            if n:
                picks.append("%s %d" % (k, n))
        print("%-40s %6d instr | %s" % (name[:40], c["_total"], ", ".join(picks)))


if __name__ == "__main__":
    main(sys.argv[1])
