#!/usr/bin/env python
"""Aggregates an ncu SASS source page (ncu -i X.ncu-rep --page source --csv --kernel-name K) by CUDA source line.

usage: ncu_lines.py <object.o> <mangled-or-substring kernel name> <source_page.csv> [top N]
The address -> line map comes from `nvdisasm --print-line-info` on the cubin embedded in the object (needs -lineinfo).
Inlined code is attributed to the innermost source line.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def line_map(obj, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    amap = {}
    in_fn = False
    cur = None
    for ln in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            in_fn = kernel in m.group(1)
            continue
        if not in_fn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            amap[int(m.group(1), 16)] = cur
    return amap


def main():
    obj, kernel, page = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    amap = line_map(obj, kernel)
    rows = list(csv.reader(open(page)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    ia, ii, it, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    base = None
    agg = defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r in rows[h + 1:]:
        if r and r[0] in ("Kernel Name", "Address"):
            break  # next kernel of a multi-kernel page
        if len(r) <= it or not r[ia]:
            continue
        a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
        if base is None:
            base = a
        key = amap.get(a - base, ("?", 0))
        vals = [int(float(r[ii] or 0)), int(float(r[it] or 0)), int(float(r[isamp] or 0))]
        for k in range(3):
            agg[key][k] += vals[k]
            tot[k] += vals[k]
    print(f"total warp-instr {tot[0]}, thread-instr {tot[1]}, samples {tot[2]}")
    srcs = {}
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in srcs:
            p = os.path.join(os.path.dirname(os.path.abspath(obj)), f)
            srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
        print(f"{100.0 * v[0] / max(tot[0], 1):5.1f}% inst {100.0 * v[2] / max(tot[2], 1):5.1f}% smp  eff {v[1] / max(v[0], 1):4.1f}  {f}:{l}  {text}")


if __name__ == "__main__":
    main()
