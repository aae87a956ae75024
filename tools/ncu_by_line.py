#!/usr/bin/env python3
"""Attribute an ncu per-instruction SASS export to source lines.

    ncu -i prof.ncu-rep --page source --csv --print-source sass > sass.csv
    python tools/ncu_by_line.py sass.csv reseq_b200/libreseq_b200.so k_simulate [top]

Joins the i-th SASS instruction of the ncu export with the i-th instruction of `nvdisasm --print-line-info`
for the same kernel of the same build (ncu's CSV has no line column) and prints, per source line, executed warp
instructions and stall samples.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def disasm_lines(so, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    out = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    lines, cur, active = [], ("?", 0), False
    for l in out.split("\n"):
        if l.startswith("//---") and ".text." in l:
            active = kernel in l
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
            lines.append(cur)
    return lines


def main():
    sass_csv, so, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(sass_csv)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    lines = disasm_lines(so, kernel)
    if len(lines) != len(data):
        print(f"warning: {len(data)} instructions in ncu export vs {len(lines)} in nvdisasm - build mismatch?", file=sys.stderr)
    agg = defaultdict(lambda: [0.0, 0.0])
    tot_i = tot_s = 0.0
    for r, ln in zip(data, lines):
        try:
            inst, smp = float(r[ix["Instructions Executed"]] or 0), float(r[ix["# Samples"]] or 0)
        except ValueError:
            continue
        agg[ln][0] += inst
        agg[ln][1] += smp
        tot_i += inst
        tot_s += smp
    print(f"total warp instructions {tot_i:.4g}, samples {tot_s:.0f}")
    byfile = defaultdict(lambda: [0.0, 0.0])
    for (f, _), (i, s) in agg.items():
        byfile[f][0] += i
        byfile[f][1] += s
    for f, (i, s) in sorted(byfile.items(), key=lambda x: -x[1][1]):
        print(f"{f:20s} inst {100 * i / tot_i:5.1f}%  samples {100 * s / tot_s:5.1f}%")
    src_cache = {}
    for (f, n), (i, s) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        path = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", f)
        if path not in src_cache:
            src_cache[path] = open(path).read().split("\n") if os.path.exists(path) else []
        text = src_cache[path][n - 1].strip()[:90] if 0 < n <= len(src_cache[path]) else ""
        print(f"{f}:{n:<5d} inst {100 * i / tot_i:5.2f}%  samples {100 * s / tot_s:5.2f}%  | {text}")


if __name__ == "__main__":
    main()
