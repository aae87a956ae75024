#!/usr/bin/env python3
"""Turns the raw ncu artefacts of a gpurun call (gpurun_out/) into the small tracked summaries under profiles/.

    python tools/summarize_profiles.py r01 gpurun_out/launches.csv gpurun_out/prof_k_simulate.ncu-rep gpurun_out/bench.json
"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic"]


def launches(tag, path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        val = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        ms = val / 1e6 if unit == "ns" else val / 1e3 if unit.startswith("us") else val
        agg[name][0] += 1
        agg[name][1] += ms
    total = sum(v[1] for v in agg.values())
    out = os.path.join(ROOT, "profiles", f"{tag}_launches.md")
    with open(out, "w") as f:
        f.write(f"# {tag}: every kernel launch of `bench.py --steps 1 --warmup 1` under ncu (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("Cold-cache, serialised launches: compare shares, not absolutes.\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| {k} | {v[0]} | {v[1]:.3f} | {100 * v[1] / total:.1f}% |\n")
    return out


def raw_metrics(tag, rep, kernel="k_simulate"):
    csv_text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(csv_text.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    picked = {h: (v, u) for h, u, v in zip(hdr, units, vals) if h in KEEP}
    out = os.path.join(ROOT, "profiles", f"{tag}_{kernel}_ncu_full.md")
    with open(out, "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none` of {kernel} (one launch of the C2 workload)\n\n| metric | value | unit |\n|---|---|---|\n")
        for h in KEEP:
            if h in picked:
                f.write(f"| {h} | {picked[h][0]} | {picked[h][1]} |\n")
    try:
        rd = float(picked["dram__bytes_read.sum"][0].replace(",", ""))
        wr = float(picked["dram__bytes_write.sum"][0].replace(",", ""))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        total = rd * scale[picked["dram__bytes_read.sum"][1]] + wr * scale[picked["dram__bytes_write.sum"][1]]
        json.dump({"kernel": kernel, "dram_bytes_per_launch": total, "source": os.path.basename(out)},
                  open(os.path.join(ROOT, "profiles", f"{kernel}_dram_bytes.json"), "w"))
    except Exception as e:   # noqa: BLE001
        print("could not derive dram traffic:", e)
    sass = os.path.join("/tmp", f"{tag}_sass.csv")
    with open(sass, "w") as f:
        f.write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout)
    byline = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_line.py"), sass, os.path.join(ROOT, "reseq_b200", "libreseq_b200.so"), kernel, "45"],
                            capture_output=True, text=True).stdout
    with open(os.path.join(ROOT, "profiles", f"{tag}_{kernel}_by_source_line.txt"), "w") as f:
        f.write(f"# {tag}: warp instructions / stall samples of {kernel} attributed to source lines (tools/ncu_by_line.py)\n" + byline)
    return out


def main():
    tag = sys.argv[1]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    print(launches(tag, sys.argv[2]))
    print(raw_metrics(tag, sys.argv[3]))
    if len(sys.argv) > 4:
        line = json.load(open(sys.argv[4]))
        json.dump(line, open(os.path.join(ROOT, "profiles", f"{tag}_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
