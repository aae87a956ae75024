#!/usr/bin/env python3
"""Turns the raw ncu artefacts of a gpurun call (gpurun_out/) into the small tracked summaries under profiles/.

    python tools/summarize_profiles.py r01c gpurun_out/launches_r01c.csv gpurun_out/bench_r01c.json \
        [--full k_spec_reads=gpurun_out/prof_r01c_k_spec_reads.ncu-rep ...] [--serial gpurun_out/serial_r01c.csv]

  <tag>_launches.md                 every kernel of one bench step: launches, total time, share, instructions, DRAM bytes
  <tag>_<kernel>_ncu_full.md        selected metrics of one `ncu --set full` capture of that kernel
  <tag>_<kernel>_by_source_line.txt warp instructions / stall samples per source line (tools/ncu_by_line.py)
  <tag>_bench.json                  the bench line of the same build
  simulate_dram_bytes.json          measured DRAM traffic of the simulate phase per step (bench.py reports it as roofline.traffic)
"""
import csv
import json
import os
import subprocess
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3, "inst": 1}


def read_launch_csv(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    launches = OrderedDict()
    for row in csv.DictReader(lines):
        d = launches.setdefault(row["ID"], {"name": row["Kernel Name"].split("(")[0].replace("void ", "")})
        d[row["Metric Name"]] = float(row["Metric Value"].replace(",", "")) * SCALE.get(row["Metric Unit"], 1)
    return list(launches.values())


def launches_md(tag, path):
    ls = read_launch_csv(path)
    # the command is `bench.py --steps 1 --warmup 1`: the second half of the launches is the timed step
    step = ls[len(ls) // 2:]
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for l in step:
        a = agg[l["name"]]
        a[0] += 1
        a[1] += l.get("gpu__time_duration.sum", 0.0)
        a[2] += l.get("smsp__inst_executed.sum", 0.0)
        a[3] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
    total = sum(v[1] for v in agg.values())
    out = os.path.join(ROOT, "profiles", f"{tag}_launches.md")
    with open(out, "w") as f:
        f.write(f"# {tag}: every kernel launch of the timed step of `bench.py --steps 1 --warmup 1` under ncu\n\n"
                "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none`.\n"
                "Launches are serialised and cold-cache under ncu (the two stream groups of the simulate phase and the bias sums on their own stream\n"
                "overlap in a normal run): compare shares, not absolutes.\n\n"
                "| kernel | launches | total ms | share | warp instructions | DRAM bytes |\n|---|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| {k} | {v[0]} | {v[1]:.3f} | {100 * v[1] / total:.1f}% | {v[2]:.3e} | {v[3]:.3e} |\n")
    spec = sum(v[3] for k, v in agg.items() if k.startswith("k_spec_"))
    return out, spec


def raw_metrics(tag, rep, kernel):
    csv_text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(csv_text.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    picked = {h: (v, u) for h, u, v in zip(hdr, units, vals) if h in KEEP}
    stalls = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): v for h, v in zip(hdr, vals) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
    out = os.path.join(ROOT, "profiles", f"{tag}_{kernel}_ncu_full.md")
    with open(out, "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on` of one {kernel} launch in the middle of a bench step (C2 workload)\n\n| metric | value | unit |\n|---|---|---|\n")
        for h in KEEP:
            if h in picked:
                f.write(f"| {h} | {picked[h][0]} | {picked[h][1]} |\n")
        f.write("\nWarp stall samples (pc sampling): " + ", ".join(f"{k} {v}" for k, v in sorted(stalls.items(), key=lambda x: -float(x[1] or 0)) if float(v or 0) > 0) + "\n")
    sass = os.path.join("/tmp", f"{tag}_{kernel}_sass.csv")
    with open(sass, "w") as f:
        f.write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout)
    byline = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_line.py"), sass, os.path.join(ROOT, "reseq_b200", "libreseq_b200.so"), kernel, "45"],
                            capture_output=True, text=True).stdout
    with open(os.path.join(ROOT, "profiles", f"{tag}_{kernel}_by_source_line.txt"), "w") as f:
        f.write(f"# {tag}: warp instructions / stall samples of {kernel} attributed to source lines (tools/ncu_by_line.py)\n" + byline)
    return out


def main():
    tag, launch_csv, bench_json = sys.argv[1:4]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    out, spec_bytes = launches_md(tag, launch_csv)
    print(out)
    traffic = {"spec": {"dram_bytes_per_step": spec_bytes, "source": os.path.basename(out),
                        "what": "dram__bytes_read.sum + dram__bytes_write.sum over every k_spec_* launch of one bench step"}}
    args = sys.argv[4:]
    i = 0
    while i < len(args):
        if args[i] == "--full":
            kernel, rep = args[i + 1].split("=")
            print(raw_metrics(tag, rep, kernel))
            i += 2
        elif args[i] == "--serial":
            ls = read_launch_csv(args[i + 1])
            traffic["serial"] = {"dram_bytes_per_step": sum(l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0) for l in ls),
                                 "ms": sum(l.get("gpu__time_duration.sum", 0) for l in ls), "source": "k_simulate (RSQ_SIM_PATH=serial), one launch"}
            i += 2
        else:
            i += 1
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "simulate_dram_bytes.json"), "w"), indent=1)
    json.dump(json.load(open(bench_json)), open(os.path.join(ROOT, "profiles", f"{tag}_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
