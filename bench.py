#!/usr/bin/env python3
"""Headline benchmark: simulated read pairs per second (2x150) for BASELINE.json config 2
("E. coli 4.6 Mbp ref, precomputed stats, 30x coverage 2x150 on 1xB200").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One step = one pass of the hot path over the whole workload: systematic-error drawing for both strands,
the (position x fragment length) scan, read generation and the ordered FASTQ gather.
  value : pairs / device time of those kernels (CUDA events: bias normalisation as far as it is not hidden behind the systematic
          errors, systematic errors, simulate, gather), inputs (reference bases, probability tables) resident in HBM
  e2e   : pairs / wall time of the C-ABI calls prepare + simulate + download with HOST buffers in and out
          (reference bases host->device, FASTQ text device->pinned host, every step)
N > 1 (torchrun): weak scaling - the reference grows to N sequences of the C2 length (N x 4.64 Mbp, 30x), whose SimBlocks are split
into N contiguous shards, one per rank/GPU (per-GPU work fixed); no collective on the data path, only pair counts and times
cross ranks (NCCL all-reduce).  Every rank still runs the prologue (normalisation, master stream, systematic errors) for the
whole reference.
--impl reference: the reference's own CPU Simulator (oracle/_ref/reseq_oracle, built from the unmodified sources)
with all host threads, each step on a bounded slice of the same workload.
"""
import argparse
import json
import lzma
import os
import re
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_LEN = 4_641_652          # E. coli K-12 MG1655 length; synthetic sequence (the real FASTA does not travel)
COVERAGE = 30.0
SEED = 42
BYTES_PER_PAIR = 1420.0      # 740 B FASTQ out + 600 B systematic errors in + 80 B reference in
BYTES_PER_POSITION = 8.25    # 0.25 B reference + 2 strands x 2 B systematic errors written and read once
ORACLE = os.path.join(ROOT, "oracle", "_ref", "reseq_oracle")
# profile150r: synthetic 2x150 profile fitted by the reference's own stats + IPF code from a synthetic SAM with a realistic InDel rate
# (~5e-5 per base); RSQ_BENCH_PROFILE=profile150 selects the InDel stress profile of the parity tests (1.6e-3 per base)
PROFILE = os.environ.get("RSQ_BENCH_PROFILE", "profile150r")


def unxz(name, tmp):
    dst = os.path.join(tmp, name[:-3])
    if not os.path.exists(dst):
        with lzma.open(os.path.join(GOLDEN, name)) as f, open(dst, "wb") as o:
            o.write(f.read())
    return dst


def workload_sequence(length=REF_LEN):
    import make_synthetic
    return make_synthetic.gen_reference([length], 1234)[0]


def workload_sequences(n, length=REF_LEN):
    """N sequences of the C2 length; the first one is the N=1 workload."""
    import make_synthetic
    return make_synthetic.gen_reference([length] * n, 1234)


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split("\n")[0]
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def shard_range(n_blocks, shard_index, shard_count):
    """Block range an engine simulates for (shard_index, shard_count): mirrors prepare() in csrc/engine.cu."""
    first = n_blocks * shard_index // shard_count
    return first, n_blocks * (shard_index + 1) // shard_count - first


def aggregate(dist, world, maxima, sums, device):
    """Max over ranks of the timings, sum over ranks of the counts (the only cross-rank traffic of the path)."""
    import torch
    t = torch.tensor(maxima, dtype=torch.float64, device=device)
    s = torch.tensor(sums, dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return t.tolist(), s.tolist()


def time_reference_cpu(length, coverage, threads, tmp):
    """Runs the reference binary on `length` bases of the workload; returns (pairs, seconds of read generation, total seconds)."""
    prof = unxz(PROFILE + ".reseq.xz", tmp)
    unxz(PROFILE + ".reseq.ipf.xz", tmp)
    fa = os.path.join(tmp, f"slice_{length}.fa")
    if not os.path.exists(fa):
        seq = workload_sequence()[:length]
        with open(fa, "w") as f:
            f.write(">ecoli_sized synthetic\n")
            for i in range(0, len(seq), 80):
                f.write(seq[i:i + 80] + "\n")
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else tmp
    r1, r2 = os.path.join(shm, f"rsq_ref_{os.getpid()}_1.fq"), os.path.join(shm, f"rsq_ref_{os.getpid()}_2.fq")
    cmd = [ORACLE, "illuminaPE", "-j", str(threads), "-s", prof, "-R", fa, "--ipfIterations", "0", "--seed", str(SEED),
           "-c", str(coverage), "-1", r1, "-2", r2]
    t0 = time.perf_counter()
    t_gen = None
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    for line in proc.stdout:
        if "Starting read generation" in line:     # Simulator.cpp:2828
            t_gen = time.perf_counter()
    proc.wait()
    t1 = time.perf_counter()
    if proc.returncode:
        raise RuntimeError("reference binary failed")
    with open(r1, "rb") as f:
        pairs = sum(chunk.count(b"\n") for chunk in iter(lambda: f.read(1 << 24), b"")) // 4
    os.remove(r1)
    os.remove(r2)
    return pairs, t1 - (t_gen or t0), t1 - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if not os.path.exists(ORACLE):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/reseq_oracle is not built on this box"}))
        return 0
    tmp = tempfile.mkdtemp(prefix="rsq_bench_ref_")
    cores = os.cpu_count() or 1
    sample_len = 1_500_000
    rates, secs = [], []
    for i in range(args.warmup + args.steps):
        pairs, gen_s, total_s = time_reference_cpu(sample_len, COVERAGE, cores, tmp)
        if i >= args.warmup:
            rates.append(pairs / gen_s)
            secs.append(gen_s)
    value = statistics.mean(rates)
    sample = f"first {sample_len} bp of the workload reference at {COVERAGE}x, -j {cores}, FASTQ to tmpfs; read-generation interval " \
             "(log line 'Starting read generation' to exit) - the scan cost is per position, so pairs/s carries over to the full genome"
    print(json.dumps({
        "impl": "reference", "metric": "simulated read-pairs/s (2x150)", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * statistics.mean(secs), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C2: 4,641,652 bp synthetic E. coli-sized reference, 2x150 synthetic profile ({PROFILE}), 30x coverage, seed 42"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    return 0


def run_b200(args):
    import torch
    import torch.distributed as dist
    import reseq_b200 as rb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    # stdout carries exactly one JSON line: NCCL prints its version banner (and torchrun children anything else) straight to file
    # descriptor 1, so the descriptor is pointed at stderr for the whole run and the line goes to a saved copy of the original
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    tmp = tempfile.mkdtemp(prefix="rsq_bench_")
    prof = rb.Profile.load_flat(unxz(PROFILE + ".flat.xz", tmp))
    seqs = [q.encode() for q in workload_sequences(world)]
    names = ["ecoli_sized synthetic"] + [f"ecoli_sized{i + 1} synthetic" for i in range(1, world)]
    eng = rb.Engine(prof, local_rank)

    def step():
        # host buffers in (reference bases), host buffers out (FASTQ text in pinned memory)
        ref = rb.Reference.from_memory(names, seqs)
        eng.prepare(ref, seed=SEED, coverage=COVERAGE, shard_index=rank, shard_count=world)
        eng.simulate()
        rep = eng.download()
        return rep.as_dict()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    reps = [step() for _ in range(args.steps)]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()

    dev_ms = sum(r["ms_bias"] + r["ms_syserr"] + r["ms_simulate"] + r["ms_gather"] for r in reps)
    sim_ms = sum(r["ms_simulate"] for r in reps)
    pairs = sum(r["pairs"] for r in reps)
    positions = sum(r["positions"] for r in reps)
    launches = sum(r["kernel_launches"] for r in reps)
    d2h = sum(r["bytes"][0] + r["bytes"][1] for r in reps)
    (dev_ms, wall, sim_ms_max), (pairs_all, positions_all, launches_all, d2h_all) = aggregate(
        dist, world, [dev_ms, wall, sim_ms], [pairs, positions, launches, d2h], "cuda")

    if rank == 0:
        peak, peak_src = hbm_peak()
        # roofline of the dominant phase on rank 0: algorithmic bytes of its launches / their event time
        alg_bytes = pairs * BYTES_PER_PAIR + positions * BYTES_PER_POSITION
        achieved = alg_bytes / (sim_ms / 1000.0) / 1e9 if sim_ms else 0.0
        spec = reps[-1]["spec_rounds"] > 0
        kernel = "k_spec_scan + k_spec_reads (all rounds of a step)" if spec else "k_simulate"
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "simulate_dram_bytes.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath))["spec" if spec else "serial"]["dram_bytes_per_step"]
            except Exception:
                traffic = None
        line = {
            "metric": "simulated read-pairs/s (2x150)", "value": pairs_all / (dev_ms / 1000.0), "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C2: 4,641,652 bp synthetic E. coli-sized reference, 2x150 synthetic profile ({PROFILE}), 30x coverage, seed 42"
                                   + (f"; x{world} sequences of that length for {world} GPUs (weak scaling: one sequence's worth of SimBlocks per GPU)" if world > 1 else ""),
                       "simulation_path": f"speculative two-phase, {reps[-1]['spec_rounds']} rounds, first depth {reps[-1]['spec_depth']}" if spec else "serial (one warp per SimBlock)",
                       "l2": "per-step working set (reference 4.6 MB + 2x9.3 MB systematic errors + 74 MB surroundings + ~340 MB FASTQ arena) exceeds the 126 MB L2; "
                             "every step re-uploads the reference and rewrites all of it",
                       "pairs_per_step": pairs_all / args.steps, "blocks_per_step": sum(r["blocks"] for r in reps) / args.steps,
                       "device_ms_breakdown_rank0": {k: sum(r[k] for r in reps) / args.steps for k in ("ms_upload", "ms_bias", "ms_syserr", "ms_simulate", "ms_gather", "ms_download")},
                       "scan_draws_per_s": sum(r["scan_draws"] for r in reps) / (sim_ms / 1000.0) if sim_ms else None},
            "e2e": {"value": pairs_all / wall, "unit": "pairs/s", "h2d_bytes_per_step": world * sum(len(q) for q in seqs), "d2h_bytes_per_step": d2h_all / args.steps,
                    "ms_per_step": 1000 * wall / args.steps},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "note": "algorithmic bytes = pairs x 1420 B + positions x 8.25 B over the event time of the simulate phase (every launch of the two kernels of a "
                                 "step); the phase is bound by the per-SimBlock serial mt19937_64 stream and dependent FP64 Draw chains (latency), not by HBM; "
                                 "measured DRAM traffic exceeds the algorithmic bytes because every read's slice of the stream (3.9 KB) goes through HBM/L2 once"},
        }
        if world == 1 and os.path.exists(ORACLE) and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            p, gen_s, _ = time_reference_cpu(REF_LEN, COVERAGE, cores, tmp)
            line["cpu_baseline"] = {"value": p / gen_s, "unit": "pairs/s", "cores": cores, "kind": "reference",
                                    "sample": f"reference binary (unmodified sources) on the whole workload ({REF_LEN} bp at {COVERAGE}x), -j {cores}, FASTQ to tmpfs, "
                                              f"read-generation interval {gen_s:.1f} s ({p} pairs)"}
        os.write(result_fd, (json.dumps(line) + "\n").encode())
    os.close(result_fd)
    if world > 1:
        dist.destroy_process_group()
    eng.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
