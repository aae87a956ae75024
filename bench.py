#!/usr/bin/env python3
"""Headline benchmark: simulated read pairs per second (2x150).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c5]

N = 1 (default workload c2): BASELINE.json config 2 - the reference's own E. coli fixture (4 641 652 bp, committed xz-compressed), 2x150
profile150r, 30x, seed 42.  The engine's output for exactly this workload has the sha256 of the reference's `-j 1` run
(tests/test_gpu_parity.py::test_c1_real_ecoli_sequence_equals_the_reference).
N > 1 (torchrun; default workload c5): STRONG scaling on one fixed workload shaped like BASELINE.json config 5 at a third of its size - synthetic
1 Gbp reference in 8 sequences, phased diploid VCF (1 SNP per kb, 1 indel <= 20 bases per 10 kb), 30x, 2x150 - split over the N engines,
which are joined into a group (rsq_engine_join_group: NCCL over NVLink): each rank uploads and prepares only the sequences of its shard, the
bias sums of CalculateBiasNormalization are computed by the owner of a sequence and all-reduced, the pair counts are all-reduced.
The N = 1 line carries the same workload's single-GPU number as "strong_ref" (one step), the reference point of the strong-scaling ratios.

One step = one pass of the hot path over the whole workload: systematic-error drawing for both strands, the (position x fragment length)
scan, read generation and the ordered FASTQ gather.
  value : pairs / device time of those kernels (CUDA events: bias normalisation as far as it is not hidden behind the systematic
          errors, systematic errors, simulate, gather; max over ranks), inputs (reference bases, probability tables) resident in HBM
  e2e   : pairs / wall time of the C-ABI calls prepare + simulate + download with HOST buffers in and out
          (reference bases host->device, FASTQ text device->host, every step; max over ranks)
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
REF_LEN = 4_641_652          # E. coli K-12 MG1655 (the reference's test/ecoli-GCF_000005845.2_ASM584v2_genomic.fa, committed as tests/golden/ecoli_GCF_000005845.2.fa.xz)
ECOLI_XZ = os.path.join(GOLDEN, "ecoli_GCF_000005845.2.fa.xz")
C5_SEQS, C5_SEQ_LEN, C5_REF_SEED, C5_VCF_SEED = 8, 125_000_000, 4321, 77   # strong-scaling workload: 1 Gbp in 8 sequences + diploid VCF
ISSUE_PEAK = 148 * 4 * 1.965e9   # warp instructions per second: 148 SMs x 4 schedulers x 1.965 GHz (the ceiling the scan works against)
SCAN_INST_PER_DRAW = 3.41        # warp instructions of k_spec_scan per scan draw (profiles/r02y_launches.md: 13.05 G per step / 3.829 G draws)
READ_INST_PER_BASE = 68.7        # warp instructions of k_spec_reads per base and read (profiles/r02y_launches.md: 9.59 G per step / 930 758 reads / 150)
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


def workload_c2():
    """(names, sequences) of config C2: the real E. coli sequence when the fixture travelled, else a synthetic one of its length."""
    if os.path.exists(ECOLI_XZ):
        name, parts = None, []
        for line in lzma.open(ECOLI_XZ, "rt"):
            if line.startswith(">"):
                name = line[1:].rstrip("\n")
            else:
                parts.append(line.strip())
        return [name], ["".join(parts)], "the reference's E. coli K-12 fixture (NC_000913.3)"
    import make_synthetic
    return ["ecoli_sized synthetic"], make_synthetic.gen_reference([REF_LEN], 1234), "synthetic E. coli-sized reference"


def workload_c5(tmp, n_seqs=C5_SEQS, seq_len=C5_SEQ_LEN):
    """(names, sequences, vcf path) of the strong-scaling workload: config C5's shape at a third of its size."""
    import make_synthetic
    seqs = make_synthetic.gen_reference([seq_len] * n_seqs, C5_REF_SEED)
    names = [f"chr{i + 1}" for i in range(n_seqs)]
    vcf = os.path.join(tmp, f"c5_{n_seqs}x{seq_len}_{os.getpid()}.vcf")
    n_rec = make_synthetic.write_vcf(vcf, names, seqs, C5_VCF_SEED)
    return names, seqs, vcf, n_rec


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
    """Even split of n_blocks simulated SimBlocks; the engine moves a boundary onto a sequence start lying within 5 % of a shard's size
    (prepare() in csrc/engine.cu reports the range it really took: rsq_sim_report.shard_first / blocks; tests/test_gpu_multi.py compares)."""
    first = n_blocks * shard_index // shard_count
    return first, n_blocks * (shard_index + 1) // shard_count - first


def aggregate(dist, world, maxima, sums, device):
    """Max over ranks of the timings, sum over ranks of the counts."""
    import torch
    t = torch.tensor(maxima, dtype=torch.float64, device=device)
    s = torch.tensor(sums, dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return t.tolist(), s.tolist()


def workload_name(kind, source=""):
    if kind == "c2":
        return f"C2: 4,641,652 bp {source}, 2x150 synthetic profile ({PROFILE}), 30x coverage, seed 42"
    return (f"C5-shaped strong-scaling workload: synthetic {C5_SEQS * C5_SEQ_LEN / 1e9:g} Gbp reference in {C5_SEQS} sequences, phased diploid VCF "
            f"(1 SNP per kb, 1 indel <= 20 bases per 10 kb), 2x150 synthetic profile ({PROFILE}), 30x coverage, seed 42")


def time_reference_cpu(kind, length, coverage, threads, tmp):
    """Runs the reference binary on the first `length` bases of the workload; returns (pairs, seconds of read generation, total seconds)."""
    import make_synthetic
    prof = unxz(PROFILE + ".reseq.xz", tmp)
    unxz(PROFILE + ".reseq.ipf.xz", tmp)
    fa = os.path.join(tmp, f"slice_{kind}_{length}.fa")
    vcf = None
    if kind == "c2":
        names, seqs, _ = workload_c2()
        seq = seqs[0][:length]
        name = names[0]
    else:
        seq = make_synthetic.gen_reference([length], C5_REF_SEED)[0]
        name = "chr1"
        vcf = os.path.join(tmp, f"slice_{length}.vcf")
        make_synthetic.write_vcf(vcf, [name], [seq], C5_VCF_SEED)
    if not os.path.exists(fa):
        with open(fa, "w") as f:
            f.write(">" + name + "\n")
            for i in range(0, len(seq), 80):
                f.write(seq[i:i + 80] + "\n")
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else tmp
    r1, r2 = os.path.join(shm, f"rsq_ref_{os.getpid()}_1.fq"), os.path.join(shm, f"rsq_ref_{os.getpid()}_2.fq")
    cmd = [ORACLE, "illuminaPE", "-j", str(threads), "-s", prof, "-R", fa, "--ipfIterations", "0", "--seed", str(SEED),
           "-c", str(coverage), "-1", r1, "-2", r2] + (["-V", vcf] if vcf else [])
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
    kind = args.workload or ("c2" if args.gpus <= 1 else "c5")
    tmp = tempfile.mkdtemp(prefix="rsq_bench_ref_")
    cores = os.cpu_count() or 1
    # C2 fits the budget whole (about 6 s per step on 16 threads); the 1 Gbp strong-scaling workload is sampled (the scan cost is per position)
    sample_len = REF_LEN if kind == "c2" else 1_500_000
    rates, secs = [], []
    for i in range(args.warmup + args.steps):
        pairs, gen_s, total_s = time_reference_cpu(kind, sample_len, COVERAGE, cores, tmp)
        if i >= args.warmup:
            rates.append(pairs / gen_s)
            secs.append(gen_s)
    value = statistics.mean(rates)
    sample = (f"the whole workload ({sample_len} bp)" if kind == "c2" else f"first {sample_len} bp of the workload reference with its share of the VCF (-V)") + \
             f" at {COVERAGE}x, -j {cores}, FASTQ to tmpfs; read-generation interval " \
             "(log line 'Starting read generation' to exit)" + ("" if kind == "c2" else " - the scan cost is per position, so pairs/s carries over to the full genome")
    source = workload_c2()[2] if kind == "c2" else ""
    print(json.dumps({
        "impl": "reference", "metric": "simulated read-pairs/s (2x150)", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * statistics.mean(secs), "higher_is_better": True,
        "scaling": "weak" if kind == "c2" else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(kind, source)},
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

    kind = args.workload or ("c2" if world == 1 else "c5")
    tmp = tempfile.mkdtemp(prefix="rsq_bench_")
    prof = rb.Profile.load_flat(unxz(PROFILE + ".flat.xz", tmp))
    eng = rb.Engine(prof, local_rank)
    if world > 1:
        # the engines of the ranks form one group: the 128-byte NCCL id travels over torch.distributed, everything else stays inside the library
        ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            ident.copy_(torch.frombuffer(bytearray(rb.group_unique_id()), dtype=torch.uint8))
        dist.broadcast(ident, 0)
        eng.join_group(bytes(ident.cpu().tolist()), rank, world)

    def make_reference(k):
        """Reference::ReadFasta (+ the VCF of -V): loaded once, like the `Reference&` the reference's Simulate() is handed."""
        if k == "c2":
            names, seqs, source = workload_c2()
            return rb.Reference.from_memory(names, [q.encode() for q in seqs]), sum(len(q) for q in seqs), source, 0
        names, seqs, vcf, n_rec = workload_c5(tmp)
        ref = rb.Reference.from_memory(names, [q.encode() for q in seqs])
        ref.load_variants(vcf)
        return ref, sum(len(q) for q in seqs), "", n_rec

    def step(ref):
        # host buffers in (reference bases), host buffers out (FASTQ text in host memory)
        eng.prepare(ref, seed=SEED, coverage=COVERAGE, shard_index=rank, shard_count=world)
        eng.simulate()
        rep = eng.download()
        return rep.as_dict()

    def measure(ref, warmup, steps):
        for _ in range(warmup):
            step(ref)
        sampler = ClockSampler(local_rank)
        sampler.start()
        barrier()
        t0 = time.perf_counter()
        reps = [step(ref) for _ in range(steps)]
        barrier()
        wall = time.perf_counter() - t0
        return reps, wall, sampler.stop()

    ref, ref_bases, source, n_variants = make_reference(kind)
    reps, wall, clocks = measure(ref, args.warmup, args.steps)

    dev_ms = sum(r["ms_bias"] + r["ms_syserr"] + r["ms_simulate"] + r["ms_gather"] for r in reps)
    sim_ms = sum(r["ms_simulate"] for r in reps)
    pairs = sum(r["pairs"] for r in reps)
    positions = sum(r["positions"] for r in reps)
    launches = sum(r["kernel_launches"] for r in reps)
    d2h = sum(r["bytes"][0] + r["bytes"][1] for r in reps)
    draws = sum(r["scan_draws"] for r in reps)
    (dev_ms, wall, sim_ms_max), (pairs_all, positions_all, launches_all, d2h_all, draws_all) = aggregate(
        dist, world, [dev_ms, wall, sim_ms], [pairs, positions, launches, d2h, draws], "cuda")
    group_pairs = sum(r["group_pairs"] for r in reps)   # the library's own all-reduce (NCCL inside the engine)

    if rank == 0:
        peak, peak_src = hbm_peak()
        # roofline of the dominant phase on rank 0: algorithmic bytes of its launches / their event time
        alg_bytes = pairs * BYTES_PER_PAIR + positions * BYTES_PER_POSITION
        achieved = alg_bytes / (sim_ms / 1000.0) / 1e9 if sim_ms else 0.0
        spec = reps[-1]["spec_rounds"] > 0
        kernel = "k_spec_scan + k_spec_reads (all rounds of a step)" if spec else "k_simulate"
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "simulate_dram_bytes.json")
        if os.path.exists(tpath) and kind == "c2":
            try:
                traffic = json.load(open(tpath))["spec" if spec else "serial"]["dram_bytes_per_step"]
            except Exception:
                traffic = None
        read_len = 150
        warp_inst = draws * SCAN_INST_PER_DRAW + pairs * 2 * read_len * READ_INST_PER_BASE
        line = {
            "metric": "simulated read-pairs/s (2x150)", "value": pairs_all / (dev_ms / 1000.0), "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak" if kind == "c2" else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(kind, source) + (f"; one fixed workload split over {world} GPUs (strong scaling, engines joined into an NCCL group)" if world > 1 else ""),
                       "simulation_path": f"speculative two-phase, {reps[-1]['spec_rounds']} rounds, first depth {reps[-1]['spec_depth']}" if spec else "serial (one warp per SimBlock)",
                       "l2": "per-step working set (reference, 2 B/base/strand systematic errors, 16 B/base surroundings, FASTQ arena) exceeds the 126 MB L2; "
                             "every step re-uploads the reference and rewrites all of it",
                       "pairs_per_step": pairs_all / args.steps, "blocks_per_step": sum(r["blocks"] for r in reps) / args.steps, "vcf_records": n_variants,
                       "group_pairs_per_step_all_reduced_in_library": group_pairs / args.steps,
                       "device_ms_breakdown_rank0": {k: sum(r[k] for r in reps) / args.steps for k in ("ms_upload", "ms_bias", "ms_syserr", "ms_simulate", "ms_gather", "ms_download")},
                       "scan_draws_per_s": draws / (sim_ms / 1000.0) if sim_ms else None,
                       "batches_rank0": reps[-1].get("batches"), "resident_bytes_per_base_rank0": reps[-1].get("resident_bytes_per_base")},
            "e2e": {"value": pairs_all / wall, "unit": "pairs/s", "h2d_bytes_per_step": ref_bases, "d2h_bytes_per_step": d2h_all / args.steps,
                    "ms_per_step": 1000 * wall / args.steps},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "secondary": {"what": "warp-instruction issue: the ceiling SURVEY.md 8(d) names for the scan (mt19937_64 regeneration + tempering) and the ordered Draw chains",
                                       "scan_draws_per_s": draws / (sim_ms / 1000.0) if sim_ms else None,
                                       "warp_inst_per_s": warp_inst / (sim_ms / 1000.0) if sim_ms else None,
                                       "issue_peak_warp_inst_per_s": ISSUE_PEAK,
                                       "frac_of_issue_peak": warp_inst / (sim_ms / 1000.0) / ISSUE_PEAK if sim_ms else None,
                                       "model": f"{SCAN_INST_PER_DRAW} warp instructions per scan draw + {READ_INST_PER_BASE} per base and read (ncu launch list profiles/r02y_launches.md; the reads kernel issues about 5 % fewer since it loads two candidates per instruction); rank 0, simulate phase"},
                         "note": "algorithmic bytes = pairs x 1420 B + positions x 8.25 B over the event time of the simulate phase (every launch of the two kernels of a "
                                 "step); the phase is bound by the per-SimBlock serial mt19937_64 stream and dependent FP64 Draw chains (latency), not by HBM"},
        }
        if world == 1 and kind == "c2" and not args.no_strong_ref:
            # the single-GPU reference point of the strong-scaling runs: one step of the N > 1 workload on this GPU
            ref = None
            ref5, bases5, _, nvar5 = make_reference("c5")
            reps5, wall5, _ = measure(ref5, 0, 1)
            r5 = reps5[0]
            dev5 = r5["ms_bias"] + r5["ms_syserr"] + r5["ms_simulate"] + r5["ms_gather"]
            line["strong_ref"] = {"workload": workload_name("c5"), "n_gpus": 1, "steps": 1, "value": r5["pairs"] / (dev5 / 1000.0), "unit": "pairs/s",
                                  "ms_per_step": dev5, "e2e": {"value": r5["pairs"] / wall5, "ms_per_step": 1000 * wall5}, "pairs": r5["pairs"], "vcf_records": nvar5,
                                  "device_ms_breakdown": {k: r5[k] for k in ("ms_upload", "ms_bias", "ms_syserr", "ms_simulate", "ms_gather", "ms_download")},
                                  "note": "bench.py --gpus N (N > 1) runs this workload split over N GPUs: value(N) / this value is the strong-scaling speed-up"}
            ref5 = None
        if world > 1 and kind == "c5":
            # the committed single-GPU measurement of this very workload (a `strong_ref` of an N = 1 line, copied into profiles/): the denominator of the
            # strong-scaling speed-up.  The N = 1 line of a scaling series is config C2, a different workload - its value is not comparable with this one.
            spath = os.path.join(ROOT, "profiles", "strong_ref_n1.json")
            if os.path.exists(spath):
                try:
                    sr = json.load(open(spath))
                    line["strong_ref"] = {"n_gpus": 1, "value": sr["value"], "unit": "pairs/s", "e2e_value": sr["e2e"]["value"], "source": sr.get("source", "profiles/strong_ref_n1.json"),
                                          "speedup_device": line["value"] / sr["value"], "speedup_e2e": line["e2e"]["value"] / sr["e2e"]["value"],
                                          "note": "same workload measured on ONE GPU of this pool with the same build (a committed measurement, not taken in this run)"}
                except Exception:
                    pass
        qpath = os.path.join(GOLDEN, "profile150q.flat.xz")
        if world == 1 and kind == "c2" and os.path.exists(qpath) and not args.no_realistic:
            # the same workload on the profile with realistic tables (every base quality 2..41, two tiles, 2.1 MB of LogArrayResult tables)
            profq = rb.Profile.load_flat(unxz("profile150q.flat.xz", tmp))
            engq = rb.Engine(profq, local_rank)
            refq = make_reference("c2")[0]

            def stepq():
                engq.prepare(refq, seed=SEED, coverage=COVERAGE)
                engq.simulate()
                return engq.download().as_dict()
            for _ in range(3):
                stepq()
            t0 = time.perf_counter()
            repsq = [stepq() for _ in range(3)]
            wallq = time.perf_counter() - t0
            devq = sum(r["ms_bias"] + r["ms_syserr"] + r["ms_simulate"] + r["ms_gather"] for r in repsq)
            pq = sum(r["pairs"] for r in repsq)
            line["realistic_profile"] = {"profile": "profile150q: 40 base qualities (2..41), 2 tiles, 2.1 MB of probability tables, InDel rate 2e-5 (tests/golden/make_golden.py); "
                                                    "parity pinned on the small reference (tests/golden/profile150q_small_sha256.json)",
                                         "steps": 3, "warmup": 3, "value": pq / (devq / 1000.0), "unit": "pairs/s", "ms_per_step": devq / 3,
                                         "e2e": {"value": pq / wallq, "ms_per_step": 1000 * wallq / 3}, "pairs_per_step": pq / 3,
                                         "device_ms_breakdown": {k: sum(r[k] for r in repsq) / 3 for k in ("ms_upload", "ms_bias", "ms_syserr", "ms_simulate", "ms_gather", "ms_download")},
                                         "spec_rounds": repsq[-1]["spec_rounds"]}
            engq.close()
        if world == 1 and kind == "c2" and not args.no_cold:
            # the cold drop-in call: engine creation + table upload + prologue + simulation + both FASTQ files written (tmpfs)
            names, seqs, _ = workload_c2()
            refc = rb.Reference.from_memory(names, [q.encode() for q in seqs])
            shm = "/dev/shm" if os.path.isdir("/dev/shm") else tmp
            outs = [os.path.join(shm, f"rsq_cold_{os.getpid()}_{k}.fq") for k in (1, 2)]
            t0 = time.perf_counter()
            repc = rb.simulate(prof, refc, outs[0], outs[1], seed=SEED, coverage=COVERAGE)
            cold = time.perf_counter() - t0
            for o in outs:
                os.remove(o)
            line["e2e_cold"] = {"value": repc.pairs / cold, "unit": "pairs/s", "seconds": cold,
                                "what": "one rsq_simulate call from nothing: engine creation, table upload, prologue, simulation, both FASTQ files written to tmpfs"}
        if world == 1 and os.path.exists(ORACLE) and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            p, gen_s, _ = time_reference_cpu(kind, REF_LEN if kind == "c2" else 1_500_000, COVERAGE, cores, tmp)
            line["cpu_baseline"] = {"value": p / gen_s, "unit": "pairs/s", "cores": cores, "kind": "reference",
                                    "sample": f"reference binary (unmodified sources) on " + (f"the whole workload ({REF_LEN} bp" if kind == "c2" else "the first 1.5 Mbp of the workload with its variants (") +
                                              f" at {COVERAGE}x), -j {cores}, FASTQ to tmpfs, read-generation interval {gen_s:.1f} s ({p} pairs)"}
        os.write(result_fd, (json.dumps(line) + "\n").encode())
    os.close(result_fd)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong-ref", action="store_true", help="N = 1: skip the single step of the strong-scaling workload")
    ap.add_argument("--no-cold", action="store_true", help="N = 1: skip the cold rsq_simulate call")
    ap.add_argument("--no-realistic", action="store_true", help="N = 1: skip the steps on the realistic-tables profile (profile150q)")
    ap.add_argument("--workload", default=None, choices=["c2", "c5"], help="default: c2 for one GPU, c5 (strong scaling) for several")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
