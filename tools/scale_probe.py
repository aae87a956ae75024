#!/usr/bin/env python3
"""Runs the hot path on larger synthetic references (GPU box): python tools/scale_probe.py <Mbp> [<Mbp> ...] [--seqs N] [--coverage C]

Prints one JSON line per size with the engine's own report (device times per stage, pairs, rounds) and the wall times of
prepare / simulate / download.  The FASTQ text stays in pinned host memory; nothing is written to disk."""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
import make_synthetic  # noqa: E402
import reseq_b200 as rb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mbp", type=float, nargs="+")
    ap.add_argument("--seqs", type=int, default=4)
    ap.add_argument("--coverage", type=float, default=30.0)
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--profile", default=None, help="golden profile base name (tests/golden/<name>.flat.xz), e.g. profile250; default: the bench profile")
    ap.add_argument("--methylation", action="store_true", help="bisulfite run: one unmethylated region of ~600 bp per 10 kb, methylation ~ U(0,1) (BASELINE config C4)")
    ap.add_argument("--vcf", action="store_true", help="variant run (BASELINE config C5): phased diploid VCF with 1 SNP per kb and 1 indel (<= 20 bases) per 10 kb")
    ap.add_argument("--gz", action="store_true", help="with --files: .fq.gz output names (device-side gzip)")
    ap.add_argument("--files", action="store_true", help="drop-in call rsq_simulate (text streamed through the writer thread; set RSQ_DISCARD_OUTPUT=1 for runs larger than the disk)")
    args = ap.parse_args()
    tmp = tempfile.mkdtemp(prefix="rsq_scale_")
    prof = rb.Profile.load_flat(bench.unxz((args.profile or bench.PROFILE) + ".flat.xz", tmp))
    eng = rb.Engine(prof, 0)
    for mbp in args.mbp:
        total = int(mbp * 1e6)
        sizes = [total // args.seqs] * args.seqs
        t0 = time.perf_counter()
        seqs = make_synthetic.gen_reference(sizes, 4321)
        ref = rb.Reference.from_memory([f"chr{i + 1} synthetic" for i in range(len(seqs))], [s.encode() for s in seqs])
        n_variants = 0
        if args.vcf:
            vcf = os.path.join(tmp, f"var_{mbp}.vcf")
            n_variants = make_synthetic.write_vcf(vcf, [f"chr{i + 1}" for i in range(len(seqs))], seqs, 77)
            t_v = time.perf_counter()
            ref.load_variants(vcf)
            print(json.dumps({"vcf_records": n_variants, "load_variants_s": round(time.perf_counter() - t_v, 2)}), flush=True)
        if args.methylation:
            import random
            rnd = random.Random(9)
            bed = os.path.join(tmp, f"meth_{mbp}.bed")
            with open(bed, "w") as f:
                for i, size in enumerate(sizes):
                    # the first region starts at 0: with a later first region Simulator::CTConversion indexes regions.at(-1) for reverse reads
                    # in front of it (std::out_of_range in the reference; reported as an error here too)
                    for start in range(0, size - 2000, 10000):
                        f.write(f"chr{i + 1}\t{start}\t{start + rnd.randint(100, 1100)}\t{rnd.random():.3f}\n")
            ref.load_methylation(bed)
        t_gen = time.perf_counter() - t0
        best = None
        if args.files:
            out = [os.path.join(tmp, "probe_R1.fq" + (".gz" if args.gz else "")), os.path.join(tmp, "probe_R2.fq" + (".gz" if args.gz else ""))]
            t0 = time.perf_counter()
            rep = rb.simulate(prof, ref, out[0], out[1], seed=42, coverage=args.coverage).as_dict()
            wall = time.perf_counter() - t0
            sizes_on_disk = [os.path.getsize(o) for o in out]
            for o in out:
                os.remove(o)
            print(json.dumps({"mbp": mbp, "seqs": args.seqs, "coverage": args.coverage, "profile": args.profile or bench.PROFILE, "methylation": args.methylation, "vcf_records": n_variants, "gz": args.gz, "gen_ref_s": round(t_gen, 1), "wall_rsq_simulate_s": round(wall, 2),
                              "pairs_per_s_e2e": round(rep["pairs"] / wall), "bytes_on_disk": sizes_on_disk, "report": rep}), flush=True)
            continue
        for _ in range(args.repeat):
            t0 = time.perf_counter()
            eng.prepare(ref, seed=42, coverage=args.coverage)
            t1 = time.perf_counter()
            eng.simulate()
            t2 = time.perf_counter()
            rep = eng.download().as_dict()
            t3 = time.perf_counter()
            rec = {"mbp": mbp, "seqs": args.seqs, "coverage": args.coverage, "vcf_records": n_variants, "gen_ref_s": round(t_gen, 1), "wall_prepare_s": round(t1 - t0, 3),
                   "wall_simulate_s": round(t2 - t1, 3), "wall_download_s": round(t3 - t2, 3), "pairs_per_s_e2e": round(rep["pairs"] / (t3 - t0)),
                   "pairs_per_s_device": round(rep["pairs"] / ((rep["ms_syserr"] + rep["ms_simulate"] + rep["ms_gather"]) / 1e3)), "report": rep}
            if best is None or rec["pairs_per_s_e2e"] > best["pairs_per_s_e2e"]:
                best = rec
        print(json.dumps(best), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
