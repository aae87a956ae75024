#!/usr/bin/env python3
"""ms_simulate of one workload under several settings of the speculative path's tuning knobs (environment variables read per batch).

    python tools/knob_probe.py [--mbp 4.64] [--seqs 1] "RSQ_SPEC_BUDGET=0.7" "RSQ_SPEC_BUDGET=1.0 RSQ_SPEC_MAX_DEPTH=32" ...
The first line is the default setting; every setting runs `--repeat` times and reports the minimum."""
import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
import make_synthetic  # noqa: E402
import reseq_b200 as rb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("settings", nargs="*")
    ap.add_argument("--mbp", type=float, default=0.0, help="synthetic reference of this size; default: the bench's C2 workload (real E. coli)")
    ap.add_argument("--seqs", type=int, default=1)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--vcf", action="store_true")
    args = ap.parse_args()
    tmp = tempfile.mkdtemp(prefix="rsq_knob_")
    prof = rb.Profile.load_flat(bench.unxz(bench.PROFILE + ".flat.xz", tmp))
    if args.mbp:
        seqs = make_synthetic.gen_reference([int(args.mbp * 1e6) // args.seqs] * args.seqs, 4321)
        ref = rb.Reference.from_memory([f"chr{i + 1} synthetic" for i in range(len(seqs))], [s.encode() for s in seqs])
        if args.vcf:
            vcf = os.path.join(tmp, "v.vcf")
            make_synthetic.write_vcf(vcf, [f"chr{i + 1}" for i in range(len(seqs))], seqs, 77)
            ref.load_variants(vcf)
    else:
        names, seqs, _ = bench.workload_c2()
        ref = rb.Reference.from_memory(names, [q.encode() for q in seqs])
    eng = rb.Engine(prof, 0)
    keys = set()
    for setting in [""] + args.settings:
        for k in keys:
            os.environ.pop(k, None)
        for kv in setting.split():
            k, v = kv.split("=")
            os.environ[k] = v
            keys.add(k)
        best = None
        for _ in range(args.repeat):
            eng.prepare(ref, seed=42, coverage=30.0)
            rep = eng.simulate().as_dict()
            if best is None or rep["ms_simulate"] + rep["ms_syserr"] < best["ms_simulate"] + best["ms_syserr"]:
                best = rep
        print(json.dumps({"setting": setting or "default", "ms_simulate": round(best["ms_simulate"], 2), "ms_syserr": round(best["ms_syserr"], 2), "syserr_passes": best.get("syserr_passes"), "rounds": best.get("spec_rounds"), "pairs": best["pairs"],
                          "launches": best.get("launches"), "depth": best.get("spec_depth")}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
