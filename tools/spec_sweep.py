#!/usr/bin/env python3
"""Sweeps the speculation depth / reads-per-warp of the two-phase simulation on the bench workload (GPU box).

    python tools/spec_sweep.py [D,R ...]        e.g.  4,8 8,8 8,16 2,32 serial
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import reseq_b200 as rb  # noqa: E402


def main():
    configs = sys.argv[1:] or ["4,8", "8,8", "8,16", "2,32", "serial"]
    tmp = tempfile.mkdtemp(prefix="rsq_sweep_")
    prof = rb.Profile.load_flat(bench.unxz(bench.PROFILE + ".flat.xz", tmp))
    eng = rb.Engine(prof, 0)
    if os.environ.get("RSQ_SWEEP_MBP"):   # larger synthetic genome (4 sequences) instead of the bench workload
        import make_synthetic
        seqs = make_synthetic.gen_reference([int(float(os.environ["RSQ_SWEEP_MBP"]) * 1e6) // 4] * 4, 4321)
        ref = rb.Reference.from_memory([f"chr{i + 1} synthetic" for i in range(4)], [s.encode() for s in seqs])
    else:
        seq = bench.workload_sequence().encode()
        ref = rb.Reference.from_memory(["ecoli_sized synthetic"], [seq])
    for cfg in configs:
        for k in ("RSQ_SIM_PATH", "RSQ_SPEC_DEPTH", "RSQ_SPEC_LANES"):
            os.environ.pop(k, None)
        if cfg == "serial":
            os.environ["RSQ_SIM_PATH"] = "serial"
        elif cfg != "auto":
            d, r = cfg.split(",")
            os.environ["RSQ_SPEC_DEPTH"], os.environ["RSQ_SPEC_LANES"] = d, r
        best = None
        for _ in range(int(os.environ.get("RSQ_SWEEP_REPEAT", "3"))):
            eng.prepare(ref, seed=bench.SEED, coverage=bench.COVERAGE)
            rep = eng.simulate().as_dict()
            if best is None or rep["ms_simulate"] < best["ms_simulate"]:
                best = rep
        print(json.dumps({"config": cfg, "ms_simulate": round(best["ms_simulate"], 2), "ms_gather": round(best["ms_gather"], 2), "rounds": best["spec_rounds"],
                          "depth": best["spec_depth"], "pairs": best["pairs"], "launches": best["kernel_launches"]}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
