#!/usr/bin/env python3
"""ms_simulate of config C2 on the realistic-tables profile (profile150q); used next to tools/knob_probe.py for A/B builds."""
import json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench
import reseq_b200 as rb
tmp = tempfile.mkdtemp()
prof = rb.Profile.load_flat(bench.unxz("profile150q.flat.xz", tmp))
names, seqs, _ = bench.workload_c2()
ref = rb.Reference.from_memory(names, [q.encode() for q in seqs])
eng = rb.Engine(prof, 0)
best = None
for _ in range(3):
    eng.prepare(ref, seed=42, coverage=30.0)
    rep = eng.simulate().as_dict()
    if best is None or rep["ms_simulate"] < best["ms_simulate"]:
        best = rep
print(json.dumps({"profile": "profile150q", "ms_simulate": round(best["ms_simulate"], 2), "ms_syserr": round(best["ms_syserr"], 2), "pairs": best["pairs"], "rounds": best["spec_rounds"]}))
