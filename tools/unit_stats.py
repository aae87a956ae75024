#!/usr/bin/env python3
"""Per-unit counters of the speculative path after one C2 run: how many rounds, reads and scan draws each SimBlock took.
    python tools/unit_stats.py        (GPU box)"""
import ctypes as C
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import reseq_b200 as rb  # noqa: E402

tmp = tempfile.mkdtemp(prefix="rsq_units_")
prof = rb.Profile.load_flat(bench.unxz(bench.PROFILE + ".flat.xz", tmp))
names, seqs, _ = bench.workload_c2()
ref = rb.Reference.from_memory(names, [q.encode() for q in seqs])
eng = rb.Engine(prof, 0)
eng.prepare(ref, seed=42, coverage=30.0)
rep = eng.simulate().as_dict()
raw = eng.fetch("spec_blocks").tobytes()
a = np.frombuffer(raw, dtype=np.uint32).reshape(-1, 20)   # SpecBlock: 14 u32, bytes[2] u64, scan_draws u64
rounds, reads = a[:, 13], a[:, 12]
draws = a[:, 18].astype(np.uint64) | (a[:, 19].astype(np.uint64) << 32)
print(json.dumps({"ms_simulate": rep["ms_simulate"], "units": int(len(a)), "rounds_hist": np.bincount(rounds).tolist(),
                  "reads_mean": float(reads.mean()), "reads_pct": [int(np.percentile(reads, p)) for p in (1, 10, 50, 90, 99, 100)],
                  "draws_mean": float(draws.mean()), "corr_rounds_reads": float(np.corrcoef(rounds, reads)[0, 1])}))
for r in sorted(set(rounds.tolist()))[-6:]:
    m = rounds == r
    print(r, int(m.sum()), "reads mean %.0f min %d max %d" % (reads[m].mean(), reads[m].min(), reads[m].max()), "draws mean %.0f" % draws[m].mean())
