#!/usr/bin/env python3
"""Host-side timeline of one C2 step (RSQ_TIMING=1 prints the stages of prepare/simulate/download to stderr)."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench
import reseq_b200 as rb
tmp = tempfile.mkdtemp()
prof = rb.Profile.load_flat(bench.unxz(bench.PROFILE + ".flat.xz", tmp))
names, seqs, _ = bench.workload_c2()
ref = rb.Reference.from_memory(names, [q.encode() for q in seqs])
eng = rb.Engine(prof, 0)
for i in range(4):
    if i == 3:
        os.environ["RSQ_TIMING"] = "1"
    t0 = time.perf_counter(); eng.prepare(ref, seed=42, coverage=30.0); t1 = time.perf_counter(); eng.simulate(); t2 = time.perf_counter(); rep = eng.download(); t3 = time.perf_counter()
    print("step %d: prepare %.1f ms simulate %.1f ms download %.1f ms | device: bias %.1f syserr %.1f sim %.1f gather %.1f dl %.1f" % (i, 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), rep.ms_bias, rep.ms_syserr, rep.ms_simulate, rep.ms_gather, rep.ms_download), flush=True)
