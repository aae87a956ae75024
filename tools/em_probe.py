#!/usr/bin/env python3
"""seqToIllumina at scale (BASELINE config C3 shape: given 300-bp fragments, error model only) on the GPU box.

    python tools/em_probe.py [records] [fragment length]

Writes a synthetic input (records cut from a synthetic reference, sparse systematic errors), runs rsq_apply_error_model on the
speculative and on the serial kernel form and prints device time of the kernels and wall time of the call (which includes
reading and parsing the ~1 kB-per-record FASTA on the host and writing the FASTQ)."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
import make_synthetic  # noqa: E402
import reseq_b200 as rb  # noqa: E402


def write_input(path, n, flen, seed=3):
    rng = np.random.default_rng(seed)
    ref = np.frombuffer(make_synthetic.gen_reference([2_000_000], 77)[0].encode(), dtype=np.uint8)
    comp = np.zeros(256, dtype=np.uint8)
    comp[list(b"ACGT")] = list(b"TGCA")
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    with open(path, "wb") as f:
        for i in range(n):
            seg = i % 2
            st = int(rng.integers(0, len(ref) - flen))
            frag = ref[st:st + flen]
            if seg:
                frag = comp[frag][::-1]
            dom = frag.copy()
            rate = np.zeros(flen, dtype=np.uint8)
            hits = np.flatnonzero(rng.random(flen) < 0.01)
            rate[hits] = rng.integers(5, 86, size=len(hits))
            dom[hits] = acgt[rng.integers(0, 4, size=len(hits))]
            f.write(b">frag%d %d;%d;" % (i, seg + 1, flen) + dom.tobytes() + b";" + (rate + 33).tobytes() + b"\n" + frag.tobytes() + b"\n")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    flen = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    tmp = tempfile.mkdtemp(prefix="rsq_em_")
    src = os.path.join(tmp, "frags.fa")
    t0 = time.perf_counter()
    write_input(src, n, flen)
    t_gen = time.perf_counter() - t0
    eng = rb.Engine(rb.Profile.load_flat(bench.unxz(bench.PROFILE + ".flat.xz", tmp)), 0)
    outs = {}
    for path in ("spec", "serial"):
        os.environ["RSQ_SIM_PATH"] = path
        out = os.path.join(tmp, f"out_{path}.fq")
        t0 = time.perf_counter()
        rep = eng.apply_error_model(src, out, 7).as_dict()
        wall = time.perf_counter() - t0
        outs[path] = out
        print(json.dumps({"path": path, "records": n, "fragment_length": flen, "gen_input_s": round(t_gen, 1), "wall_s": round(wall, 2),
                          "ms_kernels": round(rep["ms_simulate"], 1), "reads_per_s_kernels": round(n / (rep["ms_simulate"] / 1e3)),
                          "rounds": rep["spec_rounds"], "depth": rep["spec_depth"], "batches": rep["blocks"]}), flush=True)
    same = open(outs["spec"], "rb").read() == open(outs["serial"], "rb").read()
    print("identical output of both kernel forms:", same)
    eng.close()


if __name__ == "__main__":
    main()
