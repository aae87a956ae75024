#!/usr/bin/env python3
"""Throughput of the drop-in call (rsq_simulate) at BASELINE config C2's size with plain, host-gzip and device-gzip output files.
    python tools/gz_probe.py            (needs a GPU)"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import lzma  # noqa: E402

import make_synthetic  # noqa: E402
import reseq_b200 as rb  # noqa: E402


def main():
    tmp = tempfile.mkdtemp(prefix="rsq_gz_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    flat = os.path.join(tmp, "p.flat")
    with lzma.open(os.path.join(ROOT, "tests", "golden", "profile150r.flat.xz")) as f, open(flat, "wb") as o:
        o.write(f.read())
    prof = rb.Profile.load_flat(flat)
    seq = make_synthetic.gen_reference([4_641_652], 1234)[0]
    ref = rb.Reference.from_memory(["ecoli_sized synthetic"], [seq.encode()])
    out = []
    modes = (("plain", {}, ".fq"), ("plain", {}, ".fq"), ("gzip device", {"RSQ_GZIP": "device"}, ".fq.gz"), ("gzip device", {"RSQ_GZIP": "device"}, ".fq.gz"),
                              ("gzip host level 1", {"RSQ_GZIP": "host", "RSQ_GZIP_LEVEL": "1"}, ".fq.gz"),
                              ("gzip host default level", {"RSQ_GZIP": "host"}, ".fq.gz"))
    if len(sys.argv) > 2 and sys.argv[1] == "--once":   # one run of one mode (for ncu)
        modes = tuple(m for m in modes if m[0] == sys.argv[2])[:1]
    for name, env, suffix in modes:
        for k in ("RSQ_GZIP", "RSQ_GZIP_LEVEL"):
            os.environ.pop(k, None)
        os.environ.update(env)
        o1, o2 = os.path.join(tmp, "r1" + suffix), os.path.join(tmp, "r2" + suffix)
        t0 = time.perf_counter()
        rep = rb.simulate(prof, ref, o1, o2, seed=42, coverage=30.0)
        dt = time.perf_counter() - t0
        out.append({"mode": name, "seconds": round(dt, 3), "pairs": int(rep.pairs), "pairs_per_s": round(rep.pairs / dt), "text_bytes": int(rep.bytes[0] + rep.bytes[1]),
                    "file_bytes": os.path.getsize(o1) + os.path.getsize(o2), "host_threads": os.cpu_count()})
        print(json.dumps(out[-1]), flush=True)
        os.remove(o1)
        os.remove(o2)


if __name__ == "__main__":
    main()
