#!/usr/bin/env python3
"""sha256 of the reference's own FASTQ for BASELINE config C2 (4 641 652 bp synthetic reference of bench.py and
tests/test_gpu_parity.py::test_full_size_*, 30x, 2x150, seed 42, `-j 1`), one entry per golden profile.

The FASTQ itself is 2 x 170 MB and cannot be committed; its hash and pair count can.  Run in the build container:
    python tests/golden/make_fullsize_hashes.py [profile150 profile150r ...]
writes tests/golden/fullsize_c2_sha256.json.  The GPU test compares the engine's output with these hashes; with
RSQ_LIVE_ORACLE=1 it also runs oracle/_ref/reseq_oracle on the box and compares byte for byte."""
import hashlib
import json
import lzma
import os
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_synthetic  # noqa: E402

ORACLE = os.path.join(ROOT, "oracle", "_ref", "reseq_oracle")
OUT = os.path.join(HERE, "fullsize_c2_sha256.json")
SIZE, REF_SEED, SEED, COVERAGE = 4_641_652, 1234, 42, 30.0


def sha(path):
    h = hashlib.sha256()
    n = 0
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b)
            n += b.count(b"\n")
    return h.hexdigest(), n // 4


def main():
    """Arguments: profile names; "ecoli:<profile>" runs the profile on the reference's own E. coli fixture (test/ecoli-GCF_000005845.2_ASM584v2_genomic.fa,
    committed as tests/golden/ecoli_GCF_000005845.2.fa.xz) instead of the synthetic sequence and stores the entry as "ecoli_<profile>"."""
    profiles = sys.argv[1:] or ["profile150", "profile150r"]
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    with tempfile.TemporaryDirectory(prefix="rsq_full_") as tmp:
        fa = os.path.join(tmp, "ref.fa")
        seq = make_synthetic.gen_reference([SIZE], REF_SEED)[0]
        with open(fa, "w") as f:   # one line: the id the tests pass to Reference.from_memory
            f.write(">ecoli_sized synthetic\n" + seq + "\n")
        real = os.path.join(tmp, "ecoli.fa")
        with lzma.open(os.path.join(HERE, "ecoli_GCF_000005845.2.fa.xz")) as src, open(real, "wb") as dst:
            dst.write(src.read())
        for arg in profiles:
            on_real = arg.startswith("ecoli:")
            prof = arg.split(":")[-1]
            run_fa = real if on_real else fa
            stats = os.path.join(tmp, prof + ".reseq")
            for ext in (".reseq", ".reseq.ipf"):
                with lzma.open(os.path.join(HERE, prof + ext + ".xz")) as src, open(os.path.join(tmp, prof + ext), "wb") as dst:
                    dst.write(src.read())
            r1, r2 = os.path.join(tmp, prof + "_R1.fq"), os.path.join(tmp, prof + "_R2.fq")
            t0 = time.time()
            subprocess.run([ORACLE, "illuminaPE", "-j", "1", "--verbosity", "1", "-s", stats, "-R", run_fa, "--ipfIterations", "0", "--seed", str(SEED),
                            "-c", str(COVERAGE), "-1", r1, "-2", r2], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            h1, n1 = sha(r1)
            h2, n2 = sha(r2)
            assert n1 == n2
            key = ("ecoli_" if on_real else "") + prof
            res[key] = {"r1": h1, "r2": h2, "pairs": n1, "bytes": [os.path.getsize(r1), os.path.getsize(r2)], "size": SIZE, "ref_seed": None if on_real else REF_SEED,
                        "seed": SEED, "coverage": COVERAGE, "oracle_seconds_j1": round(time.time() - t0, 1)}
            os.remove(r1), os.remove(r2)
            print(key, res[key], flush=True)
            json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
