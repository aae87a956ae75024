#!/usr/bin/env python3
"""seqToIllumina on more than one batch of 10 000 input records: input and the hash of the reference's output.

The released reference dead-locks on its second batch (WriteSingleReads waits for written_blocks_, which nothing increments,
Simulator.cpp:184-213); `oracle/dump_tables errmodel` runs the reference's own SimulateErrorModelOnly with that counter preset, so that
one thread writes the batches in input order - everything else (one block_seed_gen_() seed per batch, ApplyErrorsAndQualityToFastaInput,
FlushWriteValues) is the reference's code.  The input is the committed em_frags.fa (9000 records) three times with `_<k>` appended to the
ids: 27 000 records = 3 batches.  The output is 11 MB, so only its sha256 is committed (tests/golden/em_multibatch_sha256.json).
    python tests/golden/make_em_multibatch.py
"""
import hashlib
import json
import lzma
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
COPIES, SEED = 3, 7


def write_input(path, copies=COPIES):
    lines = lzma.open(os.path.join(HERE, "em_frags.fa.xz"), "rt").read().split("\n")
    n = 0
    with open(path, "w") as f:
        for rep in range(copies):
            for line in lines:
                if line.startswith(">"):
                    parts = line.split(" ", 1)
                    line = parts[0] + f"_{rep}" + (" " + parts[1] if len(parts) > 1 else "")
                    n += 1
                if line:
                    f.write(line + "\n")
    return n


def run_oracle(dump_tables, stats, fasta, out, seed=SEED):
    subprocess.run([dump_tables, "errmodel", stats, fasta, str(seed), out], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=1200)


def main():
    with tempfile.TemporaryDirectory(prefix="rsq_em_") as tmp:
        for ext in (".reseq", ".reseq.ipf"):
            with lzma.open(os.path.join(HERE, "profile150" + ext + ".xz")) as src, open(os.path.join(tmp, "profile150" + ext), "wb") as dst:
                dst.write(src.read())
        fa, out = os.path.join(tmp, "em3.fa"), os.path.join(tmp, "em3.fq")
        n = write_input(fa)
        run_oracle(os.path.join(ROOT, "oracle", "_ref", "dump_tables"), os.path.join(tmp, "profile150.reseq"), fa, out)
        data = open(out, "rb").read()
        res = {"records": n, "copies": COPIES, "seed": SEED, "profile": "profile150", "bytes": len(data), "sha256": hashlib.sha256(data).hexdigest(),
               "input_sha256": hashlib.sha256(open(fa, "rb").read()).hexdigest()}
        json.dump(res, open(os.path.join(HERE, "em_multibatch_sha256.json"), "w"), indent=1, sort_keys=True)
        print(res)


if __name__ == "__main__":
    sys.exit(main())
