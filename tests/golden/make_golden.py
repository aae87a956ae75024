#!/usr/bin/env python3
"""Regenerates the committed fixtures of tests/golden/ with the reference's own code (oracle/_ref).

Run in the build container (needs /root/reference for the TruSeq adapter files and a built oracle):
    python tests/golden/make_golden.py

  profile150.reseq.xz / .reseq.ipf.xz   2x150 profile: synthetic SAM -> `reseq illuminaPE --statsOnly` +
                                        `--stopAfterEstimation` (reference stats + IPF code), bias fit replaced
                                        by deterministic non-trivial values (dump_tables patch)
  profile150.flat.xz                    what the reference holds in memory after Load + PrepareProcessing +
                                        Estimate + PrepareResult for that profile (dump_tables profile)
  profile150r.{reseq,reseq.ipf,flat}.xz same pipeline with a realistic InDel rate in the synthetic SAM (2e-5 per base for insertions and
                                        for deletions instead of 8e-4): the bench profile; profile150 stays the InDel stress profile
  profile150t.{reseq,reseq.ipf,flat}.xz same pipeline, `--tiles`, from a SAM with Casava-1.8 read names on three tiles and 20 % of the reads 144
                                        instead of 150 bases long: per-tile tables, tile and read-length draws
  profile150q.flat.xz + profile150q_small_sha256.json   the "realistic" bench profile: 60 000 pairs with every base quality 2..41 (40 values), two tiles, InDel rate 2e-5:
                                        2.1 MB of LogArrayResult tables (profile150r: 0.5 MB).  Its .reseq (89 MB) and .ipf (26 MB) are too large to commit: only the
                                        flat image the reference holds in memory is, plus the sha256 of the reference's FASTQ for simref_small.fa (seed 42, 20x)
  profile250.{reseq,reseq.ipf,flat}.xz  same pipeline from a SAM with 2x250 reads (the read length of BASELINE config C4), InDel rate 1e-4
  simref_small.fa                       small multi-contig reference with N runs and one too-short contig
  sim_small_seed42_R{1,2}.fq.xz         `reseq illuminaPE -j 1 --seed 42 -c 20` on simref_small.fa
  simref_small_meth.bed, sim_small_meth_seed42_R{1,2}.fq.xz   same run with `--methylation` (bisulfite C->T conversions)
  em_frags.fa.xz / em_seed7.fq.xz       seqToIllumina input and its output.  9000 records = ONE 10000-record batch on purpose:
                                        the reference never increments written_blocks_ (Simulator.cpp:184-213), so its second
                                        batch waits forever in WriteSingleReads -- larger inputs cannot be pinned against it
"""
import lzma
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
ORACLE = os.path.join(ROOT, "oracle", "_ref", "reseq_oracle")
DUMP = os.path.join(ROOT, "oracle", "_ref", "dump_tables")
SYN = os.path.join(ROOT, "tools", "make_synthetic.py")
ADAPTERS = "/root/reference/adapters/TruSeq_single"


def run(cmd, **kw):
    print("+", " ".join(cmd), flush=True)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, **kw)
    if res.returncode:
        sys.stderr.write(res.stdout)
        raise SystemExit(f"command failed: {cmd}")
    return res.stdout


def xz(src, dst):
    with open(src, "rb") as f, lzma.open(dst, "wb", preset=9 | lzma.PRESET_EXTREME) as o:
        shutil.copyfileobj(f, o)


def write_bed(path):
    """Unmethylated-region file for simref_small.fa: touching, separated and far-apart regions; chr2 does not start at 0."""
    import random
    rnd = random.Random(3)
    lines = ["track type=bedGraph name=x"]
    for name, length in (("chr1", 30000), ("chr2", 22000), ("chr4", 15000)):
        pos = 0 if name != "chr2" else 137
        while pos < length - 600:
            end = min(length, pos + rnd.randint(1, 900))
            lines.append(f"{name}\t{pos}\t{end}\t{rnd.random():.3f}")
            pos = end + rnd.choice([0, 0, 1, 50, 700, 2500])
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def main():
    tmp = tempfile.mkdtemp(prefix="rsq_golden_")
    py = sys.executable
    ref = os.path.join(tmp, "prof_ref.fa")
    run([py, SYN, "reference", ref, "--sizes", "200000,120000", "--seed", "7"])
    sam = os.path.join(tmp, "prof.sam")
    run([py, SYN, "sam", ref, sam, "--pairs", "20000", "--seed", "11"])
    raw = os.path.join(tmp, "raw.reseq")
    run([ORACLE, "illuminaPE", "-j", "8", "-b", sam, "-r", ref, "--adapterFile", ADAPTERS + ".fa", "--adapterMatrix", ADAPTERS + ".mat",
         "--statsOnly", "-S", raw])
    log = run([ORACLE, "illuminaPE", "-j", "8", "-s", raw, "-r", ref, "--stopAfterEstimation"])
    if "did not reach precision aim" in log:
        raise SystemExit("IPF did not converge for every table; the profile would be refitted on load")
    prof = os.path.join(tmp, "profile150.reseq")
    run([DUMP, "patch", raw, prof, "5"])
    shutil.copy(raw + ".ipf", prof + ".ipf")
    flat = os.path.join(tmp, "profile150.flat")
    run([DUMP, "profile", prof, flat])
    xz(prof, os.path.join(HERE, "profile150.reseq.xz"))
    xz(prof + ".ipf", os.path.join(HERE, "profile150.reseq.ipf.xz"))
    xz(flat, os.path.join(HERE, "profile150.flat.xz"))

    # bench profile: same pipeline, realistic InDel rate
    sam_r = os.path.join(tmp, "prof_r.sam")
    run([py, SYN, "sam", ref, sam_r, "--pairs", "20000", "--seed", "11", "--indel-rate", "0.00002"])
    raw_r = os.path.join(tmp, "raw_r.reseq")
    run([ORACLE, "illuminaPE", "-j", "8", "-b", sam_r, "-r", ref, "--adapterFile", ADAPTERS + ".fa", "--adapterMatrix", ADAPTERS + ".mat",
         "--statsOnly", "-S", raw_r])
    log = run([ORACLE, "illuminaPE", "-j", "8", "-s", raw_r, "-r", ref, "--stopAfterEstimation"])
    if "did not reach precision aim" in log:
        raise SystemExit("IPF did not converge for every table of the realistic profile")
    prof_r = os.path.join(tmp, "profile150r.reseq")
    run([DUMP, "patch", raw_r, prof_r, "5"])
    shutil.copy(raw_r + ".ipf", prof_r + ".ipf")
    run([DUMP, "profile", prof_r, os.path.join(tmp, "profile150r.flat")])
    for name in ("profile150r.reseq", "profile150r.reseq.ipf", "profile150r.flat"):
        xz(os.path.join(tmp, name), os.path.join(HERE, name + ".xz"))

    # three tiles, two read lengths
    sam_t = os.path.join(tmp, "prof_t.sam")
    run([py, SYN, "sam", ref, sam_t, "--pairs", "24000", "--seed", "13", "--read-len", "150", "--indel-rate", "0.0002",
         "--tiles", "1101,1102,2205", "--alt-len", "144", "--alt-frac", "0.2"])
    raw_t = os.path.join(tmp, "raw_t.reseq")
    run([ORACLE, "illuminaPE", "-j", "8", "-b", sam_t, "-r", ref, "--adapterFile", ADAPTERS + ".fa", "--adapterMatrix", ADAPTERS + ".mat",
         "--statsOnly", "--tiles", "-S", raw_t])
    log = run([ORACLE, "illuminaPE", "-j", "8", "-s", raw_t, "-r", ref, "--stopAfterEstimation"])
    if "did not reach precision aim" in log:
        raise SystemExit("IPF did not converge for every table of the tile profile")
    prof_t = os.path.join(tmp, "profile150t.reseq")
    run([DUMP, "patch", raw_t, prof_t, "5"])
    shutil.copy(raw_t + ".ipf", prof_t + ".ipf")
    run([DUMP, "profile", prof_t, os.path.join(tmp, "profile150t.flat")])
    for name in ("profile150t.reseq", "profile150t.reseq.ipf", "profile150t.flat"):
        xz(os.path.join(tmp, name), os.path.join(HERE, name + ".xz"))

    # realistic tables: 40 quality values, two tiles (IPF takes ~20 minutes on 16 threads); only the flat image and a pinned run are committed
    if os.environ.get("RSQ_GOLDEN_REALISTIC"):
        import hashlib
        import json
        sam_q = os.path.join(tmp, "prof_q.sam")
        run([py, SYN, "sam", ref, sam_q, "--pairs", "60000", "--seed", "19", "--quals", "40", "--tiles", "1101,2204", "--indel-rate", "0.00002"])
        raw_q = os.path.join(tmp, "raw_q.reseq")
        run([ORACLE, "illuminaPE", "-j", "16", "-b", sam_q, "-r", ref, "--adapterFile", ADAPTERS + ".fa", "--adapterMatrix", ADAPTERS + ".mat",
             "--statsOnly", "--tiles", "-S", raw_q])
        log = run([ORACLE, "illuminaPE", "-j", "16", "-s", raw_q, "-r", ref, "--stopAfterEstimation"])
        if "did not reach precision aim" in log:
            raise SystemExit("IPF did not converge for every table of the realistic profile")
        prof_q = os.path.join(tmp, "profile150q.reseq")
        run([DUMP, "patch", raw_q, prof_q, "5"])
        shutil.copy(raw_q + ".ipf", prof_q + ".ipf")
        run([DUMP, "profile", prof_q, os.path.join(tmp, "profile150q.flat")])
        xz(os.path.join(tmp, "profile150q.flat"), os.path.join(HERE, "profile150q.flat.xz"))
        q1, q2 = os.path.join(tmp, "q_R1.fq"), os.path.join(tmp, "q_R2.fq")
        run([ORACLE, "illuminaPE", "-j", "1", "--verbosity", "1", "-s", prof_q, "-R", os.path.join(HERE, "simref_small.fa"), "--ipfIterations", "0", "--seed", "42", "-c", "20", "-1", q1, "-2", q2])
        json.dump({"profile": "profile150q", "r1": hashlib.sha256(open(q1, "rb").read()).hexdigest(), "r2": hashlib.sha256(open(q2, "rb").read()).hexdigest(),
                   "pairs": open(q1, "rb").read().count(b"\n") // 4, "bytes": [os.path.getsize(q1), os.path.getsize(q2)]},
                  open(os.path.join(HERE, "profile150q_small_sha256.json"), "w"), indent=1, sort_keys=True)

    # 2x250 reads (config C4's read length)
    sam_l = os.path.join(tmp, "prof_250.sam")
    run([py, SYN, "sam", ref, sam_l, "--pairs", "20000", "--seed", "17", "--read-len", "250", "--indel-rate", "0.0001"])
    raw_l = os.path.join(tmp, "raw_250.reseq")
    run([ORACLE, "illuminaPE", "-j", "8", "-b", sam_l, "-r", ref, "--adapterFile", ADAPTERS + ".fa", "--adapterMatrix", ADAPTERS + ".mat",
         "--statsOnly", "-S", raw_l])
    log = run([ORACLE, "illuminaPE", "-j", "8", "-s", raw_l, "-r", ref, "--stopAfterEstimation"])
    if "did not reach precision aim" in log:
        raise SystemExit("IPF did not converge for every table of the 2x250 profile")
    prof_l = os.path.join(tmp, "profile250.reseq")
    run([DUMP, "patch", raw_l, prof_l, "5"])
    shutil.copy(raw_l + ".ipf", prof_l + ".ipf")
    run([DUMP, "profile", prof_l, os.path.join(tmp, "profile250.flat")])
    for name in ("profile250.reseq", "profile250.reseq.ipf", "profile250.flat"):
        xz(os.path.join(tmp, name), os.path.join(HERE, name + ".xz"))

    small = os.path.join(HERE, "simref_small.fa")
    run([py, SYN, "reference", small, "--sizes", "30000,22000,800,15000", "--seed", "21", "--n-rate", "0.004", "--prefix", "chr"])
    r1, r2 = os.path.join(tmp, "r1.fq"), os.path.join(tmp, "r2.fq")
    run([ORACLE, "illuminaPE", "-j", "1", "-s", prof, "-R", small, "--ipfIterations", "0", "--seed", "42", "-c", "20", "-1", r1, "-2", r2])
    xz(r1, os.path.join(HERE, "sim_small_seed42_R1.fq.xz"))
    xz(r2, os.path.join(HERE, "sim_small_seed42_R2.fq.xz"))

    # methylation (bisulfite) run on the same reference
    bed = os.path.join(HERE, "simref_small_meth.bed")
    write_bed(bed)
    m1, m2 = os.path.join(tmp, "m1.fq"), os.path.join(tmp, "m2.fq")
    run([ORACLE, "illuminaPE", "-j", "1", "-s", prof, "-R", small, "--ipfIterations", "0", "--seed", "42", "-c", "20", "--methylation", bed,
         "-1", m1, "-2", m2])
    xz(m1, os.path.join(HERE, "sim_small_meth_seed42_R1.fq.xz"))
    xz(m2, os.path.join(HERE, "sim_small_meth_seed42_R2.fq.xz"))

    frags = os.path.join(tmp, "em_frags.fa")
    run([py, SYN, "fragments", ref, "-", frags, "--n", "9000", "--len", "120", "--seed", "3"])
    emq = os.path.join(tmp, "em.fq")
    run([ORACLE, "seqToIllumina", "-j", "1", "-i", frags, "-s", prof, "--ipfIterations", "0", "--seed", "7", "-o", emq])
    xz(frags, os.path.join(HERE, "em_frags.fa.xz"))
    xz(emq, os.path.join(HERE, "em_seed7.fq.xz"))
    shutil.rmtree(tmp)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
