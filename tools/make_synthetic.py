#!/usr/bin/env python3
"""Deterministic synthetic inputs for the ReSeq hot path (no data ships with the reference).

  make_synthetic.py reference <out.fa> --sizes 200000,120000 --seed 7 [--n-rate 0.0]
      Random reference with slowly varying GC content (piecewise GC targets) so GC bias matters.

  make_synthetic.py sam <ref.fa> <out.sam> --pairs 60000 --read-len 150 --seed 11
      Position-sorted paired-end SAM "mapped" to <ref.fa>: FR pairs, log-normal fragment lengths
      (some shorter than the read -> adapter read-through, written the way an end-to-end aligner
      reports them), Markov base qualities, quality-driven substitutions, a few indels and
      motif-triggered systematic errors.  It is the input to the reference's own stats/IPF step
      (`reseq illuminaPE -b ... --statsOnly`), which is how profiles for tests and bench are made.

  make_synthetic.py fragments <ref.fa> <syserr.fq|-> <out.fa> --n 10000 --len 300 --seed 3
      seqToIllumina input: records "<id> <seg>;<fraglen>;<dom-err>;<err-rate>" (Simulator.cpp:2423-2485).
"""
import argparse
import sys

import numpy as np

ADAPTER1 = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCACATCACGATCTCGTATGCCGTCTTCTGCTTG"
ADAPTER2 = "AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGTAGATCTCGGTGGTCGCCGTATCATT"
COMP = str.maketrans("ACGTN", "TGCAN")


def revcomp(s):
    return s.translate(COMP)[::-1]


def gen_reference(sizes, seed, n_rate=0.0):
    rng = np.random.default_rng(seed)
    seqs = []
    for size in sizes:
        out = np.empty(size, dtype=np.uint8)
        pos = 0
        while pos < size:
            seg = int(rng.integers(2000, 12000))
            gc = float(rng.uniform(0.30, 0.65))
            n = min(seg, size - pos)
            p = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
            out[pos:pos + n] = rng.choice(4, size=n, p=p)
            pos += n
        s = np.frombuffer(b"ACGT", dtype=np.uint8)[out]
        if n_rate > 0:
            nruns = max(1, int(size * n_rate / 50))
            for _ in range(nruns):
                st = int(rng.integers(0, size - 100))
                ln = int(rng.integers(1, 100))
                s[st:st + ln] = ord("N")
        seqs.append(s.tobytes().decode())
    return seqs


def write_vcf(path, names, seqs, seed, snp_rate=1e-3, indel_rate=1e-4, max_indel=20, ploidy=2):
    """Phased diploid VCF for a synthetic reference (BASELINE config C5: 1 SNP per kb, 1 indel of <= 20 bases per 10 kb, 2 alleles): records
    at least 32 bases apart, genotypes 0|1, 1|0 or 1|1, insertions and deletions in equal shares.  Returns the number of records."""
    rng = np.random.default_rng(seed)
    n_rec = 0
    with open(path, "w") as f:
        f.write("##fileformat=VCFv4.2\n##source=reseq_b200 tools/make_synthetic.py\n")
        for name, s in zip(names, seqs):
            f.write(f"##contig=<ID={name},length={len(s)}>\n")
        f.write('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts0\n')
        gts = ["0|1", "1|0", "1|1"] if ploidy == 2 else ["1"]
        for name, s in zip(names, seqs):
            L = len(s)
            n = int(L * (snp_rate + indel_rate))
            pos = np.unique(rng.integers(1, max(2, (L - 64) // 32), size=n)) * 32   # sorted, >= 32 apart
            kinds = rng.random(len(pos)) < indel_rate / (snp_rate + indel_rate)
            lens = rng.integers(1, max_indel + 1, size=len(pos))
            ins = rng.random(len(pos)) < 0.5
            alt_idx = rng.integers(1, 4, size=len(pos))
            gt_idx = rng.integers(0, len(gts), size=len(pos))
            ins_bases = rng.integers(0, 4, size=int(lens[kinds & ins].sum()) + 1)
            ib = 0
            out = []
            for k in range(len(pos)):
                p = int(pos[k])
                ref = s[p]
                if ref == "N":
                    continue
                if not kinds[k]:
                    alt = "ACGT"[("ACGT".index(ref) + int(alt_idx[k])) % 4]
                    out.append(f"{name}\t{p + 1}\t.\t{ref}\t{alt}\t40\tPASS\t.\tGT\t{gts[gt_idx[k]]}\n")
                elif ins[k]:
                    n_ins = int(lens[k])
                    alt = ref + "".join("ACGT"[b] for b in ins_bases[ib:ib + n_ins])
                    ib += n_ins
                    out.append(f"{name}\t{p + 1}\t.\t{ref}\t{alt}\t40\tPASS\t.\tGT\t{gts[gt_idx[k]]}\n")
                else:
                    refd = s[p:p + 1 + int(lens[k])]
                    if "N" in refd or len(refd) < 2:
                        continue
                    out.append(f"{name}\t{p + 1}\t.\t{refd}\t{ref}\t40\tPASS\t.\tGT\t{gts[gt_idx[k]]}\n")
                if len(out) >= 100000:
                    f.write("".join(out))
                    n_rec += len(out)
                    out = []
            f.write("".join(out))
            n_rec += len(out)
    return n_rec


def write_fasta(path, seqs, prefix="synth"):
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(f">{prefix}{i + 1} synthetic contig {i + 1}\n")
            for j in range(0, len(s), 80):
                f.write(s[j:j + 80])
                f.write("\n")


def read_fasta(path):
    names, seqs, cur = [], [], []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if line.startswith(">"):
                if names:
                    seqs.append("".join(cur))
                names.append(line[1:].split(" ")[0])
                cur = []
            else:
                cur.append(line)
    seqs.append("".join(cur))
    return names, seqs


QUALS = np.array([2, 12, 16, 20, 24, 27, 30, 33, 36, 38, 40])


INDEL_RATE = 0.0008   # per base, each of insertion and deletion (gen_sam --indel-rate overrides)


def simulate_read(rng, template, read_len, seg):
    """template: the fragment strand this read is sequenced from (string, already oriented), followed by adapter.
    Returns (seq, qual, cigar-ops list of (op,len) relative to template consumption)."""
    nq = len(QUALS)
    level = int(rng.integers(nq // 2, nq))  # read-wide quality level
    step = max(1, nq // 11)   # quality levels per Markov step (1 for the default 11 levels)
    seq = []
    qual = []
    cigar = []
    tpos = 0
    state = level
    tl = len(template)

    def push(op):
        if cigar and cigar[-1][0] == op:
            cigar[-1][1] += 1
        else:
            cigar.append([op, 1])

    while len(seq) < read_len and tpos < tl:
        r = rng.random()
        if r < INDEL_RATE and 5 < len(seq) < read_len - 5:
            # insertion
            seq.append("ACGT"[int(rng.integers(0, 4))])
            qual.append(int(QUALS[max(0, state - 2)]))
            push("I")
            continue
        if r < 2 * INDEL_RATE and 5 < len(seq) < read_len - 5:
            tpos += 1
            push("D")
            continue
        # quality Markov step, degrading along the read
        drift = 0.10 + 0.25 * len(seq) / read_len + (0.05 if seg else 0.0)
        u = rng.random()
        if u < drift and state > 0:
            state = max(0, state - (int(rng.integers(1, step + 1)) if step > 1 else 1))   # (no extra draw for the default levels: the committed fixtures regenerate)
        elif u > 0.80 and state < level:
            state = min(level, state + (int(rng.integers(1, step + 1)) if step > 1 else 1))
        q = int(QUALS[state])
        base = template[tpos]
        # systematic error: after "GGC" on the read strand the next base tends to be called as G
        perr = 10 ** (-q / 10.0)
        sys_err = tpos >= 3 and template[tpos - 3:tpos] == "GGC"
        called = base
        if sys_err and base != "G" and rng.random() < 0.30:
            called = "G"
            if rng.random() < 0.7 and state > 1:
                q = int(QUALS[state - 2])
        elif rng.random() < perr:
            called = "ACGT".replace(base, "")[int(rng.integers(0, 3))]
        seq.append(called)
        qual.append(q)
        push("M")
        tpos += 1
    return "".join(seq), qual, cigar, tpos


def gen_sam(ref_path, out_path, pairs, read_len, seed, tiles=None, alt_len=0, alt_frac=0.0):
    """tiles: list of tile numbers -> Casava >= 1.8 read names (instrument:run:flowcell:lane:TILE:x:y) and a tile-dependent
    quality level; alt_len/alt_frac: that share of the reads is `alt_len` bases long instead of read_len (both lengths end up in the profile)."""
    rng = np.random.default_rng(seed)
    names, seqs = read_fasta(ref_path)
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    records = []
    gc_pref = []
    for s in seqs:
        a = np.frombuffer(s.encode(), dtype=np.uint8)
        gc_pref.append(np.concatenate([[0], np.cumsum((a == ord("G")) | (a == ord("C")))]))
    n_done = 0
    while n_done < pairs:
        rid = int(rng.choice(len(seqs), p=lens / lens.sum()))
        flen = int(np.exp(rng.normal(np.log(300), 0.35)))
        flen = max(40, min(900, flen))
        L = int(lens[rid])
        start = int(rng.integers(read_len, L - flen - read_len))
        frag = seqs[rid][start:start + flen]
        if "N" in frag:
            continue
        # GC bias: accept with a bump-shaped probability
        gc = (gc_pref[rid][start + flen] - gc_pref[rid][start]) / flen
        if rng.random() > np.exp(-((gc - 0.48) / 0.12) ** 2):
            continue
        strand = int(rng.integers(0, 2))  # 0: first read forward
        polya = "A" * 40
        fwd_template = frag + (ADAPTER1 if strand == 0 else ADAPTER2) + polya
        rev_template = revcomp(frag) + (ADAPTER2 if strand == 0 else ADAPTER1) + polya
        # the forward-mapping read is segment `strand`; the reverse-mapping read the other one
        len_f = alt_len if (alt_len and rng.random() < alt_frac) else read_len
        len_r = alt_len if (alt_len and rng.random() < alt_frac) else read_len
        fseq, fqual, fcig, _ = simulate_read(rng, fwd_template, len_f, strand)
        rseq, rqual, rcig, _ = simulate_read(rng, rev_template, len_r, 1 - strand)
        if len(fseq) < len_f or len(rseq) < len_r:
            continue

        def ref_span(cig):
            return sum(n for op, n in cig if op in "MD")

        fpos = start
        # reverse read in reference orientation
        rseq_ref = revcomp(rseq)
        rqual_ref = rqual[::-1]
        rcig_ref = rcig[::-1]
        rpos = start + flen - ref_span(rcig)
        if rpos < 0 or fpos + ref_span(fcig) > L:
            continue
        cig_f = "".join(f"{n}{op}" for op, n in fcig)
        cig_r = "".join(f"{n}{op}" for op, n in rcig_ref)
        qname = f"sim{n_done}"
        if tiles:
            tile_idx = int(rng.integers(0, len(tiles)))
            qname = f"SYN1:7:FCSYN:1:{tiles[tile_idx]}:{int(rng.integers(1000, 20000))}:{n_done + 1000}"
            if tile_idx:   # later tiles are a bit worse: the per-tile tables differ
                fqual = [max(2, q - 2 * tile_idx) if (i % 3 == 0) else q for i, q in enumerate(fqual)]
                rqual = [max(2, q - 2 * tile_idx) if (i % 3 == 0) else q for i, q in enumerate(rqual)]
                rqual_ref = rqual[::-1]
        flag_f = 1 | 2 | 32 | (64 if strand == 0 else 128)
        flag_r = 1 | 2 | 16 | (128 if strand == 0 else 64)
        tlen = flen
        qf = "".join(chr(q + 33) for q in fqual)
        qr = "".join(chr(q + 33) for q in rqual_ref)
        records.append((rid, fpos, f"{qname}\t{flag_f}\t{names[rid]}\t{fpos + 1}\t42\t{cig_f}\t=\t{rpos + 1}\t{tlen}\t{fseq}\t{qf}"))
        records.append((rid, rpos, f"{qname}\t{flag_r}\t{names[rid]}\t{rpos + 1}\t42\t{cig_r}\t=\t{fpos + 1}\t{-tlen}\t{rseq_ref}\t{qr}"))
        n_done += 1
    records.sort(key=lambda r: (r[0], r[1]))
    with open(out_path, "w") as f:
        f.write("@HD\tVN:1.0\tSO:coordinate\n")
        for n, s in zip(names, seqs):
            f.write(f"@SQ\tSN:{n}\tLN:{len(s)}\n")
        for _, _, line in records:
            f.write(line)
            f.write("\n")


def gen_fragments(ref_path, sys_path, out_path, n, flen, seed):
    rng = np.random.default_rng(seed)
    names, seqs = read_fasta(ref_path)
    s = seqs[0].replace("N", "A")
    with open(out_path, "w") as f:
        for i in range(n):
            seg = i % 2
            st = int(rng.integers(0, len(s) - flen))
            frag = s[st:st + flen]
            if seg:
                frag = revcomp(frag)
            # sparse synthetic systematic errors: ~1% of positions, rates 5..60 (even above 86 is legal)
            dom = np.frombuffer(frag.encode(), dtype=np.uint8).copy()
            rate = np.zeros(flen, dtype=np.int64)
            hits = rng.random(flen) < 0.01
            rate[hits] = rng.integers(5, 95, size=int(hits.sum()))
            dom[hits] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(hits.sum()))]
            comp = np.where(rate > 86, rate - (rate - 85) // 2, rate)
            f.write(f">frag{i} {seg + 1};{flen};{dom.tobytes().decode()};{''.join(chr(int(c) + 33) for c in comp)}\n{frag}\n")


def main():
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    a = sub.add_parser("reference")
    a.add_argument("out")
    a.add_argument("--sizes", default="200000,120000")
    a.add_argument("--seed", type=int, default=7)
    a.add_argument("--n-rate", type=float, default=0.0)
    a.add_argument("--prefix", default="synth")
    b = sub.add_parser("sam")
    b.add_argument("ref")
    b.add_argument("out")
    b.add_argument("--pairs", type=int, default=60000)
    b.add_argument("--read-len", type=int, default=150)
    b.add_argument("--seed", type=int, default=11)
    b.add_argument("--indel-rate", type=float, default=0.0008, help="per-base rate of insertions, and of deletions")
    b.add_argument("--tiles", default="", help="comma separated tile numbers (Casava 1.8 read names)")
    b.add_argument("--alt-len", type=int, default=0)
    b.add_argument("--alt-frac", type=float, default=0.0)
    b.add_argument("--quals", type=int, default=0, help="number of distinct base qualities, evenly spread over 2..41 (default: the 11 values of QUALS); 40 = every value like a real HiSeq/NovaSeq run before binning")
    c = sub.add_parser("fragments")
    c.add_argument("ref")
    c.add_argument("sys")
    c.add_argument("out")
    c.add_argument("--n", type=int, default=10000)
    c.add_argument("--len", type=int, default=300)
    c.add_argument("--seed", type=int, default=3)
    args = ap.parse_args()
    if args.cmd == "reference":
        write_fasta(args.out, gen_reference([int(x) for x in args.sizes.split(",")], args.seed, args.n_rate), args.prefix)
    elif args.cmd == "sam":
        global INDEL_RATE, QUALS
        INDEL_RATE = args.indel_rate
        if args.quals:
            QUALS = np.unique(np.round(np.linspace(2, 41, args.quals)).astype(int))
        gen_sam(args.ref, args.out, args.pairs, args.read_len, args.seed, [int(t) for t in args.tiles.split(",")] if args.tiles else None, args.alt_len, args.alt_frac)
    else:
        gen_fragments(args.ref, args.sys, args.out, args.n, args.len, args.seed)


if __name__ == "__main__":
    sys.exit(main())
