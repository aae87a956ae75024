#!/usr/bin/env python3
"""Golden fixtures for the VCF loader (reseq_b200/csrc/variants.hpp): synthetic VCFs over tests/golden/simref_small.fa and what the
UNMODIFIED reference holds in Reference::variants_ after reading them (oracle/_ref/dump_tables variants, built by oracle/Makefile).

    python tests/golden/make_variants_golden.py        (needs oracle/_ref/dump_tables, i.e. /root/reference at build time)

Writes simref_small_var*.vcf and simref_small_var*.variants.txt ("rejected" when the reference refuses the file) next to this script,
and sim_small_var{,_base}_seed42_R{1,2}.fq.xz: what `reseq illuminaPE -V <vcf>` (oracle/_ref/reseq_oracle, profile150, seed 42, coverage 20,
one thread) writes for the 5-allele and the 2-allele file - the parity target of the variant-aware kernels;
simref_small_var{,70}.varseq.txt.xz: 800 seeded calls each of the variant overload of Reference::ReferenceSequence with the reference's results."""
import lzma
import tempfile
import os
import random
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DUMP = os.path.join(ROOT, "oracle", "_ref", "dump_tables")
RESEQ = os.path.join(ROOT, "oracle", "_ref", "reseq_oracle")
REF = os.path.join(HERE, "simref_small.fa")


def header(names, n_samples):
    lines = ["##fileformat=VCFv4.2", "##source=reseq_b200 tests/golden/make_variants_golden.py"]
    lines += [f"##contig=<ID={n},length={ln}>" for n, ln in names]
    lines += ['##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">']
    lines += ["#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"s{k}" for k in range(n_samples))]
    return lines


def other(rng, base):
    return rng.choice([b for b in "ACGT" if b != base])


def records(rng, seqs, ploidies, per_seq):
    """Position-sorted, non-overlapping records of every shape ReadVariants distinguishes; N stretches are avoided (the
    reference compares REF after ReplaceN, the dump does not run ReplaceN)."""
    out = []
    for (name, _), seq in zip(NAMES, seqs):
        pos = rng.randrange(0, 40)
        for _ in range(per_seq):
            pos += rng.randrange(1, 2 * len(seq) // per_seq // 3 + 2)
            kind = rng.choice(["snp", "snp", "mnp", "ins", "del", "complex_long", "complex_short", "multi", "lower", "same"])
            rlen = {"snp": 1, "mnp": 3, "ins": 1, "del": rng.randrange(2, 9), "complex_long": 3, "complex_short": 4, "multi": 2, "lower": 1, "same": 2}[kind]
            if pos + rlen + 1 >= len(seq):
                break
            ref = seq[pos:pos + rlen]
            if "N" in ref:
                pos += rlen
                continue
            if kind == "snp":
                alts = [other(rng, ref)]
            elif kind == "mnp":
                alts = [other(rng, ref[0]) + ref[1] + other(rng, ref[2])]
            elif kind == "ins":
                alts = [ref + "".join(rng.choice("ACGT") for _ in range(rng.randrange(1, 12)))]
            elif kind == "del":
                alts = [ref[0]]
            elif kind == "complex_long":
                alts = [other(rng, ref[0]) + ref[1] + ref[2] + "".join(rng.choice("ACGT") for _ in range(2))]
            elif kind == "complex_short":
                alts = [ref[0] + other(rng, ref[1])]
            elif kind == "multi":
                alts = [ref[0] + other(rng, ref[1]), ref + "ACT", ref[0], ref + "ACT"[:2] + "G"]
            elif kind == "lower":
                alts = [other(rng, ref).lower()]
            else:
                alts = [ref[0] + other(rng, ref[1]), ref]   # second alternative equals the reference: nothing to insert for it
            gts = []
            for ploidy in ploidies:
                sep = rng.choice("|/")
                gts.append(sep.join(str(rng.randrange(0, len(alts) + 1)) for _ in range(ploidy)) + rng.choice(["", ":12", ":3:0.5"]))
            out.append((name, pos + 1, ref, ",".join(alts), gts))
            pos += rlen
    return out


def write_vcf(path, recs, n_samples, names=None):
    with open(path, "w") as f:
        f.write("\n".join(header(names or NAMES, n_samples)) + "\n")
        for name, pos, ref, alt, gts in recs:
            f.write(f"{name}\t{pos}\t.\t{ref}\t{alt}\t{30 + pos % 7}\tPASS\tDP=20\tGT" + "".join("\t" + g for g in gts) + "\n")


def dump(vcf):
    out = vcf[:-4] + ".variants.txt"
    rc = subprocess.run([DUMP, "variants", REF, vcf, out], stderr=subprocess.DEVNULL, stdout=subprocess.DEVNULL).returncode
    text = open(out).read()
    assert (rc != 0) == text.startswith("rejected"), (vcf, rc)
    return text


def write_ends_vcf(path, ids, seqs):
    """simref_small_var_ends.vcf: variants on the first and last bases of every sequence (surroundings that roll around the sequence ends, fragments
    that cannot end inside the sequence) and insertions of 30-140 bases (longer than a read; fragments that start, end or lie completely inside one)."""
    rng = random.Random(5)

    def other(b):
        return rng.choice([x for x in "ACGT" if x != b])
    lines = ["##fileformat=VCFv4.2"] + [f"##contig=<ID={n},length={len(s)}>" for n, s in zip(ids, seqs)]
    lines += ['##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">', "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts0\ts1"]
    for n, s in zip(ids, seqs):
        L = len(s)
        if L < 2000:
            continue
        pos_list = sorted(set(p for p in [0, 3, 7, 12, 18, 25, 31, 40, 500, 1001, 1999, 2000, L - 45, L - 33, L - 28, L - 21, L - 15, L - 9, L - 5, L - 2, L - 1] if 0 <= p < L))
        last = -1
        for p in pos_list:
            if p <= last:
                continue
            kind = rng.choice(["snp", "ins", "del", "longins", "multi"])
            if "N" in s[p:p + 4]:
                continue
            end = p
            if kind == "snp":
                ref = s[p]
                alt = other(ref)
            elif kind == "ins":
                ref = s[p]
                alt = ref + "".join(rng.choice("ACGT") for _ in range(rng.randrange(1, 6)))
            elif kind == "longins":
                ref = s[p]
                alt = ref + "".join(rng.choice("ACGT") for _ in range(rng.randrange(30, 140)))
            elif kind == "del":
                k = rng.randrange(2, 5)
                if p + k >= L:
                    continue
                ref = s[p:p + k]
                alt = ref[0]
                end = p + k - 1
            else:
                ref = s[p]
                alt = other(ref) + "," + ref + "GG"
            nalt = alt.count(",") + 1
            gts = ["|".join(str(rng.randrange(0, nalt + 1)) for _ in range(2)) for _ in range(2)]
            lines.append(f"{n}\t{p + 1}\t.\t{ref}\t{alt}\t30\tPASS\tDP=20\tGT\t" + "\t".join(gts))
            last = end
    open(path, "w").write("\n".join(lines) + "\n")


def main():
    global NAMES
    ids, seqs = [], []
    for line in open(REF):
        line = line.rstrip("\n")
        if line.startswith(">"):
            ids.append(line[1:].split(" ")[0])
            seqs.append([])
        else:
            seqs[-1].append(line.upper())
    seqs = ["".join(s) for s in seqs]
    NAMES = [(i, len(s)) for i, s in zip(ids, seqs)]
    write_ends_vcf(os.path.join(HERE, "simref_small_var_ends.vcf"), ids, seqs)

    rng = random.Random(20261017)
    good = records(rng, seqs, [2, 2, 1], 60)                 # three populations, 5 alleles
    write_vcf(os.path.join(HERE, "simref_small_var.vcf"), good, 3)
    wide = records(rng, seqs, [1] * 70, 12)                  # 70 haploid populations: allele bits beyond the first word
    write_vcf(os.path.join(HERE, "simref_small_var70.vcf"), wide, 70)
    base = records(rng, seqs, [2], 8)

    def variant(tag, edit, names=None):
        recs = [list(r) for r in base]
        edit(recs)
        write_vcf(os.path.join(HERE, f"simref_small_var_bad_{tag}.vcf"), recs, 1, names)

    def swap(recs):
        recs[2], recs[3] = recs[3], recs[2]

    def overlap(recs):
        recs[1][2] = seqs[0][recs[1][1] - 1:recs[2][1]]      # REF runs into the next record
        recs[1][3] = recs[1][2][0]

    def wrong_ref(recs):
        recs[4][2] = other(rng, recs[4][2][0]) + recs[4][2][1:]

    def alt_n(recs):
        recs[0][3] = "N"
        recs[0][4] = ["0|1"]

    def gt_index(recs):
        recs[5][4] = ["0|9"]

    def gt_few(recs):
        recs[3][4] = ["1"]

    def gt_many(recs):
        recs[3][4] = ["1|0|1"]

    def gt_char(recs):
        recs[3][4] = [".|1"]

    def seq_order(recs):
        recs.append(list(recs[0]))

    def past_end(recs):
        recs[7][1] = NAMES[0][1] + 5

    variant("unsorted", swap)
    variant("overlap", overlap)
    variant("wrong_ref", wrong_ref)
    variant("alt_n", alt_n)
    variant("gt_index", gt_index)
    variant("gt_few", gt_few)
    variant("gt_many", gt_many)
    variant("gt_char", gt_char)
    variant("seq_order", seq_order)
    variant("past_end", past_end)
    variant("contig_names", lambda recs: None, [("chrX", NAMES[0][1])] + NAMES[1:])
    variant("contig_count", lambda recs: None, NAMES[:3])
    write_vcf(os.path.join(HERE, "simref_small_var_base.vcf"), base, 1)   # the unedited base file: accepted

    for name in sorted(os.listdir(HERE)):
        if name.startswith("simref_small_var") and name.endswith(".vcf"):
            text = dump(os.path.join(HERE, name))
            print(f"{name}: {'rejected' if text.startswith('rejected') else str(text.count(chr(10)) - 1) + ' variants'}")


def simulate_with_reference():
    with tempfile.TemporaryDirectory() as tmp:
        for name in ("profile150.reseq", "profile150.reseq.ipf"):
            with lzma.open(os.path.join(HERE, name + ".xz")) as f, open(os.path.join(tmp, name), "wb") as o:
                o.write(f.read())
        for tag in ("var", "var_base"):
            r1, r2 = os.path.join(tmp, tag + "_R1.fq"), os.path.join(tmp, tag + "_R2.fq")
            subprocess.run([RESEQ, "illuminaPE", "-j", "1", "--verbosity", "1", "-s", os.path.join(tmp, "profile150.reseq"), "-R", REF, "--ipfIterations", "0",
                            "--seed", "42", "-c", "20", "-1", r1, "-2", r2, "-V", os.path.join(HERE, f"simref_small_{tag}.vcf")],
                           check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            for path, seg in ((r1, "R1"), (r2, "R2")):
                with open(path, "rb") as f, lzma.open(os.path.join(HERE, f"sim_small_{tag}_seed42_{seg}.fq.xz"), "wb", preset=9 | lzma.PRESET_EXTREME) as o:
                    o.write(f.read())


def spliced_sequences():
    """simref_small_var{,70}.varseq.txt.xz: seeded calls of Reference::ReferenceSequence (variant overload) with their results."""
    for tag, seed in (("var", 7), ("var70", 8)):
        with tempfile.TemporaryDirectory() as tmp:
            out = os.path.join(tmp, "calls.txt")
            subprocess.run([DUMP, "varseq", REF, os.path.join(HERE, f"simref_small_{tag}.vcf"), str(seed), "800", out], check=True)
            with open(out, "rb") as f, lzma.open(os.path.join(HERE, f"simref_small_{tag}.varseq.txt.xz"), "wb", preset=9) as o:
                o.write(f.read())


def allele_choices():
    """choose_alleles_seed11.txt.xz: 1500 seeded calls of Simulator::SelectAllele/ReverseSelection as ChooseAlleles combines them."""
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "alleles.txt")
        subprocess.run([DUMP, "alleles", "11", "1500", out], check=True)
        with open(out, "rb") as f, lzma.open(os.path.join(HERE, "choose_alleles_seed11.txt.xz"), "wb", preset=9) as o:
            o.write(f.read())


def sys_error_walks():
    """sys_error_variants_seed5.txt.xz: a seeded SimBlock chain with SysErrorVariants and 300 walks of Simulator::GetSysErrorFromBlock."""
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "walks.txt")
        subprocess.run([DUMP, "syserrvar", "5", "12", "300", out], check=True)
        with open(out, "rb") as f, lzma.open(os.path.join(HERE, "sys_error_variants_seed5.txt.xz"), "wb", preset=9) as o:
            o.write(f.read())


def bias_modifier_traces():
    """bias_mod_trace_seq{0,1}.txt.xz: Simulator's VariantBiasVarModifiers bookkeeping (PrepareBiasModForCurrentStartPos, GetPossibleAlleles,
    PrepareBiasModForCurrentFragmentLength, GetGCPercent, Start/EndVariant, CheckForInsertedBasesToStartFrom) driven over start positions
    the way SimulateFromGivenBlock does, for simref_small_var.vcf (5 alleles), with the sequence after ReplaceN in the first line."""
    # the third trace evaluates EVERY fragment length 50..124 for the starts around a substitution, two deletions and a 5-base insertion
    # (positions 3842-3893 of chr1): fragments ending inside the insertion, starts on its inserted bases
    # the fourth runs over the multi-allelic site of the 70-allele file (deletion, substitution and two different insertions at chr1:5115)
    for name, tag, seq, start, n, len_to, sparsity in (("seq0", "var", 0, 0, 1500, 400, 24), ("seq1", "var", 1, 9000, 1000, 400, 24),
                                                      ("seq0_dense", "var", 0, 3780, 120, 125, 1), ("var70", "var70", 0, 4900, 230, 300, 40)):
        with tempfile.TemporaryDirectory() as tmp:
            out = os.path.join(tmp, "trace.txt")
            subprocess.run([DUMP, "biasmod", REF, os.path.join(HERE, f"simref_small_{tag}.vcf"), "42", str(seq), str(start), str(n), "50", str(len_to), out, str(sparsity)],
                           check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            with open(out, "rb") as f, lzma.open(os.path.join(HERE, f"bias_mod_trace_{name}.txt.xz"), "wb", preset=9 | lzma.PRESET_EXTREME) as o:
                o.write(f.read())


def allele_fragment_counts():
    """fragment_counts_alleles_seed3.txt.xz: 2500 seeded (mean, dispersion parameters, allele count, uniform) -> count of the reference."""
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "negbin.txt")
        subprocess.run([DUMP, "negbin", "3", "2500", out], check=True)
        with open(out, "rb") as f, lzma.open(os.path.join(HERE, "fragment_counts_alleles_seed3.txt.xz"), "wb", preset=9) as o:
            o.write(f.read())


if __name__ == "__main__":
    main()
    simulate_with_reference()
    spliced_sequences()
    allele_choices()
    sys_error_walks()
    bias_modifier_traces()
    allele_fragment_counts()
