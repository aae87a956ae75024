"""VCF loading for `-V/--vcfSim` (reseq_b200/csrc/variants.hpp, rsq_reference_load_variants): the host half of SURVEY §8 row a6.

Pinned three ways: (1) the reference's own known-answer test for test-var.vcf (ReferenceTest.cpp:138-232 TestVariationLoading,
242-257 TestVariationPositionLoading), restated below; (2) fixtures under tests/golden/ written by the UNMODIFIED reference
(`oracle/_ref/dump_tables variants`, tests/golden/make_variants_golden.py) for synthetic VCFs with several populations, 70 alleles,
multi-allelic / MNP / InDel / complex records, and for files the reference rejects; (3) the same through the C ABI."""
import glob
import gzip
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_TEST = "/root/reference/test"
VCFS = sorted(glob.glob(os.path.join(GOLDEN, "simref_small_var*.vcf")))

# ReferenceTest.cpp:152-231: position_, var_seq_, InAllele(0), InAllele(1) of the 13 variants
KAT = [(2, "T", 0, 1), (3, "A", 0, 1), (4, "A", 0, 1), (5, "G", 0, 1), (7, "A", 0, 1), (8, "TTTTTCAGCTTTTCA", 0, 1), (11368, "T", 0, 1),
       (11370, "T", 0, 1), (953165, "G", 0, 1), (3192437, "G", 0, 1), (3192438, "", 0, 1), (3424235, "C", 1, 1), (3424236, "A", 1, 1)]
# ReferenceTest.cpp:252
KAT_POSITIONS = [0, 1, 2, 3, 4, 5, 6, 7, 8, 16, 20, 11368, 11369, 11370, 953165, 3192437, 3192438, 3424235, 3424236]


@pytest.fixture(scope="module")
def var_check(workdir):
    exe = os.path.join(workdir, "variants_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "host_twin", "variants_check.cpp"), "-lz"], check=True)
    return exe


def parse(text):
    lines = text.strip().split("\n")
    assert lines[0].startswith("alleles ")
    out = []
    for line in lines[1:]:
        seq, pos, bases, lo, hi = line.split(" ")
        out.append((int(seq), int(pos), "" if bases == "-" else bases, int(lo, 16) | (int(hi, 16) << 64)))
    return int(lines[0].split(" ")[1]), out


def test_fixture_set_is_complete():
    assert len(VCFS) == 16
    assert sum(open(v[:-4] + ".variants.txt").read().startswith("rejected") for v in VCFS) == 12    # accepted: base, var, var70, ends


@pytest.mark.skipif(not os.path.isdir(REF_TEST), reason="the reference's test data is only present in the build container")
def test_reference_known_answers(var_check):
    fa, vcf = os.path.join(REF_TEST, "ecoli-GCF_000005845.2_ASM584v2_genomic.fa"), os.path.join(REF_TEST, "test-var.vcf")
    alleles, got = parse(subprocess.run([var_check, fa, vcf], check=True, capture_output=True, text=True).stdout)
    assert alleles == 2
    assert got == [(0, p, s, a0 | (a1 << 1)) for p, s, a0, a1 in KAT]
    pos = subprocess.run([var_check, fa, vcf, "positions"], check=True, capture_output=True, text=True).stdout
    assert [int(line.split(" ")[1]) for line in pos.strip().split("\n")] == KAT_POSITIONS


@pytest.mark.parametrize("vcf", VCFS, ids=[os.path.basename(v)[len("simref_small_"):-4] for v in VCFS])
def test_loader_matches_reference_memory(var_check, vcf):
    want = open(vcf[:-4] + ".variants.txt").read()
    res = subprocess.run([var_check, os.path.join(GOLDEN, "simref_small.fa"), vcf], capture_output=True, text=True)
    assert res.stdout == want
    assert (res.returncode != 0) == want.startswith("rejected")
    if res.returncode:
        assert res.stderr.strip()   # the reference's diagnostic, not a bare failure


def test_allele_bits_beyond_the_first_word():
    alleles, got = parse(open(os.path.join(GOLDEN, "simref_small_var70.variants.txt")).read())
    assert alleles == 70 and any(bits >> 64 for _, _, _, bits in got)


def test_same_position_order_and_merging(var_check):
    """ReferenceTest.cpp:86-136 TestInsertVariant: at one position deletion < substitution < insertions by length, equal replacements merge their alleles."""
    src = os.path.join(ROOT, "tests", "host_twin", "variants_check.cpp")
    assert os.path.exists(src)
    from_fixture = parse(open(os.path.join(GOLDEN, "simref_small_var.variants.txt")).read())[1]
    by_pos = {}
    for seq, pos, bases, bits in from_fixture:
        by_pos.setdefault((seq, pos), []).append((bases, bits))
    multi = [v for v in by_pos.values() if len(v) > 1]
    assert multi
    for entries in multi:
        lengths = [len(b) for b, _ in entries]
        assert lengths == sorted(lengths) and len({b for b, _ in entries}) == len(entries)


def test_c_abi_variants(library, workdir):
    import reseq_b200 as rb
    ref = rb.Reference.load_fasta(os.path.join(GOLDEN, "simref_small.fa"))
    assert ref.num_alleles == 1
    vcf = os.path.join(GOLDEN, "simref_small_var.vcf")
    gz = os.path.join(workdir, "v.vcf.gz")
    with open(vcf, "rb") as f, gzip.open(gz, "wb") as o:
        o.write(f.read())
    alleles, want = parse(open(vcf[:-4] + ".variants.txt").read())
    for path in (vcf, gz):   # VcfFileIn reads gzip as well
        ref.load_variants(path)
        assert ref.num_alleles == alleles == 5
        got = [(s, p, b, bits) for s in range(ref.num_sequences) for p, b, bits in ref.variants(s)]
        assert got == want
    ref70 = rb.Reference.load_fasta(os.path.join(GOLDEN, "simref_small.fa"))
    ref70.load_variants(os.path.join(GOLDEN, "simref_small_var70.vcf"))
    assert [(s, p, b, bits) for s in range(4) for p, b, bits in ref70.variants(s)] == parse(open(os.path.join(GOLDEN, "simref_small_var70.variants.txt")).read())[1]


def test_c_abi_rejections_carry_the_reference_diagnostics(library):
    import reseq_b200 as rb
    ref = rb.Reference.load_fasta(os.path.join(GOLDEN, "simref_small.fa"))
    expect = {"unsorted": "not properly position sorted", "overlap": "overlaps with a previous variant", "wrong_ref": "is not identical with the specified reference",
              "alt_n": "alternative column containing ambiguous bases", "gt_index": "Variant number 9 does not exist", "gt_few": "Could not find enough alleles",
              "gt_char": "Unallowed character '.'", "contig_names": "do not match between reference(chr1) and variant(chrX)", "contig_count": "Number of contigs does not match",
              "past_end": "starts after the end of the reference sequence", "seq_order": "Found sequence id 0 after id 3", "gt_many": "Could not read vcf record"}
    for tag, text in expect.items():
        with pytest.raises(rb.RsqError) as info:
            ref.load_variants(os.path.join(GOLDEN, f"simref_small_var_bad_{tag}.vcf"))
        assert text in str(info.value), tag
    with pytest.raises(rb.RsqError):
        ref.load_variants(os.path.join(GOLDEN, "does_not_exist.vcf"))


def test_reference_made_fastq_targets_for_variant_runs():
    """Parity targets of the variant-aware kernels (not built yet): FASTQ the unmodified reference wrote with -V for the 5- and the 2-allele
    file. Checked here only for what CreateReadId promises (Simulator.cpp:596-632): `_allele<a>` with a < NumAlleles in every id, pairs in step."""
    import lzma
    import re
    for tag, alleles in (("var", 5), ("var_base", 2)):
        ids = []
        for seg in ("R1", "R2"):
            text = lzma.open(os.path.join(GOLDEN, f"sim_small_{tag}_seed42_{seg}.fq.xz")).read().decode().split("\n")
            ids.append([line for line in text[0::4] if line])
        assert len(ids[0]) == len(ids[1]) > 4000
        seen = set()
        for a, b in zip(*ids):
            m = re.match(r"@ReseqRead\d+_\d+_allele(\d+):", a)
            assert m and a.split(":")[0] == b.split(":")[0]
            seen.add(int(m.group(1)))
        assert seen == set(range(alleles))
