"""splice_reference (reseq_b200/csrc/variant_core.cuh): the variant overload of Reference::ReferenceSequence (Reference.cpp:498-567), i.e.
the variant-aware half of Simulator::GetOrgSeq. Host/device source exercised on the CPU.

Pinned by the reference's own known answers (ReferenceTest.cpp:286-334) and by 2 x 800 seeded calls answered by the unmodified reference
(`oracle/_ref/dump_tables varseq`, committed as tests/golden/simref_small_var{,70}.varseq.txt.xz)."""
import lzma
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# reference-test.fa sequence 0 (ReferenceTest.cpp:274); the calls below only touch its first and last 12 bases
SEQ0 = ("AGCTTTTCATTCTGACTGCAACGGGCAATATGTCTCTGTGTGGATTAAAAAAAGAGTGTCTGATAGCAGCTTCTGAACTGGTTACCTGCCGTGAGTAAATTAAAATTTTATTGACTTAGGTCACTAAATACTTTAACCAATATAGGCATAGCGCACAGACAGATAAAAATTACAGAGTACACAACATCCATGAAACG"
        "CATTAGCACCACCATTACCACCACCATCACCATTACCACAGGTAACGGTGCGGGCTGACGCGTACAGGAAACACAGAAAAAAGCCCGCACCTGACAGTGCGGGCTTTTTTTTTCGACCAAAGGTAACGAGGTAACAACCATGCGAGTGTTGAAGTTCGGCGGTACATCAGTGGCAAATGCAGAACGTTTTCTGCGTGTTGCC"
        "GATATTCTGGAAAGCAATGCCAGGCAGGGGCAGGTGGCCACCGTCCTCTCTGCCCCCGCCAAAATCACCAACCACCTGGTGGCGATGATTGAAAAAACCAT")
# test_variants of ReferenceTest.cpp:286-289 (+ 317): {position, var_seq, allele bits}
VARS = [(2, "-", 2), (4, "TAG", 3), (9, "C", 1)]
LAST = (499, "TAG", 3)
# (start, len, reversed, first variant, posCurrentlyAt, allele) -> expected, ReferenceTest.cpp:290-334
KAT = [((0, 12, 0, 0, 0, 0), "AGCTTAGTTCAC"), ((0, 11, 0, 0, 0, 1), "AGTTAGTTCAT"), ((4, 8, 0, 1, 0, 1), "TAGTTCAT"), ((4, 7, 0, 1, 1, 1), "AGTTCAT"),
       ((4, 6, 0, 1, 2, 1), "GTTCAT"), ((10, 12, 1, 2, 0, 0), "GTGAACTAAGCT"), ((10, 11, 1, 2, 0, 1), "ATGAACTAACT"), ((5, 7, 1, 1, 0, 0), "CTAAGCT"),
       ((5, 6, 1, 1, 2, 0), "TAAGCT"), ((5, 5, 1, 1, 1, 0), "AAGCT")]
KAT_WITH_LAST = [((500, 12, 1, 3, 0, 0), "CTATGGTTTTTT"), ((4, 1, 0, 1, 0, 1), "T"), ((4, 1, 0, 1, 1, 1), "A"), ((4, 1, 0, 1, 2, 1), "G"),
                 ((5, 1, 1, 1, 0, 0), "C"), ((5, 1, 1, 1, 2, 0), "T"), ((5, 1, 1, 1, 1, 0), "A")]


@pytest.fixture(scope="module")
def splice(workdir):
    exe = os.path.join(workdir, "variant_core_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "host_twin", "variant_core_check.cpp"), "-lz"], check=True)

    def run(commands):
        return subprocess.run([exe], input="\n".join(commands) + "\n", capture_output=True, text=True, check=True).stdout.split("\n")[:-1]
    return run


def test_reference_known_answers(splice):
    assert len(SEQ0) == 500
    cmds = ["seq " + SEQ0] + [f"var {p} {b} {bits:x} 0" for p, b, bits in VARS] + ["call 0 " + " ".join(map(str, c)) for c, _ in KAT]
    cmds += ["var {} {} {:x} 0".format(*LAST)] + ["call 0 " + " ".join(map(str, c)) for c, _ in KAT_WITH_LAST]
    assert splice(cmds) == [want for _, want in KAT + KAT_WITH_LAST]


def test_without_variants_equals_the_plain_overload(splice):
    """Reference.cpp:483-496: infix / reverse complement of the infix."""
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    got = splice(["seq " + SEQ0, "call 0 0 10 0 0 0 0", "call 0 500 10 1 -1 0 0", "call 0 137 150 0 0 0 0", "call 0 300 150 1 -1 0 0"])
    assert got[0] == "AGCTTTTCAT" and got[1] == "ATGGTTTTTT"   # ReferenceTest.cpp:279-284
    assert got[2] == SEQ0[137:287]
    assert got[3] == "".join(comp[b] for b in reversed(SEQ0[150:300]))


@pytest.mark.parametrize("tag", ["var", "var70"])
def test_seeded_calls_match_the_reference(splice, tag):
    lines = lzma.open(os.path.join(GOLDEN, f"simref_small_{tag}.varseq.txt.xz")).read().decode().strip().split("\n")
    assert len(lines) == 800
    calls = [" ".join(line.split(" ")[:8]) for line in lines]
    assert sum(c.split(" ")[6] != "0" for c in calls) > 15 and {c.split(" ")[4] for c in calls} == {"0", "1"}   # starts inside insertions, both directions
    got = splice([f"load {os.path.join(GOLDEN, 'simref_small.fa')} {os.path.join(GOLDEN, f'simref_small_{tag}.vcf')}"] + calls)
    assert got == [line.split(" ")[8] for line in lines]


def test_choose_alleles_known_answers(splice):
    """SimulatorTest.cpp:89-114 TestSelectAllele: random value 0.5 throughout gives [1, 0] for 2 and [2, 1, 3, 0] for 4 ids; the direct
    branch of ChooseAlleles draws at most half of them."""
    assert splice(["select 2", "select 4"]) == ["1", "2 1"]


def test_choose_alleles_matches_the_reference(splice):
    want = lzma.open(os.path.join(GOLDEN, "choose_alleles_seed11.txt.xz")).read().decode().strip().split("\n")
    assert len(want) == 1500 and max(int(line.split(" ")[0]) for line in want) == 256
    assert any(int(line.split(" ")[1]) > int(line.split(" ")[0]) // 2 for line in want)   # complement branch covered
    assert splice(["alleles 11 1500"]) == want


def test_sys_error_walks_match_the_reference(splice, workdir):
    """Simulator::GetSysErrorFromBlock (Simulator.cpp:240-292) over a chain of blocks with deletions, substitutions and insertions of three
    alleles, same-position and neighbouring variants, short blocks: every step of 300 walks as the unmodified reference returned it."""
    text = lzma.open(os.path.join(GOLDEN, "sys_error_variants_seed5.txt.xz")).read().decode()
    path = os.path.join(workdir, "sys_error_variants.txt")
    open(path, "w").write(text)
    want = [line for line in text.split("\n") if line.startswith("walk ")]
    assert len(want) == 300 and sum(line.startswith("v ") for line in text.split("\n")) > 200
    assert splice([f"sysfile {path}"]) == want


@pytest.mark.parametrize("seq_id,name,lines", [(0, "seq0", 53611), (1, "seq1", 36594), (0, "seq0_dense", 22564), (0, "var70", 52376)])
def test_allele_fragment_reproduces_the_reference_traces(splice, workdir, seq_id, name, lines):
    """allele_fragment (variant_core.cuh) on VariantSet::materialise: GC percent, start / end surrounding and end position of every traced
    (start, inserted base, length, allele) evaluation of the unmodified reference's VariantBiasVarModifiers (oracle/dump_tables biasmod)."""
    path = os.path.join(workdir, f"bias_mod_trace_{name}.txt")
    with lzma.open(os.path.join(GOLDEN, f"bias_mod_trace_{name}.txt.xz")) as f, open(path, "wb") as o:
        o.write(f.read())
    vcf = os.path.join(GOLDEN, "simref_small_var70.vcf" if name == "var70" else "simref_small_var.vcf")   # var70: allele bits beyond the first word
    got = splice([f"trace {os.path.join(GOLDEN, 'simref_small.fa')} {vcf} {seq_id} {path}"])
    assert got == [f"{lines} 0"]


def test_fragment_counts_per_allele_match_the_reference(splice, workdir):
    """allele_fragment_counts: GetDispersion / alleles, mean / alleles, NegativeBinomial by CDF inversion (FragmentDistributionStats.cpp:900-907,
    3602-3626) for 1..128 alleles, Poisson-limit dispersion parameters included; bit patterns in, counts out, as the unmodified reference computed them."""
    path = os.path.join(workdir, "fragment_counts_alleles.txt")
    with lzma.open(os.path.join(GOLDEN, "fragment_counts_alleles_seed3.txt.xz")) as f, open(path, "wb") as o:
        o.write(f.read())
    lines = open(path).read().strip().split("\n")
    assert len(lines) == 2500 and sum(line.split(" ")[5] != "0" for line in lines) > 300 and max(int(line.split(" ")[3]) for line in lines) == 128
    assert splice([f"negbin {path}"]) == ["2500 0"]
