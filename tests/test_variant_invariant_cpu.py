"""The design check for variant-aware simulation: what Simulator's incremental VariantBiasVarModifiers bookkeeping
(Simulator.cpp:1399-1896: gc_mod_, end_pos_shift_, surrounding_start_/end_, unhandled variants) yields per (start position, inserted start
base, fragment length, allele) equals plain lookups in the *materialised allele sequence* - the reference with the allele's variants applied -
through one coordinate map. The reference's own test states this for a handful of positions (SimulatorTest.cpp:116-195); here it is checked
on every line of two traces written by the unmodified reference (`oracle/_ref/dump_tables biasmod`, tests/golden/bias_mod_trace_seq?.txt.xz:
2500 start positions, ~90 000 (length, allele) evaluations, 5 alleles with deletions, substitutions, insertions, multi-allelic sites).

    off_a[p]   index in allele a's sequence of the first base standing for reference position p
    start      off_a[start] + start_variant_pos                       (start inside an insertion: its 2nd, 3rd, ... base)
    GC %       Percent(GC of allele_seq[start : start + length], length)            = GetGCPercent
    surroundings  forward surrounding at start, reverse surrounding at start + length - 1   = bias_mod.surrounding_start_/end_
    end        smallest p with off_a[p] >= start + length                = cur_start_position + length + end_pos_shift_
    possible alleles  all but those deleting the start base; inside an insertion only its carriers   = GetPossibleAlleles

So the kernels need per allele a GC prefix, the two per-base surrounding-bias arrays and the map - the arrays they already use for the
reference - instead of a port of the incremental state machine."""
import bisect
import lzma
import os

import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def load_variants(seq, tag="var"):
    out = []
    for line in open(os.path.join(GOLDEN, f"simref_small_{tag}.variants.txt")).read().strip().split("\n")[1:]:
        s, pos, bases, lo, hi = line.split(" ")
        if int(s) == seq:
            out.append((int(pos), "" if bases == "-" else bases, int(lo, 16) | (int(hi, 16) << 64)))
    return out


def allele_sequence(ref, variants, allele):
    by_pos = {}
    for pos, bases, bits in variants:
        if (bits >> allele) & 1:
            assert pos not in by_pos   # one variant per position and allele (overlapping records are rejected on load)
            by_pos[pos] = bases
    parts, off, n = [], [], 0
    for p, base in enumerate(ref):
        off.append(n)
        rep = by_pos.get(p, base)
        parts.append(rep)
        n += len(rep)
    off.append(n)
    return "".join(parts), off


def forward_surrounding(seq, pos):   # SurroundingBase::Set/Forward (SurroundingBase.hpp:64-81, 196-201): 3 x 10-mer codes of pos-10 .. pos+19
    length, start = len(seq), pos + len(seq) - 10
    return [sum(CODE[seq[(start + block * 10 + k) % length]] << (2 * (9 - k)) for k in range(10)) for block in range(3)]


def reverse_surrounding(seq, pos):   # the same on the reverse complement, anchored at the fragment's last base
    length = len(seq)
    start = (length - pos - 1) + length - 10
    return [sum((3 - CODE[seq[length - 1 - (start + block * 10 + k) % length]]) << (2 * (9 - k)) for k in range(10)) for block in range(3)]


@pytest.mark.parametrize("seq_id,name", [(0, "seq0"), (1, "seq1"), (0, "seq0_dense"), (0, "var70")])
def test_bias_modifiers_equal_lookups_in_the_allele_sequence(seq_id, name):
    n_alleles, tag = (70, "var70") if name == "var70" else (5, "var")
    lines = lzma.open(os.path.join(GOLDEN, f"bias_mod_trace_{name}.txt.xz")).read().decode().strip().split("\n")
    ref = lines[0].split(" ")[1]
    variants = load_variants(seq_id, tag)
    alleles = {a: allele_sequence(ref, variants, a) for a in range(n_alleles)}
    checked = {"gc": 0, "surroundings": 0, "end": 0, "alleles": 0, "inside_insertion": 0}
    start = svp = None
    for line in lines[1:]:
        t = line.split(" ")
        if t[0] == "p":
            start, first_var, svp = int(t[1]), int(t[2]), int(t[3])
            at_start = first_var < len(variants) and variants[first_var][0] == start
            want = []
            for a in range(n_alleles):
                carries = at_start and (variants[first_var][2] >> a) & 1
                skipped = at_start and ((variants[first_var][1] == "" and carries) or (variants[first_var][1] != "" and svp and not carries))   # AlleleSkipped, Simulator.h:401-413
                if not skipped:
                    want.append(a)
            assert [int(x) for x in t[4:]] == want, line
            checked["alleles"] += 1
            checked["inside_insertion"] += svp > 0
            continue
        length, allele, shift = int(t[1]), int(t[2]), int(t[3])
        aseq, off = alleles[allele]
        mod_start = off[start] + svp
        end = start + length + shift
        if end < len(ref):
            assert bisect.bisect_left(off, mod_start + length) == end, line
            checked["end"] += 1
        if mod_start < 40 or mod_start + length + 40 > len(aseq):
            continue   # circular surroundings at the sequence ends are not part of this check
        if t[4] != "-":
            gc = sum(c in "GC" for c in aseq[mod_start:mod_start + length])
            assert ((gc * 100 + length // 2) // length) & 0xff == int(t[4]), line   # utilities::Percent (utilities.hpp:450-452, 552-554)
            checked["gc"] += 1
        assert forward_surrounding(aseq, mod_start) == [int(x) for x in t[5:8]], line
        assert reverse_surrounding(aseq, mod_start + length - 1) == [int(x) for x in t[8:11]], line
        checked["surroundings"] += 1
    if name == "var70":
        assert checked["gc"] > 50000 and checked["inside_insertion"] == 6   # two different insertions at one position, three inserted bases each
    elif name == "seq0_dense":
        assert checked["gc"] > 20000 and checked["inside_insertion"] == 4
    else:
        assert checked["gc"] > 30000 and checked["end"] > 30000 and checked["alleles"] >= 1000
    if name == "seq0":
        assert checked["inside_insertion"] > 2


@pytest.mark.parametrize("tag", ["var", "var70"])
def test_spliced_sequences_are_slices_of_the_allele_sequence(tag):
    """Reference::ReferenceSequence, variant overload (what GetOrgSeq calls): forward = allele_seq[off[start] + posCurrentlyAt : + length];
    reverse = reverse complement of the `length` bases in front of off[start] (or of off[insertion] + posCurrentlyAt when the fragment ends
    inside an insertion). 2 x 800 calls answered by the unmodified reference (tests/golden/simref_small_var{,70}.varseq.txt.xz)."""
    seqs = []
    for line in open(os.path.join(GOLDEN, "simref_small.fa")):
        if line.startswith(">"):
            seqs.append([])
        else:
            seqs[-1].append(line.strip().upper())
    seqs = ["".join(s) for s in seqs]
    per_seq = {}
    for line in open(os.path.join(GOLDEN, f"simref_small_{tag}.variants.txt")).read().strip().split("\n")[1:]:
        s, pos, bases, lo, hi = line.split(" ")
        per_seq.setdefault(int(s), []).append((int(pos), "" if bases == "-" else bases, int(lo, 16) | (int(hi, 16) << 64)))
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    cache = {}
    calls = lzma.open(os.path.join(GOLDEN, f"simref_small_{tag}.varseq.txt.xz")).read().decode().strip().split("\n")
    for line in calls:
        t = line.split(" ")
        s, start, length, rev, first, first_pos, allele = map(int, t[1:8])
        if (s, allele) not in cache:
            cache[(s, allele)] = allele_sequence(seqs[s], per_seq.get(s, []), allele)
        aseq, off = cache[(s, allele)]
        if rev:
            x = off[per_seq[s][first][0]] + first_pos if first_pos else off[start]
            want = "".join(comp[c] for c in reversed(aseq[x - length:x]))
        else:
            x = off[start] + first_pos
            want = aseq[x:x + length]
        assert want == t[8], line[:80]
    assert len(calls) == 800


@pytest.mark.parametrize("seq_id,first_pos", [(0, 0), (1, 9000)])
def test_start_enumeration_with_inserted_bases(seq_id, first_pos):
    """The scan's outer loop with variants (Simulator.cpp:2287-2353 do-while + CheckForInsertedBasesToStartFrom, 1875-1896): every reference
    position is a start once (start_variant_pos 0, which covers the first base of an insertion there), then once more per further inserted
    base of every insertion at that position, longest-sorted order of Reference::InsertVariant; first_variant_id_ always names the first
    variant at or behind the start that is not yet consumed. Checked against the (start, first variant, inserted base) sequence of the trace."""
    lines = lzma.open(os.path.join(GOLDEN, f"bias_mod_trace_seq{seq_id}.txt.xz")).read().decode().strip().split("\n")
    got = [tuple(int(x) for x in line.split(" ")[1:4]) for line in lines[1:] if line.startswith("p ")]
    variants = load_variants(seq_id)
    first = 0
    while first < len(variants) and variants[first][0] < first_pos:
        first += 1
    want = []
    for start in range(first_pos, got[-1][0] + 1):
        svp = 0
        while True:
            want.append((start, first, svp))
            if first < len(variants) and variants[first][0] == start:   # CheckForInsertedBasesToStartFrom
                if svp:
                    svp += 1
                    if svp >= len(variants[first][1]):
                        svp = 0
                        first += 1
                else:
                    while first < len(variants) and variants[first][0] == start and len(variants[first][1]) < 2:
                        first += 1
                if svp == 0 and first < len(variants) and variants[first][0] == start:
                    svp = 1
            if not svp:
                break
    assert got == want
    assert seq_id != 0 or len(got) > got[-1][0] - first_pos + 1   # insertions add starts


def test_product_materialisation_matches_the_checked_model(workdir):
    """VariantSet::materialise (reseq_b200/csrc/variants.hpp) builds the allele sequence + coordinate map the relations above were checked on."""
    import subprocess
    root = os.path.dirname(GOLDEN)
    exe = os.path.join(workdir, "variants_check_m")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(root, "host_twin", "variants_check.cpp"), "-lz"], check=True)
    seqs = []
    for line in open(os.path.join(GOLDEN, "simref_small.fa")):
        if line.startswith(">"):
            seqs.append([])
        else:
            seqs[-1].append(line.strip().upper())
    seqs = ["".join(s) for s in seqs]
    for seq_id, allele in ((0, 0), (0, 4), (1, 2), (3, 1)):
        out = subprocess.run([exe, os.path.join(GOLDEN, "simref_small.fa"), os.path.join(GOLDEN, "simref_small_var.vcf"), "allele", str(seq_id), str(allele)],
                             check=True, capture_output=True, text=True).stdout.strip().split("\n")
        aseq, off = allele_sequence(seqs[seq_id], load_variants(seq_id), allele)
        assert out[0] == aseq
        assert [tuple(map(int, line.split(" "))) for line in out[1:] if not line.startswith("i ")] == [(p, off[p]) for p in range(1, len(off)) if off[p] != off[p - 1] + 1]
        for line in out[1:]:
            if line.startswith("i "):
                _, idx, pos = line.split(" ")
                assert bisect.bisect_left(off, int(idx)) == int(pos)


@pytest.mark.parametrize("seq_id,name,inside", [(0, "seq0", 0), (1, "seq1", 0), (0, "seq0_dense", 272)])
def test_end_variant_in_allele_coordinates(seq_id, name, inside):
    """VariantBiasVarModifiers::EndVariant (Simulator.h:70-87), the cursor CreateReads hands to the reverse read (Simulator.cpp:687-688):
    a fragment ending strictly inside an insertion of its allele names that insertion and the number of its bases inside the fragment;
    otherwise the last variant of the sequence in front of the end position, whatever its alleles. (A fragment lying completely inside the
    insertion it starts in - shorter than the insertion - has its own branch in the reference and is not covered by the traces.)"""
    lines = lzma.open(os.path.join(GOLDEN, f"bias_mod_trace_{name}.txt.xz")).read().decode().strip().split("\n")
    ref = lines[0].split(" ")[1]
    variants = load_variants(seq_id)
    positions = [v[0] for v in variants]
    alleles = {a: allele_sequence(ref, variants, a) for a in range(5)}
    seen_inside = 0
    start = first_var = svp = None
    for line in lines[1:]:
        t = line.split(" ")
        if t[0] == "p":
            start, first_var, svp = int(t[1]), int(t[2]), int(t[3])
            continue
        assert (int(t[11]), int(t[12])) == (first_var, svp)   # StartVariant
        if len(t) < 15:
            continue
        length, allele, shift = int(t[1]), int(t[2]), int(t[3])
        _, off = alleles[allele]
        mod_end = off[start] + svp + length
        want = None
        q = bisect.bisect_left(off, mod_end) - 1   # last reference position whose bases start in front of the fragment's end
        for vid in range(bisect.bisect_left(positions, q), bisect.bisect_right(positions, q)):
            _, bases, bits = variants[vid]
            if (bits >> allele) & 1 and off[q] < mod_end < off[q] + len(bases):
                want = (vid, mod_end - off[q])
                seen_inside += 1
        if want is None:
            want = (bisect.bisect_left(positions, start + length + shift) - 1, 0)
        assert want == (int(t[13]), int(t[14])), line
    assert seen_inside == inside
