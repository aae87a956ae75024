"""GPU parity tests for variant-aware simulation (`reseq illuminaPE -V <vcf>`, SURVEY section 8 row a6 and the variant halves of a4, a8, a9, a14, a15,
a18): through the C ABI against FASTQ the unmodified reference wrote (committed goldens) and against the reference binary run live."""
import lzma
import os

import pytest

from conftest import run_oracle_sim

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def rb(library):
    import reseq_b200
    if library.rsq_device_count() < 1:
        pytest.fail("no CUDA device: the engine has no CPU path")
    return reseq_b200


@pytest.fixture(scope="module")
def engine(rb, golden):
    eng = rb.Engine(rb.Profile.load_flat(golden["flat"]), 0)
    yield eng
    eng.close()


def _simulate(eng, ref, **kw):
    eng.prepare(ref, **kw)
    rep = eng.simulate()
    eng.download()
    return eng.output(0), eng.output(1), rep


def _golden_fastq(golden, tag):
    return [lzma.open(os.path.join(golden["dir"], f"sim_small_{tag}_seed42_R{k}.fq.xz")).read() for k in (1, 2)]


def _ref_with_vcf(rb, golden, tag):
    ref = rb.Reference.load_fasta(golden["small_ref"])
    ref.load_variants(os.path.join(golden["dir"], f"simref_small_{tag}.vcf"))
    return ref


@pytest.mark.parametrize("tag,alleles", [("var", 5), ("var_base", 2)])
@pytest.mark.parametrize("path,depth", [("serial", None), ("spec", 1), ("spec", 3), ("spec", 32), ("spec", None)])
def test_variant_goldens_bit_exact(rb, engine, golden, monkeypatch, tag, alleles, path, depth):
    """5 alleles in three populations (deletions, substitutions, insertions, multi-allelic sites) and a 2-allele file: the reference's FASTQ
    (tests/golden/make_variants_golden.py) byte for byte, on the one-warp-per-SimBlock kernel and on the speculative kernels at several depths."""
    monkeypatch.setenv("RSQ_SIM_PATH", path)
    if depth:
        monkeypatch.setenv("RSQ_SPEC_DEPTH", str(depth))
    ref = _ref_with_vcf(rb, golden, tag)
    assert ref.num_alleles == alleles
    r1, r2, rep = _simulate(engine, ref, seed=42, coverage=20.0)
    want = _golden_fastq(golden, tag)
    assert r1 == want[0] and r2 == want[1]
    assert b"_allele" in r1[:200]
    assert (rep.spec_rounds > 0) == (path == "spec")


@pytest.mark.parametrize("shards", [3, 7])
def test_variant_run_in_shards(rb, engine, golden, shards):
    ref = _ref_with_vcf(rb, golden, "var")
    want = _golden_fastq(golden, "var")
    parts = [b"", b""]
    for k in range(shards):
        r1, r2, _ = _simulate(engine, ref, seed=42, coverage=20.0, shard_index=k, shard_count=shards)
        parts[0] += r1
        parts[1] += r2
    assert parts == want


@pytest.mark.parametrize("path,batch", [("spec", 5), ("spec", 1), ("serial", 9)])
def test_variant_run_in_batches_with_surroundings_windows(rb, engine, golden, monkeypatch, path, batch):
    """What a large genome with a VCF does on one GPU: several batches of SimBlocks, each with its own window of per-position surrounding biases
    (RSQ_SUR_WINDOW forces it on the small reference): the variant-aware evaluation reads the same arrays through the window-biased pointers."""
    monkeypatch.setenv("RSQ_SIM_PATH", path)
    monkeypatch.setenv("RSQ_BATCH_UNITS", str(batch))
    monkeypatch.setenv("RSQ_SUR_WINDOW", "1")
    want = _golden_fastq(golden, "var")
    r1, r2, rep = _simulate(engine, _ref_with_vcf(rb, golden, "var"), seed=42, coverage=20.0)
    assert [r1, r2] == want
    assert rep.batches == -(-rep.blocks // batch)


def test_engine_switches_between_variant_and_plain_runs(rb, engine, golden):
    want = _golden_fastq(golden, "var")
    r1, r2, _ = _simulate(engine, _ref_with_vcf(rb, golden, "var"), seed=42, coverage=20.0)
    assert [r1, r2] == want
    p1, p2, _ = _simulate(engine, rb.Reference.load_fasta(golden["small_ref"]), seed=42, coverage=20.0)
    assert p1 == open(golden["r1"], "rb").read() and p2 == open(golden["r2"], "rb").read()
    r1, r2, _ = _simulate(engine, _ref_with_vcf(rb, golden, "var"), seed=42, coverage=20.0)
    assert [r1, r2] == want


@pytest.mark.parametrize("tag,seed,coverage,path", [("var70", 7, 15.0, "spec"), ("var70", 7, 15.0, "serial"), ("var", 1234, 28.0, "spec"), ("var_base", 5, 40.0, "spec"),
                                                    ("var_ends", 3, 60.0, "spec"), ("var_ends", 3, 60.0, "serial")])
def test_variants_against_reference_binary(rb, engine, golden, oracle, workdir, monkeypatch, tag, seed, coverage, path):
    """70 haploid populations (allele bits beyond the first 64-bit word), further seeds, and var_ends: variants on the first and last bases of every
    sequence plus insertions longer than a read (tests/golden/make_variants_golden.py::write_ends_vcf) - the reference binary's own run with -V."""
    monkeypatch.setenv("RSQ_SIM_PATH", path)
    vcf = os.path.join(golden["dir"], f"simref_small_{tag}.vcf")
    o1, o2 = run_oracle_sim(oracle, golden["reseq"], golden["small_ref"], seed, coverage, os.path.join(workdir, f"ora_{tag}_{seed}"), extra=("-V", vcf))
    r1, r2, rep = _simulate(engine, _ref_with_vcf(rb, golden, tag), seed=seed, coverage=coverage)
    assert r1 == open(o1, "rb").read() and r2 == open(o2, "rb").read()
    assert rep.pairs > 1000


@pytest.mark.parametrize("path", ["spec", "serial"])
def test_variants_with_methylation_against_reference_binary(rb, engine, golden, oracle, workdir, monkeypatch, path):
    """-V together with --methylation: the variant overload of CTConversion (Simulator.cpp:2004-2217) walks the allele's variants while it converts."""
    monkeypatch.setenv("RSQ_SIM_PATH", path)
    vcf = os.path.join(golden["dir"], "simref_small_var.vcf")
    o1, o2 = run_oracle_sim(oracle, golden["reseq"], golden["small_ref"], 42, 20.0, os.path.join(workdir, "ora_var_meth"), extra=("-V", vcf, "--methylation", golden["meth_bed"]))
    ref = _ref_with_vcf(rb, golden, "var")
    ref.load_methylation(golden["meth_bed"])
    r1, r2, _ = _simulate(engine, ref, seed=42, coverage=20.0)
    assert r1 == open(o1, "rb").read() and r2 == open(o2, "rb").read()


def test_per_allele_methylation_against_reference_binary(rb, engine, golden, oracle, workdir):
    """A methylation file with one column per allele (Reference::ReadMethylation, Reference.cpp:1231-1275; Unmethylation(seq, allele))."""
    import random
    rnd = random.Random(11)
    bed = os.path.join(workdir, "alleles.bed")
    with open(bed, "w") as f:
        for line in open(golden["meth_bed"]):
            t = line.rstrip("\n").split("\t")
            if line.startswith("track") or len(t) < 4:
                f.write(line)
            elif t[0] == "chr2":   # one sequence keeps a single column
                f.write(line)
            else:
                f.write("\t".join(t[:3] + [f"{rnd.random():.3f}" for _ in range(5)]) + "\n")
    vcf = os.path.join(golden["dir"], "simref_small_var.vcf")
    o1, o2 = run_oracle_sim(oracle, golden["reseq"], golden["small_ref"], 9, 20.0, os.path.join(workdir, "ora_var_meth5"), extra=("-V", vcf, "--methylation", bed))
    ref = _ref_with_vcf(rb, golden, "var")
    ref.load_methylation(bed)
    r1, r2, _ = _simulate(engine, ref, seed=9, coverage=20.0)
    assert r1 == open(o1, "rb").read() and r2 == open(o2, "rb").read()


def test_read_sys_error_with_variants_is_refused(rb, engine, golden, workdir):
    ref = _ref_with_vcf(rb, golden, "var")
    prof = os.path.join(workdir, "sys_for_var.fq")
    open(prof, "w").write("@x reverse\nA\n+\n!\n")
    with pytest.raises(rb.RsqError, match="readSysError together with a variant file"):
        engine.prepare(ref, seed=42, coverage=20.0, sys_error_file=prof)
