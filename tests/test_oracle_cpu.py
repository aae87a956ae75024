"""CPU-side parity: the committed fixtures really come from the reference, the glibc exp/pow port is bit exact, and
the lane-group templates (instantiated with one lane) reproduce the reference's stages and FASTQ byte for byte."""
import filecmp
import os
import subprocess
import sys

import pytest

from conftest import run_oracle_sim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TWIN_DIR = os.path.join(ROOT, "tests", "host_twin")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _compile(src, out, extra=()):
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", *extra, "-o", out, os.path.join(TWIN_DIR, src), "-lz"], check=True)
    return out


def test_exp_pow_port_matches_libm(workdir):
    """mathx.cuh vs the system libm the reference links (utilities.hpp:506, FragmentDistributionStats.cpp:3586,3604)."""
    exe = _compile("mathx_check.cpp", os.path.join(workdir, "mathx_check"), ["-mfma"])
    out = subprocess.run([exe, "3000000", "11"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "exp_mismatch=0 logit_mismatch=0 pow_mismatch=0" in out.stdout


def test_mt19937_64_jump_polynomials_match_discard(workdir):
    """csrc/mt_jump_tables.inc (tools/gen_mt_jump.py): window 2^k outputs ahead == std::mt19937_64::discard(2^k)."""
    exe = _compile("jump_check.cpp", os.path.join(workdir, "jump_check"))
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "jump_mismatches=0" in out.stdout, out.stdout


def test_chunked_ordered_sum_equals_the_chain(workdir):
    """csrc/ordered_sum.cuh: the four passes that evaluate Reference::SumBias' strictly ordered FP64 sum in parallel chunks give the bits of the
    plain chain - forced round-to-even ties, binade boundaries, zeros, 24 orders of magnitude, denormal starts, chunk lengths 1 .. 1024."""
    exe = _compile("ordered_sum_check.cpp", os.path.join(workdir, "ordered_sum_check"))
    for seed in ("5", "77", "2026"):
        out = subprocess.run([exe, seed], capture_output=True, text=True)
        assert out.returncode == 0 and "mismatches=0" in out.stdout, out.stdout


def test_golden_fastq_is_what_the_reference_writes(oracle, golden, workdir):
    r1, r2 = run_oracle_sim(oracle, golden["reseq"], golden["small_ref"], 42, 20, os.path.join(workdir, "pin"))
    assert filecmp.cmp(r1, golden["r1"], shallow=False)
    assert filecmp.cmp(r2, golden["r2"], shallow=False)


def test_golden_error_model_is_what_the_reference_writes(oracle, golden, workdir):
    out = os.path.join(workdir, "pin_em.fq")
    subprocess.run([oracle["reseq"], "seqToIllumina", "-j", "1", "--verbosity", "1", "-i", golden["em_in"], "-s", golden["reseq"],
                    "--ipfIterations", "0", "--seed", "7", "-o", out], check=True, timeout=600, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    assert filecmp.cmp(out, golden["em_out"], shallow=False)


def test_multibatch_error_model_hash_is_what_the_reference_writes(oracle, golden, workdir):
    """tests/golden/em_multibatch_sha256.json: 27 000 records = three batches through the reference's own SimulateErrorModelOnly
    (`dump_tables errmodel`: the counter its writer waits for is preset, see tests/golden/make_em_multibatch.py)."""
    import hashlib
    import json
    sys.path.insert(0, GOLDEN)
    import make_em_multibatch as mk
    want = json.load(open(os.path.join(GOLDEN, "em_multibatch_sha256.json")))
    fa, out = os.path.join(workdir, "em3.fa"), os.path.join(workdir, "em3_oracle.fq")
    assert mk.write_input(fa) == want["records"]
    assert hashlib.sha256(open(fa, "rb").read()).hexdigest() == want["input_sha256"]
    mk.run_oracle(oracle["dump"], golden["reseq"], fa, out)
    data = open(out, "rb").read()
    assert len(data) == want["bytes"] and hashlib.sha256(data).hexdigest() == want["sha256"]
    # the first batch is the single-batch golden run on the same records (ids aside): same seed, same stream
    first = data.split(b"\n")[:4 * 9000]
    gold = open(golden["em_out"], "rb").read().split(b"\n")[:4 * 9000]
    assert [x for i, x in enumerate(first) if i % 4] == [x for i, x in enumerate(gold) if i % 4]


@pytest.fixture(scope="module")
def twin(workdir):
    return _compile("twin.cpp", os.path.join(workdir, "twin"))


@pytest.mark.parametrize("seed,coverage", [(42, 20), (7, 6)])
def test_templates_match_reference_stages_and_fastq(oracle, golden, twin, workdir, seed, coverage):
    """Normalisation, thresholds, block seeds, adapter/forward/reverse systematic errors and the final FASTQ."""
    stage = os.path.join(workdir, f"stage_{seed}.flat")
    subprocess.run([oracle["dump"], "sim", golden["reseq"], golden["small_ref"], str(seed), str(coverage), stage], check=True, timeout=600,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    if (seed, coverage) == (42, 20):
        r1, r2 = golden["r1"], golden["r2"]
    else:
        r1, r2 = run_oracle_sim(oracle, golden["reseq"], golden["small_ref"], seed, coverage, os.path.join(workdir, f"o{seed}"))
    from reseq_b200.flatfile import read_flat
    st = read_flat(stage)
    n_blocks = len(st["sim.block_seed"])
    lookahead = 1 + (int(st["insert_lengths.from"][0]) + len(st["insert_lengths"])) // 1000
    prefix = os.path.join(workdir, f"twin_{seed}")
    res = subprocess.run([twin, stage, str(seed), prefix, str(n_blocks - lookahead)], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout
    assert "stage_mismatches=0" in res.stdout and "error_flag=0" in res.stdout
    assert filecmp.cmp(prefix + "_1.fq", r1, shallow=False)
    assert filecmp.cmp(prefix + "_2.fq", r2, shallow=False)


@pytest.mark.parametrize("depth", [1, 5, 32, 64])
def test_speculative_two_phase_templates_match_reference(oracle, golden, twin, workdir, depth):
    """spec_core.cuh (scan_window + ReadMachine, verification and replay) with one-lane groups: same bytes as the reference."""
    stage = os.path.join(workdir, "stage_spec.flat")
    if not os.path.exists(stage):
        subprocess.run([oracle["dump"], "sim", golden["reseq"], golden["small_ref"], "42", "20", stage], check=True, timeout=600,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    prefix = os.path.join(workdir, f"twin_spec{depth}")
    res = subprocess.run([twin, stage, "42", prefix, "66"], capture_output=True, text=True, timeout=900, env=dict(os.environ, RSQ_TWIN_SPEC=str(depth)))
    assert res.returncode == 0, res.stdout
    assert "error_flag=0" in res.stdout and "spec: depth=%d" % depth in res.stdout
    assert filecmp.cmp(prefix + "_1.fq", golden["r1"], shallow=False)
    assert filecmp.cmp(prefix + "_2.fq", golden["r2"], shallow=False)


def test_speculation_depth_scaled_by_read_density_matches_reference(oracle, golden, twin, workdir):
    """SpecCtx::mean_reads: a unit whose reads come denser than the expected number per SimBlock speculates proportionally deeper than run_depth
    (up to the capacity) - the rule that evens out the number of rounds per unit on small genomes.  Same bytes whatever the rule does."""
    stage = os.path.join(workdir, "stage_spec.flat")
    if not os.path.exists(stage):
        subprocess.run([oracle["dump"], "sim", golden["reseq"], golden["small_ref"], "42", "20", stage], check=True, timeout=600,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for mean in ("40", "400"):
        prefix = os.path.join(workdir, f"twin_hot{mean}")
        res = subprocess.run([twin, stage, "42", prefix, "66"], capture_output=True, text=True, timeout=900,
                             env=dict(os.environ, RSQ_TWIN_SPEC="48", RSQ_TWIN_MEAN_READS=mean))
        assert res.returncode == 0 and "error_flag=0" in res.stdout, res.stdout
        assert filecmp.cmp(prefix + "_1.fq", golden["r1"], shallow=False)
        assert filecmp.cmp(prefix + "_2.fq", golden["r2"], shallow=False)


@pytest.mark.parametrize("spec_depth", [None, "4"])
def test_templates_match_reference_with_tiles_and_read_lengths(oracle, golden, twin, workdir, spec_depth):
    """Three tiles and two read lengths per segment (profile150t): tile draw per pair, read-length draw per read."""
    stage = os.path.join(workdir, "stage_t.flat")
    if not os.path.exists(stage):
        subprocess.run([oracle["dump"], "sim", golden["reseq_t"], golden["small_ref"], "11", "25", stage], check=True, timeout=600,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    r1, r2 = run_oracle_sim(oracle, golden["reseq_t"], golden["small_ref"], 11, 25, os.path.join(workdir, "ora_t"))
    prefix = os.path.join(workdir, "twin_t" + (spec_depth or ""))
    env = dict(os.environ, RSQ_TWIN_SPEC=spec_depth) if spec_depth else dict(os.environ)
    res = subprocess.run([twin, stage, "11", prefix, "66"], capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0 and "error_flag=0" in res.stdout, res.stdout
    assert filecmp.cmp(prefix + "_1.fq", r1, shallow=False)
    assert filecmp.cmp(prefix + "_2.fq", r2, shallow=False)


@pytest.mark.parametrize("spec_depth,meth", [(None, False), ("6", False), (None, True), ("32", True)])
def test_templates_match_reference_with_250_base_reads(oracle, golden, twin, workdir, spec_depth, meth):
    """profile250: 2x250 reads (BASELINE config C4's read length; longer stream slices, CIGARs and record slots), with and without
    bisulfite conversion, serial and speculative form."""
    stage = os.path.join(workdir, "stage_250.flat")
    if not os.path.exists(stage):
        subprocess.run([oracle["dump"], "sim", golden["reseq_250"], golden["small_ref"], "9", "15", stage], check=True, timeout=600,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    extra = ("--methylation", golden["meth_bed"]) if meth else ()
    r1, r2 = run_oracle_sim(oracle, golden["reseq_250"], golden["small_ref"], 9, 15, os.path.join(workdir, f"ora_250_{int(meth)}"), extra=extra)
    assert {len(s) for s in open(r1, "rb").read().split(b"\n")[1::4]} == {250}
    prefix = os.path.join(workdir, f"twin_250_{spec_depth}_{int(meth)}")
    env = dict(os.environ, RSQ_TWIN_SPEC=spec_depth) if spec_depth else dict(os.environ)
    res = subprocess.run([twin, stage, "9", prefix, "66"] + ([golden["meth_bed"]] if meth else []), capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0 and "error_flag=0" in res.stdout, res.stdout
    assert filecmp.cmp(prefix + "_1.fq", r1, shallow=False)
    assert filecmp.cmp(prefix + "_2.fq", r2, shallow=False)


def test_speculative_adapter_only_pairs_match_serial_templates(oracle, golden, twin, workdir):
    """Adapter-only pairs (SimulateAdapterOnlyPairs; none in the golden profile) forced on: serial and speculative forms agree."""
    stage = os.path.join(workdir, "stage_spec.flat")
    if not os.path.exists(stage):
        subprocess.run([oracle["dump"], "sim", golden["reseq"], golden["small_ref"], "42", "20", stage], check=True, timeout=600,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    outs = []
    for name, extra in (("ao_serial", {}), ("ao_spec", {"RSQ_TWIN_SPEC": "6"}), ("ao_spec64", {"RSQ_TWIN_SPEC": "64"})):   # 64: three output slabs in flight
        prefix = os.path.join(workdir, name)
        res = subprocess.run([twin, stage, "42", prefix, "100000"], capture_output=True, text=True, timeout=900,
                             env=dict(os.environ, RSQ_TWIN_ADAPTER_ONLY="150", **extra))
        assert res.returncode == 0 and "error_flag=0" in res.stdout, res.stdout
        outs.append(prefix)
    assert b"Adapter" in open(outs[0] + "_1.fq", "rb").read()
    for other in outs[1:]:
        assert filecmp.cmp(outs[0] + "_1.fq", other + "_1.fq", shallow=False)
        assert filecmp.cmp(outs[0] + "_2.fq", other + "_2.fq", shallow=False)


def test_methylation_golden_is_what_the_reference_writes(oracle, golden, workdir):
    r1, r2 = run_oracle_sim(oracle, golden["reseq"], golden["small_ref"], 42, 20, os.path.join(workdir, "pin_meth"),
                            extra=("--methylation", golden["meth_bed"]))
    assert filecmp.cmp(r1, golden["meth_r1"], shallow=False)
    assert filecmp.cmp(r2, golden["meth_r2"], shallow=False)


@pytest.mark.parametrize("spec_depth", [None, "5"])
def test_templates_match_reference_with_methylation(oracle, golden, twin, workdir, spec_depth):
    """CTConversion (bisulfite C->T per unmethylated region, Simulator.cpp:1925-2003) through the one-lane twin, serial and speculative form."""
    stage = os.path.join(workdir, "stage_meth.flat")
    subprocess.run([oracle["dump"], "sim", golden["reseq"], golden["small_ref"], "42", "20", stage], check=True, timeout=600,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    prefix = os.path.join(workdir, "twin_meth")
    env = dict(os.environ, RSQ_TWIN_SPEC=spec_depth) if spec_depth else dict(os.environ)
    res = subprocess.run([twin, stage, "42", prefix, "66", golden["meth_bed"]], capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stdout
    assert filecmp.cmp(prefix + "_1.fq", golden["meth_r1"], shallow=False)
    assert filecmp.cmp(prefix + "_2.fq", golden["meth_r2"], shallow=False)


def test_replace_n_matches_reference(oracle, golden, workdir):
    """Reference::ReplaceN (Reference.cpp:813-891): short N runs drawn from mt19937_64(seed), long runs filled with the neighbouring
    repeat; the host code (word-wise skip over N-free stretches) against the reference's in-memory sequences."""
    import numpy as np
    from reseq_b200.flatfile import read_flat
    exe = _compile("replace_n_check.cpp", os.path.join(workdir, "replace_n_check"))
    fa = os.path.join(workdir, "n_rich.fa")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_synthetic.py"), "reference", fa, "--sizes", "40000,9000,23000", "--seed", "77",
                    "--n-rate", "0.02", "--prefix", "n"], check=True)
    for ref, seed in ((golden["small_ref"], 42), (fa, 5)):
        stage = os.path.join(workdir, f"stage_replace_n_{seed}.flat")
        subprocess.run([oracle["dump"], "sim", golden["reseq"], ref, str(seed), "5", stage], check=True, timeout=600, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        st = read_flat(stage)
        want = np.concatenate([st[k] for k in sorted((k for k in st if k.startswith("sim.ref.")), key=lambda k: int(k.split(".")[-1]))])
        got = np.frombuffer(subprocess.run([exe, ref, str(seed)], capture_output=True, check=True).stdout, dtype=np.uint8)
        assert got.size == want.size and np.array_equal(got, want.astype(np.uint8))
        assert got.max() <= 3


@pytest.mark.parametrize("tag,spec_depth", [("var", None), ("var", "1"), ("var", "8"), ("var_base", None), ("var_base", "32")])
def test_templates_match_reference_with_variants(oracle, golden, twin, workdir, tag, spec_depth):
    """-V (SURVEY section 8 row a6 + the variant halves of a4/a8/a14/a15/a18): thresholds for the file's allele count, block seeds with the variants'
    draws in the master stream, SimBlock::err_variants_ of every block and strand (SetSystematicErrorVariantsForward/Reverse), and the FASTQ
    the unmodified reference wrote for the 5-allele / 2-allele golden VCF - serial and speculative form of the templates."""
    import lzma
    vcf = os.path.join(golden["dir"], f"simref_small_{tag}.vcf")
    stage = os.path.join(workdir, f"stage_{tag}.flat")
    if not os.path.exists(stage):
        subprocess.run([oracle["dump"], "sim", golden["reseq"], golden["small_ref"], "42", "20", stage, "1000000", vcf], check=True, timeout=600,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    prefix = os.path.join(workdir, f"twin_{tag}_{spec_depth}")
    env = dict(os.environ, RSQ_TWIN_SPEC=spec_depth) if spec_depth else dict(os.environ)
    res = subprocess.run([twin, stage, "42", prefix, "66", "-", vcf], capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stdout
    assert "stage_mismatches=0" in res.stdout and "error_flag=0" in res.stdout and "strands equal the reference's" in res.stdout
    for k in (1, 2):
        want = lzma.open(os.path.join(golden["dir"], f"sim_small_{tag}_seed42_R{k}.fq.xz")).read()
        assert open(f"{prefix}_{k}.fq", "rb").read() == want


@pytest.mark.parametrize("spec_depth", [None, "6"])
def test_templates_match_reference_with_70_alleles_and_methylation(oracle, golden, twin, workdir, spec_depth):
    """70 haploid populations (allele bits in the second 64-bit word, up to 140 (allele, strand) ids per hit) together with --methylation:
    ChooseAlleles beyond two ids and the variant overload of CTConversion, against a live run of the reference binary."""
    vcf = os.path.join(golden["dir"], "simref_small_var70.vcf")
    stage = os.path.join(workdir, "stage_var70.flat")
    if not os.path.exists(stage):
        subprocess.run([oracle["dump"], "sim", golden["reseq"], golden["small_ref"], "7", "15", stage, "1000000", vcf], check=True, timeout=600,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    r1, r2 = run_oracle_sim(oracle, golden["reseq"], golden["small_ref"], 7, 15, os.path.join(workdir, "ora_var70m"), extra=("-V", vcf, "--methylation", golden["meth_bed"]))
    prefix = os.path.join(workdir, f"twin_var70m_{spec_depth}")
    env = dict(os.environ, RSQ_TWIN_SPEC=spec_depth) if spec_depth else dict(os.environ)
    res = subprocess.run([twin, stage, "7", prefix, "66", golden["meth_bed"], vcf], capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stdout
    assert "stage_mismatches=0" in res.stdout and "error_flag=0" in res.stdout
    assert filecmp.cmp(prefix + "_1.fq", r1, shallow=False)
    assert filecmp.cmp(prefix + "_2.fq", r2, shallow=False)


@pytest.mark.parametrize("spec_depth", [None, "8"])
def test_templates_match_reference_with_variants_at_sequence_ends(oracle, golden, twin, workdir, spec_depth):
    """Variants on the first and last bases of every sequence (surroundings roll around into the reference's other end and ignore variants there,
    Simulator.cpp:1601, 1706) and insertions of 30-140 bases (fragments starting, ending or lying inside one), at 60x so that they are hit."""
    vcf = os.path.join(golden["dir"], "simref_small_var_ends.vcf")
    stage = os.path.join(workdir, "stage_var_ends.flat")
    if not os.path.exists(stage):
        subprocess.run([oracle["dump"], "sim", golden["reseq"], golden["small_ref"], "3", "60", stage, "1000000", vcf], check=True, timeout=600,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    r1, r2 = run_oracle_sim(oracle, golden["reseq"], golden["small_ref"], 3, 60, os.path.join(workdir, "ora_var_ends"), extra=("-V", vcf))
    prefix = os.path.join(workdir, f"twin_var_ends_{spec_depth}")
    env = dict(os.environ, RSQ_TWIN_SPEC=spec_depth) if spec_depth else dict(os.environ)
    res = subprocess.run([twin, stage, "3", prefix, "66", "-", vcf], capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stdout
    assert "stage_mismatches=0" in res.stdout and "error_flag=0" in res.stdout
    assert filecmp.cmp(prefix + "_1.fq", r1, shallow=False)
    assert filecmp.cmp(prefix + "_2.fq", r2, shallow=False)


def test_oracle_passes_the_references_own_unit_tests(oracle, workdir):
    """`reseq test` of the oracle binary (the reference's gtest suites compiled in from its unmodified sources, main.cpp:1138-1196): three tiers,
    13 + 17 + 24 = 54 tests.  Needs the reference's test/ fixtures, so it only runs where /root/reference exists (the build container)."""
    if not os.path.isdir("/root/reference/test"):
        pytest.skip("/root/reference/test (fixtures of the reference's own tests) is absent on this box")
    res = subprocess.run([oracle["reseq"], "test"], capture_output=True, text=True, timeout=900, cwd=workdir, env=dict(os.environ, RESEQ_FOLDER="/root/reference"))
    out = res.stdout + res.stderr
    import re
    passed = [int(n) for n in re.findall(r"\[  PASSED  \] (\d+) tests", out)]
    assert res.returncode == 0, out[-2000:]
    assert "FAILED" not in out
    assert passed == [13, 17, 24], passed


@pytest.mark.parametrize("spec_depth", [None, "8"])
def test_adapter_only_pairs_match_reference(oracle, golden, twin, workdir, spec_depth):
    """Simulator::SimulateAdapterOnlyPairs (Simulator.cpp:2359-2382) against the reference itself: profile150a has InsertLengths()[0] = 700
    (oracle/dump_tables patch_adapter_only), so 3.5 % of the pairs are adapter + poly-A tail + overrun bases only, written behind the last block."""
    stage = os.path.join(workdir, "stage_ao.flat")
    if not os.path.exists(stage):
        subprocess.run([oracle["dump"], "sim", golden["reseq_a"], golden["small_ref"], "42", "20", stage], check=True, timeout=600,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    r1, r2 = run_oracle_sim(oracle, golden["reseq_a"], golden["small_ref"], 42, 20, os.path.join(workdir, "ora_ao"))
    assert open(r1, "rb").read().count(b":Adapter:") > 100
    prefix = os.path.join(workdir, f"twin_ao_{spec_depth}")
    env = dict(os.environ, RSQ_TWIN_SPEC=spec_depth) if spec_depth else dict(os.environ)
    res = subprocess.run([twin, stage, "42", prefix, "66"], capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stdout
    assert "stage_mismatches=0" in res.stdout and "error_flag=0" in res.stdout
    assert filecmp.cmp(prefix + "_1.fq", r1, shallow=False)
    assert filecmp.cmp(prefix + "_2.fq", r2, shallow=False)
