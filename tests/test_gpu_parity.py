"""GPU parity tests: everything goes through the C ABI (libreseq_b200.so); the checker is the committed golden
FASTQ made by the reference's own code and, when oracle/_ref travelled to this box, the reference binary itself."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import run_oracle_sim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

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


def test_small_golden_bit_exact(rb, engine, golden):
    ref = rb.Reference.load_fasta(golden["small_ref"])
    r1, r2, rep = _simulate(engine, ref, seed=42, coverage=20.0)
    assert r1 == open(golden["r1"], "rb").read()
    assert r2 == open(golden["r2"], "rb").read()
    assert rep.pairs == r1.count(b"\n") // 4 and rep.blocks == rep.blocks_total - 1
    assert rep.kernel_launches > 0


@pytest.mark.parametrize("path,depth", [("serial", None), ("spec", 1), ("spec", 3), ("spec", 32)])
def test_serial_and_speculative_paths_bit_exact(rb, engine, golden, monkeypatch, path, depth):
    """The one-warp-per-SimBlock kernel (k_simulate) and the speculative two-phase kernels (k_spec_scan/k_spec_reads)
    at several speculation depths write the same bytes; ~18% of this profile's reads draw an InDel, so the
    verification/replay machinery runs on every block."""
    monkeypatch.setenv("RSQ_SIM_PATH", path)
    if depth:
        monkeypatch.setenv("RSQ_SPEC_DEPTH", str(depth))
    ref = rb.Reference.load_fasta(golden["small_ref"])
    r1, r2, rep = _simulate(engine, ref, seed=42, coverage=20.0)
    assert r1 == open(golden["r1"], "rb").read()
    assert r2 == open(golden["r2"], "rb").read()
    if path == "serial":
        assert rep.spec_rounds == 0 and rep.spec_depth == 0
    else:
        assert rep.spec_depth == depth and rep.spec_rounds > 0


@pytest.mark.parametrize("path,batch", [("spec", 7), ("spec", 64), ("serial", 5)])
def test_block_batches_concatenate_to_the_whole_run(rb, engine, golden, monkeypatch, path, batch):
    """Runs whose state does not fit HBM are simulated in batches of SimBlocks whose text is copied out behind each other
    (RSQ_BATCH_UNITS forces small batches here): the bytes must not depend on the batch size."""
    monkeypatch.setenv("RSQ_SIM_PATH", path)
    monkeypatch.setenv("RSQ_BATCH_UNITS", str(batch))
    ref = rb.Reference.load_fasta(golden["small_ref"])
    r1, r2, rep = _simulate(engine, ref, seed=42, coverage=20.0)
    assert r1 == open(golden["r1"], "rb").read()
    assert r2 == open(golden["r2"], "rb").read()
    assert rep.pairs == r1.count(b"\n") // 4


@pytest.mark.parametrize("path", ["spec", "serial"])
def test_realistic_tables_profile_equals_the_reference(rb, golden, monkeypatch, path):
    """profile150q: every base quality 2..41, two tiles, 2.1 MB of probability tables (the bench's second profile).  Only its flat image is
    committed (the archives are 115 MB); the reference's FASTQ for the small reference is pinned by hash (tests/golden/make_golden.py)."""
    import json
    want = json.load(open(os.path.join(GOLDEN_DIR, "profile150q_small_sha256.json")))
    monkeypatch.setenv("RSQ_SIM_PATH", path)
    prof = rb.Profile.load_flat(golden["flat_q"])
    eng = rb.Engine(prof, 0)
    try:
        r1, r2, rep = _simulate(eng, rb.Reference.load_fasta(golden["small_ref"]), seed=42, coverage=20.0)
    finally:
        eng.close()
    assert rep.pairs == want["pairs"] and [len(r1), len(r2)] == want["bytes"]
    assert hashlib.sha256(r1).hexdigest() == want["r1"] and hashlib.sha256(r2).hexdigest() == want["r2"]


@pytest.mark.parametrize("chunk", [None, "32", "1024"])
def test_chunked_bias_sums_equal_the_chain_kernel(rb, golden, monkeypatch, chunk):
    """CalculateBiasNormalization: the chunked exact evaluation of Reference::SumBias (k_bias_chunks / _scan / _resolve, ordered_sum.cuh) gives
    the same normalisation and thresholds, bit for bit, as the kernel that adds a whole chain in order (RSQ_BIAS_PATH=chain) - on the small
    reference (chains of 800 .. 30 000 terms) and on a 6 Mbp sequence, for several chunk lengths."""
    import bench
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_synthetic
    prof = rb.Profile.load_flat(golden["flat_r"])
    refs = [rb.Reference.load_fasta(golden["small_ref"])]
    seqs = make_synthetic.gen_reference([6_000_000, 1_500_000], 99)
    refs.append(rb.Reference.from_memory(["a", "b"], [q.encode() for q in seqs]))
    eng = rb.Engine(prof, 0)
    try:
        for ref in refs:
            got = {}
            for path in ("chain", "chunks"):
                if path == "chain":
                    monkeypatch.setenv("RSQ_BIAS_PATH", "chain")
                else:
                    monkeypatch.delenv("RSQ_BIAS_PATH", raising=False)
                    if chunk:
                        monkeypatch.setenv("RSQ_BIAS_CHUNK", chunk)
                rep = eng.prepare(ref, seed=42, coverage=20.0)
                got[path] = (rep.bias_normalization, eng.fetch("thresholds").tobytes())
            assert got["chain"][0] == got["chunks"][0]
            assert got["chain"][1] == got["chunks"][1]
    finally:
        eng.close()


@pytest.mark.parametrize("path,batch", [("spec", 7), ("spec", 1), ("serial", 5)])
def test_batches_with_windows_of_surrounding_biases(rb, engine, golden, monkeypatch, path, batch):
    """Multi-batch runs keep the per-position surrounding biases (16 bytes per base) of one batch only, recomputed in front of it, with the kernels
    indexing through pointers biased by the window start (RSQ_SUR_WINDOW forces that mode on a small run): same bytes, 9 bytes per base plus one
    batch's window resident; the second simulate call on the same prepare stays in that mode."""
    monkeypatch.setenv("RSQ_SIM_PATH", path)
    monkeypatch.setenv("RSQ_BATCH_UNITS", str(batch))
    monkeypatch.setenv("RSQ_SUR_WINDOW", "1")
    ref = rb.Reference.load_fasta(golden["small_ref"])
    engine.prepare(ref, seed=42, coverage=20.0)
    for _ in range(2):
        rep = engine.simulate()
        engine.download()
        assert engine.output(0) == open(golden["r1"], "rb").read()
        assert engine.output(1) == open(golden["r2"], "rb").read()
        assert rep.batches == -(-rep.blocks // batch)
        assert 9.0 <= rep.resident_bytes_per_base < 9.0 + 16.0 * (batch * 1000 + 2000) / 70000
    with pytest.raises(rb.RsqError, match="no longer resident"):
        engine.fetch("sur_start")


@pytest.mark.parametrize("path,shards", [("spec", 1), ("serial", 1), ("spec", 3)])
def test_adapter_only_pairs_against_reference_binary(rb, golden, oracle, workdir, monkeypatch, path, shards):
    """Simulator::SimulateAdapterOnlyPairs (Simulator.cpp:2359-2382) against the reference: profile150a has InsertLengths()[0] > 0 (oracle/dump_tables
    patch_adapter_only), loaded here through the .reseq/.ipf archive reader; the adapter-only pairs follow the last block (the last shard)."""
    monkeypatch.setenv("RSQ_SIM_PATH", path)
    o1, o2 = run_oracle_sim(oracle, golden["reseq_a"], golden["small_ref"], 42, 20.0, os.path.join(workdir, "ora_ao_gpu"))
    want = [open(o1, "rb").read(), open(o2, "rb").read()]
    assert want[0].count(b":Adapter:") > 100
    eng = rb.Engine(rb.Profile.load(golden["reseq_a"], golden["ipf_a"]), 0)
    try:
        ref = rb.Reference.load_fasta(golden["small_ref"])
        got = [b"", b""]
        for k in range(shards):
            r1, r2, rep = _simulate(eng, ref, seed=42, coverage=20.0, shard_index=k, shard_count=shards)
            got[0] += r1
            got[1] += r2
        assert rep.adapter_only_pairs > 100
    finally:
        eng.close()
    assert got == want


def test_adapter_only_pairs_serial_and_speculative_agree(rb, engine, golden, monkeypatch):
    """SimulateAdapterOnlyPairs (insert length 0: reads made of adapter, poly-A tail and overrun bases).  The golden profiles
    contain no such pairs, so they are forced on and the two kernel forms are compared with each other (k_adapter_only vs the
    pseudo unit of the speculative kernels, which always runs at full depth)."""
    monkeypatch.setenv("RSQ_FORCE_ADAPTER_ONLY", "333")
    ref = rb.Reference.load_fasta(golden["small_ref"])
    outs = []
    for path, depth in (("serial", None), ("spec", "3"), ("spec", None)):
        monkeypatch.setenv("RSQ_SIM_PATH", path)
        if depth:
            monkeypatch.setenv("RSQ_SPEC_DEPTH", depth)
        else:
            monkeypatch.delenv("RSQ_SPEC_DEPTH", raising=False)
        r1, r2, rep = _simulate(engine, ref, seed=42, coverage=20.0)
        assert rep.adapter_only_pairs == 333 and r1.count(b":Adapter:") == 333
        outs.append((r1, r2))
    assert outs[0] == outs[1] == outs[2]


def test_speculation_overflow_falls_back_to_the_serial_kernel(rb, engine, golden, monkeypatch):
    """A read that needs more draws than assumed + margin cannot be finished by the read kernel; the batch is then redone by
    k_simulate.  With the margin forced to 0 the first deletion triggers that path (batches of 16 blocks: some fall back, some do not)."""
    monkeypatch.setenv("RSQ_SPEC_MARGIN", "0")
    monkeypatch.setenv("RSQ_BATCH_UNITS", "16")
    ref = rb.Reference.load_fasta(golden["small_ref"])
    r1, r2, rep = _simulate(engine, ref, seed=42, coverage=20.0)
    assert r1 == open(golden["r1"], "rb").read()
    assert r2 == open(golden["r2"], "rb").read()


@pytest.mark.parametrize("path,depth", [("spec", None), ("spec", 3), ("serial", None)])
def test_methylation_golden_bit_exact(rb, engine, golden, monkeypatch, path, depth):
    """--methylation: bisulfite C->T conversions per unmethylated region (Simulator::CTConversion), on both kernel forms."""
    monkeypatch.setenv("RSQ_SIM_PATH", path)
    if depth:
        monkeypatch.setenv("RSQ_SPEC_DEPTH", str(depth))
    ref = rb.Reference.load_fasta(golden["small_ref"])
    ref.load_methylation(golden["meth_bed"])
    r1, r2, _ = _simulate(engine, ref, seed=42, coverage=20.0)
    assert r1 == open(golden["meth_r1"], "rb").read()
    assert r2 == open(golden["meth_r2"], "rb").read()


def test_methylation_against_reference_binary(rb, engine, golden, oracle, workdir):
    ref = rb.Reference.load_fasta(golden["small_ref"])
    ref.load_methylation(golden["meth_bed"])
    r1, r2, _ = _simulate(engine, ref, seed=77, coverage=9.0)
    o1, o2 = run_oracle_sim(oracle, golden["reseq"], golden["small_ref"], 77, 9.0, os.path.join(workdir, "ora_meth"),
                            extra=("--methylation", golden["meth_bed"]))
    assert r1 == open(o1, "rb").read()
    assert r2 == open(o2, "rb").read()


def test_malformed_methylation_file_is_an_error(rb, golden, workdir):
    ref = rb.Reference.load_fasta(golden["small_ref"])
    bad = os.path.join(workdir, "bad.bed")
    open(bad, "w").write("chr1\t10\t5\t0.5\n")
    with pytest.raises(rb.RsqError, match="Third field"):
        ref.load_methylation(bad)


def test_systematic_error_profile_write_and_read_against_reference_binary(rb, engine, golden, oracle, workdir):
    """--writeSysError (Simulator::CreateSystematicErrorProfile) and --readSysError (ReadSystematicErrors)."""
    fa = os.path.join(workdir, "sysref.fa")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_synthetic.py"), "reference", fa, "--sizes", "21000,9000,12500", "--seed", "5",
                    "--prefix", "s"], check=True)
    ref = rb.Reference.load_fasta(fa)
    mine = os.path.join(workdir, "sys_mine.fq")
    engine.create_systematic_error_profile(ref, 9, mine)
    theirs = os.path.join(workdir, "sys_ref.fq")
    o1, o2 = run_oracle_sim(oracle, golden["reseq"], fa, 9, 10.0, os.path.join(workdir, "ora_sys"), extra=("--writeSysError", theirs))
    assert open(mine, "rb").read() == open(theirs, "rb").read()
    r1, r2, _ = _simulate(engine, ref, seed=9, coverage=10.0, sys_error_file=mine)
    assert r1 == open(o1, "rb").read()
    assert r2 == open(o2, "rb").read()
    # a different seed with the same error file (plain --readSysError)
    o1, o2 = run_oracle_sim(oracle, golden["reseq"], fa, 10, 10.0, os.path.join(workdir, "ora_sys2"), extra=("--readSysError", theirs))
    r1, r2, _ = _simulate(engine, ref, seed=10, coverage=10.0, sys_error_file=theirs)
    assert r1 == open(o1, "rb").read()
    assert r2 == open(o2, "rb").read()


def test_ref_bias_models_against_reference_binary(rb, engine, golden, oracle, workdir):
    """--refBias draw (consumes master-stream draws first) and --refBias file (several coverage groups)."""
    fa = os.path.join(workdir, "biasref.fa")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_synthetic.py"), "reference", fa, "--sizes", "18000,14000,9000", "--seed", "8",
                    "--prefix", "b"], check=True)
    ref = rb.Reference.load_fasta(fa)
    r1, r2, _ = _simulate(engine, ref, seed=21, coverage=10.0, ref_bias_model=2)
    o1, o2 = run_oracle_sim(oracle, golden["reseq"], fa, 21, 10.0, os.path.join(workdir, "ora_draw"), extra=("--refBias", "draw"))
    assert r1 == open(o1, "rb").read() and r2 == open(o2, "rb").read()
    bias_file = os.path.join(workdir, "bias.txt")
    open(bias_file, "w").write("b1 0.4\n>b2 some description 2.5\nb3\t1.0\n")
    r1, r2, _ = _simulate(engine, ref, seed=22, coverage=10.0, ref_bias_model=3, ref_bias_file=bias_file)
    o1, o2 = run_oracle_sim(oracle, golden["reseq"], fa, 22, 10.0, os.path.join(workdir, "ora_file"), extra=("--refBias", "file", "--refBiasFile", bias_file))
    assert r1 == open(o1, "rb").read() and r2 == open(o2, "rb").read()
    with pytest.raises(rb.RsqError, match="reference sequence biases"):
        open(bias_file, "w").write("b1 0.4\n")
        engine.prepare(ref, seed=22, coverage=10.0, ref_bias_model=3, ref_bias_file=bias_file)


@pytest.mark.parametrize("batch", [None, "9"])
def test_dropin_simulate_call_writes_files(rb, golden, workdir, monkeypatch, batch):
    if batch:
        monkeypatch.setenv("RSQ_BATCH_UNITS", batch)   # several batches through the staging buffers and the writer thread
    prof = rb.Profile.load_flat(golden["flat"])
    ref = rb.Reference.load_fasta(golden["small_ref"])
    o1, o2 = os.path.join(workdir, "d1.fq"), os.path.join(workdir, "d2.fq")
    rb.simulate(prof, ref, o1, o2, seed=42, coverage=20.0)
    assert open(o1, "rb").read() == open(golden["r1"], "rb").read()
    assert open(o2, "rb").read() == open(golden["r2"], "rb").read()


@pytest.mark.parametrize("path", ["spec", "serial"])
def test_error_model_bit_exact(engine, golden, workdir, monkeypatch, path):
    """seqToIllumina (ApplyErrorsAndQualityToFastaInput) against the reference's output, on the speculative kernels (the batch of
    input records is one unit of reads in lock step) and on the serial kernel (one warp per batch)."""
    monkeypatch.setenv("RSQ_SIM_PATH", path)
    out = os.path.join(workdir, f"em_gpu_{path}.fq")
    rep = engine.apply_error_model(golden["em_in"], out, 7)
    assert open(out, "rb").read() == open(golden["em_out"], "rb").read()
    assert rep.pairs == 9000
    assert (rep.spec_rounds > 0) == (path == "spec")


def test_error_model_batches_serial_and_speculative_agree(engine, golden, workdir, monkeypatch):
    """Several batches = several units of the speculative kernels.  The reference cannot run more than one batch (its writer waits
    for a counter nobody increments, Simulator.cpp:184-213), so the batch size is shrunk and the two kernel forms are compared."""
    monkeypatch.setenv("RSQ_EM_BATCH", "700")
    outs = []
    for path in ("serial", "spec"):
        monkeypatch.setenv("RSQ_SIM_PATH", path)
        out = os.path.join(workdir, f"em_batches_{path}.fq")
        rep = engine.apply_error_model(golden["em_in"], out, 7)
        assert rep.pairs == 9000 and rep.blocks == 13
        outs.append(open(out, "rb").read())
    assert outs[0] == outs[1] and outs[0].count(b"\n") == 4 * 9000


@pytest.mark.parametrize("path", ["spec", "serial"])
def test_error_model_three_batches_equal_the_reference(engine, golden, oracle_optional, workdir, monkeypatch, path):
    """27 000 input records = three batches of kBatchSizeErrorModelOnly, each with its own seed of the master stream: same bytes as the
    reference's own SimulateErrorModelOnly (hash in tests/golden/em_multibatch_sha256.json; on a box that has the oracle, byte for byte)."""
    import hashlib
    import json
    sys.path.insert(0, GOLDEN_DIR)
    import make_em_multibatch as mk
    want = json.load(open(os.path.join(GOLDEN_DIR, "em_multibatch_sha256.json")))
    monkeypatch.setenv("RSQ_SIM_PATH", path)
    fa, out = os.path.join(workdir, "em3.fa"), os.path.join(workdir, f"em3_{path}.fq")
    assert mk.write_input(fa) == want["records"]
    rep = engine.apply_error_model(fa, out, want["seed"])
    data = open(out, "rb").read()
    assert rep.pairs == want["records"] and rep.blocks == 3
    assert len(data) == want["bytes"] and hashlib.sha256(data).hexdigest() == want["sha256"]
    if oracle_optional:
        ref_out = os.path.join(workdir, "em3_oracle.fq")
        mk.run_oracle(oracle_optional["dump"], golden["reseq"], fa, ref_out)
        assert data == open(ref_out, "rb").read()


def test_error_model_rejects_malformed_header(engine, workdir):
    import reseq_b200
    bad = os.path.join(workdir, "bad_em.fa")
    open(bad, "w").write(">x 1;10;ACGT;!!!!\nACGTACGT\n")
    with pytest.raises(reseq_b200.RsqError):
        engine.apply_error_model(bad, os.path.join(workdir, "bad_em.fq"), 1)
    assert not os.path.exists(os.path.join(workdir, "bad_em.fq"))


@pytest.mark.parametrize("jump_min", [None, "65536"])
def test_stage_arrays_match_reference(rb, engine, golden, oracle, workdir, monkeypatch, jump_min):
    # jump_min: forces the jump-ahead form of the master stream (k_master_jump, segments generated in parallel) on this small genome
    if jump_min:
        monkeypatch.setenv("RSQ_MASTER_JUMP_MIN", jump_min)
    """Systematic errors (both strands, adapters) and thresholds against the reference's in-memory values."""
    from reseq_b200.flatfile import read_flat
    stage = os.path.join(workdir, "stage_gpu.flat")
    subprocess.run([oracle["dump"], "sim", golden["reseq"], golden["small_ref"], "42", "20", stage], check=True, timeout=600,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    st = read_flat(stage)
    ref = rb.Reference.load_fasta(golden["small_ref"])
    rep = engine.prepare(ref, seed=42, coverage=20.0)
    assert rep.bias_normalization == st["sim.bias_normalization"][0]
    assert rep.total_pairs_aim == st["sim.total_pairs"][0]
    thr = engine.fetch("thresholds", "float64")
    assert np.array_equal(thr, st["sim.thresholds.0"])
    blocks = engine.fetch("blocks", "uint64").reshape(-1, 4)   # BlockDesc: ref id | start, block id | first methylation id, seed, first variant id | pad
    assert np.array_equal(blocks[:, 2], st["sim.block_seed"])
    fwd = engine.fetch("sys_fwd").reshape(-1, 2)
    rev = engine.fetch("sys_rev").reshape(-1, 2)
    lens = [len(st[f"sim.ref.{i}"]) for i in range(4)]
    offs = np.concatenate([[0], np.cumsum(lens)])
    exp_f, exp_r = [], []
    for rid, start in zip(st["sim.block_ref"], st["sim.block_start"]):
        L = lens[rid]
        e = min(start + 1000, L)
        exp_f.append(fwd[offs[rid] + start: offs[rid] + e])
        exp_r.append(rev[offs[rid] + L - e: offs[rid] + L - start])
    assert np.array_equal(np.concatenate(exp_f).ravel(), st["sim.sys_fwd"])
    assert np.array_equal(np.concatenate(exp_r).ravel(), st["sim.sys_rev"])
    refseq = engine.fetch("reference")
    assert np.array_equal(refseq, np.concatenate([st[f"sim.ref.{i}"] for i in range(4)]))


@pytest.mark.parametrize("seed,coverage", [(7, 6.0), (123456789, 35.0)])
def test_against_reference_binary(rb, engine, golden, oracle, workdir, seed, coverage):
    ref = rb.Reference.load_fasta(golden["small_ref"])
    r1, r2, _ = _simulate(engine, ref, seed=seed, coverage=coverage)
    o1, o2 = run_oracle_sim(oracle, golden["reseq"], golden["small_ref"], seed, coverage, os.path.join(workdir, f"ora{seed}"))
    assert r1 == open(o1, "rb").read()
    assert r2 == open(o2, "rb").read()


@pytest.mark.parametrize("path", ["spec", "serial"])
def test_realistic_indel_profile_against_reference_binary(rb, golden, oracle, workdir, monkeypatch, path):
    """profile150r: the bench profile (InDel rate ~5e-5 per base as on real Illumina runs instead of the stress profile's 1.6e-3);
    here nearly every speculated read verifies, so the deep-window code paths run."""
    monkeypatch.setenv("RSQ_SIM_PATH", path)
    eng = rb.Engine(rb.Profile.load_flat(golden["flat_r"]), 0)
    try:
        ref = rb.Reference.load_fasta(golden["small_ref"])
        r1, r2, _ = _simulate(eng, ref, seed=5, coverage=25.0)
    finally:
        eng.close()
    o1, o2 = run_oracle_sim(oracle, golden["reseq_r"], golden["small_ref"], 5, 25.0, os.path.join(workdir, "ora_r_" + path))
    assert r1 == open(o1, "rb").read()
    assert r2 == open(o2, "rb").read()


@pytest.mark.parametrize("path", ["spec", "serial"])
def test_tiles_and_read_lengths_profile_against_reference_binary(rb, golden, oracle, workdir, monkeypatch, path):
    """profile150t: three tiles (a tile is drawn per pair, GeneralRandomDistributions::TileId, and selects the quality / base-call
    tables) and two read lengths per segment (GeneralRandomDistributions::ReadLength draws one per read)."""
    monkeypatch.setenv("RSQ_SIM_PATH", path)
    eng = rb.Engine(rb.Profile.load_flat(golden["flat_t"]), 0)
    try:
        ref = rb.Reference.load_fasta(golden["small_ref"])
        r1, r2, _ = _simulate(eng, ref, seed=11, coverage=25.0)
    finally:
        eng.close()
    o1, o2 = run_oracle_sim(oracle, golden["reseq_t"], golden["small_ref"], 11, 25.0, os.path.join(workdir, "ora_t_" + path))
    assert r1 == open(o1, "rb").read()
    assert r2 == open(o2, "rb").read()
    tiles = {line.split(b":")[4] for line in r1.split(b"\n")[0::4] if line}
    assert tiles == {b"1101", b"1102", b"2205"}
    assert {len(s) for s in r1.split(b"\n")[1::4]} == {144, 150}


def test_other_reference_and_prefix_against_reference_binary(rb, engine, golden, oracle, workdir):
    fa = os.path.join(workdir, "other.fa")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_synthetic.py"), "reference", fa, "--sizes", "61000,1001,2500", "--seed", "99",
                    "--n-rate", "0.01", "--prefix", "ctg"], check=True)
    ref = rb.Reference.load_fasta(fa)
    r1, r2, _ = _simulate(engine, ref, seed=5, coverage=12.0, record_base_identifier="Sim")
    o1, o2 = run_oracle_sim(oracle, golden["reseq"], fa, 5, 12.0, os.path.join(workdir, "ora_other"), extra=("--recordBaseIdentifier", "Sim"))
    assert r1 == open(o1, "rb").read()
    assert r2 == open(o2, "rb").read()


def test_num_read_pairs_option_against_reference_binary(rb, engine, golden, oracle, workdir):
    ref = rb.Reference.load_fasta(golden["small_ref"])
    r1, r2, _ = _simulate(engine, ref, seed=3, num_read_pairs=2500)
    o1, o2 = run_oracle_sim(oracle, golden["reseq"], golden["small_ref"], 3, 0, os.path.join(workdir, "ora_np"), extra=("--numReads", "2500"))
    assert r1 == open(o1, "rb").read()
    assert r2 == open(o2, "rb").read()


@pytest.mark.parametrize("switch", ["noInDelErrors", "noSubstitutionErrors", "errorMutliplier"])
def test_error_switches_against_reference_binary(rb, golden, oracle, workdir, switch):
    """--noInDelErrors / --noSubstitutionErrors / --errorMutliplier 2.5 (ProbabilityEstimates.h:1516-1549)."""
    prof = rb.Profile.load(golden["reseq"], golden["ipf"])
    extra = ("--" + switch,)
    if switch == "noInDelErrors":
        prof.remove_indel_errors()
    elif switch == "noSubstitutionErrors":
        prof.remove_substitution_errors()
    else:
        prof.change_error_rate(2.5)
        extra = ("--errorMutliplier", "2.5")
    eng = rb.Engine(prof, 0)
    ref = rb.Reference.load_fasta(golden["small_ref"])
    r1, r2, _ = _simulate(eng, ref, seed=11, coverage=8.0)
    eng.close()
    o1, o2 = run_oracle_sim(oracle, golden["reseq"], golden["small_ref"], 11, 8.0, os.path.join(workdir, "ora_" + switch), extra=extra)
    assert r1 == open(o1, "rb").read()
    assert r2 == open(o2, "rb").read()


def test_cli_dropin(golden, workdir, library):
    """reseq-b200 illuminaPE / seqToIllumina with the reference's own option names, reading X.reseq + X.reseq.ipf."""
    from reseq_b200 import build
    o1, o2 = os.path.join(workdir, "cli1.fq"), os.path.join(workdir, "cli2.fq")
    subprocess.run([build.CLI, "illuminaPE", "-j", "4", "-s", golden["reseq"], "-R", golden["small_ref"], "--ipfIterations", "0", "--seed", "42",
                    "-c", "20", "-1", o1, "-2", o2, "--verbosity", "1"], check=True, timeout=600)
    assert open(o1, "rb").read() == open(golden["r1"], "rb").read()
    assert open(o2, "rb").read() == open(golden["r2"], "rb").read()
    em = os.path.join(workdir, "cli_em.fq")
    subprocess.run([build.CLI, "seqToIllumina", "-i", golden["em_in"], "-o", em, "-s", golden["reseq"], "--ipfIterations", "0", "--seed", "7",
                    "--verbosity", "1"], check=True, timeout=600)
    assert open(em, "rb").read() == open(golden["em_out"], "rb").read()
    res = subprocess.run([build.CLI, "illuminaPE", "-s", golden["reseq"], "-R", golden["small_ref"], "-b", "x.bam"], capture_output=True, text=True)
    assert res.returncode == 1 and "does not replace" in res.stderr
    # -V: the file is read and checked like the reference does
    res = subprocess.run([build.CLI, "illuminaPE", "-s", golden["reseq"], "-R", golden["small_ref"], "-V", "x.vcf"], capture_output=True, text=True)
    assert res.returncode == 1 and "Could not open vcf file 'x.vcf'" in res.stderr
    res = subprocess.run([build.CLI, "illuminaPE", "-s", golden["reseq"], "-R", golden["small_ref"], "-V", os.path.join(golden["dir"], "simref_small_var_bad_overlap.vcf")],
                         capture_output=True, text=True)
    assert res.returncode == 1 and "overlaps with a previous variant" in res.stderr
    res = subprocess.run([build.CLI, "illuminaPE", "-s", golden["reseq"], "-R", golden["small_ref"], "--ipfIterations", "0", "--seed", "42", "-c", "20",
                          "-1", o1 + ".v", "-2", o2 + ".v", "-V", os.path.join(golden["dir"], "simref_small_var.vcf")], capture_output=True, text=True)
    import lzma
    assert res.returncode == 0, res.stderr
    for path, name in ((o1 + ".v", "sim_small_var_seed42_R1.fq.xz"), (o2 + ".v", "sim_small_var_seed42_R2.fq.xz")):
        assert open(path, "rb").read() == lzma.open(os.path.join(golden["dir"], name)).read()


@pytest.mark.parametrize("shards", [3, 7])
def test_shards_concatenate_to_the_whole_run(rb, engine, golden, shards):
    """A shard only computes systematic errors and block seeds for the sequences it has SimBlocks in; the master-stream draws
    of the other sequences are skipped by jump-ahead (master_skip).  simref_small has four sequences (30 + 22 + 0 + 15 blocks),
    so with 3 and 7 shards every shard skips at least one sequence, in front of, between and behind the ones it needs."""
    ref = rb.Reference.load_fasta(golden["small_ref"])
    w1, w2, rep = _simulate(engine, ref, seed=42, coverage=20.0)
    whole = (w1, w2, int(rep.pairs))   # the engine reuses one report object
    assert w1 == open(golden["r1"], "rb").read()
    parts1, parts2, pairs = b"", b"", 0
    for i in range(shards):
        a, b, rep = _simulate(engine, ref, seed=42, coverage=20.0, shard_index=i, shard_count=shards)
        parts1 += a
        parts2 += b
        pairs += rep.pairs
    assert parts1 == whole[0] and parts2 == whole[1] and pairs == whole[2]


def test_too_short_reference_is_an_error(rb, engine):
    ref = rb.Reference.from_memory(["tiny"], [b"ACGT" * 50])
    with pytest.raises(rb.RsqError, match="too short"):
        engine.prepare(ref, seed=1, coverage=5.0)


def test_c1_real_ecoli_sequence_equals_the_reference(rb, golden, workdir):
    """BASELINE config C1/C2 on the reference's own E. coli fixture (test/ecoli-GCF_000005845.2_ASM584v2_genomic.fa, committed xz-compressed):
    2x150 profile150r, 30x, seed 42 - sha256 of the unmodified reference's `-j 1` FASTQ (tests/golden/fullsize_c2_sha256.json, entry
    ecoli_profile150r).  The FASTA goes through the engine's own reader (80-column lines, full header line as id)."""
    import json
    import lzma
    want = json.load(open(os.path.join(golden["dir"], "fullsize_c2_sha256.json")))["ecoli_profile150r"]
    fa = os.path.join(workdir, "ecoli.fa")
    with lzma.open(os.path.join(golden["dir"], "ecoli_GCF_000005845.2.fa.xz")) as src, open(fa, "wb") as dst:
        dst.write(src.read())
    ref = rb.Reference.load_fasta(fa)
    assert ref.total_size == want["size"]
    eng = rb.Engine(rb.Profile.load_flat(golden["flat_r"]), 0)
    try:
        r1, r2, rep = _simulate(eng, ref, seed=want["seed"], coverage=want["coverage"])
    finally:
        eng.close()
    assert rep.pairs == want["pairs"] and [len(r1), len(r2)] == want["bytes"]
    assert hashlib.sha256(r1).hexdigest() == want["r1"] and hashlib.sha256(r2).hexdigest() == want["r2"]
    assert b":NC_000913.3:" in r1[:200]


def test_c1_four_pair_profile_fails_like_the_reference(rb, golden, workdir):
    """BASELINE config C1 literally (profile of test/ecoli-SRR490124-4pairs.bam, built by the reference's own stats + IPF code and committed as
    c1_ecoli4pairs.reseq*): four read pairs do not give two usable insert lengths, and the reference refuses to simulate from it -
    `Sampling insert lengths did not find at least two usable lengths.` (FragmentDistributionStats.cpp:1686).  Same answer here."""
    import lzma
    paths = []
    for ext in (".reseq", ".reseq.ipf"):
        dst = os.path.join(workdir, "c1" + ext)
        with lzma.open(os.path.join(golden["dir"], "c1_ecoli4pairs" + ext + ".xz")) as src, open(dst, "wb") as o:
            o.write(src.read())
        paths.append(dst)
    eng = rb.Engine(rb.Profile.load(*paths), 0)
    try:
        ref = rb.Reference.load_fasta(golden["small_ref"])
        with pytest.raises(rb.RsqError, match="Sampling insert lengths did not find at least two usable lengths"):
            eng.prepare(ref, seed=42, coverage=2.0)
    finally:
        eng.close()


@pytest.mark.parametrize("prof,key", [("profile150", "flat"), ("profile150r", "flat_r")])
def test_full_size_equals_the_reference(rb, golden, oracle_optional, workdir, prof, key):
    """BASELINE config C2 at full size (the workload bench.py times): the engine's FASTQ has the sha256 of what the unmodified reference
    wrote with `-j 1` for the same synthetic 4 641 652-bp sequence, profile, seed and coverage (tests/golden/fullsize_c2_sha256.json, made
    by tests/golden/make_fullsize_hashes.py).  This is the only size at which the jump-ahead master stream (segments of 2^16..2^22 draws)
    and the deep speculation windows are active.  RSQ_LIVE_ORACLE=1 additionally runs the reference binary on this box (3-4 minutes)."""
    import json
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_synthetic
    want = json.load(open(os.path.join(golden["dir"], "fullsize_c2_sha256.json")))[prof]
    seq = make_synthetic.gen_reference([want["size"]], want["ref_seed"])[0]
    ref = rb.Reference.from_memory(["ecoli_sized synthetic"], [seq.encode()])
    eng = rb.Engine(rb.Profile.load_flat(golden[key]), 0)
    try:
        r1, r2, rep = _simulate(eng, ref, seed=want["seed"], coverage=want["coverage"])
    finally:
        eng.close()
    assert rep.pairs == want["pairs"] and [len(r1), len(r2)] == want["bytes"]
    assert hashlib.sha256(r1).hexdigest() == want["r1"]
    assert hashlib.sha256(r2).hexdigest() == want["r2"]
    assert rep.spec_rounds > 0
    if oracle_optional and os.environ.get("RSQ_LIVE_ORACLE"):
        fa = os.path.join(workdir, "full_" + prof + ".fa")
        with open(fa, "w") as f:
            f.write(">ecoli_sized synthetic\n" + seq + "\n")
        o1, o2 = run_oracle_sim(oracle_optional, golden["reseq_r" if prof.endswith("r") else "reseq"], fa, want["seed"], want["coverage"], os.path.join(workdir, "full_" + prof))
        assert open(o1, "rb").read() == r1 and open(o2, "rb").read() == r2


def test_full_size_properties(rb, engine, workdir, monkeypatch):
    """BASELINE config 2 size (4.64 Mbp, 30x): determinism, record structure, pairing, pair count near the aim."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_synthetic
    seq = make_synthetic.gen_reference([4_641_652], 1234)[0]
    ref = rb.Reference.from_memory(["ecoli_sized synthetic"], [seq.encode()])
    r1, r2, rep = _simulate(engine, ref, seed=42, coverage=30.0)
    assert abs(rep.pairs - rep.total_pairs_aim) < 0.02 * rep.total_pairs_aim
    l1, l2 = r1.split(b"\n"), r2.split(b"\n")
    assert len(l1) == len(l2) == 4 * rep.pairs + 1
    ids1 = [x.split(b" ")[0] for x in l1[0:-1:4]]
    ids2 = [x.split(b" ")[0] for x in l2[0:-1:4]]
    assert ids1 == ids2 and all(i.startswith(b"@ReseqRead") for i in ids1[:1000])
    assert all(len(s) == 150 for s in l1[1:-1:4]) and all(len(q) == 150 for q in l2[3:-1:4])
    assert set(b"".join(l1[1:2000:4])) <= set(b"ACGTN")
    blocks = [int(i[len(b"@ReseqRead"):].split(b"_")[0]) for i in ids1]
    assert blocks == sorted(blocks), "records must come in block order like the reference's 1-thread run"
    h = hashlib.sha256(r1).hexdigest(), hashlib.sha256(r2).hexdigest()
    spec_rounds = rep.spec_rounds
    r1b, r2b, _ = _simulate(engine, ref, seed=42, coverage=30.0)
    assert (hashlib.sha256(r1b).hexdigest(), hashlib.sha256(r2b).hexdigest()) == h
    r1c, _, _ = _simulate(engine, ref, seed=43, coverage=30.0)
    assert hashlib.sha256(r1c).hexdigest() != h[0]
    # the serial forms (one warp per SimBlock, master stream as one recurrence) write the same bytes as the
    # speculative two-phase kernels fed by the jump-ahead master stream
    monkeypatch.setenv("RSQ_SIM_PATH", "serial")
    monkeypatch.setenv("RSQ_MASTER_JUMP_MIN", str(1 << 40))
    r1d, r2d, repd = _simulate(engine, ref, seed=42, coverage=30.0)
    assert repd.spec_rounds == 0 and spec_rounds > 0
    assert (hashlib.sha256(r1d).hexdigest(), hashlib.sha256(r2d).hexdigest()) == h
