"""Multi-GPU tests (need at least two CUDA devices; skipped otherwise): engines joined into an NCCL group through the C ABI
(rsq_simulate_multi / rsq_engine_join_group) write the bytes of the single-GPU run, i.e. of the reference's 1-thread run."""
import lzma
import os
import threading

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def rb(library):
    import reseq_b200
    if library.rsq_device_count() < 2:
        pytest.skip("needs two CUDA devices")
    return reseq_b200


def _golden(golden, tag):
    return [lzma.open(os.path.join(golden["dir"], f"sim_small_{tag}seed42_R{k}.fq.xz")).read() for k in (1, 2)]


@pytest.mark.parametrize("n_gpus", [2, 4, 8])
@pytest.mark.parametrize("tag,gz", [("", False), ("var_", False), ("", True)])
def test_simulate_multi_writes_the_single_gpu_files(rb, library, golden, workdir, n_gpus, tag, gz):
    """rsq_simulate_multi: one engine + host thread per GPU, shard files appended in order = the golden FASTQ of the reference (plain and -V run)."""
    if library.rsq_device_count() < n_gpus:
        pytest.skip(f"needs {n_gpus} CUDA devices")
    import gzip
    prof = rb.Profile.load_flat(golden["flat"])
    ref = rb.Reference.load_fasta(golden["small_ref"])
    if tag:
        ref.load_variants(os.path.join(golden["dir"], "simref_small_var.vcf"))
    ext = ".fq.gz" if gz else ".fq"
    out = [os.path.join(workdir, f"multi{n_gpus}_{tag}R{k}{ext}") for k in (1, 2)]
    rep = rb.simulate_multi(prof, ref, out[0], out[1], seed=42, n_gpus=n_gpus, coverage=20.0)
    want = _golden(golden, tag)
    for path, w in zip(out, want):
        data = gzip.open(path).read() if gz else open(path, "rb").read()
        assert data == w
    assert rep.pairs == want[0].count(b"\n") // 4
    assert not [f for f in os.listdir(workdir) if f.startswith(".rsq_shard")]


def test_group_of_engines_all_reduces_pair_counts_and_shards_the_prologue(rb, golden):
    """Two engines on two devices joined through rsq_group_unique_id / rsq_engine_join_group from two host threads (what one rank per GPU does):
    each takes its shard, the bias sums come from the owners of the sequences over NCCL (same normalisation bit for bit), group_pairs is the whole run."""
    prof = rb.Profile.load_flat(golden["flat"])
    ref = rb.Reference.load_fasta(golden["small_ref"])
    single = rb.Engine(prof, 0)
    rep1 = single.prepare(ref, seed=42, coverage=20.0)
    norm = rep1.bias_normalization
    single.simulate()
    single.download()
    whole = [single.output(0), single.output(1)]
    single.close()
    ident = rb.group_unique_id()
    res = [None, None]

    def run(rank):
        eng = rb.Engine(prof, rank)
        try:
            eng.join_group(ident, rank, 2)
            rep = eng.prepare(ref, seed=42, coverage=20.0)
            bn = rep.bias_normalization
            rep = eng.simulate()
            eng.download()
            res[rank] = (eng.output(0), eng.output(1), bn, rep.group_pairs, rep.pairs, rep.shard_first, rep.blocks)
        finally:
            eng.close()
    threads = [threading.Thread(target=run, args=(k,)) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert res[0] is not None and res[1] is not None
    assert res[0][2] == norm and res[1][2] == norm
    assert res[0][0] + res[1][0] == whole[0] and res[0][1] + res[1][1] == whole[1]
    assert res[0][3] == res[1][3] == res[0][4] + res[1][4] == whole[0].count(b"\n") // 4
    assert res[0][5] == 0 and res[1][5] == res[0][6]
    plan = rb.shard_plan(prof, ref, 2)    # the host-only function of the same split (tests/test_multirank_cpu.py checks its properties)
    assert [res[0][5], res[1][5], res[1][5] + res[1][6]] == plan
