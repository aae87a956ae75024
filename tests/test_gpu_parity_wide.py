"""GPU parity tests, second file (runs after test_gpu_parity.py): the formats and shapes either side of the golden configuration -
gzip files in and out, 2x250 reads (BASELINE config C4's read length), 300-base seqToIllumina fragments (config C3's shape).
Everything goes through the C ABI; the checker is the reference binary run on the spot and the committed golden FASTQ."""
import os
import subprocess
import sys

import pytest

from conftest import run_oracle_sim
from test_gpu_parity import _simulate, engine, rb  # noqa: F401  (fixtures)

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("gz_mode", ["host", "device"])
@pytest.mark.parametrize("batch", [None, "7"])
def test_dropin_simulate_call_gzip_files(rb, golden, workdir, monkeypatch, batch, gz_mode):
    """Output names ending in .gz are written gzip-compressed like SeqAn's SeqFileOut does (members deflated on the host cores by the
    writer threads); the reference is read from a .fa.gz.  The inflated text is the reference's golden FASTQ."""
    import gzip
    monkeypatch.setenv("RSQ_GZIP", gz_mode)   # device: k_deflate_members makes the gzip members on the GPU (deflate_core.cuh)
    if batch:
        monkeypatch.setenv("RSQ_BATCH_UNITS", batch)
    fa_gz = os.path.join(workdir, "small_ref_copy.fa.gz")
    with open(golden["small_ref"], "rb") as f, gzip.open(fa_gz, "wb") as o:
        o.write(f.read())
    prof = rb.Profile.load_flat(golden["flat"])
    ref = rb.Reference.load_fasta(fa_gz)
    o1, o2 = os.path.join(workdir, f"gz{batch}{gz_mode}_1.fq.gz"), os.path.join(workdir, f"gz{batch}{gz_mode}_2.fq.gz")
    rb.simulate(prof, ref, o1, o2, seed=42, coverage=20.0)
    assert open(o1, "rb").read(2) == b"\x1f\x8b"
    assert gzip.open(o1).read() == open(golden["r1"], "rb").read()
    assert subprocess.run(["gzip", "-dc", o2], capture_output=True, check=True).stdout == open(golden["r2"], "rb").read()
    assert os.path.getsize(o1) < 0.6 * os.path.getsize(golden["r1"])


def test_error_model_gzip_in_and_out(engine, golden, workdir):
    import gzip
    em_gz = os.path.join(workdir, "em_in.fa.gz")
    with open(golden["em_in"], "rb") as f, gzip.open(em_gz, "wb") as o:
        o.write(f.read())
    out = os.path.join(workdir, "em_out.fq.gz")
    engine.apply_error_model(em_gz, out, 7)
    assert gzip.open(out).read() == open(golden["em_out"], "rb").read()


@pytest.mark.parametrize("path", ["spec", "serial"])
def test_error_model_300_base_fragments_against_reference_binary(engine, golden, oracle, workdir, monkeypatch, path):
    """BASELINE config C3's shape: seqToIllumina on 300-base fragments (reads of 150 bases end inside the fragment, so no adapter part),
    one batch (the reference cannot run more, see the test below), checked against the reference binary run here."""
    monkeypatch.setenv("RSQ_SIM_PATH", path)
    fa = os.path.join(workdir, "c3ref.fa")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_synthetic.py"), "reference", fa, "--sizes", "90000", "--seed", "31"], check=True)
    frags = os.path.join(workdir, "c3_frags.fa")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_synthetic.py"), "fragments", fa, "-", frags, "--n", "6000", "--len", "300", "--seed", "5"], check=True)
    theirs = os.path.join(workdir, f"c3_ref_{path}.fq")
    subprocess.run([oracle["reseq"], "seqToIllumina", "-j", "1", "--verbosity", "1", "-i", frags, "-s", golden["reseq"], "--ipfIterations", "0", "--seed", "19",
                    "-o", theirs], check=True, timeout=900, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    mine = os.path.join(workdir, f"c3_mine_{path}.fq")
    rep = engine.apply_error_model(frags, mine, 19)
    assert rep.pairs == 6000
    assert open(mine, "rb").read() == open(theirs, "rb").read()


@pytest.mark.parametrize("path,meth", [("spec", False), ("serial", False), ("spec", True)])
def test_250_base_reads_against_reference_binary(rb, golden, oracle, workdir, monkeypatch, path, meth):
    """profile250: 2x250 reads, the read length of BASELINE config C4 (Drosophila, methylation BED): longer stream slices, record slots
    and CIGARs; with bisulfite conversion on the speculative path."""
    monkeypatch.setenv("RSQ_SIM_PATH", path)
    eng = rb.Engine(rb.Profile.load_flat(golden["flat_250"]), 0)
    try:
        ref = rb.Reference.load_fasta(golden["small_ref"])
        if meth:
            ref.load_methylation(golden["meth_bed"])
        r1, r2, _ = _simulate(eng, ref, seed=9, coverage=15.0)
    finally:
        eng.close()
    extra = ("--methylation", golden["meth_bed"]) if meth else ()
    o1, o2 = run_oracle_sim(oracle, golden["reseq_250"], golden["small_ref"], 9, 15.0, os.path.join(workdir, f"ora_250_{path}_{int(meth)}"), extra=extra)
    assert r1 == open(o1, "rb").read()
    assert r2 == open(o2, "rb").read()
    assert {len(s) for s in r1.split(b"\n")[1::4]} == {250}


def test_device_gzip_at_full_size(rb, golden, workdir, monkeypatch):
    """BASELINE config C2's size through the drop-in call with .gz names and the device deflate kernels: ~2 x 170 MB of text in
    ~2600 gzip members per file; the inflated files equal the plain files of the same run."""
    import gzip
    import hashlib
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_synthetic
    seq = make_synthetic.gen_reference([4_641_652], 1234)[0]
    ref = rb.Reference.from_memory(["ecoli_sized synthetic"], [seq.encode()])
    prof = rb.Profile.load_flat(golden["flat_r"])
    p1, p2 = os.path.join(workdir, "full_1.fq"), os.path.join(workdir, "full_2.fq")
    rb.simulate(prof, ref, p1, p2, seed=42, coverage=30.0)
    monkeypatch.setenv("RSQ_GZIP", "device")
    z1, z2 = os.path.join(workdir, "full_1.fq.gz"), os.path.join(workdir, "full_2.fq.gz")
    rb.simulate(prof, ref, z1, z2, seed=42, coverage=30.0)
    for plain, packed in ((p1, z1), (p2, z2)):
        want = hashlib.sha256(open(plain, "rb").read()).hexdigest()
        with gzip.open(packed) as f:
            h = hashlib.sha256()
            while True:
                chunk = f.read(1 << 24)
                if not chunk:
                    break
                h.update(chunk)
        assert h.hexdigest() == want
        assert os.path.getsize(packed) < 0.45 * os.path.getsize(plain)
    assert subprocess.run(["gzip", "-t", z1]).returncode == 0
