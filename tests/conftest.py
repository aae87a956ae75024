import lzma
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _unxz(name, dst_dir):
    dst = os.path.join(dst_dir, name[:-3])
    if not os.path.exists(dst):
        with lzma.open(os.path.join(GOLDEN, name)) as f, open(dst, "wb") as o:
            shutil.copyfileobj(f, o)
    return dst


@pytest.fixture(scope="session")
def workdir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("rsq"))


@pytest.fixture(scope="session")
def golden(workdir):
    """Decompressed committed fixtures (tests/golden/make_golden.py made them with the reference's own code)."""
    out = {"dir": GOLDEN, "small_ref": os.path.join(GOLDEN, "simref_small.fa"), "meth_bed": os.path.join(GOLDEN, "simref_small_meth.bed")}
    for key, name in (("flat", "profile150.flat.xz"), ("reseq", "profile150.reseq.xz"), ("ipf", "profile150.reseq.ipf.xz"),
                      ("r1", "sim_small_seed42_R1.fq.xz"), ("r2", "sim_small_seed42_R2.fq.xz"),
                      ("em_in", "em_frags.fa.xz"), ("em_out", "em_seed7.fq.xz"),
                      ("meth_r1", "sim_small_meth_seed42_R1.fq.xz"), ("meth_r2", "sim_small_meth_seed42_R2.fq.xz"),
                      ("flat_r", "profile150r.flat.xz"), ("reseq_r", "profile150r.reseq.xz"), ("ipf_r", "profile150r.reseq.ipf.xz"),
                      ("flat_t", "profile150t.flat.xz"), ("reseq_t", "profile150t.reseq.xz"), ("ipf_t", "profile150t.reseq.ipf.xz"),
                      ("flat_250", "profile250.flat.xz"), ("reseq_250", "profile250.reseq.xz"), ("ipf_250", "profile250.reseq.ipf.xz"),
                      ("reseq_a", "profile150a.reseq.xz"), ("flat_q", "profile150q.flat.xz")):
        out[key] = _unxz(name, workdir)
    # profile150a = profile150 with InsertLengths()[0] = 700 (oracle/dump_tables patch_adapter_only): adapter-only pairs; same .ipf (same creation time)
    out["ipf_a"] = out["reseq_a"] + ".ipf"
    if not os.path.exists(out["ipf_a"]):
        shutil.copyfile(out["ipf"], out["ipf_a"])
    return out


@pytest.fixture(scope="session")
def library():
    from reseq_b200 import build
    build.build()
    import reseq_b200
    return reseq_b200.load_library()


def _oracle_paths():
    d = os.path.join(ROOT, "oracle", "_ref")
    return os.path.join(d, "reseq_oracle"), os.path.join(d, "dump_tables")


@pytest.fixture(scope="session")
def oracle():
    """The reference built from its own sources (oracle/Makefile).  Built here when /root/reference is present;
    on the GPU box the prebuilt binaries travel with the snapshot.  Tests that need it skip when it is absent."""
    exe, dump = _oracle_paths()
    if not (os.path.exists(exe) and os.path.exists(dump)) and os.path.isdir("/root/reference"):
        subprocess.run(["make", "-j8", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    if not (os.path.exists(exe) and os.path.exists(dump)):
        pytest.skip("oracle/_ref is not built and /root/reference is absent")
    return {"reseq": exe, "dump": dump}


@pytest.fixture(scope="session")
def oracle_optional():
    exe, dump = _oracle_paths()
    return {"reseq": exe, "dump": dump} if os.path.exists(exe) and os.path.exists(dump) else None


def run_oracle_sim(oracle, profile, ref, seed, coverage, out_prefix, threads=1, extra=()):
    r1, r2 = out_prefix + "_R1.fq", out_prefix + "_R2.fq"
    cmd = [oracle["reseq"], "illuminaPE", "-j", str(threads), "--verbosity", "1", "-s", profile, "-R", ref, "--ipfIterations", "0",
           "--seed", str(seed), "-c", str(coverage), "-1", r1, "-2", r2, *extra]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=1200)
    return r1, r2
