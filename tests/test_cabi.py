"""The C-ABI library loads, exports every symbol include/reseq_b200.h declares, and refuses to run without a GPU."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "reseq_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rsq_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported(library):
    from reseq_b200 import api
    syms = header_symbols()
    assert len(syms) >= 20
    assert sorted(api.SIGNATURES) == syms, "ctypes binding and header disagree"
    nm = subprocess.run(["nm", "-D", "--defined-only", api.lib_path()], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (rsq_[a-z_0-9]+)", nm))
    assert set(syms) <= exported
    for s in syms:
        assert getattr(library, s) is not None


def test_library_contains_sm100a_kernels(library):
    from reseq_b200 import api
    out = subprocess.run(["cuobjdump", "-lelf", api.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback(library, golden):
    import reseq_b200 as rb
    if library.rsq_device_count() > 0:
        pytest.skip("a GPU is present")
    prof = rb.Profile.load_flat(golden["flat"])
    with pytest.raises(rb.RsqError, match="no usable CUDA device"):
        rb.Engine(prof)


def test_errors_are_reported(library, workdir):
    import reseq_b200 as rb
    with pytest.raises(rb.RsqError):
        rb.Profile.load_flat(os.path.join(workdir, "does_not_exist.flat"))
    with pytest.raises(rb.RsqError, match="Could not open"):
        rb.Reference.load_fasta(os.path.join(workdir, "missing.fa"))
    bad = os.path.join(workdir, "bad.flat")
    open(bad, "wb").write(b"not a flat file")
    with pytest.raises(rb.RsqError, match="RSQFLAT1"):
        rb.Profile.load_flat(bad)


def test_reference_reader_semantics(library, workdir):
    """Reference::ReadFasta: IUPAC/unknown -> N, lower case accepted, multi-line records, ids keep the description."""
    import reseq_b200 as rb
    fa = os.path.join(workdir, "r.fa")
    open(fa, "w").write(">a first\nACGTacgt\nNNRY\n>b\nUU\n")
    ref = rb.Reference.load_fasta(fa)
    assert ref.num_sequences == 2 and ref.total_size == 14
    mem = rb.Reference.from_memory(["a first", "b"], [b"ACGTacgtNNRY", b"UU"])
    assert mem.total_size == 14


def test_flat_profile_roundtrip(library, golden, workdir):
    """The host loader keeps every array of the reference's in-memory profile (dump_tables) bit for bit."""
    import numpy as np
    import reseq_b200 as rb
    from reseq_b200.flatfile import read_flat
    prof = rb.Profile.load_flat(golden["flat"])
    out = os.path.join(workdir, "roundtrip.flat")
    prof.save_flat(out)
    a, b = read_flat(golden["flat"]), read_flat(out)
    assert set(a) == set(b)
    for k in a:
        assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("suffix", ["", "_r", "_t", "_250"])
def test_reseq_archive_loader_matches_reference_memory_image(library, golden, workdir, suffix):
    """rsq_profile_load(X.reseq, X.reseq.ipf) == what the reference holds after DataStats::Load + PrepareProcessing +
    ProbabilityEstimates::Estimate(0 iterations) + PrepareResult (golden flat file written by oracle/dump_tables), for all four golden
    profiles: profile150, profile150r (bench), profile150t (three tiles, two read lengths), profile250."""
    import numpy as np
    import reseq_b200 as rb
    from reseq_b200.flatfile import read_flat
    prof = rb.Profile.load(golden["reseq" + suffix], golden["ipf" + suffix])
    out = os.path.join(workdir, f"from_archive{suffix}.flat")
    prof.save_flat(out)
    a, b = read_flat(golden["flat" + suffix]), read_flat(out)
    assert set(a) == set(b)
    for k in a:
        assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]), k


def test_archive_loader_rejects_foreign_ipf(library, golden, workdir):
    import reseq_b200 as rb
    bad = os.path.join(workdir, "foreign.ipf")
    text = open(golden["ipf"]).read()
    head, rest = text.split(" ", 3)[:3], text.split(" ", 3)[3]
    # third token after the header is stats_creation_time_: change it
    toks = rest.split(" ", 3)
    toks[2] = str(int(toks[2]) + 1)
    open(bad, "w").write(" ".join(head) + " " + " ".join(toks))
    with pytest.raises(rb.RsqError, match="creation time"):
        rb.Profile.load(golden["reseq"], bad)
