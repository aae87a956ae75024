"""Host text I/O either side of the path (reseq_b200/csrc/text_io.hpp): gzip output named like SeqAn expects it (.gz), gzip input
recognised by its magic bytes.  The checker is Python's gzip module, the `gzip` tool and, for input, the reference binary."""
import gzip
import lzma
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TWIN_DIR = os.path.join(ROOT, "tests", "host_twin")


@pytest.fixture(scope="module")
def text_io(workdir):
    exe = os.path.join(workdir, "text_io_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(TWIN_DIR, "text_io_check.cpp"), "-lz"], check=True)
    return exe


@pytest.fixture(scope="module")
def fastq_text(golden, workdir):
    text = open(golden["r1"], "rb").read()
    path = os.path.join(workdir, "tio_plain.fq")
    open(path, "wb").write(text * 3)   # ~several gzip members
    return path, text * 3


@pytest.mark.parametrize("piece,threads", [(0, 1), (0, 4), (700_001, 3), (1 << 20, 2), (100, 1)])
def test_gzip_sink_round_trip(text_io, fastq_text, workdir, piece, threads):
    path, text = fastq_text
    if piece == 100:
        text = text[:text.rfind(b"\n", 0, 20_000) + 1]
        path = os.path.join(workdir, "tio_short.fq")
        open(path, "wb").write(text)
    out = os.path.join(workdir, f"tio_{piece}_{threads}.fq.gz")
    res = subprocess.run([text_io, "compress", path, out, str(piece)], capture_output=True, text=True, env={**os.environ, "RSQ_GZIP_THREADS": str(threads)})
    assert res.returncode == 0, res.stderr
    assert f"text={len(text)}" in res.stdout and "compressed=1" in res.stdout
    assert gzip.open(out).read() == text                                               # Python's multi-member reader
    assert subprocess.run(["gzip", "-dc", out], capture_output=True, check=True).stdout == text   # the gzip tool
    assert subprocess.run([text_io, "cat", out], capture_output=True, check=True).stdout == text  # our own reader
    if piece != 100:   # every write() ends a member, so tiny writes cannot compress well
        assert os.path.getsize(out) < 0.6 * len(text)


def test_plain_sink_and_level(text_io, fastq_text, workdir):
    path, text = fastq_text
    out = os.path.join(workdir, "tio_plain_copy.fq")
    res = subprocess.run([text_io, "compress", path, out, "12345"], capture_output=True, text=True, check=True)
    assert "compressed=0" in res.stdout and open(out, "rb").read() == text
    sizes = {}
    for level in ("1", "9"):
        out = os.path.join(workdir, f"tio_l{level}.fq.gz")
        subprocess.run([text_io, "compress", path, out, "0"], check=True, capture_output=True, env={**os.environ, "RSQ_GZIP_LEVEL": level})
        assert gzip.open(out).read() == text
        sizes[level] = os.path.getsize(out)
    assert sizes["9"] < sizes["1"]


def test_empty_gzip_output_is_a_valid_stream(text_io, workdir):
    out = os.path.join(workdir, "tio_empty.fq.gz")
    subprocess.run([text_io, "empty", out], check=True)
    assert os.path.getsize(out) > 0 and gzip.open(out).read() == b""
    assert subprocess.run(["gzip", "-t", out]).returncode == 0


def test_bzip2_output_is_rejected(text_io, fastq_text, workdir):
    res = subprocess.run([text_io, "compress", fastq_text[0], os.path.join(workdir, "x.fq.bz2"), "0"], capture_output=True, text=True)
    assert res.returncode == 3 and "bzip2" in res.stderr


def test_truncated_gzip_input_is_detected(text_io, fastq_text, workdir):
    out = os.path.join(workdir, "tio_trunc_src.fq.gz")
    subprocess.run([text_io, "compress", fastq_text[0], out, "0"], check=True, capture_output=True)
    cut = os.path.join(workdir, "tio_trunc.fq.gz")
    open(cut, "wb").write(open(out, "rb").read()[:-1000])
    res = subprocess.run([text_io, "cat", cut], capture_output=True)
    assert res.returncode == 2


def test_reference_reader_takes_gzip_fasta(library, golden, workdir):
    """Reference::ReadFasta goes through SeqAn's SeqFileIn, which inflates gzip input (Reference.cpp:758-811)."""
    import reseq_b200 as rb
    fa_gz = os.path.join(workdir, "simref_small.fa.gz")
    with open(golden["small_ref"], "rb") as f, gzip.open(fa_gz, "wb") as o:
        o.write(f.read())
    plain, packed = rb.Reference.load_fasta(golden["small_ref"]), rb.Reference.load_fasta(fa_gz)
    assert packed.num_sequences == plain.num_sequences == 4 and packed.total_size == plain.total_size
    bad = os.path.join(workdir, "cut.fa.gz")
    open(bad, "wb").write(open(fa_gz, "rb").read()[:-200])
    with pytest.raises(rb.RsqError, match="corrupt or truncated"):
        rb.Reference.load_fasta(bad)


def test_reference_binary_reads_our_gzip_members(text_io, oracle, golden, workdir):
    """The reference's own SeqAn reader decodes a multi-member .gz written by TextSink: its simulation from our compressed copy of
    the FASTA equals its golden output (the drop-in direction that matters for --readSysError / reference files we write)."""
    from conftest import run_oracle_sim
    big = os.path.join(workdir, "padded.fa")
    src = open(golden["small_ref"], "rb").read()
    open(big, "wb").write(src)
    fa_gz = os.path.join(workdir, "members.fa.gz")
    subprocess.run([text_io, "compress", big, fa_gz, "20000"], check=True, capture_output=True)   # several members
    assert gzip.open(fa_gz).read() == src
    r1, r2 = run_oracle_sim(oracle, golden["reseq"], fa_gz, 42, 20, os.path.join(workdir, "ora_gzmembers"))
    assert open(r1, "rb").read() == open(golden["r1"], "rb").read()


@pytest.fixture(scope="module")
def deflate_twin(workdir):
    exe = os.path.join(workdir, "deflate_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(TWIN_DIR, "deflate_check.cpp")], check=True)
    return exe


def _fib_bytes():
    """Byte counts growing like Fibonacci numbers: an unrestricted Huffman code would be deeper than deflate's 15 bits."""
    import random
    fib = [1, 1]
    while len(fib) < 30:
        fib.append(fib[-1] + fib[-2])
    data = list(b"".join(bytes([i]) * min(f, 60000) for i, f in enumerate(fib)))
    random.Random(1).shuffle(data)
    return bytes(data)[:131072]


@pytest.mark.parametrize("case", ["fastq", "one_byte", "run", "random", "member_exact", "member_plus_one", "crc_piece_511", "crc_piece_513",
                                  "slice_edge", "length_limit"])
def test_device_deflate_member_code_on_the_cpu_twin(deflate_twin, fastq_text, workdir, case):
    """deflate_core.cuh (the code of k_deflate_members) instantiated with one thread: every member is a valid gzip member (header, one
    dynamic-Huffman block, CRC-32 built from 512-byte pieces with the GF(2) shift operator, ISIZE) and inflates to the input."""
    import random
    rnd = random.Random(7)
    data = {
        "fastq": lambda: fastq_text[1][:2_000_000],
        "one_byte": lambda: b"A",
        "run": lambda: b"I" * 300_000,
        "random": lambda: os.urandom(300_000),
        "member_exact": lambda: (b"ACGT" * 40000)[:131072],
        "member_plus_one": lambda: (b"ACGTTGCA" * 40000)[:131073],
        "crc_piece_511": lambda: bytes(rnd.choice(b"ACGT") for _ in range(511)),
        "crc_piece_513": lambda: bytes(rnd.choice(b"ACGTN!#IJ") for _ in range(513)),
        "slice_edge": lambda: bytes(rnd.choice(b"ACGT") for _ in range(16384 + 3)),
        "length_limit": _fib_bytes,
    }[case]()
    src, out = os.path.join(workdir, f"dfl_{case}.in"), os.path.join(workdir, f"dfl_{case}.gz")
    open(src, "wb").write(data)
    res = subprocess.run([deflate_twin, src, out], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert gzip.open(out).read() == data
    assert subprocess.run(["gzip", "-t", out]).returncode == 0
    if case == "fastq":
        assert os.path.getsize(out) < 0.4 * len(data)   # zlib level 1 reaches 0.31 on this text, level 6 0.23
    if case == "random":
        assert os.path.getsize(out) < 1.01 * len(data) + 400
