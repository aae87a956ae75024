"""seqToIllumina input parsing (reseq_b200/csrc/em_input.hpp, the parallel host side of rsq_apply_error_model) against a line-by-line
restatement of the reference's parsing in Simulator::ApplyErrorsAndQualityToFastaInput (Simulator.cpp:2403-2470)."""
import gzip
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TWIN_DIR = os.path.join(ROOT, "tests", "host_twin")
CODE = {**{c: 0 for c in "Aa"}, **{c: 1 for c in "Cc"}, **{c: 2 for c in "Gg"}, **{c: 3 for c in "TtUu"}}


@pytest.fixture(scope="module")
def em_check(workdir):
    exe = os.path.join(workdir, "em_input_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(TWIN_DIR, "em_input_check.cpp"), "-lz"], check=True)
    return exe


def restate(text):
    """SeqAn FASTA reading + the header checks of the reference, record by record, in file order."""
    ids, seqs = [], []
    for line in text.split("\n"):
        if line.endswith("\r"):
            line = line[:-1]
        if line.startswith(">"):
            ids.append(line[1:])
            seqs.append("")
        elif seqs:
            seqs[-1] += line.replace(" ", "").replace("\t", "")
    if not ids:
        raise ValueError("does not contain any sequences")
    out = []
    for rid, seq in zip(ids, seqs):
        n = len(seq)
        if len(rid) <= 2 * n + 2:
            raise ValueError("Read description is too short")
        end = len(rid) - 2 * n - 3
        if rid[end + 1] != ";" or rid[end + 2 + n] != ";":
            raise ValueError("not separated by a semicolon from themselves")
        dom = rid[end + 2:end + 2 + n]
        rates = []
        for ch in rid[len(rid) - n:]:
            r = (ord(ch) - 33) & 0xFF
            if r > 86:
                r = (r + r - 86) & 0xFF
            rates.append(r)
        while end and rid[end] != " ":
            end -= 1
        if end == 0:
            raise ValueError("No sequence id found")
        if rid[end + 1] not in "12":
            raise ValueError("Template segment is")
        if rid[end + 2] != ";":
            raise ValueError("template segment and fragment length are not separated")
        fl = rid[end + 3:len(rid) - 2 * n - 2]
        if not fl.strip() or not fl.lstrip().lstrip("+-").isdigit() or fl != fl.rstrip():
            raise ValueError("is not a pure integer")
        out.append((rid[:end], int(rid[end + 1]) - 1, int(fl), "".join("ACGT"[CODE.get(c, 4) & 3] for c in seq),
                    "".join("ACGTN"[CODE.get(c, 4)] for c in dom), rates))
    return out


def run(em_check, path):
    res = subprocess.run([em_check, path], capture_output=True, text=True)
    if res.returncode:
        raise ValueError(res.stderr.strip())
    lines = res.stdout.split("\n")
    recs = []
    for line in lines[1:]:
        if not line:
            continue
        recs.append(line.split(" "))   # the id may contain spaces: compare() takes the fixed fields from the end
    return lines[0], recs


def compare(em_check, path, text):
    want = restate(text)
    head, got = run(em_check, path)
    assert head.split()[1] == str(len(want))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        n = len(w[3])
        fixed = g[len(g) - n - 4:] if n else g[-4:]
        rid = " ".join(g[:len(g) - len(fixed)])
        assert rid == w[0]
        assert int(fixed[0]) == w[1] and int(fixed[1]) == w[2] and fixed[2] == w[3] and fixed[3] == w[4]
        assert [int(x) for x in fixed[4:]] == w[5]


def test_golden_input(em_check, golden):
    compare(em_check, golden["em_in"], open(golden["em_in"]).read())


def test_line_ends_blanks_wrapped_sequences_and_compressed_rates(em_check, workdir):
    recs = [
        ">r1 extra words 1;300;ACGTN;!+~}z\nACG\nTN\n",                      # wrapped sequence, rates above 86 are stored compressed
        ">r2 2;7;acgu;!!!!\r\nAC GU\r\n",                                   # CRLF, blank inside the sequence, lower case + U
        "\n>r3 1;12;NNNNNNNN;IIIIIIII\nACGT\tACGT\n\n",                      # empty lines, tab
        ">r4 2;+5;AA;!!\nRY",                                               # IUPAC codes become A; no final newline; signed length
    ]
    text = "junk before the first record\n" + "".join(recs)
    path = os.path.join(workdir, "em_mixed.fa")
    open(path, "w", newline="").write(text)
    compare(em_check, path, text)
    gz = path + ".gz"
    with gzip.open(gz, "wt", newline="") as f:
        f.write(text)
    compare(em_check, gz, text)


def test_many_records_use_every_thread(em_check, workdir):
    import random
    rnd = random.Random(5)
    parts = []
    for i in range(30000):
        n = rnd.randint(1, 40)
        seq = "".join(rnd.choice("ACGT") for _ in range(n))
        parts.append(f">rec{i} {1 + i % 2};{n + rnd.randint(0, 50)};{seq};{''.join(chr(33 + rnd.randint(0, 90)) for _ in range(n))}\n{seq}\n")
    text = "".join(parts)
    path = os.path.join(workdir, "em_many.fa")
    open(path, "w").write(text)
    compare(em_check, path, text)


@pytest.mark.parametrize("bad,message", [
    (">x 1;10;ACGT;!!!\nACGT\n", "semicolon"),
    (">x1;10;ACGT;!!!!\nACGT\n", "No sequence id found"),
    (">x 3;10;ACGT;!!!!\nACGT\n", "Template segment is 3"),
    (">x 1:10;ACGT;!!!!\nACGT\n", "not separated by a semicolon"),
    (">x 1;1o;ACGT;!!!!\nACGT\n", "not a pure integer"),
    (">x\nACGT\n", "too short"),
    ("no records here\n", "does not contain any sequences"),
])
def test_malformed_records_are_reported_first_in_file_order(em_check, workdir, bad, message):
    good = "".join(f">g{i} 1;9;AC;!!\nAC\n" for i in range(20000))
    path = os.path.join(workdir, "em_bad.fa")
    later = ">later 9;1;A;!\nA\n"   # a second, different error further down must not win
    open(path, "w").write((good + bad + good + later) if bad.startswith(">") else bad)
    with pytest.raises(ValueError, match=message):
        run(em_check, path)
