#!/usr/bin/env python3
"""First-light check on a GPU box: small golden simulation + seqToIllumina against the committed fixtures."""
import lzma
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import reseq_b200 as rb  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def unxz(name, tmp):
    dst = os.path.join(tmp, name[:-3])
    with lzma.open(os.path.join(G, name)) as f, open(dst, "wb") as o:
        o.write(f.read())
    return dst


def main():
    tmp = tempfile.mkdtemp()
    flat = unxz("profile150.flat.xz", tmp)
    prof = rb.Profile.load_flat(flat)
    ref = rb.Reference.load_fasta(os.path.join(G, "simref_small.fa"))
    t0 = time.time()
    eng = rb.Engine(prof, 0)
    print("engine create", time.time() - t0, flush=True)
    rep = eng.prepare(ref, seed=42, coverage=20.0)
    print("prepare", rep.as_dict(), flush=True)
    rep = eng.simulate()
    print("simulate", rep.as_dict(), flush=True)
    eng.download()
    ok = True
    for seg, name in ((0, "sim_small_seed42_R1.fq.xz"), (1, "sim_small_seed42_R2.fq.xz")):
        got = eng.output(seg)
        want = lzma.open(os.path.join(G, name)).read()
        same = got == want
        ok &= same
        print(f"segment {seg}: {len(got)} bytes, golden {len(want)} bytes, identical={same}", flush=True)
        if not same:
            open(os.path.join(ROOT, "gpurun_out", f"first_R{seg + 1}.fq"), "wb").write(got)
            for i, (a, b) in enumerate(zip(got.split(b"\n"), want.split(b"\n"))):
                if a != b:
                    print("first differing line", i, a[:200], b[:200])
                    break
    frags = unxz("em_frags.fa.xz", tmp)
    out = os.path.join(tmp, "em.fq")
    rep = eng.apply_error_model(frags, out, 7)
    got = open(out, "rb").read()
    want = lzma.open(os.path.join(G, "em_seed7.fq.xz")).read()
    print("error model", rep.as_dict(), "identical=", got == want, flush=True)
    ok &= got == want
    if got != want:
        for i, (a, b) in enumerate(zip(got.split(b"\n"), want.split(b"\n"))):
            if a != b:
                print("first differing line", i, a[:200], b[:200])
                break
    print("FIRST LIGHT", "OK" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    sys.exit(main())
