import os, sys, lzma, tempfile
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import reseq_b200 as rb
root = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
g = os.path.join(root, "tests", "golden")
tmp = tempfile.mkdtemp()
def unxz(n):
    d = os.path.join(tmp, n[:-3]); open(d, "wb").write(lzma.open(os.path.join(g, n)).read()); return d
prof = rb.Profile.load_flat(unxz("profile150.flat.xz"))
eng = rb.Engine(prof, 0)
ref = rb.Reference.load_fasta(os.path.join(g, "simref_small.fa"))
mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
if mode == "var":
    ref.load_variants(os.path.join(g, "simref_small_var.vcf"))
if mode == "window":
    os.environ["RSQ_SUR_WINDOW"] = "1"; os.environ["RSQ_BATCH_UNITS"] = "9"
eng.prepare(ref, seed=42, coverage=6.0)
rep = eng.simulate(); eng.download()
print(mode, rep.pairs, len(eng.output(0)))
eng.close()
