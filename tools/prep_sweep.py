import os, sys, json, tempfile
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import bench, reseq_b200 as rb
tmp = tempfile.mkdtemp()
prof = rb.Profile.load_flat(bench.unxz(bench.PROFILE + ".flat.xz", tmp))
seq = bench.workload_sequence().encode()
eng = rb.Engine(prof, 0)
ref = rb.Reference.from_memory(["ecoli_sized synthetic"], [seq])
for cfg in sys.argv[1:]:
    c, w = cfg.split(",")
    if c == "auto":
        os.environ.pop("RSQ_SYS_CHUNK", None); os.environ.pop("RSQ_SYS_WARMUP", None)
    else:
        os.environ["RSQ_SYS_CHUNK"], os.environ["RSQ_SYS_WARMUP"] = c, w
    best = None
    for _ in range(3):
        rep = eng.prepare(ref, seed=bench.SEED, coverage=bench.COVERAGE).as_dict()
        if best is None or rep["ms_syserr"] < best["ms_syserr"]:
            best = rep
    print(cfg, "ms_syserr", round(best["ms_syserr"], 2), "ms_bias", round(best["ms_bias"], 2), "passes", best["syserr_passes"], flush=True)
