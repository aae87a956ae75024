"""Builds libreseq_b200.so (CUDA engine + C ABI) in-tree with nvcc for sm_100a.

    python -m reseq_b200.build        (also called by __graft_entry__.build())
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libreseq_b200.so")
CLI = os.path.join(HERE, "reseq-b200")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-fmad=false",                      # FP64 results must match the reference's non-contracted x86-64 arithmetic
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
    "-shared",
]


def sources():
    return [os.path.join(CSRC, "engine.cu")]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for root, _, files in os.walk(CSRC):
        for f in files:
            if os.path.getmtime(os.path.join(root, f)) > t:
                return True
    inc = os.path.join(os.path.dirname(HERE), "include", "reseq_b200.h")
    return os.path.getmtime(inc) > t


def build_cli():
    """reseq-b200: the `reseq illuminaPE` / `reseq seqToIllumina` command line over the C ABI."""
    cmd = ["g++", "-O2", "-std=c++17", "-pthread", "-o", CLI, os.path.join(CSRC, "cli_main.cpp"),
           "-L" + HERE, "-lreseq_b200", "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building reseq-b200")


def build(force=False, verbose=False):
    if not force and not needs_build() and os.path.exists(CLI):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + sources() + ["-lz"]   # zlib: gzip FASTQ/FASTA/BED either side of the path
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode:
        raise RuntimeError("nvcc failed building libreseq_b200.so")
    build_cli()
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
