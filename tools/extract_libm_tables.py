#!/usr/bin/env python3
"""Regenerates reseq_b200/csrc/libm_tables.inc from the image's libm.so.6 (glibc 2.39).

The tables are located by signature (first field invln2N = 0x1.71547652b82fep+7 for __exp_data,
ln2hi/ln2lo followed by A[0] = -0.5 for __pow_log_data), not by fixed offsets.
"""
import struct
import sys

LIBM = "/lib/x86_64-linux-gnu/libm.so.6"


def find_tables(data):
    exp_off = data.find(struct.pack("<d", float.fromhex("0x1.71547652b82fep+7")))
    sig = struct.pack("<d", float.fromhex("0x1.62e42fefa3800p-1"))
    j = data.find(sig)
    pow_off = -1
    while j != -1:
        if struct.unpack_from("<d", data, j + 16)[0] == -0.5:
            pow_off = j
        j = data.find(sig, j + 1)
    return exp_off, pow_off


def arr(name, vals, per=4):
    s = f"RSQ_TABLE_QUALIFIER unsigned long long {name}[{len(vals)}] = {{\n"
    for i in range(0, len(vals), per):
        s += "  " + ", ".join(f"0x{v:016x}ULL" for v in vals[i:i + per]) + ",\n"
    return s + "};\n"


def main(out):
    data = open(LIBM, "rb").read()
    eo, po = find_tables(data)
    hdr = struct.unpack_from("<22Q", data, eo)
    tab = struct.unpack_from("<256Q", data, eo + 22 * 8)
    plhdr = struct.unpack_from("<9Q", data, po)
    pltab = struct.unpack_from("<512Q", data, po + 72)
    assert struct.unpack("<d", struct.pack("<Q", tab[1]))[0] == 1.0
    with open(out) as f:
        head = f.read().split("#pragma once")[0]
    with open(out, "w") as f:
        f.write(head + "#pragma once\n" + arr("kExpHdr", hdr[:8]) + arr("kExpTab", tab) + arr("kPowLogHdr", plhdr) + arr("kPowLogTab", pltab))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "reseq_b200/csrc/libm_tables.inc")
