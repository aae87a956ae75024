"""Reader/writer for the RSQFLAT1 tagged-array container (profile tables, per-stage dumps).

Layout: magic "RSQFLAT1", then records {u32 name_len, name, u32 dtype, u64 count, payload};
dtype 0=u8 1=u32 2=u64 3=f64 4=i64.  Written by the engine's host side (rsq_profile_save_flat),
and - for validation only - by oracle/dump_tables.cpp.
"""
import struct

import numpy as np

_DTYPES = {0: np.uint8, 1: np.uint32, 2: np.uint64, 3: np.float64, 4: np.int64}
_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


def read_flat(path):
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] != b"RSQFLAT1":
        raise ValueError(f"{path}: not an RSQFLAT1 file")
    pos = 8
    while pos < len(data):
        (nl,) = struct.unpack_from("<I", data, pos)
        pos += 4
        name = data[pos:pos + nl].decode()
        pos += nl
        dtype, count = struct.unpack_from("<IQ", data, pos)
        pos += 12
        dt = np.dtype(_DTYPES[dtype])
        out[name] = np.frombuffer(data, dtype=dt, count=count, offset=pos).copy()
        pos += count * dt.itemsize
    return out


def write_flat(path, arrays):
    with open(path, "wb") as f:
        f.write(b"RSQFLAT1")
        for name, arr in arrays.items():
            arr = np.ascontiguousarray(arr)
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)))
            f.write(nb)
            f.write(struct.pack("<IQ", _CODES[arr.dtype], arr.size))
            f.write(arr.tobytes())
