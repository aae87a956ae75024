"""reseq_b200: B200-native engine for ReSeq's per-read simulation hot path (C ABI in include/reseq_b200.h).

The Python layer is a thin ctypes binding used by the tests, bench.py and the CLI shim; all work happens in
libreseq_b200.so (CUDA, sm_100a).  There is no CPU fallback: importing works anywhere, creating an Engine
requires a CUDA device.
"""
from .api import (Engine, Profile, Reference, SimOptions, SimReport, RsqError, group_unique_id, lib_path, load_library, shard_plan, simulate, simulate_multi)  # noqa: F401
