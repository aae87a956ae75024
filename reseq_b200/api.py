"""ctypes binding of include/reseq_b200.h, mirroring reseq::Simulator's public interface
(reference reseq/Simulator.h:456-458): Engine.simulate(...) ~ Simulator::Simulate,
Engine.apply_error_model(...) ~ Simulator::SimulateErrorModelOnly."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))


class RsqError(RuntimeError):
    pass


class SimOptions(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("coverage", C.c_double), ("num_read_pairs", C.c_uint64), ("ref_bias_model", C.c_int32),
                ("record_base_identifier", C.c_char_p), ("shard_index", C.c_uint32), ("shard_count", C.c_uint32),
                ("sys_error_file", C.c_char_p), ("ref_bias_file", C.c_char_p)]


class SimReport(C.Structure):
    _fields_ = [("total_pairs_aim", C.c_uint64), ("adapter_only_pairs", C.c_uint64), ("pairs", C.c_uint64), ("bytes", C.c_uint64 * 2),
                ("blocks", C.c_uint64), ("blocks_total", C.c_uint64), ("positions", C.c_uint64), ("scan_draws", C.c_uint64),
                ("bias_normalization", C.c_double), ("syserr_passes", C.c_uint32), ("kernel_launches", C.c_uint32),
                ("ms_upload", C.c_float), ("ms_bias", C.c_float), ("ms_syserr", C.c_float), ("ms_simulate", C.c_float),
                ("ms_gather", C.c_float), ("ms_download", C.c_float),
                ("spec_rounds", C.c_uint32), ("spec_depth", C.c_uint32),
                ("group_pairs", C.c_uint64), ("group_world", C.c_uint32), ("shard_first", C.c_uint32),
                ("batches", C.c_uint32), ("reserved", C.c_uint32), ("resident_bytes_per_base", C.c_double)]

    def as_dict(self):
        out = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            out[name] = list(v) if hasattr(v, "__len__") else v
        return out


_LIB = None

# every symbol include/reseq_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "rsq_last_error": (C.c_char_p, []),
    "rsq_device_count": (C.c_int, []),
    "rsq_profile_load": (C.c_void_p, [C.c_char_p, C.c_char_p]),
    "rsq_profile_load_flat": (C.c_void_p, [C.c_char_p]),
    "rsq_profile_save_flat": (C.c_int, [C.c_void_p, C.c_char_p]),
    "rsq_profile_free": (None, [C.c_void_p]),
    "rsq_profile_remove_indel_errors": (C.c_int, [C.c_void_p]),
    "rsq_profile_remove_substitution_errors": (C.c_int, [C.c_void_p]),
    "rsq_profile_change_error_rate": (C.c_int, [C.c_void_p, C.c_double]),
    "rsq_reference_load_fasta": (C.c_void_p, [C.c_char_p]),
    "rsq_reference_from_memory": (C.c_void_p, [C.c_uint32, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_uint64)]),
    "rsq_reference_load_methylation": (C.c_int, [C.c_void_p, C.c_char_p]),
    "rsq_reference_load_variants": (C.c_int, [C.c_void_p, C.c_char_p]),
    "rsq_reference_num_alleles": (C.c_uint32, [C.c_void_p]),
    "rsq_reference_num_variants": (C.c_uint64, [C.c_void_p, C.c_uint32]),
    "rsq_reference_variants": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                         C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint8), C.c_uint64]),
    "rsq_reference_total_size": (C.c_uint64, [C.c_void_p]),
    "rsq_reference_num_sequences": (C.c_uint32, [C.c_void_p]),
    "rsq_reference_free": (None, [C.c_void_p]),
    "rsq_engine_create": (C.c_void_p, [C.c_void_p, C.c_int]),
    "rsq_engine_destroy": (None, [C.c_void_p]),
    "rsq_engine_prepare": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(SimOptions), C.POINTER(SimReport)]),
    "rsq_engine_simulate": (C.c_int, [C.c_void_p, C.POINTER(SimReport)]),
    "rsq_engine_download": (C.c_int, [C.c_void_p, C.POINTER(SimReport)]),
    "rsq_engine_output": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "rsq_engine_write": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "rsq_simulate": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(SimOptions), C.c_int, C.c_char_p, C.c_char_p, C.POINTER(SimReport)]),
    "rsq_simulate_multi": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(SimOptions), C.c_int, C.POINTER(C.c_int), C.c_char_p, C.c_char_p, C.POINTER(SimReport)]),
    "rsq_shard_plan": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]),
    "rsq_group_unique_id": (C.c_int, [C.c_void_p, C.c_uint64]),
    "rsq_engine_join_group": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "rsq_engine_leave_group": (C.c_int, [C.c_void_p]),
    "rsq_create_systematic_error_profile": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_char_p]),
    "rsq_apply_error_model": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_uint64, C.POINTER(SimReport)]),
    "rsq_engine_fetch": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]),
}


def lib_path():
    """The in-tree library; RSQ_B200_LIB points A/B measurements at another build of the same sources (tools/build_variant.sh)."""
    return os.environ.get("RSQ_B200_LIB") or os.path.join(HERE, "libreseq_b200.so")


def load_library():
    """Loads the in-tree CUDA library; raises (no fallback) when it has not been built."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise RsqError(f"{path} is missing: build it with `python -m reseq_b200.build` (there is no CPU fallback)")
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def _err(lib):
    return RsqError(lib.rsq_last_error().decode(errors="replace"))


class Profile:
    def __init__(self, handle):
        self._h = handle

    @classmethod
    def load_flat(cls, path):
        lib = load_library()
        h = lib.rsq_profile_load_flat(os.fsencode(path))
        if not h:
            raise _err(lib)
        return cls(h)

    @classmethod
    def load(cls, stats_path, ipf_path=None):
        """DataStats::Load + ProbabilityEstimates::Load/PrepareResult on X.reseq (+ X.reseq.ipf)."""
        lib = load_library()
        h = lib.rsq_profile_load(os.fsencode(stats_path), os.fsencode(ipf_path or stats_path + ".ipf"))
        if not h:
            raise _err(lib)
        return cls(h)

    def remove_indel_errors(self):
        """ProbabilityEstimates::RemoveInDelErrors (--noInDelErrors)."""
        if load_library().rsq_profile_remove_indel_errors(self._h):
            raise _err(load_library())

    def remove_substitution_errors(self):
        """ProbabilityEstimates::RemoveSubstitutionErrors (--noSubstitutionErrors)."""
        if load_library().rsq_profile_remove_substitution_errors(self._h):
            raise _err(load_library())

    def change_error_rate(self, error_multiplier):
        """ProbabilityEstimates::ChangeErrorRate (--errorMutliplier)."""
        if load_library().rsq_profile_change_error_rate(self._h, error_multiplier):
            raise _err(load_library())

    def save_flat(self, path):
        lib = load_library()
        if lib.rsq_profile_save_flat(self._h, os.fsencode(path)):
            raise _err(lib)

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.rsq_profile_free(self._h)
            self._h = None


class Reference:
    def __init__(self, handle):
        self._h = handle

    @classmethod
    def load_fasta(cls, path):
        lib = load_library()
        h = lib.rsq_reference_load_fasta(os.fsencode(path))
        if not h:
            raise _err(lib)
        return cls(h)

    @classmethod
    def from_memory(cls, ids, seqs):
        """ids: list of str; seqs: list of bytes-like ASCII sequences (host buffers)."""
        lib = load_library()
        n = len(ids)
        id_arr = (C.c_char_p * n)(*[i.encode() for i in ids])
        keep = [bytes(s) if not isinstance(s, bytes) else s for s in seqs]
        seq_arr = (C.c_char_p * n)(*keep)
        len_arr = (C.c_uint64 * n)(*[len(s) for s in keep])
        h = lib.rsq_reference_from_memory(n, id_arr, seq_arr, len_arr)
        if not h:
            raise _err(lib)
        return cls(h)

    def load_methylation(self, bed_path):
        """Reference::PrepareMethylationFile/ReadMethylation (--methylation)."""
        lib = load_library()
        if lib.rsq_reference_load_methylation(self._h, os.fsencode(bed_path)):
            raise _err(lib)

    def load_variants(self, vcf_path):
        """Reference::PrepareVariantFile/ReadFirstVariants/ReadVariants (-V/--vcfSim): loads and validates the whole VCF."""
        lib = load_library()
        if lib.rsq_reference_load_variants(self._h, os.fsencode(vcf_path)):
            raise _err(lib)

    @property
    def num_alleles(self):
        return load_library().rsq_reference_num_alleles(self._h)

    def variants(self, seq):
        """Reference::Variants(seq) as a list of (position, replacement bases, allele bit set as a Python int)."""
        lib = load_library()
        n = lib.rsq_reference_num_variants(self._h, seq)
        cap_bases = 1 << 20
        while True:
            pos, off = (C.c_uint32 * max(n, 1))(), (C.c_uint32 * (n + 1))()
            lo, hi = (C.c_uint64 * max(n, 1))(), (C.c_uint64 * max(n, 1))()
            bases = (C.c_uint8 * cap_bases)()
            if lib.rsq_reference_variants(self._h, seq, n, pos, off, lo, hi, bases, cap_bases) == 0:
                break
            if cap_bases >= 1 << 30:
                raise _err(lib)
            cap_bases <<= 4
        return [(pos[k], "".join("ACGT"[b] for b in bases[off[k]:off[k + 1]]), lo[k] | (hi[k] << 64)) for k in range(n)]

    @property
    def total_size(self):
        return load_library().rsq_reference_total_size(self._h)

    @property
    def num_sequences(self):
        return load_library().rsq_reference_num_sequences(self._h)

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.rsq_reference_free(self._h)
            self._h = None


class Engine:
    """One CUDA device holding the profile tables; prepare() + simulate() + download() = Simulator::Simulate."""

    def __init__(self, profile, device=0):
        lib = load_library()
        self._lib = lib
        self._profile = profile
        self._h = lib.rsq_engine_create(profile._h, device)
        if not self._h:
            raise _err(lib)
        self.report = SimReport()

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rsq_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def prepare(self, reference, seed, coverage=0.0, num_read_pairs=0, ref_bias_model=1, record_base_identifier=None,
                shard_index=0, shard_count=1, sys_error_file=None, ref_bias_file=None):
        opt = SimOptions(seed, coverage, num_read_pairs, ref_bias_model,
                         record_base_identifier.encode() if record_base_identifier else None, shard_index, shard_count,
                         os.fsencode(sys_error_file) if sys_error_file else None,
                         os.fsencode(ref_bias_file) if ref_bias_file else None)
        if self._lib.rsq_engine_prepare(self._h, reference._h, C.byref(opt), C.byref(self.report)):
            raise _err(self._lib)
        return self.report

    def simulate(self):
        if self._lib.rsq_engine_simulate(self._h, C.byref(self.report)):
            raise _err(self._lib)
        return self.report

    def join_group(self, unique_id, rank, world):
        """This engine becomes rank `rank` of `world` engines of one run (NCCL over NVLink): prepare() then takes shard rank/world."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        if self._lib.rsq_engine_join_group(self._h, buf, rank, world):
            raise _err(self._lib)

    def leave_group(self):
        if self._lib.rsq_engine_leave_group(self._h):
            raise _err(self._lib)

    def download(self):
        if self._lib.rsq_engine_download(self._h, C.byref(self.report)):
            raise _err(self._lib)
        return self.report

    def output(self, segment):
        """FASTQ text of first (0) / second (1) reads as bytes (copied out of the engine's pinned buffer)."""
        p = C.c_void_p()
        n = C.c_uint64()
        if self._lib.rsq_engine_output(self._h, segment, C.byref(p), C.byref(n)):
            raise _err(self._lib)
        return C.string_at(p, n.value) if n.value else b""

    def write(self, first_reads_path, second_reads_path):
        if self._lib.rsq_engine_write(self._h, os.fsencode(first_reads_path), os.fsencode(second_reads_path)):
            raise _err(self._lib)

    def create_systematic_error_profile(self, reference, seed, fastq_out):
        """Simulator::CreateSystematicErrorProfile (--writeSysError)."""
        if self._lib.rsq_create_systematic_error_profile(self._h, reference._h, seed, os.fsencode(fastq_out)):
            raise _err(self._lib)

    def apply_error_model(self, fasta_in, fastq_out, seed):
        rep = SimReport()
        if self._lib.rsq_apply_error_model(self._h, os.fsencode(fasta_in), os.fsencode(fastq_out), seed, C.byref(rep)):
            raise _err(self._lib)
        return rep

    def fetch(self, name, dtype="uint8"):
        import numpy as np
        n = C.c_uint64()
        if self._lib.rsq_engine_fetch(self._h, name.encode(), None, 0, C.byref(n)):
            raise _err(self._lib)
        buf = np.empty(n.value, dtype=np.uint8)
        if self._lib.rsq_engine_fetch(self._h, name.encode(), buf.ctypes.data_as(C.c_void_p), n.value, C.byref(n)):
            raise _err(self._lib)
        return buf.view(dtype)


def group_unique_id():
    """128 bytes that rank 0 hands to every engine of a multi-GPU group (ncclGetUniqueId)."""
    lib = load_library()
    buf = C.create_string_buffer(128)
    if lib.rsq_group_unique_id(buf, 128):
        raise _err(lib)
    return buf.raw


def shard_plan(profile, reference, shard_count):
    """rsq_shard_plan: the SimBlock boundaries of shard_count shards, [b0 = 0, b1, ..., b_count] (host only)."""
    lib = load_library()
    out = (C.c_uint32 * (shard_count + 1))()
    if lib.rsq_shard_plan(profile._h, reference._h, shard_count, out):
        raise _err(lib)
    return list(out)


def simulate_multi(profile, reference, first_reads_path, second_reads_path, seed, n_gpus, coverage=0.0, num_read_pairs=0, ref_bias_model=1,
                   record_base_identifier=None, devices=None):
    """rsq_simulate_multi: the drop-in call on n_gpus devices of this box (one engine + one host thread per GPU)."""
    lib = load_library()
    opt = SimOptions(seed, coverage, num_read_pairs, ref_bias_model,
                     record_base_identifier.encode() if record_base_identifier else None, 0, 1, None, None)
    rep = SimReport()
    dev = (C.c_int * n_gpus)(*devices) if devices else None
    if lib.rsq_simulate_multi(profile._h, reference._h, C.byref(opt), n_gpus, dev, os.fsencode(first_reads_path), os.fsencode(second_reads_path), C.byref(rep)):
        raise _err(lib)
    return rep


def simulate(profile, reference, first_reads_path, second_reads_path, seed, coverage=0.0, num_read_pairs=0, ref_bias_model=1,
             record_base_identifier=None, device=0):
    """Drop-in for Simulator::Simulate(R1, R2, ref, stats, estimates, threads, seed, num_read_pairs, coverage, ...)."""
    lib = load_library()
    opt = SimOptions(seed, coverage, num_read_pairs, ref_bias_model,
                     record_base_identifier.encode() if record_base_identifier else None, 0, 1, None, None)
    rep = SimReport()
    if lib.rsq_simulate(profile._h, reference._h, C.byref(opt), device, os.fsencode(first_reads_path), os.fsencode(second_reads_path), C.byref(rep)):
        raise _err(lib)
    return rep
