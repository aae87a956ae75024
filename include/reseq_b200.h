/* reseq_b200 -- C ABI of the B200 engine for ReSeq's per-read simulation hot path.
 *
 * The reference (schmeing/ReSeq, C++) has no FFI; its seam is the Simulator class
 * (reseq/Simulator.h:456-458).  Every entry point below names the reference interface it stands in for, so a
 * maintainer can bind it from reseq/main.cpp.  Plain pointers and sizes only; functions
 * return 0 on success (or a handle / NULL), never throw, and leave a message for rsq_last_error().
 * There is no CPU fallback: engine creation fails when no CUDA device is usable.
 */
#ifndef RESEQ_B200_H
#define RESEQ_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rsq_profile rsq_profile;     /* DataStats + ProbabilityEstimates state the simulation reads */
typedef struct rsq_reference rsq_reference; /* reseq::Reference: ids + Dna5 sequences */
typedef struct rsq_engine rsq_engine;       /* one CUDA device: tables, reference, systematic errors, output arena */

/* printErr-style diagnostics of the last failing call on this thread (reportingUtils.hpp:262-273) */
const char *rsq_last_error(void);
int rsq_device_count(void);

/* --- profile -------------------------------------------------------------------------------------
 * rsq_profile_load: DataStats::Load + PrepareProcessing (DataStats.cpp:1280-1328) and
 * ProbabilityEstimates::Load + PrepareResult (ProbabilityEstimates.cpp:961-1045) for a converged profile;
 * stats_path = X.reseq, ipf_path = X.reseq.ipf.
 * rsq_profile_load_flat / rsq_profile_save_flat: the same state as a binary cache (RSQFLAT1). */
rsq_profile *rsq_profile_load(const char *stats_path, const char *ipf_path);
rsq_profile *rsq_profile_load_flat(const char *flat_path);
int rsq_profile_save_flat(const rsq_profile *profile, const char *flat_path);
void rsq_profile_free(rsq_profile *profile);
/* ProbabilityEstimates::RemoveInDelErrors / RemoveSubstitutionErrors / ChangeErrorRate (ProbabilityEstimates.h:1516-1549),
 * i.e. the CLI switches --noInDelErrors, --noSubstitutionErrors, --errorMutliplier; apply before rsq_engine_create. */
int rsq_profile_remove_indel_errors(rsq_profile *profile);
int rsq_profile_remove_substitution_errors(rsq_profile *profile);
int rsq_profile_change_error_rate(rsq_profile *profile, double error_multiplier);

/* --- reference -----------------------------------------------------------------------------------
 * Reference::ReadFasta (Reference.cpp:758-811): IUPAC text -> Dna5 (ACGT/acgt/U -> 0..3, anything else N). */
rsq_reference *rsq_reference_load_fasta(const char *fasta_path);
rsq_reference *rsq_reference_from_memory(uint32_t n_seqs, const char *const *ids, const char *const *bases, const uint64_t *lengths);
/* Reference::PrepareMethylationFile + ReadMethylation (Reference.cpp:1132-1322), `--methylation <bed>`: extended bedGraph
 * "<sequence> <start> <end> <methylation>"; reads simulated from this reference get bisulfite C->T conversions. */
int rsq_reference_load_methylation(rsq_reference *ref, const char *bed_path);
/* Reference::PrepareVariantFile + ReadFirstVariants + ReadVariants + InsertVariant (Reference.cpp:96-113, 126-426, 1005-1078;
 * Reference.h:115-139), `-V/--vcfSim <vcf>`: contigs checked against the reference, genotype columns -> allele bit sets, records
 * split into single-position variants sorted deletion/substitution/insertion.  The whole file is read (the reference pages it in per
 * sequence, Simulator.cpp:938, 1278); a file the reference rejects anywhere is rejected here with the same diagnostics.
 * rsq_engine_prepare re-validates the REF bases that fell on an N of the unprocessed reference after ReplaceN (the reference
 * loads variants against the N-replaced sequence) and rsq_simulate / rsq_simulate_multi then simulate reads from the alleles
 * (`_allele<a>` read ids, Simulator.cpp:1700-2150).  Not supported, refused with a message: `--readSysError` together with -V. */
int rsq_reference_load_variants(rsq_reference *ref, const char *vcf_path);
uint32_t rsq_reference_num_alleles(const rsq_reference *ref);      /* Reference::NumAlleles (1 without variants) */
uint64_t rsq_reference_num_variants(const rsq_reference *ref, uint32_t seq);   /* Reference::Variants(seq).size() */
/* Reference::Variants(seq), flattened: position[k], replacement bases (codes 0..3) bases[bases_off[k] .. bases_off[k+1]),
 * allele bits 0-63 / 64-127.  bases_off needs capacity + 1 entries. */
int rsq_reference_variants(const rsq_reference *ref, uint32_t seq, uint64_t capacity, uint32_t *position, uint32_t *bases_off,
                           uint64_t *allele_lo, uint64_t *allele_hi, uint8_t *bases, uint64_t bases_capacity);
uint64_t rsq_reference_total_size(const rsq_reference *ref);       /* Reference::TotalSize */
uint32_t rsq_reference_num_sequences(const rsq_reference *ref);    /* Reference::NumberSequences */
void rsq_reference_free(rsq_reference *ref);

/* --- simulation ----------------------------------------------------------------------------------
 * Arguments of Simulator::Simulate (Simulator.h:457) that this revision supports. */
typedef struct rsq_sim_options {
	uint64_t seed;                 /* uintSeed seed */
	double coverage;               /* 0 = DataStats::CorrectedCoverage() */
	uint64_t num_read_pairs;       /* 0 = derive from coverage */
	int32_t ref_bias_model;        /* RefSeqBiasSimulation: 0 kKeep, 1 kNo, 2 kDraw, 3 kFile (ref_bias_file) */
	const char *record_base_identifier; /* NULL/"" = "ReseqRead" */
	/* sharding (one engine per GPU): simulate forward blocks [shard_index*n/shard_count, ...) of the run;
	 * shard_count 0 or 1 = whole run.  Block seeds and systematic errors are identical in every shard. */
	uint32_t shard_index;
	uint32_t shard_count;
	/* `sys_error_file` of Simulator::Simulate (--readSysError): FASTQ written by rsq_create_systematic_error_profile /
	 * `--writeSysError` (per sequence a "reverse" and a "forward" record: seq = dominant error, qual = error rate + 33,
	 * Simulator.cpp:326-335, 2562-2576).  NULL/"" = draw the systematic errors from the master stream. */
	const char *sys_error_file;
	const char *ref_bias_file;     /* `--refBiasFile`: one "<sequence id> <bias>" line per reference sequence (ref_bias_model 3) */
} rsq_sim_options;

typedef struct rsq_sim_report {
	uint64_t total_pairs_aim;      /* total_pairs_ after removing adapter-only pairs */
	uint64_t adapter_only_pairs;   /* num_adapter_only_pairs_ */
	uint64_t pairs;                /* read pairs written by this engine (shard) */
	uint64_t bytes[2];             /* FASTQ bytes of first / second reads */
	uint64_t blocks;               /* SimBlocks simulated by this engine */
	uint64_t blocks_total;         /* SimBlocks of the run (incl. the look-ahead blocks the reference never simulates) */
	uint64_t positions;            /* reference positions scanned */
	uint64_t scan_draws;           /* mt19937_64 draws consumed by the (position, fragment length) scan */
	double bias_normalization;
	uint32_t syserr_passes;        /* speculative passes of the systematic-error chains */
	uint32_t kernel_launches;      /* launches of this library's kernels during prepare + simulate */
	/* device time in milliseconds (CUDA events on the engine's stream) */
	float ms_upload, ms_bias, ms_syserr, ms_simulate, ms_gather, ms_download;
	/* speculative two-phase simulation (0/0 when the serial kernel ran): rounds of scan + read kernels, reads emitted per unit and round */
	uint32_t spec_rounds, spec_depth;
	/* multi-GPU group (rsq_engine_join_group / rsq_simulate_multi): read pairs of the whole run (all-reduced over NCCL; = pairs for a single engine),
	 * engines in the group, first SimBlock of this engine's shard */
	uint64_t group_pairs;
	uint32_t group_world, shard_first;
	/* memory plan of rsq_engine_simulate: batches the shard's blocks were simulated in, and the bytes per reference base that stay resident in HBM
	 * through the simulation (bases, G/C prefix counts, systematic errors of both strands) */
	uint32_t batches, reserved;
	double resident_bytes_per_base;
} rsq_sim_report;

rsq_engine *rsq_engine_create(const rsq_profile *profile, int device);
void rsq_engine_destroy(rsq_engine *engine);

/* --- multi-GPU ------------------------------------------------------------------------------------
 * The reference runs `-j N` worker threads over one shared block queue (Simulator.cpp:2384-2401, 2830-2836); here one engine per GPU takes a
 * contiguous range of SimBlocks.  Engines joined into a group (NCCL over NVLink, libnccl.so.2 bound at run time) split the prologue as well:
 * each uploads and prepares only the sequences its blocks lie in, the per-(sequence, fragment length) bias sums of CalculateBiasNormalization
 * (FragmentDistributionStats.cpp:3527-3533 threads the same list) are computed by the sequence's owner and all-reduced, and the read-pair counts
 * are all-reduced behind the data path (rsq_sim_report.group_pairs).  shard_index / shard_count of rsq_sim_options are then rank / world.
 * One process per GPU (torchrun, mpirun): rank 0 calls rsq_group_unique_id, ships the 128 bytes to the other ranks by any means, every rank calls
 * rsq_engine_join_group on its engine.  One process for all GPUs of a box: rsq_simulate_multi. */
/* The block ranges the engines take (host only, no GPU needed): boundaries[k] .. boundaries[k + 1] are the SimBlocks of shard k of shard_count
 * for this profile (its longest insert decides which sequences are simulated and how many look-ahead blocks end the run) and reference;
 * boundaries needs shard_count + 1 entries.  The even split, with a boundary moved onto a sequence start lying within 5 % of a shard's size.
 * rsq_engine_prepare uses the same function; rsq_sim_report.shard_first / blocks report the range an engine took. */
int rsq_shard_plan(const rsq_profile *profile, const rsq_reference *ref, uint32_t shard_count, uint32_t *boundaries);
int rsq_group_unique_id(void *id_out, uint64_t capacity /* >= 128 */);
int rsq_engine_join_group(rsq_engine *engine, const void *unique_id, int rank, int world);
int rsq_engine_leave_group(rsq_engine *engine);

/* Simulator::Simulate prologue (Simulator.cpp:2687-2826): ReplaceN, UpdateRefSeqBias, pair counts,
 * CalculateBiasNormalization, adapter + genome systematic errors, block seeds.  Copies `ref`; uploads to HBM. */
int rsq_engine_prepare(rsq_engine *engine, const rsq_reference *ref, const rsq_sim_options *opt, rsq_sim_report *report);
/* SimulationThread / SimulateFromGivenBlock / CreateReads over this shard's blocks (Simulator.cpp:2249-2401).
 * FASTQ text ends up device-resident, ordered like the reference's 1-thread run. */
int rsq_engine_simulate(rsq_engine *engine, rsq_sim_report *report);
/* Copies the FASTQ text of segment 0 (first reads) / 1 (second reads) to engine-owned pinned host memory. */
int rsq_engine_download(rsq_engine *engine, rsq_sim_report *report);
int rsq_engine_output(const rsq_engine *engine, int segment, const char **data, uint64_t *bytes);
/* Output()/Flush() (Simulator.cpp:114-230): append the downloaded text to the two files. */
int rsq_engine_write(const rsq_engine *engine, const char *first_reads_path, const char *second_reads_path);

/* Drop-in for `bool Simulator::Simulate(R1, R2, ref, stats, estimates, ...)`: prepare + simulate + download +
 * write on device `device`; on failure the output files are removed like the reference does (Simulator.cpp:2888-2893). */
int rsq_simulate(const rsq_profile *profile, const rsq_reference *ref, const rsq_sim_options *opt, int device,
                 const char *first_reads_path, const char *second_reads_path, rsq_sim_report *report);

/* The same on n_gpus devices of one box (devices: their CUDA ordinals, NULL = 0 .. n_gpus-1): one engine and one host thread per GPU, joined
 * into a group (see above); every engine streams its shard's FASTQ to disk while it simulates, shard 0 into the two files themselves, the
 * others into hidden files next to them that are appended in shard order at the end - byte for byte the files of the single-GPU call.
 * `report` holds the sums (pairs, bytes, blocks, launches) and the maxima over the engines (device times). */
int rsq_simulate_multi(const rsq_profile *profile, const rsq_reference *ref, const rsq_sim_options *opt, int n_gpus, const int *devices,
                       const char *first_reads_path, const char *second_reads_path, rsq_sim_report *report);

/* Drop-in for `bool Simulator::CreateSystematicErrorProfile(out, ref, stats, estimates, seed)` (Simulator.cpp:2597-2653,
 * `--writeSysError`): per sequence the systematic errors of the whole reverse strand, then of the whole forward strand. */
int rsq_create_systematic_error_profile(rsq_engine *engine, const rsq_reference *ref, uint64_t seed, const char *fastq_out_path);

/* Drop-in for `bool Simulator::SimulateErrorModelOnly(out, in, stats, estimates, threads, seed)`
 * (Simulator.cpp:2900-3014): FASTA records "<id> <1|2>;<fraglen>;<dom-err>;<err-rate>" -> FASTQ. */
int rsq_apply_error_model(rsq_engine *engine, const char *fasta_in_path, const char *fastq_out_path, uint64_t seed, rsq_sim_report *report);

/* Stage introspection for parity tests: copies a named device/host array ("sys_fwd", "sys_rev", "adapter_sys",
 * "blocks" (BlockDesc records, 32 bytes each), "spec_blocks" (per-unit counters of the last speculative batch, 80 bytes each), "thresholds", "sur_start", "sur_end", "reference") into dst (up to capacity bytes). */
int rsq_engine_fetch(const rsq_engine *engine, const char *name, void *dst, uint64_t capacity, uint64_t *bytes);

#ifdef __cplusplus
}
#endif
#endif
