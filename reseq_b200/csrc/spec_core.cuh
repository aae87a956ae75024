// Speculative two-phase form of the per-read simulation hot path.
//
// The reference consumes ONE mt19937_64 stream per 1000-bp SimBlock strictly in order: a scan draw per (position,
// fragment length), then - for every fragment hit - the draws of its read pairs (Simulator.cpp:2249-2357, 634-721,
// 454-594, 294-452).  The raw stream is independent of who consumes it, and a read consumes a *predictable* number of
// draws unless an InDel is drawn.  That allows splitting the work so that reads of many blocks advance in lock step:
//
//   phase A  scan_window()   lane group per block: scans, evaluates hits/fragment counts exactly as the reference does and,
//                            for every read, *assumes* its consumption (plan_read: no InDel), copies the read's slice of
//                            the stream to HBM and continues scanning behind it.  Emits <= D new reads per round.
//   phase B  ReadMachine     one lane per read, 32 reads (of several blocks) in lock step: FillRead/FillReadPart as a state
//                            machine with three draw points per base; reports the consumption it really had.
//   verify   next scan_window() call: the reads up to and including the first one whose consumption differs from the
//                            assumption are final (its own start was exact).  The stream is brought to the state right
//                            behind that prefix (replay with the measured consumptions), which becomes the committed
//                            snapshot; everything emitted behind it is discarded and scanned again from there.
//
// Results are therefore bit-identical to the serial order by construction; a wrong assumption only costs a replay.
// The same templates run in tests/host_twin with a one-lane group (CPU), the product instantiates them for warps.
#pragma once
#include "sim_core.cuh"

namespace rsq {

constexpr uint32_t kSpecNone = 0xffffffffu;
constexpr uint32_t kSpecOverflow = 0xffffffffu;   // ReadJob::consumed when the slice (assumed + margin) was too short
constexpr uint32_t kSpecMaxDepth = 64;            // reads speculated per unit and round at most (three output slabs of 32 slots)
#ifndef RSQ_SLICE_WINDOW
#define RSQ_SLICE_WINDOW 0   // 1: the reads kernel streams its slices through bulk-copy windows in shared memory (measured slower: E. coli 69.4 ms against 59.5, 100 Mbp 1172 ms against 1006 - 144 bytes of shared memory per lane cost a block per SM and every eighth word a barrier wait); 0: through four prefetch registers
#endif
constexpr uint32_t kSpecWindowBytes = RSQ_SLICE_WINDOW ? 144 : 0;   // shared memory per lane of the reads kernel for its slice window (ReadMachineT::win)
constexpr uint32_t kSpecMargin = 32;              // extra stream words copied behind the assumed consumption
constexpr uint32_t kConvSlots = 20;               // bisulfite-converted fragment ends kept per unit (>= depth / 2 + 3: one per hit of a round + the committed one)
constexpr uint32_t kErrSpecOverflow = 64;         // error flag: a read needed more than assumed + margin draws

// ----------------------------------------------------------------------------------------------------------------
// mt19937_64 as a two-generation ring: look-ahead of up to 312 words without consuming them
// ----------------------------------------------------------------------------------------------------------------
struct MtRing {
	uint64_t *w;       // 2 * kMtN words, group-shared
	uint32_t cur;      // next word, [0, 624)
	uint32_t avail;    // generated words from cur on; (cur + avail) % 312 == 0
};

template<class G> RSQ_HD void ring_generate_words(const G &g, uint64_t *w, uint32_t end){
	MtRing r; r.w = w; r.cur = 0; r.avail = 0;
	const uint32_t dst = (end >= 2u * kMtN ? end - 2u * kMtN : end) ? kMtN : 0u;   // end is 0, 312 or 624
	const uint32_t src = dst ? 0u : kMtN;
	g.sync();
	for(uint32_t base = 0; base < kMtN; base += G::kSize){
		const uint32_t i = base + g.lane();
		if(i < kMtN){
			const uint64_t x = r.w[src + i];
			const uint64_t y = (i + 1u < kMtN) ? r.w[src + i + 1u] : r.w[dst];
			const uint64_t z = (i < kMtM) ? r.w[src + i + kMtM] : r.w[dst + i - kMtM];
			const uint64_t v = (x & 0xFFFFFFFF80000000ull) | (y & 0x7FFFFFFFull);
			r.w[dst + i] = z ^ (v >> 1) ^ ((v & 1ull) ? 0xB5026F5AA96619E9ull : 0ull);
		}
		g.sync();
	}
}
#if defined(__CUDACC__)
// One copy of the regeneration loop per kernel instead of one per place that consumes the stream, for the variant-aware scan only: inlined it
// was a third of that kernel's 382 KB (31 copies), and the kernel stalls on instruction fetches (no_instruction 60 % of its stall cycles).
// The plain scan regenerates every 2.4 trips of its filter loop - there the call costs more than the smaller code gains (E. coli 59.6 ms
// against 57.6 ms).
__device__ __noinline__ void ring_generate_words_warp(uint64_t *w, uint32_t end){ WarpGroup g; ring_generate_words(g, w, end); }
#endif
template<class G> RSQ_HD void ring_generate(const G &g, MtRing &r){
#if defined(__CUDA_ARCH__)
	if(G::kCompactCode){ ring_generate_words_warp(r.w, r.cur + r.avail); }
	else{ ring_generate_words(g, r.w, r.cur + r.avail); }
#else
	ring_generate_words(g, r.w, r.cur + r.avail);
#endif
	r.avail += kMtN;
}
template<class G> RSQ_HD void ring_ensure(const G &g, MtRing &r, uint32_t n){   // n <= 312
	if(r.avail < n){ ring_generate(g, r); }
}
RSQ_HD uint64_t ring_raw(const MtRing &r, uint32_t i){   // i < avail
	uint32_t p = r.cur + i;
	if(p >= 2u * kMtN){ p -= 2u * kMtN; }
	return r.w[p];
}
RSQ_HD void ring_advance(MtRing &r, uint32_t n){
	r.cur += n; if(r.cur >= 2u * kMtN){ r.cur -= 2u * kMtN; }
	r.avail -= n;
}
template<class G> RSQ_HD void ring_skip(const G &g, MtRing &r, uint32_t n){
	while(n){
		const uint32_t m = n < static_cast<uint32_t>(kMtN) ? n : kMtN;
		ring_ensure(g, r, m);
		ring_advance(r, m);
		n -= m;
	}
}
template<class G> RSQ_HD uint64_t ring_next(const G &g, MtRing &r){
	ring_ensure(g, r, 1);
	const uint64_t x = mt_temper(ring_raw(r, 0));
	ring_advance(r, 1);
	return x;
}
struct RingSource {                // the unit's stream as the scan holds it
	MtRing &r;
	template<class G> RSQ_HD double next(const G &g){ return canonical(ring_next(g, r)); }
};
// seed state in the first half; nothing generated yet
template<class G> RSQ_HD void ring_seed(const G &g, MtRing &r, uint64_t seed){
	g.sync();
	if(g.lane() == 0){
		uint64_t x = seed;
		r.w[0] = x;
		for(int i = 1; i < kMtN; ++i){ x = 6364136223846793005ull * (x ^ (x >> 62)) + static_cast<uint64_t>(i); r.w[i] = x; }
	}
	r.cur = kMtN; r.avail = 0;
	g.sync();
}

// ----------------------------------------------------------------------------------------------------------------
// Shared data of the two phases
// ----------------------------------------------------------------------------------------------------------------
// seqToIllumina input record (Simulator::ApplyErrorsAndQualityToFastaInput): sequence, per-base systematic errors, id text
struct EmRecord { uint64_t seq_off; uint32_t len; uint32_t seg; uint32_t fragment_length; uint32_t id_off; uint32_t id_len; uint32_t pad; };

struct ReadJob {
	uint32_t ref_id;
	uint32_t start_pos, end_pos;   // fragment [start, end) on the forward strand (0,0: adapter-only pair)
	uint32_t fragment_length;
	uint32_t block_id;             // id printed in the record name
	uint32_t flags;                // bit 0 segment, bit 1 strand, bit 2 seqToIllumina record (ref_id = record index), bit 3 fragment touches variants, bits 8-23 tile index, bits 24-30 allele
	uint64_t read_number;
	uint32_t assumed;              // draws the scan assumed this read consumes
	uint32_t consumed;             // draws it consumed (phase B)
	uint32_t rec_len;              // bytes of its FASTQ record (phase B)
	uint32_t slot;                 // output slot
	uint32_t conv_index;           // bisulfite runs / fragments touching variants: which staged fragment end this read starts from (kSpecNone: the reference itself)
	int32_t var_id;                // fragments touching variants: StartVariant (read on the forward strand) / EndVariant (read on the reverse strand) ...
	uint32_t var_pos;              // ... {id, posCurrentlyAt}
	uint32_t pad;
};

struct SpecHit {                   // where inside SimulateFromGivenBlock's / CreateReads' loop nest the stream stands
	uint32_t active, in_reads, fragment_length, n_chosen, chosen0, chosen1, ci, counts_left, strand;
	uint32_t pair_stage;           // reads of the current pair already handled (0..2)
	uint32_t tile;                 // tile drawn for the current pair
	uint32_t conv_slot, conv_next; // bisulfite runs / variants: slot of this hit's staged fragment ends / next slot to hand out
	// runs with variants: the chosen (allele, strand) whose reads are being emitted
	uint32_t allele, end_pos, slow; int32_t end_var; uint32_t end_var_pos;
};
struct SpecSnap {                  // resumable state of one unit's stream
	uint32_t pos, len;
	uint32_t finished, mt_off;
	SpecHit hit;
	int32_t cur_meth;
	uint32_t first_var, start_variant_pos;   // VariantBiasVarModifiers::first_variant_id_ / start_variant_pos_
	uint64_t read_number, scan_draws;
	uint64_t mt[kMtN];
};
struct SpecBlock {
	uint32_t done;
	uint32_t snap_bank, snap_idx;  // committed snapshot: snaps[unit][bank][idx] ...
	uint32_t pending_skip;         // ... plus this many stream words (the measured consumption of the read behind it)
	uint32_t n_jobs;               // reads emitted in the last round (speculative until verified)
	uint32_t fill;                 // verified records in cur_slab
	uint32_t cur_slab, next_slab;  // output slabs (32 slots) being filled: slots fill .. of cur_slab, then next_slab, then next2_slab
	uint32_t next2_slab, pad;      // (fill < 32 verified + up to kSpecMaxDepth = 64 speculative records reach into a third slab)
	uint32_t chain_head, chain_tail;
	uint32_t reads, rounds;
	unsigned long long bytes[2];
	unsigned long long scan_draws;
};

struct SpecCtx {
	uint32_t depth;                // D: capacity per unit and round (reads emitted beyond the verified ones, <= kSpecMaxDepth); array stride
	uint32_t run_depth;            // reads to emit per unit in this round (<= depth): grows when few units are left
	uint32_t scan_budget;          // scan draws per unit and round after which the round ends even with < run_depth reads (bounds stragglers)
	uint32_t map_depth;            // read slots per unit that k_spec_reads covers (run_depth, or depth when dense units may speculate deeper)
	float mean_reads;              // expected reads of a SimBlock; > 0: a unit whose reads come denser than that speculates proportionally deeper (up to depth),
	                               // so that all units get through their block in about the same number of rounds (0: run_depth for every unit)
	uint32_t words_per_job;        // K: capacity of a read's stream slice
	uint32_t margin;               // stream words copied behind the assumed consumption (kSpecMargin; tests shrink it to force the fallback)
	uint32_t n_units;              // blocks (+ the adapter-only pseudo block) of this batch
	SpecBlock *blocks;             // [n_units]
	SpecSnap *snaps;               // [n_units][2 banks][D + 1]: in front of every emitted read + behind the last one
	ReadJob *jobs;                 // [n_units * D]
	uint64_t *words;               // [n_units * D][K] tempered stream words of every speculated read (spec_slice)
	uint8_t *conv;                 // bisulfite runs / variants: [n_units][kConvSlots][2][kMaxOrgLen] staged (spliced, converted) forward / reverse fragment ends
	uint16_t *snap_chosen;         // variants: [n_units][2 banks][D + 1][chosen_stride] the chosen (allele, strand) ids of the hit a snapshot stands in
	uint32_t chosen_stride;        // 2 * num_alleles
	// output slots, handed out in slabs of 32
	unsigned char *slots; uint32_t slot_stride, id_cap, seq_off, qual_off;
	uint32_t n_slabs; uint32_t *next_slab; uint32_t *slab_next; uint32_t *slab_count;
	uint32_t *n_done;              // units that are through (monotonic; the host polls it between batches of rounds)
	unsigned long long *stat;      // [0] reads emitted, [1] reads verified (the host tunes the depth with their ratio)
	// adapter-only pseudo block (Simulator::SimulateAdapterOnlyPairs), unit index n_blocks when present
	uint32_t n_blocks; uint32_t adapter_only_pairs; uint64_t adapter_only_seed;
	// seqToIllumina: unit u = batch u of em_batch input records with its own stream (seed em_seeds[u]); no scan, one read per record
	const EmRecord *em_recs; uint32_t em_n, em_batch; const uint64_t *em_seeds;
	const uint8_t *em_seq, *em_sys; const char *em_ids;   // bases; (dominant error, rate) pairs; id text
};

// Stream slice of job gidx: words_per_job consecutive words (a multiple of 8: the reads kernel stages them in 64-byte pieces)
RSQ_HD uint64_t *spec_slice(const SpecCtx &sp, size_t gidx){ return sp.words + gidx * sp.words_per_job; }

RSQ_HD uint32_t spec_alloc_slab(const SpecCtx &sp){
#if defined(__CUDA_ARCH__)
	const uint32_t t = atomicAdd(sp.next_slab, 1u);
#else
	const uint32_t t = (*sp.next_slab)++;
#endif
	return t < sp.n_slabs ? t : kSpecNone;
}
RSQ_HD void spec_unit_done(const SpecCtx &sp, uint32_t *done_field){   // one lane
	*done_field = 1;
#if defined(__CUDA_ARCH__)
	atomicAdd(sp.n_done, 1u);
#else
	*sp.n_done += 1;
#endif
}
RSQ_HD void spec_flag(const SimCtx &c, uint32_t f){
#if defined(__CUDA_ARCH__)
	atomicOr(c.error_flag, f);
#else
	*c.error_flag |= f;
#endif
}

RSQ_HD uint32_t discrete_lookup(const Discrete &d, double p){   // std::lower_bound part of discrete_draw
	uint32_t lo = 0, len = d.n;
	while(len > 0){
		uint32_t half = len >> 1;
		if(d.cp[lo + half] < p){ lo += half + 1; len -= half + 1; }
		else{ len = half; }
	}
	return lo;
}
// GeneralRandomDistributions::ReadLength with the uniform variate supplied (Simulator.h:185-198)
RSQ_HD uint32_t read_length_from(const SimCtx &c, uint32_t seg, uint32_t fragment_length, double u){
	const double ins = (fragment_length < c.insert_to) ? static_cast<double>(c.insert_lengths[fragment_length]) : 0.0;
	const double random_value = mul_rn(u, ins);
	double counter = 0.0;
	uint32_t row_from = 0, row_n = 0;
	const uint64_t *vals = nullptr;
	if(fragment_length >= c.rlbf_from[seg] && fragment_length < c.rlbf_to[seg]){
		const uint32_t r = fragment_length - c.rlbf_from[seg];
		row_from = c.rlbf_row_from[seg][r];
		row_n = c.rlbf_row_off[seg][r + 1] - c.rlbf_row_off[seg][r];
		vals = c.rlbf_val[seg] + c.rlbf_row_off[seg][r];
	}
	uint32_t read_len = row_from + row_n;
	while(counter <= random_value && (read_len-- > row_from)){
		counter = add_rn(counter, static_cast<double>(vals[read_len - row_from]));
	}
	return read_len & 0xffffu;
}
RSQ_HD uint32_t adapter_length(const SimCtx &c, uint32_t seg, uint32_t adapter_id){
	uint32_t len = c.adapters[seg].off[adapter_id + 1] - c.adapters[seg].off[adapter_id];
	return len > c.max_org_len ? c.max_org_len : len;
}
RSQ_HD uint32_t fragment_org_len(const SimCtx &c, uint32_t seg, uint32_t fragment_length){
	uint32_t org_len = c.read_len_to[seg] + c.max_len_deletion;
	if(fragment_length < org_len){ org_len = fragment_length; }
	if(org_len > c.max_org_len){ org_len = c.max_org_len; }
	return org_len;
}

// ----------------------------------------------------------------------------------------------------------------
// Phase A
// ----------------------------------------------------------------------------------------------------------------
// Copies n stream words to the slice (word k of the read lives at dst[k]), consuming them; returns the last one.
template<class G> RSQ_HD uint64_t emit_words(const G &g, MtRing &r, uint64_t *dst, uint32_t &k, uint32_t cap, uint32_t n, bool consume = true){
	uint64_t last = 0;
	uint32_t kk = k;
	while(n){
		const uint32_t m = n < static_cast<uint32_t>(kMtN) ? n : kMtN;
		ring_ensure(g, r, m);
		for(uint32_t i = g.lane(); i < m; i += G::kSize){
			if(kk + i < cap){ dst[kk + i] = mt_temper(ring_raw(r, i)); }
		}
		last = mt_temper(ring_raw(r, m - 1u));
		if(consume){ ring_advance(r, m); }
		kk += m; n -= m;
	}
	if(consume){ k = kk; }
	return last;
}

// Emits the slice of one read under the no-InDel hypothesis (mirrors the draw order of Simulator::FillRead) and returns
// the assumed consumption.  Any deviation of the real read is caught by the verification, so this only has to be right
// in the common case.
struct RingPos { uint32_t cur, avail, value; };   // what an out-of-line consumer of the stream hands back (the ring itself stays in the caller's registers)
template<class G> RSQ_HD uint32_t plan_read_inline(const G &g, const SimCtx &c, MtRing &r, uint64_t *dst, uint32_t cap, uint32_t margin, uint32_t seg, uint32_t fragment_length,
                                                   uint32_t given_org_len){
	uint32_t k = 0;
	uint32_t read_length = c.read_len_from[seg];
	if(1 != c.read_len_count[seg]){
		const double u = canonical(emit_words(g, r, dst, k, cap, 1));
		read_length = read_length_from(c, seg, fragment_length, u);
	}
	if(read_length > c.max_read_len){ read_length = c.max_read_len; }
	const uint32_t org_len = given_org_len != kSpecNone ? given_org_len : (fragment_length ? fragment_org_len(c, seg, fragment_length) : 0u);
	const uint32_t n_part = read_length < org_len ? read_length : org_len;
	uint32_t adapter_id = 0;
	const AdapterSet &as = c.adapters[seg];
	if(0 == n_part && as.pick.n){ adapter_id = discrete_lookup(as.pick, canonical(emit_words(g, r, dst, k, cap, 1))); }
	emit_words(g, r, dst, k, cap, 1u + 3u * n_part);      // sequence quality + (InDel, quality, base call) per base
	if(n_part < read_length){
		if(0 == adapter_id && as.pick.n){ adapter_id = discrete_lookup(as.pick, canonical(emit_words(g, r, dst, k, cap, 1))); }
		uint32_t adapter_pos = 0;
		if(0 == n_part){
			const Discrete &sc = as.start_cut[adapter_id];
			adapter_pos = (sc.n ? discrete_lookup(sc, canonical(emit_words(g, r, dst, k, cap, 1))) : 0u) + as.start_cut_from[adapter_id];
		}
		const uint32_t alen = adapter_length(c, seg, adapter_id);
		uint32_t n_ad = alen > adapter_pos ? alen - adapter_pos : 0u;
		if(n_ad > read_length - n_part){ n_ad = read_length - n_part; }
		emit_words(g, r, dst, k, cap, 3u * n_ad);
		if(n_part + n_ad < read_length){
			const uint32_t rem = read_length - n_part - n_ad;
			uint32_t tail = c.polya_from;
			if(c.polya_pick.n){ tail += discrete_lookup(c.polya_pick, canonical(emit_words(g, r, dst, k, cap, 1))); }
			tail &= 0xffffu;
			const uint32_t n_tail = tail < rem ? tail : rem;
			emit_words(g, r, dst, k, cap, n_tail + (rem - n_tail) * (c.overrun_pick.n ? 2u : 1u));
		}
	}
	uint32_t km = k;
	if(margin){ emit_words(g, r, dst, km, cap, margin, false); }   // look-ahead only: the next consumer starts at k
	return k;
}
template<class G> RSQ_HD_COLD RingPos plan_read_cold(const G &g, const SimCtx &c, uint64_t *ring_w, uint32_t cur, uint32_t avail, uint64_t *dst, uint32_t cap, uint32_t margin, uint32_t seg,
                                                     uint32_t fragment_length, uint32_t given_org_len){
	MtRing r; r.w = ring_w; r.cur = cur; r.avail = avail;
	const uint32_t k = plan_read_inline(g, c, r, dst, cap, margin, seg, fragment_length, given_org_len);
	return RingPos{r.cur, r.avail, k};
}
template<class G> RSQ_HD uint32_t plan_read(const G &g, const SimCtx &c, MtRing &r, uint64_t *dst, uint32_t cap, uint32_t margin, uint32_t seg, uint32_t fragment_length,
                                            uint32_t given_org_len = kSpecNone){
	const RingPos p = plan_read_cold(g, c, r.w, r.cur, r.avail, dst, cap, margin, seg, fragment_length, given_org_len);
	r.cur = p.cur; r.avail = p.avail;
	return p.value;
}

template<class G> RSQ_HD uint32_t first_lane(const G &g, unsigned mask){
#if defined(__CUDA_ARCH__)
	(void)g; return __ffs(mask) - 1;
#else
	(void)g; (void)mask; return 0;
#endif
}

struct ScanVarState { uint32_t first_var, start_variant_pos; const uint16_t *chosen_live; uint16_t *chosen_out; uint32_t n_chosen; };
template<class G> RSQ_HD_COLD void save_snapshot_cold(const G &g, const uint64_t *ring_w, uint32_t ring_cur, SpecSnap &out, uint32_t pos, uint32_t len, bool finished, const SpecHit hit,
                                                      int32_t cur_meth, uint64_t read_number, uint64_t draws, const ScanVarState vs){
	MtRing ring; ring.w = const_cast<uint64_t *>(ring_w); ring.cur = ring_cur; ring.avail = 1;
	const uint32_t half = ring.cur >= static_cast<uint32_t>(kMtN) ? kMtN : 0u;
	g.sync();
	for(uint32_t i = g.lane(); i < static_cast<uint32_t>(kMtN); i += G::kSize){ out.mt[i] = ring.w[half + i]; }
	if(g.lane() == 0){
		out.pos = pos; out.len = len; out.finished = finished ? 1u : 0u; out.mt_off = ring.cur - half; out.hit = hit; out.cur_meth = cur_meth;
		out.read_number = read_number; out.scan_draws = draws; out.first_var = vs.first_var; out.start_variant_pos = vs.start_variant_pos;
	}
	if(vs.chosen_out){ for(uint32_t i = g.lane(); i < vs.n_chosen; i += G::kSize){ vs.chosen_out[i] = vs.chosen_live[i]; } }
	g.sync();
}
template<class G> RSQ_HD void save_snapshot(const G &g, MtRing &ring, SpecSnap &out, uint32_t pos, uint32_t len, bool finished, const SpecHit &hit,
                                            int32_t cur_meth, uint64_t read_number, uint64_t draws, const ScanVarState &vs = ScanVarState{0, 0, nullptr, nullptr, 0}){
	ring_ensure(g, ring, 1);
	save_snapshot_cold(g, ring.w, ring.cur, out, pos, len, finished, hit, cur_meth, read_number, draws, vs);
}

// CTConversion of one staged fragment end with the unit's stream, out of line (bisulfite runs only)
template<class G> RSQ_HD_COLD RingPos ct_conversion_ring(const G &g, const SimCtx &c, uint64_t *ring_w, uint32_t cur, uint32_t avail, uint8_t *read, uint32_t read_len, uint32_t seq_id,
                                                         uint32_t start_pos, int32_t cur_methylation_start, bool reversed){
	MtRing r; r.w = ring_w; r.cur = cur; r.avail = avail;
	RingSource rng{r};
	ct_conversion(g, c, rng, read, read_len, seq_id, start_pos, cur_methylation_start, reversed);
	return RingPos{r.cur, r.avail, 0};
}
template<class G> RSQ_HD_VCOLD RingPos ct_conversion_var_ring(const G &g, const SimCtx &c, uint64_t *ring_w, uint32_t cur, uint32_t avail, uint8_t *read, uint32_t read_len, uint32_t seq_id,
                                                             uint32_t start_pos, uint32_t allele, int32_t cur_methylation_start, bool reversed, const VariantView v, int32_t first_variant,
                                                             uint32_t first_variant_pos){
	MtRing r; r.w = ring_w; r.cur = cur; r.avail = avail;
	RingSource rng{r};
	ct_conversion_var(g, c, rng, read, read_len, seq_id, start_pos, allele, cur_methylation_start, reversed, v, first_variant, first_variant_pos);
	return RingPos{r.cur, r.avail, 0};
}
// GetOrgSeq with variants, out of line
template<class G> RSQ_HD_VCOLD void splice_fragment_ends_cold(const G &g, const SimCtx &c, const VariantView v, uint32_t ref_id, uint32_t strand, uint32_t pos, uint32_t first_var,
                                                             uint32_t start_variant_pos, uint32_t fl, const VarEval e, const VarGeom geo, uint8_t *frag_fwd, uint8_t *frag_rev, uint32_t which){
	splice_fragment_ends(g, c, v, ref_id, strand, pos, first_var, start_variant_pos, fl, e, geo, frag_fwd, frag_rev, which);
}

// Links a full (or final) slab into the unit's chain.  Lane 0 only.
RSQ_HD void spec_link_slab(const SpecCtx &sp, SpecBlock &blk, uint32_t slab, uint32_t count){
	sp.slab_count[slab] = count; sp.slab_next[slab] = kSpecNone;
	if(blk.chain_tail == kSpecNone){ blk.chain_head = slab; } else{ sp.slab_next[blk.chain_tail] = slab; }
	blk.chain_tail = slab;
}

// A unit's row of threshold high words (SimCtx::thr_hi).  The kernel stages it in shared memory; the load then names the shared state space
// (a generic load through a pointer that may be global or shared cannot become LDS and costs the fast loop its overlap).
struct ThrRow {
	const uint32_t *p;
#if defined(__CUDA_ARCH__)
	uint32_t shared_addr;   // device: always staged (k_spec_scan does it for every unit that scans)
	__device__ ThrRow(const uint32_t *row, bool) : p(row), shared_addr(static_cast<uint32_t>(__cvta_generic_to_shared(row))) {}
	__device__ __forceinline__ uint32_t at(uint32_t i) const { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(shared_addr + 4u * i)); return v; }
#else
	ThrRow(const uint32_t *row, bool) : p(row) {}
	uint32_t at(uint32_t i) const { return p[i]; }
#endif
};

// One round of one unit (SimBlock, or the adapter-only pseudo block):
//   1. verify the reads phase B just ran: the prefix up to and including the first read whose consumption differs
//      from the assumption is final (its own start was exact); commit its records,
//   2. restore the stream behind that prefix: the tentative end snapshot if every assumption held, else the snapshot
//      taken in front of the deviating read plus its measured consumption,
//   3. scan on and emit up to `depth` new reads, leaving a snapshot in front of each and one behind the last.
// Snapshots live in two banks of depth + 1 entries per unit; a round reads the committed one from one bank and writes the other.
template<bool kVar, class G>
RSQ_HD void scan_window(const G &g, const SimCtx &c, const SpecCtx &sp, const BlockDesc *descs, uint32_t first_desc, uint32_t u, uint64_t *ring_mem,
                        uint16_t *chosen_live = nullptr /* runs with variants: 2 * num_alleles entries of group-shared memory */,
                        const uint32_t *thr_hi_staged = nullptr /* this unit's row of c.thr_hi when the caller staged it in shared memory */){
	SpecBlock &blk = sp.blocks[u];
	if(blk.done){ return; }
	const uint32_t D = sp.depth;
	ReadJob *jobs = sp.jobs + static_cast<size_t>(u) * D;
	SpecSnap *unit_snaps = sp.snaps + static_cast<size_t>(u) * 2u * (D + 1u);
	const uint32_t n_prev = blk.n_jobs;
	const uint32_t reads_before = blk.reads;   // verified reads of this unit before this round (lane 0 adds this round's below)
	uint32_t bank = blk.snap_bank, idx = blk.snap_idx, pending_skip = blk.pending_skip;
	uint32_t fill = blk.fill, cur_slab = blk.cur_slab, next_slab = blk.next_slab, next2_slab = blk.next2_slab;
	g.sync();
	if(n_prev){
		uint32_t first_bad = kSpecNone, bad_consumed = 0;
		for(uint32_t base = 0; base < n_prev && first_bad == kSpecNone; base += G::kSize){
			const uint32_t j = base + g.lane();
			bool bad = false;
			uint32_t cons = 0;
			if(j < n_prev){
				cons = jobs[j].consumed;
				bad = cons != jobs[j].assumed;
			}
			const unsigned mask = g.ballot(bad);
			if(mask){
				const uint32_t fl = first_lane(g, mask);
				first_bad = base + fl;
#if defined(__CUDA_ARCH__)
				bad_consumed = __shfl_sync(0xffffffffu, cons, fl);
#else
				bad_consumed = cons;
#endif
			}
		}
		const bool all_ok = first_bad == kSpecNone;
		const uint32_t v = all_ok ? n_prev : first_bad + 1u;
		if(!all_ok && bad_consumed == kSpecOverflow){
			if(g.lane() == 0){ spec_flag(c, kErrSpecOverflow); spec_unit_done(sp, &blk.done); }
			return;
		}
		// commit the records of the verified prefix: slots fill .. fill + v - 1
		uint32_t b0 = 0, b1 = 0;
		for(uint32_t j = g.lane(); j < v; j += G::kSize){
			if((jobs[j].flags & 5u) == 1u){ b1 += jobs[j].rec_len; } else{ b0 += jobs[j].rec_len; }   // seqToIllumina records all go to the one output
		}
		b0 = g.reduce_add(b0); b1 = g.reduce_add(b1);
		fill += v;
		if(g.lane() == 0){
#if defined(__CUDA_ARCH__)
			atomicAdd(sp.stat, static_cast<unsigned long long>(n_prev)); atomicAdd(sp.stat + 1, static_cast<unsigned long long>(v));
#else
			sp.stat[0] += n_prev; sp.stat[1] += v;
#endif
			blk.bytes[0] += b0; blk.bytes[1] += b1; blk.reads += v;
		}
		while(fill >= 32u){   // full slabs join the unit's chain (up to two per round)
			if(g.lane() == 0){ spec_link_slab(sp, blk, cur_slab, 32u); }
			cur_slab = next_slab; next_slab = next2_slab; next2_slab = kSpecNone; fill -= 32u;
		}
		bank ^= 1u;   // the snapshots of the previous round are in the other bank
		if(all_ok){ idx = D; pending_skip = 0; }
		else{ idx = first_bad; pending_skip = bad_consumed; }
	}
	// ---- restore the committed state ----
	const SpecSnap &snap = unit_snaps[bank * (D + 1u) + idx];
	SpecSnap *out_snaps = unit_snaps + (bank ^ 1u) * (D + 1u);
	MtRing ring; ring.w = ring_mem;
	g.sync();
	for(uint32_t i = g.lane(); i < static_cast<uint32_t>(kMtN); i += G::kSize){ ring.w[i] = snap.mt[i]; }
	ring.cur = snap.mt_off; ring.avail = kMtN - snap.mt_off;
	uint32_t pos = snap.pos, len = snap.len;
	SpecHit hit = snap.hit;
	int32_t cur_meth = snap.cur_meth;
	uint64_t read_number = snap.read_number, draws = snap.scan_draws;
	bool finished = snap.finished != 0;
	// kVar: the instantiation for runs with variants, launched for exactly those (simulate_spec_batch); what it can never reach - seqToIllumina records, the
	// plain evaluation of a hit - is compiled out of it: it is bound by instruction fetches and every kilobyte of code counts
	const bool with_var = kVar && u < sp.n_blocks;
	uint32_t first_var = snap.first_var, start_variant_pos = snap.start_variant_pos;
	uint16_t *chosen_bank_out = nullptr;
	uint32_t next_var_pos = 0xffffffffu;   // position of variant first_var (0xffffffff: none left): the per-position checks below stay in registers
	auto refresh_next_var = [&](const VariantView &vv){ next_var_pos = first_var < vv.n ? vv.position[first_var] : 0xffffffffu; };
	if(with_var){
		const uint16_t *chosen_in = sp.snap_chosen + (static_cast<size_t>(u) * 2u * (D + 1u) + bank * (D + 1u) + idx) * sp.chosen_stride;
		chosen_bank_out = sp.snap_chosen + (static_cast<size_t>(u) * 2u * (D + 1u) + (bank ^ 1u) * (D + 1u)) * sp.chosen_stride;
		for(uint32_t i = g.lane(); i < hit.n_chosen && i < sp.chosen_stride; i += G::kSize){ chosen_live[i] = chosen_in[i]; }
		if(!(sp.em_recs != nullptr) && u < sp.n_blocks){ refresh_next_var(c.var.view(descs[first_desc + u].ref_id)); }
	}
	g.sync();
	if(pending_skip){
		// the snapshot stands in front of the read whose assumption failed: walk over it with its measured consumption
		ring_skip(g, ring, pending_skip);
		if(sp.em_recs){ ++pos; } else{ ++hit.pair_stage; }
	}
	if(finished){
		if(g.lane() == 0){
			if(fill){ spec_link_slab(sp, blk, cur_slab, fill); }
			blk.scan_draws = draws; blk.n_jobs = 0; blk.fill = 0; spec_unit_done(sp, &blk.done);
		}
		return;
	}
	const bool records = !kVar && sp.em_recs != nullptr;
	const bool adapter_only = !records && u >= sp.n_blocks;
	BlockDesc b{};
	if(!adapter_only && !records){ b = descs[first_desc + u]; }
	const uint32_t L = (adapter_only || records) ? 0u : c.seq_len[b.ref_id];
	const uint64_t off = (adapter_only || records) ? 0u : c.seq_off[b.ref_id];
	const uint32_t group = (adapter_only || records) ? 0u : c.coverage_group[b.ref_id];
	const double *thr = c.thr + static_cast<size_t>(group) * c.insert_to * 2;
	const uint64_t *thr_int = c.thr_int + static_cast<size_t>(group) * c.insert_to;
	const ThrRow thr_row(thr_hi_staged ? thr_hi_staged : c.thr_hi + static_cast<size_t>(group) * c.thr_hi_stride, thr_hi_staged != nullptr);
	const double *binom_p0 = c.binom_p0 + static_cast<size_t>(group) * c.insert_to;
	const uint32_t *gcp = c.gc_prefix + off + b.ref_id;
	const uint32_t insert_from = c.insert_from, insert_to = c.insert_to;
	uint32_t end = b.start_pos + 1000u;
	if(end > L){ end = L; }
	uint32_t emitted = 0;
	bool full = false, failed = false;
	const uint64_t draws_limit = draws + sp.scan_budget;
	// the adapter-only pairs are ONE serial stream however large the run is: always speculate as deep as the buffers allow
	uint32_t D_run = adapter_only ? D : (sp.run_depth < D ? sp.run_depth : D);
	if(!adapter_only && !records && sp.mean_reads > 0.0f && D_run < D && pos >= b.start_pos + 32u){
		// reads verified so far over the positions scanned so far, extrapolated to the block: hot blocks (GC / surrounding bias) hold up to twice the
		// reads of an average one and would need as many more rounds at the same depth - the tail of a small run
		const float est = static_cast<float>(reads_before) * static_cast<float>(end - b.start_pos) / static_cast<float>(pos - b.start_pos);
		const float scale = est / sp.mean_reads;
		if(scale > 1.0f){
			const uint32_t d = static_cast<uint32_t>(static_cast<float>(D_run) * scale + 0.999f);
			D_run = d < D ? d : D;
		}
	}
	if(records){
		// seqToIllumina (Simulator::ErrorModelOnlyThread): no scan - pos is the next input record of this batch, len the end of the batch
		while(pos < len && emitted < D_run && !failed){
			const EmRecord rec = sp.em_recs[pos];
			hit.tile = 0;
			if(1 < c.num_tiles){ hit.tile = discrete_lookup(c.tile_pick, canonical(ring_next(g, ring))); }
			const uint32_t p = fill + emitted;
			uint32_t slab = p < 32u ? cur_slab : (p < 64u ? next_slab : next2_slab);
			if(slab == kSpecNone){
				if(g.lane() == 0){ slab = spec_alloc_slab(sp); }
#if defined(__CUDA_ARCH__)
				slab = __shfl_sync(0xffffffffu, slab, 0);
#endif
				if(slab == kSpecNone){ failed = true; break; }
				if(p < 32u){ cur_slab = slab; } else if(p < 64u){ next_slab = slab; } else{ next2_slab = slab; }
			}
			save_snapshot(g, ring, out_snaps[emitted], pos, len, false, hit, cur_meth, read_number, draws);
			const size_t gidx = static_cast<size_t>(u) * D + emitted;
			uint64_t *dst = spec_slice(sp, gidx);
			const uint32_t org_len = rec.len < c.max_org_len ? rec.len : c.max_org_len;
			const uint32_t assumed = plan_read(g, c, ring, dst, sp.words_per_job, sp.margin, rec.seg, rec.fragment_length, org_len);
			if(g.lane() == 0){
				ReadJob j;
				j.ref_id = pos; j.start_pos = 0; j.end_pos = 0; j.fragment_length = rec.fragment_length; j.block_id = 0;
				j.flags = rec.seg | 4u | (hit.tile << 8);
				j.read_number = 0; j.assumed = assumed; j.consumed = kSpecOverflow; j.rec_len = 0; j.slot = slab * 32u + (p & 31u);
				j.conv_index = kSpecNone; j.var_id = 0; j.var_pos = 0; j.pad = 0;
				jobs[emitted] = j;
			}
			++emitted; ++pos;
		}
		finished = pos >= len;
		full = !finished;
	}
	while(!finished && !full && !failed){
		if(hit.active){
			if(hit.in_reads){
				// CreateReads: counts_left pairs of (second read, first read)
				while(hit.counts_left && !full && !failed){
					if(hit.pair_stage == 0u){
						if(emitted >= D_run){ full = true; break; }
						++read_number;
						hit.tile = 0;
						if(1 < c.num_tiles){ hit.tile = discrete_lookup(c.tile_pick, canonical(ring_next(g, ring))); }
					}
					while(hit.pair_stage < 2u){
						if(emitted >= D_run){ full = true; break; }
						const uint32_t seg = 1u - hit.pair_stage;
						const uint32_t p = fill + emitted;
						uint32_t slab = p < 32u ? cur_slab : (p < 64u ? next_slab : next2_slab);
						if(slab == kSpecNone){
							if(g.lane() == 0){ slab = spec_alloc_slab(sp); }
#if defined(__CUDA_ARCH__)
							slab = __shfl_sync(0xffffffffu, slab, 0);
#endif
							if(slab == kSpecNone){ failed = true; break; }
							if(p < 32u){ cur_slab = slab; } else if(p < 64u){ next_slab = slab; } else{ next2_slab = slab; }
						}
						save_snapshot(g, ring, out_snaps[emitted], pos, len, false, hit, cur_meth, read_number, draws,
						              ScanVarState{first_var, start_variant_pos, chosen_live, with_var ? chosen_bank_out + static_cast<size_t>(emitted) * sp.chosen_stride : nullptr, hit.n_chosen});
						const size_t gidx = static_cast<size_t>(u) * D + emitted;
						uint64_t *dst = spec_slice(sp, gidx);
						const uint32_t assumed = plan_read(g, c, ring, dst, sp.words_per_job, sp.margin, seg, hit.fragment_length);
						if(g.lane() == 0){
							ReadJob j;
							const bool reversed = seg != hit.strand;
							const bool read_slow = with_var && (hit.slow & (reversed ? 2u : 1u)) != 0u;
							const bool staged = !adapter_only && (c.meth_loaded || read_slow);
							j.ref_id = b.ref_id; j.start_pos = adapter_only ? 0u : pos; j.end_pos = adapter_only ? 0u : (with_var ? hit.end_pos : pos + hit.fragment_length);
							j.fragment_length = hit.fragment_length; j.block_id = b.block_id; j.flags = seg | (hit.strand << 1) | (hit.tile << 8);
							j.read_number = read_number; j.assumed = assumed; j.consumed = kSpecOverflow; j.rec_len = 0; j.slot = slab * 32u + (p & 31u);
							j.conv_index = staged ? static_cast<uint32_t>((static_cast<size_t>(u) * kConvSlots + hit.conv_slot) * 2u + (reversed ? 1u : 0u)) : kSpecNone;
							j.var_id = 0; j.var_pos = 0; j.pad = 0;
							if(with_var){
								j.flags |= (hit.allele << 24) | (read_slow ? 8u : 0u);
								j.var_id = reversed ? hit.end_var : static_cast<int32_t>(first_var); j.var_pos = reversed ? hit.end_var_pos : start_variant_pos;
							}
							jobs[emitted] = j;
						}
						++emitted; ++hit.pair_stage;
					}
					if(full || failed){ break; }
					hit.pair_stage = 0; --hit.counts_left;
				}
				if(full || failed){ break; }
				hit.in_reads = 0; ++hit.ci;
				if(adapter_only){ finished = true; break; }
			}
			if(with_var && hit.ci < hit.n_chosen){
				// one chosen (allele, strand) of a hit in a run with variants (Simulator.cpp:2311-2345)
				const VariantView v = c.var.view(b.ref_id);
				const uint32_t id = chosen_live[hit.ci];
				const uint32_t strand = id & 1u, fl = hit.fragment_length;
				const uint32_t allele = nth_possible_allele(v, c.var.num_alleles, first_var, start_variant_pos, pos, id / 2u);
				auto uniform = [&]() -> double { return canonical(ring_next(g, ring)); };
				VarEval e; VarGeom geo;
				bool runaway = false;
				if(eval_allele_hit(c, v, b.ref_id, pos, first_var, start_variant_pos, fl, allele, thr[2 * fl], uniform, e, geo, runaway)){
					if(runaway && g.lane() == 0){ spec_flag(c, kErrCountRunaway); }
					if(e.counts){
						hit.allele = allele; hit.end_pos = e.end_position; hit.slow = e.slow; hit.end_var = e.end_var; hit.end_var_pos = e.end_var_pos;
						if(e.slow || c.meth_loaded){
							hit.conv_slot = hit.conv_next; hit.conv_next = (hit.conv_next + 1u) % kConvSlots;
							uint8_t *frag0 = sp.conv + ((static_cast<size_t>(u) * kConvSlots + hit.conv_slot) * 2u) * kMaxOrgLen, *frag1 = frag0 + kMaxOrgLen;
							// the ends whose reads walk variants are spliced; with methylation both ends are staged for the conversion
							splice_fragment_ends_cold(g, c, v, b.ref_id, strand, pos, first_var, start_variant_pos, fl, e, geo, frag0, frag1, c.meth_loaded ? 3u : e.slow);
							if(c.meth_loaded){
								const int32_t end_var = e.end_var;
								for(uint32_t rev = 0; rev < 2; ++rev){
									const uint32_t seg = rev ? (strand ? 0u : 1u) : (strand ? 1u : 0u);
									const RingPos rp = ct_conversion_var_ring(g, c, ring.w, ring.cur, ring.avail, rev ? frag1 : frag0, fragment_org_len(c, seg, fl), b.ref_id, rev ? e.end_position : pos, allele,
									                                          cur_meth, rev != 0, v, rev ? end_var : static_cast<int32_t>(first_var), rev ? e.end_var_pos : start_variant_pos);
									ring.cur = rp.cur; ring.avail = rp.avail;
								}
								g.sync();
							}
						}
						hit.in_reads = 1; hit.counts_left = e.counts; hit.strand = strand; hit.pair_stage = 0; continue;
					}
				}
				++hit.ci;
				continue;
			}
			if(!kVar && hit.ci < hit.n_chosen){
				const uint32_t strand = (hit.ci ? hit.chosen1 : hit.chosen0) & 1u;
				const uint32_t fl = hit.fragment_length;
				const uint32_t cur_end = pos + fl;
				if(cur_end < L){
					const double thr0 = thr[2 * fl];
					const uint32_t gc_perc = percent_u32(gcp[cur_end] - gcp[pos], fl);
					const double rv = canonical(ring_next(g, ring));
					const double adjusted_random = add_rn(thr0, mul_rn(rv, sub_rn(1.0, thr0)));
					bool runaway = false;
					const uint32_t counts = fragment_counts(c, b.ref_id, fl, gc_perc, c.sur_start[off + pos], c.sur_end[off + cur_end - 1], adjusted_random, runaway);
					if(runaway && g.lane() == 0){ spec_flag(c, kErrCountRunaway); }
					if(counts){
						if(c.meth_loaded){
							// GetOrgSeq + CTConversion of both fragment ends (Simulator.cpp:2325-2340), once per (hit, strand), in front of its reads
							hit.conv_slot = hit.conv_next; hit.conv_next = (hit.conv_next + 1u) % kConvSlots;
							for(uint32_t rev = 0; rev < 2; ++rev){
								const uint32_t seg = rev ? (strand ? 0u : 1u) : (strand ? 1u : 0u);
								const uint32_t n = fragment_org_len(c, seg, fl);
								uint8_t *frag = sp.conv + ((static_cast<size_t>(u) * kConvSlots + hit.conv_slot) * 2u + rev) * kMaxOrgLen;
								g.sync();
								for(uint32_t i = g.lane(); i < n; i += G::kSize){
									frag[i] = rev ? static_cast<uint8_t>(3 - c.ref[off + cur_end - 1 - i]) : c.ref[off + pos + i];
								}
								g.sync();
								const RingPos rp = ct_conversion_ring(g, c, ring.w, ring.cur, ring.avail, frag, n, b.ref_id, rev ? cur_end : pos, cur_meth, rev != 0);
								ring.cur = rp.cur; ring.avail = rp.avail;
							}
							g.sync();
						}
						hit.in_reads = 1; hit.counts_left = counts; hit.strand = strand; hit.pair_stage = 0; continue;
					}
				}
				++hit.ci;
				continue;
			}
			hit.active = 0;
		}
		// ---- scan until the next candidate hit (tight loop: 32 (position, length) draws per trip) ----
		uint32_t fragment_length = 0;
		uint64_t x_hit = 0;
		bool found = false;
		while(true){
			if(len >= insert_to){
				len = insert_from;
				if(with_var && next_var_pos == pos){   // CheckForInsertedBasesToStartFrom: once more from the same position per further inserted base
					const VariantView vv = c.var.view(b.ref_id);
					next_start_pass(vv, pos, first_var, start_variant_pos);
					refresh_next_var(vv);
					if(start_variant_pos){ continue; }
				}
				++pos;
				if(pos >= end){ finished = true; break; }
				if(c.meth_loaded){   // SimulateFromGivenBlock: the first region that does not end in front of this position
					const uint32_t r0 = c.meth_off[b.ref_id], nr = c.meth_off[b.ref_id + 1] - r0;
					if(cur_meth >= 0 && static_cast<uint32_t>(cur_meth) < nr && c.meth_end[r0 + cur_meth] <= pos){ ++cur_meth; }
				}
			}
			const uint32_t lane = g.lane();
			{
				// First filter, four independent trips at once (hits are rare: ~1 in several thousand draws, so the loads and the tempering of
				// the four overlap instead of waiting on each other): only the high word of the tempered draw against the high word of the
				// threshold.  A draw that fails it cannot be a hit; the trip that holds a candidate goes through the exact test below.
				uint32_t n4 = insert_to - len;
				if(n4 > 4u * G::kSize){ n4 = 4u * G::kSize; }
				ring_ensure(g, ring, n4);
				bool any_hit = false;
				// all four loads are unconditional (a branch per trip would serialise them): behind the last length of a position the ring holds
				// older words and the threshold row its padding of 0xffffffff; what they say is masked out
#pragma unroll
				for(uint32_t q = 0; q < 4u; ++q){
					const uint32_t i = q * G::kSize + lane;
					const bool cand = mt_temper_hi(ring_raw(ring, i)) >= thr_row.at(len + i);
					any_hit |= cand & (i < n4);
				}
				if(g.ballot(any_hit) == 0){
					ring_advance(ring, n4); len += n4; draws += n4;
					if(draws >= draws_limit){ full = true; break; }
					continue;
				}
			}
			ring_ensure(g, ring, G::kSize);
			uint32_t n = insert_to - len;
			if(n > static_cast<uint32_t>(G::kSize)){ n = G::kSize; }
			uint64_t x = 0;
			bool is_hit = false;
			if(lane < n){
				x = mt_temper(ring_raw(ring, lane));
				is_hit = x >= thr_int[len + lane];
			}
			const unsigned mask = g.ballot(is_hit);
			if(mask == 0){
				ring_advance(ring, n); len += n; draws += n;
				if(draws >= draws_limit){ full = true; break; }
				continue;
			}
			const uint32_t first = first_lane(g, mask);
			fragment_length = len + first;
#if defined(__CUDA_ARCH__)
			x_hit = __shfl_sync(0xffffffffu, x, first);
#else
			x_hit = x;
#endif
			ring_advance(ring, first + 1u); len = fragment_length + 1u; draws += first + 1u;
			found = true;
			break;
		}
		if(!found){ break; }
		const double probability_chosen = canonical(x_hit);
		const double thr0 = thr[2 * fragment_length], thr1 = thr[2 * fragment_length + 1];
		if(!(probability_chosen >= thr1)){ continue; }
		if(with_var){
			// DrawNumberNonZeroStrands over the alleles possible at this start + ChooseAlleles (Simulator.cpp:2302-2309)
			const VariantView v = c.var.view(b.ref_id);
			const uint32_t n_possible = next_var_pos == pos ? count_possible_alleles(v, c.var.num_alleles, first_var, start_variant_pos, pos) : c.var.num_alleles;
			const uint32_t n_pow = 2u * c.var.num_alleles + 1u;
			const double pow_term = c.binom_pow[(static_cast<size_t>(group) * insert_to + fragment_length) * n_pow + 2u * n_possible];
			const uint32_t non_zero_strands = binomial_count(2u * n_possible, sub_rn(1.0, thr0), pow_term, probability_chosen);
			if(!non_zero_strands){ continue; }
			hit.active = 1; hit.in_reads = 0; hit.fragment_length = fragment_length; hit.ci = 0; hit.counts_left = 0; hit.strand = 0; hit.pair_stage = 0; hit.tile = 0;
			hit.chosen0 = 0; hit.chosen1 = 0;
			auto uniform = [&]() -> double { return canonical(ring_next(g, ring)); };
			g.sync();
			hit.n_chosen = choose_alleles(chosen_live, non_zero_strands, 2u * n_possible, uniform);
			g.sync();
			continue;
		}
		const uint32_t non_zero_strands = binomial_count(2, sub_rn(1.0, thr0), binom_p0[fragment_length], probability_chosen);
		if(!non_zero_strands){ continue; }
		hit.active = 1; hit.in_reads = 0; hit.fragment_length = fragment_length; hit.ci = 0; hit.counts_left = 0; hit.strand = 0; hit.pair_stage = 0; hit.tile = 0;
		if(non_zero_strands <= 1){
			const double rv = canonical(ring_next(g, ring));
			hit.chosen0 = static_cast<uint32_t>(mul_rn(rv, 2.0)) & 0xffffu; hit.chosen1 = 0; hit.n_chosen = 1;
		}
		else{
			hit.chosen0 = 0; hit.chosen1 = 1; hit.n_chosen = 2;
		}
	}
	if(failed){
		if(g.lane() == 0){ spec_flag(c, kErrArenaFull); spec_unit_done(sp, &blk.done); }
		return;
	}
	// ---- tentative snapshot behind the new reads ----
	save_snapshot(g, ring, out_snaps[D], pos, len, finished, hit, cur_meth, read_number, draws,
	              ScanVarState{first_var, start_variant_pos, chosen_live, with_var ? chosen_bank_out + static_cast<size_t>(D) * sp.chosen_stride : nullptr, hit.n_chosen});
	if(g.lane() == 0){
		blk.snap_bank = bank; blk.snap_idx = idx; blk.pending_skip = pending_skip;
		blk.n_jobs = emitted; blk.fill = fill; blk.cur_slab = cur_slab; blk.next_slab = next_slab; blk.next2_slab = next2_slab; blk.rounds += 1;
		if(0 == emitted && finished){
			// nothing left to verify: the state behind the last verified read is exact and the unit is through
			if(fill){ spec_link_slab(sp, blk, cur_slab, fill); }
			blk.scan_draws = draws; blk.fill = 0; spec_unit_done(sp, &blk.done);
		}
		else{
			if(0 == emitted){
				// scan budget used up without a read: the end snapshot has nothing unverified in front of it, so it is the committed one
				blk.snap_bank = bank ^ 1u; blk.snap_idx = D; blk.pending_skip = 0;
			}
		}
	}
}

// Initial snapshot of a unit: freshly seeded stream, scan at the block start (or, for the adapter-only pseudo block,
// inside CreateReads with all its pairs left).
RSQ_HD void spec_init_unit(const SimCtx &c, const SpecCtx &sp, const BlockDesc *descs, uint32_t first_desc, uint32_t u){
	SpecBlock &blk = sp.blocks[u];
	blk.done = 0; blk.snap_bank = 0; blk.snap_idx = 0; blk.pending_skip = 0; blk.n_jobs = 0; blk.fill = 0; blk.cur_slab = kSpecNone; blk.next_slab = kSpecNone; blk.next2_slab = kSpecNone; blk.pad = 0;
	blk.chain_head = kSpecNone; blk.chain_tail = kSpecNone;
	blk.reads = 0; blk.rounds = 0; blk.bytes[0] = 0; blk.bytes[1] = 0; blk.scan_draws = 0;
	SpecSnap &s = sp.snaps[static_cast<size_t>(u) * 2u * (sp.depth + 1u)];
	const bool records = sp.em_recs != nullptr;
	const bool adapter_only = !records && u >= sp.n_blocks;
	uint64_t x = records ? sp.em_seeds[u] : (adapter_only ? sp.adapter_only_seed : descs[first_desc + u].seed);
	s.mt[0] = x;
	for(int i = 1; i < kMtN; ++i){ x = 6364136223846793005ull * (x ^ (x >> 62)) + static_cast<uint64_t>(i); s.mt[i] = x; }
	s.mt_off = kMtN;   // the whole generation is consumed: the first ring_ensure produces generation 1
	s.finished = 0; s.read_number = 0; s.scan_draws = 0; s.cur_meth = (adapter_only || records) ? 0 : descs[first_desc + u].first_meth;
	s.first_var = (adapter_only || records) ? 0u : descs[first_desc + u].first_var; s.start_variant_pos = 0;
	SpecHit h{};
	if(records){
		const uint64_t first = static_cast<uint64_t>(u) * sp.em_batch, last = first + sp.em_batch;
		s.pos = static_cast<uint32_t>(first); s.len = static_cast<uint32_t>(last < sp.em_n ? last : sp.em_n);
		if(s.pos >= s.len){ spec_unit_done(sp, &blk.done); }
	}
	else if(adapter_only){
		s.pos = 0; s.len = 0;
		h.active = 1; h.in_reads = 1; h.fragment_length = 0; h.n_chosen = 1; h.counts_left = sp.adapter_only_pairs;
		if(0 == sp.adapter_only_pairs){ spec_unit_done(sp, &blk.done); }
	}
	else{
		s.pos = descs[first_desc + u].start_pos; s.len = c.insert_from;
		if(c.meth_loaded){
			const uint32_t rid = descs[first_desc + u].ref_id, r0 = c.meth_off[rid], nr = c.meth_off[rid + 1] - r0;
			if(s.cur_meth >= 0 && static_cast<uint32_t>(s.cur_meth) < nr && c.meth_end[r0 + s.cur_meth] <= s.pos){ ++s.cur_meth; }
		}
		const uint32_t L = c.seq_len[descs[first_desc + u].ref_id];
		if(s.pos >= L){ spec_unit_done(sp, &blk.done); }
	}
	s.hit = h;
}

// ----------------------------------------------------------------------------------------------------------------
// Phase B: Simulator::FillRead / FillReadPart / CreateReadId for ONE read per lane (Simulator.cpp:454-594, 294-452, 596-632)
// ----------------------------------------------------------------------------------------------------------------
enum : uint32_t { kPhFrag = 0, kPhAdapter = 1, kPhTail = 2, kPhOverrun = 3, kPhDone = 4 };

template<bool kVar> struct ReadMachineT {
	// stream slice
	const uint64_t *words; uint32_t k, kcap; uint32_t overflow;
	// Device: the slice streams from HBM through a two-piece window in shared memory, 8 words (64 bytes) per piece, filled by bulk asynchronous
	// copies (cp.async.bulk, one mbarrier per piece) that this lane issues one piece ahead of its own consumption - no registers wait on the loads.
	//   window layout at `win` (shared-space address, 144 bytes per lane): piece 0 | piece 1 | mbarrier 0 | mbarrier 1
	uint32_t win, win_phase, win_pending, kslice;
	uint64_t w0, w1, w2, w3;   // RSQ_SLICE_WINDOW == 0: words k .. k+3, loaded ahead of their use
	// output
	uint8_t *seq_out, *qual_out; char *id; int id_len, id_cap, cigar_len;
	// job
	uint32_t seg, tile, fragment_length, record_out;
	// original sequence of the current part: base = comp ? 3 - org[step * pos] : org[step * pos]; sys[2 * pos] / [2 * pos + 1]
	const uint8_t *org; int32_t org_step; uint32_t org_comp; const uint8_t *sys;
	uint32_t org_pos, org_len;
	// Simulator::ReadFillParameter
	uint32_t read_length, read_pos, previous_indel_type, indel_pos, base_call, gc_seq, seq_qual, qual, error_rate, num_errors;
	// FillReadPart locals
	uint32_t phase; char base_cigar, cigar_element; uint32_t cigar_element_length;
	uint32_t adapter_id, tail_left;
	// per step
	uint32_t ref_base, dom_error, indel;
	// fragments touching variants: GetSysErrorFromBlock cursor over the SimBlocks' SysErrorVariants instead of sys[]
	uint32_t var_slow, allele, var_ref_id, var_reversed; SysWalk walk; const uint8_t *walk_sys;
	RSQ_HD SysWalkCtx walk_ctx(const SimCtx &c) const {
		SysWalkCtx wc;
		wc.sys = (var_reversed ? c.sys_rev : c.sys_fwd) + 2 * c.seq_off[var_ref_id]; wc.errs = var_reversed ? c.var.errs_rev : c.var.errs_fwd;
		wc.block_first = c.var.block_first + c.var.block_first_off[var_ref_id]; wc.v = c.var.view(var_ref_id); wc.L = c.seq_len[var_ref_id]; wc.reverse = var_reversed;
		return wc;
	}

#if defined(__CUDA_ARCH__) && RSQ_SLICE_WINDOW
	__device__ __forceinline__ void win_issue(uint32_t chunk){   // words [8 * chunk, 8 * chunk + 8) into piece chunk & 1
		const uint32_t h = chunk & 1u, bar = win + 128u + 8u * h;
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 64;" :: "r"(bar) : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 64, [%2];"
		             :: "r"(win + 64u * h), "l"(words + 8u * chunk), "r"(bar) : "memory");
		win_pending |= 1u << h;
	}
	__device__ __forceinline__ void win_wait(uint32_t h){
		const uint32_t bar = win + 128u + 8u * h, parity = (win_phase >> h) & 1u;
		asm volatile("{\n.reg .pred p;\nRSQ_WIN_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra RSQ_WIN_DONE;\nbra RSQ_WIN_WAIT;\nRSQ_WIN_DONE:\n}"
		             :: "r"(bar), "r"(parity) : "memory");
		win_phase ^= 1u << h; win_pending &= ~(1u << h);
	}
#endif
	RSQ_HD void load_window(){
#if defined(__CUDA_ARCH__) && RSQ_SLICE_WINDOW
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(win + 128u) : "memory");
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(win + 136u) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		win_phase = 0; win_pending = 0;
		win_issue(0);
		if(8u < kslice){ win_issue(1); }
#elif defined(__CUDA_ARCH__)
		w0 = words[k]; w1 = words[k + 1u]; w2 = words[k + 2u]; w3 = words[k + 3u];
#endif
	}
	// a lane leaves no copy in flight behind it (the shared memory may be handed to another block)
	RSQ_HD void drain_window(){
#if defined(__CUDA_ARCH__) && RSQ_SLICE_WINDOW
		if(win_pending & 1u){ win_wait(0); }
		if(win_pending & 2u){ win_wait(1); }
#endif
	}
	RSQ_HD double next_u(){
		if(k >= kcap){ overflow = 1; return 0.5; }
#if defined(__CUDA_ARCH__) && RSQ_SLICE_WINDOW
		const uint32_t chunk = k >> 3, h = chunk & 1u, i = k & 7u;
		if(i == 0u){
			// piece h starts: the other piece was consumed to its last word (its values have been used), refill it with the chunk behind this one
			if(chunk && (chunk + 1u) * 8u < kslice){ win_issue(chunk + 1u); }
			win_wait(h);
		}
		uint64_t x;
		asm volatile("ld.shared.u64 %0, [%1];" : "=l"(x) : "r"(win + 64u * h + 8u * i) : "memory");
		++k;
		return canonical(x);
#elif defined(__CUDA_ARCH__)
		const uint64_t x = w0;
		w0 = w1; w1 = w2; w2 = w3;
		w3 = (k + 4u < kcap) ? words[k + 4u] : 0ull;
		++k;
		return canonical(x);
#else
		return canonical(words[k++]);
#endif
	}
	RSQ_HD uint32_t org_base(uint32_t p) const {
		const uint32_t b = org[static_cast<int64_t>(org_step) * static_cast<int64_t>(p)];
		return org_comp ? 3u - b : b;
	}
	RSQ_HD void cigar_append(char op, uint32_t count){
		SingleLane one;
		int n = put_uint(one, id, id_len + cigar_len, id_len + kCigarCap < id_cap ? id_len + kCigarCap : id_cap, count);
		n = put_char(one, id, n, id_len + kCigarCap < id_cap ? id_len + kCigarCap : id_cap, op);
		cigar_len = n - id_len;
	}
	RSQ_HD void stage_adapter(const SimCtx &c){
		const uint32_t off = c.adapters[seg].off[adapter_id];
		org = c.adapter_seq + off; org_step = 1; org_comp = 0; sys = c.adapter_sys + 2 * static_cast<size_t>(off);
		org_len = adapter_length(c, seg, adapter_id);
		if(c.adapters[seg].off[adapter_id + 1] - off > c.max_org_len){ spec_flag(c, kErrOrgOverflow); }
	}
	RSQ_HD void gc_and_error(const SimCtx &c, uint32_t n, uint32_t &mean_error_rate){
		uint32_t gc = 0, err = 0;
		if(kVar && var_slow){
			SysWalk pre = walk;
			for(uint32_t i = 0; i < n; ++i){
				const uint32_t b = org_base(i);
				gc += (b == 1 || b == 2) ? 1u : 0u;
				uint32_t e;
				if(!sysw_plain_step(walk_sys, pre, e)){ const SysWalkCtx wc = walk_ctx(c); e = sysw_next(wc, pre, allele); }
				err += e >> 8;
			}
		}
		else{
			for(uint32_t i = 0; i < n; ++i){
				const uint32_t b = org_base(i);
				gc += (b == 1 || b == 2) ? 1u : 0u;
				err += sys[2 * i + 1];
			}
		}
		gc_seq = percent_u16(gc, n);
		mean_error_rate = divide_u32(err, n);
	}

	// FillRead up to the sequence-quality draw; returns the mean systematic error rate (its second index)
	RSQ_HD uint32_t begin(const SimCtx &c, const SpecCtx &sp, const ReadJob &j, const uint64_t *slice, unsigned char *slot, void *lane_window){
		// only assumed + margin words of the slice were written by the scan
		words = slice; k = 0; kcap = j.assumed + sp.margin < sp.words_per_job ? j.assumed + sp.margin : sp.words_per_job; overflow = 0;
		kslice = sp.words_per_job;
#if defined(__CUDA_ARCH__) && RSQ_SLICE_WINDOW
		win = static_cast<uint32_t>(__cvta_generic_to_shared(lane_window));
#else
		(void)lane_window; win = 0; win_phase = 0; win_pending = 0;
#endif
		load_window();
		seg = j.flags & 1u; tile = (j.flags >> 8) & 0xffffu; fragment_length = j.fragment_length;
		allele = (j.flags >> 24) & 0x7fu; var_slow = 0; var_ref_id = 0; var_reversed = 0; walk = SysWalk{}; walk_sys = nullptr;
		const bool strand = (j.flags >> 1) & 1u;
		id = reinterpret_cast<char *>(slot + 16); id_cap = static_cast<int>(sp.id_cap); cigar_len = 0;
		seq_out = slot + sp.seq_off; qual_out = slot + sp.qual_off;
		read_pos = 0; previous_indel_type = 0; indel_pos = 0; base_call = 5; gc_seq = 0; qual = 1; error_rate = 0; num_errors = 0; seq_qual = 0;
		read_length = c.read_len_from[seg];
		if(1 != c.read_len_count[seg]){ read_length = read_length_from(c, seg, fragment_length, next_u()); }
		if(read_length > c.max_read_len){ read_length = c.max_read_len; spec_flag(c, kErrOrgOverflow); }
		const bool record_job = (j.flags & 4u) != 0u;
		record_out = record_job ? 1u : 0u;
		// CreateReadId up to the CIGAR (everything the read itself does not change)
		if(record_job){   // seqToIllumina: the input id, a blank, then CIGAR and error count
			SingleLane one;
			const EmRecord rec = sp.em_recs[j.ref_id];
			int n = put_str(one, id, 0, id_cap, sp.em_ids + rec.id_off, rec.id_len);
			n = put_char(one, id, n, id_cap, ' ');
			id_len = n;
		}
		else{
			SingleLane one;
			uint32_t print_start = 0, print_end = 0;
			if(fragment_length){
				if(strand){ print_start = j.end_pos; print_end = j.start_pos + 1; }
				else{ print_start = j.start_pos + 1; print_end = j.end_pos; }
			}
			int n = 0;
			n = put_str(one, id, n, id_cap, c.base_id, c.base_id_len);
			n = put_uint(one, id, n, id_cap, j.block_id);
			n = put_char(one, id, n, id_cap, '_');
			n = put_uint(one, id, n, id_cap, j.read_number);
			if(kVar && print_start && c.var.loaded && 1 < c.var.num_alleles){ n = put_str(one, id, n, id_cap, "_allele", 7); n = put_uint(one, id, n, id_cap, allele); }
			n = put_char(one, id, n, id_cap, ':');
			n = put_uint(one, id, n, id_cap, print_start);
			n = put_char(one, id, n, id_cap, ':');
			if(print_start){ n = put_str(one, id, n, id_cap, c.name_blob + c.name_off[j.ref_id], c.name_off[j.ref_id + 1] - c.name_off[j.ref_id]); }
			else{ n = put_str(one, id, n, id_cap, "Adapter", 7); }
			n = put_char(one, id, n, id_cap, ':');
			n = put_uint(one, id, n, id_cap, print_end);
			n = put_char(one, id, n, id_cap, ':');
			n = put_uint(one, id, n, id_cap, c.tile_names[tile]);
			n = put_str(one, id, n, id_cap, ":1337:1337 ", 11);
			id_len = n;
		}
		// GetOrgSeq without variants
		org_len = 0; org_pos = 0; org = c.ref; org_step = 1; org_comp = 0; sys = c.sys_fwd;
		if(fragment_length && !record_job){
			org_len = fragment_org_len(c, seg, fragment_length);
			if(c.read_len_to[seg] + c.max_len_deletion > c.max_org_len && fragment_length > c.max_org_len){ spec_flag(c, kErrOrgOverflow); }
			const uint64_t off = c.seq_off[j.ref_id];
			const uint32_t L = c.seq_len[j.ref_id];
			const bool reversed = (seg != static_cast<uint32_t>(strand));
			if(!reversed){ org = c.ref + off + j.start_pos; sys = c.sys_fwd + 2 * (off + j.start_pos); }
			else{ org = c.ref + off + j.end_pos - 1; org_step = -1; org_comp = 1; sys = c.sys_rev + 2 * (off + (L - j.end_pos)); }
			if(j.conv_index != kSpecNone){ org = sp.conv + static_cast<size_t>(j.conv_index) * kMaxOrgLen; org_step = 1; org_comp = 0; }   // staged end (bisulfite-converted and / or spliced)
			if(kVar && (j.flags & 8u)){   // CreateReads with variants (Simulator.cpp:680-689): where this read starts in the block chain of its strand
				var_slow = 1; var_ref_id = j.ref_id; var_reversed = reversed ? 1u : 0u;
				const SysWalkCtx wc = walk_ctx(c);
				walk_sys = wc.sys;
				walk = reversed ? sysw_reverse_start(wc, j.start_pos / 1000u, j.end_pos, j.var_id, j.var_pos) : sysw_forward_start(wc, j.start_pos / 1000u, j.start_pos, static_cast<uint32_t>(j.var_id), j.var_pos);
			}
		}
		if(record_job){
			const EmRecord rec = sp.em_recs[j.ref_id];
			org = sp.em_seq + rec.seq_off; org_step = 1; org_comp = 0; sys = sp.em_sys + 2 * rec.seq_off;
			org_len = rec.len < c.max_org_len ? rec.len : c.max_org_len;
			if(rec.len > c.max_org_len){ spec_flag(c, kErrOrgOverflow); }
		}
		adapter_id = 0; tail_left = 0;
		const uint32_t seq_length = read_length < org_len ? read_length : org_len;
		uint32_t mean_error_rate = 0;
		if(seq_length){ gc_and_error(c, seq_length, mean_error_rate); }
		else{
			if(c.adapters[seg].pick.n){ adapter_id = discrete_lookup(c.adapters[seg].pick, next_u()); }
			var_slow = 0;
			stage_adapter(c);
			gc_and_error(c, org_len, mean_error_rate);
			org_len = 0;   // FillReadPart over the (empty) fragment part
		}
		phase = kPhFrag; base_cigar = 'M'; cigar_element = 'M'; cigar_element_length = 0;
		return mean_error_rate;
	}

	// Everything between two draw points: part ends, adapter choice, tails; leaves the lane at a state that needs a draw, or done.
	RSQ_HD void settle(const SimCtx &c){
		while(true){
			if(overflow){ phase = kPhDone; return; }
			if(phase <= kPhAdapter){
				if(read_pos < read_length && org_pos < org_len){ return; }
				if(cigar_element_length){ cigar_append(cigar_element, cigar_element_length); }
				if(read_pos >= read_length){ phase = kPhDone; return; }
				if(phase == kPhFrag){
					const AdapterSet &as = c.adapters[seg];
					if(0 == adapter_id && as.pick.n){ adapter_id = discrete_lookup(as.pick, next_u()); }
					uint32_t adapter_pos = 0;
					if(0 == read_pos){
						const Discrete &sc = as.start_cut[adapter_id];
						adapter_pos = (sc.n ? discrete_lookup(sc, next_u()) : 0u) + as.start_cut_from[adapter_id];
					}
					stage_adapter(c);
					var_slow = 0;   // adapter bases take their systematic errors from adapter_sys_error_
					org_pos = adapter_pos;
					phase = kPhAdapter; base_cigar = 'S'; cigar_element = 'S'; cigar_element_length = 0;
				}
				else{
					cigar_append('H', read_length - read_pos);
					uint32_t tail = c.polya_from;
					if(c.polya_pick.n){ tail += discrete_lookup(c.polya_pick, next_u()); }
					tail_left = tail & 0xffffu;
					phase = kPhTail;
				}
			}
			else if(phase == kPhTail){
				if(tail_left && read_pos < read_length){ return; }
				phase = kPhOverrun;
			}
			else if(phase == kPhOverrun){
				if(read_pos < read_length){ return; }
				phase = kPhDone; return;
			}
			else{ return; }
		}
	}

	RSQ_HD uint32_t previous_quality(const SimCtx &c) const { return (qual_out[read_pos - 1] - c.phred_offset) & 0xffu; }

	// CreateReadId tail + record header; returns the record length
	RSQ_HD uint32_t finish(const SimCtx &c, unsigned char *slot){
		SingleLane one;
		if(cigar_len > kCigarCap){ spec_flag(c, kErrCigarOverflow); cigar_len = kCigarCap; }
		int n = id_len + cigar_len;
		n = put_str(one, id, n, id_cap, " E", 2);
		n = put_uint(one, id, n, id_cap, num_errors);
		if(n > id_cap){ spec_flag(c, kErrRecordTooLong); n = id_cap; }
		uint32_t *hdr = reinterpret_cast<uint32_t *>(slot);
		hdr[0] = static_cast<uint32_t>(n); hdr[1] = read_length; hdr[2] = record_out ? 0u : seg; hdr[3] = 0;
		return 1u + n + 1u + read_length + 3u + read_length + 1u;
	}
};

// The lock-step body shared by the device kernel and the host twin.  DrawFn(active, table, i0, i1, i2, i3, u, zero) -> value
// is LogArrayResult::Draw for every lane whose `active` is set (all lanes of a group call it together).
template<bool kVar, class DrawFn, class AnyFn>
RSQ_HD void run_read_machine(const SimCtx &c, const SpecCtx &sp, bool have_job, const ReadJob &job, const uint64_t *slice, unsigned char *slot,
                             DrawFn &&draw_fn, AnyFn &&any_fn, uint32_t &consumed, uint32_t &rec_len, void *lane_window = nullptr /* device: 144 bytes of shared memory of this lane */){
	ReadMachineT<kVar> m;
	m.phase = kPhDone; m.overflow = 0; m.k = 0; m.var_slow = 0; m.win_pending = 0;
	uint32_t mean_error_rate = 0;
	if(have_job){ mean_error_rate = m.begin(c, sp, job, slice, slot, lane_window); }
	bool zero = false;
	{
		const uint32_t tid = have_job ? c.tab.seq_quality(m.seg, m.tile) : 0u;
		const double u = have_job ? m.next_u() : 0.0;
		uint32_t sq = draw_fn(have_job, tid, have_job ? m.gc_seq : 0u, mean_error_rate, have_job ? m.fragment_length / 10 : 0u, 0u, u, zero);
		if(have_job){
			if(zero){ sq = table_most_likely(c.tab, tid); }
			m.seq_qual = sq & 0xffu;
		}
	}
	while(true){
		if(m.phase != kPhDone){ m.settle(c); }
		if(!any_fn(m.phase != kPhDone)){ break; }
		const bool part = m.phase <= kPhAdapter;
		// --- InDel draw ---
		uint32_t t1 = 0; double u1 = 0.0;
		if(part){
			m.ref_base = m.org_base(m.org_pos);
			u1 = m.next_u();
			t1 = c.tab.indel(m.previous_indel_type, m.base_call);
		}
		uint32_t indel = draw_fn(part, t1, m.indel_pos, m.read_pos, m.gc_seq, 0u, u1, zero);
		if(part && zero){ indel = 0; }
		// --- quality draw ---
		const bool q_part = part && indel != 1u;
		const bool q_any = q_part || m.phase == kPhTail || m.phase == kPhOverrun;
		uint32_t t2 = 0; double u2 = 0.0;
		if(part && indel == 0u){
			if(kVar && m.var_slow){
				uint32_t e;
				if(!sysw_plain_step(m.walk_sys, m.walk, e)){ const SysWalkCtx wc = m.walk_ctx(c); e = sysw_next(wc, m.walk, m.allele); }
				m.dom_error = e & 0xffu; m.error_rate = e >> 8;
			}
			else{
				m.dom_error = m.sys[2 * m.org_pos];
				m.error_rate = m.sys[2 * m.org_pos + 1];
			}
		}
		if(q_any){
			u2 = m.next_u();
			t2 = c.tab.quality(m.seg, m.tile, q_part ? m.ref_base : 0u);
		}
		uint32_t q = draw_fn(q_any, t2, m.seq_qual, m.qual, m.read_pos, m.error_rate, u2, zero);
		bool need_call = false;
		uint32_t t3 = 0; double u3 = 0.0;
		if(q_any){
			if(part){
				if(indel == 0u){
					if(zero){ q = m.read_pos ? m.previous_quality(c) : table_max_value(c.tab, t2); }
					m.qual = q & 0xffu;
					m.qual_out[m.read_pos] = static_cast<uint8_t>(m.qual + c.phred_offset);
					need_call = true;
					u3 = m.next_u();
					t3 = c.tab.base_call(m.seg, m.tile, m.ref_base, m.dom_error);
				}
				else{   // insertion
					if(zero){ q = m.qual; }
					m.qual_out[m.read_pos] = static_cast<uint8_t>(c.phred_offset + q);
					m.seq_out[m.read_pos] = static_cast<uint8_t>(indel - 2u);
					if('I' == m.cigar_element){ ++m.cigar_element_length; ++m.indel_pos; }
					else{
						m.cigar_append(m.cigar_element, m.cigar_element_length);
						m.cigar_element = 'I'; m.cigar_element_length = 1; m.indel_pos = 1; m.previous_indel_type = 0;
					}
					++m.num_errors; ++m.read_pos;
				}
			}
			else{   // poly-A tail / overrun bases behind the adapter
				if(zero){ q = m.previous_quality(c); }
				m.qual = q & 0xffu;
				uint32_t b = 0;
				if(m.phase == kPhOverrun){ if(c.overrun_pick.n){ b = discrete_lookup(c.overrun_pick, m.next_u()); } }
				else{ --m.tail_left; }
				m.qual_out[m.read_pos] = static_cast<uint8_t>(m.qual + c.phred_offset);
				m.seq_out[m.read_pos] = static_cast<uint8_t>(b);
				++m.read_pos;
			}
		}
		else if(part){   // deletion
			if(kVar && m.var_slow){ const SysWalkCtx wc = m.walk_ctx(c); m.error_rate = sysw_deletion(wc, m.walk); }
			else{ m.error_rate = m.sys[2 * m.org_pos + 1]; }
			if('D' == m.cigar_element){ ++m.cigar_element_length; ++m.indel_pos; }
			else{
				m.cigar_append(m.cigar_element, m.cigar_element_length);
				m.cigar_element = 'D'; m.cigar_element_length = 1; m.indel_pos = 1; m.previous_indel_type = 1;
			}
			++m.num_errors; ++m.org_pos;
		}
		// --- base-call draw ---
		uint32_t call = draw_fn(need_call, t3, m.qual, m.read_pos, m.num_errors, m.error_rate, u3, zero);
		if(need_call){
			if(zero){ call = m.ref_base; }
			m.base_call = call;
			m.seq_out[m.read_pos] = static_cast<uint8_t>(call);
			if(m.base_cigar == m.cigar_element){ ++m.cigar_element_length; }
			else{
				m.cigar_append(m.cigar_element, m.cigar_element_length);
				m.cigar_element = m.base_cigar; m.cigar_element_length = 1; m.indel_pos = 0; m.previous_indel_type = 0;
			}
			if(call != m.ref_base){ ++m.num_errors; }
			++m.read_pos; ++m.org_pos;
		}
	}
	if(have_job){
		m.drain_window();
		rec_len = m.finish(c, slot);
		consumed = m.overflow ? kSpecOverflow : m.k;
	}
}

}  // namespace rsq
