// Variant-aware building blocks of the simulation path, written for both sides like core.cuh:
// the same source runs in the host test harness (tests/host_twin/variant_core_check.cpp) and in device code.
//
//   VariantView                      the flattened Reference::variants_ of one reference sequence (variants.hpp: FlatVariants)
//   splice_reference                 Reference::ReferenceSequence, variant overload (reference Reference.cpp:498-567): the first
//                                    `frag_length` bases of allele `allele` reading forward from `start_pos`, or the reverse
//                                    complement reading backwards from `start_pos` (exclusive), starting inside a variant's
//                                    replacement when `first_variant_pos` is set - the a8 half of Simulator::GetOrgSeq
//                                    (Simulator.cpp:1909-1914)
//   choose_alleles                   Simulator::ChooseAlleles / DrawNAlleles / SelectAllele / ReverseSelection (Simulator.cpp:1341-1397)
//                                    for any number of (allele, strand) ids; the kernels' built-in form is the 2-id case
//   sys_error_with_variants          Simulator::GetSysErrorFromBlock + IncrementBlockPos (Simulator.cpp:232-292): the systematic error of
//                                    the next base of allele `allele`, walking the per-block SysErrorVariant lists (Simulator.h:91-106)
//   allele_fragment                  what the scan needs for a hit of allele a at (start position, inserted start base, fragment length):
//                                    GC percent, start and end surrounding, end position - PrepareBiasModForCurrentStartPos/FragmentLength,
//                                    GetGCPercent and end_pos_shift_ of the reference (Simulator.cpp:1399-1873) as lookups in the
//                                    materialised allele sequence (variants.hpp: AlleleSequence)
//   allele_fragment_counts           the tail of FragmentDistributionStats::GetFragmentCounts with an allele count
//                                    (FragmentDistributionStats.cpp:3602-3626, 900-907): NB(mean / alleles, Dispersion(mean) / alleles)
//
// Further down: what the kernels use per hit and per base (VarCtx, start passes over inserted bases, possible alleles, allele walkers,
// allele_hit, the SysWalk cursor = GetSysErrorFromBlock with the reference's quirks).  k_simulate / k_spec_scan<true> / k_spec_reads<true>
// consume them (sim_core.cuh: eval_allele_hit, splice_fragment_ends, simulate_block_var; spec_core.cuh: scan_window<true>, ReadMachineT<true>).
#pragma once
#include "bias_core.cuh"

namespace rsq {

struct VariantView {
	const uint32_t *position;    // [n] sorted; same position: deletion, substitution, insertions by length
	const uint32_t *bases_off;   // [n + 1]
	const uint8_t *bases;        // replacement bases, codes 0..3
	const uint64_t *allele_lo, *allele_hi;
	uint32_t n;
	RSQ_HD bool in_allele(uint32_t var, uint32_t allele) const { return ((allele < 64 ? allele_lo[var] : allele_hi[var]) >> (allele & 63u)) & 1u; }   // Variant::InAllele
	RSQ_HD uint32_t length(uint32_t var) const { return bases_off[var + 1] - bases_off[var]; }
	RSQ_HD uint8_t base(uint32_t var, uint32_t k) const { return bases[bases_off[var] + k]; }
};

// out: codes 0..3 (a Dna5 'N' of the unprocessed reference becomes A like SeqAn's Dna5 -> Dna conversion; ReplaceN leaves none).
// first_variant / first_variant_pos: the pair {id, posCurrentlyAt} of VariantBiasVarModifiers::StartVariant / EndVariant.
// Returns the number of bases written (frag_length unless the sequence ends first).
RSQ_HD uint32_t splice_reference(uint8_t *out, const uint8_t *seq, const VariantView &vars, uint32_t start_pos, uint32_t frag_length, bool reversed,
                                 int32_t first_variant, uint32_t first_variant_pos, uint32_t allele){
	uint32_t len = 0;   // length(insert_string); the reference lets it overshoot by a replacement and cuts afterwards - here writes are clipped
	auto put = [&](uint8_t b){ if(len < frag_length){ out[len] = b & 3u; } ++len; };
	uint32_t cur_start = start_pos;
	int32_t cur_var = first_variant;
	if(reversed){
		auto put_ref_rc = [&](uint64_t from, uint32_t to){ for(uint64_t p = to; p > from; --p){ put(3u - (seq[p - 1] & 3u)); } };   // ReverseComplementorDna(infix(ref, from, to))
		if(first_variant_pos){
			for(uint32_t k = first_variant_pos; k > 0; --k){ put(3u - vars.base(cur_var, k - 1)); }   // prefix(var_seq_, posCurrentlyAt), reverse complemented
			--cur_var;
			--cur_start;
		}
		for( ; cur_var >= 0 && len < frag_length; --cur_var){
			if(vars.in_allele(cur_var, allele)){
				if(static_cast<uint64_t>(cur_start - vars.position[cur_var]) > static_cast<uint64_t>(frag_length) - len){
					put_ref_rc(static_cast<uint64_t>(cur_start) + len - frag_length, cur_start);   // variant lies behind the returned sequence
				}
				else{
					put_ref_rc(vars.position[cur_var] + 1u, cur_start);
					for(uint32_t k = vars.length(cur_var); k > 0; --k){ put(3u - vars.base(cur_var, k - 1)); }
					cur_start = vars.position[cur_var];
				}
			}
		}
		if(cur_var == -1 && len < frag_length){
			put_ref_rc(static_cast<uint64_t>(cur_start) + len - frag_length, cur_start);
		}
	}
	else{
		auto put_ref = [&](uint32_t from, uint64_t to){ for(uint64_t p = from; p < to; ++p){ put(seq[p]); } };
		if(first_variant_pos){
			for(uint32_t k = first_variant_pos; k < vars.length(cur_var); ++k){ put(vars.base(cur_var, k)); }   // suffix(var_seq_, posCurrentlyAt)
			++cur_var;
			++cur_start;
		}
		for( ; static_cast<uint32_t>(cur_var) < vars.n && len < frag_length; ++cur_var){
			if(vars.in_allele(cur_var, allele)){
				if(static_cast<uint64_t>(vars.position[cur_var] - cur_start) >= static_cast<uint64_t>(frag_length) - len){
					put_ref(cur_start, static_cast<uint64_t>(cur_start) + frag_length - len);
				}
				else{
					put_ref(cur_start, vars.position[cur_var]);
					for(uint32_t k = 0; k < vars.length(cur_var); ++k){ put(vars.base(cur_var, k)); }
					cur_start = vars.position[cur_var] + 1u;
				}
			}
		}
		if(static_cast<uint32_t>(cur_var) == vars.n && len < frag_length){
			put_ref(cur_start, static_cast<uint64_t>(cur_start) + frag_length - len);
		}
	}
	return len < frag_length ? len : frag_length;
}

// SimBlock::sys_errors_ / err_variants_ of one strand of one sequence, flattened: block b owns sys[block_start[b] .. block_start[b+1]) and the
// variants [var_first[b], var_first[b+1]); a variant's position counts from the block start; an entry is dominant error | rate << 8.
struct SysErrorVariantView {
	const uint16_t *sys;
	const uint32_t *block_start;   // [n_blocks + 1]
	const uint32_t *var_first;     // [n_blocks + 1]
	const uint32_t *position;      // [n_var]
	const uint32_t *err_off;       // [n_var + 1] into errs (var_errors_: empty = deletion, 1 = substitution, more = insertion)
	const uint16_t *errs;
	const uint64_t *allele_lo, *allele_hi;
	RSQ_HD bool in_allele(uint32_t var, uint32_t allele) const { return ((allele < 64 ? allele_lo[var] : allele_hi[var]) >> (allele & 63u)) & 1u; }
};
struct SysErrorCursor { uint32_t block, block_pos; int32_t cur_var; uint32_t var_pos; };   // block, block_pos, cur_var, var_pos of the reference

// Returns dominant error | rate << 8 and advances the cursor exactly like the reference - including its two oddities: inside an insertion
// (var_pos > 0) the values come from sys_errors_[block_pos], not from var_errors_[var_pos], and a substitution advances cur_var twice.
RSQ_HD uint16_t sys_error_with_variants(const SysErrorVariantView &v, SysErrorCursor &c, uint32_t allele){
	auto increment_block_pos = [&](){   // IncrementBlockPos (cur_var already advanced by the caller where the reference passes ++cur_var)
		if(v.block_start[c.block + 1] - v.block_start[c.block] <= ++c.block_pos){ ++c.block; c.block_pos = 0; c.cur_var = 0; }
	};
	uint16_t res = 0;
	bool no_variant = true;
	if(c.var_pos){
		no_variant = false;
		res = v.sys[v.block_start[c.block] + c.block_pos];
		const uint32_t var = v.var_first[c.block] + c.cur_var;
		if(++c.var_pos >= v.err_off[var + 1] - v.err_off[var]){
			c.var_pos = 0;
			++c.cur_var;
			increment_block_pos();
		}
	}
	else{
		while(static_cast<uint32_t>(c.cur_var) < v.var_first[c.block + 1] - v.var_first[c.block] && v.position[v.var_first[c.block] + c.cur_var] <= c.block_pos){
			const uint32_t var = v.var_first[c.block] + c.cur_var;
			if(v.in_allele(var, allele)){
				const uint32_t n_err = v.err_off[var + 1] - v.err_off[var];
				if(n_err == 0){   // deletion
					++c.cur_var;
					increment_block_pos();
				}
				else{
					no_variant = false;
					res = v.errs[v.err_off[var]];
					if(n_err == 1){   // substitution
						++c.cur_var;
						increment_block_pos();
						++c.cur_var;
					}
					else{ c.var_pos = 1; }   // insertion
					break;
				}
			}
			else{ ++c.cur_var; }
		}
	}
	if(no_variant){
		res = v.sys[v.block_start[c.block] + c.block_pos];
		increment_block_pos();
	}
	return res;
}

// One allele of one reference sequence: its materialised bases, the map from reference positions and a GC prefix over the allele's bases.
struct AlleleView {
	const uint8_t *bases;       // [n_bases]
	const uint32_t *off;        // [ref_len + 1]
	const uint32_t *gc_prefix;  // [n_bases + 1] number of G/C in bases[0 .. i)
	uint32_t n_bases, ref_len;
};
struct AlleleFragment { uint32_t gc_percent, end_position; uint32_t sur_start[3], sur_end[3]; };

// start_variant_pos: 0, or the inserted base of the insertion at `start` the fragment starts from (the allele carries that insertion).
// end_position is the reference's cur_end_position = cur_start_position + fragment_length + end_pos_shift_: the smallest reference
// position whose first allele base lies at or behind the fragment's end.
RSQ_HD void allele_fragment(const AlleleView &a, uint32_t start, uint32_t start_variant_pos, uint32_t fragment_length, AlleleFragment &f){
	const uint32_t first = a.off[start] + start_variant_pos, end = first + fragment_length;
	uint32_t lo = start, hi = a.ref_len;   // lower bound of `end` in off[start .. ref_len]
	while(lo < hi){
		const uint32_t mid = lo + (hi - lo) / 2;
		if(a.off[mid] < end){ lo = mid + 1; } else{ hi = mid; }
	}
	f.end_position = lo;
	const uint32_t gc = a.gc_prefix[end < a.n_bases ? end : a.n_bases] - a.gc_prefix[first];
	f.gc_percent = ((gc * 100u + fragment_length / 2u) / fragment_length) & 0xffu;   // utilities::Percent into uintPercent
	forward_surrounding(a.bases, a.n_bases, first, f.sur_start);
	reverse_surrounding(a.bases, a.n_bases, end - 1u, f.sur_end);
}

// mean = bias * bias_normalization of the allele's fragment; (disp_a, disp_b) = dispersion_parameters_. sim_core.cuh's fragment_counts is
// the alleles == 1 form of this (both divisions exact there). `runaway` mirrors its guard against a count that wraps uintDupCount.
RSQ_HD uint32_t allele_fragment_counts(double mean, double disp_a, double disp_b, uint32_t alleles, double probability_chosen, bool &runaway){
	double r = mean / add_rn(disp_a, mul_rn(disp_b, mean));   // BiasCalculationVectors::GetDispersion
	const double cap = mul_rn(mean, 1e10);
	if(r > cap){ r = cap; }
	const double n = static_cast<double>(alleles);
	r = r / n;
	mean = mean / n;
	const double p = mean / add_rn(mean, r);
	double probability_count = pow_glibc(sub_rn(1.0, p), r);   // NegativeBinomial
	double probability_left = sub_rn(probability_chosen, probability_count);
	uint32_t count = 0;
	while(0.0 < probability_left){
		count = (count + 1) & 0xffffu;   // uintDupCount
		probability_count = mul_rn(probability_count, mul_rn(p, add_rn(sub_rn(r, 1.0) / static_cast<double>(static_cast<int>(count)), 1.0)));
		probability_left = sub_rn(probability_left, probability_count);
		if(count == 0xffffu){ runaway = true; break; }
	}
	return count;
}

// Sampling `non_zero_strands` of `possible_strands` (= 2 x possible alleles, <= 256) ids without replacement, in the reference's order:
// up to half are drawn directly, more than half by drawing the complement and listing what is left in ascending order.
// `uniform()` is GeneralRandomDistributions::ZeroToOne on the block's stream; one value per drawn id. Returns the number of ids in `chosen`.
template<class Uniform> RSQ_HD uint32_t choose_alleles(uint16_t *chosen, uint32_t non_zero_strands, uint32_t possible_strands, Uniform &&uniform){
	uint64_t open[4] = {~0ull, ~0ull, ~0ull, ~0ull};   // reverse_selection: bit set = id not chosen yet
	auto is_open = [&](uint32_t id){ return (open[id >> 6] >> (id & 63u)) & 1ull; };
	const bool direct = non_zero_strands <= possible_strands / 2;
	const uint32_t n_draw = direct ? non_zero_strands : possible_strands - non_zero_strands;
	uint32_t n = 0;
	while(n < n_draw){
		// SelectAllele
		uint32_t id = static_cast<uint16_t>(mul_rn(uniform(), static_cast<double>(possible_strands - n)));
		uint32_t correction = 0;
		for(uint32_t k = 0; k < n; ++k){ if(chosen[k] <= id){ ++correction; } }
		while(correction){ if(is_open(++id)){ --correction; } }
		chosen[n++] = static_cast<uint16_t>(id);
		open[id >> 6] &= ~(1ull << (id & 63u));
	}
	if(direct){ return n; }
	n = 0;   // ReverseSelection
	for(uint32_t id = 0; id < possible_strands; ++id){ if(is_open(id)){ chosen[n++] = static_cast<uint16_t>(id); } }
	return n;
}


// ===================================================================================================================
// Variant-aware simulation on the engine's layout (what the kernels consume)
// ===================================================================================================================
// Reference::variants_ of all sequences, flattened (variants.hpp: FlatVariants), and what Simulator derives from them per run.
struct VarCtx {
	uint32_t loaded;                 // Reference::VariantsLoaded()
	uint32_t num_alleles;            // Reference::NumAlleles()
	const uint32_t *seq_first;       // [n_seqs + 1] first variant of each sequence
	const uint32_t *position;        // [n_var]
	const uint32_t *bases_off;       // [n_var + 1]
	const uint8_t *bases;
	const uint64_t *allele_lo, *allele_hi;
	// SysErrorVariant::var_errors_ (Simulator.h:91-106) of the forward / reverse SimBlock holding the variant: (dominant error, rate) of replacement
	// base i at 2 * (bases_off[v] + i); reverse: i counts in reverse-strand order (the order SetSystematicErrorVariantsReverse draws them in)
	const uint8_t *errs_fwd, *errs_rev;
	// SimBlock::first_variant_id_: per sequence ceil(L / 1000) + 1 entries starting at block_first_off[seq]; entry b = first variant (counted inside
	// the sequence) with position >= 1000 * b.  Forward block b starts from entry b, its reverse partner from entry b + 1 minus one.
	const uint32_t *block_first;
	const uint32_t *block_first_off; // [n_seqs]
	const double *sur_tab[3];        // fragment_surroundings_bias_: surroundings of fragments that touch a variant are evaluated per hit
	const double *binom_all;         // unused slot kept for layout stability
	RSQ_HD VariantView view(uint32_t seq) const {
		VariantView v;
		const uint32_t f = seq_first[seq];
		v.position = position + f; v.bases_off = bases_off + f; v.bases = bases; v.allele_lo = allele_lo + f; v.allele_hi = allele_hi + f; v.n = seq_first[seq + 1] - f;
		return v;
	}
};

RSQ_HD uint32_t var_lower_bound(const VariantView &v, uint32_t p){   // first variant with position >= p
	uint32_t lo = 0, hi = v.n;
	while(lo < hi){ const uint32_t mid = lo + (hi - lo) / 2; if(v.position[mid] < p){ lo = mid + 1; } else{ hi = mid; } }
	return lo;
}
// The same from a nearby index (the scan always has one: the first variant at or behind the start position): a few steps instead of a search
RSQ_HD uint32_t var_seek(const VariantView &v, uint32_t hint, uint32_t p){
	if(hint > v.n){ hint = v.n; }
	while(hint > 0 && v.position[hint - 1] >= p){ --hint; }
	while(hint < v.n && v.position[hint] < p){ ++hint; }
	return hint;
}
RSQ_HD bool var_in_range(const VariantView &v, uint32_t hint, uint32_t lo, uint32_t hi){   // any variant with lo <= position <= hi
	const uint32_t i = var_seek(v, hint, lo);
	return i < v.n && v.position[i] <= hi;
}
// The variant of `allele` at reference position p (one per position and allele), or -1.  vi: any index <= the first variant at p; left at the first variant with position >= p.
RSQ_HD int32_t allele_variant_at(const VariantView &v, uint32_t &vi, uint32_t p, uint32_t allele){
	while(vi < v.n && v.position[vi] < p){ ++vi; }
	for(uint32_t t = vi; t < v.n && v.position[t] == p; ++t){ if(v.in_allele(t, allele)){ return static_cast<int32_t>(t); } }
	return -1;
}

// Simulator::AlleleSkipped (Simulator.h:401-413): alleles that delete the start base, and - on the passes that start from an inserted base -
// alleles that do not carry that insertion
RSQ_HD bool allele_skipped(const VariantView &v, uint32_t first_var, uint32_t start_variant_pos, uint32_t pos, uint32_t allele){
	if(first_var < v.n && v.position[first_var] == pos){
		if(0 == v.length(first_var)){ return v.in_allele(first_var, allele); }
		else if(start_variant_pos){ return !v.in_allele(first_var, allele); }
	}
	return false;
}
RSQ_HD uint32_t count_possible_alleles(const VariantView &v, uint32_t num_alleles, uint32_t first_var, uint32_t start_variant_pos, uint32_t pos){
	if(!(first_var < v.n && v.position[first_var] == pos)){ return num_alleles; }
	uint32_t n = 0;
	for(uint32_t a = 0; a < num_alleles; ++a){ n += allele_skipped(v, first_var, start_variant_pos, pos, a) ? 0u : 1u; }
	return n;
}
RSQ_HD uint32_t nth_possible_allele(const VariantView &v, uint32_t num_alleles, uint32_t first_var, uint32_t start_variant_pos, uint32_t pos, uint32_t k){
	if(!(first_var < v.n && v.position[first_var] == pos)){ return k; }
	for(uint32_t a = 0; a < num_alleles; ++a){
		if(!allele_skipped(v, first_var, start_variant_pos, pos, a)){ if(0 == k){ return a; } --k; }
	}
	return 0;
}
// Simulator::CheckForInsertedBasesToStartFrom (Simulator.cpp:1875-1896): the start position is visited once more per further inserted base
RSQ_HD void next_start_pass(const VariantView &v, uint32_t pos, uint32_t &first_var, uint32_t &start_variant_pos){
	if(first_var < v.n && v.position[first_var] == pos){
		if(start_variant_pos){
			if(++start_variant_pos >= v.length(first_var)){ start_variant_pos = 0; ++first_var; }
		}
		else{
			while(first_var < v.n && v.position[first_var] == pos && 2 > v.length(first_var)){ ++first_var; }
		}
		if(0 == start_variant_pos){
			if(first_var < v.n && v.position[first_var] == pos){ start_variant_pos = 1; }
		}
	}
}

// A point between two bases of an allele's sequence: in front of reference position p, or - inside an insertion the allele carries at p -
// behind the first k (0 < k < its length) bases of that insertion `ins`.
struct AllelePoint { uint32_t p; uint32_t k; int32_t ins; };

// The next n bases of the allele behind the point.  Beyond the sequence end the reference rolls around into its own first bases and ignores
// variants there (Simulator.cpp:1601, 1706).  hint: an index near the first variant at or behind the point.
RSQ_HD_VCOLD void allele_bases_forward(const VariantView &v, const uint8_t *seq, uint32_t L, uint32_t allele, AllelePoint at, uint32_t n, uint8_t *out, uint32_t hint = 0){
	uint32_t i = 0, p = at.p;
	if(at.k){
		for(uint32_t j = at.k; j < v.length(at.ins) && i < n; ++j){ out[i++] = v.base(at.ins, j); }
		++p;
	}
	uint32_t vi = var_seek(v, hint, p);
	while(i < n){
		if(p >= L){ out[i++] = seq[(p - L) % L]; ++p; continue; }
		// reference bases up to the next position that carries a variant of any allele
		const uint32_t q = vi < v.n ? v.position[vi] : L;
		uint32_t run = (q < L ? q : L) - p;
		if(run > n - i){ run = n - i; }
		for(uint32_t k = 0; k < run; ++k){ out[i + k] = seq[p + k]; }
		i += run; p += run;
		if(i >= n || p >= L){ continue; }
		const int32_t t = allele_variant_at(v, vi, p, allele);
		if(t < 0){ out[i++] = seq[p]; }
		else{ for(uint32_t j = 0; j < v.length(t) && i < n; ++j){ out[i++] = v.base(t, j); } }
		++p;
		while(vi < v.n && v.position[vi] < p){ ++vi; }
	}
}
// The n bases of the allele in front of the point, nearest first.  In front of the sequence start: the reference's last bases, variants ignored.
RSQ_HD_VCOLD void allele_bases_backward(const VariantView &v, const uint8_t *seq, uint32_t L, uint32_t allele, AllelePoint at, uint32_t n, uint8_t *out, uint32_t hint = 0){
	uint32_t i = 0;
	if(at.k){ for(uint32_t j = at.k; j-- > 0 && i < n; ){ out[i++] = v.base(at.ins, j); } }
	int64_t q = static_cast<int64_t>(at.p) - 1;
	uint32_t vi = var_seek(v, hint, at.p);   // one behind the last variant with position < p
	while(i < n){
		if(q < 0){ out[i++] = seq[static_cast<uint32_t>((static_cast<int64_t>(L) + q % static_cast<int64_t>(L)) % static_cast<int64_t>(L))]; --q; continue; }
		// reference bases down to the next lower position that carries a variant of any allele
		const int64_t vq = vi > 0 ? static_cast<int64_t>(v.position[vi - 1]) : -1;
		uint32_t run = static_cast<uint32_t>(q - vq);
		if(run > n - i){ run = n - i; }
		for(uint32_t k = 0; k < run; ++k){ out[i + k] = seq[q - k]; }
		i += run; q -= run;
		if(i >= n || q < 0){ continue; }
		int32_t t = -1;
		for(uint32_t u = vi; u > 0 && v.position[u - 1] == static_cast<uint32_t>(q); --u){ if(v.in_allele(u - 1, allele)){ t = static_cast<int32_t>(u - 1); break; } }
		if(t < 0){ out[i++] = seq[q]; }
		else{ for(uint32_t j = v.length(t); j-- > 0 && i < n; ){ out[i++] = v.base(t, j); } }
		--q;
		while(vi > 0 && static_cast<int64_t>(v.position[vi - 1]) > q){ --vi; }
	}
}

// The same two walks for a lane group writing to memory all lanes see (staged fragment ends): the runs of reference bases between variants
// are copied by all lanes, single replacement bases by lane 0; every lane keeps the same bookkeeping.  comp: store 3 - base.
template<class G>
RSQ_HD_VCOLD void allele_bases_forward_g(const G &g, const VariantView &v, const uint8_t *seq, uint32_t L, uint32_t allele, AllelePoint at, uint32_t n, uint8_t *out, uint32_t hint){
	uint32_t i = 0, p = at.p;
	if(at.k){
		if(g.lane() == 0){ uint32_t ii = 0; for(uint32_t j = at.k; j < v.length(at.ins) && ii < n; ++j){ out[ii++] = v.base(at.ins, j); } }
		const uint32_t rest = v.length(at.ins) - at.k;
		i = rest < n ? rest : n;
		++p;
	}
	uint32_t vi = var_seek(v, hint, p);
	while(i < n){
		if(p >= L){ if(g.lane() == 0){ out[i] = seq[(p - L) % L]; } ++i; ++p; continue; }
		const uint32_t q = vi < v.n ? v.position[vi] : L;
		uint32_t run = (q < L ? q : L) - p;
		if(run > n - i){ run = n - i; }
		for(uint32_t k = g.lane(); k < run; k += G::kSize){ out[i + k] = seq[p + k]; }
		i += run; p += run;
		if(i >= n || p >= L){ continue; }
		const int32_t t = allele_variant_at(v, vi, p, allele);
		if(t < 0){ if(g.lane() == 0){ out[i] = seq[p]; } ++i; }
		else{
			const uint32_t len = v.length(t), take = len < n - i ? len : n - i;
			if(g.lane() == 0){ for(uint32_t j = 0; j < take; ++j){ out[i + j] = v.base(t, j); } }
			i += take;
		}
		++p;
		while(vi < v.n && v.position[vi] < p){ ++vi; }
	}
	g.sync();
}
template<class G>
RSQ_HD_VCOLD void allele_bases_backward_g(const G &g, const VariantView &v, const uint8_t *seq, uint32_t L, uint32_t allele, AllelePoint at, uint32_t n, uint8_t *out, uint32_t hint, bool comp){
	const uint32_t cx = comp ? 3u : 0u;   // 3 - b == 3 ^ b for two-bit codes
	uint32_t i = 0;
	if(at.k){
		if(g.lane() == 0){ uint32_t ii = 0; for(uint32_t j = at.k; j-- > 0 && ii < n; ){ out[ii++] = static_cast<uint8_t>(v.base(at.ins, j) ^ cx); } }
		i = at.k < n ? at.k : n;
	}
	int64_t q = static_cast<int64_t>(at.p) - 1;
	uint32_t vi = var_seek(v, hint, at.p);
	while(i < n){
		if(q < 0){ if(g.lane() == 0){ out[i] = static_cast<uint8_t>(seq[static_cast<uint32_t>((static_cast<int64_t>(L) + q % static_cast<int64_t>(L)) % static_cast<int64_t>(L))] ^ cx); } ++i; --q; continue; }
		const int64_t vq = vi > 0 ? static_cast<int64_t>(v.position[vi - 1]) : -1;
		uint32_t run = static_cast<uint32_t>(q - vq);
		if(run > n - i){ run = n - i; }
		for(uint32_t k = g.lane(); k < run; k += G::kSize){ out[i + k] = static_cast<uint8_t>(seq[q - k] ^ cx); }
		i += run; q -= run;
		if(i >= n || q < 0){ continue; }
		int32_t t = -1;
		for(uint32_t u = vi; u > 0 && v.position[u - 1] == static_cast<uint32_t>(q); --u){ if(v.in_allele(u - 1, allele)){ t = static_cast<int32_t>(u - 1); break; } }
		if(t < 0){ if(g.lane() == 0){ out[i] = static_cast<uint8_t>(seq[q] ^ cx); } ++i; }
		else{
			const uint32_t len = v.length(t), take = len < n - i ? len : n - i;
			if(g.lane() == 0){ for(uint32_t j = 0; j < take; ++j){ out[i + j] = static_cast<uint8_t>(v.base(t, len - 1u - j) ^ cx); } }
			i += take;
		}
		--q;
		while(vi > 0 && static_cast<int64_t>(v.position[vi - 1]) > q){ --vi; }
	}
	g.sync();
}

// What SimulateFromGivenBlock needs for one (start position, inserted start base, fragment length, allele): PrepareBiasModForCurrentStartPos /
// ...FragmentLength, GetGCPercent, end_pos_shift_ and EndVariant of the reference (Simulator.cpp:1399-1873, Simulator.h:68-88) evaluated directly on
// the allele's sequence (tests/test_variant_invariant_cpu.py states the relations and checks them on traces of the unmodified reference).
struct AlleleHit {
	uint32_t end_position;            // cur_end_position = cur_start_position + fragment_length + end_pos_shift_
	uint32_t gc_percent;              // GetGCPercent
	int32_t end_var; uint32_t end_var_pos;   // EndVariant: {id, posCurrentlyAt}
	AllelePoint end;                  // the point behind the fragment's last base
	uint32_t valid;                   // the fragment ends inside the sequence (cur_end_position < SequenceLength)
	uint32_t end_hint;                // an index near the first variant at or behind end_position
};
RSQ_HD uint32_t count_gc_bases(const VariantView &v, uint32_t var, uint32_t from, uint32_t to){
	uint32_t gc = 0;
	for(uint32_t j = from; j < to; ++j){ const uint32_t b = v.base(var, j); gc += (b == 1u || b == 2u) ? 1u : 0u; }
	return gc;
}
// gcp: G/C prefix counts of the reference sequence ([L + 1]).  first_var / start_variant_pos: VariantBiasVarModifiers::StartVariant.
RSQ_HD_VCOLD void allele_hit(const VariantView &v, const uint32_t *gcp, uint32_t L, uint32_t allele, uint32_t pos, uint32_t first_var, uint32_t start_variant_pos,
                       uint32_t fl, AlleleHit &h){
	uint32_t consumed = 0, gc = 0, p = pos;
	h.valid = 0; h.end_var = -1; h.end_var_pos = 0; h.end = AllelePoint{0, 0, -1}; h.end_position = L; h.gc_percent = 0;
	bool done = false, inside = false;
	if(start_variant_pos){
		const uint32_t len = v.length(first_var);
		const uint32_t take = (len - start_variant_pos) < fl ? (len - start_variant_pos) : fl;
		gc += count_gc_bases(v, first_var, start_variant_pos, start_variant_pos + take);
		consumed = take;
		p = pos + 1;
		if(consumed == fl){
			done = true;
			h.end_position = pos + 1;
			if(start_variant_pos + take < len){ h.end = AllelePoint{pos, start_variant_pos + take, static_cast<int32_t>(first_var)}; }
			else{ h.end = AllelePoint{pos + 1, 0, -1}; }
			// EndVariant: a fragment of one base names the start insertion (second branch), longer ones inside it fall through to the last variant in front
			// of the end position (third branch)
			if(1 == fl){ h.end_var = static_cast<int32_t>(first_var); h.end_var_pos = start_variant_pos + 1; inside = true; }
		}
	}
	uint32_t vi = var_seek(v, first_var, p);
	h.end_hint = vi;
	while(!done){
		// the allele's next variant at or behind p
		uint32_t t = vi;
		while(t < v.n && !v.in_allele(t, allele)){ ++t; }
		const uint32_t q = t < v.n ? v.position[t] : L;
		const uint32_t seg = q - p;
		if(consumed + seg >= fl){
			const uint32_t k = fl - consumed;
			gc += gcp[p + k] - gcp[p];
			h.end_position = p + k; h.end = AllelePoint{p + k, 0, -1}; h.end_hint = t;
			break;
		}
		if(q >= L){ return; }   // runs off the sequence
		gc += gcp[q] - gcp[p]; consumed += seg;
		const uint32_t len = v.length(t);
		const uint32_t take = len < fl - consumed ? len : fl - consumed;
		gc += count_gc_bases(v, t, 0, take);
		consumed += take;
		if(consumed == fl){
			h.end_position = q + 1;
			if(take < len){ h.end = AllelePoint{q, take, static_cast<int32_t>(t)}; h.end_var = static_cast<int32_t>(t); h.end_var_pos = take; inside = true; }
			else{ h.end = AllelePoint{q + 1, 0, -1}; }
			h.end_hint = t;
			break;
		}
		p = q + 1;
		vi = t + 1;
		while(vi < v.n && v.position[vi] <= q){ ++vi; }
	}
	if(h.end_position >= L){ return; }
	h.valid = 1;
	h.gc_percent = ((gc * 100u + fl / 2u) / fl) & 0xffu;   // utilities::Percent on uintSeqLen into uintPercent
	h.end_hint = var_seek(v, h.end_hint, h.end_position);
	if(!inside){ h.end_var = static_cast<int32_t>(h.end_hint) - 1; h.end_var_pos = 0; }
}

RSQ_HD uint32_t pack_10mer(const uint8_t *b){ uint32_t s = 0; for(uint32_t k = 0; k < 10; ++k){ s = (s << 2) + (b[k] & 3u); } return s; }
// Surrounding codes of an allele's fragment: bias_mod.surrounding_start_ at its first base, bias_mod.surrounding_end_ at its last base
RSQ_HD void allele_start_surrounding(const VariantView &v, const uint8_t *seq, uint32_t L, uint32_t allele, uint32_t pos, uint32_t first_var, uint32_t start_variant_pos, uint32_t code[3]){
	uint8_t w[30], back[10];
	const AllelePoint at{pos, start_variant_pos, start_variant_pos ? static_cast<int32_t>(first_var) : -1};
	allele_bases_backward(v, seq, L, allele, at, 10, back, first_var);
	for(uint32_t k = 0; k < 10; ++k){ w[k] = back[9 - k]; }
	allele_bases_forward(v, seq, L, allele, at, 20, w + 10, first_var);
	code[0] = pack_10mer(w); code[1] = pack_10mer(w + 10); code[2] = pack_10mer(w + 20);
}
RSQ_HD void allele_end_surrounding(const VariantView &v, const uint8_t *seq, uint32_t L, uint32_t allele, const AllelePoint &end, uint32_t code[3], uint32_t hint = 0){
	uint8_t w[30], fwd[10], back[20];
	allele_bases_forward(v, seq, L, allele, end, 10, fwd, hint);
	allele_bases_backward(v, seq, L, allele, end, 20, back, hint);
	for(uint32_t k = 0; k < 10; ++k){ w[k] = 3u - fwd[9 - k]; }
	for(uint32_t k = 0; k < 20; ++k){ w[10 + k] = 3u - back[k]; }
	code[0] = pack_10mer(w); code[1] = pack_10mer(w + 10); code[2] = pack_10mer(w + 20);
}

// ---------------------------------------------------------------------------------------------------------------
// Systematic errors of a read with variants: Simulator::GetSysErrorFromBlock / IncrementBlockPos (Simulator.cpp:232-292) and the deletion
// branch of FillReadPart (Simulator.cpp:380-392) on the engine's flat per-position arrays.  Forward SimBlock b of a sequence covers
// [1000 b, min(1000 (b + 1), L)); its reverse partner holds the same interval in reverse-strand order and is followed by block b - 1.
// ---------------------------------------------------------------------------------------------------------------
struct SysWalkCtx {
	const uint8_t *sys;          // sys_fwd / sys_rev of the sequence: 2 bytes per position, strand order
	const uint8_t *errs;         // VarCtx::errs_fwd / errs_rev
	const uint32_t *block_first; // VarCtx::block_first of the sequence
	VariantView v;
	uint32_t L; uint32_t reverse;
};
// block, block_pos, cur_var, var_pos of the reference + what a step needs of the current block: its size, the index of its first entry in the
// strand's array and the position of err_variants_[cur_var] (0xffffffff: none left) - refreshed whenever block or cur_var change, so that the
// steps between variants (nearly all of them) cost one load like a run without variants
struct SysWalk { uint32_t block, block_pos; int32_t cur_var; uint32_t var_pos; uint32_t size, next_bp; uint64_t base_idx; };

RSQ_HD uint32_t sysw_block_end(const SysWalkCtx &c, uint32_t b){ const uint64_t e = 1000ull * (b + 1ull); return e < c.L ? static_cast<uint32_t>(e) : c.L; }
RSQ_HD uint32_t sysw_block_size(const SysWalkCtx &c, uint32_t b){ return sysw_block_end(c, b) - 1000u * b; }
RSQ_HD uint32_t sysw_n_vars(const SysWalkCtx &c, uint32_t b){ return c.block_first[b + 1] - c.block_first[b]; }
RSQ_HD uint32_t sysw_var(const SysWalkCtx &c, uint32_t b, uint32_t k){ return c.reverse ? c.block_first[b + 1] - 1u - k : c.block_first[b] + k; }   // err_variants_.at(k)
RSQ_HD uint32_t sysw_var_position(const SysWalkCtx &c, uint32_t b, uint32_t var){ return c.reverse ? sysw_block_end(c, b) - 1u - c.v.position[var] : c.v.position[var] - 1000u * b; }
RSQ_HD void sysw_refresh_var(const SysWalkCtx &c, SysWalk &w){
	w.next_bp = 0xffffffffu;
	if(w.size != 0xffffffffu && w.cur_var >= 0 && static_cast<uint32_t>(w.cur_var) < sysw_n_vars(c, w.block)){
		w.next_bp = sysw_var_position(c, w.block, sysw_var(c, w.block, static_cast<uint32_t>(w.cur_var)));
	}
}
RSQ_HD void sysw_refresh(const SysWalkCtx &c, SysWalk &w){
	const uint32_t n_blocks = (c.L + 999u) / 1000u;
	if(w.block >= n_blocks){ w.size = 0xffffffffu; w.base_idx = 0; w.next_bp = 0xffffffffu; return; }   // block->next_block_ == NULL: nothing is read behind the chain
	w.size = sysw_block_size(c, w.block);
	w.base_idx = c.reverse ? static_cast<uint64_t>(c.L - sysw_block_end(c, w.block)) : 1000ull * w.block;
	sysw_refresh_var(c, w);
}
RSQ_HD const uint8_t *sysw_entry(const SysWalkCtx &c, const SysWalk &w){   // block->sys_errors_.at(block_pos)
	return c.sys + 2ull * (w.base_idx + w.block_pos);
}
RSQ_HD void sysw_increment(const SysWalkCtx &c, SysWalk &w){   // IncrementBlockPos (cur_var advanced by the caller where the reference passes ++cur_var)
	if(w.size <= ++w.block_pos){ w.block = c.reverse ? w.block - 1u : w.block + 1u; w.block_pos = 0; w.cur_var = 0; sysw_refresh(c, w); }
}
// One step between variants: sys: the strand's array (SysWalkCtx::sys).  Returns false when the step has to look at a variant (sysw_next).
RSQ_HD bool sysw_plain_step(const uint8_t *sys, SysWalk &w, uint32_t &res){
	if(w.var_pos || w.block_pos >= w.next_bp || w.block_pos + 1u >= w.size){ return false; }
	const uint8_t *e = sys + 2ull * (w.base_idx + w.block_pos);
	res = e[0] | (static_cast<uint32_t>(e[1]) << 8);
	++w.block_pos;
	return true;
}
// returns dominant error | rate << 8 (out of line: the steps between variants are sysw_plain_step)
RSQ_HD_VCOLD uint32_t sysw_next(const SysWalkCtx &c, SysWalk &w, uint32_t allele){
	uint32_t res = 0;
	if(sysw_plain_step(c.sys, w, res)){ return res; }
	bool no_variant = true;
	if(w.var_pos){
		no_variant = false;
		const uint8_t *e = sysw_entry(c, w);   // the reference reads sys_errors_[block_pos] here, not var_errors_[var_pos]
		res = e[0] | (static_cast<uint32_t>(e[1]) << 8);
		const uint32_t var = sysw_var(c, w.block, static_cast<uint32_t>(w.cur_var));
		if(++w.var_pos >= c.v.length(var)){
			w.var_pos = 0;
			++w.cur_var;
			sysw_increment(c, w);
		}
	}
	else{
		while(w.cur_var >= 0 && static_cast<uint32_t>(w.cur_var) < sysw_n_vars(c, w.block)){
			const uint32_t var = sysw_var(c, w.block, static_cast<uint32_t>(w.cur_var));
			if(!(sysw_var_position(c, w.block, var) <= w.block_pos)){ break; }
			if(c.v.in_allele(var, allele)){
				const uint32_t n_err = c.v.length(var);
				if(0 == n_err){   // deletion
					++w.cur_var;
					sysw_increment(c, w);
				}
				else{
					no_variant = false;
					const uint8_t *e = c.errs + 2ull * c.v.bases_off[var];
					res = e[0] | (static_cast<uint32_t>(e[1]) << 8);
					if(1 == n_err){   // substitution: the reference advances cur_var twice
						++w.cur_var;
						sysw_increment(c, w);
						++w.cur_var;
					}
					else{ w.var_pos = 1; }   // insertion
					break;
				}
			}
			else{ ++w.cur_var; }
		}
	}
	if(no_variant){
		const uint8_t *e = sysw_entry(c, w);
		res = e[0] | (static_cast<uint32_t>(e[1]) << 8);
		sysw_increment(c, w);
	}
	sysw_refresh_var(c, w);
	return res;
}
// FillReadPart's deletion branch: the error rate of sys_errors_[block_pos], then the position advances without looking at the variants
RSQ_HD_VCOLD uint32_t sysw_deletion(const SysWalkCtx &c, SysWalk &w){
	const uint32_t rate = sysw_entry(c, w)[1];
	if(w.var_pos){
		const uint32_t var = sysw_var(c, w.block, static_cast<uint32_t>(w.cur_var));
		if(++w.var_pos >= c.v.length(var)){ w.var_pos = 0; }
	}
	if(0 == w.var_pos && w.size <= ++w.block_pos){ w.block = c.reverse ? w.block - 1u : w.block + 1u; w.block_pos = 0; w.cur_var = 0; sysw_refresh(c, w); }
	return rate;
}
// CreateReads (Simulator.cpp:653-689): where the two reads of a fragment start in the block chains.  start_var / end_var: StartVariant / EndVariant.
RSQ_HD_VCOLD SysWalk sysw_forward_start(const SysWalkCtx &c, uint32_t start_block, uint32_t pos, uint32_t first_var, uint32_t start_variant_pos){
	SysWalk w;
	w.block = start_block; w.block_pos = pos - 1000u * start_block;
	w.cur_var = static_cast<int32_t>(first_var) - static_cast<int32_t>(c.block_first[start_block]); w.var_pos = start_variant_pos;
	sysw_refresh(c, w);
	return w;
}
RSQ_HD_VCOLD SysWalk sysw_reverse_start(const SysWalkCtx &c, uint32_t start_block, uint32_t end_position, int32_t end_var, uint32_t end_var_pos){
	SysWalk w;
	uint32_t b = start_block;
	while(sysw_block_end(c, b) < end_position){ ++b; }
	w.block = b; w.block_pos = sysw_block_end(c, b) - end_position;
	w.cur_var = static_cast<int32_t>(c.block_first[b + 1]) - 1 - end_var;
	w.var_pos = end_var_pos ? c.v.length(static_cast<uint32_t>(end_var)) - end_var_pos : 0u;
	sysw_refresh(c, w);
	return w;
}

} // namespace rsq
