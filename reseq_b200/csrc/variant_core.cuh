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
// Not consumed by a kernel yet: rsq_engine_prepare refuses references with variants until the scan and read kernels take alleles.
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

} // namespace rsq
