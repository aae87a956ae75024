// Exact parallel evaluation of a strictly ordered FP64 sum of non-negative terms.
//
// Reference::SumBias (Reference.cpp:622-659) adds the bias of every start position of a sequence in position order:
//     s_{i+1} = RN(s_i + x_i),  s_0 = 0,  x_i >= 0.
// FP64 addition is not associative, so the result has to be the one of exactly this order - a chain of 10^8 dependent additions for a
// chromosome (0.56 s on a B200, the part of the prologue that did not shrink with more GPUs).  The chain can be cut into chunks all
// the same:
//
//   * While s stays inside one binade [2^e, 2^(e+1)), every s_i is a multiple of u = 2^(e-52) and RN(s_i + x) = s_i + R_u(x) with
//     R_u(x) = x rounded to the nearest multiple of u - *unless* x lies exactly half way between two multiples (then round-to-even looks
//     at s_i).  Whether a step is such a tie depends on x and u only, not on s_i.
//   * So a chunk that is run from ANY start value g in the right binade yields, if it neither met a tie nor left the binade, the same
//     increments as the true chain: s_out = s_in + (o - g), all three operations exact (multiples of u below 2^(e+1)).
//
//   pass A  per chunk: plain sum of its terms from 0 (any chunk, in parallel)                        -> p_c
//   pass B  per chain: g_0 = 0, g_{c+1} = g_c + p_c   (one addition per chunk: approximate start values in the right binade)
//   pass C  per chunk: the exact chain from g_c, noting ties (TwoSum residual == half an ulp)        -> o_c, tie_c
//   pass D  per chain, in order: s == g_c: s = o_c (identical run).  Same binade before and after and no tie: s += o_c - g_c.
//           Otherwise (a binade boundary - about 50 per chain - or a tie - about ln(chunks) per chain): re-run that chunk from s.
//
// Every decision in D is a comparison of exponent fields; nothing is approximated.  tests/host_twin/ordered_sum_check.cpp runs the four
// passes against the plain chain on adversarial inputs (forced ties, boundaries, zeros, huge and tiny terms).
#pragma once
#include "core.cuh"

namespace rsq {

RSQ_HD uint32_t fp64_exponent_field(double x){
#if defined(__CUDA_ARCH__)
	return static_cast<uint32_t>(__double2hiint(x) >> 20) & 0x7ffu;
#else
	uint64_t b; memcpy(&b, &x, 8); return static_cast<uint32_t>(b >> 52) & 0x7ffu;
#endif
}
// 2^(field - 1023 - 53): half a unit in the last place of a normal number with this exponent field (field > 53)
RSQ_HD double fp64_half_ulp(uint32_t exponent_field){
	const uint64_t b = static_cast<uint64_t>(exponent_field - 53u) << 52;
#if defined(__CUDA_ARCH__)
	return __longlong_as_double(static_cast<long long>(b));
#else
	double d; memcpy(&d, &b, 8); return d;
#endif
}

struct ChunkRun { double out; uint32_t tie; };

// pass A: plain chain from 0 (also yields the chunk's largest term)
template<class Term> RSQ_HD double chunk_plain_sum(const Term &term, uint32_t begin, uint32_t end, double &max_term){
	double s = 0.0;
	for(uint32_t p = begin; p < end; ++p){
		const double x = term(p);
		if(x > max_term){ max_term = x; }
		s = add_rn(s, x);
	}
	return s;
}

// one step of pass C: t = RN(s + x); tie |= the step was a round-to-even decision (or its exponent is too small to tell)
RSQ_HD double ordered_step(double s, double x, uint32_t &tie){
	const double t = add_rn(s, x);
	// TwoSum (Knuth): s + x = t + err exactly
	const double bb = sub_rn(t, s);
	const double err = add_rn(sub_rn(s, sub_rn(t, bb)), sub_rn(x, bb));
	const uint32_t e = fp64_exponent_field(t);
	if(e <= 53u){ tie = 1u; }   // tiny sums (first terms of a chain): half an ulp is not a normal number - decide by re-running
	else{
		const double h = fp64_half_ulp(e);
		if(err == h || err == -h){ tie = 1u; }
	}
	return t;
}

// pass C over terms [begin, end) from the start value g
template<class Term> RSQ_HD ChunkRun chunk_exact_run(const Term &term, uint32_t begin, uint32_t end, double g){
	ChunkRun r{g, 0u};
	for(uint32_t p = begin; p < end; ++p){ r.out = ordered_step(r.out, term(p), r.tie); }
	return r;
}

// pass D for one chunk: the true value behind the chunk, given the true value s in front of it
template<class Term> RSQ_HD double chunk_resolve(const Term &term, uint32_t begin, uint32_t end, double s, double g, double o, uint32_t tie, uint32_t &reran){
	if(s == g){ return o; }   // the speculative run WAS the true run
	if(!tie){
		const uint32_t e = fp64_exponent_field(g);
		if(e > 53u && fp64_exponent_field(o) == e && fp64_exponent_field(s) == e){
			const double cand = add_rn(s, sub_rn(o, g));   // both exact: multiples of the binade's ulp
			if(fp64_exponent_field(cand) == e){ return cand; }
		}
	}
	++reran;
	double t = s;
	for(uint32_t p = begin; p < end; ++p){ t = add_rn(t, term(p)); }
	return t;
}

}  // namespace rsq
