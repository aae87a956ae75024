// Fragment-end surroundings and the bias-normalisation pre-pass:
//   SurroundingBase<3,10,10,int32_t>::Set / Forward / Reverse   (reference SurroundingBase.hpp:64-81,196-201)
//   SurroundingBias::Bias                                       (Surrounding.h:114-120)
//   Reference::SumBias (simulation overload)                    (Reference.cpp:622-659)
// Without variants the surrounding of a fragment end depends on the reference position only, so both
// biases are evaluated once per position (2 exp per position instead of 2 per (position, length)) and
// SumBias turns into a strictly ordered FP64 sum over precomputed factors.
#pragma once
#include "core.cuh"

namespace rsq {

// 3 x 10-mer codes of the 30 bases starting 10 before `pos`, circular in the sequence.
// Index arithmetic is done in uint32 like the reference (pos + seq_len - kStartPos + offset) % seq_len.
RSQ_HD void forward_surrounding(const uint8_t *seq, uint32_t L, uint32_t pos, uint32_t code[3]){
	const uint32_t start = pos + L - 10u;
	for(uint32_t block = 0; block < 3; ++block){
		uint32_t sur = 0;
		for(uint32_t k = 0; k < 10; ++k){
			sur = (sur << 2) + seq[(start + block * 10u + k) % L];
		}
		code[block] = sur;
	}
}
// Same on the reverse complement strand, anchored at the last base `pos` of a fragment
RSQ_HD void reverse_surrounding(const uint8_t *seq, uint32_t L, uint32_t pos, uint32_t code[3]){
	const uint32_t start = (L - pos - 1u) + L - 10u;
	for(uint32_t block = 0; block < 3; ++block){
		uint32_t sur = 0;
		for(uint32_t k = 0; k < 10; ++k){
			const uint32_t q = (start + block * 10u + k) % L;   // coordinate on the reverse complement
			sur = (sur << 2) + (3u - seq[L - 1u - q]);
		}
		code[block] = sur;
	}
}
RSQ_HD double surrounding_bias(const double *t0, const double *t1, const double *t2, const uint32_t code[3]){
	double bias = 0.0;
	bias = add_rn(bias, t2[code[2]]);
	bias = add_rn(bias, t1[code[1]]);
	bias = add_rn(bias, t0[code[0]]);
	return inv_logit2(bias);
}

// One SumBias chain: all start positions of fragments of `fragment_length` on one sequence, summed in
// position order.  general = ref_seq_bias * insert_lengths_bias[fragment_length].
RSQ_HD double sum_bias_chain(const double *sur_start, const double *sur_end, const uint32_t *gc_prefix, uint32_t L,
                             uint32_t fragment_length, double general, const double *gc_bias, double &max_bias){
	double tot = 0.0;
	double mx = max_bias;
	for(uint32_t p = 0; p + fragment_length <= L; ++p){
		const uint32_t gc = gc_prefix[p + fragment_length] - gc_prefix[p];
		double bias = mul_rn(general, gc_bias[percent_u32(gc, fragment_length)]);
		bias = mul_rn(bias, sur_start[p]);
		bias = mul_rn(bias, sur_end[p + fragment_length - 1]);
		if(bias > mx){ mx = bias; }
		tot = add_rn(tot, bias);
	}
	max_bias = mx;
	return tot;
}

}  // namespace rsq
