// Building blocks shared by every kernel of the simulation hot path:
//   * lane-group abstraction (a warp on the device; a single lane in the host test twin),
//   * std::mt19937_64 reproduced cooperatively by a lane group (libstdc++ bits/random.tcc semantics:
//     seeding, regeneration, tempering, generate_canonical<double,53>, discrete_distribution),
//   * ProbabilityEstimates' LogArrayResult<N>::Draw (reference ProbabilityEstimates.h:481-508) over
//     flattened tables.
// Everything here is templated on the group type so that the exact same source is exercised on the CPU
// by tests/host_twin (group of 1 lane) and on the GPU (group = 32-lane warp).
#pragma once
#include <stdint.h>
#include "mathx.cuh"

namespace rsq {

// ----------------------------------------------------------------------------------------------
// Lane groups
// ----------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
struct WarpGroup {
	static constexpr int kSize = 32;
	static constexpr bool kCompactCode = false;
	__device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
	__device__ __forceinline__ void sync() const { __syncwarp(); }
	__device__ __forceinline__ unsigned ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
	__device__ __forceinline__ uint32_t reduce_add(uint32_t v) const { return __reduce_add_sync(0xffffffffu, v); }
};
// the same lanes for instantiations whose code has to stay small (the variant-aware scan): rarely run loops are called, not inlined
struct WarpGroupCompact : WarpGroup { static constexpr bool kCompactCode = true; };
#endif
struct SingleLane {
	static constexpr int kSize = 1;
	static constexpr bool kCompactCode = false;
	RSQ_HD int lane() const { return 0; }
	RSQ_HD void sync() const {}
	RSQ_HD unsigned ballot(bool p) const { return p ? 1u : 0u; }
	RSQ_HD uint32_t reduce_add(uint32_t v) const { return v; }
};

// ----------------------------------------------------------------------------------------------
// Flattened probability tables (one entry per LogArrayResult<N>)
// ----------------------------------------------------------------------------------------------
struct TableDesc {
	uint32_t n0;         // par0_indeces_.size()
	uint32_t nm;         // N-1 margins
	uint32_t from[4];    // limits_[n].first
	uint32_t span[4];    // limits_[n].second - limits_[n].first
	uint32_t off[4];     // start of dim2_[n] inside the blob (in doubles)
	uint32_t par0_off;   // start of par0_indeces_ inside the par0 array
	uint32_t stride;     // doubles from one row of a margin to the next: n0, or n0 rounded up to even with a 0.0 behind it (device layout: 16-byte rows)
};

struct Tables {
	const TableDesc *desc;
	const double *blob;
	const uint32_t *par0;
	uint32_t num_tiles;
	uint32_t quality_base, seqq_base, basecall_base, domerr_base, errrate_base, indel_base;
	// table ids in the fixed family order (reference ProbabilityEstimates.h:1320-1325)
	RSQ_HD uint32_t quality(uint32_t seg, uint32_t tile, uint32_t base) const { return quality_base + (seg * num_tiles + tile) * 4 + base; }
	RSQ_HD uint32_t seq_quality(uint32_t seg, uint32_t tile) const { return seqq_base + seg * num_tiles + tile; }
	RSQ_HD uint32_t base_call(uint32_t seg, uint32_t tile, uint32_t base, uint32_t dom) const { return basecall_base + ((seg * num_tiles + tile) * 4 + base) * 5 + dom; }
	RSQ_HD uint32_t dom_error(uint32_t base, uint32_t last, uint32_t dom5) const { return domerr_base + (base * 5 + last) * 5 + dom5; }
	RSQ_HD uint32_t error_rate(uint32_t base, uint32_t dom) const { return errrate_base + base * 5 + dom; }
	RSQ_HD uint32_t indel(uint32_t type, uint32_t last_call) const { return indel_base + type * 6 + last_call; }
};

// LogArrayResult<N>::Draw.  `prob` is group-shared scratch with room for the largest n0.
// Returns par0_indeces_[ind0]; `zero_sum` mirrors the callers' `0.0 == prob_sum` test.
// Likelihood products run lane-parallel; the two sums run in the reference's exact sequential order
// (forward over ind0 for prob_sum, backwards for the cumulative search) because FP64 addition is not
// associative and the result must be bit-identical.
// AdjustIndeces (ProbabilityEstimates.h:368-380) for one margin: clamp into [from, from+span) and rebase
RSQ_HD uint32_t adjust_index(uint32_t v, uint32_t from, uint32_t span){
	const uint32_t hi = from + span - 1u;
	const uint32_t c = v < from ? from : (v > hi ? hi : v);
	return c - from;
}

// (A __noinline__ variant was measured: 128 registers and a 624-byte stack frame in k_simulate - worse.)
// `prob` must hold round_up(n0, 4) doubles: the tail is padded with +0.0 so that the ordered sum can run four
// candidates per trip without a remainder loop (x + 0.0 == x exactly for the non-negative likelihoods).
template<class G>
RSQ_HD uint32_t draw(const G &g, const Tables &t, uint32_t table_id, uint32_t i0, uint32_t i1, uint32_t i2, uint32_t i3,
                     double random_number, double *prob, bool &zero_sum){
	// only constant member indices below: the descriptor must stay in registers
	const TableDesc d = t.desc[table_id];
	const uint32_t n0 = d.n0;
	if(n0 == 0){
		zero_sum = true;
		return 0;
	}
	const bool four = d.nm > 3;
	const double *r0 = t.blob + (d.off[0] + adjust_index(i0, d.from[0], d.span[0]) * d.stride);
	const double *r1 = t.blob + (d.off[1] + adjust_index(i1, d.from[1], d.span[1]) * d.stride);
	const double *r2 = t.blob + (d.off[2] + adjust_index(i2, d.from[2], d.span[2]) * d.stride);
	const double *r3 = four ? t.blob + (d.off[3] + adjust_index(i3, d.from[3], d.span[3]) * d.stride) : r0;
	const uint32_t n4 = (n0 + 3u) & ~3u;
	g.sync();  // previous consumer of `prob` is done
	for(uint32_t i = g.lane(); i < n4; i += G::kSize){
		double p = 0.0;
		if(i < n0){
			p = r0[i];
			p = mul_rn(p, r1[i]);
			p = mul_rn(p, r2[i]);
			if(four){ p = mul_rn(p, r3[i]); }
		}
		prob[i] = p;
	}
	g.sync();
	double prob_sum = 0.0;
	for(uint32_t i = 0; i < n4; i += 4){   // left-to-right like the reference, four candidates per trip
		const double a = prob[i], b = prob[i + 1], c = prob[i + 2], e = prob[i + 3];
		prob_sum = add_rn(add_rn(add_rn(add_rn(prob_sum, a), b), c), e);
	}
	zero_sum = (0.0 == prob_sum);
	const double r = mul_rn(random_number, prob_sum);
	double sum = 0.0;
	uint32_t ind0 = n0;
	while(sum <= r && --ind0){
		sum = add_rn(sum, prob[ind0]);
	}
	return t.par0[d.par0_off + ind0];
}

RSQ_HD uint32_t table_max_value(const Tables &t, uint32_t table_id){   // LogArrayResult::MaxValue
	const TableDesc d = t.desc[table_id];
	uint32_t m = 0;
	for(uint32_t i = 0; i < d.n0; ++i){
		uint32_t v = t.par0[d.par0_off + i];
		if(v > m){ m = v; }
	}
	return m;
}
RSQ_HD uint32_t table_most_likely(const Tables &t, uint32_t table_id){  // LogArrayResult::MostLikely
	const TableDesc d = t.desc[table_id];
	return d.n0 ? t.par0[d.par0_off + d.n0 - 1] : 0;
}

// ----------------------------------------------------------------------------------------------
// std::mt19937_64
// ----------------------------------------------------------------------------------------------
constexpr int kMtN = 312;
constexpr int kMtM = 156;

RSQ_HD uint64_t mt_temper(uint64_t x){
	x ^= (x >> 29) & 0x5555555555555555ull;
	x ^= (x << 17) & 0x71D67FFFEDA60000ull;
	x ^= (x << 37) & 0xFFF7EEE000000000ull;
	x ^= (x >> 43);
	return x;
}

// High word of mt_temper(raw) on its own (the last tempering step only touches the low word): 10 integer operations instead of 17.
// The block scan tests it against the high word of the integer threshold first - a draw whose high word is too small cannot be a hit.
RSQ_HD uint32_t mt_temper_hi(uint64_t raw){
	const uint32_t hi = static_cast<uint32_t>(raw >> 32), lo = static_cast<uint32_t>(raw);
	const uint32_t yh = hi ^ ((hi >> 29) & 0x55555555u);
	const uint32_t yl = lo ^ (((lo >> 29) | (hi << 3)) & 0x55555555u);
	const uint32_t zl = yl ^ ((yl << 17) & 0xEDA60000u);
	const uint32_t zh = yh ^ (((yh << 17) | (yl >> 15)) & 0x71D67FFFu);
	return zh ^ ((zl << 5) & 0xFFF7EEE0u);
}

struct Mt {
	uint64_t *s;   // kMtN words of group-shared memory
	int idx;       // next word to hand out (kMtN => regenerate first), identical in all lanes
};

// seed(value): x[0]=value, x[i] = 6364136223846793005 * (x[i-1] ^ (x[i-1] >> 62)) + i  (strictly serial)
template<class G> RSQ_HD void mt_seed(const G &g, Mt &mt, uint64_t seed){
	g.sync();
	if(g.lane() == 0){
		uint64_t x = seed;
		mt.s[0] = x;
		for(int i = 1; i < kMtN; ++i){
			x = 6364136223846793005ull * (x ^ (x >> 62)) + static_cast<uint64_t>(i);
			mt.s[i] = x;
		}
	}
	mt.idx = kMtN;
	g.sync();
}

// One state transition (312 new words), in place.  Lanes take consecutive words; a chunk's reads all
// happen before its writes, chunks run in ascending order, which preserves the serial algorithm's
// old/new operand pattern: x[i+1] is always still old, x[(i+156)%312] is old for i<156 and new after.
template<class G> RSQ_HD void mt_regen(const G &g, Mt &mt){
	for(int base = 0; base < kMtN; base += G::kSize){
		const int i = base + g.lane();
		uint64_t v = 0;
		if(i < kMtN){
			const uint64_t x = mt.s[i];
			const uint64_t y = mt.s[i + 1 < kMtN ? i + 1 : 0];
			const uint64_t z = mt.s[i + kMtM < kMtN ? i + kMtM : i + kMtM - kMtN];
			const uint64_t w = (x & 0xFFFFFFFF80000000ull) | (y & 0x7FFFFFFFull);
			v = z ^ (w >> 1) ^ ((w & 1ull) ? 0xB5026F5AA96619E9ull : 0ull);
		}
		g.sync();
		if(i < kMtN){ mt.s[i] = v; }
	}
	g.sync();
	mt.idx = 0;
}

template<class G> RSQ_HD uint64_t mt_next(const G &g, Mt &mt){
	if(mt.idx >= kMtN){ mt_regen(g, mt); }
	return mt_temper(mt.s[mt.idx++]);
}

// std::generate_canonical<double,53>(mt19937_64) == uniform_real_distribution<double>(0,1):
// double(x) * 2^-64, with the libstdc++ clamp to nextafter(1,0)
RSQ_HD double canonical(uint64_t x){
#if defined(__CUDA_ARCH__)
	double r = __ull2double_rn(x) * 5.42101086242752217e-20;
#else
	double r = static_cast<double>(x) * 5.42101086242752217e-20;
#endif
	if(r >= 1.0){ r = 0.99999999999999989; }
	return r;
}

template<class G> RSQ_HD double mt_uniform(const G &g, Mt &mt){ return canonical(mt_next(g, mt)); }

// std::discrete_distribution::operator(): no draw when fewer than two weights were given, else
// lower_bound over the cumulative probabilities prepared on the host (last entry forced to 1.0).
struct Discrete {
	const double *cp;
	uint32_t n;   // 0 when the distribution had < 2 weights
};
template<class G> RSQ_HD uint32_t discrete_draw(const G &g, Mt &mt, const Discrete &d){
	if(d.n == 0){ return 0; }
	const double p = mt_uniform(g, mt);
	uint32_t lo = 0, len = d.n;
	while(len > 0){  // std::lower_bound
		uint32_t half = len >> 1;
		if(d.cp[lo + half] < p){ lo += half + 1; len -= half + 1; }
		else{ len = half; }
	}
	return lo;
}

// utilities::Divide / Percent with the reference's integer widths (utilities.hpp:450-452, 552-554)
RSQ_HD uint32_t divide_u32(uint32_t nom, uint32_t den){ return (nom + den / 2) / den; }
RSQ_HD uint8_t percent_u16(uint32_t nom, uint32_t den){   // T = uintReadLen (uint16_t)
	uint32_t n = (nom * 100u) & 0xffffu;
	return static_cast<uint8_t>(static_cast<uint16_t>((n + den / 2) / den));
}
RSQ_HD uint8_t percent_u32(uint32_t nom, uint32_t den){   // T = uintSeqLen (uint32_t)
	uint32_t n = nom * 100u;
	return static_cast<uint8_t>((n + den / 2) / den);
}

}  // namespace rsq
