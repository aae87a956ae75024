// The per-read simulation hot path, restated for a lane group (warp):
//   Simulator::SimulateFromGivenBlock  (reference Simulator.cpp:2249-2357)   -> simulate_block
//   Simulator::CreateReads             (Simulator.cpp:634-721)              -> create_reads
//   Simulator::FillRead / FillReadPart (Simulator.cpp:454-594, 294-452)     -> fill_read / fill_read_part
//   Simulator::CreateReadId            (Simulator.cpp:596-632)              -> format_read_id
//   FragmentDistributionStats::GetFragmentCounts / NegativeBinomial / Binomial (FragmentDistributionStats.cpp:3584-3627)
//   Simulator::SetSystematicErrors / DrawSystematicError (Simulator.h:337-382), CoverageStats::UpdateDistances
//   Simulator::ApplyErrorsAndQualityToFastaInput (Simulator.cpp:2403-2512)  -> error_model_batch
//
// One lane group owns one 1000-bp SimBlock: the block's mt19937_64 stream is consumed strictly in the
// reference's order (scan draw per (position, fragment length), then the draws of every fragment hit),
// lanes only share the work *inside* a step: 32 scan draws at a time, the candidates of a Draw, byte copies.
// Methylation (CTConversion) and variants (simulate_block_var, eval_allele_hit, splice_fragment_ends; building blocks in variant_core.cuh) are part of it.
#pragma once
#include "core.cuh"
#include "variant_core.cuh"

namespace rsq {

struct BlockDesc {
	uint32_t ref_id;
	uint32_t start_pos;
	uint32_t block_id;   // id printed in the read names (forward SimBlock::id_)
	int32_t first_meth;  // SimBlock::first_methylation_id_
	uint64_t seed;
	uint32_t first_var;  // SimBlock::first_variant_id_ (counted inside the sequence)
	uint32_t pad;
};

struct AdapterSet {            // per template segment
	uint32_t n;                // number of adapters
	const uint32_t *off;       // [n+1] offsets into adapter_seq / adapter_sys
	Discrete pick;             // discrete_distribution over SignificantCounts(seg)
	const Discrete *start_cut; // [n]
	const uint32_t *start_cut_from; // [n] StartCut(seg,id).from()
};

struct SimCtx {
	Tables tab;
	// --- profile (DataStats getters used by Simulator) ---
	uint32_t phred_offset;
	uint32_t max_len_deletion;
	uint32_t insert_from;              // max(1, InsertLengths().from())
	uint32_t insert_to;                // InsertLengths().to()
	uint32_t read_len_from[2];         // ReadLengths(seg).from()
	uint32_t read_len_to[2];           // ReadLengths(seg).to()
	uint32_t read_len_count[2];        // ReadLengths(seg).size()
	// ReadLengthsByFragmentLength(seg) for profiles with several read lengths (GeneralRandomDistributions::ReadLength)
	const uint32_t *rlbf_row_from[2];  // per fragment length: row.from()
	const uint32_t *rlbf_row_off[2];   // per fragment length: offset of row values; [n_rows+1]
	const uint64_t *rlbf_val[2];
	uint32_t rlbf_from[2], rlbf_to[2];
	const uint64_t *insert_lengths;    // dense [0, insert_to)
	uint32_t num_tiles;
	const uint16_t *tile_names;
	Discrete tile_pick;
	AdapterSet adapters[2];
	const uint8_t *adapter_seq;        // base codes
	const uint8_t *adapter_sys;        // (dominant error, rate) pairs, same offsets * 2
	Discrete polya_pick;
	uint32_t polya_from;
	Discrete overrun_pick;
	// --- fragment count model ---
	const double *ref_seq_bias;        // [n_seqs]
	const double *il_bias;             // dense [0, insert_to)
	const double *gc_bias;             // dense [0, 101)
	double disp_a, disp_b;
	double bias_normalization;
	const uint32_t *coverage_group;    // [n_seqs]
	const double *thr;                 // [group][insert_to][2]  non_zero_thresholds_
	const uint64_t *thr_int;           // [group][insert_to]     smallest raw draw x with canonical(x) >= thr[..][1] (filter only)
	const uint32_t *thr_hi;            // [group][thr_hi_stride] high words of thr_int: the speculative scan's first filter (staged in shared memory)
	uint32_t thr_hi_stride;            // insert_to + 128 entries of padding (0xffffffff), rounded up to 4 entries (16-byte rows for the bulk copy)
	const double *binom_p0;            // [group][insert_to]     pow(1-(1-thr0), 2) (Binomial's first term, host libm)
	// --- reference ---
	uint32_t n_seqs;
	const uint64_t *seq_off;           // [n_seqs] start of each sequence in the concatenated per-position arrays
	const uint32_t *seq_len;
	const uint8_t *ref;                // base codes 0..3 (after ReplaceN)
	const uint32_t *gc_prefix;         // [total + n_seqs] per sequence: prefix count of G/C, entry i = #GC in [0,i)
	const double *sur_start;           // SurroundingBias::Bias(forward surrounding at p)
	const double *sur_end;             // SurroundingBias::Bias(reverse surrounding at p)
	const uint8_t *sys_fwd;            // 2 bytes / position, forward strand order
	const uint8_t *sys_rev;            // 2 bytes / position, reverse-strand order (index L-1-p)
	const char *name_blob;             // ReferenceIdFirstPart per sequence
	const uint32_t *name_off;          // [n_seqs+1]
	const char *base_id;               // record_base_identifier_
	uint32_t base_id_len;
	uint32_t max_read_len;             // capacity of the per-read scratch
	uint32_t max_org_len;
	uint32_t *error_flag;              // set non-zero on unsupported situations (cigar overflow, runaway count)
	// --- methylation (Reference::unmethylated_regions_ / unmethylation_, allele 0) ---
	uint32_t meth_loaded;              // Reference::MethylationLoaded()
	const uint32_t *meth_off;          // [n_seqs+1] first region of each sequence
	const uint32_t *meth_start;        // region.first
	const uint32_t *meth_end;          // region.second
	const double *meth_rate;           // 1 - methylation = C->T conversion probability (allele 0, or the only column)
	uint32_t meth_alleles;             // Reference::Unmethylation(seq, allele): columns of meth_rate, 1 or Reference::NumAlleles() (sequences with one column repeat it)
	uint32_t meth_rate_stride;         // doubles between the columns of two alleles
	// --- variants (Reference::variants_, SimBlock::err_variants_) ---
	VarCtx var;
	const double *binom_pow;           // [group][insert_to][2 * num_alleles + 1]: pow(thr0, N) of Binomial for N possible strands (host libm); null without variants
};

enum : uint32_t { kErrCigarOverflow = 1, kErrCountRunaway = 2, kErrArenaFull = 4, kErrOrgOverflow = 8, kErrRecordTooLong = 16, kErrReferenceOutOfRange = 32 };

constexpr int kCigarCap = 192;
constexpr int kIdCap = 384;

// Group-shared scratch with a compile-time layout (constant offsets keep the pointers out of registers).
constexpr uint32_t kMaxN0 = 128;       // largest candidate count of any table (error-rate tables: <= 101)
constexpr uint32_t kMaxOrgLen = 384;   // read length + longest deletion, adapters, seqToIllumina fragments
constexpr uint32_t kMaxReadLen = 320;
constexpr uint32_t kScratchBytes = kMtN * 8 + kMaxN0 * 8 + 5 * kMaxOrgLen + 2 * kMaxReadLen + kCigarCap + kIdCap;

struct Scratch {          // group-shared memory
	unsigned char *base;
	RSQ_HD uint64_t *mt_words() const { return reinterpret_cast<uint64_t *>(base); }
	uint64_t *mt;         // kMtN
	double *prob;         // kMaxN0
	uint8_t *org;         // kMaxOrgLen      original bases of the current part
	uint8_t *sdom;        // kMaxOrgLen      dominant systematic error per original base
	uint8_t *srate;       // kMaxOrgLen      systematic error rate per original base
	uint8_t *seq;         // kMaxReadLen     called bases (codes)
	uint8_t *qual;        // kMaxReadLen     qualities (already + phred offset)
	char *cigar;          // kCigarCap
	char *id;             // kIdCap
	uint8_t *frag[2];     // 2 x kMaxOrgLen  bisulfite-converted forward / reverse fragment ends (methylation runs only)
};

RSQ_HD size_t scratch_bytes(uint32_t, uint32_t, uint32_t){ return kScratchBytes; }
RSQ_HD Scratch carve_scratch(unsigned char *base, uint32_t, uint32_t, uint32_t){
	Scratch s;
	s.base = base;
	s.mt = reinterpret_cast<uint64_t *>(base);
	s.prob = reinterpret_cast<double *>(base + kMtN * 8);
	s.org = base + kMtN * 8 + kMaxN0 * 8;
	s.sdom = s.org + kMaxOrgLen;
	s.srate = s.sdom + kMaxOrgLen;
	s.seq = s.srate + kMaxOrgLen;
	s.qual = s.seq + kMaxReadLen;
	s.cigar = reinterpret_cast<char *>(s.qual + kMaxReadLen);
	s.id = s.cigar + kCigarCap;
	s.frag[0] = reinterpret_cast<uint8_t *>(s.id + kIdCap);
	s.frag[1] = s.frag[0] + kMaxOrgLen;
	return s;
}

// ---------------------------------------------------------------------------------------------------
// small text helpers (lane-uniform; every lane computes the same lengths, lane 0 stores)
// ---------------------------------------------------------------------------------------------------
template<class G> RSQ_HD int put_uint(const G &g, char *dst, int pos, int cap, uint64_t v){
	int n = 1;
	if(v <= 0xffffffffull){
		uint32_t w = static_cast<uint32_t>(v);
		for(uint32_t t = w; t >= 10u; t /= 10u){ ++n; }
		if(g.lane() == 0){
			for(int i = n - 1; i >= 0; --i){ if(pos + i < cap){ dst[pos + i] = static_cast<char>('0' + w % 10u); } w /= 10u; }
		}
	}
	else{
		for(uint64_t t = v; t >= 10u; t /= 10u){ ++n; }
		if(g.lane() == 0){
			for(int i = n - 1; i >= 0; --i){ if(pos + i < cap){ dst[pos + i] = static_cast<char>('0' + v % 10u); } v /= 10u; }
		}
	}
	return pos + n;
}
template<class G> RSQ_HD int put_char(const G &g, char *dst, int pos, int cap, char c){
	if(g.lane() == 0 && pos < cap){ dst[pos] = c; }
	return pos + 1;
}
template<class G> RSQ_HD int put_str(const G &g, char *dst, int pos, int cap, const char *src, int n){
	if(g.lane() == 0){
		for(int i = 0; i < n && pos + i < cap; ++i){ dst[pos + i] = src[i]; }
	}
	return pos + n;
}

struct ReadState {               // Simulator::ReadFillParameter (Simulator.h:215-240) + cigar bookkeeping
	uint32_t read_length;
	uint32_t read_pos;
	uint32_t previous_indel_type;
	uint32_t indel_pos;
	uint32_t base_call;
	uint32_t gc_seq;
	uint32_t seq_qual;
	uint32_t qual;
	uint32_t error_rate;
	uint32_t num_errors;
	int cigar_len;               // bytes used in Scratch::cigar
};

template<class G> RSQ_HD void cigar_append(const G &g, const Scratch &s, ReadState &par, char op, uint32_t count){
	par.cigar_len = put_uint(g, s.cigar, par.cigar_len, kCigarCap, count);
	par.cigar_len = put_char(g, s.cigar, par.cigar_len, kCigarCap, op);
}

// Simulator::FillReadPart.  `org_len` bases are staged in s.org; their systematic errors in s.sdom / s.srate, or - for a read that touches
// variants - behind the cursor `w` (GetSysErrorFromBlock over the block's SysErrorVariants).
template<class G>
RSQ_HD void fill_read_part(const G &g, const SimCtx &c, const Scratch &s, Mt &mt, ReadState &par,
                           uint32_t seg, uint32_t tile, uint32_t org_pos, uint32_t org_len, char base_cigar_element,
                           const SysWalkCtx *wc = nullptr, SysWalk *w = nullptr, uint32_t allele = 0){
	uint32_t cigar_element_length = 0;
	char cigar_element = base_cigar_element;
	bool zero;
	while(par.read_pos < par.read_length && org_pos < org_len){
		const uint32_t ref_base = s.org[org_pos];
		double u = mt_uniform(g, mt);
		uint32_t indel = draw(g, c.tab, c.tab.indel(par.previous_indel_type, par.base_call), par.indel_pos, par.read_pos, par.gc_seq, 0, u, s.prob, zero);
		if(zero){ indel = 0; }

		if(indel == 0){
			uint32_t dom_error;
			if(w){ const uint32_t e = sysw_next(*wc, *w, allele); dom_error = e & 0xffu; par.error_rate = e >> 8; }
			else{ dom_error = s.sdom[org_pos]; par.error_rate = s.srate[org_pos]; }

			u = mt_uniform(g, mt);
			const uint32_t qtab = c.tab.quality(seg, tile, ref_base);
			uint32_t q = draw(g, c.tab, qtab, par.seq_qual, par.qual, par.read_pos, par.error_rate, u, s.prob, zero);
			if(zero){
				if(par.read_pos){ q = (s.qual[par.read_pos - 1] - c.phred_offset) & 0xffu; }
				else{ q = table_max_value(c.tab, qtab); }
			}
			par.qual = q & 0xffu;
			g.sync();
			if(g.lane() == 0){ s.qual[par.read_pos] = static_cast<uint8_t>(par.qual + c.phred_offset); }

			u = mt_uniform(g, mt);
			uint32_t call = draw(g, c.tab, c.tab.base_call(seg, tile, ref_base, dom_error), par.qual, par.read_pos, par.num_errors, par.error_rate, u, s.prob, zero);
			if(zero){ call = ref_base; }
			par.base_call = call;
			if(g.lane() == 0){ s.seq[par.read_pos] = static_cast<uint8_t>(call); }

			if(base_cigar_element == cigar_element){
				++cigar_element_length;
			}
			else{
				cigar_append(g, s, par, cigar_element, cigar_element_length);
				cigar_element = base_cigar_element;
				cigar_element_length = 1;
				par.indel_pos = 0;
				par.previous_indel_type = 0;
			}
			if(call != ref_base){ ++par.num_errors; }
			++par.read_pos;
			++org_pos;
		}
		else if(indel == 1){
			par.error_rate = w ? sysw_deletion(*wc, *w) : s.srate[org_pos];
			if('D' == cigar_element){
				++cigar_element_length;
				++par.indel_pos;
			}
			else{
				cigar_append(g, s, par, cigar_element, cigar_element_length);
				cigar_element = 'D';
				cigar_element_length = 1;
				par.indel_pos = 1;
				par.previous_indel_type = 1;
			}
			++par.num_errors;
			++org_pos;
		}
		else{
			u = mt_uniform(g, mt);
			uint32_t q = draw(g, c.tab, c.tab.quality(seg, tile, ref_base), par.seq_qual, par.qual, par.read_pos, par.error_rate, u, s.prob, zero);
			if(zero){ q = par.qual; }
			g.sync();
			if(g.lane() == 0){
				s.qual[par.read_pos] = static_cast<uint8_t>(c.phred_offset + q);
				s.seq[par.read_pos] = static_cast<uint8_t>(indel - 2);
			}
			if('I' == cigar_element){
				++cigar_element_length;
				++par.indel_pos;
			}
			else{
				cigar_append(g, s, par, cigar_element, cigar_element_length);
				cigar_element = 'I';
				cigar_element_length = 1;
				par.indel_pos = 1;
				par.previous_indel_type = 0;
			}
			++par.num_errors;
			++par.read_pos;
		}
	}
	if(cigar_element_length){
		cigar_append(g, s, par, cigar_element, cigar_element_length);
	}
}

// GeneralRandomDistributions::ReadLength (Simulator.h:185-198)
template<class G> RSQ_HD uint32_t draw_read_length(const G &g, const SimCtx &c, Mt &mt, uint32_t seg, uint32_t fragment_length){
	if(1 == c.read_len_count[seg]){ return c.read_len_from[seg]; }
	const double ins = (fragment_length < c.insert_to) ? static_cast<double>(c.insert_lengths[fragment_length]) : 0.0;
	const double random_value = mul_rn(mt_uniform(g, mt), ins);
	double counter = 0.0;
	uint32_t row_from = 0, row_n = 0;
	const uint64_t *vals = nullptr;
	if(fragment_length >= c.rlbf_from[seg] && fragment_length < c.rlbf_to[seg]){
		const uint32_t r = fragment_length - c.rlbf_from[seg];
		row_from = c.rlbf_row_from[seg][r];
		row_n = c.rlbf_row_off[seg][r + 1] - c.rlbf_row_off[seg][r];
		vals = c.rlbf_val[seg] + c.rlbf_row_off[seg][r];
	}
	uint32_t read_len = row_from + row_n;   // .to()
	while(counter <= random_value && (read_len-- > row_from)){
		counter = add_rn(counter, static_cast<double>(vals[read_len - row_from]));
	}
	return read_len & 0xffffu;
}

// Stage an adapter (bases + its systematic errors) as the current original sequence.
template<class G> RSQ_HD uint32_t stage_adapter(const G &g, const SimCtx &c, const Scratch &s, uint32_t seg, uint32_t adapter_id){
	const uint32_t off = c.adapters[seg].off[adapter_id];
	uint32_t len = c.adapters[seg].off[adapter_id + 1] - off;
	if(len > c.max_org_len){ len = c.max_org_len; if(g.lane() == 0){ *c.error_flag |= kErrOrgOverflow; } }
	g.sync();
	for(uint32_t i = g.lane(); i < len; i += G::kSize){
		s.org[i] = c.adapter_seq[off + i];
		s.sdom[i] = c.adapter_sys[2 * (off + i)];
		s.srate[i] = c.adapter_sys[2 * (off + i) + 1];
	}
	g.sync();
	return len;
}

// Simulator::FillRead.  On entry s.org/sdom/srate hold `org_len` bases of the fragment as this read sees it.
template<class G>
RSQ_HD void fill_read(const G &g, const SimCtx &c, const Scratch &s, Mt &mt, ReadState &par,
                      uint32_t seg, uint32_t tile, uint32_t fragment_length, uint32_t org_len,
                      const SysWalkCtx *wc = nullptr, const SysWalk *w_start = nullptr, uint32_t allele = 0){
	par.read_pos = 0; par.previous_indel_type = 0; par.indel_pos = 0; par.base_call = 5; par.gc_seq = 0;
	par.qual = 1; par.error_rate = 0; par.num_errors = 0; par.cigar_len = 0; par.seq_qual = 0;
	par.read_length = draw_read_length(g, c, mt, seg, fragment_length);
	if(par.read_length > c.max_read_len){ par.read_length = c.max_read_len; if(g.lane() == 0){ *c.error_flag |= kErrOrgOverflow; } }

	uint32_t adapter_id = 0;
	const uint32_t seq_length = par.read_length < org_len ? par.read_length : org_len;
	uint32_t mean_error_rate = 0;
	if(seq_length){
		uint32_t gc = 0, err = 0;
		for(uint32_t i = g.lane(); i < seq_length; i += G::kSize){
			const uint32_t b = s.org[i];
			gc += (b == 1 || b == 2) ? 1u : 0u;
			if(!w_start){ err += s.srate[i]; }
		}
		gc = g.reduce_add(gc);
		err = g.reduce_add(err);
		if(w_start){   // the cursor's own walk over the first seq_length bases (Simulator.cpp:483-500)
			SysWalk pre = *w_start;
			for(uint32_t i = 0; i < seq_length; ++i){ err += sysw_next(*wc, pre, allele) >> 8; }
		}
		par.gc_seq = percent_u16(gc, seq_length);
		mean_error_rate = divide_u32(err, seq_length);
	}
	else{
		adapter_id = discrete_draw(g, mt, c.adapters[seg].pick);
		const uint32_t alen = stage_adapter(g, c, s, seg, adapter_id);
		uint32_t gc = 0, err = 0;
		for(uint32_t i = g.lane(); i < alen; i += G::kSize){
			const uint32_t b = s.org[i];
			gc += (b == 1 || b == 2) ? 1u : 0u;
			err += s.srate[i];
		}
		gc = g.reduce_add(gc);
		err = g.reduce_add(err);
		par.gc_seq = percent_u16(gc, alen);
		mean_error_rate = divide_u32(err, alen);
	}

	bool zero;
	{
		const uint32_t tid = c.tab.seq_quality(seg, tile);
		const double u = mt_uniform(g, mt);
		uint32_t sq = draw(g, c.tab, tid, par.gc_seq, mean_error_rate, fragment_length / 10, 0, u, s.prob, zero);
		if(zero){ sq = table_most_likely(c.tab, tid); }
		par.seq_qual = sq & 0xffu;
	}

	{
		SysWalk w{};
		if(w_start){ w = *w_start; }
		fill_read_part(g, c, s, mt, par, seg, tile, 0, org_len, 'M', wc, w_start ? &w : nullptr, allele);
	}

	if(par.read_pos < par.read_length){
		if(0 == adapter_id){
			adapter_id = discrete_draw(g, mt, c.adapters[seg].pick);
		}
		uint32_t adapter_pos = 0;
		if(0 == par.read_pos){
			adapter_pos = discrete_draw(g, mt, c.adapters[seg].start_cut[adapter_id]) + c.adapters[seg].start_cut_from[adapter_id];
		}
		const uint32_t alen = stage_adapter(g, c, s, seg, adapter_id);
		fill_read_part(g, c, s, mt, par, seg, tile, adapter_pos, alen, 'S');

		if(par.read_pos < par.read_length){
			cigar_append(g, s, par, 'H', par.read_length - par.read_pos);
			const uint32_t qtab = c.tab.quality(seg, tile, 0);
			const uint32_t tail_length = (discrete_draw(g, mt, c.polya_pick) + c.polya_from) & 0xffffu;
			for(uint32_t pos_tail = 0; pos_tail < tail_length && par.read_pos < par.read_length; ++pos_tail){
				const double u = mt_uniform(g, mt);
				uint32_t q = draw(g, c.tab, qtab, par.seq_qual, par.qual, par.read_pos, par.error_rate, u, s.prob, zero);
				g.sync();
				if(zero){ q = (s.qual[par.read_pos - 1] - c.phred_offset) & 0xffu; }
				par.qual = q & 0xffu;
				if(g.lane() == 0){
					s.qual[par.read_pos] = static_cast<uint8_t>(par.qual + c.phred_offset);
					s.seq[par.read_pos] = 0;
				}
				++par.read_pos;
			}
			while(par.read_pos < par.read_length){
				const double u = mt_uniform(g, mt);
				uint32_t q = draw(g, c.tab, qtab, par.seq_qual, par.qual, par.read_pos, par.error_rate, u, s.prob, zero);
				g.sync();
				if(zero){ q = (s.qual[par.read_pos - 1] - c.phred_offset) & 0xffu; }
				par.qual = q & 0xffu;
				const uint32_t b = discrete_draw(g, mt, c.overrun_pick);
				if(g.lane() == 0){
					s.qual[par.read_pos] = static_cast<uint8_t>(par.qual + c.phred_offset);
					s.seq[par.read_pos] = static_cast<uint8_t>(b);
				}
				++par.read_pos;
			}
		}
	}
	if(par.cigar_len > kCigarCap && g.lane() == 0){ *c.error_flag |= kErrCigarOverflow; }
	g.sync();
}

// FragmentDistributionStats::Binomial with the first term pow(1-p, N) supplied by the host
RSQ_HD uint32_t binomial_count(uint32_t N, double p, double pow_term, double probability_chosen){
	double probability_count = pow_term;
	double probability_left = sub_rn(probability_chosen, probability_count);
	uint32_t count = 0;
	const double one_minus_p = sub_rn(1.0, p);
	while(0.0 < probability_left && count < N){
		++count;
		const double f = mul_rn(static_cast<double>(static_cast<int>(N + 1 - count)) / static_cast<double>(static_cast<int>(count)), p) / one_minus_p;
		probability_count = mul_rn(probability_count, f);
		probability_left = sub_rn(probability_left, probability_count);
	}
	return count;
}

// FragmentDistributionStats::GetFragmentCounts + NegativeBinomial + BiasCalculationVectors::GetDispersion
RSQ_HD uint32_t fragment_counts(const SimCtx &c, uint32_t ref_id, uint32_t fragment_length, uint32_t gc, double sur_start, double sur_end,
                                double probability_chosen, bool &runaway){
	double bias = mul_rn(c.ref_seq_bias[ref_id], c.il_bias[fragment_length]);   // Reference::Bias (Reference.h:167-169, 283-285)
	bias = mul_rn(bias, c.gc_bias[gc]);
	bias = mul_rn(bias, sur_start);
	bias = mul_rn(bias, sur_end);
	if(!(0.0 < bias)){ return 0; }
	double mean = mul_rn(bias, c.bias_normalization);
	double r = mean / add_rn(c.disp_a, mul_rn(c.disp_b, mean));
	const double cap = mul_rn(mean, 1e10);
	if(r > cap){ r = cap; }
	// num_alleles == 1: the two divisions by the allele count are exact
	const double p = mean / add_rn(mean, r);
	double probability_count = pow_glibc(sub_rn(1.0, p), r);
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

// GetFragmentCounts with Reference::NumAlleles() alleles (runs with variants): mean and dispersion are divided by the allele count
RSQ_HD uint32_t fragment_counts_alleles(const SimCtx &c, uint32_t ref_id, uint32_t fragment_length, uint32_t gc, double sur_start, double sur_end,
                                        double probability_chosen, uint32_t alleles, bool &runaway){
	double bias = mul_rn(c.ref_seq_bias[ref_id], c.il_bias[fragment_length]);
	bias = mul_rn(bias, c.gc_bias[gc]);
	bias = mul_rn(bias, sur_start);
	bias = mul_rn(bias, sur_end);
	if(!(0.0 < bias)){ return 0; }
	return allele_fragment_counts(mul_rn(bias, c.bias_normalization), c.disp_a, c.disp_b, alleles, probability_chosen, runaway);
}

// One chosen (allele, strand) of a hit in a run with variants (SimulateFromGivenBlock, Simulator.cpp:2311-2330): end position, GC and
// surroundings of the allele's fragment, then the count draw.  Hits with no variant of any allele within reach of the fragment and its
// surroundings take the reference's own per-position arrays; the others are evaluated on the allele's sequence (allele_hit).
struct VarEval {
	uint32_t allele, end_position, counts;
	uint32_t slow;                               // bit 0: the read on the forward strand walks variants, bit 1: the read on the reverse strand does
	int32_t end_var; uint32_t end_var_pos;       // VariantBiasVarModifiers::EndVariant (when the fragment was evaluated on the allele's sequence)
};
// Geometry of one chosen allele's fragment (no draw involved): end position, GC, both surrounding biases, EndVariant, which reads walk variants.
// Every part falls back to the reference's own arrays when no variant of any allele lies within its reach: the fragment itself (end position, GC),
// each of the two surroundings, each of the two reads.  Kept out of line: it is rare next to the scan's draws and must not cost it registers.
struct VarGeom { uint32_t valid, gc_perc; double sur_start, sur_end; AllelePoint end; uint32_t end_hint; };
// what the out-of-line evaluation reads of the run (passed by value: a reference to the kernel's SimCtx would force the whole struct into local memory)
struct VarGeomCtx { const uint8_t *seq; const uint32_t *gcp; const double *sur_start, *sur_end; const double *sur_tab0, *sur_tab1, *sur_tab2; uint32_t L, n_read_max, probe_plain; };
RSQ_HD_VCOLD void eval_allele_geometry(const VarGeomCtx gc, const VariantView v, uint32_t pos, uint32_t first_var, uint32_t start_variant_pos, uint32_t fl,
                                      uint32_t allele, VarEval &e, VarGeom &geo){
	const uint32_t L = gc.L;
	const uint32_t *gcp = gc.gcp;
	e.allele = allele; e.counts = 0; e.end_var = -1; e.end_var_pos = 0; e.slow = 0;
	geo.valid = 0;
	const uint32_t hint = var_seek(v, first_var, pos);   // first variant at or behind the start position
	const bool probe_plain = gc.probe_plain != 0u;   // timing probe only (RSQ_VAR_PROBE=plain): every hit takes the reference's arrays - wrong output
	const bool in_fragment = !probe_plain && (start_variant_pos || (hint < v.n && v.position[hint] <= pos + fl));
	geo.end_hint = hint;
	geo.end = AllelePoint{pos + fl, 0, -1};
	bool only_substitutions = in_fragment && !start_variant_pos;
	uint32_t gc_delta_plus = 0, gc_delta_minus = 0;
	if(only_substitutions){
		// fragments whose allele carries nothing but substitutions keep the reference's coordinates: end position pos + fl, G/C count corrected per base
		for(uint32_t t = hint; t < v.n && v.position[t] < pos + fl + 1u && only_substitutions; ++t){
			if(!v.in_allele(t, allele)){ continue; }
			if(v.length(t) != 1u){ only_substitutions = false; break; }
			if(v.position[t] < pos + fl){
				const uint32_t alt = v.base(t, 0), org = gc.seq[v.position[t]];
				gc_delta_plus += (alt == 1u || alt == 2u) ? 1u : 0u;
				gc_delta_minus += (org == 1u || org == 2u) ? 1u : 0u;
			}
		}
	}
	if(only_substitutions){
		e.end_position = pos + fl;
		if(!(e.end_position < L)){ return; }
		const uint32_t gcn = gcp[e.end_position] - gcp[pos] + gc_delta_plus - gc_delta_minus;
		geo.gc_perc = ((gcn * 100u + fl / 2u) / fl) & 0xffu;
		geo.end_hint = var_seek(v, hint, e.end_position);
		e.end_var = static_cast<int32_t>(geo.end_hint) - 1;
	}
	else if(in_fragment){
		AlleleHit h;
		allele_hit(v, gcp, L, allele, pos, first_var, start_variant_pos, fl, h);
		if(!h.valid){ return; }
		e.end_position = h.end_position; e.end_var = h.end_var; e.end_var_pos = h.end_var_pos;
		geo.gc_perc = h.gc_percent; geo.end = h.end; geo.end_hint = h.end_hint;
	}
	else{
		e.end_position = pos + fl;
		if(!(e.end_position < L)){ return; }
		geo.gc_perc = percent_u32(gcp[e.end_position] - gcp[pos], fl);
		e.end_var = static_cast<int32_t>(hint) - 1;
	}
	geo.valid = 1;
	const uint32_t ep = e.end_position;
	if(!probe_plain && (start_variant_pos || var_in_range(v, hint, pos >= 11u ? pos - 11u : 0u, pos + 20u))){
		uint32_t code[3];
		allele_start_surrounding(v, gc.seq, L, allele, pos, first_var, start_variant_pos, code);
		geo.sur_start = surrounding_bias(gc.sur_tab0, gc.sur_tab1, gc.sur_tab2, code);
	}
	else{ geo.sur_start = gc.sur_start[pos]; }
	if(!probe_plain && (geo.end.k || var_in_range(v, geo.end_hint, ep >= 22u ? ep - 22u : 0u, ep + 11u))){
		uint32_t code[3];
		allele_end_surrounding(v, gc.seq, L, allele, geo.end, code, geo.end_hint);
		geo.sur_end = surrounding_bias(gc.sur_tab0, gc.sur_tab1, gc.sur_tab2, code);
	}
	else{ geo.sur_end = gc.sur_end[ep - 1]; }
	if(!probe_plain){
		// which of the two reads has to walk variants (spliced bases, SysErrorVariant cursor): any variant of any allele within the read's bases
		const uint32_t n = fl + 2u < gc.n_read_max ? fl + 2u : gc.n_read_max;
		if(start_variant_pos || var_in_range(v, hint, pos, pos + n)){ e.slow |= 1u; }
		if(geo.end.k || e.end_var_pos || var_in_range(v, geo.end_hint, ep >= n ? ep - n : 0u, ep)){ e.slow |= 2u; }
	}
}
// uniform(): the next ZeroToOne of the block's stream.  Returns false when the fragment does not end inside the sequence (no draw is consumed).
template<class Uniform>
RSQ_HD bool eval_allele_hit(const SimCtx &c, const VariantView &v, uint32_t ref_id, uint32_t pos, uint32_t first_var, uint32_t start_variant_pos, uint32_t fl,
                            uint32_t allele, double thr0, Uniform &&uniform, VarEval &e, VarGeom &geo, bool &runaway){
	const uint64_t off = c.seq_off[ref_id];
	// nearly two thirds of the hits have no variant of any allele within reach of the fragment and its surroundings: the reference's arrays, no walk.
	// Both cases end in ONE draw + count evaluation (the Binomial / pow code exists once in the kernel: the variant-aware scan is bound by instruction fetches)
	const uint32_t hint = var_seek(v, first_var, pos);
	const bool near = start_variant_pos || (c.var.loaded & 2u) == 0u && ((hint < v.n && v.position[hint] <= pos + fl + 12u) || (hint > 0 && v.position[hint - 1] + 12u >= pos));
	uint32_t gc_perc;
	double sur_s, sur_e;
	if(!near){
		const uint32_t L = c.seq_len[ref_id];
		const uint32_t *gcp = c.gc_prefix + off + ref_id;
		e.allele = allele; e.counts = 0; e.end_var = static_cast<int32_t>(hint) - 1; e.end_var_pos = 0; e.slow = 0;
		e.end_position = pos + fl;
		geo.valid = 0; geo.end = AllelePoint{pos + fl, 0, -1}; geo.end_hint = hint;
		if(!(e.end_position < L)){ return false; }
		geo.valid = 1;
		gc_perc = percent_u32(gcp[e.end_position] - gcp[pos], fl); sur_s = c.sur_start[off + pos]; sur_e = c.sur_end[off + e.end_position - 1];
	}
	else{
		VarGeomCtx gc;
		gc.seq = c.ref + off; gc.gcp = c.gc_prefix + off + ref_id; gc.sur_start = c.sur_start + off; gc.sur_end = c.sur_end + off;
		gc.sur_tab0 = c.var.sur_tab[0]; gc.sur_tab1 = c.var.sur_tab[1]; gc.sur_tab2 = c.var.sur_tab[2];
		gc.L = c.seq_len[ref_id]; gc.n_read_max = (c.read_len_to[0] > c.read_len_to[1] ? c.read_len_to[0] : c.read_len_to[1]) + c.max_len_deletion + 2u;
		gc.probe_plain = (c.var.loaded & 2u) ? 1u : 0u;
		eval_allele_geometry(gc, v, pos, first_var, start_variant_pos, fl, allele, e, geo);
		if(!geo.valid){ return false; }
		gc_perc = geo.gc_perc; sur_s = geo.sur_start; sur_e = geo.sur_end;
	}
	const double rv = uniform();
	const double adjusted_random = add_rn(thr0, mul_rn(rv, sub_rn(1.0, thr0)));
	e.counts = fragment_counts_alleles(c, ref_id, fl, gc_perc, sur_s, sur_e, adjusted_random, c.var.num_alleles, runaway);
	if(!e.counts){ e.slow = 0; }
	return true;
}
// GetOrgSeq with variants (Simulator.cpp:1909-1914): the two ends of the allele's fragment - the allele's bases behind its start (StartVariant) and,
// reverse complemented, in front of its end (EndVariant) - written by the whole lane group.  An end whose read has no variant among its bases
// is copied from the reference.
template<class G>
RSQ_HD void splice_fragment_ends(const G &g, const SimCtx &c, const VariantView &v, uint32_t ref_id, uint32_t strand, uint32_t pos, uint32_t first_var, uint32_t start_variant_pos,
                                 uint32_t fl, const VarEval &e, const VarGeom &geo, uint8_t *frag_fwd, uint8_t *frag_rev, uint32_t which = 3u /* bit 0 forward end, bit 1 reverse end */){
	g.sync();
	const uint8_t *seq = c.ref + c.seq_off[ref_id];
	const uint32_t L = c.seq_len[ref_id];
	uint32_t n_fwd = c.read_len_to[strand ? 1 : 0] + c.max_len_deletion;   // forward end: the read of segment `strand`
	if(fl < n_fwd){ n_fwd = fl; }
	if(n_fwd > c.max_org_len){ n_fwd = c.max_org_len; }
	uint32_t n_rev = c.read_len_to[strand ? 0 : 1] + c.max_len_deletion;
	if(fl < n_rev){ n_rev = fl; }
	if(n_rev > c.max_org_len){ n_rev = c.max_org_len; }
	if((which & 1u) && (e.slow & 1u)){
		allele_bases_forward_g(g, v, seq, L, e.allele, AllelePoint{pos, start_variant_pos, start_variant_pos ? static_cast<int32_t>(first_var) : -1}, n_fwd, frag_fwd, first_var);
	}
	else if(which & 1u){ for(uint32_t i = g.lane(); i < n_fwd; i += G::kSize){ frag_fwd[i] = seq[pos + i]; } }
	if((which & 2u) && (e.slow & 2u)){
		// EndVariant {id, posCurrentlyAt}: inside an insertion the point lies behind its first posCurrentlyAt bases, else in front of the end position
		const AllelePoint at = e.end_var_pos ? AllelePoint{e.end_position - 1u, e.end_var_pos, e.end_var} : AllelePoint{e.end_position, 0, -1};
		allele_bases_backward_g(g, v, seq, L, e.allele, at, n_rev, frag_rev, geo.end_hint, true);
	}
	else if(which & 2u){ for(uint32_t i = g.lane(); i < n_rev; i += G::kSize){ frag_rev[i] = static_cast<uint8_t>(3u - seq[e.end_position - 1u - i]); } }
	g.sync();
}

// Simulator::CreateReadId
template<class G>
RSQ_HD int format_read_id(const G &g, const SimCtx &c, const Scratch &s, uint32_t block_number, uint64_t read_number,
                          uint32_t start_pos, uint32_t end_pos, uint32_t tile, uint32_t ref_id, const ReadState &par, uint32_t allele = 0){
	int n = 0;
	n = put_str(g, s.id, n, kIdCap, c.base_id, c.base_id_len);
	n = put_uint(g, s.id, n, kIdCap, block_number);
	n = put_char(g, s.id, n, kIdCap, '_');
	n = put_uint(g, s.id, n, kIdCap, read_number);
	if(start_pos && 1 < c.var.num_alleles && c.var.loaded){   // Simulator.cpp:612
		n = put_str(g, s.id, n, kIdCap, "_allele", 7);
		n = put_uint(g, s.id, n, kIdCap, allele);
	}
	n = put_char(g, s.id, n, kIdCap, ':');
	n = put_uint(g, s.id, n, kIdCap, start_pos);
	n = put_char(g, s.id, n, kIdCap, ':');
	if(start_pos){
		n = put_str(g, s.id, n, kIdCap, c.name_blob + c.name_off[ref_id], c.name_off[ref_id + 1] - c.name_off[ref_id]);
	}
	else{
		n = put_str(g, s.id, n, kIdCap, "Adapter", 7);
	}
	n = put_char(g, s.id, n, kIdCap, ':');
	n = put_uint(g, s.id, n, kIdCap, end_pos);
	n = put_char(g, s.id, n, kIdCap, ':');
	n = put_uint(g, s.id, n, kIdCap, tile);
	n = put_str(g, s.id, n, kIdCap, ":1337:1337 ", 11);
	g.sync();
	n = put_str(g, s.id, n, kIdCap, s.cigar, par.cigar_len < kCigarCap ? par.cigar_len : kCigarCap);
	n = put_str(g, s.id, n, kIdCap, " E", 2);
	n = put_uint(g, s.id, n, kIdCap, par.num_errors);
	if(n > kIdCap){ if(g.lane() == 0){ *c.error_flag |= kErrRecordTooLong; } n = kIdCap; }
	g.sync();
	return n;
}

// Stage the original sequence + systematic errors of one read of a fragment (Simulator::GetOrgSeq without
// variants + the block/partner-block lookup of CreateReads): forward read = prefix of the fragment on the
// forward strand, reverse read = reverse complement of its suffix with the reverse-strand errors.
template<class G>
RSQ_HD uint32_t stage_fragment_read(const G &g, const SimCtx &c, const Scratch &s, uint32_t ref_id, uint32_t seg, bool reversed,
                                    uint32_t start_pos, uint32_t end_pos, uint32_t fragment_length, const uint8_t *converted = nullptr){
	uint32_t org_len = c.read_len_to[seg] + c.max_len_deletion;
	if(fragment_length < org_len){ org_len = fragment_length; }
	if(org_len > c.max_org_len){ org_len = c.max_org_len; if(g.lane() == 0){ *c.error_flag |= kErrOrgOverflow; } }
	const uint64_t off = c.seq_off[ref_id];
	const uint32_t L = c.seq_len[ref_id];
	g.sync();
	if(!reversed){
		const uint8_t *ref = c.ref + off + start_pos;
		const uint8_t *sys = c.sys_fwd + 2 * (off + start_pos);
		for(uint32_t i = g.lane(); i < org_len; i += G::kSize){
			s.org[i] = converted ? converted[i] : ref[i];
			s.sdom[i] = sys[2 * i];
			s.srate[i] = sys[2 * i + 1];
		}
	}
	else{
		const uint8_t *ref = c.ref + off;
		const uint8_t *sys = c.sys_rev + 2 * (off + (L - end_pos));
		for(uint32_t i = g.lane(); i < org_len; i += G::kSize){
			s.org[i] = converted ? converted[i] : static_cast<uint8_t>(3 - ref[end_pos - 1 - i]);
			s.sdom[i] = sys[2 * i];
			s.srate[i] = sys[2 * i + 1];
		}
	}
	g.sync();
	return org_len;
}

// Simulator::CreateReads for `counts` copies of one fragment (start_block != NULL case), or for the
// adapter-only pairs (fragment_length == 0).
struct VarRead { uint32_t allele, slow, first_var, start_variant_pos; int32_t end_var; uint32_t end_var_pos; uint32_t start_block; };   // a hit's allele + StartVariant / EndVariant
template<class G, class Sink>
RSQ_HD void create_reads(const G &g, const SimCtx &c, const Scratch &s, Mt &mt, Sink &sink, uint32_t counts, bool strand,
                         uint32_t ref_id, uint32_t fragment_length, uint64_t &read_number, uint32_t block_id,
                         uint32_t start_position_forward, uint32_t end_position_forward, bool converted = false, const VarRead *vr = nullptr){
	uint32_t print_start = 0, print_end = 0;
	if(fragment_length){
		if(strand){ print_start = end_position_forward; print_end = start_position_forward + 1; }
		else{ print_start = start_position_forward + 1; print_end = end_position_forward; }
	}
	for(uint32_t n = counts; n--; ){
		++read_number;
		uint32_t tile = 0;
		if(1 < c.num_tiles){ tile = discrete_draw(g, mt, c.tile_pick); }
		for(uint32_t seg = 2; seg--; ){
			uint32_t org_len = 0;
			if(fragment_length){
				// block.at(strand) is the forward start block, block.at(!strand) the reverse partner of the end block
				const bool reversed = (seg != static_cast<uint32_t>(strand));
				org_len = stage_fragment_read(g, c, s, ref_id, seg, reversed, start_position_forward, end_position_forward, fragment_length, converted ? s.frag[reversed ? 1 : 0] : nullptr);
			}
			ReadState par;
			if(vr && fragment_length && (vr->slow & ((seg != static_cast<uint32_t>(strand)) ? 2u : 1u))){
				// a read with variants among its bases walks the SimBlocks' SysErrorVariants (CreateReads, Simulator.cpp:680-689)
				const bool reversed = (seg != static_cast<uint32_t>(strand));
				SysWalkCtx wc{};
				wc.sys = (reversed ? c.sys_rev : c.sys_fwd) + 2 * c.seq_off[ref_id]; wc.errs = reversed ? c.var.errs_rev : c.var.errs_fwd;
				wc.block_first = c.var.block_first + c.var.block_first_off[ref_id]; wc.v = c.var.view(ref_id); wc.L = c.seq_len[ref_id]; wc.reverse = reversed ? 1u : 0u;
				const SysWalk w = reversed ? sysw_reverse_start(wc, vr->start_block, end_position_forward, vr->end_var, vr->end_var_pos)
				                           : sysw_forward_start(wc, vr->start_block, start_position_forward, vr->first_var, vr->start_variant_pos);
				fill_read(g, c, s, mt, par, seg, tile, fragment_length, org_len, &wc, &w, vr->allele);
			}
			else{ fill_read(g, c, s, mt, par, seg, tile, fragment_length, org_len); }
			const int id_len = format_read_id(g, c, s, block_id, read_number, print_start, print_end, c.tile_names[tile], ref_id, par, vr ? vr->allele : 0u);
			sink.write_record(g, seg, s.id, id_len, s.seq, s.qual, par.read_length);
		}
		sink.pair_done(g);
	}
}

// Simulator::CTConversion without variants (Simulator.cpp:1925-2003): bisulfite C->T on one fragment end, one draw per
// C inside an unmethylated region.  `read` holds `read_len` bases; regions/rates are those of the sequence.
struct MtSource {                 // the block's stream as the serial kernel holds it
	Mt &mt;
	template<class G> RSQ_HD double next(const G &g){ return mt_uniform(g, mt); }
};
template<class G, class Rng>
RSQ_HD void ct_conversion(const G &g, const SimCtx &c, Rng &rng, uint8_t *read, uint32_t read_len, uint32_t seq_id, uint32_t start_pos,
                          int32_t cur_methylation_start, bool reversed){
	const uint32_t r0 = c.meth_off[seq_id];
	const uint32_t n_regions = c.meth_off[seq_id + 1] - r0;
	const uint32_t *first = c.meth_start + r0, *second = c.meth_end + r0;
	const double *rate = c.meth_rate + r0;
	int32_t cur_meth = cur_methylation_start;
	uint32_t read_pos = 0;   // uintReadLen
	uint32_t ref_pos = start_pos;
	auto in_range = [&](int32_t i) -> bool {   // regions.at(i) of the reference throws otherwise
		if(static_cast<uint32_t>(i) < n_regions){ return true; }
		if(g.lane() == 0){ *c.error_flag |= kErrReferenceOutOfRange; }
		return false;
	};
	auto convert = [&](){
		if(1 == read[read_pos]){
			const double u = rng.next(g);
			if(u < rate[cur_meth]){
				g.sync();
				if(g.lane() == 0){ read[read_pos] = 3; }
				g.sync();
			}
		}
	};
	if(reversed){
		while(static_cast<uint32_t>(cur_meth) < n_regions && cur_meth >= 0 && first[cur_meth] <= ref_pos){ ++cur_meth; }
		--cur_meth;
		if(cur_meth){
			if(!in_range(cur_meth)){ return; }
			if(second[cur_meth] <= ref_pos){
				read_pos = (read_pos + ref_pos - (second[cur_meth] - 1)) & 0xffffu;
				ref_pos = second[cur_meth] - 1;
			}
		}
		while(cur_meth && read_pos < read_len){
			if(!in_range(cur_meth)){ return; }
			while(ref_pos >= first[cur_meth] && read_pos < read_len){
				convert();
				--ref_pos;
				read_pos = (read_pos + 1) & 0xffffu;
			}
			if(--cur_meth){
				if(!in_range(cur_meth)){ return; }
				if(second[cur_meth] <= ref_pos){
					read_pos = (read_pos + ref_pos - (second[cur_meth] - 1)) & 0xffffu;
					ref_pos = second[cur_meth] - 1;
				}
			}
		}
	}
	else{
		if(cur_meth >= 0 && static_cast<uint32_t>(cur_meth) < n_regions && first[cur_meth] > ref_pos){
			read_pos = (read_pos + first[cur_meth] - ref_pos) & 0xffffu;
			ref_pos = first[cur_meth];
		}
		while(cur_meth >= 0 && static_cast<uint32_t>(cur_meth) < n_regions && read_pos < read_len){
			while(ref_pos < second[cur_meth] && read_pos < read_len){
				convert();
				++ref_pos;
				read_pos = (read_pos + 1) & 0xffffu;
			}
			if(static_cast<uint32_t>(++cur_meth) < n_regions){
				read_pos = (read_pos + first[cur_meth] - ref_pos) & 0xffffu;
				ref_pos = first[cur_meth];
			}
		}
	}
}

// Simulator::CTConversion, variant overload (Simulator.cpp:2004-2217): the same conversion with the walk over the reference kept in step with the
// allele's variants (insertions stand at one reference position, deletions skip one).  first_variant / first_variant_pos: StartVariant (forward
// end) or EndVariant (reverse end).  The reference reads its local `deletion` before ever writing it; in its build here the flag starts out false.
template<class G, class Rng>
RSQ_HD void ct_conversion_var(const G &g, const SimCtx &c, Rng &rng, uint8_t *read, uint32_t read_len, uint32_t seq_id, uint32_t start_pos, uint32_t allele,
                              int32_t cur_methylation_start, bool reversed, const VariantView &v, int32_t first_variant, uint32_t first_variant_pos){
	const uint32_t r0 = c.meth_off[seq_id];
	const int64_t n_regions = c.meth_off[seq_id + 1] - r0;
	const uint32_t *first = c.meth_start + r0, *second = c.meth_end + r0;
	const double *rate = c.meth_rate + r0 + (c.meth_alleles > 1 ? static_cast<size_t>(allele) * c.meth_rate_stride : 0u);
	int64_t cur_meth = cur_methylation_start;
	uint32_t read_pos = 0;   // uintReadLen
	uint32_t ref_pos = start_pos;
	int64_t cur_var = first_variant;
	const int64_t n_var = v.n;
	uint32_t var_bases_left = 0;
	bool deletion = false;
	auto var_pos = [&](int64_t i) -> uint32_t { return v.position[i]; };
	auto var_len = [&](int64_t i) -> uint32_t { return v.length(static_cast<uint32_t>(i)); };
	auto in_allele = [&](int64_t i) -> bool { return v.in_allele(static_cast<uint32_t>(i), allele); };
	auto convert = [&](){
		if(1 == read[read_pos]){
			const double u = rng.next(g);
			if(u < rate[cur_meth]){
				g.sync();
				if(g.lane() == 0){ read[read_pos] = 3; }
				g.sync();
			}
		}
	};
	auto add = [&](uint32_t n){ read_pos = (read_pos + n) & 0xffffu; };
	if(reversed){
		if(0 <= cur_var && cur_var < n_var && var_pos(cur_var) == ref_pos && 1 < var_len(cur_var) && in_allele(cur_var)){ var_bases_left = var_len(cur_var) - first_variant_pos; }
		while(cur_meth < n_regions && cur_meth >= 0 && first[cur_meth] <= ref_pos){ ++cur_meth; }
		--cur_meth;
		if(0 <= cur_meth && cur_meth < n_regions && second[cur_meth] <= ref_pos){
			if(var_bases_left){ add(var_bases_left); var_bases_left = 0; --ref_pos; --cur_var; }
			while(0 <= cur_var && second[cur_meth] <= var_pos(cur_var) && read_pos < read_len){
				if(in_allele(cur_var)){
					add(ref_pos - var_pos(cur_var));
					add(var_len(cur_var));
					ref_pos = var_pos(cur_var) - 1u;
				}
				--cur_var;
			}
			add(ref_pos - (second[cur_meth] - 1u));
			ref_pos = second[cur_meth] - 1u;
		}
		while(0 <= cur_meth && cur_meth < n_regions && read_pos < read_len){
			while(ref_pos >= first[cur_meth] && ref_pos != 0xffffffffu && read_pos < read_len){
				if(0 == var_bases_left){
					while(0 <= cur_var && var_pos(cur_var) == ref_pos && !in_allele(cur_var)){ --cur_var; }
					if(0 <= cur_var && var_pos(cur_var) == ref_pos){
						if(0 == var_len(cur_var)){ deletion = true; --cur_var; }
						else{ var_bases_left = var_len(cur_var); }
					}
				}
				if(deletion){ deletion = false; }
				else{ convert(); }
				if(var_bases_left){ if(0 == --var_bases_left){ --cur_var; } }
				if(0 == var_bases_left){ --ref_pos; }
				add(1);
			}
			if(0 <= --cur_meth && second[cur_meth] <= ref_pos){
				while(0 <= cur_var && second[cur_meth] <= var_pos(cur_var) && read_pos < read_len){
					if(in_allele(cur_var)){
						add(ref_pos - var_pos(cur_var));
						add(var_len(cur_var));
						ref_pos = var_pos(cur_var) - 1u;
					}
					--cur_var;
				}
				add(ref_pos - (second[cur_meth] - 1u));
				ref_pos = second[cur_meth] - 1u;
			}
		}
	}
	else{
		if(cur_var >= 0 && cur_var < n_var && var_pos(cur_var) == ref_pos && 1 < var_len(cur_var) && in_allele(cur_var)){ var_bases_left = var_len(cur_var) - first_variant_pos; }
		if(cur_meth >= 0 && cur_meth < n_regions && first[cur_meth] > ref_pos){
			if(var_bases_left){ add(var_bases_left); var_bases_left = 0; ++ref_pos; ++cur_var; }
			while(cur_var < n_var && first[cur_meth] > var_pos(cur_var) && read_pos < read_len){
				if(in_allele(cur_var)){
					add(var_pos(cur_var) - ref_pos);
					add(var_len(cur_var));
					ref_pos = var_pos(cur_var) + 1u;
				}
				++cur_var;
			}
			add(first[cur_meth] - ref_pos);
			ref_pos = first[cur_meth];
		}
		while(cur_meth >= 0 && cur_meth < n_regions && read_pos < read_len){
			while(ref_pos < second[cur_meth] && read_pos < read_len){
				if(0 == var_bases_left){
					while(cur_var < n_var && var_pos(cur_var) == ref_pos && !in_allele(cur_var)){ ++cur_var; }
					if(cur_var < n_var && var_pos(cur_var) == ref_pos){
						if(0 == var_len(cur_var)){ deletion = true; ++cur_var; }
						else{ var_bases_left = var_len(cur_var); }
					}
				}
				if(deletion){ deletion = false; }
				else{ convert(); }
				if(var_bases_left){ if(0 == --var_bases_left){ ++cur_var; } }
				if(0 == var_bases_left){ ++ref_pos; }
				add(1);
			}
			if(++cur_meth < n_regions && first[cur_meth] > ref_pos){
				while(cur_var < n_var && first[cur_meth] > var_pos(cur_var) && read_pos < read_len){
					if(in_allele(cur_var)){
						add(var_pos(cur_var) - ref_pos);
						add(var_len(cur_var));
						ref_pos = var_pos(cur_var) + 1u;
					}
					++cur_var;
				}
				add(first[cur_meth] - ref_pos);
				ref_pos = first[cur_meth];
			}
		}
	}
}

// Simulator::SimulateFromGivenBlock without variants / methylation
template<bool kMeth, class G, class Sink>
RSQ_HD void simulate_block(const G &g, const SimCtx &c, const Scratch &s, Sink &sink, const BlockDesc &b,
                           unsigned long long *scan_draws){
	Mt mt; mt.s = s.mt; mt.idx = kMtN;
	mt_seed(g, mt, b.seed);
	const uint32_t L = c.seq_len[b.ref_id];
	const uint64_t off = c.seq_off[b.ref_id];
	const uint32_t group = c.coverage_group[b.ref_id];
	const double *thr = c.thr + static_cast<size_t>(group) * c.insert_to * 2;
	const uint64_t *thr_int = c.thr_int + static_cast<size_t>(group) * c.insert_to;
	const double *binom_p0 = c.binom_p0 + static_cast<size_t>(group) * c.insert_to;
	const uint32_t *gcp = c.gc_prefix + off + b.ref_id;
	uint64_t read_number = 0;
	unsigned long long draws = 0;
	uint32_t end = b.start_pos + 1000u;
	if(end > L){ end = L; }
	int32_t cur_methylation_start = b.first_meth;
	for(uint32_t pos = b.start_pos; pos < end; ++pos){
		if(kMeth){
			const uint32_t r0 = c.meth_off[b.ref_id], nr = c.meth_off[b.ref_id + 1] - r0;
			if(cur_methylation_start >= 0 && static_cast<uint32_t>(cur_methylation_start) < nr && c.meth_end[r0 + cur_methylation_start] <= pos){ ++cur_methylation_start; }
		}
		uint32_t len = c.insert_from;
		while(len < c.insert_to){
			if(mt.idx >= kMtN){ mt_regen(g, mt); }
			uint32_t n = c.insert_to - len;
			if(n > static_cast<uint32_t>(G::kSize)){ n = G::kSize; }
			if(n > static_cast<uint32_t>(kMtN - mt.idx)){ n = kMtN - mt.idx; }
			const uint32_t lane = g.lane();
			uint64_t x = 0;
			bool hit = false;
			if(lane < n){
				x = mt_temper(mt.s[mt.idx + lane]);
				hit = x >= thr_int[len + lane];
			}
			const unsigned mask = g.ballot(hit);
			if(mask == 0){
				mt.idx += n; len += n; draws += n;
				continue;
			}
#if defined(__CUDA_ARCH__)
			const uint32_t first = __ffs(mask) - 1;
#else
			const uint32_t first = 0;
#endif
			const uint32_t fragment_length = len + first;
			x = mt_temper(mt.s[mt.idx + first]);
			mt.idx += first + 1; len = fragment_length + 1; draws += first + 1;

			const double probability_chosen = canonical(x);
			const double thr0 = thr[2 * fragment_length], thr1 = thr[2 * fragment_length + 1];
			if(!(probability_chosen >= thr1)){ continue; }   // exact ProbabilityAboveThreshold
			// DrawNumberNonZeroStrands: Binomial(2*#alleles, 1-thr0, u)
			const uint32_t non_zero_strands = binomial_count(2, sub_rn(1.0, thr0), binom_p0[fragment_length], probability_chosen);
			if(!non_zero_strands){ continue; }
			// ChooseAlleles with possible_strands == 2
			uint32_t chosen[2]; uint32_t n_chosen;
			if(non_zero_strands <= 1){
				const double rv = mt_uniform(g, mt);
				chosen[0] = static_cast<uint32_t>(mul_rn(rv, 2.0)) & 0xffffu;   // SelectAllele with nothing chosen yet
				n_chosen = 1;
			}
			else{
				chosen[0] = 0; chosen[1] = 1; n_chosen = 2;
			}
			for(uint32_t ci = 0; ci < n_chosen; ++ci){
				const bool strand = chosen[ci] & 1u;
				const uint32_t cur_end = pos + fragment_length;
				if(cur_end < L){
					const uint32_t gc_perc = percent_u32(gcp[cur_end] - gcp[pos], fragment_length);
					const double rv = mt_uniform(g, mt);
					const double adjusted_random = add_rn(thr0, mul_rn(rv, sub_rn(1.0, thr0)));
					bool runaway = false;
					const uint32_t counts = fragment_counts(c, b.ref_id, fragment_length, gc_perc, c.sur_start[off + pos], c.sur_end[off + cur_end - 1], adjusted_random, runaway);
					if(runaway && g.lane() == 0){ *c.error_flag |= kErrCountRunaway; }
					if(counts){
						if(kMeth){
							// GetOrgSeq + CTConversion: forward end of the `strand` read first, then the reverse end
							for(uint32_t rev = 0; rev < 2; ++rev){
								const uint32_t seg = rev ? (strand ? 0u : 1u) : (strand ? 1u : 0u);
								uint32_t n = c.read_len_to[seg] + c.max_len_deletion;
								if(fragment_length < n){ n = fragment_length; }
								if(n > c.max_org_len){ n = c.max_org_len; }
								g.sync();
								for(uint32_t i = g.lane(); i < n; i += G::kSize){
									s.frag[rev][i] = rev ? static_cast<uint8_t>(3 - c.ref[off + cur_end - 1 - i]) : c.ref[off + pos + i];
								}
								g.sync();
								MtSource rng{mt};
								ct_conversion(g, c, rng, s.frag[rev], n, b.ref_id, rev ? cur_end : pos, cur_methylation_start, rev != 0);
							}
							g.sync();
						}
						create_reads(g, c, s, mt, sink, counts, strand, b.ref_id, fragment_length, read_number, b.block_id, pos, cur_end, kMeth);
					}
				}
			}
		}
	}
	if(scan_draws && g.lane() == 0){ *scan_draws = draws; }
}

// Simulator::SimulateFromGivenBlock for a reference with variants (-V): every start position is visited once more per inserted base of the
// insertions there, the number of strands with fragments is drawn over the alleles possible at that start, the chosen (allele, strand) ids
// follow ChooseAlleles, and every chosen allele's fragment is evaluated on its own sequence.  chosen: 2 * num_alleles uint16_t of group-shared memory.
template<bool kMeth, class G, class Sink>
RSQ_HD void simulate_block_var(const G &g, const SimCtx &c, const Scratch &s, Sink &sink, const BlockDesc &b, unsigned long long *scan_draws, uint16_t *chosen){
	Mt mt; mt.s = s.mt; mt.idx = kMtN;
	mt_seed(g, mt, b.seed);
	const uint32_t L = c.seq_len[b.ref_id];
	const uint32_t group = c.coverage_group[b.ref_id];
	const double *thr = c.thr + static_cast<size_t>(group) * c.insert_to * 2;
	const uint64_t *thr_int = c.thr_int + static_cast<size_t>(group) * c.insert_to;
	const uint32_t A = c.var.num_alleles, n_pow = 2u * A + 1u;
	const double *binom_pow = c.binom_pow + static_cast<size_t>(group) * c.insert_to * n_pow;
	const VariantView v = c.var.view(b.ref_id);
	uint64_t read_number = 0;
	unsigned long long draws = 0;
	uint32_t end = b.start_pos + 1000u;
	if(end > L){ end = L; }
	int32_t cur_methylation_start = b.first_meth;
	uint32_t first_var = b.first_var, start_variant_pos = 0;
	for(uint32_t pos = b.start_pos; pos < end; ++pos){
		if(kMeth){
			const uint32_t r0 = c.meth_off[b.ref_id], nr = c.meth_off[b.ref_id + 1] - r0;
			if(cur_methylation_start >= 0 && static_cast<uint32_t>(cur_methylation_start) < nr && c.meth_end[r0 + cur_methylation_start] <= pos){ ++cur_methylation_start; }
		}
		do{
			const uint32_t n_possible = count_possible_alleles(v, A, first_var, start_variant_pos, pos);
			uint32_t len = c.insert_from;
			while(len < c.insert_to){
				if(mt.idx >= kMtN){ mt_regen(g, mt); }
				uint32_t n = c.insert_to - len;
				if(n > static_cast<uint32_t>(G::kSize)){ n = G::kSize; }
				if(n > static_cast<uint32_t>(kMtN - mt.idx)){ n = kMtN - mt.idx; }
				const uint32_t lane = g.lane();
				uint64_t x = 0;
				bool hit = false;
				if(lane < n){
					x = mt_temper(mt.s[mt.idx + lane]);
					hit = x >= thr_int[len + lane];
				}
				const unsigned mask = g.ballot(hit);
				if(mask == 0){
					mt.idx += n; len += n; draws += n;
					continue;
				}
#if defined(__CUDA_ARCH__)
				const uint32_t first = __ffs(mask) - 1;
#else
				const uint32_t first = 0;
#endif
				const uint32_t fragment_length = len + first;
				x = mt_temper(mt.s[mt.idx + first]);
				mt.idx += first + 1; len = fragment_length + 1; draws += first + 1;

				const double probability_chosen = canonical(x);
				const double thr0 = thr[2 * fragment_length], thr1 = thr[2 * fragment_length + 1];
				if(!(probability_chosen >= thr1)){ continue; }
				const uint32_t non_zero_strands = binomial_count(2u * n_possible, sub_rn(1.0, thr0), binom_pow[static_cast<size_t>(fragment_length) * n_pow + 2u * n_possible], probability_chosen);
				if(!non_zero_strands){ continue; }
				auto uniform = [&]() -> double { return mt_uniform(g, mt); };
				g.sync();
				const uint32_t n_chosen = choose_alleles(chosen, non_zero_strands, 2u * n_possible, uniform);
				g.sync();
				for(uint32_t ci = 0; ci < n_chosen; ++ci){
					const uint32_t id = chosen[ci];
					const uint32_t allele = nth_possible_allele(v, A, first_var, start_variant_pos, pos, id / 2u);
					const bool strand = id & 1u;
					VarEval e; VarGeom geo;
					bool runaway = false;
					if(!eval_allele_hit(c, v, b.ref_id, pos, first_var, start_variant_pos, fragment_length, allele, thr0, uniform, e, geo, runaway)){ continue; }
					if(runaway && g.lane() == 0){ *c.error_flag |= kErrCountRunaway; }
					if(!e.counts){ continue; }
					const bool staged = e.slow || kMeth;
					if(staged){ splice_fragment_ends(g, c, v, b.ref_id, strand, pos, first_var, start_variant_pos, fragment_length, e, geo, s.frag[0], s.frag[1]); }
					if(kMeth){
						// CTConversion, variant overload: forward end from StartVariant, reverse end from EndVariant
						const int32_t end_var = e.end_var;
						for(uint32_t rev = 0; rev < 2; ++rev){
							const uint32_t seg = rev ? (strand ? 0u : 1u) : (strand ? 1u : 0u);
							uint32_t nn = c.read_len_to[seg] + c.max_len_deletion;
							if(fragment_length < nn){ nn = fragment_length; }
							if(nn > c.max_org_len){ nn = c.max_org_len; }
							MtSource rng{mt};
							ct_conversion_var(g, c, rng, s.frag[rev], nn, b.ref_id, rev ? e.end_position : pos, allele, cur_methylation_start, rev != 0, v,
							                  rev ? end_var : static_cast<int32_t>(first_var), rev ? e.end_var_pos : start_variant_pos);
						}
						g.sync();
					}
					VarRead vr{allele, e.slow, first_var, start_variant_pos, e.end_var, e.end_var_pos, b.start_pos / 1000u};
					create_reads(g, c, s, mt, sink, e.counts, strand, b.ref_id, fragment_length, read_number, b.block_id, pos, e.end_position, staged, &vr);
				}
			}
			next_start_pass(v, pos, first_var, start_variant_pos);
		}while(start_variant_pos);
	}
	if(scan_draws && g.lane() == 0){ *scan_draws = draws; }
}

// ---------------------------------------------------------------------------------------------------
// Systematic errors (master stream).  One chain = Simulator::SetSystematicErrors over [begin,end) of a
// strand, starting from `state`; raw[] holds the two pre-generated master draws of every position.
// ---------------------------------------------------------------------------------------------------
struct SysState {
	uint32_t distance;     // distance_to_start_of_error_region_
	uint32_t start_rate;   // start_error_rate_
};

// strand base at coordinate p of a chain (forward: ref[p]; reverse: complement of ref[L-1-p]; adapter: seq[p])
RSQ_HD uint32_t chain_base(const uint8_t *seq, uint32_t L, bool reverse, uint32_t p){
	return reverse ? 3u - seq[L - 1 - p] : seq[p];
}

// utilities::DominantBase state right before position p is drawn == after Update at p-1
// (utilities.hpp:229-293): counts over the previous min(p,5) bases, ties -> base closest to p.
// hist: the previous min(p,5) bases, 2 bits each, most recent in the low bits; n: how many are valid
RSQ_HD uint32_t dominant_from_window(uint32_t hist, uint32_t n, uint32_t carried){
	if(n == 0){ return carried; }
	uint32_t cnt = 0;   // four 8-bit counters
	for(uint32_t k = 0; k < n; ++k){ cnt += 1u << (8u * ((hist >> (2u * k)) & 3u)); }
	uint32_t mx = cnt & 0xffu;
	for(uint32_t b = 1; b < 4; ++b){ const uint32_t v = (cnt >> (8u * b)) & 0xffu; if(v > mx){ mx = v; } }
	uint32_t base;
	do{ base = hist & 3u; hist >>= 2; }while(((cnt >> (8u * base)) & 0xffu) != mx);
	return base;
}
RSQ_HD void window_before(const uint8_t *seq, uint32_t L, bool reverse, uint32_t p, uint32_t &hist, uint32_t &n){
	const uint32_t lo = p > 5 ? p - 5 : 0;
	hist = 0; n = p - lo;
	for(uint32_t q = lo; q < p; ++q){ hist = ((hist << 2) | chain_base(seq, L, reverse, q)) & 0x3ffu; }
}
RSQ_HD uint32_t dominant_before(const uint8_t *seq, uint32_t L, bool reverse, uint32_t p, uint32_t carried){
	uint32_t hist, n;
	window_before(seq, L, reverse, p, hist, n);
	return dominant_from_window(hist, n, carried);
}

// raw draw k (0: dominant error, 1: error rate) of chain coordinate p.  Reverse-strand and adapter chains own a
// contiguous run of the master stream (2 draws per base); the forward strand is interleaved with one block
// seed per 1000 bases (Simulator::CreateBlock draws the seed, then the block's 2*1000 values).
// With variants every SimBlock is followed by the draws of its variants' replacement bases (SetSystematicErrorVariantsForward / Reverse, 2 per
// base), so the blocks' places in the stream come from a table: blk_off[b] = index of block b's seed (forward) / of its first position draw (reverse).
RSQ_HD uint64_t chain_raw(const uint64_t *raw, bool seed_interleaved, uint32_t p, uint32_t k, const uint64_t *blk_off = nullptr, uint32_t L = 0, bool reverse = false){
	if(blk_off){
		if(reverse){
			const uint32_t f = L - 1u - p, b = f / 1000u;
			const uint64_t e = 1000ull * (b + 1ull) < L ? 1000ull * (b + 1ull) : L;
			return raw[blk_off[b] + 2ull * (e - 1ull - f) + k];
		}
		const uint32_t b = p / 1000u;
		return raw[blk_off[b] + 1ull + 2ull * (p - 1000u * b) + k];
	}
	const size_t i = seed_interleaved ? static_cast<size_t>(p / 1000u) * 2001u + 1u + 2u * (p % 1000u) + k : 2u * static_cast<size_t>(p) + k;
	return raw[i];
}
// CoverageStats::UpdateDistances (CoverageStats.cpp:379-396)
RSQ_HD void update_distances(SysState &st, uint32_t error_rate, uint32_t reset_distance){
	if(st.distance){
		if(st.start_rate < error_rate){ st.distance = 0; st.start_rate = error_rate; }
		else if(++st.distance >= reset_distance){ st.distance = 0; st.start_rate = 0; }
	}
	else if(error_rate){ st.distance = 1; st.start_rate = error_rate; }
}
// (distance_to_start_of_error_region_, start_error_rate_) in front of SimBlock b of a strand - what CreateBlock / CreateUnit copy into
// tmp_distance_to_start_of_error_region / tmp_start_error_rate for the block's variants (Simulator.cpp:990-993, 1236-1239)
RSQ_HD bool chain_block_start(uint32_t L, bool reverse, uint32_t p, uint32_t &block){
	const uint32_t f = reverse ? L - 1u - p : p;
	block = f / 1000u;
	return reverse ? ((f + 1u) % 1000u == 0u || f + 1u == L) : (f % 1000u == 0u);
}

template<class G>
RSQ_HD SysState sys_error_chain(const G &g, const Tables &tab, double *prob, const uint8_t *seq, uint32_t L, bool reverse,
                                uint32_t begin, uint32_t end, SysState st, uint32_t carried_dom, uint32_t sys_gc_range, uint32_t reset_distance,
                                const uint64_t *raw, bool seed_interleaved, uint8_t *out /* indexed by chain coordinate; null: warm-up only */,
                                const uint64_t *blk_off = nullptr, uint32_t *bstate = nullptr){
	// GC window state before `begin` (Simulator::UpdateGC): previous min(begin, range) bases
	uint32_t gc_bases = begin < sys_gc_range ? begin : sys_gc_range;
	uint32_t gc = 0;
	for(uint32_t q = begin - gc_bases; q < begin; ++q){
		const uint32_t b = chain_base(seq, L, reverse, q);
		gc += (b == 1 || b == 2) ? 1u : 0u;
	}
	uint32_t last_base = begin ? chain_base(seq, L, reverse, begin - 1) : 4u;
	uint32_t hist, nwin;
	window_before(seq, L, reverse, begin, hist, nwin);
	bool zero;
	for(uint32_t p = begin; p < end; ++p){
		const uint32_t ref_base = chain_base(seq, L, reverse, p);
		const uint32_t dom_base = dominant_from_window(hist, nwin, carried_dom);
		const uint32_t gc_percent = gc_bases ? percent_u16(gc, gc_bases) : 50u;
		const uint32_t dist = (st.distance + 9) / 10;
		if(bstate && out){
			uint32_t blk;
			if(chain_block_start(L, reverse, p, blk) && g.lane() == 0){ bstate[blk] = st.distance | (st.start_rate << 24); }
		}
		const double u1 = canonical(chain_raw(raw, seed_interleaved, p, 0, blk_off, L, reverse));
		const double u2 = canonical(chain_raw(raw, seed_interleaved, p, 1, blk_off, L, reverse));
		uint32_t dom_error = draw(g, tab, tab.dom_error(ref_base, last_base, dom_base), dist, gc_percent, st.start_rate, 0, u1, prob, zero);
		if(zero){ dom_error = 4; }
		uint32_t error_rate = draw(g, tab, tab.error_rate(ref_base, dom_error), dist, gc_percent, st.start_rate, 0, u2, prob, zero);
		if(zero){ error_rate = 0; }
		error_rate &= 0xffu;
		if(out && g.lane() == 0){
			out[2 * static_cast<size_t>(p)] = static_cast<uint8_t>(dom_error);
			out[2 * static_cast<size_t>(p) + 1] = static_cast<uint8_t>(error_rate);
		}
		last_base = ref_base;
		hist = ((hist << 2) | ref_base) & 0x3ffu;
		if(nwin < 5){ ++nwin; }
		update_distances(st, error_rate, reset_distance);
		// UpdateGC
		if(ref_base == 1 || ref_base == 2){ ++gc; }
		if(gc_bases < sys_gc_range){ ++gc_bases; }
		else{
			const uint32_t ob = chain_base(seq, L, reverse, p - gc_bases);
			if(ob == 1 || ob == 2){ --gc; }
		}
	}
	return st;
}


// Simulator::SetSystematicErrorVariantsForward / Reverse (Simulator.cpp:1011-1147, 769-909) for the variants of ONE SimBlock of one strand:
// walks the block's error rates from the state in front of the block (variants cannot start an error region, and the state stands still while a
// variant's bases are drawn), takes the GC content of the sys_gc_range bases in front of the variant (the variants' own bases do not count),
// and draws (dominant error, rate) per replacement base with the context the host prepared (variant_syserr.hpp: base | last base << 2 |
// dominant base << 5).  raw: the block's variant draws in the master stream, two per base.
struct VarDrawCtx {
	const uint8_t *ctx;          // context bytes of this strand (VariantSysContext::fwd / rev)
	uint8_t *errs;               // VarCtx::errs_fwd / errs_rev (out)
	const uint8_t *sys;          // sys_fwd / sys_rev of the sequence
	const uint32_t *gcp;         // G/C prefix counts of the sequence
	const uint32_t *block_first; // VarCtx::block_first of the sequence
	VariantView v;
	uint32_t L, reverse, sys_gc_range, reset_distance;
};
template<class G>
RSQ_HD void draw_variant_errors_block(const G &g, const Tables &tab, double *prob, const VarDrawCtx &c, uint32_t block, SysState st, const uint64_t *raw){
	SysWalkCtx w{}; w.sys = c.sys; w.errs = nullptr; w.block_first = c.block_first; w.v = c.v; w.L = c.L; w.reverse = c.reverse;
	const uint32_t n_vars = sysw_n_vars(w, block);
	uint32_t cursor = 0;
	uint64_t r = 0;
	bool zero;
	SysWalk at{}; at.block = block; at.cur_var = -1;
	sysw_refresh(w, at);
	for(uint32_t k = 0; k < n_vars; ++k){
		const uint32_t var = sysw_var(w, block, k);
		const uint32_t vp = sysw_var_position(w, block, var);
		for(; cursor < vp; ++cursor){
			at.block_pos = cursor;
			update_distances(st, sysw_entry(w, at)[1], c.reset_distance);
		}
		const uint32_t P = c.v.position[var];
		uint32_t gc_bases, gc;
		if(c.reverse){
			const uint32_t rev_pos = c.L - P - 1u;
			gc_bases = rev_pos < c.sys_gc_range ? rev_pos : c.sys_gc_range;
			gc = c.gcp[P + 1u + gc_bases] - c.gcp[P + 1u];
		}
		else{
			gc_bases = P < c.sys_gc_range ? P : c.sys_gc_range;
			gc = c.gcp[P] - c.gcp[P - gc_bases];
		}
		const uint32_t gc_percent = gc_bases ? percent_u16(gc, gc_bases) : 50u;
		const uint32_t dist = (st.distance + 9) / 10;
		const uint32_t len = c.v.length(var);
		for(uint32_t i = 0; i < len; ++i){
			const uint32_t cb = c.ctx[c.v.bases_off[var] + i];
			const uint32_t base = cb & 3u, last_base = (cb >> 2) & 7u, dom_base = (cb >> 5) & 3u;
			const double u1 = canonical(raw[r]), u2 = canonical(raw[r + 1]);
			r += 2;
			uint32_t dom_error = draw(g, tab, tab.dom_error(base, last_base, dom_base), dist, gc_percent, st.start_rate, 0, u1, prob, zero);
			if(zero){ dom_error = 4; }
			uint32_t error_rate = draw(g, tab, tab.error_rate(base, dom_error), dist, gc_percent, st.start_rate, 0, u2, prob, zero);
			if(zero){ error_rate = 0; }
			if(g.lane() == 0){
				c.errs[2ull * (c.v.bases_off[var] + i)] = static_cast<uint8_t>(dom_error);
				c.errs[2ull * (c.v.bases_off[var] + i) + 1] = static_cast<uint8_t>(error_rate & 0xffu);
			}
		}
	}
}

}  // namespace rsq
