// TEST HARNESS (not part of the product): runs the lane-group templates of reseq_b200/csrc/*.cuh with a
// one-lane group on the CPU, so that their logic can be checked against the reference-built oracle without
// a GPU.  The shipped library never contains or calls this; the GPU tests check the very same templates
// instantiated for 32-lane warps.
//
//   twin <stage.flat from oracle/_ref/dump_tables sim> <seed> <out_prefix> [max_blocks] [methylation.bed or -] [variants.vcf]
// Recomputes: ReplaceN'd reference (taken from the dump), surroundings bias + normalisation + thresholds,
// master stream -> adapter / reverse / forward systematic errors + block seeds, then simulates every block
// and writes <out_prefix>_1.fq / _2.fq.  Prints mismatches against the dump's stage values.
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "../../reseq_b200/csrc/host_profile.hpp"
#include "../../reseq_b200/csrc/sim_core.cuh"
#include "../../reseq_b200/csrc/spec_core.cuh"
#include "../../reseq_b200/csrc/bias_core.cuh"
#include "../../reseq_b200/csrc/variant_syserr.hpp"

using namespace rsq;

struct StringSink {
	std::string out[2];
	uint64_t pairs = 0;
	template<class G> void write_record(const G &, uint32_t seg, const char *id, int id_len, const uint8_t *seq, const uint8_t *qual, uint32_t n){
		std::string &o = out[seg];
		o += '@'; o.append(id, id_len); o += '\n';
		for(uint32_t i = 0; i < n; ++i){ o += "ACGTN"[seq[i] > 4 ? 4 : seq[i]]; }
		o += "\n+\n";
		o.append(reinterpret_cast<const char *>(qual), n);
		o += '\n';
	}
	template<class G> void pair_done(const G &){ ++pairs; }
};

struct DeviceLikeStorage {   // owns everything SimCtx points to
	std::vector<TableDesc> desc; std::vector<double> blob; std::vector<uint32_t> par0;
	std::vector<double> cps;  // all discrete cumulative arrays
	std::vector<Discrete> start_cut[2]; std::vector<uint32_t> start_cut_from[2], adapter_off[2];
	std::vector<uint8_t> adapter_seq, adapter_sys;
	std::vector<double> il_bias, gc_bias;
	std::vector<uint64_t> insert_lengths, seq_off; std::vector<uint32_t> seq_len, gc_prefix, name_off;
	std::vector<uint8_t> ref, sys_fwd, sys_rev;
	std::vector<double> sur_start, sur_end;
	std::string names;
	std::vector<uint16_t> tile_names;
	std::vector<uint32_t> rlbf_row_from[2], rlbf_row_off[2]; std::vector<uint64_t> rlbf_val[2];
	uint32_t error_flag = 0;
	uint32_t max_n0 = 0;
};

static Discrete add_cp(DeviceLikeStorage &st, std::vector<std::vector<double>> &keep, const std::vector<double> &cp){
	keep.push_back(cp);
	Discrete d; d.n = cp.size(); d.cp = keep.back().data();
	(void)st;
	return d;
}

int main(int argc, char **argv){
	if(argc < 4){ fprintf(stderr, "usage: twin <stage.flat> <seed> <out_prefix> [max_blocks]\n"); return 2; }
	FlatFile f; f.load(argv[1]);
	const uint64_t seed = strtoull(argv[2], nullptr, 10);
	const std::string prefix = argv[3];
	const size_t max_blocks = argc > 4 ? strtoull(argv[4], nullptr, 10) : static_cast<size_t>(-1);
	const char *bed = (argc > 5 && std::string(argv[5]) != "-") ? argv[5] : nullptr;
	const char *vcf = argc > 6 ? argv[6] : nullptr;
	Profile p; p.from_flat(f);
	int bad = 0;

	// ---- reference (already N-replaced by the oracle dump) ----
	Genome g;
	for(size_t s = 0; f.has("sim.ref." + std::to_string(s)); ++s){
		const auto &a = f.get("sim.ref." + std::to_string(s));
		g.seqs.emplace_back(a.as<uint8_t>(), a.as<uint8_t>() + a.count);
		const auto &idr = f.get("sim.ref_id." + std::to_string(s));
		g.ids.emplace_back(reinterpret_cast<const char *>(idr.bytes.data()), idr.count);
	}
	if(vcf){ g.read_variants(vcf); }
	if(bed){ g.read_methylation(bed); }
	const uint32_t num_alleles = g.variants.num_alleles;
	FlatVariants flat_vars; VariantSysContext var_ctx; std::vector<uint8_t> errs_fwd, errs_rev; std::vector<uint32_t> block_first, block_first_off;
	DeviceLikeStorage st;
	std::vector<uint32_t> moff{0}, mstart, mend; std::vector<double> mrate;
	std::vector<std::vector<double>> cp_keep; cp_keep.reserve(4096);
	SimCtx c{};
	// tables
	for(const auto &h : p.tables){
		TableDesc d{}; d.n0 = h.par0.size(); d.nm = h.nm; d.stride = d.n0;   // (the engine pads rows to 16 bytes; the one-lane Draw reads either layout)
		for(uint32_t n = 0; n < h.nm; ++n){ d.from[n] = h.from[n]; d.span[n] = h.to[n] - h.from[n]; d.off[n] = st.blob.size(); st.blob.insert(st.blob.end(), h.dim2[n].begin(), h.dim2[n].end()); }
		d.par0_off = st.par0.size(); st.par0.insert(st.par0.end(), h.par0.begin(), h.par0.end());
		st.desc.push_back(d);
		if(d.n0 > st.max_n0){ st.max_n0 = d.n0; }
	}
	const uint32_t T = p.num_tiles;
	c.tab.desc = st.desc.data(); c.tab.blob = st.blob.data(); c.tab.par0 = st.par0.data(); c.tab.num_tiles = T;
	c.tab.quality_base = 0; c.tab.seqq_base = 8 * T; c.tab.basecall_base = 10 * T; c.tab.domerr_base = 50 * T;
	c.tab.errrate_base = 50 * T + 100; c.tab.indel_base = 50 * T + 120;
	c.phred_offset = p.phred_quality_offset; c.max_len_deletion = p.max_len_deletion;
	c.insert_from = std::max<uint64_t>(1, p.insert_lengths.from); c.insert_to = p.insert_lengths.to();
	for(int seg = 0; seg < 2; ++seg){
		c.read_len_from[seg] = p.read_lengths[seg].from; c.read_len_to[seg] = p.read_lengths[seg].to(); c.read_len_count[seg] = p.read_lengths[seg].size();
		const auto &rl = p.read_lengths_by_fragment_length[seg];
		c.rlbf_from[seg] = rl.from; c.rlbf_to[seg] = rl.to();
		st.rlbf_row_off[seg].push_back(0);
		for(const auto &row : rl.v){ st.rlbf_row_from[seg].push_back(row.from); st.rlbf_val[seg].insert(st.rlbf_val[seg].end(), row.v.begin(), row.v.end()); st.rlbf_row_off[seg].push_back(st.rlbf_val[seg].size()); }
		c.rlbf_row_from[seg] = st.rlbf_row_from[seg].data(); c.rlbf_row_off[seg] = st.rlbf_row_off[seg].data(); c.rlbf_val[seg] = st.rlbf_val[seg].data();
	}
	st.insert_lengths.assign(c.insert_to, 0); st.il_bias.assign(c.insert_to, 0.0); st.gc_bias.assign(101, 0.0);
	for(uint32_t i = 0; i < c.insert_to; ++i){ st.insert_lengths[i] = p.insert_lengths[i]; st.il_bias[i] = p.insert_lengths_bias[i]; }
	for(uint32_t i = 0; i < 101; ++i){ st.gc_bias[i] = p.gc_fragment_content_bias[i]; }
	c.insert_lengths = st.insert_lengths.data(); c.il_bias = st.il_bias.data(); c.gc_bias = st.gc_bias.data();
	c.num_tiles = p.tiles.size(); st.tile_names = p.tiles; c.tile_names = st.tile_names.data();
	c.tile_pick = add_cp(st, cp_keep, discrete_cp(p.tile_abundance.begin(), p.tile_abundance.end()));
	c.polya_pick = add_cp(st, cp_keep, discrete_cp(p.polya_tail_length.v.begin(), p.polya_tail_length.v.end()));
	c.polya_from = p.polya_tail_length.from;
	c.overrun_pick = add_cp(st, cp_keep, discrete_cp(p.overrun_bases.begin(), p.overrun_bases.end() - 1));
	for(int seg = 0; seg < 2; ++seg){
		st.adapter_off[seg].push_back(st.adapter_seq.size());
		for(size_t a = 0; a < p.adapter_seqs[seg].size(); ++a){
			for(char ch : p.adapter_seqs[seg][a]){ st.adapter_seq.push_back(Genome::code(ch) & 3); }
			st.adapter_off[seg].push_back(st.adapter_seq.size());
			st.start_cut[seg].push_back(add_cp(st, cp_keep, discrete_cp(p.adapter_start_cut[seg][a].v.begin(), p.adapter_start_cut[seg][a].v.end())));
			st.start_cut_from[seg].push_back(p.adapter_start_cut[seg][a].from);
		}
	}
	st.adapter_sys.assign(2 * st.adapter_seq.size(), 0);
	for(int seg = 0; seg < 2; ++seg){
		c.adapters[seg].n = p.adapter_seqs[seg].size(); c.adapters[seg].off = st.adapter_off[seg].data();
		c.adapters[seg].pick = add_cp(st, cp_keep, discrete_cp(p.adapter_significant_count[seg].begin(), p.adapter_significant_count[seg].end()));
		c.adapters[seg].start_cut = st.start_cut[seg].data(); c.adapters[seg].start_cut_from = st.start_cut_from[seg].data();
	}
	c.adapter_seq = st.adapter_seq.data(); c.adapter_sys = st.adapter_sys.data();
	c.ref_seq_bias = p.ref_seq_bias.data(); c.disp_a = p.dispersion_parameters[0]; c.disp_b = p.dispersion_parameters[1];
	// reference arrays
	c.n_seqs = g.seqs.size();
	uint64_t total = 0;
	st.name_off.push_back(0);
	for(size_t s = 0; s < g.seqs.size(); ++s){
		st.seq_off.push_back(total); st.seq_len.push_back(g.seqs[s].size()); total += g.seqs[s].size();
		st.ref.insert(st.ref.end(), g.seqs[s].begin(), g.seqs[s].end());
		st.names += g.first_part(s); st.name_off.push_back(st.names.size());
	}
	// gc_prefix of sequence s occupies L+1 entries starting at seq_off[s]+s
	st.gc_prefix.resize(total + g.seqs.size() + 1);
	for(size_t s = 0; s < g.seqs.size(); ++s){
		uint32_t *gp = st.gc_prefix.data() + st.seq_off[s] + s;
		uint32_t acc = 0; gp[0] = 0;
		for(size_t i = 0; i < g.seqs[s].size(); ++i){ uint8_t b = g.seqs[s][i]; acc += (b == 1 || b == 2); gp[i + 1] = acc; }
	}
	st.sur_start.assign(total, 0.0); st.sur_end.assign(total, 0.0);
	for(size_t s = 0; s < g.seqs.size(); ++s){
		const uint8_t *seq = g.seqs[s].data(); const uint32_t L = g.seqs[s].size();
		for(uint32_t pos = 0; pos < L; ++pos){
			uint32_t code[3];
			forward_surrounding(seq, L, pos, code);
			st.sur_start[st.seq_off[s] + pos] = surrounding_bias(p.fragment_surroundings_bias[0].data(), p.fragment_surroundings_bias[1].data(), p.fragment_surroundings_bias[2].data(), code);
			reverse_surrounding(seq, L, pos, code);
			st.sur_end[st.seq_off[s] + pos] = surrounding_bias(p.fragment_surroundings_bias[0].data(), p.fragment_surroundings_bias[1].data(), p.fragment_surroundings_bias[2].data(), code);
		}
	}
	c.seq_off = st.seq_off.data(); c.seq_len = st.seq_len.data(); c.ref = st.ref.data(); c.gc_prefix = st.gc_prefix.data();
	c.sur_start = st.sur_start.data(); c.sur_end = st.sur_end.data();
	c.name_blob = st.names.data(); c.name_off = st.name_off.data();
	c.base_id = "ReseqRead"; c.base_id_len = 9;
	c.max_read_len = kMaxReadLen; c.max_org_len = kMaxOrgLen;
	c.error_flag = &st.error_flag;

	if(bed){
		for(size_t i = 0; i < g.seqs.size(); ++i){
			for(size_t r = 0; r < g.unmethylated_regions[i].size(); ++r){ mstart.push_back(g.unmethylated_regions[i][r].first); mend.push_back(g.unmethylated_regions[i][r].second); mrate.push_back(g.unmethylation[i].at(r)); }
			moff.push_back(mstart.size());
		}
		mstart.push_back(0); mend.push_back(0); mrate.push_back(0.0);
		c.meth_alleles = g.methylation_alleles_max > 1 ? g.methylation_alleles_max : 1; c.meth_rate_stride = mrate.size();
		if(c.meth_alleles > 1){
			const size_t stride = mrate.size();
			mrate.resize(stride * c.meth_alleles, 0.0);
			for(uint32_t a = 1; a < c.meth_alleles; ++a){
				for(size_t i = 0; i < g.seqs.size(); ++i){
					for(size_t r = 0; r < g.unmethylated_regions[i].size(); ++r){
						const auto &cols = g.unmethylation_alleles[i];
						mrate[a * stride + moff[i] + r] = cols.size() > 1 ? cols.at(a).at(r) : g.unmethylation[i].at(r);
					}
				}
			}
		}
		c.meth_loaded = 1; c.meth_off = moff.data(); c.meth_start = mstart.data(); c.meth_end = mend.data(); c.meth_rate = mrate.data();
	}

	// ---- normalisation (CalculateBiasNormalization) ----
	const uint64_t total_pairs = f.scalar_i("sim.total_pairs");
	Spline spline;
	if(!spline.get_sample_positions(p.insert_lengths)){ fprintf(stderr, "no sample positions\n"); return 1; }
	std::vector<BiasParam> params;
	for(uint32_t r = p.ref_seq_bias.size(); r--; ){
		if(0.0 != p.ref_seq_bias[r]){
			for(auto fl : spline.sample_positions){ if(fl <= g.seqs[r].size()){ params.push_back({r, fl}); } }
		}
	}
	std::vector<double> sums(params.size()), maxb(params.size(), 0.0);
	for(size_t i = 0; i < params.size(); ++i){
		const uint32_t r = params[i].ref_id, fl = params[i].fragment_length;
		double mx = 0.0;
		sums[i] = sum_bias_chain(c.sur_start + st.seq_off[r], c.sur_end + st.seq_off[r], c.gc_prefix + st.seq_off[r] + r, g.seqs[r].size(), fl,
		                         p.ref_seq_bias[r] * p.insert_lengths_bias[fl], c.gc_bias, mx);
		maxb[i] = mx;
	}
	Normalization norm;
	finish_normalization(norm, p, p.ref_seq_bias, spline, params, sums, maxb, total_pairs, num_alleles, vcf != nullptr);
	{
		const double ref_norm = f.scalar_d("sim.bias_normalization");
		if(ref_norm != norm.bias_normalization){ printf("MISMATCH bias_normalization: oracle %a twin %a\n", ref_norm, norm.bias_normalization); ++bad; }
		for(uint32_t grp = 0; grp < norm.num_groups; ++grp){
			auto rt = f.vec_f64("sim.thresholds." + std::to_string(grp));
			size_t nbad = 0;
			for(size_t i = 0; i < rt.size(); ++i){ if(rt[i] != norm.thresholds[static_cast<size_t>(grp) * c.insert_to * 2 + i]){ if(nbad < 3){ printf("MISMATCH threshold[%u][%zu]: %a vs %a\n", grp, i, rt[i], norm.thresholds[static_cast<size_t>(grp) * c.insert_to * 2 + i]); } ++nbad; } }
			if(nbad){ printf("MISMATCH thresholds group %u: %zu entries\n", grp, nbad); ++bad; }
		}
	}
	c.bias_normalization = norm.bias_normalization; c.coverage_group = norm.coverage_groups.data();
	c.thr = norm.thresholds.data(); c.thr_int = norm.thr_int.data(); c.binom_p0 = norm.binom_p0.data(); c.thr_hi = norm.thr_hi.data(); c.thr_hi_stride = norm.thr_hi_stride;
	if(vcf){
		// Reference::variants_ flattened + SimBlock::first_variant_id_ of every block + the host half of SetSystematicErrorVariants*
		flat_vars = g.variants.flatten();
		flat_vars.position.push_back(0); flat_vars.allele_lo.push_back(0); flat_vars.allele_hi.push_back(0); flat_vars.bases.push_back(0);
		var_ctx = variant_sys_context(g.seqs, g.variants, flat_vars);
		errs_fwd.assign(2 * flat_vars.bases.size() + 2, 0); errs_rev.assign(2 * flat_vars.bases.size() + 2, 0);
		for(size_t s = 0; s < g.seqs.size(); ++s){
			block_first_off.push_back(block_first.size());
			const uint32_t nb = (g.seqs[s].size() + 999) / 1000;
			const auto &vars = g.variants.variants[s];
			uint32_t v = 0;
			for(uint32_t b = 0; b <= nb; ++b){
				while(v < vars.size() && vars[v].position < 1000ull * b){ ++v; }
				block_first.push_back(b == nb ? vars.size() : v);
			}
		}
		c.var.loaded = 1; c.var.num_alleles = num_alleles;
		c.var.seq_first = flat_vars.seq_first.data(); c.var.position = flat_vars.position.data(); c.var.bases_off = flat_vars.bases_off.data(); c.var.bases = flat_vars.bases.data();
		c.var.allele_lo = flat_vars.allele_lo.data(); c.var.allele_hi = flat_vars.allele_hi.data();
		c.var.errs_fwd = errs_fwd.data(); c.var.errs_rev = errs_rev.data(); c.var.block_first = block_first.data(); c.var.block_first_off = block_first_off.data();
		for(int b = 0; b < 3; ++b){ c.var.sur_tab[b] = p.fragment_surroundings_bias[b].data(); }
		c.binom_pow = norm.binom_pow.data();
	}

	// ---- master stream ----
	const uint32_t sys_gc_range = f.scalar_i("sim.sys_gc_range");
	{
		uint64_t reads = 0, sum_len = 0;
		for(int seg = 2; seg--; ){ for(auto len = p.read_lengths[seg].from; len < p.read_lengths[seg].to(); ++len){ reads += p.read_lengths[seg][len]; sum_len += p.read_lengths[seg][len] * len; } }
		const uint32_t mine = ((sum_len + reads / 2) / reads) / 2;
		if(mine != sys_gc_range){ printf("MISMATCH sys_gc_range %u vs %u\n", mine, sys_gc_range); ++bad; }
	}
	std::mt19937_64 master(seed);
	SingleLane lane;
	std::vector<double> prob(st.max_n0 + 4);
	uint32_t carried = 0;
	for(int seg = 2; seg--; ){
		for(size_t a = p.adapter_count_sum[seg].size(); a--; ){
			if(!p.adapter_count_sum[seg][a]){ continue; }
			const uint32_t off = st.adapter_off[seg][a], len = st.adapter_off[seg][a + 1] - off;
			std::vector<uint64_t> raw(2 * len);
			for(auto &r : raw){ r = master(); }
			sys_error_chain(lane, c.tab, prob.data(), st.adapter_seq.data() + off, len, false, 0, len, SysState{0, 0}, carried, sys_gc_range, p.reset_distance, raw.data(), false, st.adapter_sys.data() + 2 * off);
			carried = dominant_before(st.adapter_seq.data() + off, len, false, len, carried);
			const auto &ra = f.get("sim.adapter_sys_error." + std::to_string(seg) + "." + std::to_string(a));
			if(ra.count != 2 * len || memcmp(ra.as<uint8_t>(), st.adapter_sys.data() + 2 * off, 2 * len) != 0){ printf("MISMATCH adapter sys error seg %d adapter %zu\n", seg, a); ++bad; }
		}
	}
	st.sys_fwd.assign(2 * total, 0); st.sys_rev.assign(2 * total, 0);
	std::vector<BlockDesc> blocks;
	uint32_t next_block_id = 1;
	for(size_t s = 0; s < g.seqs.size(); ++s){
		const uint32_t L = g.seqs[s].size();
		if(L < c.insert_to){ continue; }
		const uint32_t nb = (L + 999) / 1000;
		const uint8_t *seq = g.seqs[s].data();
		if(vcf){
			// master stream of a unit with variants: nb seeds of the reverse blocks; per reverse block (last one first) 2 draws per position, then
			// 2 per replacement base of its variants; per forward block its seed, 2 per position, 2 per replacement base
			const uint32_t *bf = block_first.data() + block_first_off[s];
			const uint32_t vf = flat_vars.seq_first[s];
			std::vector<uint64_t> rev_off(nb), fwd_off(nb), vb(nb);
			for(uint32_t b = 0; b < nb; ++b){ vb[b] = flat_vars.bases_off[vf + bf[b + 1]] - flat_vars.bases_off[vf + bf[b]]; }
			uint64_t at = nb;
			for(uint32_t b = nb; b--; ){ rev_off[b] = at; at += 2ull * (std::min<uint64_t>(1000ull * (b + 1), L) - 1000ull * b) + 2 * vb[b]; }
			for(uint32_t b = 0; b < nb; ++b){ fwd_off[b] = at; at += 1 + 2ull * (std::min<uint64_t>(1000ull * (b + 1), L) - 1000ull * b) + 2 * vb[b]; }
			std::vector<uint64_t> raw(at);
			for(auto &r : raw){ r = master(); }
			std::vector<uint32_t> bstate_rev(nb, 0), bstate_fwd(nb, 0);
			sys_error_chain(lane, c.tab, prob.data(), seq, L, true, 0, L, SysState{0, 0}, carried, sys_gc_range, p.reset_distance, raw.data(), false, st.sys_rev.data() + 2 * st.seq_off[s], rev_off.data(), bstate_rev.data());
			carried = dominant_before(seq, L, true, L, carried);
			sys_error_chain(lane, c.tab, prob.data(), seq, L, false, 0, L, SysState{0, 0}, carried, sys_gc_range, p.reset_distance, raw.data(), true, st.sys_fwd.data() + 2 * st.seq_off[s], fwd_off.data(), bstate_fwd.data());
			carried = dominant_before(seq, L, false, L, carried);
			for(int strand = 0; strand < 2; ++strand){
				VarDrawCtx dc{}; dc.ctx = strand ? var_ctx.rev.data() : var_ctx.fwd.data(); dc.errs = strand ? errs_rev.data() : errs_fwd.data();
				dc.sys = (strand ? st.sys_rev.data() : st.sys_fwd.data()) + 2 * st.seq_off[s]; dc.gcp = st.gc_prefix.data() + st.seq_off[s] + s; dc.block_first = bf;
				dc.v = c.var.view(s); dc.L = L; dc.reverse = strand; dc.sys_gc_range = sys_gc_range; dc.reset_distance = p.reset_distance;
				for(uint32_t b = 0; b < nb; ++b){
					const uint32_t bs = (strand ? bstate_rev : bstate_fwd)[b];
					const uint64_t size = std::min<uint64_t>(1000ull * (b + 1), L) - 1000ull * b;
					draw_variant_errors_block(lane, c.tab, prob.data(), dc, b, SysState{bs & 0xffffffu, bs >> 24}, raw.data() + (strand ? rev_off[b] + 2 * size : fwd_off[b] + 1 + 2 * size));
				}
			}
			for(uint32_t b = 0; b < nb; ++b){
				BlockDesc d{}; d.ref_id = s; d.start_pos = b * 1000; d.block_id = next_block_id++; d.first_meth = g.first_methylation_id(s, b * 1000); d.seed = raw[fwd_off[b]]; d.first_var = bf[b];
				blocks.push_back(d);
			}
			continue;
		}
		std::vector<uint64_t> raw(2 * static_cast<size_t>(nb) + 4 * static_cast<size_t>(L));
		for(auto &r : raw){ r = master(); }
		SysState end_rev = sys_error_chain(lane, c.tab, prob.data(), seq, L, true, 0, L, SysState{0, 0}, carried, sys_gc_range, p.reset_distance, raw.data() + nb, false, st.sys_rev.data() + 2 * st.seq_off[s]);
		(void)end_rev;
		carried = dominant_before(seq, L, true, L, carried);
		sys_error_chain(lane, c.tab, prob.data(), seq, L, false, 0, L, SysState{0, 0}, carried, sys_gc_range, p.reset_distance, raw.data() + nb + 2 * static_cast<size_t>(L), true, st.sys_fwd.data() + 2 * st.seq_off[s]);
		carried = dominant_before(seq, L, false, L, carried);
		for(uint32_t b = 0; b < nb; ++b){
			BlockDesc d{}; d.ref_id = s; d.start_pos = b * 1000; d.block_id = next_block_id++; d.first_meth = g.first_methylation_id(s, b * 1000); d.seed = raw[nb + 2 * static_cast<size_t>(L) + static_cast<size_t>(b) * 2001];
			blocks.push_back(d);
		}
	}
	c.sys_fwd = st.sys_fwd.data(); c.sys_rev = st.sys_rev.data();
	{
		const auto &rs = f.get("sim.block_seed"); const auto &rf = f.get("sim.sys_fwd"); const auto &rr = f.get("sim.sys_rev");
		size_t nb = std::min<size_t>(rs.count, blocks.size()), sb = 0;
		for(size_t i = 0; i < nb; ++i){ if(rs.as<uint64_t>()[i] != blocks[i].seed){ ++sb; } }
		if(sb || rs.count != blocks.size()){ printf("MISMATCH block seeds: %zu of %zu differ (oracle %" PRIu64 " blocks, twin %zu)\n", sb, nb, rs.count, blocks.size()); ++bad; }
		// oracle dump: forward errors concatenated block by block (= position order); reverse errors per forward block interval in reverse-strand order
		size_t fb = 0, fi = 0;
		for(const auto &bd : blocks){
			const uint32_t L = st.seq_len[bd.ref_id]; const uint32_t e = std::min(bd.start_pos + 1000, L);
			for(uint32_t q = bd.start_pos; q < e; ++q){
				for(int k = 0; k < 2; ++k, ++fi){
					if(fi < rf.count && rf.as<uint8_t>()[fi] != st.sys_fwd[2 * (st.seq_off[bd.ref_id] + q) + k]){ if(fb < 5){ printf("  sys_fwd diff at ref %u pos %u\n", bd.ref_id, q); } ++fb; }
				}
			}
		}
		if(fb || fi != rf.count){ printf("MISMATCH sys_fwd: %zu bytes differ (%zu vs %" PRIu64 ")\n", fb, fi, rf.count); ++bad; }
		size_t rb = 0, ri = 0;
		for(const auto &bd : blocks){
			const uint32_t L = st.seq_len[bd.ref_id]; const uint32_t e = std::min(bd.start_pos + 1000, L);
			for(uint32_t q = L - e; q < L - bd.start_pos; ++q){
				for(int k = 0; k < 2; ++k, ++ri){
					if(ri < rr.count && rr.as<uint8_t>()[ri] != st.sys_rev[2 * (st.seq_off[bd.ref_id] + q) + k]){ ++rb; }
				}
			}
		}
		if(rb || ri != rr.count){ printf("MISMATCH sys_rev: %zu bytes differ (%zu vs %" PRIu64 ")\n", rb, ri, rr.count); ++bad; }
		if(vcf && f.has("sim.err_variants")){
			// SimBlock::err_variants_ of every block as the reference built them: position_, allele bits and the drawn var_errors_
			const auto &rec = f.get("sim.err_variants"); const auto &er = f.get("sim.err_variant_errors");
			const int64_t *r = rec.as<int64_t>(); const uint8_t *ev = er.as<uint8_t>();
			size_t eoff = 0, vbad = 0, seen = 0;
			std::vector<uint32_t> next_in_block(2 * blocks.size(), 0);
			for(size_t i = 0; i + 5 < rec.count; i += 6, ++seen){
				const size_t bi = r[i]; const int strand = r[i + 1];
				const BlockDesc &bd = blocks.at(bi);
				SysWalkCtx w{}; w.block_first = block_first.data() + block_first_off[bd.ref_id]; w.v = c.var.view(bd.ref_id); w.L = st.seq_len[bd.ref_id]; w.reverse = strand;
				const uint32_t b = bd.start_pos / 1000, k = next_in_block[2 * bi + strand]++;
				bool ok = k < sysw_n_vars(w, b);
				if(ok){
					const uint32_t var = sysw_var(w, b, k);
					ok = sysw_var_position(w, b, var) == static_cast<uint32_t>(r[i + 2]) && w.v.allele_lo[var] == static_cast<uint64_t>(r[i + 3]) && w.v.allele_hi[var] == static_cast<uint64_t>(r[i + 4]) && w.v.length(var) == static_cast<uint32_t>(r[i + 5]);
					if(ok){ ok = 0 == memcmp(ev + eoff, (strand ? errs_rev.data() : errs_fwd.data()) + 2 * static_cast<size_t>(w.v.bases_off[var]), 2 * r[i + 5]); }
				}
				if(!ok){ if(vbad < 5){ printf("  err_variant diff: block %zu strand %d variant %u (position_ %" PRId64 ")\n", bi, strand, k, r[i + 2]); } ++vbad; }
				eoff += 2 * r[i + 5];
			}
			for(size_t bi = 0; bi < blocks.size(); ++bi){
				SysWalkCtx w{}; w.block_first = block_first.data() + block_first_off[blocks[bi].ref_id];
				const uint32_t b = blocks[bi].start_pos / 1000;
				if(next_in_block[2 * bi] != sysw_n_vars(w, b) || next_in_block[2 * bi + 1] != sysw_n_vars(w, b)){ if(vbad < 5){ printf("  err_variant count differs in block %zu\n", bi); } ++vbad; }
			}
			const auto &ff = f.get("sim.block_first_variant_fwd"); const auto &fr = f.get("sim.block_first_variant_rev");
			for(size_t bi = 0; bi < blocks.size() && bi < ff.count; ++bi){
				const uint32_t *bf = block_first.data() + block_first_off[blocks[bi].ref_id]; const uint32_t b = blocks[bi].start_pos / 1000;
				if(ff.as<int64_t>()[bi] != static_cast<int64_t>(bf[b]) || fr.as<int64_t>()[bi] != static_cast<int64_t>(bf[b + 1]) - 1){ if(vbad < 5){ printf("  first_variant_id_ differs in block %zu: %" PRId64 " %" PRId64 " vs %u %d\n", bi, ff.as<int64_t>()[bi], fr.as<int64_t>()[bi], bf[b], static_cast<int>(bf[b + 1]) - 1); } ++vbad; }
			}
			if(vbad){ printf("MISMATCH err_variants: %zu of %zu differ\n", vbad, seen); ++bad; }
			else{ printf("err_variants: %zu variants x 2 strands equal the reference's\n", seen / 2); }
		}
	}

	// ---- simulate ----
	std::vector<unsigned char> scratch_mem(scratch_bytes(st.max_n0, c.max_org_len, c.max_read_len));
	Scratch s = carve_scratch(scratch_mem.data(), st.max_n0, c.max_org_len, c.max_read_len);
	StringSink sink;
	FILE *o1 = fopen((prefix + "_1.fq").c_str(), "wb"), *o2 = fopen((prefix + "_2.fq").c_str(), "wb");
	size_t nsim = std::min(max_blocks, blocks.size());
	unsigned long long total_draws = 0;
	int64_t n_adapter_only = f.scalar_i("sim.num_adapter_only_pairs");
	if(getenv("RSQ_TWIN_ADAPTER_ONLY")){ n_adapter_only = atoll(getenv("RSQ_TWIN_ADAPTER_ONLY")); }   // serial-vs-speculative consistency checks
	// the adapter-only pairs follow the last simulated block: all blocks, or all but the look-ahead blocks the reference creates and never simulates
	const size_t lookahead_blocks = 1 + c.insert_to / 1000;
	const bool run_adapter_only = n_adapter_only && (nsim == blocks.size() || nsim + lookahead_blocks == blocks.size());
	const char *spec_env = getenv("RSQ_TWIN_SPEC");
	if(spec_env){
		// two-phase speculative form (spec_core.cuh) with one-lane groups: rounds of scan_window + ReadMachine
		SpecCtx sp{};
		sp.depth = std::max(1, atoi(spec_env));
		sp.run_depth = sp.depth;
		if(getenv("RSQ_TWIN_MEAN_READS")){ sp.mean_reads = static_cast<float>(atof(getenv("RSQ_TWIN_MEAN_READS"))); sp.run_depth = std::max(1u, sp.depth / 3); }   // dense units speculate deeper than run_depth
		sp.scan_budget = getenv("RSQ_TWIN_BUDGET") ? atoi(getenv("RSQ_TWIN_BUDGET")) : 4000000000u;
		const uint32_t max_rl = std::max(c.read_len_to[0], c.read_len_to[1]);
		sp.words_per_job = (3 * max_rl + 8 + kSpecMargin + 7u) & ~7u; sp.margin = kSpecMargin;
		sp.n_blocks = nsim;
		const bool with_adapter_only = run_adapter_only;
		sp.n_units = nsim + (with_adapter_only ? 1 : 0);
		sp.adapter_only_pairs = with_adapter_only ? n_adapter_only : 0;
		sp.adapter_only_seed = with_adapter_only ? master() : 0;
		std::vector<SpecBlock> sblocks(sp.n_units); std::vector<SpecSnap> snaps(2 * static_cast<size_t>(sp.n_units) * (sp.depth + 1));
		std::vector<ReadJob> jobs(static_cast<size_t>(sp.n_units) * sp.depth);
		std::vector<uint64_t> words(jobs.size() * sp.words_per_job);
		std::vector<uint8_t> conv((bed || vcf) ? static_cast<size_t>(sp.n_units) * kConvSlots * 2 * kMaxOrgLen : 0);
		sp.blocks = sblocks.data(); sp.snaps = snaps.data(); sp.jobs = jobs.data(); sp.words = words.data(); sp.conv = conv.data();
		std::vector<uint16_t> snap_chosen(vcf ? 2 * static_cast<size_t>(sp.n_units) * (sp.depth + 1) * 2 * num_alleles : 0), chosen_live(2 * num_alleles + 2);
		sp.snap_chosen = snap_chosen.data(); sp.chosen_stride = 2 * num_alleles;
		sp.id_cap = kIdCap; sp.seq_off = 16 + sp.id_cap; sp.qual_off = sp.seq_off + ((max_rl + 3) & ~3u); sp.slot_stride = (sp.qual_off + max_rl + 15) & ~15u;
		sp.n_slabs = sp.n_units * 40 + 64;   // 32 reads each: enough for 60x runs of the small fixtures
		std::vector<unsigned char> slots(static_cast<size_t>(sp.n_slabs) * 32 * sp.slot_stride);
		std::vector<uint32_t> slab_next(sp.n_slabs), slab_count(sp.n_slabs);
		uint32_t next_slab = 0, n_done = 0;
		sp.slots = slots.data(); sp.next_slab = &next_slab; sp.slab_next = slab_next.data(); sp.slab_count = slab_count.data(); sp.n_done = &n_done; unsigned long long spec_stat[2] = {0, 0}; sp.stat = spec_stat;
		for(uint32_t u = 0; u < sp.n_units; ++u){ spec_init_unit(c, sp, blocks.data(), 0, u); }
		std::vector<uint64_t> ring_mem(2 * kMtN);
		uint32_t rounds = 0; uint64_t jobs_run = 0, jobs_ok = 0;
		auto draw_fn = [&](bool active, uint32_t table, uint32_t i0, uint32_t i1, uint32_t i2, uint32_t i3, double u, bool &zero) -> uint32_t {
			if(!active){ return 0; }
			return draw(lane, c.tab, table, i0, i1, i2, i3, u, prob.data(), zero);
		};
		auto any_fn = [](bool p){ return p; };
		while(true){
			if(getenv("RSQ_TWIN_VARY_DEPTH")){ sp.run_depth = 1 + (rounds * 7) % sp.depth; }   // the product grows the depth as units finish
			for(uint32_t u = 0; u < sp.n_units; ++u){ if(vcf){ scan_window<true>(lane, c, sp, blocks.data(), 0, u, ring_mem.data(), chosen_live.data()); } else{ scan_window<false>(lane, c, sp, blocks.data(), 0, u, ring_mem.data(), chosen_live.data()); } }
			if(n_done == sp.n_units){ break; }
			++rounds;
			for(uint32_t u = 0; u < sp.n_units; ++u){
				if(sblocks[u].done){ continue; }
				for(uint32_t i = 0; i < sblocks[u].n_jobs; ++i){
					const size_t gidx = static_cast<size_t>(u) * sp.depth + i;
					ReadJob &j = jobs[gidx];
					const uint64_t *slice = spec_slice(sp, gidx);
					if(vcf){ run_read_machine<true>(c, sp, true, j, slice, sp.slots + static_cast<size_t>(j.slot) * sp.slot_stride, draw_fn, any_fn, j.consumed, j.rec_len); }
					else{ run_read_machine<false>(c, sp, true, j, slice, sp.slots + static_cast<size_t>(j.slot) * sp.slot_stride, draw_fn, any_fn, j.consumed, j.rec_len); }
					++jobs_run; jobs_ok += j.consumed == j.assumed;
				}
			}
		}
		for(uint32_t u = 0; u < sp.n_units; ++u){
			uint64_t bytes[2] = {0, 0};
			for(uint32_t slab = sblocks[u].chain_head; slab != kSpecNone; slab = slab_next[slab]){
				for(uint32_t k = 0; k < slab_count[slab]; ++k){
					const unsigned char *slot = sp.slots + (static_cast<size_t>(slab) * 32 + k) * sp.slot_stride;
					const uint32_t *hdr = reinterpret_cast<const uint32_t *>(slot);
					sink.write_record(lane, hdr[2], reinterpret_cast<const char *>(slot + 16), hdr[0], slot + sp.seq_off, slot + sp.qual_off, hdr[1]);
					bytes[hdr[2]] += 1 + hdr[0] + 1 + hdr[1] + 3 + hdr[1] + 1;
					if(hdr[2] == 0){ ++sink.pairs; }
				}
			}
			if(bytes[0] != sblocks[u].bytes[0] || bytes[1] != sblocks[u].bytes[1]){ printf("MISMATCH unit %u byte counts\n", u); ++bad; }
			total_draws += sblocks[u].scan_draws;
			fwrite(sink.out[0].data(), 1, sink.out[0].size(), o1); fwrite(sink.out[1].data(), 1, sink.out[1].size(), o2);
			sink.out[0].clear(); sink.out[1].clear();
		}
		printf("spec: depth=%u rounds=%u reads_run=%llu assumption_held=%llu slabs=%u\n", sp.depth, rounds, (unsigned long long)jobs_run, (unsigned long long)jobs_ok, next_slab);
	}
	else{
		for(size_t i = 0; i < nsim; ++i){
			unsigned long long d = 0;
			std::vector<uint16_t> chosen(2 * num_alleles + 2);
			if(vcf){ if(bed){ simulate_block_var<true>(lane, c, s, sink, blocks[i], &d, chosen.data()); } else{ simulate_block_var<false>(lane, c, s, sink, blocks[i], &d, chosen.data()); } }
			else if(bed){ simulate_block<true>(lane, c, s, sink, blocks[i], &d); } else{ simulate_block<false>(lane, c, s, sink, blocks[i], &d); }
			total_draws += d;
			fwrite(sink.out[0].data(), 1, sink.out[0].size(), o1); fwrite(sink.out[1].data(), 1, sink.out[1].size(), o2);
			sink.out[0].clear(); sink.out[1].clear();
		}
		if(run_adapter_only){
			Mt mt; mt.s = s.mt; mt.idx = kMtN;
			mt_seed(lane, mt, master());
			uint64_t read_number = 0;
			create_reads(lane, c, s, mt, sink, n_adapter_only, false, 0, 0, read_number, 0, 0, 0);
			fwrite(sink.out[0].data(), 1, sink.out[0].size(), o1); fwrite(sink.out[1].data(), 1, sink.out[1].size(), o2);
		}
	}
	fclose(o1); fclose(o2);
	printf("blocks=%zu pairs=%" PRIu64 " scan_draws=%llu error_flag=%u stage_mismatches=%d\n", nsim, sink.pairs, total_draws, st.error_flag, bad);
	return bad ? 1 : 0;
}
