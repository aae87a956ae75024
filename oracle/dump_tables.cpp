// TEST INFRASTRUCTURE (oracle side) -- never linked or executed by the product.
//
// Links the UNMODIFIED reference objects built by oracle/Makefile and writes what the reference holds
// in memory right before / while it simulates, as a flat tagged binary ("RSQFLAT1"), so that
//   * the product's own profile loader + table flattening can be byte-compared with the reference's
//     (DataStats::Load + PrepareProcessing, ProbabilityEstimates::Estimate + PrepareResult), and
//   * the device kernels for the master-stream systematic errors / block seeds / bias normalisation
//     have per-stage ground truth (Simulator::Simulate prologue, Simulator.cpp:2687-2826).
//
//   dump_tables profile <stats.reseq> <out.flat>                      profile + LogArrayResult tables
//   dump_tables patch <in.reseq> <out.reseq> <seed>                    synthetic GC / surroundings / dispersion biases
//   dump_tables errmodel <stats.reseq> <in.fa> <seed> <out.fq>           Simulator::SimulateErrorModelOnly for inputs of several batches (see the mode)
//   dump_tables patch_adapter_only <in.reseq> <out.reseq> <count>      sets InsertLengths()[0] = count: pairs made of adapters only (Simulator::SimulateAdapterOnlyPairs)
//   dump_tables sim <stats.reseq> <ref.fa> <seed> <coverage> <out.flat> [max_blocks] [in.vcf]   + normalisation, thresholds, seeds, sys-errors
//                                                                     (with a VCF: thresholds for its allele count, first_variant_id_ and err_variants_ of every block)
//   dump_tables variants <ref.fa> <in.vcf> <out.txt>                  Reference::variants_ after reading the whole VCF
//   dump_tables alleles <seed> <n> <out.txt>                          n seeded calls of Simulator::ChooseAlleles + the chosen ids
//   dump_tables syserrvar <seed> <n_blocks> <n_walks> <out.txt>       seeded SimBlock chain with SysErrorVariants + walks of GetSysErrorFromBlock
//   dump_tables biasmod <ref.fa> <in.vcf> <seed> <seq> <from> <n> <len_from> <len_to> <out.txt> [sparsity]   trace of VariantBiasVarModifiers over start positions
//   dump_tables negbin <seed> <n> <out.txt>                           seeded fragment counts per allele: GetDispersion + NegativeBinomial as GetFragmentCounts combines them
//   dump_tables varseq <ref.fa> <in.vcf> <seed> <n> <out.txt>         n seeded calls of Reference::ReferenceSequence (variant overload) + results
//
// Private members are reached by re-declaring access for this translation unit only.
#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <mutex>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include <seqan/bam_io.h>
#include <seqan/seq_io.h>
#include <seqan/vcf_io.h>
#include <seqan/modifier.h>
#include <boost/archive/text_iarchive.hpp>
#include <boost/archive/text_oarchive.hpp>
#include <gtest/gtest.h>

#define private public
#define protected public
#include "Simulator.h"
#undef private
#undef protected

namespace reseq{ uint16_t kVerbosityLevel = 1; } // defined in main.cpp:22 in the reference binary
using namespace reseq;

namespace {
struct FlatWriter {
	FILE *f;
	explicit FlatWriter(const char *path) : f(fopen(path, "wb")) {
		if(!f){ perror(path); exit(2); }
		fwrite("RSQFLAT1", 1, 8, f);
	}
	~FlatWriter(){ fclose(f); }
	void raw(const std::string &name, uint32_t dtype, uint64_t count, const void *data, size_t elem){
		uint32_t nl = name.size();
		fwrite(&nl, 4, 1, f); fwrite(name.data(), 1, nl, f);
		fwrite(&dtype, 4, 1, f); fwrite(&count, 8, 1, f);
		if(count){ fwrite(data, elem, count, f); }
	}
	void u8(const std::string &n, const std::vector<uint8_t> &v){ raw(n, 0, v.size(), v.data(), 1); }
	void u32(const std::string &n, const std::vector<uint32_t> &v){ raw(n, 1, v.size(), v.data(), 4); }
	void u64(const std::string &n, const std::vector<uint64_t> &v){ raw(n, 2, v.size(), v.data(), 8); }
	void f64(const std::string &n, const std::vector<double> &v){ raw(n, 3, v.size(), v.data(), 8); }
	void i64(const std::string &n, const std::vector<int64_t> &v){ raw(n, 4, v.size(), v.data(), 8); }
	void str(const std::string &n, const std::string &s){ raw(n, 0, s.size(), s.data(), 1); }
	void s64(const std::string &n, int64_t v){ i64(n, std::vector<int64_t>{v}); }
	void d(const std::string &n, double v){ f64(n, std::vector<double>{v}); }
	// Vect<T> as {from, values...}
	template<class T> void vect(const std::string &n, const Vect<T> &v){
		s64(n + ".from", v.from());
		std::vector<uint64_t> vals;
		for(auto i=v.from(); i<v.to(); ++i){ vals.push_back(v[i]); }
		u64(n, vals);
	}
	void vectd(const std::string &n, const Vect<double> &v){
		s64(n + ".from", v.from());
		std::vector<double> vals;
		for(auto i=v.from(); i<v.to(); ++i){ vals.push_back(v[i]); }
		f64(n, vals);
	}
	// Vect<Vect<u64>> as rows: from, per-row {from, count} + concatenated values
	void vect2(const std::string &n, const Vect<Vect<uint64_t>> &v){
		s64(n + ".from", v.from());
		std::vector<int64_t> rows; std::vector<uint64_t> vals;
		for(auto i=v.from(); i<v.to(); ++i){
			rows.push_back(v[i].from()); rows.push_back(v[i].to() - v[i].from());
			for(auto j=v[i].from(); j<v[i].to(); ++j){ vals.push_back(v[i][j]); }
		}
		i64(n + ".rows", rows);
		u64(n, vals);
	}
};

struct TableDump {
	std::vector<int64_t> desc;   // per table: n0, nm, from[4], to[4], off[4] (doubles), par0_off  -> 15 entries (+1 pad)
	std::vector<double> blob;
	std::vector<uint32_t> par0;
	template<uintMarginId N> void add(const ProbabilityEstimatesSubClasses::LogArrayResult<N> &r){
		int64_t e[16] = {0};
		e[0] = r.par0_indeces_.size();
		e[1] = N-1;
		for(uintMarginId n=0; n<N-1; ++n){
			e[2+n] = r.limits_.at(n).first;
			e[6+n] = r.limits_.at(n).second;
			e[10+n] = blob.size();
			blob.insert(blob.end(), r.dim2_.at(n).begin(), r.dim2_.at(n).end());
		}
		e[14] = par0.size();
		for(auto p : r.par0_indeces_){ par0.push_back(p); }
		desc.insert(desc.end(), e, e+16);
	}
};

void DumpProfile(FlatWriter &w, const DataStats &stats, const ProbabilityEstimates &est){
	// ---- DataStats members the simulation reads (Simulator.cpp, FragmentDistributionStats.cpp:3504-3627) ----
	for(int seg=0; seg<2; ++seg){
		std::string s = std::to_string(seg);
		w.vect("read_lengths." + s, stats.ReadLengths(seg));
		w.vect2("read_lengths_by_fragment_length." + s, stats.ReadLengthsByFragmentLength(seg));
		w.vect2("non_mapped_read_lengths_by_fragment_length." + s, stats.NonMappedReadLengthsByFragmentLength(seg));
		std::vector<uint64_t> counts(stats.Adapters().Counts(seg).begin(), stats.Adapters().Counts(seg).end());
		w.u64("adapter.count_sum." + s, counts);
		std::vector<uint64_t> sig(stats.Adapters().SignificantCounts(seg).begin(), stats.Adapters().SignificantCounts(seg).end());
		w.u64("adapter.significant_count." + s, sig);
		w.s64("adapter.n." + s, stats.adapters_.seqs_.at(seg).size());
		for(size_t a=0; a<stats.adapters_.seqs_.at(seg).size(); ++a){
			std::string seq;
			for(auto c : stats.adapters_.seqs_.at(seg).at(a)){ seq += static_cast<char>(c); }
			w.str("adapter.seq." + s + "." + std::to_string(a), seq);
			w.vect("adapter.start_cut." + s + "." + std::to_string(a), stats.Adapters().StartCut(seg, a));
		}
	}
	w.vect("adapter.polya_tail_length", stats.Adapters().PolyATailLength());
	w.u64("adapter.overrun_bases", std::vector<uint64_t>(stats.Adapters().OverrunBases().begin(), stats.Adapters().OverrunBases().end()));
	w.s64("phred_quality_offset", stats.PhredQualityOffset());
	w.s64("total_number_reads", stats.TotalNumberReads());
	w.d("corrected_coverage", stats.CorrectedCoverage());
	w.s64("creation_time", stats.CreationTime());
	w.s64("reset_distance", stats.Coverage().reset_distance_);
	w.s64("max_len_deletion", stats.Errors().MaxLenDeletion());
	w.u64("tiles.tiles", std::vector<uint64_t>(stats.Tiles().Tiles().begin(), stats.Tiles().Tiles().end()));
	w.u64("tiles.abundance", std::vector<uint64_t>(stats.Tiles().Abundance().begin(), stats.Tiles().Abundance().end()));

	const auto &fd = stats.FragmentDistribution();
	w.vect("insert_lengths", fd.insert_lengths_);
	w.f64("ref_seq_bias", fd.ref_seq_bias_);
	w.vectd("insert_lengths_bias", fd.insert_lengths_bias_);
	w.vectd("gc_fragment_content_bias", fd.gc_fragment_content_bias_);
	for(int b=0; b<3; ++b){
		w.f64("fragment_surroundings_bias." + std::to_string(b), fd.fragment_surroundings_bias_.bias_.at(b));
	}
	w.f64("dispersion_parameters", std::vector<double>(fd.dispersion_parameters_.begin(), fd.dispersion_parameters_.end()));

	// ---- LogArrayResult tables in the fixed family order used by the engine ----
	TableDump t;
	const size_t tiles = est.quality_result_.at(0).size();
	w.s64("tab.num_tiles", tiles);
	for(int seg=0; seg<2; ++seg) for(size_t tile=0; tile<tiles; ++tile) for(int b=0; b<4; ++b) t.add(est.quality_result_.at(seg).at(tile).at(b));
	for(int seg=0; seg<2; ++seg) for(size_t tile=0; tile<tiles; ++tile) t.add(est.sequence_quality_result_.at(seg).at(tile));
	for(int seg=0; seg<2; ++seg) for(size_t tile=0; tile<tiles; ++tile) for(int b=0; b<4; ++b) for(int d=0; d<5; ++d) t.add(est.base_call_result_.at(seg).at(tile).at(b).at(d));
	for(int b=0; b<4; ++b) for(int l=0; l<5; ++l) for(int d=0; d<5; ++d) t.add(est.dom_error_result_.at(b).at(l).at(d));
	for(int b=0; b<4; ++b) for(int d=0; d<5; ++d) t.add(est.error_rate_result_.at(b).at(d));
	for(int ty=0; ty<2; ++ty) for(int c=0; c<6; ++c) t.add(est.indels_result_.at(ty).at(c));
	w.i64("tab.desc", t.desc);
	w.f64("tab.blob", t.blob);
	w.u32("tab.par0", t.par0);
}

bool LoadAll(DataStats &stats, ProbabilityEstimates &est, const std::string &stats_file){
	if(!stats.Load(stats_file.c_str())){ return false; }
	stats.PrepareProcessing();
	std::string ipf = stats_file + ".ipf";
	// same call main.cpp:841 makes for `--ipfIterations 0` with an existing .ipf
	if(!est.Estimate(stats, 0, 5, 1, ipf.c_str(), ipf.c_str())){ return false; }
	est.PrepareResult();
	return true;
}
}

int main(int argc, char **argv){
	if(argc < 4){
		std::cerr << "usage: dump_tables profile <stats.reseq> <out> | sim <stats.reseq> <ref.fa> <seed> <coverage> <out> [max_blocks]" << std::endl;
		return 2;
	}
	std::string mode = argv[1];
	kVerbosityLevel = 1;
	if(mode == "profile"){
		DataStats stats(NULL);
		ProbabilityEstimates est;
		if(!LoadAll(stats, est, argv[2])){ return 1; }
		FlatWriter w(argv[3]);
		DumpProfile(w, stats, est);
		return 0;
	}
	if(mode == "patch" && argc >= 5){
		// Replace the (failed / uniform) bias fit of a tiny synthetic data set by a deterministic non-trivial one and
		// write the profile back with the reference's own DataStats::Save, so that the fragment-count model
		// (GC spline values, surroundings, dispersion) is exercised by the parity tests.
		DataStats stats(NULL);
		if(!stats.Load(argv[2])){ return 1; }
		std::mt19937_64 gen(std::stoull(argv[4]));
		auto &fd = stats.fragment_distribution_;
		for(uint32_t gc = 0; gc <= 100; ++gc){
			double x = (static_cast<double>(gc) - 48.0) / 22.0;
			fd.gc_fragment_content_bias_[gc] = std::round((0.15 + 1.6 * std::exp(-x * x)) * 256.0) / 256.0;
		}
		std::array<double, 4*Surrounding::Length()> separated;
		for(auto &v : separated){ v = (static_cast<double>(gen() % 17) - 8.0) / 16.0; }
		fd.fragment_surroundings_bias_.CombinePositions(separated);
		fd.dispersion_parameters_ = {{0.25, 0.75}};
		if(!stats.Save(argv[3])){ return 1; }
		return 0;
	}
	if(mode == "patch_adapter_only" && argc >= 5){
		// None of the synthetic data sets has read pairs without a fragment, so their profiles never send the reference into
		// SimulateAdapterOnlyPairs (Simulator.cpp:2359-2382).  This gives a profile such pairs and writes it back with DataStats::Save.
		DataStats stats(NULL);
		if(!stats.Load(argv[2])){ return 1; }
		stats.fragment_distribution_.insert_lengths_[0] = std::stoull(argv[4]);
		if(!stats.Save(argv[3])){ return 1; }
		return 0;
	}
	if(mode == "variants" && argc >= 5){
		// Reference::PrepareVariantFile + ReadFirstVariants + ReadVariants (Reference.cpp:126-426, 1005-1078), as Simulate calls
		// them (Simulator.cpp:2750-2751, 938, 1278), but for all sequences at once. One text line per variant:
		// "<seq> <position> <var_seq or -> <allele bits 0-63, hex> <allele bits 64-127, hex>"; exit code 1 + no lines when the reference rejects the file.
		Reference ref;
		if(!ref.ReadFasta(argv[2])){ return 1; }
		std::ofstream out(argv[4]);
		if(!ref.PrepareVariantFile(argv[3]) || !ref.ReadFirstVariants() || !ref.ReadVariants(ref.NumberSequences())){
			out << "rejected\n";
			return 1;
		}
		out << "alleles " << ref.NumAlleles() << "\n";
		for(uintRefSeqId s = 0; s < ref.NumberSequences(); ++s){
			for(const auto &v : ref.Variants(s)){
				std::string vs;
				for(auto b : v.var_seq_){ vs += static_cast<char>(b); }
				char bits[64];
				snprintf(bits, sizeof(bits), "%llx %llx", static_cast<unsigned long long>(v.allele_[0]), static_cast<unsigned long long>(v.allele_[1]));
				out << s << ' ' << v.position_ << ' ' << (vs.empty() ? std::string("-") : vs) << ' ' << bits << "\n";
			}
		}
		return 0;
	}
	if(mode == "negbin"){
		// The tail of FragmentDistributionStats::GetFragmentCounts (FragmentDistributionStats.cpp:3617-3626) with an allele count:
		// dispersion = GetDispersion(mean, a, b) / alleles; mean /= alleles; NegativeBinomial(mean/(mean+dispersion), dispersion, u).
		// Line: "<mean bits> <a bits> <b bits> <alleles> <u bits> <count>" (doubles as hex bit patterns).
		std::mt19937_64 gen(std::stoull(argv[2]));
		std::ofstream out(argv[4]);
		std::uniform_real_distribution<double> zero_to_one(0.0, 1.0);
		auto bits = [](double v){ uint64_t b; std::memcpy(&b, &v, 8); char t[20]; snprintf(t, sizeof t, "%016llx", static_cast<unsigned long long>(b)); return std::string(t); };
		for(size_t call = 0; call < std::stoull(argv[3]); ++call){
			const double mean0 = std::exp(-9.0 + 13.0 * zero_to_one(gen));   // 1e-4 .. 55
			const double a = call % 5 == 0 ? 0.0 : zero_to_one(gen), b = call % 5 == 0 ? 1e-100 : 0.1 + 2.0 * zero_to_one(gen);
			const uintAlleleId alleles = call % 3 == 0 ? 1 : 1 + gen() % 128;
			const double thr = zero_to_one(gen);
			const double u = thr + zero_to_one(gen) * (1 - thr);   // adjusted_random of SimulateFromGivenBlock
			double mean = mean0;
			double dispersion = BiasCalculationVectors::GetDispersion(mean, a, b) / alleles;
			mean /= alleles;
			const uintDupCount count = FragmentDistributionStats::NegativeBinomial(mean / (mean + dispersion), dispersion, u);
			out << bits(mean0) << ' ' << bits(a) << ' ' << bits(b) << ' ' << alleles << ' ' << bits(u) << ' ' << count << "\n";
		}
		return 0;
	}
	if(mode == "alleles"){
		// Simulator::ChooseAlleles (Simulator.cpp:1386-1397) on one stream seeded with <seed>: per call possible_strands = 2 x alleles
		// (1..128 alleles) and 1..possible_strands non-zero strands. Line: "<possible_strands> <non_zero_strands> <ids...>".
		Simulator sim;
		std::mt19937_64 gen(std::stoull(argv[2]));
		std::ofstream out(argv[4]);
		std::vector<uintAlleleId> chosen;
		std::vector<bool> reverse_selection;
		std::uniform_real_distribution<double> zero_to_one(0.0, 1.0);
		for(size_t call = 0; call < std::stoull(argv[3]); ++call){
			const uintAlleleId alleles = call % 7 == 0 ? 1 + gen() % 128 : 1 + gen() % 6;
			const uintAlleleId possible = 2 * alleles;
			const uintAlleleId non_zero = 1 + gen() % possible;
			// ChooseAlleles needs a GeneralRandomDistributions only for ZeroToOne (Simulator.h:209-211: uniform_real_distribution<double>(0,1));
			// its two halves are public through `#define private public`, so they are driven with the same distribution directly.
			chosen.clear();
			reverse_selection.clear();
			reverse_selection.resize(possible, true);
			const uintAlleleId n_draw = non_zero <= possible / 2 ? non_zero : possible - non_zero;
			while(chosen.size() < n_draw){ sim.SelectAllele(chosen, reverse_selection, possible, zero_to_one(gen)); }
			if(non_zero > possible / 2){ sim.ReverseSelection(chosen, reverse_selection, possible); }
			out << possible << ' ' << non_zero;
			for(auto id : chosen){ out << ' ' << id; }
			out << "\n";
		}
		return 0;
	}
	if(mode == "syserrvar" && argc >= 6){
		// Simulator::GetSysErrorFromBlock (Simulator.cpp:240-292) over a chain of SimBlocks with seeded sys_errors_ / err_variants_ (3 alleles).
		// "b <len> <n variants>", "s <4 hex digits per position: dominant error, rate>", "v <position> <allele bits> <n errors> <hex>" describe the
		// chain; "walk <block> <block_pos> <cur_var> <allele> <steps> <4 hex digits per step>" is what the reference returns step by step.
		Simulator sim;
		std::mt19937_64 gen(std::stoull(argv[2]));
		const size_t n_blocks = std::stoull(argv[3]);
		std::ofstream out(argv[5]);
		auto hex4 = [](unsigned dom, unsigned rate){ char b[8]; snprintf(b, sizeof b, "%02x%02x", dom, rate); return std::string(b); };
		std::vector<Simulator::SimBlock *> blocks;
		for(size_t b = 0; b < n_blocks; ++b){
			auto *blk = new Simulator::SimBlock(b, b * 1000, NULL, 0);
			const size_t len = b + 1 == n_blocks ? 1 + gen() % 1000 : (b % 5 == 3 ? 1 + gen() % 40 : 1000);
			for(size_t p = 0; p < len; ++p){ blk->sys_errors_.emplace_back(seqan::Dna5(gen() % 5), gen() % 101); }
			uintSeqLen pos = gen() % 30;
			while(pos < len){
				std::vector<std::pair<seqan::Dna5, uintPercent>> errs;
				const unsigned kind = gen() % 4;
				const size_t n_err = kind == 0 ? 0 : (kind == 3 ? 2 + gen() % 5 : 1);
				for(size_t k = 0; k < n_err; ++k){ errs.emplace_back(seqan::Dna5(gen() % 5), gen() % 101); }
				blk->err_variants_.emplace_back(pos, errs, std::array<uintAlleleBitArray, 2>{{1 + gen() % 7, 0}});
				pos += gen() % 4 == 0 ? 0 : (gen() % 3 == 0 ? 1 : 1 + gen() % 60);   // same position, neighbours, gaps
			}
			if(!blocks.empty()){ blocks.back()->next_block_ = blk; }
			blocks.push_back(blk);
			out << "b " << len << ' ' << blk->err_variants_.size() << "\ns ";
			for(auto &e : blk->sys_errors_){ out << hex4(static_cast<unsigned>(seqan::ordValue(e.first)), e.second); }
			out << "\n";
			for(auto &v : blk->err_variants_){
				out << "v " << v.position_ << ' ' << v.allele_[0] << ' ' << v.var_errors_.size() << ' ';
				for(auto &e : v.var_errors_){ out << hex4(static_cast<unsigned>(seqan::ordValue(e.first)), e.second); }
				out << "\n";
			}
		}
		for(size_t w = 0; w < std::stoull(argv[4]); ++w){
			const size_t b0 = gen() % (n_blocks - 1);
			const Simulator::SimBlock *block = blocks[b0];
			uintSeqLen block_pos = gen() % block->sys_errors_.size();
			intVariantId cur_var = 0;
			while(cur_var < static_cast<intVariantId>(block->err_variants_.size()) && block->err_variants_.at(cur_var).position_ < block_pos){ ++cur_var; }
			uintSeqLen var_pos = 0;
			const uintAlleleId allele = gen() % 3;
			const size_t steps = 50 + gen() % 400;
			out << "walk " << b0 << ' ' << block_pos << ' ' << cur_var << ' ' << allele << ' ' << steps << ' ';
			for(size_t k = 0; k < steps && block; ++k){
				seqan::Dna5 dom;
				uintPercent rate;
				sim.GetSysErrorFromBlock(dom, rate, block, block_pos, cur_var, var_pos, allele);
				out << hex4(static_cast<unsigned>(seqan::ordValue(dom)), rate);
			}
			out << "\n";
		}
		return 0;
	}
	if(mode == "biasmod" && argc >= 11){
		// The variant bookkeeping of Simulator::SimulateFromGivenBlock (Simulator.cpp:2287-2353) without the random stream of a block: for
		// every start position of [from, from+n) the do-while over inserted start bases, PrepareBiasModForCurrentStartPos, GetPossibleAlleles,
		// then - for a seeded sparse ascending choice of fragment lengths and alleles, the way hits arrive - PrepareBiasModForCurrentFragmentLength,
		// GetGCPercent, StartVariant/EndVariant, and CheckForInsertedBasesToStartFrom at the end of a pass.
		//   "s <sequence after ReplaceN>"
		//   "p <start> <first_variant_id> <start_variant_pos> <possible alleles...>"
		//   "f <length> <allele> <end_pos_shift> <gc_perc or -> <start surrounding x3> <end surrounding x3> <StartVariant id pos> <EndVariant id pos>"
		Reference ref;
		if(!ref.ReadFasta(argv[2])){ return 1; }
		const uintSeed seed = std::stoull(argv[4]);
		ref.ReplaceN(seed);   // Simulator.cpp:2690 runs before the VCF is opened (2750)
		if(!ref.PrepareVariantFile(argv[3]) || !ref.ReadFirstVariants() || !ref.ReadVariants(ref.NumberSequences())){ return 1; }
		std::mt19937_64 gen(seed);
		const uintRefSeqId seq = std::stoul(argv[5]);
		const uintSeqLen from = std::stoul(argv[6]), n_pos = std::stoul(argv[7]), len_from = std::stoul(argv[8]), len_to = std::stoul(argv[9]);
		std::ofstream out(argv[10]);
		const uint64_t sparsity = argc > 11 ? std::stoull(argv[11]) : 24;   // one fragment length in <sparsity> is evaluated (1 = every length)
		Simulator sim;
		const auto &vars = ref.Variants(seq);
		intVariantId first_var = 0;
		while(first_var < static_cast<intVariantId>(vars.size()) && vars.at(first_var).position_ < from){ ++first_var; }
		Simulator::VariantBiasVarModifiers bias_mod(first_var, ref.NumAlleles());
		Surrounding surrounding_start;
		ref.ForwardSurrounding(surrounding_start, seq, (0 < from ? from - 1 : ref.SequenceLength(seq) - 1));
		std::vector<uintAlleleId> possible_alleles;
		out << "s ";   // the sequence the reference works on (after ReplaceN)
		for(auto b : ref.ReferenceSequence(seq)){ out << static_cast<char>(b); }
		out << "\n";
		for(uintSeqLen start = from; start < from + n_pos && start < ref.SequenceLength(seq); ++start){
			surrounding_start.UpdateForward(ref.ReferenceSequence(seq), start);
			do{
				sim.PrepareBiasModForCurrentStartPos(bias_mod, seq, ref, start, len_from, surrounding_start);
				sim.GetPossibleAlleles(possible_alleles, ref, bias_mod, start, seq);
				out << "p " << start << ' ' << bias_mod.first_variant_id_ << ' ' << bias_mod.start_variant_pos_;
				for(auto a : possible_alleles){ out << ' ' << a; }
				out << "\n";
				for(uintSeqLen len = len_from; len < len_to; ++len){
					if(gen() % sparsity){ continue; }
					for(auto allele : possible_alleles){
						if(gen() % 2){ continue; }
						sim.PrepareBiasModForCurrentFragmentLength(bias_mod, seq, ref, start, len, allele);
						const uintSeqLen end = start + len + bias_mod.end_pos_shift_.at(allele);
						out << "f " << len << ' ' << allele << ' ' << bias_mod.end_pos_shift_.at(allele) << ' ';
						if(end < ref.SequenceLength(seq)){ out << static_cast<unsigned>(sim.GetGCPercent(bias_mod, seq, ref, end, len, allele)); }
						else{ out << '-'; }
						for(auto v : bias_mod.surrounding_start_.at(allele).sur_){ out << ' ' << v; }
						for(auto v : bias_mod.surrounding_end_.at(allele).sur_){ out << ' ' << v; }
						const auto sv = bias_mod.StartVariant();
						out << ' ' << sv.first << ' ' << sv.second;
						if(end < ref.SequenceLength(seq)){
							const auto ev = bias_mod.EndVariant(vars, end, allele);
							out << ' ' << ev.first << ' ' << ev.second;
						}
						out << "\n";
					}
				}
				sim.CheckForInsertedBasesToStartFrom(bias_mod, seq, start, ref);
			} while(bias_mod.start_variant_pos_);
		}
		return 0;
	}
	if(mode == "varseq" && argc >= 7){
		// Reference::ReferenceSequence(insert_string, seq, start, frag_length, reversed, variants, {first variant, posCurrentlyAt}, allele)
		// (Reference.cpp:498-567) on seeded arguments of the kind Simulator::GetOrgSeq passes (Simulator.cpp:1909-1914): the first variant
		// at/after the start (forward) or the last one before it (reverse), or a start inside the replacement of an insertion.
		// One line per call: "call <seq> <start> <len> <reversed> <first variant> <posCurrentlyAt> <allele> <result>".
		Reference ref;
		if(!ref.ReadFasta(argv[2])){ return 1; }
		if(!ref.PrepareVariantFile(argv[3]) || !ref.ReadFirstVariants() || !ref.ReadVariants(ref.NumberSequences())){ return 1; }
		std::mt19937_64 gen(std::stoull(argv[4]));
		const size_t n_calls = std::stoull(argv[5]);
		std::ofstream out(argv[6]);
		for(size_t call = 0; call < n_calls; ++call){
			const uintRefSeqId s = gen() % ref.NumberSequences();
			const auto &vars = ref.Variants(s);
			if(vars.empty()){ --call; continue; }
			const uintSeqLen L = ref.SequenceLength(s);
			const bool reversed = gen() & 1;
			uintSeqLen len = (gen() % 8 == 0) ? 1 + gen() % 6 : 20 + gen() % 300;
			uintAlleleId allele = gen() % ref.NumAlleles();
			uintSeqLen start;
			intVariantId first;
			uintSeqLen first_pos = 0;
			const intVariantId pick = gen() % vars.size();
			if(gen() % 3 == 0 && length(vars.at(pick).var_seq_) > 1){
				// start inside an insertion's replacement (the allele must carry it)
				allele = vars.at(pick).FirstAllele();
				first = pick;
				first_pos = 1 + gen() % (length(vars.at(pick).var_seq_) - 1);
				start = reversed ? vars.at(pick).position_ + 1 : vars.at(pick).position_;
			}
			else{
				// start near a variant so that variants fall inside the returned sequence
				const uintSeqLen jitter = gen() % 64;
				if(reversed){
					start = std::min<uintSeqLen>(L, vars.at(pick).position_ + 1 + jitter);
					first = -1;
					for(intVariantId v = 0; v < static_cast<intVariantId>(vars.size()) && vars.at(v).position_ < start; ++v){ first = v; }
				}
				else{
					start = vars.at(pick).position_ > jitter ? vars.at(pick).position_ - jitter : 0;
					first = vars.size();
					for(intVariantId v = vars.size(); v-- && vars.at(v).position_ >= start; ){ first = v; }
				}
			}
			if(reversed ? start < len + 400 : start + len + 400 > L){ --call; continue; }   // keep the walk inside the sequence
			bool has_n = false;
			for(uintSeqLen p = reversed ? start - len - 400 : start; p < (reversed ? start : start + len + 400); ++p){ has_n |= ref.ReferenceSequence(s)[p] == 'N'; }
			if(has_n){ --call; continue; }   // the simulator only sees the reference after ReplaceN
			seqan::DnaString res;
			ref.ReferenceSequence(res, s, start, len, reversed, vars, {first, first_pos}, allele);
			out << "call " << s << ' ' << start << ' ' << len << ' ' << reversed << ' ' << first << ' ' << first_pos << ' ' << allele << ' ' << res << "\n";
		}
		return 0;
	}
	if(mode == "errmodel" && argc >= 6){
		// Simulator::SimulateErrorModelOnly (Simulator.cpp:2900-3014) on inputs of more than one batch of kBatchSizeErrorModelOnly records.
		// The reference as released waits for written_blocks_ >= cur_block in WriteSingleReads (Simulator.cpp:184-192) and never increments
		// written_blocks_, so `reseq seqToIllumina` dead-locks on its second batch.  Everything else of the call - one block_seed_gen_() seed
		// per batch in input order, ApplyErrorsAndQualityToFastaInput, FlushWriteValues - is the reference's own code; with the counter preset
		// the wait is never entered and one thread writes the batches in order.
		DataStats stats(NULL);
		ProbabilityEstimates est;
		if(!LoadAll(stats, est, argv[2])){ return 1; }
		Simulator sim;
		sim.written_blocks_ = std::numeric_limits<uintFragCount>::max();
		return sim.SimulateErrorModelOnly(argv[5], argv[3], stats, est, 1, std::stoull(argv[4])) ? 0 : 1;
	}
	if(mode == "sim" && argc >= 7){
		Reference ref;
		if(!ref.ReadFasta(argv[3])){ return 1; }
		DataStats stats(NULL);
		ProbabilityEstimates est;
		if(!LoadAll(stats, est, argv[2])){ return 1; }
		uintSeed seed = std::stoull(argv[4]);
		double coverage = std::stod(argv[5]);
		size_t max_blocks = argc > 7 ? std::stoull(argv[7]) : static_cast<size_t>(-1);
		const char *vcf = argc > 8 ? argv[8] : nullptr;
		FlatWriter w(argv[6]);

		// --- Simulator::Simulate prologue, statement by statement (Simulator.cpp:2687-2797) ---
		Simulator sim;
		sim.block_seed_gen_.seed(seed);
		ref.ReplaceN(seed);
		if(!stats.FragmentDistribution().UpdateRefSeqBias(RefSeqBiasSimulation::kNo, "", ref, sim.block_seed_gen_)){ return 1; }
		DumpProfile(w, stats, est); // after UpdateRefSeqBias so ref_seq_bias matches the simulated reference
		sim.record_base_identifier_ = "ReseqRead";
		uintFragCount reads(0);
		uintRefLenCalc sum_read_length(0);
		for(auto seg=2; seg--;){
			for(auto len = stats.ReadLengths(seg).from(); len < stats.ReadLengths(seg).to(); ++len){
				reads += stats.ReadLengths(seg).at(len);
				sum_read_length += stats.ReadLengths(seg).at(len) * len;
			}
		}
		double average_read_length = static_cast<double>(sum_read_length)/reads;
		auto total_ref_size = ref.TotalSize();
		double adapter_part = Simulator::CoveragePropLostFromAdapters(stats);
		sim.total_pairs_ = Simulator::CoverageToNumberPairs(coverage, total_ref_size, average_read_length, adapter_part);
		sim.num_adapter_only_pairs_ = round( static_cast<double>(sim.total_pairs_)*stats.FragmentDistribution().InsertLengths()[0]/(stats.TotalNumberReads()/2) );
		sim.total_pairs_ -= sim.num_adapter_only_pairs_;
		sim.simulation_error_ = false;
		if(vcf){   // Simulator.cpp:2747-2759
			if(!ref.PrepareVariantFile(vcf) || !ref.ReadFirstVariants()){ return 1; }
			sim.sys_dom_base_per_allele_.resize(ref.NumAlleles());
		}
		w.s64("sim.num_alleles", ref.NumAlleles());
		sim.bias_normalization_ = stats.FragmentDistribution().CalculateBiasNormalization(sim.coverage_groups_, sim.non_zero_thresholds_, ref, 1, sim.total_pairs_);
		sim.sys_gc_range_ = utilities::Divide(sum_read_length,reads)/2;
		for(auto seg=2; seg--;){
			sim.adapter_sys_error_.at(seg).resize( stats.Adapters().Counts(seg).size() );
			for(auto seq=stats.Adapters().Counts(seg).size(); seq--;){
				if( stats.Adapters().Counts(seg).at(seq) ){
					sim.ResetSystematicErrorCounters(ref);
					sim.SetSystematicErrors(sim.adapter_sys_error_.at(seg).at(seq), stats.Adapters().Sequence(seg, seq), 0, length(stats.Adapters().Sequence(seg, seq)), stats, est);
				}
			}
		}
		sim.sys_from_file_ = false;

		w.d("sim.average_read_length", average_read_length);
		w.d("sim.adapter_part", adapter_part);
		w.s64("sim.total_pairs", sim.total_pairs_);
		w.s64("sim.num_adapter_only_pairs", sim.num_adapter_only_pairs_);
		w.d("sim.bias_normalization", sim.bias_normalization_);
		w.s64("sim.sys_gc_range", sim.sys_gc_range_);
		w.u32("sim.coverage_groups", std::vector<uint32_t>(sim.coverage_groups_.begin(), sim.coverage_groups_.end()));
		w.s64("sim.num_groups", sim.non_zero_thresholds_.size());
		for(size_t g=0; g<sim.non_zero_thresholds_.size(); ++g){
			std::vector<double> thr;
			for(auto &t : sim.non_zero_thresholds_.at(g)){ thr.push_back(t.at(0)); thr.push_back(t.at(1)); }
			w.f64("sim.thresholds." + std::to_string(g), thr);
		}
		for(int seg=0; seg<2; ++seg){
			for(size_t a=0; a<sim.adapter_sys_error_.at(seg).size(); ++a){
				std::vector<uint8_t> e;
				for(auto &p : sim.adapter_sys_error_.at(seg).at(a)){ e.push_back(static_cast<uint8_t>(p.first)); e.push_back(p.second); }
				w.u8("sim.adapter_sys_error." + std::to_string(seg) + "." + std::to_string(a), e);
			}
		}
		// replaced reference (after ReplaceN), concatenated codes
		for(uintRefSeqId s=0; s<ref.NumberSequences(); ++s){
			std::vector<uint8_t> codes;
			codes.reserve(ref.SequenceLength(s));
			for(auto c : ref.ReferenceSequence(s)){ codes.push_back(static_cast<uint8_t>(c)); }
			w.u8("sim.ref." + std::to_string(s), codes);
			w.str("sim.ref_id." + std::to_string(s), std::string(seqan::toCString(ref.ReferenceId(s))));
		}

		// --- all blocks, in creation order (Simulator.cpp:2823-2826 + GetNextBlock) ---
		std::vector<uint64_t> seeds; std::vector<int64_t> ids, starts, refids;
		std::vector<uint8_t> fwd, rev;
		// SysErrorVariants of every block, forward then its reverse partner: per variant {block index, strand (0 forward, 1 reverse), position_,
		// allele bits 0-63, allele bits 64-127, number of var_errors_}, the (dominant error, rate) pairs concatenated in var_errs
		std::vector<int64_t> var_rec, first_var_fwd, first_var_rev; std::vector<uint8_t> var_errs;
		size_t nblocks = 0;
		while(nblocks < max_blocks && sim.CreateBlock(ref, stats, est)){
			++nblocks;
			// the reference pages the variants of later sequences in from GetNextBlock (Simulator.cpp:1276-1285); CreateUnit loads what it needs itself (938)
		}
		size_t bi = 0;
		for(auto unit = sim.first_unit_; unit; unit = unit->next_unit_){
			for(Simulator::SimBlock *b = unit->first_block_; b; b = b->next_block_, ++bi){
				seeds.push_back(b->seed_); ids.push_back(b->id_); starts.push_back(b->start_pos_); refids.push_back(unit->ref_seq_id_);
				for(auto &p : b->sys_errors_){ fwd.push_back(static_cast<uint8_t>(p.first)); fwd.push_back(p.second); }
				// partner (reverse) block: its sys_errors_ run in reverse-strand order for the same forward interval
				for(auto &p : b->partner_block_->sys_errors_){ rev.push_back(static_cast<uint8_t>(p.first)); rev.push_back(p.second); }
				first_var_fwd.push_back(b->first_variant_id_); first_var_rev.push_back(b->partner_block_->first_variant_id_);
				for(int strand = 0; strand < 2; ++strand){
					for(auto &v : (strand ? b->partner_block_ : b)->err_variants_){
						var_rec.push_back(bi); var_rec.push_back(strand); var_rec.push_back(v.position_);
						var_rec.push_back(static_cast<int64_t>(v.allele_[0])); var_rec.push_back(static_cast<int64_t>(v.allele_[1])); var_rec.push_back(v.var_errors_.size());
						for(auto &e : v.var_errors_){ var_errs.push_back(static_cast<uint8_t>(e.first)); var_errs.push_back(e.second); }
					}
				}
			}
		}
		w.i64("sim.err_variants", var_rec);
		w.u8("sim.err_variant_errors", var_errs);
		w.i64("sim.block_first_variant_fwd", first_var_fwd);
		w.i64("sim.block_first_variant_rev", first_var_rev);
		w.u64("sim.block_seed", seeds);
		w.i64("sim.block_id", ids);
		w.i64("sim.block_start", starts);
		w.i64("sim.block_ref", refids);
		w.u8("sim.sys_fwd", fwd);
		w.u8("sim.sys_rev", rev);
		w.s64("sim.adapter_only_seed_next", 0);
		return 0;
	}
	return 2;
}
