// Loader for the reference's own profile files: X.reseq (DataStats) and X.reseq.ipf (ProbabilityEstimates),
// both Boost.Serialization text archives, followed by the derived state the simulation uses:
//   DataStats::Load + PrepareProcessing            reference DataStats.cpp:1280-1328, 698-703
//   AdapterStats::SumCounts / PrepareSimulation     AdapterStats.cpp:841-908
//   ErrorStats::PrepareSimulation                   ErrorStats.cpp:202-209
//   ProbabilityEstimates::Load + PrepareResult      ProbabilityEstimates.cpp:961-1045
//   LogIPF::FullExpansion, LogArrayCalc::Expand     ProbabilityEstimates.h:1004-1036, 251-289
//   LogArrayResult::GetResults / ImputeMissingValues ProbabilityEstimates.h:386-479
//
// Text archive token rules (Boost 1.6x/1.7x text_oarchive): header "22 serialization::archive <ver>", blank
// separated tokens; a class type (anything with serialize(), std::pair, std::array, std::vector<non-arithmetic>)
// writes "<tracking> <version>" the first time that exact C++ type is met; vectors write count + item_version;
// arrays write N; strings write length + raw characters.  The member lists below restate the serialize()
// members of the reference classes (file:line given per struct) - they ARE the file format.
//
// Not reproduced: the IPF refit the reference runs when a loaded table has not reached the precision aim
// (ProbabilityEstimates.h:1050-1168).  Such profiles are rejected with a message instead of being refitted.
#pragma once
#include <cstdlib>
#include <set>
#include <utility>
#include "host_profile.hpp"

namespace rsq {
namespace archive {

class TextIn {
	std::vector<char> buf_;
	const char *p_ = nullptr, *end_ = nullptr;
	std::set<const void *> seen_;
	std::string path_;
	uint64_t library_version_ = 0;

	void skip_ws(){ while(p_ < end_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\r' || *p_ == '\t')){ ++p_; } }
	[[noreturn]] void fail(const char *what) const { throw std::runtime_error("Could not load '" + path_ + "': input stream error (" + what + ")"); }
	template<class T> static const void *type_key(){ static const char k = 0; return &k; }

public:
	explicit TextIn(const std::string &path) : path_(path) {
		std::ifstream f(path, std::ios::binary | std::ios::ate);
		if(!f){ throw std::runtime_error("File '" + path + "' does not exists or no read permission given."); }
		const std::streamsize n = f.tellg();
		f.seekg(0);
		buf_.resize(static_cast<size_t>(n) + 1);
		if(n && !f.read(buf_.data(), n)){ fail("read"); }
		buf_[n] = '\0';
		p_ = buf_.data(); end_ = p_ + n;
		const uint64_t siglen = u64();
		skip_ws();
		if(siglen != 22 || static_cast<size_t>(end_ - p_) < 22 || std::string(p_, 22) != "serialization::archive"){ fail("invalid signature"); }
		p_ += 22;
		// Archive library version (Boost 1.53 .. 1.8x write 10 .. 20).  The token rules below were checked against archives of version 17
		// written by the oracle build's own text_oarchive stand-in (oracle/boost_shim), never against a file from a real Boost build:
		// anything outside the range whose text dialect is documented to be the same is refused instead of being parsed on a guess.
		library_version_ = u64();
		if(library_version_ < 10 || library_version_ > 20){
			throw std::runtime_error("Could not load '" + path_ + "': Boost archive library version " + std::to_string(library_version_) + " is outside 10..20, the text-archive dialect this loader restates");
		}
	}
	uint64_t library_version() const { return library_version_; }
	// Self-check after the last member: every serialize() list restated here must have consumed the file exactly.  Left-over tokens mean the
	// file was written by another revision of the reference (other members) or in another archive dialect.
	void expect_end(){
		skip_ws();
		if(p_ < end_){ throw std::runtime_error("Could not load '" + path_ + "': " + std::to_string(end_ - p_) + " bytes are left behind the last member (written by another ReSeq revision or Boost archive dialect?)"); }
	}
	uint64_t u64(){
		skip_ws();
		if(p_ >= end_){ fail("unexpected end"); }
		char *e = nullptr;
		const unsigned long long v = std::strtoull(p_, &e, 10);
		if(e == p_){ fail("integer"); }
		p_ = e;
		return v;
	}
	int64_t i64(){
		skip_ws();
		if(p_ >= end_){ fail("unexpected end"); }
		char *e = nullptr;
		const long long v = std::strtoll(p_, &e, 10);
		if(e == p_){ fail("integer"); }
		p_ = e;
		return v;
	}
	double f64(){
		skip_ws();
		if(p_ >= end_){ fail("unexpected end"); }
		char *e = nullptr;
		const double v = std::strtod(p_, &e);
		if(e == p_){ fail("float"); }
		p_ = e;
		return v;
	}
	template<class T> void class_info(){
		if(seen_.insert(type_key<T>()).second){ (void)u64(); (void)u64(); }
	}

	template<class T> typename std::enable_if<std::is_integral<T>::value && std::is_unsigned<T>::value>::type get(T &v){ v = static_cast<T>(u64()); }
	template<class T> typename std::enable_if<std::is_integral<T>::value && std::is_signed<T>::value>::type get(T &v){ v = static_cast<T>(i64()); }
	void get(double &v){ v = f64(); }
	void get(std::string &s){
		const uint64_t n = u64();
		if(p_ < end_){ ++p_; }   // separating blank
		if(static_cast<uint64_t>(end_ - p_) < n){ fail("string"); }
		s.assign(p_, n);
		p_ += n;
	}
	template<class T> void get(std::vector<T> &v){
		if(!std::is_arithmetic<T>::value){ class_info<std::vector<T>>(); }
		const uint64_t n = u64();
		(void)u64();   // item_version
		v.clear();
		v.resize(n);
		for(auto &e : v){ get(e); }
	}
	void get(std::vector<bool> &v){
		const uint64_t n = u64();
		(void)u64();
		v.assign(n, false);
		for(uint64_t i = 0; i < n; ++i){ v[i] = u64() != 0; }
	}
	template<class T, size_t N> void get(std::array<T, N> &a){
		class_info<std::array<T, N>>();
		if(u64() != N){ fail("array size mismatch"); }
		for(auto &e : a){ get(e); }
	}
	template<class A, class B> void get(std::pair<A, B> &pr){
		class_info<std::pair<A, B>>();
		get(pr.first); get(pr.second);
	}
	template<class T> typename std::enable_if<std::is_class<T>::value>::type get(T &t){
		class_info<T>();
		t.serialize(*this);
	}
	template<class T> TextIn &operator&(T &t){ get(t); return *this; }
};

// reseq::Vect<T> (Vect.hpp:21,40-42)
template<class T> struct AVect {
	std::pair<size_t, std::vector<T>> vec;
	void serialize(TextIn &ar){ ar & vec; }
	size_t from() const { return vec.first; }
	size_t to() const { return vec.first + vec.second.size(); }
};
// reseq::SeqQualityStats<T> (SeqQualityStats.hpp:19-21)
template<class T> struct ASeqQual { AVect<T> qualities; void serialize(TextIn &ar){ ar & qualities; } };

typedef uint64_t u64;
typedef uint16_t u16;

struct AAdapterStats {                       // AdapterStats.h:59-92
	std::array<std::vector<std::string>, 2> names;
	std::vector<std::vector<bool>> combinations;
	std::vector<std::vector<AVect<AVect<u64>>>> counts;
	std::array<std::vector<AVect<u64>>, 2> start_cut;
	AVect<u64> polya_tail_length;
	std::array<u64, 5> overrun_bases;
	std::array<std::vector<std::string>, 2> seqs;
	void serialize(TextIn &ar){ ar & names & combinations & counts & start_cut & polya_tail_length & overrun_bases & seqs; }
};
struct ACoverageStats {                      // CoverageStats.h:281-315
	uint32_t coverage_threshold = 0, reset_distance = 0;
	typedef std::array<std::array<std::array<AVect<AVect<u64>>, 4>, 5>, 4> DomErr;
	typedef std::array<std::array<AVect<AVect<u64>>, 5>, 4> ErrRate;
	DomErr de[6];
	ErrRate er[6];
	AVect<u16> block_error_rate, block_percent_systematic;
	AVect<u64> systematic_error_p_values, coverage;
	std::array<AVect<u64>, 2> stranded[4];
	AVect<u64> error_coverage[4];
	AVect<AVect<u64>> error_coverage_stranded[3];
	void serialize(TextIn &ar){
		ar & coverage_threshold & reset_distance;
		for(auto &x : de){ ar & x; }
		for(auto &x : er){ ar & x; }
		ar & block_error_rate & block_percent_systematic & systematic_error_p_values & coverage;
		for(auto &x : stranded){ ar & x; }
		for(auto &x : error_coverage){ ar & x; }
		for(auto &x : error_coverage_stranded){ ar & x; }
	}
};
struct AErrorStats {                         // ErrorStats.h:78-96
	typedef std::array<std::array<std::array<AVect<AVect<AVect<u64>>>, 5>, 4>, 2> PerTile;
	typedef std::array<std::array<AVect<AVect<u64>>, 6>, 2> InDel;
	PerTile per_tile[7];
	InDel indel[6];   // [0] = indel_by_indel_pos_
	std::array<AVect<u64>, 2> errors_per_read;
	std::array<std::array<std::array<std::array<AVect<u64>, 6>, 5>, 4>, 2> called_bases_by_base_quality_per_previous_called_base;
	void serialize(TextIn &ar){
		for(auto &x : per_tile){ ar & x; }
		for(auto &x : indel){ ar & x; }
		ar & errors_per_read & called_bases_by_base_quality_per_previous_called_base;
	}
};
struct ADuplicationStats { AVect<u64> duplication_number; void serialize(TextIn &ar){ ar & duplication_number; } };   // FragmentDuplicationStats.h:33-35
struct ASurroundingCount { std::array<std::vector<u64>, 3> counts; void serialize(TextIn &ar){ ar & counts; } };     // Surrounding.h:63-65
struct ASurroundingBias { std::array<std::vector<double>, 3> bias; void serialize(TextIn &ar){ ar & bias; } };       // Surrounding.h:89-91
struct AFragmentDistributionStats {          // FragmentDistributionStats.h:440-456
	std::vector<u64> abundance;
	AVect<u64> insert_lengths, gc_fragment_content;
	ASurroundingCount fragment_surroundings;
	AVect<AVect<u64>> site_count;
	std::array<std::array<AVect<u64>, 4>, 2> outskirt_content;
	std::vector<double> ref_seq_bias;
	AVect<double> insert_lengths_bias, gc_fragment_content_bias;
	ASurroundingBias fragment_surroundings_bias;
	std::array<double, 2> dispersion_parameters;
	void serialize(TextIn &ar){
		ar & abundance & insert_lengths & gc_fragment_content & fragment_surroundings & site_count & outskirt_content
		   & ref_seq_bias & insert_lengths_bias & gc_fragment_content_bias & fragment_surroundings_bias & dispersion_parameters;
	}
};
struct AQualityStats {                       // QualityStats.h:156-197 (types l.65-107)
	typedef AVect<AVect<AVect<u64>>> V3;
	typedef AVect<AVect<ASeqQual<u64>>> VSQ;
	std::array<std::array<std::array<VSQ, 5>, 4>, 2> base_quality_stats_per_tile_per_error_reference;
	std::array<std::array<std::array<V3, 5>, 4>, 2> error_rate_for_position, base_quality_for_error_rate;
	std::array<std::array<V3, 4>, 2> ref7[7];
	std::array<VSQ, 2> sequence_quality_mean_for_gc_per_tile_reference;
	std::array<V3, 2> seq5[5];
	std::array<std::array<V3, 5>, 2> base_quality_for_sequence_per_tile, base_quality_for_preceding_quality_per_tile;
	std::array<std::array<VSQ, 5>, 2> base_quality_stats_per_tile;
	std::array<std::array<V3, 5>, 2> raw3[3];
	std::array<AVect<ASeqQual<u64>>, 2> base_quality_stats_per_strand;
	std::array<std::array<VSQ, 5>, 2> sequence_quality_for_base_per_tile;
	V3 sequence_quality_mean_paired_per_tile;
	std::array<VSQ, 2> sequence_quality_mean_for_gc_per_tile;
	std::array<AVect<u64>, 2> sq6[6];
	std::array<AVect<AVect<u64>>, 2> sequence_quality_content;
	AVect<AVect<u64>> homoquality_distribution;
	std::array<std::array<ASeqQual<u64>, 5>, 2> nucleotide_quality;
	void serialize(TextIn &ar){
		ar & base_quality_stats_per_tile_per_error_reference & error_rate_for_position & base_quality_for_error_rate;
		for(auto &x : ref7){ ar & x; }
		ar & sequence_quality_mean_for_gc_per_tile_reference;
		for(auto &x : seq5){ ar & x; }
		ar & base_quality_for_sequence_per_tile & base_quality_for_preceding_quality_per_tile & base_quality_stats_per_tile;
		for(auto &x : raw3){ ar & x; }
		ar & base_quality_stats_per_strand & sequence_quality_for_base_per_tile & sequence_quality_mean_paired_per_tile & sequence_quality_mean_for_gc_per_tile;
		for(auto &x : sq6){ ar & x; }
		ar & sequence_quality_content & homoquality_distribution & nucleotide_quality;
	}
};
struct ATileStats { std::vector<u16> tiles; std::vector<u64> abundance; void serialize(TextIn &ar){ ar & tiles & abundance; } };   // TileStats.h:42-52
struct ADataStats {                          // DataStats.h:180-212
	AAdapterStats adapters; ACoverageStats coverage; AErrorStats errors; ADuplicationStats duplicates;
	AFragmentDistributionStats fragment_distribution; AQualityStats qualities; ATileStats tiles;
	u64 creation_time = 0;
	std::array<AVect<u64>, 2> read_lengths;
	std::array<AVect<AVect<u64>>, 2> read_lengths_by_fragment_length, non_mapped_read_lengths_by_fragment_length;
	uint8_t phred_quality_offset = 0, minimum_quality = 0, maximum_quality = 0;
	u16 minimum_read_length_on_reference = 0, maximum_read_length_on_reference = 0;
	double corrected_coverage = 0;
	void serialize(TextIn &ar){
		ar & adapters & coverage & errors & duplicates & fragment_distribution & qualities & tiles;
		ar & creation_time & read_lengths & read_lengths_by_fragment_length & non_mapped_read_lengths_by_fragment_length;
		ar & phred_quality_offset & minimum_quality & maximum_quality & minimum_read_length_on_reference & maximum_read_length_on_reference & corrected_coverage;
		// the plotting-only members that follow in the file are not needed and not read
	}
};

// ProbabilityEstimates (ProbabilityEstimates.h:1475-1483), LogIPF (955-967), LogArrayCalc (111-114)
template<unsigned N> struct ALogArrayCalc {
	static const unsigned kM = N * (N - 1) / 2;
	std::array<std::vector<double>, kM> dim2;
	std::array<uint32_t, N> dim_size;
	void serialize(TextIn &ar){ ar & dim2 & dim_size; }
};
template<unsigned N> struct ALogIPF {
	static const unsigned kM = N * (N - 1) / 2;
	uint32_t steps = 0, needed_updates = 0;
	double precision = 0;
	std::array<double, kM> margin_precision;
	u16 last_margin = 0;
	std::array<uint32_t, kM> last_update;
	std::array<u16, kM> update_dist;
	ALogArrayCalc<N> estimates;
	std::array<std::vector<uint32_t>, N> dim_indices, initial_dim_indices_reduced, dim_indices_reduced;
	void serialize(TextIn &ar){
		ar & steps & needed_updates & precision & margin_precision & last_margin & last_update & update_dist & estimates
		   & dim_indices & initial_dim_indices_reduced & dim_indices_reduced;
	}
};
struct AProbabilityEstimates {
	u64 stats_creation_time = 0;
	std::array<std::vector<std::array<ALogIPF<5>, 4>>, 2> quality;
	std::array<std::vector<ALogIPF<4>>, 2> sequence_quality;
	std::array<std::vector<std::array<std::array<ALogIPF<5>, 5>, 4>>, 2> base_call;
	std::array<std::array<std::array<ALogIPF<4>, 5>, 5>, 4> dom_error;
	std::array<std::array<ALogIPF<4>, 5>, 4> error_rate;
	std::array<std::array<ALogIPF<4>, 6>, 2> indels;
	void serialize(TextIn &ar){ ar & stats_creation_time & quality & sequence_quality & base_call & dom_error & error_rate & indels; }
};

template<class T> OffsetVec<T> to_offset(const AVect<T> &v){ OffsetVec<T> o; o.from = v.vec.first; o.v = v.vec.second; return o; }
inline OffsetVec<OffsetVec<u64>> to_offset2(const AVect<AVect<u64>> &v){
	OffsetVec<OffsetVec<u64>> o; o.from = v.vec.first;
	for(const auto &r : v.vec.second){ o.v.push_back(to_offset(r)); }
	return o;
}

// LogIPF::FullExpansion + LogArrayResult::GetResults + ImputeMissingValues
template<unsigned N> HostTable make_result(ALogIPF<N> ipf, double precision_aim, const std::string &what){
	if(ipf.steps && ipf.precision > precision_aim){
		throw std::runtime_error("probability table '" + what + "' has not reached the precision aim (" + std::to_string(ipf.precision * 100) +
		                         "%): the reference would continue the iterative proportional fitting on load; finish the fit with `reseq illuminaPE --stopAfterEstimation` first");
	}
	// FullExpansion
	bool expansion_necessary = false;
	for(unsigned n = N; n-- && !expansion_necessary; ){
		for(size_t ind = ipf.dim_indices_reduced[n].size(); ind--; ){ if(ind != ipf.dim_indices_reduced[n][ind]){ expansion_necessary = true; break; } }
	}
	for(unsigned n = N; n-- && !expansion_necessary; ){
		for(size_t ind = ipf.initial_dim_indices_reduced[n].size(); ind--; ){ if(ind != ipf.initial_dim_indices_reduced[n][ind]){ expansion_necessary = true; break; } }
	}
	if(expansion_necessary){
		std::array<std::vector<uint32_t>, N> reduced = ipf.initial_dim_indices_reduced, count;
		for(unsigned d = 0; d < N; ++d){ for(auto &b : reduced[d]){ b = ipf.dim_indices_reduced[d].at(b); } }          // CombineDimIndices
		for(unsigned n = N; n--; ){                                                                                 // ReconstructDimIndicesCount
			count[n].assign(*std::max_element(reduced[n].begin(), reduced[n].end()) + 1, 0);
			for(auto ind : reduced[n]){ ++count[n].at(ind); }
		}
		// LogArrayCalc::Expand(reduced, count)
		std::array<std::vector<double>, N> mult;
		for(unsigned n = N; n--; ){
			mult[n].resize(reduced[n].size());
			for(size_t i = reduced[n].size(); i--; ){ mult[n][i] = std::pow(1.0 / count[n].at(reduced[n][i]), 1.0 / (N - 1)); }
		}
		unsigned dim_a = N, dim_b = N - 1;
		for(unsigned n = ALogIPF<N>::kM; n--; ){
			if(--dim_a == dim_b){ --dim_b; dim_a = N - 1; }
			std::vector<double> old_values = std::move(ipf.estimates.dim2[n]);
			ipf.estimates.dim2[n].assign(reduced[dim_a].size() * reduced[dim_b].size(), 0.0);
			for(size_t i = reduced[dim_a].size(); i--; ){
				for(size_t j = reduced[dim_b].size(); j--; ){
					ipf.estimates.dim2[n][i * reduced[dim_b].size() + j] = old_values.at(reduced[dim_a][i] * count[dim_b].size() + reduced[dim_b][j]) * mult[dim_a][i] * mult[dim_b][j];
				}
			}
		}
		for(unsigned n = N; n--; ){ ipf.estimates.dim_size[n] = reduced[n].size(); }
	}
	// GetResults
	HostTable h;
	h.nm = N - 1;
	const auto &di = ipf.dim_indices;
	const size_t n0 = di[0].size();
	if(!n0){ return h; }
	std::vector<std::pair<double, uint32_t>> order(n0);
	for(size_t j = n0; j--; ){ order[j] = {0.0, static_cast<uint32_t>(j)}; }
	for(unsigned n = N - 1; n--; ){
		const unsigned dim_a = n + 1;
		for(size_t j = n0; j--; ){
			double sum = 0.0;
			for(size_t i = di[dim_a].size(); i--; ){ sum += ipf.estimates.dim2[n].at(i * n0 + j); }
			order[j].first += sum / di[dim_a].size();
		}
	}
	std::sort(order.begin(), order.end());
	std::vector<uint32_t> par0(n0);
	for(size_t k = n0; k--; ){ par0.at(order[k].second) = k; }
	for(unsigned n = N - 1; n--; ){
		uint32_t lo = UINT32_MAX, hi = 0;
		for(auto ind : di[n + 1]){ if(ind < lo){ lo = ind; } if(ind > hi){ hi = ind; } }
		h.from[n] = lo; h.to[n] = hi + 1;
	}
	for(unsigned n = N - 1; n--; ){
		const unsigned dim_a = n + 1;
		h.dim2[n].assign(static_cast<size_t>(h.to[n] - h.from[n]) * n0, 0.0);
		for(size_t i = di[dim_a].size(); i--; ){
			for(size_t j = n0; j--; ){
				h.dim2[n].at((di[dim_a][i] - h.from[n]) * n0 + par0[j]) = ipf.estimates.dim2[n].at(i * n0 + j);
			}
		}
	}
	h.par0.resize(n0);
	for(size_t k = n0; k--; ){ h.par0[k] = di[0].at(order[k].second); }
	// ImputeMissingValues
	for(unsigned n = N - 1; n--; ){
		uint32_t last = 0;
		for(uint32_t i = 1; i < h.to[n] - h.from[n]; ++i){
			bool filled = false;
			for(size_t j = 0; j < n0; ++j){ if(0.0 != h.dim2[n][i * n0 + j]){ filled = true; break; } }
			if(filled){
				for(uint32_t imp = last + 1; imp < i; ++imp){
					for(size_t j = 0; j < n0; ++j){
						h.dim2[n][imp * n0 + j] = h.dim2[n][last * n0 + j] * (imp - last) / (i - last) + h.dim2[n][i * n0 + j] * (i - imp) / (i - last);
					}
				}
				last = i;
			}
		}
	}
	return h;
}

}  // namespace archive

inline void load_reseq_profile(Profile &p, const char *stats_path, const char *ipf_path, double ipf_precision_percent = 5.0){
	using namespace archive;
	std::unique_ptr<ADataStats> ds(new ADataStats);
	{
		TextIn in(stats_path);
		in & *ds;   // (no expect_end: the plotting-only members behind corrected_coverage_ are deliberately not restated)
	}
	// --- DataStats members + PrepareProcessing ---
	for(int seg = 0; seg < 2; ++seg){
		p.read_lengths[seg] = to_offset(ds->read_lengths[seg]);
		p.read_lengths_by_fragment_length[seg] = to_offset2(ds->read_lengths_by_fragment_length[seg]);
		p.non_mapped_read_lengths_by_fragment_length[seg] = to_offset2(ds->non_mapped_read_lengths_by_fragment_length[seg]);
		p.adapter_seqs[seg] = ds->adapters.seqs[seg];
		p.adapter_start_cut[seg].clear();
		for(const auto &v : ds->adapters.start_cut[seg]){ p.adapter_start_cut[seg].push_back(to_offset(v)); }
	}
	p.phred_quality_offset = ds->phred_quality_offset;
	p.corrected_coverage = ds->corrected_coverage;
	p.creation_time = ds->creation_time;
	p.reset_distance = ds->coverage.reset_distance;
	p.tiles = ds->tiles.tiles;
	p.tile_abundance = ds->tiles.abundance;
	p.polya_tail_length = to_offset(ds->adapters.polya_tail_length);
	p.overrun_bases = ds->adapters.overrun_bases;
	p.insert_lengths = to_offset(ds->fragment_distribution.insert_lengths);
	p.ref_seq_bias = ds->fragment_distribution.ref_seq_bias;
	p.insert_lengths_bias = to_offset(ds->fragment_distribution.insert_lengths_bias);
	p.gc_fragment_content_bias = to_offset(ds->fragment_distribution.gc_fragment_content_bias);
	p.fragment_surroundings_bias = ds->fragment_distribution.fragment_surroundings_bias.bias;
	p.dispersion_parameters = ds->fragment_distribution.dispersion_parameters;
	// PrepareGeneral: total_number_reads_
	p.total_number_reads = 0;
	for(int seg = 0; seg < 2; ++seg){ for(auto v : ds->read_lengths[seg].vec.second){ p.total_number_reads += v; } }
	// AdapterStats::SumCounts
	{
		const auto &A = ds->adapters;
		for(int seg = 2; seg--; ){ p.adapter_count_sum[seg].assign(A.start_cut[seg].size(), 0); }
		auto common = [](const std::string &a, const std::string &b){ uint16_t k = 0; while(k < std::min(a.size(), b.size()) && a[k] == b[k]){ ++k; } return k; };
		uint16_t before_a1 = 0;
		for(size_t a1 = A.counts.size(); a1--; ){
			const uint16_t after_a1 = a1 ? common(A.seqs[0].at(a1), A.seqs[0].at(a1 - 1)) : 0;
			uint16_t before_a2 = 0;
			for(size_t a2 = A.counts.at(0).size(); a2--; ){
				const uint16_t after_a2 = a2 ? common(A.seqs[1].at(a2), A.seqs[1].at(a2 - 1)) : 0;
				const auto &c = A.counts[a1].at(a2);
				u64 sum = 0;
				for(size_t pos1 = std::max<uint16_t>(std::max(before_a1, after_a1), static_cast<uint16_t>(c.from())); pos1 < c.to(); ++pos1){
					const auto &row = c.vec.second.at(pos1 - c.from());
					for(size_t pos2 = std::max<uint16_t>(std::max(before_a2, after_a2), static_cast<uint16_t>(row.from())); pos2 < row.to(); ++pos2){
						sum += row.vec.second.at(pos2 - row.from());
					}
				}
				p.adapter_count_sum[0].at(a1) += sum;
				p.adapter_count_sum[1].at(a2) += sum;
				before_a2 = after_a2;
			}
			before_a1 = after_a1;
		}
		// PrepareSimulation: adapters below 10% of the most frequent one are not simulated
		for(int seg = 2; seg--; ){
			const auto &cs = p.adapter_count_sum[seg];
			p.adapter_significant_count[seg].assign(cs.size(), 0);
			if(cs.empty()){ continue; }
			const u64 threshold = std::ceil(*std::max_element(cs.begin(), cs.end()) * 0.1);
			for(size_t i = cs.size(); i--; ){ p.adapter_significant_count[seg][i] = cs[i] < threshold ? 0 : cs[i]; }
		}
	}
	// ErrorStats::PrepareSimulation
	p.max_len_deletion = 0;
	for(const auto &v : ds->errors.indel[0].at(1)){ if(v.to() > p.max_len_deletion){ p.max_len_deletion = v.to(); } }
	ds.reset();

	// --- ProbabilityEstimates ---
	std::unique_ptr<AProbabilityEstimates> pe(new AProbabilityEstimates);
	{
		TextIn in(ipf_path);
		in & *pe;
		in.expect_end();
	}
	if(pe->stats_creation_time != p.creation_time){
		throw std::runtime_error(std::string("'") + ipf_path + "' was estimated for a different statistics file (creation time mismatch): the reference would discard it and refit");
	}
	const double aim = ipf_precision_percent / 100;
	const size_t T = pe->quality[0].size();
	p.num_tiles = T;
	p.tables.clear();
	for(int seg = 0; seg < 2; ++seg) for(size_t t = 0; t < T; ++t) for(int b = 0; b < 4; ++b) p.tables.push_back(make_result<5>(pe->quality[seg].at(t)[b], aim, "quality"));
	for(int seg = 0; seg < 2; ++seg) for(size_t t = 0; t < T; ++t) p.tables.push_back(make_result<4>(pe->sequence_quality[seg].at(t), aim, "sequence quality"));
	for(int seg = 0; seg < 2; ++seg) for(size_t t = 0; t < T; ++t) for(int b = 0; b < 4; ++b) for(int d = 0; d < 5; ++d) p.tables.push_back(make_result<5>(pe->base_call[seg].at(t)[b][d], aim, "base call"));
	for(int b = 0; b < 4; ++b) for(int l = 0; l < 5; ++l) for(int d = 0; d < 5; ++d) p.tables.push_back(make_result<4>(pe->dom_error[b][l][d], aim, "dominant error"));
	for(int b = 0; b < 4; ++b) for(int d = 0; d < 5; ++d) p.tables.push_back(make_result<4>(pe->error_rate[b][d], aim, "error rate"));
	for(int ty = 0; ty < 2; ++ty) for(int c = 0; c < 6; ++c) p.tables.push_back(make_result<4>(pe->indels[ty][c], aim, "indels"));
}

}  // namespace rsq
