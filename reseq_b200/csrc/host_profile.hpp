// Host-side model of what the simulation needs from a ReSeq profile (DataStats + ProbabilityEstimates)
// and from the reference genome, plus the Simulate() prologue arithmetic that stays on the host
// (pair counts, spline interpolation of the normalisation, non-zero thresholds).
//
// Mirrors (reference file:line):
//   DataStats getters used by Simulator                       DataStats.h:224-257
//   Simulator::CoveragePropLostFromAdapters / CoverageToNumberPairs   Simulator.cpp:90-108
//   FragmentDistributionStats::CalculateBiasNormalization      FragmentDistributionStats.cpp:3504-3582
//   InsertLengthSpline::{GetSamplePositions,GetSampledValues,PrepareInsertLengthSpline,SetStartingParameters,
//                        FillInWithFittedRatios}               FragmentDistributionStats.cpp:1217-1330,1539-1567,1656-1705
//   BiasCalculationVectors::{PrepareSplines,GetSplineCoefficients}     FragmentDistributionStats.cpp:418-496
//   FragmentDistributionStats::SplitCoverageGroups / CalculateNonZeroThreshold   FragmentDistributionStats.cpp:2909-2976
//   Reference::ReadFasta / ReplaceN                            Reference.cpp:758-891
//   std::discrete_distribution::param_type::_M_initialize      libstdc++ bits/random.tcc:2657-2680
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <numeric>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include "text_io.hpp"
#include "variants.hpp"

namespace rsq {

// ---------------------------------------------------------------------------------------------------
// RSQFLAT1 container
// ---------------------------------------------------------------------------------------------------
struct FlatArray {
	uint32_t dtype = 0;   // 0=u8 1=u32 2=u64 3=f64 4=i64
	uint64_t count = 0;
	std::vector<unsigned char> bytes;
	template<class T> const T *as() const { return reinterpret_cast<const T *>(bytes.data()); }
};

struct FlatFile {
	std::map<std::string, FlatArray> arrays;
	std::vector<std::string> order;

	static size_t elem(uint32_t dtype){ return dtype == 0 ? 1 : dtype == 1 ? 4 : 8; }

	void load(const std::string &path){
		std::ifstream f(path, std::ios::binary);
		if(!f){ throw std::runtime_error("cannot open " + path); }
		char magic[8];
		f.read(magic, 8);
		if(!f || std::memcmp(magic, "RSQFLAT1", 8) != 0){ throw std::runtime_error(path + " is not an RSQFLAT1 file"); }
		while(true){
			uint32_t nl;
			f.read(reinterpret_cast<char *>(&nl), 4);
			if(!f){ break; }
			std::string name(nl, '\0');
			f.read(&name[0], nl);
			FlatArray a;
			f.read(reinterpret_cast<char *>(&a.dtype), 4);
			f.read(reinterpret_cast<char *>(&a.count), 8);
			a.bytes.resize(a.count * elem(a.dtype));
			if(a.count){ f.read(reinterpret_cast<char *>(a.bytes.data()), a.bytes.size()); }
			if(!f){ throw std::runtime_error(path + ": truncated record " + name); }
			order.push_back(name);
			arrays.emplace(name, std::move(a));
		}
	}
	void save(const std::string &path) const {
		std::ofstream f(path, std::ios::binary);
		if(!f){ throw std::runtime_error("cannot write " + path); }
		f.write("RSQFLAT1", 8);
		for(const auto &name : order){
			const FlatArray &a = arrays.at(name);
			uint32_t nl = name.size();
			f.write(reinterpret_cast<const char *>(&nl), 4);
			f.write(name.data(), nl);
			f.write(reinterpret_cast<const char *>(&a.dtype), 4);
			f.write(reinterpret_cast<const char *>(&a.count), 8);
			f.write(reinterpret_cast<const char *>(a.bytes.data()), a.bytes.size());
		}
	}
	bool has(const std::string &n) const { return arrays.count(n) != 0; }
	const FlatArray &get(const std::string &n) const {
		auto it = arrays.find(n);
		if(it == arrays.end()){ throw std::runtime_error("flat profile lacks array '" + n + "'"); }
		return it->second;
	}
	template<class T> void put(const std::string &n, uint32_t dtype, const T *data, uint64_t count){
		FlatArray a;
		a.dtype = dtype; a.count = count;
		a.bytes.resize(count * sizeof(T));
		if(count){ std::memcpy(a.bytes.data(), data, a.bytes.size()); }
		if(!arrays.count(n)){ order.push_back(n); }
		arrays[n] = std::move(a);
	}
	int64_t scalar_i(const std::string &n) const { return get(n).as<int64_t>()[0]; }
	double scalar_d(const std::string &n) const { return get(n).as<double>()[0]; }
	std::vector<uint64_t> vec_u64(const std::string &n) const { const auto &a = get(n); return std::vector<uint64_t>(a.as<uint64_t>(), a.as<uint64_t>() + a.count); }
	std::vector<double> vec_f64(const std::string &n) const { const auto &a = get(n); return std::vector<double>(a.as<double>(), a.as<double>() + a.count); }
};

// reseq::Vect<T>: values with an index offset; out-of-range reads give 0 (Vect.hpp:188-195)
template<class T> struct OffsetVec {
	uint64_t from = 0;
	std::vector<T> v;
	uint64_t to() const { return from + v.size(); }
	size_t size() const { return v.size(); }
	T operator[](uint64_t i) const { return (i >= from && i < to()) ? v[i - from] : T(0); }
};

struct HostTable {            // one LogArrayResult<N> (ProbabilityEstimates.h:351-557)
	uint32_t nm = 0;
	std::array<uint32_t, 4> from{{0, 0, 0, 0}}, to{{0, 0, 0, 0}};
	std::array<std::vector<double>, 4> dim2;
	std::vector<uint32_t> par0;
};

struct Profile {
	// DataStats
	std::array<OffsetVec<uint64_t>, 2> read_lengths;
	std::array<OffsetVec<OffsetVec<uint64_t>>, 2> read_lengths_by_fragment_length, non_mapped_read_lengths_by_fragment_length;
	uint32_t phred_quality_offset = 33;
	uint64_t total_number_reads = 0;
	double corrected_coverage = 0;
	uint64_t creation_time = 0;
	uint32_t reset_distance = 0;
	uint32_t max_len_deletion = 0;
	std::vector<uint16_t> tiles;
	std::vector<uint64_t> tile_abundance;
	// AdapterStats
	std::array<std::vector<std::string>, 2> adapter_seqs;                 // "ACGT" strings
	std::array<std::vector<uint64_t>, 2> adapter_count_sum, adapter_significant_count;
	std::array<std::vector<OffsetVec<uint64_t>>, 2> adapter_start_cut;
	OffsetVec<uint64_t> polya_tail_length;
	std::array<uint64_t, 5> overrun_bases{{0, 0, 0, 0, 0}};
	// FragmentDistributionStats
	OffsetVec<uint64_t> insert_lengths;
	std::vector<double> ref_seq_bias;
	OffsetVec<double> insert_lengths_bias, gc_fragment_content_bias;
	std::array<std::vector<double>, 3> fragment_surroundings_bias;
	std::array<double, 2> dispersion_parameters{{0, 0}};
	// ProbabilityEstimates results, fixed family order (see core.cuh Tables)
	uint32_t num_tiles = 1;
	std::vector<HostTable> tables;

	static OffsetVec<uint64_t> read_vect(const FlatFile &f, const std::string &n){
		OffsetVec<uint64_t> o;
		o.from = f.scalar_i(n + ".from");
		o.v = f.vec_u64(n);
		return o;
	}
	static OffsetVec<double> read_vectd(const FlatFile &f, const std::string &n){
		OffsetVec<double> o;
		o.from = f.scalar_i(n + ".from");
		o.v = f.vec_f64(n);
		return o;
	}
	static OffsetVec<OffsetVec<uint64_t>> read_vect2(const FlatFile &f, const std::string &n){
		OffsetVec<OffsetVec<uint64_t>> o;
		o.from = f.scalar_i(n + ".from");
		const auto &rows = f.get(n + ".rows");
		const auto &vals = f.get(n);
		size_t pos = 0;
		for(uint64_t r = 0; r < rows.count / 2; ++r){
			OffsetVec<uint64_t> row;
			row.from = rows.as<int64_t>()[2 * r];
			uint64_t cnt = rows.as<int64_t>()[2 * r + 1];
			row.v.assign(vals.as<uint64_t>() + pos, vals.as<uint64_t>() + pos + cnt);
			pos += cnt;
			o.v.push_back(std::move(row));
		}
		return o;
	}

	void from_flat(const FlatFile &f){
		for(int seg = 0; seg < 2; ++seg){
			const std::string s = std::to_string(seg);
			read_lengths[seg] = read_vect(f, "read_lengths." + s);
			read_lengths_by_fragment_length[seg] = read_vect2(f, "read_lengths_by_fragment_length." + s);
			non_mapped_read_lengths_by_fragment_length[seg] = read_vect2(f, "non_mapped_read_lengths_by_fragment_length." + s);
			adapter_count_sum[seg] = f.vec_u64("adapter.count_sum." + s);
			adapter_significant_count[seg] = f.vec_u64("adapter.significant_count." + s);
			const int64_t n = f.scalar_i("adapter.n." + s);
			adapter_seqs[seg].clear(); adapter_start_cut[seg].clear();
			for(int64_t a = 0; a < n; ++a){
				const auto &sa = f.get("adapter.seq." + s + "." + std::to_string(a));
				adapter_seqs[seg].emplace_back(reinterpret_cast<const char *>(sa.bytes.data()), sa.count);
				adapter_start_cut[seg].push_back(read_vect(f, "adapter.start_cut." + s + "." + std::to_string(a)));
			}
		}
		polya_tail_length = read_vect(f, "adapter.polya_tail_length");
		{ auto o = f.vec_u64("adapter.overrun_bases"); for(int i = 0; i < 5; ++i){ overrun_bases[i] = o.at(i); } }
		phred_quality_offset = f.scalar_i("phred_quality_offset");
		total_number_reads = f.scalar_i("total_number_reads");
		corrected_coverage = f.scalar_d("corrected_coverage");
		creation_time = f.scalar_i("creation_time");
		reset_distance = f.scalar_i("reset_distance");
		max_len_deletion = f.scalar_i("max_len_deletion");
		{ auto t = f.vec_u64("tiles.tiles"); tiles.assign(t.begin(), t.end()); }
		tile_abundance = f.vec_u64("tiles.abundance");
		insert_lengths = read_vect(f, "insert_lengths");
		ref_seq_bias = f.vec_f64("ref_seq_bias");
		insert_lengths_bias = read_vectd(f, "insert_lengths_bias");
		gc_fragment_content_bias = read_vectd(f, "gc_fragment_content_bias");
		for(int b = 0; b < 3; ++b){ fragment_surroundings_bias[b] = f.vec_f64("fragment_surroundings_bias." + std::to_string(b)); }
		{ auto d = f.vec_f64("dispersion_parameters"); dispersion_parameters = {{d.at(0), d.at(1)}}; }
		num_tiles = f.scalar_i("tab.num_tiles");
		const auto &desc = f.get("tab.desc");
		const auto &blob = f.get("tab.blob");
		const auto &par0 = f.get("tab.par0");
		tables.clear();
		for(uint64_t t = 0; t < desc.count / 16; ++t){
			const int64_t *e = desc.as<int64_t>() + 16 * t;
			HostTable h;
			h.nm = e[1];
			const uint64_t n0 = e[0];
			for(uint32_t n = 0; n < h.nm; ++n){
				h.from[n] = e[2 + n]; h.to[n] = e[6 + n];
				const uint64_t cnt = static_cast<uint64_t>(h.to[n] - h.from[n]) * n0;
				h.dim2[n].assign(blob.as<double>() + e[10 + n], blob.as<double>() + e[10 + n] + cnt);
			}
			h.par0.assign(par0.as<uint32_t>() + e[14], par0.as<uint32_t>() + e[14] + n0);
			tables.push_back(std::move(h));
		}
		const size_t expect = 8 * num_tiles + 2 * num_tiles + 40 * num_tiles + 100 + 20 + 12;
		if(tables.size() != expect){ throw std::runtime_error("flat profile: unexpected number of probability tables"); }
	}

	void to_flat(FlatFile &f) const {
		auto put_i = [&](const std::string &n, int64_t v){ f.put(n, 4, &v, 1); };
		auto put_d = [&](const std::string &n, double v){ f.put(n, 3, &v, 1); };
		auto put_vect = [&](const std::string &n, const OffsetVec<uint64_t> &o){ put_i(n + ".from", o.from); f.put(n, 2, o.v.data(), o.v.size()); };
		auto put_vectd = [&](const std::string &n, const OffsetVec<double> &o){ put_i(n + ".from", o.from); f.put(n, 3, o.v.data(), o.v.size()); };
		auto put_vect2 = [&](const std::string &n, const OffsetVec<OffsetVec<uint64_t>> &o){
			put_i(n + ".from", o.from);
			std::vector<int64_t> rows; std::vector<uint64_t> vals;
			for(const auto &r : o.v){ rows.push_back(r.from); rows.push_back(r.v.size()); vals.insert(vals.end(), r.v.begin(), r.v.end()); }
			f.put(n + ".rows", 4, rows.data(), rows.size());
			f.put(n, 2, vals.data(), vals.size());
		};
		for(int seg = 0; seg < 2; ++seg){
			const std::string s = std::to_string(seg);
			put_vect("read_lengths." + s, read_lengths[seg]);
			put_vect2("read_lengths_by_fragment_length." + s, read_lengths_by_fragment_length[seg]);
			put_vect2("non_mapped_read_lengths_by_fragment_length." + s, non_mapped_read_lengths_by_fragment_length[seg]);
			f.put("adapter.count_sum." + s, 2, adapter_count_sum[seg].data(), adapter_count_sum[seg].size());
			f.put("adapter.significant_count." + s, 2, adapter_significant_count[seg].data(), adapter_significant_count[seg].size());
			put_i("adapter.n." + s, adapter_seqs[seg].size());
			for(size_t a = 0; a < adapter_seqs[seg].size(); ++a){
				f.put("adapter.seq." + s + "." + std::to_string(a), 0, adapter_seqs[seg][a].data(), adapter_seqs[seg][a].size());
				put_vect("adapter.start_cut." + s + "." + std::to_string(a), adapter_start_cut[seg][a]);
			}
		}
		put_vect("adapter.polya_tail_length", polya_tail_length);
		f.put("adapter.overrun_bases", 2, overrun_bases.data(), 5);
		put_i("phred_quality_offset", phred_quality_offset);
		put_i("total_number_reads", total_number_reads);
		put_d("corrected_coverage", corrected_coverage);
		put_i("creation_time", creation_time);
		put_i("reset_distance", reset_distance);
		put_i("max_len_deletion", max_len_deletion);
		{ std::vector<uint64_t> t(tiles.begin(), tiles.end()); f.put("tiles.tiles", 2, t.data(), t.size()); }
		f.put("tiles.abundance", 2, tile_abundance.data(), tile_abundance.size());
		put_vect("insert_lengths", insert_lengths);
		f.put("ref_seq_bias", 3, ref_seq_bias.data(), ref_seq_bias.size());
		put_vectd("insert_lengths_bias", insert_lengths_bias);
		put_vectd("gc_fragment_content_bias", gc_fragment_content_bias);
		for(int b = 0; b < 3; ++b){ f.put("fragment_surroundings_bias." + std::to_string(b), 3, fragment_surroundings_bias[b].data(), fragment_surroundings_bias[b].size()); }
		f.put("dispersion_parameters", 3, dispersion_parameters.data(), 2);
		put_i("tab.num_tiles", num_tiles);
		std::vector<int64_t> desc; std::vector<double> blob; std::vector<uint32_t> par0;
		for(const auto &h : tables){
			int64_t e[16] = {0};
			e[0] = h.par0.size(); e[1] = h.nm;
			for(uint32_t n = 0; n < h.nm; ++n){
				e[2 + n] = h.from[n]; e[6 + n] = h.to[n]; e[10 + n] = blob.size();
				blob.insert(blob.end(), h.dim2[n].begin(), h.dim2[n].end());
			}
			e[14] = par0.size();
			par0.insert(par0.end(), h.par0.begin(), h.par0.end());
			desc.insert(desc.end(), e, e + 16);
		}
		f.put("tab.desc", 4, desc.data(), desc.size());
		f.put("tab.blob", 3, blob.data(), blob.size());
		f.put("tab.par0", 1, par0.data(), par0.size());
	}
};

// std::discrete_distribution's cumulative probabilities (empty when < 2 weights: no draw is consumed)
inline std::vector<double> discrete_cp(const std::vector<double> &weights){
	std::vector<double> prob(weights);
	if(prob.size() < 2){ return {}; }
	const double sum = std::accumulate(prob.begin(), prob.end(), 0.0);
	for(auto &p : prob){ p /= sum; }
	std::vector<double> cp;
	cp.reserve(prob.size());
	std::partial_sum(prob.begin(), prob.end(), std::back_inserter(cp));
	cp[cp.size() - 1] = 1.0;
	return cp;
}
template<class It> inline std::vector<double> discrete_cp(It b, It e){
	std::vector<double> w;
	for(; b != e; ++b){ w.push_back(static_cast<double>(*b)); }
	return discrete_cp(w);
}

// ---------------------------------------------------------------------------------------------------
// Reference genome
// ---------------------------------------------------------------------------------------------------
struct Genome {
	std::vector<std::string> ids;               // full header lines
	std::vector<std::vector<uint8_t>> seqs;     // Dna5 codes: A0 C1 G2 T3 N4

	// Reference::variants_ / num_alleles_, filled by read_variants() (variants.hpp)
	VariantSet variants;
	void read_variants(const std::string &path){
		std::vector<std::string> first_parts;
		for(size_t i = 0; i < ids.size(); ++i){ first_parts.push_back(first_part(i)); }
		variants.read(path, first_parts, seqs);
	}

	// Reference::unmethylated_regions_ / unmethylation_, filled by read_methylation(): unmethylation = the first column (allele 0),
	// unmethylation_alleles[seq][allele] = every column when a sequence's lines carry one value per allele (Reference.cpp:1231-1275: 1 or NumAlleles())
	bool methylation_loaded = false;
	std::vector<std::vector<std::pair<uint32_t, uint32_t>>> unmethylated_regions;
	std::vector<std::vector<double>> unmethylation;
	std::vector<std::vector<std::vector<double>>> unmethylation_alleles;
	uint32_t methylation_alleles_max = 1;

	// Reference::PrepareMethylationFile + ReadMethylation (Reference.cpp:1132-1322) for a single-allele run:
	// extended bedGraph lines "<sequence> <start> <end> <methylation>", grouped by sequence in reference order.
	void read_methylation(const std::string &path){
		TextInput in(path);   // BedFileIn reads gzip-compressed files as well
		if(!in.is_open()){ throw std::runtime_error("Unable to open methylation file " + path); }
		std::istream &f = in.stream();
		std::string line;
		if(!std::getline(f, line)){ throw std::runtime_error("Methylation file is empty: " + path); }
		while((line.empty() || !line.compare(0, 5, "track")) && std::getline(f, line));
		if(f.fail()){ throw std::runtime_error("Methylation file only contains track lines: " + path); }
		std::string cur_seq = line.substr(0, line.find_first_of(" \t"));
		unmethylated_regions.assign(seqs.size(), {});
		unmethylation.assign(seqs.size(), {});
		unmethylation_alleles.assign(seqs.size(), {});
		methylation_alleles_max = 1;
		const uint32_t allowed = variants.loaded() ? variants.num_alleles : 1;   // NumAlleles() when the VCF was loaded first (the reference opens it first, Simulator.cpp:2747-2772)
		bool eof = false;
		for(size_t sid = 0; sid < seqs.size(); ++sid){
			if(eof || first_part(sid) != cur_seq){ continue; }
			auto &regions = unmethylated_regions[sid];
			while(!f.fail()){
				size_t first_space = line.find_first_not_of(" \t", cur_seq.size() + 1);
				size_t second_space = line.find_first_of(" \t", first_space);
				long long v;
				try{ v = std::stoll(line.substr(first_space, second_space)); }
				catch(const std::exception &e){ throw std::runtime_error("Could not convert second field to int for line:\n" + line); }
				if(regions.empty()){ if(v < 0){ throw std::runtime_error("Second field is negative in line:\n" + line); } }
				else if(v < regions.back().second){ throw std::runtime_error("Region is overlapping with previous region in line:\n" + line); }
				if(v >= static_cast<long long>(seqs[sid].size())){ throw std::runtime_error("Second field is larger than sequence length:\n" + line); }
				const uint32_t region_start = v;
				first_space = line.find_first_not_of(" \t", second_space);
				second_space = line.find_first_of(" \t", first_space);
				try{ v = std::stoll(line.substr(first_space, second_space)); }
				catch(const std::exception &e){ throw std::runtime_error("Could not convert third field to int for line:\n" + line); }
				if(v <= region_start){ throw std::runtime_error("Third field is smaller than second field in line:\n" + line); }
				if(v > static_cast<long long>(seqs[sid].size())){ throw std::runtime_error("Third field is larger than sequence length:\n" + line); }
				regions.emplace_back(region_start, static_cast<uint32_t>(v));
				uint32_t allele = 0;
				first_space = line.find_first_not_of(" \t", second_space);
				auto &cols = unmethylation_alleles[sid];
				const uint32_t num_alleles_seq = regions.size() == 1 ? allowed : static_cast<uint32_t>(cols.size());   // first line of a sequence: up to NumAlleles(); later: as in its first line
				while(first_space < line.size()){
					if(allele >= num_alleles_seq){
						if(allele >= allowed){ throw std::runtime_error("More alleles specified than in variant file [" + std::to_string(allowed) + "] in line:\n" + line); }
						throw std::runtime_error("More alleles specified than in last line [" + std::to_string(num_alleles_seq) + "] in line:\n" + line);
					}
					second_space = line.find_first_of(" \t", first_space);
					double d;
					try{ d = std::stod(line.substr(first_space, second_space)); }
					catch(const std::exception &e){ throw std::runtime_error("Could not convert field " + std::to_string(4 + allele) + " to double for line:\n" + line); }
					if(0.0 > d || d > 1.0){ throw std::runtime_error("Field " + std::to_string(4 + allele) + " is not between 0 and 1:\n" + line); }
					if(cols.size() <= allele){ cols.resize(allele + 1); }
					cols[allele].push_back(1.0 - d);
					if(0 == allele){ unmethylation[sid].push_back(1.0 - d); }
					++allele;
					first_space = line.find_first_not_of(" \t", second_space);
				}
				if(regions.size() == 1){
					if(1 != allele && allowed != allele){ throw std::runtime_error(std::to_string(allele) + " alleles specified (must be either 1 or same as in variant file[" + std::to_string(allowed) + "]) in line:\n" + line); }
					cols.resize(allele);
				}
				else if(cols.size() != allele){ throw std::runtime_error(std::to_string(allele) + " alleles specified (must be either identical in all lines of a sequence [" + std::to_string(cols.size()) + "]) in line:\n" + line); }
				methylation_alleles_max = std::max<uint32_t>(methylation_alleles_max, allele);
				while(std::getline(f, line) && line.empty());
				if(!f.fail()){
					const size_t sp = line.find_first_of(" \t");
					if(line.compare(0, sp, cur_seq)){ cur_seq = line.substr(0, sp); break; }
				}
			}
			if(f.fail()){
				if(!f.eof()){ throw std::runtime_error("Could not read methylation file for reference sequence: " + first_part(sid)); }
				eof = true;
			}
		}
		methylation_loaded = true;
	}

	// SimBlock::first_methylation_id_ of the forward block starting at `start_pos` (Simulator.cpp:965-977, 1214-1220)
	int32_t first_methylation_id(size_t sid, uint32_t start_pos) const {
		if(!methylation_loaded || 0 == start_pos){ return 0; }
		const auto &regions = unmethylated_regions[sid];
		int32_t first_meth = 0;   // reverse partner of the previous block: last region starting before start_pos
		while(static_cast<size_t>(first_meth) < regions.size() && regions[first_meth].first < start_pos){ ++first_meth; }
		int32_t id = first_meth - 1;
		if(0 > id || regions.at(id).second <= start_pos){ ++id; }
		return id;
	}

	static uint8_t code(char ch){
		switch(ch){
		case 'A': case 'a': return 0;
		case 'C': case 'c': return 1;
		case 'G': case 'g': return 2;
		case 'T': case 't': case 'U': case 'u': return 3;
		default: return 4;
		}
	}
	// code() over a whole sequence: table lookup, split over the host cores for chromosome-sized inputs
	static void encode(const char *bases, size_t n, uint8_t *out){
		uint8_t table[256];
		for(int ch = 0; ch < 256; ++ch){ table[ch] = code(static_cast<char>(ch)); }
		auto run = [&](size_t lo, size_t hi){ for(size_t k = lo; k < hi; ++k){ out[k] = table[static_cast<unsigned char>(bases[k])]; } };
		const size_t kPerThread = 1u << 20;
		const size_t n_threads = std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), n / kPerThread);
		if(n_threads < 2){ run(0, n); return; }
		std::vector<std::thread> pool;
		for(size_t t = 0; t < n_threads; ++t){ pool.emplace_back(run, n * t / n_threads, n * (t + 1) / n_threads); }
		for(auto &th : pool){ th.join(); }
	}
	std::string first_part(size_t i) const { return ids[i].substr(0, ids[i].find(' ')); }
	uint64_t total_size() const { uint64_t s = 0; for(const auto &q : seqs){ s += q.size(); } return s; }

	void read_fasta(const std::string &path){
		TextInput in(path);   // SeqFileIn reads gzip-compressed FASTA as well (recognised by its magic bytes)
		if(!in.is_open()){ throw std::runtime_error("Could not open " + path + " for reading."); }
		std::istream &f = in.stream();
		ids.clear(); seqs.clear();
		std::string line;
		while(std::getline(f, line)){
			if(!line.empty() && line.back() == '\r'){ line.pop_back(); }
			if(!line.empty() && line[0] == '>'){
				ids.push_back(line.substr(1));
				seqs.emplace_back();
			}
			else if(!seqs.empty()){
				auto &s = seqs.back();
				for(char ch : line){
					if(ch != ' ' && ch != '\t'){ s.push_back(code(ch)); }
				}
			}
		}
		if(in.corrupt()){ throw std::runtime_error("Could not read " + path + ": corrupt or truncated gzip stream."); }
		if(seqs.empty()){ throw std::runtime_error(path + " does not contain any reference sequences."); }
	}

	// Reference::ReplaceN with the same libstdc++ generator/distribution objects the reference uses
	void replace_n(uint64_t seed){
		std::mt19937_64 rgen;
		rgen.seed(seed);
		std::uniform_int_distribution<> rdis(0, 3);
		const uint32_t kMinNToReplaceNWithRepeat = 100;
		for(auto &seq : seqs){
			const uint32_t len = seq.size();
			for(uint32_t start = 0; start < len; ){
				if(seq[start] > 3){
					uint32_t end = start;
					while(++end < len && seq[end] > 3);
					if(end - start < kMinNToReplaceNWithRepeat){
						for(auto pos = start; pos < end; ++pos){ seq[pos] = rdis(rgen); }
					}
					else{
						std::vector<uint8_t> rep;
						if(2 > start){
							if(end + 4 > len){
								for(uint32_t pos = 4; pos--; ){ rep.push_back(rdis(rgen)); }
							}
							else{
								rep.assign(seq.begin() + end, seq.begin() + end + 4);
								for(uint32_t pos = 4; --pos; ){
									if(rep[pos] > 3){ rep[pos] = rdis(rgen); }
								}
							}
						}
						else{
							if(end + 2 > len){
								if(4 > start){
									for(uint32_t pos = 4; pos--; ){ rep.push_back(rdis(rgen)); }
								}
								else{
									rep.assign(seq.begin() + start - 4, seq.begin() + start);
								}
							}
							else{
								rep.insert(rep.end(), seq.begin() + end, seq.begin() + end + 2);
								rep.insert(rep.end(), seq.begin() + start - 2, seq.begin() + start);
								if(rep[1] > 3){ rep[1] = rdis(rgen); }
							}
						}
						for(auto pos = start; pos < end; ++pos){ seq[pos] = rep[(pos - start) % 4]; }
					}
					start = end;
				}
				else{
					++start;
					// eight bases at a time while none of them is an N (code 4: bit 2 set)
					while(start + 8 <= len){
						uint64_t w;
						std::memcpy(&w, seq.data() + start, 8);
						if(w & 0x0404040404040404ull){ break; }
						start += 8;
					}
				}
			}
		}
	}
};

// ---------------------------------------------------------------------------------------------------
// Simulate() prologue arithmetic
// ---------------------------------------------------------------------------------------------------
inline double coverage_prop_lost_from_adapters(const Profile &p){   // Simulator.cpp:90-104
	uint64_t adapter_bases = 0, total_bases = 0;
	for(int seg = 2; seg--; ){
		const auto &rl = p.read_lengths_by_fragment_length[seg];
		const auto &nm = p.non_mapped_read_lengths_by_fragment_length[seg];
		for(uint64_t frag_len = rl.from; frag_len < rl.to(); ++frag_len){
			const auto &row = rl.v[frag_len - rl.from];
			for(uint64_t read_len = row.from; read_len < row.to(); ++read_len){
				const uint64_t cnt = row.v[read_len - row.from];
				total_bases += cnt * read_len;
				if(frag_len < read_len){
					const uint64_t non_mapped = (frag_len >= nm.from && frag_len < nm.to()) ? nm.v[frag_len - nm.from][read_len] : 0;
					adapter_bases += (cnt - non_mapped) * (read_len - frag_len);
					adapter_bases += non_mapped * read_len;
				}
			}
		}
	}
	return static_cast<double>(adapter_bases) / total_bases;
}

inline uint64_t coverage_to_number_pairs(double coverage, uint64_t total_ref_size, double average_read_length, double adapter_part){
	return std::round(coverage * total_ref_size / average_read_length / 2 / (1 - adapter_part));
}

struct Spline {               // InsertLengthSpline, interpolation use only
	std::vector<uint32_t> sample_positions, knots;
	std::vector<double> sampled_values, pars;
	std::array<std::vector<std::vector<double>>, 3> lin_comb;

	bool get_sample_positions(const OffsetVec<uint64_t> &insert_lengths){
		const uint32_t kDist = 20;
		auto at = [&](uint32_t i) -> uint64_t { return insert_lengths.v.at(i - insert_lengths.from); };
		uint32_t first_sample = std::max<size_t>(1, insert_lengths.from);
		while(first_sample < insert_lengths.to() && 0 == at(first_sample)){ ++first_sample; }
		uint32_t num_samples = 0, hit_zero = 0;
		for(uint32_t len = first_sample; len < insert_lengths.to(); len += kDist){
			if(hit_zero){
				if(at(len) >= 10){
					num_samples += (len - hit_zero) / kDist + 1;
					hit_zero = 0;
				}
			}
			else{
				if(at(len) > 0){ ++num_samples; }
				else{ hit_zero = len; }
			}
		}
		if(2 > num_samples){ return false; }
		sample_positions.resize(num_samples);
		sample_positions.at(0) = first_sample;
		uint32_t found_zeros = 0;
		for(uint32_t k = 1; k < num_samples - found_zeros; ++k){
			sample_positions.at(k) = sample_positions.at(k - 1) + kDist;
			while(0 == at(sample_positions.at(k))){
				++found_zeros;
				sample_positions.at(k) += kDist;
			}
		}
		sample_positions.resize(num_samples - found_zeros);
		return true;
	}

	void prepare(){            // DistributeStartingKnots + PrepareInsertLengthSpline (PrepareSplines)
		knots = sample_positions;
		const size_t n = knots.size();
		for(int dim = 3; dim--; ){
			lin_comb[dim].assign(n - 1, std::vector<double>(n, 0.0));
		}
		std::vector<double> l(n, 0.0), mu(n - 1, 0.0), h(n - 1, 0.0);
		std::vector<std::vector<double>> c(n, std::vector<double>(n, 0.0)), z(n, std::vector<double>(n, 0.0));
		std::vector<std::vector<double>> beta(n - 1);
		for(size_t k = 1; k < beta.size(); ++k){ beta[k].assign(n, 0.0); }
		auto &b = lin_comb[0];
		auto &d = lin_comb[2];
		for(size_t i = 0; i < h.size(); ++i){ h[i] = knots[i + 1] - knots[i]; }
		for(size_t k = 1; k < beta.size(); ++k){
			beta[k][k + 1] = 3 / h[k];
			beta[k][k] = -3 / h[k] - 3 / h[k - 1];
			beta[k][k - 1] = 3 / h[k - 1];
		}
		l[0] = 0.0; mu[0] = 0.0;
		for(size_t k = 1; k < beta.size(); ++k){
			l[k] = 2 * (knots[k + 1] - knots[k - 1]) - h[k - 1] * mu[k - 1];
			mu[k] = h[k] / l[k];
			for(size_t ai = 0; ai < n; ++ai){
				z[k][ai] = (beta[k][ai] - h[k - 1] * z[k - 1][ai]) / l[k];
			}
		}
		l[n - 1] = 1.0;
		for(size_t i = h.size(); i--; ){
			for(size_t ai = 0; ai < n; ++ai){
				c[i][ai] = z[i][ai] - mu[i] * c[i + 1][ai];
				b[i][ai] = -h[i] * (c[i + 1][ai] + 2 * c[i][ai]) / 3;
				d[i][ai] = (c[i + 1][ai] - c[i][ai]) / 3 / h[i];
			}
			b[i][i + 1] += 1 / h[i];
			b[i][i] -= 1 / h[i];
		}
		for(size_t i = h.size(); i--; ){
			for(size_t ai = 0; ai < n; ++ai){ lin_comb[1][i][ai] = c[i][ai]; }
		}
	}

	void set_starting_parameters(){
		pars.resize(knots.size() + 1);
		pars[0] = 1.0;
		uint32_t s = 0;
		for(size_t k = 0; k < knots.size(); ++k){
			while(knots[k] > sample_positions.at(s)){ ++s; }
			if(knots[k] == sample_positions[s]){ pars[k + 1] = sampled_values[s]; }
			else{ pars[k + 1] = 0.5 * (sampled_values[s - 1] + sampled_values[s]); }
			if(pars[k + 1] > 0.0){ pars[k + 1] = std::log(pars[k + 1]); }
			else{ pars[k + 1] = std::log(1e-10); }
		}
	}

	void coefficients(double &a, double &b, double &c, double &d, size_t k) const {
		a = pars.at(k + 1);
		b = 0.0; c = 0.0; d = 0.0;
		for(size_t ai = 1; ai < pars.size(); ++ai){
			b += pars[ai] * lin_comb[0][k][ai - 1];
			c += pars[ai] * lin_comb[1][k][ai - 1];
			d += pars[ai] * lin_comb[2][k][ai - 1];
		}
	}

	// InterpolateNormalizationWithSpline
	void interpolate(std::vector<double> &normalization, const OffsetVec<double> &il_bias){
		auto bias_at = [&](uint32_t i) -> double { return il_bias.v.at(i - il_bias.from); };
		sampled_values.resize(sample_positions.size());
		for(size_t s = 0; s < sample_positions.size(); ++s){
			sampled_values[s] = normalization.at(sample_positions[s]) / bias_at(sample_positions[s]);
		}
		prepare();
		set_starting_parameters();
		for(uint32_t len = 1; len < knots.at(0); ++len){ normalization.at(len) = 0.0; }
		size_t k = 0;
		double a = 0, b = 0, c = 0, d = 0;
		for(; k < knots.size() - 1; ++k){
			coefficients(a, b, c, d, k);
			normalization.at(knots[k]) = bias_at(knots[k]) * std::exp(a);
			for(uint32_t len = knots[k] + 1; len < knots[k + 1]; ++len){
				const uint32_t cur_len = len - knots[k];
				normalization.at(len) = bias_at(len) * std::exp(a + b * cur_len + c * cur_len * cur_len + d * cur_len * cur_len * cur_len);
			}
		}
		normalization.at(knots[k]) = bias_at(knots[k]) * std::exp(pars.at(k + 1));
		const uint32_t cur_len = knots[k] - knots[k - 1];
		const double slope = (b + c * cur_len);
		for(uint32_t len = knots[k] + 1; len < normalization.size(); ++len){
			normalization.at(len) = bias_at(len) * std::exp(pars.at(k + 1) + (len - knots[k]) * slope);
		}
	}
};

inline double get_dispersion(double bias, double a, double b){     // BiasCalculationVectors::GetDispersion
	double r = bias / (a + b * bias);
	if(r > bias * 1e10){ r = bias * 1e10; }
	return r;
}

inline uint32_t split_coverage_groups(std::vector<uint32_t> &coverage_groups, const std::vector<double> &ref_seq_bias){
	std::vector<std::pair<double, uint32_t>> sorted;
	for(auto r = ref_seq_bias.size(); r--; ){ sorted.emplace_back(ref_seq_bias[r], r); }
	std::sort(sorted.begin(), sorted.end());
	coverage_groups.assign(ref_seq_bias.size(), 0);
	double group_start = sorted.front().first;
	uint32_t group = 0;
	for(auto &b : sorted){
		if(b.first > 2 * group_start){ group_start = b.first; ++group; }
		coverage_groups.at(b.second) = group;
	}
	return group + 1;
}

struct BiasParam { uint32_t ref_id; uint32_t fragment_length; };

struct Normalization {
	double bias_normalization = 0.0;
	std::vector<uint32_t> coverage_groups;
	uint32_t num_groups = 0;
	std::vector<double> thresholds;     // [group][len][2]
	std::vector<double> binom_p0;       // [group][len]   pow(1-(1-thr0), 2 * alleles): Binomial's first term when every allele is possible
	std::vector<uint64_t> thr_int;      // [group][len]
	std::vector<uint32_t> thr_hi;       // [group][thr_hi_stride] high words of thr_int (rows padded to 16 bytes with 0xffffffff)
	uint32_t thr_hi_stride = 0;
	std::vector<double> binom_pow;      // [group][len][2 * alleles + 1]   pow(1-(1-thr0), N) for N possible strands (runs with variants only)
};

// Smallest raw 64-bit draw x whose canonical value (double(x)*2^-64, clamped) is >= thr; UINT64_MAX with
// `never` semantics is handled by the exact floating-point re-check in the kernel.
inline uint64_t raw_threshold(double thr){
	auto canon = [](uint64_t x){ double r = static_cast<double>(x) * 5.42101086242752217e-20; if(r >= 1.0){ r = std::nextafter(1.0, 0.0); } return r; };
	if(!(canon(UINT64_MAX) >= thr)){ return UINT64_MAX; }   // never reachable (filter lets UINT64_MAX through; exact check rejects)
	uint64_t lo = 0, hi = UINT64_MAX;                       // invariant: canon(hi) >= thr
	while(lo < hi){
		const uint64_t mid = lo + (hi - lo) / 2;
		if(canon(mid) >= thr){ hi = mid; } else{ lo = mid + 1; }
	}
	return lo;
}

// Everything of CalculateBiasNormalization after the per-(ref, length) sums are known.
// sums/max_bias are indexed like `params` (1-thread order: ref id descending, sample length ascending).
inline bool finish_normalization(Normalization &out, const Profile &p, const std::vector<double> &ref_seq_bias, const Spline &spline_in, const std::vector<BiasParam> &params,
                                 const std::vector<double> &sums, const std::vector<double> &max_bias, uint64_t total_reads, uint32_t num_alleles = 1, bool with_variants = false){
	Spline spline = spline_in;
	const uint32_t to = p.insert_lengths.to();
	out.num_groups = split_coverage_groups(out.coverage_groups, ref_seq_bias);
	std::vector<std::vector<std::array<double, 2>>> thr(out.num_groups, std::vector<std::array<double, 2>>(to, {{0.0, 0.0}}));
	std::vector<double> norm_by_len(to, 0.0), tmp_norm(to, 0.0);
	for(size_t i = 0; i < params.size(); ++i){
		tmp_norm.at(params[i].fragment_length) += sums[i];
		auto &m = thr.at(out.coverage_groups.at(params[i].ref_id)).at(params[i].fragment_length)[0];
		if(max_bias[i] > m){ m = max_bias[i]; }
	}
	for(auto len = to; len--; ){ norm_by_len[len] += tmp_norm[len]; }
	spline.interpolate(norm_by_len, p.insert_lengths_bias);
	auto bias_at = [&](uint32_t i) -> double { return p.insert_lengths_bias.v.at(i - p.insert_lengths_bias.from); };
	for(auto &group : thr){
		double max_ratio = 0.0;
		for(size_t s = 0; s < spline.sample_positions.size(); ++s){
			const uint32_t fl = spline.sample_positions[s];
			const double ratio = group.at(fl)[0] / bias_at(fl);
			if(ratio > max_ratio){ max_ratio = ratio; }
		}
		for(size_t s = 1; s < spline.sample_positions.size(); ++s){
			for(uint32_t fl = spline.sample_positions[s - 1] + 1; fl < spline.sample_positions[s]; ++fl){
				group.at(fl)[0] = max_ratio * bias_at(fl);
			}
		}
		for(uint32_t fl = spline.sample_positions.back() + 1; fl < group.size(); ++fl){
			group.at(fl)[0] = max_ratio * bias_at(fl);
		}
	}
	double normalization = 0.0;
	for(auto n : norm_by_len){ normalization += n; }
	const double full = total_reads / (normalization * 2);
	out.thresholds.assign(static_cast<size_t>(out.num_groups) * to * 2, 1.0);
	out.binom_p0.assign(static_cast<size_t>(out.num_groups) * to, 1.0);
	out.thr_int.assign(static_cast<size_t>(out.num_groups) * to, UINT64_MAX);
	out.binom_pow.assign(with_variants ? static_cast<size_t>(out.num_groups) * to * (2 * num_alleles + 1) : 0, 1.0);
	for(uint32_t g = 0; g < out.num_groups; ++g){
		for(uint32_t len = 0; len < to; ++len){
			auto &t = thr[g][len];
			if(0.0 == t[0]){ t[0] = 1.0; t[1] = 1.0; }
			else{
				// CalculateNonZeroThreshold(full, max_bias, NumAlleles())
				double max_mean = full * t[0];
				double max_dispersion = get_dispersion(max_mean, p.dispersion_parameters[0], p.dispersion_parameters[1]) / num_alleles;
				max_mean /= num_alleles;
				t[0] = std::pow(max_dispersion / (max_dispersion + max_mean), max_dispersion);
				t[1] = std::pow(t[0], 2 * num_alleles);
			}
			const size_t i = static_cast<size_t>(g) * to + len;
			out.thresholds[2 * i] = t[0];
			out.thresholds[2 * i + 1] = t[1];
			const double pp = 1 - t[0];                // DrawNumberNonZeroStrands: p = 1 - zero_probability
			out.binom_p0[i] = std::pow(1 - pp, static_cast<uint16_t>(2 * num_alleles));     // Binomial: pow(1-p, N) with N = 2*#alleles (uintAlleleId)
			out.thr_int[i] = raw_threshold(t[1]);
			if(with_variants){   // fewer strands are possible where alleles delete the start base or do not carry the insertion a fragment starts in
				for(uint32_t n = 0; n <= 2 * num_alleles; ++n){ out.binom_pow[i * (2 * num_alleles + 1) + n] = std::pow(1 - pp, static_cast<uint16_t>(n)); }
			}
		}
	}
	out.thr_hi_stride = (to + 128u + 3u) & ~3u;   // + 128 entries that no draw reaches: the scan's four-trip filter reads up to 127 entries past the longest insert
	out.thr_hi.assign(static_cast<size_t>(out.num_groups) * out.thr_hi_stride, 0xffffffffu);
	for(uint32_t g = 0; g < out.num_groups; ++g){
		for(uint32_t len = 0; len < to; ++len){ out.thr_hi[static_cast<size_t>(g) * out.thr_hi_stride + len] = static_cast<uint32_t>(out.thr_int[static_cast<size_t>(g) * to + len] >> 32); }
	}
	out.bias_normalization = full;
	return full != 0.0;
}

}  // namespace rsq
