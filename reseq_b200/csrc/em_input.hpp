// Input of seqToIllumina (Simulator::ApplyErrorsAndQualityToFastaInput, Simulator.cpp:2403-2512): FASTA records whose header carries
// "<id> <1|2>;<fragment length>;<dominant systematic errors>;<systematic error rates>".  At BASELINE config C3's size the file has
// 10 M records of ~1 kB, so the host side works on the whole text at once and in parallel: record starts are found by all cores,
// every record is measured and validated (pass 1), offsets are prefix sums, and the arrays the kernels read are filled in place
// (pass 2).  The checks and their order per record are the reference's; of several bad records the first one in the file is reported.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include "host_profile.hpp"
#include "spec_core.cuh"
#include "text_io.hpp"

namespace rsq {

struct EmInput {
	std::vector<EmRecord> recs;
	std::vector<uint8_t> seq, dom, rate;   // per base: code (DnaString semantics: non-ACGTU -> A), dominant error, error rate
	std::string ids;                       // record ids behind each other
	uint32_t max_len = 0, max_id_len = 0;
};

namespace em_detail {
template<class F> inline void parallel_ranges(size_t n, size_t min_per_thread, F &&f){   // f(thread index, lo, hi)
	size_t t = std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), std::max<size_t>(1, n / std::max<size_t>(1, min_per_thread)));
	if(t < 2){ f(0, 0, n); return; }
	std::vector<std::thread> pool;
	for(size_t i = 0; i < t; ++i){ pool.emplace_back([&, i]{ f(i, n * i / t, n * (i + 1) / t); }); }
	for(auto &th : pool){ th.join(); }
}
struct Measured { size_t head_lo, head_hi, body_lo, body_hi; uint32_t len, id_len, seg, fragment_length; };
}  // namespace em_detail

inline EmInput read_em_input(const std::string &path){
	using namespace em_detail;
	TextInput in(path);
	if(!in.is_open()){ throw std::runtime_error("Could not open '" + path + "' for reading."); }
	std::string text;
	{
		std::istream &f = in.stream();
		std::vector<char> chunk(64u << 20);
		while(f){ f.read(chunk.data(), static_cast<std::streamsize>(chunk.size())); text.append(chunk.data(), static_cast<size_t>(f.gcount())); }
		if(in.corrupt()){ throw std::runtime_error("Could not read " + path + ": corrupt or truncated gzip stream."); }
	}
	// record starts: '>' at the beginning of a line
	std::vector<std::vector<size_t>> found(std::max(1u, std::thread::hardware_concurrency()) + 1);
	parallel_ranges(text.size(), 4u << 20, [&](size_t t, size_t lo, size_t hi){
		auto &v = found[t];
		for(size_t i = lo; i < hi; ++i){
			const char *p = static_cast<const char *>(std::memchr(text.data() + i, '>', hi - i));
			if(!p){ break; }
			i = static_cast<size_t>(p - text.data());
			if(i == 0 || text[i - 1] == '\n'){ v.push_back(i); }
		}
	});
	std::vector<size_t> starts;
	for(const auto &v : found){ starts.insert(starts.end(), v.begin(), v.end()); }
	if(starts.empty()){ throw std::runtime_error(path + " does not contain any sequences."); }
	const size_t n = starts.size();
	std::vector<Measured> ms(n);
	struct Failure { size_t rec = static_cast<size_t>(-1); std::string what; };
	std::vector<Failure> failures(found.size());
	auto id_at = [&](const Measured &m, size_t k){ return text[m.head_lo + k]; };
	// pass 1: measure and validate
	parallel_ranges(n, 4096, [&](size_t t, size_t lo, size_t hi){
		for(size_t r = lo; r < hi; ++r){
			Measured &m = ms[r];
			const size_t rec_end = r + 1 < n ? starts[r + 1] : text.size();
			const char *eol = static_cast<const char *>(std::memchr(text.data() + starts[r], '\n', rec_end - starts[r]));
			m.head_lo = starts[r] + 1;
			m.head_hi = eol ? static_cast<size_t>(eol - text.data()) : rec_end;
			m.body_lo = eol ? m.head_hi + 1 : rec_end; m.body_hi = rec_end;
			if(m.head_hi > m.head_lo && text[m.head_hi - 1] == '\r'){ --m.head_hi; }
			// sequence characters: every line without its trailing '\r', blanks and tabs dropped
			uint32_t L = 0;
			for(size_t i = m.body_lo; i < m.body_hi; ++i){
				const char ch = text[i];
				if(ch == '\n' || ch == ' ' || ch == '\t'){ continue; }
				if(ch == '\r' && (i + 1 == m.body_hi || text[i + 1] == '\n')){ continue; }
				++L;
			}
			m.len = L;
			const size_t id_size = m.head_hi - m.head_lo;
			auto id_text_of = [&]{ return text.substr(m.head_lo, id_size); };   // only needed for messages
			auto fail = [&](const std::string &what){ if(failures[t].rec == static_cast<size_t>(-1)){ failures[t].rec = r; failures[t].what = what; } };
			if(id_size <= 2 * static_cast<size_t>(L) + 2){ fail("Read description is too short to contain systematic error information and a sequence id: " + id_text_of()); break; }
			size_t end_pos = id_size - 2 * static_cast<size_t>(L) - 3;
			if(';' != id_at(m, end_pos + 1) || ';' != id_at(m, end_pos + 2 + L)){ fail("The two systematic error entries are not separated by a semicolon from themselves or the rest of the ReSeq information: " + id_text_of()); break; }
			while(end_pos && ' ' != id_at(m, end_pos)){ --end_pos; }
			if(0 == end_pos){ fail("No sequence id found that is separated by a space from the ReSeq information: " + id_text_of()); break; }
			m.id_len = static_cast<uint32_t>(end_pos);
			if('1' == id_at(m, end_pos + 1)){ m.seg = 0; }
			else if('2' == id_at(m, end_pos + 1)){ m.seg = 1; }
			else{ fail(std::string("Template segment is ") + id_at(m, end_pos + 1) + " not 1 or 2: " + id_text_of()); break; }
			if(';' != id_at(m, end_pos + 2)){ fail("The template segment and fragment length are not separated by a semicolon: " + id_text_of()); break; }
			const std::string fl = text.substr(m.head_lo + end_pos + 3, id_size - 2 * static_cast<size_t>(L) - 2 - (end_pos + 3));
			size_t used = 0; int v = 0;
			try{ v = std::stoi(fl, &used); }catch(...){ used = 0; }
			if(used < fl.size() || fl.empty()){ fail("Fragment length '" + fl + "' is not a pure integer: " + id_text_of()); break; }
			m.fragment_length = static_cast<uint32_t>(v);
		}
	});
	{
		const Failure *first = nullptr;
		for(const auto &f : failures){ if(f.rec != static_cast<size_t>(-1) && (!first || f.rec < first->rec)){ first = &f; } }
		if(first){ throw std::runtime_error(first->what); }
	}
	// offsets
	EmInput out;
	out.recs.resize(n);
	uint64_t seq_total = 0, id_total = 0;
	for(size_t r = 0; r < n; ++r){
		EmRecord &rec = out.recs[r];
		rec = EmRecord{};
		rec.seq_off = seq_total; rec.len = ms[r].len; rec.seg = ms[r].seg; rec.fragment_length = ms[r].fragment_length;
		if(id_total + ms[r].id_len > 0xffffffffull){ throw std::runtime_error("record ids of " + path + " exceed 4 GiB"); }
		rec.id_off = static_cast<uint32_t>(id_total); rec.id_len = ms[r].id_len;
		seq_total += ms[r].len; id_total += ms[r].id_len;
		out.max_len = std::max(out.max_len, ms[r].len); out.max_id_len = std::max(out.max_id_len, ms[r].id_len);
	}
	out.seq.resize(seq_total); out.dom.resize(seq_total); out.rate.resize(seq_total); out.ids.resize(id_total);
	// pass 2: fill
	parallel_ranges(n, 4096, [&](size_t, size_t lo, size_t hi){
		for(size_t r = lo; r < hi; ++r){
			const Measured &m = ms[r]; const EmRecord &rec = out.recs[r];
			const size_t id_size = m.head_hi - m.head_lo, L = m.len;
			uint8_t *sq = out.seq.data() + rec.seq_off;
			size_t k = 0;
			for(size_t i = m.body_lo; i < m.body_hi; ++i){
				const char ch = text[i];
				if(ch == '\n' || ch == ' ' || ch == '\t'){ continue; }
				if(ch == '\r' && (i + 1 == m.body_hi || text[i + 1] == '\n')){ continue; }
				sq[k++] = Genome::code(ch) & 3;
			}
			const size_t dom_at = id_size - 2 * L - 3 + 2, rate_at = id_size - L;
			for(size_t pos = 0; pos < L; ++pos){
				out.dom[rec.seq_off + pos] = Genome::code(id_at(m, dom_at + pos));
				uint8_t rate = static_cast<uint8_t>(id_at(m, rate_at + pos) - 33);
				if(86 < rate){ rate += rate - 86; }
				out.rate[rec.seq_off + pos] = rate;
			}
			std::memcpy(&out.ids[rec.id_off], text.data() + m.head_lo, rec.id_len);
		}
	});
	return out;
}

}  // namespace rsq
