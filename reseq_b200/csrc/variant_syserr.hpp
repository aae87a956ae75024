// Host half of Simulator::SetSystematicErrorVariantsForward / ...Reverse (reference Simulator.cpp:1011-1147, 769-909).
//
// The systematic errors of a variant's replacement bases are drawn with the (reference base, last base, dominant base) of the FIRST allele
// carrying the variant.  Last base and dominant base follow a per-allele memory (sys_last_var_pos_per_allele_, sys_last_base_per_allele_,
// sys_dom_base_per_allele_: utilities::DominantBaseWithMemory, utilities.hpp:302-351) that is carried from variant to variant over the
// whole strand - a strictly sequential walk, but one that only looks at the reference and the variant list, never at a random draw.  So
// it runs here, once per run, and leaves one context byte per replacement base and strand:  base | last base << 2 | dominant base << 5.
// The draws themselves (which need the error rates around the variant and the master stream) are draw_variant_errors_block() in
// variant_core.cuh, on the device.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <limits>
#include <thread>
#include <vector>

#include "variants.hpp"

namespace rsq {

// utilities::DominantBase + DominantBaseWithMemory, statement by statement (the sums in seq_content_ are NOT always consistent with memory_:
// Clear() keeps dom_base_, Set() adds to whatever seq_content_ holds - both are reproduced by keeping the same state)
struct DomBaseMemory {
	uint8_t memory[8]; uint32_t len = 0;      // memory_ (at most kLastX + 2 = 7 entries)
	uint16_t content[5] = {0, 0, 0, 0, 0};    // DominantBase::seq_content_
	uint8_t dom = 0;                          // DominantBase::dom_base_
	static constexpr uint32_t kLastX = 5;

	void clear(){ for(auto &c : content){ c = 0; } len = 0; }
	uint8_t get() const { return dom; }
	void find_dominant(uint32_t cur_pos){   // DominantBase::FindDominant(memory_, cur_pos)
		uint32_t max_content = 0;
		for(uint32_t n = 4; n--; ){ if(content[n] > max_content){ max_content = content[n]; } }
		if(0 == max_content){
			if(len <= cur_pos || 4 == memory[cur_pos]){ dom = 0; }
			else{ dom = memory[cur_pos] & 3u; }
		}
		else{
			uint32_t pos = cur_pos;
			while(pos > 0 && max_content != content[memory[--pos]]){}
			dom = memory[pos] & 3u;
		}
	}
	// Set(seq, cur_pos): memory_ = the last min(kLastX, cur_pos) + 1 bases of seq ending at cur_pos; at(p) = seq[p]
	template<class At> void set(At at, uint32_t cur_pos){
		len = (kLastX < cur_pos ? kLastX : cur_pos) + 1;
		for(uint32_t mem_pos = len; mem_pos--; ){ memory[mem_pos] = at(cur_pos + mem_pos + 1 - len); }
		const uint32_t cp = len - 1;   // DominantBase::Set(memory_, cp)
		for(uint32_t pos = (kLastX < cp ? cp - kLastX : 0); pos < cp; ++pos){ ++content[memory[pos]]; }
		find_dominant(cp);
	}
	void update(uint8_t base){
		if(len > kLastX + 1){ for(uint32_t i = 1; i < len; ++i){ memory[i - 1] = memory[i]; } --len; }
		memory[len++] = base;
		if(1 < len){   // DominantBase::Update(memory_[len - 2], memory_, len - 2)
			const uint32_t last_pos = len - 2;
			++content[memory[last_pos]];
			if(kLastX <= last_pos){ --content[memory[last_pos - kLastX]]; }
			find_dominant(last_pos + 1);
		}
		else{   // DominantBase::Set(memory_, 0)
			find_dominant(0);
		}
	}
};

struct VariantSysContext {
	// one byte per replacement base, indexed like FlatVariants::bases; reverse strand: in the order the bases are drawn (last base of var_seq first, complemented)
	std::vector<uint8_t> fwd, rev;
};

inline uint8_t pack_var_ctx(uint8_t base, uint8_t last, uint8_t dom){ return static_cast<uint8_t>((base & 3u) | ((last & 7u) << 2) | ((dom & 3u) << 5)); }

// One sequence, one strand.  seq: Dna5 codes after ReplaceN; out: context bytes at bases_off (of this sequence's variants) relative to `out_base`.
inline void variant_sys_context_strand(const std::vector<uint8_t> &seq, const std::vector<Variant> &vars, uint32_t num_alleles, bool reverse,
                                       const std::vector<uint32_t> &bases_off /* per variant of this sequence, absolute */, uint8_t *out){
	const uint32_t L = seq.size();
	const uint32_t kLastX = DomBaseMemory::kLastX;
	std::vector<uint32_t> last_var_pos(num_alleles, std::numeric_limits<uint32_t>::max());   // ResetSystematicErrorCounters(ref)
	std::vector<uint8_t> last_base_of(num_alleles, 4);
	std::vector<DomBaseMemory> dom(num_alleles);
	auto comp = [](uint8_t b) -> uint8_t { return static_cast<uint8_t>(3u - (b & 3u)); };   // Complement::Dna (a Dna5 N converts to A first; none left after ReplaceN)
	auto fwd_at = [&](uint32_t p) -> uint8_t { return seq[p]; };
	auto rev_at = [&](uint32_t p) -> uint8_t { const uint8_t b = seq[L - 1 - p]; return b > 3 ? b : static_cast<uint8_t>(3u - b); };   // ConstDna5StringReverseComplement
	auto first_allele = [&](const Variant &v) -> uint32_t { for(uint32_t a = 0; a < VariantSet::kMaxAlleles; ++a){ if(v.in_allele(a)){ return a; } } return 0; };
	if(!reverse){
		for(size_t var_id = 0; var_id < vars.size(); ++var_id){
			const Variant &var = vars[var_id];
			const uint32_t chosen = first_allele(var), len = var.var_seq.size();
			uint8_t last_base;
			if(last_var_pos[chosen] < L && last_var_pos[chosen] + 1 == var.position){ last_base = last_base_of[chosen]; }
			else{ last_base = var.position ? seq[var.position - 1] : 4; }
			if(last_var_pos[chosen] < L && last_var_pos[chosen] + kLastX >= var.position){
				for(uint32_t pos = last_var_pos[chosen] + 1; pos < var.position; ++pos){ dom[chosen].update(seq[pos]); }
			}
			else{
				dom[chosen].clear();
				if(var.position){ dom[chosen].set(fwd_at, var.position - 1); }
			}
			for(uint32_t vpos = 0; vpos < len; ++vpos){
				const uint8_t base = var.var_seq[vpos];
				dom[chosen].update(base);
				out[bases_off[var_id] + vpos] = pack_var_ctx(base, last_base, dom[chosen].get());
				last_base = base;
			}
			uint32_t ref_allele = num_alleles;
			if(last_var_pos[chosen] >= L || static_cast<uint64_t>(last_var_pos[chosen]) + kLastX < static_cast<uint64_t>(var.position) + len){ ref_allele = chosen; }
			for(uint32_t allele = 0; allele < num_alleles; ++allele){
				if(!var.in_allele(allele)){ continue; }
				if(allele != chosen){
					if(last_var_pos[allele] < L && static_cast<uint64_t>(last_var_pos[allele]) + kLastX >= static_cast<uint64_t>(var.position) + len){
						for(uint32_t pos = last_var_pos[allele] + 1; pos < var.position; ++pos){ dom[allele].update(seq[pos]); }
						for(uint32_t vpos = 0; vpos < len; ++vpos){ dom[allele].update(var.var_seq[vpos]); }
					}
					else if(ref_allele < num_alleles){ dom[allele] = dom[ref_allele]; }
					else{
						ref_allele = allele;
						dom[allele].clear();
						if(var.position){ dom[allele].set(fwd_at, var.position - 1); }
						for(uint32_t vpos = 0; vpos < len; ++vpos){ dom[allele].update(var.var_seq[vpos]); }
					}
				}
				last_var_pos[allele] = var.position;
				last_base_of[allele] = last_base;
			}
		}
	}
	else{
		for(size_t var_id = vars.size(); var_id--; ){
			const Variant &var = vars[var_id];
			const uint32_t chosen = first_allele(var), len = var.var_seq.size();
			const uint32_t rev_pos = L - var.position - 1;
			uint8_t last_base;
			if(last_var_pos[chosen] < L && last_var_pos[chosen] == var.position + 1){ last_base = last_base_of[chosen]; }
			else{ last_base = (var.position + 1 < L) ? comp(seq[var.position + 1]) : 4; }
			if(last_var_pos[chosen] < L && last_var_pos[chosen] <= var.position + kLastX){
				for(uint32_t pos = last_var_pos[chosen] - 1; pos > var.position; --pos){ dom[chosen].update(comp(seq[pos])); }
			}
			else{
				dom[chosen].clear();
				if(rev_pos){ dom[chosen].set(rev_at, rev_pos - 1); }
			}
			uint32_t k = 0;
			for(uint32_t vpos = len; vpos--; ){
				const uint8_t base = comp(var.var_seq[vpos]);
				dom[chosen].update(base);
				out[bases_off[var_id] + k++] = pack_var_ctx(base, last_base, dom[chosen].get());
				last_base = base;
			}
			uint32_t ref_allele = num_alleles;
			if(last_var_pos[chosen] >= L || static_cast<uint64_t>(last_var_pos[chosen]) + len > static_cast<uint64_t>(var.position) + kLastX){ ref_allele = chosen; }
			for(uint32_t allele = 0; allele < num_alleles; ++allele){
				if(!var.in_allele(allele)){ continue; }
				if(allele != chosen){
					if(last_var_pos[allele] < L && static_cast<uint64_t>(last_var_pos[allele]) + len <= static_cast<uint64_t>(var.position) + kLastX){
						for(uint32_t pos = last_var_pos[allele] - 1; pos > var.position; --pos){ dom[allele].update(comp(seq[pos])); }
						for(uint32_t vpos = len; vpos--; ){ dom[allele].update(comp(var.var_seq[vpos])); }
					}
					else if(ref_allele < num_alleles){ dom[allele] = dom[ref_allele]; }
					else{
						ref_allele = allele;
						dom[allele].clear();
						if(rev_pos){ dom[allele].set(rev_at, rev_pos - 1); }
						for(uint32_t vpos = len; vpos--; ){ dom[allele].update(comp(var.var_seq[vpos])); }
					}
				}
				last_var_pos[allele] = var.position;
				last_base_of[allele] = last_base;
			}
		}
	}
}

// All sequences, both strands (independent of each other: one host thread per (sequence, strand) up to the core count).
inline VariantSysContext variant_sys_context(const std::vector<std::vector<uint8_t>> &seqs, const VariantSet &vs, const FlatVariants &flat){
	VariantSysContext out;
	out.fwd.assign(flat.bases.size() + 1, 0); out.rev.assign(flat.bases.size() + 1, 0);
	struct Job { uint32_t seq; bool reverse; };
	std::vector<Job> jobs;
	for(uint32_t s = 0; s < seqs.size(); ++s){ if(!vs.variants[s].empty()){ jobs.push_back({s, false}); jobs.push_back({s, true}); } }
	auto run = [&](const Job &j){
		std::vector<uint32_t> off(flat.bases_off.begin() + flat.seq_first[j.seq], flat.bases_off.begin() + flat.seq_first[j.seq + 1]);
		variant_sys_context_strand(seqs[j.seq], vs.variants[j.seq], vs.num_alleles, j.reverse, off, j.reverse ? out.rev.data() : out.fwd.data());
	};
	const size_t n_threads = std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), jobs.size());
	if(n_threads < 2){ for(const auto &j : jobs){ run(j); } return out; }
	std::vector<std::thread> pool;
	std::atomic<size_t> next{0};
	for(size_t t = 0; t < n_threads; ++t){ pool.emplace_back([&]{ for(size_t i = next++; i < jobs.size(); i = next++){ run(jobs[i]); } }); }
	for(auto &th : pool){ th.join(); }
	return out;
}

}  // namespace rsq
