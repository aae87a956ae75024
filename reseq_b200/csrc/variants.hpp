// Variant (VCF) loading for the simulation path: the host half of SURVEY §8 row a6.
//
// Mirrors Reference::PrepareVariantFile, ReadFirstVariants, ReadVariants and InsertVariant
// (reference: reseq/Reference.cpp:96-113, 126-426, 1005-1078; reseq/Reference.h:24-63, 115-139) together with the
// parts of SeqAn's VCF reader they rely on (seqan/vcf_io/read_vcf.h:67-99 contig names from "##contig=<ID=...>",
// 104-160 header, 170-225 records: tab-split columns, POS-1, sample columns after FORMAT).
//
// What comes out is what Simulator consumes: per reference sequence a position-sorted list of single-position
// variants {position, replacement bases (empty = deletion, >1 = insertion after the base), allele bit set}, flattened
// into arrays a kernel can index (VariantSet::flatten). The reference pages sequences in while it simulates
// (Simulator.cpp:938, 1278) and frees them behind itself; HBM holds the whole set, so the file is read once. A file
// the reference would reject at any point of its run is rejected here up front with the same diagnostics.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "text_io.hpp"

namespace rsq {

struct Variant {                       // Reference::Variant (Reference.h:24-63)
	uint32_t position;
	std::vector<uint8_t> var_seq;      // codes 0..3
	std::array<uint64_t, 2> allele;    // bit a of word a/64: variant present in allele a (kMaxAlleles = 128)
	bool in_allele(uint32_t a) const { return (allele[a / 64] >> (a % 64)) & 1; }
};

struct FlatVariants {                  // device layout: sequences back to back, variants of a sequence sorted by position
	std::vector<uint32_t> seq_first;   // [n_seqs + 1] first variant id of each sequence
	std::vector<uint32_t> position;    // [n_var]
	std::vector<uint32_t> bases_off;   // [n_var + 1] into bases
	std::vector<uint8_t> bases;        // replacement bases, codes 0..3
	std::vector<uint64_t> allele_lo, allele_hi;   // [n_var]
};

// The materialised sequence of one allele of one reference sequence and the coordinate map the variant-aware scan works with
// (tests/test_variant_invariant_cpu.py: GC, surroundings, end position and fragment ends of an allele are plain lookups here).
struct AlleleSequence {
	std::vector<uint8_t> bases;    // the reference with the allele's variants applied
	std::vector<uint32_t> off;     // [L + 1] index in `bases` of the first base standing for reference position p (off[L] = size)
	// smallest reference position p with off[p] >= index: the reference's cur_end_position for a fragment ending at `index` (exclusive)
	uint32_t ref_position(uint32_t index) const { return std::lower_bound(off.begin(), off.end(), index) - off.begin(); }
};

class VariantSet {
public:
	static constexpr uint32_t kMaxAlleles = 128;          // Reference::Variant::kMaxAlleles
	static constexpr uint32_t kMaxErrorsShownPerFile = 50; // Reference.h:71

	uint32_t num_alleles = 1;          // Reference::num_alleles_
	uint32_t num_populations = 0;      // genotype columns of the first record
	std::vector<std::vector<Variant>> variants;             // Reference::variants_
	std::vector<std::vector<uint32_t>> variant_positions;   // Reference::variant_positions_ (positions_only)
	uint64_t ref_checks_deferred = 0;  // REF bases that lie on N of the unprocessed reference (compared after ReplaceN)
	struct DeferredCheck { uint32_t seq, position; uint8_t base; };
	std::vector<DeferredCheck> deferred;   // those bases: the reference reads the VCF after ReplaceN (Simulator.cpp:2690, 2750), so the engine re-checks them then
	// Throws like read() does when a REF base that stood on an N of the unprocessed reference differs from what ReplaceN put there.
	void check_deferred(const std::vector<std::vector<uint8_t>> &replaced) const {
		for(const auto &d : deferred){
			if(replaced.at(d.seq).at(d.position) != d.base){
				throw std::runtime_error(std::string("The specified reference in vcf file '") + "ACGTN"[d.base] + "' is not identical with the specified reference sequence " + std::to_string(d.seq) +
				                         " at position " + std::to_string(d.position) + ": '" + "ACGTN"[replaced[d.seq][d.position]] + "' (a base ReplaceN filled in for an N).");
			}
		}
	}
	std::string diagnostics;           // what the reference prints through printErr, one line each

	bool loaded() const { return !variants.empty() || !variant_positions.empty(); }

	// seq_ids: Reference::ReferenceIdFirstPart of every sequence; seqs: Dna5 codes (A0 C1 G2 T3 N4).
	// Throws std::runtime_error carrying the diagnostics when the reference would return false.
	void read(const std::string &path, const std::vector<std::string> &seq_ids, const std::vector<std::vector<uint8_t>> &seqs, bool positions_only = false){
		variants.clear(); variant_positions.clear(); diagnostics.clear(); ref_checks_deferred = 0; deferred.clear(); num_alleles = 1; num_populations = 0;
		TextInput in(path);   // VcfFileIn opens gzip-compressed files as well
		if(!in.is_open()){ fail("Could not open vcf file '" + path + "'."); }
		std::istream &f = in.stream();

		// --- readHeader + CheckVcf (PrepareVariantFile) ---
		std::vector<std::string> contigs;
		size_t n_samples = 0;
		std::string line;
		bool have_line = false;
		while(std::getline(f, line)){
			strip_cr(line);
			if(line.empty() || line[0] != '#'){ have_line = true; break; }
			if(line.size() > 1 && line[1] == '#'){
				const size_t eq = line.find('=');
				if(eq == std::string::npos){ fail("Could not prepare vcf file '" + path + "' for record readin: header line without '='"); }
				if(line.compare(2, eq - 2, "contig") == 0){ contigs.push_back(contig_id(line.substr(eq + 1), path)); }
			}
			else{
				if(line.compare(1, 5, "CHROM") != 0){ fail("Could not prepare vcf file '" + path + "' for record readin: Invalid line with samples."); }
				const auto fields = split_tabs(line);
				if(fields.size() < 8){ fail("Could not prepare vcf file '" + path + "' for record readin: Not enough fields."); }
				for(size_t i = 8; i < fields.size(); ++i){
					if(i == 8 && fields[i] == "FORMAT"){ continue; }
					++n_samples;
				}
			}
		}
		uint32_t errors = 0;
		if(contigs.size() != seq_ids.size()){
			err("Number of contigs does not match between reference(" + std::to_string(seq_ids.size()) + ") and variant(" + std::to_string(contigs.size()) + ") file.");
			++errors;
		}
		for(size_t con = 0; con < std::min(contigs.size(), seq_ids.size()); ++con){
			if(errors < 20 && contigs[con] != seq_ids[con]){
				err("Contigs at position " + std::to_string(con) + " do not match between reference(" + seq_ids[con] + ") and variant(" + contigs[con] + ") file.");
				++errors;
			}
		}
		if(errors){ fail_collected(); }
		if(!have_line){ fail("Vcf file '" + path + "' has no records."); }

		// --- ReadFirstVcfRecord + allele count (ReadFirstVariants) ---
		Record rec;
		if(!parse_record(line, contigs, n_samples, rec)){ fail("Could not read first vcf record: " + parse_error_); }
		if(!positions_only){
			num_alleles = 0;
			for(const auto &genotype : rec.genotypes){
				for(size_t pos = 0; pos < genotype.size() && genotype[pos] != ':'; ++pos){
					if(genotype[pos] == '|' || genotype[pos] == '/'){ ++num_alleles; }
				}
				++num_alleles;
			}
			num_populations = rec.genotypes.size();
			if(num_alleles > kMaxAlleles){
				fail("Currently only " + std::to_string(kMaxAlleles) + " alleles are supported, but file has " + std::to_string(num_alleles) + ".");
			}
			variants.assign(seqs.size(), {});
		}
		else{
			variant_positions.assign(seqs.size(), {});
		}

		// --- ReadVariants over the whole file ---
		const uint32_t n_seqs = seqs.size();
		uint32_t start_pos = 0, end_pos = 0;
		uint32_t old_ref_id = std::numeric_limits<uint32_t>::max();
		uint32_t read_for_num_sequences = 0;   // read_variation_for_num_sequences_
		std::vector<uint16_t> allele(num_alleles);
		std::vector<std::array<uint64_t, 2>> gt_has_var;
		std::vector<uint32_t> alt_start_pos;
		bool stop = false;
		auto count_error = [&](const std::string &msg){
			err(msg);
			if(++errors >= kMaxErrorsShownPerFile){
				err("Maximum number of errors reached. Additional errors are not shown for this file.");
				stop = true;
			}
		};
		while(!stop){
			const std::string where = "Variant starting in reference sequence " + std::to_string(rec.rid) + " at position ";
			if(rec.rid >= n_seqs){
				count_error(where + std::to_string(start_pos) + " does not belong to an existing reference sequence.");
			}
			else if(static_cast<uint32_t>(rec.begin_pos) >= seqs[rec.rid].size()){
				count_error(where + std::to_string(start_pos) + " starts after the end of the reference sequence.");
			}
			else{
				start_pos = rec.begin_pos;
				if(old_ref_id == rec.rid){
					if(start_pos < end_pos){ count_error(where + std::to_string(start_pos) + " overlaps with a previous variant."); }
				}
				else{ old_ref_id = rec.rid; }
				if(stop){ break; }
				end_pos = start_pos + rec.ref.size();
				std::vector<uint8_t> vcf_ref(rec.ref.size());
				bool ref_has_n = false;
				for(size_t k = 0; k < rec.ref.size(); ++k){ vcf_ref[k] = dna5(rec.ref[k]); ref_has_n |= vcf_ref[k] > 3; }
				if(ref_has_n){
					count_error(where + std::to_string(start_pos) + " has an reference column containing ambiguous bases (e.g. N). Please change or remove them, but make sure the reference file stays consistent with this column.");
				}
				else{
					const auto &s = seqs[rec.rid];
					bool same = end_pos <= s.size();
					for(uint32_t k = 0; same && k < vcf_ref.size(); ++k){
						if(s[start_pos + k] > 3){ ++ref_checks_deferred; deferred.push_back({rec.rid, start_pos + k, vcf_ref[k]}); }   // ReplaceN runs first in the reference (Simulator.cpp:2690, 2750)
						else if(s[start_pos + k] != vcf_ref[k]){ same = false; }
					}
					if(!same){
						std::string have;
						for(uint32_t k = start_pos; k < end_pos && k < s.size(); ++k){ have += "ACGTN"[s[k]]; }
						count_error("The specified reference in vcf file '" + upper(vcf_ref) + "' is not identical with the specified reference sequence " + std::to_string(rec.rid) + " at position " + std::to_string(start_pos) + ": '" + have + "'.");
					}
				}
				if(stop){ break; }

				if(positions_only){
					for(uint32_t pos = start_pos; pos < end_pos; ++pos){ variant_positions[rec.rid].push_back(pos); }
				}
				else{
					// genotypes: which alternative every allele carries
					bool tmp_success = true;
					uint32_t cur_allele = 0;
					for(const auto &genotype : rec.genotypes){
						if(cur_allele >= num_alleles){
							tmp_success = false;
							count_error("Found to many alleles in genotype definition '" + join_genotypes(rec) + "'");
							break;
						}
						uint16_t chosen_var = 0;
						bool overflow = false;
						for(size_t pos = 0; pos < genotype.size() && genotype[pos] != ':' && !stop; ++pos){
							const char ch = genotype[pos];
							if(ch == '|' || ch == '/'){
								if(cur_allele >= num_alleles){ overflow = true; break; }   // std::vector::at throws in the reference
								allele[cur_allele++] = chosen_var;
								chosen_var = 0;
							}
							else if(ch >= '0' && ch <= '9'){
								chosen_var = static_cast<uint16_t>(chosen_var * 10 + (ch - 48));
							}
							else{
								tmp_success = false;
								count_error(std::string("Unallowed character '") + ch + "' in genotype definition '" + genotype + "'");
							}
						}
						if(stop){ break; }
						if(overflow || cur_allele >= num_alleles){
							count_error("Could not read vcf record: more alleles in genotype definition '" + join_genotypes(rec) + "' than in the first record");
							stop = true;
							break;
						}
						allele[cur_allele++] = tmp_success ? chosen_var : 0;
					}
					if(stop){ break; }
					if(cur_allele < num_alleles){
						tmp_success = false;
						count_error("Could not find enough alleles in genotype definition '" + join_genotypes(rec) + "'");
						if(stop){ break; }
					}

					if(tmp_success){
						gt_has_var.clear();
						alt_start_pos.clear();
						alt_start_pos.push_back(0);
						uint16_t chosen_var = 1;   // 0 is the reference sequence
						auto carriers = [&](bool last){
							std::array<uint64_t, 2> bits{{0, 0}};
							for(uint32_t a = num_alleles; a--; ){   // allele 0 in the rightmost bit
								bits[a / 64] <<= 1;
								if(allele[a] == chosen_var){ ++bits[a / 64]; }
								else if(last && allele[a] > chosen_var){
									if(++errors <= kMaxErrorsShownPerFile){
										err("Variant number " + std::to_string(allele[a]) + " does not exist for sequence id " + std::to_string(rec.rid) + " and position " + std::to_string(rec.begin_pos));
									}
									if(errors >= kMaxErrorsShownPerFile){ err("Maximum number of errors reached. Additional errors are not shown for this file."); }
								}
							}
							gt_has_var.push_back(bits);
							++chosen_var;
						};
						uint32_t pos;
						for(pos = 0; pos < rec.alt.size(); ++pos){
							if(rec.alt[pos] == ','){
								alt_start_pos.push_back(pos + 1);
								carriers(false);
							}
						}
						alt_start_pos.push_back(pos + 1);   // one after the end, in line with "one after the ','"
						carriers(true);

						// one entry per reference position of the record
						for(pos = 0; pos < vcf_ref.size(); ++pos){
							for(uint32_t n_alt = 0; n_alt < gt_has_var.size(); ++n_alt){
								if(!(gt_has_var[n_alt][0] | gt_has_var[n_alt][1])){ continue; }
								const uint32_t alt_len = alt_start_pos[n_alt + 1] - 1 - alt_start_pos[n_alt];
								std::vector<uint8_t> inserted;
								if(pos + 1 == vcf_ref.size() && pos + 1 < alt_len){   // insertion
									for(uint32_t k = alt_start_pos[n_alt] + pos; k < alt_start_pos[n_alt + 1] - 1; ++k){ inserted.push_back(dna5(rec.alt[k])); }
								}
								else if(pos < alt_len){   // base mutation (compared as characters: SeqAn's CompareType of Dna5 and char is char)
									const char alt_ch = rec.alt[alt_start_pos[n_alt] + pos];
									if("ACGTN"[vcf_ref[pos]] != alt_ch){ inserted.push_back(dna5(alt_ch)); }
									else{ continue; }
								}
								// else: deletion, empty replacement
								bool has_n = false;
								for(uint8_t b : inserted){ has_n |= b > 3; }
								if(has_n){
									if(++errors <= kMaxErrorsShownPerFile){
										err(where + std::to_string(start_pos) + " has an alternative column containing ambiguous bases (e.g. N). Please change or remove them.");
									}
									if(errors >= kMaxErrorsShownPerFile){ err("Maximum number of errors reached. Additional errors are not shown for this file."); }
								}
								else{
									insert_variant(rec.rid, start_pos + pos, inserted, gt_has_var[n_alt]);
								}
							}
						}
					}
				}
			}
			if(stop){ break; }

			// next record (sortedness checks of the reference)
			bool got = false;
			while(std::getline(f, line)){
				strip_cr(line);
				got = true;
				break;
			}
			if(!got){ break; }
			if(!parse_record(line, contigs, n_samples, rec)){
				err("Could not read vcf record: " + parse_error_);
				++errors;
				break;
			}
			if(rec.rid < read_for_num_sequences){
				count_error("Variant file is not properly position sorted. Found sequence id " + std::to_string(rec.rid) + " after id " + std::to_string(read_for_num_sequences));
			}
			else if(rec.rid == read_for_num_sequences){
				if(static_cast<uint32_t>(rec.begin_pos) < start_pos && rec.begin_pos >= 0){
					count_error("Variant file is not properly position sorted. Found in sequence id " + std::to_string(rec.rid) + " position " + std::to_string(rec.begin_pos) + " after position " + std::to_string(start_pos));
				}
			}
			else{ read_for_num_sequences = rec.rid; }
		}
		if(in.corrupt()){ err("Could not read vcf record: corrupt or truncated gzip stream."); ++errors; }
		if(errors){ variants.clear(); variant_positions.clear(); fail_collected(); }
	}

	// Applies the variants of `allele` to `seq` (one variant per position and allele: overlapping records are rejected by read()).
	AlleleSequence materialise(uint32_t seq_id, const std::vector<uint8_t> &seq, uint32_t allele) const {
		AlleleSequence a;
		a.bases.reserve(seq.size() + seq.size() / 64);
		a.off.resize(seq.size() + 1);
		const auto &vars = variants.at(seq_id);
		size_t v = 0;
		for(uint32_t p = 0; p < seq.size(); ++p){
			a.off[p] = a.bases.size();
			const Variant *mine = nullptr;
			for(; v < vars.size() && vars[v].position == p; ++v){
				if(vars[v].in_allele(allele)){
					if(mine){ throw std::runtime_error("two variants of one allele at position " + std::to_string(p)); }
					mine = &vars[v];
				}
			}
			if(mine){ a.bases.insert(a.bases.end(), mine->var_seq.begin(), mine->var_seq.end()); }
			else{ a.bases.push_back(seq[p]); }
		}
		a.off[seq.size()] = a.bases.size();
		return a;
	}

	FlatVariants flatten() const {
		FlatVariants o;
		o.seq_first.push_back(0);
		o.bases_off.push_back(0);
		for(const auto &per_seq : variants){
			for(const auto &v : per_seq){
				o.position.push_back(v.position);
				o.bases.insert(o.bases.end(), v.var_seq.begin(), v.var_seq.end());
				o.bases_off.push_back(o.bases.size());
				o.allele_lo.push_back(v.allele[0]);
				o.allele_hi.push_back(v.allele[1]);
			}
			o.seq_first.push_back(o.position.size());
		}
		return o;
	}

	// Reference::InsertVariant (Reference.h:115-139): same position sorted deletion / substitution / insertion by
	// length; an identical replacement only adds its alleles.
	void insert_variant(uint32_t seq, uint32_t position, const std::vector<uint8_t> &var_seq, const std::array<uint64_t, 2> &allele){
		auto &vars = variants.at(seq);
		size_t insert_at = vars.size();
		size_t var = vars.size();
		while(var > 0 && vars[--var].position == position){
			if(vars[var].var_seq == var_seq){
				vars[var].allele[0] |= allele[0];
				vars[var].allele[1] |= allele[1];
				return;
			}
			else if(vars[var].var_seq.size() > var_seq.size()){ --insert_at; }
		}
		vars.insert(vars.begin() + insert_at, Variant{position, var_seq, allele});
	}

private:
	struct Record {
		uint32_t rid = 0;
		int32_t begin_pos = 0;
		std::string ref, alt;
		std::vector<std::string> genotypes;
	};
	std::string parse_error_;

	static void strip_cr(std::string &line){ if(!line.empty() && line.back() == '\r'){ line.pop_back(); } }
	static uint8_t dna5(char ch){   // SeqAn's char -> Dna5 table: ACGT/acgt (U/u as T), everything else N
		switch(ch){
			case 'A': case 'a': return 0;
			case 'C': case 'c': return 1;
			case 'G': case 'g': return 2;
			case 'T': case 't': case 'U': case 'u': return 3;
			default: return 4;
		}
	}
	static std::string upper(const std::vector<uint8_t> &codes){ std::string s; for(uint8_t c : codes){ s += "ACGTN"[c]; } return s; }
	static std::vector<std::string> split_tabs(const std::string &line){
		std::vector<std::string> out;
		size_t from = 0;
		while(true){
			const size_t tab = line.find('\t', from);
			out.push_back(line.substr(from, tab == std::string::npos ? std::string::npos : tab - from));
			if(tab == std::string::npos){ break; }
			from = tab + 1;
		}
		return out;
	}
	static std::string join_genotypes(const Record &rec){ std::string s; for(const auto &g : rec.genotypes){ s += ' '; s += g; } return s; }

	// _readVcfContig (read_vcf.h:67-99): value is "<ID=name,length=...>"; keys are scanned until "ID"
	std::string contig_id(const std::string &value, const std::string &path){
		size_t p = (!value.empty() && value[0] == '<') ? 1 : 0;
		while(p < value.size()){
			const size_t eq = value.find('=', p);
			if(eq == std::string::npos){ break; }
			if(value.compare(p, eq - p, "ID") == 0){
				const size_t end = value.find_first_of(",>", eq + 1);
				const std::string name = value.substr(eq + 1, end == std::string::npos ? std::string::npos : end - eq - 1);
				if(name.empty()){ fail("Could not prepare vcf file '" + path + "' for record readin: Contig ID value not found in header."); }
				return name;
			}
			const size_t sep = value.find_first_of(",>", eq);
			if(sep == std::string::npos){ break; }
			p = sep + 1;
		}
		fail("Could not prepare vcf file '" + path + "' for record readin: Contig ID key not found in header.");
		return {};
	}

	// readRecord (read_vcf.h:170-225); unknown contig names are appended to the name store and so get ids past the header's
	bool parse_record(const std::string &line, std::vector<std::string> &contigs, size_t n_samples, Record &rec){
		const auto fields = split_tabs(line);
		if(fields.size() < 8 + n_samples){ parse_error_ = "Not enough values in a line."; return false; }
		size_t rid = 0;
		while(rid < contigs.size() && contigs[rid] != fields[0]){ ++rid; }
		if(rid == contigs.size()){ contigs.push_back(fields[0]); }
		rec.rid = rid;
		const std::string &p = fields[1];
		size_t k = 0;
		bool neg = false;
		if(k < p.size() && (p[k] == '-' || p[k] == '+')){ neg = p[k] == '-'; ++k; }
		if(k == p.size()){ parse_error_ = "Unable to convert '" + p + "' into int."; return false; }
		int64_t v = 0;
		for(; k < p.size(); ++k){
			if(p[k] < '0' || p[k] > '9' || v > std::numeric_limits<int32_t>::max()){ parse_error_ = "Unable to convert '" + p + "' into int."; return false; }
			v = v * 10 + (p[k] - '0');
		}
		if(v > std::numeric_limits<int32_t>::max()){ parse_error_ = "Unable to convert '" + p + "' into int."; return false; }
		rec.begin_pos = static_cast<int32_t>(neg ? -v : v) - 1;
		rec.ref = fields[3];
		rec.alt = fields[4];
		if(fields[5] != "."){
			try{ size_t used = 0; (void)std::stof(fields[5], &used); if(used != fields[5].size()){ throw std::invalid_argument(""); } }
			catch(const std::exception &){ parse_error_ = "Unable to convert '" + fields[5] + "' into float."; return false; }
		}
		rec.genotypes.clear();
		const size_t first_sample = fields.size() > 8 + n_samples ? 9 : 8;
		for(size_t i = first_sample; i < fields.size(); ++i){ rec.genotypes.push_back(fields[i]); }
		return true;
	}

	void err(const std::string &msg){ diagnostics += msg; diagnostics += '\n'; }
	[[noreturn]] void fail(const std::string &msg){ err(msg); fail_collected(); }
	[[noreturn]] void fail_collected(){
		std::string what = diagnostics;
		while(!what.empty() && what.back() == '\n'){ what.pop_back(); }
		throw std::runtime_error(what);
	}
};

} // namespace rsq
