// TEST HARNESS: splice_reference (reseq_b200/csrc/variant_core.cuh) driven by text commands on stdin, one result line per call.
//   load <ref.fa> <in.vcf>                      sequences + variants through Genome::read_fasta / VariantSet::read (flattened layout)
//   seq <ACGT...>                               a single sequence 0 without variants
//   var <position> <bases or -> <lo> <hi>       append a variant to sequence 0 (hex allele words)
//   alleles <seed> <n>                          replays `oracle/dump_tables alleles <seed> <n>`: same stream, same argument draws, choose_alleles
//   sysfile <path>                              chain description of `oracle/dump_tables syserrvar` -> its "walk" lines recomputed by sys_error_with_variants
//   trace <ref.fa> <in.vcf> <seq> <trace file>  `oracle/dump_tables biasmod` trace: every "f" line recomputed by allele_fragment on VariantSet::materialise
//                                               of the traced (post-ReplaceN) sequence; prints "<lines checked> <mismatches>" (+ the first mismatches)
//   negbin <path>                               lines of `oracle/dump_tables negbin` recomputed by allele_fragment_counts: "<lines> <mismatches>"
//   call <seq> <start> <len> <reversed> <first variant> <posCurrentlyAt> <allele>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <random>
#include <sstream>
#include "../../reseq_b200/csrc/host_profile.hpp"
#include "../../reseq_b200/csrc/variant_core.cuh"

int main(){
	rsq::Genome g;
	std::vector<rsq::FlatVariants> flat;   // one per sequence
	std::string line;
	try{
		while(std::getline(std::cin, line)){
			std::istringstream in(line);
			std::string cmd;
			in >> cmd;
			if(cmd == "load"){
				std::string fa, vcf;
				in >> fa >> vcf;
				g = rsq::Genome();
				g.read_fasta(fa);
				g.read_variants(vcf);
				flat.clear();
				for(const auto &per_seq : g.variants.variants){
					rsq::VariantSet one;
					one.variants.push_back(per_seq);
					flat.push_back(one.flatten());
				}
			}
			else if(cmd == "seq"){
				std::string bases;
				in >> bases;
				g = rsq::Genome();
				g.ids.push_back("s");
				g.seqs.emplace_back(bases.size());
				rsq::Genome::encode(bases.data(), bases.size(), g.seqs[0].data());
				flat.assign(1, rsq::FlatVariants());
				flat[0].bases_off.push_back(0);
			}
			else if(cmd == "var"){
				uint32_t pos; std::string bases, lo, hi;
				in >> pos >> bases >> lo >> hi;
				auto &f = flat.at(0);
				f.position.push_back(pos);
				if(bases != "-"){ for(char ch : bases){ f.bases.push_back(rsq::Genome::code(ch)); } }
				f.bases_off.push_back(f.bases.size());
				f.allele_lo.push_back(std::stoull(lo, nullptr, 16));
				f.allele_hi.push_back(std::stoull(hi, nullptr, 16));
			}
			else if(cmd == "alleles"){
				uint64_t seed, n_calls;
				in >> seed >> n_calls;
				std::mt19937_64 gen(seed);
				std::uniform_real_distribution<double> zero_to_one(0.0, 1.0);   // GeneralRandomDistributions::ZeroToOne
				for(uint64_t call = 0; call < n_calls; ++call){
					const uint32_t alleles = call % 7 == 0 ? 1 + gen() % 128 : 1 + gen() % 6;
					const uint32_t possible = 2 * alleles;
					const uint32_t non_zero = 1 + gen() % possible;
					uint16_t chosen[256];
					const uint32_t n = rsq::choose_alleles(chosen, non_zero, possible, [&](){ return zero_to_one(gen); });
					printf("%u %u", possible, non_zero);
					for(uint32_t k = 0; k < n; ++k){ printf(" %u", chosen[k]); }
					putchar('\n');
				}
			}
			else if(cmd == "sysfile"){
				std::string path;
				in >> path;
				std::ifstream f(path);
				std::vector<uint16_t> sys, errs;
				std::vector<uint32_t> block_start{0}, var_first{0}, position, err_off{0};
				std::vector<uint64_t> lo, hi;
				auto entry = [](const std::string &hex, size_t k){ return static_cast<uint16_t>(std::stoul(hex.substr(4 * k, 2), nullptr, 16) | (std::stoul(hex.substr(4 * k + 2, 2), nullptr, 16) << 8)); };
				std::string l;
				while(std::getline(f, l)){
					std::istringstream li(l);
					std::string tag;
					li >> tag;
					if(tag == "b"){
						uint32_t len, n_var;
						li >> len >> n_var;
						block_start.push_back(block_start.back() + len);
						var_first.push_back(var_first.back() + n_var);
					}
					else if(tag == "s"){
						std::string hex;
						li >> hex;
						for(size_t k = 0; k < hex.size() / 4; ++k){ sys.push_back(entry(hex, k)); }
					}
					else if(tag == "v"){
						uint32_t pos, n_err; uint64_t bits; std::string hex;
						li >> pos >> bits >> n_err >> hex;
						position.push_back(pos); lo.push_back(bits); hi.push_back(0);
						for(size_t k = 0; k < n_err; ++k){ errs.push_back(entry(hex, k)); }
						err_off.push_back(errs.size());
					}
					else if(tag == "walk"){
						rsq::SysErrorVariantView view{sys.data(), block_start.data(), var_first.data(), position.data(), err_off.data(), errs.data(), lo.data(), hi.data()};
						rsq::SysErrorCursor c{0, 0, 0, 0};
						uint32_t allele, steps;
						li >> c.block >> c.block_pos >> c.cur_var >> allele >> steps;
						printf("walk %u %u %d %u %u ", c.block, c.block_pos, c.cur_var, allele, steps);
						for(uint32_t k = 0; k < steps && c.block + 1 < block_start.size(); ++k){
							const uint16_t e = rsq::sys_error_with_variants(view, c, allele);
							printf("%02x%02x", e & 0xffu, e >> 8);
						}
						putchar('\n');
					}
				}
			}
			else if(cmd == "trace"){
				std::string fa, vcf, path; uint32_t seq;
				in >> fa >> vcf >> seq >> path;
				rsq::Genome tg;
				tg.read_fasta(fa);
				tg.read_variants(vcf);   // checked against the raw reference; the traced sequence below only differs where that one has N
				std::ifstream f(path);
				std::string l;
				std::getline(f, l);
				std::vector<uint8_t> bases(l.size() - 2);
				rsq::Genome::encode(l.data() + 2, bases.size(), bases.data());
				std::vector<rsq::AlleleSequence> alleles;
				std::vector<std::vector<uint32_t>> gc_prefix;
				for(uint32_t a = 0; a < tg.variants.num_alleles; ++a){
					alleles.push_back(tg.variants.materialise(seq, bases, a));
					std::vector<uint32_t> pre(alleles.back().bases.size() + 1, 0);
					for(size_t k = 0; k < alleles.back().bases.size(); ++k){ pre[k + 1] = pre[k] + (alleles.back().bases[k] == 1 || alleles.back().bases[k] == 2); }
					gc_prefix.push_back(std::move(pre));
				}
				uint32_t start = 0, svp = 0;
				uint64_t checked = 0, bad = 0;
				while(std::getline(f, l)){
					std::istringstream li(l);
					std::string tag;
					li >> tag;
					if(tag == "p"){ int32_t first_var; li >> start >> first_var >> svp; continue; }
					uint32_t len, allele; int32_t shift; std::string gc;
					uint32_t sur[6];
					li >> len >> allele >> shift >> gc;
					for(auto &v : sur){ li >> v; }
					const auto &as = alleles.at(allele);
					rsq::AlleleView view{as.bases.data(), as.off.data(), gc_prefix[allele].data(), static_cast<uint32_t>(as.bases.size()), static_cast<uint32_t>(bases.size())};
					const uint32_t first = as.off[start] + svp;
					if(first < 40 || first + len + 40 > as.bases.size()){ continue; }   // circular surroundings at the sequence ends: not part of this check
					rsq::AlleleFragment fr;
					rsq::allele_fragment(view, start, svp, len, fr);
					bool ok = fr.end_position == start + len + shift && (gc == "-" || fr.gc_percent == std::stoul(gc));
					for(int k = 0; k < 3; ++k){ ok = ok && fr.sur_start[k] == sur[k] && fr.sur_end[k] == sur[3 + k]; }
					++checked;
					if(!ok && ++bad <= 3){ fprintf(stderr, "mismatch at start %u +%u: %s\n", start, svp, l.c_str()); }
				}
				printf("%llu %llu\n", (unsigned long long)checked, (unsigned long long)bad);
			}
			else if(cmd == "negbin"){
				std::string path;
				in >> path;
				std::ifstream f(path);
				std::string m, a, b, u;
				uint32_t alleles, want;
				uint64_t n = 0, bad = 0;
				auto dbl = [](const std::string &hex){ const uint64_t bits = std::stoull(hex, nullptr, 16); double v; memcpy(&v, &bits, 8); return v; };
				while(f >> m >> a >> b >> alleles >> u >> want){
					bool runaway = false;
					const uint32_t got = rsq::allele_fragment_counts(dbl(m), dbl(a), dbl(b), alleles, dbl(u), runaway);
					++n;
					if(got != want && ++bad <= 3){ fprintf(stderr, "mismatch: %s %s %s %u %s -> %u, reference %u\n", m.c_str(), a.c_str(), b.c_str(), alleles, u.c_str(), got, want); }
				}
				printf("%llu %llu\n", (unsigned long long)n, (unsigned long long)bad);
			}
			else if(cmd == "select"){   // SimulatorTest::TestSelectAllele: every id of `possible` drawn with random value 0.5
				uint32_t possible;
				in >> possible;
				uint16_t chosen[256];
				// (the direct branch draws at most half of the ids: the first half of the reference's known answer)
				const uint32_t n = rsq::choose_alleles(chosen, possible / 2, possible, [](){ return 0.5; });
				for(uint32_t k = 0; k < n; ++k){ printf(k ? " %u" : "%u", chosen[k]); }
				putchar('\n');
			}
			else if(cmd == "call"){
				uint32_t s, start, len, reversed, first_pos, allele; int32_t first;
				in >> s >> start >> len >> reversed >> first >> first_pos >> allele;
				const auto &f = flat.at(s);
				rsq::VariantView view{f.position.data(), f.bases_off.data(), f.bases.data(), f.allele_lo.data(), f.allele_hi.data(), static_cast<uint32_t>(f.position.size())};
				std::vector<uint8_t> out(len + 1, 9);
				const uint32_t n = rsq::splice_reference(out.data(), g.seqs.at(s).data(), view, start, len, reversed, first, first_pos, allele);
				if(out[len] != 9){ throw std::runtime_error("splice_reference wrote past frag_length"); }
				std::string text;
				for(uint32_t k = 0; k < n; ++k){ text += "ACGT"[out[k]]; }
				puts(text.c_str());
			}
		}
	}
	catch(const std::exception &ex){ fprintf(stderr, "%s\n", ex.what()); return 1; }
	return 0;
}
