// TEST HARNESS: Genome::read_fasta + VariantSet::read (reseq_b200/csrc/variants.hpp) -> the same text `oracle/dump_tables variants`
// writes from the reference's Reference::variants_ ("alleles N", then "<seq> <position> <var_seq or -> <bits lo> <bits hi>").
// With "allele <seq> <a>" as 4th-6th argument: VariantSet::materialise of that allele - the sequence, then "<p> <off[p]>" for every
// position where the map changes slope, then for 16 fixed indices "<index> <ref_position(index)>".
// With a 4th argument "positions": Reference::variant_positions_ (ReadFirstVariantPositions) as "<seq> <position>" lines.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../../reseq_b200/csrc/host_profile.hpp"
#include "../../reseq_b200/csrc/variants.hpp"

int main(int argc, char **argv){
	if(argc < 3){ fprintf(stderr, "usage: variants_check <ref.fa> <in.vcf[.gz]> [positions]\n"); return 64; }
	try{
		rsq::Genome g;
		g.read_fasta(argv[1]);
		std::vector<std::string> ids;
		for(size_t i = 0; i < g.seqs.size(); ++i){ ids.push_back(g.first_part(i)); }
		rsq::VariantSet vs;
		const bool positions = argc > 3 && !strcmp(argv[3], "positions");
		vs.read(argv[2], ids, g.seqs, positions);
		if(argc > 5 && !strcmp(argv[3], "allele")){
			const uint32_t seq = atoi(argv[4]), allele = atoi(argv[5]);
			const rsq::AlleleSequence a = vs.materialise(seq, g.seqs.at(seq), allele);
			for(uint8_t b : a.bases){ putchar("ACGTN"[b]); }
			putchar('\n');
			for(size_t p = 1; p < a.off.size(); ++p){ if(a.off[p] != a.off[p - 1] + 1){ printf("%zu %u\n", p, a.off[p]); } }
			for(uint32_t k = 0; k < 16; ++k){ const uint32_t idx = a.bases.size() * k / 16 + k; printf("i %u %u\n", idx, a.ref_position(idx)); }
			return 0;
		}
		if(positions){
			for(size_t s = 0; s < vs.variant_positions.size(); ++s){
				for(uint32_t p : vs.variant_positions[s]){ printf("%zu %u\n", s, p); }
			}
			return 0;
		}
		printf("alleles %u\n", vs.num_alleles);
		const rsq::FlatVariants f = vs.flatten();   // printed from the flattened (device) layout
		for(size_t s = 0; s + 1 < f.seq_first.size(); ++s){
			for(uint32_t v = f.seq_first[s]; v < f.seq_first[s + 1]; ++v){
				printf("%zu %u ", s, f.position[v]);
				if(f.bases_off[v] == f.bases_off[v + 1]){ putchar('-'); }
				for(uint32_t k = f.bases_off[v]; k < f.bases_off[v + 1]; ++k){ putchar("ACGT"[f.bases[k]]); }
				printf(" %llx %llx\n", (unsigned long long)f.allele_lo[v], (unsigned long long)f.allele_hi[v]);
			}
		}
	}
	catch(const std::exception &ex){ fprintf(stderr, "%s\n", ex.what()); printf("rejected\n"); return 1; }
	return 0;
}
