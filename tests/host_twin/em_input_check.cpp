// TEST HARNESS: read_em_input (reseq_b200/csrc/em_input.hpp) -> a plain text dump of what the kernels would get.
#include <cstdio>
#include "../../reseq_b200/csrc/em_input.hpp"

int main(int argc, char **argv){
	if(argc != 2){ fprintf(stderr, "usage: em_input_check <frags.fa[.gz]>\n"); return 64; }
	try{
		const rsq::EmInput in = rsq::read_em_input(argv[1]);
		printf("records %zu max_len %u max_id_len %u\n", in.recs.size(), in.max_len, in.max_id_len);
		for(const auto &r : in.recs){
			printf("%.*s %u %u ", (int)r.id_len, in.ids.data() + r.id_off, r.seg, r.fragment_length);
			for(uint32_t k = 0; k < r.len; ++k){ putchar("ACGTN"[in.seq[r.seq_off + k]]); }
			putchar(' ');
			for(uint32_t k = 0; k < r.len; ++k){ putchar("ACGTN"[in.dom[r.seq_off + k]]); }
			for(uint32_t k = 0; k < r.len; ++k){ printf(" %u", in.rate[r.seq_off + k]); }
			putchar('\n');
		}
	}
	catch(const std::exception &ex){ fprintf(stderr, "%s\n", ex.what()); return 1; }
	return 0;
}
