// TEST HARNESS: Genome::read_fasta + Genome::replace_n (host_profile.hpp) -> the Dna5 codes of every sequence, concatenated, on stdout.
#include <cstdio>
#include <cstdlib>
#include "../../reseq_b200/csrc/host_profile.hpp"

int main(int argc, char **argv){
	if(argc != 3){ fprintf(stderr, "usage: replace_n_check <ref.fa[.gz]> <seed>\n"); return 64; }
	try{
		rsq::Genome g;
		g.read_fasta(argv[1]);
		g.replace_n(strtoull(argv[2], nullptr, 10));
		for(const auto &q : g.seqs){ fwrite(q.data(), 1, q.size(), stdout); }
	}
	catch(const std::exception &ex){ fprintf(stderr, "%s\n", ex.what()); return 1; }
	return 0;
}
