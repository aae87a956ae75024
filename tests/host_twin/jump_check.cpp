// TEST HARNESS: the jump-ahead polynomials of reseq_b200/csrc/mt_jump_tables.inc against std::mt19937_64::discard.
// Tempering is linear, so the identity  out[n+J+j] = XOR_{i in g_J} out[n+i+j]  holds for the generator's outputs too.
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>
#include "../../reseq_b200/csrc/mt_jump_tables.inc"

int main(){
	int bad = 0;
	for(int t = 0; t < kMtJumpTables; ++t){
		for(uint64_t seed : {42ull, 5489ull, 0xdeadbeefcafeull}){
			std::mt19937_64 a(seed), b(seed);
			std::vector<uint64_t> outs(19937 + 312);
			for(auto &o : outs){ o = a(); }
			b.discard(1ull << kMtJumpLog2[t]);
			for(int j = 0; j < 312; ++j){
				uint64_t acc = 0;
				for(int w = 0; w < 312; ++w){
					uint64_t bits = kMtJumpPoly[t][w];
					while(bits){ acc ^= outs[w * 64 + __builtin_ctzll(bits) + j]; bits &= bits - 1; }
				}
				if(acc != b()){ ++bad; }
			}
		}
	}
	printf("jump_mismatches=%d tables=%d\n", bad, kMtJumpTables);
	return bad ? 1 : 0;
}
