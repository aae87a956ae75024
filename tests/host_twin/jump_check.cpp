// TEST HARNESS: the jump-ahead polynomials of reseq_b200/csrc/mt_jump_tables.inc.
//   * tables up to 2^22: the window they produce == the window the plain mt19937_64 recurrence reaches after 2^k steps, and the
//     first outputs behind it == std::mt19937_64 after discard(2^k) (tempering is a bijection per word);
//   * larger tables, inductively: jumping twice by 2^(k-1) == jumping once by 2^k.
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>
#include "../../reseq_b200/csrc/mt_jump_tables.inc"

typedef std::vector<uint64_t> Window;   // 312 consecutive untempered words x[n .. n+312)

static void extend(std::vector<uint64_t> &x, size_t count){   // x[k+312] = x[k+156] ^ twist(x[k], x[k+1])
	for(size_t c = 0; c < count; ++c){
		const size_t k = x.size() - 312;
		const uint64_t y = (x[k] & 0xFFFFFFFF80000000ull) | (x[k + 1] & 0x7FFFFFFFull);
		x.push_back(x[k + 156] ^ (y >> 1) ^ ((y & 1ull) ? 0xB5026F5AA96619E9ull : 0ull));
	}
}
static Window jump(const Window &w, int table){   // what k_master_jump_gen + k_master_jump_xor compute
	std::vector<uint64_t> seq(w);
	extend(seq, 19937 + 8);
	Window out(312, 0);
	for(int word = 0; word < 312; ++word){
		uint64_t bits = kMtJumpPoly[table][word];
		while(bits){
			const int i = word * 64 + __builtin_ctzll(bits);
			for(int j = 0; j < 312; ++j){ out[j] ^= seq[i + j]; }
			bits &= bits - 1;
		}
	}
	return out;
}
static uint64_t temper(uint64_t x){
	x ^= (x >> 29) & 0x5555555555555555ull; x ^= (x << 17) & 0x71D67FFFEDA60000ull; x ^= (x << 37) & 0xFFF7EEE000000000ull; x ^= (x >> 43);
	return x;
}

int main(){
	int bad = 0;
	for(uint64_t seed : {42ull, 0xdeadbeefcafeull}){
		// window of generated words right behind the seed state (the device never jumps from the seed words themselves)
		std::vector<uint64_t> x(312);
		x[0] = seed;
		for(int i = 1; i < 312; ++i){ x[i] = 6364136223846793005ull * (x[i - 1] ^ (x[i - 1] >> 62)) + static_cast<uint64_t>(i); }
		extend(x, 312 + (1u << 22) + 312);
		const Window w0(x.begin() + 312, x.begin() + 624);
		Window prev;
		for(int t = 0; t < kMtJumpTables; ++t){
			const int k = kMtJumpLog2[t];
			const Window w = jump(w0, t);
			if(k <= 22){
				for(int j = 0; j < 312; ++j){ if(w[j] != x[312 + (1ull << k) + j]){ ++bad; } }
				std::mt19937_64 gen(seed);
				gen.discard(1ull << k);
				for(int j = 0; j < 8; ++j){ if(temper(w[j]) != gen()){ ++bad; } }
			}
			else{
				const Window twice = jump(prev, t - 1);   // prev = jump(w0, t-1)
				if(kMtJumpLog2[t - 1] != k - 1 || twice != w){ ++bad; }
			}
			prev = w;
		}
	}
	printf("jump_mismatches=%d tables=%d\n", bad, kMtJumpTables);
	return bad ? 1 : 0;
}
