// ordered_sum.cuh against the plain chain: the four passes (plain chunk sums, start values, exact runs, in-order resolve) must give the bits of
// s = RN(RN(RN(0 + x_0) + x_1) + ...) for any non-negative terms.  g++ -O2 -std=c++17 -ffp-contract=off
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>
#include "../../reseq_b200/csrc/ordered_sum.cuh"

using namespace rsq;

struct VecTerm { const double *v; double operator()(uint32_t p) const { return v[p]; } };

static uint64_t bits(double d){ uint64_t b; memcpy(&b, &d, 8); return b; }

static bool run_case(const std::vector<double> &x, uint32_t K, uint64_t &reran_total, uint64_t &chunks_total){
	const uint32_t n = x.size();
	double plain = 0.0;
	for(double v : x){ plain = plain + v; }
	VecTerm term{x.data()};
	const uint32_t n_chunks = (n + K - 1) / K;
	std::vector<double> p(n_chunks), g(n_chunks), o(n_chunks);
	std::vector<uint32_t> tie(n_chunks);
	double mx = 0.0;
	for(uint32_t c = 0; c < n_chunks; ++c){ p[c] = chunk_plain_sum(term, c * K, std::min(n, (c + 1) * K), mx); }       // pass A
	double acc = 0.0;
	for(uint32_t c = 0; c < n_chunks; ++c){ g[c] = acc; acc = acc + p[c]; }                                              // pass B
	for(uint32_t c = 0; c < n_chunks; ++c){ const ChunkRun r = chunk_exact_run(term, c * K, std::min(n, (c + 1) * K), g[c]); o[c] = r.out; tie[c] = r.tie; }   // pass C
	double s = 0.0; uint32_t reran = 0;
	for(uint32_t c = 0; c < n_chunks; ++c){ s = chunk_resolve(term, c * K, std::min(n, (c + 1) * K), s, g[c], o[c], tie[c], reran); }   // pass D
	reran_total += reran; chunks_total += n_chunks;
	double want_max = 0.0; for(double v : x){ if(v > want_max){ want_max = v; } }
	if(bits(s) != bits(plain) || bits(mx) != bits(want_max)){
		printf("MISMATCH n=%u K=%u: %a vs %a (max %a vs %a)\n", n, K, s, plain, mx, want_max);
		return false;
	}
	return true;
}

int main(int argc, char **argv){
	const uint64_t seed = argc > 1 ? strtoull(argv[1], nullptr, 10) : 5;
	std::mt19937_64 rng(seed);
	std::uniform_real_distribution<double> uni(0.0, 1.0);
	uint64_t bad = 0, cases = 0, reran = 0, chunks = 0;
	const uint32_t Ks[] = {1, 7, 64, 256, 1024};
	for(int rep = 0; rep < 60; ++rep){
		const uint32_t n = rep < 50 ? 1 + rng() % 20000 : 400000 + rng() % 400000;
		std::vector<double> x(n);
		const int kind = rep % 6;
		double run = 0.0;
		for(uint32_t i = 0; i < n; ++i){
			double v;
			switch(kind){
			case 0: v = uni(rng); break;                                             // same magnitude (the bias terms)
			case 1: v = std::ldexp(uni(rng), static_cast<int>(rng() % 80) - 40); break;   // 24 orders of magnitude
			case 2: v = (rng() % 4 == 0) ? 0.0 : uni(rng) * 1e-3; break;                // zeros in between
			case 3: {                                                                   // forced ties: half an ulp of the running sum plus a multiple
				if(run > 0.0 && rng() % 3 == 0){
					int e; std::frexp(run, &e);   // run = m * 2^e, m in [0.5, 1): ulp = 2^(e-53)
					v = std::ldexp(1.0, e - 54) + std::ldexp(static_cast<double>(rng() % 1000), e - 53);
				}
				else{ v = uni(rng); }
				break; }
			case 4: v = (rng() % 50 == 0) ? run * (1.0 + uni(rng)) : uni(rng) * 1e-6; break;   // jumps over binade boundaries
			default: v = std::ldexp(static_cast<double>(rng() % 16), -2); break;        // exactly representable steps: everything ties or is exact
			}
			x[i] = v;
			run = run + v;
		}
		for(uint32_t K : Ks){
			if(n > 100000 && K < 64){ continue; }
			++cases;
			if(!run_case(x, K, reran, chunks)){ ++bad; }
		}
	}
	// denormal and tiny starts
	{
		std::vector<double> x(5000);
		for(auto &v : x){ v = std::ldexp(uni(rng), -1070 + static_cast<int>(rng() % 30)); }
		for(uint32_t K : Ks){ ++cases; if(!run_case(x, K, reran, chunks)){ ++bad; } }
	}
	printf("cases=%llu mismatches=%llu chunks=%llu reran=%llu\n", (unsigned long long)cases, (unsigned long long)bad, (unsigned long long)chunks, (unsigned long long)reran);
	return bad ? 1 : 0;
}
