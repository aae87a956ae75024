// Host-compiled check of reseq_b200/csrc/mathx.cuh against the system libm (the one the reference links).
// usage: mathx_check <n> <seed>   -> prints mismatch counts; exit 0 iff none
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../../reseq_b200/csrc/mathx.cuh"

int main(int argc, char **argv){
	size_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1000000;
	uint64_t seed = argc > 2 ? strtoull(argv[2], nullptr, 10) : 1;
	std::mt19937_64 gen(seed);
	std::uniform_real_distribution<double> u01(0.0, 1.0);
	size_t bad_exp = 0, bad_pow = 0, bad_logit = 0;
	volatile double sink = 0;
	for(size_t i = 0; i < n; ++i){
		// exp: arguments as in InvLogit2 (sums of three logit-scale biases), plus wide-range and special magnitudes
		double x;
		switch(i % 4){
		case 0: x = (u01(gen) - 0.5) * 8.0; break;
		case 1: x = (u01(gen) - 0.5) * 1600.0; break;
		case 2: x = std::ldexp(u01(gen) - 0.5, static_cast<int>(gen() % 80) - 70); break;
		default: x = -u01(gen) * 745.2; break;
		}
		double e0 = std::exp(x), e1 = rsq::exp_glibc(x);
		if(rsq::as_u64(e0) != rsq::as_u64(e1)){ if(bad_exp < 5) printf("exp(%a): libm %a port %a\n", x, e0, e1); ++bad_exp; }
		double l0 = 2/(1+std::exp(-x)), l1 = rsq::inv_logit2(x);
		if(rsq::as_u64(l0) != rsq::as_u64(l1)){ ++bad_logit; }
		// pow: base 1-p in [0,1], exponent = dispersion r over many magnitudes (NegativeBinomial), plus thresholds^(2n)
		double base, ex;
		switch(i % 5){
		case 0: base = u01(gen); ex = u01(gen) * 10; break;
		case 1: base = 1.0 - std::ldexp(u01(gen), -static_cast<int>(gen() % 50)); ex = std::ldexp(u01(gen), static_cast<int>(gen() % 60)); break;
		case 2: base = std::ldexp(u01(gen), -static_cast<int>(gen() % 1070)); ex = u01(gen) * 3; break;
		case 3: base = u01(gen); ex = std::ldexp(u01(gen), static_cast<int>(gen() % 140) - 70); break;
		default: base = 1.0 - u01(gen) * 1e-3; ex = static_cast<double>(1 + gen() % 256); break;
		}
		double p0 = std::pow(base, ex), p1 = rsq::pow_glibc(base, ex);
		if(rsq::as_u64(p0) != rsq::as_u64(p1)){ if(bad_pow < 5) printf("pow(%a,%a): libm %a port %a\n", base, ex, p0, p1); ++bad_pow; }
		sink = sink + p1;
	}
	// edge values
	const double xs[] = {0.0, 1.0, 0.5, 1e-320, 1e-300, 0.9999999999999999, 2.2250738585072014e-308};
	const double ys[] = {0.0, 1.0, 2.0, 1e-30, 1e30, 1e300, 0.5, 1e19, 9.3e18, 1e-20};
	for(double x : xs) for(double y : ys){
		double p0 = std::pow(x, y), p1 = rsq::pow_glibc(x, y);
		if(rsq::as_u64(p0) != rsq::as_u64(p1)){ printf("pow(%a,%a): libm %a port %a\n", x, y, p0, p1); ++bad_pow; }
	}
	printf("n=%zu exp_mismatch=%zu logit_mismatch=%zu pow_mismatch=%zu\n", n, bad_exp, bad_logit, bad_pow);
	return (bad_exp || bad_pow || bad_logit) ? 1 : 0;
}
