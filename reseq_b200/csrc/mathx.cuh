// Bit-exact device restatement of glibc 2.39 exp() and pow() (x86-64, FMA code path) for the two call
// sites on the simulation hot path:
//   utilities::InvLogit2            -> exp(-bias)                 (reference utilities.hpp:505-507)
//   FragmentDistributionStats::NegativeBinomial -> pow(1-p, r)   (reference FragmentDistributionStats.cpp:3604)
// The reference links the system libm; on every CPU with FMA the ifunc resolver selects __exp_fma /
// __pow_fma.  The operation order below (every fma, every plain mul/add) follows the instruction
// sequence of those two functions in this image's libm.so.6 (algorithm: Szabolcs Nagy's exp/pow from
// ARM optimized-routines as adopted by glibc >= 2.28, sysdeps/ieee754/dbl-64/e_exp.c, e_pow.c), so
// results agree bit for bit; tests/test_mathx.py checks >1e7 arguments per function against libm.
//
// Compiles as host code too (test twin); the product only uses the device instantiation.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define RSQ_HD __host__ __device__ __forceinline__
#define RSQ_HD_NOINLINE __host__ __device__
#define RSQ_HD_COLD __host__ __device__ __forceinline__   // (out-of-line variants of these were measured: slower - call overhead and spills outweigh the smaller code)
// helpers that only the variant-aware instantiations call (A/B build switch: out of line keeps those kernels' code smaller)
#ifdef RSQ_VCOLD_NOINLINE
#define RSQ_HD_VCOLD __host__ __device__ __noinline__
#else
#define RSQ_HD_VCOLD __host__ __device__ __forceinline__
#endif
#else
#define RSQ_HD_VCOLD inline __attribute__((noinline))
#define RSQ_HD inline
#define RSQ_HD_NOINLINE
#define RSQ_HD_COLD inline __attribute__((noinline))
#endif

namespace rsq {

#if defined(__CUDA_ARCH__)
#define RSQ_TABLE_QUALIFIER __device__ const
#else
#define RSQ_TABLE_QUALIFIER static const
#endif
#include "libm_tables.inc"
#undef RSQ_TABLE_QUALIFIER

RSQ_HD double as_double(uint64_t u){
#if defined(__CUDA_ARCH__)
	return __longlong_as_double(static_cast<long long>(u));
#else
	double d; __builtin_memcpy(&d, &u, 8); return d;
#endif
}
RSQ_HD uint64_t as_u64(double d){
#if defined(__CUDA_ARCH__)
	return static_cast<uint64_t>(__double_as_longlong(d));
#else
	uint64_t u; __builtin_memcpy(&u, &d, 8); return u;
#endif
}
RSQ_HD double fma_rn(double a, double b, double c){
#if defined(__CUDA_ARCH__)
	return __fma_rn(a, b, c);
#else
	return __builtin_fma(a, b, c);
#endif
}
// Plain IEEE operations that must never be contracted (translation units are built with
// -fmad=false / -ffp-contract=off as well; the intrinsics make the intent explicit on device).
RSQ_HD double mul_rn(double a, double b){
#if defined(__CUDA_ARCH__)
	return __dmul_rn(a, b);
#else
	return a * b;
#endif
}
RSQ_HD double add_rn(double a, double b){
#if defined(__CUDA_ARCH__)
	return __dadd_rn(a, b);
#else
	return a + b;
#endif
}
RSQ_HD double sub_rn(double a, double b){
#if defined(__CUDA_ARCH__)
	return __dsub_rn(a, b);
#else
	return a - b;
#endif
}

namespace detail {
RSQ_HD double ehdr(int i){ return as_double(kExpHdr[i]); }
// exp's tail when the scale factor cannot be represented directly (512 <= |x| < 1024)
RSQ_HD double exp_specialcase(double tmp, uint64_t sbits, uint64_t ki){
	if((ki & 0x80000000ull) == 0){
		sbits -= 1009ull << 52;
		double scale = as_double(sbits);
		return mul_rn(as_double(0x7f00000000000000ull) /*0x1p1009*/, fma_rn(scale, tmp, scale));
	}
	sbits += 1022ull << 52;
	double scale = as_double(sbits);
	double st = mul_rn(scale, tmp);
	double y = add_rn(scale, st);
	if(y < 1.0){
		double lo = add_rn(sub_rn(scale, y), st);
		double hi = add_rn(1.0, y);
		lo = add_rn(add_rn(sub_rn(1.0, hi), y), lo);
		y = sub_rn(add_rn(hi, lo), 1.0);
		if(y == 0.0){ y = 0.0; }
	}
	return mul_rn(as_double(0x0010000000000000ull) /*0x1p-1022*/, y);
}

// Shared core of exp(x) and pow's exp_inline(x, xtail): returns 2^(k/N)*exp(r)
RSQ_HD double exp_core(double x, double xtail, bool with_tail, uint32_t abstop, uint64_t sign_bias){
	const double InvLn2N = ehdr(0), Shift = ehdr(1), NegLn2hiN = ehdr(2), NegLn2loN = ehdr(3);
	const double C2 = ehdr(4), C3 = ehdr(5), C4 = ehdr(6), C5 = ehdr(7);
	double kd = fma_rn(x, InvLn2N, Shift);
	uint64_t ki = as_u64(kd);
	kd = sub_rn(kd, Shift);
	double r = fma_rn(kd, NegLn2hiN, x);
	r = fma_rn(kd, NegLn2loN, r);
	if(with_tail){ r = add_rn(r, xtail); }
	uint64_t idx = 2 * (ki & 127);
	uint64_t top = (ki + sign_bias) << 45;
	double tail = as_double(kExpTab[idx]);
	uint64_t sbits = kExpTab[idx + 1] + top;
	double a = fma_rn(r, C3, C2);
	double t = add_rn(r, tail);
	double r2 = mul_rn(r, r);
	double b = fma_rn(r, C5, C4);
	double s1 = fma_rn(a, r2, t);
	double r4 = mul_rn(r2, r2);
	double tmp = fma_rn(r4, b, s1);
	if(abstop == 0){ return exp_specialcase(tmp, sbits, ki); }
	double scale = as_double(sbits);
	return fma_rn(scale, tmp, scale);
}
}  // namespace detail

// exp(x), glibc 2.39 __exp_fma
RSQ_HD_NOINLINE double exp_glibc(double x){
	uint64_t ix = as_u64(x);
	uint32_t abstop = static_cast<uint32_t>(ix >> 52) & 0x7ff;
	if(abstop - 0x3c9u >= 0x3fu){
		if(abstop - 0x3c9u >= 0x80000000u){
			return add_rn(1.0, x);  // |x| < 2^-54
		}
		if(abstop >= 0x409u){  // |x| >= 1024, inf, nan
			if(ix == 0xfff0000000000000ull){ return 0.0; }
			if(abstop >= 0x7ffu){ return add_rn(1.0, x); }
			if(ix >> 63){ return 0.0; }                       // underflow
			return as_double(0x7ff0000000000000ull);          // overflow
		}
		abstop = 0;
	}
	return detail::exp_core(x, 0.0, false, abstop, 0);
}

// pow(x, y) for x >= 0 (the reference only raises probabilities in [0,1] to positive powers);
// glibc 2.39 __pow_fma.  Negative / NaN bases are outside the hot path's domain and return NaN.
RSQ_HD_NOINLINE double pow_glibc(double x, double y){
	uint64_t ix = as_u64(x), iy = as_u64(y);
	uint32_t topx = static_cast<uint32_t>(ix >> 52), topy = static_cast<uint32_t>(iy >> 52);
	if(topx - 1u >= 0x7feu || (topy & 0x7ffu) - 0x3beu >= 0x80u){
		// y is 0, inf or nan
		if(2 * iy - 1 >= 2 * 0x7ff0000000000000ull - 1){
			if(2 * iy == 0){ return 1.0; }
			if(ix == 0x3ff0000000000000ull){ return 1.0; }
			if(2 * ix > 2 * 0x7ff0000000000000ull || 2 * iy > 2 * 0x7ff0000000000000ull){ return add_rn(x, y); }
			if(2 * ix == 2 * 0x3ff0000000000000ull){ return 1.0; }
			if((2 * ix < 2 * 0x3ff0000000000000ull) == !(iy >> 63)){ return 0.0; }
			return mul_rn(y, y);
		}
		// x is 0, inf or nan
		if(2 * ix - 1 >= 2 * 0x7ff0000000000000ull - 1){
			double x2 = mul_rn(x, x);
			return (iy >> 63) ? 1.0 / x2 : x2;
		}
		if(ix >> 63){ return as_double(0x7ff8000000000000ull); }
		if((topy & 0x7ffu) - 0x3beu >= 0x80u){
			if(ix == 0x3ff0000000000000ull){ return 1.0; }
			if((topy & 0x7ffu) < 0x3beu){
				return ix > 0x3ff0000000000000ull ? add_rn(1.0, y) : sub_rn(1.0, y);
			}
			return ((ix > 0x3ff0000000000000ull) == (topy < 0x800u)) ? as_double(0x7ff0000000000000ull) : 0.0;
		}
		if(topx == 0){
			ix = as_u64(mul_rn(x, 4503599627370496.0 /*0x1p52*/));
			ix &= 0x7fffffffffffffffull;
			ix -= 52ull << 52;
		}
	}

	// log_inline
	const double Ln2hi = as_double(kPowLogHdr[0]), Ln2lo = as_double(kPowLogHdr[1]);
	const double A0 = as_double(kPowLogHdr[2]), A1 = as_double(kPowLogHdr[3]), A2 = as_double(kPowLogHdr[4]),
	             A3 = as_double(kPowLogHdr[5]), A4 = as_double(kPowLogHdr[6]), A5 = as_double(kPowLogHdr[7]), A6 = as_double(kPowLogHdr[8]);
	uint64_t tmp = ix - 0x3fe6955500000000ull;
	int i = static_cast<int>((tmp >> 45) & 127);
	int k = static_cast<int>(static_cast<int64_t>(tmp) >> 52);
	uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
	double z = as_double(iz);
	double kd = static_cast<double>(k);
	double invc = as_double(kPowLogTab[4 * i]), logc = as_double(kPowLogTab[4 * i + 2]), logctail = as_double(kPowLogTab[4 * i + 3]);
	double t1 = fma_rn(kd, Ln2hi, logc);
	double lo1 = fma_rn(kd, Ln2lo, logctail);
	double r = fma_rn(z, invc, -1.0);
	double ar = mul_rn(r, A0);
	double p1 = fma_rn(r, A2, A1);
	double p2 = fma_rn(r, A4, A3);
	double t2 = add_rn(r, t1);
	double lo2 = add_rn(sub_rn(t1, t2), r);
	double ar2 = mul_rn(r, ar);
	double ar3 = mul_rn(r, ar2);
	double lo3 = fma_rn(ar, r, -ar2);
	double hi = add_rn(t2, ar2);
	double p3 = fma_rn(r, A6, A5);
	double lo4 = add_rn(sub_rn(t2, hi), ar2);
	double q = fma_rn(p3, ar2, p2);
	double s = fma_rn(ar2, q, p1);
	double lsum = add_rn(add_rn(add_rn(lo1, lo2), lo3), lo4);
	double lo = fma_rn(ar3, s, lsum);
	double lhi = add_rn(hi, lo);
	double llo = add_rn(sub_rn(hi, lhi), lo);

	double ehi = mul_rn(y, lhi);
	double elo = fma_rn(y, llo, fma_rn(lhi, y, -ehi));

	// exp_inline
	uint32_t abstop = static_cast<uint32_t>(as_u64(ehi) >> 52) & 0x7ff;
	if(abstop - 0x3c9u >= 0x3fu){
		if(abstop - 0x3c9u >= 0x80000000u){
			return add_rn(1.0, ehi);
		}
		if(abstop >= 0x409u){
			return (as_u64(ehi) >> 63) ? 0.0 : as_double(0x7ff0000000000000ull);
		}
		abstop = 0;
	}
	return detail::exp_core(ehi, elo, true, abstop, 0);
}

// utilities::InvLogit2 (reference utilities.hpp:505-507): 2/(1+exp(-bias))
RSQ_HD double inv_logit2(double bias){
	return 2.0 / add_rn(1.0, exp_glibc(-bias));
}

}  // namespace rsq
