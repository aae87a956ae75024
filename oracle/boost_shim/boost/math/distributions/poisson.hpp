// boost::math poisson cdf stand-in (oracle build only; used during stats creation, not simulation).
#ifndef RSQ_SHIM_BOOST_POISSON_HPP
#define RSQ_SHIM_BOOST_POISSON_HPP
#include <cmath>
namespace boost { namespace math {
template<class T=double> class poisson_distribution {
	T mean_;
public:
	explicit poisson_distribution(T mean=1) : mean_(mean) {}
	T mean() const { return mean_; }
};
typedef poisson_distribution<double> poisson;
namespace shim_detail {
	// regularised upper incomplete gamma Q(a, x) for integer a>=1: sum_{i<a} e^-x x^i / i!
	inline double gamma_q_int(double a, double x){
		if(x <= 0) return 1.0;
		double term = std::exp(-x), sum = term;
		for(double i=1; i<a; ++i){ term *= x/i; sum += term; }
		return sum > 1.0 ? 1.0 : sum;
	}
}
template<class T> inline T cdf(const poisson_distribution<T> &d, const T &k){
	if(k < 0) return 0;
	return shim_detail::gamma_q_int(std::floor(k)+1, d.mean());
}
template<class T, class K> inline T cdf(const poisson_distribution<T> &d, const K &k){ return cdf(d, static_cast<T>(k)); }
}}
#endif
