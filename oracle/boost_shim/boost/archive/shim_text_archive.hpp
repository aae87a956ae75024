// Minimal text archive in the Boost.Serialization text_{i,o}archive token dialect
// (test infrastructure for the oracle build; not product code).
//
// Token rules implemented (Boost 1.6x/1.7x behaviour, archive library version 17):
//   * header: "22 serialization::archive 17"
//   * every token is separated by one blank
//   * class types (anything with a serialize() member, std::pair, std::array, std::map and
//     std::vector<T> for non-arithmetic T) emit "<tracking> <version>" == "0 0" the first time
//     the type is met and nothing afterwards
//   * std::vector<T>: count, item_version (0), items; std::vector<bool>: count, item_version, items
//   * std::array<T,N>: N, items
//   * std::pair: first, second;  std::map: count, item_version, pairs
//   * std::string: length, blank, raw characters
//   * 8-bit integers are written as numbers, bool as 0/1, floating point with 17 significant
//     digits in scientific notation
#ifndef RSQ_SHIM_TEXT_ARCHIVE_HPP
#define RSQ_SHIM_TEXT_ARCHIVE_HPP

#include <array>
#include <cstdint>
#include <iomanip>
#include <istream>
#include <limits>
#include <map>
#include <ostream>
#include <set>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <typeindex>
#include <utility>
#include <vector>

#include <boost/serialization/access.hpp>

namespace boost { namespace archive {

class archive_exception : public std::runtime_error {
public:
	explicit archive_exception(const std::string &what) : std::runtime_error(what) {}
};

namespace shim_detail {
	template<class T> struct is_arith_vector : std::false_type {};
	template<class T, class A> struct is_arith_vector<std::vector<T, A>> : std::integral_constant<bool, std::is_arithmetic<T>::value> {};
}

class text_oarchive {
	std::ostream &os_;
	std::set<std::type_index> seen_;
	bool first_token_ = true;

	void sep(){ if(first_token_){ first_token_ = false; } else { os_.put(' '); } }
	template<class T> void class_info(){
		if(seen_.insert(std::type_index(typeid(T))).second){ sep(); os_ << 0; sep(); os_ << 0; }
	}

	template<class T> typename std::enable_if<std::is_integral<T>::value && sizeof(T)==1 && !std::is_same<T,bool>::value>::type put(const T &v){ sep(); os_ << static_cast<int>(v); }
	void put(const bool &v){ sep(); os_ << (v ? 1 : 0); }
	template<class T> typename std::enable_if<std::is_integral<T>::value && (sizeof(T)>1)>::type put(const T &v){ sep(); os_ << v; }
	template<class T> typename std::enable_if<std::is_floating_point<T>::value>::type put(const T &v){
		sep(); os_ << std::setprecision(std::numeric_limits<T>::digits10 + 2) << std::scientific << v;
	}
	void put(const std::string &s){ sep(); os_ << s.size(); os_.put(' '); os_.write(s.data(), s.size()); }

	template<class T, class A> void put(const std::vector<T, A> &v){
		if(!shim_detail::is_arith_vector<std::vector<T, A>>::value){ class_info<std::vector<T, A>>(); }
		put(static_cast<std::size_t>(v.size()));
		put(static_cast<unsigned int>(0));
		for(const auto &e : v){ put(e); }
	}
	template<class A> void put(const std::vector<bool, A> &v){
		put(static_cast<std::size_t>(v.size()));
		put(static_cast<unsigned int>(0));
		for(bool e : v){ put(e); }
	}
	template<class T, std::size_t N> void put(const std::array<T, N> &a){
		class_info<std::array<T, N>>();
		put(static_cast<std::size_t>(N));
		for(const auto &e : a){ put(e); }
	}
	template<class F, class S> void put(const std::pair<F, S> &p){
		class_info<std::pair<F, S>>();
		put(p.first); put(p.second);
	}
	template<class K, class V, class C, class A> void put(const std::map<K, V, C, A> &m){
		class_info<std::map<K, V, C, A>>();
		put(static_cast<std::size_t>(m.size()));
		put(static_cast<unsigned int>(0));
		for(const auto &e : m){ put(e); }
	}
	template<class T> typename std::enable_if<std::is_class<T>::value>::type put(const T &t){
		class_info<T>();
		boost::serialization::access::serialize(*this, const_cast<T &>(t), 0);
	}

public:
	typedef std::false_type is_loading;
	typedef std::true_type is_saving;

	explicit text_oarchive(std::ostream &os) : os_(os) {
		os_ << "22 serialization::archive 17";
		first_token_ = false;
	}
	template<class T> text_oarchive &operator<<(const T &t){ put(t); return *this; }
	template<class T> text_oarchive &operator&(const T &t){ put(t); return *this; }
};

class text_iarchive {
	std::istream &is_;
	std::set<std::type_index> seen_;

	void fail(const char *what){ throw archive_exception(std::string("input stream error: ") + what); }
	template<class T> void class_info(){
		if(seen_.insert(std::type_index(typeid(T))).second){
			unsigned int tracking, version;
			if(!(is_ >> tracking >> version)){ fail("class info"); }
		}
	}

	template<class T> typename std::enable_if<std::is_integral<T>::value && sizeof(T)==1 && !std::is_same<T,bool>::value>::type get(T &v){ int tmp; if(!(is_ >> tmp)){ fail("int8"); } v = static_cast<T>(tmp); }
	void get(bool &v){ int tmp; if(!(is_ >> tmp)){ fail("bool"); } v = (tmp != 0); }
	template<class T> typename std::enable_if<std::is_integral<T>::value && (sizeof(T)>1)>::type get(T &v){ if(!(is_ >> v)){ fail("integer"); } }
	template<class T> typename std::enable_if<std::is_floating_point<T>::value>::type get(T &v){
		// operator>> does not accept inf/nan; go through strtod
		std::string tok; if(!(is_ >> tok)){ fail("float"); }
		v = static_cast<T>(std::strtod(tok.c_str(), nullptr));
	}
	void get(std::string &s){
		std::size_t n; if(!(is_ >> n)){ fail("string length"); }
		is_.get(); // separating blank
		s.resize(n);
		if(n && !is_.read(&s[0], n)){ fail("string"); }
	}

	template<class T, class A> void get(std::vector<T, A> &v){
		if(!shim_detail::is_arith_vector<std::vector<T, A>>::value){ class_info<std::vector<T, A>>(); }
		std::size_t n; unsigned int item_version;
		get(n); get(item_version);
		v.clear(); v.resize(n);
		for(auto &e : v){ get(e); }
	}
	template<class A> void get(std::vector<bool, A> &v){
		std::size_t n; unsigned int item_version;
		get(n); get(item_version);
		v.clear(); v.reserve(n);
		for(std::size_t i=0; i<n; ++i){ bool b; get(b); v.push_back(b); }
	}
	template<class T, std::size_t N> void get(std::array<T, N> &a){
		class_info<std::array<T, N>>();
		std::size_t n; get(n);
		if(n != N){ fail("array size mismatch"); }
		for(auto &e : a){ get(e); }
	}
	template<class F, class S> void get(std::pair<F, S> &p){
		class_info<std::pair<F, S>>();
		get(const_cast<typename std::remove_const<F>::type &>(p.first)); get(p.second);
	}
	template<class K, class V, class C, class A> void get(std::map<K, V, C, A> &m){
		class_info<std::map<K, V, C, A>>();
		std::size_t n; unsigned int item_version;
		get(n); get(item_version);
		m.clear();
		for(std::size_t i=0; i<n; ++i){ std::pair<K, V> e; get(e); m.insert(m.end(), e); }
	}
	template<class T> typename std::enable_if<std::is_class<T>::value>::type get(T &t){
		class_info<T>();
		boost::serialization::access::serialize(*this, t, 0);
	}

public:
	typedef std::true_type is_loading;
	typedef std::false_type is_saving;

	explicit text_iarchive(std::istream &is) : is_(is) {
		std::size_t siglen; std::string sig; unsigned int libver;
		if(!(is_ >> siglen >> sig >> libver) || sig != "serialization::archive"){
			throw archive_exception("invalid signature");
		}
	}
	template<class T> text_iarchive &operator>>(T &t){ get(t); return *this; }
	template<class T> text_iarchive &operator&(T &t){ get(t); return *this; }
};

}}
#endif
