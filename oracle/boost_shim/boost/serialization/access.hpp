// Minimal stand-in for Boost.Serialization (test infrastructure only; see oracle/README.md).
// Written from the published Boost text-archive token rules; Boost itself is not available offline.
#ifndef RSQ_SHIM_BOOST_SERIALIZATION_ACCESS_HPP
#define RSQ_SHIM_BOOST_SERIALIZATION_ACCESS_HPP
namespace boost { namespace serialization {
class access {
public:
	template<class Archive, class T> static void serialize(Archive &ar, T &t, const unsigned int version){
		t.serialize(ar, version);
	}
};
}}
#endif
