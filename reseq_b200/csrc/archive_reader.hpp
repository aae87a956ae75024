// Loader for the reference's own profile files (X.reseq + X.reseq.ipf, Boost text archives).
#pragma once
#include "host_profile.hpp"
namespace rsq {
inline void load_reseq_profile(Profile &, const char *, const char *){
	throw std::runtime_error("loading .reseq/.ipf archives is not implemented yet; use a flat profile");
}
}
