// boost::filesystem stand-in over POSIX calls (oracle build only; C++14, no std::filesystem).
#ifndef RSQ_SHIM_BOOST_FILESYSTEM_HPP
#define RSQ_SHIM_BOOST_FILESYSTEM_HPP
#include <string>
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>
namespace boost { namespace filesystem {
class path {
	std::string p_;
public:
	path() {}
	path(const std::string &p) : p_(p) {}
	path(const char *p) : p_(p) {}
	const std::string &string() const { return p_; }
	const char *c_str() const { return p_.c_str(); }
	bool empty() const { return p_.empty(); }
	path parent_path() const {
		auto pos = p_.find_last_of('/');
		if(std::string::npos == pos){ return path(); }
		while(pos > 1 && '/' == p_[pos-1]){ --pos; }
		return path(p_.substr(0, 0 == pos ? 1 : pos));
	}
};
inline bool exists(const path &p){ struct stat st; return 0 == ::stat(p.c_str(), &st); }
inline bool create_directories(const path &p){
	if(p.empty() || exists(p)){ return false; }
	create_directories(p.parent_path());
	return 0 == ::mkdir(p.c_str(), 0777);
}
inline bool remove(const path &p){
	if(!exists(p)){ return false; }
	if(0 == ::unlink(p.c_str())){ return true; }
	return 0 == ::rmdir(p.c_str());
}
}}
#endif
