#include <boost/serialization/access.hpp>
#include <boost/archive/shim_text_archive.hpp>
