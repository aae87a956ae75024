#include <boost/archive/shim_text_archive.hpp>
