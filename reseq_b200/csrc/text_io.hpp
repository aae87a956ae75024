// Host text I/O either side of the simulation path: the files the reference opens through SeqAn (SeqFileOut / SeqFileIn /
// BedFileIn: Simulator.cpp:137, 2543, 2580; Reference.cpp:774, 1132) may be gzip-compressed.  SeqAn picks the *output*
// compression from the file name (".gz" -> gzip, ".bgzf" -> BGZF, ".bz2" -> bzip2; seqan/stream/stream_base.h:120-161,
// virtual_stream.h:493-507) and recognises compressed *input* by its magic bytes.  Here:
//   * TextSink   - output file; ".gz" / ".bgzf" names are written as a sequence of independent gzip members (a valid gzip stream
//                  that `gzip -d`, zlib's gzread and SeqAn's own reader all decode to the identical text), the members of one
//                  write() compressed in parallel on the host cores; ".bz2" is rejected loudly.
//   * open_text  - input stream; gzip input (magic 1f 8b, any number of members) is inflated on the fly.
// Host C++ only (zlib); no device code.
#pragma once
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <istream>
#include <memory>
#include <stdexcept>
#include <streambuf>
#include <string>
#include <thread>
#include <vector>

namespace rsq {

inline bool name_ends_with(const std::string &name, const char *suffix){
	const size_t n = std::strlen(suffix);
	if(name.size() < n){ return false; }
	for(size_t i = 0; i < n; ++i){
		if(std::tolower(static_cast<unsigned char>(name[name.size() - n + i])) != suffix[i]){ return false; }
	}
	return true;
}

class TextSink {
public:
	static constexpr size_t kMemberBytes = 1u << 20;   // text per gzip member: the unit of parallel compression

	TextSink() = default;
	TextSink(const TextSink &) = delete;
	TextSink &operator=(const TextSink &) = delete;
	~TextSink(){ if(f_){ std::fclose(f_); } }

	// Creates / truncates the file (append = true: continues it; appended gzip members continue a gzip stream).
	// Returns false when it cannot be opened (the caller words the message like the reference).
	bool open(const std::string &path, bool append = false){
		if(name_ends_with(path, ".bz2")){ throw std::runtime_error("bzip2-compressed output ('" + path + "') is not supported by this engine; use a .gz or an uncompressed file name"); }
		gz_ = name_ends_with(path, ".gz") || name_ends_with(path, ".bgzf");
		path_ = path; members_ = 0; text_bytes_ = 0; file_bytes_ = 0;
		level_ = Z_DEFAULT_COMPRESSION;   // what SeqAn's zip_stream uses (seqan/stream/iostream_zip.h:285)
		if(const char *env = std::getenv("RSQ_GZIP_LEVEL")){ level_ = std::max(0, std::min(9, std::atoi(env))); }
		threads_ = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
		if(const char *env = std::getenv("RSQ_GZIP_THREADS")){ threads_ = static_cast<unsigned>(std::max(1, std::atoi(env))); }
		f_ = std::fopen(path.c_str(), append ? "ab" : "wb");
		if(f_ && append && std::ftell(f_) > 0){ members_ = 1; }   // no empty member behind existing content
		return f_ != nullptr;
	}
	bool is_open() const { return f_ != nullptr; }
	bool compressed() const { return gz_; }
	uint64_t text_bytes() const { return text_bytes_; }
	uint64_t file_bytes() const { return file_bytes_; }
	const std::string &path() const { return path_; }

	// Appends n bytes of text; false on a write error.
	bool write(const void *data, size_t n){
		if(!f_){ return false; }
		text_bytes_ += n;
		if(!gz_){ file_bytes_ += n; return std::fwrite(data, 1, n, f_) == n; }
		const unsigned char *p = static_cast<const unsigned char *>(data);
		const size_t n_members = (n + kMemberBytes - 1) / kMemberBytes;
		// waves of `threads_` members: each worker deflates whole members, the wave is then written in order
		std::vector<std::vector<unsigned char>> out(std::min<size_t>(n_members, 4 * threads_));
		for(size_t first = 0; first < n_members; first += out.size()){
			const size_t count = std::min(out.size(), n_members - first);
			std::atomic<size_t> next{0};
			std::atomic<bool> ok{true};
			auto work = [&]{
				for(size_t k; (k = next.fetch_add(1)) < count; ){
					const size_t off = (first + k) * kMemberBytes;
					if(!deflate_member(p + off, std::min(kMemberBytes, n - off), out[k])){ ok = false; }
				}
			};
			const unsigned n_threads = static_cast<unsigned>(std::min<size_t>(threads_, count));
			std::vector<std::thread> pool;
			for(unsigned t = 1; t < n_threads; ++t){ pool.emplace_back(work); }
			work();
			for(auto &t : pool){ t.join(); }
			if(!ok){ return false; }
			for(size_t k = 0; k < count; ++k){
				if(std::fwrite(out[k].data(), 1, out[k].size(), f_) != out[k].size()){ return false; }
				file_bytes_ += out[k].size(); ++members_;
			}
		}
		return true;
	}

	// Appends gzip members that were made elsewhere (the device deflate kernels): `n` bytes of file data standing for `text` bytes.
	bool write_members(const void *data, size_t n, uint64_t text, uint64_t members){
		if(!f_ || !gz_){ return false; }
		text_bytes_ += text; file_bytes_ += n; members_ += members;
		return std::fwrite(data, 1, n, f_) == n;
	}

	// Flushes and closes; false on a write error.  An empty compressed file still gets one (empty) gzip member.
	bool close(){
		if(!f_){ return true; }
		bool ok = true;
		if(gz_ && !members_){
			std::vector<unsigned char> m;
			ok = deflate_member(nullptr, 0, m) && std::fwrite(m.data(), 1, m.size(), f_) == m.size();
			file_bytes_ += m.size();
		}
		if(std::fclose(f_)){ ok = false; }
		f_ = nullptr;
		return ok;
	}

private:
	bool deflate_member(const unsigned char *src, size_t n, std::vector<unsigned char> &dst) const {
		z_stream z; std::memset(&z, 0, sizeof z);
		if(deflateInit2(&z, level_, Z_DEFLATED, 15 + 16 /* gzip wrapper */, 8, Z_DEFAULT_STRATEGY) != Z_OK){ return false; }
		dst.resize(deflateBound(&z, n) + 32);
		static const unsigned char none = 0;
		z.next_in = const_cast<unsigned char *>(n ? src : &none); z.avail_in = static_cast<uInt>(n);
		z.next_out = dst.data(); z.avail_out = static_cast<uInt>(dst.size());
		const int rc = deflate(&z, Z_FINISH);
		const size_t produced = dst.size() - z.avail_out;
		deflateEnd(&z);
		if(rc != Z_STREAM_END){ return false; }
		dst.resize(produced);
		return true;
	}

	FILE *f_ = nullptr;
	bool gz_ = false;
	int level_ = Z_DEFAULT_COMPRESSION;
	unsigned threads_ = 1;
	uint64_t members_ = 0, text_bytes_ = 0, file_bytes_ = 0;
	std::string path_;
};

// streambuf over zlib's gzFile (reads plain files transparently as well, but those take the ifstream route below)
class GzInBuf : public std::streambuf {
public:
	explicit GzInBuf(const std::string &path) : buf_(1u << 20) {
		f_ = gzopen(path.c_str(), "rb");
		if(f_){ gzbuffer(f_, 1u << 18); }
	}
	~GzInBuf() override { if(f_){ gzclose(f_); } }
	bool is_open() const { return f_ != nullptr; }
	bool failed() const { return failed_; }
protected:
	int_type underflow() override {
		if(gptr() < egptr()){ return traits_type::to_int_type(*gptr()); }
		if(!f_){ return traits_type::eof(); }
		const int n = gzread(f_, buf_.data(), static_cast<unsigned>(buf_.size()));
		if(n < 0){ failed_ = true; return traits_type::eof(); }
		if(n == 0){ int err = 0; gzerror(f_, &err); if(err != Z_OK && err != Z_STREAM_END){ failed_ = true; } return traits_type::eof(); }
		setg(buf_.data(), buf_.data(), buf_.data() + n);
		return traits_type::to_int_type(*gptr());
	}
private:
	gzFile f_ = nullptr;
	std::vector<char> buf_;
	bool failed_ = false;
};

class TextInput {   // owns the buffer behind the stream it hands out
public:
	// binary = false keeps the text-mode default of the std::ifstream it replaces
	explicit TextInput(const std::string &path){
		unsigned char magic[2] = {0, 0};
		if(FILE *probe = std::fopen(path.c_str(), "rb")){
			const size_t got = std::fread(magic, 1, 2, probe);
			std::fclose(probe);
			if(got == 2 && magic[0] == 0x1f && magic[1] == 0x8b){
				gz_.reset(new GzInBuf(path));
				if(gz_->is_open()){ stream_.reset(new std::istream(gz_.get())); }
				return;
			}
		}
		else{ return; }
		std::unique_ptr<std::ifstream> f(new std::ifstream(path));
		if(f->is_open()){ stream_ = std::move(f); }
	}
	bool is_open() const { return stream_ != nullptr; }
	bool compressed() const { return gz_ != nullptr; }
	// true when the decompressor met a corrupt or truncated stream (the text read so far must not be trusted)
	bool corrupt() const { return gz_ && gz_->failed(); }
	std::istream &stream(){ return *stream_; }
private:
	std::unique_ptr<GzInBuf> gz_;
	std::unique_ptr<std::istream> stream_;
};

}  // namespace rsq
