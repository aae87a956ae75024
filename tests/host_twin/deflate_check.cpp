// TEST HARNESS: reseq_b200/csrc/deflate_core.cuh instantiated with one thread (SerialCta) - the same member code the kernel
// k_deflate_members runs.   deflate_check <in> <out.gz>   writes <in> as a multi-member gzip file; the test inflates it with zlib.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <memory>
#include <string>
#include <vector>
#include "../../reseq_b200/csrc/deflate_core.cuh"

int main(int argc, char **argv){
	if(argc != 3){ fprintf(stderr, "usage: deflate_check <in> <out.gz>\n"); return 64; }
	using namespace rsq::dfl;
	std::ifstream f(argv[1], std::ios::binary);
	const std::string file_text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
	// the member code reads whole aligned words (WordStream): keep the text in a word buffer with slack on both sides, at an
	// odd offset so that unaligned streams are exercised
	std::vector<uint32_t> backing(file_text.size() / 4 + 8);
	char *text_ptr = reinterpret_cast<char *>(backing.data()) + 5;
	std::memcpy(text_ptr, file_text.data(), file_text.size());
	struct View { const char *p; size_t n; const char *data() const { return p; } size_t size() const { return n; } } text{text_ptr, file_text.size()};
	std::vector<uint32_t> crc_table(256), crc_shift(32);
	crc_make_table(crc_table.data());
	crc_make_shift_operator(crc_shift.data(), kCrcPiece);
	std::unique_ptr<Shared> sh(new Shared);
	std::vector<uint32_t> tokens(kMember), slot(kSlotWords);
	FILE *o = fopen(argv[2], "wb");
	if(!o){ return 1; }
	size_t total = 0, n_tokens = 0;
	SerialCta cta;
	for(size_t off = 0; off < text.size(); off += kMember){
		const uint32_t n = static_cast<uint32_t>(std::min<size_t>(kMember, text.size() - off));
		std::memset(slot.data(), 0xAB, slot.size() * 4);   // the member code has to clear what it uses
		const uint32_t bytes = deflate_member(cta, *sh, reinterpret_cast<const uint8_t *>(text.data()) + off, n, slot.data(), tokens.data(), crc_table.data(), crc_shift.data());
		if(bytes > kSlotBytes){ fprintf(stderr, "member overflows its slot\n"); return 2; }
		fwrite(slot.data(), 1, bytes, o);
		total += bytes;
		for(uint32_t sl = 0; sl < (n + kSlice - 1) / kSlice; ++sl){ n_tokens += sh->slice_tokens[sl]; }
	}
	fclose(o);
	printf("text=%zu file=%zu ratio=%.3f tokens=%zu\n", text.size(), total, text.size() ? double(total) / text.size() : 0.0, n_tokens);
	return 0;
}
