// Device-side gzip of FASTQ text (the step right behind the simulation path).
//
// The text of a batch is cut into members of kMember bytes; one CTA turns one member into a complete gzip member (RFC 1952
// header, ONE dynamic-Huffman deflate block per RFC 1951, CRC-32 + ISIZE trailer), so the members concatenate to a standard
// multi-member gzip stream - the same kind of file the host path (text_io.hpp) writes with zlib.  Phases of a member:
//   parse    a warp per slice of kSlice bytes: every lane looks its 4-byte hash up in the slice's table (the reads of one SimBlock
//            overlap, so the same strand's previous read a few hundred bytes back is the usual hit), also tries distance 1 (runs
//            in quality strings), measures its match; the warp then walks the 32 positions greedily and appends tokens
//   count    token histogram (literal/length and distance alphabets) in shared memory; CRC-32 of 512-byte pieces in parallel
//   codes    one thread: length-limited Huffman code lengths (two-queue construction, Kraft repair), canonical codes
//   emit     bit length of every token, warp scans give bit positions, bits are OR-ed into the zeroed output slot
//   finish   end-of-block code, padding, combined CRC-32 (GF(2) operator for 512 zero bytes), ISIZE, member size
// Everything is written against a small CTA policy (threads, barrier, warp shuffles, atomics) that is instantiated twice:
// DeviceCta (the kernel in engine.cu) and SerialCta (one thread; the CPU test twin tests/host_twin/deflate_check.cpp).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define DFL_HD __host__ __device__ __forceinline__
#else
#define DFL_HD inline
#endif

namespace rsq {
namespace dfl {

constexpr uint32_t kMember = 128u << 10;        // text bytes per gzip member
constexpr uint32_t kSlices = 8;                 // parse slices per member (one warp each on the device)
constexpr uint32_t kSlice = kMember / kSlices;  // 16 KiB: match distances stay far below deflate's 32 KiB window
constexpr uint32_t kHashBits = 12;
constexpr uint32_t kHashSize = 1u << kHashBits;
constexpr uint32_t kMinMatch = 4, kMaxMatch = 258;
constexpr uint32_t kLit = 286, kDist = 30, kMaxBits = 15;
constexpr uint32_t kCrcPiece = 512;
constexpr uint32_t kSlotBytes = 2u * kMember + 256u;   // output slot of a member: <= 15 bits per token + headers, rounded up
constexpr uint32_t kSlotWords = kSlotBytes / 4u;
constexpr uint32_t kHeaderWords = 64;   // gzip header + block header + code length table: 10 B + 1338 bits < 256 B

struct Shared {
	uint32_t freq_lit[288], freq_dist[32];
	uint16_t code_lit[288], code_dist[32];
	uint8_t len_lit[288], len_dist[32];
	uint32_t crc_table[256];
	uint32_t crc_part[kMember / kCrcPiece];
	uint32_t slice_tokens[kSlices];
	uint64_t slice_bits[kSlices], slice_pos[kSlices];
	uint64_t eob_pos;
	// scratch of the code construction (one thread)
	uint16_t h_sym[288]; uint32_t h_w[576]; uint16_t h_parent[576]; uint8_t h_depth[576];
	uint16_t hash[kSlices][kHashSize];
};

#if defined(__CUDACC__)
struct DeviceCta {
	__device__ __forceinline__ uint32_t tid() const { return threadIdx.x; }
	__device__ __forceinline__ uint32_t size() const { return blockDim.x; }
	__device__ __forceinline__ uint32_t lane() const { return threadIdx.x & 31u; }
	__device__ __forceinline__ uint32_t warp() const { return threadIdx.x >> 5; }
	__device__ __forceinline__ uint32_t warps() const { return blockDim.x >> 5; }
	static constexpr uint32_t kLanes = 32;
	__device__ __forceinline__ void sync() const { __syncthreads(); }
	__device__ __forceinline__ void sync_warp() const { __syncwarp(); }
	__device__ __forceinline__ uint32_t shfl(uint32_t v, uint32_t src) const { return __shfl_sync(0xffffffffu, v, src); }
	__device__ __forceinline__ uint32_t shfl_up(uint32_t v, uint32_t d) const { return __shfl_up_sync(0xffffffffu, v, d); }
	__device__ __forceinline__ uint32_t ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
	__device__ __forceinline__ void atomic_add(uint32_t *p, uint32_t v) const { atomicAdd(p, v); }
	__device__ __forceinline__ void atomic_or(uint32_t *p, uint32_t v) const { atomicOr(p, v); }
};
#endif
struct SerialCta {
	uint32_t tid() const { return 0; }
	uint32_t size() const { return 1; }
	uint32_t lane() const { return 0; }
	uint32_t warp() const { return 0; }
	uint32_t warps() const { return 1; }
	static constexpr uint32_t kLanes = 1;
	void sync() const {}
	void sync_warp() const {}
	uint32_t shfl(uint32_t v, uint32_t) const { return v; }
	uint32_t shfl_up(uint32_t v, uint32_t) const { return v; }
	uint32_t ballot(bool p) const { return p ? 1u : 0u; }
	void atomic_add(uint32_t *p, uint32_t v) const { *p += v; }
	void atomic_or(uint32_t *p, uint32_t v) const { *p |= v; }
};

// ---- tokens: literal = byte; match = 1<<31 | (dist-1) << 8 | (len-3) ----
DFL_HD uint32_t token_match(uint32_t len, uint32_t dist){ return 0x80000000u | ((dist - 1u) << 8) | (len - 3u); }

DFL_HD uint32_t floor_log2(uint32_t x){   // x > 0
#if defined(__CUDA_ARCH__)
	return 31u - static_cast<uint32_t>(__clz(static_cast<int>(x)));
#else
	return 31u - static_cast<uint32_t>(__builtin_clz(x));
#endif
}
// length 3..258 -> (symbol, number of extra bits, extra value)   (RFC 1951 3.2.5)
DFL_HD void length_symbol(uint32_t len, uint32_t &sym, uint32_t &nextra, uint32_t &extra){
	const uint32_t l = len - 3u;
	if(l < 8u){ sym = 257u + l; nextra = 0; extra = 0; return; }
	if(len == 258u){ sym = 285u; nextra = 0; extra = 0; return; }
	const uint32_t b = floor_log2(l);
	sym = 257u + 4u * (b - 1u) + ((l >> (b - 2u)) & 3u);
	nextra = b - 2u; extra = l & ((1u << nextra) - 1u);
}
// distance 1..32768 -> (symbol, number of extra bits, extra value)
DFL_HD void distance_symbol(uint32_t dist, uint32_t &sym, uint32_t &nextra, uint32_t &extra){
	const uint32_t x = dist - 1u;
	if(x < 4u){ sym = x; nextra = 0; extra = 0; return; }
	const uint32_t b = floor_log2(x);
	sym = 2u * b + ((x >> (b - 1u)) & 1u);
	nextra = b - 1u; extra = x & ((1u << nextra) - 1u);
}
DFL_HD uint32_t reverse_bits(uint32_t code, uint32_t n){
	uint32_t r = 0;
	for(uint32_t i = 0; i < n; ++i){ r = (r << 1) | ((code >> i) & 1u); }
	return r;
}

// ORs `nbits` (<= 56) bits of `value` into the little-endian bit stream at bit position `pos`
template<class C> DFL_HD void put_bits(const C &c, uint32_t *out, uint64_t pos, uint64_t value, uint32_t nbits){
	if(!nbits){ return; }
	const uint64_t w = pos >> 5; const uint32_t s = static_cast<uint32_t>(pos & 31u);
	c.atomic_or(out + w, static_cast<uint32_t>(value << s));
	if(s + nbits > 32u){
		const uint64_t rest = value >> (32u - s);   // s may be 0: shift by 32 of a 64-bit value is defined
		c.atomic_or(out + w + 1, static_cast<uint32_t>(rest));
		if(s + nbits > 64u){ c.atomic_or(out + w + 2, static_cast<uint32_t>(rest >> 32)); }
	}
}

// ---- CRC-32 (the gzip polynomial, reflected 0xEDB88320) ----
inline void crc_make_table(uint32_t *table){
	for(uint32_t n = 0; n < 256; ++n){
		uint32_t c = n;
		for(int k = 0; k < 8; ++k){ c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1; }
		table[n] = c;
	}
}
DFL_HD uint32_t crc_update(const uint32_t *table, uint32_t crc, const uint8_t *p, uint32_t n){   // crc: finalised value in, finalised value out
	uint32_t c = crc ^ 0xFFFFFFFFu;
	for(uint32_t i = 0; i < n; ++i){ c = table[(c ^ p[i]) & 0xFFu] ^ (c >> 8); }
	return c ^ 0xFFFFFFFFu;
}
DFL_HD uint32_t gf2_times(const uint32_t *mat, uint32_t vec){
	uint32_t sum = 0;
	for(uint32_t i = 0; vec; vec >>= 1, ++i){ if(vec & 1u){ sum ^= mat[i]; } }
	return sum;
}
// operator that advances a finalised CRC over `zero_bytes` zero bytes (a power of two): crc(A || B) = M_|B| crc(A) ^ crc(B)
inline void crc_make_shift_operator(uint32_t *mat, uint32_t zero_bytes){
	uint32_t even[32], odd[32];
	odd[0] = 0xEDB88320u;
	for(uint32_t n = 1, row = 1; n < 32; ++n, row <<= 1){ odd[n] = row; }           // one zero bit
	auto square = [](uint32_t *sq, const uint32_t *m){ for(int n = 0; n < 32; ++n){ sq[n] = gf2_times(m, m[n]); } };
	square(even, odd);   // two zero bits
	square(odd, even);   // four zero bits
	// odd = 4 bits; three more squarings give one byte
	uint32_t *cur = odd, *nxt = even;
	square(nxt, cur); { uint32_t *t = cur; cur = nxt; nxt = t; }   // 8 bits = 1 byte
	for(uint32_t b = 1; b < zero_bytes; b <<= 1){ square(nxt, cur); uint32_t *t = cur; cur = nxt; nxt = t; }
	for(int n = 0; n < 32; ++n){ mat[n] = cur[n]; }
}

// ---- Huffman code lengths, limited to max_bits; one thread ----
DFL_HD void huffman_lengths(Shared &s, const uint32_t *freq, uint32_t n, uint32_t max_bits, uint8_t *len_out){
	uint32_t m = 0;
	for(uint32_t i = 0; i < n; ++i){ len_out[i] = 0; if(freq[i]){ s.h_sym[m] = static_cast<uint16_t>(i); s.h_w[m] = freq[i]; ++m; } }
	if(m == 0){ return; }
	if(m == 1){ len_out[s.h_sym[0]] = 1; return; }
	// insertion sort by weight, ascending (m <= 286)
	for(uint32_t i = 1; i < m; ++i){
		const uint32_t w = s.h_w[i]; const uint16_t sy = s.h_sym[i];
		uint32_t j = i;
		while(j > 0 && s.h_w[j - 1] > w){ s.h_w[j] = s.h_w[j - 1]; s.h_sym[j] = s.h_sym[j - 1]; --j; }
		s.h_w[j] = w; s.h_sym[j] = sy;
	}
	// two queues: leaves [0, m), internal nodes [m, 2m-1) in creation order (their weights are non-decreasing)
	uint32_t leaf = 0, inner = m, made = m;
	for(uint32_t k = 0; k + 1 < m; ++k){
		uint32_t pick[2];
		for(int t = 0; t < 2; ++t){
			if(leaf < m && (inner >= made || s.h_w[leaf] <= s.h_w[inner])){ pick[t] = leaf++; }
			else{ pick[t] = inner++; }
		}
		s.h_w[made] = s.h_w[pick[0]] + s.h_w[pick[1]];
		s.h_parent[pick[0]] = static_cast<uint16_t>(made); s.h_parent[pick[1]] = static_cast<uint16_t>(made);
		++made;
	}
	const uint32_t root = made - 1;
	s.h_depth[root] = 0;
	for(uint32_t i = root; i-- > 0; ){ const uint32_t d = s.h_depth[s.h_parent[i]] + 1u; s.h_depth[i] = static_cast<uint8_t>(d > 60u ? 60u : d); }
	uint32_t num[64];
	for(uint32_t i = 0; i < 64; ++i){ num[i] = 0; }
	for(uint32_t i = 0; i < m; ++i){ ++num[s.h_depth[i]]; }
	// fold lengths above the limit, then repair the Kraft sum (the well-known heuristic of zlib-like encoders)
	for(uint32_t i = max_bits + 1; i < 64; ++i){ num[max_bits] += num[i]; num[i] = 0; }
	uint32_t total = 0;
	for(uint32_t i = max_bits; i > 0; --i){ total += num[i] << (max_bits - i); }
	while(total != (1u << max_bits)){
		--num[max_bits];
		for(uint32_t i = max_bits - 1; i > 0; --i){ if(num[i]){ --num[i]; num[i + 1] += 2; break; } }
		--total;
	}
	// the rarest symbols get the longest codes
	uint32_t idx = 0;
	for(uint32_t l = max_bits; l > 0; --l){ for(uint32_t c = 0; c < num[l]; ++c){ len_out[s.h_sym[idx++]] = static_cast<uint8_t>(l); } }
}
DFL_HD void canonical_codes(const uint8_t *len, uint32_t n, uint16_t *code_out){   // bit-reversed, ready for LSB-first emission
	uint32_t count[16], next[16];
	for(uint32_t i = 0; i < 16; ++i){ count[i] = 0; }
	for(uint32_t i = 0; i < n; ++i){ ++count[len[i]]; }
	count[0] = 0;
	uint32_t code = 0;
	for(uint32_t b = 1; b < 16; ++b){ code = (code + count[b - 1]) << 1; next[b] = code; }
	for(uint32_t i = 0; i < n; ++i){ code_out[i] = len[i] ? static_cast<uint16_t>(reverse_bits(next[len[i]]++, len[i])) : 0; }
}

// bits of one token (value LSB-first, <= 48 bits)
DFL_HD void token_bits(const Shared &s, uint32_t tok, uint64_t &value, uint32_t &nbits){
	if(!(tok & 0x80000000u)){ value = s.code_lit[tok]; nbits = s.len_lit[tok]; return; }
	uint32_t ls, ln, lx, ds, dn, dx;
	length_symbol((tok & 0xFFu) + 3u, ls, ln, lx);
	distance_symbol(((tok >> 8) & 0x7FFFu) + 1u, ds, dn, dx);
	value = s.code_lit[ls]; nbits = s.len_lit[ls];
	value |= static_cast<uint64_t>(lx) << nbits; nbits += ln;
	value |= static_cast<uint64_t>(s.code_dist[ds]) << nbits; nbits += s.len_dist[ds];
	value |= static_cast<uint64_t>(dx) << nbits; nbits += dn;
}

DFL_HD uint32_t load4(const uint8_t *p){ return p[0] | (static_cast<uint32_t>(p[1]) << 8) | (static_cast<uint32_t>(p[2]) << 16) | (static_cast<uint32_t>(p[3]) << 24); }
DFL_HD uint32_t popc(uint32_t x){
#if defined(__CUDA_ARCH__)
	return static_cast<uint32_t>(__popc(x));
#else
	return static_cast<uint32_t>(__builtin_popcount(x));
#endif
}
DFL_HD uint32_t first_bit(uint32_t x){   // x != 0
#if defined(__CUDA_ARCH__)
	return static_cast<uint32_t>(__ffs(static_cast<int>(x))) - 1u;
#else
	return static_cast<uint32_t>(__builtin_ctz(x));
#endif
}
// Unaligned 32-bit little-endian reads from aligned words: a stream keeps the aligned word it stands in and pulls the next one.
// Reads whole aligned words, i.e. up to 3 bytes in front of / behind the range it is asked about (the text buffers are
// allocations of whole words, so these bytes exist).
struct WordStream {
	const uint32_t *w; uint32_t cur, shift;
	DFL_HD explicit WordStream(const uint8_t *p){
		const uintptr_t a = reinterpret_cast<uintptr_t>(p);
		w = reinterpret_cast<const uint32_t *>(a & ~static_cast<uintptr_t>(3)); shift = static_cast<uint32_t>(a & 3u) * 8u; cur = *w;
	}
	DFL_HD uint32_t next(){
		if(!shift){ const uint32_t v = cur; cur = *++w; return v; }   // (reads one word ahead; see above)
		const uint32_t nxt = *++w;
		const uint32_t v = (cur >> shift) | (nxt << (32u - shift));
		cur = nxt;
		return v;
	}
};
DFL_HD uint32_t match_length(const uint8_t *a, const uint8_t *b, uint32_t max_len){
	WordStream sa(a), sb(b);
	uint32_t n = 0;
	while(n < max_len){
		const uint32_t x = sa.next() ^ sb.next();
		if(x){ n += first_bit(x) >> 3; break; }
		n += 4u;
	}
	return n < max_len ? n : max_len;
}

// One member: in[0, n) (1 <= n <= kMember) -> out slot (kSlotWords words); tokens: kMember words of scratch owned by this CTA.
// crc_table / crc_shift: tables of crc_make_table / crc_make_shift_operator(kCrcPiece) in global memory.  Returns the member's
// size in bytes (valid in thread 0).
template<class C> DFL_HD uint32_t deflate_member(const C &c, Shared &s, const uint8_t *in, uint32_t n, uint32_t *out, uint32_t *tokens,
                                                 const uint32_t *crc_table, const uint32_t *crc_shift){
	const uint32_t tid = c.tid(), nt = c.size();
	// ---- reset ----
	for(uint32_t i = tid; i < 288; i += nt){ s.freq_lit[i] = 0; }
	for(uint32_t i = tid; i < 32; i += nt){ s.freq_dist[i] = 0; }
	for(uint32_t i = tid; i < 256; i += nt){ s.crc_table[i] = crc_table[i]; }
	for(uint32_t i = tid; i < kSlices * kHashSize; i += nt){ (&s.hash[0][0])[i] = 0; }
	for(uint32_t i = tid; i < kHeaderWords; i += nt){ out[i] = 0; }   // the rest of the slot is cleared once the member's size is known
	c.sync();
	// ---- parse ----
	const uint32_t n_slices = (n + kSlice - 1) / kSlice;
	for(uint32_t sl = c.warp(); sl < n_slices; sl += c.warps()){
		const uint32_t s0 = sl * kSlice, s1 = (s0 + kSlice < n) ? s0 + kSlice : n;
		uint16_t *table = s.hash[sl];
		uint32_t *tok = tokens + s0;
		uint32_t count = 0;
		const uint32_t lane = c.lane();
		for(uint32_t p = s0; p < s1; ){
			const uint32_t pos = p + lane;
			uint32_t len = 0, dist = 0, h = 0, cand = 0;
			const bool hashed = pos + kMinMatch <= s1;
			if(hashed){ h = (load4(in + pos) * 2654435761u) >> (32u - kHashBits); cand = table[h]; }
			c.sync_warp();
			if(hashed){ table[h] = static_cast<uint16_t>(pos - s0 + 1u); }
			c.sync_warp();
			if(pos < s1){
				const uint32_t max_len = (s1 - pos < kMaxMatch) ? s1 - pos : kMaxMatch;
				if(cand){ const uint32_t cp = s0 + cand - 1u; if(cp < pos){ len = match_length(in + cp, in + pos, max_len); dist = pos - cp; } }
				if(pos > s0 && len < max_len && in[pos - 1] == in[pos]){ const uint32_t l1 = match_length(in + pos - 1, in + pos, max_len); if(l1 > len){ len = l1; dist = 1; } }
				if(len < kMinMatch){ len = 0; }
			}
			// greedy walk over the window, from match to match (everything between two taken matches is a literal);
			// every lane follows the same path, the taken positions then write their tokens side by side
			const uint32_t in_window = (s1 - p < C::kLanes) ? s1 - p : C::kLanes;
			const uint32_t valid = in_window >= 32u ? 0xffffffffu : (1u << in_window) - 1u;
			const uint32_t with_match = c.ballot(len != 0);
			uint32_t taken = 0, cur = 0;
			while(cur < in_window){
				const uint32_t rest = with_match & ~((1u << cur) - 1u);
				if(!rest){ taken |= valid & ~((1u << cur) - 1u); cur = in_window; break; }
				const uint32_t m = first_bit(rest);
				taken |= (m >= 31u ? 0xffffffffu : (1u << (m + 1u)) - 1u) & ~((1u << cur) - 1u);
				cur = m + c.shfl(len, m);
			}
			if((taken >> lane) & 1u){ tok[count + popc(taken & ((1u << lane) - 1u))] = len ? token_match(len, dist) : in[pos]; }
			count += popc(taken);
			p += cur;
		}
		if(c.lane() == 0){ s.slice_tokens[sl] = count; }
	}
	c.sync();
	// ---- count: token histogram, CRC-32 of the pieces ----
	for(uint32_t sl = 0; sl < n_slices; ++sl){
		const uint32_t *tok = tokens + sl * kSlice; const uint32_t count = s.slice_tokens[sl];
		for(uint32_t i = tid; i < count; i += nt){
			const uint32_t t = tok[i];
			if(!(t & 0x80000000u)){ c.atomic_add(&s.freq_lit[t], 1u); }
			else{
				uint32_t sy, ne, ex;
				length_symbol((t & 0xFFu) + 3u, sy, ne, ex); c.atomic_add(&s.freq_lit[sy], 1u);
				distance_symbol(((t >> 8) & 0x7FFFu) + 1u, sy, ne, ex); c.atomic_add(&s.freq_dist[sy], 1u);
			}
		}
	}
	const uint32_t full_pieces = n / kCrcPiece;
	for(uint32_t i = tid; i < full_pieces; i += nt){ s.crc_part[i] = crc_update(s.crc_table, 0u, in + i * kCrcPiece, kCrcPiece); }
	c.sync();
	// ---- codes, headers (one thread) ----
	if(tid == 0){
		s.freq_lit[256] = 1;   // end of block
		huffman_lengths(s, s.freq_lit, kLit, kMaxBits, s.len_lit);
		huffman_lengths(s, s.freq_dist, kDist, kMaxBits, s.len_dist);
		canonical_codes(s.len_lit, kLit, s.code_lit);
		canonical_codes(s.len_dist, kDist, s.code_dist);
		uint64_t pos = 0;
		const uint8_t gz[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 0xff};   // deflate, no flags, no mtime, unknown OS
		for(int i = 0; i < 10; ++i){ put_bits(c, out, pos, gz[i], 8); pos += 8; }
		put_bits(c, out, pos, 1u | (2u << 1), 3); pos += 3;                       // BFINAL = 1, BTYPE = dynamic
		put_bits(c, out, pos, kLit - 257u, 5); pos += 5;                          // HLIT
		put_bits(c, out, pos, kDist - 1u, 5); pos += 5;                           // HDIST
		put_bits(c, out, pos, 19u - 4u, 4); pos += 4;                             // HCLEN: all 19 code length codes
		// code length alphabet: symbols 0..15 with 4 bits each (a complete code), 16/17/18 (repeats) unused
		const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
		for(int i = 0; i < 19; ++i){ put_bits(c, out, pos, order[i] < 16 ? 4u : 0u, 3); pos += 3; }
		for(uint32_t i = 0; i < kLit; ++i){ put_bits(c, out, pos, reverse_bits(s.len_lit[i], 4), 4); pos += 4; }
		for(uint32_t i = 0; i < kDist; ++i){ put_bits(c, out, pos, reverse_bits(s.len_dist[i], 4), 4); pos += 4; }
		s.eob_pos = pos;   // first data bit, for now
	}
	c.sync();
	// ---- bit lengths of the slices ----
	for(uint32_t sl = c.warp(); sl < n_slices; sl += c.warps()){
		const uint32_t *tok = tokens + sl * kSlice; const uint32_t count = s.slice_tokens[sl];
		uint64_t bits = 0;
		for(uint32_t i = c.lane(); i < count; i += C::kLanes){ uint64_t v; uint32_t nb; token_bits(s, tok[i], v, nb); bits += nb; }
		uint32_t lo = static_cast<uint32_t>(bits);   // < 2^32: a slice has at most 16 Ki tokens of <= 48 bits
		for(uint32_t d = 1; d < C::kLanes; d <<= 1){ const uint32_t o = c.shfl_up(lo, d); if(c.lane() >= d){ lo += o; } }
		const uint32_t sum = c.shfl(lo, C::kLanes - 1u);
		if(c.lane() == 0){ s.slice_bits[sl] = sum; }
	}
	c.sync();
	if(tid == 0){
		uint64_t pos = s.eob_pos;
		for(uint32_t sl = 0; sl < n_slices; ++sl){ s.slice_pos[sl] = pos; pos += s.slice_bits[sl]; }
		s.eob_pos = pos;
	}
	c.sync();
	{	// data bits + end of block (<= 15) + padding + trailer (64) + the two words put_bits may touch behind its position
		const uint32_t end_word = static_cast<uint32_t>((s.eob_pos + 15u + 7u + 64u) >> 5) + 3u;
		for(uint32_t i = kHeaderWords + tid; i < end_word && i < kSlotWords; i += nt){ out[i] = 0; }
	}
	c.sync();
	// ---- emit ----
	for(uint32_t sl = c.warp(); sl < n_slices; sl += c.warps()){
		const uint32_t *tok = tokens + sl * kSlice; const uint32_t count = s.slice_tokens[sl];
		uint64_t base = s.slice_pos[sl];
		for(uint32_t i0 = 0; i0 < count; i0 += C::kLanes){
			const uint32_t i = i0 + c.lane();
			uint64_t v = 0; uint32_t nb = 0;
			if(i < count){ token_bits(s, tok[i], v, nb); }
			uint32_t incl = nb;
			for(uint32_t d = 1; d < C::kLanes; d <<= 1){ const uint32_t o = c.shfl_up(incl, d); if(c.lane() >= d){ incl += o; } }
			put_bits(c, out, base + incl - nb, v, nb);
			base += c.shfl(incl, C::kLanes - 1u);
		}
	}
	c.sync();
	// ---- finish ----
	uint32_t member_bytes = 0;
	if(tid == 0){
		uint64_t pos = s.eob_pos;
		put_bits(c, out, pos, s.code_lit[256], s.len_lit[256]); pos += s.len_lit[256];
		pos = (pos + 7u) & ~static_cast<uint64_t>(7u);
		uint32_t crc = 0;
		for(uint32_t i = 0; i < full_pieces; ++i){ crc = gf2_times(crc_shift, crc) ^ s.crc_part[i]; }
		crc = crc_update(s.crc_table, crc, in + full_pieces * kCrcPiece, n - full_pieces * kCrcPiece);
		put_bits(c, out, pos, crc, 32); pos += 32;
		put_bits(c, out, pos, n, 32); pos += 32;
		member_bytes = static_cast<uint32_t>(pos >> 3);
	}
	c.sync();
	return member_bytes;
}

}  // namespace dfl
}  // namespace rsq
