// Which SimBlocks a shard of a run simulates (host only; rsq_engine_prepare and rsq_shard_plan both call this).
//
// The reference has no shards: Simulator::Simulate (Simulator.cpp:2687-2860) walks the 1000-base SimBlocks of all sequences of at least
// MaxInsertLength bases in order and leaves the last 1 + MaxInsertLength / 1000 of them to the look-ahead.  Every block owns its own
// seed, so any partition of the simulated blocks into consecutive ranges reproduces the reference's output when the ranges' outputs
// are concatenated.  The even split is moved onto the first block of a sequence when one starts within 5 % of a shard's size: a shard
// holding only a sliver of a sequence would still need that sequence's whole systematic-error chains.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace rsq {

struct ShardPlan {
	std::vector<uint32_t> seq_first_block, seq_blocks;   // per sequence; 0 blocks = shorter than the longest insert, skipped
	uint32_t blocks_total = 0;                            // all SimBlocks
	uint32_t blocks_simulated = 0;                        // without the look-ahead blocks at the end
	uint32_t shard_count = 1;

	uint64_t boundary(uint32_t k) const {
		if(k == 0){ return 0; }
		if(k >= shard_count){ return blocks_simulated; }
		const uint64_t tol = std::max<uint64_t>(1, blocks_simulated / (20ull * shard_count));
		const uint64_t even = static_cast<uint64_t>(blocks_simulated) * k / shard_count;
		uint64_t best = even, best_d = tol + 1;
		for(size_t i = 0; i < seq_blocks.size(); ++i){
			if(!seq_blocks[i]){ continue; }
			const uint64_t f = seq_first_block[i];
			const uint64_t d = f > even ? f - even : even - f;
			if(f > 0 && f < blocks_simulated && d < best_d){ best = f; best_d = d; }
		}
		return best;
	}
	uint64_t first(uint32_t k) const { return boundary(k); }
	uint64_t count(uint32_t k) const { const uint64_t lo = boundary(k); return std::max<uint64_t>(boundary(k + 1), lo) - lo; }
	bool needs(uint32_t k, size_t seq) const {
		const uint64_t lo = boundary(k), hi = std::max<uint64_t>(boundary(k + 1), lo);
		return seq_blocks[seq] && hi > lo && seq_first_block[seq] < hi && lo < static_cast<uint64_t>(seq_first_block[seq]) + seq_blocks[seq];
	}
};

// lengths: sequence lengths in reference order; insert_to: InsertLengths().to() (sequences shorter than it are not simulated)
inline ShardPlan make_shard_plan(const uint64_t *lengths, size_t n_seqs, uint32_t insert_to, uint32_t shard_count){
	ShardPlan p;
	p.shard_count = shard_count ? shard_count : 1;
	p.seq_first_block.assign(n_seqs, 0); p.seq_blocks.assign(n_seqs, 0);
	for(size_t i = 0; i < n_seqs; ++i){
		if(lengths[i] < insert_to){ continue; }
		p.seq_first_block[i] = p.blocks_total;
		p.seq_blocks[i] = static_cast<uint32_t>((lengths[i] + 999) / 1000);
		p.blocks_total += p.seq_blocks[i];
	}
	const uint32_t lookahead = 1 + insert_to / 1000;
	p.blocks_simulated = p.blocks_total > lookahead ? p.blocks_total - lookahead : 0;
	return p;
}

}  // namespace rsq
