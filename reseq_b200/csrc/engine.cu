// CUDA engine behind include/reseq_b200.h: kernels (sm_100a) + host orchestration of the Simulate() path.
//
// Device data layout (all in HBM, per engine):
//   tables      TableDesc[] + FP64 blob + par0 (every LogArrayResult, ~1-7 MB: L2 resident)
//   reference   1 byte / base (Dna codes after ReplaceN), concatenated sequences
//   gc_prefix   u32 / base (+1 per sequence)          -> fragment GC in O(1)
//   sur_start, sur_end  f64 / base                    -> SurroundingBias::Bias per fragment end (multi-batch runs: one batch's window, see simulate())
//   sys_fwd, sys_rev    2 bytes / base / strand       -> (dominant error, rate) of SetSystematicErrors
//   master      raw mt19937_64 outputs of one SimUnit (2*blocks + 4*L words), reused per sequence
//   blocks      BlockDesc[] (seed, ref, start, id, first methylation region, first variant)
//   variants    runs with -V: flattened Reference::variants_, per-block first variant, systematic errors + context bytes of replacement bases
//   per batch of SimBlocks (speculative path, spec_core.cuh): snapshots, read jobs, stream slices, record slots in slabs of 32
//               chained per block; the FASTQ text of a batch is gathered in block order into one of two device buffers and
//               pulled to the host by the writer threads (ChunkWriter) while the next batch is simulated
//   arena       serial path only: fixed-size chunks of FASTQ text, per (block, segment) chains
// Kernels: k_surroundings, k_gc_*, k_bias_chunks/_scan/_resolve (k_sum_bias: chain form), k_master_seed/_stream/_jump_gen/_jump_xor, k_sys_chunks_lanes (k_sys_chunks) + k_sys_check,
//          k_build_blocks, k_var_sys_errors, k_spec_init/_scan<kVar>/_reads<kVar>/_block_out/_gather (product path), k_deflate_* (gzip output), k_simulate + k_adapter_only + k_gather and
//          k_error_model (serial forms: cross-check and fallback), k_block_offsets.
#include <cuda_runtime.h>
#include <type_traits>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <sys/mman.h>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <memory>
#include <string>
#include <vector>
#include "../../include/reseq_b200.h"
#include "host_profile.hpp"
#include "sim_core.cuh"
#include "spec_core.cuh"
#include "bias_core.cuh"
#include "archive_reader.hpp"
#include "deflate_core.cuh"
#include "em_input.hpp"
#include "variant_syserr.hpp"
#include "shard_plan.hpp"
#include "ordered_sum.cuh"

#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only: the library is bound at run time (libnccl.so.2), so single-GPU use does not need it

namespace rsq {

// NCCL entry points used by the multi-GPU data plane (engines of one run joined into a group: rsq_engine_join_group)
struct NcclApi {
	decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
	decltype(&ncclCommInitRank) CommInitRank = nullptr;
	decltype(&ncclCommDestroy) CommDestroy = nullptr;
	decltype(&ncclAllReduce) AllReduce = nullptr;
	decltype(&ncclBroadcast) Broadcast = nullptr;
	decltype(&ncclGetErrorString) GetErrorString = nullptr;
	bool ok = false;
};
static NcclApi &nccl_api(){
	static NcclApi api;
	static std::once_flag once;
	std::call_once(once, []{
		void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);   // the copy a host program (e.g. PyTorch) already loaded is reused
		if(!h){ h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL); }
		if(!h){ return; }
		api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
		api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
		api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
		api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
		api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(dlsym(h, "ncclBroadcast"));
		api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
		api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.Broadcast && api.GetErrorString;
	});
	return api;
}
#define RSQ_NCCL(call) do{ ncclResult_t r_ = (call); if(r_ != ncclSuccess){ throw std::runtime_error(std::string(#call) + ": " + nccl_api().GetErrorString(r_)); } }while(0)

static thread_local std::string g_last_error;
static void set_error(const char *fmt, ...){
	char buf[2048];
	va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
	g_last_error = buf;
}

// RSQ_TIMING=1: wall-clock stage log on stderr (host-side costs are invisible to the CUDA-event times of the report)
static void stage_log(const char *what){
	static const bool on = getenv("RSQ_TIMING") != nullptr;
	if(!on){ return; }
	static auto last = std::chrono::steady_clock::now();
	const auto now = std::chrono::steady_clock::now();
	fprintf(stderr, "[rsq timing] %-40s +%.3f s\n", what, std::chrono::duration<double>(now - last).count());
	last = now;
}

#define RSQ_CUDA(call) do{ cudaError_t e_ = (call); if(e_ != cudaSuccess){ throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e_)); } }while(0)

template<class T> struct DevBuf {
	T *p = nullptr; size_t n = 0, cap = 0;
	DevBuf() = default;
	DevBuf(const DevBuf &) = delete; DevBuf &operator=(const DevBuf &) = delete;
	~DevBuf(){ release(); }
	void release(){ if(p){ cudaFree(p); p = nullptr; n = 0; cap = 0; } }
	// Buffers are recycled between runs: cudaMalloc/cudaFree synchronise the device and cost milliseconds each.
	void alloc(size_t count){
		if(count > cap){ release(); RSQ_CUDA(cudaMalloc(&p, count * sizeof(T))); cap = count; }
		n = count;
	}
	void upload(const std::vector<T> &v, cudaStream_t s){ alloc(v.size()); if(v.size()){ RSQ_CUDA(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s)); } }
	void upload(const T *v, size_t count, cudaStream_t s){ alloc(count); if(count){ RSQ_CUDA(cudaMemcpyAsync(p, v, count * sizeof(T), cudaMemcpyHostToDevice, s)); } }
	void zero(cudaStream_t s){ if(n){ RSQ_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); } }
};

// ---------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------
constexpr uint32_t kNone = 0xffffffffu;
constexpr int kWarpsPerCta = 4;

// positions [p0, p0 + n) of one sequence; sur_start / sur_end point at the entry of position p0
__global__ void k_surroundings(const uint8_t *seq, uint32_t L, uint32_t p0, uint32_t n, const double *t0, const double *t1, const double *t2,
                               double *sur_start, double *sur_end){
	const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
	if(idx >= n){ return; }
	const uint32_t pos = p0 + idx;
	uint32_t code[3];
	forward_surrounding(seq, L, pos, code);
	sur_start[idx] = surrounding_bias(t0, t1, t2, code);
	reverse_surrounding(seq, L, pos, code);
	sur_end[idx] = surrounding_bias(t0, t1, t2, code);
}

// G/C prefix counts of one sequence: entry i = number of G/C in [0, i).  Three small passes
// (tile counts, scan of the tile counts, tile-local scan) instead of shipping 4 bytes/base over PCIe.
constexpr uint32_t kGcTile = 4096;
__global__ void k_gc_tile_counts(const uint8_t *seq, uint32_t L, uint32_t *tile_sums){
	__shared__ uint32_t red[8];
	const uint32_t lo = blockIdx.x * kGcTile;
	uint32_t cnt = 0;
	for(uint32_t i = lo + threadIdx.x; i < min(L, lo + kGcTile); i += blockDim.x){ const uint8_t b = seq[i]; cnt += (b == 1 || b == 2); }
	cnt = __reduce_add_sync(0xffffffffu, cnt);
	if((threadIdx.x & 31) == 0){ red[threadIdx.x >> 5] = cnt; }
	__syncthreads();
	if(threadIdx.x == 0){ uint32_t t = 0; for(uint32_t w = 0; w < (blockDim.x >> 5); ++w){ t += red[w]; } tile_sums[blockIdx.x] = t; }
}
__global__ void k_gc_scan_tiles(uint32_t *tile_sums, uint32_t n_tiles){
	// single CTA, sequential over chunks of blockDim.x tiles
	__shared__ uint32_t buf[1024];
	__shared__ uint32_t carry;
	if(threadIdx.x == 0){ carry = 0; }
	__syncthreads();
	for(uint32_t base = 0; base < n_tiles; base += blockDim.x){
		const uint32_t i = base + threadIdx.x;
		const uint32_t v = i < n_tiles ? tile_sums[i] : 0;
		buf[threadIdx.x] = v;
		__syncthreads();
		for(uint32_t o = 1; o < blockDim.x; o <<= 1){
			const uint32_t add = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
			__syncthreads();
			buf[threadIdx.x] += add;
			__syncthreads();
		}
		if(i < n_tiles){ tile_sums[i] = carry + buf[threadIdx.x] - v; }   // exclusive
		__syncthreads();
		if(threadIdx.x == blockDim.x - 1){ carry += buf[threadIdx.x]; }
		__syncthreads();
	}
}
__global__ void k_gc_prefix(const uint8_t *seq, uint32_t L, const uint32_t *tile_offsets, uint32_t *prefix /*L+1*/){
	// one warp per tile: 32 positions at a time, warp-inclusive scan by shuffles
	const uint32_t lo = blockIdx.x * kGcTile, hi = min(L, lo + kGcTile);
	const uint32_t lane = threadIdx.x;
	uint32_t run = tile_offsets[blockIdx.x];
	if(blockIdx.x == 0 && lane == 0){ prefix[0] = 0; }
	for(uint32_t base = lo; base < hi; base += 32){
		const uint32_t i = base + lane;
		uint32_t v = 0;
		if(i < hi){ const uint8_t b = seq[i]; v = (b == 1 || b == 2); }
		for(int o = 1; o < 32; o <<= 1){ const uint32_t t = __shfl_up_sync(0xffffffffu, v, o); if(lane >= o){ v += t; } }
		if(i < hi){ prefix[i + 1] = run + v; }
		run += __shfl_sync(0xffffffffu, v, 31);
	}
}

struct BiasParamDev { uint32_t ref_id, fragment_length; double general; };

// One CTA per (sequence, sampled fragment length).  The FP64 sum has to follow the reference's sequential order
// (Reference::SumBias), so one warp adds the terms one after the other - a pure chain of dependent DADDs fed from shared
// memory - while the other three warps evaluate the terms of the next 384 positions (coalesced loads, several tiles in
// flight per warp: the chain never waits for HBM/L2 latency).
constexpr uint32_t kBiasProducers = 3, kBiasTilesPerWarp = 4, kBiasSuper = kBiasProducers * kBiasTilesPerWarp * 32;   // 384 positions
__global__ void __launch_bounds__(128)
k_sum_bias(const BiasParamDev *params, uint32_t n_params, const uint64_t *seq_off, const uint32_t *seq_len,
           const double *sur_start, const double *sur_end, const uint32_t *gc_prefix, const double *gc_bias,
           double *sums, double *max_bias){
	__shared__ double s_gc[101];
	__shared__ __align__(16) double s_terms[2][kBiasSuper];
	__shared__ double s_max[kBiasProducers];
	const uint32_t i = blockIdx.x;
	if(i >= n_params){ return; }
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const BiasParamDev p = params[i];
	const uint64_t off = seq_off[p.ref_id];
	const uint32_t L = seq_len[p.ref_id], fl = p.fragment_length;
	const double *ss = sur_start + off, *se = sur_end + off + fl - 1;
	const uint32_t *gp = gc_prefix + off + p.ref_id;
	const uint32_t n_pos = L - fl + 1;
	for(uint32_t k = threadIdx.x; k < 101; k += blockDim.x){ s_gc[k] = gc_bias[k]; }
	__syncthreads();
	double tot = 0.0, mx = 0.0;
	auto produce = [&](uint32_t super_base, double *dst){   // warps 1..3: 4 tiles each
		struct Raw { uint32_t g0, g1; double a, b; };
		Raw r[kBiasTilesPerWarp];
#pragma unroll
		for(uint32_t t = 0; t < kBiasTilesPerWarp; ++t){
			const uint32_t pos = super_base + ((warp - 1u) * kBiasTilesPerWarp + t) * 32u + lane;
			r[t] = Raw{0, 0, 0.0, 0.0};
			if(pos < n_pos){ r[t].g0 = gp[pos]; r[t].g1 = gp[pos + fl]; r[t].a = ss[pos]; r[t].b = se[pos]; }
		}
#pragma unroll
		for(uint32_t t = 0; t < kBiasTilesPerWarp; ++t){
			const uint32_t slot = ((warp - 1u) * kBiasTilesPerWarp + t) * 32u + lane;
			double bias = 0.0;   // positions behind the last one add +0.0: exact
			if(super_base + slot < n_pos){
				bias = mul_rn(p.general, s_gc[percent_u32(r[t].g1 - r[t].g0, fl)]);
				bias = mul_rn(bias, r[t].a);
				bias = mul_rn(bias, r[t].b);
				if(bias > mx){ mx = bias; }
			}
			dst[slot] = bias;
		}
	};
	if(warp){ produce(0, s_terms[0]); }
	__syncthreads();
	uint32_t buf = 0;
	for(uint32_t base = 0; base < n_pos; base += kBiasSuper){
		if(warp){
			if(base + kBiasSuper < n_pos){ produce(base + kBiasSuper, s_terms[buf ^ 1u]); }
		}
		else{
			const double2 *t2 = reinterpret_cast<const double2 *>(s_terms[buf]);
			for(uint32_t k = 0; k < kBiasSuper / 2; k += 16){
				double2 v[16];
#pragma unroll
				for(int q = 0; q < 16; ++q){ v[q] = t2[k + q]; }
#pragma unroll
				for(int q = 0; q < 16; ++q){ tot = add_rn(add_rn(tot, v[q].x), v[q].y); }
			}
		}
		__syncthreads();
		buf ^= 1u;
	}
	for(int o = 16; o; o >>= 1){ mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
	if(warp && lane == 0){ s_max[warp - 1u] = mx; }
	__syncthreads();
	if(threadIdx.x == 0){ sums[i] = tot; max_bias[i] = fmax(fmax(s_max[0], s_max[1]), s_max[2]); }
}

// ---------------------------------------------------------------------------------------------------
// The same sums without the chain (ordered_sum.cuh): chunks of a chain run in parallel from approximate start values, an in-order pass
// joins them exactly.  One thread owns one chunk; a warp evaluates the terms of 32 chunks x 32 positions with coalesced loads into shared
// memory, then every lane adds the 32 terms of its own chunk in order.
// ---------------------------------------------------------------------------------------------------
struct BiasChain { uint32_t ref_id, fragment_length; double general; uint32_t n_pos, n_chunks; uint32_t chunk_first, cta_first; };
struct BiasTerm {   // one term of Reference::SumBias (same expression as sum_bias_chain)
	const double *ss, *se; const uint32_t *gp; const double *gc_bias; uint32_t fl; double general;
	__device__ __forceinline__ double operator()(uint32_t p) const {
		double bias = mul_rn(general, gc_bias[percent_u32(gp[p + fl] - gp[p], fl)]);
		bias = mul_rn(bias, ss[p]);
		return mul_rn(bias, se[p]);
	}
};
__device__ __forceinline__ BiasTerm bias_term_of(const BiasChain &ch, const uint64_t *seq_off, const double *sur_start, const double *sur_end, const uint32_t *gc_prefix, const double *gc_bias){
	const uint64_t off = seq_off[ch.ref_id];
	return BiasTerm{sur_start + off, sur_end + off + ch.fragment_length - 1, gc_prefix + off + ch.ref_id, gc_bias, ch.fragment_length, ch.general};
}
constexpr uint32_t kBiasChunkCta = 128;   // chunks per CTA (one per thread)
// kSpec = false: pass A (plain chunk sums + chunk maxima); true: pass C (exact runs from the start values in `start`)
template<bool kSpec> __global__ void __launch_bounds__(kBiasChunkCta)
k_bias_chunks(const BiasChain *chains, uint32_t n_chains, uint32_t K, const uint64_t *seq_off, const double *sur_start, const double *sur_end,
              const uint32_t *gc_prefix, const double *gc_bias, const double *start, double *out, double *chunk_max, uint32_t *tie){
	__shared__ double s_gc[101];
	__shared__ double s_terms[kBiasChunkCta / 32][32][33];
	// which chain this CTA belongs to (few chains: linear search)
	uint32_t ci = 0;
	while(ci + 1 < n_chains && chains[ci + 1].cta_first <= blockIdx.x){ ++ci; }
	const BiasChain ch = chains[ci];
	for(uint32_t k = threadIdx.x; k < 101; k += blockDim.x){ s_gc[k] = gc_bias[k]; }
	__syncthreads();
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const BiasTerm term = bias_term_of(ch, seq_off, sur_start, sur_end, gc_prefix, s_gc);
	const uint32_t c_warp = (blockIdx.x - ch.cta_first) * kBiasChunkCta + warp * 32u;   // first chunk of this warp
	if(c_warp >= ch.n_chunks){ return; }
	const uint32_t my_chunk = c_warp + lane;
	const bool mine = my_chunk < ch.n_chunks;
	double acc = (kSpec && mine) ? start[ch.chunk_first + my_chunk] : 0.0, mx = 0.0;
	uint32_t tied = 0;
	for(uint32_t j = 0; j < K; j += 32u){
		__syncwarp();
#pragma unroll 4
		for(uint32_t cc = 0; cc < 32u; ++cc){
			const uint32_t chunk = c_warp + cc;
			const uint64_t p = static_cast<uint64_t>(chunk) * K + j + lane;
			double x = 0.0;   // behind the chain's last position: + 0.0 is exact and never a tie
			if(chunk < ch.n_chunks && p < ch.n_pos){ x = term(static_cast<uint32_t>(p)); }
			s_terms[warp][cc][lane] = x;
		}
		__syncwarp();
		if(mine){
#pragma unroll 8
			for(uint32_t i = 0; i < 32u; ++i){
				const double x = s_terms[warp][lane][i];
				if(kSpec){ acc = ordered_step(acc, x, tied); }
				else{ if(x > mx){ mx = x; } acc = add_rn(acc, x); }
			}
		}
	}
	if(mine){
		out[ch.chunk_first + my_chunk] = acc;
		if(kSpec){ tie[ch.chunk_first + my_chunk] = tied; } else{ chunk_max[ch.chunk_first + my_chunk] = mx; }
	}
}
// pass B: a warp per chain - start values g_c = sum of the plain chunk sums in front of chunk c (one addition per chunk), and the chain's maximum
__global__ void k_bias_scan(const BiasChain *chains, uint32_t n_chains, const double *plain, const double *chunk_max, double *start, double *max_out){
	const uint32_t ci = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
	if(ci >= n_chains){ return; }
	const BiasChain ch = chains[ci];
	double g = 0.0, mx = 0.0;
	for(uint32_t base = 0; base < ch.n_chunks; base += 32u){
		const uint32_t c = base + lane;
		const double pv = c < ch.n_chunks ? plain[ch.chunk_first + c] : 0.0;
		const double mv = c < ch.n_chunks ? chunk_max[ch.chunk_first + c] : 0.0;
		mx = fmax(mx, mv);
		double mine = 0.0;
		for(uint32_t i = 0; i < 32u; ++i){
			if(i == lane){ mine = g; }
			g = add_rn(g, __shfl_sync(0xffffffffu, pv, i));
		}
		if(c < ch.n_chunks){ start[ch.chunk_first + c] = mine; }
	}
	for(int o = 16; o; o >>= 1){ mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
	if(lane == 0){ max_out[ci] = mx; }
}
// pass D: a warp per chain walks its chunks in order (all lanes run the same scalar code; the loads of 32 chunks are coalesced)
__global__ void k_bias_resolve(const BiasChain *chains, uint32_t n_chains, uint32_t K, const uint64_t *seq_off, const double *sur_start, const double *sur_end,
                               const uint32_t *gc_prefix, const double *gc_bias, const double *start, const double *out, const uint32_t *tie,
                               double *sums, uint32_t *reran_out){
	const uint32_t ci = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
	if(ci >= n_chains){ return; }
	const BiasChain ch = chains[ci];
	const BiasTerm term = bias_term_of(ch, seq_off, sur_start, sur_end, gc_prefix, gc_bias);
	double s = 0.0;
	uint32_t reran = 0;
	for(uint32_t base = 0; base < ch.n_chunks; base += 32u){
		const uint32_t c = base + lane;
		const bool have = c < ch.n_chunks;
		const double gv = have ? start[ch.chunk_first + c] : 0.0, ov = have ? out[ch.chunk_first + c] : 0.0;
		const uint32_t tv = have ? tie[ch.chunk_first + c] : 0u;
		const uint32_t n = ch.n_chunks - base < 32u ? ch.n_chunks - base : 32u;
		for(uint32_t i = 0; i < n; ++i){
			const double g = __shfl_sync(0xffffffffu, gv, i), o = __shfl_sync(0xffffffffu, ov, i);
			const uint32_t t = __shfl_sync(0xffffffffu, tv, i);
			const uint64_t begin = static_cast<uint64_t>(base + i) * K;
			const uint64_t end = begin + K < ch.n_pos ? begin + K : ch.n_pos;
			s = chunk_resolve(term, static_cast<uint32_t>(begin), static_cast<uint32_t>(end), s, g, o, t, reran);
		}
	}
	if(lane == 0){ sums[ci] = s; reran_out[ci] = reran; }
}

// Continuation of the master mt19937_64 (Simulator::block_seed_gen_): state[0..311] is any window of 312
// consecutive words of the MT sequence x[], state[312] how many of them were already handed out; `n` further
// outputs are appended to out.  x[m+312] = x[m+156] ^ twist(x[m], x[m+1]) only looks >= 156 words back, so one
// CTA advances 156 words per step through a 4 x 156 word ring in shared memory (one barrier per step).
constexpr int kMasterThreads = 160;
constexpr int kMasterStateWords = kMtN + 1;
// CTA b continues the stream from state_in[b] for n words into out + b * n; the state behind them goes to state_out[b] (if given).
__global__ void __launch_bounds__(kMasterThreads) k_master_stream(const uint64_t *state_in_all, uint64_t *state_out_all, uint64_t *out_all, uint64_t n){
	__shared__ uint64_t ring[4 * kMtM];
	const uint64_t *state_in = state_in_all + static_cast<size_t>(blockIdx.x) * kMasterStateWords;
	uint64_t *state = state_out_all ? state_out_all + static_cast<size_t>(blockIdx.x) * kMasterStateWords : nullptr;
	uint64_t *out = out_all + static_cast<size_t>(blockIdx.x) * n;
	const uint32_t j = threadIdx.x;
	for(uint32_t i = j; i < kMtN; i += blockDim.x){ ring[i] = state_in[i]; }
	uint32_t idx = static_cast<uint32_t>(state_in[kMtN]);
	__syncthreads();
	uint64_t done = 0;
	{
		const uint64_t avail = kMtN - idx;
		const uint64_t take = avail < n ? avail : n;
		for(uint64_t i = j; i < take; i += blockDim.x){ out[i] = mt_temper(ring[idx + i]); }
		done = take; idx += static_cast<uint32_t>(take);
	}
	uint32_t k = 0, last_take = 0;
	while(done < n){
		const uint32_t s0 = (k & 3u) * kMtM, s1 = ((k + 1u) & 3u) * kMtM, s2 = ((k + 2u) & 3u) * kMtM;
		const uint64_t left = n - done;
		last_take = left < static_cast<uint64_t>(kMtM) ? static_cast<uint32_t>(left) : kMtM;
		if(j < kMtM){
			const uint64_t a = ring[s0 + j];
			const uint64_t b = (j + 1u < kMtM) ? ring[s0 + j + 1u] : ring[s1];
			const uint64_t c = ring[s1 + j];
			const uint64_t w = (a & 0xFFFFFFFF80000000ull) | (b & 0x7FFFFFFFull);
			const uint64_t v = c ^ (w >> 1) ^ ((w & 1ull) ? 0xB5026F5AA96619E9ull : 0ull);
			ring[s2 + j] = v;
			if(j < last_take){ out[done + j] = mt_temper(v); }
		}
		done += last_take;
		++k;
		__syncthreads();
	}
	if(!state){ return; }
	// persist the window made of the two most recent segments
	if(k){
		const uint32_t s0 = (k & 3u) * kMtM, s1 = ((k + 1u) & 3u) * kMtM;
		for(uint32_t i = j; i < kMtM; i += blockDim.x){ state[i] = ring[s0 + i]; state[kMtM + i] = ring[s1 + i]; }
		if(j == 0){ state[kMtN] = kMtM + last_take; }
	}
	else{
		for(uint32_t i = j; i < kMtN; i += blockDim.x){ state[i] = ring[i]; }
		if(j == 0){ state[kMtN] = idx; }
	}
}

// Jump-ahead of the master stream by J = 2^k words (tools/gen_mt_jump.py): the window J steps ahead is the XOR of the
// windows at the offsets given by the bits of g_J(t) = t^J mod phi(t).
//   k_master_jump_gen  one CTA: the 19937 + 312 words behind the current window (128 dependent steps of 156 words, in shared
//                      memory), copied to HBM/L2; clears the output window
//   k_master_jump_xor  39 CTAs: CTA b takes 8 words (512 bits) of the polynomial, thread j XORs seq[i + j] over their set bits i
//                      and merges its part into output word j with one atomic XOR
constexpr int kJumpSeqWords = 19937 + kMtN + 7;
constexpr int kJumpPolyWordsPerCta = 8;
#include "mt_jump_tables.inc"
__global__ void __launch_bounds__(kMasterThreads) k_master_jump_gen(const uint64_t *state_in, uint64_t *seq_out, uint64_t *state_out){
	extern __shared__ __align__(16) uint64_t jump_seq[];
	const uint32_t t = threadIdx.x;
	for(uint32_t i = t; i < kMtN; i += blockDim.x){ jump_seq[i] = state_in[i]; state_out[i] = 0; }
	if(t == 0){ state_out[kMtN] = state_in[kMtN]; }
	__syncthreads();
	for(uint32_t base = 0; base + kMtN < static_cast<uint32_t>(kJumpSeqWords); base += kMtM){
		const uint32_t i = base + t;
		if(t < kMtM && i + kMtN < static_cast<uint32_t>(kJumpSeqWords)){
			const uint64_t w = (jump_seq[i] & 0xFFFFFFFF80000000ull) | (jump_seq[i + 1] & 0x7FFFFFFFull);
			jump_seq[i + kMtN] = jump_seq[i + kMtM] ^ (w >> 1) ^ ((w & 1ull) ? 0xB5026F5AA96619E9ull : 0ull);
		}
		__syncthreads();
	}
	for(uint32_t i = t; i < static_cast<uint32_t>(kJumpSeqWords); i += blockDim.x){ seq_out[i] = jump_seq[i]; }
}
__global__ void __launch_bounds__(320) k_master_jump_xor(const uint64_t *seq, const uint64_t *poly, uint64_t *state_out){
	const uint32_t j = threadIdx.x;
	if(j >= kMtN){ return; }
	uint64_t acc = 0;
	const uint32_t w0 = blockIdx.x * kJumpPolyWordsPerCta;
	for(uint32_t w = w0; w < w0 + kJumpPolyWordsPerCta && w < static_cast<uint32_t>(kMtN); ++w){
		uint64_t bits = poly[w];
		while(bits){
			const uint32_t i = w * 64u + static_cast<uint32_t>(__ffsll(static_cast<long long>(bits)) - 1);
			acc ^= seq[i + j];
			bits &= bits - 1ull;
		}
	}
	if(acc){ atomicXor(reinterpret_cast<unsigned long long *>(state_out + j), static_cast<unsigned long long>(acc)); }
}

__global__ void k_master_seed(uint64_t *state, uint64_t seed){
	if(threadIdx.x == 0 && blockIdx.x == 0){
		uint64_t x = seed;
		state[0] = x;
		for(int i = 1; i < kMtN; ++i){ x = 6364136223846793005ull * (x ^ (x >> 62)) + static_cast<uint64_t>(i); state[i] = x; }
		state[kMtN] = kMtN;
	}
}

struct SysChain {
	const uint8_t *seq; uint32_t L; uint32_t reverse; const uint64_t *raw; uint32_t seed_interleaved; uint8_t *out; uint32_t carried_dom; uint32_t pad;
	const uint64_t *blk_off;   // runs with variants: where every SimBlock's draws start in raw (chain_raw); null otherwise
	uint32_t *bstate;          // runs with variants: distance state in front of every SimBlock (for the draws of its variants)
};
constexpr uint32_t kSysCheckpoint = 64;   // positions between two checkpoints of a chunk's Markov state (power of two)
struct SysChunk {
	uint32_t chain; uint32_t begin; uint32_t end; uint32_t warm_from;
	uint32_t in_dist, in_rate, out_dist, out_rate;
	uint32_t dirty; uint32_t first_of_chain; uint32_t computed /* ran once: its checkpoints are valid */, pad1;
};

__global__ void k_sys_chunks(Tables tab, const SysChain *chains, SysChunk *chunks, uint32_t n_chunks, uint32_t max_n0,
                             uint32_t sys_gc_range, uint32_t reset_distance){
	extern __shared__ double prob_all[];
	WarpGroup g;
	const uint32_t warp = threadIdx.x >> 5;
	double *prob = prob_all + static_cast<size_t>(warp) * ((max_n0 + 3) & ~3u);
	const uint32_t c = blockIdx.x * (blockDim.x >> 5) + warp;
	if(c >= n_chunks){ return; }
	SysChunk ck = chunks[c];
	if(!ck.dirty){ return; }
	const SysChain ch = chains[ck.chain];
	SysState st{ck.in_dist, ck.in_rate};
	if(ck.warm_from < ck.begin){
		st = sys_error_chain(g, tab, prob, ch.seq, ch.L, ch.reverse != 0, ck.warm_from, ck.begin, SysState{0, 0}, ch.carried_dom, sys_gc_range, reset_distance, ch.raw, ch.seed_interleaved != 0, nullptr, ch.blk_off);
		if(g.lane() == 0){ chunks[c].in_dist = st.distance; chunks[c].in_rate = st.start_rate; chunks[c].warm_from = ck.begin; }
	}
	st = sys_error_chain(g, tab, prob, ch.seq, ch.L, ch.reverse != 0, ck.begin, ck.end, st, ch.carried_dom, sys_gc_range, reset_distance, ch.raw, ch.seed_interleaved != 0, ch.out, ch.blk_off, ch.bstate);
	if(g.lane() == 0){ chunks[c].out_dist = st.distance; chunks[c].out_rate = st.start_rate; chunks[c].dirty = 0; }
}

// After a pass: a chunk whose assumed start state differs from its predecessor's end state must be redone.
__global__ void k_sys_check(SysChunk *chunks, uint32_t n_chunks, uint32_t *n_dirty){
	const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
	if(c >= n_chunks || chunks[c].first_of_chain){ return; }
	const SysChunk prev = chunks[c - 1];
	SysChunk &me = chunks[c];
	if(me.in_dist != prev.out_dist || me.in_rate != prev.out_rate){
		me.in_dist = prev.out_dist; me.in_rate = prev.out_rate; me.dirty = 1;
		atomicAdd(n_dirty, 1u);
	}
}

__global__ void k_build_blocks(BlockDesc *blocks, uint32_t first, uint32_t nb, uint32_t ref_id, uint32_t first_block_id, const uint64_t *fwd_raw,
                               const int32_t *first_meth /* per block of the run, or null */, uint32_t seed_stride /* 2001, or 1 with --readSysError */,
                               const uint64_t *fwd_off = nullptr /* runs with variants: index of every block's seed in fwd_raw */, const uint32_t *block_first = nullptr){
	const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
	if(b >= nb){ return; }
	BlockDesc d; d.ref_id = ref_id; d.start_pos = b * 1000u; d.block_id = first_block_id + b; d.first_meth = first_meth ? first_meth[first + b] : 0;
	d.seed = fwd_off ? fwd_raw[fwd_off[b]] : fwd_raw[static_cast<size_t>(b) * seed_stride];
	d.first_var = block_first ? block_first[b] : 0u; d.pad = 0;
	blocks[first + b] = d;
}

// Simulator::SetSystematicErrorVariantsForward / Reverse (sim_core.cuh: draw_variant_errors_block): one thread per (SimBlock, strand) of a sequence.
// The draws are few (two per replacement base), the walk over the block's error rates in front of every variant is the work.
__global__ void __launch_bounds__(128)
k_var_sys_errors(Tables tab, VarDrawCtx fwd, VarDrawCtx rev, uint32_t nb, const uint32_t *bstate_fwd, const uint32_t *bstate_rev, const uint64_t *raw,
                 const uint64_t *fwd_off, const uint64_t *rev_off){
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= 2u * nb){ return; }
	const uint32_t strand = i >= nb ? 1u : 0u, b = strand ? i - nb : i;
	const VarDrawCtx &dc = strand ? rev : fwd;
	if(dc.block_first[b + 1] == dc.block_first[b]){ return; }
	double prob[kMaxN0 + 4];
	SingleLane one;
	const uint32_t bs = (strand ? bstate_rev : bstate_fwd)[b];
	const uint64_t e = 1000ull * (b + 1ull) < dc.L ? 1000ull * (b + 1ull) : dc.L;
	const uint64_t size = e - 1000ull * b;
	draw_variant_errors_block(one, tab, prob, dc, b, SysState{bs & 0xffffffu, bs >> 24}, raw + (strand ? rev_off[b] + 2ull * size : fwd_off[b] + 1ull + 2ull * size));
}

struct Arena {
	unsigned char *data; uint32_t chunk_bytes; uint32_t n_chunks;
	uint32_t *next_free; uint32_t *chunk_next; uint32_t *chunk_used; uint32_t *error_flag;
};

struct DeviceSink {
	Arena a;
	// per segment state kept in scalars (no dynamically indexed arrays -> stays in registers)
	uint32_t cur0, cur1, used0, used1, head0, head1;
	unsigned long long bytes0, bytes1;
	uint32_t pairs;
	__device__ void init(const Arena &arena){ a = arena; cur0 = cur1 = kNone; used0 = used1 = 0; head0 = head1 = kNone; bytes0 = bytes1 = 0; pairs = 0; }
	__device__ void write_record(const WarpGroup &g, uint32_t seg, const char *id, int id_len, const uint8_t *seq, const uint8_t *qual, uint32_t n){
		const uint32_t rec = 1u + id_len + 1u + n + 3u + n + 1u;
		if(rec > a.chunk_bytes){ if(g.lane() == 0){ atomicOr(a.error_flag, kErrRecordTooLong); } return; }
		uint32_t cur = seg ? cur1 : cur0, used = seg ? used1 : used0;
		if(cur == kNone || used + rec > a.chunk_bytes){
			uint32_t idx = kNone;
			if(g.lane() == 0){
				idx = atomicAdd(a.next_free, 1u);
				if(idx >= a.n_chunks){ atomicOr(a.error_flag, kErrArenaFull); idx = kNone; }
				else{
					a.chunk_next[idx] = kNone; a.chunk_used[idx] = 0;
					if(cur != kNone){ a.chunk_next[cur] = idx; }
				}
			}
			idx = __shfl_sync(0xffffffffu, idx, 0);
			if(idx == kNone){ return; }
			if(cur == kNone){ if(seg){ head1 = idx; } else{ head0 = idx; } }
			cur = idx; used = 0;
		}
		unsigned char *dst = a.data + static_cast<size_t>(cur) * a.chunk_bytes + used;
		const uint32_t o_seq = 2u + id_len, o_plus = o_seq + n, o_qual = o_plus + 3u;
		for(uint32_t i = g.lane(); i < rec; i += 32){
			unsigned char ch;
			if(i == 0){ ch = '@'; }
			else if(i < 1u + id_len){ ch = static_cast<unsigned char>(id[i - 1]); }
			else if(i < o_seq){ ch = '\n'; }
			else if(i < o_plus){ const uint32_t b = seq[i - o_seq]; ch = static_cast<unsigned char>(b > 3u ? 'N' : (0x54474341u >> (8u * b)) & 0xffu); }
			else if(i < o_qual){ ch = (i == o_plus + 1u) ? '+' : '\n'; }
			else if(i < o_qual + n){ ch = qual[i - o_qual]; }
			else{ ch = '\n'; }
			dst[i] = ch;
		}
		used += rec;
		if(seg){ cur1 = cur; used1 = used; bytes1 += rec; } else{ cur0 = cur; used0 = used; bytes0 += rec; }
		if(g.lane() == 0){ a.chunk_used[cur] = used; }
		g.sync();
	}
	__device__ void pair_done(const WarpGroup &){ ++pairs; }
};

struct BlockOut { uint32_t head[2]; unsigned long long bytes[2]; uint32_t pairs; uint32_t pad; unsigned long long scan_draws; };

__global__ void k_spec_block_out(SpecCtx sp, BlockOut *out){
	const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
	if(u >= sp.n_units){ return; }
	const SpecBlock &b = sp.blocks[u];
	BlockOut o; o.head[0] = b.chain_head; o.head[1] = b.chain_head; o.bytes[0] = b.bytes[0]; o.bytes[1] = b.bytes[1];
	o.pairs = sp.em_recs ? b.reads : b.reads / 2u; o.pad = b.rounds; o.scan_draws = b.scan_draws;   // seqToIllumina counts reads
	out[u] = o;
}

template<bool kMeth>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 8)   // 8 CTAs x 4 warps = 32 blocks in flight per SM (<= 64 registers)
k_simulate(SimCtx c, const BlockDesc *blocks, uint32_t first_block, uint32_t n_blocks, Arena arena, BlockOut *out, uint32_t *next_block,
           uint32_t max_n0, uint32_t scratch_per_warp){
	extern __shared__ __align__(16) unsigned char smem[];
	WarpGroup g;
	const uint32_t warp = threadIdx.x >> 5;
	Scratch s = carve_scratch(smem + static_cast<size_t>(warp) * scratch_per_warp, max_n0, c.max_org_len, c.max_read_len);
	while(true){
		uint32_t i = 0;
		if(g.lane() == 0){ i = atomicAdd(next_block, 1u); }
		i = __shfl_sync(0xffffffffu, i, 0);
		if(i >= n_blocks){ break; }
		DeviceSink sink; sink.init(arena);
		unsigned long long draws = 0;
		const BlockDesc b = blocks[first_block + i];
		if(c.var.loaded){
			// runs with variants: the chosen (allele, strand) ids of a hit live behind the warp's scratch
			simulate_block_var<kMeth>(g, c, s, sink, b, &draws, reinterpret_cast<uint16_t *>(smem + static_cast<size_t>(kWarpsPerCta) * scratch_per_warp) + static_cast<size_t>(warp) * 2u * c.var.num_alleles);
		}
		else{ simulate_block<kMeth>(g, c, s, sink, b, &draws); }
		draws = __shfl_sync(0xffffffffu, draws, 0);
		if(g.lane() == 0){
			BlockOut o; o.head[0] = sink.head0; o.head[1] = sink.head1; o.bytes[0] = sink.bytes0; o.bytes[1] = sink.bytes1;
			o.pairs = sink.pairs; o.pad = 0; o.scan_draws = draws;
			out[i] = o;
		}
	}
}

// ---------------------------------------------------------------------------------------------------
// Speculative two-phase path (spec_core.cuh): k_spec_scan (warp per unit) and k_spec_reads (lane per read) alternate
// in rounds until every unit has verified all its reads; k_spec_gather assembles the FASTQ text from the record slots.
// ---------------------------------------------------------------------------------------------------
__global__ void k_spec_init(SimCtx c, SpecCtx sp, const BlockDesc *descs, uint32_t first_desc){
	const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
	if(u < sp.n_units){ spec_init_unit(c, sp, descs, first_desc, u); }
}

// --- bulk asynchronous copies (TMA unit, cp.async.bulk -> SASS UBLKCP) completing on an mbarrier -------------------------------------
// Contiguous global -> shared staging without registers: one lane arms the barrier with the byte count and issues the copy, every lane
// waits on the barrier's phase.  Source, destination and size are multiples of 16 bytes.
__device__ __forceinline__ uint32_t smem_addr(const void *p){ return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals){
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(arrivals) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // visible to the async proxy before a copy names the barrier
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes){
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst_shared, const void *src_global, uint32_t bytes, uint64_t *bar){
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(smem_addr(dst_shared)), "l"(src_global), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity){
	asm volatile("{\n.reg .pred p;\nRSQ_MBAR_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra RSQ_MBAR_DONE;\nbra RSQ_MBAR_WAIT;\nRSQ_MBAR_DONE:\n}"
	             :: "r"(smem_addr(bar)), "r"(parity) : "memory");
}

template<bool kVar>
__device__ __forceinline__ void spec_scan_body(const SimCtx &c, const SpecCtx &sp, const BlockDesc *descs, uint32_t first_desc, uint32_t unit_first, uint32_t unit_end){
	__shared__ uint64_t rings[kWarpsPerCta][2 * kMtN];
	__shared__ __align__(8) uint64_t thr_bar[kWarpsPerCta];
	// dynamic: per warp its unit's row of threshold high words (c.thr_hi_stride entries, 16-byte rows), then - runs with variants - 2 * num_alleles chosen (allele, strand) ids
	extern __shared__ __align__(16) unsigned char scan_dyn[];
	typename std::conditional<kVar, WarpGroupCompact, WarpGroup>::type g;
	const uint32_t warp = threadIdx.x >> 5;
	const uint32_t u = unit_first + blockIdx.x * kWarpsPerCta + warp;
	if(u >= unit_end){ return; }
	if(sp.blocks[u].done){ return; }
	// stage the threshold row of this unit's coverage group: the scan reads one entry per draw and nothing else from global memory
	uint32_t *thr_row = reinterpret_cast<uint32_t *>(scan_dyn) + static_cast<size_t>(warp) * c.thr_hi_stride;
	const bool scanning = sp.em_recs == nullptr && u < sp.n_blocks;
	if(scanning){
		const uint32_t group = c.coverage_group[descs[first_desc + u].ref_id];
		if((threadIdx.x & 31u) == 0u){
			mbar_init(&thr_bar[warp], 1);
			mbar_expect_tx(&thr_bar[warp], c.thr_hi_stride * 4u);
			bulk_copy_g2s(thr_row, c.thr_hi + static_cast<size_t>(group) * c.thr_hi_stride, c.thr_hi_stride * 4u, &thr_bar[warp]);
		}
		__syncwarp();
		mbar_wait(&thr_bar[warp], 0);
	}
	uint16_t *chosen = reinterpret_cast<uint16_t *>(scan_dyn + static_cast<size_t>(kWarpsPerCta) * c.thr_hi_stride * 4u) + static_cast<size_t>(warp) * sp.chosen_stride;
	scan_window<kVar>(g, c, sp, descs, first_desc, u, rings[warp], chosen, scanning ? thr_row : nullptr);
}
// The two contexts are __grid_constant__: the out-of-line helpers (snapshots, read plans, allele evaluation ...) take them by reference straight from
// the constant bank instead of forcing a 1 KB copy into local memory; out of line they are because the scan's code has to stay within the
// instruction cache (with everything inlined the variant instantiation was 384 KB of SASS and stalled on instruction fetches).
template<bool kVar> __global__ void k_spec_scan(const __grid_constant__ SimCtx c, const __grid_constant__ SpecCtx sp, const BlockDesc *descs, uint32_t first_desc, uint32_t unit_first, uint32_t unit_end);
#ifdef RSQ_SCAN_MINBLOCKS   // A/B builds (tools/build_variant.sh): the register budget the compiler works with
#define RSQ_SCAN_BOUNDS kWarpsPerCta * 32, RSQ_SCAN_MINBLOCKS
#else
#define RSQ_SCAN_BOUNDS kWarpsPerCta * 32
#endif
template<> __global__ void __launch_bounds__(RSQ_SCAN_BOUNDS)
k_spec_scan<false>(const __grid_constant__ SimCtx c, const __grid_constant__ SpecCtx sp, const BlockDesc *descs, uint32_t first_desc, uint32_t unit_first, uint32_t unit_end){ spec_scan_body<false>(c, sp, descs, first_desc, unit_first, unit_end); }
#ifndef RSQ_VAR_SCAN_MINBLOCKS
#define RSQ_VAR_SCAN_MINBLOCKS 4   // 118 registers, 16 warps per SM: 200 Mbp + VCF 4382 ms against 4737 ms with 5 blocks (96 registers, spills) and 4841 ms with 3
#endif
template<> __global__ void __launch_bounds__(kWarpsPerCta * 32, RSQ_VAR_SCAN_MINBLOCKS)
k_spec_scan<true>(const __grid_constant__ SimCtx c, const __grid_constant__ SpecCtx sp, const BlockDesc *descs, uint32_t first_desc, uint32_t unit_first, uint32_t unit_end){ spec_scan_body<true>(c, sp, descs, first_desc, unit_first, unit_end); }

// LogArrayResult::Draw for up to 32 independent reads at once (lanes 0 .. n_rows-1 own one read each).  The likelihood
// products of every read are computed cooperatively (the lanes of a group of 8/16/32 take consecutive candidates of one
// read: coalesced table rows) and parked in shared memory; each lane then runs the two strictly ordered FP64 sums of
// its own read.
#ifndef RSQ_PROD_UNROLL
#define RSQ_PROD_UNROLL 4   // (measured: 2 -> E. coli 55.0 ms / realistic tables 87.4 ms / 100 Mbp 983 ms; 4 -> 52.3 / 73.7 / 918) candidate PAIRS per lane whose four 16-byte table-row loads are in flight together (each trip of the product loop waits one L2 round trip)
#endif
constexpr int kProdUnroll = RSQ_PROD_UNROLL;
#ifndef RSQ_COOP_WIDE_FROM
#define RSQ_COOP_WIDE_FROM 20
#endif
constexpr uint32_t kCoopWideFrom = RSQ_COOP_WIDE_FROM;   // candidate lists longer than this get four lanes per read
#ifndef RSQ_COOP_WIDE_SHIFT
#define RSQ_COOP_WIDE_SHIFT 3   // more than 32 candidates: eight lanes per read (profile150q: 84.5 ms; four lanes 94.7, sixteen 87.8, two 122.5)
#endif
constexpr uint32_t kCoopWideShift = RSQ_COOP_WIDE_SHIFT;
#ifndef RSQ_COOP_MID_SHIFT
#define RSQ_COOP_MID_SHIFT 2
#endif
constexpr uint32_t kCoopMidShift = RSQ_COOP_MID_SHIFT;
__device__ __forceinline__ uint32_t coop_draw(const Tables &t, double *buf, uint32_t stride, uint32_t n_rows, bool active, uint32_t table_id,
                                               uint32_t i0, uint32_t i1, uint32_t i2, uint32_t i3, double u, bool &zero){
	const unsigned amask = __ballot_sync(0xffffffffu, active);
	zero = false;
	if(amask == 0){ return 0; }
	const uint32_t lane = threadIdx.x & 31;
	uint32_t n0 = 0, nm = 0, par0_off = 0, o0 = 0, o1 = 0, o2 = 0, o3 = 0;
	if(active){
		const TableDesc d = t.desc[table_id];
		n0 = d.n0; nm = d.nm; par0_off = d.par0_off;
		o0 = d.off[0] + adjust_index(i0, d.from[0], d.span[0]) * d.stride;
		o1 = d.off[1] + adjust_index(i1, d.from[1], d.span[1]) * d.stride;
		o2 = d.off[2] + adjust_index(i2, d.from[2], d.span[2]) * d.stride;
		o3 = nm > 3 ? d.off[3] + adjust_index(i3, d.from[3], d.span[3]) * d.stride : o0;
	}
	const uint32_t maxn = __reduce_max_sync(0xffffffffu, n0);
	if(maxn == 0){ zero = true; return 0; }
	const uint32_t n4 = (maxn + 3u) & ~3u;
	__syncwarp();
	// 32 / rows-per-pass lanes work on every read of a pass at the same time (one trip of shuffles, all reads of the pass in
	// flight together); 8 or 16 reads per pass, 32 reads per warp take two passes of 16
#ifdef RSQ_COOP_LPR_SHIFT   // A/B builds: 2^shift lanes per read and pass (fewer reads per pass = fewer distinct table rows, i.e. L1 wavefronts, per load instruction)
	const uint32_t lpr_shift = RSQ_COOP_LPR_SHIFT, lpr = 1u << lpr_shift, pass_rows = 32u >> lpr_shift;
#else
	// two lanes per read for short candidate lists, four or eight for long ones (a quality draw of a profile with 40 quality values): more lanes per read mean
	// fewer distinct table rows - L1 wavefronts, the pipe this kernel saturates - per load instruction, but idle lanes when the list is short
	// (E. coli: 57.9 ms with two lanes on profile150r against 65.6 with four; 122.5 against 103.3 ms on profile150q)
	uint32_t lpr_shift = n4 > 32u ? kCoopWideShift : (n4 > kCoopWideFrom ? kCoopMidShift : 1u);
	if((32u >> lpr_shift) > n_rows){ lpr_shift = 2u; }   // 8 reads per warp: one pass of 8
	const uint32_t lpr = 1u << lpr_shift, pass_rows = 32u >> lpr_shift;
#endif
	for(uint32_t first = 0; first < n_rows; first += pass_rows){
		const uint32_t src = first + (lane >> lpr_shift), i = lane & (lpr - 1u);
		const uint32_t sn0 = __shfl_sync(0xffffffffu, n0, src);
		const uint32_t s0 = __shfl_sync(0xffffffffu, o0, src), s1 = __shfl_sync(0xffffffffu, o1, src);
		const uint32_t s2 = __shfl_sync(0xffffffffu, o2, src), s3 = __shfl_sync(0xffffffffu, o3, src);
		const bool four = __shfl_sync(0xffffffffu, nm, src) > 3u;
		double *row = buf + src * stride;
		if((amask >> src) & 1u){
			// two candidates per lane and trip: rows are 16-byte aligned pairs with a 0.0 behind an odd last candidate (the product of the padding is
			// the +0.0 the sums expect there); a pair that starts behind the list is not loaded at all
#pragma unroll kProdUnroll
			for(uint32_t idx = 2u * i; idx < n4; idx += 2u * lpr){
				double2 p = make_double2(0.0, 0.0);
				if(idx < sn0){
					p = __ldg(reinterpret_cast<const double2 *>(t.blob + s0 + idx));
					const double2 b = __ldg(reinterpret_cast<const double2 *>(t.blob + s1 + idx));
					const double2 cc = __ldg(reinterpret_cast<const double2 *>(t.blob + s2 + idx));
					p.x = mul_rn(mul_rn(p.x, b.x), cc.x); p.y = mul_rn(mul_rn(p.y, b.y), cc.y);
					if(four){ const double2 e = __ldg(reinterpret_cast<const double2 *>(t.blob + s3 + idx)); p.x = mul_rn(p.x, e.x); p.y = mul_rn(p.y, e.y); }
				}
				row[idx] = p.x; row[idx + 1u] = p.y;
			}
		}
	}
	__syncwarp();
	const double *row = buf + lane * stride;
	uint32_t result = 0;
	if(active){
		double prob_sum = 0.0;
		for(uint32_t k = 0; k < n4; k += 4){
			const double a = row[k], b = row[k + 1], cc = row[k + 2], e = row[k + 3];
			prob_sum = add_rn(add_rn(add_rn(add_rn(prob_sum, a), b), cc), e);
		}
		zero = (0.0 == prob_sum) || n0 == 0u;
		if(n0){
			const double r = mul_rn(u, prob_sum);
			double sum = 0.0;
			uint32_t ind0 = n0;
			// reverse cumulative search, four candidates per trip: the loads leave the dependent add chain
			while(ind0 >= 5u && sum <= r){
				const double a = row[ind0 - 1u], b = row[ind0 - 2u], cc = row[ind0 - 3u], e = row[ind0 - 4u];
				const double s1 = add_rn(sum, a), s2 = add_rn(s1, b), s3 = add_rn(s2, cc), s4 = add_rn(s3, e);
				if(!(s1 <= r)){ ind0 -= 1u; sum = s1; }
				else if(!(s2 <= r)){ ind0 -= 2u; sum = s2; }
				else if(!(s3 <= r)){ ind0 -= 3u; sum = s3; }
				else{ ind0 -= 4u; sum = s4; }
			}
			while(sum <= r && --ind0){ sum = add_rn(sum, row[ind0]); }
			result = __ldg(t.par0 + par0_off + ind0);
		}
	}
	return result;
}

// Simulator::SetSystematicErrors for `lanes_per_warp` chunks per warp in lock step (one lane per chunk, the other lanes
// help with the likelihood products): the same chain as sys_error_chain() in sim_core.cuh - warm-up from a reset state,
// then the chunk itself - with both Draws of a position going through coop_draw.
__global__ void __launch_bounds__(128)
k_sys_chunks_lanes(Tables tab, const SysChain *chains, SysChunk *chunks, uint32_t n_chunks, uint32_t stride, uint32_t lanes_per_warp,
                   uint32_t sys_gc_range, uint32_t reset_distance, uint32_t *checkpoints, uint32_t cp_per_chunk){
	extern __shared__ __align__(16) unsigned char smem[];
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	double *buf = reinterpret_cast<double *>(smem) + static_cast<size_t>(warp) * lanes_per_warp * stride;
	const uint32_t c = (blockIdx.x * (blockDim.x >> 5) + warp) * lanes_per_warp + lane;
	bool have = lane < lanes_per_warp && c < n_chunks;
	SysChunk ck{};
	if(have){ ck = chunks[c]; have = ck.dirty != 0; }
	if(!__any_sync(0xffffffffu, have)){ return; }
	SysChain ch{};
	if(have){ ch = chains[ck.chain]; }
	const bool reverse = ch.reverse != 0, interleaved = ch.seed_interleaved != 0;
	const bool warming = have && ck.warm_from < ck.begin;
	uint32_t p = have ? ck.warm_from : 0u, end = have ? ck.end : 0u;
	SysState st{warming ? 0u : ck.in_dist, warming ? 0u : ck.in_rate};
	// A chunk that is run again (its predecessor ended in another state than assumed) stops as soon as its Markov state equals the one its last
	// run had at the same position (checkpoints every kSysCheckpoint positions): everything behind that point is what it was, the draws being
	// a function of the position.  A fix-up pass then costs one checkpoint interval instead of a whole chunk.
	const bool rerun = have && ck.computed != 0;
	bool converged = false;
	uint32_t *cp = checkpoints + static_cast<size_t>(have ? c : 0u) * cp_per_chunk;
	// state in front of p: GC window (Simulator::UpdateGC), last base, dominant-base window
	uint32_t gc_bases = 0, gc = 0, last_base = 4u, hist = 0, nwin = 0;
	if(have){
		gc_bases = p < sys_gc_range ? p : sys_gc_range;
		for(uint32_t q = p - gc_bases; q < p; ++q){
			const uint32_t b = chain_base(ch.seq, ch.L, reverse, q);
			gc += (b == 1 || b == 2) ? 1u : 0u;
		}
		last_base = p ? chain_base(ch.seq, ch.L, reverse, p - 1) : 4u;
		window_before(ch.seq, ch.L, reverse, p, hist, nwin);
	}
	bool zero;
	while(true){
		bool active = have && p < end;
		if(active && p > ck.begin && ((p - ck.begin) & (kSysCheckpoint - 1u)) == 0u){
			const uint32_t k = (p - ck.begin) / kSysCheckpoint - 1u;
			const uint32_t packed = st.distance | (st.start_rate << 24);
			if(rerun && cp[k] == packed){ converged = true; have = false; active = false; }
			else{ cp[k] = packed; }
		}
		if(!__any_sync(0xffffffffu, active)){ break; }
		uint32_t ref_base = 0, dom_base = 0, gc_percent = 50u, dist = 0, t1 = 0;
		double u1 = 0.0, u2 = 0.0;
		if(active){
			if(p == ck.begin && warming){ ck.in_dist = st.distance; ck.in_rate = st.start_rate; }
			ref_base = chain_base(ch.seq, ch.L, reverse, p);
			dom_base = dominant_from_window(hist, nwin, ch.carried_dom);
			gc_percent = gc_bases ? percent_u16(gc, gc_bases) : 50u;
			dist = (st.distance + 9) / 10;
			if(ch.bstate && p >= ck.begin){
				uint32_t blk;
				if(chain_block_start(ch.L, reverse, p, blk)){ ch.bstate[blk] = st.distance | (st.start_rate << 24); }
			}
			u1 = canonical(chain_raw(ch.raw, interleaved, p, 0, ch.blk_off, ch.L, reverse));
			u2 = canonical(chain_raw(ch.raw, interleaved, p, 1, ch.blk_off, ch.L, reverse));
			t1 = tab.dom_error(ref_base, last_base, dom_base);
		}
		uint32_t dom_error = coop_draw(tab, buf, stride, lanes_per_warp, active, t1, dist, gc_percent, st.start_rate, 0, u1, zero);
		if(zero){ dom_error = 4; }
		const uint32_t t2 = active ? tab.error_rate(ref_base, dom_error) : 0u;
		uint32_t error_rate = coop_draw(tab, buf, stride, lanes_per_warp, active, t2, dist, gc_percent, st.start_rate, 0, u2, zero);
		if(zero){ error_rate = 0; }
		error_rate &= 0xffu;
		if(active){
			if(p >= ck.begin){
				ch.out[2 * static_cast<size_t>(p)] = static_cast<uint8_t>(dom_error);
				ch.out[2 * static_cast<size_t>(p) + 1] = static_cast<uint8_t>(error_rate);
			}
			last_base = ref_base;
			hist = ((hist << 2) | ref_base) & 0x3ffu;
			if(nwin < 5){ ++nwin; }
			if(st.distance){   // CoverageStats::UpdateDistances
				if(st.start_rate < error_rate){ st.distance = 0; st.start_rate = error_rate; }
				else if(++st.distance >= reset_distance){ st.distance = 0; st.start_rate = 0; }
			}
			else if(error_rate){ st.distance = 1; st.start_rate = error_rate; }
			if(ref_base == 1 || ref_base == 2){ ++gc; }
			if(gc_bases < sys_gc_range){ ++gc_bases; }
			else{
				const uint32_t ob = chain_base(ch.seq, ch.L, reverse, p - gc_bases);
				if(ob == 1 || ob == 2){ --gc; }
			}
			++p;
		}
	}
	if(have){
		SysChunk &o = chunks[c];
		if(warming){ o.in_dist = ck.in_dist; o.in_rate = ck.in_rate; o.warm_from = ck.begin; }
		o.out_dist = st.distance; o.out_rate = st.start_rate; o.dirty = 0; o.computed = 1;
	}
	else if(converged){ chunks[c].dirty = 0; }   // the state at the chunk's end is the one of its last run
}

// Lanes 0 .. lanes_per_warp-1 of a warp own one read each; the other lanes only help with the likelihood products.
// Few reads per warp = short latency per round (small genomes), 32 = fewest instructions per read (large ones).
constexpr int kSpecReadWarps = 4;
template<bool kVar>
__device__ __forceinline__ void spec_reads_body(const SimCtx &c, const SpecCtx &sp, uint32_t stride, uint32_t lanes_per_warp, uint32_t unit_first, uint32_t unit_end){
	extern __shared__ __align__(16) unsigned char smem[];
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	double *buf = reinterpret_cast<double *>(smem) + static_cast<size_t>(warp) * lanes_per_warp * stride;
	const size_t gwarp = static_cast<size_t>(blockIdx.x) * kSpecReadWarps + warp;
	// dense over (unit, read of this round): reads beyond run_depth do not exist in this round - except for the adapter-only
	// pseudo unit (always full depth), whose further reads are appended behind the regular ones of its group
	const size_t dense = gwarp * lanes_per_warp + lane;
	const size_t regular = static_cast<size_t>(unit_end - unit_first) * sp.map_depth;
	uint32_t u, k;
	if(dense < regular){ u = unit_first + static_cast<uint32_t>(dense / sp.map_depth); k = static_cast<uint32_t>(dense % sp.map_depth); }
	else{ u = sp.n_blocks; k = sp.map_depth + static_cast<uint32_t>(dense - regular); }
	const bool in_range = dense < regular || (sp.n_units > sp.n_blocks && unit_end == sp.n_units && k < sp.depth);
	const size_t gidx = static_cast<size_t>(u) * sp.depth + k;
	bool have = false;
	ReadJob job{};
	if(lane < lanes_per_warp && in_range){
		const SpecBlock &b = sp.blocks[u];
		have = !b.done && k < b.n_jobs;
		if(have){ job = sp.jobs[gidx]; }
	}
	if(!__any_sync(0xffffffffu, have)){ return; }
	const uint64_t *slice = spec_slice(sp, gidx);
	unsigned char *slot = sp.slots + static_cast<size_t>(have ? job.slot : 0u) * sp.slot_stride;
	// behind the product rows of all warps: one slice window per read-owning lane
	unsigned char *window = smem + static_cast<size_t>(kSpecReadWarps) * lanes_per_warp * stride * sizeof(double) + (static_cast<size_t>(warp) * lanes_per_warp + (lane < lanes_per_warp ? lane : 0u)) * kSpecWindowBytes;
	uint32_t consumed = 0, rec_len = 0;
	auto draw_fn = [&](bool active, uint32_t table, uint32_t i0, uint32_t i1, uint32_t i2, uint32_t i3, double u, bool &zero) -> uint32_t {
		return coop_draw(c.tab, buf, stride, lanes_per_warp, active, table, i0, i1, i2, i3, u, zero);
	};
	auto any_fn = [](bool p) -> bool { return __any_sync(0xffffffffu, p); };
	run_read_machine<kVar>(c, sp, have, job, slice, slot, draw_fn, any_fn, consumed, rec_len, window);
	if(have){ sp.jobs[gidx].consumed = consumed; sp.jobs[gidx].rec_len = rec_len; }
}
template<bool kVar> __global__ void k_spec_reads(SimCtx c, SpecCtx sp, uint32_t stride, uint32_t lanes_per_warp, uint32_t unit_first, uint32_t unit_end);
#ifdef RSQ_READS_MINBLOCKS
#define RSQ_READS_BOUNDS kSpecReadWarps * 32, RSQ_READS_MINBLOCKS
#else
#define RSQ_READS_BOUNDS kSpecReadWarps * 32
#endif
template<> __global__ void __launch_bounds__(RSQ_READS_BOUNDS)
k_spec_reads<false>(SimCtx c, SpecCtx sp, uint32_t stride, uint32_t lanes_per_warp, uint32_t unit_first, uint32_t unit_end){ spec_reads_body<false>(c, sp, stride, lanes_per_warp, unit_first, unit_end); }
#ifndef RSQ_VAR_READS_MINBLOCKS
#define RSQ_VAR_READS_MINBLOCKS 4   // 122 registers: 4200 ms together with the scan's 4 (5 blocks / 96 registers: 4382 ms)
#endif
template<> __global__ void __launch_bounds__(kSpecReadWarps * 32, RSQ_VAR_READS_MINBLOCKS)
k_spec_reads<true>(SimCtx c, SpecCtx sp, uint32_t stride, uint32_t lanes_per_warp, uint32_t unit_first, uint32_t unit_end){ spec_reads_body<true>(c, sp, stride, lanes_per_warp, unit_first, unit_end); }

// FASTQ text of one (unit, segment): walks the unit's slab chain, a warp assembles one record at a time.
__global__ void __launch_bounds__(128)
k_spec_gather(SpecCtx sp, const unsigned long long *offsets, unsigned char *dst0, unsigned char *dst1){
	__shared__ uint32_t rec_off[33];
	const uint32_t n = sp.n_units;
	const uint32_t unit = blockIdx.x >> 1, seg = blockIdx.x & 1u;
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	unsigned char *dst = (seg ? dst1 : dst0) + offsets[static_cast<size_t>(seg) * (n + 1) + unit];
	for(uint32_t slab = sp.blocks[unit].chain_head; slab != kSpecNone; slab = sp.slab_next[slab]){
		const uint32_t count = sp.slab_count[slab];
		const unsigned char *base = sp.slots + static_cast<size_t>(slab) * 32u * sp.slot_stride;
		if(warp == 0){
			uint32_t len = 0;
			if(lane < count){
				const uint32_t *hdr = reinterpret_cast<const uint32_t *>(base + static_cast<size_t>(lane) * sp.slot_stride);
				if(hdr[2] == seg){ len = 1u + hdr[0] + 1u + hdr[1] + 3u + hdr[1] + 1u; }
			}
			uint32_t incl = len;
			for(int o = 1; o < 32; o <<= 1){ const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if(lane >= static_cast<uint32_t>(o)){ incl += v; } }
			rec_off[lane] = incl - len;
			if(lane == 31){ rec_off[32] = incl; }
		}
		__syncthreads();
		for(uint32_t k = warp; k < count; k += 4){
			const unsigned char *slot = base + static_cast<size_t>(k) * sp.slot_stride;
			const uint32_t *hdr = reinterpret_cast<const uint32_t *>(slot);
			if(hdr[2] != seg){ continue; }
			const uint32_t id_len = hdr[0], rl = hdr[1];
			const uint32_t rec = 1u + id_len + 1u + rl + 3u + rl + 1u;
			const uint32_t o_seq = 2u + id_len, o_plus = o_seq + rl, o_qual = o_plus + 3u;
			const unsigned char *id = slot + 16, *seq = slot + sp.seq_off, *qual = slot + sp.qual_off;
			unsigned char *d = dst + rec_off[k];
			for(uint32_t i = lane; i < rec; i += 32){
				unsigned char ch;
				if(i == 0){ ch = '@'; }
				else if(i < 1u + id_len){ ch = id[i - 1]; }
				else if(i < o_seq){ ch = '\n'; }
				else if(i < o_plus){ const uint32_t b = seq[i - o_seq]; ch = static_cast<unsigned char>(b > 3u ? 'N' : (0x54474341u >> (8u * b)) & 0xffu); }
				else if(i < o_qual){ ch = (i == o_plus + 1u) ? '+' : '\n'; }
				else if(i < o_qual + rl){ ch = qual[i - o_qual]; }
				else{ ch = '\n'; }
				d[i] = ch;
			}
		}
		dst += rec_off[32];
		__syncthreads();
	}
}

// Adapter-only pairs (Simulator::SimulateAdapterOnlyPairs): one stream, one warp, appended as slot `slot`.
__global__ void k_adapter_only(SimCtx c, uint64_t seed, uint32_t count, Arena arena, BlockOut *out, uint32_t slot, uint32_t max_n0){
	extern __shared__ __align__(16) unsigned char smem[];
	WarpGroup g;
	Scratch s = carve_scratch(smem, max_n0, c.max_org_len, c.max_read_len);
	DeviceSink sink; sink.init(arena);
	Mt mt; mt.s = s.mt; mt.idx = kMtN;
	mt_seed(g, mt, seed);
	uint64_t read_number = 0;
	create_reads(g, c, s, mt, sink, count, false, 0, 0, read_number, 0, 0, 0);
	if(g.lane() == 0){
		BlockOut o; o.head[0] = sink.head0; o.head[1] = sink.head1; o.bytes[0] = sink.bytes0; o.bytes[1] = sink.bytes1;
		o.pairs = sink.pairs; o.pad = 0; o.scan_draws = 0;
		out[slot] = o;
	}
}

// Exclusive prefix sums of the per-slot byte counts (one CTA; slots <= a few million).
// Device-side gzip of the FASTQ text (deflate_core.cuh): persistent CTAs, one gzip member of dfl::kMember text bytes at a time.
__global__ void __launch_bounds__(256) k_deflate_members(const uint8_t *text, uint64_t n_bytes, uint32_t n_members, uint32_t *slots, uint32_t *tokens,
                                                         uint32_t *sizes, const uint32_t *crc_tables /*[256 + 32]*/){
	extern __shared__ __align__(16) unsigned char dfl_smem[];
	dfl::Shared &sh = *reinterpret_cast<dfl::Shared *>(dfl_smem);
	const dfl::DeviceCta cta;
	for(uint32_t m = blockIdx.x; m < n_members; m += gridDim.x){
		const uint64_t off = static_cast<uint64_t>(m) * dfl::kMember;
		const uint32_t n = static_cast<uint32_t>(n_bytes - off < dfl::kMember ? n_bytes - off : dfl::kMember);
		const uint32_t bytes = dfl::deflate_member(cta, sh, text + off, n, slots + static_cast<size_t>(m) * dfl::kSlotWords,
		                                           tokens + static_cast<size_t>(blockIdx.x) * dfl::kMember, crc_tables, crc_tables + 256);
		if(threadIdx.x == 0){ sizes[m] = bytes; }
	}
}
// Members behind each other: CTA m copies its slot to the sum of the sizes in front of it.
__global__ void __launch_bounds__(256) k_deflate_compact(const uint32_t *slots, const uint32_t *sizes, uint32_t n_members, uint8_t *out, unsigned long long *total){
	__shared__ unsigned long long offset;
	const uint32_t m = blockIdx.x;
	if(threadIdx.x == 0){
		unsigned long long o = 0;
		for(uint32_t i = 0; i < m; ++i){ o += sizes[i]; }
		offset = o;
		if(m + 1 == n_members){ *total = o + sizes[m]; }
	}
	__syncthreads();
	const uint32_t n = sizes[m];
	const uint8_t *src = reinterpret_cast<const uint8_t *>(slots + static_cast<size_t>(m) * dfl::kSlotWords);
	uint8_t *dst = out + offset;
	const uint32_t head = static_cast<uint32_t>((4u - (offset & 3u)) & 3u) < n ? static_cast<uint32_t>((4u - (offset & 3u)) & 3u) : n;
	for(uint32_t i = threadIdx.x; i < head; i += blockDim.x){ dst[i] = src[i]; }
	// aligned 32-bit stores; the source is read bytewise (its alignment differs from the destination's)
	const uint32_t words = (n - head) / 4u;
	for(uint32_t w = threadIdx.x; w < words; w += blockDim.x){
		const uint8_t *q = src + head + 4u * w;
		reinterpret_cast<uint32_t *>(dst + head)[w] = q[0] | (static_cast<uint32_t>(q[1]) << 8) | (static_cast<uint32_t>(q[2]) << 16) | (static_cast<uint32_t>(q[3]) << 24);
	}
	for(uint32_t i = head + 4u * words + threadIdx.x; i < n; i += blockDim.x){ dst[i] = src[i]; }
}

__global__ void k_block_offsets(const BlockOut *out, uint32_t n, unsigned long long *offsets /*[2][n+1]*/, unsigned long long *totals /*[4]: bytes0, bytes1, pairs, draws*/){
	__shared__ unsigned long long part[4][1024];
	const uint32_t t = threadIdx.x, nt = blockDim.x;
	const uint32_t per = (n + nt - 1) / nt;
	const uint32_t lo = min(n, t * per), hi = min(n, lo + per);
	unsigned long long s0 = 0, s1 = 0, sp = 0, sd = 0;
	for(uint32_t i = lo; i < hi; ++i){ s0 += out[i].bytes[0]; s1 += out[i].bytes[1]; sp += out[i].pairs; sd += out[i].scan_draws; }
	part[0][t] = s0; part[1][t] = s1; part[2][t] = sp; part[3][t] = sd;
	__syncthreads();
	if(t == 0){
		unsigned long long a0 = 0, a1 = 0, ap = 0, ad = 0;
		for(uint32_t k = 0; k < nt; ++k){
			unsigned long long v0 = part[0][k], v1 = part[1][k];
			part[0][k] = a0; part[1][k] = a1; a0 += v0; a1 += v1; ap += part[2][k]; ad += part[3][k];
		}
		totals[0] = a0; totals[1] = a1; totals[2] = ap; totals[3] = ad;
		offsets[n] = a0; offsets[(n + 1) + n] = a1;
	}
	__syncthreads();
	unsigned long long a0 = part[0][t], a1 = part[1][t];
	for(uint32_t i = lo; i < hi; ++i){
		offsets[i] = a0; offsets[(n + 1) + i] = a1;
		a0 += out[i].bytes[0]; a1 += out[i].bytes[1];
	}
}

__global__ void k_gather(const BlockOut *out, uint32_t n, Arena arena, const unsigned long long *offsets, unsigned char *dst0, unsigned char *dst1){
	const uint32_t slot = blockIdx.x >> 1, seg = blockIdx.x & 1u;
	if(slot >= n){ return; }
	unsigned char *dst = (seg ? dst1 : dst0) + offsets[static_cast<size_t>(seg) * (n + 1) + slot];
	uint32_t chunk = out[slot].head[seg];
	while(chunk != kNone){
		const uint32_t used = arena.chunk_used[chunk];
		const unsigned char *src = arena.data + static_cast<size_t>(chunk) * arena.chunk_bytes;
		for(uint32_t i = threadIdx.x; i < used; i += blockDim.x){ dst[i] = src[i]; }
		dst += used;
		chunk = arena.chunk_next[chunk];
	}
}

// seqToIllumina: one warp per batch of records (Simulator::ErrorModelOnlyThread + ApplyErrorsAndQualityToFastaInput)

struct EmSink {
	DeviceSink inner;
	__device__ void write_record(const WarpGroup &g, uint32_t, const char *id, int id_len, const uint8_t *seq, const uint8_t *qual, uint32_t n){ inner.write_record(g, 0, id, id_len, seq, qual, n); }
};

__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_error_model(SimCtx c, const EmRecord *recs, uint32_t n_recs, uint32_t batch_size, const uint64_t *batch_seeds, uint32_t n_batches,
              const uint8_t *seqs, const uint8_t *sdom, const uint8_t *srate, const char *ids, Arena arena, BlockOut *out, uint32_t *next_batch,
              uint32_t max_n0, uint32_t scratch_per_warp){
	extern __shared__ __align__(16) unsigned char smem[];
	WarpGroup g;
	const uint32_t warp = threadIdx.x >> 5;
	Scratch s = carve_scratch(smem + static_cast<size_t>(warp) * scratch_per_warp, max_n0, c.max_org_len, c.max_read_len);
	while(true){
		uint32_t bi = 0;
		if(g.lane() == 0){ bi = atomicAdd(next_batch, 1u); }
		bi = __shfl_sync(0xffffffffu, bi, 0);
		if(bi >= n_batches){ break; }
		DeviceSink sink; sink.init(arena);
		Mt mt; mt.s = s.mt; mt.idx = kMtN;
		mt_seed(g, mt, batch_seeds[bi]);
		const uint32_t lo = bi * batch_size, hi = min(n_recs, lo + batch_size);
		for(uint32_t r = lo; r < hi; ++r){
			const EmRecord rec = recs[r];
			uint32_t org_len = rec.len;
			if(org_len > c.max_org_len){ org_len = c.max_org_len; if(g.lane() == 0){ atomicOr(c.error_flag, kErrOrgOverflow); } }
			g.sync();
			for(uint32_t i = g.lane(); i < org_len; i += 32){
				s.org[i] = seqs[rec.seq_off + i]; s.sdom[i] = sdom[rec.seq_off + i]; s.srate[i] = srate[rec.seq_off + i];
			}
			g.sync();
			uint32_t tile = 0;
			if(1 < c.num_tiles){ tile = discrete_draw(g, mt, c.tile_pick); }
			ReadState par;
			fill_read(g, c, s, mt, par, rec.seg, tile, rec.fragment_length, org_len);
			// id + " " + cigar + " E" + errors
			int n = 0;
			n = put_str(g, s.id, n, kIdCap, ids + rec.id_off, rec.id_len);
			n = put_char(g, s.id, n, kIdCap, ' ');
			n = put_str(g, s.id, n, kIdCap, s.cigar, par.cigar_len < kCigarCap ? par.cigar_len : kCigarCap);
			n = put_str(g, s.id, n, kIdCap, " E", 2);
			n = put_uint(g, s.id, n, kIdCap, par.num_errors);
			if(n > kIdCap){ if(g.lane() == 0){ atomicOr(c.error_flag, kErrRecordTooLong); } n = kIdCap; }
			g.sync();
			sink.write_record(g, 0, s.id, n, s.seq, s.qual, par.read_length);
			sink.pair_done(g);
		}
		if(g.lane() == 0){
			BlockOut o; o.head[0] = sink.head0; o.head[1] = kNone; o.bytes[0] = sink.bytes0; o.bytes[1] = 0; o.pairs = sink.pairs; o.pad = 0; o.scan_draws = 0;
			out[bi] = o;
		}
	}
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
// Large host buffer that is neither pinned nor zero-filled (either costs longer than a run at this size); 2 MB pages where the kernel grants them.
struct HostBig {
	char *p = nullptr; size_t cap = 0;
	HostBig() = default; HostBig(const HostBig &) = delete; HostBig &operator=(const HostBig &) = delete;
	~HostBig(){ free(p); }
	char *data(){ return p; } const char *data() const { return p; } size_t size() const { return cap; }
	void resize(size_t n){   // keeps the old content
		if(n <= cap){ return; }
		void *q = nullptr;
		if(posix_memalign(&q, 2u << 20, (n + (2u << 20) - 1) & ~static_cast<size_t>((2u << 20) - 1))){ throw std::runtime_error("out of host memory for the FASTQ text"); }
		madvise(q, n, MADV_HUGEPAGE);
		if(p){ std::memcpy(q, p, cap); free(p); }
		p = static_cast<char *>(q); cap = n;
	}
};

struct PinnedBuf {
	char *p = nullptr; size_t cap = 0;
	~PinnedBuf(){ if(p){ cudaFreeHost(p); } }
	void ensure(size_t n){ if(n > cap){ if(p){ cudaFreeHost(p); p = nullptr; } RSQ_CUDA(cudaMallocHost(&p, n ? n : 1)); cap = n; } }
};

struct EventTimer {
	cudaEvent_t a, b; cudaStream_t s;
	explicit EventTimer(cudaStream_t st) : s(st){ cudaEventCreate(&a); cudaEventCreate(&b); }
	~EventTimer(){ cudaEventDestroy(a); cudaEventDestroy(b); }
	void start(){ cudaEventRecord(a, s); }
	float stop(){ cudaEventRecord(b, s); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }
};

}  // namespace rsq

using namespace rsq;

struct rsq_profile { Profile p; };
struct rsq_reference { Genome g; };

struct rsq_engine {
	int device = 0;
	cudaStream_t stream = nullptr;
	cudaStream_t stream2 = nullptr;        // bias sums overlap with the master stream / systematic errors
	cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
	PinnedBuf h_bias_results;
	Profile prof;
	uint32_t max_n0 = 0;
	uint32_t launches = 0;
	// tables + profile on device
	DevBuf<TableDesc> d_desc; DevBuf<double> d_blob; DevBuf<uint32_t> d_par0;
	DevBuf<double> d_cp; DevBuf<Discrete> d_start_cut[2]; DevBuf<uint32_t> d_start_cut_from[2], d_adapter_off[2];
	DevBuf<uint8_t> d_adapter_seq, d_adapter_sys;
	DevBuf<double> d_il_bias, d_gc_bias, d_ref_seq_bias, d_sur_tab[3];
	DevBuf<uint64_t> d_insert_lengths;
	DevBuf<uint16_t> d_tile_names;
	DevBuf<uint32_t> d_rlbf_row_from[2], d_rlbf_row_off[2]; DevBuf<uint64_t> d_rlbf_val[2];
	DevBuf<char> d_base_id;
	DevBuf<uint32_t> d_error_flag;
	std::vector<uint32_t> h_adapter_off[2];
	std::vector<uint8_t> h_adapter_seq;
	SimCtx ctx{};
	// run state
	Genome genome;
	std::vector<double> run_ref_seq_bias;
	Normalization norm;
	DevBuf<uint64_t> d_seq_off; DevBuf<uint32_t> d_seq_len, d_gc_prefix, d_name_off, d_cov_group;
	DevBuf<uint8_t> d_ref, d_sys_fwd, d_sys_rev;
	DevBuf<double> d_sur_start, d_sur_end, d_thr, d_binom_p0;
	DevBuf<BiasChain> d_bias_chains; DevBuf<double> d_bias_plain, d_bias_start, d_bias_out, d_bias_cmax; DevBuf<uint32_t> d_bias_tie, d_bias_reran;   // chunked SumBias (ordered_sum.cuh)
	DevBuf<uint64_t> d_thr_int; DevBuf<uint32_t> d_thr_hi;
	DevBuf<char> d_names;
	DevBuf<uint64_t> d_master_state, d_master, d_jump_states, d_jump_poly, d_jump_seq, d_jump_scratch;
	DevBuf<BlockDesc> d_blocks;
	DevBuf<SysChain> d_sys_chains; DevBuf<SysChunk> d_sys_chunks; DevBuf<uint32_t> d_sys_dirty, d_gc_tiles, d_sys_checkpoints;
	DevBuf<BiasParamDev> d_bias_params; DevBuf<double> d_bias_sums, d_bias_max;
	DevBuf<uint32_t> d_meth_off, d_meth_start, d_meth_end; DevBuf<double> d_meth_rate; DevBuf<int32_t> d_block_meth;
	// variants (Reference::variants_ flattened, SimBlock::first_variant_id_, SysErrorVariant::var_errors_ of both strands)
	DevBuf<uint32_t> d_var_seq_first, d_var_position, d_var_bases_off, d_var_block_first, d_var_block_first_off, d_var_bstate;
	DevBuf<uint8_t> d_var_bases, d_var_errs_fwd, d_var_errs_rev, d_var_ctx_fwd, d_var_ctx_rev;
	DevBuf<uint64_t> d_var_allele_lo, d_var_allele_hi, d_var_blk_off;
	DevBuf<double> d_binom_pow;
	DevBuf<uint16_t> d_spec_chosen;
	uint32_t num_alleles = 1; bool with_var = false;
	// multi-GPU group (rsq_engine_join_group): this engine is rank group_rank of group_world engines of ONE run; prepare() then only uploads and
	// prepares the sequences its shard holds blocks in, the per-(sequence, length) bias sums are computed by the sequence's owner and all-reduced
	ncclComm_t comm = nullptr; int group_rank = 0, group_world = 1;
	DevBuf<double> d_group_scratch;
	uint64_t group_pairs = 0;
	PinnedBuf h_ref_stage;
	std::vector<uint64_t> h_seq_off;
	uint64_t total_size = 0;
	uint32_t n_blocks_total = 0, n_blocks_sim = 0, shard_first = 0, shard_n = 0, spec_units_last = 0;
	rsq::ShardPlan plan;   // blocks of every sequence (shard_plan.hpp)
	uint64_t sur_window_max = 0;
	bool sur_windowed = false;   // simulate(): the per-position surrounding biases are held per batch instead of for the whole reference
	bool shard_has_adapter_only = false;
	uint64_t adapter_only_seed = 0;
	uint64_t total_pairs = 0, adapter_only_pairs = 0;
	uint32_t sys_gc_range = 0, syserr_passes = 0;
	bool prepared = false;
	// output
	DevBuf<unsigned char> d_arena;
	DevBuf<uint32_t> d_chunk_next, d_chunk_used, d_next_free, d_next_block;
	DevBuf<BlockOut> d_block_out;
	DevBuf<unsigned long long> d_offsets, d_totals;
	uint64_t out_bytes[2] = {0, 0};
	uint64_t out_pairs = 0, out_draws = 0;
	PinnedBuf h_out[2];
	bool downloaded = false;
	// batched output: device text of two batches in flight, copy stream, optional file sink (rsq_simulate)
	DevBuf<unsigned char> d_out_batch[2][2];
	PinnedBuf h_ring[4]; cudaEvent_t ev_ring[4] = {nullptr, nullptr, nullptr, nullptr};
	HostBig h_big[2];   // text of runs that arrive in several batches (ordinary memory: pinning tens of GB takes longer than the run)
	cudaStream_t copy_stream = nullptr, copy_stream2 = nullptr; cudaEvent_t ev_out[2] = {nullptr, nullptr};
	// device-side gzip (.gz sinks unless RSQ_GZIP=host): per writer a slot per member, token scratch per CTA, the compacted members
	DevBuf<uint32_t> d_dfl_slots[2], d_dfl_tokens[2], d_dfl_sizes[2], d_dfl_crc; DevBuf<uint8_t> d_dfl_out[2]; DevBuf<unsigned long long> d_dfl_total[2];
	TextSink *sink_files[2] = {nullptr, nullptr};   // rsq_simulate: the two FASTQ files (plain or gzip by name)
	bool streamed_to_host = false; int last_par = 0;
	double reusable_bytes() const;
	// speculative two-phase path
	uint32_t max_n0_reads = 0;             // largest candidate count of the tables FillRead draws from
	uint32_t max_name_len = 0, em_max_id_len = 0;
	DevBuf<SpecBlock> d_spec_blocks; DevBuf<SpecSnap> d_spec_snaps; DevBuf<ReadJob> d_spec_jobs;
	DevBuf<uint64_t> d_spec_words; DevBuf<unsigned char> d_spec_slots; DevBuf<uint8_t> d_spec_conv; DevBuf<uint32_t> d_slab_next, d_slab_count, d_spec_counters;
	PinnedBuf h_spec_counters;
	uint32_t spec_rounds = 0, spec_depth = 0;
	std::vector<cudaStream_t> spec_streams; std::vector<cudaEvent_t> spec_events;

	~rsq_engine(){ if(comm && nccl_api().ok){ nccl_api().CommDestroy(comm); } for(int i = 0; i < 4; ++i){ if(ev_ring[i]){ cudaEventDestroy(ev_ring[i]); } } for(int i = 0; i < 2; ++i){ if(ev_out[i]){ cudaEventDestroy(ev_out[i]); } } if(copy_stream){ cudaStreamDestroy(copy_stream); } if(copy_stream2){ cudaStreamDestroy(copy_stream2); } for(auto ev : spec_events){ cudaEventDestroy(ev); } for(auto st : spec_streams){ cudaStreamDestroy(st); } if(ev_fork){ cudaEventDestroy(ev_fork); } if(ev_join){ cudaEventDestroy(ev_join); } if(stream2){ cudaStreamDestroy(stream2); } if(stream){ cudaStreamDestroy(stream); } }
};

double rsq_engine::reusable_bytes() const {
	double b = 0;
	b += d_spec_blocks.cap * sizeof(rsq::SpecBlock) + d_spec_snaps.cap * sizeof(rsq::SpecSnap) + d_spec_jobs.cap * sizeof(rsq::ReadJob) + d_spec_words.cap * 8.0;
	b += d_spec_slots.cap + (d_slab_next.cap + d_slab_count.cap) * 4.0 + d_arena.cap + (d_chunk_next.cap + d_chunk_used.cap) * 4.0;
	for(int i = 0; i < 2; ++i){ for(int j = 0; j < 2; ++j){ b += d_out_batch[i][j].cap; } }
	return b;
}

namespace rsq {

static const uint32_t kChunkBytes = 4096;

static void upload_profile(rsq_engine &e){
	const Profile &p = e.prof;
	cudaStream_t s = e.stream;
	std::vector<TableDesc> desc; std::vector<double> blob; std::vector<uint32_t> par0;
	e.max_n0 = 0;
	for(const auto &h : p.tables){
		TableDesc d{}; d.n0 = h.par0.size(); d.nm = h.nm;
		d.stride = (d.n0 + 1u) & ~1u;   // rows of 16-byte pairs: the cooperative product loads two candidates per instruction (half the L1 wavefronts)
		for(uint32_t n = 0; n < h.nm; ++n){
			d.from[n] = h.from[n]; d.span[n] = h.to[n] - h.from[n]; d.off[n] = blob.size();
			for(uint32_t r = 0; r < d.span[n]; ++r){
				blob.insert(blob.end(), h.dim2[n].begin() + static_cast<size_t>(r) * d.n0, h.dim2[n].begin() + static_cast<size_t>(r + 1) * d.n0);
				if(d.stride != d.n0){ blob.push_back(0.0); }
			}
		}
		d.par0_off = par0.size(); par0.insert(par0.end(), h.par0.begin(), h.par0.end());
		desc.push_back(d);
		e.max_n0 = std::max(e.max_n0, d.n0);
	}
	e.max_n0_reads = 0;
	for(size_t i = 0; i < desc.size(); ++i){   // every family but dom_error_result_ / error_rate_result_ (systematic errors only)
		if(i < 50u * p.num_tiles || i >= 50u * p.num_tiles + 120u){ e.max_n0_reads = std::max(e.max_n0_reads, desc[i].n0); }
	}
	e.d_desc.upload(desc, s); e.d_blob.upload(blob, s); e.d_par0.upload(par0, s);
	SimCtx &c = e.ctx;
	const uint32_t T = p.num_tiles;
	c.tab.desc = e.d_desc.p; c.tab.blob = e.d_blob.p; c.tab.par0 = e.d_par0.p; c.tab.num_tiles = T;
	c.tab.quality_base = 0; c.tab.seqq_base = 8 * T; c.tab.basecall_base = 10 * T; c.tab.domerr_base = 50 * T;
	c.tab.errrate_base = 50 * T + 100; c.tab.indel_base = 50 * T + 120;
	c.phred_offset = p.phred_quality_offset; c.max_len_deletion = p.max_len_deletion;
	c.insert_from = std::max<uint64_t>(1, p.insert_lengths.from); c.insert_to = p.insert_lengths.to();

	// discrete distributions: all cumulative arrays in one device buffer
	std::vector<double> cps;
	struct Pending { Discrete *target; size_t off; uint32_t n; };
	std::vector<Discrete> sc[2];
	auto add = [&](const std::vector<double> &cp) -> std::pair<size_t, uint32_t> { size_t off = cps.size(); cps.insert(cps.end(), cp.begin(), cp.end()); return {off, static_cast<uint32_t>(cp.size())}; };
	auto tile_cp = add(discrete_cp(p.tile_abundance.begin(), p.tile_abundance.end()));
	auto polya_cp = add(discrete_cp(p.polya_tail_length.v.begin(), p.polya_tail_length.v.end()));
	auto overrun_cp = add(discrete_cp(p.overrun_bases.begin(), p.overrun_bases.end() - 1));
	std::pair<size_t, uint32_t> pick_cp[2];
	std::vector<std::pair<size_t, uint32_t>> sc_cp[2];
	std::vector<uint32_t> sc_from[2];
	e.h_adapter_seq.clear();
	for(int seg = 0; seg < 2; ++seg){
		pick_cp[seg] = add(discrete_cp(p.adapter_significant_count[seg].begin(), p.adapter_significant_count[seg].end()));
		e.h_adapter_off[seg].clear();
		e.h_adapter_off[seg].push_back(e.h_adapter_seq.size());
		for(size_t a = 0; a < p.adapter_seqs[seg].size(); ++a){
			for(char ch : p.adapter_seqs[seg][a]){ e.h_adapter_seq.push_back(Genome::code(ch) & 3); }
			e.h_adapter_off[seg].push_back(e.h_adapter_seq.size());
			sc_cp[seg].push_back(add(discrete_cp(p.adapter_start_cut[seg][a].v.begin(), p.adapter_start_cut[seg][a].v.end())));
			sc_from[seg].push_back(p.adapter_start_cut[seg][a].from);
		}
	}
	e.d_cp.upload(cps, s);
	auto mk = [&](std::pair<size_t, uint32_t> x){ Discrete d; d.cp = e.d_cp.p + x.first; d.n = x.second; return d; };
	c.tile_pick = mk(tile_cp); c.polya_pick = mk(polya_cp); c.overrun_pick = mk(overrun_cp);
	c.polya_from = p.polya_tail_length.from;
	e.d_adapter_seq.upload(e.h_adapter_seq, s);
	e.d_adapter_sys.alloc(2 * e.h_adapter_seq.size() + 2); e.d_adapter_sys.zero(s);
	for(int seg = 0; seg < 2; ++seg){
		for(auto x : sc_cp[seg]){ sc[seg].push_back(mk(x)); }
		e.d_start_cut[seg].upload(sc[seg], s); e.d_start_cut_from[seg].upload(sc_from[seg], s); e.d_adapter_off[seg].upload(e.h_adapter_off[seg], s);
		c.adapters[seg].n = p.adapter_seqs[seg].size(); c.adapters[seg].off = e.d_adapter_off[seg].p; c.adapters[seg].pick = mk(pick_cp[seg]);
		c.adapters[seg].start_cut = e.d_start_cut[seg].p; c.adapters[seg].start_cut_from = e.d_start_cut_from[seg].p;
		c.read_len_from[seg] = p.read_lengths[seg].from; c.read_len_to[seg] = p.read_lengths[seg].to(); c.read_len_count[seg] = p.read_lengths[seg].size();
		const auto &rl = p.read_lengths_by_fragment_length[seg];
		c.rlbf_from[seg] = rl.from; c.rlbf_to[seg] = rl.to();
		std::vector<uint32_t> row_from, row_off{0}; std::vector<uint64_t> vals;
		for(const auto &row : rl.v){ row_from.push_back(row.from); vals.insert(vals.end(), row.v.begin(), row.v.end()); row_off.push_back(vals.size()); }
		e.d_rlbf_row_from[seg].upload(row_from, s); e.d_rlbf_row_off[seg].upload(row_off, s); e.d_rlbf_val[seg].upload(vals, s);
		c.rlbf_row_from[seg] = e.d_rlbf_row_from[seg].p; c.rlbf_row_off[seg] = e.d_rlbf_row_off[seg].p; c.rlbf_val[seg] = e.d_rlbf_val[seg].p;
	}
	c.adapter_seq = e.d_adapter_seq.p; c.adapter_sys = e.d_adapter_sys.p;
	std::vector<uint64_t> il(c.insert_to, 0); std::vector<double> ilb(c.insert_to, 0.0), gcb(101, 0.0);
	for(uint32_t i = 0; i < c.insert_to; ++i){ il[i] = p.insert_lengths[i]; ilb[i] = p.insert_lengths_bias[i]; }
	for(uint32_t i = 0; i < 101; ++i){ gcb[i] = p.gc_fragment_content_bias[i]; }
	e.d_insert_lengths.upload(il, s); e.d_il_bias.upload(ilb, s); e.d_gc_bias.upload(gcb, s);
	c.insert_lengths = e.d_insert_lengths.p; c.il_bias = e.d_il_bias.p; c.gc_bias = e.d_gc_bias.p;
	for(int b = 0; b < 3; ++b){ e.d_sur_tab[b].upload(p.fragment_surroundings_bias[b], s); }
	c.num_tiles = p.tiles.size(); e.d_tile_names.upload(p.tiles, s); c.tile_names = e.d_tile_names.p;
	c.disp_a = p.dispersion_parameters[0]; c.disp_b = p.dispersion_parameters[1];
	uint32_t max_adapter = 0;
	for(int seg = 0; seg < 2; ++seg){ for(const auto &a : p.adapter_seqs[seg]){ max_adapter = std::max<uint32_t>(max_adapter, a.size()); } }
	const uint32_t need_read = std::max(c.read_len_to[0], c.read_len_to[1]);
	if(e.max_n0 > kMaxN0 || need_read > kMaxReadLen || std::max(need_read + c.max_len_deletion, max_adapter) > kMaxOrgLen){
		throw std::runtime_error("profile exceeds the engine's compile-time limits (candidates per table <= " + std::to_string(kMaxN0) + ", read length < " +
		                         std::to_string(kMaxReadLen) + ", read + longest deletion / adapter length <= " + std::to_string(kMaxOrgLen) + ")");
	}
	c.max_read_len = kMaxReadLen;   // capacities of the per-warp scratch (sim_core.cuh)
	c.max_org_len = kMaxOrgLen;
	e.d_error_flag.alloc(1); e.d_error_flag.zero(s);
	c.error_flag = e.d_error_flag.p;
	RSQ_CUDA(cudaStreamSynchronize(s));
}

static uint32_t read_error_flag(rsq_engine &e){
	uint32_t f = 0;
	RSQ_CUDA(cudaMemcpyAsync(&f, e.d_error_flag.p, 4, cudaMemcpyDeviceToHost, e.stream));
	RSQ_CUDA(cudaStreamSynchronize(e.stream));
	return f;
}

static std::string describe_flag(uint32_t f){
	std::string s;
	if(f & kErrCigarOverflow){ s += "CIGAR text exceeds the per-read buffer; "; }
	if(f & kErrCountRunaway){ s += "fragment count did not terminate (uintDupCount overflow); "; }
	if(f & kErrArenaFull){ s += "output arena exhausted; "; }
	if(f & kErrOrgOverflow){ s += "read or adapter longer than the staged buffers; "; }
	if(f & kErrRecordTooLong){ s += "FASTQ record longer than an output chunk; "; }
	if(f & kErrReferenceOutOfRange){ s += "the reference implementation indexes its methylation regions out of range for this input (std::out_of_range in Simulator::CTConversion); "; }
	return s;
}

// Appends n outputs of the master stream to `out` and leaves its state behind them.  Long runs are split into segments of
// J = 2^k words whose start states come from jump-ahead (k_master_jump, one after the other: ~0.1 ms each) and which are then
// generated by one CTA each at the same time; short runs and the remainder use the serial recurrence.
static void master_generate(rsq_engine &e, uint64_t *out, uint64_t n){
	cudaStream_t s = e.stream;
	uint64_t min_words = 1ull << 20;
	if(const char *env = getenv("RSQ_MASTER_JUMP_MIN")){ min_words = std::max<long long>(1 << 16, atoll(env)); }
	if(n < min_words + 1024){
		k_master_stream<<<1, kMasterThreads, 0, s>>>(e.d_master_state.p, e.d_master_state.p, out, n); ++e.launches;
		return;
	}
	// the first words come from the serial recurrence: afterwards the state window holds generated words only (the jump
	// identity does not cover the unused low bits of the very first seed word)
	const uint64_t head = 1024;
	k_master_stream<<<1, kMasterThreads, 0, s>>>(e.d_master_state.p, e.d_master_state.p, out, head); ++e.launches;
	out += head; n -= head;
	// segment length: a jump costs ~40 us, a segment of J words ~J * 0.83 ns; tables exist for 2^16, 2^19, 2^22
	int table = 0; double best = -1.0;
	for(int t = 0; t < kMtJumpTables; ++t){
		const double J = static_cast<double>(1ull << kMtJumpLog2[t]);
		if(J > static_cast<double>(n)){ continue; }
		const double cost = std::floor(n / J) * 40.0 + J * 0.00083;
		if(best < 0.0 || cost < best){ best = cost; table = t; }
	}
	const uint64_t J = 1ull << kMtJumpLog2[table];
	const uint64_t segments = n / J;
	if(!e.d_jump_poly.p){
		std::vector<uint64_t> polys(static_cast<size_t>(kMtJumpTables) * kMtN);
		std::memcpy(polys.data(), kMtJumpPoly, polys.size() * 8);
		e.d_jump_poly.upload(polys, s);
		RSQ_CUDA(cudaFuncSetAttribute(k_master_jump_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, kJumpSeqWords * 8));
		e.d_jump_seq.alloc(kJumpSeqWords);
	}
	e.d_jump_states.alloc((segments + 1) * kMasterStateWords);
	RSQ_CUDA(cudaMemcpyAsync(e.d_jump_states.p, e.d_master_state.p, kMasterStateWords * 8, cudaMemcpyDeviceToDevice, s));
	for(uint64_t g = 0; g < segments; ++g){
		k_master_jump_gen<<<1, kMasterThreads, kJumpSeqWords * 8, s>>>(e.d_jump_states.p + g * kMasterStateWords, e.d_jump_seq.p, e.d_jump_states.p + (g + 1) * kMasterStateWords);
		k_master_jump_xor<<<(kMtN + kJumpPolyWordsPerCta - 1) / kJumpPolyWordsPerCta, 320, 0, s>>>(e.d_jump_seq.p, e.d_jump_poly.p + static_cast<size_t>(table) * kMtN,
		                                                                                          e.d_jump_states.p + (g + 1) * kMasterStateWords);
	}
	k_master_stream<<<static_cast<unsigned>(segments), kMasterThreads, 0, s>>>(e.d_jump_states.p, nullptr, out, J);
	e.launches += 2 * static_cast<uint32_t>(segments) + 1;
	// remainder from the state behind the last segment
	RSQ_CUDA(cudaMemcpyAsync(e.d_master_state.p, e.d_jump_states.p + segments * kMasterStateWords, kMasterStateWords * 8, cudaMemcpyDeviceToDevice, s));
	k_master_stream<<<1, kMasterThreads, 0, s>>>(e.d_master_state.p, e.d_master_state.p, out + segments * J, n - segments * J); ++e.launches;
}

// Advances the master stream by n outputs without producing them: the remainder below 1024 (and the first 1024, so that the
// window holds generated words) by the serial recurrence into scratch, the rest bit by bit with the jump polynomials for 2^10 .. 2^36.
static void master_skip(rsq_engine &e, uint64_t n){
	cudaStream_t s = e.stream;
	if(!n){ return; }
	uint64_t serial = n % 1024;
	if(n >= 1024){ serial += 1024; }
	if(serial){
		e.d_jump_scratch.alloc(2048);
		k_master_stream<<<1, kMasterThreads, 0, s>>>(e.d_master_state.p, e.d_master_state.p, e.d_jump_scratch.p, serial); ++e.launches;
	}
	uint64_t rest = n - serial;
	if(!rest){ return; }
	if(rest >> (kMtJumpLog2[kMtJumpTables - 1] + 1)){ throw std::runtime_error("master stream skip beyond the largest jump polynomial"); }
	if(!e.d_jump_poly.p){
		std::vector<uint64_t> polys(static_cast<size_t>(kMtJumpTables) * kMtN);
		std::memcpy(polys.data(), kMtJumpPoly, polys.size() * 8);
		e.d_jump_poly.upload(polys, s);
		RSQ_CUDA(cudaFuncSetAttribute(k_master_jump_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, kJumpSeqWords * 8));
		e.d_jump_seq.alloc(kJumpSeqWords);
	}
	e.d_jump_states.alloc(2 * kMasterStateWords);
	for(int t = 0; t < kMtJumpTables; ++t){
		if(!((rest >> kMtJumpLog2[t]) & 1ull)){ continue; }
		k_master_jump_gen<<<1, kMasterThreads, kJumpSeqWords * 8, s>>>(e.d_master_state.p, e.d_jump_seq.p, e.d_jump_states.p);
		k_master_jump_xor<<<(kMtN + kJumpPolyWordsPerCta - 1) / kJumpPolyWordsPerCta, 320, 0, s>>>(e.d_jump_seq.p, e.d_jump_poly.p + static_cast<size_t>(t) * kMtN, e.d_jump_states.p);
		RSQ_CUDA(cudaMemcpyAsync(e.d_master_state.p, e.d_jump_states.p, kMasterStateWords * 8, cudaMemcpyDeviceToDevice, s));
		e.launches += 2;
	}
}

// Systematic errors of a set of chains: speculative chunks + exact fix-up passes.
static void run_sys_chains(rsq_engine &e, const std::vector<SysChain> &chains, const std::vector<std::pair<uint32_t, uint32_t>> &chain_len_known,
                           uint32_t chunk_len, uint32_t warmup, uint32_t &passes){
	{
		// a pass lasts as long as one chunk: short chunks for small genomes (enough of them to fill the machine), long ones
		// (less warm-up overhead) once there are plenty
		uint64_t total = 0; for(const auto &cl : chain_len_known){ total += cl.first; }
		int dev_sms = 0; RSQ_CUDA(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, e.device));
		const uint64_t want = total / (static_cast<uint64_t>(dev_sms) * 64) + 1;   // E. coli: chunks of 1024 (18.2 ms for the phase, five passes) against 2048 (22.3 ms, three passes)
		uint32_t len = 1024; while(len < chunk_len && len < want){ len *= 2; }
		if(const char *env = getenv("RSQ_SYS_CHUNK")){ len = std::max(256, atoi(env)); }
		chunk_len = len; warmup = std::min(warmup, std::max<uint32_t>(len / 2, 512));
		if(const char *env = getenv("RSQ_SYS_WARMUP")){ warmup = atoi(env); }
	}
	std::vector<SysChunk> chunks;
	for(uint32_t ci = 0; ci < chains.size(); ++ci){
		const uint32_t L = chain_len_known[ci].first;
		for(uint32_t b = 0; b < L; b += chunk_len){
			SysChunk k{}; k.chain = ci; k.begin = b; k.end = std::min(L, b + chunk_len);
			k.warm_from = b > warmup ? b - warmup : 0; if(b == 0){ k.warm_from = 0; }
			k.dirty = 1; k.first_of_chain = (b == 0);
			chunks.push_back(k);
		}
	}
	if(chunks.empty()){ return; }
	DevBuf<SysChain> &d_chains = e.d_sys_chains; d_chains.upload(chains, e.stream);
	DevBuf<SysChunk> &d_chunks = e.d_sys_chunks; d_chunks.upload(chunks, e.stream);
	DevBuf<uint32_t> &d_dirty = e.d_sys_dirty; d_dirty.alloc(1);
	const uint32_t cp_per_chunk = chunk_len / kSysCheckpoint + 1;
	e.d_sys_checkpoints.alloc(static_cast<size_t>(chunks.size()) * cp_per_chunk);
	const uint32_t n = chunks.size();
	// one lane per chunk (lock step, likelihood products shared by the warp); RSQ_SYS_PATH=warp keeps the one-warp-per-chunk kernel
	const bool lanes_path = !(getenv("RSQ_SYS_PATH") && std::string(getenv("RSQ_SYS_PATH")) == "warp");
	const int warps = 4;
	const uint32_t lanes = n <= 2368u * 8u ? 8u : 16u;
	const uint32_t stride = ((e.max_n0 + 3u) & ~3u) + 1u;
	const size_t shmem = lanes_path ? static_cast<size_t>(warps) * lanes * stride * 8 : static_cast<size_t>(warps) * ((e.max_n0 + 3) & ~3u) * 8;
	if(lanes_path){ RSQ_CUDA(cudaFuncSetAttribute(k_sys_chunks_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(shmem))); }
	uint32_t dirty = n;
	passes = 0;
	while(dirty){
		if(lanes_path){
			const uint32_t per_cta = warps * lanes;
			k_sys_chunks_lanes<<<(n + per_cta - 1) / per_cta, warps * 32, shmem, e.stream>>>(e.ctx.tab, d_chains.p, d_chunks.p, n, stride, lanes, e.sys_gc_range, e.prof.reset_distance, e.d_sys_checkpoints.p, cp_per_chunk);
		}
		else{
			k_sys_chunks<<<(n + warps - 1) / warps, warps * 32, shmem, e.stream>>>(e.ctx.tab, d_chains.p, d_chunks.p, n, e.max_n0, e.sys_gc_range, e.prof.reset_distance);
		}
		d_dirty.zero(e.stream);
		k_sys_check<<<(n + 255) / 256, 256, 0, e.stream>>>(d_chunks.p, n, d_dirty.p);
		e.launches += 2;
		RSQ_CUDA(cudaMemcpyAsync(&dirty, d_dirty.p, 4, cudaMemcpyDeviceToHost, e.stream));
		RSQ_CUDA(cudaStreamSynchronize(e.stream));
		RSQ_CUDA(cudaGetLastError());
		++passes;
		if(passes > n + 2){ throw std::runtime_error("systematic error chains did not converge"); }
	}
}

// FASTQ written by CreateSystematicErrorProfile: (id, dominant errors, compressed rates) per record
struct SysErrorRecord { std::string id, dom, rate; };
static std::vector<SysErrorRecord> read_sys_error_file(const std::string &path){
	TextInput in(path);
	if(!in.is_open()){ throw std::runtime_error("Could not open '" + path + "' for reading."); }
	std::istream &f = in.stream();
	std::vector<SysErrorRecord> recs;
	std::string id, seq, plus, qual;
	while(std::getline(f, id)){
		if(id.empty()){ continue; }
		if(id[0] != '@' || !std::getline(f, seq) || !std::getline(f, plus) || !std::getline(f, qual) || plus.empty() || plus[0] != '+'){
			throw std::runtime_error("Could not read systematic error profile '" + path + "': malformed fastq record");
		}
		recs.push_back({id.substr(1), seq, qual});
	}
	if(in.corrupt()){ throw std::runtime_error("Could not read systematic error profile '" + path + "': corrupt or truncated gzip stream"); }
	return recs;
}
// Simulator::ReadSystematicErrors (Simulator.h:326-335)
static void decode_sys_errors(const SysErrorRecord &r, uint8_t *out){
	for(size_t pos = 0; pos < r.dom.size(); ++pos){
		uint8_t rate = static_cast<uint8_t>(r.rate[pos] - 33);
		if(86 < rate){ rate += rate - 86; }
		out[2 * pos] = Genome::code(r.dom[pos]);
		out[2 * pos + 1] = rate;
	}
}

// Reference::SumBias for a list of (sequence, fragment length) chains on `st`: sums[k], max[k] per chain.  Chunked exact evaluation
// (ordered_sum.cuh); RSQ_BIAS_PATH=chain keeps the one-chain-per-CTA kernel (cross-check).
static void run_sum_bias(rsq_engine &e, const std::vector<BiasParamDev> &params, double *d_sums, double *d_max, cudaStream_t st){
	if(params.empty()){ return; }
	if(const char *env = getenv("RSQ_BIAS_PATH")){
		if(std::string(env) == "chain"){
			e.d_bias_params.upload(params, st);
			k_sum_bias<<<params.size(), 128, 0, st>>>(e.d_bias_params.p, params.size(), e.d_seq_off.p, e.d_seq_len.p, e.d_sur_start.p, e.d_sur_end.p, e.d_gc_prefix.p, e.d_gc_bias.p, d_sums, d_max);
			++e.launches;
			return;
		}
	}
	std::vector<BiasChain> chains(params.size());
	uint32_t K = 256;
	for(const auto &p : params){ if(e.genome.seqs[p.ref_id].size() >= (1u << 25)){ K = 1024; } }   // long chains: fewer chunks for the in-order pass
	if(const char *env = getenv("RSQ_BIAS_CHUNK")){ K = std::max(32, atoi(env) & ~31); }
	uint64_t chunk_total = 0, cta_total = 0;
	for(size_t k = 0; k < params.size(); ++k){
		BiasChain &c = chains[k];
		c.ref_id = params[k].ref_id; c.fragment_length = params[k].fragment_length; c.general = params[k].general;
		c.n_pos = static_cast<uint32_t>(e.genome.seqs[c.ref_id].size()) - c.fragment_length + 1u;
		c.n_chunks = (c.n_pos + K - 1) / K;
		c.chunk_first = static_cast<uint32_t>(chunk_total); c.cta_first = static_cast<uint32_t>(cta_total);
		chunk_total += c.n_chunks; cta_total += (c.n_chunks + kBiasChunkCta - 1) / kBiasChunkCta;
	}
	if(chunk_total >= (1ull << 32)){ throw std::runtime_error("bias normalisation: too many chunks"); }
	e.d_bias_chains.upload(chains, st);
	e.d_bias_plain.alloc(chunk_total); e.d_bias_start.alloc(chunk_total); e.d_bias_out.alloc(chunk_total); e.d_bias_cmax.alloc(chunk_total);
	e.d_bias_tie.alloc(chunk_total); e.d_bias_reran.alloc(chains.size());
	const uint32_t n = chains.size();
	k_bias_chunks<false><<<static_cast<unsigned>(cta_total), kBiasChunkCta, 0, st>>>(e.d_bias_chains.p, n, K, e.d_seq_off.p, e.d_sur_start.p, e.d_sur_end.p, e.d_gc_prefix.p, e.d_gc_bias.p,
	                                                                             nullptr, e.d_bias_plain.p, e.d_bias_cmax.p, nullptr);
	k_bias_scan<<<(n + 3) / 4, 128, 0, st>>>(e.d_bias_chains.p, n, e.d_bias_plain.p, e.d_bias_cmax.p, e.d_bias_start.p, d_max);
	k_bias_chunks<true><<<static_cast<unsigned>(cta_total), kBiasChunkCta, 0, st>>>(e.d_bias_chains.p, n, K, e.d_seq_off.p, e.d_sur_start.p, e.d_sur_end.p, e.d_gc_prefix.p, e.d_gc_bias.p,
	                                                                            e.d_bias_start.p, e.d_bias_out.p, nullptr, e.d_bias_tie.p);
	k_bias_resolve<<<(n + 3) / 4, 128, 0, st>>>(e.d_bias_chains.p, n, K, e.d_seq_off.p, e.d_sur_start.p, e.d_sur_end.p, e.d_gc_prefix.p, e.d_gc_bias.p,
	                                            e.d_bias_start.p, e.d_bias_out.p, e.d_bias_tie.p, d_sums, e.d_bias_reran.p);
	e.launches += 4;
}

static void prepare(rsq_engine &e, const Genome &ref_in, const rsq_sim_options &opt, rsq_sim_report *rep){
	cudaStream_t s = e.stream;
	const Profile &p = e.prof;
	SimCtx &c = e.ctx;
	EventTimer tm(s);
	e.prepared = false; e.downloaded = false; e.launches = 0; e.sur_windowed = false;
	e.d_error_flag.zero(s);

	// --- host: ReplaceN, ref-seq bias, pair counts (Simulator.cpp:2687-2743) ---
	stage_log("prepare: start");
	e.genome = ref_in;
	e.genome.replace_n(opt.seed);
	stage_log("prepare: genome copy + ReplaceN");
	Genome &g = e.genome;
	e.with_var = g.variants.loaded();
	e.num_alleles = e.with_var ? g.variants.num_alleles : 1;
	if(e.with_var){ g.variants.check_deferred(g.seqs); }   // REF columns that stood on an N: the reference opens the VCF after ReplaceN (Simulator.cpp:2690, 2750)
	if(e.with_var && g.methylation_loaded && g.methylation_alleles_max > 1 && g.methylation_alleles_max != e.num_alleles){
		throw std::runtime_error(std::to_string(g.methylation_alleles_max) + " alleles specified (must be either 1 or same as in variant file[" + std::to_string(e.num_alleles) + "]) in the methylation file");
	}
	// FragmentDistributionStats::UpdateRefSeqBias (FragmentDistributionStats.cpp:3352-3502)
	e.run_ref_seq_bias = p.ref_seq_bias;
	uint64_t master_draws_before = 0;     // master-stream outputs consumed on the host (kDraw only)
	switch(opt.ref_bias_model){
	case 0:   // kKeep, falling through to kNo when the biases do not match the reference
		if(e.run_ref_seq_bias.size() == g.seqs.size()){ break; }
		// fallthrough
	case 1:   // kNo
		e.run_ref_seq_bias.assign(g.seqs.size(), 1.0);
		break;
	case 2: { // kDraw: with replacement from the old biases, one draw per sequence (descending), from the MASTER stream
		struct CountingMt {
			std::mt19937_64 gen; uint64_t n = 0;
			typedef uint64_t result_type;
			static constexpr result_type min(){ return std::mt19937_64::min(); }
			static constexpr result_type max(){ return std::mt19937_64::max(); }
			result_type operator()(){ ++n; return gen(); }
		} cm;
		cm.gen.seed(opt.seed);
		const std::vector<double> old_bias = e.run_ref_seq_bias;
		e.run_ref_seq_bias.assign(g.seqs.size(), 0.0);
		std::uniform_int_distribution<uint32_t> rdist(0, old_bias.size());
		for(auto seq = e.run_ref_seq_bias.size(); seq--; ){
			const uint32_t pick = rdist(cm);
			if(pick >= old_bias.size()){ throw std::runtime_error("refBias draw: the reference draws index " + std::to_string(pick) + " of " + std::to_string(old_bias.size()) + " stored biases (std::out_of_range in UpdateRefSeqBias) for this seed"); }
			e.run_ref_seq_bias[seq] = old_bias[pick];
		}
		master_draws_before = cm.n;
		break;
	}
	case 3: { // kFile: "<identifier> <bias>" per line
		e.run_ref_seq_bias.assign(g.seqs.size(), 0.0);
		const std::string bias_file = opt.ref_bias_file ? opt.ref_bias_file : "";
		std::ifstream fbias(bias_file);
		if(!fbias.is_open()){ throw std::runtime_error("Unable to open reference bias file " + bias_file); }
		std::map<std::string, uint32_t> ids;
		for(size_t i = 0; i < g.seqs.size(); ++i){ ids.emplace(g.first_part(i), i); }
		std::vector<bool> found(g.seqs.size(), false);
		std::string line; uint32_t nline = 0; bool empty_line = false; uint32_t errors = 0;
		while(std::getline(fbias, line)){
			if(empty_line){ ++errors; continue; }
			++nline;
			if(line.empty()){ empty_line = true; continue; }
			const auto sep = line.find_last_of(" \t");
			if(sep == std::string::npos){ ++errors; continue; }
			double bias = 0.0;
			try{ bias = std::stod(line.substr(sep + 1)); }catch(...){ ++errors; bias = 0.0; }
			if(0.0 > bias){ ++errors; }
			auto id_len = line.find(' ');
			if(id_len == std::string::npos){ id_len = sep; }
			uint16_t id_start = 0;
			if('>' == line.at(0)){ ++id_start; --id_len; }
			auto it = ids.find(line.substr(id_start, id_len));
			if(it != ids.end()){ e.run_ref_seq_bias.at(it->second) = bias; found.at(it->second) = true; }
		}
		for(size_t i = 0; i < g.seqs.size(); ++i){ if(!found[i]){ ++errors; } }
		if(errors){ throw std::runtime_error("Error reading in reference sequence biases from " + bias_file + " (malformed line, negative bias or missing reference sequence)"); }
		break;
	}
	default:
		throw std::runtime_error("Unknown option chosen for reference sequence bias");
	}
	uint64_t reads = 0, sum_read_length = 0;
	for(int seg = 2; seg--; ){
		for(auto len = p.read_lengths[seg].from; len < p.read_lengths[seg].to(); ++len){ reads += p.read_lengths[seg][len]; sum_read_length += p.read_lengths[seg][len] * len; }
	}
	const double average_read_length = static_cast<double>(sum_read_length) / reads;
	e.total_size = g.total_size();
	const double adapter_part = coverage_prop_lost_from_adapters(p);
	double coverage = opt.coverage;
	if(opt.num_read_pairs){ e.total_pairs = opt.num_read_pairs; }
	else{
		if(0.0 == coverage){ coverage = p.corrected_coverage; }
		e.total_pairs = coverage_to_number_pairs(coverage, e.total_size, average_read_length, adapter_part);
	}
	e.adapter_only_pairs = std::round(static_cast<double>(e.total_pairs) * p.insert_lengths[0] / (p.total_number_reads / 2));
	if(const char *env = getenv("RSQ_FORCE_ADAPTER_ONLY")){ e.adapter_only_pairs = std::min<uint64_t>(e.total_pairs, atoll(env)); }   // test hook: the golden profiles have none
	e.total_pairs -= e.adapter_only_pairs;
	e.sys_gc_range = static_cast<uint32_t>((sum_read_length + reads / 2) / reads) / 2;

	// --- this engine's shard of the run (block range), and which sequences it has to hold ---
	// Shard boundaries: the even split of the simulated blocks, moved onto the first block of a sequence when one starts within 5 % of a shard's
	// size - a shard that only holds a sliver of a sequence would still need that sequence's whole systematic-error chains.
	const bool grouped = e.comm != nullptr;
	const uint32_t shard_count_pre = grouped ? static_cast<uint32_t>(e.group_world) : (opt.shard_count ? opt.shard_count : 1);
	const uint32_t shard_index_pre = grouped ? static_cast<uint32_t>(e.group_rank) : opt.shard_index;
	if(shard_index_pre >= shard_count_pre){ throw std::runtime_error("shard_index out of range"); }
	std::vector<uint64_t> seq_lengths(g.seqs.size());
	for(size_t i = 0; i < g.seqs.size(); ++i){ seq_lengths[i] = g.seqs[i].size(); }
	const ShardPlan plan = make_shard_plan(seq_lengths.data(), seq_lengths.size(), c.insert_to, shard_count_pre);   // shard_plan.hpp (also behind rsq_shard_plan)
	if(!plan.blocks_total){ throw std::runtime_error("All reference sequences are too short for simulating."); }
	e.plan = plan;
	e.shard_first = static_cast<uint32_t>(plan.first(shard_index_pre));
	e.shard_n = static_cast<uint32_t>(plan.count(shard_index_pre));
	// a sequence is needed by the shards that hold blocks of it; in a group its bias sums are computed by the first of them (its owner)
	auto shard_needs = [&](uint32_t k, size_t i) -> bool { return plan.needs(k, i); };
	std::vector<uint8_t> needed(g.seqs.size(), 1);
	std::vector<int32_t> owner(g.seqs.size(), 0);
	if(grouped){
		for(size_t i = 0; i < g.seqs.size(); ++i){
			needed[i] = shard_needs(shard_index_pre, i) ? 1 : 0;
			owner[i] = -1;
			for(uint32_t k = 0; k < shard_count_pre && owner[i] < 0; ++k){ if(shard_needs(k, i)){ owner[i] = static_cast<int32_t>(k); } }
			if(owner[i] < 0){ owner[i] = static_cast<int32_t>(i % shard_count_pre); }   // sequences nobody simulates (too short, look-ahead only): spread their sums
			if(owner[i] == static_cast<int32_t>(shard_index_pre)){ needed[i] = 1; }
		}
	}

	// --- upload reference: the sequences this engine needs, laid out behind each other ---
	stage_log("prepare: ref bias, counts");
	tm.start();
	std::vector<uint64_t> seq_off; std::vector<uint32_t> seq_len, name_off{0}; std::string names;
	uint64_t total = 0;
	for(size_t i = 0; i < g.seqs.size(); ++i){
		seq_off.push_back(total); seq_len.push_back(g.seqs[i].size()); total += needed[i] ? g.seqs[i].size() : 0;
		names += g.first_part(i); name_off.push_back(names.size());
	}
	e.h_seq_off = seq_off;
	e.h_ref_stage.ensure(total + 1);
	for(size_t i = 0; i < g.seqs.size(); ++i){ if(needed[i]){ std::memcpy(e.h_ref_stage.p + seq_off[i], g.seqs[i].data(), g.seqs[i].size()); } }
	stage_log("prepare: pinned staging of the reference");
	e.d_ref.alloc(total + 1);
	RSQ_CUDA(cudaMemcpyAsync(e.d_ref.p, e.h_ref_stage.p, total, cudaMemcpyHostToDevice, s));
	e.d_gc_prefix.alloc(total + g.seqs.size() + 1);
	{ uint32_t max_tiles = 1; for(const auto &q : g.seqs){ max_tiles = std::max<uint32_t>(max_tiles, (q.size() + kGcTile - 1) / kGcTile); } e.d_gc_tiles.alloc(max_tiles); }
	for(size_t i = 0; i < g.seqs.size(); ++i){
		const uint32_t L = needed[i] ? g.seqs[i].size() : 0;
		const uint32_t tiles = (L + kGcTile - 1) / kGcTile;
		uint32_t *gp = e.d_gc_prefix.p + seq_off[i] + i;
		if(!tiles){ RSQ_CUDA(cudaMemsetAsync(gp, 0, 4, s)); continue; }
		k_gc_tile_counts<<<tiles, 256, 0, s>>>(e.d_ref.p + seq_off[i], L, e.d_gc_tiles.p);
		k_gc_scan_tiles<<<1, 1024, 0, s>>>(e.d_gc_tiles.p, tiles);
		k_gc_prefix<<<tiles, 32, 0, s>>>(e.d_ref.p + seq_off[i], L, e.d_gc_tiles.p, gp);
		e.launches += 3;
	}
	e.d_seq_off.upload(seq_off, s); e.d_seq_len.upload(seq_len, s); e.d_name_off.upload(name_off, s);
	e.d_names.upload(names.data(), names.size() + 1, s);
	e.max_name_len = 0;
	for(size_t i = 0; i + 1 < name_off.size(); ++i){ e.max_name_len = std::max(e.max_name_len, name_off[i + 1] - name_off[i]); }
	e.d_ref_seq_bias.upload(e.run_ref_seq_bias, s);
	std::string base_id = (opt.record_base_identifier && opt.record_base_identifier[0]) ? opt.record_base_identifier : "ReseqRead";
	e.d_base_id.upload(base_id.data(), base_id.size() + 1, s);
	c.n_seqs = g.seqs.size(); c.seq_off = e.d_seq_off.p; c.seq_len = e.d_seq_len.p; c.ref = e.d_ref.p; c.gc_prefix = e.d_gc_prefix.p;
	c.name_blob = e.d_names.p; c.name_off = e.d_name_off.p; c.base_id = e.d_base_id.p; c.base_id_len = base_id.size();
	c.ref_seq_bias = e.d_ref_seq_bias.p;
	e.d_sur_start.alloc(total); e.d_sur_end.alloc(total);
	c.sur_start = e.d_sur_start.p; c.sur_end = e.d_sur_end.p;
	// --- variants: flat lists, first variant of every SimBlock, host half of SetSystematicErrorVariants* (variant_syserr.hpp) ---
	c.var = VarCtx{};
	c.binom_pow = nullptr;
	FlatVariants fv; std::vector<uint32_t> block_first, block_first_off;
	if(e.with_var){
		fv = g.variants.flatten();
		const VariantSysContext vctx = variant_sys_context(g.seqs, g.variants, fv);
		for(size_t i = 0; i < g.seqs.size(); ++i){
			block_first_off.push_back(block_first.size());
			const uint32_t nb = (g.seqs[i].size() + 999) / 1000;
			const auto &vars = g.variants.variants[i];
			uint32_t v = 0;
			for(uint32_t b = 0; b <= nb; ++b){
				while(v < vars.size() && vars[v].position < 1000ull * b){ ++v; }
				block_first.push_back(b == nb ? vars.size() : v);
			}
		}
		fv.position.push_back(0); fv.allele_lo.push_back(0); fv.allele_hi.push_back(0); fv.bases.push_back(0);   // never empty
		e.d_var_seq_first.upload(fv.seq_first, s); e.d_var_position.upload(fv.position, s); e.d_var_bases_off.upload(fv.bases_off, s); e.d_var_bases.upload(fv.bases, s);
		e.d_var_allele_lo.upload(fv.allele_lo, s); e.d_var_allele_hi.upload(fv.allele_hi, s);
		e.d_var_block_first.upload(block_first, s); e.d_var_block_first_off.upload(block_first_off, s);
		e.d_var_ctx_fwd.upload(vctx.fwd, s); e.d_var_ctx_rev.upload(vctx.rev, s);
		e.d_var_errs_fwd.alloc(2 * fv.bases.size() + 2); e.d_var_errs_fwd.zero(s); e.d_var_errs_rev.alloc(2 * fv.bases.size() + 2); e.d_var_errs_rev.zero(s);
		RSQ_CUDA(cudaStreamSynchronize(s));   // vctx is a local
		c.var.loaded = 1; c.var.num_alleles = e.num_alleles;
		if(const char *env = getenv("RSQ_VAR_PROBE")){ if(std::string(env) == "plain"){ c.var.loaded |= 2u; } }   // timing probe (wrong output): see eval_allele_hit
		c.var.seq_first = e.d_var_seq_first.p; c.var.position = e.d_var_position.p; c.var.bases_off = e.d_var_bases_off.p; c.var.bases = e.d_var_bases.p;
		c.var.allele_lo = e.d_var_allele_lo.p; c.var.allele_hi = e.d_var_allele_hi.p; c.var.errs_fwd = e.d_var_errs_fwd.p; c.var.errs_rev = e.d_var_errs_rev.p;
		c.var.block_first = e.d_var_block_first.p; c.var.block_first_off = e.d_var_block_first_off.p;
		for(int b = 0; b < 3; ++b){ c.var.sur_tab[b] = e.d_sur_tab[b].p; }
		stage_log("prepare: variants flattened and uploaded");
	}
	if(rep){ rep->ms_upload = tm.stop(); } else { tm.stop(); }

	// --- CalculateBiasNormalization: surroundings + per (ref, sampled length) sums on device ---
	tm.start();
	for(size_t i = 0; i < g.seqs.size(); ++i){
		const uint32_t L = needed[i] ? g.seqs[i].size() : 0;
		if(!L){ continue; }
		k_surroundings<<<(L + 255) / 256, 256, 0, s>>>(e.d_ref.p + seq_off[i], L, 0, L, e.d_sur_tab[0].p, e.d_sur_tab[1].p, e.d_sur_tab[2].p, e.d_sur_start.p + seq_off[i], e.d_sur_end.p + seq_off[i]);
		++e.launches;
	}
	Spline spline;
	if(!spline.get_sample_positions(p.insert_lengths)){ throw std::runtime_error("Sampling insert lengths did not find at least two usable lengths."); }
	std::vector<BiasParam> params; std::vector<BiasParamDev> dparams;
	const std::vector<double> &rsb = e.run_ref_seq_bias;
	for(uint32_t r = rsb.size(); r--; ){
		if(0.0 != rsb[r]){
			for(auto fl : spline.sample_positions){
				if(fl <= g.seqs[r].size()){ params.push_back({r, fl}); dparams.push_back({r, fl, rsb[r] * p.insert_lengths_bias[fl]}); }
			}
		}
	}
	// The ordered sums run on a second stream: they only feed the thresholds, which nothing needs before k_simulate,
	// so they overlap with the master stream and the systematic-error chains below.
	std::vector<double> sums(params.size(), 0.0), maxb(params.size(), 0.0);
	if(!params.empty() && !grouped){
		DevBuf<double> &d_sums = e.d_bias_sums, &d_max = e.d_bias_max; d_sums.alloc(params.size()); d_max.alloc(params.size());
		e.h_bias_results.ensure(2 * params.size() * sizeof(double));
		RSQ_CUDA(cudaEventRecord(e.ev_fork, s));
		RSQ_CUDA(cudaStreamWaitEvent(e.stream2, e.ev_fork, 0));
		run_sum_bias(e, dparams, d_sums.p, d_max.p, e.stream2);
		RSQ_CUDA(cudaMemcpyAsync(e.h_bias_results.p, d_sums.p, sums.size() * 8, cudaMemcpyDeviceToHost, e.stream2));
		RSQ_CUDA(cudaMemcpyAsync(e.h_bias_results.p + sums.size() * 8, d_max.p, maxb.size() * 8, cudaMemcpyDeviceToHost, e.stream2));
		RSQ_CUDA(cudaEventRecord(e.ev_join, e.stream2));
	}
	if(!params.empty() && grouped){
		// The chains of a sequence run on the engine that owns it (the first shard holding blocks of it); every engine contributes its sums and
		// maxima to one zero-initialised array and an all-reduce over NVLink hands everyone the whole set (x + 0.0 == x exactly, so the sum of the
		// contributions is the owner's value bit for bit).  The reference threads the same (sequence, length) list (FragmentDistributionStats.cpp:3527-3533).
		std::vector<BiasParamDev> mine; std::vector<uint32_t> mine_index;
		for(size_t k = 0; k < params.size(); ++k){ if(owner[params[k].ref_id] == static_cast<int32_t>(shard_index_pre)){ mine.push_back(dparams[k]); mine_index.push_back(k); } }
		DevBuf<double> &d_all = e.d_group_scratch; d_all.alloc(2 * params.size());
		e.h_bias_results.ensure(2 * params.size() * sizeof(double));
		RSQ_CUDA(cudaEventRecord(e.ev_fork, s));
		RSQ_CUDA(cudaStreamWaitEvent(e.stream2, e.ev_fork, 0));
		RSQ_CUDA(cudaMemsetAsync(d_all.p, 0, 2 * params.size() * 8, e.stream2));
		if(!mine.empty()){
			DevBuf<double> &d_sums = e.d_bias_sums, &d_max = e.d_bias_max; d_sums.alloc(mine.size()); d_max.alloc(mine.size());
			run_sum_bias(e, mine, d_sums.p, d_max.p, e.stream2);
			std::vector<double> hs(mine.size()), hm(mine.size());
			RSQ_CUDA(cudaMemcpyAsync(hs.data(), d_sums.p, hs.size() * 8, cudaMemcpyDeviceToHost, e.stream2));
			RSQ_CUDA(cudaMemcpyAsync(hm.data(), d_max.p, hm.size() * 8, cudaMemcpyDeviceToHost, e.stream2));
			RSQ_CUDA(cudaStreamSynchronize(e.stream2));
			std::vector<double> contrib(2 * params.size(), 0.0);
			for(size_t k = 0; k < mine.size(); ++k){ contrib[mine_index[k]] = hs[k]; contrib[params.size() + mine_index[k]] = hm[k]; }
			RSQ_CUDA(cudaMemcpyAsync(d_all.p, contrib.data(), contrib.size() * 8, cudaMemcpyHostToDevice, e.stream2));
			RSQ_CUDA(cudaStreamSynchronize(e.stream2));
		}
		RSQ_NCCL(nccl_api().AllReduce(d_all.p, d_all.p, 2 * params.size(), ncclDouble, ncclSum, e.comm, e.stream2));
		RSQ_CUDA(cudaMemcpyAsync(e.h_bias_results.p, d_all.p, 2 * params.size() * 8, cudaMemcpyDeviceToHost, e.stream2));
		RSQ_CUDA(cudaEventRecord(e.ev_join, e.stream2));
	}
	float ms_bias = tm.stop();

	// --- methylation regions (Reference::ReadMethylation) ---
	c.meth_loaded = g.methylation_loaded ? 1u : 0u;
	std::vector<int32_t> block_meth;
	if(g.methylation_loaded){
		std::vector<uint32_t> moff{0}, mstart, mend; std::vector<double> mrate;
		for(size_t i = 0; i < g.seqs.size(); ++i){
			for(size_t r = 0; r < g.unmethylated_regions[i].size(); ++r){
				mstart.push_back(g.unmethylated_regions[i][r].first); mend.push_back(g.unmethylated_regions[i][r].second); mrate.push_back(g.unmethylation[i].at(r));
			}
			moff.push_back(mstart.size());
		}
		mstart.push_back(0); mend.push_back(0); mrate.push_back(0.0);
		// Reference::Unmethylation(seq, allele): one column per allele when the file gives them (sequences with a single column repeat it)
		c.meth_alleles = g.methylation_alleles_max > 1 ? g.methylation_alleles_max : 1; c.meth_rate_stride = static_cast<uint32_t>(mrate.size());
		if(c.meth_alleles > 1){
			const size_t stride = mrate.size();
			mrate.resize(stride * c.meth_alleles, 0.0);
			for(uint32_t a = 1; a < c.meth_alleles; ++a){
				for(size_t i = 0; i < g.seqs.size(); ++i){
					for(size_t r = 0; r < g.unmethylated_regions[i].size(); ++r){
						const auto &cols = g.unmethylation_alleles[i];
						mrate[a * stride + moff[i] + r] = cols.size() > 1 ? cols.at(a).at(r) : g.unmethylation[i].at(r);
					}
				}
			}
		}
		e.d_meth_off.upload(moff, s); e.d_meth_start.upload(mstart, s); e.d_meth_end.upload(mend, s); e.d_meth_rate.upload(mrate, s);
		c.meth_off = e.d_meth_off.p; c.meth_start = e.d_meth_start.p; c.meth_end = e.d_meth_end.p; c.meth_rate = e.d_meth_rate.p;
		for(size_t i = 0; i < g.seqs.size(); ++i){
			const uint32_t L = g.seqs[i].size();
			if(L < c.insert_to){ continue; }
			for(uint32_t b = 0; b < (L + 999) / 1000; ++b){ block_meth.push_back(g.first_methylation_id(i, b * 1000)); }
		}
		e.d_block_meth.upload(block_meth, s);
	}

	// --- master stream: adapter systematic errors, then per unit reverse strand / seeds / forward strand ---
	tm.start();
	e.d_master_state.alloc(kMtN + 1);
	k_master_seed<<<1, 32, 0, s>>>(e.d_master_state.p, opt.seed); ++e.launches;
	if(master_draws_before){   // outputs the host consumed for --refBias draw
		e.d_master.alloc(master_draws_before);
		k_master_stream<<<1, kMasterThreads, 0, s>>>(e.d_master_state.p, e.d_master_state.p, e.d_master.p, master_draws_before); ++e.launches;
	}
	uint32_t carried = 0;
	uint32_t passes_total = 0;
	{
		std::vector<SysChain> chains; std::vector<std::pair<uint32_t, uint32_t>> lens;
		uint64_t n_draws = 0;
		struct Ad { int seg; size_t a; uint64_t raw_off; };
		std::vector<Ad> order;
		for(int seg = 2; seg--; ){
			for(size_t a = p.adapter_count_sum[seg].size(); a--; ){
				if(!p.adapter_count_sum[seg][a]){ continue; }
				const uint32_t len = e.h_adapter_off[seg][a + 1] - e.h_adapter_off[seg][a];
				order.push_back({seg, a, n_draws});
				n_draws += 2ull * len;
			}
		}
		if(n_draws){
			e.d_master.alloc(n_draws);
			k_master_stream<<<1, kMasterThreads, 0, s>>>(e.d_master_state.p, e.d_master_state.p, e.d_master.p, n_draws); ++e.launches;
			for(const auto &o : order){
				const uint32_t off = e.h_adapter_off[o.seg][o.a], len = e.h_adapter_off[o.seg][o.a + 1] - off;
				SysChain ch{}; ch.seq = e.d_adapter_seq.p + off; ch.L = len; ch.reverse = 0; ch.raw = e.d_master.p + o.raw_off; ch.seed_interleaved = 0;
				ch.out = e.d_adapter_sys.p + 2 * off; ch.carried_dom = carried;
				chains.push_back(ch); lens.push_back({len, 0});
				carried = dominant_before(e.h_adapter_seq.data() + off, len, false, len, carried);
			}
			uint32_t passes = 0;
			run_sys_chains(e, chains, lens, 1u << 30, 0, passes);
		}
	}
	e.d_sys_fwd.alloc(2 * total + 2); e.d_sys_rev.alloc(2 * total + 2);
	c.sys_fwd = e.d_sys_fwd.p; c.sys_rev = e.d_sys_rev.p;
	const bool from_file = opt.sys_error_file && opt.sys_error_file[0];
	std::vector<SysErrorRecord> sys_records;
	if(from_file){ sys_records = read_sys_error_file(opt.sys_error_file); }
	size_t next_record = 0;
	uint64_t max_unit_draws = 0; uint32_t nb_total = 0, nb_max = 0;
	if(from_file && e.with_var){
		throw std::runtime_error("--readSysError together with a variant file is not supported by this engine revision (the distance state in front of every SimBlock comes from the systematic-error chains)");
	}
	// master-stream draws of a unit: one seed per reverse block, 2 per position and strand, one seed per forward block - and with variants 2 per
	// replacement base and strand (SetSystematicErrorVariantsReverse / Forward, Simulator.cpp:836-844, 1075-1083)
	auto variant_bases = [&](size_t i) -> uint64_t { return e.with_var ? fv.bases_off[fv.seq_first[i + 1]] - fv.bases_off[fv.seq_first[i]] : 0ull; };
	for(size_t i = 0; i < g.seqs.size(); ++i){
		const uint32_t L = g.seqs[i].size();
		if(L < c.insert_to){ continue; }
		const uint32_t nb = (L + 999) / 1000;
		nb_total += nb; nb_max = std::max(nb_max, nb);
		max_unit_draws = std::max<uint64_t>(max_unit_draws, 2ull * nb + (from_file ? 0ull : 4ull * L + 4ull * variant_bases(i)));
	}
	if(!nb_total){ throw std::runtime_error("All reference sequences are too short for simulating."); }
	e.d_master.alloc(max_unit_draws + 1);
	e.d_blocks.alloc(nb_total);
	if(e.with_var){ e.d_var_blk_off.alloc(2ull * nb_max); e.d_var_bstate.alloc(2ull * nb_max); }
	uint32_t next_block_id = 1, first = 0;
	std::vector<uint8_t> decoded;
	// this engine's shard of the run (computed above): systematic errors and block seeds are only needed for the sequences it has blocks in
	// (fragments never span sequences); the master-stream draws of the others are skipped by jump-ahead
	const uint64_t shard_lo = e.shard_first, shard_hi = static_cast<uint64_t>(e.shard_first) + e.shard_n;
	for(size_t i = 0; i < g.seqs.size(); ++i){
		const uint32_t L = g.seqs[i].size();
		if(L < c.insert_to){ continue; }
		const uint32_t nb = (L + 999) / 1000;
		const uint64_t n_draws = 2ull * nb + (from_file ? 0ull : 4ull * L + 4ull * variant_bases(i));
		const bool needed_here = shard_count_pre == 1 || (e.shard_n && first < shard_hi && shard_lo < static_cast<uint64_t>(first) + nb);
		if(!needed_here){
			if(from_file){ next_record += 2; }   // the two records of this sequence stay unread
			master_skip(e, n_draws);
			const uint8_t *hseq = g.seqs[i].data();
			carried = dominant_before(hseq, L, true, L, carried);
			carried = dominant_before(hseq, L, false, L, carried);
			next_block_id += nb; first += nb;
			continue;
		}
		master_generate(e, e.d_master.p, n_draws);
		if(from_file){
			// CreateUnit: LoadSysErrorRecord (reverse strand) ... LoadSysErrorRecord (forward strand), strictly in file order
			for(int strand = 0; strand < 2; ++strand){
				if(next_record >= sys_records.size()){ throw std::runtime_error("Could not read systematic error profile for reference sequence '" + g.ids[i] + "': end of file"); }
				const SysErrorRecord &r = sys_records[next_record++];
				if(r.dom.size() != L || r.rate.size() != L){
					throw std::runtime_error("Systematic error profile '" + r.id + "' (length " + std::to_string(r.dom.size()) + ") does not match reference sequence '" + g.ids[i] +
					                         "' (length " + std::to_string(L) + "). Wrong file or order incorrect?");
				}
				decoded.resize(2ull * L);
				decode_sys_errors(r, decoded.data());
				RSQ_CUDA(cudaMemcpyAsync((strand ? e.d_sys_fwd.p : e.d_sys_rev.p) + 2 * seq_off[i], decoded.data(), 2ull * L, cudaMemcpyHostToDevice, s));
				RSQ_CUDA(cudaStreamSynchronize(s));
			}
			k_build_blocks<<<(nb + 127) / 128, 128, 0, s>>>(e.d_blocks.p, first, nb, i, next_block_id, e.d_master.p + nb, g.methylation_loaded ? e.d_block_meth.p : nullptr, 1u); ++e.launches;
		}
		else if(e.with_var){
			// every SimBlock's draws are followed by those of its variants: the blocks' places in the unit's stream come from tables
			const uint8_t *hseq = g.seqs[i].data();
			const uint32_t *bf = block_first.data() + block_first_off[i];
			const uint32_t vf = fv.seq_first[i];
			std::vector<uint64_t> blk_off(2ull * nb);   // [0, nb): reverse blocks (first position draw), [nb, 2 nb): forward blocks (seed)
			uint64_t at = nb;
			auto vb = [&](uint32_t b) -> uint64_t { return fv.bases_off[vf + bf[b + 1]] - fv.bases_off[vf + bf[b]]; };
			auto size_of = [&](uint32_t b) -> uint64_t { return std::min<uint64_t>(1000ull * (b + 1), L) - 1000ull * b; };
			for(uint32_t b = nb; b--; ){ blk_off[b] = at; at += 2 * size_of(b) + 2 * vb(b); }
			for(uint32_t b = 0; b < nb; ++b){ blk_off[nb + b] = at; at += 1 + 2 * size_of(b) + 2 * vb(b); }
			if(at != n_draws){ throw std::runtime_error("internal error: master stream layout of a unit with variants"); }
			RSQ_CUDA(cudaMemcpyAsync(e.d_var_blk_off.p, blk_off.data(), blk_off.size() * 8, cudaMemcpyHostToDevice, s));
			RSQ_CUDA(cudaMemsetAsync(e.d_var_bstate.p, 0, 2ull * nb * 4, s));
			std::vector<SysChain> chains(2); std::vector<std::pair<uint32_t, uint32_t>> lens{{L, 0}, {L, 0}};
			chains[0].seq = e.d_ref.p + seq_off[i]; chains[0].L = L; chains[0].reverse = 1; chains[0].raw = e.d_master.p; chains[0].seed_interleaved = 0;
			chains[0].out = e.d_sys_rev.p + 2 * seq_off[i]; chains[0].carried_dom = carried; chains[0].blk_off = e.d_var_blk_off.p; chains[0].bstate = e.d_var_bstate.p;
			carried = dominant_before(hseq, L, true, L, carried);
			chains[1].seq = e.d_ref.p + seq_off[i]; chains[1].L = L; chains[1].reverse = 0; chains[1].raw = e.d_master.p; chains[1].seed_interleaved = 1;
			chains[1].out = e.d_sys_fwd.p + 2 * seq_off[i]; chains[1].carried_dom = carried; chains[1].blk_off = e.d_var_blk_off.p + nb; chains[1].bstate = e.d_var_bstate.p + nb;
			carried = dominant_before(hseq, L, false, L, carried);
			uint32_t passes = 0;
			run_sys_chains(e, chains, lens, 8192, 1024, passes);
			passes_total = std::max(passes_total, passes);
			if(fv.seq_first[i + 1] > vf){
				VarDrawCtx dfw{}, drv{};
				dfw.ctx = e.d_var_ctx_fwd.p; dfw.errs = e.d_var_errs_fwd.p; dfw.sys = e.d_sys_fwd.p + 2 * seq_off[i]; dfw.gcp = e.d_gc_prefix.p + seq_off[i] + i;
				dfw.block_first = e.d_var_block_first.p + block_first_off[i];
				dfw.v.position = e.d_var_position.p + vf; dfw.v.bases_off = e.d_var_bases_off.p + vf; dfw.v.bases = e.d_var_bases.p; dfw.v.allele_lo = e.d_var_allele_lo.p + vf; dfw.v.allele_hi = e.d_var_allele_hi.p + vf;
				dfw.v.n = fv.seq_first[i + 1] - vf; dfw.L = L; dfw.reverse = 0; dfw.sys_gc_range = e.sys_gc_range; dfw.reset_distance = p.reset_distance;
				drv = dfw; drv.ctx = e.d_var_ctx_rev.p; drv.errs = e.d_var_errs_rev.p; drv.sys = e.d_sys_rev.p + 2 * seq_off[i]; drv.reverse = 1;
				k_var_sys_errors<<<(2 * nb + 127) / 128, 128, 0, s>>>(c.tab, dfw, drv, nb, e.d_var_bstate.p + nb, e.d_var_bstate.p, e.d_master.p, e.d_var_blk_off.p + nb, e.d_var_blk_off.p); ++e.launches;
			}
			k_build_blocks<<<(nb + 127) / 128, 128, 0, s>>>(e.d_blocks.p, first, nb, i, next_block_id, e.d_master.p, g.methylation_loaded ? e.d_block_meth.p : nullptr, 0u,
			                                              e.d_var_blk_off.p + nb, e.d_var_block_first.p + block_first_off[i]); ++e.launches;
		}
		else{
			const uint8_t *hseq = g.seqs[i].data();
			std::vector<SysChain> chains(2); std::vector<std::pair<uint32_t, uint32_t>> lens{{L, 0}, {L, 0}};
			chains[0].seq = e.d_ref.p + seq_off[i]; chains[0].L = L; chains[0].reverse = 1; chains[0].raw = e.d_master.p + nb; chains[0].seed_interleaved = 0;
			chains[0].out = e.d_sys_rev.p + 2 * seq_off[i]; chains[0].carried_dom = carried;
			carried = dominant_before(hseq, L, true, L, carried);
			chains[1].seq = e.d_ref.p + seq_off[i]; chains[1].L = L; chains[1].reverse = 0; chains[1].raw = e.d_master.p + nb + 2ull * L; chains[1].seed_interleaved = 1;
			chains[1].out = e.d_sys_fwd.p + 2 * seq_off[i]; chains[1].carried_dom = carried;
			carried = dominant_before(hseq, L, false, L, carried);
			uint32_t passes = 0;
			run_sys_chains(e, chains, lens, 8192, 1024, passes);
			passes_total = std::max(passes_total, passes);
			k_build_blocks<<<(nb + 127) / 128, 128, 0, s>>>(e.d_blocks.p, first, nb, i, next_block_id, e.d_master.p + nb + 2ull * L, g.methylation_loaded ? e.d_block_meth.p : nullptr, 2001u); ++e.launches;
		}
		RSQ_CUDA(cudaStreamSynchronize(s));   // d_master is reused by the next unit
		next_block_id += nb; first += nb;
	}
	e.n_blocks_total = nb_total;
	const uint32_t lookahead = 1 + c.insert_to / 1000;    // blocks the reference creates but never hands to a worker (Simulator.cpp:1298-1303)
	e.n_blocks_sim = nb_total > lookahead ? nb_total - lookahead : 0;
	e.adapter_only_seed = 0;
	if(e.adapter_only_pairs){
		k_master_stream<<<1, kMasterThreads, 0, s>>>(e.d_master_state.p, e.d_master_state.p, e.d_master.p, 1); ++e.launches;
		RSQ_CUDA(cudaMemcpyAsync(&e.adapter_only_seed, e.d_master.p, 8, cudaMemcpyDeviceToHost, s));
		RSQ_CUDA(cudaStreamSynchronize(s));
	}
	e.syserr_passes = passes_total;
	const float ms_syserr = tm.stop();
	// --- join the bias sums, finish CalculateBiasNormalization on the host (spline, thresholds: same libm as the reference) ---
	tm.start();
	if(!params.empty()){
		RSQ_CUDA(cudaEventSynchronize(e.ev_join));
		std::memcpy(sums.data(), e.h_bias_results.p, sums.size() * 8);
		std::memcpy(maxb.data(), e.h_bias_results.p + sums.size() * 8, maxb.size() * 8);
	}
	if(!finish_normalization(e.norm, p, rsb, spline, params, sums, maxb, e.total_pairs, e.num_alleles, e.with_var)){ throw std::runtime_error("bias normalisation is zero"); }
	e.d_thr.upload(e.norm.thresholds, s); e.d_thr_int.upload(e.norm.thr_int, s); e.d_thr_hi.upload(e.norm.thr_hi, s); e.d_binom_p0.upload(e.norm.binom_p0, s); e.d_cov_group.upload(e.norm.coverage_groups, s);
	if(e.with_var){ e.d_binom_pow.upload(e.norm.binom_pow, s); c.binom_pow = e.d_binom_pow.p; }
	c.thr = e.d_thr.p; c.thr_int = e.d_thr_int.p; c.thr_hi = e.d_thr_hi.p; c.thr_hi_stride = e.norm.thr_hi_stride; c.binom_p0 = e.d_binom_p0.p; c.coverage_group = e.d_cov_group.p; c.bias_normalization = e.norm.bias_normalization;
	RSQ_CUDA(cudaStreamSynchronize(s));
	ms_bias += tm.stop();
	if(rep){ rep->ms_bias = ms_bias; rep->bias_normalization = e.norm.bias_normalization; rep->ms_syserr = ms_syserr; }
	stage_log("prepare: device stages (bias, master stream, systematic errors)");
	const uint32_t sc = shard_count_pre, si = shard_index_pre;
	if(static_cast<uint64_t>(e.shard_first) + e.shard_n > e.n_blocks_sim){ throw std::runtime_error("internal error: shard range beyond the simulated blocks"); }
	e.shard_has_adapter_only = (si + 1 == sc) && e.adapter_only_pairs;
	RSQ_CUDA(cudaGetLastError());
	const uint32_t flag = read_error_flag(e);
	if(flag){ throw std::runtime_error("device reported: " + describe_flag(flag)); }
	if(rep){
		rep->total_pairs_aim = e.total_pairs; rep->adapter_only_pairs = e.adapter_only_pairs; rep->blocks_total = e.n_blocks_total;
		rep->syserr_passes = e.syserr_passes; rep->kernel_launches = e.launches;
	}
	e.prepared = true;
}

// Simulator::CreateSystematicErrorProfile: every sequence, whole reverse strand then whole forward strand, each chain from
// reset counters with a contiguous run of the master stream (no block seeds, no adapters).
static void create_sys_profile(rsq_engine &e, const Genome &g, uint64_t seed, const char *out_path){
	cudaStream_t s = e.stream;
	e.launches = 0; e.d_error_flag.zero(s);
	for(const auto &q : g.seqs){ for(uint8_t b : q){ if(b > 3){ throw std::runtime_error("Reference contains ambiguous bases(e.g. N). Please replace them, remove them from the scaffolds or split scaffolds into contigs."); } } }
	// The reference never initialises sys_gc_range_ on this path (it is only set inside Simulate / SimulateErrorModelOnly,
	// Simulator.cpp:2782, 2962; CreateSystematicErrorProfile runs on a fresh Simulator object): the GC window length is
	// whatever the stack held.  In the reference build of this image every value >= ~2*10^4 reproduces its output; we use
	// the largest uintReadLen, i.e. "all bases seen so far, at most 65535".
	e.sys_gc_range = 65535;
	TextSink o;
	if(!o.open(out_path)){ throw std::runtime_error(std::string("Could not open '") + out_path + "' for writing."); }
	{
		e.d_master_state.alloc(kMtN + 1);
		k_master_seed<<<1, 32, 0, s>>>(e.d_master_state.p, seed); ++e.launches;
		uint32_t carried = 0;
		std::vector<uint8_t> host; std::string text;
		for(size_t i = 0; i < g.seqs.size(); ++i){
			const uint32_t L = g.seqs[i].size();
			if(!L){ const std::string t = "@" + g.ids[i] + " reverse\n\n+\n\n@" + g.ids[i] + " forward\n\n+\n\n"; if(!o.write(t.data(), t.size())){ throw std::runtime_error("Could not write systematic error profile"); } continue; }
			e.d_ref.alloc(L + 1);
			RSQ_CUDA(cudaMemcpyAsync(e.d_ref.p, g.seqs[i].data(), L, cudaMemcpyHostToDevice, s));
			e.d_master.alloc(4ull * L);
			master_generate(e, e.d_master.p, 4ull * L);
			e.d_sys_rev.alloc(2ull * L + 2); e.d_sys_fwd.alloc(2ull * L + 2);
			std::vector<SysChain> chains(2); std::vector<std::pair<uint32_t, uint32_t>> lens{{L, 0}, {L, 0}};
			chains[0].seq = e.d_ref.p; chains[0].L = L; chains[0].reverse = 1; chains[0].raw = e.d_master.p; chains[0].out = e.d_sys_rev.p; chains[0].carried_dom = carried;
			carried = dominant_before(g.seqs[i].data(), L, true, L, carried);
			chains[1].seq = e.d_ref.p; chains[1].L = L; chains[1].reverse = 0; chains[1].raw = e.d_master.p + 2ull * L; chains[1].out = e.d_sys_fwd.p; chains[1].carried_dom = carried;
			carried = dominant_before(g.seqs[i].data(), L, false, L, carried);
			uint32_t passes = 0;
			run_sys_chains(e, chains, lens, 8192, 1024, passes);
			host.resize(2ull * L);
			for(int strand = 0; strand < 2; ++strand){
				RSQ_CUDA(cudaMemcpyAsync(host.data(), strand ? e.d_sys_fwd.p : e.d_sys_rev.p, 2ull * L, cudaMemcpyDeviceToHost, s));
				RSQ_CUDA(cudaStreamSynchronize(s));
				text.assign(1, '@'); text += g.ids[i]; text += strand ? " forward\n" : " reverse\n";
				const size_t seq_at = text.size();
				text.resize(seq_at + 2ull * L + 4);
				char *dom = &text[seq_at], *qual = dom + L + 3;
				for(uint32_t pos = 0; pos < L; ++pos){
					dom[pos] = "ACGTN"[host[2 * pos] > 4 ? 4 : host[2 * pos]];
					uint8_t q = host[2 * pos + 1];
					if(86 < q){ q -= (q - 85) / 2; }      // WriteOutSystematicErrorProfile (Simulator.cpp:2569-2575)
					qual[pos] = static_cast<char>(q + 33);
				}
				dom[L] = '\n'; dom[L + 1] = '+'; dom[L + 2] = '\n'; qual[L] = '\n';
				if(!o.write(text.data(), text.size())){ throw std::runtime_error("Could not write systematic error profile"); }
			}
		}
		const uint32_t flag = read_error_flag(e);
		if(flag){ throw std::runtime_error("device reported: " + describe_flag(flag)); }
	}
	if(!o.close()){ throw std::runtime_error("Could not write systematic error profile"); }
	e.prepared = false;
}

static void setup_arena(rsq_engine &e, Arena &a, uint64_t expected_bytes, uint32_t slots){
	uint64_t n_chunks = expected_bytes / kChunkBytes + 2ull * slots + 64;
	e.d_arena.alloc(n_chunks * kChunkBytes);
	e.d_chunk_next.alloc(n_chunks); e.d_chunk_used.alloc(n_chunks);
	e.d_next_free.alloc(1); e.d_next_free.zero(e.stream);
	a.data = e.d_arena.p; a.chunk_bytes = kChunkBytes; a.n_chunks = n_chunks; a.next_free = e.d_next_free.p;
	a.chunk_next = e.d_chunk_next.p; a.chunk_used = e.d_chunk_used.p; a.error_flag = e.d_error_flag.p;
}

static void gather(rsq_engine &e, const Arena &a, uint32_t slots, rsq_sim_report *rep){
	cudaStream_t s = e.stream;
	EventTimer tm(s);
	tm.start();
	e.d_offsets.alloc(2ull * (slots + 1)); e.d_totals.alloc(4);
	k_block_offsets<<<1, 1024, 0, s>>>(e.d_block_out.p, slots, e.d_offsets.p, e.d_totals.p); ++e.launches;
	unsigned long long totals[4];
	RSQ_CUDA(cudaMemcpyAsync(totals, e.d_totals.p, sizeof totals, cudaMemcpyDeviceToHost, s));
	RSQ_CUDA(cudaStreamSynchronize(s));
	e.out_bytes[0] = totals[0]; e.out_bytes[1] = totals[1]; e.out_pairs = totals[2]; e.out_draws = totals[3];
	e.d_out_batch[0][0].alloc(totals[0] + 16); e.d_out_batch[0][1].alloc(totals[1] + 16);
	e.last_par = 0; e.streamed_to_host = false;
	if(slots){ k_gather<<<2 * slots, 128, 0, s>>>(e.d_block_out.p, slots, a, e.d_offsets.p, e.d_out_batch[0][0].p, e.d_out_batch[0][1].p); ++e.launches; }
	const float ms = tm.stop();
	if(rep){ rep->ms_gather = ms; }
	RSQ_CUDA(cudaGetLastError());
}

static void fill_simulate_report(rsq_engine &e, rsq_sim_report *rep, float ms_sim){
	if(!rep){ return; }
	rep->ms_simulate = ms_sim; rep->pairs = e.out_pairs; rep->bytes[0] = e.out_bytes[0]; rep->bytes[1] = e.out_bytes[1];
	rep->blocks = e.shard_n; rep->scan_draws = e.out_draws; rep->kernel_launches = e.launches;
	rep->spec_rounds = e.spec_rounds; rep->spec_depth = e.spec_depth;
	uint64_t positions = 0;
	std::vector<BlockDesc> hb(e.shard_n);
	if(e.shard_n){ RSQ_CUDA(cudaMemcpy(hb.data(), e.d_blocks.p + e.shard_first, e.shard_n * sizeof(BlockDesc), cudaMemcpyDeviceToHost)); }
	for(const auto &b : hb){ positions += std::min<uint32_t>(1000, e.genome.seqs[b.ref_id].size() - b.start_pos); }
	rep->positions = positions;
}

// Speculative two-phase path: rounds of k_spec_scan + k_spec_reads until every unit is done (see spec_core.cuh).
// One batch of this shard's blocks [u_begin, u_begin + u_count) (+ the adapter-only pairs behind the last one): on return the
// FASTQ text of the batch is in e.d_out_batch[par][0/1], its sizes in `res`.  false: a read outran the speculation margin.
struct BatchResult { uint64_t bytes[2] = {0, 0}; uint64_t pairs = 0, draws = 0; float ms_sim = 0, ms_gather = 0; uint32_t rounds = 0, depth = 0; };
static bool simulate_spec_batch(rsq_engine &e, uint32_t u_begin, uint32_t u_count, bool with_adapter_only, uint32_t depth_cap, int par, BatchResult &res,
                                const SpecCtx *records = nullptr){
	cudaStream_t s = e.stream;
	SimCtx &c = e.ctx;
	int dev_sms = 0; RSQ_CUDA(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, e.device));
	const uint32_t first_desc = e.shard_first + u_begin;
	SpecCtx sp{};
	sp.n_blocks = u_count;
	sp.n_units = u_count + (with_adapter_only ? 1 : 0);
	sp.adapter_only_pairs = with_adapter_only ? e.adapter_only_pairs : 0;
	sp.adapter_only_seed = e.adapter_only_seed;
	if(records){   // seqToIllumina: the units are batches of input records
		sp.em_recs = records->em_recs; sp.em_n = records->em_n; sp.em_batch = records->em_batch; sp.em_seeds = records->em_seeds;
		sp.em_seq = records->em_seq; sp.em_sys = records->em_sys; sp.em_ids = records->em_ids;
	}
	const uint32_t max_rl = std::max(c.read_len_to[0], c.read_len_to[1]);   // ReadLengths().to(): one past the longest read
	sp.words_per_job = (3 * max_rl + 8 + kSpecMargin + 7u) & ~7u;   // whole 64-byte pieces (ReadMachineT::win_issue)
	sp.margin = getenv("RSQ_SPEC_MARGIN") ? std::min<uint32_t>(kSpecMargin, atoi(getenv("RSQ_SPEC_MARGIN"))) : kSpecMargin;   // tests: 0 forces the serial fallback
	// Depth (reads speculated per unit and round) and reads per warp are chosen per batch of rounds from a cost model:
	//   a round costs max(latency, throughput) with latency ~ 1.0 ms + 0.12 ms per read of depth (one lock-step pass over a read
	//   + the scan in front of `depth` reads) and throughput ~ 3.2e-5 ms per read in flight plus 3 reads' worth of fixed work per
	//   unit (verification, snapshot restore/save) on a B200 (measured at 4.6 and 100 Mbp: 5.3 ms per round of 4641 x 32 reads);
	//   it verifies (1 - p^d) / (1 - p) reads per unit, p = measured share of reads whose assumption held.
	// Small runs are latency bound (few reads per warp), large ones throughput bound; how deep to speculate mostly depends on p.
	const uint64_t warps_cap = static_cast<uint64_t>(dev_sms) * 16;
	const int fixed_depth = getenv("RSQ_SPEC_DEPTH") ? std::min<int>(depth_cap, std::max(1, atoi(getenv("RSQ_SPEC_DEPTH")))) : 0;
	const int fixed_lanes = getenv("RSQ_SPEC_LANES") ? atoi(getenv("RSQ_SPEC_LANES")) : 0;
	uint32_t depth = depth_cap;   // capacity per unit (array stride)
	if(fixed_depth){ depth = fixed_depth; }
	if(const char *env = getenv("RSQ_SPEC_MAX_DEPTH")){ depth = std::max<uint32_t>(fixed_depth, std::min<int>(depth_cap, std::max(1, atoi(env)))); }
	sp.depth = depth;
	const double draws_per_read = records ? 0.0 : static_cast<double>(e.total_size) * (c.insert_to - c.insert_from) / (2.0 * std::max<uint64_t>(1, e.total_pairs));
	const double budget_factor = getenv("RSQ_SPEC_BUDGET") ? atof(getenv("RSQ_SPEC_BUDGET")) : 1.5;
	double lat_fixed = 1.0, lat_per_read = 0.12;
	if(const char *env = getenv("RSQ_SPEC_LAT")){ sscanf(env, "%lf,%lf", &lat_fixed, &lat_per_read); }   // tuning runs
	const bool trace = getenv("RSQ_SPEC_TRACE") != nullptr;
	// Option (RSQ_SPEC_CAP=64 RSQ_SPEC_HOT=1): a unit whose reads come denser than average speculates proportionally deeper (capacity 64 instead of 32),
	// so that the hot SimBlocks (up to 1.6 x the reads of an average one on E. coli) do not need as many more rounds.  Measured on E. coli: 62.0 ms
	// against 52.2 ms without - the wider slot range per unit costs more than the few units that finish a round earlier gain.  Off.
	const bool hot_scaling = depth_cap > 32 && !records && getenv("RSQ_SPEC_HOT") && atoi(getenv("RSQ_SPEC_HOT")) == 1;
	sp.mean_reads = hot_scaling && e.n_blocks_sim ? static_cast<float>(2.0 * e.total_pairs / e.n_blocks_sim) : 0.0f;
	uint32_t lanes = 32;
	auto choose = [&](uint64_t active, double p_hold){
		uint32_t best_d = 2; double best = -1.0;
		static const uint32_t cand[] = {2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64};
		for(uint32_t d : cand){
			if(d > depth || (hot_scaling && d > 32u)){ break; }
			const double prog = p_hold >= 0.9999 ? d : (1.0 - std::pow(p_hold, static_cast<double>(d))) / (1.0 - p_hold);
			const double latency = records ? 0.6 + 0.01 * d : lat_fixed + lat_per_read * d;   // seqToIllumina units have no scan in front of their reads
			const double cost = std::max(latency, 3.2e-5 * static_cast<double>(active) * (3.0 + d));
			if(prog / cost > best){ best = prog / cost; best_d = d; }
		}
		if(fixed_depth){ best_d = fixed_depth; }
		sp.run_depth = best_d;
		sp.map_depth = hot_scaling ? sp.depth : best_d;
		const double budget = budget_factor * best_d * draws_per_read + 4096.0;
		sp.scan_budget = budget > 4.0e9 ? 4000000000u : static_cast<uint32_t>(budget);
		// reads per warp: as few as keep all reads of the round resident at once
		const uint64_t reads = active * best_d;
		lanes = reads <= warps_cap * 8 ? 8 : (reads <= warps_cap * 16 ? 16 : 32);   // 32 once the reads of a round no longer fit one wave anyway
		if(fixed_lanes){ lanes = fixed_lanes >= 32 ? 32 : (fixed_lanes >= 16 ? 16 : 8); }
	};
	choose(sp.n_units, 0.99);   // before the first poll has measured it: InDels are rare in every profile seen so far (E. coli: 58 ms starting deep, 60 ms starting at depth 8)
	res.depth = sp.run_depth;
	const size_t n_jobs = static_cast<size_t>(sp.n_units) * depth;
	const size_t n_tiles = (n_jobs + 31) / 32;
	e.d_spec_blocks.alloc(sp.n_units); e.d_spec_snaps.alloc(2 * static_cast<size_t>(sp.n_units) * (depth + 1)); e.d_spec_jobs.alloc(n_tiles * 32);
	e.d_spec_words.alloc(n_tiles * 32 * sp.words_per_job);
	sp.blocks = e.d_spec_blocks.p; sp.snaps = e.d_spec_snaps.p; sp.jobs = e.d_spec_jobs.p; sp.words = e.d_spec_words.p;
	const bool with_var = c.var.loaded != 0 && !records;
	if(c.meth_loaded || with_var){ e.d_spec_conv.alloc(static_cast<size_t>(sp.n_units) * kConvSlots * 2 * kMaxOrgLen); sp.conv = e.d_spec_conv.p; }
	sp.chosen_stride = with_var ? 2 * c.var.num_alleles : 0;
	if(with_var){ e.d_spec_chosen.alloc(2 * static_cast<size_t>(sp.n_units) * (depth + 1) * sp.chosen_stride); sp.snap_chosen = e.d_spec_chosen.p; }
	const size_t scan_shmem = static_cast<size_t>(kWarpsPerCta) * c.thr_hi_stride * sizeof(uint32_t) + (with_var ? static_cast<size_t>(kWarpsPerCta) * sp.chosen_stride * sizeof(uint16_t) : 0);
	const uint32_t id_prefix = records ? e.em_max_id_len + 1 : c.base_id_len + 10 + 1 + 20 + (with_var ? 10 : 0) + 1 + 10 + 1 + std::max<uint32_t>(e.max_name_len, 7) + 1 + 10 + 1 + 5 + 11;
	sp.id_cap = std::min<uint32_t>(kIdCap, (id_prefix + kCigarCap + 2 + 10 + 15) & ~15u);
	sp.seq_off = 16 + sp.id_cap; sp.qual_off = sp.seq_off + ((max_rl + 3) & ~3u); sp.slot_stride = (sp.qual_off + max_rl + 15) & ~15u;
	e.d_spec_counters.alloc(8); e.h_spec_counters.ensure(8 * sizeof(uint32_t));
	sp.next_slab = e.d_spec_counters.p; sp.n_done = e.d_spec_counters.p + 1; sp.stat = reinterpret_cast<unsigned long long *>(e.d_spec_counters.p + 2);
	// units are split into independent groups on their own streams: the scan of one group overlaps the reads of another
	uint32_t n_groups = sp.n_units <= 20000u ? 4 : 2;   // small runs: shorter kernels, more of them in flight (E. coli: 62.7 ms with four groups, 64.1 with two)
	if(const char *env = getenv("RSQ_SPEC_GROUPS")){ n_groups = std::min(4, std::max(1, atoi(env))); }
	n_groups = std::max<uint32_t>(1, std::min<uint32_t>(n_groups, sp.n_units));
	while(e.spec_streams.size() < n_groups){
		cudaStream_t st; RSQ_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)); e.spec_streams.push_back(st);
		cudaEvent_t ev; RSQ_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); e.spec_events.push_back(ev);
	}
	const uint32_t stride = ((e.max_n0_reads + 3u) & ~3u) + 1u;
	const size_t shmem_reads_max = static_cast<size_t>(kSpecReadWarps) * 32 * (stride * sizeof(double) + kSpecWindowBytes);
	auto scan_kernel = with_var ? k_spec_scan<true> : k_spec_scan<false>;
	auto reads_kernel = with_var ? k_spec_reads<true> : k_spec_reads<false>;
	RSQ_CUDA(cudaFuncSetAttribute(reads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(shmem_reads_max)));
	if(scan_shmem > 24 * 1024){ RSQ_CUDA(cudaFuncSetAttribute(scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(scan_shmem))); }   // 20 KB of static rings on top
	const double share = e.n_blocks_sim ? static_cast<double>(u_count) / e.n_blocks_sim : 0.0;
	uint64_t expected_reads = records ? records->em_n + 2048ull : static_cast<uint64_t>(2.0 * (e.total_pairs * share + sp.adapter_only_pairs) * 1.15) + 2048;
	float ms_sim = 0;
	uint32_t rounds = 0;
	for(int attempt = 0; ; ++attempt){
		sp.n_slabs = static_cast<uint32_t>(expected_reads / 32 + 3ull * sp.n_units + 64);
		e.d_spec_slots.alloc(static_cast<size_t>(sp.n_slabs) * 32 * sp.slot_stride);
		e.d_slab_next.alloc(sp.n_slabs); e.d_slab_count.alloc(sp.n_slabs);
		sp.slots = e.d_spec_slots.p; sp.slab_next = e.d_slab_next.p; sp.slab_count = e.d_slab_count.p;
		e.d_spec_counters.zero(s);
		choose(sp.n_units, 0.99);   // before the first poll has measured it: InDels are rare in every profile seen so far (E. coli: 58 ms starting deep, 60 ms starting at depth 8)
		EventTimer tm(s);
		tm.start();
		rounds = 0;
		if(sp.n_units){
			k_spec_init<<<(sp.n_units + 127) / 128, 128, 0, s>>>(c, sp, e.d_blocks.p, first_desc); ++e.launches;
			RSQ_CUDA(cudaEventRecord(e.ev_fork, s));
			for(uint32_t gi = 0; gi < n_groups; ++gi){ RSQ_CUDA(cudaStreamWaitEvent(e.spec_streams[gi], e.ev_fork, 0)); }
			volatile uint32_t *h_done = reinterpret_cast<volatile uint32_t *>(e.h_spec_counters.p) + 1;
			volatile unsigned long long *h_stat = reinterpret_cast<volatile unsigned long long *>(e.h_spec_counters.p + 8);
			unsigned long long last_emitted = 0, last_verified = 0;
			uint32_t rounds_per_poll = 2;   // the first poll comes early: it measures how often the consumption assumption holds
			while(true){
				for(uint32_t r = 0; r < rounds_per_poll; ++r){
					for(uint32_t gi = 0; gi < n_groups; ++gi){
						const uint32_t u0 = static_cast<uint64_t>(sp.n_units) * gi / n_groups, u1 = static_cast<uint64_t>(sp.n_units) * (gi + 1) / n_groups;
						if(u1 == u0){ continue; }
						cudaStream_t gs = e.spec_streams[gi];
						scan_kernel<<<(u1 - u0 + kWarpsPerCta - 1) / kWarpsPerCta, kWarpsPerCta * 32, scan_shmem, gs>>>(c, sp, e.d_blocks.p, first_desc, u0, u1);
						const size_t extra = (sp.n_units > sp.n_blocks && u1 == sp.n_units && sp.depth > sp.map_depth) ? sp.depth - sp.map_depth : 0;
						const size_t warps = (static_cast<size_t>(u1 - u0) * sp.map_depth + extra + lanes - 1) / lanes;
						const size_t shmem_reads = static_cast<size_t>(kSpecReadWarps) * lanes * (stride * sizeof(double) + kSpecWindowBytes);
						reads_kernel<<<static_cast<unsigned>((warps + kSpecReadWarps - 1) / kSpecReadWarps), kSpecReadWarps * 32, shmem_reads, gs>>>(c, sp, stride, lanes, u0, u1);
						e.launches += 2;
					}
				}
				rounds += rounds_per_poll;
				rounds_per_poll = 4;
				// every group has queued its rounds: join, then look at the number of finished units
				for(uint32_t gi = 0; gi < n_groups; ++gi){
					RSQ_CUDA(cudaEventRecord(e.spec_events[gi], e.spec_streams[gi]));
					RSQ_CUDA(cudaStreamWaitEvent(s, e.spec_events[gi], 0));
				}
				RSQ_CUDA(cudaMemcpyAsync(e.h_spec_counters.p, e.d_spec_counters.p, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
				RSQ_CUDA(cudaEventRecord(e.ev_fork, s));
				RSQ_CUDA(cudaStreamSynchronize(s));
				if(trace){ fprintf(stderr, "[spec] rounds=%u run_depth=%u lanes=%u budget=%u done=%u/%u emitted=%llu verified=%llu\n", rounds, sp.run_depth, lanes, sp.scan_budget, *h_done, sp.n_units, (unsigned long long)h_stat[0], (unsigned long long)h_stat[1]); }
				if(*h_done >= sp.n_units){ break; }
				{
					const unsigned long long em = h_stat[0], ve = h_stat[1];
					// share of reads whose assumption held, from the reads behind the first one of each unit and round (that one is always exact)
					double p_hold = 0.9;
					if(em > last_emitted){
						const double frac = static_cast<double>(ve - last_verified) / static_cast<double>(em - last_emitted);   // verified share at the depth just run
						const double d = sp.run_depth;
						// invert (1 - p^d) / ((1 - p) d) = frac by bisection
						double lo = 0.0, hi = 1.0;
						for(int it = 0; it < 40; ++it){
							const double mid = 0.5 * (lo + hi);
							const double f = mid >= 0.9999 ? 1.0 : (1.0 - std::pow(mid, d)) / ((1.0 - mid) * d);
							if(f < frac){ lo = mid; } else{ hi = mid; }
						}
						p_hold = 0.5 * (lo + hi);
					}
					last_emitted = em; last_verified = ve;
					choose(sp.n_units - *h_done, p_hold);
				}
				for(uint32_t gi = 0; gi < n_groups; ++gi){ RSQ_CUDA(cudaStreamWaitEvent(e.spec_streams[gi], e.ev_fork, 0)); }
				if(rounds > 100000000u){ throw std::runtime_error("speculative simulation does not terminate"); }
			}
		}
		ms_sim = tm.stop();
		RSQ_CUDA(cudaGetLastError());
		const uint32_t flag = read_error_flag(e);
		if(flag == kErrArenaFull && attempt < 3){
			expected_reads *= 2; e.d_error_flag.zero(s);
			continue;
		}
		if(flag & kErrSpecOverflow){
			// a read drew more InDels than the look-ahead margin covers: redo the run on the serial kernel
			e.d_error_flag.zero(s);
			return false;
		}
		if(flag){ throw std::runtime_error("device reported: " + describe_flag(flag)); }
		break;
	}
	res.rounds = rounds; res.ms_sim = ms_sim; e.spec_units_last = sp.n_units;
	// ordered FASTQ text
	EventTimer tm(s);
	tm.start();
	const uint32_t slots = sp.n_units;
	e.d_block_out.alloc(slots + 1);
	e.d_offsets.alloc(2ull * (slots + 1)); e.d_totals.alloc(4);
	if(slots){ k_spec_block_out<<<(slots + 127) / 128, 128, 0, s>>>(sp, e.d_block_out.p); ++e.launches; }
	k_block_offsets<<<1, 1024, 0, s>>>(e.d_block_out.p, slots, e.d_offsets.p, e.d_totals.p); ++e.launches;
	unsigned long long totals[4];
	RSQ_CUDA(cudaMemcpyAsync(totals, e.d_totals.p, sizeof totals, cudaMemcpyDeviceToHost, s));
	RSQ_CUDA(cudaStreamSynchronize(s));
	res.bytes[0] = totals[0]; res.bytes[1] = totals[1]; res.pairs = totals[2]; res.draws = totals[3];
	e.d_out_batch[par][0].alloc(totals[0] + 16); e.d_out_batch[par][1].alloc(totals[1] + 16);
	if(slots){ k_spec_gather<<<2 * slots, 128, 0, s>>>(sp, e.d_offsets.p, e.d_out_batch[par][0].p, e.d_out_batch[par][1].p); ++e.launches; }
	res.ms_gather = tm.stop();
	RSQ_CUDA(cudaGetLastError());
	return true;
}

// The same batch on the serial kernel (one warp per SimBlock): bisulfite runs, RSQ_SIM_PATH=serial, fallback of the speculative path.
static void simulate_serial_batch(rsq_engine &e, uint32_t u_begin, uint32_t u_count, bool with_adapter_only, int par, BatchResult &res){
	cudaStream_t s = e.stream;
	SimCtx &c = e.ctx;
	const bool meth = c.meth_loaded != 0;
	const uint32_t slots = u_count + (with_adapter_only ? 1 : 0);
	const uint32_t scratch = scratch_bytes(e.max_n0, c.max_org_len, c.max_read_len);
	const size_t shmem = static_cast<size_t>(scratch) * kWarpsPerCta + (c.var.loaded ? static_cast<size_t>(kWarpsPerCta) * 2 * c.var.num_alleles * sizeof(uint16_t) : 0);
	auto kernel = meth ? k_simulate<true> : k_simulate<false>;
	RSQ_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(shmem)));
	RSQ_CUDA(cudaFuncSetAttribute(k_adapter_only, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(scratch)));
	int dev_sms = 0; RSQ_CUDA(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, e.device));
	int ctas_per_sm = 0; RSQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kernel, kWarpsPerCta * 32, shmem));
	if(ctas_per_sm < 1){ throw std::runtime_error("k_simulate does not fit on an SM"); }
	// expected output: this batch's share of the pairs, generously padded
	const double share = e.n_blocks_sim ? static_cast<double>(u_count) / e.n_blocks_sim : 0.0;
	uint64_t expected = static_cast<uint64_t>((e.total_pairs * share + (with_adapter_only ? e.adapter_only_pairs : 0) + 1000) * (2.0 * (2.0 * c.max_read_len + 160.0)) * 1.3);
	for(int attempt = 0; ; ++attempt){
		Arena a;
		setup_arena(e, a, expected, slots);
		e.d_block_out.alloc(slots + 1);
		RSQ_CUDA(cudaMemsetAsync(e.d_block_out.p, 0, (slots + 1) * sizeof(BlockOut), s));
		e.d_next_block.alloc(1); e.d_next_block.zero(s);
		EventTimer tm(s);
		tm.start();
		if(u_count){
			const uint32_t ctas = std::min<uint32_t>((u_count + kWarpsPerCta - 1) / kWarpsPerCta, dev_sms * ctas_per_sm);
			kernel<<<ctas, kWarpsPerCta * 32, shmem, s>>>(c, e.d_blocks.p, e.shard_first + u_begin, u_count, a, e.d_block_out.p, e.d_next_block.p, e.max_n0, scratch);
			++e.launches;
		}
		if(with_adapter_only){
			k_adapter_only<<<1, 32, scratch, s>>>(c, e.adapter_only_seed, e.adapter_only_pairs, a, e.d_block_out.p, u_count, e.max_n0);
			++e.launches;
		}
		res.ms_sim = tm.stop();
		RSQ_CUDA(cudaGetLastError());
		const uint32_t flag = read_error_flag(e);
		if(flag == kErrArenaFull && attempt < 3){
			expected *= 2; e.d_error_flag.zero(s);
			continue;
		}
		if(flag){ throw std::runtime_error("device reported: " + describe_flag(flag)); }
		// ordered FASTQ text
		EventTimer tg(s);
		tg.start();
		e.d_offsets.alloc(2ull * (slots + 1)); e.d_totals.alloc(4);
		k_block_offsets<<<1, 1024, 0, s>>>(e.d_block_out.p, slots, e.d_offsets.p, e.d_totals.p); ++e.launches;
		unsigned long long totals[4];
		RSQ_CUDA(cudaMemcpyAsync(totals, e.d_totals.p, sizeof totals, cudaMemcpyDeviceToHost, s));
		RSQ_CUDA(cudaStreamSynchronize(s));
		res.bytes[0] = totals[0]; res.bytes[1] = totals[1]; res.pairs = totals[2]; res.draws = totals[3];
		e.d_out_batch[par][0].alloc(totals[0] + 16); e.d_out_batch[par][1].alloc(totals[1] + 16);
		if(slots){ k_gather<<<2 * slots, 128, 0, s>>>(e.d_block_out.p, slots, a, e.d_offsets.p, e.d_out_batch[par][0].p, e.d_out_batch[par][1].p); ++e.launches; }
		res.ms_gather = tg.stop();
		RSQ_CUDA(cudaGetLastError());
		break;
	}
}

// Host side of the output: the text of the batches arrives in order.  A writer thread copies it out of the device buffers in
// chunks of <= kRingChunk bytes through a small ring of pinned staging buffers (pinning host memory costs ~0.6 s per GB, so the
// ring is allocated once and kept small; the copy of chunk k+1 runs while chunk k is consumed) and appends the chunks to the
// FASTQ files (rsq_simulate) or copies them behind each other into ordinary host memory (engine API, runs of several batches)
// - all of it while the GPU is busy with the next batch.
constexpr size_t kRingChunk = 64u << 20;
constexpr int kRingSlots = 4;
constexpr size_t kDeflatePiece = 32u << 20;   // text per launch of the deflate kernels (256 members)
struct ChunkWriter {   // one per segment (first / second reads): two files, two threads
	int device = 0;
	cudaStream_t copy_stream = nullptr;
	TextSink *f = nullptr;              // file sink (compresses on the host cores when the name asks for it)
	HostBig *mem = nullptr;             // memory sink (sized in advance)
	std::thread th;
	std::mutex m; std::condition_variable cv;
	struct Batch { const unsigned char *src; uint64_t bytes, dst_off; cudaEvent_t ready; };
	std::deque<Batch> q; bool stop = false; std::string error;
	uint64_t batches_done = 0;
	bool discard = getenv("RSQ_DISCARD_OUTPUT") != nullptr;   // throughput probes of runs larger than the disk: the text reaches host memory, not the file
	PinnedBuf *ring = nullptr; cudaEvent_t *ev = nullptr; int n_slots = 2;
	// device-side gzip: the writer launches the deflate kernels on its copy stream and only the members cross PCIe
	struct Deflate { uint32_t *slots, *tokens, *sizes; const uint32_t *crc; uint8_t *out; unsigned long long *total; int ctas; } dfl = {};
	bool device_gzip = false;
	void deflate_batch(const Batch &j){
		for(uint64_t off = 0; off < j.bytes && error.empty(); off += kDeflatePiece){
			const uint64_t n = std::min<uint64_t>(kDeflatePiece, j.bytes - off);
			const uint32_t n_members = static_cast<uint32_t>((n + dfl::kMember - 1) / dfl::kMember);
			k_deflate_members<<<std::min<uint32_t>(n_members, dfl.ctas), 256, sizeof(dfl::Shared), copy_stream>>>(j.src + off, n, n_members, dfl.slots, dfl.tokens, dfl.sizes, dfl.crc);
			k_deflate_compact<<<n_members, 256, 0, copy_stream>>>(dfl.slots, dfl.sizes, n_members, dfl.out, dfl.total);
			unsigned long long total = 0;
			if(cudaMemcpyAsync(&total, dfl.total, sizeof total, cudaMemcpyDeviceToHost, copy_stream) != cudaSuccess || cudaStreamSynchronize(copy_stream) != cudaSuccess){
				error = std::string("compressing FASTQ text on the device failed: ") + cudaGetErrorString(cudaGetLastError()); return;
			}
			for(uint64_t done = 0; done < total; done += kRingChunk){
				const uint64_t part = std::min<uint64_t>(kRingChunk, total - done);
				if(cudaMemcpyAsync(ring[0].p, dfl.out + done, part, cudaMemcpyDeviceToHost, copy_stream) != cudaSuccess || cudaStreamSynchronize(copy_stream) != cudaSuccess){
					error = "copying compressed FASTQ text to the host failed"; return;
				}
				if(!discard && !f->write_members(ring[0].p, part, done ? 0 : n, done ? 0 : n_members)){ error = "Could not write records to the output file"; return; }
			}
		}
	}
	void consume(int slot, uint64_t n, uint64_t dst){
		if(cudaEventSynchronize(ev[slot]) != cudaSuccess){ error = "copying FASTQ text to the host failed"; return; }
		if(mem){ std::memcpy(mem->data() + dst, ring[slot].p, n); }
		else if(!discard && error.empty() && !f->write(ring[slot].p, n)){ error = "Could not write records to the output file"; }
	}
	void run(){
		cudaSetDevice(device);
		while(true){
			Batch j;
			{ std::unique_lock<std::mutex> l(m); cv.wait(l, [&]{ return stop || !q.empty(); }); if(q.empty()){ return; } j = q.front(); q.pop_front(); }
			cudaStreamWaitEvent(copy_stream, j.ready, 0);
			if(device_gzip){
				deflate_batch(j);
				{ std::lock_guard<std::mutex> l(m); ++batches_done; }
				cv.notify_all();
				continue;
			}
			struct Piece { int slot; uint64_t n, dst; };
			std::deque<Piece> inflight;
			int slot = 0;
			for(uint64_t off = 0; off < j.bytes; off += kRingChunk){
				const uint64_t n = std::min<uint64_t>(kRingChunk, j.bytes - off);
				if(static_cast<int>(inflight.size()) == n_slots){ const Piece p = inflight.front(); inflight.pop_front(); consume(p.slot, p.n, p.dst); }
				if(cudaMemcpyAsync(ring[slot].p, j.src + off, n, cudaMemcpyDeviceToHost, copy_stream) != cudaSuccess){ error = "copying FASTQ text to the host failed"; }
				cudaEventRecord(ev[slot], copy_stream);
				inflight.push_back({slot, n, j.dst_off + off});
				slot = (slot + 1) % n_slots;
			}
			while(!inflight.empty()){ const Piece p = inflight.front(); inflight.pop_front(); consume(p.slot, p.n, p.dst); }
			{ std::lock_guard<std::mutex> l(m); ++batches_done; }
			cv.notify_all();
		}
	}
};

static void simulate(rsq_engine &e, rsq_sim_report *rep){
	if(!e.prepared){ throw std::runtime_error("rsq_engine_prepare has not been called"); }
	cudaStream_t s = e.stream;
	SimCtx &c = e.ctx;
	e.downloaded = false; e.streamed_to_host = false;
	e.spec_rounds = 0; e.spec_depth = 0;
	e.out_bytes[0] = e.out_bytes[1] = 0; e.out_pairs = 0; e.out_draws = 0;
	const bool meth = c.meth_loaded != 0;
	// RSQ_SIM_PATH=serial keeps the one-warp-per-SimBlock kernel (parity tests run both)
	const char *path = getenv("RSQ_SIM_PATH");
	bool spec = !(path && std::string(path) == "serial");
	// ---- batches: as many blocks as HBM holds speculation state, record slots and two output buffers for ----
	int dev_sms = 0; RSQ_CUDA(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, e.device));
	const uint32_t max_rl = std::max(c.read_len_to[0], c.read_len_to[1]);
	const double reads_per_block = e.n_blocks_sim ? 2.0 * e.total_pairs / e.n_blocks_sim : 0.0;
	const double rec_bytes = 2.0 * max_rl + 160.0;
	auto unit_bytes = [&](uint32_t depth){
		return 2.0 * (depth + 1) * sizeof(SpecSnap) + depth * (sizeof(ReadJob) + 8.0 * (3 * max_rl + 8 + kSpecMargin))
		       + 1.2 * reads_per_block * (rec_bytes + 250.0) + 2.0 * 1.1 * reads_per_block * rec_bytes + 256.0 + ((meth || c.var.loaded) ? kConvSlots * 2.0 * kMaxOrgLen : 0.0)
		       + (c.var.loaded ? 2.0 * (depth + 1) * 4.0 * c.var.num_alleles : 0.0);
	};
	size_t free_b = 0, total_b = 0; RSQ_CUDA(cudaMemGetInfo(&free_b, &total_b));
	double budget = 0.85 * (static_cast<double>(free_b) + e.reusable_bytes());
	uint32_t depth_cap = 32;
	uint64_t per_batch = e.shard_n;
	// A run that needs several batches does not keep the surrounding biases of the whole reference (16 bytes per base, needed everywhere only by the bias
	// sums of the prologue): every batch gets the window of positions its blocks can touch, recomputed in front of it (k_surroundings, microseconds).
	// What stays resident per base is then 1 (bases) + 4 (G/C prefix) + 2 x 2 (systematic errors) bytes.  RSQ_KEEP_STAGES=1 keeps the arrays (stage fetches).
	double window_bytes = 0.0;
	e.sur_window_max = 0;
	if(e.sur_windowed || ((unit_bytes(32) * e.shard_n > budget || getenv("RSQ_SUR_WINDOW")) && !getenv("RSQ_KEEP_STAGES"))){   // (a second simulate call on the same prepare stays windowed)
		RSQ_CUDA(cudaDeviceSynchronize());
		e.d_sur_start.release(); e.d_sur_end.release();
		c.sur_start = nullptr; c.sur_end = nullptr;
		e.sur_windowed = true;
		window_bytes = 16.0 * 1000.0;
		RSQ_CUDA(cudaMemGetInfo(&free_b, &total_b));
		budget = 0.85 * (static_cast<double>(free_b) + e.reusable_bytes());
	}

	// (twice the depth per round, RSQ_SPEC_CAP=64, was measured on E. coli: four rounds fewer, but each scan and each lock-step pass over the reads
	// of a round takes as much longer - 65 ms instead of 64; the capacity stays an option for profiles with rarer InDels)
	if(const char *env = getenv("RSQ_SPEC_CAP")){ if(!meth && !c.var.loaded && atoi(env) > 32){ depth_cap = kSpecMaxDepth; } }
	if((unit_bytes(32) + window_bytes) * e.shard_n > budget){ depth_cap = 16; per_batch = std::max<uint64_t>(1024, static_cast<uint64_t>((budget - 16.0 * (c.insert_to + 4096.0)) / (unit_bytes(16) + window_bytes))); }
	if(const char *env = getenv("RSQ_BATCH_UNITS")){ per_batch = std::max(1, atoi(env)); }
	const uint32_t n_batches = e.shard_n ? static_cast<uint32_t>((e.shard_n + per_batch - 1) / per_batch) : 1;
	const bool to_files = e.sink_files[0] != nullptr;
	const bool stream_host = !to_files && n_batches > 1;
	if(!e.copy_stream){ RSQ_CUDA(cudaStreamCreateWithFlags(&e.copy_stream, cudaStreamNonBlocking)); }
	for(int i = 0; i < 2; ++i){ if(!e.ev_out[i]){ RSQ_CUDA(cudaEventCreateWithFlags(&e.ev_out[i], cudaEventDisableTiming)); } }
	ChunkWriter writer[2];
	const bool streaming = to_files || stream_host;
	// declared before any writer thread starts: a throw between the two thread starts (an allocation for segment 1) must still stop and join segment 0's thread
	struct Joiner { ChunkWriter *w; bool on; ~Joiner(){ for(int seg = 0; on && seg < 2; ++seg){ if(w[seg].th.joinable()){ { std::lock_guard<std::mutex> l(w[seg].m); w[seg].stop = true; } w[seg].cv.notify_all(); w[seg].th.join(); } } } } joiner{writer, streaming};
	if(streaming){
		if(!e.copy_stream2){ RSQ_CUDA(cudaStreamCreateWithFlags(&e.copy_stream2, cudaStreamNonBlocking)); }
		for(int i = 0; i < kRingSlots; ++i){ e.h_ring[i].ensure(kRingChunk); if(!e.ev_ring[i]){ RSQ_CUDA(cudaEventCreateWithFlags(&e.ev_ring[i], cudaEventDisableTiming)); } }
		const double est = (e.total_pairs * (e.n_blocks_sim ? static_cast<double>(e.shard_n) / e.n_blocks_sim : 0.0) + e.adapter_only_pairs) * (max_rl * 2.0 + 90.0) * 1.02 + (1 << 20);
		for(int seg = 0; seg < 2; ++seg){
			ChunkWriter &w = writer[seg];
			w.ring = e.h_ring + 2 * seg; w.ev = e.ev_ring + 2 * seg; w.device = e.device; w.copy_stream = seg ? e.copy_stream2 : e.copy_stream;
			if(to_files){
				w.f = e.sink_files[seg];
				const char *gz_mode = getenv("RSQ_GZIP");
				if(w.f->compressed() && !(gz_mode && std::string(gz_mode) == "host")){   // RSQ_GZIP=host: zlib on the host cores (text_io.hpp)
					if(!e.d_dfl_crc.p){
						std::vector<uint32_t> tables(256 + 32);
						dfl::crc_make_table(tables.data()); dfl::crc_make_shift_operator(tables.data() + 256, dfl::kCrcPiece);
						e.d_dfl_crc.upload(tables, s);
						RSQ_CUDA(cudaStreamSynchronize(s));
						RSQ_CUDA(cudaFuncSetAttribute(k_deflate_members, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(dfl::Shared))));
					}
					const size_t piece_members = kDeflatePiece / dfl::kMember;
					e.d_dfl_slots[seg].alloc(piece_members * dfl::kSlotWords); e.d_dfl_out[seg].alloc(piece_members * dfl::kSlotBytes);
					// one wave: a CTA per member of a piece where the SMs hold them (3 CTAs per SM by shared memory), 512 KiB of token scratch each
					const int deflate_ctas = static_cast<int>(std::min<size_t>(piece_members, 2 * static_cast<size_t>(dev_sms)));
					e.d_dfl_tokens[seg].alloc(static_cast<size_t>(deflate_ctas) * dfl::kMember); e.d_dfl_sizes[seg].alloc(piece_members); e.d_dfl_total[seg].alloc(1);
					w.dfl = {e.d_dfl_slots[seg].p, e.d_dfl_tokens[seg].p, e.d_dfl_sizes[seg].p, e.d_dfl_crc.p, e.d_dfl_out[seg].p, e.d_dfl_total[seg].p, deflate_ctas};
					w.device_gzip = true;
				}
			}
			else{ e.h_big[seg].resize(static_cast<size_t>(est)); w.mem = &e.h_big[seg]; }
			w.th = std::thread([&w]{ w.run(); });
		}
	}
	auto wait_batches = [&](uint64_t n){ for(int seg = 0; seg < 2; ++seg){ std::unique_lock<std::mutex> l(writer[seg].m); writer[seg].cv.wait(l, [&]{ return writer[seg].batches_done >= n; }); } };
	float ms_sim = 0, ms_gather = 0;
	for(uint32_t b = 0; b < n_batches; ++b){
		const uint32_t u_begin = static_cast<uint32_t>(std::min<uint64_t>(e.shard_n, b * per_batch));
		const uint32_t u_count = static_cast<uint32_t>(std::min<uint64_t>(per_batch, e.shard_n - u_begin));
		const bool with_ao = e.shard_has_adapter_only && b + 1 == n_batches;
		const int par = b & 1;
		if(streaming && b >= 2){ wait_batches(b - 1); }   // the text of batch b-2 has left this device buffer
		BatchResult res;
		bool done = false;
		stage_log("simulate: batch start");
		if(e.sur_windowed && u_count){
			// positions the blocks [B0, B1) of this batch can touch: per sequence from the first block's start to insert_to behind the last block's end
			const uint64_t B0 = static_cast<uint64_t>(e.shard_first) + u_begin, B1 = B0 + u_count;
			struct Piece { size_t seq; uint32_t p0, n; };
			std::vector<Piece> pieces;
			uint64_t G0 = 0, G1 = 0;
			for(size_t i = 0; i < e.plan.seq_blocks.size(); ++i){
				const uint64_t first = e.plan.seq_first_block[i], last = first + e.plan.seq_blocks[i];
				if(!e.plan.seq_blocks[i] || last <= B0 || first >= B1){ continue; }
				const uint32_t L = e.genome.seqs[i].size();
				const uint32_t p0 = static_cast<uint32_t>((std::max(B0, first) - first) * 1000u);
				const uint32_t p1 = static_cast<uint32_t>(std::min<uint64_t>(L, (std::min(B1, last) - first) * 1000ull + c.insert_to));
				if(pieces.empty()){ G0 = e.h_seq_off[i] + p0; }
				G1 = e.h_seq_off[i] + p1;
				pieces.push_back({i, p0, p1 - p0});
			}
			e.d_sur_start.alloc(G1 - G0 + 1); e.d_sur_end.alloc(G1 - G0 + 1);
			e.sur_window_max = std::max<uint64_t>(e.sur_window_max, G1 - G0);
			for(const Piece &pc : pieces){
				const uint64_t at = e.h_seq_off[pc.seq] + pc.p0 - G0;
				k_surroundings<<<(pc.n + 255) / 256, 256, 0, s>>>(e.d_ref.p + e.h_seq_off[pc.seq], static_cast<uint32_t>(e.genome.seqs[pc.seq].size()), pc.p0, pc.n,
				                                                 e.d_sur_tab[0].p, e.d_sur_tab[1].p, e.d_sur_tab[2].p, e.d_sur_start.p + at, e.d_sur_end.p + at);
				++e.launches;
			}
			// the kernels index with the offset of the whole layout: bias the pointers by the window's start
			c.sur_start = reinterpret_cast<const double *>(reinterpret_cast<uintptr_t>(e.d_sur_start.p) - G0 * sizeof(double));
			c.sur_end = reinterpret_cast<const double *>(reinterpret_cast<uintptr_t>(e.d_sur_end.p) - G0 * sizeof(double));
		}
		if(spec){
			done = simulate_spec_batch(e, u_begin, u_count, with_ao, depth_cap, par, res);
			if(done){ e.spec_rounds += res.rounds; if(!e.spec_depth){ e.spec_depth = res.depth; } }
		}
		if(!done){ simulate_serial_batch(e, u_begin, u_count, with_ao, par, res); }
		ms_sim += res.ms_sim; ms_gather += res.ms_gather;
		stage_log("simulate: batch kernels done");
		if(streaming){
			RSQ_CUDA(cudaEventRecord(e.ev_out[par], s));
			for(int seg = 0; seg < 2; ++seg){
				ChunkWriter &w = writer[seg];
				if(w.mem && e.out_bytes[seg] + res.bytes[seg] > w.mem->size()){   // the estimate was too small: grow once the writer is idle
					wait_batches(b);
					w.mem->resize(static_cast<size_t>((e.out_bytes[seg] + res.bytes[seg]) * 1.3) + (1 << 20));
				}
				ChunkWriter::Batch job{e.d_out_batch[par][seg].p, res.bytes[seg], e.out_bytes[seg], e.ev_out[par]};
				{ std::lock_guard<std::mutex> l(w.m); w.q.push_back(job); }
				w.cv.notify_all();
			}
		}
		e.out_bytes[0] += res.bytes[0]; e.out_bytes[1] += res.bytes[1]; e.out_pairs += res.pairs; e.out_draws += res.draws;
		e.last_par = par;
	}
	stage_log("simulate: last batch delivered");
	if(streaming){
		wait_batches(n_batches);
		for(int seg = 0; seg < 2; ++seg){ if(!writer[seg].error.empty()){ throw std::runtime_error(writer[seg].error); } }
		e.streamed_to_host = true;
	}
	stage_log("simulate: copies and writer drained");
	if(rep){ rep->ms_gather = ms_gather; }
	fill_simulate_report(e, rep, ms_sim);
	e.group_pairs = e.out_pairs;
	if(e.comm){   // read pairs of the whole run: the one number the engines of a group exchange behind the data path
		unsigned long long *d_cnt = reinterpret_cast<unsigned long long *>(e.d_totals.p);
		unsigned long long mine = e.out_pairs;
		RSQ_CUDA(cudaMemcpyAsync(d_cnt, &mine, 8, cudaMemcpyHostToDevice, s));
		RSQ_NCCL(nccl_api().AllReduce(d_cnt, d_cnt, 1, ncclUint64, ncclSum, e.comm, s));
		RSQ_CUDA(cudaMemcpyAsync(&mine, d_cnt, 8, cudaMemcpyDeviceToHost, s));
		RSQ_CUDA(cudaStreamSynchronize(s));
		e.group_pairs = mine;
	}
	if(rep){
		rep->group_pairs = e.group_pairs; rep->group_world = e.comm ? e.group_world : 1; rep->shard_first = e.shard_first;
		rep->batches = n_batches;
		// bases + G/C prefix counts + systematic errors of both strands (2 x 2 bytes) + surrounding biases (the whole reference, or the largest batch window)
		const double per_base = static_cast<double>(e.total_size) * (1.0 + 4.0 + 4.0) + 16.0 * static_cast<double>(e.sur_windowed ? e.sur_window_max : e.total_size);
		rep->resident_bytes_per_base = e.total_size ? per_base / static_cast<double>(e.total_size) : 0.0;
	}
	stage_log("simulate: report");
}

static void download(rsq_engine &e, rsq_sim_report *rep){
	cudaStream_t s = e.stream;
	EventTimer tm(s);
	tm.start();
	if(!e.streamed_to_host){   // single batch: its text is still on the device
		for(int seg = 0; seg < 2; ++seg){
			e.h_out[seg].ensure(e.out_bytes[seg] + 1);
			if(e.out_bytes[seg]){ RSQ_CUDA(cudaMemcpyAsync(e.h_out[seg].p, e.d_out_batch[e.last_par][seg].p, e.out_bytes[seg], cudaMemcpyDeviceToHost, s)); }
		}
	}
	const float ms = tm.stop();
	if(rep){ rep->ms_download = ms; }
	e.downloaded = true;
}

// seqToIllumina host part: FASTA records -> device batches
static void apply_error_model(rsq_engine &e, const char *in_path, const char *out_path, uint64_t seed, rsq_sim_report *rep){
	cudaStream_t s = e.stream;
	const Profile &p = e.prof;
	SimCtx &c = e.ctx;
	e.launches = 0; e.d_error_flag.zero(s);
	// read records (SeqAn FASTA semantics: id = header without '>', sequence = DnaString: non-ACGTU -> A); em_input.hpp
	EmInput input = read_em_input(in_path);
	std::vector<EmRecord> &recs = input.recs;
	std::vector<uint8_t> &hseq = input.seq, &hdom = input.dom, &hrate = input.rate; std::string &hid = input.ids;
	const uint32_t max_len = input.max_len;
	if(max_len > kMaxOrgLen){ throw std::runtime_error("input fragments longer than " + std::to_string(kMaxOrgLen) + " bases are not supported by this build"); }
	// sys_gc_range + adapter systematic errors from the master stream, then one seed per 10000-record batch
	// The reference never initialises sys_gc_range_ on this path (it is only set inside Simulate / SimulateErrorModelOnly,
	// Simulator.cpp:2782, 2962; CreateSystematicErrorProfile runs on a fresh Simulator object): the GC window length is
	// whatever the stack held.  In the reference build of this image every value >= ~2*10^4 reproduces its output; we use
	// the largest uintReadLen, i.e. "all bases seen so far, at most 65535".
	e.sys_gc_range = 65535;
	e.d_master_state.alloc(kMtN + 1);
	k_master_seed<<<1, 32, 0, s>>>(e.d_master_state.p, seed); ++e.launches;
	uint32_t carried = 0;
	{
		std::vector<SysChain> chains; std::vector<std::pair<uint32_t, uint32_t>> lens; uint64_t n_draws = 0;
		struct Ad { int seg; size_t a; uint64_t raw_off; }; std::vector<Ad> order;
		for(int seg = 2; seg--; ){ for(size_t a = p.adapter_count_sum[seg].size(); a--; ){ if(!p.adapter_count_sum[seg][a]){ continue; } order.push_back({seg, a, n_draws}); n_draws += 2ull * (e.h_adapter_off[seg][a + 1] - e.h_adapter_off[seg][a]); } }
		if(n_draws){
			e.d_master.alloc(n_draws);
			k_master_stream<<<1, kMasterThreads, 0, s>>>(e.d_master_state.p, e.d_master_state.p, e.d_master.p, n_draws); ++e.launches;
			for(const auto &o : order){
				const uint32_t off = e.h_adapter_off[o.seg][o.a], len = e.h_adapter_off[o.seg][o.a + 1] - off;
				SysChain ch{}; ch.seq = e.d_adapter_seq.p + off; ch.L = len; ch.raw = e.d_master.p + o.raw_off; ch.out = e.d_adapter_sys.p + 2 * off; ch.carried_dom = carried;
				chains.push_back(ch); lens.push_back({len, 0});
				carried = dominant_before(e.h_adapter_seq.data() + off, len, false, len, carried);
			}
			uint32_t passes = 0;
			run_sys_chains(e, chains, lens, 1u << 30, 0, passes);
		}
	}
	// ErrorModelOnlyThread hands out batches of 10000 records, one master-stream seed each (RSQ_EM_BATCH: smaller batches for the
	// serial-vs-speculative consistency test; the reference itself cannot get past its first batch)
	const uint32_t batch = getenv("RSQ_EM_BATCH") ? std::max(1, atoi(getenv("RSQ_EM_BATCH"))) : 10000;
	const uint32_t n_batches = (recs.size() + batch - 1) / batch;
	DevBuf<uint64_t> d_seeds; d_seeds.alloc(n_batches);
	k_master_stream<<<1, kMasterThreads, 0, s>>>(e.d_master_state.p, e.d_master_state.p, d_seeds.p, n_batches); ++e.launches;
	DevBuf<EmRecord> d_recs; d_recs.upload(recs, s);
	DevBuf<uint8_t> d_seq, d_dom, d_rate; d_seq.upload(hseq, s); d_dom.upload(hdom, s); d_rate.upload(hrate, s);
	DevBuf<char> d_ids; d_ids.upload(hid.data(), hid.size() + 1, s);
	float ms_sim = 0;
	bool done = false;
	{
		// speculative two-phase kernels: every batch is a unit whose reads run in lock step with those of the other batches
		// (RSQ_SIM_PATH=serial: one warp per batch)
		const char *path = getenv("RSQ_SIM_PATH");
		if(!(path && std::string(path) == "serial")){
			std::vector<uint8_t> hsys(2 * hdom.size());
			for(size_t i = 0; i < hdom.size(); ++i){ hsys[2 * i] = hdom[i]; hsys[2 * i + 1] = hrate[i]; }
			DevBuf<uint8_t> d_sys; d_sys.upload(hsys, s);
			SpecCtx em{};
			em.em_recs = d_recs.p; em.em_n = static_cast<uint32_t>(recs.size()); em.em_batch = batch; em.em_seeds = d_seeds.p;
			em.em_seq = d_seq.p; em.em_sys = d_sys.p; em.em_ids = d_ids.p;
			e.em_max_id_len = 0; for(const auto &r : recs){ e.em_max_id_len = std::max(e.em_max_id_len, r.id_len); }
			if(e.em_max_id_len + 1 + kCigarCap + 12 <= static_cast<uint32_t>(kIdCap)){
				BatchResult res;
				done = simulate_spec_batch(e, 0, n_batches, false, 32, 0, res, &em);
				if(done){
					ms_sim = res.ms_sim; e.last_par = 0; e.streamed_to_host = false;
					e.out_bytes[0] = res.bytes[0]; e.out_bytes[1] = res.bytes[1]; e.out_pairs = res.pairs; e.out_draws = 0;
					e.spec_rounds = res.rounds; e.spec_depth = res.depth;
					if(rep){ rep->ms_gather = res.ms_gather; rep->spec_rounds = res.rounds; rep->spec_depth = res.depth; }
				}
			}
			RSQ_CUDA(cudaStreamSynchronize(s));   // d_sys goes out of scope
		}
	}
	if(!done){
		const uint32_t scratch = scratch_bytes(e.max_n0, c.max_org_len, c.max_read_len);
		const size_t shmem = static_cast<size_t>(scratch) * kWarpsPerCta;
		RSQ_CUDA(cudaFuncSetAttribute(k_error_model, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(shmem)));
		Arena a;
		uint64_t expected = static_cast<uint64_t>(recs.size() * (2.0 * c.max_read_len + 200.0 + hid.size() / std::max<size_t>(1, recs.size())) * 1.3);
		for(int attempt = 0; ; ++attempt){
			setup_arena(e, a, expected, n_batches);
			e.d_block_out.alloc(n_batches + 1);
			RSQ_CUDA(cudaMemsetAsync(e.d_block_out.p, 0, (n_batches + 1) * sizeof(BlockOut), s));
			e.d_next_block.alloc(1); e.d_next_block.zero(s);
			EventTimer tm(s); tm.start();
			k_error_model<<<(n_batches + kWarpsPerCta - 1) / kWarpsPerCta, kWarpsPerCta * 32, shmem, s>>>(c, d_recs.p, recs.size(), batch, d_seeds.p, n_batches, d_seq.p, d_dom.p, d_rate.p, d_ids.p, a, e.d_block_out.p, e.d_next_block.p, e.max_n0, scratch);
			++e.launches;
			ms_sim = tm.stop();
			RSQ_CUDA(cudaGetLastError());
			const uint32_t flag = read_error_flag(e);
			if(flag == kErrArenaFull && attempt < 3){ expected *= 2; e.d_error_flag.zero(s); continue; }
			if(flag){ throw std::runtime_error("device reported: " + describe_flag(flag)); }
			break;
		}
		gather(e, a, n_batches, rep);
	}
	download(e, rep);
	RSQ_CUDA(cudaStreamSynchronize(s));
	TextSink o;
	if(!o.open(out_path)){ throw std::runtime_error(std::string("Could not open '") + out_path + "' for writing."); }
	if(!o.write(e.h_out[0].p, e.out_bytes[0]) || !o.close()){ throw std::runtime_error(std::string("Could not write records to '") + out_path + "'"); }
	if(rep){ rep->ms_simulate = ms_sim; rep->pairs = e.out_pairs; rep->bytes[0] = e.out_bytes[0]; rep->bytes[1] = 0; rep->blocks = n_batches; rep->kernel_launches = e.launches; }
}

}  // namespace rsq

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
#define RSQ_TRY try{
#define RSQ_CATCH(ret) }catch(const std::exception &ex){ set_error("%s", ex.what()); return ret; }catch(...){ set_error("unknown error"); return ret; }

extern "C" {

const char *rsq_last_error(void){ return g_last_error.c_str(); }

int rsq_device_count(void){ int n = 0; if(cudaGetDeviceCount(&n) != cudaSuccess){ return 0; } return n; }

rsq_profile *rsq_profile_load_flat(const char *flat_path){
	RSQ_TRY
	std::unique_ptr<rsq_profile> p(new rsq_profile);
	FlatFile f; f.load(flat_path);
	p->p.from_flat(f);
	return p.release();
	RSQ_CATCH(nullptr)
}

rsq_profile *rsq_profile_load(const char *stats_path, const char *ipf_path){
	RSQ_TRY
	std::unique_ptr<rsq_profile> p(new rsq_profile);
	load_reseq_profile(p->p, stats_path, ipf_path);
	return p.release();
	RSQ_CATCH(nullptr)
}

int rsq_profile_save_flat(const rsq_profile *profile, const char *flat_path){
	RSQ_TRY
	FlatFile f; profile->p.to_flat(f); f.save(flat_path);
	return 0;
	RSQ_CATCH(1)
}

void rsq_profile_free(rsq_profile *profile){ delete profile; }

// LogArrayResult::SetPar0 (ProbabilityEstimates.h:547-556)
static void set_par0(HostTable &h, uint32_t value){
	h.par0.assign(1, value);
	for(uint32_t n = 0; n < h.nm; ++n){ h.dim2[n].assign(h.to[n] - h.from[n], 1.0); }
}

int rsq_profile_remove_indel_errors(rsq_profile *profile){
	RSQ_TRY
	const uint32_t T = profile->p.num_tiles, base = 50 * T + 120;
	for(uint32_t i = 0; i < 12; ++i){ set_par0(profile->p.tables.at(base + i), 0); }
	return 0;
	RSQ_CATCH(1)
}

int rsq_profile_remove_substitution_errors(rsq_profile *profile){
	RSQ_TRY
	const uint32_t T = profile->p.num_tiles, base = 10 * T;
	for(uint32_t st = 0; st < 2 * T; ++st){ for(uint32_t b = 0; b < 4; ++b){ for(uint32_t d = 0; d < 5; ++d){ set_par0(profile->p.tables.at(base + (st * 4 + b) * 5 + d), b); } } }
	return 0;
	RSQ_CATCH(1)
}

int rsq_profile_change_error_rate(rsq_profile *profile, double error_multiplier){
	RSQ_TRY
	if(!(error_multiplier > 0.0)){ throw std::runtime_error("--error_multiplier must be a positive value."); }
	const uint32_t T = profile->p.num_tiles, base = 10 * T;
	const double multiplier = 1.0 / error_multiplier;
	for(uint32_t st = 0; st < 2 * T; ++st){ for(uint32_t b = 0; b < 4; ++b){ for(uint32_t d = 0; d < 5; ++d){
		HostTable &h = profile->p.tables.at(base + (st * 4 + b) * 5 + d);   // LogArrayResult::ModifyPar0(ref_base, 1/multiplier)
		size_t col = 0;
		while(col < h.par0.size() && h.par0[col] != b){ ++col; }
		if(col < h.par0.size()){ for(size_t i = col; i < h.dim2[0].size(); i += h.par0.size()){ h.dim2[0][i] *= multiplier; } }
	} } }
	return 0;
	RSQ_CATCH(1)
}

rsq_reference *rsq_reference_load_fasta(const char *fasta_path){
	RSQ_TRY
	std::unique_ptr<rsq_reference> r(new rsq_reference);
	r->g.read_fasta(fasta_path);
	return r.release();
	RSQ_CATCH(nullptr)
}

rsq_reference *rsq_reference_from_memory(uint32_t n_seqs, const char *const *ids, const char *const *bases, const uint64_t *lengths){
	RSQ_TRY
	std::unique_ptr<rsq_reference> r(new rsq_reference);
	for(uint32_t i = 0; i < n_seqs; ++i){
		r->g.ids.emplace_back(ids[i]);
		std::vector<uint8_t> s(lengths[i]);
		Genome::encode(bases[i], lengths[i], s.data());
		r->g.seqs.push_back(std::move(s));
	}
	if(!n_seqs){ throw std::runtime_error("reference does not contain any sequences"); }
	return r.release();
	RSQ_CATCH(nullptr)
}

int rsq_reference_load_methylation(rsq_reference *ref, const char *bed_path){
	RSQ_TRY
	ref->g.read_methylation(bed_path);
	return 0;
	RSQ_CATCH(1)
}

int rsq_reference_load_variants(rsq_reference *ref, const char *vcf_path){
	RSQ_TRY
	ref->g.read_variants(vcf_path);
	return 0;
	RSQ_CATCH(1)
}
uint32_t rsq_reference_num_alleles(const rsq_reference *ref){ return ref->g.variants.num_alleles; }
uint64_t rsq_reference_num_variants(const rsq_reference *ref, uint32_t seq){
	return seq < ref->g.variants.variants.size() ? ref->g.variants.variants[seq].size() : 0;
}
int rsq_reference_variants(const rsq_reference *ref, uint32_t seq, uint64_t capacity, uint32_t *position, uint32_t *bases_off,
                           uint64_t *allele_lo, uint64_t *allele_hi, uint8_t *bases, uint64_t bases_capacity){
	RSQ_TRY
	if(seq >= ref->g.variants.variants.size()){ throw std::runtime_error("rsq_reference_variants: no variants loaded for this sequence"); }
	const auto &vars = ref->g.variants.variants[seq];
	uint64_t n_bases = 0;
	for(const auto &v : vars){ n_bases += v.var_seq.size(); }
	if(vars.size() > capacity || n_bases > bases_capacity){ throw std::runtime_error("rsq_reference_variants: destination too small"); }
	uint32_t off = 0;
	for(size_t k = 0; k < vars.size(); ++k){
		position[k] = vars[k].position;
		bases_off[k] = off;
		allele_lo[k] = vars[k].allele[0];
		allele_hi[k] = vars[k].allele[1];
		for(uint8_t b : vars[k].var_seq){ bases[off++] = b; }
	}
	bases_off[vars.size()] = off;
	return 0;
	RSQ_CATCH(1)
}

uint64_t rsq_reference_total_size(const rsq_reference *ref){ return ref->g.total_size(); }
uint32_t rsq_reference_num_sequences(const rsq_reference *ref){ return ref->g.seqs.size(); }
void rsq_reference_free(rsq_reference *ref){ delete ref; }

rsq_engine *rsq_engine_create(const rsq_profile *profile, int device){
	RSQ_TRY
	int n = 0;
	if(cudaGetDeviceCount(&n) != cudaSuccess || n <= device || device < 0){ throw std::runtime_error("no usable CUDA device (this engine has no CPU path)"); }
	RSQ_CUDA(cudaSetDevice(device));
	std::unique_ptr<rsq_engine> e(new rsq_engine);
	e->device = device;
	RSQ_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
	RSQ_CUDA(cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking));
	RSQ_CUDA(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
	RSQ_CUDA(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
	e->prof = profile->p;
	upload_profile(*e);
	return e.release();
	RSQ_CATCH(nullptr)
}

void rsq_engine_destroy(rsq_engine *engine){ if(engine){ cudaSetDevice(engine->device); delete engine; } }

int rsq_shard_plan(const rsq_profile *profile, const rsq_reference *ref, uint32_t shard_count, uint32_t *boundaries){
	RSQ_TRY
	if(!profile || !ref || !boundaries || !shard_count){ throw std::runtime_error("rsq_shard_plan: null argument or zero shards"); }
	const Genome &g = ref->g;
	std::vector<uint64_t> lengths(g.seqs.size());
	for(size_t i = 0; i < g.seqs.size(); ++i){ lengths[i] = g.seqs[i].size(); }
	const ShardPlan plan = make_shard_plan(lengths.data(), lengths.size(), static_cast<uint32_t>(profile->p.insert_lengths.to()), shard_count);
	for(uint32_t k = 0; k <= shard_count; ++k){ boundaries[k] = static_cast<uint32_t>(std::max<uint64_t>(plan.boundary(k), k ? boundaries[k - 1] : 0)); }
	return 0;
	RSQ_CATCH(1)
}

int rsq_group_unique_id(void *id_out, uint64_t capacity){
	RSQ_TRY
	if(!nccl_api().ok){ throw std::runtime_error("libnccl.so.2 could not be loaded: engines cannot be joined into a multi-GPU group"); }
	if(capacity < sizeof(ncclUniqueId)){ throw std::runtime_error("rsq_group_unique_id: buffer too small (" + std::to_string(sizeof(ncclUniqueId)) + " bytes needed)"); }
	ncclUniqueId id;
	RSQ_NCCL(nccl_api().GetUniqueId(&id));
	std::memcpy(id_out, &id, sizeof id);
	return 0;
	RSQ_CATCH(1)
}

int rsq_engine_join_group(rsq_engine *engine, const void *unique_id, int rank, int world){
	RSQ_TRY
	if(!nccl_api().ok){ throw std::runtime_error("libnccl.so.2 could not be loaded: engines cannot be joined into a multi-GPU group"); }
	if(world < 1 || rank < 0 || rank >= world){ throw std::runtime_error("rsq_engine_join_group: rank out of range"); }
	RSQ_CUDA(cudaSetDevice(engine->device));
	if(engine->comm){ nccl_api().CommDestroy(engine->comm); engine->comm = nullptr; }
	ncclUniqueId id;
	std::memcpy(&id, unique_id, sizeof id);
	RSQ_NCCL(nccl_api().CommInitRank(&engine->comm, world, id, rank));
	engine->group_rank = rank; engine->group_world = world;
	return 0;
	RSQ_CATCH(1)
}

int rsq_engine_leave_group(rsq_engine *engine){
	RSQ_TRY
	RSQ_CUDA(cudaSetDevice(engine->device));
	if(engine->comm){ nccl_api().CommDestroy(engine->comm); engine->comm = nullptr; }
	engine->group_rank = 0; engine->group_world = 1;
	return 0;
	RSQ_CATCH(1)
}

int rsq_engine_prepare(rsq_engine *engine, const rsq_reference *ref, const rsq_sim_options *opt, rsq_sim_report *report){
	RSQ_TRY
	RSQ_CUDA(cudaSetDevice(engine->device));
	if(report){ std::memset(report, 0, sizeof *report); }
	prepare(*engine, ref->g, *opt, report);
	return 0;
	RSQ_CATCH(1)
}

int rsq_engine_simulate(rsq_engine *engine, rsq_sim_report *report){
	RSQ_TRY
	RSQ_CUDA(cudaSetDevice(engine->device));
	simulate(*engine, report);
	return 0;
	RSQ_CATCH(1)
}

int rsq_engine_download(rsq_engine *engine, rsq_sim_report *report){
	RSQ_TRY
	RSQ_CUDA(cudaSetDevice(engine->device));
	download(*engine, report);
	return 0;
	RSQ_CATCH(1)
}

int rsq_engine_output(const rsq_engine *engine, int segment, const char **data, uint64_t *bytes){
	RSQ_TRY
	if(!engine->downloaded){ throw std::runtime_error("rsq_engine_download has not been called"); }
	if(segment < 0 || segment > 1){ throw std::runtime_error("segment must be 0 or 1"); }
	*data = engine->streamed_to_host ? engine->h_big[segment].data() : engine->h_out[segment].p; *bytes = engine->out_bytes[segment];
	return 0;
	RSQ_CATCH(1)
}

int rsq_engine_write(const rsq_engine *engine, const char *first_reads_path, const char *second_reads_path){
	RSQ_TRY
	if(!engine->downloaded){ throw std::runtime_error("rsq_engine_download has not been called"); }
	const char *paths[2] = {first_reads_path, second_reads_path};
	for(int seg = 0; seg < 2; ++seg){
		TextSink o;
		if(!o.open(paths[seg], true)){ throw std::runtime_error(std::string("Could not open '") + paths[seg] + "' for writing."); }
		if(!o.write(engine->streamed_to_host ? engine->h_big[seg].data() : engine->h_out[seg].p, engine->out_bytes[seg]) || !o.close()){
			throw std::runtime_error(std::string("Could not write records to '") + paths[seg] + "'");
		}
	}
	return 0;
	RSQ_CATCH(1)
}

int rsq_simulate(const rsq_profile *profile, const rsq_reference *ref, const rsq_sim_options *opt, int device,
                 const char *first_reads_path, const char *second_reads_path, rsq_sim_report *report){
	stage_log("rsq_simulate: enter");
	rsq_engine *e = rsq_engine_create(profile, device);
	if(!e){ return 1; }
	stage_log("rsq_simulate: engine created");
	TextSink sinks[2];
	try{
		const char *paths[2] = {first_reads_path, second_reads_path};
		for(int seg = 0; seg < 2; ++seg){ if(!sinks[seg].open(paths[seg])){ set_error("Could not open '%s' for writing.", paths[seg]); rsq_engine_destroy(e); return 1; } }
	}
	catch(const std::exception &ex){ set_error("%s", ex.what()); rsq_engine_destroy(e); return 1; }
	rsq_sim_report local; rsq_sim_report *rep = report ? report : &local;
	int rc = rsq_engine_prepare(e, ref, opt, rep);
	if(!rc){
		// the batches of the run are appended to the two files by a writer thread while the GPU works on the next one
		// (plain text, or gzip members compressed on the host cores when the file name ends in .gz)
		e->sink_files[0] = &sinks[0]; e->sink_files[1] = &sinks[1];
		rc = rsq_engine_simulate(e, rep);
		e->sink_files[0] = e->sink_files[1] = nullptr;
	}
	for(int seg = 0; seg < 2; ++seg){ if(!sinks[seg].close() && !rc){ set_error("Could not write records to '%s'", seg ? second_reads_path : first_reads_path); rc = 1; } }
	if(rc){ remove(first_reads_path); remove(second_reads_path); }
	rsq_engine_destroy(e);
	stage_log("rsq_simulate: engine destroyed");
	return rc;
}

// Appends the file `from` to the open descriptor of `to` (64 MiB at a time); false on an I/O error.
static bool append_file(FILE *to, const std::string &from){
	FILE *in = std::fopen(from.c_str(), "rb");
	if(!in){ return false; }
	std::vector<char> buf(64u << 20);
	bool ok = true;
	while(ok){
		const size_t n = std::fread(buf.data(), 1, buf.size(), in);
		if(!n){ ok = !std::ferror(in); break; }
		ok = std::fwrite(buf.data(), 1, n, to) == n;
	}
	std::fclose(in);
	return ok;
}

int rsq_simulate_multi(const rsq_profile *profile, const rsq_reference *ref, const rsq_sim_options *opt, int n_gpus, const int *devices,
                       const char *first_reads_path, const char *second_reads_path, rsq_sim_report *report){
	if(n_gpus <= 1){ return rsq_simulate(profile, ref, opt, devices ? devices[0] : 0, first_reads_path, second_reads_path, report); }
	RSQ_TRY
	stage_log("rsq_simulate_multi: enter");
	int have = 0;
	if(cudaGetDeviceCount(&have) != cudaSuccess || have < n_gpus){ throw std::runtime_error("rsq_simulate_multi: " + std::to_string(n_gpus) + " CUDA devices requested, " + std::to_string(have) + " visible"); }
	unsigned char id[128];
	if(rsq_group_unique_id(id, sizeof id)){ throw std::runtime_error(g_last_error); }
	// Shard 0 writes the two files themselves, the others hidden files next to them (same extension: gzip or plain is chosen by name) that are
	// appended in shard order once everything is through - the bytes of the 1-thread run of the reference.
	auto shard_path = [&](const char *path, int k) -> std::string {
		if(k == 0){ return path; }
		const std::string p = path;
		const size_t slash = p.find_last_of('/');
		const std::string dir = slash == std::string::npos ? "" : p.substr(0, slash + 1), base = slash == std::string::npos ? p : p.substr(slash + 1);
		return dir + ".rsq_shard" + std::to_string(k) + "_" + base;
	};
	std::vector<rsq_sim_report> reps(n_gpus);
	std::vector<std::string> errors(n_gpus);
	std::vector<std::thread> pool;
	for(int k = 0; k < n_gpus; ++k){
		pool.emplace_back([&, k]{
			const int dev = devices ? devices[k] : k;
			rsq_engine *e = rsq_engine_create(profile, dev);
			if(!e){ errors[k] = g_last_error; }
			// every engine has to enter the communicator, or the others wait forever: a failed creation ends the run through the error below
			if(e && rsq_engine_join_group(e, id, k, n_gpus)){ errors[k] = g_last_error; }
			TextSink sinks[2];
			const std::string p1 = shard_path(first_reads_path, k), p2 = shard_path(second_reads_path, k);
			if(errors[k].empty()){
				try{
					if(!sinks[0].open(p1)){ errors[k] = "Could not open '" + p1 + "' for writing."; }
					else if(!sinks[1].open(p2)){ errors[k] = "Could not open '" + p2 + "' for writing."; }
				}
				catch(const std::exception &ex){ errors[k] = ex.what(); }
			}
			if(e && errors[k].empty()){
				int rc = rsq_engine_prepare(e, ref, opt, &reps[k]);
				if(!rc){
					e->sink_files[0] = &sinks[0]; e->sink_files[1] = &sinks[1];
					rc = rsq_engine_simulate(e, &reps[k]);
					e->sink_files[0] = e->sink_files[1] = nullptr;
				}
				if(rc){ errors[k] = g_last_error; }
			}
			for(int seg = 0; seg < 2; ++seg){ if(sinks[seg].is_open() && !sinks[seg].close() && errors[k].empty()){ errors[k] = "Could not write records to '" + (seg ? p2 : p1) + "'"; } }
			if(e){ rsq_engine_destroy(e); }
		});
	}
	for(auto &th : pool){ th.join(); }
	std::string error;
	for(int k = 0; k < n_gpus && error.empty(); ++k){ error = errors[k]; }
	if(error.empty()){
		const char *paths[2] = {first_reads_path, second_reads_path};
		for(int seg = 0; seg < 2 && error.empty(); ++seg){
			FILE *to = std::fopen(paths[seg], "ab");
			if(!to){ error = std::string("Could not open '") + paths[seg] + "' for writing."; break; }
			for(int k = 1; k < n_gpus && error.empty(); ++k){ if(!append_file(to, shard_path(paths[seg], k))){ error = std::string("Could not write records to '") + paths[seg] + "'"; } }
			if(std::fclose(to) && error.empty()){ error = std::string("Could not write records to '") + paths[seg] + "'"; }
		}
	}
	for(int k = 1; k < n_gpus; ++k){ remove(shard_path(first_reads_path, k).c_str()); remove(shard_path(second_reads_path, k).c_str()); }
	if(!error.empty()){ remove(first_reads_path); remove(second_reads_path); throw std::runtime_error(error); }
	if(report){
		*report = reps[0];
		for(int k = 1; k < n_gpus; ++k){
			report->pairs += reps[k].pairs; report->bytes[0] += reps[k].bytes[0]; report->bytes[1] += reps[k].bytes[1]; report->blocks += reps[k].blocks;
			report->positions += reps[k].positions; report->scan_draws += reps[k].scan_draws; report->kernel_launches += reps[k].kernel_launches;
			report->ms_upload = std::max(report->ms_upload, reps[k].ms_upload); report->ms_bias = std::max(report->ms_bias, reps[k].ms_bias);
			report->ms_syserr = std::max(report->ms_syserr, reps[k].ms_syserr); report->ms_simulate = std::max(report->ms_simulate, reps[k].ms_simulate);
			report->ms_gather = std::max(report->ms_gather, reps[k].ms_gather); report->spec_rounds = std::max(report->spec_rounds, reps[k].spec_rounds);
		}
	}
	stage_log("rsq_simulate_multi: done");
	return 0;
	RSQ_CATCH(1)
}

int rsq_create_systematic_error_profile(rsq_engine *engine, const rsq_reference *ref, uint64_t seed, const char *fastq_out_path){
	RSQ_TRY
	RSQ_CUDA(cudaSetDevice(engine->device));
	try{ create_sys_profile(*engine, ref->g, seed, fastq_out_path); }
	catch(...){ remove(fastq_out_path); throw; }
	return 0;
	RSQ_CATCH(1)
}

int rsq_apply_error_model(rsq_engine *engine, const char *fasta_in_path, const char *fastq_out_path, uint64_t seed, rsq_sim_report *report){
	RSQ_TRY
	RSQ_CUDA(cudaSetDevice(engine->device));
	if(report){ std::memset(report, 0, sizeof *report); }
	try{ apply_error_model(*engine, fasta_in_path, fastq_out_path, seed, report); }
	catch(...){ remove(fastq_out_path); throw; }
	return 0;
	RSQ_CATCH(1)
}

int rsq_engine_fetch(const rsq_engine *engine, const char *name, void *dst, uint64_t capacity, uint64_t *bytes){
	RSQ_TRY
	RSQ_CUDA(cudaSetDevice(engine->device));
	const std::string n = name;
	const void *src = nullptr; uint64_t nb = 0; bool host = false;
	std::vector<uint64_t> seeds;
	if(n == "sys_fwd"){ src = engine->d_sys_fwd.p; nb = 2 * engine->total_size; }
	else if(n == "sys_rev"){ src = engine->d_sys_rev.p; nb = 2 * engine->total_size; }
	else if(n == "adapter_sys"){ src = engine->d_adapter_sys.p; nb = 2 * engine->h_adapter_seq.size(); }
	else if(n == "sur_start"){ src = engine->d_sur_start.p; nb = 8 * engine->total_size; }
	else if(n == "sur_end"){ src = engine->d_sur_end.p; nb = 8 * engine->total_size; }
	else if(n == "reference"){ src = engine->d_ref.p; nb = engine->total_size; }
	else if(n == "thresholds"){ src = engine->norm.thresholds.data(); nb = 8 * engine->norm.thresholds.size(); host = true; }
	else if(n == "blocks"){ src = engine->d_blocks.p; nb = sizeof(BlockDesc) * engine->n_blocks_total; }
	else if(n == "spec_blocks"){ src = engine->d_spec_blocks.p; nb = sizeof(SpecBlock) * engine->spec_units_last; }   // per-unit counters of the last speculative batch (rounds, reads, scan draws): tuning
	else{ throw std::runtime_error("unknown stage array '" + n + "'"); }
	if((n == "sur_start" || n == "sur_end") && engine->sur_windowed){ throw std::runtime_error("stage array '" + n + "' is no longer resident: a multi-batch simulate call keeps one batch's window only (RSQ_KEEP_STAGES=1 keeps the whole arrays)"); }
	if(bytes){ *bytes = nb; }
	const uint64_t cnt = std::min(nb, capacity);
	if(dst && cnt){
		if(host){ std::memcpy(dst, src, cnt); }
		else{ RSQ_CUDA(cudaMemcpy(dst, src, cnt, cudaMemcpyDeviceToHost)); }
	}
	return 0;
	RSQ_CATCH(1)
}

}  // extern "C"
