// gc.cu — GPU garbage collection / compaction of the node pool.
//
// Replaces NodePoolThreadedGC::ThreadedGC (include/hashdag/NodePoolThreadedGC.hpp:72-103 forward mark, :276-348
// backward compact/re-hash, :372-403 entry).  Same contract: afterwards the pool holds exactly the nodes reachable
// from the given roots plus the filled nodes (NodePool.hpp:54 "should be preserved when GC"), pointers change, the
// remapped roots are returned.  The reference's page-free precedence quirk is not reproduced: parity
// is canonical (same DAG, no unreachable node left).
//
//   mark (top-down)     level lists of unique reachable pointers; children are deduplicated through a scratch
//                       open-addressing set (atomicCAS on the pointer value).
//   rebuild (bottom-up) into a FRESH word space allocated beside the old one: per level gather the marked nodes with their
//                       children remapped through the level-below map, find-or-insert them with the edit path's batched
//                       upsert (hash -> bucket -> lock-free append), record old -> new in a hash map.  The new arrays
//                       replace the old ones only when every node found a place (a full bucket -> HD_ERR_OVERFLOW with
//                       the pool untouched); hd_pool_words_dev / hd_pool_bucket_words_dev change across a successful GC.
#include "common.cuh"

#include <algorithm>
#include <cstdlib>

namespace hd {

constexpr int kGcBlock = 256;
constexpr uint32_t kEmpty = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t gc_hash(uint32_t v) {
	v ^= v >> 16;
	v *= 0x7feb352du;
	v ^= v >> 15;
	v *= 0x846ca68bu;
	v ^= v >> 16;
	return v;
}

// set insert; returns true when `key` was not present
__device__ __forceinline__ bool set_insert(uint32_t *keys, uint32_t mask, uint32_t key) {
	uint32_t slot = gc_hash(key) & mask;
	for (;;) {
		const uint32_t cur = keys[slot];
		if (cur == key)
			return false;
		if (cur == kEmpty) {
			const uint32_t prev = atomicCAS(&keys[slot], kEmpty, key);
			if (prev == kEmpty)
				return true;
			if (prev == key)
				return false;
		}
		slot = (slot + 1u) & mask;
	}
}
__device__ __forceinline__ uint32_t map_find_slot(const uint32_t *keys, uint32_t mask, uint32_t key) {
	uint32_t slot = gc_hash(key) & mask;
	while (keys[slot] != key)
		slot = (slot + 1u) & mask;
	return slot;
}

// mark: children of the level-l list that were not seen yet are appended to the level-(l+1) list
__global__ void __launch_bounds__(kGcBlock) k_gc_expand(const uint32_t *__restrict__ words, const uint32_t *__restrict__ list,
                                                        uint32_t n, uint32_t *keys, uint32_t mask, uint32_t *next,
                                                        uint32_t *next_count) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t i = t >> 3, c = t & 7u;
	bool fresh = false;
	uint32_t child = kEmpty;
	if (i < n) {
		const uint32_t ptr = list[i];
		const uint32_t m = words[ptr] & 0xFFu;
		if (m >> c & 1u) {
			child = words[ptr + 1u + __popc(m & ((1u << c) - 1u))];
			fresh = set_insert(keys, mask, child);
		}
	}
	const uint32_t vote = __ballot_sync(0xFFFFFFFFu, fresh);
	if (!vote)
		return;
	const uint32_t lane = threadIdx.x & 31u;
	uint32_t base = 0;
	if (lane == 0)
		base = atomicAdd(next_count, __popc(vote));
	base = __shfl_sync(0xFFFFFFFFu, base, 0);
	if (fresh)
		next[base + __popc(vote & ((1u << lane) - 1u))] = child;
}

__global__ void k_gc_seed(uint32_t *keys, uint32_t mask, const uint32_t *__restrict__ seeds, uint32_t n_seeds, uint32_t *list,
                          uint32_t *count) {
	if (blockIdx.x || threadIdx.x)
		return;
	for (uint32_t i = 0; i < n_seeds; ++i)
		if (seeds[i] != kNull && set_insert(keys, mask, seeds[i]))
			list[(*count)++] = seeds[i];
}

// rebuild: copy marked nodes out of the pool, children remapped through the map of the level below
__global__ void __launch_bounds__(kGcBlock) k_gc_gather(const uint32_t *__restrict__ words, const uint32_t *__restrict__ list,
                                                        uint32_t n, bool is_leaf, const uint32_t *__restrict__ child_keys,
                                                        const uint32_t *__restrict__ child_vals, uint32_t child_mask,
                                                        uint32_t *cand) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const uint32_t ptr = list[i];
	if (is_leaf) {
		*reinterpret_cast<uint2 *>(cand + size_t(i) * 2u) = *reinterpret_cast<const uint2 *>(words + ptr);
		return;
	}
	const uint32_t m = words[ptr] & 0xFFu, k = __popc(m);
	uint32_t *dst = cand + size_t(i) * 9u;
	dst[0] = m;
	for (uint32_t j = 0; j < k; ++j)
		dst[1 + j] = child_vals[map_find_slot(child_keys, child_mask, words[ptr + 1u + j])];
}

__global__ void __launch_bounds__(kGcBlock) k_gc_map_fill(const uint32_t *__restrict__ list, const uint32_t *__restrict__ result,
                                                          uint32_t n, const uint32_t *__restrict__ keys, uint32_t *vals,
                                                          uint32_t mask) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
		vals[map_find_slot(keys, mask, list[i])] = result[i];
}

__global__ void k_gc_lookup(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint32_t mask, uint32_t *io,
                            uint32_t n) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && io[i] != kNull)
		io[i] = vals[map_find_slot(keys, mask, io[i])];
}

template <typename T> static cudaError_t gmalloc(T **p, uint64_t count, cudaStream_t s) {
	return cudaMallocAsync(reinterpret_cast<void **>(p), std::max<uint64_t>(count, 1) * sizeof(T), s);
}
static inline uint32_t gblocks(uint64_t threads) { return uint32_t((threads + kGcBlock - 1) / kGcBlock); }
static inline uint64_t table_size(uint64_t n) {
	uint64_t t = 64;
	while (t < n * 2)
		t <<= 1;
	return t;
}

struct GcLevel {
	uint32_t *list = nullptr; // unique reachable pointers of the level
	uint32_t n = 0;
	uint32_t *keys = nullptr, *vals = nullptr; // set during mark, map old -> new during rebuild
	uint32_t mask = 0;
};

// number of kNull results of a batched upsert (a full bucket: the rebuilt node could not be stored)
__global__ void __launch_bounds__(kGcBlock) k_gc_count_null(const uint32_t *__restrict__ result, uint32_t n, uint32_t *count) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	const bool bad = i < n && result[i] == kNull;
	const uint32_t vote = __ballot_sync(0xFFFFFFFFu, bad);
	if (vote && (threadIdx.x & 31u) == 0)
		atomicAdd(count, __popc(vote));
}

namespace {
// Everything hd_gc allocates; released on every exit path (the early returns of HD_CUDA_TRY included).
struct GcState {
	hd_pool *p;
	cudaStream_t s;
	std::vector<GcLevel> lv;
	uint32_t *count_dev = nullptr, *seeds_dev = nullptr, *cand = nullptr, *result = nullptr, *io = nullptr;
	uint32_t *new_words = nullptr, *new_bw = nullptr; // the compacted pool is built here; swapped in only on success
	uint32_t *old_words = nullptr, *old_bw = nullptr;
	bool swapped_in = false, committed = false;
	~GcState() {
		for (auto &l : lv) {
			if (l.list)
				cudaFreeAsync(l.list, s);
			if (l.keys)
				cudaFreeAsync(l.keys, s);
			if (l.vals)
				cudaFreeAsync(l.vals, s);
		}
		uint32_t *tmp[] = {count_dev, seeds_dev, cand, result, io};
		for (uint32_t *q : tmp)
			if (q)
				cudaFreeAsync(q, s);
		cudaStreamSynchronize(s);
		if (swapped_in && !committed) // failure after the swap: the old pool is untouched, put it back
			p->words = old_words, p->bucket_words = old_bw;
		if (committed)
			cudaFree(old_words), cudaFree(old_bw);
		else
			cudaFree(new_words), cudaFree(new_bw);
	}
};
} // namespace

} // namespace hd

using namespace hd;

extern "C" hd_status hd_gc(hd_pool *p, const uint32_t *roots, uint32_t n_roots, uint32_t *new_roots,
                           uint64_t *reachable_nodes) {
	if (!p || (!roots && n_roots) || (!new_roots && n_roots))
		return HD_ERR_INVALID;
	HD_CUDA_TRY(cudaSetDevice(p->device));
	hd_status st = ensure_filled(p);
	if (st != HD_OK)
		return st;
	const Geometry &g = p->geo;
	const uint32_t L = g.node_levels;
	cudaStream_t s = p->stream;
	GcState G{p, s};
	std::vector<GcLevel> &lv = G.lv;
	lv.resize(L);
	HD_CUDA_TRY(gmalloc(&G.count_dev, 1, s));
	HD_CUDA_TRY(gmalloc(&G.seeds_dev, n_roots + 1, s));
	uint64_t total = 0;

	// ---- mark (forward pass, NodePoolThreadedGC.hpp:72-103) ----
	for (uint32_t l = 0; l < L; ++l) {
		// capacity: children of the previous level + the filled node (+ the roots at level 0)
		const uint64_t cap = l == 0 ? uint64_t(n_roots) + 1 : uint64_t(lv[l - 1].n) * 8 + 1;
		if (cap > 0xFFFFFFF0ull) {
			set_error("gc: level %u too large", l);
			return HD_ERR_OVERFLOW;
		}
		const uint64_t ts = table_size(cap);
		lv[l].mask = uint32_t(ts - 1);
		HD_CUDA_TRY(gmalloc(&lv[l].list, cap, s));
		HD_CUDA_TRY(gmalloc(&lv[l].keys, ts, s));
		HD_CUDA_TRY(cudaMemsetAsync(lv[l].keys, 0xFF, ts * 4, s));
		HD_CUDA_TRY(cudaMemsetAsync(G.count_dev, 0, 4, s));
		std::vector<uint32_t> seeds;
		if (l == 0)
			seeds.assign(roots, roots + n_roots);
		seeds.push_back(p->filled[l]);
		HD_CUDA_TRY(cudaMemcpyAsync(G.seeds_dev, seeds.data(), seeds.size() * 4, cudaMemcpyHostToDevice, s));
		HD_CUDA_TRY(cudaStreamSynchronize(s)); // seeds is a stack vector
		k_gc_seed<<<1, 32, 0, s>>>(lv[l].keys, lv[l].mask, G.seeds_dev, uint32_t(seeds.size()), lv[l].list, G.count_dev);
		HD_LAUNCH_CHECK();
		if (l > 0 && lv[l - 1].n) {
			k_gc_expand<<<gblocks(uint64_t(lv[l - 1].n) * 8), kGcBlock, 0, s>>>(p->words, lv[l - 1].list, lv[l - 1].n, lv[l].keys,
			                                                                 lv[l].mask, lv[l].list, G.count_dev);
			HD_LAUNCH_CHECK();
		}
		HD_CUDA_TRY(cudaMemcpyAsync(&lv[l].n, G.count_dev, 4, cudaMemcpyDeviceToHost, s));
		HD_CUDA_TRY(cudaStreamSynchronize(s));
		total += lv[l].n;
	}

	// ---- rebuild (backward pass, NodePoolThreadedGC.hpp:276-348) into a FRESH word space ----
	// Remapped child pointers change every inner node's hash, so bucket loads are re-randomised and a bucket of a nearly
	// full pool can overflow during the rebuild.  The compacted pool is therefore built beside the old one and swapped
	// in only when every node found a place; on overflow the call returns HD_ERR_OVERFLOW and the pool is unchanged.
	HD_CUDA_TRY(cudaMalloc(&G.new_words, g.total_words * sizeof(uint32_t)));
	HD_CUDA_TRY(cudaMalloc(&G.new_bw, size_t(g.total_buckets) * sizeof(uint32_t)));
	HD_CUDA_TRY(cudaMemsetAsync(G.new_words, 0, g.total_words * sizeof(uint32_t), s));
	HD_CUDA_TRY(cudaMemsetAsync(G.new_bw, 0, size_t(g.total_buckets) * sizeof(uint32_t), s));
	G.old_words = p->words, G.old_bw = p->bucket_words;
	p->tt_invalidate(); // every pointer changes
	p->words = G.new_words, p->bucket_words = G.new_bw; // upsert_batch_dev appends into the pool's current arrays
	G.swapped_in = true;
	for (uint32_t l = L; l-- > 0;) {
		const bool is_leaf = l == L - 1;
		const uint32_t n = lv[l].n, stride = is_leaf ? 2u : 9u;
		HD_CUDA_TRY(gmalloc(&G.cand, uint64_t(n) * stride, s));
		HD_CUDA_TRY(gmalloc(&G.result, n, s));
		if (n) {
			k_gc_gather<<<gblocks(n), kGcBlock, 0, s>>>(G.old_words, lv[l].list, n, is_leaf, is_leaf ? nullptr : lv[l + 1].keys,
			                                          is_leaf ? nullptr : lv[l + 1].vals, is_leaf ? 0u : lv[l + 1].mask, G.cand);
			HD_LAUNCH_CHECK();
		}
		st = upsert_batch_dev(p, l, n, stride, G.cand, G.result);
		if (st != HD_OK)
			return st;
		if (n) { // a full bucket leaves kNull in result[]: publishing it would put Null children under set mask bits
			uint32_t lost = 0;
			HD_CUDA_TRY(cudaMemsetAsync(G.count_dev, 0, 4, s));
			k_gc_count_null<<<gblocks(n), kGcBlock, 0, s>>>(G.result, n, G.count_dev);
			HD_LAUNCH_CHECK();
			HD_CUDA_TRY(cudaMemcpyAsync(&lost, G.count_dev, 4, cudaMemcpyDeviceToHost, s));
			HD_CUDA_TRY(cudaStreamSynchronize(s));
			// test hook: HD_GC_INJECT_OVERFLOW_LEVEL=l pretends level l lost a node (tests/test_gpu_gc_io.py checks that the
			// pool survives a failed compaction untouched)
			if (const char *inj = getenv("HD_GC_INJECT_OVERFLOW_LEVEL"))
				if (uint32_t(atoi(inj)) == l)
					lost = std::max(lost, 1u);
			if (lost) {
				set_error("gc: %u node(s) of level %u found their bucket full during the rebuild; the pool is unchanged", lost, l);
				return HD_ERR_OVERFLOW;
			}
		}
		// the mark set of this level becomes its old -> new map
		HD_CUDA_TRY(gmalloc(&lv[l].vals, uint64_t(lv[l].mask) + 1, s));
		if (n) {
			k_gc_map_fill<<<gblocks(n), kGcBlock, 0, s>>>(lv[l].list, G.result, n, lv[l].keys, lv[l].vals, lv[l].mask);
			HD_LAUNCH_CHECK();
		}
		cudaFreeAsync(G.cand, s), cudaFreeAsync(G.result, s);
		G.cand = G.result = nullptr;
	}

	// ---- remap roots and filled nodes (tiny lookup kernels) ----
	std::vector<uint32_t> filled(L);
	std::vector<uint32_t> mapped(std::max(n_roots, 1u), kNull);
	{
		const uint32_t n_io = std::max(n_roots, 1u);
		HD_CUDA_TRY(gmalloc(&G.io, n_io, s));
		for (uint32_t l = 0; l < L; ++l) {
			HD_CUDA_TRY(cudaMemcpyAsync(G.io, &p->filled[l], 4, cudaMemcpyHostToDevice, s));
			k_gc_lookup<<<1, 32, 0, s>>>(lv[l].keys, lv[l].vals, lv[l].mask, G.io, 1);
			HD_LAUNCH_CHECK();
			HD_CUDA_TRY(cudaMemcpyAsync(&filled[l], G.io, 4, cudaMemcpyDeviceToHost, s));
			HD_CUDA_TRY(cudaStreamSynchronize(s));
		}
		if (n_roots) {
			HD_CUDA_TRY(cudaMemcpyAsync(G.io, roots, size_t(n_roots) * 4, cudaMemcpyHostToDevice, s));
			k_gc_lookup<<<gblocks(n_roots), kGcBlock, 0, s>>>(lv[0].keys, lv[0].vals, lv[0].mask, G.io, n_roots);
			HD_LAUNCH_CHECK();
			HD_CUDA_TRY(cudaMemcpyAsync(mapped.data(), G.io, size_t(n_roots) * 4, cudaMemcpyDeviceToHost, s));
			HD_CUDA_TRY(cudaStreamSynchronize(s));
		}
	}
	uint32_t root_new = p->root;
	if (p->root != kNull) {
		// the pool's published root follows if it is one of the GC roots, otherwise it is dropped
		root_new = kNull;
		for (uint32_t i = 0; i < n_roots; ++i)
			if (roots[i] == p->root)
				root_new = mapped[i];
	}
	// every pointer changed: replicas need the whole pool again, after clearing theirs
	HD_CUDA_TRY(cudaMemsetAsync(p->bucket_synced, 0, size_t(g.total_buckets) * 4, s));
	HD_CUDA_TRY(cudaStreamSynchronize(s));
	// ---- commit: nothing below can fail with the new pool half-published ----
	G.committed = true;
	for (uint32_t i = 0; i < n_roots; ++i)
		new_roots[i] = mapped[i];
	p->root = root_new;
	p->needs_full_resync = true;
	edit_scratch_free(p); // the low-latency edit path caches the pool's array addresses (arena locks, captured graph)
	st = set_filled(p, filled);
	if (st != HD_OK)
		return st;
	if (reachable_nodes)
		*reachable_nodes = total;
	return HD_OK;
}
