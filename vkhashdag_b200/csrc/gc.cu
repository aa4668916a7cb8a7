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
//   rebuild (bottom-up) per level: gather the marked nodes with their children remapped through the level-below map,
//                       zero the used part of the level's buckets, find-or-insert everything again with the edit
//                       path's batched upsert (hash -> bucket -> lock-free append), record old -> new in a hash map.
#include "common.cuh"

#include <algorithm>

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

// zero the used prefix of every bucket of one level and reset its cursor (one CTA per bucket, grid-stride)
__global__ void __launch_bounds__(kGcBlock) k_gc_clear_level(uint32_t *words, uint32_t *bucket_words, uint32_t first_bucket,
                                                             uint32_t n_buckets, uint32_t bucket_shift) {
	for (uint32_t b = blockIdx.x; b < n_buckets; b += gridDim.x) {
		const uint32_t bucket = first_bucket + b, used = bucket_words[bucket];
		uint32_t *base = words + (size_t(bucket) << bucket_shift);
		for (uint32_t i = threadIdx.x; i < used; i += blockDim.x)
			base[i] = 0u;
		__syncthreads();
		if (threadIdx.x == 0)
			bucket_words[bucket] = 0u;
	}
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
	std::vector<GcLevel> lv(L);
	uint32_t *count_dev = nullptr, *seeds_dev = nullptr;
	HD_CUDA_TRY(gmalloc(&count_dev, 1, s));
	HD_CUDA_TRY(gmalloc(&seeds_dev, n_roots + 1, s));
	uint64_t total = 0;
	auto cleanup = [&]() {
		for (auto &l : lv) {
			if (l.list)
				cudaFreeAsync(l.list, s);
			if (l.keys)
				cudaFreeAsync(l.keys, s);
			if (l.vals)
				cudaFreeAsync(l.vals, s);
		}
		cudaFreeAsync(count_dev, s), cudaFreeAsync(seeds_dev, s);
		cudaStreamSynchronize(s);
	};

	// ---- mark (forward pass, NodePoolThreadedGC.hpp:72-103) ----
	for (uint32_t l = 0; l < L; ++l) {
		// capacity: children of the previous level + the filled node (+ the roots at level 0)
		const uint64_t cap = l == 0 ? uint64_t(n_roots) + 1 : uint64_t(lv[l - 1].n) * 8 + 1;
		if (cap > 0xFFFFFFF0ull) {
			set_error("gc: level %u too large", l);
			cleanup();
			return HD_ERR_OVERFLOW;
		}
		const uint64_t ts = table_size(cap);
		lv[l].mask = uint32_t(ts - 1);
		HD_CUDA_TRY(gmalloc(&lv[l].list, cap, s));
		HD_CUDA_TRY(gmalloc(&lv[l].keys, ts, s));
		HD_CUDA_TRY(cudaMemsetAsync(lv[l].keys, 0xFF, ts * 4, s));
		HD_CUDA_TRY(cudaMemsetAsync(count_dev, 0, 4, s));
		std::vector<uint32_t> seeds;
		if (l == 0)
			seeds.assign(roots, roots + n_roots);
		seeds.push_back(p->filled[l]);
		HD_CUDA_TRY(cudaMemcpyAsync(seeds_dev, seeds.data(), seeds.size() * 4, cudaMemcpyHostToDevice, s));
		HD_CUDA_TRY(cudaStreamSynchronize(s)); // seeds is a stack vector
		k_gc_seed<<<1, 32, 0, s>>>(lv[l].keys, lv[l].mask, seeds_dev, uint32_t(seeds.size()), lv[l].list, count_dev);
		HD_LAUNCH_CHECK();
		if (l > 0 && lv[l - 1].n) {
			k_gc_expand<<<gblocks(uint64_t(lv[l - 1].n) * 8), kGcBlock, 0, s>>>(p->words, lv[l - 1].list, lv[l - 1].n, lv[l].keys,
			                                                                 lv[l].mask, lv[l].list, count_dev);
			HD_LAUNCH_CHECK();
		}
		HD_CUDA_TRY(cudaMemcpyAsync(&lv[l].n, count_dev, 4, cudaMemcpyDeviceToHost, s));
		HD_CUDA_TRY(cudaStreamSynchronize(s));
		total += lv[l].n;
	}

	// ---- rebuild (backward pass, NodePoolThreadedGC.hpp:276-348) ----
	for (uint32_t l = L; l-- > 0;) {
		const bool is_leaf = l == L - 1;
		const uint32_t n = lv[l].n, stride = is_leaf ? 2u : 9u;
		uint32_t *cand = nullptr, *result = nullptr;
		HD_CUDA_TRY(gmalloc(&cand, uint64_t(n) * stride, s));
		HD_CUDA_TRY(gmalloc(&result, n, s));
		if (n) {
			k_gc_gather<<<gblocks(n), kGcBlock, 0, s>>>(p->words, lv[l].list, n, is_leaf, is_leaf ? nullptr : lv[l + 1].keys,
			                                          is_leaf ? nullptr : lv[l + 1].vals, is_leaf ? 0u : lv[l + 1].mask, cand);
			HD_LAUNCH_CHECK();
		}
		const uint32_t nb = 1u << g.bucket_bits[l];
		k_gc_clear_level<<<std::min<uint32_t>(nb, 148u * 16u), kGcBlock, 0, s>>>(p->words, p->bucket_words, g.level_base[l], nb,
		                                                                      g.bucket_shift());
		HD_LAUNCH_CHECK();
		st = upsert_batch_dev(p, l, n, stride, cand, result);
		if (st != HD_OK) {
			cudaFreeAsync(cand, s), cudaFreeAsync(result, s);
			cleanup();
			return st;
		}
		// the mark set of this level becomes its old -> new map
		HD_CUDA_TRY(gmalloc(&lv[l].vals, uint64_t(lv[l].mask) + 1, s));
		if (n) {
			k_gc_map_fill<<<gblocks(n), kGcBlock, 0, s>>>(lv[l].list, result, n, lv[l].keys, lv[l].vals, lv[l].mask);
			HD_LAUNCH_CHECK();
		}
		cudaFreeAsync(cand, s), cudaFreeAsync(result, s);
	}

	// ---- remap roots and filled nodes (tiny lookup kernels) ----
	std::vector<uint32_t> filled(L);
	{
		uint32_t *io = nullptr;
		const uint32_t n_io = std::max(n_roots, 1u);
		HD_CUDA_TRY(gmalloc(&io, n_io, s));
		for (uint32_t l = 0; l < L; ++l) {
			HD_CUDA_TRY(cudaMemcpyAsync(io, &p->filled[l], 4, cudaMemcpyHostToDevice, s));
			k_gc_lookup<<<1, 32, 0, s>>>(lv[l].keys, lv[l].vals, lv[l].mask, io, 1);
			HD_LAUNCH_CHECK();
			HD_CUDA_TRY(cudaMemcpyAsync(&filled[l], io, 4, cudaMemcpyDeviceToHost, s));
			HD_CUDA_TRY(cudaStreamSynchronize(s));
		}
		if (n_roots) {
			HD_CUDA_TRY(cudaMemcpyAsync(io, roots, size_t(n_roots) * 4, cudaMemcpyHostToDevice, s));
			k_gc_lookup<<<gblocks(n_roots), kGcBlock, 0, s>>>(lv[0].keys, lv[0].vals, lv[0].mask, io, n_roots);
			HD_LAUNCH_CHECK();
			HD_CUDA_TRY(cudaMemcpyAsync(new_roots, io, size_t(n_roots) * 4, cudaMemcpyDeviceToHost, s));
			HD_CUDA_TRY(cudaStreamSynchronize(s));
		}
		cudaFreeAsync(io, s);
	}
	uint32_t root_new = p->root;
	if (p->root != kNull) {
		// the pool's published root follows if it is one of the GC roots, otherwise it is dropped
		root_new = kNull;
		for (uint32_t i = 0; i < n_roots; ++i)
			if (roots[i] == p->root)
				root_new = new_roots[i];
	}
	cleanup();
	if (st != HD_OK)
		return st;
	st = set_filled(p, filled);
	if (st != HD_OK)
		return st;
	p->root = root_new;
	// every pointer changed: replicas need the whole pool again, after clearing theirs
	HD_CUDA_TRY(cudaMemsetAsync(p->bucket_synced, 0, size_t(g.total_buckets) * 4, s));
	HD_CUDA_TRY(cudaStreamSynchronize(s));
	p->needs_full_resync = true;
	if (reachable_nodes)
		*reachable_nodes = total;
	return HD_OK;
}
