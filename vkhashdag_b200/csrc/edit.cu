// edit.cu — kernel family #2: batched, level-synchronous edit rebuild on the GPU (sm_100a).
//
// Replaces the libfork CPU editor: NodePoolBase::Edit / edit_node / edit_leaf / upsert_node
// (include/hashdag/NodePool.hpp:159-212,319-417) under NodePoolThreadedEdit (NodePoolThreadedEdit.hpp:27-126).
//
// The reference recurses depth-first, one edit at a time.  Here a whole batch of edits is applied in ONE
// breadth-first pass:
//   top-down  (k_down)      every work item (node that some edit must enter) classifies its 8 children against
//                           its ordered edit list (EditNode, NodePool.hpp:345-360).  A kFill/kClear replaces the
//                           child by the filled node / Null and discards the earlier edits of the list (exactly
//                           what sequential application would leave); kNotAffected edits are dropped; the kProceed
//                           edits that remain form the child's list.  Children with an empty list are final.
//   leaves    (k_leaf)      one warp per 4x4x4 leaf: 64 EditVoxel evaluations per listed edit (NodePool.hpp:319-343).
//   bottom-up (k_assemble)  re-pack changed nodes (NodePool.hpp:362-396), then find-or-insert them:
//       k_dedup   in-batch dedup through a scratch open-addressing table (one winner per distinct content),
//       k_upsert  one warp per winner: coalesced scan of the bucket (find_node, NodePool.hpp:79-132) and, on a
//                 miss, a lock-free CAS reservation in the bucket honouring the no-page-straddle rule
//                 (append_node, NodePool.hpp:134-157),
//       k_resolve losers take their winner's pointer; results flow to the parent's child slots.
// Phase separation makes the append log safe without locks: a bucket scan can only miss nodes whose content
// differs from the scanner's (equal contents were merged by k_dedup).  A scan may still observe a node another warp
// is in the middle of writing: for inner nodes a half-written node can never equal a candidate (unwritten child
// slots are 0, no child pointer is 0); a leaf word CAN be 0, so leaves are published with a single 64-bit store
// (found the hard way: (z0, 0) half-written aliased the valid leaf (z0, 0) in ~10 % of cold runs).
// The final voxel set equals sequential application of the
// batch, hence the canonical DAG is identical to the reference's; pointer values are not.
#include "common.cuh"
#include "editors.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace hd {
namespace cg = cooperative_groups;

constexpr uint32_t kPending = 0xFFFFFFFEu; // placeholder in child_new until the child item reports
constexpr int kBlock = 256;

struct DevCounters {
	unsigned long long stats[8]; // visited_nodes, visited_leaves, upserts, appended_nodes, appended_words, overflow, -, scan_words
	uint32_t next_items;
	uint32_t next_entries;
	uint32_t has_long; // some list of the level just expanded is longer than 32 entries -> k_down_long must run
	uint32_t has_huge; // ... longer than kHugeList entries -> k_down_huge (a CTA per list) must run
	uint32_t root_out;
	uint32_t error; // 1 = scratch overflow
	// low-latency path (one CUDA graph, no host round trips): item / list-entry counts of every level stay on the device
	uint32_t lvl_items[HD_MAX_NODE_LEVELS];
	uint32_t lvl_entries[HD_MAX_NODE_LEVELS];
	uint32_t lvl_long[HD_MAX_NODE_LEVELS]; // fused kernel: some item of the level carries a list longer than 32 entries
	// phase boundaries of the fused kernel (%globaltimer, ns; CTA 0 / thread 0): [0] start, then one stamp per finished
	// phase.  Printed by HD_EDIT_FAST_TRACE=1 (a profiling aid: ncu sees the cooperative kernel as one launch).
	uint32_t n_stamps;
	uint32_t pad_;
	unsigned long long stamp_ns[2 * HD_MAX_NODE_LEVELS + 8];
};
__device__ __forceinline__ void phase_stamp(DevCounters *ctr) {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	const uint32_t i = ctr->n_stamps;
	if (i < 2 * HD_MAX_NODE_LEVELS + 8)
		ctr->stamp_ns[i] = t, ctr->n_stamps = i + 1;
}

// One BFS level of work items (device arrays, SoA).
struct LevelView {
	uint32_t n;          // items (host-known); ignored when n_dev is set
	const uint32_t *n_dev; // device-resident item count (low-latency path) or NULL
	const uint32_t *err_dev; // with n_dev: DevCounters::error — once a queue overflowed, every later kernel sees 0 items
	__device__ __forceinline__ uint32_t count() const {
		if (!n_dev)
			return n;
		return *(volatile const uint32_t *)err_dev ? 0u : min(*(volatile const uint32_t *)n_dev, cap);
	}
	uint32_t cap;        // item capacity
	uint32_t cap_entries;
	uint32_t *cur;       // current node pointer of the item (after Fill/Clear substitution)
	uint64_t *pos;       // packed node coordinates x | y<<21 | z<<42
	uint32_t *list_off;  // first entry in `lists`
	uint32_t *list_len;
	uint32_t *parent;    // (parent item << 3) | child slot; 0xFFFFFFFF for the root item
	uint32_t *result;    // new pointer of the item
	uint32_t *child_new; // [n*8] new child pointers (inner levels)
	uint32_t *lists;     // edit indices, ascending (= application order)
	// upsert candidates
	uint32_t *cand;      // [n*stride] packed node words
	uint8_t *state;      // 0 = final, 1 = candidate (winner or loser), 2 = winner
	uint32_t *winner;    // winner item of a loser
};

// Low-latency path for small batches (the interactive brush of src/main.cpp:214-238 is ONE editor per call): every
// level's work queue lives in a fixed arena, the item counts stay on the device, and the whole rebuild — root
// classification, every k_down, the leaf pass, every level's dedup / find-or-insert / resolve — is ONE instantiated
// CUDA graph.  A call costs one graph launch and one 16-byte-aligned counter read-back instead of ~2 host round trips
// per level.  A queue that turns out too small sets DevCounters::error before anything touches the pool (k_upsert
// checks it), and the call falls back to the exactly-sized general path.
constexpr uint32_t kFastMaxEdits = 1024; // batches up to this many editors take the one-launch path; lists longer than 32
                                         // entries (only near the root) are filtered serially by the child's thread there
struct FastDyn {
	uint32_t root, n_edits, pad[2];
	hd_edit_desc edits[kFastMaxEdits];
};
struct FastPath {
	bool tried = false, ok = false;
	std::vector<LevelView> lv;
	char *arena = nullptr;
	uint32_t *locks = nullptr; // one word per bucket (the reference stripes 1024 mutexes, src/DAGNodePool.hpp:43,56)
	FastDyn *dyn_host = nullptr, *dyn_dev = nullptr;
	DevCounters *ctr_host = nullptr;
	uint32_t *iota_dev = nullptr;
	cudaGraph_t graph = nullptr;
	cudaGraphExec_t exec = nullptr;
	uint32_t kernels = 0;
	// mode 1 (default): the fused cooperative kernel; mode 2 (HD_EDIT_FAST=2): the same phases as a CUDA graph
	int mode = 1;
	FastDyn *dyn_host_dev = nullptr; // device-side addresses of the mapped host blocks
	DevCounters *ctr_host_dev = nullptr;
	uint32_t fused_grid = 0;
	uint32_t pending_edits = 0; // editors of the call in flight between fast_edit_begin and fast_edit_end
};

struct EditScratch {
	DevCounters *ctr = nullptr;
	uint32_t *filled_dev = nullptr;
	bool fast_scan = true;
	FastPath fast;
	uint32_t last_path = 0; // hd_edit_last_path
};

__device__ __forceinline__ uint64_t pack_pos(uint32_t x, uint32_t y, uint32_t z) {
	return uint64_t(x) | (uint64_t(y) << 21) | (uint64_t(z) << 42);
}
__device__ __forceinline__ void unpack_pos(uint64_t p, uint32_t &x, uint32_t &y, uint32_t &z) {
	x = uint32_t(p) & 0x1FFFFFu, y = uint32_t(p >> 21) & 0x1FFFFFu, z = uint32_t(p >> 42) & 0x1FFFFFu;
}

// ---- hashing: include/hashdag/Hasher.hpp:21-49 -------------------------------------------------------
__device__ __forceinline__ uint32_t rotl32(uint32_t v, int s) { return __funnelshift_l(v, v, s); }
__device__ __forceinline__ uint32_t hash_inner(const uint32_t *w, uint32_t n) {
	uint32_t h = 0;
	for (uint32_t i = 0; i < n; ++i) {
		uint32_t k = w[i] * 0xcc9e2d51u;
		k = rotl32(k, 15) * 0x1b873593u;
		h = rotl32(h ^ k, 13) * 5u + 0xe6546b64u;
	}
	h ^= n;
	return fmix32(h);
}
__device__ __forceinline__ uint32_t hash_leaf(uint32_t w0, uint32_t w1) {
	uint64_t h = uint64_t(w0) | (uint64_t(w1) << 32);
	h ^= h >> 33;
	h *= 0xff51afd7ed558ccdull;
	h ^= h >> 33;
	h *= 0xc4ceb9fe1a85ec53ull;
	h ^= h >> 33;
	return uint32_t(h);
}
__device__ __forceinline__ uint64_t mix64(uint64_t h) {
	h ^= h >> 30;
	h *= 0xbf58476d1ce4e5b9ull;
	h ^= h >> 27;
	h *= 0x94d049bb133111ebull;
	h ^= h >> 31;
	return h;
}

// ---- edit-list filtering for one (node, child) pair ------------------------------------------------------
// Scans the parent's list from the back: the last kFill/kClear decides the child's base pointer, the kProceed
// edits after it survive.  Lists of <= 32 entries are evaluated once (decisions cached in a mask).
struct Filtered {
	uint32_t cur, count, start, keep;
};
// An edit that cannot change a subtree in its current state: digging empty space, or filling a filled node
// (the reference recursion returns the same pointer in both cases, NodePool.hpp:340-342,391-395).
__device__ __forceinline__ bool is_noop(uint32_t kind, uint32_t cur, uint32_t filled_ptr) {
	return kind == HD_EDIT_SPHERE_DIG ? cur == kNull : cur == filled_ptr;
}
template <bool kTerrain>
__device__ inline Filtered filter_list(const hd_edit_desc *__restrict__ edits, const uint32_t *__restrict__ list,
                                       uint32_t len, uint32_t bits, uint32_t x, uint32_t y, uint32_t z, uint32_t cur,
                                       uint32_t filled_ptr) {
	Filtered f{cur, 0u, 0u, 0u};
	int j = int(len) - 1;
	for (; j >= 0; --j) {
		const EditType t = edit_node<kTerrain>(edits[list[j]], bits, x, y, z);
		if (t == kFill) {
			f.cur = filled_ptr;
			break;
		}
		if (t == kClear) {
			f.cur = kNull;
			break;
		}
		if (t == kProceed) {
			++f.count;
			if (len <= 32)
				f.keep |= 1u << j;
		}
	}
	f.start = uint32_t(j + 1);
	// drop leading edits that are no-ops on the base state
	if (f.count && (f.cur == kNull || f.cur == filled_ptr)) {
		if (len <= 32) {
			while (f.keep) {
				const uint32_t k = __ffs(f.keep) - 1;
				if (!is_noop(edits[list[k]].kind, f.cur, filled_ptr))
					break;
				f.keep &= f.keep - 1;
				--f.count;
				f.start = k + 1;
			}
		} else {
			while (f.count) {
				const hd_edit_desc &e = edits[list[f.start]];
				const bool proceed = edit_node<kTerrain>(e, bits, x, y, z) == kProceed;
				if (proceed && !is_noop(e.kind, f.cur, filled_ptr))
					break;
				if (proceed)
					--f.count;
				++f.start;
			}
		}
	}
	return f;
}
template <bool kTerrain>
__device__ inline void write_list(const hd_edit_desc *__restrict__ edits, const uint32_t *__restrict__ list, uint32_t len,
                                  uint32_t bits, uint32_t x, uint32_t y, uint32_t z, const Filtered &f, uint32_t *dst) {
	if (len <= 32) {
		uint32_t keep = f.keep;
		while (keep) {
			const uint32_t j = __ffs(keep) - 1;
			keep &= keep - 1;
			*dst++ = list[j];
		}
	} else {
		for (uint32_t j = f.start; j < len; ++j)
			if (edit_node<kTerrain>(edits[list[j]], bits, x, y, z) == kProceed)
				*dst++ = list[j];
	}
}

// Allocate one slot in the next level (and `count` list entries): CTA-aggregated — ONE pair of atomics per CTA and trip.
// (Per-warp aggregation put 2 x 7.5 M atomics on one cache line at the level above the leaves of the cfg3 batch:
// ~8 of the kernel's 11.7 ms were the L2 atomic unit serialising them.)
// Must be called by every thread of the CTA (threads without work pass want = false); `s_alloc` is 2 x 17 words of
// shared memory.
constexpr uint32_t kMaxWarps = 16;
__device__ __forceinline__ bool alloc_item(DevCounters *ctr, uint32_t *items_ctr, uint32_t *entries_ctr, bool want,
                                           uint32_t count, uint32_t cap, uint32_t cap_entries, uint32_t &item,
                                           uint32_t &entry_off, uint32_t *s_alloc) {
	const uint32_t full = 0xFFFFFFFFu;
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31u) >> 5;
	const uint32_t wants = __ballot_sync(full, want);
	uint32_t c = want ? count : 0u, scan = c; // inclusive warp scan of the list lengths
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t v = __shfl_up_sync(full, scan, d);
		if (lane >= uint32_t(d))
			scan += v;
	}
	uint32_t *s_items = s_alloc, *s_entries = s_alloc + kMaxWarps + 1;
	if (lane == 31u)
		s_items[warp] = __popc(wants), s_entries[warp] = scan;
	__syncthreads();
	if (threadIdx.x == 0) { // exclusive scan over the warps' totals, one reservation for the whole CTA
		uint32_t ti = 0, te = 0;
		for (uint32_t w = 0; w < nwarps; ++w) {
			const uint32_t a = s_items[w], b = s_entries[w];
			s_items[w] = ti, s_entries[w] = te;
			ti += a, te += b;
		}
		uint32_t bi = 0, be = 0;
		if (ti) {
			bi = atomicAdd(items_ctr, ti);
			be = atomicAdd(entries_ctr, te);
		}
		s_items[kMaxWarps] = bi, s_entries[kMaxWarps] = be;
	}
	__syncthreads();
	item = s_items[kMaxWarps] + s_items[warp] + __popc(wants & ((1u << lane) - 1u));
	entry_off = s_entries[kMaxWarps] + s_entries[warp] + scan - c;
	__syncthreads(); // the next trip overwrites s_alloc
	if (want && (item >= cap || entry_off + count > cap_entries)) {
		ctr->error = 1;
		return false;
	}
	return want;
}

// The same reservation per WARP (ballot, warp scan, one pair of atomics by lane 0): no CTA barrier, so the warps of a CTA run
// on independently and their load chains overlap.  For the one-launch path, whose levels hold at most 2^19 items (tens of
// thousands of atomics, not millions).  Must be called by all 32 lanes.
__device__ __forceinline__ bool alloc_item_warp(DevCounters *ctr, uint32_t *items_ctr, uint32_t *entries_ctr, bool want,
                                                uint32_t count, uint32_t cap, uint32_t cap_entries, uint32_t &item,
                                                uint32_t &entry_off) {
	const uint32_t full = 0xFFFFFFFFu, lane = threadIdx.x & 31u;
	const uint32_t wants = __ballot_sync(full, want);
	if (wants == 0u)
		return false;
	uint32_t c = want ? count : 0u, scan = c;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t v = __shfl_up_sync(full, scan, d);
		if (lane >= uint32_t(d))
			scan += v;
	}
	uint32_t bi = 0, be = 0;
	if (lane == 31u) {
		bi = atomicAdd(items_ctr, uint32_t(__popc(wants)));
		be = atomicAdd(entries_ctr, scan);
	}
	bi = __shfl_sync(full, bi, 31), be = __shfl_sync(full, be, 31);
	item = bi + __popc(wants & ((1u << lane) - 1u));
	entry_off = be + scan - c;
	if (want && (item >= cap || entry_off + count > cap_entries)) {
		ctr->error = 1;
		return false;
	}
	return want;
}

// ---- warp-cooperative filtering for long lists (all 32 lanes pass the same arguments) -------------------------------
struct WarpFiltered {
	uint32_t cur, count, start;
};
__device__ inline WarpFiltered warp_filter_count(const hd_edit_desc *__restrict__ edits, const uint32_t *__restrict__ list,
                                                 uint32_t len, uint32_t bits, uint32_t x, uint32_t y, uint32_t z,
                                                 uint32_t cur, uint32_t filled_ptr) {
	const uint32_t lane = threadIdx.x & 31u, full = 0xFFFFFFFFu;
	WarpFiltered f{cur, 0u, 0u};
	// 128 entries at a time, from the back of the list: the four classifications of a lane are independent, so their
	// list -> descriptor load chains are in flight together (one chunk of 32 per trip made the root pass of a 10 000-edit
	// batch a chain of 626 serialised global round trips: 0.53 ms on one warp)
	for (uint32_t done = 0; done < len; done += 128u) {
		uint32_t t[4];
#pragma unroll
		for (uint32_t q = 0; q < 4u; ++q) {
			const uint32_t k = done + q * 32u + lane;
			t[q] = k < len ? uint32_t(edit_node(edits[list[len - 1u - k]], bits, x, y, z)) : uint32_t(kNotAffected);
		}
#pragma unroll
		for (uint32_t q = 0; q < 4u; ++q) {
			const uint32_t term = __ballot_sync(full, t[q] == kFill || t[q] == kClear);
			const uint32_t proc = __ballot_sync(full, t[q] == kProceed);
			if (term) {
				const uint32_t first = __ffs(term) - 1u; // the terminal edit nearest to the end of the list
				f.count += __popc(proc & ((1u << first) - 1u));
				f.cur = __shfl_sync(full, t[q], first) == kFill ? filled_ptr : kNull;
				f.start = len - (done + q * 32u + first);
				return f;
			}
			f.count += __popc(proc);
		}
	}
	return f;
}
// writes the kProceed edits of list[start..len) in order; then drops leading no-ops.  Returns {skip, remaining}.
__device__ inline uint2 warp_filter_write(const hd_edit_desc *__restrict__ edits, const uint32_t *__restrict__ list,
                                          uint32_t len, uint32_t bits, uint32_t x, uint32_t y, uint32_t z,
                                          const WarpFiltered &f, uint32_t filled_ptr, uint32_t *dst) {
	const uint32_t lane = threadIdx.x & 31u, full = 0xFFFFFFFFu;
	uint32_t out = 0;
	for (uint32_t base = f.start; base < len; base += 128u) {
		bool keep[4];
		uint32_t e[4];
#pragma unroll
		for (uint32_t q = 0; q < 4u; ++q) {
			const uint32_t idx = base + q * 32u + lane;
			e[q] = idx < len ? list[idx] : 0u;
			keep[q] = idx < len && edit_node(edits[e[q]], bits, x, y, z) == kProceed;
		}
#pragma unroll
		for (uint32_t q = 0; q < 4u; ++q) {
			const uint32_t proc = __ballot_sync(full, keep[q]);
			if (keep[q])
				dst[out + __popc(proc & ((1u << lane) - 1u))] = e[q];
			out += __popc(proc);
		}
	}
	__syncwarp();
	uint32_t skip = 0;
	if (lane == 0 && (f.cur == kNull || f.cur == filled_ptr))
		while (skip < out && is_noop(edits[dst[skip]].kind, f.cur, filled_ptr))
			++skip;
	skip = __shfl_sync(full, skip, 0);
	return make_uint2(skip, out - skip);
}

// Root classification (edit_switch on the root, NodePool.hpp:405-413).  One warp.
// `dyn` (low-latency path): {root, n_edits} live in device memory so that one instantiated graph serves every call.
__device__ __forceinline__ void phase_root(const Geometry &g, const hd_edit_desc *__restrict__ edits, uint32_t n_edits,
                                           const uint32_t *__restrict__ iota, const uint32_t *__restrict__ filled,
                                           uint32_t root, const LevelView &out, DevCounters *ctr) {
	const uint32_t bits = g.voxel_level();
	const WarpFiltered f = warp_filter_count(edits, iota, n_edits, bits, 0, 0, 0, root, filled[0]);
	uint2 r = make_uint2(0u, 0u);
	if (f.count)
		r = warp_filter_write(edits, iota, n_edits, bits, 0, 0, 0, f, filled[0], out.lists);
	if ((threadIdx.x & 31u) != 0)
		return;
	if (r.y == 0) {
		ctr->root_out = f.cur;
		return;
	}
	ctr->next_items = 1;
	ctr->next_entries = r.x + r.y;
	ctr->lvl_items[0] = 1;
	ctr->lvl_entries[0] = r.x + r.y;
	ctr->lvl_long[0] = r.y > 32u ? 1u : 0u;
	out.cur[0] = f.cur;
	out.pos[0] = 0;
	out.parent[0] = 0xFFFFFFFFu;
	out.list_off[0] = r.x;
	out.list_len[0] = r.y;
}
__global__ void k_root(Geometry g, const hd_edit_desc *__restrict__ edits, uint32_t n_edits,
                       const uint32_t *__restrict__ iota, const uint32_t *__restrict__ filled, uint32_t root,
                       LevelView out, DevCounters *ctr, const uint32_t *__restrict__ dyn) {
	if (dyn)
		root = dyn[0], n_edits = dyn[1];
	phase_root(g, edits, n_edits, iota, filled, root, out, ctr);
}

// ---- CTA-cooperative filtering for huge lists (the first levels of a 10 000-edit batch) --------------------------------
// One warp walking a 10 000-entry list costs 0.5 ms per (item, child) and there are only 1 + 8 + 64 of them: a whole CTA
// takes a list instead.  Three passes over the list (the classification is recomputed rather than stored): the last
// kFill / kClear, the number of kProceed edits behind it, and their ordered compaction (one CTA-wide scan per chunk).
constexpr uint32_t kHugeList = 1024u;
constexpr int kHugeThreads = 256;
struct CtaFiltered {
	uint32_t cur, count, start;
};
__device__ inline CtaFiltered cta_filter_count(const hd_edit_desc *__restrict__ edits, const uint32_t *__restrict__ list, uint32_t len,
                                               uint32_t bits, uint32_t x, uint32_t y, uint32_t z, uint32_t cur, uint32_t filled_ptr,
                                               uint32_t *s_tmp /* 2 words */) {
	if (threadIdx.x == 0)
		s_tmp[0] = 0u, s_tmp[1] = 0u;
	__syncthreads();
	uint32_t best = 0u; // (index + 1) << 1 | is_fill of the last terminal edit this thread saw
	for (uint32_t k = threadIdx.x; k < len; k += blockDim.x) {
		const EditType t = edit_node(edits[list[k]], bits, x, y, z);
		if (t == kFill || t == kClear)
			best = ((k + 1u) << 1) | (t == kFill ? 1u : 0u);
	}
	if (best)
		atomicMax(&s_tmp[0], best);
	__syncthreads();
	const uint32_t term = s_tmp[0];
	CtaFiltered f{cur, 0u, term >> 1};
	if (term)
		f.cur = (term & 1u) ? filled_ptr : kNull;
	uint32_t mine = 0u;
	for (uint32_t k = f.start + threadIdx.x; k < len; k += blockDim.x)
		mine += edit_node(edits[list[k]], bits, x, y, z) == kProceed ? 1u : 0u;
	for (int d = 16; d; d >>= 1)
		mine += __shfl_xor_sync(0xFFFFFFFFu, mine, d);
	if ((threadIdx.x & 31u) == 0u && mine)
		atomicAdd(&s_tmp[1], mine);
	__syncthreads();
	f.count = s_tmp[1];
	__syncthreads();
	return f;
}
// ordered compaction of the kProceed edits of list[start..len) into dst, then the leading no-ops are dropped: {skip, remaining}
__device__ inline uint2 cta_filter_write(const hd_edit_desc *__restrict__ edits, const uint32_t *__restrict__ list, uint32_t len,
                                         uint32_t bits, uint32_t x, uint32_t y, uint32_t z, const CtaFiltered &f, uint32_t filled_ptr,
                                         uint32_t *dst, uint32_t *s_scan /* kHugeThreads / 32 + 2 words */) {
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
	uint32_t base = 0u;
	for (uint32_t c0 = f.start; c0 < len; c0 += blockDim.x) {
		const uint32_t idx = c0 + threadIdx.x;
		const uint32_t e = idx < len ? list[idx] : 0u;
		const bool keep = idx < len && edit_node(edits[e], bits, x, y, z) == kProceed;
		const uint32_t vote = __ballot_sync(0xFFFFFFFFu, keep);
		if (lane == 0u)
			s_scan[warp] = __popc(vote);
		__syncthreads();
		uint32_t off = 0u, total = 0u;
		for (uint32_t w = 0; w < nw; ++w) {
			const uint32_t c = s_scan[w];
			off += w < warp ? c : 0u, total += c;
		}
		if (keep)
			dst[base + off + __popc(vote & ((1u << lane) - 1u))] = e;
		base += total;
		__syncthreads();
	}
	__threadfence_block();
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t skip = 0u;
		if (f.cur == kNull || f.cur == filled_ptr)
			while (skip < base && is_noop(edits[dst[skip]].kind, f.cur, filled_ptr))
				++skip;
		s_scan[nw] = skip;
	}
	__syncthreads();
	const uint32_t skip = s_scan[nw];
	__syncthreads();
	return make_uint2(skip, base - skip);
}

// Root classification for batches of more than kHugeList edits: one CTA instead of one warp (phase_root otherwise).
__global__ void __launch_bounds__(kHugeThreads) k_root_cta(Geometry g, const hd_edit_desc *__restrict__ edits, uint32_t n_edits,
                                                           const uint32_t *__restrict__ iota, const uint32_t *__restrict__ filled,
                                                           uint32_t root, LevelView out, DevCounters *ctr) {
	__shared__ uint32_t s_tmp[2], s_scan[kHugeThreads / 32 + 2];
	const uint32_t bits = g.voxel_level();
	const CtaFiltered f = cta_filter_count(edits, iota, n_edits, bits, 0, 0, 0, root, filled[0], s_tmp);
	uint2 r = make_uint2(0u, 0u);
	if (f.count)
		r = cta_filter_write(edits, iota, n_edits, bits, 0, 0, 0, f, filled[0], out.lists, s_scan);
	if (threadIdx.x != 0)
		return;
	if (r.y == 0) {
		ctr->root_out = f.cur;
		return;
	}
	ctr->next_items = 1;
	ctr->next_entries = r.x + r.y;
	ctr->lvl_items[0] = 1;
	ctr->lvl_entries[0] = r.x + r.y;
	out.cur[0] = f.cur;
	out.pos[0] = 0;
	out.parent[0] = 0xFFFFFFFFu;
	out.list_off[0] = r.x;
	out.list_len[0] = r.y;
}

// Top-down expansion of the items whose edit list is longer than kHugeList: one CTA per (item, child).
__global__ void __launch_bounds__(kHugeThreads) k_down_huge(Geometry g, uint32_t level, const uint32_t *__restrict__ words,
                                                            const hd_edit_desc *__restrict__ edits,
                                                            const uint32_t *__restrict__ filled, LevelView in, LevelView out,
                                                            DevCounters *ctr) {
	__shared__ uint32_t s_tmp[2], s_scan[kHugeThreads / 32 + 2], s_off;
	const uint32_t item = blockIdx.x >> 3, c = blockIdx.x & 7u;
	if (item >= in.n)
		return;
	const uint32_t len = in.list_len[item];
	if (len <= kHugeList)
		return; // k_down / k_down_long
	const uint32_t cur = in.cur[item];
	uint32_t child = kNull;
	if (cur != kNull) {
		const uint32_t mask = words[cur];
		if (mask >> c & 1u)
			child = words[cur + 1u + __popc(mask & ((1u << c) - 1u))];
	}
	uint32_t x, y, z;
	unpack_pos(in.pos[item], x, y, z);
	x = (x << 1) | (c & 1u), y = (y << 1) | ((c >> 1) & 1u), z = (z << 1) | ((c >> 2) & 1u);
	const uint32_t bits = g.voxel_level() - (level + 1u);
	const uint32_t *list = in.lists + in.list_off[item];
	const uint32_t fp = filled[level + 1u];
	const CtaFiltered f = cta_filter_count(edits, list, len, bits, x, y, z, child, fp, s_tmp);
	uint32_t result = f.cur;
	if (f.count) {
		if (threadIdx.x == 0)
			s_off = atomicAdd(&ctr->next_entries, f.count);
		__syncthreads();
		const uint32_t entry_off = s_off;
		if (entry_off + f.count > out.cap_entries) {
			if (threadIdx.x == 0)
				ctr->error = 1;
			return;
		}
		const uint2 r = cta_filter_write(edits, list, len, bits, x, y, z, f, fp, out.lists + entry_off, s_scan);
		if (r.y) {
			if (threadIdx.x == 0) {
				const uint32_t slot = atomicAdd(&ctr->next_items, 1u);
				if (slot >= out.cap)
					ctr->error = 1;
				else {
					out.cur[slot] = f.cur;
					out.pos[slot] = pack_pos(x, y, z);
					out.parent[slot] = (item << 3) | c;
					out.list_off[slot] = entry_off + r.x;
					out.list_len[slot] = r.y;
					if (r.y > 32u)
						ctr->has_long = 1;
					if (r.y > kHugeList)
						ctr->has_huge = 1;
				}
			}
			result = kPending;
		}
	}
	if (threadIdx.x == 0)
		in.child_new[size_t(item) * 8u + c] = result;
}

// Top-down expansion of the (rare) items whose edit list is longer than 32: one warp per (item, child).
__global__ void __launch_bounds__(kBlock) k_down_long(Geometry g, uint32_t level, const uint32_t *__restrict__ words,
                                                      const hd_edit_desc *__restrict__ edits,
                                                      const uint32_t *__restrict__ filled, LevelView in, LevelView out,
                                                      DevCounters *ctr) {
	const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
	const uint32_t item = w >> 3, c = w & 7u;
	if (item >= in.n) // host-sized launch only (lists longer than 32 never take the low-latency path)
		return;
	const uint32_t len = in.list_len[item];
	if (len <= 32u || (len > kHugeList && !in.n_dev))
		return; // done by k_down / k_down_huge (host-driven path)
	const uint32_t cur = in.cur[item];
	uint32_t child = kNull;
	if (cur != kNull) {
		const uint32_t mask = words[cur];
		if (mask >> c & 1u)
			child = words[cur + 1u + __popc(mask & ((1u << c) - 1u))];
	}
	uint32_t x, y, z;
	unpack_pos(in.pos[item], x, y, z);
	x = (x << 1) | (c & 1u), y = (y << 1) | ((c >> 1) & 1u), z = (z << 1) | ((c >> 2) & 1u);
	const uint32_t bits = g.voxel_level() - (level + 1u);
	const uint32_t *list = in.lists + in.list_off[item];
	const uint32_t fp = filled[level + 1u];
	const WarpFiltered f = warp_filter_count(edits, list, len, bits, x, y, z, child, fp);
	uint32_t result = f.cur;
	if (f.count) {
		uint32_t entry_off = 0;
		if (lane == 0)
			entry_off = atomicAdd(&ctr->next_entries, f.count);
		entry_off = __shfl_sync(0xFFFFFFFFu, entry_off, 0);
		if (entry_off + f.count > out.cap_entries) {
			if (lane == 0)
				ctr->error = 1;
			return;
		}
		const uint2 r = warp_filter_write(edits, list, len, bits, x, y, z, f, fp, out.lists + entry_off);
		if (r.y) {
			if (lane == 0) {
				const uint32_t slot = atomicAdd(&ctr->next_items, 1u);
				if (slot >= out.cap)
					ctr->error = 1;
				else {
					out.cur[slot] = f.cur;
					out.pos[slot] = pack_pos(x, y, z);
					out.parent[slot] = (item << 3) | c;
					out.list_off[slot] = entry_off + r.x;
					out.list_len[slot] = r.y;
					if (r.y > 32u)
						ctr->has_long = 1;
				}
			}
			result = kPending;
		}
	}
	if (lane == 0)
		in.child_new[size_t(item) * 8u + c] = result;
}

// Top-down expansion: thread per (item, child); whole warps stride over the level (one trip when the host sized the
// grid from a known item count, a grid-stride loop over the device-resident count on the low-latency path).
// (tid0, nthreads) = this thread's index among, and the number of, the threads that share the level: the whole grid, or
// one CTA when the fused kernel walks a small level alone.
// kWarpAlloc (fused kernel): queue slots are reserved per warp (alloc_item_warp) and lists longer than 32 entries are left
// to phase_down_long; otherwise per CTA, and `in.n_dev` set (the CUDA-graph variant) means this thread walks long lists itself.
template <bool kTerrain, bool kWarpAlloc = false>
__device__ __forceinline__ void phase_down(const Geometry &g, uint32_t level /* of `in` */,
                                           const uint32_t *__restrict__ words, const hd_edit_desc *__restrict__ edits,
                                           const uint32_t *__restrict__ filled, const LevelView &in, const LevelView &out,
                                           DevCounters *ctr, uint32_t *items_ctr, uint32_t *entries_ctr, uint32_t tid0,
                                           uint32_t nthreads, uint32_t *s_alloc, uint32_t n_known = 0xFFFFFFFFu) {
	// n_known: the caller already holds the level's item count (the solo stage of the fused kernel keeps it in shared memory)
	const uint32_t n8 = (n_known != 0xFFFFFFFFu ? n_known : in.count()) * 8u;
	const uint32_t bits = g.voxel_level() - (level + 1u);
	const bool serial_long = !kWarpAlloc && in.n_dev != nullptr;
	// whole CTAs (alloc_item synchronises) or whole warps (alloc_item_warp shuffles) iterate together
	for (uint32_t t = tid0; t - (kWarpAlloc ? (threadIdx.x & 31u) : threadIdx.x) < n8; t += nthreads) {
		const uint32_t item = t >> 3, c = t & 7u;
		const bool valid = t < n8;
		Filtered f{kNull, 0u, 0u, 0u};
		uint32_t x = 0, y = 0, z = 0, len = 0;
		const uint32_t *list = nullptr;
		if (valid) {
			const uint32_t cur = in.cur[item];
			uint32_t child = kNull;
			if (cur != kNull) { // get_unpacked_node_array, NodePool.hpp:278-309
				const uint32_t mask = words[cur];
				if (mask >> c & 1u)
					child = words[cur + 1u + __popc(mask & ((1u << c) - 1u))];
			}
			unpack_pos(in.pos[item], x, y, z);
			x = (x << 1) | (c & 1u), y = (y << 1) | ((c >> 1) & 1u), z = (z << 1) | ((c >> 2) & 1u); // NodeCoord.hpp:18-28
			list = in.lists + in.list_off[item];
			len = in.list_len[item];
			// lists longer than 32: k_down_long on the host-driven path; on the one-launch path (device-resident counts)
			// this thread walks them itself — they only occur in the few levels next to the root
			if (len <= 32u || serial_long)
				f = filter_list<kTerrain>(edits, list, len, bits, x, y, z, child, filled[level + 1u]);
		}
		uint32_t slot, entry_off;
		const bool made = kWarpAlloc ? alloc_item_warp(ctr, items_ctr, entries_ctr, valid && f.count != 0, f.count, out.cap,
		                                               out.cap_entries, slot, entry_off)
		                             : alloc_item(ctr, items_ctr, entries_ctr, valid && f.count != 0, f.count, out.cap,
		                                          out.cap_entries, slot, entry_off, s_alloc);
		if (!valid || (len > 32u && !serial_long))
			continue; // long lists: k_down_long / phase_down_long
		if (made) {
			out.cur[slot] = f.cur;
			out.pos[slot] = pack_pos(x, y, z);
			out.parent[slot] = (item << 3) | c;
			out.list_off[slot] = entry_off;
			out.list_len[slot] = f.count;
			write_list<kTerrain>(edits, list, len, bits, x, y, z, f, out.lists + entry_off);
			in.child_new[size_t(item) * 8u + c] = kPending;
		} else {
			in.child_new[size_t(item) * 8u + c] = f.cur;
		}
	}
}
// Fused kernel: the (item, child) pairs whose parent carries a list longer than 32 entries — the levels next to the root of
// a batch of dozens to a thousand editors.  A warp per pair filters the list cooperatively (warp_filter_*); walking such
// lists with one thread per pair made the top levels of a 100-editor batch 0.45 of its 1.15 ms.  (w0, nw) = this warp's
// index among, and the number of, the warps sharing the level.  Small levels stride over pairs; on large levels a warp
// checks 32 items' list lengths at a time and only visits the long ones.
__device__ __forceinline__ void phase_down_long(const Geometry &g, uint32_t level, const uint32_t *__restrict__ words,
                                                const hd_edit_desc *__restrict__ edits, const uint32_t *__restrict__ filled,
                                                const LevelView &in, const LevelView &out, DevCounters *ctr, uint32_t *items_ctr,
                                                uint32_t *entries_ctr, uint32_t *long_flag, uint32_t w0, uint32_t nw) {
	const uint32_t lane = threadIdx.x & 31u, full = 0xFFFFFFFFu;
	const uint32_t n = in.count(), bits = g.voxel_level() - (level + 1u), fp = filled[level + 1u];
	auto do_pair = [&](uint32_t item, uint32_t c, uint32_t len) {
		const uint32_t cur = in.cur[item];
		uint32_t child = kNull;
		if (cur != kNull) {
			const uint32_t mask = words[cur];
			if (mask >> c & 1u)
				child = words[cur + 1u + __popc(mask & ((1u << c) - 1u))];
		}
		uint32_t x, y, z;
		unpack_pos(in.pos[item], x, y, z);
		x = (x << 1) | (c & 1u), y = (y << 1) | ((c >> 1) & 1u), z = (z << 1) | ((c >> 2) & 1u);
		const uint32_t *list = in.lists + in.list_off[item];
		const WarpFiltered f = warp_filter_count(edits, list, len, bits, x, y, z, child, fp);
		uint32_t result = f.cur;
		if (f.count) {
			uint32_t entry_off = 0;
			if (lane == 0)
				entry_off = atomicAdd(entries_ctr, f.count);
			entry_off = __shfl_sync(full, entry_off, 0);
			if (entry_off + f.count > out.cap_entries) {
				if (lane == 0)
					ctr->error = 1;
				return;
			}
			const uint2 r = warp_filter_write(edits, list, len, bits, x, y, z, f, fp, out.lists + entry_off);
			if (r.y) {
				if (lane == 0) {
					const uint32_t slot = atomicAdd(items_ctr, 1u);
					if (slot >= out.cap)
						ctr->error = 1;
					else {
						out.cur[slot] = f.cur;
						out.pos[slot] = pack_pos(x, y, z);
						out.parent[slot] = (item << 3) | c;
						out.list_off[slot] = entry_off + r.x;
						out.list_len[slot] = r.y;
						if (r.y > 32u)
							*long_flag = 1u;
					}
				}
				result = kPending;
			}
		}
		if (lane == 0)
			in.child_new[size_t(item) * 8u + c] = result;
	};
	if (n * 8u <= nw * 8u) { // few pairs: one warp each
		for (uint32_t w = w0; w < n * 8u; w += nw) {
			const uint32_t len = in.list_len[w >> 3];
			if (len > 32u)
				do_pair(w >> 3, w & 7u, len);
		}
	} else { // many items, few of them long: 32 list lengths per trip
		for (uint32_t base = w0 * 32u; base < n; base += nw * 32u) {
			const uint32_t mine = base + lane < n ? in.list_len[base + lane] : 0u;
			uint32_t todo = __ballot_sync(full, mine > 32u);
			while (todo) {
				const uint32_t k = __ffs(todo) - 1u;
				todo &= todo - 1u;
				const uint32_t len = __shfl_sync(full, mine, k);
				for (uint32_t c = 0; c < 8u; ++c)
					do_pair(base + k, c, len);
			}
		}
	}
}

template <bool kTerrain>
__global__ void __launch_bounds__(kBlock) k_down(Geometry g, uint32_t level, const uint32_t *__restrict__ words,
                                                 const hd_edit_desc *__restrict__ edits,
                                                 const uint32_t *__restrict__ filled, LevelView in, LevelView out,
                                                 DevCounters *ctr, uint32_t *items_ctr, uint32_t *entries_ctr) {
	__shared__ uint32_t s_alloc[2 * (kMaxWarps + 1)];
	phase_down<kTerrain>(g, level, words, edits, filled, in, out, ctr, items_ctr, entries_ctr, blockIdx.x * blockDim.x + threadIdx.x,
	                     gridDim.x * blockDim.x, s_alloc);
}

// Leaf pass: one warp per 4x4x4 leaf (a persistent grid-stride variant measured 10-20 % slower: the per-leaf work is
// short and uniform, so the hardware CTA scheduler balances it better); lane l owns voxels l and l+32 (NodeCoord::GetLeafCoord, NodeCoord.hpp:32-43).
// kTerrain = false: the batch holds no terrain edit — the generator is compiled out and the kernel fits 32 registers
// (8 CTAs of 256 threads per SM; the pass is latency-bound on its list -> descriptor loads, so resident warps matter).
template <bool kTerrain>
__global__ void __launch_bounds__(kBlock, kTerrain ? 4 : 6) k_leaf(Geometry g, const uint32_t *__restrict__ words,
                                                                    const hd_edit_desc *__restrict__ edits, LevelView lv,
                                                                    DevCounters *ctr) {
	// A warp takes 32 consecutive items at a time: lane i fetches item i's queue record and old leaf (coalesced loads, 32
	// independent leaf gathers in flight), then the warp walks the 32 leaves one after the other with the record handed
	// round by shuffles, and lane i writes item i's result back (coalesced again).  One-warp-per-item loads cost a 32-byte
	// sector per 4-byte field; this way the per-leaf loads are the edit list and its descriptors only.
	const uint32_t lane = threadIdx.x & 31u, full = 0xFFFFFFFFu, n = lv.count();
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32u; base < n; base += warps * 32u) {
		const uint32_t mine = base + lane;
		uint32_t m_cur = kNull, m_off = 0, m_len = 0, m_w0 = 0, m_w1 = 0;
		uint64_t m_pos = 0;
		if (mine < n) {
			m_cur = lv.cur[mine], m_off = lv.list_off[mine], m_len = lv.list_len[mine], m_pos = lv.pos[mine];
			if (m_cur != kNull) {
				const uint2 w = *reinterpret_cast<const uint2 *>(words + m_cur);
				m_w0 = w.x, m_w1 = w.y;
			}
		}
		uint32_t o_res = m_cur, o_n0 = 0, o_n1 = 0;
		uint8_t o_st = 0;
		const uint32_t count = min(32u, n - base);
		for (uint32_t k = 0; k < count; ++k) {
			const uint32_t cur = __shfl_sync(full, m_cur, k), len = __shfl_sync(full, m_len, k);
			const uint32_t *list = lv.lists + __shfl_sync(full, m_off, k);
			const uint32_t w0 = __shfl_sync(full, m_w0, k), w1 = __shfl_sync(full, m_w1, k);
			uint32_t x, y, z;
			unpack_pos(__shfl_sync(full, m_pos, k), x, y, z);
			const uint32_t vx = (x << 2) | ((lane >> 2) & 2u) | (lane & 1u);
			const uint32_t vy = (y << 2) | ((lane >> 3) & 2u) | ((lane >> 1) & 1u);
			const uint32_t vz = (z << 2) | ((lane >> 2) & 1u); // voxel l: z1 = 0; voxel l+32: z1 = 1 (adds 2)
			bool a = w0 >> lane & 1u, b = w1 >> lane & 1u;
			for (uint32_t j = 0; j < len; ++j) {
				const hd_edit_desc &e = edits[list[j]];
				bool ia, ib;
				if (kTerrain && e.kind == HD_EDIT_TERRAIN_FILL && terrain_leaf_pair(e, x << 2, z << 2, vx, vy, vz, ia, ib)) { // warp-uniform
					a = a || ia, b = b || ib;
					continue;
				}
				edit_voxel_pair<kTerrain>(e, vx, vy, vz, a, b);
			}
			const uint32_t n0 = __ballot_sync(full, a), n1 = __ballot_sync(full, b);
			if (lane == k && (n0 != w0 || n1 != w1)) { // changed
				if ((n0 | n1) == 0u)
					o_res = kNull;
				else
					o_st = 1, o_n0 = n0, o_n1 = n1;
			}
		}
		if (mine < n) {
			if (o_st)
				*reinterpret_cast<uint2 *>(lv.cand + size_t(mine) * 2u) = make_uint2(o_n0, o_n1);
			lv.state[mine] = o_st;
			lv.result[mine] = o_res;
		}
	}
}

// Leaf pass for batches without a terrain edit: like k_leaf, but a HALF-warp works on a leaf (lane owns 4 voxels), so a
// trip of the inner loop finishes two leaves and the per-leaf overhead (record shuffles, coordinate unpacking, loop
// control, descriptor loads) is paid once per two.  Leaf bit i = [z1 y1 x1 z0 y0 x0] (NodeCoord.hpp:32-43): the lane's
// low four bits are [x1 z0 y0 x0], its voxels are (y1, z1) in {0,1}^2.
__global__ void __launch_bounds__(kBlock, 5) k_leaf_half(Geometry g, const uint32_t *__restrict__ words,
                                                         const hd_edit_desc *__restrict__ edits, LevelView lv, DevCounters *ctr) {
	const uint32_t lane = threadIdx.x & 31u, l16 = lane & 15u, half = lane >> 4, full = 0xFFFFFFFFu, n = lv.count();
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32u; base < n; base += warps * 32u) {
		const uint32_t mine = base + lane;
		uint32_t m_cur = kNull, m_off = 0, m_len = 0, m_w0 = 0, m_w1 = 0;
		uint64_t m_pos = 0;
		if (mine < n) {
			m_cur = lv.cur[mine], m_off = lv.list_off[mine], m_len = lv.list_len[mine], m_pos = lv.pos[mine];
			if (m_cur != kNull) {
				const uint2 w = *reinterpret_cast<const uint2 *>(words + m_cur);
				m_w0 = w.x, m_w1 = w.y;
			}
		}
		uint32_t o_res = m_cur, o_n0 = 0, o_n1 = 0;
		uint8_t o_st = 0;
		const uint32_t count = min(32u, n - base);
		for (uint32_t k = 0; k < count; k += 2u) {
			const uint32_t src = k + half; // items beyond `count` have m_len == 0 and m_cur == kNull: nothing happens
			const uint32_t len = __shfl_sync(full, m_len, src);
			const uint32_t *list = lv.lists + __shfl_sync(full, m_off, src);
			const uint32_t w0 = __shfl_sync(full, m_w0, src), w1 = __shfl_sync(full, m_w1, src);
			uint32_t x, y, z;
			unpack_pos(__shfl_sync(full, m_pos, src), x, y, z);
			const uint32_t vx = (x << 2) | ((l16 >> 2) & 2u) | (l16 & 1u);
			const uint32_t vy = (y << 2) | ((l16 >> 1) & 1u); // + 2 for y1
			const uint32_t vz = (z << 2) | ((l16 >> 2) & 1u); // + 2 for z1
			bool v[4] = {bool(w0 >> l16 & 1u), bool(w0 >> (l16 + 16u) & 1u), bool(w1 >> l16 & 1u), bool(w1 >> (l16 + 16u) & 1u)};
			for (uint32_t j = 0; j < len; ++j)
				edit_voxel_quad(edits[list[j]], vx, vy, vz, v);
			const uint32_t b0 = __ballot_sync(full, v[0]), b1 = __ballot_sync(full, v[1]);
			const uint32_t b2 = __ballot_sync(full, v[2]), b3 = __ballot_sync(full, v[3]);
			if (lane == k || lane == k + 1u) { // the owner of the leaf half `lane - k` worked on
				const uint32_t sh = (lane - k) << 4;
				const uint32_t n0 = ((b0 >> sh) & 0xFFFFu) | ((b1 >> sh) << 16), n1 = ((b2 >> sh) & 0xFFFFu) | ((b3 >> sh) << 16);
				if (n0 != m_w0 || n1 != m_w1) { // changed
					if ((n0 | n1) == 0u)
						o_res = kNull;
					else
						o_st = 1, o_n0 = n0, o_n1 = n1;
				}
			}
		}
		if (mine < n) {
			if (o_st)
				*reinterpret_cast<uint2 *>(lv.cand + size_t(mine) * 2u) = make_uint2(o_n0, o_n1);
			lv.state[mine] = o_st;
			lv.result[mine] = o_res;
		}
	}
}

// Leaf pass, ONE THREAD PER LEAF (batches without a terrain edit; HD_EDIT_LEAF=half brings k_leaf_half back).
// The cooperative kernels above spend most of their 78 warp instructions per leaf on plumbing (record shuffles, coordinate
// unpacking per lane, four ballots, the owner's re-assembly) and evaluate 4 voxels per lane; here a lane owns a whole
// 4x4x4 leaf: its record loads are coalesced across the warp, nothing is exchanged between lanes, and a sphere edit
// (EditVoxel, main.cpp:127-142) costs two instructions per voxel — the squares (dx+i)^2, (dy+j)^2, (dz+k)^2 are formed
// once (12 multiplies, 16 sums), then per voxel one three-input add r2 - X[i] - YZ[j,k] whose SIGN is the "outside" flag
// and one funnel shift that moves that sign bit into the 64-bit mask.  The 32-bit form is exact while every |d| < 26 000
// (3 * 26 003^2 < 2^31; r2 is clamped to 2^31 - 1, beyond which every such voxel is inside anyway); farther from the
// centre, and for AABBs, the voxels are evaluated one by one with the reference's 64-bit arithmetic.
__device__ __forceinline__ void leaf_mask_generic(const hd_edit_desc &e, uint32_t bx, uint32_t by, uint32_t bz, uint32_t &in0,
                                                  uint32_t &in1) {
	in0 = in1 = 0u;
	for (uint32_t i = 0; i < 64u; ++i) { // leaf bit i = [z1 y1 x1 z0 y0 x0] (NodeCoord.hpp:32-43)
		const uint32_t vx = bx | ((i >> 2) & 2u) | (i & 1u), vy = by | ((i >> 3) & 2u) | ((i >> 1) & 1u),
		               vz = bz | ((i >> 4) & 2u) | ((i >> 2) & 1u);
		if (voxel_in_range<false>(e, vx, vy, vz)) {
			if (i < 32u)
				in0 |= 1u << i;
			else
				in1 |= 1u << (i - 32u);
		}
	}
}
// EditVoxel of one edit over all 64 voxels of the leaf at voxel origin (bx, by, bz): (n0, n1) in, (n0, n1) out.
__device__ __forceinline__ void leaf_apply(const hd_edit_desc &e, uint32_t bx, uint32_t by, uint32_t bz, uint32_t &n0, uint32_t &n1) {
	const uint2 k0 = *reinterpret_cast<const uint2 *>(&e.kind), k1 = *reinterpret_cast<const uint2 *>(&e.p0[1]);
	const uint32_t kind = k0.x;
	uint32_t in0, in1;
	const int32_t dx = int32_t(bx - k0.y), dy = int32_t(by - k1.x), dz = int32_t(bz - k1.y);
	const uint32_t far = max(max(uint32_t(abs(dx)), uint32_t(abs(dx + 3))),
	                         max(max(uint32_t(abs(dy)), uint32_t(abs(dy + 3))), max(uint32_t(abs(dz)), uint32_t(abs(dz + 3)))));
	if ((kind == HD_EDIT_SPHERE_FILL || kind == HD_EDIT_SPHERE_DIG) && far < 26000u) {
		const uint64_t r2 = e.r2;
		const int32_t r = int32_t(r2 > 0x7FFFFFFFull ? 0x7FFFFFFFu : uint32_t(r2));
		int32_t X[4], YZ[16];
#pragma unroll
		for (int i = 0; i < 4; ++i)
			X[i] = (dx + i) * (dx + i);
#pragma unroll
		for (int a = 0; a < 4; ++a)
#pragma unroll
			for (int b = 0; b < 4; ++b)
				YZ[a * 4 + b] = (dy + a) * (dy + a) + (dz + b) * (dz + b);
		uint32_t m0 = 0u, m1 = 0u; // bit = 1: OUTSIDE (r2 - d^2 < 0)
#pragma unroll
		for (int i = 31; i >= 0; --i) {
			const int xi = ((i >> 3) & 1) * 2 + (i & 1), yi = ((i >> 4) & 1) * 2 + ((i >> 1) & 1), zi = (i >> 2) & 1;
			m0 = __funnelshift_l(uint32_t(r - X[xi] - YZ[yi * 4 + zi]), m0, 1);
			m1 = __funnelshift_l(uint32_t(r - X[xi] - YZ[yi * 4 + zi + 2]), m1, 1);
		}
		in0 = ~m0, in1 = ~m1;
	} else {
		leaf_mask_generic(e, bx, by, bz, in0, in1);
	}
	if (kind == HD_EDIT_SPHERE_DIG)
		n0 &= ~in0, n1 &= ~in1;
	else
		n0 |= in0, n1 |= in1;
}

__global__ void __launch_bounds__(kBlock) k_leaf_lane(Geometry g, const uint32_t *__restrict__ words,
                                                      const hd_edit_desc *__restrict__ edits, LevelView lv, DevCounters *ctr) {
	const uint32_t n = lv.count();
	for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < n; item += gridDim.x * blockDim.x) {
		const uint32_t cur = lv.cur[item], len = lv.list_len[item];
		const uint32_t *list = lv.lists + lv.list_off[item];
		uint32_t x, y, z, w0 = 0u, w1 = 0u;
		unpack_pos(lv.pos[item], x, y, z);
		if (cur != kNull) {
			const uint2 w = *reinterpret_cast<const uint2 *>(words + cur);
			w0 = w.x, w1 = w.y;
		}
		uint32_t n0 = w0, n1 = w1;
		for (uint32_t j = 0; j < len; ++j)
			leaf_apply(edits[list[j]], x << 2, y << 2, z << 2, n0, n1);
		uint32_t res = cur;
		uint8_t st = 0;
		if (n0 != w0 || n1 != w1) { // changed (edit_leaf tail, NodePool.hpp:336-342)
			if ((n0 | n1) == 0u)
				res = kNull;
			else {
				st = 1;
				*reinterpret_cast<uint2 *>(lv.cand + size_t(item) * 2u) = make_uint2(n0, n1);
			}
		}
		lv.state[item] = st;
		lv.result[item] = res;
	}
}

// Last inner level and the leaf pass in ONE kernel (batches without a terrain edit, lists of <= 32 entries): the children of
// the items of level L-2 are classified (EditNode on each leaf's box) and the surviving edits applied to the leaves
// (leaf_apply) right here, instead of writing a 24-byte queue record plus list per leaf for k_leaf_lane to read back.
// Only CHANGED, non-empty leaves get a slot in the (compact) leaf level that the find-or-insert then works on: the cfg3 batch
// visits 138.6 M leaves and changes a tenth of them, so the dedup / bucket-sort / resolve kernels of the leaf level shrink
// with it.  Unchanged and emptied leaves report straight into the parent's child_new.
// Round 2, second pass.  ncu on the cfg3 batch: 5.3 G warp instructions, issue slots 76 % busy, DRAM 8 % — the kernel is bound
// by the instructions of the CLASSIFICATION (EditNode of every child against every edit of the parent's list, ~45 instructions
// each, 277 M (item, child) pairs), not by the leaves (compacting the proceeding pairs so that leaf_apply ran with full warps
// raised the lanes per instruction from 20.7 to 27.5 and saved nothing).  So a thread now takes a whole ITEM: the eight
// children share three planes per axis, edit_node8 forms the per-axis terms once and classifies all eight (~8 instructions
// per child and edit), and the list is walked once instead of eight times.  The proceeding children are then compacted
// through shared memory and the CTA's threads apply the surviving edits to one leaf each, with full warps.
// An item's eight children against its edit list (<= 32 entries), last edit first — filter_list's semantics for all eight at
// once: the last kFill / kClear replaces the child and ends its scan (kProceed edits before it are dropped), kNotAffected
// edits are ignored, and leading edits that cannot change the child in its base state (is_noop) are dropped.
// curc[c] comes in as the child's old pointer and goes out as its base pointer; keep[c] = surviving edits (bit j = list[j]).
__device__ __forceinline__ void classify_children(const hd_edit_desc *__restrict__ edits, const uint32_t *__restrict__ list,
                                                  uint32_t len, uint32_t bits, uint32_t x, uint32_t y, uint32_t z,
                                                  uint32_t filled_ptr, uint32_t (&curc)[8], uint32_t (&keep)[8]) {
#pragma unroll
	for (int c = 0; c < 8; ++c)
		keep[c] = 0u;
	uint32_t open = 0xFFu; // children whose scan has not met a terminal edit yet
	for (int j = int(len) - 1; j >= 0 && open; --j) {
		const uint32_t types = edit_node8(edits[list[j]], bits, x, y, z);
#pragma unroll
		for (int c = 0; c < 8; ++c) {
			const uint32_t ty = (types >> (2 * c)) & 3u;
			if ((open >> c & 1u) && ty != kNotAffected) {
				if (ty == kProceed)
					keep[c] |= 1u << j;
				else
					curc[c] = ty == kFill ? filled_ptr : kNull, open &= ~(1u << c);
			}
		}
	}
#pragma unroll
	for (int c = 0; c < 8; ++c)
		if (keep[c] && (curc[c] == kNull || curc[c] == filled_ptr))
			while (keep[c] && is_noop(edits[list[__ffs(keep[c]) - 1]].kind, curc[c], filled_ptr))
				keep[c] &= keep[c] - 1u;
}

// Top-down expansion of one inner level with a thread per ITEM (batches without a terrain edit; items whose list is longer
// than 32 entries are left to k_down_long / k_down_huge): what phase_down does with a thread per (item, child) pair, with
// the classification shared between the eight children (classify_children) and ONE pair of atomics per CTA for the
// queue slots and list entries of everything the CTA's items produce.
constexpr uint32_t kDownItems = 128;
__global__ void __launch_bounds__(kDownItems) k_down_items(Geometry g, uint32_t level, const uint32_t *__restrict__ words,
                                                           const hd_edit_desc *__restrict__ edits,
                                                           const uint32_t *__restrict__ filled, LevelView in, LevelView out,
                                                           DevCounters *ctr, uint32_t *items_ctr, uint32_t *entries_ctr) {
	__shared__ uint32_t s_wi[kMaxWarps + 1], s_we[kMaxWarps + 1], s_base[2];
	const uint32_t n = in.count(), bits = g.voxel_level() - (level + 1u), filled_ptr = filled[level + 1u];
	const uint32_t item = blockIdx.x * kDownItems + threadIdx.x;
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = kDownItems >> 5;
	uint32_t curc[8], keep[8];
#pragma unroll
	for (int c = 0; c < 8; ++c)
		curc[c] = kNull, keep[c] = 0u;
	uint32_t x = 0, y = 0, z = 0, want = 0u, n_entries = 0u;
	const uint32_t *list = nullptr;
	if (item < n && in.list_len[item] <= 32u) {
		const uint32_t cur = in.cur[item];
		if (cur != kNull) { // get_unpacked_node_array, NodePool.hpp:278-309
			const uint32_t mask = words[cur];
			uint32_t k = 1u;
#pragma unroll
			for (int c = 0; c < 8; ++c)
				if (mask >> c & 1u)
					curc[c] = words[cur + k++];
		}
		unpack_pos(in.pos[item], x, y, z);
		list = in.lists + in.list_off[item];
		classify_children(edits, list, in.list_len[item], bits, x, y, z, filled_ptr, curc, keep);
#pragma unroll
		for (int c = 0; c < 8; ++c) {
			if (keep[c])
				want |= 1u << c, n_entries += __popc(keep[c]);
			else
				in.child_new[size_t(item) * 8u + c] = curc[c];
		}
	}
	// ---- queue slots and list entries for the CTA: warp scans, one pair of atomics ----
	const uint32_t n_items = __popc(want);
	uint32_t ii = n_items, ie = n_entries;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, ii, d), b = __shfl_up_sync(0xFFFFFFFFu, ie, d);
		if (lane >= uint32_t(d))
			ii += a, ie += b;
	}
	if (lane == 31u)
		s_wi[warp] = ii, s_we[warp] = ie;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t ti = 0, te = 0;
		for (uint32_t w = 0; w < nwarps; ++w) {
			const uint32_t a = s_wi[w], b = s_we[w];
			s_wi[w] = ti, s_we[w] = te;
			ti += a, te += b;
		}
		uint32_t bi = 0, be = 0;
		if (ti) {
			bi = atomicAdd(items_ctr, ti);
			be = atomicAdd(entries_ctr, te);
			if (bi + ti > out.cap || be + te > out.cap_entries)
				ctr->error = 1, bi = 0xFFFFFFFFu;
		}
		s_base[0] = bi, s_base[1] = be;
	}
	__syncthreads();
	if (!want || s_base[0] == 0xFFFFFFFFu)
		return;
	uint32_t slot = s_base[0] + s_wi[warp] + ii - n_items, entry = s_base[1] + s_we[warp] + ie - n_entries;
#pragma unroll
	for (int c = 0; c < 8; ++c) {
		if (!(want >> c & 1u))
			continue;
		out.cur[slot] = curc[c];
		out.pos[slot] = pack_pos((x << 1) | uint32_t(c & 1), (y << 1) | uint32_t((c >> 1) & 1), (z << 1) | uint32_t(c >> 2)); // NodeCoord.hpp:18-28
		out.parent[slot] = (item << 3) | uint32_t(c);
		out.list_off[slot] = entry;
		out.list_len[slot] = __popc(keep[c]);
		for (uint32_t kp = keep[c]; kp; kp &= kp - 1u)
			out.lists[entry++] = list[__ffs(kp) - 1];
		in.child_new[size_t(item) * 8u + c] = kPending;
		++slot;
	}
}

constexpr uint32_t kLeafItems = 128; // items (threads) per CTA of k_down_leaf: up to 1 024 proceeding leaves in shared memory
__global__ void __launch_bounds__(kLeafItems) k_down_leaf(Geometry g, const uint32_t *__restrict__ words,
                                                          const hd_edit_desc *__restrict__ edits, const uint32_t *__restrict__ filled,
                                                          LevelView in, LevelView out, DevCounters *ctr) {
	__shared__ uint32_t s_alloc[2 * (kMaxWarps + 1)];
	__shared__ uint32_t s_pair[kLeafItems * 8], s_cur[kLeafItems * 8], s_keep[kLeafItems * 8], s_warp[kMaxWarps + 1];
	const uint32_t level = g.node_levels - 2u, n = in.count(), filled_leaf = filled[g.node_levels - 1u];
	const uint32_t bits = g.voxel_level() - (level + 1u);
	const uint32_t item = blockIdx.x * kLeafItems + threadIdx.x;
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = kLeafItems >> 5;
	// ---- phase 1: classify the item's eight children against its edit list (classify_children) ----
	uint32_t curc[8], keep[8];
#pragma unroll
	for (int c = 0; c < 8; ++c)
		curc[c] = kNull, keep[c] = 0u;
	uint32_t x = 0, y = 0, z = 0, want = 0u; // want: children with surviving edits
	const uint32_t *list = nullptr;
	if (item < n) {
		const uint32_t cur = in.cur[item];
		if (cur != kNull) { // get_unpacked_node_array, NodePool.hpp:278-309
			const uint32_t mask = words[cur];
			uint32_t k = 1u;
#pragma unroll
			for (int c = 0; c < 8; ++c)
				if (mask >> c & 1u)
					curc[c] = words[cur + k++];
		}
		unpack_pos(in.pos[item], x, y, z);
		list = in.lists + in.list_off[item];
		classify_children(edits, list, in.list_len[item], bits, x, y, z, filled_leaf, curc, keep);
#pragma unroll
		for (int c = 0; c < 8; ++c) {
			if (keep[c])
				want |= 1u << c;
			else
				in.child_new[size_t(item) * 8u + c] = curc[c];
		}
	}
	// ---- compaction of the proceeding children of the CTA ----
	const uint32_t mine = __popc(want);
	uint32_t incl = mine;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
		if (lane >= uint32_t(d))
			incl += v;
	}
	if (lane == 31u)
		s_warp[warp] = incl;
	__syncthreads();
	uint32_t before = 0u, n_work = 0u;
	for (uint32_t w = 0; w < nwarps; ++w) {
		const uint32_t k = s_warp[w];
		before += w < warp ? k : 0u, n_work += k;
	}
	{
		uint32_t at = before + incl - mine;
#pragma unroll
		for (int c = 0; c < 8; ++c)
			if (want >> c & 1u)
				s_pair[at] = (item << 3) | uint32_t(c), s_cur[at] = curc[c], s_keep[at] = keep[c], ++at;
	}
	__syncthreads();
	if (threadIdx.x == 0 && n_work)
		atomicAdd(&ctr->stats[1], (unsigned long long)n_work);
	// ---- phase 2: one proceeding leaf per thread and trip (EditVoxel, NodePool.hpp:319-343) ----
	for (uint32_t k0 = 0; k0 < n_work; k0 += kLeafItems) { // uniform trip count: alloc_item synchronises the CTA
		const uint32_t k = k0 + threadIdx.x;
		bool changed = false;
		uint32_t n0 = 0u, n1 = 0u, base_ptr = kNull, pair = 0u;
		if (k < n_work) {
			pair = s_pair[k], base_ptr = s_cur[k];
			const uint32_t it = pair >> 3, c = pair & 7u;
			uint32_t px, py, pz;
			unpack_pos(in.pos[it], px, py, pz);
			px = (px << 1) | (c & 1u), py = (py << 1) | ((c >> 1) & 1u), pz = (pz << 1) | ((c >> 2) & 1u); // NodeCoord.hpp:18-28
			const uint32_t *lst = in.lists + in.list_off[it];
			uint32_t w0 = 0u, w1 = 0u;
			if (base_ptr != kNull) {
				const uint2 w = *reinterpret_cast<const uint2 *>(words + base_ptr);
				w0 = w.x, w1 = w.y;
			}
			n0 = w0, n1 = w1;
			for (uint32_t kp = s_keep[k]; kp; kp &= kp - 1u)
				leaf_apply(edits[lst[__ffs(kp) - 1]], px << 2, py << 2, pz << 2, n0, n1);
			if (n0 == w0 && n1 == w1)
				in.child_new[pair] = base_ptr; // unchanged (edit_leaf tail, NodePool.hpp:336-342)
			else if ((n0 | n1) == 0u)
				in.child_new[pair] = kNull;
			else
				changed = true;
		}
		uint32_t slot, entry_off;
		if (alloc_item(ctr, &ctr->next_items, &ctr->next_entries, changed, 0u, out.cap, out.cap_entries, slot, entry_off, s_alloc)) {
			out.cur[slot] = base_ptr; // the bucket-full fallback of upsert_leaf (NodePool.hpp:195)
			out.parent[slot] = pair;  // (item << 3) | child
			out.state[slot] = 1; // every leaf of the compact level is a candidate
			*reinterpret_cast<uint2 *>(out.cand + size_t(slot) * 2u) = make_uint2(n0, n1);
			in.child_new[pair] = kPending;
		}
	}
}

// Bottom-up re-pack of inner items (edit_node tail, NodePool.hpp:384-395).  Thread per item.
__device__ __forceinline__ void assemble_item(const uint32_t *__restrict__ words, const LevelView &lv, uint32_t item) {
	const uint32_t cur = lv.cur[item];
	uint32_t old_mask = 0;
	if (cur != kNull)
		old_mask = words[cur];
	uint32_t packed[9], k = 1, mask = 0, oi = 1;
	bool changed = false;
	for (uint32_t c = 0; c < 8; ++c) {
		uint32_t oldc = kNull;
		if (old_mask >> c & 1u)
			oldc = words[cur + oi++];
		const uint32_t newc = lv.child_new[size_t(item) * 8u + c];
		changed |= newc != oldc;
		if (newc != kNull) {
			mask |= 1u << c;
			packed[k++] = newc;
		}
	}
	packed[0] = mask;
	uint8_t st = 0;
	uint32_t res = cur;
	if (changed) {
		if (mask == 0)
			res = kNull;
		else {
			st = 1;
			for (uint32_t i = 0; i < k; ++i)
				lv.cand[size_t(item) * 9u + i] = packed[i];
		}
	}
	lv.state[item] = st;
	lv.result[item] = res;
}
__global__ void __launch_bounds__(kBlock) k_assemble(const uint32_t *__restrict__ words, LevelView lv) {
	const uint32_t n = lv.count();
	for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < n; item += gridDim.x * blockDim.x)
		assemble_item(words, lv, item);
}

// In-batch dedup: one winner per distinct candidate content.  Thread per item.
// bkt / count (optional): a winner also works out its bucket and counts itself there — what k_bucket_count does in a pass
// of its own over the candidates' words, which this thread already holds.
__global__ void __launch_bounds__(kBlock) k_dedup(uint32_t n, uint32_t stride, bool is_leaf,
                                                  const uint32_t *__restrict__ cand, uint8_t *state, uint32_t *winner,
                                                  uint32_t *table, uint32_t table_mask, const uint32_t *n_dev,
                                                  const uint32_t *err_dev, uint32_t *bkt = nullptr, uint32_t *count = nullptr,
                                                  uint32_t bucket_mask = 0u) {
	if (n_dev)
		n = *err_dev ? 0u : *n_dev;
	for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < n; item += gridDim.x * blockDim.x) {
		if (state[item] == 0)
			continue;
		const uint32_t *me = cand + size_t(item) * stride;
		const uint32_t nw = is_leaf ? 2u : 1u + __popc(me[0] & 0xFFu);
		uint64_t h = 0x9e3779b97f4a7c15ull;
		for (uint32_t i = 0; i < nw; ++i)
			h = mix64(h ^ me[i]) + i;
		uint32_t slot = uint32_t(h >> 20) & table_mask;
		for (;;) {
			uint32_t v = table[slot];
			if (v == 0u) {
				v = atomicCAS(&table[slot], 0u, item + 1u);
				if (v == 0u) {
					state[item] = 2; // winner
					if (bkt) { // NodePool.hpp:163-164
						const uint32_t b = (is_leaf ? hash_leaf(me[0], me[1]) : hash_inner(me, nw)) & bucket_mask;
						bkt[item] = b;
						atomicAdd(&count[b], 1u);
					}
					break;
				}
			}
			const uint32_t *other = cand + size_t(v - 1u) * stride;
			bool same = true;
			for (uint32_t i = 0; i < nw && same; ++i)
				same = other[i] == me[i];
			if (same) {
				winner[item] = v - 1u;
				break;
			}
			slot = (slot + 1u) & table_mask;
		}
	}
}

// Find-or-insert of every winner: warps stride over the items (persistent grid), one warp per item at a time.
// find_node + append_node, NodePool.hpp:79-157,159-212.
__global__ void __launch_bounds__(kBlock) k_upsert(Geometry g, uint32_t level, bool fast_scan, uint32_t n,
                                                   uint32_t stride, const uint32_t *__restrict__ cand,
                                                   const uint8_t *__restrict__ state, const uint32_t *__restrict__ fallback,
                                                   uint32_t *result, uint32_t *words, uint32_t *bucket_words,
                                                   DevCounters *ctr, const uint32_t *n_dev) {
	if (n_dev) { // low-latency path: the count lives on the device; a scratch overflow upstream must not touch the pool
		if (*reinterpret_cast<volatile uint32_t *>(&ctr->error))
			return;
		n = *n_dev;
	}
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	const bool is_leaf = level == g.node_levels - 1u;
	const uint32_t wpp = g.words_per_page(), wpb = g.words_per_bucket();
	unsigned long long st_upserts = 0, st_nodes = 0, st_words = 0, st_overflow = 0, st_scan = 0; // lane 0 only

	for (uint32_t item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < n; item += warps) {
		if (state[item] != 2)
			continue;
		const uint32_t *me = cand + size_t(item) * stride;
		const uint32_t c0 = me[0], c1 = me[1];
		const uint32_t nw = is_leaf ? 2u : 1u + __popc(c0 & 0xFFu);
		const uint32_t h = is_leaf ? hash_leaf(c0, c1) : hash_inner(me, nw);
		const uint32_t bucket = g.level_base[level] + (h & ((1u << g.bucket_bits[level]) - 1u)); // NodePool.hpp:163-164
		const uint32_t base = bucket << g.bucket_shift();
		const uint32_t bw = *reinterpret_cast<volatile uint32_t *>(bucket_words + bucket);

		uint32_t found = kNull;
		if (is_leaf) {
			// leaves are 2-word aligned: each lane compares one pair per 64-word chunk; 4 chunks are loaded before any
			// is tested so every lane keeps 4 independent 8-byte loads in flight (the scan is latency-bound otherwise)
			for (uint32_t off = 0; off < bw && found == kNull; off += 256u) {
				uint2 w[4];
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					const uint32_t p = off + k * 64u + lane * 2u;
					w[k] = p + 2u <= bw ? *reinterpret_cast<const uint2 *>(words + base + p) : make_uint2(0u, 0u);
				}
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					const uint32_t m = __ballot_sync(0xFFFFFFFFu, w[k].x == c0 && w[k].y == c1);
					if (m && found == kNull)
						found = base + off + k * 64u + (__ffs(m) - 1u) * 2u;
				}
			}
		} else if (fast_scan) {
			// A word <= 0xFF inside a used bucket region is always a node header (child pointers are >= 256
			// here), so every position can be tested independently: header match, then the children.
			// 4 x 128 B are loaded per lane-step before testing (4 independent loads in flight).
			for (uint32_t off = 0; off < bw && found == kNull; off += 128u) {
				uint32_t w[4];
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					const uint32_t p = off + k * 32u + lane;
					w[k] = p + nw <= bw ? words[base + p] : 0u;
				}
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					const uint32_t p = off + k * 32u + lane;
					bool hit = w[k] == c0; // c0 != 0, so out-of-range slots (0) never match
					if (hit)
						for (uint32_t i = 1; i < nw && hit; ++i)
							hit = words[base + p + i] == me[i];
					const uint32_t m = __ballot_sync(0xFFFFFFFFu, hit);
					if (m && found == kNull)
						found = base + off + k * 32u + __ffs(m) - 1u;
				}
			}
		} else {
			if (lane == 0) { // tiny configs where a child pointer may look like a header: sequential walk
				for (uint32_t page = 0; page < bw && found == kNull; page += wpp) {
					const uint32_t end = min(page + wpp, bw);
					for (uint32_t it = page; nw <= end - it;) {
						const uint32_t hw = words[base + it] & 0xFFu;
						if (hw == 0u)
							break;
						const uint32_t sz = 1u + __popc(hw);
						bool same = sz == nw;
						for (uint32_t i = 0; i < nw && same; ++i)
							same = words[base + it + i] == me[i];
						if (same) {
							found = base + it;
							break;
						}
						it += sz;
					}
				}
			}
			found = __shfl_sync(0xFFFFFFFFu, found, 0);
		}

		if (lane == 0) {
			if (found == kNull) {
				// append_node: lock-free reservation; a node never straddles a page, the skipped tail stays zero
				uint32_t old = bw;
				for (;;) {
					const uint32_t off = old & (wpp - 1u);
					const uint32_t at = off + nw > wpp ? (old | (wpp - 1u)) + 1u : old;
					if (at + nw > wpb) { // bucket full: NodePool.hpp:137-139,195 -> keep the old node
						found = fallback ? fallback[item] : kNull;
						++st_overflow;
						break;
					}
					const uint32_t prev = atomicCAS(bucket_words + bucket, old, at + nw);
					if (prev == old) {
						// Concurrent scanners of this bucket may observe the node half-written.  For inner nodes that
						// is harmless (an unwritten child slot is 0 and no child pointer is 0), but a leaf word CAN
						// legitimately be 0: a half-written leaf (z0, 0) would equal the valid leaf (z0, 0) of another
						// candidate.  Leaves are therefore published with ONE 64-bit store (8-byte aligned, atomic):
						// a scanner sees (0, 0) — never a stored leaf — or the whole leaf.
						if (is_leaf)
							*reinterpret_cast<uint2 *>(words + base + at) = make_uint2(c0, c1);
						else
							for (uint32_t i = 0; i < nw; ++i)
								words[base + at + i] = me[i];
						found = base + at;
						++st_nodes;
						st_words += at + nw - old;
						break;
					}
					old = prev;
				}
			}
			result[item] = found;
			++st_upserts;
			st_scan += bw;
		}
		__syncwarp();
	}
	if (lane == 0 && st_upserts) {
		atomicAdd(&ctr->stats[2], st_upserts);
		atomicAdd(&ctr->stats[7], st_scan);
		if (st_nodes) {
			atomicAdd(&ctr->stats[3], st_nodes);
			atomicAdd(&ctr->stats[4], st_words);
		}
		if (st_overflow)
			atomicAdd(&ctr->stats[5], st_overflow);
	}
}

// ---- low-latency path: fused bottom-up step ------------------------------------------------------------------------
// One warp per work item does what k_leaf / k_assemble + k_dedup + k_upsert + k_resolve do in four launches: build the
// item's new node, find-or-insert it, report the pointer to the parent.  Without a dedup phase two warps may carry equal
// contents, so the insert is the reference's own threaded protocol (upsert_node<true>, NodePool.hpp:172-211): lock-free
// find over [0, bucket_words), take the bucket's lock, re-find over what was appended meanwhile, append, publish
// bucket_words, unlock.  Scans read through L2 (__ldcg: another SM's append is never hidden by a stale L1 line) and a
// writer fences before it publishes bucket_words, so every word below a bucket_words value a scanner read is complete.
// kLoads = 4-byte loads per lane and trip.  Measured on the one-launch kernel: 16 (512 words per trip) instead of 4 takes 5 % off
// a 100-editor batch (level 14 of the cfg3 scene, 83 000 items x 1 200 words: 256 -> 220 us) and nothing off a single brush.
// Dropped: verifying header matches one candidate at a time with the whole warp instead of every lane verifying its own —
// the header is just the 8-bit child mask, dozens of nodes of a bucket share it, and the verifications become a chain of
// dependent round trips (the same level 263 -> 402 us).
template <int kLoads>
__device__ __forceinline__ uint32_t warp_find_t(const uint32_t *words, uint32_t base, uint32_t from, uint32_t to,
                                                const uint32_t *me /* shared */, uint32_t nw, bool is_leaf, bool fast_scan,
                                                uint32_t wpp) {
	const uint32_t lane = threadIdx.x & 31u, full = 0xFFFFFFFFu;
	const uint32_t c0 = me[0], c1 = me[1];
	uint32_t found = kNull;
	if (is_leaf) {
		for (uint32_t off = from & ~1u; off < to && found == kNull; off += kLoads * 32u) {
			uint2 w[kLoads / 2];
#pragma unroll
			for (int k = 0; k < kLoads / 2; ++k) {
				const uint32_t q = off + k * 64u + lane * 2u;
				w[k] = q + 2u <= to ? __ldcg(reinterpret_cast<const uint2 *>(words + base + q)) : make_uint2(0u, 0u);
			}
#pragma unroll
			for (int k = 0; k < kLoads / 2; ++k) {
				const uint32_t m = __ballot_sync(full, w[k].x == c0 && w[k].y == c1);
				if (m && found == kNull)
					found = base + off + k * 64u + (__ffs(m) - 1u) * 2u;
			}
		}
	} else if (fast_scan) {
		for (uint32_t off = from; off < to && found == kNull; off += kLoads * 32u) {
			uint32_t w[kLoads];
#pragma unroll
			for (int k = 0; k < kLoads; ++k) {
				const uint32_t q = off + k * 32u + lane;
				w[k] = q + nw <= to ? __ldcg(words + base + q) : 0u;
			}
#pragma unroll
			for (int k = 0; k < kLoads; ++k) {
				const uint32_t q = off + k * 32u + lane;
				bool hit = w[k] == c0; // c0 != 0: out-of-range slots (0) never match
				if (hit) { // header match (rare unless it is the node): all child words in flight at once, then compare
					uint32_t v[8];
#pragma unroll
					for (uint32_t i = 1; i <= 8u; ++i)
						v[i - 1] = i < nw ? __ldcg(words + base + q + i) : 0u;
#pragma unroll
					for (uint32_t i = 1; i <= 8u; ++i)
						hit = hit && (i >= nw || v[i - 1] == me[i]);
				}
				const uint32_t m = __ballot_sync(full, hit);
				if (m && found == kNull)
					found = base + off + k * 32u + __ffs(m) - 1u;
			}
		}
	} else {
		if (lane == 0) { // tiny configs (a child pointer may look like a header): walk node by node, page by page
			for (uint32_t it = from; it < to && found == kNull;) {
				const uint32_t end = min((it | (wpp - 1u)) + 1u, to);
				while (nw <= end - it) {
					const uint32_t hw = __ldcg(words + base + it) & 0xFFu;
					if (hw == 0u)
						break;
					const uint32_t sz = 1u + __popc(hw);
					bool same = sz == nw;
					for (uint32_t i = 0; i < nw && same; ++i)
						same = __ldcg(words + base + it + i) == me[i];
					if (same) {
						found = base + it;
						break;
					}
					it += sz;
				}
				it = end;
			}
		}
		found = __shfl_sync(full, found, 0);
	}
	return found;
}

// 512 words per trip in the one-launch kernel (FusedArgs::wide_scan; 5 % on a 100-editor batch), 128 elsewhere
__device__ __forceinline__ uint32_t warp_find(const uint32_t *words, uint32_t base, uint32_t from, uint32_t to,
                                              const uint32_t *me /* shared */, uint32_t nw, bool is_leaf, bool fast_scan,
                                              uint32_t wpp, bool wide = false) {
	return wide && to - from > 128u ? warp_find_t<16>(words, base, from, to, me, nw, is_leaf, fast_scan, wpp)
	                                : warp_find_t<4>(words, base, from, to, me, nw, is_leaf, fast_scan, wpp);
}
struct UpStats {
	uint32_t upserts, nodes, words, overflow;
	unsigned long long scan;
};
__device__ __forceinline__ uint32_t warp_upsert(const Geometry &g, uint32_t level, bool fast_scan, const uint32_t *me,
                                                uint32_t nw, uint32_t fallback, uint32_t *words, uint32_t *bucket_words,
                                                uint32_t *locks, UpStats &st, uint32_t lock_mask = 0xFFFFFFFFu, bool wide = false) {
	const uint32_t lane = threadIdx.x & 31u, full = 0xFFFFFFFFu;
	const bool is_leaf = level == g.node_levels - 1u;
	const uint32_t wpp = g.words_per_page(), wpb = g.words_per_bucket();
	const uint32_t h = is_leaf ? hash_leaf(me[0], me[1]) : hash_inner(me, nw);
	const uint32_t bucket = g.level_base[level] + (h & ((1u << g.bucket_bits[level]) - 1u)); // NodePool.hpp:163-164
	const uint32_t base = bucket << g.bucket_shift();
	volatile uint32_t *bwp = bucket_words + bucket;
	const uint32_t bw = *bwp;
	uint32_t found = warp_find(words, base, 0u, bw, me, nw, is_leaf, fast_scan, wpp, wide);
	st.upserts += 1, st.scan += bw;
	if (found != kNull)
		return found;
	if (lane == 0) {
		while (atomicCAS(locks + (bucket & lock_mask), 0u, 1u) != 0u)
			__nanosleep(32);
		__threadfence();
	}
	__syncwarp(full);
	const uint32_t bw2 = *bwp;
	if (bw2 > bw) { // somebody appended between the scan and the lock: look at the new tail only (NodePool.hpp:187-192)
		found = warp_find(words, base, bw, bw2, me, nw, is_leaf, fast_scan, wpp, wide);
		st.scan += bw2 - bw;
	}
	if (found == kNull) {
		const uint32_t off = bw2 & (wpp - 1u);
		const uint32_t at = off + nw > wpp ? (bw2 | (wpp - 1u)) + 1u : bw2; // never straddle a page; the tail stays zero
		if (at + nw > wpb) { // bucket full: keep the old node (NodePool.hpp:137-139,195)
			found = fallback;
			st.overflow += 1;
		} else {
			if (lane < nw)
				words[base + at + lane] = me[lane];
			__threadfence();
			__syncwarp(full);
			if (lane == 0)
				*bwp = at + nw;
			found = base + at;
			st.nodes += 1, st.words += at + nw - bw2;
		}
	}
	__syncwarp(full);
	if (lane == 0) {
		__threadfence();
		atomicExch(locks + (bucket & lock_mask), 0u);
	}
	return found;
}

// One bottom-up level of the low-latency path.  Leaf level: the item's 64 voxels (k_leaf); inner levels: re-pack from
// the children's reports (k_assemble).  Then find-or-insert and report to the parent's child slot.
__device__ __forceinline__ void phase_up(const Geometry &g, uint32_t level, bool fast_scan, uint32_t *words,
                                         uint32_t *bucket_words, uint32_t *locks, const hd_edit_desc *__restrict__ edits,
                                         const LevelView &lv, uint32_t *parent_child_new, DevCounters *ctr,
                                         uint32_t (*s_cand)[12], uint32_t tid0, uint32_t nthreads,
                                         uint32_t lock_mask = 0xFFFFFFFFu, bool wide = false) {
	// lock_mask: `locks` holds one word per bucket (all ones), or — when one CTA finishes the small levels alone and nobody
	// else touches the pool — a few words of shared memory indexed by the bucket's low bits (no global atomics)
	const uint32_t lane = threadIdx.x & 31u, full = 0xFFFFFFFFu, n = lv.count();
	const uint32_t warps = nthreads >> 5;
	const bool is_leaf = level == g.node_levels - 1u;
	uint32_t *me = s_cand[threadIdx.x >> 5];
	UpStats st{0u, 0u, 0u, 0u, 0ull};
	for (uint32_t item = tid0 >> 5; item < n; item += warps) {
		const uint32_t cur = lv.cur[item];
		uint32_t res = cur, nw = 0;
		bool insert = false;
		if (is_leaf) {
			const uint32_t *list = lv.lists + lv.list_off[item];
			const uint32_t len = lv.list_len[item];
			uint32_t x, y, z;
			unpack_pos(lv.pos[item], x, y, z);
			uint32_t w0 = 0, w1 = 0;
			if (cur != kNull) {
				const uint2 w = *reinterpret_cast<const uint2 *>(words + cur);
				w0 = w.x, w1 = w.y;
			}
			const uint32_t vx = (x << 2) | ((lane >> 2) & 2u) | (lane & 1u);
			const uint32_t vy = (y << 2) | ((lane >> 3) & 2u) | ((lane >> 1) & 1u);
			const uint32_t vz = (z << 2) | ((lane >> 2) & 1u);
			bool a = w0 >> lane & 1u, b = w1 >> lane & 1u;
			for (uint32_t j = 0; j < len; ++j) {
				const hd_edit_desc &e = edits[list[j]];
				edit_voxel_pair<false>(e, vx, vy, vz, a, b);
			}
			const uint32_t n0 = __ballot_sync(full, a), n1 = __ballot_sync(full, b);
			if (n0 != w0 || n1 != w1) {
				if ((n0 | n1) == 0u)
					res = kNull;
				else {
					insert = true, nw = 2;
					if (lane == 0)
						me[0] = n0, me[1] = n1;
				}
			}
		} else {
			uint32_t newc = kNull, oldc = kNull;
			if (lane < 8u) {
				newc = lv.child_new[size_t(item) * 8u + lane];
				if (cur != kNull) {
					const uint32_t old_mask = words[cur];
					if (old_mask >> lane & 1u)
						oldc = words[cur + 1u + __popc(old_mask & ((1u << lane) - 1u))];
				}
			}
			const uint32_t changed = __ballot_sync(full, newc != oldc);
			const uint32_t mask = __ballot_sync(full, newc != kNull);
			if (changed) {
				if (mask == 0u)
					res = kNull;
				else {
					insert = true, nw = 1u + __popc(mask);
					if (lane == 0)
						me[0] = mask;
					if (newc != kNull)
						me[1u + __popc(mask & ((1u << lane) - 1u))] = newc;
				}
			}
		}
		if (insert) {
			__syncwarp(full);
			res = warp_upsert(g, level, fast_scan, me, nw, cur, words, bucket_words, locks, st, lock_mask, wide);
			__syncwarp(full);
		}
		if (lane == 0) {
			const uint32_t par = lv.parent[item];
			if (par == 0xFFFFFFFFu)
				ctr->root_out = res;
			else
				parent_child_new[par] = res;
		}
	}
	if (lane == 0 && st.upserts) {
		atomicAdd(&ctr->stats[2], (unsigned long long)st.upserts);
		atomicAdd(&ctr->stats[7], st.scan);
		if (st.nodes) {
			atomicAdd(&ctr->stats[3], (unsigned long long)st.nodes);
			atomicAdd(&ctr->stats[4], (unsigned long long)st.words);
		}
		if (st.overflow)
			atomicAdd(&ctr->stats[5], (unsigned long long)st.overflow);
	}
}
__global__ void __launch_bounds__(kBlock) k_up(Geometry g, uint32_t level, bool fast_scan, uint32_t *words,
                                               uint32_t *bucket_words, uint32_t *locks,
                                               const hd_edit_desc *__restrict__ edits, LevelView lv,
                                               uint32_t *parent_child_new, DevCounters *ctr) {
	__shared__ uint32_t s_cand[kBlock / 32][12];
	phase_up(g, level, fast_scan, words, bucket_words, locks, edits, lv, parent_child_new, ctr, s_cand,
	         blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// The whole rebuild of a small batch in ONE cooperative launch: grid-wide barriers separate the levels, and the levels
// that hold only a handful of items (every level above the brush's footprint: 10-12 of 16 at 2^17) are walked by CTA 0
// alone with __syncthreads between them, so they cost a block barrier instead of a launch or a grid barrier each.
// Call parameters come in through mapped host memory and the counters go back the same way: no copy commands at all.
constexpr int kFusedThreads = 512;
constexpr uint32_t kSoloDown = 128; // items one CTA expands alone (8 threads per item)
constexpr uint32_t kSoloUp = 32;    // items one CTA finishes alone (one warp per item)
struct FusedArgs {
	Geometry g;
	uint32_t *words, *bucket_words, *locks;
	const uint32_t *iota, *filled;
	const FastDyn *dyn_host; // mapped pinned host memory
	FastDyn *dyn_dev;
	DevCounters *ctr, *ctr_host; // ctr_host: mapped pinned host memory
	LevelView lv[HD_MAX_NODE_LEVELS];
	bool fast_scan;
	bool wide_scan; // 512-word scan trips in the bottom-up phases (default; HD_EDIT_SCAN_WIDE=0: 128 words, for A/B runs)
};
__global__ void __launch_bounds__(kFusedThreads) k_edit_fused(const __grid_constant__ FusedArgs a) {
	__shared__ uint32_t s_cand[kFusedThreads / 32][12];
	__shared__ uint32_t s_alloc[2 * (kMaxWarps + 1)];
	__shared__ uint32_t s_first_big;
	// the levels CTA 0 walks alone keep what every trip needs on chip: the queue counters of the level being produced, the
	// editors' descriptors and the bucket locks (each of them was a dependent L2 round trip per level)
	__shared__ uint32_t s_ctr[2], s_locks[64];
	__shared__ hd_edit_desc s_edits[32];
	cg::grid_group grid = cg::this_grid();
	const Geometry &g = a.g;
	const uint32_t L = g.node_levels;
	const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
	DevCounters *ctr = a.ctr;
	const hd_edit_desc *edits = a.dyn_dev->edits;
	volatile uint32_t *items = ctr->lvl_items;

	// ---- stage A (CTA 0): fetch the call, classify the root, walk the small top levels ----
	if (blockIdx.x == 0) {
		for (uint32_t i = threadIdx.x; i < sizeof(DevCounters) / 4; i += blockDim.x)
			reinterpret_cast<uint32_t *>(ctr)[i] = 0u;
		// only the header and the descriptors in use cross PCIe (a brush call is 56 bytes, not the 40 KB the struct can hold)
		const uint32_t n_edits = min(reinterpret_cast<const volatile uint32_t *>(a.dyn_host)[1], kFastMaxEdits);
		const uint32_t n_words = 4u + n_edits * uint32_t(sizeof(hd_edit_desc) / 4);
		for (uint32_t i = threadIdx.x; i < n_words; i += blockDim.x) { // up to 32 editors also stay in shared memory
			const uint32_t w = reinterpret_cast<const volatile uint32_t *>(a.dyn_host)[i];
			reinterpret_cast<uint32_t *>(a.dyn_dev)[i] = w;
			if (i >= 4u && n_edits <= 32u)
				reinterpret_cast<uint32_t *>(s_edits)[i - 4u] = w;
		}
		if (threadIdx.x < 64)
			s_locks[threadIdx.x] = 0u;
		__syncthreads();
		const hd_edit_desc *edits_solo = n_edits <= 32u ? s_edits : edits;
		if (threadIdx.x == 0)
			phase_stamp(ctr);
		if (threadIdx.x < 32)
			phase_root(g, edits_solo, a.dyn_dev->n_edits, a.iota, a.filled, a.dyn_dev->root, a.lv[0], ctr);
		__syncthreads();
		uint32_t l = 0, n_cur = items[0];
		// (levels whose lists are longer than 32 entries go to the grid: a warp per (item, child) pair; only the root's
		// flag can be set here, phase_down_long is what sets the others)
		if (*(volatile uint32_t *)&ctr->lvl_long[0] == 0u)
			for (; l + 1 < L && n_cur != 0u && n_cur <= kSoloDown; ++l) {
				if (threadIdx.x == 0)
					s_ctr[0] = 0u, s_ctr[1] = 0u;
				__syncthreads();
				phase_down<false, true>(g, l, a.words, edits_solo, a.filled, a.lv[l], a.lv[l + 1], ctr, &s_ctr[0], &s_ctr[1],
				                        threadIdx.x, blockDim.x, s_alloc, n_cur);
				__syncthreads();
				const uint32_t n_next = s_ctr[0], e_next = s_ctr[1];
				if (threadIdx.x == 0)
					ctr->lvl_items[l + 1] = n_next, ctr->lvl_entries[l + 1] = e_next;
				__syncthreads(); // s_ctr is reset by the next trip
				n_cur = n_next;
				if (n_next > a.lv[l + 1].cap || e_next > a.lv[l + 1].cap_entries) { // queue overflow: DevCounters::error is set
					++l;
					break;
				}
			}
		__threadfence(); // the grid reads the queues and counts written above
		if (threadIdx.x == 0)
			ctr->next_items = l, phase_stamp(ctr); // first level the whole grid expands
	}
	grid.sync();
	// ---- stage B (grid): the remaining top-down levels ----
	for (uint32_t l = *(volatile uint32_t *)&ctr->next_items; l + 1 < L; ++l) {
		phase_down<false, true>(g, l, a.words, edits, a.filled, a.lv[l], a.lv[l + 1], ctr, &ctr->lvl_items[l + 1],
		                        &ctr->lvl_entries[l + 1], gtid, gthreads, s_alloc);
		if (*(volatile uint32_t *)&ctr->lvl_long[l])
			phase_down_long(g, l, a.words, edits, a.filled, a.lv[l], a.lv[l + 1], ctr, &ctr->lvl_items[l + 1],
			                &ctr->lvl_entries[l + 1], &ctr->lvl_long[l + 1], gtid >> 5, gthreads >> 5);
		grid.sync();
		if (gtid == 0)
			phase_stamp(ctr);
	}
	// ---- stage C (grid): bottom-up over the levels that are worth the grid; then CTA 0 finishes the small top ----
	if (threadIdx.x == 0) {
		uint32_t solo = 0; // levels [0, solo) are small all the way up
		while (solo < L && items[solo] <= kSoloUp)
			++solo;
		s_first_big = solo;
	}
	__syncthreads();
	const uint32_t solo = s_first_big;
	for (uint32_t l = L; l-- > solo;) {
		phase_up(g, l, a.fast_scan, a.words, a.bucket_words, a.locks, edits, a.lv[l], l ? a.lv[l - 1].child_new : nullptr, ctr, s_cand,
		         gtid, gthreads, 0xFFFFFFFFu, a.wide_scan);
		grid.sync();
		if (gtid == 0)
			phase_stamp(ctr);
	}
	if (blockIdx.x != 0)
		return;
	{
		// every other CTA has left (or waits in nothing: there is no grid barrier below): shared-memory locks and descriptors
		const uint32_t n_edits = min(a.dyn_dev->n_edits, kFastMaxEdits);
		const hd_edit_desc *edits_solo = n_edits <= 32u ? s_edits : edits;
		for (uint32_t l = solo; l-- > 0;) {
			phase_up(g, l, a.fast_scan, a.words, a.bucket_words, s_locks, edits_solo, a.lv[l], l ? a.lv[l - 1].child_new : nullptr, ctr,
			         s_cand, threadIdx.x, blockDim.x, 63u, a.wide_scan);
			__syncthreads();
		}
	}
	if (threadIdx.x == 0)
		phase_stamp(ctr);
	__syncthreads();
	__threadfence();
	for (uint32_t i = threadIdx.x; i < sizeof(DevCounters) / 4; i += blockDim.x)
		reinterpret_cast<volatile uint32_t *>(a.ctr_host)[i] = __ldcg(reinterpret_cast<const uint32_t *>(ctr) + i);
	__threadfence_system();
}

// Losers copy their winner's pointer; every item reports to its parent's child slot (or the root output).
__global__ void __launch_bounds__(kBlock) k_resolve(LevelView lv, uint32_t *parent_child_new, DevCounters *ctr) {
	const uint32_t n = lv.count();
	for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < n; item += gridDim.x * blockDim.x) {
		uint32_t res = lv.result[item];
		if (lv.state[item] == 1)
			res = lv.result[lv.winner[item]];
		const uint32_t par = lv.parent[item];
		if (par == 0xFFFFFFFFu)
			ctr->root_out = res;
		else
			parent_child_new[par] = res;
	}
}

__global__ void k_resolve_flat(uint32_t n, const uint8_t *__restrict__ state, const uint32_t *__restrict__ winner,
                               uint32_t *result) {
	const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
	if (item < n && state[item] == 1)
		result[item] = result[winner[item]];
}

__global__ void k_iota(uint32_t *out, uint32_t n) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
		out[i] = i;
}
__global__ void k_fill_u8(uint8_t *out, uint32_t n, uint8_t v) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
		out[i] = v;
}

// ------------------------------------------------------------------------------------------------ host side
static inline uint32_t grid_for(uint64_t threads) { return uint32_t((threads + kBlock - 1) / kBlock); }
// persistent grid for warp-per-item kernels: a multiple of the SM count, capped by the work
static inline uint32_t upsert_grid(hd_pool *p, uint32_t n) {
	const int sms = std::max(p->sm_count, 1);
	const uint64_t need = (uint64_t(n) * 32 + kBlock - 1) / kBlock;
	return uint32_t(std::min<uint64_t>(need, uint64_t(sms) * 8));
}

template <typename T> static cudaError_t amalloc(T **p, uint64_t count, cudaStream_t s) {
	return cudaMallocAsync(reinterpret_cast<void **>(p), std::max<uint64_t>(count, 1) * sizeof(T), s);
}

struct LevelAlloc {
	LevelView v{};
	cudaStream_t s = nullptr;
	cudaError_t init(uint32_t cap, uint32_t cap_entries, bool leaf, cudaStream_t stream) {
		s = stream;
		v.cap = cap, v.cap_entries = cap_entries, v.n = 0;
		cudaError_t e;
		if ((e = amalloc(&v.cur, cap, s)) || (e = amalloc(&v.pos, cap, s)) || (e = amalloc(&v.list_off, cap, s)) ||
		    (e = amalloc(&v.list_len, cap, s)) || (e = amalloc(&v.parent, cap, s)) || (e = amalloc(&v.lists, cap_entries, s)))
			return e;
		(void)leaf;
		return cudaSuccess;
	}
	// the compact leaf level k_down_leaf fills: fallback pointer, parent slot and the new leaf words of CHANGED leaves only
	cudaError_t init_compact_leaves(uint32_t cap, cudaStream_t stream) {
		s = stream;
		v.cap = cap, v.cap_entries = 0, v.n = 0;
		cudaError_t e;
		if ((e = amalloc(&v.cur, cap, s)) || (e = amalloc(&v.parent, cap, s)) || (e = amalloc(&v.cand, uint64_t(cap) * 2, s)) ||
		    (e = amalloc(&v.state, cap, s)))
			return e;
		return cudaSuccess;
	}
	// arrays only needed once the item count is known
	cudaError_t init_up(bool leaf, bool compact = false) { // compact: cand and state already exist (init_compact_leaves)
		cudaError_t e;
		if ((e = amalloc(&v.result, v.n, s)) || (e = amalloc(&v.winner, v.n, s)))
			return e;
		if (!compact && ((e = amalloc(&v.state, v.n, s)) || (e = amalloc(&v.cand, uint64_t(v.n) * (leaf ? 2 : 9), s))))
			return e;
		if (!leaf && (e = amalloc(&v.child_new, uint64_t(v.n) * 8, s)))
			return e;
		return cudaSuccess;
	}
	void release() {
		void *ptrs[] = {v.cur, v.pos, v.list_off, v.list_len, v.parent, v.lists, v.result, v.state, v.winner, v.cand, v.child_new};
		for (void *p : ptrs)
			if (p)
				cudaFreeAsync(p, s);
		v = LevelView{};
	}
};

static hd_status scratch_init(hd_pool *p) {
	if (p->edit)
		return HD_OK;
	auto *s = new EditScratch();
	// p->edit is published only when every allocation succeeded: a half-initialised scratch would make the next call skip
	// this function and dereference null device pointers
	cudaError_t e = cudaMalloc(&s->ctr, sizeof(DevCounters));
	if (e == cudaSuccess)
		e = cudaMemset(s->ctr, 0, sizeof(DevCounters));
	if (e == cudaSuccess)
		e = cudaMalloc(&s->filled_dev, sizeof(uint32_t) * HD_MAX_NODE_LEVELS);
	if (e != cudaSuccess) {
		cudaFree(s->ctr), cudaFree(s->filled_dev);
		delete s;
		HD_CUDA_TRY(e);
	}
	// child pointers of any level >= 1 are >= (buckets at level 0) << bucket_shift
	s->fast_scan = (uint64_t(1) << (p->geo.bucket_bits[0] + p->geo.bucket_shift())) >= 256ull;
	p->edit = s;
	return HD_OK;
}

hd_status edit_scratch_free(hd_pool *p) {
	if (!p->edit)
		return HD_OK;
	cudaFree(p->edit->ctr);
	cudaFree(p->edit->filled_dev);
	FastPath &f = p->edit->fast;
	if (f.exec)
		cudaGraphExecDestroy(f.exec);
	if (f.graph)
		cudaGraphDestroy(f.graph);
	cudaFree(f.arena), cudaFree(f.dyn_dev), cudaFree(f.iota_dev);
	cudaFreeHost(f.dyn_host), cudaFreeHost(f.ctr_host);
	delete p->edit;
	p->edit = nullptr;
	return HD_OK;
}

// ---- bucket-grouped find-or-insert (large batches) ------------------------------------------------------------------
// With tens of candidates per bucket, k_upsert re-reads every bucket once per candidate (34 GB of scans for the cfg3
// batch).  Here the winners are counting-sorted by bucket and ONE CTA owns a bucket: it stages the used prefix in
// shared memory once, its warps test all candidates of the bucket against that image, and the misses are appended
// by the same CTA — no CAS, no other writer, nothing half-written for anybody to see.
constexpr uint32_t kMiss = 0xFFFFFFFDu;
constexpr int kGroupThreads = 128;
constexpr uint32_t kGroupMaxWords = 8192; // bucket images up to 32 KB are staged; larger buckets use k_upsert

__global__ void __launch_bounds__(kBlock) k_bucket_scatter(uint32_t n, const uint8_t *__restrict__ state,
                                                           const uint32_t *__restrict__ bkt, const uint32_t *__restrict__ offset,
                                                           uint32_t *fill, uint32_t *order) {
	const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
	if (item >= n || state[item] != 2)
		return;
	const uint32_t b = bkt[item];
	order[offset[b] + atomicAdd(&fill[b], 1u)] = item;
}

__device__ __forceinline__ uint32_t image_hash2(uint32_t w0, uint32_t w1) {
	uint32_t h = (w0 * 0x9E3779B1u) ^ (w1 * 0x85EBCA77u);
	return h ^ (h >> 15);
}
__device__ __forceinline__ uint32_t image_hashn(const uint32_t *w, uint32_t n) {
	uint32_t h = n;
	for (uint32_t i = 0; i < n; ++i)
		h = (h ^ w[i]) * 0x9E3779B1u;
	return h ^ (h >> 15);
}
__device__ __forceinline__ uint32_t image_hashn_reg(const uint32_t (&w)[9], uint32_t n) { // same value, registers only
	uint32_t h = n;
#pragma unroll
	for (uint32_t i = 0; i < 9u; ++i)
		if (i < n)
			h = (h ^ w[i]) * 0x9E3779B1u;
	return h ^ (h >> 15);
}

// The content index of a staged bucket holds 16-bit entries (node position + 1 <= kGroupMaxWords): the kernel's time follows the
// number of CTAs an SM can hold (measured: +4 KB of shared memory per CTA = 11 -> 9 resident = +0.9 ms on the cfg3 batch), and
// halving the table takes the CTA from 20.5 to 16.5 KB.  Insert = CAS on the 32-bit word that holds the half-word.
__device__ __forceinline__ uint32_t index_insert16(uint16_t *tab, uint32_t slot, uint32_t val) { // 0: inserted, else the occupant
	uint32_t *w = reinterpret_cast<uint32_t *>(tab) + (slot >> 1);
	const uint32_t sh = (slot & 1u) << 4;
	uint32_t old = *reinterpret_cast<volatile uint32_t *>(w);
	for (;;) {
		const uint32_t cur = (old >> sh) & 0xFFFFu;
		if (cur)
			return cur;
		const uint32_t seen = atomicCAS(w, old, old | (val << sh));
		if (seen == old)
			return 0u;
		old = seen;
	}
}

__global__ void __launch_bounds__(kGroupThreads) k_upsert_grouped(Geometry g, uint32_t level, bool fast_scan, uint32_t stride,
                                                                  const uint32_t *__restrict__ cand,
                                                                  const uint32_t *__restrict__ fallback, uint32_t *result,
                                                                  uint32_t *words, uint32_t *bucket_words,
                                                                  const uint32_t *__restrict__ offset,
                                                                  const uint32_t *__restrict__ count,
                                                                  const uint32_t *__restrict__ order, DevCounters *ctr) {
	extern __shared__ uint32_t img[]; // the bucket's used prefix
	const uint32_t b = blockIdx.x, cnt = count[b];
	if (cnt == 0)
		return;
	const uint32_t first = offset[b];
	const bool is_leaf = level == g.node_levels - 1u;
	const uint32_t bucket = g.level_base[level] + b, base = bucket << g.bucket_shift();
	const uint32_t bw = bucket_words[bucket], wpp = g.words_per_page(), wpb = g.words_per_bucket();
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	// open-addressing index over the image: 0 = empty, else node position + 1.  Only stored nodes are indexed (at most bw / 2
	// of them), so the table is the next power of two >= bw slots (<= 50 % full) instead of always wpb: with 2^20 buckets at
	// the 8^3-node level of the cfg3 batch a bucket holds ~600 words and a dozen candidates, and zeroing 2 048 slots word by
	// word was the longest phase of the CTA.  Staging and zeroing move 16 bytes per thread (the bucket base is 8 KB-aligned
	// and everything behind bw is zero).
	uint16_t *tab = reinterpret_cast<uint16_t *>(img + wpb); // wpb half-words
	uint32_t tsize = 64u;
	while (tsize < bw)
		tsize <<= 1;
	const uint32_t tmask = tsize - 1u;
	const bool indexed = is_leaf || fast_scan;
	// The first chunk's candidates are fetched BEFORE the image is staged: order -> candidate words is a chain of two
	// dependent global round trips that used to start only after the staging and the index build, and a CTA of the
	// 8^3-node level lives for a handful of such round trips (half a million CTAs with a dozen candidates each: ncu shows
	// the kernel waiting on barriers and loads, not issuing).
	// Order of issue: the candidate's index (order[]) first, then the staging loads, then — with the index back — ALL `stride`
	// words of the candidate without looking at its header (a load that waits for the header would stall the thread before
	// the next one is issued), and the index build runs in the shadow of those loads.
	uint32_t pre_item = 0, pre_w[9];
#pragma unroll
	for (uint32_t i = 0; i < 9u; ++i)
		pre_w[i] = 0u;
	const bool pre = indexed && threadIdx.x < cnt;
	if (pre)
		pre_item = order[first + threadIdx.x];
	{
		const uint4 *src = reinterpret_cast<const uint4 *>(words + base);
		uint4 *dst = reinterpret_cast<uint4 *>(img);
		for (uint32_t i = threadIdx.x; i < (bw + 3u) / 4u; i += kGroupThreads)
			dst[i] = src[i];
		if (pre) {
			const uint32_t *me = cand + size_t(pre_item) * stride;
#pragma unroll
			for (uint32_t i = 0; i < 9u; ++i)
				if (i < stride)
					pre_w[i] = me[i];
		}
		if (indexed) {
			uint4 *t4 = reinterpret_cast<uint4 *>(tab);
			for (uint32_t i = threadIdx.x; i < tsize / 8u; i += kGroupThreads) // tsize >= 64 half-words
				t4[i] = make_uint4(0u, 0u, 0u, 0u);
		}
	}
	__syncthreads();

	// phase A: every candidate of the bucket against the staged image (read-only).  A linear scan per candidate made
	// this kernel shared-memory-bandwidth-bound (72 candidates x 8 KB per bucket at the 8^3-node level of the cfg3 batch),
	// so the image is indexed once by content hash and every candidate costs one or two probes.
	if (indexed) {
		if (is_leaf) {
			for (uint32_t q = threadIdx.x * 2u; q + 2u <= bw; q += kGroupThreads * 2u) {
				const uint32_t w0 = img[q], w1 = img[q + 1u];
				if ((w0 | w1) == 0u)
					continue; // page-tail padding; an all-zero leaf is never stored (NodePool.hpp:337)
				uint32_t slot = image_hash2(w0, w1) & tmask;
				while (index_insert16(tab, slot, q + 1u) != 0u)
					slot = (slot + 1u) & tmask;
			}
		} else {
			// a word <= 0xFF inside the used region is always a node header (child pointers are >= 256 here, padding is 0).
			// One word in six is a header, so testing and hashing in the same loop ran the hash (the expensive part) at five
			// lanes of 32 — ncu: 10.7 lanes per instruction over the kernel.  The header positions are compacted first and
			// hashed by full warps.
			uint16_t *s_hdr = tab + wpb; // wpb / 2 entries of the dynamic allocation, behind the index
			__shared__ uint32_t s_nhdr;
			if (threadIdx.x == 0)
				s_nhdr = 0u;
			__syncthreads();
			for (uint32_t q0 = 0; q0 < bw; q0 += kGroupThreads) {
				const uint32_t q = q0 + threadIdx.x;
				bool hdr = false;
				if (q < bw) {
					const uint32_t hw = img[q];
					hdr = hw - 1u < 0xFFu && q + 1u + __popc(hw) <= bw;
				}
				const uint32_t votes = __ballot_sync(0xFFFFFFFFu, hdr);
				uint32_t at = 0u;
				if (lane == 0 && votes)
					at = atomicAdd(&s_nhdr, uint32_t(__popc(votes)));
				at = __shfl_sync(0xFFFFFFFFu, at, 0);
				if (hdr)
					s_hdr[at + __popc(votes & ((1u << lane) - 1u))] = uint16_t(q);
			}
			__syncthreads();
			const uint32_t nh = s_nhdr;
			for (uint32_t i = threadIdx.x; i < nh; i += kGroupThreads) {
				const uint32_t q = s_hdr[i], nw = 1u + __popc(img[q]);
				uint32_t slot = image_hashn(img + q, nw) & tmask;
				while (index_insert16(tab, slot, q + 1u) != 0u)
					slot = (slot + 1u) & tmask;
			}
		}
		__syncthreads();
		// Candidates in chunks of one per thread: (A) look up, (B1) thread 0 places the chunk's misses one after the other
		// from their sizes in shared memory (the no-page-straddle rule makes placement sequential, but it never touches
		// global memory), (B2) every thread writes its own node.  The former append loop — one warp, one miss at a time,
		// each with its own global round trips — was what this kernel spent its time in.
		__shared__ uint32_t s_nw[kGroupThreads], s_at[kGroupThreads], s_cur, s_app, s_ovf;
		if (threadIdx.x == 0)
			s_cur = bw, s_app = 0u, s_ovf = 0u;
		for (uint32_t c0 = 0; c0 < cnt; c0 += kGroupThreads) {
			const uint32_t c = c0 + threadIdx.x;
			uint32_t item = 0, nw = 0, found = kMiss;
			uint32_t w[9];
			if (c < cnt) {
				if (c0 == 0u) { // fetched while the image was staged; words behind the node's own are ignored
					item = pre_item;
					w[0] = pre_w[0], w[1] = pre_w[1];
					nw = is_leaf ? 2u : 1u + __popc(w[0] & 0xFFu);
#pragma unroll
					for (uint32_t i = 2; i < 9u; ++i)
						w[i] = i < nw ? pre_w[i] : 0u;
				} else {
					item = order[first + c];
					const uint32_t *me = cand + size_t(item) * stride;
					w[0] = me[0], w[1] = me[1];
					nw = is_leaf ? 2u : 1u + __popc(w[0] & 0xFFu);
#pragma unroll
					for (uint32_t i = 2; i < 9u; ++i)
						w[i] = i < nw ? me[i] : 0u;
				}
				uint32_t slot = (is_leaf ? image_hash2(w[0], w[1]) : image_hashn_reg(w, nw)) & tmask;
				for (uint32_t v; (v = tab[slot]) != 0u; slot = (slot + 1u) & tmask) {
					const uint32_t q = v - 1u;
					bool same = img[q] == w[0] && img[q + 1u] == w[1];
#pragma unroll
					for (uint32_t i = 2; i < 9u; ++i)
						same = same && (i >= nw || img[q + i] == w[i]);
					if (same) {
						found = base + q;
						break;
					}
				}
			}
			s_nw[threadIdx.x] = (c < cnt && found == kMiss) ? nw : 0u;
			__syncthreads();
			if (threadIdx.x == 0) { // append_node placement, NodePool.hpp:134-157, in list order
				uint32_t cur = s_cur, app = 0, ovf = 0;
				const uint32_t m = min(uint32_t(kGroupThreads), cnt - c0);
				for (uint32_t i = 0; i < m; ++i) {
					const uint32_t k = s_nw[i];
					if (k == 0u)
						continue;
					const uint32_t off = cur & (wpp - 1u);
					const uint32_t at = off + k > wpp ? (cur | (wpp - 1u)) + 1u : cur; // never straddle a page; the tail stays zero
					if (at + k > wpb) { // bucket full: keep the old node (NodePool.hpp:137-139,195)
						s_at[i] = kMiss;
						++ovf;
						continue;
					}
					s_at[i] = at;
					cur = at + k;
					++app;
				}
				s_cur = cur, s_app += app, s_ovf += ovf;
			}
			__syncthreads();
			if (c < cnt) {
				if (found == kMiss) {
					const uint32_t at = s_at[threadIdx.x];
					if (at == kMiss)
						found = fallback ? fallback[item] : kNull;
					else {
#pragma unroll
						for (uint32_t i = 0; i < 9u; ++i)
							if (i < nw)
								words[base + at + i] = w[i];
						found = base + at;
					}
				}
				result[item] = found;
			}
			__syncthreads(); // s_nw / s_at are rewritten by the next chunk
		}
		if (threadIdx.x == 0) {
			bucket_words[bucket] = s_cur;
			atomicAdd(&ctr->stats[2], (unsigned long long)cnt);
			atomicAdd(&ctr->stats[7], (unsigned long long)bw);
			if (s_app) {
				atomicAdd(&ctr->stats[3], (unsigned long long)s_app);
				atomicAdd(&ctr->stats[4], (unsigned long long)(s_cur - bw));
			}
			if (s_ovf)
				atomicAdd(&ctr->stats[5], (unsigned long long)s_ovf);
		}
		return;
	} else {
		for (uint32_t c = warp; c < cnt; c += kGroupThreads / 32) { // tiny configs: sequential walk by lane 0
			const uint32_t item = order[first + c];
			const uint32_t *me = cand + size_t(item) * stride;
			const uint32_t nw = 1u + __popc(me[0] & 0xFFu);
			uint32_t found = kMiss;
			if (lane == 0) {
				for (uint32_t page = 0; page < bw && found == kMiss; page += wpp) {
					const uint32_t end = min(page + wpp, bw);
					for (uint32_t it = page; nw <= end - it;) {
						const uint32_t hw = img[it] & 0xFFu;
						if (hw == 0u)
							break;
						const uint32_t sz = 1u + __popc(hw);
						bool same = sz == nw;
						for (uint32_t i = 0; i < nw && same; ++i)
							same = img[it + i] == me[i];
						if (same) {
							found = base + it;
							break;
						}
						it += sz;
					}
				}
				result[item] = found;
			}
		}
	}
	__syncthreads();

	// phase B: warp 0 appends the misses in list order (append_node, NodePool.hpp:134-157); this CTA is the only writer
	if (warp != 0)
		return;
	uint32_t cur = bw, appended = 0, overflow = 0;
	for (uint32_t c = 0; c < cnt; ++c) {
		const uint32_t item = order[first + c];
		if (result[item] != kMiss)
			continue;
		const uint32_t *me = cand + size_t(item) * stride;
		const uint32_t nw = is_leaf ? 2u : 1u + __popc(me[0] & 0xFFu);
		const uint32_t off = cur & (wpp - 1u);
		const uint32_t at = off + nw > wpp ? (cur | (wpp - 1u)) + 1u : cur; // never straddle a page; the tail stays zero
		if (at + nw > wpb) { // bucket full: keep the old node (NodePool.hpp:137-139,195)
			if (lane == 0)
				result[item] = fallback ? fallback[item] : kNull;
			++overflow;
			continue;
		}
		if (lane < nw)
			words[base + at + lane] = me[lane];
		if (lane == 0)
			result[item] = base + at;
		cur = at + nw;
		++appended;
	}
	if (lane == 0) {
		bucket_words[bucket] = cur;
		atomicAdd(&ctr->stats[2], (unsigned long long)cnt);
		atomicAdd(&ctr->stats[7], (unsigned long long)bw);
		if (appended) {
			atomicAdd(&ctr->stats[3], (unsigned long long)appended);
			atomicAdd(&ctr->stats[4], (unsigned long long)(cur - bw));
		}
		if (overflow)
			atomicAdd(&ctr->stats[5], (unsigned long long)overflow);
	}
}

// Debug aid (HD_EDIT_VERIFY=1): after a level's upsert, every candidate's pointer must hold exactly its content.
__global__ void k_verify(uint32_t n, uint32_t stride, bool is_leaf, const uint32_t *cand, const uint8_t *state,
                         const uint32_t *winner, const uint32_t *result, const uint32_t *words, uint32_t *errs) {
	const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
	if (item >= n || state[item] == 0)
		return;
	const uint32_t *me = cand + size_t(item) * stride;
	const uint32_t nw = is_leaf ? 2u : 1u + __popc(me[0] & 0xFFu);
	const uint32_t ptr = state[item] == 1 ? result[winner[item]] : result[item];
	bool ok = ptr != kNull;
	for (uint32_t i = 0; i < nw && ok; ++i)
		ok = words[ptr + i] == me[i];
	if (!ok) {
		const uint32_t k = atomicAdd(&errs[0], 1u);
		if (k < 4)
			errs[1 + k * 4] = item, errs[2 + k * 4] = ptr, errs[3 + k * 4] = state[item], errs[4 + k * 4] = me[0];
	}
}

// dedup + find-or-insert + resolve over n candidates (state!=0) at `level`.
static hd_status run_upsert(hd_pool *p, uint32_t level, uint32_t n, uint32_t stride, const uint32_t *cand,
                            uint8_t *state, uint32_t *winner, const uint32_t *fallback, uint32_t *result) {
	if (n == 0)
		return HD_OK;
	EditScratch *s = p->edit;
	const bool is_leaf = level == p->geo.node_levels - 1;
	uint64_t tsize = 64;
	while (tsize < uint64_t(n) * 2)
		tsize <<= 1;
	// large batches: group the winners by bucket (one staged image and one writer per bucket); small ones and pools
	// whose buckets do not fit a 32 KB image: one warp per winner with a lock-free append
	static const int grouped_min = getenv("HD_EDIT_GROUPED_MIN") ? atoi(getenv("HD_EDIT_GROUPED_MIN")) : 16384;
	const uint32_t nb = 1u << p->geo.bucket_bits[level];
	const bool grouped = n >= uint32_t(grouped_min) && p->geo.words_per_bucket() <= kGroupMaxWords;
	uint32_t *bkt = nullptr, *count = nullptr, *offset = nullptr, *order = nullptr;
	if (grouped) { // the dedup pass counts the winners per bucket on its way
		HD_CUDA_TRY(amalloc(&bkt, n, p->stream));
		HD_CUDA_TRY(amalloc(&order, n, p->stream));
		HD_CUDA_TRY(amalloc(&count, nb, p->stream));
		HD_CUDA_TRY(amalloc(&offset, nb, p->stream));
		HD_CUDA_TRY(cudaMemsetAsync(count, 0, size_t(nb) * 4, p->stream));
	}
	uint32_t *table = nullptr;
	HD_CUDA_TRY(amalloc(&table, tsize, p->stream));
	HD_CUDA_TRY(cudaMemsetAsync(table, 0, tsize * 4, p->stream));
	k_dedup<<<grid_for(n), kBlock, 0, p->stream>>>(n, stride, is_leaf, cand, state, winner, table, uint32_t(tsize - 1), nullptr, nullptr, bkt,
	                                               count, nb - 1u);
	HD_LAUNCH_CHECK();
	HD_CUDA_TRY(cudaFreeAsync(table, p->stream));
	if (grouped) {
		hd_status ss = exclusive_scan(p, count, offset, nb);
		if (ss != HD_OK)
			return ss;
		HD_CUDA_TRY(cudaMemsetAsync(count, 0, size_t(nb) * 4, p->stream)); // reused as the per-bucket fill cursor...
		k_bucket_scatter<<<grid_for(n), kBlock, 0, p->stream>>>(n, state, bkt, offset, count, order);
		HD_LAUNCH_CHECK();
		// ...which ends up holding the per-bucket candidate count again
		// dynamic shared memory: the bucket image (4 bytes per bucket word), its content index (2) and the header positions (1):
		// 14 KB for 2 048-word buckets
		static const size_t gsm_pad = getenv("HD_EDIT_GROUPED_SMEM_PAD") ? size_t(atoi(getenv("HD_EDIT_GROUPED_SMEM_PAD"))) : 0; // residency experiments
		const size_t gsm = size_t(p->geo.words_per_bucket()) * 7 + gsm_pad;
		if (!p->grouped_smem_attr && gsm > 48u * 1024u) { // a per-device attribute: kept in the pool, not in a process static
			HD_CUDA_TRY(cudaFuncSetAttribute(k_upsert_grouped, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kGroupMaxWords) * 7 + int(gsm_pad)));
			p->grouped_smem_attr = true;
		}
		k_upsert_grouped<<<nb, kGroupThreads, gsm, p->stream>>>(
		    p->geo, level, s->fast_scan, stride, cand, fallback, result, p->words, p->bucket_words, offset, count, order, s->ctr);
		HD_LAUNCH_CHECK();
		cudaFreeAsync(bkt, p->stream), cudaFreeAsync(order, p->stream), cudaFreeAsync(count, p->stream), cudaFreeAsync(offset, p->stream);
	} else {
		k_upsert<<<upsert_grid(p, n), kBlock, 0, p->stream>>>(p->geo, level, s->fast_scan, n, stride, cand, state, fallback, result,
		                                                       p->words, p->bucket_words, s->ctr, nullptr);
		HD_LAUNCH_CHECK();
	}
	static const bool verify = getenv("HD_EDIT_VERIFY") != nullptr;
	if (verify) { // an overflowed bucket legitimately yields the fallback pointer and is reported here too
		uint32_t *errs = nullptr, host[17];
		HD_CUDA_TRY(amalloc(&errs, 17, p->stream));
		HD_CUDA_TRY(cudaMemsetAsync(errs, 0, sizeof(host), p->stream));
		k_verify<<<grid_for(n), kBlock, 0, p->stream>>>(n, stride, is_leaf, cand, state, winner, result, p->words, errs);
		HD_CUDA_TRY(cudaMemcpyAsync(host, errs, sizeof(host), cudaMemcpyDeviceToHost, p->stream));
		HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
		cudaFreeAsync(errs, p->stream);
		if (host[0])
			fprintf(stderr, "[hd verify] level %u: %u of %u candidates do not hold their content; first: item %u ptr %u state %u mask %#x\n",
			        level, host[0], n, host[1], host[2], host[3], host[4]);
	}
	return HD_OK;
}

// Find-or-insert of n packed nodes already on the device (all are candidates); result[i] = pointer.  Used by the GC.
hd_status upsert_batch_dev(hd_pool *p, uint32_t level, uint32_t n, uint32_t stride, const uint32_t *cand_dev,
                           uint32_t *result_dev) {
	if (n == 0)
		return HD_OK;
	hd_status st = scratch_init(p);
	if (st != HD_OK)
		return st;
	cudaStream_t s = p->stream;
	uint32_t *winner = nullptr;
	uint8_t *state = nullptr;
	HD_CUDA_TRY(amalloc(&winner, n, s));
	HD_CUDA_TRY(amalloc(&state, n, s));
	HD_CUDA_TRY(cudaMemsetAsync(result_dev, 0xFF, uint64_t(n) * 4, s));
	k_fill_u8<<<grid_for(n), kBlock, 0, s>>>(state, n, 1);
	HD_LAUNCH_CHECK();
	st = run_upsert(p, level, n, stride, cand_dev, state, winner, nullptr, result_dev);
	if (st == HD_OK) {
		k_resolve_flat<<<grid_for(n), kBlock, 0, s>>>(n, state, winner, result_dev);
		HD_LAUNCH_CHECK();
	}
	cudaFreeAsync(winner, s), cudaFreeAsync(state, s);
	return st;
}

// install new filled-node pointers (after a GC moved them)
hd_status set_filled(hd_pool *p, const std::vector<uint32_t> &filled) {
	hd_status st = scratch_init(p);
	if (st != HD_OK)
		return st;
	p->filled = filled;
	HD_CUDA_TRY(cudaMemcpyAsync(p->edit->filled_dev, p->filled.data(), p->filled.size() * 4, cudaMemcpyHostToDevice, p->stream));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	return HD_OK;
}

hd_status ensure_filled(hd_pool *p) { // make_filled_node_pointers, NodePool.hpp:240-262
	if (!p->filled.empty())
		return HD_OK;
	HD_CUDA_TRY(cudaSetDevice(p->device));
	hd_status st = scratch_init(p);
	if (st != HD_OK)
		return st;
	const uint32_t L = p->geo.node_levels;
	std::vector<uint32_t> filled(L, kNull);
	uint32_t prev = kNull;
	for (uint32_t l = L; l-- > 0;) {
		uint32_t node[9];
		uint32_t words_each;
		if (l == L - 1)
			node[0] = node[1] = 0xFFFFFFFFu, words_each = 2;
		else {
			node[0] = 0xFFu, words_each = 9;
			for (int i = 1; i < 9; ++i)
				node[i] = prev;
		}
		st = hd_upsert_nodes(p, l, node, words_each, 1, &filled[l]);
		if (st != HD_OK)
			return st;
		if (filled[l] == kNull) {
			set_error("bucket full while creating filled nodes");
			return HD_ERR_OVERFLOW;
		}
		prev = filled[l];
	}
	p->filled = filled;
	HD_CUDA_TRY(cudaMemcpyAsync(p->edit->filled_dev, p->filled.data(), L * 4, cudaMemcpyHostToDevice, p->stream));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	return HD_OK;
}

static hd_status read_counters(hd_pool *p, DevCounters &host) {
	HD_CUDA_TRY(cudaMemcpyAsync(&host, p->edit->ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, p->stream));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	return HD_OK;
}

static hd_status edit_batch_impl(hd_pool *p, uint32_t root_in, const hd_edit_desc *edits_host, uint32_t n_edits,
                                 uint32_t *root_out, hd_edit_stats *stats, std::vector<LevelAlloc> &levels,
                                 hd_edit_desc *&edits_dev, uint32_t *&iota) {
	const Geometry &g = p->geo;
	EditScratch *s = p->edit;
	cudaStream_t st = p->stream;
	const uint32_t L = g.node_levels;

	HD_CUDA_TRY(amalloc(&edits_dev, n_edits, st));
	HD_CUDA_TRY(cudaMemcpyAsync(edits_dev, edits_host, sizeof(hd_edit_desc) * n_edits, cudaMemcpyHostToDevice, st));
	HD_CUDA_TRY(amalloc(&iota, n_edits, st));
	k_iota<<<grid_for(n_edits), kBlock, 0, st>>>(iota, n_edits);
	HD_LAUNCH_CHECK();
	HD_CUDA_TRY(cudaMemsetAsync(s->ctr, 0, sizeof(DevCounters), st));

	bool terrain = false; // kernels for batches without a terrain edit have the generator compiled out
	for (uint32_t i = 0; i < n_edits && !terrain; ++i)
		terrain = edits_host[i].kind == HD_EDIT_TERRAIN_FILL;
	levels.resize(L);
	HD_CUDA_TRY(levels[0].init(1, n_edits, L == 1, st));
	if (n_edits > kHugeList)
		k_root_cta<<<1, kHugeThreads, 0, st>>>(g, edits_dev, n_edits, iota, s->filled_dev, root_in, levels[0].v, s->ctr);
	else
		k_root<<<1, 32, 0, st>>>(g, edits_dev, n_edits, iota, s->filled_dev, root_in, levels[0].v, s->ctr, nullptr);
	HD_LAUNCH_CHECK();
	DevCounters host{};
	hd_status rs = read_counters(p, host);
	if (rs != HD_OK)
		return rs;
	bool long_lists = n_edits > 32; // does the level about to be expanded hold a list longer than 32?
	bool huge_lists = n_edits > kHugeList; // ... longer than kHugeList?
	if (host.next_items == 0) { // the root was not entered at all
		*root_out = host.root_out;
		if (stats)
			memset(stats, 0, sizeof(*stats));
		return HD_OK;
	}
	levels[0].v.n = 1;

	// ---- top-down ----
	uint32_t deepest = 0;
	bool fused_leaves = false;
	uint64_t leaves_evaluated = 0;
	for (uint32_t l = 0; l + 1 < L; ++l) {
		LevelAlloc &in = levels[l];
		HD_CUDA_TRY(in.init_up(false));
		const uint64_t cap = uint64_t(in.v.n) * 8, cap_e = uint64_t(host.next_entries) * 8;
		if (cap > 0xFFFFFFF0ull || cap_e > 0xFFFFFFF0ull) {
			set_error("edit batch too large for one pass (%llu items)", (unsigned long long)cap);
			return HD_ERR_OVERFLOW;
		}
		LevelAlloc &out = levels[l + 1];
		// reset per-level cursors (stats keep accumulating)
		HD_CUDA_TRY(cudaMemsetAsync(&s->ctr->next_items, 0, 4 * sizeof(uint32_t), st));
		static const bool fuse_off = getenv("HD_EDIT_LEAF") != nullptr; // "half" / "lane": the separate leaf kernels
		static const bool pair_threads = getenv("HD_EDIT_DOWN_PAIRS") != nullptr; // the thread-per-(item, child) k_down, for A/B runs
		if (l + 2 == L && !terrain && !long_lists && !fuse_off) {
			// ---- last inner level + leaves fused (k_down_leaf): the leaf level holds the CHANGED leaves only ----
			HD_CUDA_TRY(out.init_compact_leaves(uint32_t(cap), st));
			k_down_leaf<<<(in.v.n + kLeafItems - 1u) / kLeafItems, kLeafItems, 0, st>>>(g, p->words, edits_dev, s->filled_dev, in.v, out.v,
			                                                                          s->ctr);
			HD_LAUNCH_CHECK();
			rs = read_counters(p, host);
			if (rs != HD_OK)
				return rs;
			if (host.error) {
				set_error("edit scratch overflow at the leaf level");
				return HD_ERR_OVERFLOW;
			}
			out.v.n = host.next_items;
			leaves_evaluated = host.stats[1];
			fused_leaves = true;
			if (out.v.n) {
				deepest = l + 1;
				HD_CUDA_TRY(out.init_up(true, true));
			}
			break;
		}
		HD_CUDA_TRY(out.init(uint32_t(cap), uint32_t(cap_e), l + 2 == L, st));
		if (terrain)
			k_down<true><<<grid_for(uint64_t(in.v.n) * 8), kBlock, 0, st>>>(g, l, p->words, edits_dev, s->filled_dev, in.v, out.v,
			                                                                s->ctr, &s->ctr->next_items, &s->ctr->next_entries);
		else if (pair_threads)
			k_down<false><<<grid_for(uint64_t(in.v.n) * 8), kBlock, 0, st>>>(g, l, p->words, edits_dev, s->filled_dev, in.v, out.v,
			                                                                 s->ctr, &s->ctr->next_items, &s->ctr->next_entries);
		else // a thread per item, the eight children classified together
			k_down_items<<<(in.v.n + kDownItems - 1u) / kDownItems, kDownItems, 0, st>>>(
			    g, l, p->words, edits_dev, s->filled_dev, in.v, out.v, s->ctr, &s->ctr->next_items, &s->ctr->next_entries);
		HD_LAUNCH_CHECK();
		if (long_lists) {
			k_down_long<<<grid_for(uint64_t(in.v.n) * 8 * 32), kBlock, 0, st>>>(g, l, p->words, edits_dev, s->filled_dev,
			                                                                 in.v, out.v, s->ctr);
			HD_LAUNCH_CHECK();
		}
		if (huge_lists && uint64_t(in.v.n) * 8 <= 0x7FFFFFFFull) {
			k_down_huge<<<in.v.n * 8u, kHugeThreads, 0, st>>>(g, l, p->words, edits_dev, s->filled_dev, in.v, out.v, s->ctr);
			HD_LAUNCH_CHECK();
		}
		rs = read_counters(p, host);
		if (rs != HD_OK)
			return rs;
		if (host.error) {
			set_error("edit scratch overflow at level %u", l + 1);
			return HD_ERR_OVERFLOW;
		}
		out.v.n = host.next_items;
		long_lists = host.has_long != 0;
		huge_lists = host.has_huge != 0;
		if (out.v.n == 0)
			break;
		deepest = l + 1;
	}

	// ---- leaves ----
	if (deepest == L - 1 && fused_leaves) { // the leaves were evaluated by k_down_leaf; only the changed ones are here
		LevelAlloc &lv = levels[L - 1];
		rs = run_upsert(p, L - 1, lv.v.n, 2, lv.v.cand, lv.v.state, lv.v.winner, lv.v.cur, lv.v.result);
		if (rs != HD_OK)
			return rs;
		k_resolve<<<grid_for(lv.v.n), kBlock, 0, st>>>(lv.v, levels[L - 2].v.child_new, s->ctr);
		HD_LAUNCH_CHECK();
	} else if (deepest == L - 1) {
		LevelAlloc &lv = levels[L - 1];
		HD_CUDA_TRY(lv.init_up(true));
		// one warp per 32 leaves; at least ~8 warps per SM-slot worth of grid so that small levels still spread out
		const uint32_t leaf_grid = grid_for(uint64_t((lv.v.n + 31u) / 32u) * 32u);
		static const bool use_half = getenv("HD_EDIT_LEAF") && !strcmp(getenv("HD_EDIT_LEAF"), "half");
		if (terrain)
			k_leaf<true><<<leaf_grid, kBlock, 0, st>>>(g, p->words, edits_dev, lv.v, s->ctr);
		else if (use_half)
			k_leaf_half<<<leaf_grid, kBlock, 0, st>>>(g, p->words, edits_dev, lv.v, s->ctr);
		else
			k_leaf_lane<<<grid_for(lv.v.n), kBlock, 0, st>>>(g, p->words, edits_dev, lv.v, s->ctr);
		HD_LAUNCH_CHECK();
		rs = run_upsert(p, L - 1, lv.v.n, 2, lv.v.cand, lv.v.state, lv.v.winner, lv.v.cur, lv.v.result);
		if (rs != HD_OK)
			return rs;
		k_resolve<<<grid_for(lv.v.n), kBlock, 0, st>>>(lv.v, L >= 2 ? levels[L - 2].v.child_new : nullptr, s->ctr);
		HD_LAUNCH_CHECK();
	}
	// ---- bottom-up ----
	for (uint32_t l = std::min(deepest, L - 2) + 1; l-- > 0;) {
		if (L == 1)
			break;
		LevelAlloc &lv = levels[l];
		if (lv.v.n == 0)
			continue;
		k_assemble<<<grid_for(lv.v.n), kBlock, 0, st>>>(p->words, lv.v);
		HD_LAUNCH_CHECK();
		rs = run_upsert(p, l, lv.v.n, 9, lv.v.cand, lv.v.state, lv.v.winner, lv.v.cur, lv.v.result);
		if (rs != HD_OK)
			return rs;
		k_resolve<<<grid_for(lv.v.n), kBlock, 0, st>>>(lv.v, l ? levels[l - 1].v.child_new : nullptr, s->ctr);
		HD_LAUNCH_CHECK();
	}
	rs = read_counters(p, host);
	if (rs != HD_OK)
		return rs;
	*root_out = host.root_out;
	if (stats) {
		stats->visited_nodes = 0;
		for (uint32_t l = 0; l + 1 < L; ++l)
			stats->visited_nodes += levels[l].v.n;
		stats->visited_leaves = fused_leaves ? leaves_evaluated : levels[L - 1].v.n;
		stats->upserts = host.stats[2];
		stats->appended_nodes = host.stats[3];
		stats->appended_words = host.stats[4];
		stats->overflow_count = host.stats[5];
		stats->in_range_voxels = 0;
		stats->scan_words = host.stats[7];
	}
	return HD_OK;
}

// ---- low-latency path ------------------------------------------------------------------------------------------
static uint32_t fast_grid(hd_pool *p, uint64_t threads) {
	const int sms = std::max(p->sm_count, 1);
	return uint32_t(std::max<uint64_t>(1, std::min<uint64_t>((threads + kBlock - 1) / kBlock, uint64_t(sms) * 8)));
}

// Enqueue the whole rebuild on the pool's stream with device-resident counts (captured once into a graph).
static hd_status fast_enqueue(hd_pool *p) {
	const Geometry &g = p->geo;
	EditScratch *s = p->edit;
	FastPath &f = s->fast;
	cudaStream_t st = p->stream;
	const uint32_t L = g.node_levels;
	HD_CUDA_TRY(cudaMemcpyAsync(f.dyn_dev, f.dyn_host, sizeof(FastDyn), cudaMemcpyHostToDevice, st));
	HD_CUDA_TRY(cudaMemsetAsync(s->ctr, 0, sizeof(DevCounters), st));
	k_root<<<1, 32, 0, st>>>(g, f.dyn_dev->edits, 0, f.iota_dev, s->filled_dev, 0, f.lv[0], s->ctr, &f.dyn_dev->root);
	HD_LAUNCH_CHECK();
	for (uint32_t l = 0; l + 1 < L; ++l) {
		k_down<false><<<fast_grid(p, uint64_t(f.lv[l].cap) * 8), kBlock, 0, st>>>(g, l, p->words, f.dyn_dev->edits, s->filled_dev, f.lv[l],
		                                                                 f.lv[l + 1], s->ctr, &s->ctr->lvl_items[l + 1],
		                                                                 &s->ctr->lvl_entries[l + 1]);
		HD_LAUNCH_CHECK();
	}
	for (uint32_t l = L; l-- > 0;) {
		const LevelView &v = f.lv[l];
		k_up<<<fast_grid(p, uint64_t(v.cap) * 32), kBlock, 0, st>>>(g, l, s->fast_scan, p->words, p->bucket_words, f.locks,
		                                                          f.dyn_dev->edits, v, l ? f.lv[l - 1].child_new : nullptr, s->ctr);
		HD_LAUNCH_CHECK();
	}
	HD_CUDA_TRY(cudaMemcpyAsync(f.ctr_host, s->ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, st));
	return HD_OK;
}

static hd_status fast_build(hd_pool *p) {
	const Geometry &g = p->geo;
	EditScratch *s = p->edit;
	FastPath &f = s->fast;
	const uint32_t L = g.node_levels;
	static const uint32_t cap_max = getenv("HD_EDIT_FAST_CAP") ? uint32_t(atoi(getenv("HD_EDIT_FAST_CAP"))) : (1u << 19);
	f.lv.assign(L, LevelView{});
	// pass 1 sizes the arena, pass 2 hands out the pointers
	size_t total = 0;
	for (int pass = 0; pass < 2; ++pass) {
		size_t off = 0;
		auto take = [&](size_t bytes) {
			char *ptr = pass ? f.arena + off : nullptr;
			off += (bytes + 255) & ~size_t(255);
			return ptr;
		};
		for (uint32_t l = 0; l < L; ++l) {
			const uint64_t full = l >= 11 ? ~0ull : 1ull << (3 * l); // 8^l nodes exist at level l
			const uint32_t cap = uint32_t(std::min<uint64_t>(full, cap_max));
			const uint32_t cap_e = uint32_t(std::min<uint64_t>(uint64_t(cap) * kFastMaxEdits, uint64_t(cap_max) * 4));
			const bool leaf = l == L - 1;
			LevelView &v = f.lv[l];
			v.n = 0, v.cap = cap, v.cap_entries = cap_e;
			v.n_dev = &s->ctr->lvl_items[l], v.err_dev = &s->ctr->error;
			v.cur = (uint32_t *)take(size_t(cap) * 4), v.pos = (uint64_t *)take(size_t(cap) * 8);
			v.list_off = (uint32_t *)take(size_t(cap) * 4), v.list_len = (uint32_t *)take(size_t(cap) * 4);
			v.parent = (uint32_t *)take(size_t(cap) * 4);
			v.lists = (uint32_t *)take(size_t(cap_e) * 4);
			v.child_new = leaf ? nullptr : (uint32_t *)take(size_t(cap) * 8 * 4);
		}
		f.locks = (uint32_t *)take(size_t(g.total_buckets) * 4);
		total = off;
		if (!pass) {
			HD_CUDA_TRY(cudaMalloc(&f.arena, total));
			HD_CUDA_TRY(cudaMemset(f.arena, 0, total)); // the bucket locks start (and always end) released
		}
	}
	HD_CUDA_TRY(cudaMalloc(&f.dyn_dev, sizeof(FastDyn)));
	HD_CUDA_TRY(cudaMalloc(&f.iota_dev, kFastMaxEdits * 4));
	HD_CUDA_TRY(cudaHostAlloc(&f.dyn_host, sizeof(FastDyn), cudaHostAllocMapped));
	HD_CUDA_TRY(cudaHostAlloc(&f.ctr_host, sizeof(DevCounters), cudaHostAllocMapped));
	HD_CUDA_TRY(cudaHostGetDevicePointer(&f.dyn_host_dev, f.dyn_host, 0));
	HD_CUDA_TRY(cudaHostGetDevicePointer(&f.ctr_host_dev, f.ctr_host, 0));
	uint32_t iota[kFastMaxEdits];
	for (uint32_t i = 0; i < kFastMaxEdits; ++i)
		iota[i] = i;
	HD_CUDA_TRY(cudaMemcpy(f.iota_dev, iota, sizeof(iota), cudaMemcpyHostToDevice));
	memset(f.dyn_host, 0, sizeof(FastDyn));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	f.mode = getenv("HD_EDIT_FAST") && atoi(getenv("HD_EDIT_FAST")) == 2 ? 2 : 1;
	if (f.mode == 1) {
		int coop = 0, per_sm = 0, sms = 0;
		HD_CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, p->device));
		HD_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device));
		HD_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_edit_fused, kFusedThreads, 0));
		if (coop && per_sm > 0) {
			f.fused_grid = uint32_t(sms) * uint32_t(std::min(per_sm, 2));
			return HD_OK;
		}
		f.mode = 2; // no cooperative launch on this device: same phases as a graph
	}
	const uint64_t before = g_launches.load();
	HD_CUDA_TRY(cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeRelaxed));
	const hd_status es = fast_enqueue(p);
	const cudaError_t ce = cudaStreamEndCapture(p->stream, &f.graph);
	f.kernels = uint32_t(g_launches.load() - before);
	g_launches.store(before); // captured, not launched
	if (es != HD_OK)
		return es;
	HD_CUDA_TRY(ce);
	HD_CUDA_TRY(cudaGraphInstantiate(&f.exec, f.graph, 0));
	return HD_OK;
}

// The one-launch rebuild in two halves, so that a caller can run other work while the kernel is in flight (the colour edit
// overlaps its octree pass, color.cu).  fast_edit_begin enqueues the rebuild on the pool's stream when the batch qualifies
// (*launched = true) and returns without waiting; fast_edit_end waits for it and reports: *handled = true when the call
// was served, false when a work queue turned out too small — nothing was written to the pool then and the general path
// must redo the call.
// share_gpu: launch one CTA per SM instead of two — two 512-thread CTAs of 64 registers take an SM's whole register file,
// and nothing the caller enqueues on another stream could run beside the rebuild.
hd_status fast_edit_begin(hd_pool *p, uint32_t root_in, const hd_edit_desc *edits, uint32_t n, bool *launched, bool share_gpu) {
	*launched = false;
	static const bool enabled = !(getenv("HD_EDIT_FAST") && atoi(getenv("HD_EDIT_FAST")) == 0); // 0 general, 1 fused, 2 graph
	EditScratch *s = p->edit;
	FastPath &f = s->fast;
	const uint32_t L = p->geo.node_levels;
	if (!enabled || n > kFastMaxEdits || L < 2)
		return HD_OK;
	for (uint32_t i = 0; i < n; ++i) // terrain fills touch a large share of the world: always the general path
		if (edits[i].kind == HD_EDIT_TERRAIN_FILL)
			return HD_OK;
	if (!f.tried) {
		f.tried = true;
		f.ok = fast_build(p) == HD_OK;
		if (!f.ok) {
			cudaGetLastError(); // e.g. no memory for the arena: the general path needs none of it
			cudaStreamSynchronize(p->stream);
		}
	}
	if (!f.ok)
		return HD_OK;
	f.dyn_host->root = root_in, f.dyn_host->n_edits = n;
	memcpy(f.dyn_host->edits, edits, sizeof(hd_edit_desc) * n);
	if (f.mode == 1) {
		FusedArgs a{};
		a.g = p->geo;
		a.words = p->words, a.bucket_words = p->bucket_words, a.locks = f.locks;
		a.iota = f.iota_dev, a.filled = s->filled_dev;
		a.dyn_host = f.dyn_host_dev, a.dyn_dev = f.dyn_dev;
		a.ctr = s->ctr, a.ctr_host = f.ctr_host_dev;
		for (uint32_t l = 0; l < L; ++l)
			a.lv[l] = f.lv[l];
		a.fast_scan = s->fast_scan;
		static const bool wide_scan = !(getenv("HD_EDIT_SCAN_WIDE") && atoi(getenv("HD_EDIT_SCAN_WIDE")) == 0);
		a.wide_scan = wide_scan;
		void *params[] = {&a};
		// One CTA per SM for a brush-sized call (a handful of editors): the grid barriers are cheaper with half the CTAs, and a
		// single r <= 128 brush has no level that needs more than 148 x 16 warps (measured, r = 2 / 32 / 128 / 256: 0.173 / 0.201 /
		// 0.229 / 0.274 ms with two CTAs per SM, 0.162 / 0.184 / 0.216 / 0.283 with one; 100 editors need both).
		uint32_t grid = share_gpu || n <= 4u ? std::min<uint32_t>(f.fused_grid, uint32_t(p->sm_count > 0 ? p->sm_count : 148)) : f.fused_grid;
		static const uint32_t forced_grid = getenv("HD_EDIT_FUSED_GRID") ? uint32_t(atoi(getenv("HD_EDIT_FUSED_GRID"))) : 0u;
		if (forced_grid) // A/B knob: fewer CTAs make the grid barriers cheaper and the big levels slower
			grid = std::min(grid, std::max(forced_grid, 1u));
		HD_CUDA_TRY(cudaLaunchCooperativeKernel((const void *)k_edit_fused, dim3(grid), dim3(kFusedThreads), params, 0, p->stream));
		g_launches.fetch_add(1, std::memory_order_relaxed);
	} else {
		HD_CUDA_TRY(cudaGraphLaunch(f.exec, p->stream));
		g_launches.fetch_add(f.kernels, std::memory_order_relaxed);
	}
	f.pending_edits = n;
	*launched = true;
	return HD_OK;
}

hd_status fast_edit_end(hd_pool *p, uint32_t *root_out, hd_edit_stats *stats, bool *handled) {
	*handled = false;
	EditScratch *s = p->edit;
	FastPath &f = s->fast;
	const uint32_t L = p->geo.node_levels, n = f.pending_edits;
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	const DevCounters &c = *f.ctr_host;
	if (c.error)
		return HD_OK; // some queue was too small; nothing was written to the pool
	*handled = true;
	s->last_path = uint32_t(f.mode);
	*root_out = c.root_out;
	static const bool trace = getenv("HD_EDIT_FAST_TRACE") && atoi(getenv("HD_EDIT_FAST_TRACE"));
	if (trace && f.mode == 1 && c.n_stamps) {
		fprintf(stderr, "[hd] fused edit: %u editors, items per level:", n);
		for (uint32_t l = 0; l < L; ++l)
			fprintf(stderr, " %u", c.lvl_items[l]);
		fprintf(stderr, "\n[hd]   phase ends (us since kernel start; solo top-down | grid top-down... | grid bottom-up... | solo bottom-up):");
		for (uint32_t i = 1; i < c.n_stamps; ++i)
			fprintf(stderr, " %.1f", double(c.stamp_ns[i] - c.stamp_ns[0]) * 1e-3);
		fprintf(stderr, "\n");
	}
	if (stats) {
		memset(stats, 0, sizeof(*stats));
		for (uint32_t l = 0; l + 1 < L; ++l)
			stats->visited_nodes += c.lvl_items[l];
		stats->visited_leaves = c.lvl_items[L - 1];
		stats->upserts = c.stats[2], stats->appended_nodes = c.stats[3], stats->appended_words = c.stats[4];
		stats->overflow_count = c.stats[5], stats->scan_words = c.stats[7];
	}
	return HD_OK;
}

// the scratch both halves and the general path need (hd_edit_batch's preamble)
hd_status edit_prepare(hd_pool *p) {
	hd_status st = scratch_init(p);
	if (st == HD_OK)
		st = ensure_filled(p); // NodePool.hpp:408
	if (st == HD_OK)
		p->edit->last_path = 0;
	return st;
}

// Returns HD_OK with *handled = true when the call was served here; *handled = false sends it to the general path.
static hd_status fast_edit(hd_pool *p, uint32_t root_in, const hd_edit_desc *edits, uint32_t n, uint32_t *root_out,
                           hd_edit_stats *stats, bool *handled) {
	*handled = false;
	bool launched = false;
	hd_status st = fast_edit_begin(p, root_in, edits, n, &launched, false);
	if (st != HD_OK || !launched)
		return st;
	return fast_edit_end(p, root_out, stats, handled);
}

// hd_selftest_edit_node8: edit_node8 against eight edit_node calls on pseudo-random editors and nodes that stress the
// boundaries — boxes that touch / contain / straddle the editor, spheres around and far from the node (the 32-bit and the
// 64-bit paths), radii from 0 to beyond the world, every level from leaves (bits 2) to 2^16-voxel nodes, coordinates up to 2^21.
__global__ void k_selftest_edit_node8(uint32_t n, unsigned long long *bad) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n)
		return;
	uint32_t r = fmix32(t * 2654435761u + 0x9E3779B9u);
	auto next = [&]() { return r = fmix32(r + 0x85EBCA6Bu); };
	const uint32_t bits = 2u + next() % 15u;            // child level: 4 .. 65 536 voxels per node edge
	const uint32_t span = 21u - (bits + 1u);              // parent positions keep every coordinate below 2^21
	const uint32_t x = next() & ((1u << span) - 1u), y = next() & ((1u << span) - 1u), z = next() & ((1u << span) - 1u);
	const uint32_t s = 1u << bits, L[3] = {(x << 1) << bits, (y << 1) << bits, (z << 1) << bits};
	hd_edit_desc d{};
	d.kind = next() % 3u; // AABB, sphere fill, sphere dig
	const uint32_t mode = next() % 4u;
	for (int a = 0; a < 3; ++a) {
		// anchor near one of the node's planes (mode 0-2) or anywhere in the world (mode 3), jittered by up to +-2 voxels / a node
		const uint32_t plane = L[a] + s * (next() % 3u);
		const int32_t jitter = mode == 0u ? int32_t(next() % 5u) - 2 : mode == 1u ? int32_t(next() % (2u * s + 1u)) - int32_t(s) : 0;
		uint32_t p = mode == 3u ? next() & 0x1FFFFFu : uint32_t(max(0, int32_t(plane) + jitter));
		d.p0[a] = p;
		d.p1[a] = p + (next() % 4u == 0u ? 0u : 1u + next() % (3u * s));
	}
	const uint32_t rk = next() % 6u;
	const uint64_t rad = rk == 0u ? 0ull : rk == 1u ? uint64_t(next() % 8u) : rk == 2u ? uint64_t(next() % (4u * s)) : rk == 3u ? uint64_t(s)
	                   : rk == 4u ? uint64_t(next() & 0x3FFFFFu) : uint64_t(next());
	d.r2 = rk == 3u ? rad * rad * (next() % 4u) : rad * rad + next() % 3u;
	const uint32_t got = edit_node8(d, bits, x, y, z);
	uint32_t want = 0u;
	for (uint32_t c = 0; c < 8u; ++c)
		want |= uint32_t(edit_node<false>(d, bits, (x << 1) | (c & 1u), (y << 1) | ((c >> 1) & 1u), (z << 1) | (c >> 2))) << (2u * c);
	if (got != want)
		atomicAdd(bad, 1ull);
}

} // namespace hd

using namespace hd;

extern "C" {

hd_status hd_edit_batch(hd_pool *p, uint32_t root_in, const hd_edit_desc *edits, uint32_t n, uint32_t *root_out,
                        hd_edit_stats *stats) {
	if (!p || !root_out || (!edits && n))
		return HD_ERR_INVALID;
	*root_out = root_in;
	if (stats)
		memset(stats, 0, sizeof(*stats));
	if (n == 0)
		return HD_OK;
	for (uint32_t i = 0; i < n; ++i)
		if (edits[i].kind > HD_EDIT_TERRAIN_FILL) {
			set_error("unknown edit kind %u", edits[i].kind);
			return HD_ERR_INVALID;
		}
	HD_CUDA_TRY(cudaSetDevice(p->device));
	hd_status st = edit_prepare(p);
	if (st != HD_OK)
		return st;
	bool handled = false;
	st = fast_edit(p, root_in, edits, n, root_out, stats, &handled);
	if (st != HD_OK || handled) {
		if (st != HD_OK)
			*root_out = root_in;
		return st;
	}
	std::vector<LevelAlloc> levels;
	hd_edit_desc *edits_dev = nullptr;
	uint32_t *iota = nullptr;
	st = edit_batch_impl(p, root_in, edits, n, root_out, stats, levels, edits_dev, iota);
	for (auto &l : levels)
		l.release();
	if (edits_dev)
		cudaFreeAsync(edits_dev, p->stream);
	if (iota)
		cudaFreeAsync(iota, p->stream);
	cudaStreamSynchronize(p->stream);
	if (st != HD_OK)
		*root_out = root_in;
	return st;
}

uint32_t hd_edit_last_path(const hd_pool *p) { return p && p->edit ? p->edit->last_path : 0u; }

hd_status hd_selftest_edit_node8(int device, uint32_t n_cases, uint64_t *mismatches) {
	if (!mismatches)
		return HD_ERR_INVALID;
	HD_CUDA_TRY(cudaSetDevice(device));
	unsigned long long *bad = nullptr;
	HD_CUDA_TRY(cudaMalloc(&bad, sizeof(*bad)));
	ScopeExit guard{[&]() { cudaFree(bad); }};
	HD_CUDA_TRY(cudaMemset(bad, 0, sizeof(*bad)));
	k_selftest_edit_node8<<<(n_cases + 255u) / 256u, 256>>>(n_cases, bad);
	HD_LAUNCH_CHECK();
	unsigned long long host = 0;
	HD_CUDA_TRY(cudaMemcpy(&host, bad, sizeof(host), cudaMemcpyDeviceToHost));
	*mismatches = host;
	return HD_OK;
}

hd_status hd_upsert_nodes(hd_pool *p, uint32_t level, const uint32_t *nodes, uint32_t words_each, uint32_t n,
                          uint32_t *out_ptrs) {
	if (!p || !nodes || !out_ptrs || level >= p->geo.node_levels || words_each < 2 || words_each > 9)
		return HD_ERR_INVALID;
	const bool is_leaf = level == p->geo.node_levels - 1;
	if (is_leaf && words_each != 2)
		return HD_ERR_INVALID;
	if (n == 0)
		return HD_OK;
	if (!is_leaf)
		for (uint32_t i = 0; i < n; ++i) {
			const uint32_t m = nodes[size_t(i) * words_each];
			if (m == 0 || m > 0xFFu || 1u + uint32_t(__builtin_popcount(m)) > words_each)
				return HD_ERR_INVALID;
		}
	HD_CUDA_TRY(cudaSetDevice(p->device));
	hd_status st = scratch_init(p);
	if (st != HD_OK)
		return st;
	cudaStream_t s = p->stream;
	uint32_t *cand = nullptr, *winner = nullptr, *result = nullptr;
	uint8_t *state = nullptr;
	HD_CUDA_TRY(amalloc(&cand, uint64_t(n) * words_each, s));
	HD_CUDA_TRY(amalloc(&winner, n, s));
	HD_CUDA_TRY(amalloc(&result, n, s));
	HD_CUDA_TRY(amalloc(&state, n, s));
	HD_CUDA_TRY(cudaMemcpyAsync(cand, nodes, uint64_t(n) * words_each * 4, cudaMemcpyHostToDevice, s));
	HD_CUDA_TRY(cudaMemsetAsync(result, 0xFF, uint64_t(n) * 4, s));
	k_fill_u8<<<grid_for(n), kBlock, 0, s>>>(state, n, 1);
	HD_LAUNCH_CHECK();
	st = run_upsert(p, level, n, words_each, cand, state, winner, nullptr, result);
	if (st == HD_OK) {
		k_resolve_flat<<<grid_for(n), kBlock, 0, s>>>(n, state, winner, result);
		hd::g_launches.fetch_add(1);
		HD_CUDA_TRY(cudaMemcpyAsync(out_ptrs, result, uint64_t(n) * 4, cudaMemcpyDeviceToHost, s));
	}
	cudaFreeAsync(cand, s), cudaFreeAsync(winner, s), cudaFreeAsync(result, s), cudaFreeAsync(state, s);
	HD_CUDA_TRY(cudaStreamSynchronize(s));
	return st;
}

hd_status hd_pool_used_words(hd_pool *p, uint64_t *out) {
	if (!p || !out)
		return HD_ERR_INVALID;
	std::vector<uint32_t> bw(p->geo.total_buckets);
	hd_status s = hd_pool_read_bucket_words(p, 0, bw.data(), p->geo.total_buckets);
	if (s != HD_OK)
		return s;
	uint64_t sum = 0;
	for (uint32_t v : bw)
		sum += v;
	*out = sum;
	return HD_OK;
}

} // extern "C"
