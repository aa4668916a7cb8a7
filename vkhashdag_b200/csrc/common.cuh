// common.cuh — shared host/device declarations of libhashdag_b200.so (sm_100a only, no CPU fallback).
#pragma once
#include "../../include/hashdag_b200.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

namespace hd {

constexpr uint32_t kNull = HD_NULL_NODE;

// ---- error plumbing -----------------------------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define HD_CUDA_TRY(expr)                                                                                              \
	do {                                                                                                               \
		cudaError_t _e = (expr);                                                                                       \
		if (_e != cudaSuccess) {                                                                                       \
			hd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);                 \
			return _e == cudaErrorMemoryAllocation ? HD_ERR_OOM : HD_ERR_CUDA;                                         \
		}                                                                                                              \
	} while (0)

#define HD_LAUNCH_CHECK()                                                                                              \
	do {                                                                                                               \
		hd::g_launches.fetch_add(1, std::memory_order_relaxed);                                                        \
		HD_CUDA_TRY(cudaGetLastError());                                                                               \
	} while (0)

// runs `f` on every exit path of the enclosing scope (the early returns of HD_CUDA_TRY included)
struct ScopeExit {
	std::function<void()> f;
	~ScopeExit() {
		if (f)
			f();
	}
};

// ---- pool geometry (include/hashdag/Config.hpp:15-57), passed to kernels by value -----------------
struct Geometry {
	uint32_t word_bits_per_page;
	uint32_t page_bits_per_bucket;
	uint32_t node_levels;
	uint32_t bucket_bits[HD_MAX_NODE_LEVELS];
	uint32_t level_base[HD_MAX_NODE_LEVELS]; // Config.hpp:33-38
	uint32_t total_buckets;
	uint64_t total_words;

	__host__ __device__ uint32_t words_per_page() const { return 1u << word_bits_per_page; }
	__host__ __device__ uint32_t bucket_shift() const { return word_bits_per_page + page_bits_per_bucket; }
	__host__ __device__ uint32_t words_per_bucket() const { return 1u << bucket_shift(); }
	__host__ __device__ uint32_t voxel_level() const { return node_levels + 1u; }
	__host__ __device__ uint32_t level_of_bucket(uint32_t bucket) const {
		uint32_t l = 0;
		while (l + 1 < node_levels && bucket >= level_base[l + 1])
			++l;
		return l;
	}
};

bool make_geometry(const hd_config &cfg, Geometry &g);

// hd_tile_shard ownership map (include/hashdag_b200.h): the rank's `local`-th tile -> tile coordinates
__host__ __device__ __forceinline__ void tile_of(uint32_t local, uint32_t rank, uint32_t world, uint32_t tiles_x, uint32_t &tx,
                                                 uint32_t &ty) {
	if (tiles_x % world != 0u) {
		const uint32_t t = local * world + rank;
		tx = t % tiles_x, ty = t / tiles_x;
	} else {
		const uint32_t per_row = tiles_x / world;
		ty = local / per_row;
		tx = (local % per_row) * world + (rank + world - ty % world) % world;
	}
}

struct EditScratch; // edit.cu

} // namespace hd

// The opaque pool handle of the C ABI.
struct hd_pool {
	hd_config cfg{};
	hd::Geometry geo{};
	int device = 0;
	int sm_count = 0;               // SMs of `device` (grids are sized per device, not per process)
	bool grouped_smem_attr = false; // cudaFuncSetAttribute(k_upsert_grouped, MaxDynamicSharedMemorySize) done on `device`
	cudaStream_t stream = nullptr;

	uint32_t *words = nullptr;        // flat word space, SURVEY App. A.1 (device)
	uint32_t *bucket_words = nullptr; // used words per bucket incl. page padding (device)
	uint32_t *bucket_synced = nullptr; // bucket_words at the last hd_dirty_reset (device)

	uint32_t *color_nodes = nullptr, *color_leaves = nullptr; // DAGColorPool buffers (device)
	uint64_t color_node_words = 0, color_leaf_words = 0; // used words (host copy, refreshed after colour edits)
	uint64_t color_node_cap = 0, color_leaf_cap = 0;     // allocated words
	uint32_t color_root = HD_COLOR_NULL, color_leaf_level = 0;
	uint32_t *color_ctr = nullptr; // device: [0] = nodes used, [1] = leaf words used, [2] = out-of-space flag
	cudaStream_t color_stream = nullptr; // hd_edit_color runs its octree pass here while the geometry rebuild is in flight
	// colour replica sync (sync.cu): nodes and appended leaf chunks are append-only, chunks rewritten in place since the
	// last sync are listed on the device (color.cu appends to the list)
	uint64_t color_synced_node_words = 0, color_synced_leaf_words = 0;
	bool color_dirty = false;       // colour root / buffers changed since the last hd_dirty_reset
	bool color_full_resync = false; // replicas must take the whole colour pool (hd_color_upload, dirty list overflow)
	uint32_t *color_dirty_list = nullptr, *color_dirty_ctr = nullptr; // device: chunk word indices; [0] = count, [1] = overflow

	uint32_t root = HD_NULL_NODE;
	bool needs_full_resync = false; // set by hd_gc: replicas must clear before applying the next sync
	std::vector<uint32_t> filled; // m_filled_node_pointers (NodePool.hpp:54)

	// trace staging (device) for the host-pointer entry points
	uint32_t *stage_rgba = nullptr, *stage_iters = nullptr, *stage_fetches = nullptr;
	hd_hit_record *stage_hits = nullptr;
	uint64_t stage_pixels = 0;
	hd_trace_params *params_dev = nullptr;
	// per-column / per-row NDC coordinates of the pixel centres for the current frame size (trace.cu: ray_tables)
	float *ray_table = nullptr;
	uint64_t ray_cap = 0;
	uint32_t ray_w = 0, ray_h = 0;
	// staged top levels of the DAG below `tt_root` (trace.cu: trace_table_*): node levels [0, tt_levels) unfolded into a
	// dense array of {child reference, child's mask} pairs, read with one 64-bit load per descent.  Nodes are immutable
	// and buckets append-only, so a table stays valid for its root until the word space is rewritten (tt_invalidate()).
	uint2 *tt_entries = nullptr;    // [tt_cap_nodes * 8]
	uint32_t *tt_masks = nullptr;   // [tt_cap_nodes] child mask of every staged node
	uint32_t *tt_list[2] = {nullptr, nullptr}; // BFS frontier (pool pointers), ping-pong
	uint32_t *tt_count = nullptr;   // device counter
	uint32_t tt_cap_nodes = 0, tt_root = HD_NULL_NODE, tt_levels = 0, tt_nodes = 0;
	bool tt_valid = false;
	uint32_t tt_seen_root = HD_NULL_NODE, tt_seen_frames = 0; // auto mode: build once a root has been traced twice
	void tt_invalidate() { tt_valid = false, tt_seen_root = HD_NULL_NODE, tt_seen_frames = 0; }

	uint32_t *persist_ctr = nullptr; // chunk counter of the persistent trace kernel (trace.cu: trace_persist_kernel)

	// pipelined frames (hd_trace_submit / hd_trace_collect): two slots, copy stream overlaps the next trace
	cudaStream_t copy_stream = nullptr;
	uint32_t *pipe_rgba[2] = {nullptr, nullptr};
	uint64_t pipe_pixels[2] = {0, 0};
	uint64_t pipe_sig[2][3] = {{0, 0, 0}, {0, 0, 0}}; // frame / shard geometry the slot's padding was last cleared for
	cudaEvent_t pipe_traced[2] = {nullptr, nullptr}, pipe_done[2] = {nullptr, nullptr};
	bool pipe_busy[2] = {false, false};

	hd::EditScratch *edit = nullptr;

	// pick ray (hd_traverse_ray): result block in mapped pinned host memory, written by the kernel itself
	float *pick_host = nullptr, *pick_host_dev = nullptr;

	// dirty-range scratch
	uint32_t *dirty_scratch = nullptr; // [0]=n_ranges, [1]=payload words, then per-range data
	uint64_t dirty_scratch_bytes = 0;
};

namespace hd {
hd_status edit_scratch_free(hd_pool *pool);
// the one-launch rebuild in two halves (edit.cu): prepare, enqueue without waiting, wait + report
hd_status edit_prepare(hd_pool *pool);
hd_status fast_edit_begin(hd_pool *pool, uint32_t root_in, const hd_edit_desc *edits, uint32_t n, bool *launched, bool share_gpu);
hd_status fast_edit_end(hd_pool *pool, uint32_t *root_out, hd_edit_stats *stats, bool *handled);
hd_status ensure_filled(hd_pool *pool);
hd_status upsert_batch_dev(hd_pool *pool, uint32_t level, uint32_t n, uint32_t stride, const uint32_t *cand_dev,
                           uint32_t *result_dev);
hd_status set_filled(hd_pool *pool, const std::vector<uint32_t> &filled);
// device-wide exclusive prefix sum over u32 (color.cu); in == out is allowed
hd_status exclusive_scan(hd_pool *pool, const uint32_t *in, uint32_t *out, uint64_t n, cudaStream_t stream = nullptr);
constexpr uint32_t kColorDirtyCap = 1u << 20; // entries of hd_pool::color_dirty_list
hd_status ensure_color_storage(hd_pool *p, uint64_t node_words, uint64_t leaf_words); // color.cu
} // namespace hd
