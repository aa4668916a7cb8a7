// oracle/ref_harness.cpp — TEST INFRASTRUCTURE.  Thin C wrapper around the UNMODIFIED reference headers.
//
// Compiled by oracle/Makefile with -I/root/reference/include (+ vendored glm / libfork) into
// oracle/_ref/libhashdag_ref.so.  Nothing from the reference tree is copied into this repository: the headers
// are included where they lie.  The resulting library is (a) what pins the CPU restatement in
// hashdag_oracle.cpp, (b) the generator of tests/golden/*.json and (c) the `"kind": "reference"` CPU baseline
// of bench.py.  It is never loaded by the product.
//
// What is written here (because the reference only has it inside its Vulkan app) and which lines it follows:
//   * RefPool        — an in-memory pool satisfying the NodePool/ThreadedNodePool concepts
//                      (include/hashdag/NodePool.hpp:22-35), storage contract of src/DAGNodePool.hpp:41-74.
//   * editors        — AABB / sphere predicates with the semantics of src/main.cpp:32-150, terrain generator
//                      of oracle/terrain.h; all driven by the POD hd_edit_desc.
//   * RefColorPool   — Vulkan-free colour octree with the word layout of src/DAGColorPool.hpp:23-34,66-113,139-204
//                      implementing the VBROctree concept (include/hashdag/VBROctree.hpp:19-32).
#include "../include/hashdag_b200.h"
#include "terrain.h"

#include <hashdag/NodePool.hpp>
#include <hashdag/NodePoolThreadedEdit.hpp>
#include <hashdag/NodePoolThreadedGC.hpp>
#include <hashdag/NodePoolTraversal.hpp>
#include <hashdag/VBREditor.hpp>

#include <libfork/schedule/busy_pool.hpp>
#include <parallel_hashmap/phmap.h>

#include <array>
#include <atomic>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <sys/mman.h>
#include <thread>
#include <vector>

using hashdag::EditType;
using Cfg = hashdag::Config<uint32_t>;
using Coord = hashdag::NodeCoord<uint32_t>;
using NPtr = hashdag::NodePointer<uint32_t>;

namespace ref_phmap { // as src/DAGNodePool.hpp:26-29
template <typename K, typename V> using flat_hash_map = phmap::flat_hash_map<K, V>;
template <typename K> using flat_hash_set = phmap::flat_hash_set<K>;
} // namespace ref_phmap

// ------------------------------------------------------------------------------------------------ pool
struct RefPool final : public hashdag::NodePoolBase<RefPool, uint32_t>,
                       public hashdag::NodePoolTraversal<RefPool, uint32_t>,
                       public hashdag::NodePoolThreadedEdit<RefPool, uint32_t>,
                       public hashdag::NodePoolThreadedGC<RefPool, uint32_t, ref_phmap::flat_hash_map, ref_phmap::flat_hash_set> {
	using WordSpanHasher = hashdag::MurmurHasher32;
	uint32_t *memory = nullptr;
	uint64_t total_words = 0;
	std::vector<uint32_t> bucket_words;
	std::array<std::mutex, 1024> mutexes;

	explicit RefPool(const Cfg &cfg) : hashdag::NodePoolBase<RefPool, uint32_t>(cfg) {
		uint64_t buckets = 0;
		for (uint32_t b : cfg.bucket_bits_each_level)
			buckets += 1ull << b;
		total_words = buckets << cfg.GetWordBitsPerBucket();
		memory = static_cast<uint32_t *>(mmap(nullptr, total_words * 4, PROT_READ | PROT_WRITE,
		                                      MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0));
		bucket_words.assign(buckets, 0);
	}
	~RefPool() final { munmap(memory, total_words * 4); }

	std::mutex &GetBucketRefMutex(uint32_t bucket) { return mutexes[bucket % mutexes.size()]; }
	uint32_t &GetBucketRefWords(uint32_t bucket) { return bucket_words[bucket]; }
	const uint32_t *ReadPage(uint32_t page) const { return memory + (uint64_t(page) << GetConfig().word_bits_per_page); }
	void ZeroPage(uint32_t page, uint32_t off, uint32_t n) {
		uint32_t *p = memory + (uint64_t(page) << GetConfig().word_bits_per_page) + off;
		std::fill(p, p + n, 0u);
	}
	void WritePage(uint32_t page, uint32_t off, std::span<const uint32_t> w) {
		std::copy(w.begin(), w.end(), memory + (uint64_t(page) << GetConfig().word_bits_per_page) + off);
	}
	void FreePage(uint32_t) {} // GCNodePool concept (NodePool.hpp:37-38); the flat mapping keeps its pages
};

// --------------------------------------------------------------------------------------------- editors
struct BoxU {
	glm::u32vec3 lb, ub;
};
static inline BoxU bounds(const Cfg &config, const Coord &coord) {
	return {coord.GetLowerBoundAtLevel(config.GetVoxelLevel()), coord.GetUpperBoundAtLevel(config.GetVoxelLevel())};
}

struct AABBEd { // semantics of src/main.cpp:32-70
	glm::u32vec3 lo, hi;
	hashdag::VBRColor color;
	EditType EditNode(const Cfg &config, const Coord &coord, NPtr) const {
		BoxU b = bounds(config, coord);
		if (glm::any(glm::lessThanEqual(b.ub, lo)) || glm::any(glm::greaterThanEqual(b.lb, hi)))
			return EditType::kNotAffected;
		if (glm::all(glm::greaterThanEqual(b.lb, lo)) && glm::all(glm::lessThanEqual(b.ub, hi)))
			return EditType::kFill;
		return EditType::kProceed;
	}
	EditType EditNode(const Cfg &config, const Coord &coord, NPtr ptr, hashdag::VBRColor &final_color) const {
		EditType t = EditNode(config, coord, {});
		final_color = (t == EditType::kFill || !ptr || final_color == color) ? color : hashdag::VBRColor{};
		return t;
	}
	bool In(const Coord &c) const {
		return glm::all(glm::greaterThanEqual(c.pos, lo)) && glm::all(glm::lessThan(c.pos, hi));
	}
	bool EditVoxel(const Cfg &, const Coord &c, bool voxel) const { return voxel || In(c); }
	bool EditVoxel(const Cfg &, const Coord &c, bool voxel, hashdag::VBRColor &col) const {
		bool in = In(c);
		col = in || !voxel ? color : col;
		return voxel || in;
	}
};

enum class Mode { kFill, kDig, kPaint };
template <Mode M> struct SphereEd { // semantics of src/main.cpp:72-150
	glm::u32vec3 center;
	uint64_t r2;
	hashdag::VBRColor color;
	EditType EditNode(const Cfg &config, const Coord &coord, NPtr) const {
		BoxU b = bounds(config, coord);
		glm::i64vec3 lo = glm::i64vec3(b.lb) - glm::i64vec3(center), hi = glm::i64vec3(b.ub) - glm::i64vec3(center);
		glm::u64vec3 lo2 = lo * lo, hi2 = hi * hi;
		glm::u64vec3 mx = glm::max(lo2, hi2);
		if (mx.x + mx.y + mx.z <= r2)
			return M == Mode::kDig ? EditType::kClear : EditType::kFill;
		uint64_t mn = 0;
		for (int i = 0; i < 3; ++i) {
			if (lo[i] > 0)
				mn += lo2[i];
			if (hi[i] < 0)
				mn += hi2[i];
		}
		return mn > r2 ? EditType::kNotAffected : EditType::kProceed;
	}
	EditType EditNode(const Cfg &config, const Coord &coord, NPtr ptr, hashdag::VBRColor &final_color) const {
		static_assert(M != Mode::kDig);
		EditType t = EditNode(config, coord, {});
		if (t == EditType::kFill) {
			final_color = color;
			if constexpr (M == Mode::kPaint)
				t = EditType::kNotAffected;
		} else if (!ptr || final_color == color)
			final_color = color;
		else
			final_color = {};
		if constexpr (M == Mode::kPaint)
			if (!ptr)
				t = EditType::kNotAffected;
		return t;
	}
	bool In(const Coord &c) const {
		glm::i64vec3 d = glm::i64vec3(c.pos) - glm::i64vec3(center);
		return uint64_t(d.x * d.x + d.y * d.y + d.z * d.z) <= r2;
	}
	bool EditVoxel(const Cfg &, const Coord &c, bool voxel) const {
		if constexpr (M == Mode::kPaint)
			return voxel;
		bool in = In(c);
		return M == Mode::kFill ? (voxel || in) : (voxel && !in);
	}
	bool EditVoxel(const Cfg &, const Coord &c, bool voxel, hashdag::VBRColor &col) const {
		static_assert(M != Mode::kDig);
		bool in = In(c);
		col = in || !voxel ? color : col;
		return M == Mode::kFill ? voxel || in : voxel;
	}
};

struct TerrainEd { // synthetic scene generator, oracle/terrain.h
	terrain::Params tp;
	EditType EditNode(const Cfg &config, const Coord &coord, NPtr) const {
		BoxU b = bounds(config, coord);
		const int ext = terrain::extent_class(tp, b.lb.x, b.lb.z, config.GetVoxelLevel() - coord.level);
		if (ext == 0)
			return EditType::kNotAffected;
		uint32_t hmin, hmax;
		terrain::height_bounds(tp, b.lb.x, b.lb.z, config.GetVoxelLevel() - coord.level, hmin, hmax);
		if (b.lb.y >= hmax)
			return EditType::kNotAffected;
		if (ext == 2 && b.ub.y <= hmin)
			return EditType::kFill;
		return EditType::kProceed;
	}
	bool EditVoxel(const Cfg &, const Coord &c, bool voxel) const {
		struct Cache {
			uint32_t x[16], z[16], h[16], valid = 0;
			const TerrainEd *owner = nullptr;
			uint32_t seed = 0;
		};
		thread_local Cache cache;
		if (cache.owner != this || cache.seed != tp.seed)
			cache.valid = 0, cache.owner = this, cache.seed = tp.seed;
		uint32_t s = (c.pos.x & 3u) | ((c.pos.z & 3u) << 2);
		uint32_t h;
		if ((cache.valid >> s & 1u) && cache.x[s] == c.pos.x && cache.z[s] == c.pos.z)
			h = cache.h[s];
		else {
			h = terrain::height(tp, c.pos.x, c.pos.z);
			cache.valid |= 1u << s, cache.x[s] = c.pos.x, cache.z[s] = c.pos.z, cache.h[s] = h;
		}
		return voxel || (terrain::in_extent(tp, c.pos.x, c.pos.z) && c.pos.y < h);
	}
};

// ----------------------------------------------------------------------------------------- colour pool
template <typename T> struct View { // VBRContainer over memory that never moves
	const T *p = nullptr;
	size_t n = 0;
	const T &operator[](size_t i) const { return p[i]; }
	size_t size() const { return n; }
	bool empty() const { return n == 0; }
};

struct RefColorPool {
	struct Pointer { // src/DAGColorPool.hpp:23-34
		enum class Tag { kNode = 0, kColor, kLeaf, kNull };
		uint32_t pointer;
		Pointer() : Pointer(Tag::kNull, 0u) {}
		Pointer(Tag tag, uint32_t data) : pointer{(uint32_t(tag) << 30u) | data} {}
		Tag GetTag() const { return Tag(pointer >> 30u); }
		uint32_t GetData() const { return pointer & 0x3FFFFFFFu; }
		bool operator==(const Pointer &r) const { return pointer == r.pointer; }
	};
	using Node = std::array<Pointer, 8>;

	uint32_t leaf_level;
	// fixed-capacity arenas (addresses must stay valid while writers hold views)
	uint32_t *nodes = nullptr, *leaves = nullptr;
	uint64_t node_cap_words, leaf_cap_words;
	std::atomic<uint64_t> node_count{0}, leaf_words{0};
	std::mutex mtx;
	Pointer root{};

	RefColorPool(uint32_t leaf_level_, uint64_t node_cap, uint64_t leaf_cap)
	    : leaf_level(leaf_level_), node_cap_words(node_cap * 8), leaf_cap_words(leaf_cap) {
		nodes = static_cast<uint32_t *>(mmap(nullptr, node_cap_words * 4, PROT_READ | PROT_WRITE,
		                                     MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0));
		leaves = static_cast<uint32_t *>(mmap(nullptr, leaf_cap_words * 4, PROT_READ | PROT_WRITE,
		                                      MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0));
	}
	~RefColorPool() {
		munmap(nodes, node_cap_words * 4);
		munmap(leaves, leaf_cap_words * 4);
	}

	Pointer GetChild(Pointer ptr, auto idx) const { // DAGColorPool.hpp:139-143
		if (ptr.GetTag() == Pointer::Tag::kNode) {
			Pointer c;
			c.pointer = nodes[(uint64_t(ptr.GetData()) << 3) | uint32_t(idx)];
			return c;
		}
		return ptr.GetTag() == Pointer::Tag::kColor ? ptr : Pointer{};
	}
	static hashdag::VBRColor GetFill(Pointer ptr) { // :144-146
		return ptr.GetTag() == Pointer::Tag::kColor ? hashdag::VBRColor{hashdag::RGB8Color{ptr.GetData()}}
		                                            : hashdag::VBRColor{};
	}
	Pointer SetNode(Pointer ptr, std::span<const Pointer, 8> ch) { // :147-163
		bool all_null = true, all_same_color = ch[0].GetTag() == Pointer::Tag::kColor;
		for (const Pointer &c : ch) {
			all_null &= c.GetTag() == Pointer::Tag::kNull;
			all_same_color &= c == ch[0];
		}
		if (all_null)
			return {};
		if (all_same_color)
			return ch[0];
		if (ptr.GetTag() == Pointer::Tag::kNode) {
			const uint32_t *cur = nodes + (uint64_t(ptr.GetData()) << 3);
			bool same = true;
			for (int i = 0; i < 8; ++i)
				same &= cur[i] == ch[i].pointer;
			if (same)
				return ptr;
		}
		uint64_t id = node_count.fetch_add(1);
		if ((id + 1) * 8 > node_cap_words)
			return ptr;
		for (int i = 0; i < 8; ++i)
			nodes[(id << 3) | i] = ch[i].pointer;
		return Pointer{Pointer::Tag::kNode, uint32_t(id)};
	}
	static Pointer ClearNode(Pointer) { return Pointer{}; }
	static Pointer FillNode(Pointer, hashdag::VBRColor color) { // :165-167
		return Pointer{Pointer::Tag::kColor, hashdag::RGB8Color{color.Get()}.GetData()};
	}

	hashdag::VBRChunk<uint32_t, View> fetch(uint64_t idx) const { // fetch_leaf_chunk, :98-113
		uint32_t macro = leaves[idx], blocks = leaves[idx + 1], bitw = leaves[idx + 2];
		idx += 3;
		View<hashdag::VBRMacroBlock> mv{reinterpret_cast<const hashdag::VBRMacroBlock *>(leaves + idx), macro};
		idx += uint64_t(macro) * 2;
		View<hashdag::VBRBlockHeader> bv{reinterpret_cast<const hashdag::VBRBlockHeader *>(leaves + idx), blocks};
		idx += uint64_t(blocks) * 2;
		View<uint32_t> wv{leaves + idx, bitw};
		return hashdag::VBRChunk<uint32_t, View>{mv, bv, hashdag::VBRBitset<uint32_t, View>{wv}};
	}
	void store(uint64_t idx, const hashdag::VBRChunk<uint32_t, hashdag::VBRWriterContainer> &c) { // write_leaf_chunk :66-96
		leaves[idx++] = uint32_t(c.GetMacroBlocks().size());
		leaves[idx++] = uint32_t(c.GetBlockHeaders().size());
		leaves[idx++] = uint32_t(c.GetWeightBits().GetWords().size());
		for (const auto &m : c.GetMacroBlocks())
			leaves[idx++] = m.first_block, leaves[idx++] = m.weight_start;
		for (const auto &b : c.GetBlockHeaders())
			leaves[idx++] = b.colors, leaves[idx++] = b.packed_14_2_16;
		for (uint32_t w : c.GetWeightBits().GetWords())
			leaves[idx++] = w;
	}
	hashdag::VBRChunk<uint32_t, View> GetLeaf(Pointer ptr) const { // :169-172
		return ptr.GetTag() == Pointer::Tag::kLeaf ? fetch(uint64_t(ptr.GetData()) + 1) : hashdag::VBRChunk<uint32_t, View>{};
	}
	Pointer SetLeaf(Pointer ptr, hashdag::VBRChunk<uint32_t, hashdag::VBRWriterContainer> &&chunk) { // :173-204
		uint64_t data = chunk.GetMacroBlocks().size() * 2 + chunk.GetBlockHeaders().size() * 2 +
		                chunk.GetWeightBits().GetWords().size() + 4;
		uint64_t append = (data & 1) ? data + 1 : data;
		if (ptr.GetTag() == Pointer::Tag::kLeaf) { // keep_history = false (main.cpp:209)
			uint64_t idx = ptr.GetData(), block = leaves[idx];
			if (data <= block) {
				store(idx + 1, chunk);
				return ptr;
			}
			append = std::max(block << 1, append);
		}
		uint64_t idx = leaf_words.fetch_add(append);
		if (idx + append > leaf_cap_words || idx + append >= (1ull << 30))
			return ptr;
		leaves[idx] = uint32_t(append);
		store(idx + 1, chunk);
		return Pointer{Pointer::Tag::kLeaf, uint32_t(idx)};
	}
	uint32_t GetLeafLevel() const { return leaf_level; }
};
static_assert(hashdag::VBROctree<RefColorPool, uint32_t>);

// ------------------------------------------------------------------------------------------ C wrapper
struct ref_pool {
	std::unique_ptr<RefPool> pool;
};
struct ref_color_pool {
	std::unique_ptr<RefColorPool> pool;
};

static lf::busy_pool *get_busy_pool(uint32_t n) {
	static std::mutex m;
	static std::map<uint32_t, std::unique_ptr<lf::busy_pool>> pools;
	std::lock_guard<std::mutex> lock(m);
	auto &p = pools[n];
	if (!p)
		p = std::make_unique<lf::busy_pool>(n);
	return p.get();
}

template <typename Ed> static uint32_t run_stateless(RefPool &pool, uint32_t root, Ed ed, uint32_t threads, uint32_t max_task_level) {
	hashdag::StatelessEditorWrapper<uint32_t, Ed> w{.editor = ed};
	if (threads == 0)
		return *pool.Edit(NPtr{root}, w); // NodePool.hpp:415
	return *pool.ThreadedEdit(get_busy_pool(threads), NPtr{root}, w, max_task_level); // NodePoolThreadedEdit.hpp:121
}

static Cfg to_cfg(const hd_config *c) {
	Cfg cfg;
	cfg.word_bits_per_page = c->word_bits_per_page;
	cfg.page_bits_per_bucket = c->page_bits_per_bucket;
	cfg.bucket_bits_each_level.assign(c->bucket_bits_each_level, c->bucket_bits_each_level + c->node_levels);
	return cfg;
}

extern "C" {

uint32_t ref_hash_inner(const uint32_t *w, uint32_t n) { return hashdag::MurmurHasher32{}(std::span<const uint32_t>(w, n)); }
uint32_t ref_hash_leaf(const uint32_t *w) { return hashdag::MurmurHasher32{}(std::span<const uint32_t, 2>(w, 2)); }

int ref_config_from_default(const hd_default_config *dc, hd_config *out) {
	hashdag::DefaultConfig<uint32_t> d;
	d.level_count = dc->level_count, d.top_level_count = dc->top_level_count;
	d.word_bits_per_page = dc->word_bits_per_page, d.page_bits_per_bucket = dc->page_bits_per_bucket;
	d.bucket_bits_per_top_level = dc->bucket_bits_per_top_level;
	d.bucket_bits_per_bottom_level = dc->bucket_bits_per_bottom_level;
	Cfg cfg = d();
	std::memset(out, 0, sizeof(*out));
	out->word_bits_per_page = cfg.word_bits_per_page, out->page_bits_per_bucket = cfg.page_bits_per_bucket;
	out->node_levels = cfg.GetNodeLevels();
	for (uint32_t l = 0; l < cfg.GetNodeLevels(); ++l)
		out->bucket_bits_each_level[l] = cfg.bucket_bits_each_level[l];
	return Cfg::Validate(cfg) ? 0 : 1;
}
// geometry: [node_levels, words/page, words/bucket, total buckets, total pages, total words, level bases...]
void ref_config_geometry(const hd_config *c, uint64_t *out) {
	Cfg cfg = to_cfg(c);
	out[0] = cfg.GetNodeLevels(), out[1] = cfg.GetWordsPerPage(), out[2] = cfg.GetWordsPerBucket();
	out[3] = cfg.GetTotalBuckets(), out[4] = cfg.GetTotalPages(), out[5] = cfg.GetTotalWords();
	auto bases = cfg.GetLevelBaseBucketIndices();
	for (size_t i = 0; i < bases.size(); ++i)
		out[6 + i] = bases[i];
}

ref_pool *ref_pool_create(const hd_config *c) {
	Cfg cfg = to_cfg(c);
	if (!Cfg::Validate(cfg))
		return nullptr;
	auto *p = new ref_pool();
	p->pool = std::make_unique<RefPool>(cfg);
	return p;
}
void ref_pool_destroy(ref_pool *p) { delete p; }
uint32_t *ref_pool_words(ref_pool *p) { return p->pool->memory; }
uint32_t *ref_pool_bucket_words(ref_pool *p) { return p->pool->bucket_words.data(); }
uint64_t ref_pool_total_words(ref_pool *p) { return p->pool->total_words; }
uint32_t ref_pool_total_buckets(ref_pool *p) { return uint32_t(p->pool->bucket_words.size()); }

uint32_t ref_upsert(ref_pool *p, uint32_t level, const uint32_t *node, uint32_t n, uint32_t fallback) {
	RefPool &pool = *p->pool;
	if (level == pool.GetConfig().GetNodeLevels() - 1 && n == 2)
		return *pool.upsert_leaf<false>(level, std::span<const uint32_t, 2>(node, 2), NPtr{fallback});
	return *pool.upsert_inner_node<false>(level, std::span<const uint32_t>(node, n), NPtr{fallback});
}
void ref_filled_nodes(ref_pool *p, uint32_t *out) {
	p->pool->make_filled_node_pointers();
	for (size_t i = 0; i < p->pool->m_filled_node_pointers.size(); ++i)
		out[i] = *p->pool->m_filled_node_pointers[i];
}

// threads == 0: serial Edit; otherwise ThreadedEdit with lf::busy_pool(threads) (main.cpp:159,214-216)
uint32_t ref_edit(ref_pool *p, uint32_t root, const hd_edit_desc *d, uint32_t threads, uint32_t max_task_level) {
	RefPool &pool = *p->pool;
	glm::u32vec3 a{d->p0[0], d->p0[1], d->p0[2]}, b{d->p1[0], d->p1[1], d->p1[2]};
	switch (d->kind) {
	case HD_EDIT_AABB_FILL:
		return run_stateless(pool, root, AABBEd{a, b, {}}, threads, max_task_level);
	case HD_EDIT_SPHERE_FILL:
		return run_stateless(pool, root, SphereEd<Mode::kFill>{a, d->r2, {}}, threads, max_task_level);
	case HD_EDIT_SPHERE_DIG:
		return run_stateless(pool, root, SphereEd<Mode::kDig>{a, d->r2, {}}, threads, max_task_level);
	case HD_EDIT_TERRAIN_FILL:
		return run_stateless(pool, root, TerrainEd{terrain::from_desc(d->aux, d->p0, d->p1)}, threads,
		                     max_task_level);
	}
	return root;
}
uint32_t ref_edit_batch(ref_pool *p, uint32_t root, const hd_edit_desc *d, uint32_t n, uint32_t threads,
                        uint32_t max_task_level) {
	for (uint32_t i = 0; i < n; ++i)
		root = ref_edit(p, root, d + i, threads, max_task_level);
	return root;
}

// NodePoolThreadedGC::ThreadedGC (NodePoolThreadedGC.hpp:394-397): returns the relocated root
uint32_t ref_gc(ref_pool *p, uint32_t root, uint32_t threads) {
	return *p->pool->ThreadedGC(get_busy_pool(threads ? threads : 1), NPtr{root});
}

int ref_traverse(ref_pool *p, uint32_t root, const float o[3], const float d[3], float out[3]) {
	auto r = p->pool->Traversal<float>(NPtr{root}, glm::vec3(o[0], o[1], o[2]), glm::vec3(d[0], d[1], d[2]));
	if (!r)
		return 0;
	out[0] = r->x, out[1] = r->y, out[2] = r->z;
	return 1;
}

// Primary rays of a frame through the reference's own host tracer (Traversal<float>: no LOD, no colour).
// Rays are generated as shader/src/trace.frag:366-377 does.  hit: u8 per pixel, pos: 3 floats per pixel (may be NULL).
uint64_t ref_trace_frame_host(ref_pool *p, const hd_trace_params *P, uint32_t row_begin, uint32_t row_end,
                              uint32_t row_step, uint32_t n_threads, uint8_t *hit, float *pos) {
	RefPool &pool = *p->pool;
	std::atomic<uint64_t> hits{0};
	if (!n_threads)
		n_threads = 1;
	if (!row_step)
		row_step = 1;
	glm::vec3 o(P->pos[0], P->pos[1], P->pos[2]), look(P->look[0], P->look[1], P->look[2]),
	    side(P->side[0], P->side[1], P->side[2]), up(P->up[0], P->up[1], P->up[2]);
	auto work = [&](uint32_t tid) {
		uint64_t h = 0;
		uint32_t k = 0;
		for (uint32_t y = row_begin; y < row_end; y += row_step, ++k) {
			if (k % n_threads != tid)
				continue;
			for (uint32_t x = 0; x < P->width; ++x) {
				glm::vec2 c = glm::vec2(float(x) + 0.5f, float(y) + 0.5f) / glm::vec2(float(P->width), float(P->height));
				c = c * 2.0f - 1.0f;
				glm::vec3 d = glm::normalize(look - side * c.x - up * c.y);
				auto r = pool.Traversal<float>(NPtr{P->dag_root}, o, d);
				size_t at = size_t(y) * P->width + x;
				if (hit)
					hit[at] = r ? 1 : 0;
				if (pos && r)
					pos[at * 3] = r->x, pos[at * 3 + 1] = r->y, pos[at * 3 + 2] = r->z;
				h += r ? 1 : 0;
			}
		}
		hits += h;
	};
	if (n_threads == 1)
		work(0);
	else {
		std::vector<std::thread> th;
		for (uint32_t t = 0; t < n_threads; ++t)
			th.emplace_back(work, t);
		for (auto &t : th)
			t.join();
	}
	return hits.load();
}

// ---- colour ----
ref_color_pool *ref_color_pool_create(uint32_t leaf_level, uint64_t node_capacity, uint64_t leaf_word_capacity) {
	auto *c = new ref_color_pool();
	c->pool = std::make_unique<RefColorPool>(leaf_level, node_capacity, leaf_word_capacity);
	return c;
}
void ref_color_pool_destroy(ref_color_pool *c) { delete c; }
uint32_t ref_color_root(ref_color_pool *c) { return c->pool->root.pointer; }
uint32_t *ref_color_nodes(ref_color_pool *c) { return c->pool->nodes; }
uint32_t *ref_color_leaves(ref_color_pool *c) { return c->pool->leaves; }
uint64_t ref_color_node_words(ref_color_pool *c) { return c->pool->node_count.load() * 8; }
uint64_t ref_color_leaf_words(ref_color_pool *c) { return c->pool->leaf_words.load(); }

// Coloured edit through VBREditorWrapper (VBREditor.hpp:26-107), serial Edit.  paint != 0 selects
// SphereEditor<kPaint>.  Updates the colour root (main.cpp:239-243) and returns the new node root.
uint32_t ref_edit_color_mt(ref_pool *p, ref_color_pool *c, uint32_t root, const hd_edit_desc *d, uint32_t rgb8, int paint,
                           uint32_t threads);
uint32_t ref_edit_color(ref_pool *p, ref_color_pool *c, uint32_t root, const hd_edit_desc *d, uint32_t rgb8, int paint) {
	return ref_edit_color_mt(p, c, root, d, rgb8, paint, 0);
}
// threads != 0: ThreadedEdit with max_task_level = the colour leaf level, exactly as src/main.cpp:214-223 calls it
uint32_t ref_edit_color_mt(ref_pool *p, ref_color_pool *c, uint32_t root, const hd_edit_desc *d, uint32_t rgb8, int paint,
                           uint32_t threads) {
	RefPool &pool = *p->pool;
	RefColorPool &cp = *c->pool;
	hashdag::VBRColor color{hashdag::RGB8Color{rgb8}};
	glm::u32vec3 a{d->p0[0], d->p0[1], d->p0[2]}, b{d->p1[0], d->p1[1], d->p1[2]};
	auto run = [&](auto ed) -> uint32_t {
		using Ed = decltype(ed);
		hashdag::VBREditorWrapper<uint32_t, Ed, RefColorPool> w{.editor = ed, .p_octree = &cp, .octree_root = cp.root};
		auto done = [&](NPtr new_root, auto &&state) -> uint32_t {
			cp.root = state.octree_node;
			return *new_root;
		};
		if (threads)
			return pool.ThreadedEdit(get_busy_pool(threads), NPtr{root}, w, cp.GetLeafLevel(), done);
		return pool.Edit(NPtr{root}, w, done);
	};
	if (d->kind == HD_EDIT_AABB_FILL)
		return run(AABBEd{a, b, color});
	if (d->kind == HD_EDIT_SPHERE_FILL && paint)
		return run(SphereEd<Mode::kPaint>{a, d->r2, color});
	if (d->kind == HD_EDIT_SPHERE_FILL)
		return run(SphereEd<Mode::kFill>{a, d->r2, color});
	return root;
}

// Colour of voxel (x,y,z) through the reference's own decoder (VBRChunkIterator, VBRColor.hpp:292-366).
// Returns 0 when the octree holds no colour there.
int ref_color_at(ref_color_pool *c, uint32_t voxel_level, uint32_t x, uint32_t y, uint32_t z, float out[3]) {
	RefColorPool &cp = *c->pool;
	using P = RefColorPool::Pointer;
	P ptr = cp.root;
	for (uint32_t l = 0; l < cp.leaf_level; ++l) {
		if (ptr.GetTag() != P::Tag::kNode)
			break;
		uint32_t sh = voxel_level - 1 - l;
		ptr = cp.GetChild(ptr, ((x >> sh) & 1u) | (((y >> sh) & 1u) << 1) | (((z >> sh) & 1u) << 2));
	}
	if (ptr.GetTag() == P::Tag::kColor) {
		glm::vec3 v = RefColorPool::GetFill(ptr).Get();
		out[0] = v.x, out[1] = v.y, out[2] = v.z;
		return 1;
	}
	if (ptr.GetTag() != P::Tag::kLeaf)
		return 0;
	uint32_t bits = voxel_level - cp.leaf_level, m = (1u << bits) - 1u;
	uint32_t sx = x & m, sy = y & m, sz = z & m, idx = 0;
	for (uint32_t b = 0; b < bits; ++b) // Morton order of the edit recursion
		idx |= (((sx >> b) & 1u) << (3 * b)) | (((sy >> b) & 1u) << (3 * b + 1)) | (((sz >> b) & 1u) << (3 * b + 2));
	hashdag::VBRChunkIterator<uint32_t, View> it{cp.GetLeaf(ptr)};
	it.Jump(idx);
	glm::vec3 v = it.GetColor().Get();
	out[0] = v.x, out[1] = v.y, out[2] = v.z;
	return 1;
}

} // extern "C"
