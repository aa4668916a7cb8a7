// hashdag_b200.hpp — C++ host layer over the C ABI (hashdag_b200.h), header-only, C++17.
//
// Mirrors the operator interface of the reference's include/hashdag for the traversal + edit path so call sites
// read the same (names, argument meaning, Null/NodePointer conventions, no exceptions):
//   Config / DefaultConfig            include/hashdag/Config.hpp:15-75
//   NodePointer / NodeCoord           include/hashdag/NodePointer.hpp:13-31, NodeCoord.hpp:14-94
//   EditType / IterateType            include/hashdag/Editor.hpp:18, test/test.cpp:36-52
//   AABBEditor / SphereEditor<Mode>   src/main.cpp:32-150 (as POD descriptors: device predicates are compiled in)
//   DAGNodePool::{Create, GetConfig, Edit, ThreadedEdit, ThreadedGC, Traversal<float>, Iterate, SetRoot, GetRoot, Flush}
//                                     src/DAGNodePool.hpp:84-99, NodePool.hpp:404-417, NodePoolThreadedEdit.hpp:104-126,
//                                     NodePoolTraversal.hpp:93-256
// All compute happens in libhashdag_b200.so on the GPU; this header only marshals.
#pragma once
#include "hashdag_b200.h"

#include <algorithm>
#include <array>
#include <cstdint>
#include <memory>
#include <optional>
#include <unordered_map>
#include <utility>
#include <variant>
#include <vector>

namespace hashdag_b200 {

struct UVec3 {
	uint32_t x{}, y{}, z{};
	bool operator==(const UVec3 &r) const { return x == r.x && y == r.y && z == r.z; }
};
struct Vec3 {
	float x{}, y{}, z{};
};

template <typename Word = uint32_t> struct Config { // include/hashdag/Config.hpp:15-57
	Word word_bits_per_page{};
	Word page_bits_per_bucket{};
	std::vector<Word> bucket_bits_each_level;

	Word GetWordsPerPage() const { return Word(1u) << word_bits_per_page; }
	Word GetPagesPerBucket() const { return Word(1u) << page_bits_per_bucket; }
	Word GetWordsPerBucket() const { return Word(1u) << (word_bits_per_page + page_bits_per_bucket); }
	Word GetBucketsAtLevel(Word level) const { return Word(1u) << bucket_bits_each_level[level]; }
	Word GetNodeLevels() const { return Word(bucket_bits_each_level.size()); }
	Word GetLeafLevel() const { return GetNodeLevels(); }
	Word GetVoxelLevel() const { return GetNodeLevels() + 1u; }
	Word GetResolution() const { return Word(1u) << GetVoxelLevel(); }
	hd_config ToC() const {
		hd_config c{};
		c.word_bits_per_page = word_bits_per_page, c.page_bits_per_bucket = page_bits_per_bucket;
		c.node_levels = GetNodeLevels();
		for (Word l = 0; l < GetNodeLevels() && l < HD_MAX_NODE_LEVELS; ++l)
			c.bucket_bits_each_level[l] = bucket_bits_each_level[l];
		return c;
	}
	Word GetTotalBuckets() const {
		hd_config c = ToC();
		return hd_config_total_buckets(&c);
	}
	uint64_t GetTotalWords() const {
		hd_config c = ToC();
		return hd_config_total_words(&c);
	}
	static bool Validate(const Config &config) {
		if (config.GetNodeLevels() > HD_MAX_NODE_LEVELS)
			return false;
		hd_config c = config.ToC();
		return hd_config_validate(&c) != 0;
	}
};

template <typename Word = uint32_t> struct DefaultConfig { // Config.hpp:59-75
	uint32_t level_count = 17, top_level_count = 9;
	Word word_bits_per_page = 9, page_bits_per_bucket = 2, bucket_bits_per_top_level = 10,
	     bucket_bits_per_bottom_level = 16;
	Config<Word> operator()() const {
		Config<Word> c;
		c.word_bits_per_page = word_bits_per_page, c.page_bits_per_bucket = page_bits_per_bucket;
		for (uint32_t l = 0; l + 1 < level_count; ++l)
			c.bucket_bits_each_level.push_back(l < top_level_count ? bucket_bits_per_top_level
			                                                       : bucket_bits_per_bottom_level);
		return c;
	}
};

template <typename Word = uint32_t> class NodePointer { // NodePointer.hpp:13-31
	Word m_node;

public:
	constexpr NodePointer() : m_node(Word(-1)) {}
	constexpr NodePointer(Word node) : m_node{node} {}
	constexpr bool HasValue() const { return m_node != Word(-1); }
	constexpr explicit operator bool() const { return HasValue(); }
	constexpr bool operator==(NodePointer r) const { return m_node == r.m_node; }
	constexpr bool operator!=(NodePointer r) const { return m_node != r.m_node; }
	constexpr Word Value() const { return m_node; }
	constexpr Word operator*() const { return m_node; }
	constexpr static NodePointer Null() { return NodePointer{}; }
};

template <typename Word = uint32_t> struct NodeCoord { // NodeCoord.hpp:14-94 (integer part)
	Word level{};
	UVec3 pos{};
	NodeCoord GetChildCoord(Word i) const {
		return {level + 1, {(pos.x << 1) | (i & 1u), (pos.y << 1) | ((i >> 1) & 1u), (pos.z << 1) | ((i >> 2) & 1u)}};
	}
	NodeCoord GetLeafCoord(Word i) const {
		return {level + 2,
		        {(pos.x << 2) | ((i >> 2) & 2u) | (i & 1u), (pos.y << 2) | ((i >> 3) & 2u) | ((i >> 1) & 1u),
		         (pos.z << 2) | ((i >> 4) & 2u) | ((i >> 2) & 1u)}};
	}
	UVec3 GetLowerBoundAtLevel(Word at) const { return {pos.x << (at - level), pos.y << (at - level), pos.z << (at - level)}; }
	UVec3 GetUpperBoundAtLevel(Word at) const {
		return {(pos.x + 1) << (at - level), (pos.y + 1) << (at - level), (pos.z + 1) << (at - level)};
	}
};

enum class EditType { kNotAffected, kProceed, kFill, kClear }; // Editor.hpp:18
enum class IterateType { kProceed, kStop };                    // test/test.cpp:40-45
enum class EditMode { kFill, kDig, kPaint };                   // main.cpp:72

// ---- editors: the reference's structs reduced to their parameters (src/main.cpp:32-150) ----
// `color` is the editor's hashdag::VBRColor as RGB8 (r in the low byte); it only matters when the editor is applied through
// VBREditorWrapper (a colour-aware edit), exactly like in the reference.
struct AABBEditor {
	UVec3 aabb_min, aabb_max;
	uint32_t color = 0;
	static constexpr bool kPaint = false;
	hd_edit_desc Desc() const {
		hd_edit_desc d{};
		d.kind = HD_EDIT_AABB_FILL;
		d.p0[0] = aabb_min.x, d.p0[1] = aabb_min.y, d.p0[2] = aabb_min.z;
		d.p1[0] = aabb_max.x, d.p1[1] = aabb_max.y, d.p1[2] = aabb_max.z;
		return d;
	}
};
template <EditMode Mode = EditMode::kFill> struct SphereEditor {
	UVec3 center{};
	uint64_t r2{};
	uint32_t color = 0;
	static constexpr bool kPaint = Mode == EditMode::kPaint;
	static constexpr EditMode kMode = Mode;
	hd_edit_desc Desc() const {
		hd_edit_desc d{};
		d.kind = Mode == EditMode::kDig ? HD_EDIT_SPHERE_DIG : HD_EDIT_SPHERE_FILL; // kPaint: the sphere shape, paint flag set by the wrapper
		d.p0[0] = center.x, d.p0[1] = center.y, d.p0[2] = center.z;
		d.r2 = r2;
		return d;
	}
};

// Editor wrappers with the reference's names and NodeState members (Editor.hpp:42-57, VBREditor.hpp:26-36).  The state the
// reference hands to on_edit_done(root, state) is the ROOT node's state: std::monostate for a stateless edit, the colour
// octree pointer (`octree_node`) for a colour-aware one (src/main.cpp:214-238 builds its EditResult from exactly that).
template <typename Editor_T> struct StatelessEditorWrapper {
	Editor_T editor;
	using NodeState = std::monostate;
};
struct VBRNodeState {
	uint32_t octree_node = HD_COLOR_NULL; // DAGColorPool::Pointer of the new colour root (tag << 30 | data)
};
template <typename Editor_T> struct VBREditorWrapper {
	Editor_T editor;
	using NodeState = VBRNodeState;
};
struct TerrainEditor { // synthetic scene generator
	uint32_t seed, base, first_cell_bits, octaves, first_amplitude, extent_bits;
	static constexpr bool kPaint = false;
	static TerrainEditor ForLevel(uint32_t voxel_level, uint32_t seed = 0x5EED, uint32_t octaves = 4, uint32_t amp_div = 8,
	                              uint32_t extent_bits = 0) {
		const uint32_t bits = extent_bits ? extent_bits : voxel_level, ext = 1u << bits;
		return {seed, ext / 4, bits - 2, octaves, ext / amp_div, extent_bits};
	}
	hd_edit_desc Desc() const {
		hd_edit_desc d{};
		d.kind = HD_EDIT_TERRAIN_FILL, d.aux = seed;
		d.p0[0] = base, d.p0[1] = first_cell_bits, d.p0[2] = octaves;
		d.p1[0] = first_amplitude, d.p1[1] = extent_bits;
		return d;
	}
};

class DAGNodePool {
	hd_pool *m_pool{};
	Config<uint32_t> m_config;
	hd_edit_stats m_last_stats{};
	hd_status m_last_status{HD_OK};

	explicit DAGNodePool(Config<uint32_t> config) : m_config{std::move(config)} {}

public:
	DAGNodePool(const DAGNodePool &) = delete;
	DAGNodePool &operator=(const DAGNodePool &) = delete;
	~DAGNodePool() { hd_pool_destroy(m_pool); }

	// DAGNodePool::Create (src/DAGNodePool.cpp:9-47): nullptr on failure, like the reference
	static std::unique_ptr<DAGNodePool> Create(Config<uint32_t> config, int device = 0) {
		if (!Config<uint32_t>::Validate(config))
			return nullptr;
		std::unique_ptr<DAGNodePool> p{new DAGNodePool(std::move(config))};
		hd_config c = p->m_config.ToC();
		if (hd_pool_create(&c, device, &p->m_pool) != HD_OK)
			return nullptr;
		return p;
	}
	const Config<uint32_t> &GetConfig() const { return m_config; }
	hd_pool *Handle() const { return m_pool; }
	const hd_edit_stats &GetLastEditStats() const { return m_last_stats; }
	hd_status GetLastStatus() const { return m_last_status; }

	void SetRoot(NodePointer<uint32_t> root) { hd_pool_set_root(m_pool, *root); }
	NodePointer<uint32_t> GetRoot() const { return hd_pool_get_root(m_pool); }
	void Flush() { hd_sync(m_pool); } // DAGNodePool::Flush: device memory IS the pool; replicas use hd_dirty_*

	// NodePoolBase::Edit (NodePool.hpp:405-417).  On any failure the old root is returned (the reference's
	// silent-fallback convention); GetLastStatus()/GetLastEditStats().overflow_count tell why.
	// A bare editor is applied statelessly (the reference's stateless_edit, main.cpp:231-234); VBREditorWrapper{editor}
	// makes it the colour-aware vbr_edit (main.cpp:224-230); on_edit_done receives (new root, root NodeState).
	template <typename Editor_T> NodePointer<uint32_t> Edit(NodePointer<uint32_t> root, const Editor_T &editor) {
		return Edit(root, editor, [](NodePointer<uint32_t> r, auto &&) { return r; });
	}
	template <typename Editor_T, typename OnDone>
	auto Edit(NodePointer<uint32_t> root, const StatelessEditorWrapper<Editor_T> &w, OnDone &&on_edit_done) {
		return Edit(root, w.editor, std::forward<OnDone>(on_edit_done));
	}
	template <typename Editor_T, typename OnDone>
	auto Edit(NodePointer<uint32_t> root, const VBREditorWrapper<Editor_T> &w, OnDone &&on_edit_done) {
		VBRNodeState state{GetColorRoot()};
		NodePointer<uint32_t> out = EditColor(root, w.editor, w.editor.color, Editor_T::kPaint);
		if (m_last_status == HD_OK)
			state.octree_node = GetColorRoot();
		return on_edit_done(out, std::move(state));
	}
	template <typename Editor_T, typename OnDone> auto Edit(NodePointer<uint32_t> root, const Editor_T &editor, OnDone &&on_edit_done) {
		static_assert(!Editor_T::kPaint, "SphereEditor<kPaint> only edits colour: apply it through VBREditorWrapper "
		                                 "(the reference only ever calls it via vbr_edit, src/main.cpp:267-271)");
		hd_edit_desc d = editor.Desc();
		return on_edit_done(EditBatch(root, &d, 1), std::monostate{});
	}
	// NodePoolThreadedEdit::ThreadedEdit (NodePoolThreadedEdit.hpp:104-126).  The thread pool and task level are
	// accepted for source compatibility and ignored: the GPU pass has its own scheduling.
	template <typename Editor_T>
	NodePointer<uint32_t> ThreadedEdit(void * /*lf::busy_pool* */, NodePointer<uint32_t> root, const Editor_T &editor,
	                                   uint32_t /*max_task_level*/ = uint32_t(-1)) {
		return Edit(root, editor);
	}
	template <typename Editor_T, typename OnDone>
	auto ThreadedEdit(void *, NodePointer<uint32_t> root, const Editor_T &editor, uint32_t /*max_task_level*/, OnDone &&on_edit_done) {
		return Edit(root, editor, std::forward<OnDone>(on_edit_done));
	}
	// n edits applied in index order in ONE GPU pass
	NodePointer<uint32_t> EditBatch(NodePointer<uint32_t> root, const hd_edit_desc *edits, uint32_t n) {
		uint32_t out = *root;
		m_last_status = hd_edit_batch(m_pool, *root, edits, n, &out, &m_last_stats);
		return m_last_status == HD_OK ? NodePointer<uint32_t>{out} : root;
	}
	// Which implementation served the last Edit / EditBatch: 0 = general level-synchronous pass, 1 = the one-launch
	// low-latency path batches of <= 32 sphere/AABB editors take (the interactive brush), 2 = the same as a CUDA graph.
	uint32_t GetLastEditPath() const { return hd_edit_last_path(m_pool); }

	// NodePoolTraversal::Traversal<float> (NodePoolTraversal.hpp:93-256), the pick ray of main.cpp:320-321
	template <typename F = float> std::optional<Vec3> Traversal(NodePointer<uint32_t> root, Vec3 o, Vec3 d) const {
		static_assert(sizeof(F) == 4, "fp32 only: the stack depth is the float mantissa width");
		const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
		float out[3];
		int hit = 0;
		if (hd_traverse_ray(m_pool, *root, oo, dd, &hit, out) != HD_OK || !hit)
			return std::nullopt;
		return Vec3{out[0], out[1], out[2]};
	}

	// Iterate, reconstructed from its only surviving use (test/test.cpp:36-52,173-198): IterateNode(coord, node) ->
	// kProceed/kStop for every visited node, IterateVoxel(coord, bool) for all 64 voxels of every reached leaf at
	// voxel-level coordinates; a Null root iterates nothing.  Callbacks run depth first in child order like the
	// reference's recursion; nodes come back from the device a SUBTREE at a time (hd_pool_read_subtree: kIterateDepth
	// levels under the first node that is not cached yet), not one copy per node.
	static constexpr uint32_t kIterateDepth = 5, kIterateCapacity = 1u << 16;
	template <typename Iterator_T> void Iterate(NodePointer<uint32_t> root, Iterator_T *p_iterator) const {
		NodeCache cache;
		iterate_node(root, NodeCoord<uint32_t>{}, p_iterator, cache);
	}
	// NodePoolThreadedGC::ThreadedGC (NodePoolThreadedGC.hpp:394-403): compacts the pool on the GPU, returns the
	// relocated root(s).  The thread pool argument is accepted for source compatibility and ignored.
	NodePointer<uint32_t> ThreadedGC(void * /*lf::busy_pool* */, NodePointer<uint32_t> root) {
		uint32_t in = *root, out = *root;
		m_last_status = hd_gc(m_pool, &in, 1, &out, nullptr);
		return m_last_status == HD_OK ? NodePointer<uint32_t>{out} : root;
	}
	std::vector<NodePointer<uint32_t>> ThreadedGC(void *, std::vector<NodePointer<uint32_t>> roots) {
		std::vector<uint32_t> in(roots.size()), out(roots.size());
		for (size_t i = 0; i < roots.size(); ++i)
			in[i] = *roots[i];
		m_last_status = hd_gc(m_pool, in.data(), uint32_t(in.size()), out.data(), nullptr);
		if (m_last_status == HD_OK)
			for (size_t i = 0; i < roots.size(); ++i)
				roots[i] = NodePointer<uint32_t>{out[i]};
		return roots;
	}

	// Colour-aware edits: vbr_edit(...) of src/main.cpp:224-230 for AABBEditor / SphereEditor<kFill> with a colour, and
	// SphereEditor<kPaint> (paint = true).  ConfigureColor = DAGColorPool::Config::leaf_level + SetRoot.  Returns the new
	// node root; GetColorRoot() is the new colour root (EditResult of main.cpp:161-164).
	bool ConfigureColor(uint32_t leaf_level, uint32_t color_root = HD_COLOR_NULL) {
		return (m_last_status = hd_color_config(m_pool, leaf_level, color_root)) == HD_OK;
	}
	uint32_t GetColorRoot() const { return hd_color_root(m_pool); }
	template <typename Editor_T>
	NodePointer<uint32_t> EditColor(NodePointer<uint32_t> root, const Editor_T &editor, uint32_t rgb8, bool paint = false) {
		hd_edit_desc d = editor.Desc();
		uint32_t out = *root, color_root = HD_COLOR_NULL;
		m_last_status = hd_edit_color(m_pool, *root, &d, rgb8, paint ? 1u : 0u, &out, &color_root, &m_last_stats);
		return m_last_status == HD_OK ? NodePointer<uint32_t>{out} : root;
	}

	// pool file (no reference counterpart)
	bool Save(const char *path) { return (m_last_status = hd_pool_save(m_pool, path)) == HD_OK; }
	static std::unique_ptr<DAGNodePool> Load(const char *path, int device = 0) {
		hd_pool *h = nullptr;
		if (hd_pool_load(path, device, &h) != HD_OK)
			return nullptr;
		hd_config c{};
		hd_pool_get_config(h, &c);
		Config<uint32_t> cfg;
		cfg.word_bits_per_page = c.word_bits_per_page, cfg.page_bits_per_bucket = c.page_bits_per_bucket;
		cfg.bucket_bits_each_level.assign(c.bucket_bits_each_level, c.bucket_bits_each_level + c.node_levels);
		std::unique_ptr<DAGNodePool> p{new DAGNodePool(std::move(cfg))};
		p->m_pool = h;
		return p;
	}

	// Frame trace: TracePass::CmdExecute + trace.frag main() (TracePass.cpp:106-139); host output planes
	hd_status Trace(const hd_trace_params &params, const hd_trace_outputs &host_out) const {
		return hd_trace(m_pool, &params, &host_out);
	}

private:
	using NodeCache = std::unordered_map<uint32_t, std::array<uint32_t, 9>>;
	const std::array<uint32_t, 9> &read_node(uint32_t ptr, uint32_t level, NodeCache &cache) const {
		auto it = cache.find(ptr);
		if (it != cache.end())
			return it->second;
		if (cache.size() > (1u << 20)) // bound the host copy on huge trees; what is dropped is simply read again
			cache.clear();
		std::vector<hd_node_record> rec(kIterateCapacity);
		uint32_t n = 0;
		hd_pool_read_subtree(m_pool, ptr, level, kIterateDepth, rec.data(), kIterateCapacity, &n); // truncated is fine
		for (uint32_t i = 0; i < n; ++i) {
			std::array<uint32_t, 9> w{};
			std::copy(rec[i].words, rec[i].words + 9, w.begin());
			cache.emplace(rec[i].ptr, w);
		}
		return cache.at(ptr);
	}
	template <typename Iterator_T>
	void iterate_node(NodePointer<uint32_t> node, NodeCoord<uint32_t> coord, Iterator_T *it, NodeCache &cache) const {
		if (it->IterateNode(coord, node) == IterateType::kStop || !node)
			return;
		if (coord.level == m_config.GetNodeLevels() - 1) {
			const std::array<uint32_t, 9> leaf = read_node(*node, coord.level, cache);
			for (uint32_t i = 0; i < 64; ++i)
				it->IterateVoxel(coord.GetLeafCoord(i), (leaf[i >> 5] >> (i & 31u)) & 1u);
			return;
		}
		const std::array<uint32_t, 9> w = read_node(*node, coord.level, cache); // by value: the cache may be rebuilt below
		const uint32_t mask = w[0] & 0xFFu;
		uint32_t k = 1;
		for (uint32_t i = 0; i < 8; ++i) {
			NodePointer<uint32_t> child = (mask >> i & 1u) ? NodePointer<uint32_t>{w[k++]} : NodePointer<uint32_t>::Null();
			iterate_node(child, coord.GetChildCoord(i), it, cache);
		}
	}
};

} // namespace hashdag_b200
