// oracle/hashdag_oracle.cpp — TEST INFRASTRUCTURE.  CPU restatement of the reference's traversal + edit path.
//
// This file is the parity ORACLE for the CUDA product.  It is NOT part of the product: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.  The product
// (vkhashdag_b200/csrc) never links or calls it and fails loudly without its CUDA library.
//
// Parity pinning: every function below cites the reference file:line it follows (paths relative to the
// reference tree, AdamYuan/VkHashDAG).  The restatement is checked (tests/test_oracle_vs_ref.py, run where
// /root/reference exists) against oracle/_ref — the reference's own headers compiled in place — and against
// the committed known-answer vectors in tests/golden/ that were produced by the reference's code.
//
// Build: g++ -std=c++17 -O2 -ffp-contract=off -fno-fast-math (NOT -Ofast: the reference's Release flag would
// void bit-exact fp parity).
#include "../include/hashdag_b200.h"
#include "terrain.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/mman.h>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

constexpr uint32_t kNull = 0xFFFFFFFFu;

// ------------------------------------------------------------------------------------------------
// Config geometry — include/hashdag/Config.hpp:15-57
// ------------------------------------------------------------------------------------------------
struct Geometry {
	hd_config cfg{};
	uint32_t level_base[HD_MAX_NODE_LEVELS]{}; // Config.hpp:33-38
	uint32_t total_buckets = 0;
	uint64_t total_words = 0;
	uint32_t words_per_page() const { return 1u << cfg.word_bits_per_page; }
	uint32_t words_per_bucket() const { return 1u << (cfg.word_bits_per_page + cfg.page_bits_per_bucket); }
	uint32_t bucket_shift() const { return cfg.word_bits_per_page + cfg.page_bits_per_bucket; }
	uint32_t node_levels() const { return cfg.node_levels; }
	uint32_t voxel_level() const { return cfg.node_levels + 1; } // Config.hpp:31
};

bool make_geometry(const hd_config &cfg, Geometry &g) {
	if (cfg.node_levels == 0 || cfg.node_levels > HD_MAX_NODE_LEVELS || cfg.word_bits_per_page < 4)
		return false;
	g.cfg = cfg;
	uint64_t buckets = 0;
	for (uint32_t l = 0; l < cfg.node_levels; ++l) {
		g.level_base[l] = uint32_t(buckets);
		buckets += 1ull << cfg.bucket_bits_each_level[l];
	}
	uint64_t words = buckets << (cfg.word_bits_per_page + cfg.page_bits_per_bucket);
	if (words - 1 > 0xFFFFFFFEull) // Config.hpp:55
		return false;
	g.total_buckets = uint32_t(buckets);
	g.total_words = words;
	return true;
}

// ------------------------------------------------------------------------------------------------
// Hasher — include/hashdag/Hasher.hpp:21-49
// ------------------------------------------------------------------------------------------------
inline uint32_t rotl32(uint32_t v, int s) { return (v << s) | (v >> (32 - s)); }

uint32_t hash_inner(const uint32_t *w, uint32_t n) { // Hasher.hpp:22-39 (length mixed in WORDS, seed 0)
	uint32_t h = 0;
	for (uint32_t i = 0; i < n; ++i) {
		uint32_t k = w[i] * 0xcc9e2d51u;
		k = rotl32(k, 15) * 0x1b873593u;
		h = rotl32(h ^ k, 13) * 5u + 0xe6546b64u;
	}
	h ^= n;
	h ^= h >> 16;
	h *= 0x85ebca6bu;
	h ^= h >> 13;
	h *= 0xc2b2ae35u;
	h ^= h >> 16;
	return h;
}

uint32_t hash_leaf(const uint32_t *w) { // Hasher.hpp:40-48 (fmix64, truncated)
	uint64_t h = uint64_t(w[0]) | (uint64_t(w[1]) << 32);
	h ^= h >> 33;
	h *= 0xff51afd7ed558ccdull;
	h ^= h >> 33;
	h *= 0xc4ceb9fe1a85ec53ull;
	h ^= h >> 33;
	return uint32_t(h);
}

// ------------------------------------------------------------------------------------------------
// Pool storage — the contract of src/DAGNodePool.hpp:41-74 with one lazily-zeroed flat mapping
// ------------------------------------------------------------------------------------------------
struct Stats {
	uint64_t edit_nodes = 0, edit_leaves = 0, upserts = 0, appended_nodes = 0, appended_words = 0, overflow = 0;
	uint64_t scan_words = 0, read_words = 0;
};

} // namespace

struct orc_pool {
	Geometry g;
	uint32_t *words = nullptr; // flat address space, SURVEY App. A.1
	std::vector<uint32_t> bucket_words;
	std::vector<uint32_t> filled; // NodePool.hpp:54
	Stats st;
};

namespace {

// find_node_in_span + find_node — NodePool.hpp:79-132.  `is_leaf` selects the node-size rule
// (leaf: 2 words, NodePool.hpp:236; inner: 1+popcount(mask&0xFF), 0 for a zero mask = page padding, :214-226).
uint32_t find_node(orc_pool &p, bool is_leaf, uint32_t bucket, uint32_t bucket_words, const uint32_t *node,
                   uint32_t n) {
	const Geometry &g = p.g;
	const uint32_t wpp = g.words_per_page();
	uint32_t off = 0;
	const uint32_t base = bucket << g.bucket_shift();
	while (off < bucket_words) {
		uint32_t page_end = std::min((off & ~(wpp - 1)) + wpp, bucket_words);
		// scan one page span [off, page_end)
		uint32_t it = off;
		while (n <= page_end - it) {
			const uint32_t *q = p.words + base + it;
			uint32_t sz = is_leaf ? 2u : (uint8_t(q[0]) ? 1u + uint32_t(__builtin_popcount(uint8_t(q[0]))) : 0u);
			if (sz == 0)
				break;
			if (sz == n && std::equal(node, node + n, q)) {
				p.st.scan_words += it + n;
				return base + it;
			}
			it += sz;
		}
		off = (off & ~(wpp - 1)) + wpp;
	}
	p.st.scan_words += bucket_words;
	return kNull;
}

// append_node — NodePool.hpp:134-157
uint32_t append_node(orc_pool &p, uint32_t bucket, uint32_t &bucket_words, const uint32_t *node, uint32_t n) {
	const Geometry &g = p.g;
	const uint32_t wpp = g.words_per_page();
	if (bucket_words + n > g.words_per_bucket())
		return kNull;
	uint32_t slot = bucket_words >> g.cfg.word_bits_per_page, off = bucket_words & (wpp - 1);
	const uint32_t base = bucket << g.bucket_shift();
	if (off + n > wpp) { // would straddle: zero the tail, go to next page
		std::fill(p.words + base + (slot << g.cfg.word_bits_per_page) + off,
		          p.words + base + ((slot + 1) << g.cfg.word_bits_per_page), 0u);
		++slot, off = 0;
	}
	uint32_t at = (slot << g.cfg.word_bits_per_page) | off;
	std::copy(node, node + n, p.words + base + at);
	p.st.appended_words += at + n - bucket_words;
	p.st.appended_nodes += 1;
	bucket_words = at + n;
	return base + at;
}

// upsert_node<false> — NodePool.hpp:159-212
uint32_t upsert(orc_pool &p, uint32_t level, const uint32_t *node, uint32_t n, uint32_t fallback) {
	const Geometry &g = p.g;
	const bool is_leaf = level == g.node_levels() - 1;
	uint32_t h = is_leaf ? hash_leaf(node) : hash_inner(node, n);
	uint32_t bucket = g.level_base[level] + (h & ((1u << g.cfg.bucket_bits_each_level[level]) - 1u));
	p.st.upserts++;
	uint32_t &bw = p.bucket_words[bucket];
	uint32_t found = find_node(p, is_leaf, bucket, bw, node, n);
	if (found != kNull)
		return found;
	uint32_t app = append_node(p, bucket, bw, node, n);
	if (app != kNull)
		return app;
	p.st.overflow++;
	return fallback;
}

// make_filled_node_pointers — NodePool.hpp:240-262
void make_filled(orc_pool &p) {
	if (!p.filled.empty())
		return;
	uint32_t L = p.g.node_levels();
	p.filled.assign(L, kNull);
	uint32_t leaf[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
	p.filled[L - 1] = upsert(p, L - 1, leaf, 2, kNull);
	for (uint32_t l = L - 1; l-- > 0;) {
		uint32_t c = p.filled[l + 1];
		uint32_t node[9] = {0xFFu, c, c, c, c, c, c, c, c};
		p.filled[l] = upsert(p, l, node, 9, kNull);
	}
}

// ------------------------------------------------------------------------------------------------
// Editors — src/main.cpp:32-150 as POD descriptors; EditType of include/hashdag/Editor.hpp:18
// ------------------------------------------------------------------------------------------------
enum EditType { kNotAffected, kProceed, kFill, kClear };

struct Box { // node bounds at voxel level: NodeCoord::GetLower/UpperBoundAtLevel, NodeCoord.hpp:86-93
	uint32_t lb[3], ub[3];
};

inline Box node_box(uint32_t voxel_level, uint32_t level, const uint32_t pos[3]) {
	uint32_t bits = voxel_level - level;
	Box b;
	for (int i = 0; i < 3; ++i)
		b.lb[i] = pos[i] << bits, b.ub[i] = (pos[i] + 1u) << bits;
	return b;
}

struct HeightCache { // per-thread cache of the terrain height (pure function of x,z)
	uint32_t tag_x[16], tag_z[16], h[16];
	uint32_t valid = 0;
};

struct EditorCtx {
	hd_edit_desc d;
	uint32_t voxel_level;
	terrain::Params tp;
	mutable HeightCache hc;
	explicit EditorCtx(const hd_edit_desc &desc, uint32_t vl) : d(desc), voxel_level(vl) {
		tp = terrain::from_desc(d.aux, d.p0, d.p1);
	}

	uint32_t terrain_height(uint32_t x, uint32_t z) const {
		uint32_t s = (x & 3u) | ((z & 3u) << 2);
		if ((hc.valid >> s & 1u) && hc.tag_x[s] == x && hc.tag_z[s] == z)
			return hc.h[s];
		uint32_t h = terrain::height(tp, x, z);
		hc.valid |= 1u << s, hc.tag_x[s] = x, hc.tag_z[s] = z, hc.h[s] = h;
		return h;
	}

	EditType edit_node(uint32_t level, const uint32_t pos[3]) const {
		Box b = node_box(voxel_level, level, pos);
		switch (d.kind) {
		case HD_EDIT_AABB_FILL: { // main.cpp:35-46
			bool out = false, in = true;
			for (int i = 0; i < 3; ++i) {
				out |= b.ub[i] <= d.p0[i] || b.lb[i] >= d.p1[i];
				in &= b.lb[i] >= d.p0[i] && b.ub[i] <= d.p1[i];
			}
			return out ? kNotAffected : (in ? kFill : kProceed);
		}
		case HD_EDIT_SPHERE_FILL:
		case HD_EDIT_SPHERE_DIG: { // main.cpp:77-106 (i64 differences, u64 squares)
			uint64_t max_n2 = 0, min_n2 = 0;
			for (int i = 0; i < 3; ++i) {
				int64_t lo = int64_t(b.lb[i]) - int64_t(d.p0[i]), hi = int64_t(b.ub[i]) - int64_t(d.p0[i]);
				uint64_t lo2 = uint64_t(lo * lo), hi2 = uint64_t(hi * hi);
				max_n2 += std::max(lo2, hi2);
				if (lo > 0)
					min_n2 += lo2;
				if (hi < 0)
					min_n2 += hi2;
			}
			if (max_n2 <= d.r2)
				return d.kind == HD_EDIT_SPHERE_DIG ? kClear : kFill;
			return min_n2 > d.r2 ? kNotAffected : kProceed;
		}
		case HD_EDIT_TERRAIN_FILL: {
			const int ext = terrain::extent_class(tp, b.lb[0], b.lb[2], voxel_level - level);
			if (ext == 0)
				return kNotAffected;
			uint32_t hmin, hmax;
			terrain::height_bounds(tp, b.lb[0], b.lb[2], voxel_level - level, hmin, hmax);
			if (b.lb[1] >= hmax)
				return kNotAffected;
			if (ext == 2 && b.ub[1] <= hmin)
				return kFill;
			return kProceed;
		}
		}
		return kNotAffected;
	}

	bool in_range(const uint32_t v[3]) const {
		switch (d.kind) {
		case HD_EDIT_AABB_FILL: // main.cpp:57-59
			return v[0] >= d.p0[0] && v[1] >= d.p0[1] && v[2] >= d.p0[2] && v[0] < d.p1[0] && v[1] < d.p1[1] &&
			       v[2] < d.p1[2];
		case HD_EDIT_SPHERE_FILL:
		case HD_EDIT_SPHERE_DIG: { // main.cpp:127-132
			int64_t dx = int64_t(v[0]) - int64_t(d.p0[0]), dy = int64_t(v[1]) - int64_t(d.p0[1]),
			        dz = int64_t(v[2]) - int64_t(d.p0[2]);
			return uint64_t(dx * dx + dy * dy + dz * dz) <= d.r2;
		}
		case HD_EDIT_TERRAIN_FILL:
			return terrain::in_extent(tp, v[0], v[2]) && v[1] < terrain_height(v[0], v[2]);
		}
		return false;
	}

	bool edit_voxel(const uint32_t v[3], bool voxel) const { // main.cpp:60-63,133-142
		bool in = in_range(v);
		return d.kind == HD_EDIT_SPHERE_DIG ? (voxel && !in) : (voxel || in);
	}
};

// ------------------------------------------------------------------------------------------------
// Edit recursion — NodePool.hpp:264-417
// ------------------------------------------------------------------------------------------------
void unpack_node(const orc_pool &p, uint32_t ptr, uint32_t out[9]) { // NodePool.hpp:278-309
	out[0] = 0;
	for (int i = 1; i < 9; ++i)
		out[i] = kNull;
	if (ptr == kNull)
		return;
	const uint32_t *q = p.words + ptr;
	out[0] = q[0];
	uint32_t k = 1;
	for (uint32_t i = 0; i < 8; ++i)
		if (q[0] >> i & 1u)
			out[1 + i] = q[k++];
}

uint32_t edit_node(orc_pool &p, const EditorCtx &ed, uint32_t ptr, uint32_t level, const uint32_t pos[3]);

uint32_t edit_leaf(orc_pool &p, const EditorCtx &ed, uint32_t ptr, uint32_t level, const uint32_t pos[3]) {
	// NodePool.hpp:319-343; voxel i of the leaf: NodeCoord::GetLeafCoord, NodeCoord.hpp:32-43
	p.st.edit_leaves++;
	uint32_t leaf[2] = {0, 0};
	if (ptr != kNull)
		leaf[0] = p.words[ptr], leaf[1] = p.words[ptr + 1], p.st.read_words += 2;
	bool changed = false;
	for (uint32_t i = 0; i < 64; ++i) {
		uint32_t v[3] = {(pos[0] << 2) | ((i >> 2) & 2u) | (i & 1u), (pos[1] << 2) | ((i >> 3) & 2u) | ((i >> 1) & 1u),
		                 (pos[2] << 2) | ((i >> 4) & 2u) | ((i >> 2) & 1u)};
		bool voxel = leaf[i >> 5] >> (i & 31u) & 1u;
		bool nv = ed.edit_voxel(v, voxel);
		if (nv != voxel)
			changed = true, leaf[i >> 5] ^= 1u << (i & 31u);
	}
	if (!changed)
		return ptr;
	if (leaf[0] == 0 && leaf[1] == 0)
		return kNull;
	return upsert(p, level, leaf, 2, ptr);
}

uint32_t edit_switch(orc_pool &p, const EditorCtx &ed, uint32_t ptr, uint32_t level, const uint32_t pos[3]) {
	// NodePool.hpp:345-360
	switch (ed.edit_node(level, pos)) {
	case kClear:
		return kNull;
	case kFill:
		return p.filled[level];
	case kNotAffected:
		return ptr;
	default:
		return edit_node(p, ed, ptr, level, pos);
	}
}

uint32_t edit_node(orc_pool &p, const EditorCtx &ed, uint32_t ptr, uint32_t level, const uint32_t pos[3]) {
	// NodePool.hpp:362-396
	if (level == p.g.node_levels() - 1)
		return edit_leaf(p, ed, ptr, level, pos);
	p.st.edit_nodes++;
	uint32_t node[9];
	unpack_node(p, ptr, node);
	if (ptr != kNull)
		p.st.read_words += 1 + __builtin_popcount(node[0]);
	bool changed = false;
	for (uint32_t i = 0; i < 8; ++i) {
		uint32_t cpos[3] = {(pos[0] << 1) | (i & 1u), (pos[1] << 1) | ((i >> 1) & 1u), (pos[2] << 1) | ((i >> 2) & 1u)};
		uint32_t oldc = node[1 + i];
		uint32_t newc = edit_switch(p, ed, oldc, level + 1, cpos);
		changed |= newc != oldc;
		node[1 + i] = newc;
		if ((newc != kNull) != (oldc != kNull))
			node[0] ^= 1u << i;
	}
	if (!changed)
		return ptr;
	if (node[0] == 0)
		return kNull;
	uint32_t packed[9], k = 1; // get_packed_node_inplace, NodePool.hpp:272-277
	packed[0] = node[0];
	for (uint32_t i = 0; i < 8; ++i)
		if (node[0] >> i & 1u)
			packed[k++] = node[1 + i];
	return upsert(p, level, packed, k, ptr);
}

uint32_t edit(orc_pool &p, uint32_t root, const hd_edit_desc &d) { // NodePoolBase::Edit, NodePool.hpp:405-417
	make_filled(p);
	EditorCtx ed(d, p.g.voxel_level());
	uint32_t pos[3] = {0, 0, 0};
	return edit_switch(p, ed, root, 0, pos);
}

// ------------------------------------------------------------------------------------------------
// Canonicaliser + voxel queries over ANY flat word array (oracle pool, oracle/_ref pool or GPU readback)
// ------------------------------------------------------------------------------------------------
inline uint64_t mix64(uint64_t h) {
	h ^= h >> 30;
	h *= 0xbf58476d1ce4e5b9ull;
	h ^= h >> 27;
	h *= 0x94d049bb133111ebull;
	h ^= h >> 31;
	return h;
}

struct Canon {
	const uint32_t *words;
	uint32_t node_levels;
	std::vector<std::unordered_map<uint32_t, uint64_t>> memo;       // per level: pointer -> content hash
	std::vector<std::unordered_map<uint64_t, uint32_t>> by_content; // per level: content hash -> multiplicity
	std::vector<std::unordered_map<uint32_t, uint64_t>> voxels;     // per level: pointer -> set voxel count

	uint64_t hash(uint32_t ptr, uint32_t level) {
		auto it = memo[level].find(ptr);
		if (it != memo[level].end())
			return it->second;
		uint64_t h;
		if (level == node_levels - 1) {
			h = mix64(0x1eafull ^ (uint64_t(words[ptr]) | uint64_t(words[ptr + 1]) << 32));
		} else {
			uint32_t mask = words[ptr] & 0xFFu;
			h = mix64(0x1234567ull + mask + (uint64_t(level) << 32));
			uint32_t k = 1;
			for (uint32_t i = 0; i < 8; ++i)
				if (mask >> i & 1u)
					h = mix64(h ^ (hash(words[ptr + k++], level + 1) + 0x9e3779b97f4a7c15ull * (i + 1)));
		}
		memo[level].emplace(ptr, h);
		by_content[level][h]++;
		return h;
	}
	uint64_t count(uint32_t ptr, uint32_t level) {
		if (level == node_levels - 1)
			return uint64_t(__builtin_popcount(words[ptr]) + __builtin_popcount(words[ptr + 1]));
		auto it = voxels[level].find(ptr);
		if (it != voxels[level].end())
			return it->second;
		uint32_t mask = words[ptr] & 0xFFu, k = 1;
		uint64_t c = 0;
		for (uint32_t i = 0; i < 8; ++i)
			if (mask >> i & 1u)
				c += count(words[ptr + k++], level + 1);
		voxels[level].emplace(ptr, c);
		return c;
	}
};

// ------------------------------------------------------------------------------------------------
// Tracer — include/hashdag/NodePoolTraversal.hpp:93-256 (host) and shader/src/trace.frag:56-402 (GPU shader)
// fp32 only, no contraction; min/max written out so both sides of the parity test use the same selection rule.
// ------------------------------------------------------------------------------------------------
inline float fbits(uint32_t u) {
	float f;
	std::memcpy(&f, &u, 4);
	return f;
}
inline uint32_t ubits(float f) {
	uint32_t u;
	std::memcpy(&u, &f, 4);
	return u;
}
inline float fmin2(float a, float b) { return b < a ? b : a; } // glm::min / GLSL min
inline float fmax2(float a, float b) { return a < b ? b : a; } // glm::max / GLSL max

struct Scene {
	const uint32_t *nodes;
	const uint32_t *color_nodes;
	const uint32_t *color_leaves;
	const float *beam = nullptr; // BEAM_OPTIMIZATION: coarse start-t image (beam.frag), bw x bh texels
	uint32_t bw = 0, bh = 0;
};

struct March {
	bool hit;
	float pos[3];      // mirrored-space cube corner at exit
	float scale_exp2;  // cube size at exit
	uint32_t scale;    // may have wrapped past 22 on a miss
	uint32_t octant;   // octant_mask
	float t_min, t_max;
	float o[3], d[3];  // shifted origin (o+1) and epsilon-clamped direction
	float t_coef[3], t_bias[3];
	uint32_t iter;
	uint64_t fetches;  // 32-bit node words read (F of SURVEY §8d)
};

constexpr uint32_t kStack = 23; // trace.frag:52-54, NodePoolTraversal.hpp:104

// Shared core of Traversal<float> (NodePoolTraversal.hpp:100-243) and DAG_RayMarch (trace.frag:82-221).
// use_lod=false reproduces the host function (no projection cut-off).
void march(const uint32_t *nodes, uint32_t root, uint32_t leaf_level, bool use_lod, float proj_factor, float proj_bias,
           const float o_in[3], const float d_in[3], March &m) {
	const float eps = fbits((127u - kStack) << 23); // 2^-23 == FLT_EPSILON (trace.frag:88, Traversal.hpp:103)
	uint32_t stack[kStack];
	for (int i = 0; i < 3; ++i) {
		m.o[i] = o_in[i] + 1.0f;
		float d = d_in[i];
		m.d[i] = std::fabs(d) > eps ? d : (d >= 0 ? eps : -eps);
		m.t_coef[i] = 1.0f / -std::fabs(m.d[i]);
		m.t_bias[i] = m.t_coef[i] * m.o[i];
	}
	uint32_t octant = 0;
	for (int i = 0; i < 3; ++i)
		if (m.d[i] > 0.0f)
			octant ^= 1u << i, m.t_bias[i] = 3.0f * m.t_coef[i] - m.t_bias[i];

	const float *tc = m.t_coef, *tb = m.t_bias;
	float t_min = fmax2(fmax2(2.0f * tc[0] - tb[0], 2.0f * tc[1] - tb[1]), 2.0f * tc[2] - tb[2]);
	float t_max = fmin2(fmin2(tc[0] - tb[0], tc[1] - tb[1]), tc[2] - tb[2]);
	float h = t_max;
	t_min = fmax2(t_min, 0.0f);
	t_max = fmin2(t_max, 1.0f);

	uint32_t parent = root, child_bits = 0, idx = 0;
	float pos[3] = {1.0f, 1.0f, 1.0f};
	for (int i = 0; i < 3; ++i)
		if (1.5f * tc[i] - tb[i] > t_min)
			idx ^= 1u << i, pos[i] = 1.5f;

	uint32_t scale = kStack - 1;
	float scale_exp2 = 0.5f;
	const uint32_t leaf_scale = kStack - leaf_level;
	uint32_t iter = 0;
	uint64_t fetches = 0;

	for (;;) {
		++iter;
		if (child_bits == 0u) {
			if (scale > leaf_scale)
				child_bits = nodes[parent], fetches += 1;
			else if (scale == leaf_scale) { // DAG_GetLeafFirstChildBits, trace.frag:56-67
				uint32_t l0 = nodes[parent], l1 = nodes[parent + 1];
				fetches += 2;
				for (uint32_t b = 0; b < 4; ++b) {
					child_bits |= ((l0 >> (8 * b)) & 0xFFu) ? (1u << b) : 0u;
					child_bits |= ((l1 >> (8 * b)) & 0xFFu) ? (16u << b) : 0u;
				}
			} else
				child_bits = parent;
		}
		float t_corner[3] = {pos[0] * tc[0] - tb[0], pos[1] * tc[1] - tb[1], pos[2] * tc[2] - tb[2]};
		float tc_max = fmin2(fmin2(t_corner[0], t_corner[1]), t_corner[2]);
		uint32_t child_shift = idx ^ octant, child_mask = 1u << child_shift;

		if ((child_bits & child_mask) != 0 && t_min <= t_max) {
			float half = scale_exp2 * 0.5f;
			float t_center[3] = {half * tc[0] + t_corner[0], half * tc[1] + t_corner[1], half * tc[2] + t_corner[2]};
			if (scale < leaf_scale || (use_lod && scale_exp2 * proj_factor < tc_max + proj_bias))
				break;
			if (tc_max < h)
				stack[scale] = parent;
			h = tc_max;
			if (scale > leaf_scale)
				parent = nodes[parent + 1u + uint32_t(__builtin_popcount(child_bits & (child_mask - 1u)))];
			else
				parent = (nodes[parent + (child_shift >> 2)] >> ((child_shift & 3u) << 3)) & 0xFFu;
			fetches += 1;
			idx = 0;
			--scale;
			scale_exp2 = half;
			for (int i = 0; i < 3; ++i)
				if (t_center[i] > t_min)
					idx ^= 1u << i, pos[i] += scale_exp2;
			child_bits = 0;
			continue;
		}
		uint32_t step_mask = 0;
		for (int i = 0; i < 3; ++i)
			if (t_corner[i] <= tc_max)
				step_mask ^= 1u << i, pos[i] -= scale_exp2;
		t_min = tc_max;
		idx ^= step_mask;
		if ((idx & step_mask) != 0) {
			uint32_t differing = 0;
			for (int i = 0; i < 3; ++i)
				if (step_mask >> i & 1u)
					differing |= ubits(pos[i]) ^ ubits(pos[i] + scale_exp2);
			scale = differing ? 31u - uint32_t(__builtin_clz(differing)) : 0xFFFFFFFFu; // findMSB(0) = -1
			if (scale >= kStack)
				break;
			scale_exp2 = fbits((scale - kStack + 127u) << 23);
			parent = stack[scale];
			uint32_t sh[3];
			for (int i = 0; i < 3; ++i)
				sh[i] = ubits(pos[i]) >> scale, pos[i] = fbits(sh[i] << scale);
			idx = (sh[0] & 1u) | ((sh[1] & 1u) << 1) | ((sh[2] & 1u) << 2);
			h = 0.0f;
			child_bits = 0;
		}
	}
	m.hit = scale < kStack && t_min <= t_max;
	m.scale = scale, m.scale_exp2 = scale_exp2, m.octant = octant, m.t_min = t_min, m.t_max = t_max;
	m.iter = iter, m.fetches = fetches;
	for (int i = 0; i < 3; ++i)
		m.pos[i] = pos[i];
}

// Colour decode — trace.frag:272-364 over the DAGColorPool layout (src/DAGColorPool.hpp:23-52,66-113)
inline uint32_t morton32(uint32_t x, uint32_t y, uint32_t z) { // trace.frag:272-278
	auto spread = [](uint32_t u) {
		u = (u | (u << 16)) & 0x030000FFu;
		u = (u | (u << 8)) & 0x0300F00Fu;
		u = (u | (u << 4)) & 0x030C30C3u;
		u = (u | (u << 2)) & 0x09249249u;
		return u;
	};
	return spread(x) | (spread(y) << 1) | (spread(z) << 2);
}
struct Rgb {
	float r, g, b;
};
inline Rgb unorm4x8(uint32_t d) { return {float(d & 0xFFu) / 255.0f, float((d >> 8) & 0xFFu) / 255.0f, float((d >> 16) & 0xFFu) / 255.0f}; }
inline Rgb rgb565(uint32_t c) { // trace.frag:299-302
	return {float(c & 0x1Fu) / 31.0f, float((c >> 5) & 0x3Fu) / 63.0f, float((c >> 11) & 0x1Fu) / 31.0f};
}

Rgb leaf_color(const uint32_t *lv, uint32_t idx, const uint32_t sub[3], uint64_t &fetches) { // trace.frag:304-349
	uint32_t macro_cnt = lv[idx + 1], block_cnt = lv[idx + 2];
	fetches += 2;
	uint32_t macro_off = idx + 4, block_off = macro_off + (macro_cnt << 1), weight_off = block_off + (block_cnt << 1);
	uint32_t vox_id = morton32(sub[0], sub[1], sub[2]), macro_id = vox_id >> 14;
	if (macro_id >= macro_cnt)
		return {0, 0, 0};
	uint32_t mx = lv[macro_off + (macro_id << 1)], my = lv[(macro_off + (macro_id << 1)) | 1u];
	fetches += 2;
	block_off += mx << 1;
	if (macro_id + 1 < macro_cnt)
		block_cnt = lv[macro_off + ((macro_id + 1) << 1)] - mx, fetches += 1;
	else
		block_cnt = block_cnt - mx;
	vox_id &= 0x3FFFu;
	if (block_cnt == 0)
		return {0, 0, 0};
	for (uint32_t it = 0; it <= 14 && block_cnt != 0; ++it) { // first block with voxel_index_offset > vox_id
		uint32_t step = block_cnt >> 1;
		fetches += 1;
		if ((lv[(block_off + (step << 1)) | 1u] >> 18) <= vox_id)
			block_cnt -= step + 1, block_off += (step + 1) << 1;
		else
			block_cnt = step;
	}
	block_off -= 2;
	uint32_t bx = lv[block_off], by = lv[block_off | 1u];
	fetches += 2;
	uint32_t bpw = (by >> 16) & 3u;
	if (bpw == 0)
		return unorm4x8(bx);
	vox_id -= by >> 18;
	uint32_t bit_id = my + (by & 0xFFFFu) + vox_id * bpw;
	uint32_t bit_off = bit_id & 31u, w; // Color_GetWeight, trace.frag:287-298
	uint32_t w0 = lv[weight_off + (bit_id >> 5)] >> bit_off;
	fetches += 1;
	if (bit_off + bpw <= 32)
		w = w0 & ((1u << bpw) - 1u);
	else {
		uint32_t w1 = lv[weight_off + (bit_id >> 5) + 1] & ((1u << (bit_off + bpw - 32u)) - 1u);
		fetches += 1;
		w = w0 | (w1 << (32u - bit_off));
	}
	float alpha = float(w) / float((1u << bpw) - 1u);
	Rgb a = rgb565(bx), b = rgb565(bx >> 16); // mix(a,b,t) = a*(1-t) + b*t
	float ia = 1.0f - alpha;
	return {a.r * ia + b.r * alpha, a.g * ia + b.g * alpha, a.b * ia + b.b * alpha};
}

Rgb color_fetch(const Scene &s, uint32_t root, uint32_t voxel_level, uint32_t leaf_level, const uint32_t vox[3],
                uint64_t &fetches) { // trace.frag:351-364
	uint32_t ptr = root;
	for (uint32_t l = 0; l < leaf_level; ++l) {
		uint32_t tag = ptr >> 30, data = ptr & 0x3FFFFFFFu;
		if (tag != 0)
			return unorm4x8(data);
		uint32_t sh = voxel_level - 1u - l;
		uint32_t c = ((vox[0] >> sh) & 1u) | (((vox[1] >> sh) & 1u) << 1) | (((vox[2] >> sh) & 1u) << 2);
		ptr = s.color_nodes[(ptr << 3) | c];
		fetches += 1;
	}
	uint32_t tag = ptr >> 30, data = ptr & 0x3FFFFFFFu;
	if (tag == 2u) {
		uint32_t m = (1u << (voxel_level - leaf_level)) - 1u;
		uint32_t sub[3] = {vox[0] & m, vox[1] & m, vox[2] & m};
		return leaf_color(s.color_leaves, data, sub, fetches);
	}
	return unorm4x8(data);
}

inline void normalize3(float v[3]) { // glm: v * inversesqrt(dot(v,v)), dot summed x+y+z
	float dot = (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2];
	float inv = 1.0f / std::sqrt(dot);
	v[0] *= inv, v[1] *= inv, v[2] *= inv;
}

// Pinned sine for the heat map (trace.frag:372): GLSL leaves sin() precision implementation-defined, so the
// oracle and the kernel both use this polynomial (argument range here is [-3, 2]).
inline float pinned_sin(float x) {
	const float pi = 3.14159274f, half_pi = 1.57079637f;
	if (x > half_pi)
		x = pi - x;
	else if (x < -half_pi)
		x = -pi - x;
	float x2 = x * x;
	float p = -2.50521084e-08f;
	p = p * x2 + 2.75573192e-06f;
	p = p * x2 + -1.98412701e-04f;
	p = p * x2 + 8.33333377e-03f;
	p = p * x2 + -1.66666672e-01f;
	return x + (x * x2) * p;
}
inline uint32_t to_unorm8(float x) {
	x = x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x);
	return uint32_t(x * 255.0f + 0.5f);
}
inline uint32_t pack_rgba8(float r, float g, float b) {
	return to_unorm8(r) | (to_unorm8(g) << 8) | (to_unorm8(b) << 16) | 0xFF000000u;
}

struct PixelOut {
	uint32_t rgba8;
	hd_hit_record rec;
	uint32_t iter;
	uint64_t fetches;
};

// One fragment: trace.frag main() :374-402 + DAG_RayMarch epilogue :223-269
void shade_pixel(const Scene &s, const hd_trace_params &P, uint32_t px, uint32_t py, PixelOut &out) {
	// ray generation, trace.frag:366-377 (gl_FragCoord = pixel centre)
	float cx = (float(px) + 0.5f) / float(P.width), cy = (float(py) + 0.5f) / float(P.height);
	cx = cx * 2.0f - 1.0f, cy = cy * 2.0f - 1.0f;
	float d[3];
	for (int i = 0; i < 3; ++i)
		d[i] = (P.look[i] - P.side[i] * cx) - P.up[i] * cy;
	normalize3(d);

	March m{};
	bool hit = false;
	bool marched = P.dag_root != kNull;
	if (s.beam) { // trace.frag:384-389 — MIN-reduction linear sampler = min of the 2x2 texel footprint, x0.98
		const float u = (float(px) + 0.5f) / float(P.width) * float(s.bw) - 0.5f;
		const float w = (float(py) + 0.5f) / float(P.height) * float(s.bh) - 0.5f;
		int i0 = int(std::floor(u)), j0 = int(std::floor(w));
		int i1 = std::min(i0 + 1, int(s.bw) - 1), j1 = std::min(j0 + 1, int(s.bh) - 1);
		i0 = std::max(i0, 0), j0 = std::max(j0, 0);
		float beam = fmin2(fmin2(s.beam[size_t(j0) * s.bw + i0], s.beam[size_t(j0) * s.bw + i1]),
		                   fmin2(s.beam[size_t(j1) * s.bw + i0], s.beam[size_t(j1) * s.bw + i1]));
		beam = beam * 0.98f;
		marched = marched && !std::isinf(beam);
		if (marched) {
			const float o2[3] = {P.pos[0] + beam * d[0], P.pos[1] + beam * d[1], P.pos[2] + beam * d[2]};
			march(s.nodes, P.dag_root, P.dag_leaf_level, true, P.proj_factor, beam, o2, d, m);
			hit = m.hit;
		}
	} else if (marched) {
		march(s.nodes, P.dag_root, P.dag_leaf_level, true, P.proj_factor, 0.0f, P.pos, d, m);
		hit = m.hit;
	}
	out.iter = m.iter, out.fetches = m.fetches;
	out.rec = {{0, 0, 0}, 0};
	float norm[3] = {0, 0, 0};
	uint32_t vox[3] = {0, 0, 0}, vox_size_log2 = 0;
	if (marched) {
		// normal, trace.frag:223-234
		float tcn[3];
		for (int i = 0; i < 3; ++i)
			tcn[i] = m.t_coef[i] * (m.pos[i] + m.scale_exp2) - m.t_bias[i];
		if (tcn[0] > tcn[1] && tcn[0] > tcn[2])
			norm[0] = -1;
		else if (tcn[1] > tcn[2])
			norm[1] = -1;
		else
			norm[2] = -1;
		for (int i = 0; i < 3; ++i)
			if ((m.octant >> i & 1u) == 0u)
				norm[i] = -norm[i];
		if (hit) { // voxel size & position, trace.frag:236-246
			const uint32_t voxel_level = P.dag_leaf_level + 1u, voxel_scale = kStack - voxel_level;
			vox_size_log2 = m.scale - voxel_scale;
			uint32_t vox_size = 1u << vox_size_log2;
			for (int i = 0; i < 3; ++i) {
				vox[i] = (ubits(m.pos[i]) & 0x7FFFFFu) >> voxel_scale;
				if (m.octant >> i & 1u)
					vox[i] = (1u << voxel_level) - vox_size - vox[i];
			}
		}
	}
	Rgb col{0, 0, 0};
	if (hit && P.type == 0) {
		col = color_fetch(s, P.color_root, P.voxel_level, P.color_leaf_level, vox, out.fetches);
	} else if (hit) {
		uint64_t dummy = 0; // parity record always carries the fetched colour
		col = color_fetch(s, P.color_root, P.voxel_level, P.color_leaf_level, vox, dummy);
	}
	if (hit) {
		out.rec.vox[0] = vox[0], out.rec.vox[1] = vox[1], out.rec.vox[2] = vox[2];
		out.rec.packed = 0x80000000u | (vox_size_log2 << 24) | (pack_rgba8(col.r, col.g, col.b) & 0xFFFFFFu);
	}
	if (P.type == 0) { // trace.frag:395-397
		float L[3] = {4.0f, 5.0f, 3.0f};
		normalize3(L);
		float dt = (norm[0] * L[0] + norm[1] * L[1]) + norm[2] * L[2];
		float diffuse = fmax2(dt, 0.0f) * 0.5f + 0.5f;
		out.rgba8 = hit ? pack_rgba8(diffuse * col.r, diffuse * col.g, diffuse * col.b) : pack_rgba8(0, 0, 0);
	} else if (P.type == 1) { // trace.frag:398-399
		out.rgba8 = hit ? pack_rgba8(norm[0] * 0.5f + 0.5f, norm[1] * 0.5f + 0.5f, norm[2] * 0.5f + 0.5f)
		                : pack_rgba8(0, 0, 0);
	} else { // trace.frag:372,401
		float x = float(m.iter) / 128.0f;
		x = x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x);
		float a = x * 3.0f;
		out.rgba8 = pack_rgba8(pinned_sin(a - 1.0f) * 0.5f + 0.5f, pinned_sin(a - 2.0f) * 0.5f + 0.5f,
		                       pinned_sin(a - 3.0f) * 0.5f + 0.5f);
	}
}

} // namespace

// =================================================================================================
// C entry points (ctypes-friendly)
// =================================================================================================
extern "C" {

uint32_t orc_hash_inner(const uint32_t *w, uint32_t n) { return hash_inner(w, n); }
uint32_t orc_hash_leaf(const uint32_t *w) { return hash_leaf(w); }

int orc_config_from_default(const hd_default_config *dc, hd_config *out) { // Config.hpp:66-74
	std::memset(out, 0, sizeof(*out));
	if (dc->level_count < 2 || dc->level_count - 1 > HD_MAX_NODE_LEVELS)
		return 1;
	out->word_bits_per_page = dc->word_bits_per_page;
	out->page_bits_per_bucket = dc->page_bits_per_bucket;
	out->node_levels = dc->level_count - 1;
	for (uint32_t l = 0; l + 1 < dc->level_count; ++l)
		out->bucket_bits_each_level[l] =
		    l < dc->top_level_count ? dc->bucket_bits_per_top_level : dc->bucket_bits_per_bottom_level;
	return 0;
}

orc_pool *orc_pool_create(const hd_config *cfg) {
	auto *p = new orc_pool();
	if (!make_geometry(*cfg, p->g)) {
		delete p;
		return nullptr;
	}
	void *m = mmap(nullptr, p->g.total_words * 4, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE,
	               -1, 0);
	if (m == MAP_FAILED) {
		delete p;
		return nullptr;
	}
	p->words = static_cast<uint32_t *>(m);
	p->bucket_words.assign(p->g.total_buckets, 0);
	return p;
}
void orc_pool_destroy(orc_pool *p) {
	if (!p)
		return;
	munmap(p->words, p->g.total_words * 4);
	delete p;
}
uint32_t *orc_pool_words(orc_pool *p) { return p->words; }
uint32_t *orc_pool_bucket_words(orc_pool *p) { return p->bucket_words.data(); }
uint64_t orc_pool_total_words(const orc_pool *p) { return p->g.total_words; }
uint32_t orc_pool_total_buckets(const orc_pool *p) { return p->g.total_buckets; }
uint32_t orc_pool_level_base(const orc_pool *p, uint32_t level) { return p->g.level_base[level]; }

uint32_t orc_upsert(orc_pool *p, uint32_t level, const uint32_t *node, uint32_t n, uint32_t fallback) {
	return upsert(*p, level, node, n, fallback);
}
void orc_filled_nodes(orc_pool *p, uint32_t *out) {
	make_filled(*p);
	std::copy(p->filled.begin(), p->filled.end(), out);
}
uint32_t orc_edit(orc_pool *p, uint32_t root, const hd_edit_desc *d) { return edit(*p, root, *d); }
uint32_t orc_edit_batch(orc_pool *p, uint32_t root, const hd_edit_desc *d, uint32_t n) {
	for (uint32_t i = 0; i < n; ++i)
		root = edit(*p, root, d[i]);
	return root;
}
// stats: [edit_nodes, edit_leaves, upserts, appended_nodes, appended_words, overflow, scan_words, read_words]
void orc_pool_stats(const orc_pool *p, uint64_t out[8]) {
	const Stats &s = p->st;
	uint64_t v[8] = {s.edit_nodes, s.edit_leaves,  s.upserts,    s.appended_nodes,
	                 s.appended_words, s.overflow, s.scan_words, s.read_words};
	std::copy(v, v + 8, out);
}
void orc_pool_reset_stats(orc_pool *p) { p->st = Stats{}; }

// Number of voxels an edit's VoxelInRange covers (implementation-independent "edited voxels" unit).
uint64_t orc_in_range_voxels(const hd_edit_desc *d, uint32_t voxel_level) {
	const uint64_t res = 1ull << voxel_level;
	switch (d->kind) {
	case HD_EDIT_AABB_FILL: {
		uint64_t v = 1;
		for (int i = 0; i < 3; ++i) {
			uint64_t lo = d->p0[i], hi = std::min<uint64_t>(d->p1[i], res);
			v *= hi > lo ? hi - lo : 0;
		}
		return v;
	}
	case HD_EDIT_SPHERE_FILL:
	case HD_EDIT_SPHERE_DIG: {
		// count lattice points of the ball clipped to the world, one (x,y) column at a time
		uint64_t total = 0;
		int64_t r = int64_t(std::sqrt(double(d->r2))) + 1;
		for (int64_t dx = -r; dx <= r; ++dx) {
			int64_t x = int64_t(d->p0[0]) + dx;
			if (x < 0 || x >= int64_t(res))
				continue;
			for (int64_t dy = -r; dy <= r; ++dy) {
				int64_t y = int64_t(d->p0[1]) + dy;
				if (y < 0 || y >= int64_t(res))
					continue;
				int64_t rem = int64_t(d->r2) - dx * dx - dy * dy;
				if (rem < 0)
					continue;
				int64_t dz = int64_t(std::sqrt(double(rem)));
				while (dz * dz > rem)
					--dz;
				while ((dz + 1) * (dz + 1) <= rem)
					++dz;
				int64_t z0 = std::max<int64_t>(0, int64_t(d->p0[2]) - dz),
				        z1 = std::min<int64_t>(int64_t(res) - 1, int64_t(d->p0[2]) + dz);
				if (z1 >= z0)
					total += uint64_t(z1 - z0 + 1);
			}
		}
		return total;
	}
	case HD_EDIT_TERRAIN_FILL: {
		terrain::Params tp = terrain::from_desc(d->aux, d->p0, d->p1);
		uint64_t total = 0;
		const uint64_t ext = tp.extent_bits ? std::min<uint64_t>(res, 1ull << tp.extent_bits) : res;
		for (uint64_t z = 0; z < ext; ++z)
			for (uint64_t x = 0; x < ext; ++x)
				total += std::min<uint64_t>(terrain::height(tp, uint32_t(x), uint32_t(z)), res);
		return total;
	}
	}
	return 0;
}
uint32_t orc_terrain_height(const hd_edit_desc *d, uint32_t x, uint32_t z) {
	terrain::Params tp = terrain::from_desc(d->aux, d->p0, d->p1);
	return terrain::height(tp, x, z);
}

// Canonical description of the DAG reachable from `root` in a flat word array.
// out = [content_hash, unique_by_pointer, unique_by_content, set_voxels, per-level unique_by_pointer x node_levels]
void orc_canonical(const uint32_t *words, uint32_t node_levels, uint32_t root, uint64_t *out) {
	Canon c{words, node_levels, {}, {}, {}};
	c.memo.resize(node_levels), c.by_content.resize(node_levels), c.voxels.resize(node_levels);
	for (uint32_t i = 0; i < 4 + node_levels; ++i)
		out[i] = 0;
	if (root == kNull)
		return;
	out[0] = c.hash(root, 0);
	for (uint32_t l = 0; l < node_levels; ++l) {
		out[1] += c.memo[l].size();
		out[2] += c.by_content[l].size();
		out[4 + l] = c.memo[l].size();
	}
	out[3] = c.count(root, 0);
}

// The same canonical description, level-synchronous and multi-threaded, for bench-scale pools (10^7..10^8 reachable
// nodes, where the memoised recursion above needs minutes and several GB of hash maps).  Reachable pointers are marked
// in one bitmap per level (top-down), ranked by prefix popcounts, and content hashes / voxel counts are computed
// bottom-up into dense arrays indexed by rank.  Values are identical to orc_canonical's (tests/test_oracle_golden.py).
extern "C++" {
namespace {
template <class F> void parallel_for(uint64_t n, uint32_t threads, uint64_t grain, F f) {
	if (n == 0)
		return;
	threads = std::max(1u, std::min<uint32_t>(threads, uint32_t((n + grain - 1) / grain)));
	std::atomic<uint64_t> next{0};
	auto work = [&]() {
		for (;;) {
			uint64_t b = next.fetch_add(grain);
			if (b >= n)
				return;
			f(b, std::min(n, b + grain));
		}
	};
	std::vector<std::thread> ts;
	for (uint32_t t = 1; t < threads; ++t)
		ts.emplace_back(work);
	work();
	for (auto &t : ts)
		t.join();
}

struct LevelSet {
	uint64_t base = 0, n_words = 0; // word range of the level in the pool
	uint64_t *bits = nullptr;       // one bit per pool word of the level
	std::vector<uint32_t> prefix;   // set bits before each 64-bit bitmap word
	uint64_t count = 0;
	uint64_t rank(uint32_t ptr) const {
		const uint64_t i = ptr - base;
		return prefix[i >> 6] + uint64_t(__builtin_popcountll(bits[i >> 6] & ((1ull << (i & 63)) - 1ull)));
	}
	template <class F> void for_each(uint32_t threads, F f) const { // f(ptr, rank)
		parallel_for((n_words + 63) / 64, threads, 1u << 14, [&](uint64_t b, uint64_t e) {
			for (uint64_t w = b; w < e; ++w) {
				uint64_t m = bits[w], r = prefix[w];
				while (m) {
					const uint32_t k = uint32_t(__builtin_ctzll(m));
					m &= m - 1;
					f(uint32_t(base + (w << 6) + k), r++);
				}
			}
		});
	}
};

uint64_t count_distinct(std::vector<uint64_t> &v, uint32_t threads) {
	if (v.size() < (1u << 16)) {
		std::sort(v.begin(), v.end());
		return uint64_t(std::unique(v.begin(), v.end()) - v.begin());
	}
	// partition by the top 8 bits, sort the partitions in parallel
	std::vector<uint64_t> cnt(257, 0), out(v.size());
	for (uint64_t x : v)
		++cnt[(x >> 56) + 1];
	for (int i = 0; i < 256; ++i)
		cnt[i + 1] += cnt[i];
	std::vector<uint64_t> pos(cnt.begin(), cnt.end() - 1);
	for (uint64_t x : v)
		out[pos[x >> 56]++] = x;
	std::atomic<uint64_t> distinct{0};
	parallel_for(256, threads, 1, [&](uint64_t b, uint64_t e) {
		for (uint64_t i = b; i < e; ++i) {
			std::sort(out.begin() + cnt[i], out.begin() + cnt[i + 1]);
			distinct += uint64_t(std::unique(out.begin() + cnt[i], out.begin() + cnt[i + 1]) - (out.begin() + cnt[i]));
		}
	});
	return distinct;
}
} // namespace
} // extern "C++"

void orc_canonical_fast(const uint32_t *words, const hd_config *cfg, uint32_t root, uint32_t threads, uint64_t *out) {
	Geometry g;
	const uint32_t L = cfg->node_levels;
	for (uint32_t i = 0; i < 4 + L; ++i)
		out[i] = 0;
	if (!make_geometry(*cfg, g) || root == kNull)
		return;
	threads = threads ? threads : std::max(1u, std::thread::hardware_concurrency());
	std::vector<LevelSet> S(L);
	for (uint32_t l = 0; l < L; ++l) {
		S[l].base = uint64_t(g.level_base[l]) << g.bucket_shift();
		S[l].n_words = (1ull << cfg->bucket_bits_each_level[l]) << g.bucket_shift();
		S[l].bits = static_cast<uint64_t *>(std::calloc((S[l].n_words + 63) / 64, 8));
	}
	auto finish_level = [&](LevelSet &s) {
		const uint64_t nw = (s.n_words + 63) / 64;
		s.prefix.resize(nw);
		uint64_t acc = 0;
		for (uint64_t w = 0; w < nw; ++w)
			s.prefix[w] = uint32_t(acc), acc += uint64_t(__builtin_popcountll(s.bits[w]));
		s.count = acc;
	};
	// top-down marking
	S[0].bits[(root - S[0].base) >> 6] |= 1ull << ((root - S[0].base) & 63);
	for (uint32_t l = 0; l < L; ++l) {
		finish_level(S[l]);
		if (l + 1 == L)
			break;
		LevelSet &c = S[l + 1];
		S[l].for_each(threads, [&](uint32_t ptr, uint64_t) {
			const uint32_t mask = words[ptr] & 0xFFu;
			for (uint32_t k = 1, n = 1 + uint32_t(__builtin_popcount(mask)); k < n; ++k) {
				const uint64_t i = words[ptr + k] - c.base;
				__atomic_fetch_or(&c.bits[i >> 6], 1ull << (i & 63), __ATOMIC_RELAXED);
			}
		});
	}
	// bottom-up hashes and voxel counts (Canon::hash / Canon::count above, value for value)
	std::vector<uint64_t> H_child, V_child, H, V;
	for (uint32_t l = L; l-- > 0;) {
		const LevelSet &s = S[l];
		H.assign(s.count, 0), V.assign(s.count, 0);
		if (l == L - 1) {
			s.for_each(threads, [&](uint32_t ptr, uint64_t r) {
				H[r] = mix64(0x1eafull ^ (uint64_t(words[ptr]) | uint64_t(words[ptr + 1]) << 32));
				V[r] = uint64_t(__builtin_popcount(words[ptr]) + __builtin_popcount(words[ptr + 1]));
			});
		} else {
			const LevelSet &c = S[l + 1];
			s.for_each(threads, [&](uint32_t ptr, uint64_t r) {
				const uint32_t mask = words[ptr] & 0xFFu;
				uint64_t h = mix64(0x1234567ull + mask + (uint64_t(l) << 32)), v = 0;
				uint32_t k = 1;
				for (uint32_t i = 0; i < 8; ++i)
					if (mask >> i & 1u) {
						const uint64_t cr = c.rank(words[ptr + k++]);
						h = mix64(h ^ (H_child[cr] + 0x9e3779b97f4a7c15ull * (i + 1)));
						v += V_child[cr];
					}
				H[r] = h, V[r] = v;
			});
		}
		out[1] += s.count, out[4 + l] = s.count;
		if (l == 0)
			out[0] = H[0], out[3] = V[0];
		H_child = H, V_child.swap(V);
		out[2] += count_distinct(H, threads);
	}
	for (auto &s : S)
		std::free(s.bits);
}

// Nodes physically stored in a pool (walk of every bucket, node sizes as find_node_in_span steps them,
// NodePool.hpp:79-91): out[level] = stored node count.  Used to prove a GC left no unreachable node behind.
void orc_count_stored_nodes(const uint32_t *words, const uint32_t *bucket_words, const hd_config *cfg, uint64_t *out) {
	Geometry g;
	if (!make_geometry(*cfg, g))
		return;
	const uint32_t wpp = g.words_per_page();
	for (uint32_t l = 0; l < g.node_levels(); ++l) {
		out[l] = 0;
		const bool is_leaf = l == g.node_levels() - 1;
		for (uint32_t b = 0; b < (1u << cfg->bucket_bits_each_level[l]); ++b) {
			const uint32_t bucket = g.level_base[l] + b, bw = bucket_words[bucket];
			const uint32_t *base = words + (size_t(bucket) << g.bucket_shift());
			for (uint32_t page = 0; page < bw; page += wpp) {
				const uint32_t end = std::min(page + wpp, bw);
				for (uint32_t it = page; it < end;) {
					const uint32_t sz = is_leaf ? 2u : (uint8_t(base[it]) ? 1u + uint32_t(__builtin_popcount(uint8_t(base[it]))) : 0u);
					if (sz == 0)
						break;
					++out[l];
					it += sz;
				}
			}
		}
	}
}

// Voxel lookup (the reconstructed Iterate intent, test/test.cpp:36-52,173-198): is voxel (x,y,z) set?
int orc_voxel_get(const uint32_t *words, uint32_t node_levels, uint32_t root, uint32_t x, uint32_t y, uint32_t z) {
	if (root == kNull)
		return 0;
	uint32_t ptr = root, voxel_level = node_levels + 1;
	for (uint32_t l = 0; l + 1 < node_levels; ++l) {
		uint32_t sh = voxel_level - 1 - l;
		uint32_t c = ((x >> sh) & 1u) | (((y >> sh) & 1u) << 1) | (((z >> sh) & 1u) << 2);
		uint32_t mask = words[ptr] & 0xFFu;
		if (!(mask >> c & 1u))
			return 0;
		ptr = words[ptr + 1 + __builtin_popcount(mask & ((1u << c) - 1u))];
	}
	uint32_t i = (x & 1u) | ((y & 1u) << 1) | ((z & 1u) << 2) | ((x & 2u) << 2) | ((y & 2u) << 3) | ((z & 2u) << 4);
	return (words[ptr + (i >> 5)] >> (i & 31u)) & 1u;
}

// Host pick ray — NodePoolTraversal::Traversal<float>, NodePoolTraversal.hpp:93-256
int orc_traverse(const uint32_t *words, uint32_t node_levels, uint32_t root, const float o[3], const float d[3],
                 float out[3]) {
	if (root == kNull)
		return 0;
	March m{};
	march(words, root, node_levels, false, 0.0f, 0.0f, o, d, m);
	if (!m.hit)
		return 0;
	for (int i = 0; i < 3; ++i) {
		float pos = m.pos[i];
		if (m.octant >> i & 1u) // undo mirroring, Traversal.hpp:246-251
			pos = 3.0f - m.scale_exp2 - pos;
		float v = m.o[i] + m.t_min * m.d[i];                 // o + t_min*d
		v = fmin2(fmax2(v, pos), pos + m.scale_exp2) - 1.0f; // glm::clamp = min(max(x,lo),hi)
		out[i] = v;
	}
	return 1;
}

// Frame trace — trace.frag main() for rows [row_begin,row_end) step row_step, n_threads workers (rows interleaved).
// Outputs are full-frame row-major planes (any may be NULL); returns the number of 32-bit words fetched (F·rays).
uint64_t orc_trace_frame_beam(const uint32_t *nodes, const uint32_t *color_nodes, const uint32_t *color_leaves,
                              const hd_trace_params *P, uint32_t row_begin, uint32_t row_end, uint32_t row_step,
                              uint32_t n_threads, uint32_t *rgba8, hd_hit_record *hits, uint32_t *iters, const float *beam,
                              uint32_t bw, uint32_t bh);
uint64_t orc_trace_frame(const uint32_t *nodes, const uint32_t *color_nodes, const uint32_t *color_leaves,
                         const hd_trace_params *P, uint32_t row_begin, uint32_t row_end, uint32_t row_step,
                         uint32_t n_threads, uint32_t *rgba8, hd_hit_record *hits, uint32_t *iters) {
	return orc_trace_frame_beam(nodes, color_nodes, color_leaves, P, row_begin, row_end, row_step, n_threads, rgba8, hits,
	                            iters, nullptr, 0, 0);
}

// Coarse beam image — shader/src/beam.frag:59-223 with the parameter block of src/rg/BeamPass.cpp:81-111
// (width/height = beam image size, proj_factor = the pass's own).  out[bw*bh] floats, +inf where the beam misses.
void orc_beam_frame(const uint32_t *nodes, const hd_trace_params *B, float *out) {
	for (uint32_t y = 0; y < B->height; ++y)
		for (uint32_t x = 0; x < B->width; ++x) {
			float cx = (float(x) + 0.5f) / float(B->width), cy = (float(y) + 0.5f) / float(B->height);
			cx = cx * 2.0f - 1.0f, cy = cy * 2.0f - 1.0f;
			float d[3];
			for (int i = 0; i < 3; ++i)
				d[i] = (B->look[i] - B->side[i] * cx) - B->up[i] * cy;
			normalize3(d);
			float t = fbits(0x7F800000u);
			if (B->dag_root != kNull) {
				March m{};
				march(nodes, B->dag_root, B->dag_leaf_level, true, B->proj_factor, 0.0f, B->pos, d, m);
				if (m.hit)
					t = fmax2(m.t_min - m.scale_exp2, 0.0f); // beam.frag:219
			}
			out[size_t(y) * B->width + x] = t;
		}
}

// trace.frag compiled with BEAM_OPTIMIZATION (beam == NULL: the plain shader)
uint64_t orc_trace_frame_beam(const uint32_t *nodes, const uint32_t *color_nodes, const uint32_t *color_leaves,
                              const hd_trace_params *P, uint32_t row_begin, uint32_t row_end, uint32_t row_step,
                              uint32_t n_threads, uint32_t *rgba8, hd_hit_record *hits, uint32_t *iters, const float *beam,
                              uint32_t bw, uint32_t bh) {
	Scene s{nodes, color_nodes, color_leaves, beam, bw, bh};
	std::atomic<uint64_t> total{0};
	if (n_threads == 0)
		n_threads = 1;
	if (row_step == 0)
		row_step = 1;
	auto work = [&](uint32_t tid) {
		uint64_t f = 0;
		uint32_t k = 0;
		for (uint32_t y = row_begin; y < row_end; y += row_step, ++k) {
			if (k % n_threads != tid)
				continue;
			for (uint32_t x = 0; x < P->width; ++x) {
				PixelOut po{};
				shade_pixel(s, *P, x, y, po);
				size_t at = size_t(y) * P->width + x;
				if (rgba8)
					rgba8[at] = po.rgba8;
				if (hits)
					hits[at] = po.rec;
				if (iters)
					iters[at] = po.iter;
				f += po.fetches;
			}
		}
		total += f;
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
	return total.load();
}

// Primary rays of a frame through the HOST tracer (Traversal<float>: no LOD, no colour), rays generated as
// trace.frag:366-377.  Mirror of ref_trace_frame_host in oracle/ref_harness.cpp; pins ray generation + traversal.
uint64_t orc_trace_frame_host(const uint32_t *nodes, uint32_t node_levels, const hd_trace_params *P, uint32_t row_begin,
                              uint32_t row_end, uint32_t row_step, uint32_t n_threads, uint8_t *hit, float *pos) {
	std::atomic<uint64_t> hits{0};
	if (!n_threads)
		n_threads = 1;
	if (!row_step)
		row_step = 1;
	auto work = [&](uint32_t tid) {
		uint64_t h = 0;
		uint32_t k = 0;
		for (uint32_t y = row_begin; y < row_end; y += row_step, ++k) {
			if (k % n_threads != tid)
				continue;
			for (uint32_t x = 0; x < P->width; ++x) {
				float cx = (float(x) + 0.5f) / float(P->width), cy = (float(y) + 0.5f) / float(P->height);
				cx = cx * 2.0f - 1.0f, cy = cy * 2.0f - 1.0f;
				float d[3], out[3];
				for (int i = 0; i < 3; ++i)
					d[i] = (P->look[i] - P->side[i] * cx) - P->up[i] * cy;
				normalize3(d);
				int r = orc_traverse(nodes, node_levels, P->dag_root, P->pos, d, out);
				size_t at = size_t(y) * P->width + x;
				if (hit)
					hit[at] = uint8_t(r);
				if (pos && r)
					pos[at * 3] = out[0], pos[at * 3 + 1] = out[1], pos[at * 3 + 2] = out[2];
				h += uint64_t(r);
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

// Colour lookup alone (checked against the reference's VBR iterator in oracle/_ref).  out = RGB floats.
void orc_color_fetch(const uint32_t *color_nodes, const uint32_t *color_leaves, uint32_t root, uint32_t voxel_level,
                     uint32_t leaf_level, uint32_t x, uint32_t y, uint32_t z, float out[3]) {
	Scene s{nullptr, color_nodes, color_leaves};
	uint32_t vox[3] = {x, y, z};
	uint64_t f = 0;
	Rgb c = color_fetch(s, root, voxel_level, leaf_level, vox, f);
	out[0] = c.r, out[1] = c.g, out[2] = c.b;
}

} // extern "C"
