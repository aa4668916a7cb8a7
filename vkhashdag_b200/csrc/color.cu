// color.cu — device colour pool and colour-aware edits on the GPU.
//
// Replaces, for the editors the reference ships (AABBEditor, SphereEditor<kFill>, SphereEditor<kPaint>,
// src/main.cpp:32-150), the colour side of an edit: VBREditorWrapper (include/hashdag/VBREditor.hpp:26-107) threading a
// colour-octree pointer and a VBRChunkWriter (VBRColor.hpp:368-487) through the edit recursion, and DAGColorPool's
// SetNode / FillNode / SetLeaf (src/DAGColorPool.hpp:147-204).  Word layout is the reference's.
//
// The reference rewrites each touched colour leaf as a sequential Morton-order stream.  Here the stream is never
// materialised sequentially: the colour every voxel of a touched leaf ends up with is a pure function of the OLD
// geometry path above it, the old colour chunk and the editor (k_voxels walks that chain per voxel), and the chunk
// the writer would have produced is the canonical run-length encoding of that colour sequence (a block starts where
// (colors, bits_per_weight) changes or a 2^14-voxel macro block starts — VBRChunkWriter::append, VBRColor.hpp:384-409),
// which two prefix sums and one emit kernel produce in parallel.  Above the colour-leaf level the octree is rebuilt
// level-synchronously (k_cdown / k_cup) with SetNode's collapse rules.
#include "common.cuh"
#include "editors.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace hd {

constexpr int kCB = 256;
constexpr uint32_t kTagNode = 0u, kTagColor = 1u, kTagLeaf = 2u, kTagNull = 3u; // src/DAGColorPool.hpp:26
constexpr uint32_t kMacroBits = 14u;                                          // VBRColor.hpp:46-49
constexpr uint32_t kBlkBits = 9u;      // the block path works on 8 x 8 x 8 voxel blocks (512 consecutive Morton indices)
constexpr uint32_t kPerVoxel = 0xFFFFFFFFu; // block key: not one colour, evaluate voxel by voxel (no RGB8 colour has a high byte)
constexpr uint32_t kCPending = 0xFFFFFFFEu; // not a valid pointer: tag 3 with data != 0 never occurs otherwise

__device__ __forceinline__ uint32_t ctag(uint32_t p) { return p >> 30; }
__device__ __forceinline__ uint32_t cdata(uint32_t p) { return p & 0x3FFFFFFFu; }

// FillNode stores RGB8Color{color.Get()}: u8 -> float/255 -> *255 -> truncation (Color.hpp:20-25, DAGColorPool.hpp:165-167)
__host__ __device__ inline uint32_t fill_node_rgb8(uint32_t rgb8) {
	uint32_t out = 0;
	for (int i = 0; i < 3; ++i) {
		const float f = float((rgb8 >> (8 * i)) & 0xFFu) / 255.0f;
		out |= (uint32_t(uint8_t(f * 255.0f)) & 0xFFu) << (8 * i);
	}
	return out;
}

struct ColorEdit {
	hd_edit_desc d;
	uint32_t rgb8;    // editor colour (VBRColor{RGB8Color}: colors = rgb8, bits_per_weight = 0)
	uint32_t paint;   // SphereEditor<kPaint>
	uint32_t voxel_level, node_levels, leaf_level;
};

// Editor::EditNode(config, coord, ptr, VBRColor &final_color) of src/main.cpp:47-56,107-126.
// fill = GetFill(octree pointer): has_fill / fill_rgb8.  Returns the edit type; color_set = "final_color has a value".
__device__ inline EditType color_edit_node(const ColorEdit &e, uint32_t bits, uint32_t x, uint32_t y, uint32_t z, bool ptr_null,
                                           bool has_fill, uint32_t fill_rgb8, bool &color_set) {
	EditType t = edit_node(e.d, bits, x, y, z);
	const bool same = has_fill && fill_rgb8 == e.rgb8;
	if (e.paint) {
		if (t == kFill) {
			color_set = true;
			t = kNotAffected;
		} else
			color_set = ptr_null || same;
		if (ptr_null)
			t = kNotAffected;
	} else
		color_set = t == kFill || ptr_null || same;
	return t;
}

// raw colour of voxel `m` (Morton index inside the colour leaf) in the chunk at `idx` — the lookup of trace.frag:304-349
// without the decode: colors word, bits per weight, weight.
__device__ inline void chunk_lookup(const uint32_t *__restrict__ lv, uint32_t idx, uint32_t m, uint32_t &colors, uint32_t &bpw,
                                    uint32_t &weight) {
	const uint32_t macro_cnt = lv[idx + 1], block_cnt_all = lv[idx + 2];
	const uint32_t macro_off = idx + 4, block_base = macro_off + (macro_cnt << 1), weight_off = block_base + (block_cnt_all << 1);
	const uint32_t macro_id = m >> kMacroBits;
	const uint32_t mx = lv[macro_off + (macro_id << 1)], my = lv[macro_off + (macro_id << 1) + 1];
	uint32_t block_off = block_base + (mx << 1);
	uint32_t block_cnt = macro_id + 1 < macro_cnt ? lv[macro_off + ((macro_id + 1) << 1)] - mx : block_cnt_all - mx;
	const uint32_t vid = m & ((1u << kMacroBits) - 1u);
	while (block_cnt) { // first block with voxel_index_offset > vid
		const uint32_t step = block_cnt >> 1;
		if ((lv[block_off + (step << 1) + 1] >> 18) <= vid)
			block_cnt -= step + 1, block_off += (step + 1) << 1;
		else
			block_cnt = step;
	}
	block_off -= 2;
	const uint32_t bx = lv[block_off], by = lv[block_off + 1];
	colors = bx, bpw = (by >> 16) & 3u, weight = 0;
	if (bpw) {
		const uint32_t bit = my + (by & 0xFFFFu) + (vid - (by >> 18)) * bpw, o = bit & 31u;
		uint32_t w = lv[weight_off + (bit >> 5)] >> o;
		if (o + bpw > 32)
			w |= lv[weight_off + (bit >> 5) + 1] << (32u - o);
		weight = w & ((1u << bpw) - 1u);
	}
}

// ---- octree part (levels <= leaf_level): BFS work items ----------------------------------------------------------
struct CItems {
	uint32_t n, cap;
	const uint32_t *n_dev; // device-resident item count (the one-sync octree pass) or NULL: then `n` is host-known
	__device__ __forceinline__ uint32_t count() const { return n_dev ? min(*n_dev, cap) : n; }
	uint32_t *geom;   // old geometry pointer of the node
	uint32_t *oct;    // colour-octree pointer of the node before the edit
	uint64_t *pos;    // x | y<<21 | z<<42
	uint32_t *parent; // (parent item << 3) | slot, 0xFFFFFFFF for the root
	uint32_t *child;  // [n*8] resulting child pointers (inner items)
	uint32_t *result; // resulting pointer of the item
};

__device__ __forceinline__ uint32_t oct_child(const uint32_t *__restrict__ cnodes, uint32_t p, uint32_t c) {
	const uint32_t t = ctag(p); // DAGColorPool::GetChild, DAGColorPool.hpp:139-143
	return t == kTagNode ? cnodes[(size_t(cdata(p)) << 3) | c] : (t == kTagColor ? p : HD_COLOR_NULL);
}

// classify one node of the octree part: returns 0 = final (ptr_out), 1 = becomes an inner item, 2 = becomes a leaf item
__device__ inline int classify_oct(const ColorEdit &e, uint32_t level, uint32_t x, uint32_t y, uint32_t z, uint32_t geom,
                                   uint32_t oct, uint32_t &ptr_out) {
	const bool has_fill = ctag(oct) == kTagColor;
	bool color_set;
	const EditType t = color_edit_node(e, e.voxel_level - level, x, y, z, geom == kNull, has_fill, cdata(oct), color_set);
	if (color_set) { // VBREditor.hpp:52-55: FillNode, final
		ptr_out = (kTagColor << 30) | fill_node_rgb8(e.rgb8);
		return 0;
	}
	if (t == kClear) { // VBREditor.hpp:56-57 (no colour editor returns kClear today; kept for completeness)
		ptr_out = HD_COLOR_NULL;
		return 0;
	}
	if (t == kProceed)
		return level == e.leaf_level ? 2 : 1;
	ptr_out = oct;
	return 0;
}

__global__ void k_croot(ColorEdit e, uint32_t geom_root, uint32_t oct_root, CItems inner, CItems leaf, uint32_t *counts,
                        uint32_t *root_out) {
	if (threadIdx.x || blockIdx.x)
		return;
	uint32_t res = oct_root;
	const int k = classify_oct(e, 0, 0, 0, 0, geom_root, oct_root, res);
	if (k == 0) {
		*root_out = res;
		return;
	}
	CItems &dst = k == 1 ? inner : leaf;
	counts[k - 1] = 1;
	dst.geom[0] = geom_root, dst.oct[0] = oct_root, dst.pos[0] = 0, dst.parent[0] = 0xFFFFFFFFu;
}

// thread per (inner item at `level`, child): VBREditorWrapper::EditNode for levels <= leaf_level (VBREditor.hpp:37-59)
__device__ __forceinline__ void cdown_pair(const ColorEdit &e, uint32_t level, const uint32_t *__restrict__ words,
                                           const uint32_t *__restrict__ cnodes, const CItems &in, const CItems &inner,
                                           const CItems &leaf, uint32_t *counts, uint32_t t) {
	const uint32_t item = t >> 3, c = t & 7u;
	const uint32_t g = in.geom[item];
	uint32_t child = kNull;
	if (g != kNull) {
		const uint32_t mask = words[g];
		if (mask >> c & 1u)
			child = words[g + 1u + __popc(mask & ((1u << c) - 1u))];
	}
	const uint32_t oct = oct_child(cnodes, in.oct[item], c);
	const uint64_t p = in.pos[item];
	const uint32_t x = ((uint32_t(p) & 0x1FFFFFu) << 1) | (c & 1u), y = ((uint32_t(p >> 21) & 0x1FFFFFu) << 1) | ((c >> 1) & 1u),
	               z = ((uint32_t(p >> 42) & 0x1FFFFFu) << 1) | ((c >> 2) & 1u);
	uint32_t res = oct;
	const int k = classify_oct(e, level + 1u, x, y, z, child, oct, res);
	if (k == 0) {
		in.child[size_t(item) * 8u + c] = res;
		return;
	}
	const CItems &dst = k == 1 ? inner : leaf;
	const uint32_t slot = atomicAdd(&counts[k - 1], 1u);
	if (slot >= dst.cap) {
		counts[3] = 1;
		in.child[size_t(item) * 8u + c] = oct;
		return;
	}
	dst.geom[slot] = child, dst.oct[slot] = oct;
	dst.pos[slot] = uint64_t(x) | (uint64_t(y) << 21) | (uint64_t(z) << 42);
	dst.parent[slot] = (item << 3) | c;
	in.child[size_t(item) * 8u + c] = kCPending;
}
__global__ void __launch_bounds__(kCB) k_cdown(ColorEdit e, uint32_t level, const uint32_t *__restrict__ words,
                                               const uint32_t *__restrict__ cnodes, CItems in, CItems inner, CItems leaf,
                                               uint32_t *counts) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if ((t >> 3) < in.count())
		cdown_pair(e, level, words, cnodes, in, inner, leaf, counts, t);
}

// The whole octree descent of a brush-sized edit in ONE launch: root classification, then the levels one after the other
// inside a single CTA (a brush touches a handful of octree nodes per level; a launch per level was a quarter of the
// colour pass's launches).  lvl_counts block l = what produced level l: [0] inner items, [1] leaves, [3] overflow.
struct CLevels {
	CItems lv[HD_MAX_NODE_LEVELS + 1];
};
constexpr int kOctThreads = 1024;
__global__ void __launch_bounds__(kOctThreads) k_cdown_all(ColorEdit e, uint32_t geom_root, uint32_t oct_root,
                                                           const uint32_t *__restrict__ words, const uint32_t *cnodes,
                                                           const __grid_constant__ CLevels L, CItems leaf, uint32_t *lvl_counts,
                                                           uint32_t *root_out) {
	const uint32_t LL = e.leaf_level;
	if (threadIdx.x == 0) {
		uint32_t res = oct_root;
		const int k = classify_oct(e, 0, 0, 0, 0, geom_root, oct_root, res);
		if (k == 0)
			*root_out = res;
		else {
			const CItems &dst = k == 1 ? L.lv[0] : leaf;
			lvl_counts[k - 1] = 1;
			dst.geom[0] = geom_root, dst.oct[0] = oct_root, dst.pos[0] = 0, dst.parent[0] = 0xFFFFFFFFu;
		}
	}
	__threadfence_block();
	__syncthreads();
	for (uint32_t l = 0; l < LL; ++l) {
		const uint32_t n = min(*(volatile const uint32_t *)(lvl_counts + 4u * l), L.lv[l].cap);
		if (n == 0u)
			break;
		for (uint32_t t = threadIdx.x; t < n * 8u; t += blockDim.x)
			cdown_pair(e, l, words, cnodes, L.lv[l], L.lv[l + 1u], leaf, lvl_counts + 4u * (l + 1u), t);
		__threadfence_block();
		__syncthreads();
	}
}

// JoinNode above the leaf level: DAGColorPool::SetNode (DAGColorPool.hpp:147-163).  Thread per inner item.
__device__ __forceinline__ void cup_item(const CItems &it, uint32_t *cnodes, uint32_t *ctr, uint64_t node_cap,
                                         uint32_t *parent_child, uint32_t *root_out, uint32_t i) {
	uint32_t ch[8];
	bool all_null = true, all_same = true;
	for (int c = 0; c < 8; ++c) {
		ch[c] = it.child[size_t(i) * 8u + c];
		all_null &= ctag(ch[c]) == kTagNull;
		all_same &= ch[c] == ch[0];
	}
	const uint32_t old = it.oct[i];
	uint32_t res;
	if (all_null)
		res = HD_COLOR_NULL;
	else if (ctag(ch[0]) == kTagColor && all_same)
		res = ch[0];
	else {
		bool same = ctag(old) == kTagNode;
		if (same)
			for (int c = 0; c < 8 && same; ++c)
				same = cnodes[(size_t(cdata(old)) << 3) | c] == ch[c];
		if (same)
			res = old;
		else {
			const uint32_t id = atomicAdd(&ctr[0], 1u);
			if ((uint64_t(id) + 1) * 8 > node_cap) {
				ctr[2] = 1;
				res = old; // out of space: keep the old pointer (DAGColorPool.hpp:161)
			} else {
				for (int c = 0; c < 8; ++c)
					cnodes[(size_t(id) << 3) | c] = ch[c];
				res = (kTagNode << 30) | id;
			}
		}
	}
	const uint32_t par = it.parent[i];
	if (par == 0xFFFFFFFFu)
		*root_out = res;
	else
		parent_child[par] = res;
}
__global__ void __launch_bounds__(kCB) k_cup(CItems it, uint32_t *cnodes, uint32_t *ctr, uint64_t node_cap, uint32_t *parent_child,
                                             uint32_t *root_out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < it.count())
		cup_item(it, cnodes, ctr, node_cap, parent_child, root_out, i);
}
// every inner level from `top` - 1 down to the root in one launch (one CTA; item counts are host-known by now)
__global__ void __launch_bounds__(kOctThreads) k_cup_all(const __grid_constant__ CLevels L, uint32_t top, uint32_t *cnodes,
                                                         uint32_t *ctr, uint64_t node_cap, uint32_t *root_out) {
	for (uint32_t l = top; l-- > 0u;) {
		const uint32_t n = L.lv[l].n;
		for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
			cup_item(L.lv[l], cnodes, ctr, node_cap, l ? L.lv[l - 1u].child : nullptr, root_out, i);
		__threadfence_block();
		__syncthreads();
	}
}

// Exclusive prefix sums of TWO arrays of up to 32 K entries in one launch of one CTA (the block path's run and weight-bit
// counts of a brush that touches a few colour leaves): tiles of 4 096 entries, four consecutive entries per thread
// (coalesced), a running carry between tiles.  Larger inputs take exclusive_scan (a first version walked 256 strided
// entries per thread and cost 110 - 230 us on a 27-leaf brush).
constexpr uint32_t kSmallScan = 1u << 15;
__global__ void __launch_bounds__(1024) k_scan2_small(const uint32_t *__restrict__ in0, uint32_t *out0,
                                                      const uint32_t *__restrict__ in1, uint32_t *out1, uint32_t n) {
	__shared__ uint32_t s_warp[2][32], s_carry[2];
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0)
		s_carry[0] = s_carry[1] = 0u;
	__syncthreads();
	for (uint32_t base = 0; base < n; base += 4096u) {
		const uint32_t i0 = base + threadIdx.x * 4u;
		uint32_t a[4], b[4];
#pragma unroll
		for (int k = 0; k < 4; ++k)
			a[k] = i0 + k < n ? in0[i0 + k] : 0u, b[k] = i0 + k < n ? in1[i0 + k] : 0u;
		const uint32_t sa = a[0] + a[1] + a[2] + a[3], sb = b[0] + b[1] + b[2] + b[3];
		uint32_t ia = sa, ib = sb;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, ia, d), y = __shfl_up_sync(0xFFFFFFFFu, ib, d);
			if (lane >= uint32_t(d))
				ia += x, ib += y;
		}
		if (lane == 31u)
			s_warp[0][warp] = ia, s_warp[1][warp] = ib;
		__syncthreads();
		if (warp == 0u) {
			const uint32_t ta = s_warp[0][lane], tb = s_warp[1][lane];
			uint32_t xa = ta, xb = tb;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, xa, d), y = __shfl_up_sync(0xFFFFFFFFu, xb, d);
				if (lane >= uint32_t(d))
					xa += x, xb += y;
			}
			s_warp[0][lane] = xa - ta, s_warp[1][lane] = xb - tb; // exclusive over the warps
		}
		__syncthreads();
		uint32_t ra = s_carry[0] + s_warp[0][warp] + ia - sa, rb = s_carry[1] + s_warp[1][warp] + ib - sb;
#pragma unroll
		for (int k = 0; k < 4; ++k)
			if (i0 + k < n) {
				out0[i0 + k] = ra, out1[i0 + k] = rb;
				ra += a[k], rb += b[k];
			}
		__syncthreads();
		if (threadIdx.x == 1023u)
			s_carry[0] = ra, s_carry[1] = rb; // the last thread's running sums = carry + the tile's totals
		__syncthreads();
	}
}

// ---- leaf part: per-voxel colour, then canonical VBR encoding -------------------------------------------------------
// thread per (leaf item, Morton index): the colour the reference's writer would emit for that voxel
// (VBREditorWrapper::EditNode below the leaf level + EditVoxel, VBREditor.hpp:60-80; editors main.cpp:47-69,107-149).
__global__ void __launch_bounds__(kCB) k_voxels(ColorEdit e, const uint32_t *__restrict__ words, const uint32_t *__restrict__ cleaves,
                                                CItems leaf, uint32_t first_leaf, uint32_t n_leaves, uint32_t sbits,
                                                uint32_t *colors, uint8_t *bw, const uint32_t *__restrict__ block_list,
                                                uint32_t n_listed) {
	const uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
	uint32_t li, m;
	if (block_list) { // block path: only the voxels of the listed (mixed) 8^3 blocks are evaluated, output index = t
		const uint32_t slot = uint32_t(t >> kBlkBits);
		if (slot >= n_listed)
			return;
		const uint32_t blk = block_list[slot], bbits = sbits - kBlkBits;
		li = blk >> bbits, m = ((blk & ((1u << bbits) - 1u)) << kBlkBits) | (uint32_t(t) & ((1u << kBlkBits) - 1u));
	} else
		li = uint32_t(t >> sbits), m = uint32_t(t & ((1ull << sbits) - 1ull));
	if (li >= n_leaves)
		return;
	const uint32_t item = first_leaf + li;
	const uint32_t oct = leaf.oct[item];
	const bool has_fill = ctag(oct) == kTagColor, has_src = ctag(oct) == kTagLeaf;
	const uint32_t fill_rgb8 = cdata(oct), src = cdata(oct);
	const uint64_t p = leaf.pos[item];
	uint32_t x = uint32_t(p) & 0x1FFFFFu, y = uint32_t(p >> 21) & 0x1FFFFFu, z = uint32_t(p >> 42) & 0x1FFFFFu;
	uint32_t g = leaf.geom[item];
	uint32_t out_c = 0, out_b = 0, out_w = 0;
	bool done = false;
	// geometry levels below the colour leaf: nodes at levels leaf_level+1 .. node_levels-1
	for (uint32_t L = e.leaf_level + 1u; L < e.node_levels && !done; ++L) {
		const uint32_t c = (m >> (3u * (e.voxel_level - L))) & 7u;
		uint32_t child = kNull;
		if (g != kNull) {
			const uint32_t mask = words[g];
			if (mask >> c & 1u)
				child = words[g + 1u + __popc(mask & ((1u << c) - 1u))];
		}
		x = (x << 1) | (c & 1u), y = (y << 1) | ((c >> 1) & 1u), z = (z << 1) | ((c >> 2) & 1u);
		bool color_set;
		const EditType et = color_edit_node(e, e.voxel_level - L, x, y, z, child == kNull, has_fill, fill_rgb8, color_set);
		if (color_set) { // writer->Push(color, subtree)
			out_c = e.rgb8, done = true;
		} else if (et != kProceed) { // writer->Copy(subtree, fill)
			if (has_src)
				chunk_lookup(cleaves, src, m, out_c, out_b, out_w);
			else
				out_c = has_fill ? fill_rgb8 : 0u;
			done = true;
		}
		g = child;
	}
	if (!done) { // EditVoxel through writer->Edit (VBREditor.hpp:71-80; main.cpp:64-69,143-149)
		const uint32_t i = m & 63u;
		const uint32_t vx = (x << 2) | ((i >> 2) & 2u) | (i & 1u), vy = (y << 2) | ((i >> 3) & 2u) | ((i >> 1) & 1u),
		               vz = (z << 2) | ((i >> 4) & 2u) | ((i >> 2) & 1u);
		const bool voxel = (words[g + (i >> 5)] >> (i & 31u)) & 1u;
		if (voxel_in_range(e.d, vx, vy, vz) || !voxel)
			out_c = e.rgb8;
		else if (has_src)
			chunk_lookup(cleaves, src, m, out_c, out_b, out_w);
		else
			out_c = has_fill ? fill_rgb8 : 0u;
	}
	colors[t] = out_c;
	bw[t] = uint8_t(out_b | (out_w << 2));
}

// block-start flags and weight bit counts (VBRChunkWriter::append: a block starts at a macro block or where
// (colors, bits_per_weight) differ from the previous voxel)
__global__ void __launch_bounds__(kCB) k_flags(const uint32_t *__restrict__ colors, const uint8_t *__restrict__ bw, uint64_t n,
                                               uint32_t sbits, uint32_t *flag, uint32_t *bits) {
	const uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
	if (t >= n)
		return;
	const uint32_t b = bw[t] & 3u;
	const uint32_t m = uint32_t(t & ((1ull << sbits) - 1ull)); // Morton index inside the leaf (m == 0 starts a chunk)
	bool start = (m & ((1u << kMacroBits) - 1u)) == 0u;
	if (!start)
		start = colors[t] != colors[t - 1] || b != (bw[t - 1] & 3u);
	flag[t] = start ? 1u : 0u;
	bits[t] = b;
}

// ---- exclusive prefix sum over u32 (three-kernel scan, 1024 elements per CTA) ----------------------------------------
__global__ void __launch_bounds__(256) k_scan_block(const uint32_t *__restrict__ in, uint32_t *out, uint32_t *block_sums, uint64_t n) {
	__shared__ uint32_t s[1024];
	__shared__ uint32_t warp_sums[8];
	const uint64_t base = uint64_t(blockIdx.x) * 1024u;
	uint32_t v[4], sum = 0;
	for (int k = 0; k < 4; ++k) {
		const uint64_t i = base + threadIdx.x * 4u + k;
		v[k] = i < n ? in[i] : 0u;
		sum += v[k];
	}
	// exclusive scan of per-thread sums across the CTA
	uint32_t incl = sum;
	const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
		if (lane >= uint32_t(d))
			incl += o;
	}
	if (lane == 31)
		warp_sums[w] = incl;
	__syncthreads();
	uint32_t woff = 0;
	for (uint32_t k = 0; k < w; ++k)
		woff += warp_sums[k];
	uint32_t run = woff + incl - sum;
	for (int k = 0; k < 4; ++k) {
		const uint64_t i = base + threadIdx.x * 4u + k;
		if (i < n)
			out[i] = run;
		run += v[k];
	}
	if (threadIdx.x == 255 && block_sums)
		block_sums[blockIdx.x] = run;
	(void)s;
}
__global__ void __launch_bounds__(256) k_scan_add(uint32_t *out, const uint32_t *__restrict__ block_offsets, uint64_t n) {
	const uint64_t i = uint64_t(blockIdx.x) * 1024u + threadIdx.x * 4u;
	const uint32_t o = block_offsets[blockIdx.x];
	for (int k = 0; k < 4; ++k)
		if (i + k < n)
			out[i + k] += o;
}

hd_status exclusive_scan(hd_pool *p, const uint32_t *in, uint32_t *out, uint64_t n, cudaStream_t stream) {
	if (n == 0)
		return HD_OK;
	cudaStream_t st = stream ? stream : p->stream;
	const uint64_t blocks = (n + 1023) / 1024;
	uint32_t *sums = nullptr;
	HD_CUDA_TRY(cudaMallocAsync(&sums, std::max<uint64_t>(blocks, 1) * 4, st));
	k_scan_block<<<uint32_t(blocks), 256, 0, st>>>(in, out, sums, n);
	HD_LAUNCH_CHECK();
	if (blocks > 1) {
		hd_status s = exclusive_scan(p, sums, sums, blocks, st);
		if (s != HD_OK)
			return s;
		k_scan_add<<<uint32_t(blocks), 256, 0, st>>>(out, sums, n);
		HD_LAUNCH_CHECK();
	}
	HD_CUDA_TRY(cudaFreeAsync(sums, st));
	return HD_OK;
}

// words each leaf of the batch may have to append (0 when it will be rewritten in place); summed into *total
__global__ void k_leaf_size(CItems leaf, uint32_t first_leaf, uint32_t n_leaves, uint32_t sbits, uint32_t ebits, const uint32_t *__restrict__ flag,
                            const uint32_t *__restrict__ fscan, const uint32_t *__restrict__ bits, const uint32_t *__restrict__ bscan,
                            const uint32_t *__restrict__ cleaves, unsigned long long *total) {
	const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
	if (li >= n_leaves)
		return;
	// the scan arrays hold 2^ebits elements per leaf: one per voxel (ebits == sbits) or one per 8^3 block (block path)
	const uint64_t S = 1ull << sbits, a = uint64_t(li) << ebits, last = a + (1ull << ebits) - 1;
	const uint32_t block_cnt = fscan[last] + flag[last] - fscan[a], total_bits = bscan[last] + bits[last] - bscan[a];
	const uint32_t data = uint32_t((S + (1u << kMacroBits) - 1u) >> kMacroBits) * 2u + block_cnt * 2u + ((total_bits + 31u) >> 5) + 4u;
	uint32_t append = (data & 1u) ? data + 1u : data;
	const uint32_t old = leaf.oct[first_leaf + li];
	if (ctag(old) == kTagLeaf) {
		const uint32_t block = cleaves[cdata(old)];
		append = data <= block ? 0u : max(block << 1, append);
	}
	if (append)
		atomicAdd(total, (unsigned long long)append);
}

// per leaf: sizes, allocation (in place if it fits, else append with capacity doubling: DAGColorPool::SetLeaf,
// DAGColorPool.hpp:173-204 with keep_history = false), header words; thread per leaf
__global__ void k_leaf_alloc(CItems leaf, uint32_t first_leaf, uint32_t n_leaves, uint32_t sbits, uint32_t ebits, const uint32_t *__restrict__ flag,
                             const uint32_t *__restrict__ fscan, const uint32_t *__restrict__ bits, const uint32_t *__restrict__ bscan,
                             uint32_t *cleaves, uint32_t *ctr, uint64_t leaf_cap, uint32_t *chunk_idx) {
	const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
	if (li >= n_leaves)
		return;
	const uint64_t S = 1ull << sbits, a = uint64_t(li) << ebits, last = a + (1ull << ebits) - 1;
	const uint32_t block_cnt = fscan[last] + flag[last] - fscan[a];
	const uint32_t total_bits = bscan[last] + bits[last] - bscan[a];
	const uint32_t macro_cnt = uint32_t((S + (1u << kMacroBits) - 1u) >> kMacroBits);
	const uint32_t weight_words = (total_bits + 31u) >> 5;
	const uint32_t data = macro_cnt * 2u + block_cnt * 2u + weight_words + 4u;
	uint32_t append = (data & 1u) ? data + 1u : data;
	const uint32_t item = first_leaf + li, old = leaf.oct[item];
	uint32_t idx = 0xFFFFFFFFu;
	if (ctag(old) == kTagLeaf) {
		const uint32_t oi = cdata(old), block = cleaves[oi];
		if (data <= block)
			idx = oi; // rewrite in place
		else
			append = max(block << 1, append);
	}
	uint32_t res;
	if (idx == 0xFFFFFFFFu) {
		idx = atomicAdd(&ctr[1], append);
		if (uint64_t(idx) + append > leaf_cap || uint64_t(idx) + append >= (1ull << 30)) {
			ctr[2] = 1;
			chunk_idx[li] = 0xFFFFFFFFu;
			leaf.result[item] = old; // out of space: keep the old pointer (DAGColorPool.hpp:196-198)
			return;
		}
		cleaves[idx] = append;
	}
	res = (kTagLeaf << 30) | idx;
	cleaves[idx + 1] = macro_cnt, cleaves[idx + 2] = block_cnt, cleaves[idx + 3] = weight_words;
	chunk_idx[li] = idx;
	leaf.result[item] = res;
}

// replica sync: chunks that were rewritten IN PLACE below the already-synchronised part of the leaf array are not covered
// by the append range, so their word indices go on the pool's dirty list (sync.cu packs them)
__global__ void __launch_bounds__(kCB) k_color_mark_dirty(const uint32_t *__restrict__ chunk_idx, uint32_t n_leaves, uint32_t synced_words,
                                                          uint32_t *list, uint32_t *ctr, uint32_t cap) {
	const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
	if (li >= n_leaves)
		return;
	const uint32_t idx = chunk_idx[li];
	if (idx == 0xFFFFFFFFu || idx >= synced_words)
		return;
	const uint32_t slot = atomicAdd(&ctr[0], 1u);
	if (slot < cap)
		list[slot] = idx;
	else
		ctr[1] = 1u;
}

__global__ void __launch_bounds__(kCB) k_zero_weights(uint32_t n_leaves, const uint32_t *__restrict__ chunk_idx, uint32_t *cleaves) {
	for (uint32_t li = blockIdx.x; li < n_leaves; li += gridDim.x) {
		const uint32_t idx = chunk_idx[li];
		if (idx == 0xFFFFFFFFu)
			continue;
		const uint32_t off = idx + 4u + cleaves[idx + 1] * 2u + cleaves[idx + 2] * 2u, n = cleaves[idx + 3];
		for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
			cleaves[off + i] = 0u;
	}
}

// macro blocks, block headers and weight bits of every chunk (VBRMacroBlock / VBRBlockHeader, VBRColor.hpp:230-250)
__global__ void __launch_bounds__(kCB) k_emit(uint64_t n, uint32_t sbits, const uint32_t *__restrict__ colors,
                                              const uint8_t *__restrict__ bw, const uint32_t *__restrict__ flag,
                                              const uint32_t *__restrict__ fscan, const uint32_t *__restrict__ bscan,
                                              const uint32_t *__restrict__ chunk_idx, uint32_t *cleaves) {
	const uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
	if (t >= n)
		return;
	const uint32_t li = uint32_t(t >> sbits), idx = chunk_idx[li];
	if (idx == 0xFFFFFFFFu)
		return;
	const uint64_t a = uint64_t(li) << sbits;
	const uint32_t m = uint32_t(t - a);
	const uint32_t macro_cnt = cleaves[idx + 1], block_cnt = cleaves[idx + 2];
	const uint32_t macro_off = idx + 4u, block_off = macro_off + macro_cnt * 2u, weight_off = block_off + block_cnt * 2u;
	const uint32_t blk = fscan[t] - fscan[a], bit = bscan[t] - bscan[a]; // block index / weight bit index inside the chunk
	const uint64_t macro_first = a + (uint64_t(m >> kMacroBits) << kMacroBits);
	const uint32_t macro_bit = bscan[macro_first] - bscan[a];
	if ((m & ((1u << kMacroBits) - 1u)) == 0u) {
		cleaves[macro_off + ((m >> kMacroBits) << 1)] = blk;      // first_block
		cleaves[macro_off + ((m >> kMacroBits) << 1) + 1] = bit; // weight_start
	}
	const uint32_t b = bw[t] & 3u;
	if (flag[t]) {
		cleaves[block_off + (blk << 1)] = colors[t];
		cleaves[block_off + (blk << 1) + 1] = ((m & ((1u << kMacroBits) - 1u)) << 18) | (b << 16) | (bit - macro_bit);
	}
	if (b) { // VBRBitsetWriter::Push(weight, bits): LSB first, may straddle two words
		const uint32_t w = uint32_t(bw[t]) >> 2, o = bit & 31u;
		atomicOr(&cleaves[weight_off + (bit >> 5)], w << o);
		if (o + b > 32u)
			atomicOr(&cleaves[weight_off + (bit >> 5) + 1], w >> (32u - o));
	}
}

// ---- block path: re-encode at 8^3-block granularity ----------------------------------------------------------------------
// The voxel path above touches every voxel of every touched colour leaf (2 M voxels per 128^3 leaf) although most of a leaf
// ends up as long single-colour runs: everything the editor covers, every EMPTY region (an empty voxel takes the editor's
// colour, main.cpp:64-69,143-149) and every untouched solid region that was one colour before.  Here a thread first classifies
// each 8^3 block (512 consecutive Morton indices) by walking the old geometry and the editor's node rule down to the block:
// either the whole block is ONE colour with no weights (its key), or it is listed for the per-voxel evaluation.  Block starts,
// prefix sums, sizes and the emit then run over blocks; only listed blocks are expanded to voxels.  The chunk produced is the
// same canonical encoding, word for word (tests/test_gpu_color_edit.py compares against the reference's writer).

// is [m0, m0 + 512) covered by ONE run without weights of the old chunk?  (a block never straddles a macro block)
__device__ inline bool chunk_uniform_run(const uint32_t *__restrict__ lv, uint32_t idx, uint32_t m0, uint32_t &colors) {
	const uint32_t macro_cnt = lv[idx + 1], block_cnt_all = lv[idx + 2];
	const uint32_t macro_off = idx + 4, block_base = macro_off + (macro_cnt << 1);
	const uint32_t macro_id = m0 >> kMacroBits;
	const uint32_t mx = lv[macro_off + (macro_id << 1)];
	uint32_t block_off = block_base + (mx << 1);
	uint32_t block_cnt = macro_id + 1 < macro_cnt ? lv[macro_off + ((macro_id + 1) << 1)] - mx : block_cnt_all - mx;
	const uint32_t end_off = block_off + (block_cnt << 1);
	const uint32_t vid = m0 & ((1u << kMacroBits) - 1u);
	while (block_cnt) { // first run with voxel_index_offset > vid
		const uint32_t step = block_cnt >> 1;
		if ((lv[block_off + (step << 1) + 1] >> 18) <= vid)
			block_cnt -= step + 1, block_off += (step + 1) << 1;
		else
			block_cnt = step;
	}
	// block_off = the run after the one that holds vid; that run must start at or behind the end of the block
	if (block_off < end_off && (lv[block_off + 1] >> 18) < vid + (1u << kBlkBits))
		return false;
	const uint32_t by = lv[block_off - 1];
	colors = lv[block_off - 2];
	return ((by >> 16) & 3u) == 0u;
}

struct BlockArrays {
	uint32_t *key;      // colour of a one-colour block, or kPerVoxel
	uint32_t *first_c, *last_c; // colors word of the block's first / last voxel
	uint8_t *first_b, *last_b;  // bits per weight of the first / last voxel
	uint32_t *runs;     // run starts inside the block (after k_block_flags: including the start at the block's first voxel)
	uint32_t *bits;     // weight bits of the block
	uint32_t *slot;     // index of a per-voxel block in the compact list
	uint8_t *start;     // a run starts at the block's first voxel
};

// thread per (leaf, 8^3 block): classification (the walk of k_voxels, stopped at the block's node)
__global__ void __launch_bounds__(kCB) k_cblock(ColorEdit e, const uint32_t *__restrict__ words, const uint32_t *__restrict__ cleaves,
                                                CItems leaf, uint32_t first_leaf, uint32_t n_leaves, uint32_t sbits, BlockArrays B,
                                                uint32_t *list, uint32_t *list_count) {
	const uint32_t bbits = sbits - kBlkBits;
	const uint32_t blk = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t li = blk >> bbits;
	uint32_t key = kPerVoxel;
	const bool valid = li < n_leaves;
	if (valid) {
		const uint32_t m0 = (blk & ((1u << bbits) - 1u)) << kBlkBits, item = first_leaf + li;
		const uint32_t oct = leaf.oct[item];
		const bool has_fill = ctag(oct) == kTagColor, has_src = ctag(oct) == kTagLeaf;
		const uint32_t fill_rgb8 = cdata(oct), src = cdata(oct);
		const uint64_t p = leaf.pos[item];
		uint32_t x = uint32_t(p) & 0x1FFFFFu, y = uint32_t(p >> 21) & 0x1FFFFFu, z = uint32_t(p >> 42) & 0x1FFFFFu;
		uint32_t g = leaf.geom[item];
		const uint32_t block_level = e.voxel_level - 3u;
		for (uint32_t L = e.leaf_level + 1u; L <= block_level; ++L) {
			const uint32_t c = (m0 >> (3u * (e.voxel_level - L))) & 7u;
			uint32_t child = kNull;
			if (g != kNull) {
				const uint32_t mask = words[g];
				if (mask >> c & 1u)
					child = words[g + 1u + __popc(mask & ((1u << c) - 1u))];
			}
			x = (x << 1) | (c & 1u), y = (y << 1) | ((c >> 1) & 1u), z = (z << 1) | ((c >> 2) & 1u);
			bool color_set;
			const EditType et = color_edit_node(e, e.voxel_level - L, x, y, z, child == kNull, has_fill, fill_rgb8, color_set);
			if (color_set) { // writer->Push(color, subtree): the whole block
				key = e.rgb8;
				break;
			}
			if (et != kProceed) { // writer->Copy(subtree, fill): one colour only if the old chunk has one run there
				if (!has_src)
					key = has_fill ? fill_rgb8 : 0u;
				else {
					uint32_t c_old;
					if (chunk_uniform_run(cleaves, src, m0, c_old))
						key = c_old;
				}
				break;
			}
			g = child;
		}
		B.key[blk] = key;
		if (key != kPerVoxel) {
			B.first_c[blk] = B.last_c[blk] = key;
			B.first_b[blk] = B.last_b[blk] = 0;
			B.runs[blk] = 0u, B.bits[blk] = 0u;
		}
	}
	// compact list of the per-voxel blocks (warp-aggregated)
	const bool want = valid && key == kPerVoxel;
	const uint32_t vote = __ballot_sync(0xFFFFFFFFu, want), lane = threadIdx.x & 31u;
	if (vote) {
		uint32_t base = 0;
		if (lane == uint32_t(__ffs(vote) - 1))
			base = atomicAdd(list_count, __popc(vote));
		base = __shfl_sync(0xFFFFFFFFu, base, __ffs(vote) - 1);
		if (want) {
			const uint32_t sl = base + __popc(vote & ((1u << lane) - 1u));
			list[sl] = blk, B.slot[blk] = sl;
		}
	}
}

// warp per listed block: first / last key, run starts inside the block, weight bits (from the per-voxel arrays)
__global__ void __launch_bounds__(kCB) k_block_summary(uint32_t n_listed, const uint32_t *__restrict__ list, const uint32_t *__restrict__ colors,
                                                       const uint8_t *__restrict__ bw, BlockArrays B) {
	const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
	if (slot >= n_listed)
		return;
	const uint32_t base = slot << kBlkBits;
	uint32_t starts = 0, bits = 0;
	for (uint32_t v = lane; v < (1u << kBlkBits); v += 32u) { // coalesced: consecutive lanes, consecutive voxels
		const uint32_t c = colors[base + v], b = bw[base + v] & 3u;
		bits += b;
		if (v)
			starts += (c != colors[base + v - 1] || b != (bw[base + v - 1] & 3u)) ? 1u : 0u;
	}
	for (int d = 16; d; d >>= 1)
		starts += __shfl_xor_sync(0xFFFFFFFFu, starts, d), bits += __shfl_xor_sync(0xFFFFFFFFu, bits, d);
	if (lane == 0) {
		const uint32_t blk = list[slot], last = base + (1u << kBlkBits) - 1u;
		B.first_c[blk] = colors[base], B.first_b[blk] = bw[base] & 3u;
		B.last_c[blk] = colors[last], B.last_b[blk] = bw[last] & 3u;
		B.runs[blk] = starts, B.bits[blk] = bits;
	}
}

// thread per block: does a run start at the block's first voxel?  (macro block starts, or the key differs from the previous voxel)
__global__ void __launch_bounds__(kCB) k_block_flags(uint32_t n_blocks, uint32_t sbits, BlockArrays B) {
	const uint32_t blk = blockIdx.x * blockDim.x + threadIdx.x;
	if (blk >= n_blocks)
		return;
	const uint32_t b = blk & ((1u << (sbits - kBlkBits)) - 1u);
	bool start = ((b << kBlkBits) & ((1u << kMacroBits) - 1u)) == 0u;
	if (!start)
		start = B.first_c[blk] != B.last_c[blk - 1] || B.first_b[blk] != B.last_b[blk - 1];
	B.start[blk] = start ? 1 : 0;
	B.runs[blk] += start ? 1u : 0u;
}

// thread per block: macro-block table entries, and the header of a one-colour block that starts a run
__global__ void __launch_bounds__(kCB) k_emit_blocks(uint32_t n_blocks, uint32_t sbits, BlockArrays B, const uint32_t *__restrict__ rscan,
                                                     const uint32_t *__restrict__ wscan, const uint32_t *__restrict__ chunk_idx,
                                                     uint32_t *cleaves) {
	const uint32_t blk = blockIdx.x * blockDim.x + threadIdx.x;
	if (blk >= n_blocks)
		return;
	const uint32_t bbits = sbits - kBlkBits, li = blk >> bbits, idx = chunk_idx[li];
	if (idx == 0xFFFFFFFFu)
		return;
	const uint32_t a = li << bbits, m0 = (blk - a) << kBlkBits;
	const uint32_t macro_cnt = cleaves[idx + 1];
	const uint32_t macro_off = idx + 4u, block_off = macro_off + macro_cnt * 2u;
	const uint32_t run = rscan[blk] - rscan[a], bit = wscan[blk] - wscan[a];
	if ((m0 & ((1u << kMacroBits) - 1u)) == 0u) {
		cleaves[macro_off + ((m0 >> kMacroBits) << 1)] = run;     // first_block
		cleaves[macro_off + ((m0 >> kMacroBits) << 1) + 1] = bit; // weight_start
	}
	if (B.key[blk] != kPerVoxel && B.start[blk]) {
		const uint32_t macro_first = a + ((m0 >> kMacroBits) << (kMacroBits - kBlkBits));
		cleaves[block_off + (run << 1)] = B.key[blk];
		cleaves[block_off + (run << 1) + 1] = ((m0 & ((1u << kMacroBits) - 1u)) << 18) | (bit - (wscan[macro_first] - wscan[a]));
	}
}

// CTA of 512 threads per listed block: headers and weight bits of its voxels (k_emit restricted to one block, with the
// run / bit indices inside the block from a CTA-wide scan)
__global__ void __launch_bounds__(1u << kBlkBits) k_emit_listed(uint32_t sbits, const uint32_t *__restrict__ list, const uint32_t *__restrict__ colors,
                                                                const uint8_t *__restrict__ bw, BlockArrays B, const uint32_t *__restrict__ rscan,
                                                                const uint32_t *__restrict__ wscan, const uint32_t *__restrict__ chunk_idx,
                                                                uint32_t *cleaves) {
	__shared__ uint32_t s_runs[17], s_bits[17];
	const uint32_t slot = blockIdx.x, v = threadIdx.x, lane = v & 31u, warp = v >> 5;
	const uint32_t blk = list[slot], bbits = sbits - kBlkBits, li = blk >> bbits, idx = chunk_idx[li];
	const uint32_t t = (slot << kBlkBits) + v;
	const uint32_t c = colors[t], b = bw[t] & 3u, w = uint32_t(bw[t]) >> 2;
	const bool flag = v ? (c != colors[t - 1] || b != (bw[t - 1] & 3u)) : B.start[blk] != 0;
	// exclusive scans of (flag, b) over the 512 threads
	uint32_t fr = flag ? 1u : 0u, fb = b;
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, fr, d), q = __shfl_up_sync(0xFFFFFFFFu, fb, d);
		if (lane >= uint32_t(d))
			fr += o, fb += q;
	}
	if (lane == 31u)
		s_runs[warp + 1] = fr, s_bits[warp + 1] = fb;
	__syncthreads();
	if (v == 0) {
		s_runs[0] = s_bits[0] = 0u;
		for (int k = 1; k <= 16; ++k)
			s_runs[k] += s_runs[k - 1], s_bits[k] += s_bits[k - 1];
	}
	__syncthreads();
	if (idx == 0xFFFFFFFFu)
		return;
	const uint32_t a = li << bbits, m = ((blk - a) << kBlkBits) | v;
	const uint32_t run = rscan[blk] - rscan[a] + s_runs[warp] + fr - (flag ? 1u : 0u);
	const uint32_t bit = wscan[blk] - wscan[a] + s_bits[warp] + fb - b;
	const uint32_t macro_cnt = cleaves[idx + 1], block_cnt = cleaves[idx + 2];
	const uint32_t macro_off = idx + 4u, block_off = macro_off + macro_cnt * 2u, weight_off = block_off + block_cnt * 2u;
	if (flag) {
		const uint32_t macro_first = a + ((m >> kMacroBits) << (kMacroBits - kBlkBits));
		cleaves[block_off + (run << 1)] = c;
		cleaves[block_off + (run << 1) + 1] = ((m & ((1u << kMacroBits) - 1u)) << 18) | (b << 16) | (bit - (wscan[macro_first] - wscan[a]));
	}
	if (b) {
		const uint32_t o = bit & 31u;
		atomicOr(&cleaves[weight_off + (bit >> 5)], w << o);
		if (o + b > 32u)
			atomicOr(&cleaves[weight_off + (bit >> 5) + 1], w >> (32u - o));
	}
}

__global__ void k_leaf_report(CItems leaf, uint32_t *parent_child, uint32_t *root_out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= leaf.count())
		return;
	const uint32_t par = leaf.parent[i];
	if (par == 0xFFFFFFFFu)
		*root_out = leaf.result[i];
	else
		parent_child[par] = leaf.result[i];
}

template <typename T> static cudaError_t cmalloc(T **p, uint64_t count, cudaStream_t s) {
	return cudaMallocAsync(reinterpret_cast<void **>(p), std::max<uint64_t>(count, 1) * sizeof(T), s);
}
static inline uint32_t cblocks(uint64_t threads) { return uint32_t((threads + kCB - 1) / kCB); }

struct CLevel {
	CItems v{};
	cudaStream_t s = nullptr;
	cudaError_t init(uint32_t cap, bool inner, cudaStream_t stream) {
		s = stream, v.cap = cap, v.n = 0;
		cudaError_t e;
		if ((e = cmalloc(&v.geom, cap, s)) || (e = cmalloc(&v.oct, cap, s)) || (e = cmalloc(&v.pos, cap, s)) ||
		    (e = cmalloc(&v.parent, cap, s)) || (e = cmalloc(&v.result, cap, s)))
			return e;
		if (inner && (e = cmalloc(&v.child, uint64_t(cap) * 8, s)))
			return e;
		return cudaSuccess;
	}
	bool owned = true;
	// the same arrays carved out of one arena (the one-sync octree pass allocates every level up front)
	static size_t arena_bytes(uint32_t cap, bool inner) {
		return ((size_t(cap) * (4 + 4 + 8 + 4 + 4 + (inner ? 32 : 0))) + 255) & ~size_t(255);
	}
	void init_from(char *&ptr, uint32_t cap, bool inner, cudaStream_t stream) {
		s = stream, v = CItems{}, v.cap = cap, owned = false;
		char *q = ptr;
		v.pos = reinterpret_cast<uint64_t *>(q), q += size_t(cap) * 8;
		v.geom = reinterpret_cast<uint32_t *>(q), q += size_t(cap) * 4;
		v.oct = reinterpret_cast<uint32_t *>(q), q += size_t(cap) * 4;
		v.parent = reinterpret_cast<uint32_t *>(q), q += size_t(cap) * 4;
		v.result = reinterpret_cast<uint32_t *>(q), q += size_t(cap) * 4;
		if (inner)
			v.child = reinterpret_cast<uint32_t *>(q);
		ptr += arena_bytes(cap, inner);
	}
	void release() {
		void *ptrs[] = {v.geom, v.oct, v.pos, v.parent, v.result, v.child};
		if (owned)
			for (void *q : ptrs)
				if (q)
					cudaFreeAsync(q, s);
		v = CItems{}, owned = true;
	}
};

hd_status ensure_color_storage(hd_pool *p, uint64_t node_words, uint64_t leaf_words) {
	cudaStream_t s = p->stream;
	if (!p->color_ctr) {
		HD_CUDA_TRY(cudaMalloc(&p->color_ctr, 4 * sizeof(uint32_t)));
		HD_CUDA_TRY(cudaMemset(p->color_ctr, 0, 4 * sizeof(uint32_t)));
	}
	if (!p->color_dirty_list) {
		HD_CUDA_TRY(cudaMalloc(&p->color_dirty_list, size_t(kColorDirtyCap) * sizeof(uint32_t)));
		HD_CUDA_TRY(cudaMalloc(&p->color_dirty_ctr, 2 * sizeof(uint32_t)));
		HD_CUDA_TRY(cudaMemset(p->color_dirty_ctr, 0, 2 * sizeof(uint32_t)));
	}
	auto grow = [&](uint32_t *&buf, uint64_t &cap, uint64_t used, uint64_t need) -> hd_status {
		if (need + 8 <= cap)
			return HD_OK;
		const uint64_t ncap = std::max<uint64_t>(need * 2 + (1u << 16), cap * 2);
		uint32_t *nb = nullptr;
		HD_CUDA_TRY(cudaMalloc(&nb, ncap * 4));
		HD_CUDA_TRY(cudaMemsetAsync(nb, 0, ncap * 4, s));
		if (p->color_stream) // kernels of a colour pass running beside a rebuild may still read (or have written) the old buffer
			HD_CUDA_TRY(cudaStreamSynchronize(p->color_stream));
		if (buf && used)
			HD_CUDA_TRY(cudaMemcpyAsync(nb, buf, used * 4, cudaMemcpyDeviceToDevice, s));
		HD_CUDA_TRY(cudaStreamSynchronize(s));
		cudaFree(buf);
		buf = nb, cap = ncap;
		return HD_OK;
	};
	hd_status st = grow(p->color_nodes, p->color_node_cap, p->color_node_words, node_words);
	if (st != HD_OK)
		return st;
	return grow(p->color_leaves, p->color_leaf_cap, p->color_leaf_words, leaf_words);
}

} // namespace hd

using namespace hd;

extern "C" {

hd_status hd_color_upload(hd_pool *p, const uint32_t *nodes, uint64_t node_words, const uint32_t *leaves, uint64_t leaf_words) {
	if (!p || (!nodes && node_words) || (!leaves && leaf_words) || node_words % 8)
		return HD_ERR_INVALID;
	HD_CUDA_TRY(cudaSetDevice(p->device));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	p->color_node_words = p->color_leaf_words = 0;
	hd_status st = ensure_color_storage(p, node_words, leaf_words);
	if (st != HD_OK)
		return st;
	if (node_words)
		HD_CUDA_TRY(cudaMemcpy(p->color_nodes, nodes, node_words * 4, cudaMemcpyHostToDevice));
	if (leaf_words)
		HD_CUDA_TRY(cudaMemcpy(p->color_leaves, leaves, leaf_words * 4, cudaMemcpyHostToDevice));
	p->color_node_words = node_words, p->color_leaf_words = leaf_words;
	const uint32_t ctr[4] = {uint32_t(node_words / 8), uint32_t(leaf_words), 0u, 0u};
	HD_CUDA_TRY(cudaMemcpy(p->color_ctr, ctr, sizeof(ctr), cudaMemcpyHostToDevice));
	return HD_OK;
}

hd_status hd_color_config(hd_pool *p, uint32_t leaf_level, uint32_t color_root) {
	if (!p || leaf_level + 2 > p->geo.node_levels)
		return HD_ERR_INVALID; // the colour leaf level must lie above the geometry leaf nodes
	p->color_leaf_level = leaf_level, p->color_root = color_root;
	p->color_dirty = true;
	return HD_OK;
}
uint32_t hd_color_root(const hd_pool *p) { return p ? p->color_root : HD_COLOR_NULL; }
uint32_t hd_color_leaf_level(const hd_pool *p) { return p ? p->color_leaf_level : 0; }

hd_status hd_color_sizes(hd_pool *p, uint64_t *node_words, uint64_t *leaf_words) {
	if (!p || !node_words || !leaf_words)
		return HD_ERR_INVALID;
	*node_words = p->color_node_words, *leaf_words = p->color_leaf_words;
	return HD_OK;
}
hd_status hd_color_read(hd_pool *p, uint32_t *nodes, uint64_t node_words, uint32_t *leaves, uint64_t leaf_words) {
	if (!p || node_words > p->color_node_words || leaf_words > p->color_leaf_words)
		return HD_ERR_INVALID;
	HD_CUDA_TRY(cudaSetDevice(p->device));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	if (node_words)
		HD_CUDA_TRY(cudaMemcpy(nodes, p->color_nodes, node_words * 4, cudaMemcpyDeviceToHost));
	if (leaf_words)
		HD_CUDA_TRY(cudaMemcpy(leaves, p->color_leaves, leaf_words * 4, cudaMemcpyDeviceToHost));
	return HD_OK;
}

hd_status hd_edit_color(hd_pool *p, uint32_t root_in, const hd_edit_desc *edit, uint32_t rgb8, uint32_t paint, uint32_t *root_out,
                        uint32_t *color_root_out, hd_edit_stats *stats) {
	if (!p || !edit || !root_out || !color_root_out)
		return HD_ERR_INVALID;
	*root_out = root_in, *color_root_out = p->color_root;
	const bool sphere = edit->kind == HD_EDIT_SPHERE_FILL, aabb = edit->kind == HD_EDIT_AABB_FILL;
	if (!(sphere || aabb) || (paint && !sphere) || p->color_leaf_level == 0 || p->color_leaf_level + 2 > p->geo.node_levels) {
		set_error("hd_edit_color: AABB fill, sphere fill or sphere paint with a configured colour pool (hd_color_config)");
		return HD_ERR_INVALID;
	}
	HD_CUDA_TRY(cudaSetDevice(p->device));
	cudaStream_t s = p->stream; // replaced by the colour stream below when the geometry edit runs beside this pass
	const Geometry &g = p->geo;
	ColorEdit e{};
	e.d = *edit, e.rgb8 = rgb8 & 0xFFFFFFu, e.paint = paint ? 1u : 0u;
	e.voxel_level = g.voxel_level(), e.node_levels = g.node_levels, e.leaf_level = p->color_leaf_level;
	const uint32_t LL = e.leaf_level, sbits = 3u * (e.voxel_level - LL);
	if (sbits > 30u) {
		set_error("colour leaf too large (%u voxel bits)", sbits);
		return HD_ERR_INVALID;
	}
	hd_status st = ensure_color_storage(p, p->color_node_words, p->color_leaf_words);
	if (st != HD_OK)
		return st;
	// ---- geometry first (SphereEditor<kPaint> leaves the voxels alone, main.cpp:133-136): the node pool is append-only, so
	// a failure here (bucket overflow, CUDA error) leaves both pools exactly as they were, and the colour pass below still
	// sees the OLD geometry under root_in.  Only the in-place rewrite of leaf chunks cannot be rolled back once it starts.
	// Round 2b: the one-launch rebuild is only ENQUEUED here (fast_edit_begin) and this pass runs beside it on a stream of
	// its own — it reads the old geometry, which the rebuild never touches — up to its first irreversible write, where
	// finish_geometry() waits for the rebuild and gives up (both pools untouched) if it failed.
	uint32_t new_root = root_in;
	bool geometry_in_flight = false, geometry_done = paint != 0u;
	if (!paint) {
		static const bool overlap = !(getenv("HD_COLOR_OVERLAP") && atoi(getenv("HD_COLOR_OVERLAP")) == 0);
		if (overlap) {
			st = edit_prepare(p);
			if (st == HD_OK)
				st = fast_edit_begin(p, root_in, edit, 1, &geometry_in_flight, true);
			if (st != HD_OK)
				return st;
		}
		if (geometry_in_flight) {
			if (!p->color_stream)
				HD_CUDA_TRY(cudaStreamCreateWithFlags(&p->color_stream, cudaStreamNonBlocking));
			s = p->color_stream;
		} else {
			st = hd_edit_batch(p, root_in, edit, 1, &new_root, stats);
			if (st != HD_OK)
				return st;
			geometry_done = true;
		}
	} else if (stats)
		memset(stats, 0, sizeof(*stats));
	auto finish_geometry = [&]() -> hd_status {
		if (geometry_done)
			return HD_OK;
		geometry_done = true;
		bool handled = false;
		hd_status gs = fast_edit_end(p, &new_root, stats, &handled);
		geometry_in_flight = false;
		if (gs == HD_OK && !handled) // a work queue was too small (nothing written): the general path redoes the edit
			gs = hd_edit_batch(p, root_in, edit, 1, &new_root, stats);
		return gs;
	};
	// whatever happens below, the rebuild must not be left in flight behind the caller's back
	ScopeExit wait_geometry{[&]() {
		if (geometry_in_flight)
			cudaStreamSynchronize(p->stream);
	}};

	// ---- colour pass over the OLD geometry (node memory is immutable) ----
	std::vector<CLevel> inner(LL + 1);
	CLevel leaf;
	uint32_t *counts = nullptr, *root_dev = nullptr; // counts: [inner created, leaves created, -, overflow]
	HD_CUDA_TRY(cmalloc(&counts, 4, s));
	HD_CUDA_TRY(cmalloc(&root_dev, 1, s));
	bool cleaned = false;
	auto cleanup = [&]() {
		if (cleaned)
			return;
		cleaned = true;
		for (auto &l : inner)
			l.release();
		leaf.release();
		cudaFreeAsync(counts, s), cudaFreeAsync(root_dev, s);
		cudaStreamSynchronize(s);
	};
	ScopeExit guard{cleanup}; // the HD_CUDA_TRY early returns below release the stream-ordered scratch as well
	uint32_t hc[4] = {0, 0, 0, 0};
	auto read_counts = [&]() -> hd_status {
		HD_CUDA_TRY(cudaMemcpyAsync(hc, counts, sizeof(hc), cudaMemcpyDeviceToHost, s));
		HD_CUDA_TRY(cudaStreamSynchronize(s));
		return HD_OK;
	};
	const uint32_t col_root_in = p->color_root;
	HD_CUDA_TRY(cudaMemcpyAsync(root_dev, &col_root_in, 4, cudaMemcpyHostToDevice, s));
	uint32_t deepest = 0;
	// ---- octree levels, optimistic: ONE host round trip.  A brush touches a handful of octree nodes per level, so every
	// level gets a fixed-capacity queue up front, the kernels read their item counts from device memory, and all the counts
	// come back together.  A queue that turns out too small (a huge edit) is detected there; nothing has been written to the
	// colour pool yet, so the pass is simply redone level by level with exact sizes (the code below it).
	char *arena = nullptr;
	uint32_t *lvl_counts = nullptr;
	ScopeExit free_arena{[&]() {
		if (arena)
			cudaFreeAsync(arena, s);
		if (lvl_counts)
			cudaFreeAsync(lvl_counts, s);
	}};
	bool sized = false;
	static const bool one_sync = !(getenv("HD_COLOR_ONE_SYNC") && atoi(getenv("HD_COLOR_ONE_SYNC")) == 0);
	if (one_sync) {
		constexpr uint32_t kInnerCap = 2048, kLeafCap = 8192;
		auto cap_of = [&](uint32_t l) { return l >= 4 ? kInnerCap : std::min<uint32_t>(kInnerCap, 1u << (3 * l)); };
		size_t bytes = CLevel::arena_bytes(kLeafCap, false);
		for (uint32_t l = 0; l <= LL; ++l)
			bytes += CLevel::arena_bytes(cap_of(l), true);
		HD_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void **>(&arena), bytes, s));
		HD_CUDA_TRY(cmalloc(&lvl_counts, 4 * (LL + 2), s));
		HD_CUDA_TRY(cudaMemsetAsync(lvl_counts, 0, size_t(4) * (LL + 2) * 4, s));
		char *q = arena;
		for (uint32_t l = 0; l <= LL; ++l) {
			inner[l].init_from(q, cap_of(l), true, s);
			inner[l].v.n_dev = lvl_counts + 4 * l; // block l: what produced level l ([0] inner items, [1] leaves, [3] overflow)
		}
		leaf.init_from(q, kLeafCap, false, s);
		leaf.v.n_dev = lvl_counts + 4 * LL + 1;
		{
			CLevels all{};
			for (uint32_t l = 0; l <= LL; ++l)
				all.lv[l] = inner[l].v;
			k_cdown_all<<<1, kOctThreads, 0, s>>>(e, root_in, col_root_in, p->words, p->color_nodes, all, leaf.v, lvl_counts, root_dev);
			HD_LAUNCH_CHECK();
		}
		std::vector<uint32_t> hcs(4 * (LL + 2));
		HD_CUDA_TRY(cudaMemcpyAsync(hcs.data(), lvl_counts, hcs.size() * 4, cudaMemcpyDeviceToHost, s));
		HD_CUDA_TRY(cudaStreamSynchronize(s));
		sized = true;
		for (uint32_t l = 0; l <= LL && sized; ++l)
			sized = hcs[4 * l + 3] == 0 && hcs[4 * l] <= cap_of(l);
		const uint32_t n_leaves = LL == 0 ? hcs[1] : hcs[4 * LL + 1];
		sized = sized && n_leaves <= kLeafCap;
		if (sized) {
			for (uint32_t l = 0; l <= LL; ++l) {
				inner[l].v.n = hcs[4 * l], inner[l].v.n_dev = nullptr;
				if (inner[l].v.n)
					deepest = l;
			}
			leaf.v.n = n_leaves, leaf.v.n_dev = nullptr;
		} else {
			for (auto &l : inner)
				l.release();
			leaf.release();
		}
	}
	if (!sized) {
	// ---- octree levels, exact: one round trip per level ----
	// leaf items can be created from any octree level's expansion: size the leaf list for the worst case lazily
	// (grown by re-running is avoided: capacity = 8 x the widest inner level, which bounds the leaves of one level;
	// leaves only appear at level LL, i.e. from the expansion of level LL-1, or the root when LL == 0)
	HD_CUDA_TRY(inner[0].init(1, true, s));
	HD_CUDA_TRY(leaf.init(1, false, s));
	HD_CUDA_TRY(cudaMemsetAsync(counts, 0, 16, s));
	k_croot<<<1, 32, 0, s>>>(e, root_in, col_root_in, inner[0].v, leaf.v, counts, root_dev);
	HD_LAUNCH_CHECK();
	st = read_counts();
	if (st != HD_OK) {
		cleanup();
		return st;
	}
	inner[0].v.n = hc[0], leaf.v.n = hc[1];
	deepest = 0;
	for (uint32_t l = 0; l < LL && inner[l].v.n; ++l) {
		const uint64_t cap = uint64_t(inner[l].v.n) * 8;
		if (cap > 0x7FFFFFF0ull) {
			cleanup();
			return HD_ERR_OVERFLOW;
		}
		HD_CUDA_TRY(inner[l + 1].init(uint32_t(cap), true, s));
		if (l + 1 == LL) {
			leaf.release();
			HD_CUDA_TRY(leaf.init(uint32_t(cap), false, s));
		}
		HD_CUDA_TRY(cudaMemsetAsync(counts, 0, 8, s));
		k_cdown<<<cblocks(cap), kCB, 0, s>>>(e, l, p->words, p->color_nodes, inner[l].v, inner[l + 1].v, leaf.v, counts);
		HD_LAUNCH_CHECK();
		st = read_counts();
		if (st != HD_OK || hc[3]) {
			cleanup();
			return st != HD_OK ? st : HD_ERR_OVERFLOW;
		}
		inner[l + 1].v.n = hc[0];
		if (l + 1 == LL)
			leaf.v.n = hc[1];
		deepest = l + 1;
	}
	}

	const uint32_t n_leaf = leaf.v.n;
	uint64_t leaf_voxels = 0;
	static const bool block_path = !(getenv("HD_COLOR_BLOCKS") && atoi(getenv("HD_COLOR_BLOCKS")) == 0);
	if (n_leaf && block_path && sbits > kBlkBits) {
		// ---- block path: classify 8^3 blocks, expand only the mixed ones to voxels (see k_cblock) ----
		const uint32_t bbits = sbits - kBlkBits;
		const uint32_t per_batch = std::max<uint32_t>(1u, (1u << 24) >> bbits);
		for (uint32_t first = 0; first < n_leaf; first += per_batch) {
			const uint32_t nl = std::min(per_batch, n_leaf - first), nb = nl << bbits;
			BlockArrays B{};
			uint32_t *list = nullptr, *list_count = nullptr, *rscan = nullptr, *wscan = nullptr, *chunk_idx = nullptr, *colors = nullptr;
			uint8_t *bw = nullptr;
			unsigned long long *total = nullptr;
			// one stream-ordered allocation for the per-block arrays (sixteen of them cost more host time than the kernels run)
			char *block_arena = nullptr;
			ScopeExit free_batch{[&]() {
				if (block_arena)
					cudaFreeAsync(block_arena, s);
				if (colors)
					cudaFreeAsync(colors, s);
				if (bw)
					cudaFreeAsync(bw, s);
			}};
			{
				const size_t w4 = (size_t(nb) * 4 + 255) & ~size_t(255), w1 = (size_t(nb) + 255) & ~size_t(255);
				const size_t bytes = 10 * w4 + 3 * w1 + ((size_t(nl) * 4 + 255) & ~size_t(255)) + 512;
				HD_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void **>(&block_arena), bytes, s));
				char *q = block_arena;
				auto take4 = [&]() { uint32_t *r = reinterpret_cast<uint32_t *>(q); q += w4; return r; };
				auto take1 = [&]() { uint8_t *r = reinterpret_cast<uint8_t *>(q); q += w1; return r; };
				B.key = take4(), B.first_c = take4(), B.last_c = take4(), B.runs = take4(), B.bits = take4(), B.slot = take4();
				list = take4(), rscan = take4(), wscan = take4();
				B.first_b = take1(), B.last_b = take1(), B.start = take1();
				chunk_idx = reinterpret_cast<uint32_t *>(q), q += (size_t(nl) * 4 + 255) & ~size_t(255);
				total = reinterpret_cast<unsigned long long *>(q), q += 256;
				list_count = reinterpret_cast<uint32_t *>(q);
				HD_CUDA_TRY(cudaMemsetAsync(total, 0, 512, s)); // total and list_count
			}
			k_cblock<<<cblocks(nb), kCB, 0, s>>>(e, p->words, p->color_leaves, leaf.v, first, nl, sbits, B, list, list_count);
			HD_LAUNCH_CHECK();
			uint32_t n_listed = 0;
			HD_CUDA_TRY(cudaMemcpyAsync(&n_listed, list_count, 4, cudaMemcpyDeviceToHost, s));
			HD_CUDA_TRY(cudaStreamSynchronize(s));
			const uint64_t nv = uint64_t(n_listed) << kBlkBits;
			leaf_voxels += nv;
			if (n_listed) {
				HD_CUDA_TRY(cmalloc(&colors, nv, s));
				HD_CUDA_TRY(cmalloc(&bw, nv, s));
				k_voxels<<<cblocks(nv), kCB, 0, s>>>(e, p->words, p->color_leaves, leaf.v, first, nl, sbits, colors, bw, list, n_listed);
				HD_LAUNCH_CHECK();
				k_block_summary<<<cblocks(uint64_t(n_listed) * 32u), kCB, 0, s>>>(n_listed, list, colors, bw, B);
				HD_LAUNCH_CHECK();
			}
			k_block_flags<<<cblocks(nb), kCB, 0, s>>>(nb, sbits, B);
			HD_LAUNCH_CHECK();
			if (nb <= kSmallScan) {
				k_scan2_small<<<1, 1024, 0, s>>>(B.runs, rscan, B.bits, wscan, nb);
				HD_LAUNCH_CHECK();
			} else {
				st = exclusive_scan(p, B.runs, rscan, nb, s);
				if (st == HD_OK)
					st = exclusive_scan(p, B.bits, wscan, nb, s);
				if (st != HD_OK)
					return st;
			}
			unsigned long long need = 0;
			k_leaf_size<<<cblocks(nl), kCB, 0, s>>>(leaf.v, first, nl, sbits, bbits, B.runs, rscan, B.bits, wscan, p->color_leaves, total);
			HD_LAUNCH_CHECK();
			HD_CUDA_TRY(cudaMemcpyAsync(&need, total, 8, cudaMemcpyDeviceToHost, s));
			HD_CUDA_TRY(cudaStreamSynchronize(s));
			st = finish_geometry(); // everything up to here only filled scratch arrays; chunks are rewritten in place from here on
			if (st == HD_OK)
				st = ensure_color_storage(p, p->color_node_words, p->color_leaf_words + need);
			if (st != HD_OK)
				return st;
			k_leaf_alloc<<<cblocks(nl), kCB, 0, s>>>(leaf.v, first, nl, sbits, bbits, B.runs, rscan, B.bits, wscan, p->color_leaves, p->color_ctr,
			                                        p->color_leaf_cap, chunk_idx);
			HD_LAUNCH_CHECK();
			k_color_mark_dirty<<<cblocks(nl), kCB, 0, s>>>(chunk_idx, nl, uint32_t(std::min<uint64_t>(p->color_synced_leaf_words, 0xFFFFFFFFull)),
			                                              p->color_dirty_list, p->color_dirty_ctr, kColorDirtyCap);
			HD_LAUNCH_CHECK();
			k_zero_weights<<<std::min<uint32_t>(nl, 1184u), kCB, 0, s>>>(nl, chunk_idx, p->color_leaves);
			HD_LAUNCH_CHECK();
			k_emit_blocks<<<cblocks(nb), kCB, 0, s>>>(nb, sbits, B, rscan, wscan, chunk_idx, p->color_leaves);
			HD_LAUNCH_CHECK();
			if (n_listed) {
				k_emit_listed<<<n_listed, 1u << kBlkBits, 0, s>>>(sbits, list, colors, bw, B, rscan, wscan, chunk_idx, p->color_leaves);
				HD_LAUNCH_CHECK();
			}
			if (first + per_batch < n_leaf) { // another batch follows: its storage check needs the words used so far
				uint32_t ctr[4];
				HD_CUDA_TRY(cudaMemcpyAsync(ctr, p->color_ctr, sizeof(ctr), cudaMemcpyDeviceToHost, s));
				HD_CUDA_TRY(cudaStreamSynchronize(s));
				p->color_leaf_words = ctr[1];
			}
		}
		k_leaf_report<<<cblocks(n_leaf), kCB, 0, s>>>(leaf.v, LL ? inner[LL - 1].v.child : nullptr, root_dev);
		HD_LAUNCH_CHECK();
	} else if (n_leaf) {
		const uint32_t per_batch = std::max<uint32_t>(1u, uint32_t((1ull << 26) >> sbits));
		for (uint32_t first = 0; first < n_leaf; first += per_batch) {
			const uint32_t nl = std::min(per_batch, n_leaf - first);
			const uint64_t n = uint64_t(nl) << sbits;
			leaf_voxels += n;
			uint32_t *colors = nullptr, *flag = nullptr, *bits = nullptr, *fscan = nullptr, *bscan = nullptr, *chunk_idx = nullptr;
			uint8_t *bw = nullptr;
			HD_CUDA_TRY(cmalloc(&colors, n, s));
			HD_CUDA_TRY(cmalloc(&bw, n, s));
			HD_CUDA_TRY(cmalloc(&flag, n, s));
			HD_CUDA_TRY(cmalloc(&bits, n, s));
			HD_CUDA_TRY(cmalloc(&fscan, n, s));
			HD_CUDA_TRY(cmalloc(&bscan, n, s));
			HD_CUDA_TRY(cmalloc(&chunk_idx, nl, s));
			k_voxels<<<cblocks(n), kCB, 0, s>>>(e, p->words, p->color_leaves, leaf.v, first, nl, sbits, colors, bw, nullptr, 0u);
			HD_LAUNCH_CHECK();
			k_flags<<<cblocks(n), kCB, 0, s>>>(colors, bw, n, sbits, flag, bits);
			HD_LAUNCH_CHECK();
			st = exclusive_scan(p, flag, fscan, n, s);
			if (st == HD_OK)
				st = exclusive_scan(p, bits, bscan, n, s);
			if (st == HD_OK) { // how many words will this batch append?  grow the leaf array once, before allocating
				unsigned long long *total = nullptr, need = 0;
				HD_CUDA_TRY(cmalloc(&total, 1, s));
				HD_CUDA_TRY(cudaMemsetAsync(total, 0, 8, s));
				k_leaf_size<<<cblocks(nl), kCB, 0, s>>>(leaf.v, first, nl, sbits, sbits, flag, fscan, bits, bscan, p->color_leaves, total);
				HD_LAUNCH_CHECK();
				HD_CUDA_TRY(cudaMemcpyAsync(&need, total, 8, cudaMemcpyDeviceToHost, s));
				HD_CUDA_TRY(cudaStreamSynchronize(s));
				cudaFreeAsync(total, s);
				st = finish_geometry();
				if (st == HD_OK)
					st = ensure_color_storage(p, p->color_node_words, p->color_leaf_words + need);
			}
			if (st == HD_OK) {
				k_leaf_alloc<<<cblocks(nl), kCB, 0, s>>>(leaf.v, first, nl, sbits, sbits, flag, fscan, bits, bscan, p->color_leaves, p->color_ctr,
				                                        p->color_leaf_cap, chunk_idx);
				HD_LAUNCH_CHECK();
				k_color_mark_dirty<<<cblocks(nl), kCB, 0, s>>>(chunk_idx, nl, uint32_t(std::min<uint64_t>(p->color_synced_leaf_words, 0xFFFFFFFFull)),
				                                              p->color_dirty_list, p->color_dirty_ctr, kColorDirtyCap);
				HD_LAUNCH_CHECK();
				k_zero_weights<<<std::min<uint32_t>(nl, 1184u), kCB, 0, s>>>(nl, chunk_idx, p->color_leaves);
				HD_LAUNCH_CHECK();
				k_emit<<<cblocks(n), kCB, 0, s>>>(n, sbits, colors, bw, flag, fscan, bscan, chunk_idx, p->color_leaves);
				HD_LAUNCH_CHECK();
				uint32_t ctr[4];
				HD_CUDA_TRY(cudaMemcpyAsync(ctr, p->color_ctr, sizeof(ctr), cudaMemcpyDeviceToHost, s));
				HD_CUDA_TRY(cudaStreamSynchronize(s));
				p->color_leaf_words = ctr[1];
			}
			cudaFreeAsync(colors, s), cudaFreeAsync(bw, s), cudaFreeAsync(flag, s), cudaFreeAsync(bits, s);
			cudaFreeAsync(fscan, s), cudaFreeAsync(bscan, s), cudaFreeAsync(chunk_idx, s);
			if (st != HD_OK) {
				cleanup();
				return st;
			}
		}
		k_leaf_report<<<cblocks(n_leaf), kCB, 0, s>>>(leaf.v, LL ? inner[LL - 1].v.child : nullptr, root_dev);
		HD_LAUNCH_CHECK();
	}

	// ---- octree nodes bottom-up (SetNode) ----
	st = finish_geometry(); // edits that touch no colour leaf get here with the rebuild still in flight
	if (st != HD_OK) {
		cleanup();
		return st;
	}
	{
		uint64_t new_nodes = 0;
		for (uint32_t l = 0; l <= deepest && l < LL; ++l)
			new_nodes += inner[l].v.n;
		st = ensure_color_storage(p, p->color_node_words + new_nodes * 8, p->color_leaf_words);
		if (st != HD_OK) {
			cleanup();
			return st;
		}
	}
	if (sized) { // brush-sized edit: all levels in one launch
		CLevels all{};
		const uint32_t top = std::min(std::min(deepest, LL ? LL - 1 : 0) + 1, LL);
		for (uint32_t l = 0; l < top; ++l)
			all.lv[l] = inner[l].v;
		if (top) {
			k_cup_all<<<1, kOctThreads, 0, s>>>(all, top, p->color_nodes, p->color_ctr, p->color_node_cap, root_dev);
			HD_LAUNCH_CHECK();
		}
	} else
		for (uint32_t l = std::min(deepest, LL ? LL - 1 : 0) + 1; l-- > 0;) {
			if (l >= LL || inner[l].v.n == 0)
				continue;
			k_cup<<<cblocks(inner[l].v.n), kCB, 0, s>>>(inner[l].v, p->color_nodes, p->color_ctr, p->color_node_cap,
			                                          l ? inner[l - 1].v.child : nullptr, root_dev);
			HD_LAUNCH_CHECK();
		}
	uint32_t ctr[4], new_color_root;
	HD_CUDA_TRY(cudaMemcpyAsync(ctr, p->color_ctr, sizeof(ctr), cudaMemcpyDeviceToHost, s));
	HD_CUDA_TRY(cudaMemcpyAsync(&new_color_root, root_dev, 4, cudaMemcpyDeviceToHost, s));
	HD_CUDA_TRY(cudaStreamSynchronize(s));
	cleanup();
	p->color_node_words = uint64_t(ctr[0]) * 8, p->color_leaf_words = ctr[1];
	if (ctr[2]) {
		set_error("colour pool out of space");
		return HD_ERR_OOM;
	}
	p->color_root = new_color_root;
	p->color_dirty = true;
	*color_root_out = new_color_root;

	*root_out = new_root;
	if (stats)
		stats->in_range_voxels = leaf_voxels; // colour voxels re-encoded (diagnostic)
	return HD_OK;
}

} // extern "C"
