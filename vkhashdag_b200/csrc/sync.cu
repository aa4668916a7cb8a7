// sync.cu — dirty-range tracking and replica synchronisation (multi-GPU: pool replicated).
//
// Replaces DAGNodePool::Flush (src/DAGNodePool.cpp:56-85) and its per-page m_page_write_ranges bookkeeping
// (src/DAGNodePool.hpp:44-46,62-69).  Buckets are append-only between GCs, so "what changed since the last
// sync" is exactly [bucket_synced[b], bucket_words[b]) for every bucket: no write log is needed.  The editing
// rank packs those runs into ONE staging buffer, the caller broadcasts it (one NCCL broadcast over NVLink),
// replicas scatter it and publish the root last.
//
// Staging layout (u32 words), 8-word header:
//   [0] n_ranges  [1] payload_words  [2] root  [3] flags (bit 0: clear the replica first, bit 1: a colour section follows,
//   bit 2: the colour section is the WHOLE colour pool)  [4] n_color_ranges  [5] color_payload_words  [6] color_root
//   [7] color_leaf_level
// then n_ranges x {word_offset, word_count, payload_offset}, the node payload, and — when flags bit 1 is set — the colour
// section: [color_node_words][color_leaf_words] (totals after the edit), n_color_ranges x {tagged_offset, word_count,
// payload_offset} (tagged_offset bit 31: 1 = leaf array word index, 0 = node array NODE index), the colour payload.
// Colour nodes and appended leaf chunks are append-only; chunks rewritten in place (DAGColorPool::SetLeaf with
// keep_history = false, src/DAGColorPool.hpp:173-204) come from the dirty list color.cu keeps on the device.
#include "common.cuh"

#include <algorithm>
#include <cstring>

namespace hd {

constexpr uint32_t kHeaderWords = 8;
constexpr uint32_t kFlagClear = 1u, kFlagColor = 2u, kFlagColorFull = 4u;
constexpr uint32_t kBigRange = 1u << 15; // colour ranges above this many words are copied by the whole grid

__global__ void k_dirty_scan(const uint32_t *__restrict__ cur, const uint32_t *__restrict__ synced, uint32_t n_buckets,
                             uint32_t bucket_shift, uint32_t *header, uint32_t *ranges, uint32_t capacity) {
	const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= n_buckets)
		return;
	const uint32_t c = cur[b], s = synced[b];
	if (c <= s)
		return;
	const uint32_t idx = atomicAdd(&header[0], 1u);
	const uint32_t poff = atomicAdd(&header[1], c - s);
	if (ranges && idx < capacity) {
		ranges[idx * 3 + 0] = (b << bucket_shift) + s;
		ranges[idx * 3 + 1] = c - s;
		ranges[idx * 3 + 2] = poff;
	}
}

// one CTA per range at a time (grid-stride).  n_ranges comes from the device header; the triples start right behind
// the header and the payload right behind the triples, so no host round trip separates the scan from the gather.
// A staging buffer that cannot hold everything is left with its header only (the host sees the sizes and reports it).
__global__ void k_dirty_gather(const uint32_t *__restrict__ words, uint32_t *stg, uint32_t cap_words) {
	const uint32_t n = stg[0];
	if (uint64_t(kHeaderWords) + uint64_t(n) * 3u + stg[1] > cap_words)
		return;
	const uint32_t *ranges = stg + kHeaderWords;
	uint32_t *payload = stg + kHeaderWords + size_t(n) * 3u;
	for (uint32_t r = blockIdx.x; r < n; r += gridDim.x) {
		const uint32_t off = ranges[r * 3], cnt = ranges[r * 3 + 1], poff = ranges[r * 3 + 2];
		for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x)
			payload[poff + i] = words[off + i];
	}
}

// Applies (or, with validate_only, just checks) the range triples.  Every triple is bounds-checked before anything is
// written: it must stay inside one bucket of the pool and inside the payload; a bad triple raises *bad and is skipped, so a
// truncated or corrupt blob can never write outside the pool or push bucket_words past a bucket's capacity.
__global__ void k_dirty_scatter(uint32_t *words, uint32_t *bucket_words, uint32_t *bucket_synced, uint32_t bucket_shift,
                                uint64_t total_words, const uint32_t *__restrict__ header, const uint32_t *__restrict__ ranges,
                                const uint32_t *__restrict__ payload, bool validate_only, uint32_t *bad) {
	const uint32_t n = header[0], payload_words = header[1];
	for (uint32_t r = blockIdx.x; r < n; r += gridDim.x) {
		const uint32_t off = ranges[r * 3], cnt = ranges[r * 3 + 1], poff = ranges[r * 3 + 2];
		const uint32_t in_bucket = off & ((1u << bucket_shift) - 1u);
		const bool ok = cnt != 0u && uint64_t(off) + cnt <= total_words && uint64_t(in_bucket) + cnt <= (1ull << bucket_shift) &&
		                uint64_t(poff) + cnt <= payload_words;
		if (!ok) {
			if (threadIdx.x == 0)
				atomicOr(bad, 1u);
			continue;
		}
		if (validate_only)
			continue;
		for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x)
			words[off + i] = payload[poff + i];
		if (threadIdx.x == 0) {
			const uint32_t bucket = off >> bucket_shift;
			bucket_words[bucket] = in_bucket + cnt;
			bucket_synced[bucket] = in_bucket + cnt;
		}
	}
}

// ---- colour section -------------------------------------------------------------------------------------------------
// triples of the in-place rewritten chunks: {leaf-tagged word index, allocated words (chunk word 0), payload offset}
__global__ void k_color_chunk_ranges(const uint32_t *__restrict__ list, uint32_t n, const uint32_t *__restrict__ cleaves,
                                     uint64_t leaf_words, uint32_t *triples, uint32_t *payload_ctr) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const uint32_t idx = list[i];
	uint32_t cnt = idx < leaf_words ? cleaves[idx] : 0u;
	if (uint64_t(idx) + cnt > leaf_words)
		cnt = 0u;
	triples[i * 3] = 0x80000000u | idx, triples[i * 3 + 1] = cnt, triples[i * 3 + 2] = atomicAdd(payload_ctr, cnt);
}

// copy between the colour buffers and the payload for every triple; pack = buffers -> payload, otherwise payload -> buffers
// (bounds-checked against the buffer capacities).  Small ranges: one CTA each; big ranges: the whole grid.
__global__ void k_color_copy(uint32_t *cnodes, uint32_t *cleaves, uint64_t node_cap, uint64_t leaf_cap, const uint32_t *__restrict__ triples,
                             uint32_t n, uint32_t *payload, uint32_t payload_words, bool pack, uint32_t *bad) {
	for (uint32_t r = 0; r < n; ++r) {
		const uint32_t tagged = triples[r * 3], cnt = triples[r * 3 + 1], poff = triples[r * 3 + 2];
		const bool leaf = tagged >> 31;
		const uint64_t off = leaf ? uint64_t(tagged & 0x7FFFFFFFu) : uint64_t(tagged) << 3;
		const bool big = cnt > kBigRange;
		if (!big && r % gridDim.x != blockIdx.x)
			continue;
		if (off + cnt > (leaf ? leaf_cap : node_cap) || uint64_t(poff) + cnt > payload_words) {
			if (threadIdx.x == 0)
				atomicOr(bad, 1u);
			continue;
		}
		uint32_t *buf = (leaf ? cleaves : cnodes) + off;
		const uint32_t first = big ? blockIdx.x * blockDim.x + threadIdx.x : threadIdx.x;
		const uint32_t step = big ? gridDim.x * blockDim.x : blockDim.x;
		if (pack)
			for (uint32_t i = first; i < cnt; i += step)
				payload[poff + i] = buf[i];
		else
			for (uint32_t i = first; i < cnt; i += step)
				buf[i] = payload[poff + i];
	}
}

static hd_status ensure_dirty_scratch(hd_pool *p, uint64_t bytes) {
	if (p->dirty_scratch_bytes >= bytes)
		return HD_OK;
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	cudaFree(p->dirty_scratch);
	p->dirty_scratch = nullptr, p->dirty_scratch_bytes = 0;
	HD_CUDA_TRY(cudaMalloc(&p->dirty_scratch, bytes));
	p->dirty_scratch_bytes = bytes;
	return HD_OK;
}

static hd_status scan_counts(hd_pool *p, uint32_t host_header[kHeaderWords]) {
	hd_status s = ensure_dirty_scratch(p, kHeaderWords * 4);
	if (s != HD_OK)
		return s;
	HD_CUDA_TRY(cudaMemsetAsync(p->dirty_scratch, 0, kHeaderWords * 4, p->stream));
	const uint32_t nb = p->geo.total_buckets;
	k_dirty_scan<<<(nb + 255) / 256, 256, 0, p->stream>>>(p->bucket_words, p->bucket_synced, nb, p->geo.bucket_shift(),
	                                                      p->dirty_scratch, nullptr, 0);
	HD_LAUNCH_CHECK();
	HD_CUDA_TRY(cudaMemcpyAsync(host_header, p->dirty_scratch, kHeaderWords * 4, cudaMemcpyDeviceToHost, p->stream));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	return HD_OK;
}

} // namespace hd

using namespace hd;

extern "C" {

hd_status hd_dirty_count(hd_pool *p, uint32_t *n_ranges, uint64_t *packed_bytes) {
	if (!p)
		return HD_ERR_INVALID;
	HD_CUDA_TRY(cudaSetDevice(p->device));
	uint32_t h[kHeaderWords];
	hd_status s = scan_counts(p, h);
	if (s != HD_OK)
		return s;
	if (n_ranges)
		*n_ranges = h[0];
	if (packed_bytes)
		*packed_bytes = (uint64_t(kHeaderWords) + uint64_t(h[0]) * 3 + h[1]) * 4;
	return HD_OK;
}

hd_status hd_dirty_ranges(hd_pool *p, hd_dirty_range *out, uint32_t capacity, uint32_t *n_out) {
	if (!p || !n_out || (!out && capacity))
		return HD_ERR_INVALID;
	HD_CUDA_TRY(cudaSetDevice(p->device));
	uint32_t h[kHeaderWords];
	hd_status s = scan_counts(p, h);
	if (s != HD_OK)
		return s;
	*n_out = h[0];
	if (h[0] == 0 || capacity == 0)
		return HD_OK;
	s = ensure_dirty_scratch(p, (kHeaderWords + uint64_t(h[0]) * 3) * 4);
	if (s != HD_OK)
		return s;
	HD_CUDA_TRY(cudaMemsetAsync(p->dirty_scratch, 0, kHeaderWords * 4, p->stream));
	const uint32_t nb = p->geo.total_buckets;
	k_dirty_scan<<<(nb + 255) / 256, 256, 0, p->stream>>>(p->bucket_words, p->bucket_synced, nb, p->geo.bucket_shift(),
	                                                      p->dirty_scratch, p->dirty_scratch + kHeaderWords, h[0]);
	HD_LAUNCH_CHECK();
	std::vector<uint32_t> tri(size_t(h[0]) * 3);
	HD_CUDA_TRY(cudaMemcpyAsync(tri.data(), p->dirty_scratch + kHeaderWords, tri.size() * 4, cudaMemcpyDeviceToHost,
	                            p->stream));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	for (uint32_t i = 0; i < h[0] && i < capacity; ++i)
		out[i] = hd_dirty_range{tri[i * 3], tri[i * 3 + 1]};
	return HD_OK;
}

// Appends the colour section behind the node payload (at word `base`).  Runs only when the colour pool changed since the
// last hd_dirty_reset, and then costs two extra host round trips (the dirty-list count and the chunk payload size).
static hd_status color_pack(hd_pool *p, uint32_t *stg, uint64_t cap_words, uint64_t base, uint32_t *flags, uint32_t *n_ranges_out,
                            uint32_t *payload_out, uint64_t *section_words) {
	cudaStream_t st = p->stream;
	uint32_t dc[2] = {0, 0};
	if (p->color_dirty_ctr) {
		HD_CUDA_TRY(cudaMemcpyAsync(dc, p->color_dirty_ctr, sizeof(dc), cudaMemcpyDeviceToHost, st));
		HD_CUDA_TRY(cudaStreamSynchronize(st));
	}
	const bool full = p->color_full_resync || dc[1] != 0u;
	const uint64_t node0 = full ? 0 : p->color_synced_node_words, leaf0 = full ? 0 : p->color_synced_leaf_words;
	const uint64_t node_cnt = p->color_node_words - node0, leaf_cnt = p->color_leaf_words - leaf0;
	if (node_cnt > 0xFFFFFFFFull || leaf_cnt > 0xFFFFFFFFull) {
		set_error("colour sync: range too large for one blob");
		return HD_ERR_OVERFLOW;
	}
	const uint32_t n_chunks = full ? 0u : std::min(dc[0], kColorDirtyCap);
	std::vector<uint32_t> head; // the two totals + the (at most two) append triples, built on the host
	head.push_back(uint32_t(p->color_node_words / 8)), head.push_back(uint32_t(p->color_leaf_words));
	uint32_t poff = 0, n_app = 0;
	if (node_cnt)
		head.insert(head.end(), {uint32_t(node0 / 8), uint32_t(node_cnt), poff}), poff += uint32_t(node_cnt), ++n_app;
	if (leaf_cnt)
		head.insert(head.end(), {0x80000000u | uint32_t(leaf0), uint32_t(leaf_cnt), poff}), poff += uint32_t(leaf_cnt), ++n_app;
	uint32_t chunk_payload = 0, *tri = nullptr, *ctr = nullptr;
	if (n_chunks) { // triples of the rewritten chunks into scratch first: their payload size decides whether the blob fits
		hd_status s = ensure_dirty_scratch(p, (uint64_t(n_chunks) * 3 + 4) * 4);
		if (s != HD_OK)
			return s;
		ctr = p->dirty_scratch, tri = p->dirty_scratch + 4;
		HD_CUDA_TRY(cudaMemcpyAsync(ctr, &poff, 4, cudaMemcpyHostToDevice, st));
		k_color_chunk_ranges<<<(n_chunks + 255) / 256, 256, 0, st>>>(p->color_dirty_list, n_chunks, p->color_leaves, p->color_leaf_words,
		                                                            tri, ctr);
		HD_LAUNCH_CHECK();
		uint32_t end = 0;
		HD_CUDA_TRY(cudaMemcpyAsync(&end, ctr, 4, cudaMemcpyDeviceToHost, st));
		HD_CUDA_TRY(cudaStreamSynchronize(st));
		chunk_payload = end - poff;
	}
	const uint32_t n_ranges = n_app + n_chunks;
	const uint64_t payload = uint64_t(poff) + chunk_payload;
	*section_words = 2 + uint64_t(n_ranges) * 3 + payload;
	*n_ranges_out = n_ranges, *payload_out = uint32_t(payload);
	*flags |= kFlagColor | (full ? kFlagColorFull : 0u);
	if (payload > 0xFFFFFFFFull || base + *section_words > cap_words)
		return HD_OK; // the caller reports the size that is needed
	uint32_t *sec = stg + base, *triples = sec + 2, *pay = triples + size_t(n_ranges) * 3;
	HD_CUDA_TRY(cudaMemcpyAsync(sec, head.data(), head.size() * 4, cudaMemcpyHostToDevice, st));
	HD_CUDA_TRY(cudaStreamSynchronize(st)); // head is a stack vector
	if (n_chunks)
		HD_CUDA_TRY(cudaMemcpyAsync(triples + size_t(n_app) * 3, tri, size_t(n_chunks) * 12, cudaMemcpyDeviceToDevice, st));
	uint32_t off = 0;
	if (node_cnt) {
		HD_CUDA_TRY(cudaMemcpyAsync(pay, p->color_nodes + node0, node_cnt * 4, cudaMemcpyDeviceToDevice, st));
		off += uint32_t(node_cnt);
	}
	if (leaf_cnt)
		HD_CUDA_TRY(cudaMemcpyAsync(pay + off, p->color_leaves + leaf0, leaf_cnt * 4, cudaMemcpyDeviceToDevice, st));
	if (n_chunks) {
		hd_status s = ensure_dirty_scratch(p, (uint64_t(n_chunks) * 3 + 4) * 4); // (unchanged: the bad flag lives in word 1)
		if (s != HD_OK)
			return s;
		HD_CUDA_TRY(cudaMemsetAsync(p->dirty_scratch + 1, 0, 4, st));
		k_color_copy<<<148u * 4u, 256, 0, st>>>(p->color_nodes, p->color_leaves, p->color_node_cap, p->color_leaf_cap,
		                                        triples + size_t(n_app) * 3, n_chunks, pay, uint32_t(payload), true, p->dirty_scratch + 1);
		HD_LAUNCH_CHECK();
	}
	return HD_OK;
}

hd_status hd_dirty_pack_dev(hd_pool *p, void *staging_dev, uint64_t capacity_bytes, uint64_t *packed_bytes) {
	if (!p || !staging_dev || !packed_bytes || capacity_bytes < kHeaderWords * 4)
		return HD_ERR_INVALID;
	HD_CUDA_TRY(cudaSetDevice(p->device));
	// ONE pass, ONE host synchronisation (the interactive loop pays this every frame): the scan writes the range
	// triples straight behind the header, the gather reads n_ranges from the device header to place the payload
	// behind them, and only the header comes back.  A staging buffer that turns out too small is reported
	// with the size that is needed; nothing in the pool changes, so the caller grows the buffer and calls again.
	uint32_t *stg = static_cast<uint32_t *>(staging_dev);
	const uint64_t cap_words = std::min<uint64_t>(capacity_bytes / 4, 0xFFFFFFFFull);
	const uint32_t cap_ranges = uint32_t((cap_words - kHeaderWords) / 3);
	uint32_t head[kHeaderWords] = {0u, 0u, p->root, p->needs_full_resync ? kFlagClear : 0u, 0u, 0u, p->color_root, p->color_leaf_level};
	HD_CUDA_TRY(cudaMemcpyAsync(stg, head, sizeof(head), cudaMemcpyHostToDevice, p->stream));
	const uint32_t nb = p->geo.total_buckets;
	k_dirty_scan<<<(nb + 255) / 256, 256, 0, p->stream>>>(p->bucket_words, p->bucket_synced, nb, p->geo.bucket_shift(), stg,
	                                                      stg + kHeaderWords, cap_ranges);
	HD_LAUNCH_CHECK();
	k_dirty_gather<<<148u * 4u, 256, 0, p->stream>>>(p->words, stg, uint32_t(cap_words));
	HD_LAUNCH_CHECK();
	uint32_t h[kHeaderWords];
	HD_CUDA_TRY(cudaMemcpyAsync(h, stg, sizeof(h), cudaMemcpyDeviceToHost, p->stream));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	uint64_t need_words = uint64_t(kHeaderWords) + uint64_t(h[0]) * 3 + h[1];
	if (p->color_dirty || p->color_full_resync) { // the colour pool changed: its delta rides in the same blob
		uint64_t section = 0;
		hd_status cs = color_pack(p, stg, cap_words, need_words, &h[3], &h[4], &h[5], &section);
		if (cs != HD_OK)
			return cs;
		need_words += section;
		HD_CUDA_TRY(cudaMemcpyAsync(stg + 3, &h[3], 12, cudaMemcpyHostToDevice, p->stream));
		HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	}
	*packed_bytes = need_words * 4;
	if (need_words * 4 > capacity_bytes) {
		set_error("staging buffer too small: need %llu bytes", (unsigned long long)(need_words * 4));
		return HD_ERR_OVERFLOW;
	}
	return HD_OK;
}

static hd_status color_apply(hd_pool *p, const uint32_t *sec, const uint32_t h[kHeaderWords], uint32_t *bad_dev) {
	cudaStream_t st = p->stream;
	uint32_t totals[2];
	HD_CUDA_TRY(cudaMemcpyAsync(totals, sec, sizeof(totals), cudaMemcpyDeviceToHost, st));
	HD_CUDA_TRY(cudaStreamSynchronize(st));
	const uint64_t node_words = uint64_t(totals[0]) * 8, leaf_words = totals[1];
	if (totals[0] >= (1u << 30) || leaf_words > (1ull << 30)) { // 30-bit node ids / leaf word indices (DAGColorPool.hpp:23-34)
		set_error("colour section: totals out of range");
		return HD_ERR_INVALID;
	}
	if (h[3] & kFlagColorFull)
		p->color_node_words = p->color_leaf_words = 0; // nothing of the old buffers needs to survive a re-allocation
	hd_status s = ensure_color_storage(p, node_words, leaf_words);
	if (s != HD_OK)
		return s;
	if (h[4]) {
		const uint32_t *triples = sec + 2;
		uint32_t *payload = const_cast<uint32_t *>(triples + size_t(h[4]) * 3);
		k_color_copy<<<148u * 4u, 256, 0, st>>>(p->color_nodes, p->color_leaves, std::min<uint64_t>(p->color_node_cap, node_words),
		                                        std::min<uint64_t>(p->color_leaf_cap, leaf_words), triples, h[4], payload, h[5], false,
		                                        bad_dev);
		HD_LAUNCH_CHECK();
	}
	const uint32_t ctr[4] = {totals[0], totals[1], 0u, 0u};
	HD_CUDA_TRY(cudaMemcpyAsync(p->color_ctr, ctr, sizeof(ctr), cudaMemcpyHostToDevice, st));
	HD_CUDA_TRY(cudaStreamSynchronize(st)); // ctr is on the stack
	p->color_node_words = p->color_synced_node_words = node_words;
	p->color_leaf_words = p->color_synced_leaf_words = leaf_words;
	return HD_OK;
}

static hd_status dirty_apply(hd_pool *p, const void *staging_dev, uint64_t packed_bytes, bool validate_first) {
	if (!p || !staging_dev || packed_bytes < kHeaderWords * 4)
		return HD_ERR_INVALID;
	HD_CUDA_TRY(cudaSetDevice(p->device));
	const uint32_t *stg = static_cast<const uint32_t *>(staging_dev);
	uint32_t h[kHeaderWords];
	HD_CUDA_TRY(cudaMemcpyAsync(h, stg, sizeof(h), cudaMemcpyDeviceToHost, p->stream));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	const bool color = h[3] & kFlagColor;
	const uint64_t node_part = uint64_t(kHeaderWords) + uint64_t(h[0]) * 3 + h[1];
	const uint64_t total = node_part + (color ? 2 + uint64_t(h[4]) * 3 + h[5] : 0);
	if (total * 4 > packed_bytes || h[0] > p->geo.total_buckets || (!color && (h[4] || h[5]))) {
		set_error("staging buffer truncated or malformed");
		return HD_ERR_INVALID;
	}
	hd_status s = ensure_dirty_scratch(p, kHeaderWords * 4);
	if (s != HD_OK)
		return s;
	uint32_t *bad_dev = p->dirty_scratch, bad = 0;
	HD_CUDA_TRY(cudaMemsetAsync(bad_dev, 0, 4, p->stream));
	const uint32_t *ranges = stg + kHeaderWords, *payload = ranges + size_t(h[0]) * 3;
	const uint32_t grid = std::min<uint32_t>(std::max(h[0], 1u), 148u * 8u);
	if (validate_first && h[0]) { // untrusted input (a pool file): check every triple before the first write
		k_dirty_scatter<<<grid, 256, 0, p->stream>>>(p->words, p->bucket_words, p->bucket_synced, p->geo.bucket_shift(), p->geo.total_words,
		                                           stg, ranges, payload, true, bad_dev);
		HD_LAUNCH_CHECK();
		HD_CUDA_TRY(cudaMemcpyAsync(&bad, bad_dev, 4, cudaMemcpyDeviceToHost, p->stream));
		HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
		if (bad) {
			set_error("packed ranges fall outside the pool, a bucket or the payload");
			return HD_ERR_INVALID;
		}
	}
	if (h[3] & kFlagClear) { // the sender compacted its pool: every pointer changed, start from an empty replica
		hd_status cs = hd_pool_clear(p);
		if (cs != HD_OK)
			return cs;
	}
	if (h[0]) {
		p->tt_invalidate(); // ranges from a sender are appends, but nothing here proves it: drop the staged top levels
		k_dirty_scatter<<<grid, 256, 0, p->stream>>>(p->words, p->bucket_words, p->bucket_synced, p->geo.bucket_shift(), p->geo.total_words,
		                                           stg, ranges, payload, false, bad_dev);
		HD_LAUNCH_CHECK();
	}
	if (color) {
		s = color_apply(p, stg + node_part, h, bad_dev);
		if (s != HD_OK)
			return s;
	}
	HD_CUDA_TRY(cudaMemcpyAsync(&bad, bad_dev, 4, cudaMemcpyDeviceToHost, p->stream));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	if (bad) { // bad triples were skipped, nothing was written out of bounds; the root is NOT published
		set_error("packed ranges fall outside the pool, a bucket or the payload");
		return HD_ERR_INVALID;
	}
	if (color)
		p->color_root = h[6], p->color_leaf_level = h[7];
	p->root = h[2]; // publish the root last (src/main.cpp:281-296 ordering contract)
	return HD_OK;
}

hd_status hd_dirty_apply_dev(hd_pool *p, const void *staging_dev, uint64_t packed_bytes) {
	return dirty_apply(p, staging_dev, packed_bytes, false);
}

// ---- pool serialisation (the reference has no on-disk format) ---------------------------------------
// File = FileHeader, then the same packed staging blob replica sync uses (every non-empty bucket's used prefix as one
// range, root in the blob header), then the two colour buffers.
constexpr uint32_t kFileVersion = 2; // 2: 8-word blob header
struct FileHeader {
	char magic[8]; // "HDAGB200"
	uint32_t version; // kFileVersion
	hd_config cfg;
	uint64_t blob_bytes, color_node_words, color_leaf_words;
	uint32_t color_root, color_leaf_level; // DAGColorPool root + Config::leaf_level
};

hd_status hd_pool_save(hd_pool *p, const char *path) {
	if (!p || !path)
		return HD_ERR_INVALID;
	HD_CUDA_TRY(cudaSetDevice(p->device));
	const uint32_t nb = p->geo.total_buckets;
	uint32_t *zeros = nullptr, *stg = nullptr;
	ScopeExit guard{[&]() { cudaFree(zeros), cudaFree(stg); }};
	HD_CUDA_TRY(cudaMalloc(&zeros, size_t(nb) * 4));
	HD_CUDA_TRY(cudaMemsetAsync(zeros, 0, size_t(nb) * 4, p->stream));
	hd_status s = ensure_dirty_scratch(p, kHeaderWords * 4);
	if (s != HD_OK)
		return s;
	HD_CUDA_TRY(cudaMemsetAsync(p->dirty_scratch, 0, kHeaderWords * 4, p->stream));
	k_dirty_scan<<<(nb + 255) / 256, 256, 0, p->stream>>>(p->bucket_words, zeros, nb, p->geo.bucket_shift(),
	                                                      p->dirty_scratch, nullptr, 0);
	HD_LAUNCH_CHECK();
	uint32_t h[kHeaderWords];
	HD_CUDA_TRY(cudaMemcpyAsync(h, p->dirty_scratch, sizeof(h), cudaMemcpyDeviceToHost, p->stream));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	const uint64_t blob = (uint64_t(kHeaderWords) + uint64_t(h[0]) * 3 + h[1]) * 4;
	HD_CUDA_TRY(cudaMalloc(&stg, blob));
	HD_CUDA_TRY(cudaMemsetAsync(stg, 0, kHeaderWords * 4, p->stream));
	k_dirty_scan<<<(nb + 255) / 256, 256, 0, p->stream>>>(p->bucket_words, zeros, nb, p->geo.bucket_shift(), stg,
	                                                      stg + kHeaderWords, h[0]);
	HD_LAUNCH_CHECK();
	if (h[0]) {
		k_dirty_gather<<<std::min<uint32_t>(h[0], 148u * 8u), 256, 0, p->stream>>>(p->words, stg, uint32_t(blob / 4));
		HD_LAUNCH_CHECK();
	}
	std::vector<uint32_t> host, cn, cl;
	try {
		host.resize(blob / 4), cn.resize(p->color_node_words), cl.resize(p->color_leaf_words);
	} catch (...) { // nothing may throw through the C ABI
		set_error("pool save: out of host memory");
		return HD_ERR_OOM;
	}
	HD_CUDA_TRY(cudaMemcpyAsync(host.data(), stg, blob, cudaMemcpyDeviceToHost, p->stream));
	if (!cn.empty())
		HD_CUDA_TRY(cudaMemcpyAsync(cn.data(), p->color_nodes, cn.size() * 4, cudaMemcpyDeviceToHost, p->stream));
	if (!cl.empty())
		HD_CUDA_TRY(cudaMemcpyAsync(cl.data(), p->color_leaves, cl.size() * 4, cudaMemcpyDeviceToHost, p->stream));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	host[2] = p->root;
	FileHeader fh{};
	memcpy(fh.magic, "HDAGB200", 8);
	fh.version = kFileVersion, fh.cfg = p->cfg, fh.blob_bytes = blob;
	fh.color_node_words = cn.size(), fh.color_leaf_words = cl.size();
	fh.color_root = p->color_root, fh.color_leaf_level = p->color_leaf_level;
	FILE *f = fopen(path, "wb");
	if (!f) {
		set_error("cannot open %s for writing", path);
		return HD_ERR_INVALID;
	}
	bool ok = fwrite(&fh, sizeof(fh), 1, f) == 1 && fwrite(host.data(), 1, blob, f) == blob;
	ok = ok && (cn.empty() || fwrite(cn.data(), 4, cn.size(), f) == cn.size());
	ok = ok && (cl.empty() || fwrite(cl.data(), 4, cl.size(), f) == cl.size());
	ok = (fclose(f) == 0) && ok;
	if (!ok) {
		set_error("short write to %s", path);
		return HD_ERR_INVALID;
	}
	return HD_OK;
}

// a colour pointer must reference storage that exists (tag 0: node id, tag 2: leaf chunk word index)
static bool color_root_valid(uint32_t root, uint64_t node_words, uint64_t leaf_words) {
	const uint32_t tag = root >> 30, data = root & 0x3FFFFFFFu;
	if (tag == 0u)
		return (uint64_t(data) + 1) * 8 <= node_words;
	if (tag == 2u)
		return uint64_t(data) + 4 <= leaf_words;
	return true;
}

hd_status hd_pool_load(const char *path, int device, hd_pool **out) {
	if (!path || !out)
		return HD_ERR_INVALID;
	*out = nullptr;
	FILE *f = fopen(path, "rb");
	if (!f) {
		set_error("cannot open %s", path);
		return HD_ERR_INVALID;
	}
	// The file is untrusted: every size is checked against the real file length before anything is allocated, and the
	// allocations themselves cannot throw through the C ABI.
	FileHeader fh{};
	std::vector<uint32_t> blob, cn, cl;
	bool ok = fseek(f, 0, SEEK_END) == 0;
	const long file_len = ok ? ftell(f) : -1;
	ok = ok && file_len >= long(sizeof(fh)) && fseek(f, 0, SEEK_SET) == 0;
	ok = ok && fread(&fh, sizeof(fh), 1, f) == 1 && memcmp(fh.magic, "HDAGB200", 8) == 0 && fh.version == kFileVersion &&
	     fh.blob_bytes >= kHeaderWords * 4 && fh.blob_bytes % 4 == 0;
	if (ok) {
		const uint64_t rest = uint64_t(file_len) - sizeof(fh);
		ok = fh.blob_bytes <= rest && fh.color_node_words % 8 == 0 && fh.color_node_words <= (rest - fh.blob_bytes) / 4 &&
		     fh.color_leaf_words <= (rest - fh.blob_bytes) / 4 - fh.color_node_words &&
		     fh.blob_bytes + 4 * (fh.color_node_words + fh.color_leaf_words) == rest;
		ok = ok && color_root_valid(fh.color_root, fh.color_node_words, fh.color_leaf_words) &&
		     (fh.color_leaf_level == 0 || fh.color_leaf_level + 2 <= fh.cfg.node_levels);
	}
	if (ok) {
		try {
			blob.resize(fh.blob_bytes / 4), cn.resize(fh.color_node_words), cl.resize(fh.color_leaf_words);
		} catch (...) {
			fclose(f);
			set_error("pool load: out of host memory");
			return HD_ERR_OOM;
		}
		ok = fread(blob.data(), 1, fh.blob_bytes, f) == fh.blob_bytes;
		ok = ok && (cn.empty() || fread(cn.data(), 4, cn.size(), f) == cn.size());
		ok = ok && (cl.empty() || fread(cl.data(), 4, cl.size(), f) == cl.size());
	}
	fclose(f);
	if (!ok) {
		set_error("%s is not a valid hashdag_b200 pool file (bad magic/version, or sizes that do not match the file length)", path);
		return HD_ERR_INVALID;
	}
	hd_pool *p = nullptr;
	hd_status s = hd_pool_create(&fh.cfg, device, &p);
	if (s != HD_OK)
		return s;
	uint32_t *stg = nullptr;
	cudaError_t e = cudaMalloc(&stg, fh.blob_bytes);
	if (e == cudaSuccess)
		e = cudaMemcpy(stg, blob.data(), fh.blob_bytes, cudaMemcpyHostToDevice);
	if (e != cudaSuccess) {
		set_error("pool load: %s", cudaGetErrorString(e));
		cudaFree(stg);
		hd_pool_destroy(p);
		return HD_ERR_CUDA;
	}
	s = dirty_apply(p, stg, fh.blob_bytes, true); // every range triple is bounds-checked before the first write
	cudaFree(stg);
	if (s == HD_OK && (!cn.empty() || !cl.empty()))
		s = hd_color_upload(p, cn.data(), cn.size(), cl.data(), cl.size());
	if (s == HD_OK)
		p->color_root = fh.color_root, p->color_leaf_level = fh.color_leaf_level;
	if (s != HD_OK) {
		hd_pool_destroy(p);
		return s;
	}
	*out = p;
	return HD_OK;
}

hd_status hd_dirty_reset(hd_pool *p) {
	if (!p)
		return HD_ERR_INVALID;
	HD_CUDA_TRY(cudaSetDevice(p->device));
	HD_CUDA_TRY(cudaMemcpyAsync(p->bucket_synced, p->bucket_words, size_t(p->geo.total_buckets) * 4,
	                            cudaMemcpyDeviceToDevice, p->stream));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	p->needs_full_resync = false;
	// the colour pool is in sync as well: the append ranges start here, the list of rewritten chunks is empty again
	p->color_synced_node_words = p->color_node_words, p->color_synced_leaf_words = p->color_leaf_words;
	p->color_dirty = p->color_full_resync = false;
	if (p->color_dirty_ctr)
		HD_CUDA_TRY(cudaMemsetAsync(p->color_dirty_ctr, 0, 2 * sizeof(uint32_t), p->stream));
	return HD_OK;
}

} // extern "C"
