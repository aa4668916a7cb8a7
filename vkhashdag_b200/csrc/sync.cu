// sync.cu — dirty-range tracking and replica synchronisation (multi-GPU: pool replicated).
//
// Replaces DAGNodePool::Flush (src/DAGNodePool.cpp:56-85) and its per-page m_page_write_ranges bookkeeping
// (src/DAGNodePool.hpp:44-46,62-69).  Buckets are append-only between GCs, so "what changed since the last
// sync" is exactly [bucket_synced[b], bucket_words[b]) for every bucket: no write log is needed.  The editing
// rank packs those runs into ONE staging buffer, the caller broadcasts it (one NCCL broadcast over NVLink),
// replicas scatter it and publish the root last.
//
// Staging layout (u32 words): [n_ranges][payload_words][root][0] then n_ranges x {word_offset, word_count,
// payload_offset} then the payload.
#include "common.cuh"

#include <algorithm>
#include <cstring>

namespace hd {

constexpr uint32_t kHeaderWords = 4;

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

__global__ void k_dirty_scatter(uint32_t *words, uint32_t *bucket_words, uint32_t *bucket_synced, uint32_t bucket_shift,
                                const uint32_t *__restrict__ header, const uint32_t *__restrict__ ranges,
                                const uint32_t *__restrict__ payload) {
	const uint32_t n = header[0];
	for (uint32_t r = blockIdx.x; r < n; r += gridDim.x) {
		const uint32_t off = ranges[r * 3], cnt = ranges[r * 3 + 1], poff = ranges[r * 3 + 2];
		for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x)
			words[off + i] = payload[poff + i];
		if (threadIdx.x == 0) {
			const uint32_t bucket = off >> bucket_shift;
			const uint32_t end = (off & ((1u << bucket_shift) - 1u)) + cnt;
			bucket_words[bucket] = end;
			bucket_synced[bucket] = end;
		}
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

hd_status hd_dirty_pack_dev(hd_pool *p, void *staging_dev, uint64_t capacity_bytes, uint64_t *packed_bytes) {
	if (!p || !staging_dev || !packed_bytes || capacity_bytes < kHeaderWords * 4)
		return HD_ERR_INVALID;
	HD_CUDA_TRY(cudaSetDevice(p->device));
	// ONE pass, ONE host synchronisation (the interactive loop pays this every frame): the scan writes the range
	// triples straight behind the header, the gather reads n_ranges from the device header to place the payload
	// behind them, and only the 16-byte header comes back.  A staging buffer that turns out too small is reported
	// with the size that is needed; nothing in the pool changes, so the caller grows the buffer and calls again.
	uint32_t *stg = static_cast<uint32_t *>(staging_dev);
	const uint64_t cap_words = std::min<uint64_t>(capacity_bytes / 4, 0xFFFFFFFFull);
	const uint32_t cap_ranges = uint32_t((cap_words - kHeaderWords) / 3);
	const uint32_t head[kHeaderWords] = {0u, 0u, p->root, p->needs_full_resync ? 1u : 0u}; // [3] = 1: clear first (after a GC)
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
	const uint64_t need = (uint64_t(kHeaderWords) + uint64_t(h[0]) * 3 + h[1]) * 4;
	*packed_bytes = need;
	if (need > capacity_bytes) {
		set_error("staging buffer too small: need %llu bytes", (unsigned long long)need);
		return HD_ERR_OVERFLOW;
	}
	return HD_OK;
}

hd_status hd_dirty_apply_dev(hd_pool *p, const void *staging_dev, uint64_t packed_bytes) {
	if (!p || !staging_dev || packed_bytes < kHeaderWords * 4)
		return HD_ERR_INVALID;
	HD_CUDA_TRY(cudaSetDevice(p->device));
	const uint32_t *stg = static_cast<const uint32_t *>(staging_dev);
	uint32_t h[kHeaderWords];
	HD_CUDA_TRY(cudaMemcpyAsync(h, stg, sizeof(h), cudaMemcpyDeviceToHost, p->stream));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	if ((uint64_t(kHeaderWords) + uint64_t(h[0]) * 3 + h[1]) * 4 > packed_bytes) {
		set_error("staging buffer truncated");
		return HD_ERR_INVALID;
	}
	if (h[3] == 1u) { // the sender compacted its pool: every pointer changed, start from an empty replica
		hd_status cs = hd_pool_clear(p);
		if (cs != HD_OK)
			return cs;
	}
	if (h[0]) {
		const uint32_t *ranges = stg + kHeaderWords, *payload = ranges + size_t(h[0]) * 3;
		k_dirty_scatter<<<std::min<uint32_t>(h[0], 148u * 8u), 256, 0, p->stream>>>(
		    p->words, p->bucket_words, p->bucket_synced, p->geo.bucket_shift(), stg, ranges, payload);
		HD_LAUNCH_CHECK();
	}
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	p->root = h[2]; // publish the root last (src/main.cpp:281-296 ordering contract)
	return HD_OK;
}

// ---- pool serialisation (the reference has no on-disk format) ---------------------------------------
// File = FileHeader, then the same packed staging blob replica sync uses (every non-empty bucket's used prefix as one
// range, root in the blob header), then the two colour buffers.
struct FileHeader {
	char magic[8]; // "HDAGB200"
	uint32_t version;
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
	std::vector<uint32_t> host(blob / 4);
	HD_CUDA_TRY(cudaMemcpyAsync(host.data(), stg, blob, cudaMemcpyDeviceToHost, p->stream));
	std::vector<uint32_t> cn(p->color_node_words), cl(p->color_leaf_words);
	if (!cn.empty())
		HD_CUDA_TRY(cudaMemcpyAsync(cn.data(), p->color_nodes, cn.size() * 4, cudaMemcpyDeviceToHost, p->stream));
	if (!cl.empty())
		HD_CUDA_TRY(cudaMemcpyAsync(cl.data(), p->color_leaves, cl.size() * 4, cudaMemcpyDeviceToHost, p->stream));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	cudaFree(zeros), cudaFree(stg);
	host[2] = p->root;
	FileHeader fh{};
	memcpy(fh.magic, "HDAGB200", 8);
	fh.version = 1, fh.cfg = p->cfg, fh.blob_bytes = blob;
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

hd_status hd_pool_load(const char *path, int device, hd_pool **out) {
	if (!path || !out)
		return HD_ERR_INVALID;
	*out = nullptr;
	FILE *f = fopen(path, "rb");
	if (!f) {
		set_error("cannot open %s", path);
		return HD_ERR_INVALID;
	}
	FileHeader fh{};
	std::vector<uint32_t> blob, cn, cl;
	bool ok = fread(&fh, sizeof(fh), 1, f) == 1 && memcmp(fh.magic, "HDAGB200", 8) == 0 && fh.version == 1 &&
	          fh.blob_bytes >= kHeaderWords * 4 && fh.blob_bytes % 4 == 0;
	if (ok) {
		blob.resize(fh.blob_bytes / 4), cn.resize(fh.color_node_words), cl.resize(fh.color_leaf_words);
		ok = fread(blob.data(), 1, fh.blob_bytes, f) == fh.blob_bytes;
		ok = ok && (cn.empty() || fread(cn.data(), 4, cn.size(), f) == cn.size());
		ok = ok && (cl.empty() || fread(cl.data(), 4, cl.size(), f) == cl.size());
	}
	fclose(f);
	if (!ok) {
		set_error("%s is not a valid hashdag_b200 pool file", path);
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
	s = hd_dirty_apply_dev(p, stg, fh.blob_bytes);
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
	return HD_OK;
}

} // extern "C"
