// pool.cu — device pool lifetime, host<->device mirror interop and config helpers of the C ABI.
//
// Replaces src/DAGNodePool.{hpp,cpp} (host page array + sparse VkBuffer + Flush) with ONE flat device
// allocation: pointer = global word index exactly as in the reference, so reference-built
// pages upload verbatim and GPU-built pools read back into a reference-compatible host mirror.
#include "common.cuh"

#include <cstdarg>
#include <cstring>

namespace hd {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}

bool make_geometry(const hd_config &cfg, Geometry &g) {
	// include/hashdag/Config.hpp:48-56 Validate + :33-46 derived sizes
	if (cfg.node_levels == 0 || cfg.node_levels > HD_MAX_NODE_LEVELS || cfg.word_bits_per_page < 4)
		return false;
	if (cfg.word_bits_per_page + cfg.page_bits_per_bucket > 31)
		return false;
	g.word_bits_per_page = cfg.word_bits_per_page;
	g.page_bits_per_bucket = cfg.page_bits_per_bucket;
	g.node_levels = cfg.node_levels;
	uint64_t buckets = 0;
	for (uint32_t l = 0; l < HD_MAX_NODE_LEVELS; ++l) {
		g.bucket_bits[l] = l < cfg.node_levels ? cfg.bucket_bits_each_level[l] : 0;
		g.level_base[l] = uint32_t(buckets);
		if (l < cfg.node_levels) {
			if (cfg.bucket_bits_each_level[l] > 31)
				return false;
			buckets += 1ull << cfg.bucket_bits_each_level[l];
		}
	}
	uint64_t words = buckets << (cfg.word_bits_per_page + cfg.page_bits_per_bucket);
	if (buckets > 0xFFFFFFFFull || words - 1 > 0xFFFFFFFEull)
		return false;
	g.total_buckets = uint32_t(buckets);
	g.total_words = words;
	return true;
}

} // namespace hd

using namespace hd;

namespace hd {
// hd_pool_read_subtree: one CTA walks the subtree level by level; the records of a level are the frontier of the next.
__global__ void __launch_bounds__(256) k_read_subtree(const uint32_t *__restrict__ words, uint32_t node_levels, uint32_t root,
                                                      uint32_t level, uint32_t depth, hd_node_record *out, uint32_t capacity,
                                                      uint32_t *n_out /* [0] = records, [1] = truncated */) {
	__shared__ uint32_t s_begin, s_end, s_next;
	if (threadIdx.x == 0) {
		s_begin = 0, s_end = 1, s_next = 1;
		out[0].ptr = root, out[0].level = level;
		n_out[1] = 0;
	}
	__syncthreads();
	for (uint32_t d = 0;; ++d) {
		const uint32_t begin = s_begin, end = s_end;
		for (uint32_t i = begin + threadIdx.x; i < end; i += blockDim.x) {
			hd_node_record &r = out[i];
			const bool leaf = r.level == node_levels - 1u;
			const uint32_t first = words[r.ptr];
			const uint32_t n = leaf ? 2u : 1u + __popc(first & 0xFFu);
			r.n_words = n;
			r.words[0] = first;
			for (uint32_t k = 1; k < n; ++k)
				r.words[k] = words[r.ptr + k];
			for (uint32_t k = n; k < 9u; ++k)
				r.words[k] = 0u;
			if (!leaf && d + 1 < depth) {
				const uint32_t slot = atomicAdd(&s_next, n - 1u);
				for (uint32_t k = 1; k < n; ++k)
					if (slot + k - 1u < capacity)
						out[slot + k - 1u].ptr = r.words[k], out[slot + k - 1u].level = r.level + 1u;
					else
						n_out[1] = 1u;
			}
		}
		__syncthreads();
		if (threadIdx.x == 0)
			s_begin = end, s_end = min(s_next, capacity), s_next = min(s_next, capacity);
		__syncthreads();
		if (s_begin == s_end)
			break;
	}
	if (threadIdx.x == 0)
		n_out[0] = s_end;
}
} // namespace hd

extern "C" {

const char *hd_version(void) { return "hashdag_b200 0.1 (sm_100a)"; }
const char *hd_last_error(void) { return g_err; }

int hd_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess)
		return 0;
	return n;
}

hd_status hd_config_from_default(const hd_default_config *dc, hd_config *out) {
	if (!dc || !out || dc->level_count < 2 || dc->level_count - 1 > HD_MAX_NODE_LEVELS)
		return HD_ERR_INVALID;
	memset(out, 0, sizeof(*out));
	out->word_bits_per_page = dc->word_bits_per_page;
	out->page_bits_per_bucket = dc->page_bits_per_bucket;
	out->node_levels = dc->level_count - 1;
	for (uint32_t l = 0; l + 1 < dc->level_count; ++l)
		out->bucket_bits_each_level[l] =
		    l < dc->top_level_count ? dc->bucket_bits_per_top_level : dc->bucket_bits_per_bottom_level;
	return HD_OK;
}

int hd_config_validate(const hd_config *cfg) {
	Geometry g;
	return cfg && make_geometry(*cfg, g) ? 1 : 0;
}
uint32_t hd_config_total_buckets(const hd_config *cfg) {
	Geometry g;
	return cfg && make_geometry(*cfg, g) ? g.total_buckets : 0;
}
uint64_t hd_config_total_words(const hd_config *cfg) {
	Geometry g;
	return cfg && make_geometry(*cfg, g) ? g.total_words : 0;
}
uint32_t hd_config_level_base_bucket(const hd_config *cfg, uint32_t level) {
	Geometry g;
	return cfg && make_geometry(*cfg, g) && level < g.node_levels ? g.level_base[level] : 0;
}

hd_status hd_pool_create(const hd_config *cfg, int device, hd_pool **out) {
	if (!cfg || !out)
		return HD_ERR_INVALID;
	*out = nullptr;
	Geometry g;
	if (!make_geometry(*cfg, g)) {
		set_error("config fails Config::Validate");
		return HD_ERR_INVALID;
	}
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
		set_error("no CUDA device: libhashdag_b200 has no CPU fallback");
		return HD_ERR_NO_DEVICE;
	}
	if (device < 0 || device >= n)
		return HD_ERR_INVALID;
	HD_CUDA_TRY(cudaSetDevice(device));
	hd_pool *p = new hd_pool();
	p->cfg = *cfg;
	p->geo = g;
	p->device = device;
	cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device);
	cudaError_t e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
	if (e == cudaSuccess)
		e = cudaMalloc(&p->words, g.total_words * sizeof(uint32_t));
	if (e == cudaSuccess)
		e = cudaMalloc(&p->bucket_words, size_t(g.total_buckets) * sizeof(uint32_t));
	if (e == cudaSuccess)
		e = cudaMalloc(&p->bucket_synced, size_t(g.total_buckets) * sizeof(uint32_t));
	if (e == cudaSuccess)
		e = cudaMalloc(&p->params_dev, sizeof(hd_trace_params));
	// The whole word space starts zeroed and nothing outside [0, bucket_words) of a bucket is ever written, so
	// page-tail padding (NodePool.hpp:145-152) is zero by construction (a device-side editor
	// must own its padding).
	if (e == cudaSuccess)
		e = cudaMemsetAsync(p->words, 0, g.total_words * sizeof(uint32_t), p->stream);
	if (e == cudaSuccess)
		e = cudaMemsetAsync(p->bucket_words, 0, size_t(g.total_buckets) * sizeof(uint32_t), p->stream);
	if (e == cudaSuccess)
		e = cudaMemsetAsync(p->bucket_synced, 0, size_t(g.total_buckets) * sizeof(uint32_t), p->stream);
	if (e == cudaSuccess)
		e = cudaStreamSynchronize(p->stream);
	if (e != cudaSuccess) {
		set_error("pool allocation failed: %s", cudaGetErrorString(e));
		hd_pool_destroy(p);
		return e == cudaErrorMemoryAllocation ? HD_ERR_OOM : HD_ERR_CUDA;
	}
	// keep stream-ordered scratch cached between edits
	cudaMemPool_t mp;
	if (cudaDeviceGetDefaultMemPool(&mp, device) == cudaSuccess) {
		uint64_t thr = UINT64_MAX;
		cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &thr);
	}
	*out = p;
	return HD_OK;
}

void hd_pool_destroy(hd_pool *p) {
	if (!p)
		return;
	cudaSetDevice(p->device);
	if (p->stream)
		cudaStreamSynchronize(p->stream);
	edit_scratch_free(p);
	cudaFree(p->words);
	cudaFree(p->bucket_words);
	cudaFree(p->bucket_synced);
	cudaFree(p->color_nodes);
	cudaFree(p->color_leaves);
	cudaFree(p->color_ctr);
	cudaFree(p->color_dirty_list);
	cudaFree(p->color_dirty_ctr);
	cudaFree(p->stage_rgba);
	cudaFree(p->stage_iters);
	cudaFree(p->stage_fetches);
	cudaFree(p->stage_hits);
	cudaFree(p->params_dev);
	cudaFree(p->ray_table);
	cudaFree(p->persist_ctr);
	cudaFree(p->tt_entries), cudaFree(p->tt_masks), cudaFree(p->tt_list[0]), cudaFree(p->tt_list[1]), cudaFree(p->tt_count);
	cudaFree(p->dirty_scratch);
	cudaFreeHost(p->pick_host);
	for (int i = 0; i < 2; ++i) {
		cudaFree(p->pipe_rgba[i]);
		if (p->pipe_traced[i])
			cudaEventDestroy(p->pipe_traced[i]);
		if (p->pipe_done[i])
			cudaEventDestroy(p->pipe_done[i]);
	}
	if (p->copy_stream)
		cudaStreamDestroy(p->copy_stream);
	if (p->color_stream)
		cudaStreamDestroy(p->color_stream);
	if (p->stream)
		cudaStreamDestroy(p->stream);
	delete p;
}

hd_status hd_host_alloc(uint64_t bytes, int write_combined, void **out) {
	if (!out || bytes == 0)
		return HD_ERR_INVALID;
	*out = nullptr;
	HD_CUDA_TRY(cudaHostAlloc(out, bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
	return HD_OK;
}
hd_status hd_host_free(void *ptr) {
	if (ptr)
		HD_CUDA_TRY(cudaFreeHost(ptr));
	return HD_OK;
}

hd_status hd_pool_get_config(const hd_pool *p, hd_config *out) {
	if (!p || !out)
		return HD_ERR_INVALID;
	*out = p->cfg;
	return HD_OK;
}

hd_status hd_pool_clear(hd_pool *p) {
	if (!p)
		return HD_ERR_INVALID;
	HD_CUDA_TRY(cudaSetDevice(p->device));
	HD_CUDA_TRY(cudaMemsetAsync(p->words, 0, p->geo.total_words * sizeof(uint32_t), p->stream));
	HD_CUDA_TRY(cudaMemsetAsync(p->bucket_words, 0, size_t(p->geo.total_buckets) * 4, p->stream));
	HD_CUDA_TRY(cudaMemsetAsync(p->bucket_synced, 0, size_t(p->geo.total_buckets) * 4, p->stream));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	p->filled.clear();
	p->root = HD_NULL_NODE;
	p->tt_invalidate();
	return HD_OK;
}

hd_status hd_pool_set_root(hd_pool *p, uint32_t root) {
	if (!p)
		return HD_ERR_INVALID;
	p->root = root;
	return HD_OK;
}
uint32_t hd_pool_get_root(const hd_pool *p) { return p ? p->root : HD_NULL_NODE; }
void *hd_pool_words_dev(hd_pool *p) { return p ? p->words : nullptr; }
void *hd_pool_bucket_words_dev(hd_pool *p) { return p ? p->bucket_words : nullptr; }
void *hd_pool_stream(hd_pool *p) { return p ? (void *)p->stream : nullptr; }

hd_status hd_pool_upload_words(hd_pool *p, uint32_t off, const uint32_t *src, uint32_t count) {
	if (!p || (!src && count) || uint64_t(off) + count > p->geo.total_words)
		return HD_ERR_INVALID;
	HD_CUDA_TRY(cudaSetDevice(p->device));
	HD_CUDA_TRY(cudaMemcpyAsync(p->words + off, src, size_t(count) * 4, cudaMemcpyHostToDevice, p->stream));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	p->tt_invalidate(); // arbitrary words were rewritten
	return HD_OK;
}
hd_status hd_pool_read_words(hd_pool *p, uint32_t off, uint32_t *dst, uint32_t count) {
	if (!p || (!dst && count) || uint64_t(off) + count > p->geo.total_words)
		return HD_ERR_INVALID;
	HD_CUDA_TRY(cudaSetDevice(p->device));
	HD_CUDA_TRY(cudaMemcpyAsync(dst, p->words + off, size_t(count) * 4, cudaMemcpyDeviceToHost, p->stream));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	return HD_OK;
}
hd_status hd_pool_read_subtree(hd_pool *p, uint32_t root, uint32_t level, uint32_t depth, hd_node_record *out, uint32_t capacity,
                               uint32_t *n_out) {
	if (!p || !out || !n_out || capacity == 0 || depth == 0 || level >= p->geo.node_levels)
		return HD_ERR_INVALID;
	*n_out = 0;
	if (root == HD_NULL_NODE)
		return HD_OK;
	if (uint64_t(root) + 2 > p->geo.total_words)
		return HD_ERR_INVALID;
	HD_CUDA_TRY(cudaSetDevice(p->device));
	hd_node_record *dev = nullptr;
	uint32_t *cnt = nullptr, h[2] = {0, 0};
	ScopeExit guard{[&]() {
		if (dev)
			cudaFreeAsync(dev, p->stream);
		if (cnt)
			cudaFreeAsync(cnt, p->stream);
	}};
	HD_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void **>(&dev), size_t(capacity) * sizeof(hd_node_record), p->stream));
	HD_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void **>(&cnt), 8, p->stream));
	k_read_subtree<<<1, 256, 0, p->stream>>>(p->words, p->geo.node_levels, root, level, depth, dev, capacity, cnt);
	HD_LAUNCH_CHECK();
	HD_CUDA_TRY(cudaMemcpyAsync(h, cnt, 8, cudaMemcpyDeviceToHost, p->stream));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	HD_CUDA_TRY(cudaMemcpyAsync(out, dev, size_t(h[0]) * sizeof(hd_node_record), cudaMemcpyDeviceToHost, p->stream));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	*n_out = h[0];
	if (h[1]) {
		set_error("hd_pool_read_subtree: more than %u nodes, the deepest levels are missing", capacity);
		return HD_ERR_OVERFLOW;
	}
	return HD_OK;
}
hd_status hd_pool_upload_bucket_words(hd_pool *p, uint32_t first, const uint32_t *src, uint32_t count) {
	if (!p || (!src && count) || uint64_t(first) + count > p->geo.total_buckets)
		return HD_ERR_INVALID;
	HD_CUDA_TRY(cudaSetDevice(p->device));
	HD_CUDA_TRY(cudaMemcpyAsync(p->bucket_words + first, src, size_t(count) * 4, cudaMemcpyHostToDevice, p->stream));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	return HD_OK;
}
hd_status hd_pool_read_bucket_words(hd_pool *p, uint32_t first, uint32_t *dst, uint32_t count) {
	if (!p || (!dst && count) || uint64_t(first) + count > p->geo.total_buckets)
		return HD_ERR_INVALID;
	HD_CUDA_TRY(cudaSetDevice(p->device));
	HD_CUDA_TRY(cudaMemcpyAsync(dst, p->bucket_words + first, size_t(count) * 4, cudaMemcpyDeviceToHost, p->stream));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	return HD_OK;
}

hd_status hd_pool_filled_nodes(hd_pool *p, uint32_t *out) {
	if (!p || !out)
		return HD_ERR_INVALID;
	hd_status s = ensure_filled(p);
	if (s != HD_OK)
		return s;
	for (uint32_t l = 0; l < p->geo.node_levels; ++l)
		out[l] = p->filled[l];
	return HD_OK;
}

hd_status hd_sync(hd_pool *p) {
	if (!p)
		return HD_ERR_INVALID;
	HD_CUDA_TRY(cudaSetDevice(p->device));
	HD_CUDA_TRY(cudaStreamSynchronize(p->stream));
	return HD_OK;
}

uint64_t hd_kernel_launches(void) { return g_launches.load(); }

} // extern "C"
