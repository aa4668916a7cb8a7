/*
 * hashdag_b200.h — C ABI of the B200-native HashDAG engine (libhashdag_b200.so).
 *
 * This is the drop-in boundary for the traversal + edit path of AdamYuan/VkHashDAG.  The reference
 * has no FFI of its own: its seam is C++20 concepts inside one binary.  Each entry
 * point below names the reference interface it replaces (paths relative to the reference tree).
 * Templates cannot cross a C ABI, so editors cross as POD descriptors (hd_edit_desc) that select a
 * compiled-in device predicate.
 *
 * Conventions: every call returns an hd_status (0 = OK) and never throws; pointers are plain host
 * pointers unless the name says "_dev" / the function says "device"; one pool per CUDA device;
 * calls on one pool are externally serialised, except that hd_trace* may overlap an edit on another
 * stream (node memory is append-only between GCs, include/hashdag/NodePool.hpp:134-157).
 */
#ifndef HASHDAG_B200_H
#define HASHDAG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HD_NULL_NODE 0xFFFFFFFFu /* include/hashdag/NodePointer.hpp:18 */
#define HD_MAX_NODE_LEVELS 22u   /* float-mantissa stack: include/hashdag/NodePoolTraversal.hpp:104 */
#define HD_COLOR_NULL 0xC0000000u /* src/DAGColorPool.hpp:23-34, tag 3 */

typedef enum hd_status {
	HD_OK = 0,
	HD_ERR_INVALID = 1,  /* bad argument / config fails Config::Validate (include/hashdag/Config.hpp:48-56) */
	HD_ERR_CUDA = 2,     /* a CUDA runtime call failed; see hd_last_error() */
	HD_ERR_OOM = 3,      /* device allocation failed */
	HD_ERR_OVERFLOW = 4, /* edit scratch (work queues) exhausted; pool unchanged beyond appended garbage */
	HD_ERR_NO_DEVICE = 5 /* no CUDA device: there is no CPU fallback in the product */
} hd_status;

/* include/hashdag/Config.hpp:15-57 — pool geometry.  node_levels = bucket_bits_each_level.size(). */
typedef struct hd_config {
	uint32_t word_bits_per_page;
	uint32_t page_bits_per_bucket;
	uint32_t node_levels;
	uint32_t bucket_bits_each_level[HD_MAX_NODE_LEVELS];
} hd_config;

/* include/hashdag/Config.hpp:59-75 — DefaultConfig{}() */
typedef struct hd_default_config {
	uint32_t level_count, top_level_count;
	uint32_t word_bits_per_page, page_bits_per_bucket;
	uint32_t bucket_bits_per_top_level, bucket_bits_per_bottom_level;
} hd_default_config;

/* Editors (src/main.cpp:32-150) as POD.  Coordinates are voxel-level integers.
 *  AABB_FILL    : p0 = aabb_min, p1 = aabb_max (exclusive)         main.cpp:32-70
 *  SPHERE_FILL  : p0 = center, r2                                  main.cpp:72-150 (EditMode::kFill)
 *  SPHERE_DIG   : p0 = center, r2                                  (EditMode::kDig)
 *  TERRAIN_FILL : integer value-noise height field (synthetic scene generator):
 *                 aux = seed, p0 = {base_height, first_cell_bits, octaves}, p1 = {first_amplitude, extent_bits, 0}
 *                 (extent_bits != 0: only columns x,z < 2^extent_bits carry terrain — a patch of a larger world);
 *                 octave o has lattice cell 2^(first_cell_bits-2o) voxels and amplitude first_amplitude>>(2o);
 *                 voxel (x,y,z) becomes solid iff y < height(x,z).
 */
typedef enum hd_edit_kind {
	HD_EDIT_AABB_FILL = 0,
	HD_EDIT_SPHERE_FILL = 1,
	HD_EDIT_SPHERE_DIG = 2,
	HD_EDIT_TERRAIN_FILL = 3
} hd_edit_kind;

typedef struct hd_edit_desc {
	uint32_t kind;
	uint32_t p0[3];
	uint32_t p1[3];
	uint32_t aux;
	uint64_t r2;
} hd_edit_desc;

/* Counters of one hd_edit_batch call (all monotone sums over the batch). */
typedef struct hd_edit_stats {
	uint64_t visited_nodes;   /* inner work items expanded (edit_node calls, NodePool.hpp:362) */
	uint64_t visited_leaves;  /* leaf work items (edit_leaf calls, NodePool.hpp:319) */
	uint64_t upserts;         /* upsert_node calls issued (NodePool.hpp:159) */
	uint64_t appended_nodes;  /* nodes that missed and were appended */
	uint64_t appended_words;  /* words appended including page padding */
	uint64_t overflow_count;  /* full buckets hit (NodePool.hpp:137-139,195): >0 voids parity */
	uint64_t in_range_voxels; /* reserved (computed by callers analytically) */
	uint64_t scan_words;      /* bucket words the GPU lookup actually read */
} hd_edit_stats;

/* src/rg/TracePass.cpp:10-18 tracer_pass::PC_Data == shader/src/trace.frag:15-23 push constants (84 B). */
typedef struct hd_trace_params {
	float pos[3], look[3], side[3], up[3];
	uint32_t width, height;
	uint32_t voxel_level;
	uint32_t dag_root, dag_leaf_level;
	uint32_t color_root, color_leaf_level;
	float proj_factor;
	uint32_t type; /* 0 diffuse*colour, 1 normal, 2 iteration heat (trace.frag:395-401) */
} hd_trace_params;

/* Per-pixel parity record (16 B): vox_pos (trace.frag:240-246) and
 * packed = hit<<31 | log2(vox_size)<<24 | fetched colour as RGB8 (r in bits 0..7), 0 when !hit. */
typedef struct hd_hit_record {
	uint32_t vox[3];
	uint32_t packed;
} hd_hit_record;

/* Output planes of a trace call; any pointer may be NULL to skip that plane. */
typedef struct hd_trace_outputs {
	uint32_t *rgba8;       /* shaded frame, trace.frag:395-401 converted to UNORM8 (r low byte, a = 255) */
	hd_hit_record *hits;   /* parity records */
	uint32_t *iters;       /* loop iterations, trace.frag:129 */
	uint32_t *fetches;     /* 32-bit node/colour words the reference algorithm reads for this ray (F of SURVEY §8d);
	                          requesting it selects an instrumented kernel variant — not for timed runs */
} hd_trace_outputs;

/* Screen-tile sharding of one frame over `world` GPUs (tiles of tile_w x tile_h px, tiles_x per row).
 * Ownership interleaves the ranks in BOTH directions so that cheap (sky) and expensive (silhouette) regions spread evenly:
 *   tiles_x % world != 0:  tile t = ty*tiles_x + tx belongs to rank t % world, local index t / world (row-major round
 *                          robin; the row length already shifts the pattern from row to row);
 *   tiles_x % world == 0:  tile (tx, ty) belongs to rank (tx + ty) % world, local index ty*(tiles_x/world) + tx/world
 *                          (plain round robin would give every rank the same columns in every row).
 * A rank's outputs are tile-major: index = local*tile_w*tile_h + ly*tile_w + lx.  hd_tile_shard_locate inverts the map. */
typedef struct hd_tile_shard {
	uint32_t tile_w, tile_h;
	uint32_t rank, world;
} hd_tile_shard;

/* One contiguous run of words changed since the last hd_dirty_reset — the analogue of
 * DAGNodePool::m_page_write_ranges (src/DAGNodePool.hpp:44-46,62-69). */
typedef struct hd_dirty_range {
	uint32_t word_offset;
	uint32_t word_count;
} hd_dirty_range;

typedef struct hd_pool hd_pool;

/* ---- library ---- */
const char *hd_version(void);
const char *hd_last_error(void);
int hd_device_count(void);

/* ---- config helpers: include/hashdag/Config.hpp ---- */
hd_status hd_config_from_default(const hd_default_config *dc, hd_config *out); /* Config.hpp:66-74 */
int hd_config_validate(const hd_config *cfg);                                  /* Config.hpp:48-56 */
uint32_t hd_config_total_buckets(const hd_config *cfg);                        /* Config.hpp:39-44 */
uint64_t hd_config_total_words(const hd_config *cfg);                          /* Config.hpp:46 */
uint32_t hd_config_level_base_bucket(const hd_config *cfg, uint32_t level);    /* Config.hpp:33-38 */

/* ---- pool: replaces DAGNodePool::Create / ctor (src/DAGNodePool.hpp:84-92, DAGNodePool.cpp:9-47) ---- */
hd_status hd_pool_create(const hd_config *cfg, int device, hd_pool **out);
void hd_pool_destroy(hd_pool *pool);
hd_status hd_pool_get_config(const hd_pool *pool, hd_config *out); /* NodePoolBase::GetConfig, NodePool.hpp:404 */
hd_status hd_pool_clear(hd_pool *pool);                            /* forget all nodes (fresh pool) */
/* root bookkeeping: DAGNodePool::SetRoot/GetRoot (src/DAGNodePool.hpp:96-97) */
hd_status hd_pool_set_root(hd_pool *pool, uint32_t root);
uint32_t hd_pool_get_root(const hd_pool *pool);
/* raw device pointers (for NCCL / torch interop); words = flat uint32 address space */
void *hd_pool_words_dev(hd_pool *pool);
void *hd_pool_bucket_words_dev(hd_pool *pool);
void *hd_pool_stream(hd_pool *pool); /* cudaStream_t all pool work is enqueued on */

/* host <-> device mirror interop: the ReadPage/WritePage callbacks (src/DAGNodePool.hpp:58-69) and
 * DAGNodePool::Flush (src/DAGNodePool.cpp:56-85). */
hd_status hd_pool_upload_words(hd_pool *pool, uint32_t word_offset, const uint32_t *src, uint32_t count);
hd_status hd_pool_read_words(hd_pool *pool, uint32_t word_offset, uint32_t *dst, uint32_t count);
/* One node as the host sees it after read_node/unpack (NodePool.hpp:264-317): inner = [mask, present children...]
 * (n_words = 1 + popcount), leaf = the two voxel words (n_words = 2). */
typedef struct hd_node_record {
	uint32_t ptr, level, n_words;
	uint32_t words[9];
} hd_node_record;
/* Breadth-first read-back of the subtree under `root` (a node of `level`), at most `depth` levels deep, in ONE device
 * pass and one copy — what a host-side visitor such as NodePoolBase::Iterate (test/test.cpp:36-52,173-198) needs instead
 * of one ReadPage round trip per node.  Records come level by level; a shared node appears once per parent that reaches it.
 * *n_out = records written (<= capacity); HD_ERR_OVERFLOW when the subtree has more nodes than `capacity` (the
 * records written are still valid, deeper nodes are missing). */
hd_status hd_pool_read_subtree(hd_pool *pool, uint32_t root, uint32_t level, uint32_t depth, hd_node_record *out,
                               uint32_t capacity, uint32_t *n_out);
hd_status hd_pool_upload_bucket_words(hd_pool *pool, uint32_t first_bucket, const uint32_t *src, uint32_t count);
hd_status hd_pool_read_bucket_words(hd_pool *pool, uint32_t first_bucket, uint32_t *dst, uint32_t count);
/* m_filled_node_pointers (NodePool.hpp:54,240-262); made on first use like make_filled_node_pointers() */
hd_status hd_pool_filled_nodes(hd_pool *pool, uint32_t *out_ptrs /* [node_levels] */);

/* ---- edit: replaces NodePoolBase::Edit (NodePool.hpp:405-417) and
 *      NodePoolThreadedEdit::ThreadedEdit (NodePoolThreadedEdit.hpp:104-126) ----
 * Applies edits[0..n) in index order (same final voxel set, hence the same canonical DAG, as n sequential
 * reference Edit calls) in ONE level-synchronous GPU pass.  stats may be NULL. */
hd_status hd_edit_batch(hd_pool *pool, uint32_t root_in, const hd_edit_desc *edits, uint32_t n,
                        uint32_t *root_out, hd_edit_stats *stats);
/* find-or-insert of explicit nodes: upsert_inner_node / upsert_leaf (NodePool.hpp:228-238), used by tests.
 * nodes = n packed nodes, each `words_each` words (2 for a leaf level). out_ptrs[n]. */
hd_status hd_upsert_nodes(hd_pool *pool, uint32_t level, const uint32_t *nodes, uint32_t words_each, uint32_t n,
                          uint32_t *out_ptrs);
/* Which path served the last hd_edit_batch call on this pool.  0 = the general level-synchronous path (any batch size).
 * 1 = the low-latency path: batches of <= 32 sphere/AABB editors — the interactive brush of src/main.cpp:214-238 — run as
 * ONE cooperative kernel (grid barriers between levels, small levels walked by one CTA, work-queue counts resident on the
 * device, parameters and counters through mapped host memory): one launch and one synchronisation per call.
 * 2 = the same phases as a pre-instantiated CUDA graph (HD_EDIT_FAST=2).  HD_EDIT_FAST=0 disables 1 and 2.
 * A fixed-size work queue that would overflow sends the call to the general path before anything is written.
 * Results are identical on every path. */
uint32_t hd_edit_last_path(const hd_pool *pool);

/* ---- colour pool: the DAGColorPool buffers the tracer reads (src/DAGColorPool.hpp:43-52; bindings 1,2 of
 *      shader/src/trace.frag:6-7). ---- */
hd_status hd_color_upload(hd_pool *pool, const uint32_t *color_nodes, uint64_t node_words, const uint32_t *color_leaves,
                          uint64_t leaf_words);

/* Colour pool state and colour-aware edits on the GPU (row N2).  hd_color_config tells the pool the colour leaf level
 * (DAGColorPool::Config::leaf_level, src/main.cpp:204-211) and the current colour root (DAGColorPool::SetRoot).
 * hd_edit_color replaces vbr_edit(...) of src/main.cpp:224-230 for the editors the reference ships:
 *   edit->kind == HD_EDIT_AABB_FILL                      AABBEditor with a colour        (main.cpp:47-69)
 *   edit->kind == HD_EDIT_SPHERE_FILL, paint == 0        SphereEditor<kFill>             (main.cpp:107-149)
 *   edit->kind == HD_EDIT_SPHERE_FILL, paint != 0        SphereEditor<kPaint> (geometry untouched)
 * It performs the geometry edit AND the colour update (VBREditorWrapper, include/hashdag/VBREditor.hpp:26-107; leaf
 * chunks as VBRChunkWriter would emit them, VBRColor.hpp:368-487; SetNode/FillNode/SetLeaf, src/DAGColorPool.hpp:147-204)
 * and returns both new roots.  rgb8 = 0xBBGGRR.  hd_color_read copies the buffers back (sizes from hd_color_sizes). */
hd_status hd_color_config(hd_pool *pool, uint32_t leaf_level, uint32_t color_root);
uint32_t hd_color_root(const hd_pool *pool);
uint32_t hd_color_leaf_level(const hd_pool *pool);
hd_status hd_color_sizes(hd_pool *pool, uint64_t *node_words, uint64_t *leaf_words);
hd_status hd_color_read(hd_pool *pool, uint32_t *nodes, uint64_t node_words, uint32_t *leaves, uint64_t leaf_words);
hd_status hd_edit_color(hd_pool *pool, uint32_t root_in, const hd_edit_desc *edit, uint32_t rgb8, uint32_t paint,
                        uint32_t *root_out, uint32_t *color_root_out, hd_edit_stats *stats);

/* ---- trace: replaces TracePass::CmdExecute + shader/src/trace.frag main() (TracePass.cpp:106-139) ----
 * Outputs are HOST pointers (copied back inside the call) for hd_trace / hd_trace_tiles and DEVICE pointers for
 * the *_dev variants, which only enqueue on hd_pool_stream() and do not synchronise. */
hd_status hd_trace(hd_pool *pool, const hd_trace_params *params, const hd_trace_outputs *host_out);
hd_status hd_trace_dev(hd_pool *pool, const hd_trace_params *params, const hd_trace_outputs *dev_out);
hd_status hd_trace_tiles(hd_pool *pool, const hd_trace_params *params, const hd_tile_shard *shard,
                         const hd_trace_outputs *host_out);
hd_status hd_trace_tiles_dev(hd_pool *pool, const hd_trace_params *params, const hd_tile_shard *shard,
                             const hd_trace_outputs *dev_out);
/* Beam optimisation (row N3): replaces BeamPass::CmdExecute + shader/src/beam.frag (src/rg/BeamPass.cpp:81-111) and
 * trace.frag compiled with BEAM_OPTIMIZATION (trace.frag:384-389).  beam_params is the beam pass's own block: same
 * camera, width/height = ceil(W/8) x ceil(H/8), proj_factor with the 8-pixel tolerance.  hd_beam_dev writes the
 * coarse start-t image (float, +inf = miss); hd_trace_with_beam* start every ray at 0.98 x the minimum of the 2x2
 * beam texels around the pixel.  host_beam (may be NULL) receives the beam image. */
hd_status hd_beam_dev(hd_pool *pool, const hd_trace_params *beam_params, float *beam_dev);
hd_status hd_trace_with_beam_dev(hd_pool *pool, const hd_trace_params *params, const float *beam_dev, uint32_t beam_w,
                                 uint32_t beam_h, const hd_trace_outputs *dev_out);
hd_status hd_trace_with_beam(hd_pool *pool, const hd_trace_params *params, const hd_trace_params *beam_params,
                             const hd_trace_outputs *host_out, float *host_beam);
/* Pipelined frames: submit enqueues the trace of one frame and the asynchronous copy of its shaded rgba8 plane into
 * host_rgba8 (pinned host memory for full overlap) on a copy stream; collect blocks until that frame has landed.
 * Two slots (0/1) may be in flight, so the read-back of frame k overlaps the trace of frame k+1 (a frame loop with
 * kFrameCount frames in flight, src/main.cpp:20,389-403).  shard may be NULL (full frame, row-major). */
hd_status hd_trace_submit(hd_pool *pool, const hd_trace_params *params, const hd_tile_shard *shard, uint32_t *host_rgba8,
                          uint32_t slot);
hd_status hd_trace_collect(hd_pool *pool, uint32_t slot);
/* pixels a rank owns under a shard (size of its output planes) */
uint64_t hd_tile_shard_pixels(const hd_trace_params *params, const hd_tile_shard *shard);
/* Tile coordinates of the rank's `local`-th tile (host-only helper for stitching / display); HD_ERR_INVALID past the end. */
hd_status hd_tile_shard_locate(const hd_trace_params *params, const hd_tile_shard *shard, uint32_t local, uint32_t *tile_x,
                               uint32_t *tile_y);
/* single pick ray: NodePoolTraversal::Traversal<float> (NodePoolTraversal.hpp:93-256), main.cpp:320-321.
 * Returns hit flag in *out_hit and the float hit position in out_pos[3]. */
hd_status hd_traverse_ray(hd_pool *pool, uint32_t root, const float o[3], const float d[3], int *out_hit,
                          float out_pos[3]);

/* ---- replica sync: replaces DAGNodePool::Flush for the multi-GPU case ----
 * Buckets are append-only, so the dirty set is [words at last sync, bucket_words) of every bucket.  The editing
 * rank packs it into one staging buffer (u32 words, 8-word header)
 *   [n_ranges][payload_words][root][flags][n_color_ranges][color_payload_words][color_root][color_leaf_level]
 *   n_ranges x {word_offset, word_count, payload_offset}  payload...  [colour section]
 * which the caller broadcasts (one NCCL broadcast over NVLink); replicas apply it with a scatter kernel that also
 * advances their bucket_words, and publish the root last.  flags: bit 0 = clear the replica first (after hd_gc), bit 1 =
 * a colour section follows the node payload (the DAGColorPool delta since the last hd_dirty_reset: appended nodes,
 * appended leaf chunks, chunks rewritten in place — replaces DAGColorPool::Flush, src/main.cpp:245-251), bit 2 = that
 * section is the whole colour pool.  The total size is a function of the header alone (replica.packed_bytes).
 * hd_dirty_apply_dev bounds-checks every range against the pool, its bucket and the payload before writing and returns
 * HD_ERR_INVALID (root not published) for a malformed blob. */
hd_status hd_dirty_count(hd_pool *pool, uint32_t *n_ranges, uint64_t *packed_bytes);
hd_status hd_dirty_ranges(hd_pool *pool, hd_dirty_range *out, uint32_t capacity, uint32_t *n_out);
hd_status hd_dirty_pack_dev(hd_pool *pool, void *staging_dev, uint64_t capacity_bytes, uint64_t *packed_bytes);
hd_status hd_dirty_apply_dev(hd_pool *pool, const void *staging_dev, uint64_t packed_bytes);
hd_status hd_dirty_reset(hd_pool *pool);

/* ---- garbage collection: replaces NodePoolThreadedGC::ThreadedGC (NodePoolThreadedGC.hpp:372-403; row N1) ----
 * Keeps exactly the nodes reachable from roots[0..n) plus the filled nodes, re-hashes them into compacted buckets and
 * returns the remapped roots (pointers change, the DAG does not).  The pool's published root follows when it is one
 * of `roots`.  Replicas must be re-synchronised afterwards (the next hd_dirty_pack_dev carries a clear-first flag). */
hd_status hd_gc(hd_pool *pool, const uint32_t *roots, uint32_t n_roots, uint32_t *new_roots, uint64_t *reachable_nodes);

/* ---- serialisation (no reference counterpart: it rebuilds its scene procedurally at start; row N4) ----
 * One file holds the config, the used prefix of every bucket, bucket_words, the root and the colour buffers. */
hd_status hd_pool_save(hd_pool *pool, const char *path);
hd_status hd_pool_load(const char *path, int device, hd_pool **out);

/* ---- pinned host buffers for frame read-back ----
 * hd_trace_submit copies straight into the caller's buffer; it should be page-locked.  write_combined = 1 allocates it
 * write-combined (cudaHostAllocWriteCombined): the DMA engine does not snoop the CPU caches, which raises the read-back
 * rate when several GPUs write into the same socket's memory at once, at the price of slow CPU READS of the buffer (fine
 * for a frame that goes on to a display / encoder, wrong for one the CPU post-processes). */
hd_status hd_host_alloc(uint64_t bytes, int write_combined, void **out);
hd_status hd_host_free(void *ptr);

/* ---- introspection used by tests / benches ---- */
hd_status hd_pool_used_words(hd_pool *pool, uint64_t *out); /* sum of bucket_words */
hd_status hd_sync(hd_pool *pool);
uint64_t hd_kernel_launches(void); /* kernels this library has launched so far (bench gpu_launches) */
/* Staged top levels of the trace kernel (the compact side table of the node levels next to the root; replaces the pool reads
 * of trace.frag:132,157-158 there): the root it was built for (HD_NULL_NODE when there is none), how many node levels and
 * how many nodes it holds.  HD_TRACE_TABLE = 0 / 1 / 2 (never, the default / once a root is traced twice / always) selects the policy. */
hd_status hd_trace_table_info(hd_pool *pool, uint32_t *root, uint32_t *levels, uint32_t *nodes);
/* Exhaustive check of the trace kernel's exact-arithmetic shortcuts against the IEEE operations they stand for
 * (trace.frag:83-88, 276-277, 340-356 are written with plain `/` and `sqrt`): the unchecked reciprocal over every float in
 * [2^-51, 2^51), the unchecked square root over every float in [2^-100, 2^100), x / D for every 0 <= x <= D, D in {255, 63, 31,
 * 7, 3}, and fma(h, c, t) == h * c + t for sampled power-of-two h.  *mismatches = number of differing results (0 = exact). */
hd_status hd_selftest_exact_arith(int device, uint64_t *mismatches);
/* The batched edit classifies the eight children of a node with shared per-axis terms (editors.cuh: edit_node8) where the
 * reference calls EditNode once per child (NodePool.hpp:345-360, editors src/main.cpp:35-46,77-106): n_cases pseudo-random editors
 * and nodes around the decision boundaries (touching boxes, near and far spheres, 32- and 64-bit paths, every level) through both;
 * *mismatches = nodes whose eight classifications differ (0 = identical). */
hd_status hd_selftest_edit_node8(int device, uint32_t n_cases, uint64_t *mismatches);

#ifdef __cplusplus
}
#endif
#endif /* HASHDAG_B200_H */
