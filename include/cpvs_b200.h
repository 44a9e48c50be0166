/* cpvs_b200 -- C ABI of the B200-native compact-precomputed-voxelized-shadow path.
 *
 * This header is the drop-in boundary: plain C, opaque handles, plain pointers and sizes. Each entry
 * point names the reference interface (TobiasRp/cpvs, path:line) it stands in for. The reference has
 * no FFI layer of its own -- its boundary is the C++ class API of cpvs_lib (CMakeLists.txt:36) -- so
 * the functions below are what a binding for that class API binds; the headers under include/cpvs/ put the
 * reference's class names back on top for C++ callers, and INTEGRATION.md shows the maintainer-side
 * change.
 *
 * Everything runs on one CUDA device per context (sm_100a). There is no CPU fallback: a call either
 * runs the CUDA path or returns an error code; cpvs_last_error() describes the failure.
 *
 * Conventions
 *   - depth maps: float32, square, side a power of two, row-major, one channel
 *     (Image<float>, src/Image.h:35-40).
 *   - `mem`: CPVS_MEM_HOST pointers are copied to / from the device inside the call;
 *     CPVS_MEM_DEVICE pointers are used in place on the context's stream.
 *   - all work is enqueued on the context's stream; calls that return sizes or host data
 *     synchronise that stream, the rest are asynchronous.
 *   - NodeVisibility values (src/CompressedShadow.h:22-26): 0 shadow, 1 visible (lit), 2 partial.
 */
#ifndef CPVS_B200_H
#define CPVS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CPVS_API __attribute__((visibility("default")))
#else
#define CPVS_API
#endif

#define CPVS_OK 0
#define CPVS_EINVAL 1    /* bad argument: not square / not a power of two / too small / null */
#define CPVS_ENOMEM 2    /* device or host allocation failed */
#define CPVS_ECUDA 3     /* CUDA runtime error, or no sm_100-class device */
#define CPVS_EOVERFLOW 4 /* a DAG would need more than 2^32 words (32-bit offsets are the format) */
#define CPVS_EINTERNAL 5 /* internal consistency check failed */

#define CPVS_MEM_HOST 0
#define CPVS_MEM_DEVICE 1

#define CPVS_SHADOW 0
#define CPVS_VISIBLE 1
#define CPVS_PARTIAL 2

#define CPVS_MAX_LEVELS 32

/* Phases of cpvs_shadow_create timed with CUDA events on the context's stream (cpvs_shadow_info.phase_ms).
 * Phases marked (1 kernel) bracket exactly one kernel launch. */
#define CPVS_PHASE_COUNT 0        /* closed-form node counts of the column + host read-back (absent when sizes are predicted or cached) */
#define CPVS_PHASE_EXPAND 1       /* breadth-first expansion of all inner levels */
#define CPVS_PHASE_LEAVES 2       /* leaf build (1 kernel) */
#define CPVS_PHASE_LEAF_TABLE 3   /* leaf level: distinct-count sketch read-out, table sizing + clear */
#define CPVS_PHASE_LEAF_INSERT 4  /* leaf level: hash-table insert (1 kernel) */
#define CPVS_PHASE_LEAF_RESOLVE 5 /* leaf level: rank scan (1 kernel) on a side stream, concurrent with phase 6 */
#define CPVS_PHASE_INNER_MERGE 6  /* inner levels: inserts on the main stream (rank scans beside them), join */
#define CPVS_PHASE_BASES 7        /* level bases (+ host read-back of sizes when the DAG's allocation was not predicted) */
#define CPVS_PHASE_EMIT_INNER 8   /* emission after the bases: inner levels (and the leaves, unless they were written during the merge) */
#define CPVS_PHASE_EMIT_LEAVES 9  /* compressed leaves (1 kernel); concurrent with phase 6 (predicted allocation) or inside phase 8 */
#define CPVS_NUM_PHASES 10

/* Grid sentinels written by CompressedShadowContainer::createTopLevelGrid
 * (src/CompressedShadowContainer.cpp:8-9). The lookup tests these same values. */
#define CPVS_GRID_CELL_SHADOWED 0x0FFFFFFFu
#define CPVS_GRID_CELL_VISIBLE 0x0FFFFFFEu

typedef struct cpvs_ctx cpvs_ctx;
typedef struct cpvs_minmax cpvs_minmax;
typedef struct cpvs_shadow cpvs_shadow;
typedef struct cpvs_container cpvs_container;

/* ---- context ------------------------------------------------------------------------------- */

/* One context per GPU; owns a stream and the stream-ordered scratch pool. */
CPVS_API int cpvs_ctx_create(int device, cpvs_ctx** out);
CPVS_API int cpvs_ctx_destroy(cpvs_ctx* ctx);
/* Run on a caller stream (cudaStream_t passed as void*); NULL restores the context's own stream. */
CPVS_API int cpvs_ctx_set_stream(cpvs_ctx* ctx, void* cuda_stream);
CPVS_API void* cpvs_ctx_get_stream(const cpvs_ctx* ctx);
CPVS_API int cpvs_ctx_synchronize(cpvs_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
CPVS_API uint64_t cpvs_ctx_launch_count(const cpvs_ctx* ctx);
/* Grows the device's stream-ordered memory pool by `bytes` now, so that the finished DAGs of the following
 * builds (kept alive by a tile-grid driver) are carved from memory the pool already owns instead of each
 * paying a fresh device allocation inside the build. Optional; HBM is 180 GB, a 4x4x4 grid of 16K^2 terrain
 * tiles keeps 3.8 GB of DAG words. */
CPVS_API int cpvs_ctx_reserve(cpvs_ctx* ctx, uint64_t bytes);
/* The opposite: releases everything the context (and the contexts it created for a grid worker) keeps for recycling -- pyramid
 * and DAG blocks, staging buffers -- and trims the pool. Handles stay valid; the next builds allocate again. Waits for the
 * context's streams. */
CPVS_API int cpvs_ctx_trim(cpvs_ctx* ctx);
/* cpvs_shadow_create sizes a build from the previous build of the same shape (side, z tile, leafmasks) on this context:
 * scratch, grids and the DAG's allocation come from those numbers plus head room (predicted + predicted >> headroom_shift),
 * so the build runs without asking the device for sizes first. Every kernel stays inside its capacities; a build that
 * outgrows them is redone with exact counts, so results never depend on the prediction. enabled = 0 always counts first.
 * Defaults: enabled, headroom_shift 3 (CPVS_PREDICT / CPVS_HEADROOM_SHIFT in the environment override them at creation). */
CPVS_API int cpvs_ctx_set_prediction(cpvs_ctx* ctx, int enabled, uint32_t headroom_shift);
typedef struct cpvs_ctx_stats {
	uint64_t predicted_builds;  /* builds started on predicted sizes */
	uint64_t exact_builds;      /* builds that counted first (incl. the rebuilds below) */
	uint64_t overflow_rebuilds; /* predicted builds that outgrew a node capacity and were redone */
	uint64_t reemissions;       /* builds whose predicted DAG allocation was too small: emitted again into an exact one */
} cpvs_ctx_stats;
CPVS_API int cpvs_ctx_get_stats(const cpvs_ctx* ctx, cpvs_ctx_stats* out);
/* Message for the last non-OK status on the calling thread. */
CPVS_API const char* cpvs_last_error(void);
CPVS_API const char* cpvs_version(void);

/* ---- MinMaxHierarchy (src/MinMaxHierarchy.h:23-72, src/MinMaxHierarchy.cpp:9-97) ------------- */

/* MinMaxHierarchy::MinMaxHierarchy(const ImageF&) (src/MinMaxHierarchy.cpp:9-27).
 * n: side of the depth map (power of two, >= 2). A CPVS_MEM_DEVICE depth map is borrowed, not
 * copied: it must stay alive and unchanged for as long as the hierarchy is used. */
CPVS_API int cpvs_minmax_build(cpvs_ctx* ctx, const float* depth, int n, int mem, cpvs_minmax** out);
/* The same for a hierarchy whose builds will be cut into z_tile_num z-slices (createShadowTiles,
 * src/DeferredRenderer.cpp:150-163): the construction pass also leaves the depth map re-encoded for the per-column leaf
 * builder of exactly that slicing (1 byte per texel). Purely a speed hint: cpvs_shadow_create accepts any z_tile_num on
 * any hierarchy and falls back to reading the depth map. cpvs_minmax_build is z_tile_num = 1; 0 prepares nothing. */
CPVS_API int cpvs_minmax_build_tiled(cpvs_ctx* ctx, const float* depth, int n, int mem, uint32_t z_tile_num, cpvs_minmax** out);
CPVS_API int cpvs_minmax_destroy(cpvs_minmax* mm);
/* getNumLevels() (src/MinMaxHierarchy.h:60-62): log2(n) + 1. */
CPVS_API int cpvs_minmax_num_levels(const cpvs_minmax* mm);
CPVS_API int cpvs_minmax_size(const cpvs_minmax* mm);
/* getLevel(level) (src/MinMaxHierarchy.h:67-72) copied to the host: level 0 is n*n depths, level k>=1
 * is (n>>k)^2 interleaved (min,max) pairs -- the layout getMin/getMax index (src/MinMaxHierarchy.h:31-55). */
CPVS_API int cpvs_minmax_level(const cpvs_minmax* mm, int level, float* out_host);
/* Device time of the build (CUDA events on the context's stream): whole call, and the fused base
 * kernel that produces levels 1..5 alone (0 when the map is too small for it). Synchronises. */
CPVS_API int cpvs_minmax_timing(const cpvs_minmax* mm, float* total_ms, float* base_kernel_ms);
/* cs::createChildmask(minMax, level, offset) (src/CompressedShadowUtil.cpp:20-54) for one node: the
 * 16-bit child mask (2 bits per child x | y<<1 | z<<2: 00 shadow, 01 lit, 10 partial) of the node whose
 * children are texels (x..x+1, y..y+1) of pyramid level `level` at depth slices z..z+1. For
 * known-answer tests; the builder classifies whole levels at once. */
CPVS_API int cpvs_minmax_childmask(const cpvs_minmax* mm, uint32_t level, uint32_t x, uint32_t y, uint32_t z, uint32_t z_tile_num,
		uint32_t* out);
/* Device pointer of a level (same layout), for zero-copy consumers. */
CPVS_API const float* cpvs_minmax_level_device(const cpvs_minmax* mm, int level);

/* ---- CompressedShadow (src/CompressedShadow.h:20-126) ---------------------------------------- */

typedef struct cpvs_shadow_info {
	uint32_t num_levels;       /* getNumLevels() (src/CompressedShadow.h:71) */
	uint32_t leafmasks;        /* 1 if level 2 holds 64-bit leafmasks (src/CompressedShadow.cpp:20-27) */
	uint32_t total_visibility; /* getTotalVisibility() (src/CompressedShadow.cpp:66-72) */
	uint32_t predicted;        /* 1: the build ran on sizes predicted from the previous build of the same shape (no count pass) */
	uint64_t words;                      /* getDAG().size() */
	uint64_t svo_nodes[CPVS_MAX_LEVELS]; /* nodes per level before merging, index = level */
	uint64_t dag_nodes[CPVS_MAX_LEVELS]; /* nodes per level after merging, index = level */
	uint64_t dag_words[CPVS_MAX_LEVELS]; /* compressed words per level */
	float build_ms;                      /* device time of the create call (CUDA events) */
	float phase_ms[CPVS_NUM_PHASES];     /* see CPVS_PHASE_* */
} cpvs_shadow_info;

/* CompressedShadow::create(const MinMaxHierarchy&, zTileIndex, zTileNum)
 * (src/CompressedShadow.h:48-49, src/CompressedShadow.cpp:49-59): SVO construction, common-subtree
 * merge and pointer compression, all on the device. The result is word-for-word the reference's
 * getDAG(). `leafmasks` selects the reference's compile-time LEAFMASKS switch
 * (src/CompressedShadow.cpp:17); as there, maps smaller than 16^2 never use leafmasks.
 * z_tile_num is passed by value (the reference keeps it in a file-static). */
CPVS_API int cpvs_shadow_create(cpvs_ctx* ctx, const cpvs_minmax* mm, uint32_t z_tile_index, uint32_t z_tile_num,
		int leafmasks, cpvs_shadow** out);
/* The same build without waiting for it: when the sizes can be predicted (cpvs_ctx_set_prediction; whole-volume builds
 * of a shape this context has built before) the call returns as soon as the kernels are enqueued, so a caller can
 * keep several builds in flight -- the next frame's on this context, or others on other contexts of the same GPU,
 * whose kernels then fill the gaps of this one's latency-bound phases. Otherwise it behaves like cpvs_shadow_create.
 * Every other call on the handle waits for the build first; cpvs_shadow_wait does only that and returns the build's
 * status. `mm` (and a borrowed device depth map) must stay alive until then. */
CPVS_API int cpvs_shadow_create_async(cpvs_ctx* ctx, const cpvs_minmax* mm, uint32_t z_tile_index, uint32_t z_tile_num,
		int leafmasks, cpvs_shadow** out);
CPVS_API int cpvs_shadow_wait(cpvs_shadow* s);
/* CompressedShadow::create(const ShadowMap*, ...) (src/CompressedShadow.cpp:61-64): temporary hierarchy. */
CPVS_API int cpvs_shadow_create_from_depth(cpvs_ctx* ctx, const float* depth, int n, int mem, uint32_t z_tile_index,
		uint32_t z_tile_num, int leafmasks, cpvs_shadow** out);
CPVS_API int cpvs_shadow_destroy(cpvs_shadow* s);
CPVS_API int cpvs_shadow_info_get(const cpvs_shadow* s, cpvs_shadow_info* info);
/* getDAG() (src/CompressedShadow.h:75) copied to the host; `out_host` holds info.words words. */
CPVS_API int cpvs_shadow_copy_dag(const cpvs_shadow* s, uint32_t* out_host);
CPVS_API const uint32_t* cpvs_shadow_dag_device(const cpvs_shadow* s);
/* CompressedShadow::traverse(vec3 ndc, bool tryLeafmasks) (src/CompressedShadow.cpp:404-463) for
 * `count` points: ndc = count x (x,y,z) floats in [-1,1]^3, out = count NodeVisibility bytes. Points
 * outside the cube are clamped to it. try_leafmasks must match how the DAG was built
 * (info.leafmasks): the mismatch is undefined behaviour in the reference and CPVS_EINVAL here. */
CPVS_API int cpvs_shadow_lookup_ndc(const cpvs_shadow* s, const float* ndc, int64_t count, int mem, int try_leafmasks,
		uint8_t* out);

/* ---- CompressedShadowContainer (src/CompressedShadowContainer.h:18-92) + shader/traverse.cs ----- */

/* CompressedShadowContainer(uint length) (src/CompressedShadowContainer.h:21-25): length^3 cells. */
CPVS_API int cpvs_container_create(cpvs_ctx* ctx, uint32_t length, cpvs_container** out);
CPVS_API int cpvs_container_destroy(cpvs_container* c);
/* set(unique_ptr<CompressedShadow>, x, y, z) (src/CompressedShadowContainer.h:34-39). The container
 * copies what it needs; the caller keeps ownership of `s` and may destroy it afterwards. */
CPVS_API int cpvs_container_set(cpvs_container* c, const cpvs_shadow* s, uint32_t x, uint32_t y, uint32_t z);
/* Same, for a cell built elsewhere (another GPU / rank): the finished DAG words are handed over. */
CPVS_API int cpvs_container_set_dag(cpvs_container* c, const uint32_t* words, uint64_t count, int mem, uint32_t num_levels,
		int leafmasks, uint32_t x, uint32_t y, uint32_t z);
/* copyToGPU() (src/CompressedShadowContainer.cpp:31-46): combineDAGs (:52-69) + createTopLevelGrid
 * (:71-91) into two device buffers. Every cell must have been set. */
CPVS_API int cpvs_container_finalize(cpvs_container* c);
CPVS_API int cpvs_container_info(const cpvs_container* c, uint64_t* dag_words, uint32_t* grid_cells, uint32_t* dag_levels,
		uint32_t* grid_levels);
/* Host copies of the combined DAG and of the grid (either pointer may be NULL). */
CPVS_API int cpvs_container_copy(const cpvs_container* c, uint32_t* dag_out_host, uint32_t* grid_out_host);
/* traverse.cs traverse() (shader/traverse.cs:75-133) on NDC points: path over the whole virtual
 * volume, grid cell pick, sentinels, DAG descent. out = count NodeVisibility bytes (0 or 1). */
CPVS_API int cpvs_container_lookup_ndc(const cpvs_container* c, const float* ndc, int64_t count, int mem, uint8_t* out);
/* evaluate(positionsWS, lightViewProj, visibilities) (src/CompressedShadowContainer.cpp:93-124,
 * shader/traverse.cs:135-149): positions = width*height rgba32f texels (xyz used), light_view_proj =
 * column-major mat4 (glm::value_ptr order), visibilities = width*height r8 texels (0 or 255). */
CPVS_API int cpvs_container_evaluate(const cpvs_container* c, const float* positions, uint32_t width, uint32_t height, int mem,
		const float light_view_proj[16], uint8_t* visibilities);
/* evaluate() on the textures themselves: CUDA surface objects (cudaSurfaceObject_t) over the arrays behind the rgba32f position
 * texture and the r8 visibility texture the reference binds to image units 0 and 1 (src/CompressedShadowContainer.cpp:100-107).
 * Asynchronous on the context's stream. With CUDA-GL interop the arrays are those of the renderer's own textures
 * (cudaGraphicsGLRegisterImage + cudaGraphicsSubResourceGetMappedArray): cpvs_container_evaluate_gl below does exactly that
 * and is compiled when the library is built with -DCPVS_WITH_GL (GL headers and a current GL context are needed). */
CPVS_API int cpvs_container_evaluate_surface(const cpvs_container* c, unsigned long long positions_surface,
		unsigned long long visibilities_surface, uint32_t width, uint32_t height, const float light_view_proj[16]);
#ifdef CPVS_WITH_GL
/* The drop-in for CompressedShadowContainer::evaluate(const Texture2D&, const mat4&, Texture2D*) in
 * DeferredRenderer::doAllShading (src/DeferredRenderer.cpp:320-347): GL texture names of the G-buffer's position texture
 * (rgba32f) and of the visibility texture (r8). */
CPVS_API int cpvs_container_evaluate_gl(const cpvs_container* c, unsigned int positions_texture, unsigned int visibilities_texture,
		uint32_t width, uint32_t height, const float light_view_proj[16]);
#endif
/* On-disk container (the reference has no serialisation and rebuilds on every launch).
 * File = 64-byte header {"CPVSDAG2", version, length, dag_levels, grid_levels, leafmasks, dag_words,
 * grid_cells, fnv64 over header fields + grid + DAG words} + grid words + DAG words, little endian. Loading validates
 * the header fields against each other and the file size, the checksum, and every grid entry against the DAG's extent.
 * A loaded container is finalized and ready for lookups; its cells cannot be re-set (CPVS_EINVAL). */
CPVS_API int cpvs_container_save(const cpvs_container* c, const char* path);
CPVS_API int cpvs_container_load(cpvs_ctx* ctx, const char* path, cpvs_container** out);
/* setFilterSize (src/CompressedShadowContainer.h:71-73). The reference plumbs the value into its shader
 * (`uniform int filterSize`, shader/traverse.cs:16-17) and never uses it; here evaluate() filters: percentage-closer
 * filtering over size x size voxels of the pixel's depth slice, centred on its voxel, clamped to the volume. 1 (the
 * default) is the reference's single lookup and writes 0 / 255; larger sizes write round(255 * lit / taps). size <= 64. */
CPVS_API int cpvs_container_set_filter_size(cpvs_container* c, uint32_t size);

/* A container put together from cells that already sit in device memory -- possibly on other GPUs of the box (the words
 * are fetched with cudaMemcpyPeerAsync): combineDAGs + createTopLevelGrid (src/CompressedShadowContainer.cpp:52-91) for a
 * tile grid whose cells were built by several workers. cells: length^3 entries in container order
 * ((z * length + y) * length + x); a cell of one word may leave words_device NULL (the word is its root mask). The result
 * is finalized; its cells cannot be re-set. */
typedef struct cpvs_cell_part {
	uint64_t words;
	uint32_t root_mask;           /* first word of the cell's DAG: 0x0000 all shadow, 0x5555 all lit, else partial */
	int32_t device;               /* CUDA device of words_device */
	const uint32_t* words_device;
} cpvs_cell_part;
CPVS_API int cpvs_container_assemble(cpvs_ctx* ctx, uint32_t length, uint32_t num_levels, int leafmasks, const cpvs_cell_part* cells,
		cpvs_container** out);

/* ---- tile grids over the GPUs of one box (src/DeferredRenderer.cpp:150-235) --------------------------
 * renderWithTiles cuts the light frustum into length x length depth tiles, createShadowTiles builds `length` z-slice DAGs
 * from each tile's hierarchy, precomputeShadows moves the container to the GPU. Here the xy tiles are sharded over GPUs:
 * nothing crosses GPUs while building, the host gathers the cells' sizes and runs createTopLevelGrid's scan, and the
 * finished words are replicated with peer copies for the lookups. No NCCL.
 *
 * cpvs_grid_build does all of it inside one process, one host thread per GPU. The worker calls underneath are exported
 * for callers that run one process per GPU (torchrun): each process drives its own worker and exchanges the few bytes of
 * costs and sizes through whatever host-side channel it has. */
typedef struct cpvs_grid_desc {
	uint32_t length;    /* cells per axis, a power of two: length^2 xy tiles, `length` z-slices each */
	int32_t tile;       /* side of one depth tile (power of two) */
	int32_t leafmasks;
	int32_t scene;      /* CPVS_SCENE_*: tiles are generated on the owning GPU (cpvs_depth_generate); -1: fetch */
	int (*fetch)(void* user, uint32_t x, uint32_t y, float* host_out); /* scene < 0: writes tile (x, y), tile*tile floats, returns 0 */
	void* user;
} cpvs_grid_desc;
typedef struct cpvs_grid_cell {
	uint32_t index;      /* cell index in the container: (z * length + y) * length + x */
	uint32_t num_levels;
	uint64_t words;
	uint32_t root_mask;
	int32_t device;
	const uint32_t* words_device;
	uint64_t svo_nodes, dag_nodes;
} cpvs_grid_cell;
typedef struct cpvs_grid_worker cpvs_grid_worker;
CPVS_API int cpvs_grid_worker_create(cpvs_ctx* ctx, const cpvs_grid_desc* desc, cpvs_grid_worker** out);
CPVS_API int cpvs_grid_worker_destroy(cpvs_grid_worker* w);
/* Cost of `count` xy tiles (pairs x, y) for the ownership: the SVO nodes of all their cells from the closed-form count --
 * over a hierarchy of the tile resampled at an eighth of its resolution when the scene is generated on the device (nothing
 * stays resident), else over the tile's own hierarchy, which then stays resident for cpvs_grid_worker_build (_release
 * drops tiles given to another worker). */
CPVS_API int cpvs_grid_worker_estimate(cpvs_grid_worker* w, const uint32_t* xy, int count, uint64_t* cost_out);
CPVS_API int cpvs_grid_worker_release(cpvs_grid_worker* w, const uint32_t* xy, int count);
/* createShadowTiles (src/DeferredRenderer.cpp:150-163) for `count` xy tiles: hierarchy + one DAG per z-slice, kept on the GPU. */
CPVS_API int cpvs_grid_worker_build(cpvs_grid_worker* w, const uint32_t* xy, int count);
/* The same for tiles handed out one by one (a shared queue): `next` fills (x, y) and returns 1, or returns 0 when there are no
 * more. It is called one tile ahead: the next tile's depth, pyramid and node counts are enqueued before the current tile's
 * slices, so that the host never waits for the GPU between tiles. */
typedef int (*cpvs_next_tile_fn)(void* user, uint32_t* x, uint32_t* y);
CPVS_API int cpvs_grid_worker_build_from(cpvs_grid_worker* w, cpvs_next_tile_fn next, void* user);
CPVS_API int cpvs_grid_worker_num_cells(const cpvs_grid_worker* w);
/* The finished cells (returns their number, < 0 on error). */
CPVS_API int cpvs_grid_worker_cells(const cpvs_grid_worker* w, cpvs_grid_cell* out, int capacity);
/* Device time (CUDA events on the worker's stream) of everything _estimate and _build did so far, from depth tiles resident
 * in device memory; producing the depth tiles (generator or callback + copy) is clocked separately. */
CPVS_API float cpvs_grid_worker_device_ms(const cpvs_grid_worker* w);
CPVS_API float cpvs_grid_worker_depth_ms(const cpvs_grid_worker* w);
/* One process per GPU: packs the finished cells into one block of plain device memory and returns its CUDA IPC handle
 * (64 bytes) plus, per cell in the order of cpvs_grid_worker_cells, the first word inside the block (returns the number of
 * cells, < 0 on error). Another process maps the block with cpvs_ipc_open (on its own device: the words are then fetched
 * peer to peer) and feeds cpvs_container_assemble; the block lives until the worker is destroyed. */
CPVS_API int cpvs_grid_worker_export(cpvs_grid_worker* w, unsigned char handle[64], uint64_t* offsets, int capacity);
CPVS_API int cpvs_ipc_open(const unsigned char handle[64], int device, void** out);
/* The same cells copied to host memory instead (out_host == NULL: only offsets and the count): the cheaper medium when
 * eight processes would otherwise each have to set up peer access to seven others. Returns the number of cells. */
CPVS_API int cpvs_grid_worker_copy_cells(const cpvs_grid_worker* w, uint32_t* out_host, uint64_t capacity_words, uint64_t* offsets, int capacity);
CPVS_API int cpvs_ipc_close(int device, void* ptr);
/* Ownership by cost, longest tile first to the least loaded worker; owner_in (may be NULL) breaks ties in favour of the
 * worker that already holds the tile's hierarchy. Pure host code. */
CPVS_API int cpvs_grid_assign(const uint64_t* cost, int num_tiles, int num_workers, const int* owner_in, int* owner_out);

#define CPVS_GRID_MAX_DEVICES 16
typedef struct cpvs_grid_stats {
	uint32_t devices, cells, one_word_cells, moved_tiles;
	uint64_t dag_words, svo_nodes, dag_nodes, launches;
	float build_ms_max;                       /* device time of the slowest GPU (estimates + builds, from resident depth tiles) */
	float build_ms[CPVS_GRID_MAX_DEVICES];    /* per GPU */
	float depth_ms[CPVS_GRID_MAX_DEVICES];    /* per GPU: producing the depth tiles (not part of the build metric) */
	uint32_t tiles[CPVS_GRID_MAX_DEVICES];    /* xy tiles built per GPU */
	float build_wall_ms, gather_ms, replicate_ms, wall_ms; /* host clock: builds on all GPUs; gather of sizes; replication; all */
} cpvs_grid_stats;
typedef struct cpvs_grid cpvs_grid;
/* The whole grid on `devices` (1..16 GPUs of this box). replicate != 0: every GPU gets a copy of the container for the
 * lookups, else only devices[0]. */
CPVS_API int cpvs_grid_build(const int* devices, int num_devices, const cpvs_grid_desc* desc, int replicate, cpvs_grid** out);
CPVS_API int cpvs_grid_destroy(cpvs_grid* g);
CPVS_API int cpvs_grid_stats_get(const cpvs_grid* g, cpvs_grid_stats* out);
/* The container on devices[index] (borrowed; index 0 always exists). */
CPVS_API cpvs_container* cpvs_grid_container(const cpvs_grid* g, int index);
/* traverse.cs over the grid for host points: the batch is split by rows over the GPUs holding a replica. */
CPVS_API int cpvs_grid_lookup_ndc(const cpvs_grid* g, const float* ndc_host, int64_t count, uint8_t* out_host);

/* ---- device-resident depth source ---------------------------------------------------------
 * Replaces the reference's render + glGetTexImage read-back of one light-frustum tile
 * (src/ShadowMap.cpp:23-30, src/DeferredRenderer.cpp:173-177) for the synthetic scenes of SURVEY.md 8d:
 * writes the n x n window (tile_x, tile_y) of the (n * tiles_per_side)^2 virtual map into device memory on
 * the context's stream, byte-identical to the host generator (cpvs_b200/synth). kind: CPVS_SCENE_PLANE,
 * CPVS_SCENE_CITY or CPVS_SCENE_TERRAIN_DEV (SURVEY.md's terrain on a deterministic sin/cos shared by host and device;
 * the libm terrain itself, scene 1, has no device twin: CPVS_EINVAL). n: multiple of 4. */
#define CPVS_SCENE_PLANE 0
#define CPVS_SCENE_CITY 2
#define CPVS_SCENE_TERRAIN_DEV 3
CPVS_API int cpvs_depth_generate(cpvs_ctx* ctx, int kind, int n, int tile_x, int tile_y, int tiles_per_side, float* depth_device);

#ifdef __cplusplus
}
#endif
#endif /* CPVS_B200_H */
