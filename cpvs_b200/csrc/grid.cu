// Tile-grid driver in C++: the caller contract of the reference's DeferredRenderer::renderWithTiles / createShadowTiles /
// precomputeShadows (reference src/DeferredRenderer.cpp:150-235) spread over the GPUs of one box.
//
// The virtual shadow map is a length x length grid of depth tiles; every xy tile yields `length` z-slice DAGs from one
// pyramid, and the cells are independent. A *worker* (one per GPU, one host thread each) produces the depth
// tiles it owns in device memory, builds their pyramids and cells, and keeps the finished DAGs in its GPU's memory. Nothing
// crosses GPUs while building: the host gathers (words, root mask) per cell, runs createTopLevelGrid's exclusive scan
// (reference src/CompressedShadowContainer.cpp:71-91), and the words are then replicated with cudaMemcpyPeerAsync into one
// container per GPU for the lookups, whose batches are split by rows. No NCCL anywhere.
//
// Ownership follows the cost of the tiles, not their position: with few tiles per GPU (configs[2]: 16 tiles on 8 GPUs) every
// worker first builds the pyramids of a round-robin share and counts their nodes (closed form, one launch per tile), the host
// assigns tiles longest-first to the least loaded worker, and a tile that changes hands is produced again by its new
// owner; with many tiles per GPU (configs[4]: 256 tiles) the workers simply pull the next tile from a shared queue.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <new>
#include <thread>

#include "handles.h"

using namespace cpvs;

namespace cpvs {
int containerFromParts(cpvs_ctx* ctx, u32 length, u32 numLevels, int leafmasks, const cpvs_cell_part* parts, cpvs_container** out);  // capi.cu
}  // namespace cpvs

namespace {
struct WorkerTile {
	u32 x = 0, y = 0;
	float* depth = nullptr;  // device, owned; released once the tile is built
	cpvs_minmax* mm = nullptr;
	u64 cost = 0;
	bool built = false;
	bool queued = false;  // taken by a build call that is still working on it
	cudaEvent_t evDepth0 = nullptr, evDepth1 = nullptr;  // around the production of the depth tile, which runs alone on the GPU
	std::vector<cpvs_shadow*> cells;  // z = 0 .. length-1
};
double nowMs() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

struct cpvs_grid_worker {
	cpvs_ctx* ctx = nullptr;
	// More contexts of the same GPU: the z-slices of a tile are independent builds from one hierarchy, so those that are
	// real builds take turns on the contexts and run side by side (the kernels of one fill the latency-bound phases of the
	// others). lanes[0] == ctx.
	std::vector<cpvs_ctx*> lanes;
	std::vector<cudaEvent_t> evLane;  // [k]: the slices enqueued on lane k so far are done
	std::vector<bool> laneUsed;
	u32 nextLane = 0;
	cudaEvent_t depthFence = nullptr;  // the latest depth tile is done (an evDepth1); the lanes wait for it before their slices
	cudaEvent_t evStage = nullptr;     // the copy out of hostStage is done
	cudaEvent_t evJoin = nullptr;
	float depthMs = 0.f;  // device time spent producing depth tiles (not part of the build metric: it starts from resident depth)
	float depthInSpanMs = 0.f;  // the part of it inside the build call in progress
	cpvs_grid_desc desc{};
	std::vector<WorkerTile> tiles;
	float* hostStage = nullptr;  // pinned, one tile (fetch callback)
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	float deviceMs = 0.f;  // device time of everything the worker enqueued for estimates and builds
	u32* exported = nullptr;  // cpvs_grid_worker_export: all cells in one cudaMalloc'ed block other processes can map
	u32 built = 0;
	int status = CPVS_OK;
	std::string error;
};

namespace {

WorkerTile* findTile(cpvs_grid_worker* w, u32 x, u32 y) {
	for (WorkerTile& t : w->tiles)
		if (t.x == x && t.y == y) return &t;
	return nullptr;
}

void releaseInputs(cpvs_grid_worker* w, WorkerTile& t) {
	if (t.mm) cpvs_minmax_destroy(t.mm);
	t.mm = nullptr;
	ctxFree(w->ctx, t.depth);
	t.depth = nullptr;
	if (t.evDepth0) {
		if (w->depthFence == t.evDepth1) w->depthFence = nullptr;
		cudaEventDestroy(t.evDepth0);
		cudaEventDestroy(t.evDepth1);
		t.evDepth0 = t.evDepth1 = nullptr;
	}
}

// Depth tile in device memory (generated there, or fetched from the caller and copied).
int produceDepth(cpvs_grid_worker* w, WorkerTile& t) {
	if (t.depth) return CPVS_OK;
	cpvs_ctx* ctx = w->ctx;
	const cpvs_grid_desc& d = w->desc;
	const size_t texels = (size_t)d.tile * d.tile;
	CPVS_CUDA(cudaSetDevice(ctx->device));
	CPVS_CUDA(ctxAlloc(ctx, reinterpret_cast<void**>(&t.depth), texels * sizeof(float)));
	if (d.scene >= 0) return cpvs_depth_generate(ctx, d.scene, d.tile, (int)t.x, (int)t.y, (int)d.length, t.depth);
	if (!d.fetch) return fail(CPVS_EINVAL, "cpvs_grid: neither a scene nor a fetch callback");
	if (!w->hostStage) {
		CPVS_CUDA(cudaMallocHost(reinterpret_cast<void**>(&w->hostStage), texels * sizeof(float)));
		CPVS_CUDA(cudaEventCreateWithFlags(&w->evStage, cudaEventDisableTiming));
	} else {
		CPVS_CUDA(cudaEventSynchronize(w->evStage));  // the previous tile's copy out of the staging buffer
	}
	if (d.fetch(d.user, t.x, t.y, w->hostStage) != 0) return fail(CPVS_EINVAL, "cpvs_grid: fetch callback failed for tile (%u,%u)", t.x, t.y);
	CPVS_CUDA(cudaMemcpyAsync(t.depth, w->hostStage, texels * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
	CPVS_CUDA(cudaEventRecord(w->evStage, ctx->stream));
	return CPVS_OK;
}

// + its pyramid, prepared for `length` z-slices.
int produceTile(cpvs_grid_worker* w, WorkerTile& t) {
	if (t.mm) return CPVS_OK;
	if (int rc = produceDepth(w, t)) return rc;
	return cpvs_minmax_build_tiled(w->ctx, t.depth, w->desc.tile, CPVS_MEM_DEVICE, w->desc.length, &t.mm);
}

struct DeviceTimer {  // accumulates device time between construction and destruction into the worker
	cpvs_grid_worker* w;
	explicit DeviceTimer(cpvs_grid_worker* worker) : w(worker) { cudaEventRecord(w->ev0, w->ctx->stream); }
	~DeviceTimer() {
		cudaEventRecord(w->ev1, w->ctx->stream);
		float ms = 0.f;
		if (cudaEventSynchronize(w->ev1) == cudaSuccess && cudaEventElapsedTime(&ms, w->ev0, w->ev1) == cudaSuccess) w->deviceMs += ms;
	}
};

}  // namespace

extern "C" {

int cpvs_grid_worker_create(cpvs_ctx* ctx, const cpvs_grid_desc* desc, cpvs_grid_worker** out) {
	if (!ctx || !desc || !out) return fail(CPVS_EINVAL, "cpvs_grid_worker_create: NULL argument");
	*out = nullptr;
	if (!isPow2(desc->length) || desc->length > 64) return fail(CPVS_EINVAL, "cpvs_grid: length %u must be a power of two <= 64", desc->length);
	if (desc->tile < 16 || !isPow2((u64)desc->tile) || (u64)desc->tile * desc->length > (1ull << 23))
		return fail(CPVS_EINVAL, "cpvs_grid: tile side %d with %u slices", desc->tile, desc->length);
	cpvs_grid_worker* w = new (std::nothrow) cpvs_grid_worker();
	if (!w) return fail(CPVS_ENOMEM, "cpvs_grid_worker_create: host allocation");
	w->ctx = ctx;
	w->desc = *desc;
	cudaSetDevice(ctx->device);
	if (cudaEventCreate(&w->ev0) != cudaSuccess || cudaEventCreate(&w->ev1) != cudaSuccess ||
			cudaEventCreateWithFlags(&w->evJoin, cudaEventDisableTiming) != cudaSuccess) {
		cpvs_grid_worker_destroy(w);
		return fail(CPVS_ECUDA, "cpvs_grid_worker_create: events");
	}
	// (measured on the 256K^2 city, 4096 cells, one B200: 505 ms with two lanes, 473 with three, 461 with four, 451 with six; the
	// 64K^2 terrain, whose tiles have two real slices, does not care)
	int numLanes = 4;
	if (const char* e = getenv("CPVS_GRID_LANES")) numLanes = atoi(e) < 1 ? 1 : (atoi(e) > 8 ? 8 : atoi(e));
	w->lanes.push_back(ctx);
	while ((int)w->lanes.size() < numLanes) {
		cpvs_ctx* next = siblingContext(w->lanes.back());  // kept by the context: warm for the next worker
		if (!next) {
			cpvs_grid_worker_destroy(w);
			return fail(CPVS_ECUDA, "cpvs_grid_worker_create: context %d: %s", (int)w->lanes.size(), cpvs_last_error());
		}
		w->lanes.push_back(next);
	}
	w->evLane.assign(w->lanes.size(), nullptr);
	w->laneUsed.assign(w->lanes.size(), false);
	for (cudaEvent_t& e : w->evLane)
		if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
			cpvs_grid_worker_destroy(w);
			return fail(CPVS_ECUDA, "cpvs_grid_worker_create: events");
		}
	*out = w;
	return CPVS_OK;
}

int cpvs_grid_worker_destroy(cpvs_grid_worker* w) {
	if (!w) return CPVS_OK;
	cudaSetDevice(w->ctx->device);
	for (WorkerTile& t : w->tiles) {
		releaseInputs(w, t);
		for (cpvs_shadow* s : t.cells) cpvs_shadow_destroy(s);
	}
	if (w->hostStage) cudaFreeHost(w->hostStage);
	if (w->exported) cudaFree(w->exported);
	if (w->evJoin) cudaEventDestroy(w->evJoin);
	if (w->evStage) cudaEventDestroy(w->evStage);
	for (cudaEvent_t e : w->evLane)
		if (e) cudaEventDestroy(e);
	if (w->ev0) cudaEventDestroy(w->ev0);
	if (w->ev1) cudaEventDestroy(w->ev1);
	delete w;
	return CPVS_OK;
}

// Cost of tiles, for the ownership. Scenes generated on the device are sampled at an eighth of the tile's resolution (same
// window, same slicing: its leaves are the tile's level-5 nodes) and the nodes of all z-slices of that small hierarchy are
// counted -- a proxy that costs a few launches and keeps nothing resident. Tiles handed over by a callback cannot be
// resampled: their own hierarchy is built and counted, and stays resident for cpvs_grid_worker_build. xy: count pairs (x, y).
int cpvs_grid_worker_estimate(cpvs_grid_worker* w, const uint32_t* xy, int count, uint64_t* costOut) {
	if (!w || (count > 0 && (!xy || !costOut))) return fail(CPVS_EINVAL, "cpvs_grid_worker_estimate: NULL argument");
	DeviceTimer timer(w);
	cpvs_ctx* ctx = w->ctx;
	const cpvs_grid_desc& d = w->desc;
	const bool proxy = d.scene >= 0 && d.tile >= 1024;
	const int side = proxy ? d.tile / 8 : d.tile;
	int L = 1;
	while ((1 << (L - 1)) < side) ++L;
	const bool useLeaf = d.leafmasks && (L - 3) >= 2;
	const int minLevel = useLeaf ? 2 : 0;
	float* small = nullptr;
	if (proxy) CPVS_CUDA(ctxAlloc(ctx, reinterpret_cast<void**>(&small), (size_t)side * side * sizeof(float)));
	int rc = CPVS_OK;
	for (int i = 0; i < count && rc == CPVS_OK; ++i) {
		cpvs_minmax* mm = nullptr;
		WorkerTile* t = nullptr;
		if (proxy) {
			rc = cpvs_depth_generate(ctx, d.scene, side, (int)xy[2 * i], (int)xy[2 * i + 1], (int)d.length, small);
			if (rc == CPVS_OK) rc = cpvs_minmax_build_tiled(ctx, small, side, CPVS_MEM_DEVICE, 0, &mm);
		} else {
			t = findTile(w, xy[2 * i], xy[2 * i + 1]);
			if (!t) {
				w->tiles.emplace_back();
				t = &w->tiles.back();
				t->x = xy[2 * i];
				t->y = xy[2 * i + 1];
			}
			rc = produceTile(w, *t);
			mm = t->mm;
		}
		if (rc == CPVS_OK && !useLeaf) rc = ensureLowLevels(mm, 1);
		const u64* counts = nullptr;
		if (rc == CPVS_OK) rc = columnCountsOf(ctx, mm, d.length, minLevel, &counts);
		if (rc == CPVS_OK) {
			u64 cost = 0;
			for (u32 z = 0; z < d.length; ++z) {
				const u64* c = counts + (size_t)z * kMaxLevels;
				if (c[L - 3] == 0) {  // a one-word cell
					cost += 1;
					continue;
				}
				for (int l = minLevel; l <= L - 3; ++l) cost += c[l] * (useLeaf && l == 2 ? 3u : 2u);  // leaves weigh more: built, hashed, expanded
			}
			costOut[i] = cost + ((u64)side * side >> 4);  // + the pyramid pass over the tile
			if (t) t->cost = costOut[i];
		}
		if (proxy && mm) cpvs_minmax_destroy(mm);
	}
	if (small) ctxFree(ctx, small);
	return rc;
}

// Drops tiles that were estimated here but went to another worker.
int cpvs_grid_worker_release(cpvs_grid_worker* w, const uint32_t* xy, int count) {
	if (!w) return fail(CPVS_EINVAL, "cpvs_grid_worker_release: NULL argument");
	cudaSetDevice(w->ctx->device);
	for (int i = 0; i < count; ++i) {
		WorkerTile* t = findTile(w, xy[2 * i], xy[2 * i + 1]);
		if (t && !t->built) {
			releaseInputs(w, *t);
			w->tiles.erase(w->tiles.begin() + (t - w->tiles.data()));
		}
	}
	return CPVS_OK;
}

}  // extern "C"

namespace {

int tileMinLevel(const cpvs_grid_worker* w) {
	int L = 1;
	while ((1 << (L - 1)) < w->desc.tile) ++L;
	return w->desc.leafmasks && (L - 3) >= 2 ? 2 : 0;
}

size_t tileIndex(cpvs_grid_worker* w, u32 x, u32 y) {
	if (WorkerTile* t = findTile(w, x, y)) return (size_t)(t - w->tiles.data());
	w->tiles.emplace_back();
	w->tiles.back().x = x;
	w->tiles.back().y = y;
	return w->tiles.size() - 1;
}

// Enqueues, on the worker's own stream, everything a tile's slices start from: the depth tile (alone on the GPU, between two
// events: its time is not the build's), the pyramid, and the node counts of the tile's column with their read-back.
int prepareTile(cpvs_grid_worker* w, WorkerTile& t) {
	cpvs_ctx* ctx = w->ctx;
	if (!t.depth) {
		for (size_t k = 1; k < w->lanes.size(); ++k)
			if (w->laneUsed[k]) CPVS_CUDA(cudaStreamWaitEvent(ctx->stream, w->evLane[k], 0));
		CPVS_CUDA(cudaEventCreate(&t.evDepth0));
		CPVS_CUDA(cudaEventCreate(&t.evDepth1));
		CPVS_CUDA(cudaEventRecord(t.evDepth0, ctx->stream));
		if (int rc = produceDepth(w, t)) return rc;
		CPVS_CUDA(cudaEventRecord(t.evDepth1, ctx->stream));
		w->depthFence = t.evDepth1;
	}
	if (int rc = produceTile(w, t)) return rc;
	return columnCountsBegin(ctx, t.mm, w->desc.length, tileMinLevel(w));
}

// One DAG per z-slice; the slices that are real builds take turns on the lanes.
int enqueueSlices(cpvs_grid_worker* w, WorkerTile& t) {
	const u32 len = w->desc.length;
	t.cells.assign(len, nullptr);
	if (w->depthFence)
		for (size_t k = 1; k < w->lanes.size(); ++k) CPVS_CUDA(cudaStreamWaitEvent(w->lanes[k]->stream, w->depthFence, 0));
	for (u32 z = 0; z < len; ++z) {
		if (int rc = cpvs_shadow_create_async(w->lanes[w->nextLane], t.mm, z, len, w->desc.leafmasks, &t.cells[z])) return rc;
		if (t.cells[z]->pending) {  // (one-word slices cost nothing)
			w->laneUsed[w->nextLane] = true;
			w->nextLane = (w->nextLane + 1) % (u32)w->lanes.size();
		}
	}
	for (size_t k = 1; k < w->lanes.size(); ++k)
		if (w->laneUsed[k]) CPVS_CUDA(cudaEventRecord(w->evLane[k], w->lanes[k]->stream));
	return CPVS_OK;
}

int finishTile(cpvs_grid_worker* w, WorkerTile& t) {
	for (cpvs_shadow* s : t.cells)
		if (int rc = shadowWaitBegin(s)) return rc;
	for (cpvs_shadow* s : t.cells)
		if (int rc = cpvs_shadow_wait(s)) return rc;
	if (t.evDepth0) {
		float ms = 0.f;
		CPVS_CUDA(cudaEventSynchronize(t.evDepth1));
		CPVS_CUDA(cudaEventElapsedTime(&ms, t.evDepth0, t.evDepth1));
		w->depthMs += ms;
		w->depthInSpanMs += ms;
	}
	releaseInputs(w, t);
	t.built = true;
	t.queued = false;
	++w->built;
	return CPVS_OK;
}

struct TileList {
	const uint32_t* xy;
	int count, at;
};
int nextFromList(void* user, uint32_t* x, uint32_t* y) {
	TileList* l = static_cast<TileList*>(user);
	if (l->at >= l->count) return 0;
	*x = l->xy[2 * l->at];
	*y = l->xy[2 * l->at + 1];
	++l->at;
	return 1;
}

}  // namespace

extern "C" {

// createShadowTiles for the xy tiles `next` hands out (returns 0 when there are no more): depth tile (unless still resident from
// the estimate), pyramid, one DAG per z-slice. Two tiles are in flight: the next tile is asked for, and its depth, pyramid and
// counts are enqueued, before the slices of the current one, so the host never waits for the counts and the GPU never for the host.
int cpvs_grid_worker_build_from(cpvs_grid_worker* w, cpvs_next_tile_fn next, void* user) {
	if (!w || !next) return fail(CPVS_EINVAL, "cpvs_grid_worker_build_from: NULL argument");
	cpvs_ctx* ctx = w->ctx;
	CPVS_CUDA(cudaSetDevice(ctx->device));
	const size_t none = ~(size_t)0;
	auto take = [&](size_t* idx) {
		*idx = none;
		uint32_t x = 0, y = 0;
		while (next(user, &x, &y)) {
			if (x >= w->desc.length || y >= w->desc.length) return fail(CPVS_EINVAL, "cpvs_grid_worker_build: tile (%u,%u) of %u", x, y, w->desc.length);
			const size_t i = tileIndex(w, x, y);
			if (w->tiles[i].built || w->tiles[i].queued) continue;  // (a tile named twice is built once)
			w->tiles[i].queued = true;
			*idx = i;
			break;
		}
		return (int)CPVS_OK;
	};
	size_t prev = none, cur = none, nxt = none;
	if (int rc = take(&cur)) return rc;
	if (cur == none) return CPVS_OK;
	w->depthInSpanMs = 0.f;
	const bool trace = getenv("CPVS_TRACE") != nullptr;  // host time per step of the loop
	double tPrepare = 0, tEnqueue = 0, tFinish = 0, t0 = trace ? nowMs() : 0;
	auto lap = [&](double& into) {
		if (!trace) return;
		const double t = nowMs();
		into += t - t0;
		t0 = t;
	};
	CPVS_CUDA(cudaEventRecord(w->ev0, ctx->stream));
	if (int rc = prepareTile(w, w->tiles[cur])) return rc;
	while (cur != none) {
		if (int rc = take(&nxt)) return rc;
		if (nxt != none)
			if (int rc = prepareTile(w, w->tiles[nxt])) return rc;
		lap(tPrepare);
		if (int rc = enqueueSlices(w, w->tiles[cur])) return rc;
		lap(tEnqueue);
		if (prev != none)
			if (int rc = finishTile(w, w->tiles[prev])) return rc;
		lap(tFinish);
		prev = cur;
		cur = nxt;
	}
	if (int rc = finishTile(w, w->tiles[prev])) return rc;
	lap(tFinish);
	if (trace) fprintf(stderr, "[cpvs trace] grid worker: host %.2f ms preparing tiles, %.2f ms enqueueing slices, %.2f ms finishing tiles\n", tPrepare, tEnqueue, tFinish);
	for (size_t k = 1; k < w->lanes.size(); ++k)
		if (w->laneUsed[k]) CPVS_CUDA(cudaStreamWaitEvent(ctx->stream, w->evLane[k], 0));
	CPVS_CUDA(cudaEventRecord(w->ev1, ctx->stream));
	CPVS_CUDA(cudaEventSynchronize(w->ev1));
	float ms = 0.f;
	CPVS_CUDA(cudaEventElapsedTime(&ms, w->ev0, w->ev1));
	w->deviceMs += ms - w->depthInSpanMs;
	return CPVS_OK;
}

int cpvs_grid_worker_build(cpvs_grid_worker* w, const uint32_t* xy, int count) {
	if (!w || (count > 0 && !xy)) return fail(CPVS_EINVAL, "cpvs_grid_worker_build: NULL argument");
	TileList list{xy, count, 0};
	return cpvs_grid_worker_build_from(w, nextFromList, &list);
}

int cpvs_grid_worker_num_cells(const cpvs_grid_worker* w) { return w ? (int)(w->built * w->desc.length) : 0; }

// The finished cells of this worker: index = cell index in the container ((z * length + y) * length + x).
int cpvs_grid_worker_cells(const cpvs_grid_worker* w, cpvs_grid_cell* out, int capacity) {
	if (!w || !out) return fail(CPVS_EINVAL, "cpvs_grid_worker_cells: NULL argument");
	int n = 0;
	const u32 len = w->desc.length;
	for (const WorkerTile& t : w->tiles) {
		if (!t.built) continue;
		for (u32 z = 0; z < len; ++z) {
			if (n >= capacity) return fail(CPVS_EINVAL, "cpvs_grid_worker_cells: capacity %d", capacity);
			const cpvs_shadow* s = t.cells[z];
			cpvs_grid_cell& c = out[n++];
			c.index = (z * len + t.y) * len + t.x;
			c.words = s->info.words;
			c.num_levels = s->info.num_levels;
			c.root_mask = s->info.total_visibility == CPVS_SHADOW ? 0u : (s->info.total_visibility == CPVS_VISIBLE ? 0x5555u : 0xAAAAu);
			c.device = w->ctx->device;
			c.words_device = s->dag;
			c.svo_nodes = c.dag_nodes = 0;
			for (int l = 0; l < CPVS_MAX_LEVELS; ++l) {
				c.svo_nodes += s->info.svo_nodes[l];
				c.dag_nodes += s->info.dag_nodes[l];
			}
		}
	}
	return n;
}

float cpvs_grid_worker_device_ms(const cpvs_grid_worker* w) { return w ? w->deviceMs : 0.f; }
float cpvs_grid_worker_depth_ms(const cpvs_grid_worker* w) { return w ? w->depthMs : 0.f; }

// One process per GPU: the finished cells are packed into one block of plain device memory (stream-ordered pool memory
// cannot be shared) whose CUDA IPC handle another process opens with cpvs_ipc_open; offsets[i] = first word of the i-th cell
// of cpvs_grid_worker_cells inside the block. The block lives until the worker is destroyed.
int cpvs_grid_worker_export(cpvs_grid_worker* w, unsigned char handle[64], uint64_t* offsets, int capacity) {
	if (!w || !handle || !offsets) return fail(CPVS_EINVAL, "cpvs_grid_worker_export: NULL argument");
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the handle travels as 64 bytes");
	CPVS_CUDA(cudaSetDevice(w->ctx->device));
	u64 total = 0;
	int n = 0;
	for (const WorkerTile& t : w->tiles)
		if (t.built)
			for (const cpvs_shadow* s : t.cells) {
				if (n >= capacity) return fail(CPVS_EINVAL, "cpvs_grid_worker_export: capacity %d", capacity);
				offsets[n++] = total;
				total += s->info.words;
			}
	if (w->exported) CPVS_CUDA(cudaFree(w->exported));
	w->exported = nullptr;
	CPVS_CUDA(cudaMalloc(reinterpret_cast<void**>(&w->exported), (total ? total : 1) * sizeof(u32)));
	n = 0;
	for (const WorkerTile& t : w->tiles)
		if (t.built)
			for (const cpvs_shadow* s : t.cells)
				CPVS_CUDA(cudaMemcpyAsync(w->exported + offsets[n++], s->dag, s->info.words * sizeof(u32), cudaMemcpyDeviceToDevice, w->ctx->stream));
	CPVS_CUDA(cudaStreamSynchronize(w->ctx->stream));
	cudaIpcMemHandle_t h;
	CPVS_CUDA(cudaIpcGetMemHandle(&h, w->exported));
	std::memcpy(handle, &h, 64);
	return n;
}

// The finished cells copied to host memory, one after the other (offsets[i] = first word of the i-th cell of
// cpvs_grid_worker_cells): the exchange medium of callers that run one process per GPU and share a host (a file in shared
// memory), where mapping eight processes' device memory into each other costs seconds of peer-access set-up.
int cpvs_grid_worker_copy_cells(const cpvs_grid_worker* w, uint32_t* outHost, uint64_t capacityWords, uint64_t* offsets, int capacity) {
	if (!w || !offsets || (capacityWords && !outHost)) return fail(CPVS_EINVAL, "cpvs_grid_worker_copy_cells: NULL argument");
	CPVS_CUDA(cudaSetDevice(w->ctx->device));
	u64 total = 0;
	int n = 0;
	for (const WorkerTile& t : w->tiles)
		if (t.built)
			for (const cpvs_shadow* s : t.cells) {
				if (n >= capacity) return fail(CPVS_EINVAL, "cpvs_grid_worker_copy_cells: capacity %d", capacity);
				offsets[n++] = total;
				total += s->info.words;
			}
	if (!outHost) return n;  // sizes only
	if (total > capacityWords) return fail(CPVS_EINVAL, "cpvs_grid_worker_copy_cells: %llu words, room for %llu", (unsigned long long)total, (unsigned long long)capacityWords);
	n = 0;
	for (const WorkerTile& t : w->tiles)
		if (t.built)
			for (const cpvs_shadow* s : t.cells)
				CPVS_CUDA(cudaMemcpyAsync(outHost + offsets[n++], s->dag, s->info.words * sizeof(u32), cudaMemcpyDeviceToHost, w->ctx->stream));
	CPVS_CUDA(cudaStreamSynchronize(w->ctx->stream));
	return n;
}

int cpvs_ipc_open(const unsigned char handle[64], int device, void** out) {
	if (!handle || !out) return fail(CPVS_EINVAL, "cpvs_ipc_open: NULL argument");
	cudaIpcMemHandle_t h;
	std::memcpy(&h, handle, 64);
	CPVS_CUDA(cudaSetDevice(device));
	CPVS_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
	return CPVS_OK;
}

int cpvs_ipc_close(int device, void* ptr) {
	if (!ptr) return CPVS_OK;
	CPVS_CUDA(cudaSetDevice(device));
	CPVS_CUDA(cudaIpcCloseMemHandle(ptr));
	return CPVS_OK;
}

// Longest-processing-time-first: tiles in order of falling cost, each to the least loaded worker; among equally loaded
// workers the one that already holds the tile (its pyramid is resident there). owner_in may be NULL.
int cpvs_grid_assign(const uint64_t* cost, int numTiles, int numWorkers, const int* ownerIn, int* ownerOut) {
	if (!cost || !ownerOut || numTiles < 0 || numWorkers < 1) return fail(CPVS_EINVAL, "cpvs_grid_assign: bad arguments");
	std::vector<int> order(numTiles);
	for (int i = 0; i < numTiles; ++i) order[i] = i;
	std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
	std::vector<u64> load(numWorkers, 0);
	for (int t : order) {
		int best = 0;
		for (int wk = 1; wk < numWorkers; ++wk)
			if (load[wk] < load[best]) best = wk;
		if (ownerIn && ownerIn[t] >= 0 && ownerIn[t] < numWorkers && load[ownerIn[t]] == load[best]) best = ownerIn[t];
		ownerOut[t] = best;
		load[best] += cost[t];
	}
	return CPVS_OK;
}

}  // extern "C"

// ---- the whole grid on several GPUs of one process ----------------------------------------------------------------------

struct cpvs_grid {
	cpvs_grid_desc desc{};
	std::vector<int> devices;
	std::vector<cpvs_ctx*> ctxs;
	std::vector<cpvs_container*> containers;  // one per device (replicated) or only [0]
	cpvs_grid_stats stats{};
};

namespace {

int initialOwner(u32 x, u32 y, u32 length, int workers) { return (int)((y * length + (x + y) % length) % (u32)workers); }

template <typename F>
int runOnWorkers(std::vector<cpvs_grid_worker*>& workers, F fn) {
	std::vector<std::thread> threads;
	for (size_t i = 0; i < workers.size(); ++i)
		threads.emplace_back([&, i]() {
			cpvs_grid_worker* w = workers[i];
			if (w->status != CPVS_OK) return;
			const int rc = fn((int)i, w);
			if (rc != CPVS_OK) {
				w->status = rc;
				w->error = cpvs_last_error();
			}
		});
	for (std::thread& t : threads) t.join();
	for (cpvs_grid_worker* w : workers)
		if (w->status != CPVS_OK) return fail(w->status, "cpvs_grid_build (device %d): %s", w->ctx->device, w->error.c_str());
	return CPVS_OK;
}

}  // namespace

extern "C" {

int cpvs_grid_destroy(cpvs_grid* g) {
	if (!g) return CPVS_OK;
	for (cpvs_container* c : g->containers) cpvs_container_destroy(c);
	for (cpvs_ctx* c : g->ctxs) cpvs_ctx_destroy(c);
	delete g;
	return CPVS_OK;
}

int cpvs_grid_build(const int* devices, int numDevices, const cpvs_grid_desc* desc, int replicate, cpvs_grid** out) {
	if (!devices || numDevices < 1 || numDevices > CPVS_GRID_MAX_DEVICES || !desc || !out)
		return fail(CPVS_EINVAL, "cpvs_grid_build: bad arguments (1..%d devices)", CPVS_GRID_MAX_DEVICES);
	*out = nullptr;
	cpvs_grid* g = new (std::nothrow) cpvs_grid();
	if (!g) return fail(CPVS_ENOMEM, "cpvs_grid_build: host allocation");
	g->desc = *desc;
	g->devices.assign(devices, devices + numDevices);
	std::vector<cpvs_grid_worker*> workers;
	auto cleanup = [&](int rc) {
		for (cpvs_grid_worker* w : workers) cpvs_grid_worker_destroy(w);
		cpvs_grid_destroy(g);
		return rc;
	};
	for (int d = 0; d < numDevices; ++d) {
		cpvs_ctx* ctx = nullptr;
		if (int rc = cpvs_ctx_create(devices[d], &ctx)) return cleanup(rc);
		g->ctxs.push_back(ctx);
		cpvs_grid_worker* w = nullptr;
		if (int rc = cpvs_grid_worker_create(ctx, desc, &w)) return cleanup(rc);
		workers.push_back(w);
		// the kept DAGs and the scratch arenas come out of memory the pool already owns (see cpvs_ctx_reserve)
		const double scale = ((double)desc->tile / 16384.0) * ((double)desc->tile / 16384.0);
		const double share = (double)desc->length * desc->length / numDevices;
		cpvs_ctx_reserve(ctx, (uint64_t)(std::min(share * 320e6 * scale, 8.0 * 1073741824.0) + 6e9 * scale));
	}
	// peer access for the replication (best effort: cudaMemcpyPeerAsync stages through the host without it)
	for (int a = 0; a < numDevices; ++a)
		for (int b = 0; b < numDevices; ++b)
			if (a != b) {
				int can = 0;
				if (cudaDeviceCanAccessPeer(&can, devices[a], devices[b]) == cudaSuccess && can) {
					cudaSetDevice(devices[a]);
					if (cudaDeviceEnablePeerAccess(devices[b], 0) == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
				}
			}
	const u32 len = desc->length;
	const int numTiles = (int)(len * len);
	const double wall0 = nowMs();
	std::vector<uint32_t> xy(2 * (size_t)numTiles);
	std::vector<int> owner(numTiles);
	for (u32 y = 0; y < len; ++y)
		for (u32 x = 0; x < len; ++x) {
			const int t = (int)(y * len + x);  // the reference's loop order: y outer, x inner (src/DeferredRenderer.cpp:170-171)
			xy[2 * t] = x;
			xy[2 * t + 1] = y;
			owner[t] = initialOwner(x, y, len, numDevices);
		}
	cpvs_grid_stats& st = g->stats;
	st.devices = (uint32_t)numDevices;
	// Few tiles per GPU: the tiles are first costed (each worker a round-robin share) and then handed out longest first -- but
	// still from a shared queue, so that a worker whose tile turned out heavier than estimated simply comes back later and
	// finds the lighter ones. Many tiles per GPU: the queue alone, in the reference's loop order.
	const bool costAware = numDevices > 1 && numTiles <= 4 * numDevices;
	std::vector<int> order(numTiles);
	for (int t = 0; t < numTiles; ++t) order[t] = t;
	if (costAware) {
		std::vector<uint64_t> cost(numTiles, 0);
		const int rc = runOnWorkers(workers, [&](int i, cpvs_grid_worker* w) {
			for (int t = 0; t < numTiles; ++t)
				if (owner[t] == i)
					if (int e = cpvs_grid_worker_estimate(w, &xy[2 * t], 1, &cost[t])) return e;
			return (int)CPVS_OK;
		});
		if (rc) return cleanup(rc);
		std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
	}
	{
		const std::vector<int> startOwner = owner;
		std::atomic<int> next(0);
		struct Queue {
			std::atomic<int>* next;
			const std::vector<int>* order;
			const std::vector<uint32_t>* xy;
			std::vector<int>* owner;
			int worker;
		};
		const int rc = runOnWorkers(workers, [&](int i, cpvs_grid_worker* w) {
			Queue q{&next, &order, &xy, &owner, i};
			return cpvs_grid_worker_build_from(w, [](void* user, uint32_t* x, uint32_t* y) {
				Queue* q = static_cast<Queue*>(user);
				const int k = q->next->fetch_add(1);
				if (k >= (int)q->order->size()) return 0;
				const int t = (*q->order)[k];
				(*q->owner)[t] = q->worker;
				*x = (*q->xy)[2 * t];
				*y = (*q->xy)[2 * t + 1];
				return 1;
			}, &q);
		});
		if (rc) return cleanup(rc);
		for (int t = 0; t < numTiles; ++t) {
			if (owner[t] != startOwner[t]) {
				++st.moved_tiles;
				cpvs_grid_worker_release(workers[startOwner[t]], &xy[2 * t], 1);
			}
		}
	}
	st.build_wall_ms = (float)(nowMs() - wall0);

	// host-side gather of sizes: the only cross-GPU step of the build
	const double gather0 = nowMs();
	std::vector<cpvs_cell_part> parts((size_t)len * len * len);
	std::vector<bool> have(parts.size(), false);
	u32 numLevels = 0;
	for (int d = 0; d < numDevices; ++d) {
		std::vector<cpvs_grid_cell> cells((size_t)cpvs_grid_worker_num_cells(workers[d]) + 1);
		const int n = cpvs_grid_worker_cells(workers[d], cells.data(), (int)cells.size());
		if (n < 0) return cleanup(CPVS_EINTERNAL);
		st.tiles[d] = workers[d]->built;
		st.build_ms[d] = workers[d]->deviceMs;
		st.build_ms_max = std::max(st.build_ms_max, workers[d]->deviceMs);
		st.launches += cpvs_ctx_launch_count(workers[d]->ctx);
		st.depth_ms[d] = workers[d]->depthMs;
		for (int i = 0; i < n; ++i) {
			const cpvs_grid_cell& c = cells[i];
			parts[c.index] = cpvs_cell_part{c.words, c.root_mask, c.device, c.words_device};
			have[c.index] = true;
			numLevels = c.num_levels;
			st.svo_nodes += c.svo_nodes;
			st.dag_nodes += c.dag_nodes;
			st.dag_words += c.words;
			if (c.words == 1) ++st.one_word_cells;
		}
	}
	for (size_t i = 0; i < have.size(); ++i)
		if (!have[i]) return cleanup(fail(CPVS_EINTERNAL, "cpvs_grid_build: cell %zu was never built", i));
	st.cells = (uint32_t)parts.size();
	st.gather_ms = (float)(nowMs() - gather0);

	// replication for the lookups: one container per GPU (or only on the first), filled with peer copies
	const double rep0 = nowMs();
	const int copies = replicate ? numDevices : 1;
	g->containers.assign(copies, nullptr);
	{
		std::vector<std::thread> threads;
		std::vector<int> rcs(copies, CPVS_OK);
		std::vector<std::string> errs(copies);
		const bool useLeaf = desc->leafmasks && numLevels >= 5;
		for (int d = 0; d < copies; ++d)
			threads.emplace_back([&, d]() {
				rcs[d] = containerFromParts(g->ctxs[d], len, numLevels, useLeaf ? 1 : 0, parts.data(), &g->containers[d]);
				if (rcs[d] != CPVS_OK) errs[d] = cpvs_last_error();
			});
		for (std::thread& t : threads) t.join();
		for (int d = 0; d < copies; ++d)
			if (rcs[d] != CPVS_OK) return cleanup(fail(rcs[d], "cpvs_grid_build (replicate to device %d): %s", devices[d], errs[d].c_str()));
	}
	st.replicate_ms = (float)(nowMs() - rep0);
	st.wall_ms = (float)(nowMs() - wall0);
	for (cpvs_grid_worker* w : workers) cpvs_grid_worker_destroy(w);  // the cells now live in the containers
	workers.clear();
	*out = g;
	return CPVS_OK;
}

int cpvs_grid_stats_get(const cpvs_grid* g, cpvs_grid_stats* out) {
	if (!g || !out) return fail(CPVS_EINVAL, "cpvs_grid_stats_get: NULL argument");
	*out = g->stats;
	return CPVS_OK;
}

cpvs_container* cpvs_grid_container(const cpvs_grid* g, int index) {
	return (g && index >= 0 && index < (int)g->containers.size()) ? g->containers[index] : nullptr;
}

// traverse.cs over the whole grid: the batch is split by rows over the GPUs that hold a replica.
int cpvs_grid_lookup_ndc(const cpvs_grid* g, const float* ndcHost, int64_t count, uint8_t* outHost) {
	if (!g || g->containers.empty() || count < 0 || (count > 0 && (!ndcHost || !outHost))) return fail(CPVS_EINVAL, "cpvs_grid_lookup_ndc: bad arguments");
	const int parts = (int)g->containers.size();
	std::vector<std::thread> threads;
	std::vector<int> rcs(parts, CPVS_OK);
	for (int d = 0; d < parts; ++d)
		threads.emplace_back([&, d]() {
			const int64_t lo = count * d / parts, hi = count * (d + 1) / parts;
			rcs[d] = cpvs_container_lookup_ndc(g->containers[d], ndcHost + 3 * lo, hi - lo, CPVS_MEM_HOST, outHost + lo);
		});
	for (std::thread& t : threads) t.join();
	for (int rc : rcs)
		if (rc != CPVS_OK) return fail(rc, "cpvs_grid_lookup_ndc: a device failed");
	return CPVS_OK;
}

}  // extern "C"
