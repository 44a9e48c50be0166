// C ABI of cpvs_b200 (include/cpvs_b200.h): handle management and the host-side orchestration of the
// device pipeline  depth -> pyramid -> SVO levels -> bottom-up merge -> compressed DAG -> lookups.
//
// No CPU fallback lives here: every entry point either drives the CUDA kernels or returns an error.
#include <cstdlib>
#include <cstring>
#include <new>

#include "../synth/scene.h"
#include "handles.h"

using namespace cpvs;

namespace {
thread_local std::string gLastError;

// A context runs a build on five streams plus a copy stream, two or more contexts share a GPU, and the driver maps all
// streams of a process onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8): streams that share a queue wait for each
// other's kernels (measured: the copy of a finished DAG, on its own stream, waited 2.3 ms for the next tile's builds). The
// variable is read when the process's CUDA context is created, so it is set when the library is loaded -- unless the process
// has chosen a value itself; a process that has already initialised CUDA keeps what it had (slower, never wrong).
__attribute__((constructor)) void moreHardwareQueues() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }
}  // namespace

namespace cpvs {
constexpr size_t kCacheMinBytes = 4u << 20, kCacheMaxBytes = 24ull << 30, kCacheMaxBlocks = 24;

cudaError_t ctxAlloc(cpvs_ctx* ctx, void** out, size_t bytes) {
	*out = nullptr;
	if (!bytes) bytes = 1;
	if (bytes >= kCacheMinBytes) {
		std::lock_guard<std::mutex> guard(ctx->cacheLock);
		size_t best = ctx->freeBlocks.size();
		for (size_t i = 0; i < ctx->freeBlocks.size(); ++i) {
			const size_t have = ctx->freeBlocks[i].second;
			if (have >= bytes && have - bytes <= bytes / 4 && (best == ctx->freeBlocks.size() || have < ctx->freeBlocks[best].second)) best = i;
		}
		if (best != ctx->freeBlocks.size()) {
			*out = ctx->freeBlocks[best].first;
			ctx->liveBlocks[*out] = ctx->freeBlocks[best].second;
			ctx->cachedBytes -= ctx->freeBlocks[best].second;
			ctx->freeBlocks.erase(ctx->freeBlocks.begin() + best);
			return cudaSuccess;
		}
	}
	cudaError_t e = cudaMallocAsync(out, bytes, ctx->stream);
	if (e == cudaErrorMemoryAllocation) {  // give back what this context keeps for recycling, then once more
		cudaGetLastError();
		std::vector<void*> drop;
		{
			std::lock_guard<std::mutex> guard(ctx->cacheLock);
			for (auto& block : ctx->freeBlocks) drop.push_back(block.first);
			ctx->freeBlocks.clear();
			ctx->cachedBytes = 0;
		}
		for (void* p : drop) cudaFreeAsync(p, ctx->stream);
		e = cudaMallocAsync(out, bytes, ctx->stream);
	}
	if (e == cudaSuccess && bytes >= kCacheMinBytes) {
		std::lock_guard<std::mutex> guard(ctx->cacheLock);
		ctx->liveBlocks[*out] = bytes;
	}
	return e;
}

void ctxAdopt(cpvs_ctx* ctx, void* p, size_t bytes) {
	if (!p || bytes < kCacheMinBytes) return;
	std::lock_guard<std::mutex> guard(ctx->cacheLock);
	ctx->liveBlocks[p] = bytes;
}

void ctxFree(cpvs_ctx* ctx, void* p) {
	if (!p) return;
	std::lock_guard<std::mutex> guard(ctx->cacheLock);
	const auto it = ctx->liveBlocks.find(p);
	if (it == ctx->liveBlocks.end()) {
		cudaFreeAsync(p, ctx->stream);
		return;
	}
	ctx->freeBlocks.emplace_back(p, it->second);
	ctx->cachedBytes += it->second;
	ctx->liveBlocks.erase(it);
	while (!ctx->freeBlocks.empty() && (ctx->cachedBytes > kCacheMaxBytes || ctx->freeBlocks.size() > kCacheMaxBlocks)) {
		cudaFreeAsync(ctx->freeBlocks.front().first, ctx->stream);
		ctx->cachedBytes -= ctx->freeBlocks.front().second;
		ctx->freeBlocks.erase(ctx->freeBlocks.begin());
	}
}

cpvs_ctx* siblingContext(cpvs_ctx* ctx) {
	std::lock_guard<std::mutex> guard(ctx->cacheLock);
	if (!ctx->sibling) {
		if (cpvs_ctx_create(ctx->device, &ctx->sibling) != CPVS_OK) return nullptr;
		ctx->sibling->family = ctx->family;
		ctx->sibling->predictSizes = ctx->predictSizes;
		ctx->sibling->headroomShift = ctx->headroomShift;
		ctx->sibling->stagingMaxWords = ctx->stagingMaxWords;
		ctx->sibling->leafColumns = ctx->leafColumns;
	}
	return ctx->sibling;
}

int fail(int code, const char* fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	gLastError = buf;
	return code;
}
}  // namespace cpvs

namespace {

// Stream-ordered scratch allocations released together when the call ends.
struct Scratch {
	cudaStream_t stream;
	std::vector<void*> blocks;
	explicit Scratch(cudaStream_t s) : stream(s) {}
	~Scratch() {
		for (void* p : blocks) cudaFreeAsync(p, stream);
	}
	template <typename T>
	cudaError_t alloc(T** out, u64 count) {
		void* p = nullptr;
		cudaError_t e = cudaMallocAsync(&p, (count ? count : 1) * sizeof(T), stream);
		if (e == cudaSuccess) blocks.push_back(p);
		*out = static_cast<T*>(p);
		return e;
	}
};

}  // namespace

extern "C" {

const char* cpvs_last_error(void) { return gLastError.c_str(); }
const char* cpvs_version(void) { return "cpvs_b200 0.1 (sm_100a)"; }

int cpvs_ctx_create(int device, cpvs_ctx** out) {
	if (!out) return fail(CPVS_EINVAL, "cpvs_ctx_create: out is NULL");
	*out = nullptr;
	int count = 0;
	CPVS_CUDA(cudaGetDeviceCount(&count));
	if (device < 0 || device >= count) return fail(CPVS_EINVAL, "cpvs_ctx_create: device %d of %d", device, count);
	cudaDeviceProp prop;
	CPVS_CUDA(cudaGetDeviceProperties(&prop, device));
	if (prop.major != 10) return fail(CPVS_ECUDA, "cpvs_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
	CPVS_CUDA(cudaSetDevice(device));
	cudaMemPool_t pool;
	CPVS_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
	unsigned long long keep = ~0ull;  // keep freed scratch cached in the pool between calls
	CPVS_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
	// An allocation never waits for a release that is still queued on another stream (a finished DAG's allocation on the copy
	// stream would otherwise wait for the builds queued behind it); it takes memory that is free now, or new memory.
	int off = 0;
	CPVS_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &off));
	cpvs_ctx* ctx = new (std::nothrow) cpvs_ctx;
	if (!ctx) return fail(CPVS_ENOMEM, "cpvs_ctx_create: host allocation");
	ctx->device = device;
	ctx->launches = 0;
	ctx->arena = nullptr;
	ctx->arenaBytes = 0;
	ctx->scalars = nullptr;
	ctx->own = ctx->aux = ctx->aux2 = ctx->aux3 = ctx->aux4 = ctx->copyStream = nullptr;
	ctx->cachedBytes = 0;
	ctx->stagingWords = 0;
	ctx->familyArenaBytes = 0;
	ctx->dagFreeBytes = 0;
	ctx->sibling = nullptr;
	ctx->family = ctx;
	ctx->predictedBuilds = ctx->exactBuilds = ctx->overflowRebuilds = ctx->reemissions = 0;
	cudaEvent_t* plain[] = {&ctx->evFork, &ctx->evJoin, &ctx->evJoin3, &ctx->evClear, &ctx->evCols, &ctx->evLeafRanked, &ctx->evLeafEmitted, &ctx->evCopyFree};
	for (cudaEvent_t* e : plain) *e = nullptr;
	ctx->buildSerial = 0;
	{
		const char* v = std::getenv("CPVS_LEAF_COLUMNS");
		ctx->leafColumns = (v && v[0] >= '0' && v[0] <= '2') ? v[0] - '0' : 1;
		const char* p = std::getenv("CPVS_PREDICT");
		ctx->predictSizes = (p && p[0] == '0') ? 0 : 1;
		const char* h = std::getenv("CPVS_HEADROOM_SHIFT");  // capacity = predicted + (predicted >> shift); tests use 30 to provoke overflows
		ctx->headroomShift = (h && h[0] >= '0' && h[0] <= '9') ? (unsigned)std::atoi(h) : 3u;
		if (ctx->headroomShift > 40) ctx->headroomShift = 40;
		const char* m = std::getenv("CPVS_STAGING_MAX_WORDS");  // tests: 0 sends every build down the unstaged paths
		ctx->stagingMaxWords = (m && m[0] >= '0' && m[0] <= '9') ? std::strtoull(m, nullptr, 10) : (1ull << 29);
	}
	cudaError_t e = cudaStreamCreateWithFlags(&ctx->own, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&ctx->scalars), kNumScalars * sizeof(u64));
	for (int i = 0; i < 4 && e == cudaSuccess; ++i) {
		u64* slot = nullptr;
		e = cudaMallocHost(reinterpret_cast<void**>(&slot), kNumScalars * sizeof(u64));
		if (e == cudaSuccess) {
			ctx->readbackAll.push_back(slot);
			ctx->readbackFree.push_back(slot);
		}
	}
	int prioLeast = 0, prioGreatest = 0;
	if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&prioLeast, &prioGreatest);
	if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->aux, cudaStreamNonBlocking, prioGreatest);
	if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->copyStream, cudaStreamNonBlocking, prioGreatest);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->aux2, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->aux3, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->aux4, cudaStreamNonBlocking);
	for (cudaEvent_t* ev : plain)
		if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
	ctx->stream = ctx->own;
	if (e != cudaSuccess) {
		cpvs_ctx_destroy(ctx);  // releases whatever was created before the failure
		return fail(CPVS_ECUDA, "cpvs_ctx_create: %s", cudaGetErrorString(e));
	}
	*out = ctx;
	return CPVS_OK;
}

int cpvs_ctx_destroy(cpvs_ctx* ctx) {
	if (!ctx) return CPVS_OK;
	if (ctx->sibling) cpvs_ctx_destroy(ctx->sibling);
	ctx->sibling = nullptr;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	if (ctx->arena) cudaFreeAsync(ctx->arena, ctx->stream);
	for (auto& block : ctx->freeBlocks) cudaFreeAsync(block.first, ctx->stream);
	ctx->freeBlocks.clear();
	cudaStreamSynchronize(ctx->stream);
	if (ctx->scalars) cudaFree(ctx->scalars);
	for (u64* slot : ctx->readbackAll) cudaFreeHost(slot);
	for (u64* buf : ctx->countBuffers) cudaFreeHost(buf);
	for (auto& block : ctx->stagingFree) cudaFreeAsync(block.first, ctx->stream);
	ctx->stagingFree.clear();
	for (auto& block : ctx->dagFree) cudaFreeAsync(block.first, ctx->copyStream);
	ctx->dagFree.clear();
	if (ctx->copyStream) cudaStreamSynchronize(ctx->copyStream);
	for (cudaStream_t st : {ctx->aux, ctx->aux2, ctx->aux3, ctx->aux4, ctx->copyStream, ctx->own})
		if (st) cudaStreamDestroy(st);
	for (cudaEvent_t ev : {ctx->evFork, ctx->evJoin, ctx->evJoin3, ctx->evClear, ctx->evCols, ctx->evLeafRanked, ctx->evLeafEmitted, ctx->evCopyFree})
		if (ev) cudaEventDestroy(ev);
	delete ctx;
	return CPVS_OK;
}

int cpvs_ctx_set_stream(cpvs_ctx* ctx, void* cuda_stream) {
	if (!ctx) return fail(CPVS_EINVAL, "cpvs_ctx_set_stream: ctx is NULL");
	std::lock_guard<std::mutex> guard(ctx->buildLock);
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);  // the arena is ordered on the old stream
	ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own;
	return CPVS_OK;
}
void* cpvs_ctx_get_stream(const cpvs_ctx* ctx) { return ctx ? ctx->stream : nullptr; }

int cpvs_ctx_synchronize(cpvs_ctx* ctx) {
	if (!ctx) return fail(CPVS_EINVAL, "cpvs_ctx_synchronize: ctx is NULL");
	CPVS_CUDA(cudaSetDevice(ctx->device));
	CPVS_CUDA(cudaStreamSynchronize(ctx->stream));
	return CPVS_OK;
}
int cpvs_ctx_reserve(cpvs_ctx* ctx, uint64_t bytes) {
	if (!ctx) return fail(CPVS_EINVAL, "cpvs_ctx_reserve: NULL context");
	if (!bytes) return CPVS_OK;
	CPVS_CUDA(cudaSetDevice(ctx->device));
	void* p = nullptr;
	CPVS_CUDA(cudaMallocAsync(&p, bytes, ctx->stream));
	CPVS_CUDA(cudaFreeAsync(p, ctx->stream));  // stays cached: the pool's release threshold is unlimited
	CPVS_CUDA(cudaStreamSynchronize(ctx->stream));
	return CPVS_OK;
}

int cpvs_ctx_trim(cpvs_ctx* ctx) {
	if (!ctx) return fail(CPVS_EINVAL, "cpvs_ctx_trim: NULL context");
	CPVS_CUDA(cudaSetDevice(ctx->device));
	for (cpvs_ctx* c = ctx; c; c = c->sibling) {
		std::lock_guard<std::mutex> build(c->buildLock);  // no build is being enqueued on this context meanwhile
		std::vector<void*> onBuild, onCopy;
		{
			std::lock_guard<std::mutex> guard(c->cacheLock);
			for (auto& block : c->freeBlocks) onBuild.push_back(block.first);
			c->freeBlocks.clear();
			c->cachedBytes = 0;
			for (auto& block : c->stagingFree) onBuild.push_back(block.first);
			c->stagingFree.clear();
			for (auto& block : c->dagFree) onCopy.push_back(block.first);
			c->dagFree.clear();
			c->dagFreeBytes = 0;
		}
		for (void* p : onBuild) cudaFreeAsync(p, c->stream);
		for (void* p : onCopy) cudaFreeAsync(p, c->copyStream);
		CPVS_CUDA(cudaStreamSynchronize(c->stream));
		CPVS_CUDA(cudaStreamSynchronize(c->copyStream));
	}
	cudaMemPool_t pool;
	CPVS_CUDA(cudaDeviceGetDefaultMemPool(&pool, ctx->device));
	CPVS_CUDA(cudaMemPoolTrimTo(pool, 0));
	return CPVS_OK;
}

uint64_t cpvs_ctx_launch_count(const cpvs_ctx* ctx) { return ctx ? ctx->launches + cpvs_ctx_launch_count(ctx->sibling) : 0; }

int cpvs_ctx_set_prediction(cpvs_ctx* ctx, int enabled, uint32_t headroomShift) {
	if (!ctx) return fail(CPVS_EINVAL, "cpvs_ctx_set_prediction: NULL context");
	std::lock_guard<std::mutex> guard(ctx->buildLock);
	ctx->predictSizes = enabled ? 1 : 0;
	ctx->headroomShift = headroomShift > 40 ? 40 : headroomShift;
	return CPVS_OK;
}

int cpvs_ctx_get_stats(const cpvs_ctx* ctx, cpvs_ctx_stats* out) {
	if (!ctx || !out) return fail(CPVS_EINVAL, "cpvs_ctx_get_stats: NULL argument");
	std::memset(out, 0, sizeof(*out));
	for (const cpvs_ctx* c = ctx; c; c = c->sibling) {  // with the contexts chained to it (cpvs::siblingContext)
		out->predicted_builds += c->predictedBuilds;
		out->exact_builds += c->exactBuilds;
		out->overflow_rebuilds += c->overflowRebuilds;
		out->reemissions += c->reemissions;
	}
	return CPVS_OK;
}

/* ---- MinMaxHierarchy ------------------------------------------------------------------------ */

int cpvs_minmax_build(cpvs_ctx* ctx, const float* depth, int n, int mem, cpvs_minmax** out) {
	return cpvs_minmax_build_tiled(ctx, depth, n, mem, 1, out);
}

int cpvs_minmax_build_tiled(cpvs_ctx* ctx, const float* depth, int n, int mem, uint32_t zTileNum, cpvs_minmax** out) {
	if (!ctx || !depth || !out) return fail(CPVS_EINVAL, "cpvs_minmax_build: NULL argument");
	if ((u64)(n > 0 ? n : 0) * zTileNum > (1ull << 23)) return fail(CPVS_EINVAL, "cpvs_minmax_build_tiled: %u z-slices of side %d", zTileNum, n);
	*out = nullptr;
	if (n < 2 || !isPow2((u64)n) || n > (1 << 19)) return fail(CPVS_EINVAL, "cpvs_minmax_build: side %d is not a power of two in [2, 2^19]", n);
	if (mem != CPVS_MEM_HOST && mem != CPVS_MEM_DEVICE) return fail(CPVS_EINVAL, "cpvs_minmax_build: mem %d", mem);
	CPVS_CUDA(cudaSetDevice(ctx->device));
	cpvs_minmax* mm = new (std::nothrow) cpvs_minmax();
	if (!mm) return fail(CPVS_ENOMEM, "cpvs_minmax_build: host allocation");
	mm->ownedDepth = nullptr;
	mm->levelStorage = nullptr;
	mm->residue = nullptr;
	mm->residueTiles = 0;
	mm->evStart = mm->evBase = mm->evStop = nullptr;
	for (int k = 0; k < kMaxLevels; ++k) mm->level[k] = nullptr;
	mm->ctx = ctx;
	mm->n = n;
	int levels = 1;
	while ((1 << (levels - 1)) < n) ++levels;  // log2(n) + 1 (src/MinMaxHierarchy.cpp:17, .h:60-62)
	mm->numLevels = levels;

	u64 offsets[kMaxLevels] = {0}, total = 0;
	for (int k = 1; k < levels; ++k) {
		offsets[k] = total;
		const u64 side = (u64)n >> k;
		total += (side * side * 2 + 63) & ~63ull;  // floats, each level 256-byte aligned
	}
	// Column residues for the per-column leaf builder: where that builder is used (maps >= 8192^2 whose last build on this
	// context took it, or always when it is forced), one byte per texel written by the base kernel saves it the second pass
	// over the 4-byte depths.
	bool wantResidue = zTileNum > 0 && n >= 128 && (ctx->leafColumns == 2 || (ctx->leafColumns == 1 && n >= 8192));
	if (wantResidue && ctx->leafColumns == 1) {
		std::lock_guard<std::mutex> guard(ctx->buildLock);
		for (int side : ctx->noColumnSides) wantResidue = wantResidue && side != n;
	}
	cudaError_t e = ctxAlloc(ctx, reinterpret_cast<void**>(&mm->levelStorage), total * sizeof(float));
	if (e == cudaSuccess && wantResidue) {
		e = ctxAlloc(ctx, reinterpret_cast<void**>(&mm->residue), (u64)n * n);
		mm->residueTiles = zTileNum;
	}
	if (e == cudaSuccess && mem == CPVS_MEM_HOST) {
		e = ctxAlloc(ctx, reinterpret_cast<void**>(&mm->ownedDepth), (u64)n * n * sizeof(float));
		if (e == cudaSuccess) e = cudaMemcpyAsync(mm->ownedDepth, depth, (u64)n * n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
	}
	if (e != cudaSuccess) {
		ctxFree(ctx, mm->levelStorage);
		ctxFree(ctx, mm->residue);
		ctxFree(ctx, mm->ownedDepth);
		delete mm;
		return fail(e == cudaErrorMemoryAllocation ? CPVS_ENOMEM : CPVS_ECUDA, "cpvs_minmax_build: %s", cudaGetErrorString(e));
	}
	mm->level[0] = mem == CPVS_MEM_HOST ? mm->ownedDepth : depth;
	float* lv[kMaxLevels] = {nullptr};
	for (int k = 1; k < levels; ++k) {
		lv[k] = mm->levelStorage + offsets[k];
		mm->level[k] = lv[k];
	}
	lv[0] = const_cast<float*>(mm->level[0]);
	cudaEventCreate(&mm->evStart);
	cudaEventCreate(&mm->evBase);
	cudaEventCreate(&mm->evStop);
	cudaEventRecord(mm->evStart, ctx->stream);
	mm->columnSlices = 0;
	mm->columnMinLevel = -1;
	mm->lowLevelsBuilt = n < 128;  // small maps take the generic path, which writes every level
	ctx->launches += launchPyramid(mm->level[0], n, lv, levels, false, mm->residue, mm->residueTiles, mm->evBase, ctx->stream);
	cudaEventRecord(mm->evStop, ctx->stream);
	e = cudaGetLastError();
	if (e != cudaSuccess) {
		cpvs_minmax_destroy(mm);
		return fail(CPVS_ECUDA, "pyramid launch: %s", cudaGetErrorString(e));
	}
	*out = mm;
	return CPVS_OK;
}

int cpvs_minmax_destroy(cpvs_minmax* mm) {
	if (!mm) return CPVS_OK;
	cudaSetDevice(mm->ctx->device);
	ctxFree(mm->ctx, mm->levelStorage);
	ctxFree(mm->ctx, mm->residue);
	ctxFree(mm->ctx, mm->ownedDepth);
	if (mm->countsPinned) {  // counts that were begun and never used
		cudaEventSynchronize(mm->evCounts);
		releaseCountsBuffer(mm);
	}
	if (mm->evCounts) cudaEventDestroy(mm->evCounts);
	if (mm->evStart) {
		cudaEventDestroy(mm->evStart);
		cudaEventDestroy(mm->evBase);
		cudaEventDestroy(mm->evStop);
	}
	delete mm;
	return CPVS_OK;
}

}  // extern "C"

namespace cpvs {
// Levels 1 and 2 on demand (accessors, cs::createChildmask, leafmask-less builds).
int ensureLowLevels(const cpvs_minmax* cmm, int level) {
	cpvs_minmax* mm = const_cast<cpvs_minmax*>(cmm);
	if (level < 1 || level > 2) return CPVS_OK;
	std::lock_guard<std::mutex> guard(mm->lowLock);
	if (mm->lowLevelsBuilt) return CPVS_OK;
	CPVS_CUDA(cudaSetDevice(mm->ctx->device));
	float* lv[kMaxLevels];
	for (int k = 0; k < kMaxLevels; ++k) lv[k] = const_cast<float*>(mm->level[k]);
	mm->ctx->launches += launchPyramidLowLevels(mm->level[0], mm->n, lv, mm->ctx->stream);
	CPVS_CUDA(cudaGetLastError());
	CPVS_CUDA(cudaStreamSynchronize(mm->ctx->stream));  // other contexts / streams may read them next
	mm->lowLevelsBuilt = true;
	return CPVS_OK;
}
}  // namespace cpvs

extern "C" {

int cpvs_minmax_timing(const cpvs_minmax* mm, float* totalMs, float* baseKernelMs) {
	if (!mm) return fail(CPVS_EINVAL, "cpvs_minmax_timing: NULL argument");
	CPVS_CUDA(cudaSetDevice(mm->ctx->device));
	CPVS_CUDA(cudaEventSynchronize(mm->evStop));
	float total = 0.f, base = 0.f;
	CPVS_CUDA(cudaEventElapsedTime(&total, mm->evStart, mm->evStop));
	CPVS_CUDA(cudaEventElapsedTime(&base, mm->evStart, mm->evBase));
	if (totalMs) *totalMs = total;
	if (baseKernelMs) *baseKernelMs = mm->n >= 128 ? base : 0.f;
	return CPVS_OK;
}

int cpvs_minmax_num_levels(const cpvs_minmax* mm) { return mm ? mm->numLevels : 0; }
int cpvs_minmax_size(const cpvs_minmax* mm) { return mm ? mm->n : 0; }
const float* cpvs_minmax_level_device(const cpvs_minmax* mm, int level) {
	if (!mm || level < 0 || level >= mm->numLevels || ensureLowLevels(mm, level) != CPVS_OK) return nullptr;
	return mm->level[level];
}

int cpvs_minmax_level(const cpvs_minmax* mm, int level, float* out_host) {
	if (!mm || !out_host) return fail(CPVS_EINVAL, "cpvs_minmax_level: NULL argument");
	if (level < 0 || level >= mm->numLevels) return fail(CPVS_EINVAL, "cpvs_minmax_level: level %d of %d", level, mm->numLevels);
	CPVS_CUDA(cudaSetDevice(mm->ctx->device));
	if (int rc = ensureLowLevels(mm, level)) return rc;
	const u64 side = (u64)mm->n >> level;
	const u64 bytes = side * side * (level == 0 ? 1 : 2) * sizeof(float);
	CPVS_CUDA(cudaMemcpyAsync(out_host, mm->level[level], bytes, cudaMemcpyDeviceToHost, mm->ctx->stream));
	CPVS_CUDA(cudaStreamSynchronize(mm->ctx->stream));
	return CPVS_OK;
}

int cpvs_minmax_childmask(const cpvs_minmax* mm, uint32_t level, uint32_t x, uint32_t y, uint32_t z, uint32_t zTileNum, uint32_t* out) {
	if (!mm || !out) return fail(CPVS_EINVAL, "cpvs_minmax_childmask: NULL argument");
	const u32 side = level < (u32)mm->numLevels ? ((u32)mm->n >> level) : 0;
	if (side < 2 || x + 1 >= side || y + 1 >= side || zTileNum == 0) return fail(CPVS_EINVAL, "cpvs_minmax_childmask: node (%u,%u) outside level %u", x, y, level);
	cpvs_ctx* ctx = mm->ctx;
	CPVS_CUDA(cudaSetDevice(ctx->device));
	if (int rc = ensureLowLevels(mm, (int)level)) return rc;
	std::lock_guard<std::mutex> guard(ctx->buildLock);
	const PyramidView pyr = pyramidView(mm);
	u32* dOut = reinterpret_cast<u32*>(ctx->scalars + kNumScalars - 1);
	ctx->launches += launchChildmask(pyr, (int)level, zTileNum, x, y, z, dOut, ctx->stream);
	CPVS_CUDA(cudaMemcpyAsync(out, dOut, sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
	CPVS_CUDA(cudaStreamSynchronize(ctx->stream));
	return CPVS_OK;
}

/* ---- CompressedShadow (cpvs_shadow_create lives in build.cu) ------------------------------------- */

// Accessors of a shadow created asynchronously wait for the build first.
#define CPVS_READY(s)                                                        \
	do {                                                                     \
		if (int _rc = cpvs_shadow_wait(const_cast<cpvs_shadow*>(s))) return _rc; \
	} while (0)

int cpvs_shadow_info_get(const cpvs_shadow* s, cpvs_shadow_info* info) {
	if (!s || !info) return fail(CPVS_EINVAL, "cpvs_shadow_info_get: NULL argument");
	CPVS_READY(s);
	*info = s->info;
	return CPVS_OK;
}

int cpvs_shadow_copy_dag(const cpvs_shadow* s, uint32_t* out_host) {
	if (!s || !out_host) return fail(CPVS_EINVAL, "cpvs_shadow_copy_dag: NULL argument");
	CPVS_READY(s);
	CPVS_CUDA(cudaSetDevice(s->ctx->device));
	CPVS_CUDA(cudaMemcpyAsync(out_host, s->dag, s->info.words * sizeof(u32), cudaMemcpyDeviceToHost, s->ctx->stream));
	CPVS_CUDA(cudaStreamSynchronize(s->ctx->stream));
	return CPVS_OK;
}

const uint32_t* cpvs_shadow_dag_device(const cpvs_shadow* s) {
	if (!s || cpvs_shadow_wait(const_cast<cpvs_shadow*>(s)) != CPVS_OK) return nullptr;
	return s->dag;
}

}  // extern "C"

namespace {

// Shared by the single-DAG and container lookups: stage host buffers if needed, launch, copy back.
template <typename Launch>
int runLookup(cpvs_ctx* ctx, const float* in, u64 inFloats, int mem, unsigned char* out, u64 outBytes, Launch launch) {
	cudaStream_t st = ctx->stream;
	if (mem == CPVS_MEM_DEVICE) {
		ctx->launches += launch(in, out);
		CPVS_CUDA(cudaGetLastError());
		return CPVS_OK;
	}
	Scratch scratch(st);
	float* dIn;
	unsigned char* dOut;
	CPVS_CUDA(scratch.alloc(&dIn, inFloats));
	CPVS_CUDA(scratch.alloc(&dOut, outBytes));
	CPVS_CUDA(cudaMemcpyAsync(dIn, in, inFloats * sizeof(float), cudaMemcpyHostToDevice, st));
	ctx->launches += launch(dIn, dOut);
	CPVS_CUDA(cudaGetLastError());
	CPVS_CUDA(cudaMemcpyAsync(out, dOut, outBytes, cudaMemcpyDeviceToHost, st));
	CPVS_CUDA(cudaStreamSynchronize(st));
	return CPVS_OK;
}

}  // namespace

extern "C" {

int cpvs_shadow_lookup_ndc(const cpvs_shadow* s, const float* ndc, int64_t count, int mem, int tryLeafmasks, uint8_t* out) {
	if (!s || (count > 0 && (!ndc || !out))) return fail(CPVS_EINVAL, "cpvs_shadow_lookup_ndc: NULL argument");
	if (count < 0) return fail(CPVS_EINVAL, "cpvs_shadow_lookup_ndc: count %lld", (long long)count);
	if (count == 0) return CPVS_OK;
	CPVS_READY(s);
	// The descent must match the layout: following a leafmask DAG below level 3 reads leaf words as pointers, and the other
	// way round -- undefined in the reference, an error here.
	if ((tryLeafmasks != 0) != (s->info.leafmasks != 0))
		return fail(CPVS_EINVAL, "cpvs_shadow_lookup_ndc: tryLeafmasks=%d on a DAG built %s leafmasks (SURVEY.md T2)", tryLeafmasks ? 1 : 0,
				s->info.leafmasks ? "with" : "without");
	CPVS_CUDA(cudaSetDevice(s->ctx->device));
	cudaStream_t st = s->ctx->stream;
	LookupDag d{s->dag, nullptr, s->info.num_levels, 0, tryLeafmasks ? 1 : 0, nullptr, 0};
	{  // the private lookup copy and the shortcut grid, built once
		cpvs_shadow* ms = const_cast<cpvs_shadow*>(s);
		std::lock_guard<std::mutex> guard(ms->skipLock);
		if (d.leafmasks && !ms->index.tried)
			if (int rc = buildLookupIndex(s->ctx, s->dag, s->info.words, std::vector<u64>(1, 0), std::vector<u64>(1, s->info.words), std::vector<u32>(),
						s->info.num_levels, 0, &ms->index))
				return rc;
		if (ms->index.valid) {
			d.dag = ms->index.nodes;
			d.leafCodes = ms->index.codes;
		}
		u32*& skip = ms->index.valid ? ms->index.skip : ms->skip;
		u32& skipLevels = ms->index.valid ? ms->index.skipLevels : ms->skipLevels;
		if (!skip) {
			const u32 g = skipLevelsFor(d.dagLevels, d.leafmasks, 0);
			if (g) {
				CPVS_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&skip), (sizeof(u32) << (3 * g)), st));
				d.skipLevels = g;
				s->ctx->launches += launchBuildSkipGrid(d, skip, st);
				skipLevels = g;
			}
		}
		d.skip = skip;
		d.skipLevels = skipLevels;
	}
	return runLookup(s->ctx, ndc, (u64)count * 3, mem, out, (u64)count,
			[&](const float* in, unsigned char* o) { return launchLookupNdc(d, in, count, o, st); });
}

/* ---- CompressedShadowContainer ---------------------------------------------------------------- */

int cpvs_container_create(cpvs_ctx* ctx, uint32_t length, cpvs_container** out) {
	if (!ctx || !out) return fail(CPVS_EINVAL, "cpvs_container_create: NULL argument");
	*out = nullptr;
	if (!isPow2(length) || length > 64) return fail(CPVS_EINVAL, "cpvs_container_create: length %u must be a power of two <= 64", length);
	cpvs_container* c = new (std::nothrow) cpvs_container;
	if (!c) return fail(CPVS_ENOMEM, "cpvs_container_create: host allocation");
	c->ctx = ctx;
	c->length = length;
	c->filterSize = 1;
	c->cells.resize((size_t)length * length * length);
	*out = c;
	return CPVS_OK;
}

}  // extern "C"

namespace {
// What the lookups of a finished container walk: the private lookup copy (lookup_index.cu) where it could be built, else the DAG
// words themselves; either way with a shortcut grid over the top levels.
// hostGrid: the cell table; cellWords: words per cell in container order.
int prepareContainerLookups(cpvs_container* c, const std::vector<u32>& hostGrid, const std::vector<u64>& cellWords) {
	cpvs_ctx* ctx = c->ctx;
	cudaStream_t st = ctx->stream;
	if (c->leafmasks) {
		std::vector<u64> cellStart(cellWords.size());
		u64 offset = 0;
		for (size_t i = 0; i < cellWords.size(); ++i) {
			cellStart[i] = offset;
			offset += cellWords[i];
		}
		if (int rc = buildLookupIndex(ctx, c->dag, c->dagWords, cellStart, cellWords, hostGrid, c->dagLevels, c->gridLevels, &c->index)) return rc;
	}
	c->skipLevels = skipLevelsFor(c->dagLevels, c->leafmasks, c->gridLevels);
	if (c->skipLevels) {
		u32** skip = c->index.valid ? &c->index.skip : &c->skip;
		CPVS_CUDA(cudaMallocAsync(reinterpret_cast<void**>(skip), sizeof(u32) << (3 * (c->gridLevels + c->skipLevels)), st));
		LookupDag d{c->index.valid ? c->index.nodes : c->dag, c->index.valid ? c->index.grid : c->grid, c->dagLevels, c->gridLevels, c->leafmasks, nullptr,
				c->skipLevels};
		if (c->index.valid) d.leafCodes = c->index.codes;
		ctx->launches += launchBuildSkipGrid(d, *skip, st);
		CPVS_CUDA(cudaGetLastError());
		c->index.skipLevels = c->skipLevels;
	}
	return CPVS_OK;
}

LookupDag containerView(const cpvs_container* c) {
	if (c->index.valid) {
		LookupDag d{c->index.nodes, c->index.grid, c->dagLevels, c->gridLevels, c->leafmasks, c->index.skip, c->skipLevels};
		d.leafCodes = c->index.codes;
		return d;
	}
	return LookupDag{c->dag, c->grid, c->dagLevels, c->gridLevels, c->leafmasks, c->skip, c->skipLevels};
}
}  // namespace

extern "C" {

static void releaseContainerBuffers(cpvs_container* c) {
	freeLookupIndex(c->ctx, &c->index);
	c->index.tried = false;
	if (c->dag) cudaFreeAsync(c->dag, c->ctx->stream);
	if (c->grid) cudaFreeAsync(c->grid, c->ctx->stream);
	if (c->skip) cudaFreeAsync(c->skip, c->ctx->stream);
	c->dag = c->grid = c->skip = nullptr;
	c->skipLevels = 0;
	c->finalized = false;
}

int cpvs_container_destroy(cpvs_container* c) {
	if (!c) return CPVS_OK;
	cudaSetDevice(c->ctx->device);
	releaseContainerBuffers(c);
	for (ContainerCell& cell : c->cells)
		if (cell.words) cudaFreeAsync(cell.words, c->ctx->stream);
	delete c;
	return CPVS_OK;
}

int cpvs_container_set_dag(cpvs_container* c, const uint32_t* words, uint64_t count, int mem, uint32_t numLevels, int leafmasks, uint32_t x,
		uint32_t y, uint32_t z) {
	if (!c || !words || !count) return fail(CPVS_EINVAL, "cpvs_container_set_dag: NULL or empty DAG");
	if (x >= c->length || y >= c->length || z >= c->length)  // assert of src/CompressedShadowContainer.h:35
		return fail(CPVS_EINVAL, "cpvs_container_set_dag: cell (%u,%u,%u) outside length %u", x, y, z, c->length);
	if (c->loaded) return fail(CPVS_EINVAL, "cpvs_container_set_dag: a container loaded from a file cannot be modified");
	CPVS_CUDA(cudaSetDevice(c->ctx->device));
	cudaStream_t st = c->ctx->stream;
	ContainerCell& cell = c->cells[((size_t)z * c->length + y) * c->length + x];  // src/CompressedShadowContainer.h:37-38
	if (cell.words) CPVS_CUDA(cudaFreeAsync(cell.words, st));
	cell = ContainerCell();
	CPVS_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&cell.words), count * sizeof(u32), st));
	CPVS_CUDA(cudaMemcpyAsync(cell.words, words, count * sizeof(u32), mem == CPVS_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
	CPVS_CUDA(cudaMemcpyAsync(&cell.rootMask, cell.words, sizeof(u32), cudaMemcpyDeviceToHost, st));
	CPVS_CUDA(cudaStreamSynchronize(st));
	cell.count = count;
	cell.numLevels = numLevels;
	cell.leafmasks = leafmasks;
	cell.set = true;
	if (c->finalized) releaseContainerBuffers(c);
	return CPVS_OK;
}

int cpvs_container_set(cpvs_container* c, const cpvs_shadow* s, uint32_t x, uint32_t y, uint32_t z) {
	if (!c || !s) return fail(CPVS_EINVAL, "cpvs_container_set: NULL argument");
	CPVS_READY(s);
	if (s->ctx->device != c->ctx->device) return fail(CPVS_EINVAL, "cpvs_container_set: shadow lives on another device; use cpvs_container_set_dag");
	// a one-word DAG is stored without synchronising its builder's stream: order this stream behind that store
	if (s->ready) CPVS_CUDA(cudaStreamWaitEvent(c->ctx->stream, s->ready, 0));
	return cpvs_container_set_dag(c, s->dag, s->info.words, CPVS_MEM_DEVICE, s->info.num_levels, (int)s->info.leafmasks, x, y, z);
}

int cpvs_container_finalize(cpvs_container* c) {
	if (!c) return fail(CPVS_EINVAL, "cpvs_container_finalize: NULL argument");
	if (c->loaded) return CPVS_OK;  // a loaded container is final
	CPVS_CUDA(cudaSetDevice(c->ctx->device));
	cudaStream_t st = c->ctx->stream;
	u64 total = 0;
	for (size_t i = 0; i < c->cells.size(); ++i) {
		const ContainerCell& cell = c->cells[i];
		if (!cell.set) return fail(CPVS_EINVAL, "cpvs_container_finalize: cell %zu was never set", i);
		if (cell.numLevels != c->cells[0].numLevels || cell.leafmasks != c->cells[0].leafmasks)
			return fail(CPVS_EINVAL, "cpvs_container_finalize: cell %zu differs in levels/leafmasks from cell 0 (src/CompressedShadowContainer.cpp:39-40)", i);
		total += cell.count;
	}
	if (total > (1ull << 32)) return fail(CPVS_EOVERFLOW, "combined DAG needs %llu words; offsets are 32-bit", (unsigned long long)total);
	releaseContainerBuffers(c);
	std::vector<u32> grid(c->cells.size());
	CPVS_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&c->dag), total * sizeof(u32), st));
	CPVS_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&c->grid), grid.size() * sizeof(u32), st));
	u64 offset = 0;
	for (size_t i = 0; i < c->cells.size(); ++i) {  // combineDAGs (:52-69) + createTopLevelGrid (:71-91)
		const ContainerCell& cell = c->cells[i];
		grid[i] = cell.rootMask == 0u ? CPVS_GRID_CELL_SHADOWED : (cell.rootMask == 0x5555u ? CPVS_GRID_CELL_VISIBLE : (u32)offset);
		CPVS_CUDA(cudaMemcpyAsync(c->dag + offset, cell.words, cell.count * sizeof(u32), cudaMemcpyDeviceToDevice, st));
		offset += cell.count;
	}
	CPVS_CUDA(cudaMemcpyAsync(c->grid, grid.data(), grid.size() * sizeof(u32), cudaMemcpyHostToDevice, st));
	CPVS_CUDA(cudaStreamSynchronize(st));
	c->dagWords = total;
	c->dagLevels = c->cells[0].numLevels;  // src/CompressedShadowContainer.cpp:40
	c->gridLevels = 0;                     // log8(#cells) (:42-43), exact here
	while ((1u << c->gridLevels) < c->length) ++c->gridLevels;
	c->leafmasks = c->cells[0].leafmasks;
	std::vector<u64> cellWords(c->cells.size());
	for (size_t i = 0; i < c->cells.size(); ++i) cellWords[i] = c->cells[i].count;
	if (int rc = prepareContainerLookups(c, grid, cellWords)) return rc;
	c->finalized = true;
	return CPVS_OK;
}

}  // extern "C"

namespace {
struct WordPatch {
	u64 offset;
	u32 value;
};
__global__ void patchWordsKernel(u32* __restrict__ dag, const WordPatch* __restrict__ patches, u32 count) {
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count) dag[patches[i].offset] = patches[i].value;
}
}  // namespace

namespace cpvs {
// combineDAGs + createTopLevelGrid from cells resident on any GPU of the box.
int containerFromParts(cpvs_ctx* ctx, u32 length, u32 numLevels, int leafmasks, const cpvs_cell_part* parts, cpvs_container** out) {
	*out = nullptr;
	cpvs_container* c = nullptr;
	if (int rc = cpvs_container_create(ctx, length, &c)) return rc;
	c->loaded = true;  // no per-cell copies: the cells cannot be re-set
	CPVS_CUDA(cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	const size_t numCells = c->cells.size();
	u64 total = 0;
	for (size_t i = 0; i < numCells; ++i) total += parts[i].words;
	if (total == 0 || total > (1ull << 32)) {
		cpvs_container_destroy(c);
		return fail(CPVS_EOVERFLOW, "combined DAG needs %llu words; offsets are 32-bit", (unsigned long long)total);
	}
	std::vector<u32> grid(numCells);
	std::vector<WordPatch> patches;
	cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&c->dag), total * sizeof(u32), st);
	if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void**>(&c->grid), numCells * sizeof(u32), st);
	u64 offset = 0;
	for (size_t i = 0; i < numCells && e == cudaSuccess; ++i) {
		const cpvs_cell_part& p = parts[i];
		grid[i] = p.root_mask == 0u ? CPVS_GRID_CELL_SHADOWED : (p.root_mask == 0x5555u ? CPVS_GRID_CELL_VISIBLE : (u32)offset);
		if (p.words == 1 && (!p.words_device || p.root_mask == 0u || p.root_mask == 0x5555u))
			patches.push_back(WordPatch{offset, p.root_mask});  // (thousands of one-word cells in a tall grid: one kernel stores them all)
		else if (!p.words_device)
			e = cudaErrorInvalidValue;
		else if (p.device == ctx->device)
			e = cudaMemcpyAsync(c->dag + offset, p.words_device, p.words * sizeof(u32), cudaMemcpyDeviceToDevice, st);
		else
			e = cudaMemcpyPeerAsync(c->dag + offset, ctx->device, p.words_device, p.device, p.words * sizeof(u32), st);
		offset += p.words;
	}
	WordPatch* dPatches = nullptr;
	if (e == cudaSuccess && !patches.empty()) {
		e = cudaMallocAsync(reinterpret_cast<void**>(&dPatches), patches.size() * sizeof(WordPatch), st);
		if (e == cudaSuccess) e = cudaMemcpyAsync(dPatches, patches.data(), patches.size() * sizeof(WordPatch), cudaMemcpyHostToDevice, st);
		if (e == cudaSuccess) {
			patchWordsKernel<<<(unsigned)((patches.size() + 255) / 256), 256, 0, st>>>(c->dag, dPatches, (u32)patches.size());
			++ctx->launches;
			e = cudaGetLastError();
		}
	}
	if (e == cudaSuccess) e = cudaMemcpyAsync(c->grid, grid.data(), numCells * sizeof(u32), cudaMemcpyHostToDevice, st);
	c->dagWords = total;
	c->dagLevels = numLevels;
	c->gridLevels = 0;
	while ((1u << c->gridLevels) < length) ++c->gridLevels;
	c->leafmasks = leafmasks;
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // host staging goes away; the sources may be released by the caller
	if (dPatches) cudaFreeAsync(dPatches, st);
	if (e != cudaSuccess) {
		cpvs_container_destroy(c);
		return fail(CPVS_ECUDA, "cpvs_container_assemble: %s", cudaGetErrorString(e));
	}
	{
		std::vector<u64> cellWords(numCells);
		for (size_t i = 0; i < numCells; ++i) cellWords[i] = parts[i].words;
		if (int rc = prepareContainerLookups(c, grid, cellWords)) {
			cpvs_container_destroy(c);
			return rc;
		}
	}
	c->finalized = true;
	*out = c;
	return CPVS_OK;
}
}  // namespace cpvs

extern "C" {

int cpvs_container_assemble(cpvs_ctx* ctx, uint32_t length, uint32_t numLevels, int leafmasks, const cpvs_cell_part* cells, cpvs_container** out) {
	if (!ctx || !cells || !out) return fail(CPVS_EINVAL, "cpvs_container_assemble: NULL argument");
	if (numLevels <= 3 || numLevels >= (u32)kMaxLevels) return fail(CPVS_EINVAL, "cpvs_container_assemble: %u levels", numLevels);
	return containerFromParts(ctx, length, numLevels, leafmasks ? 1 : 0, cells, out);
}

int cpvs_container_info(const cpvs_container* c, uint64_t* dagWords, uint32_t* gridCells, uint32_t* dagLevels, uint32_t* gridLevels) {
	if (!c || !c->finalized) return fail(CPVS_EINVAL, "cpvs_container_info: container not finalized");
	if (dagWords) *dagWords = c->dagWords;
	if (gridCells) *gridCells = (u32)c->cells.size();
	if (dagLevels) *dagLevels = c->dagLevels;
	if (gridLevels) *gridLevels = c->gridLevels;
	return CPVS_OK;
}

int cpvs_container_copy(const cpvs_container* c, uint32_t* dagOut, uint32_t* gridOut) {
	if (!c || !c->finalized) return fail(CPVS_EINVAL, "cpvs_container_copy: container not finalized");
	CPVS_CUDA(cudaSetDevice(c->ctx->device));
	if (dagOut) CPVS_CUDA(cudaMemcpyAsync(dagOut, c->dag, c->dagWords * sizeof(u32), cudaMemcpyDeviceToHost, c->ctx->stream));
	if (gridOut) CPVS_CUDA(cudaMemcpyAsync(gridOut, c->grid, c->cells.size() * sizeof(u32), cudaMemcpyDeviceToHost, c->ctx->stream));
	CPVS_CUDA(cudaStreamSynchronize(c->ctx->stream));
	return CPVS_OK;
}

int cpvs_container_lookup_ndc(const cpvs_container* c, const float* ndc, int64_t count, int mem, uint8_t* out) {
	if (!c || !c->finalized) return fail(CPVS_EINVAL, "cpvs_container_lookup_ndc: container not finalized");
	if (count < 0 || (count > 0 && (!ndc || !out))) return fail(CPVS_EINVAL, "cpvs_container_lookup_ndc: bad arguments");
	if (count == 0) return CPVS_OK;
	CPVS_CUDA(cudaSetDevice(c->ctx->device));
	const LookupDag d = containerView(c);
	cudaStream_t st = c->ctx->stream;
	return runLookup(c->ctx, ndc, (u64)count * 3, mem, out, (u64)count,
			[&](const float* in, unsigned char* o) { return launchLookupNdc(d, in, count, o, st); });
}

int cpvs_container_evaluate(const cpvs_container* c, const float* positions, uint32_t width, uint32_t height, int mem, const float m[16],
		uint8_t* visibilities) {
	if (!c || !c->finalized) return fail(CPVS_EINVAL, "cpvs_container_evaluate: container not finalized");
	if (!positions || !m || !visibilities) return fail(CPVS_EINVAL, "cpvs_container_evaluate: NULL argument");
	const long long count = (long long)width * height;
	if (count == 0) return CPVS_OK;
	CPVS_CUDA(cudaSetDevice(c->ctx->device));
	const LookupDag d = containerView(c);
	cudaStream_t st = c->ctx->stream;
	return runLookup(c->ctx, positions, (u64)count * 4, mem, visibilities, (u64)count,
			[&](const float* in, unsigned char* o) { return launchEvaluate(d, in, width, height, m, (int)c->filterSize, o, st); });
}

// evaluate() on the textures themselves (src/CompressedShadowContainer.cpp:100-107 binds image units 0 = rgba32f positions and
// 1 = r8 visibilities): CUDA surface objects over the arrays behind them.
int cpvs_container_evaluate_surface(const cpvs_container* c, unsigned long long positions, unsigned long long visibilities, uint32_t width,
		uint32_t height, const float m[16]) {
	if (!c || !c->finalized) return fail(CPVS_EINVAL, "cpvs_container_evaluate_surface: container not finalized");
	if (!positions || !visibilities || !m) return fail(CPVS_EINVAL, "cpvs_container_evaluate_surface: NULL argument");
	if (!width || !height) return CPVS_OK;
	CPVS_CUDA(cudaSetDevice(c->ctx->device));
	const LookupDag d = containerView(c);
	c->ctx->launches += launchEvaluateSurface(d, positions, visibilities, width, height, m, (int)c->filterSize, c->ctx->stream);
	CPVS_CUDA(cudaGetLastError());
	return CPVS_OK;
}

#ifdef CPVS_WITH_GL
}  // extern "C"
// CUDA-GL interop for DeferredRenderer::doAllShading (src/DeferredRenderer.cpp:320-347): the G-buffer position texture
// (rgba32f, :80) and the visibility texture (r8, :18) are registered once, mapped per frame, and evaluated in place of the
// glDispatchCompute of CompressedShadowContainer::evaluate (src/CompressedShadowContainer.cpp:93-124). Needs GL headers and a
// GL context current on the calling thread; not compiled in this repository's default build (no GL on the build machine).
#include <cuda_gl_interop.h>
extern "C" {
int cpvs_container_evaluate_gl(const cpvs_container* c, unsigned int positionsTexture, unsigned int visibilitiesTexture, uint32_t width, uint32_t height,
		const float m[16]) {
	if (!c || !c->finalized) return fail(CPVS_EINVAL, "cpvs_container_evaluate_gl: container not finalized");
	CPVS_CUDA(cudaSetDevice(c->ctx->device));
	cudaStream_t st = c->ctx->stream;
	cudaGraphicsResource_t res[2] = {nullptr, nullptr};
	CPVS_CUDA(cudaGraphicsGLRegisterImage(&res[0], positionsTexture, 0x0DE1 /* GL_TEXTURE_2D */, cudaGraphicsRegisterFlagsReadOnly));
	cudaError_t e = cudaGraphicsGLRegisterImage(&res[1], visibilitiesTexture, 0x0DE1, cudaGraphicsRegisterFlagsSurfaceLoadStore);
	if (e == cudaSuccess) e = cudaGraphicsMapResources(2, res, st);
	cudaSurfaceObject_t surf[2] = {0, 0};
	for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
		cudaArray_t array = nullptr;
		e = cudaGraphicsSubResourceGetMappedArray(&array, res[i], 0, 0);
		cudaResourceDesc desc = {};
		desc.resType = cudaResourceTypeArray;
		desc.res.array.array = array;
		if (e == cudaSuccess) e = cudaCreateSurfaceObject(&surf[i], &desc);
	}
	int rc = CPVS_OK;
	if (e == cudaSuccess) rc = cpvs_container_evaluate_surface(c, surf[0], surf[1], width, height, m);
	for (cudaSurfaceObject_t s : surf)
		if (s) cudaDestroySurfaceObject(s);
	if (res[0] && res[1]) cudaGraphicsUnmapResources(2, res, st);
	for (cudaGraphicsResource_t r : res)
		if (r) cudaGraphicsUnregisterResource(r);
	if (e != cudaSuccess) return fail(CPVS_ECUDA, "cpvs_container_evaluate_gl: %s", cudaGetErrorString(e));
	return rc;
}
#endif

namespace {
struct ContainerFileHeader {
	char magic[8];
	u32 version, length, dagLevels, gridLevels, leafmasks, reserved;
	u64 dagWords, gridCells, fnv64;
	u64 pad;
};
static_assert(sizeof(ContainerFileHeader) == 64, "on-disk header is 64 bytes");
u64 fnv64Words(u64 h, const u32* w, u64 n) {
	for (u64 i = 0; i < n; ++i) {
		h ^= w[i];
		h *= 1099511628211ull;
	}
	return h;
}
// Checksum over everything the lookup trusts: the header fields, the grid and the DAG words.
u64 containerChecksum(const ContainerFileHeader& h, const u32* grid, const u32* dag) {
	const u32 fields[] = {h.version, h.length, h.dagLevels, h.gridLevels, h.leafmasks, (u32)h.dagWords, (u32)(h.dagWords >> 32),
			(u32)h.gridCells, (u32)(h.gridCells >> 32)};
	u64 sum = fnv64Words(14695981039346656037ull, fields, sizeof(fields) / sizeof(fields[0]));
	sum = fnv64Words(sum, grid, h.gridCells);
	return fnv64Words(sum, dag, h.dagWords);
}
}  // namespace

int cpvs_container_save(const cpvs_container* c, const char* path) {
	if (!c || !c->finalized || !path) return fail(CPVS_EINVAL, "cpvs_container_save: container not finalized or NULL path");
	std::vector<u32> dag, grid;
	try {
		dag.resize(c->dagWords);
		grid.resize(c->cells.size());
	} catch (const std::bad_alloc&) {
		return fail(CPVS_ENOMEM, "cpvs_container_save: host staging of %llu words", (unsigned long long)c->dagWords);
	}
	int rc = cpvs_container_copy(c, dag.data(), grid.data());
	if (rc != CPVS_OK) return rc;
	ContainerFileHeader h;
	std::memset(&h, 0, sizeof(h));
	std::memcpy(h.magic, "CPVSDAG2", 8);
	h.version = 2;
	h.length = c->length;
	h.dagLevels = c->dagLevels;
	h.gridLevels = c->gridLevels;
	h.leafmasks = (u32)c->leafmasks;
	h.dagWords = c->dagWords;
	h.gridCells = grid.size();
	h.fnv64 = containerChecksum(h, grid.data(), dag.data());
	FILE* f = std::fopen(path, "wb");
	if (!f) return fail(CPVS_EINVAL, "cpvs_container_save: cannot open %s", path);
	const bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1 && std::fwrite(grid.data(), sizeof(u32), grid.size(), f) == grid.size() &&
					std::fwrite(dag.data(), sizeof(u32), dag.size(), f) == dag.size();
	std::fclose(f);
	return ok ? CPVS_OK : fail(CPVS_EINVAL, "cpvs_container_save: short write to %s", path);
}

int cpvs_container_load(cpvs_ctx* ctx, const char* path, cpvs_container** out) {
	if (!ctx || !path || !out) return fail(CPVS_EINVAL, "cpvs_container_load: NULL argument");
	*out = nullptr;
	FILE* f = std::fopen(path, "rb");
	if (!f) return fail(CPVS_EINVAL, "cpvs_container_load: cannot open %s", path);
	ContainerFileHeader h;
	std::vector<u32> dag, grid;
	// Every header field the lookup indexes with is validated against the others and against the file's size before
	// anything is allocated; grid entries are checked against the DAG's extent afterwards.
	bool ok = std::fread(&h, sizeof(h), 1, f) == 1 && std::memcmp(h.magic, "CPVSDAG2", 8) == 0 && h.version == 2 && isPow2(h.length) &&
			  h.length <= 64 && h.gridCells == (u64)h.length * h.length * h.length && h.dagWords > 0 && h.dagWords <= (1ull << 32) &&
			  h.dagLevels > 3 && h.dagLevels < (u32)kMaxLevels && h.leafmasks <= 1 && (1u << h.gridLevels) == h.length &&
			  (!h.leafmasks || h.dagLevels >= 5);
	if (ok) {
		ok = std::fseek(f, 0, SEEK_END) == 0;
		const long long size = ok ? (long long)std::ftell(f) : -1;
		ok = ok && size == (long long)(sizeof(h) + (h.gridCells + h.dagWords) * sizeof(u32)) && std::fseek(f, (long)sizeof(h), SEEK_SET) == 0;
	}
	if (ok) {
		try {
			grid.resize(h.gridCells);
			dag.resize(h.dagWords);
		} catch (const std::bad_alloc&) {
			std::fclose(f);
			return fail(CPVS_ENOMEM, "cpvs_container_load: host staging of %llu words", (unsigned long long)h.dagWords);
		}
		ok = std::fread(grid.data(), sizeof(u32), grid.size(), f) == grid.size() && std::fread(dag.data(), sizeof(u32), dag.size(), f) == dag.size() &&
			 containerChecksum(h, grid.data(), dag.data()) == h.fnv64;
		for (size_t i = 0; ok && i < grid.size(); ++i)
			ok = grid[i] == CPVS_GRID_CELL_SHADOWED || grid[i] == CPVS_GRID_CELL_VISIBLE || grid[i] < h.dagWords;
	}
	std::fclose(f);
	if (!ok) return fail(CPVS_EINVAL, "cpvs_container_load: %s is not a valid container file (header, size, checksum or grid)", path);
	cpvs_container* c = nullptr;
	int rc = cpvs_container_create(ctx, h.length, &c);
	if (rc != CPVS_OK) return rc;
	c->loaded = true;
	cudaStream_t st = ctx->stream;
	cudaError_t e = cudaSetDevice(ctx->device);
	if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void**>(&c->dag), dag.size() * sizeof(u32), st);
	if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void**>(&c->grid), grid.size() * sizeof(u32), st);
	if (e == cudaSuccess) e = cudaMemcpyAsync(c->dag, dag.data(), dag.size() * sizeof(u32), cudaMemcpyHostToDevice, st);
	if (e == cudaSuccess) e = cudaMemcpyAsync(c->grid, grid.data(), grid.size() * sizeof(u32), cudaMemcpyHostToDevice, st);
	c->dagWords = h.dagWords;
	c->dagLevels = h.dagLevels;
	c->gridLevels = h.gridLevels;
	c->leafmasks = (int)h.leafmasks;
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // the host staging buffers go away
	if (e != cudaSuccess) {
		cpvs_container_destroy(c);
		return fail(CPVS_ECUDA, "cpvs_container_load: %s", cudaGetErrorString(e));
	}
	{
		// words per cell from the cell table: a cell with a DAG starts at its grid entry, a uniform cell is one word
		// (createTopLevelGrid advances the running offset for every cell, src/CompressedShadowContainer.cpp:88)
		std::vector<u64> cellWords(grid.size(), 1);
		size_t last = grid.size();
		u64 end = h.dagWords;
		for (size_t i = grid.size(); i-- > 0;) {
			if (grid[i] == CPVS_GRID_CELL_SHADOWED || grid[i] == CPVS_GRID_CELL_VISIBLE) continue;
			const u64 uniformBehind = (last == grid.size() ? grid.size() : last) - i - 1;
			cellWords[i] = end - grid[i] - uniformBehind;
			end = grid[i];
			last = i;
		}
		bool consistent = true;
		u64 sum = 0;
		for (size_t i = 0; i < grid.size(); ++i) {
			consistent = consistent && cellWords[i] >= 1 && cellWords[i] <= h.dagWords;
			sum += cellWords[i];
		}
		if (!consistent || sum != h.dagWords) {
			cpvs_container_destroy(c);
			return fail(CPVS_EINVAL, "cpvs_container_load: %s: the cell table does not tile the DAG words", path);
		}
		if (int rc = prepareContainerLookups(c, grid, cellWords)) {
			cpvs_container_destroy(c);
			return rc;
		}
	}
	c->finalized = true;
	*out = c;
	return CPVS_OK;
}

int cpvs_container_set_filter_size(cpvs_container* c, uint32_t size) {
	if (!c) return fail(CPVS_EINVAL, "cpvs_container_set_filter_size: NULL argument");
	if (size < 1 || size > 64) return fail(CPVS_EINVAL, "cpvs_container_set_filter_size: %u (1..64)", size);
	c->filterSize = size;
	return CPVS_OK;
}

int cpvs_depth_generate(cpvs_ctx* ctx, int kind, int n, int tileX, int tileY, int tilesPerSide, float* depthDevice) {
	if (!ctx || !depthDevice) return fail(CPVS_EINVAL, "cpvs_depth_generate: NULL argument");
	if (n < 4 || (n & 3) || tilesPerSide < 1 || tileX < 0 || tileY < 0 || tileX >= tilesPerSide || tileY >= tilesPerSide)
		return fail(CPVS_EINVAL, "cpvs_depth_generate: bad window (n=%d tile=%d,%d of %d)", n, tileX, tileY, tilesPerSide);
	const long long gn = (long long)n * tilesPerSide;
	if (gn > (1ll << 24)) return fail(CPVS_EINVAL, "cpvs_depth_generate: virtual side %lld exceeds 2^24 (float texel coordinates)", gn);
	const long long gx0 = (long long)tileX * n, gy0 = (long long)tileY * n;
	CPVS_CUDA(cudaSetDevice(ctx->device));
	if (kind == CPVS_SCENE_PLANE) {
		ctx->launches += launchPlaneDepth(depthDevice, n, gx0, gy0, gn, ctx->stream);
	} else if (kind == CPVS_SCENE_TERRAIN_DEV) {
		ctx->launches += launchTerrainDevDepth(depthDevice, n, gx0, gy0, gn, ctx->stream);
	} else if (kind == CPVS_SCENE_CITY) {
		std::vector<CityBoxDev> boxes;
		cpvs_synth::forEachCityBox(gn, gx0, gy0, n, [&](const cpvs_synth::CityBox& b) { boxes.push_back(CityBoxDev{b.x0, b.y0, b.x1, b.y1, b.z}); });
		CityBoxDev* dBoxes = nullptr;
		if (!boxes.empty()) {
			CPVS_CUDA(cudaMallocAsync(&dBoxes, boxes.size() * sizeof(CityBoxDev), ctx->stream));
			// pageable source: the copy is staged before the call returns, so `boxes` may go out of scope
			CPVS_CUDA(cudaMemcpyAsync(dBoxes, boxes.data(), boxes.size() * sizeof(CityBoxDev), cudaMemcpyHostToDevice, ctx->stream));
		}
		ctx->launches += launchCityDepth(depthDevice, n, dBoxes, (int)boxes.size(), cpvs_synth::kCityFarPlane, ctx->stream);
		if (dBoxes) CPVS_CUDA(cudaFreeAsync(dBoxes, ctx->stream));
	} else {
		return fail(CPVS_EINVAL, "cpvs_depth_generate: scene %d has no device generator", kind);
	}
	CPVS_CUDA(cudaGetLastError());
	return CPVS_OK;
}

}  // extern "C"
