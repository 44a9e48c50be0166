// C ABI of cpvs_b200 (include/cpvs_b200.h): handle management and the host-side orchestration of the
// device pipeline  depth -> pyramid -> SVO levels -> bottom-up merge -> compressed DAG -> lookups.
//
// No CPU fallback lives here: every entry point either drives the CUDA kernels or returns an error.
#include "../../include/cpvs_b200.h"

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../synth/scene.h"
#include "kernels.h"

using namespace cpvs;

namespace {

thread_local std::string gLastError;

int fail(int code, const char* fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	gLastError = buf;
	return code;
}

#define CPVS_CUDA(expr)                                                                                      \
	do {                                                                                                     \
		cudaError_t _e = (expr);                                                                             \
		if (_e != cudaSuccess)                                                                               \
			return fail(_e == cudaErrorMemoryAllocation ? CPVS_ENOMEM : CPVS_ECUDA, "%s: %s (%s:%d)", #expr, \
					cudaGetErrorString(_e), __FILE__, __LINE__);                                             \
	} while (0)

inline bool isPow2(u64 v) { return v && !(v & (v - 1)); }
inline u64 pow2AtLeast(u64 v) {
	u64 p = 1;
	while (p < v) p <<= 1;
	return p;
}

__global__ void storeU64Kernel(u64* dst, u64 value) { *dst = value; }
__global__ void storeU32Kernel(u32* dst, u32 value) { *dst = value; }

// CPVS_TRACE=1: host wall-clock between orchestration steps, to stderr.
struct HostTrace {
	bool on;
	std::chrono::steady_clock::time_point last;
	HostTrace() : on(std::getenv("CPVS_TRACE") != nullptr), last(std::chrono::steady_clock::now()) {}
	void mark(const char* what) {
		if (!on) return;
		const auto now = std::chrono::steady_clock::now();
		fprintf(stderr, "[cpvs trace] %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(now - last).count());
		last = now;
	}
};

}  // namespace

struct cpvs_ctx {
	int device;
	cudaStream_t own;
	cudaStream_t stream;
	u64 launches;
	// Scratch arena for cpvs_shadow_create: one device allocation, grown when a build needs more and
	// kept between calls, carved by bump pointer -- no allocator traffic in steady state. Builds on one
	// context are serialised by `buildLock` (use one context per host thread for concurrent builds).
	std::mutex buildLock;
	char* arena;
	size_t arenaBytes;
	u64* scalars;  // 192 device words: per-level counters of the build in flight
	u64* hostScalars;  // pinned mirror for the two read-backs of a build (pageable copies are staged)
	// High-priority side stream for a short chain of small kernels that is independent of a bulk kernel
	// on the main stream (the inner levels' emission next to the leaf level's): its CTAs are scheduled
	// ahead of the bulk kernel's queued ones. Fork/join through the two events.
	cudaStream_t aux;
	cudaEvent_t evFork, evJoin, evAuxStart;
	// Normal-priority side stream for the per-level rank scans, which only the final emission needs and
	// which therefore run next to the following levels' inserts.
	cudaStream_t aux2, aux3;  // the ranks of different levels are independent: alternate between the two
	cudaEvent_t evRankStart, evRankStop, evJoin3, evClear, evCols;
	int leafColumns;  // leaves built per column: 1 = where it pays (default), 0 = never, 2 = always (CPVS_LEAF_COLUMNS; tests)
	unsigned experiments;  // kExperiment* bits (CPVS_EXPERIMENTS): unmeasured kernel variants, off by default
	// "early-bases" (experimental): the bottom level's rank runs on a stream of its own, and the level bases, the size
	// read-back and the DAG allocation only wait for its first two kernels (the sizes), not for the third (the writes).
	cudaStream_t aux4;
	cudaEvent_t evBottomSized, evBottomRanked;
	// "leaf-fp64" (experimental): the leaf groups found by fingerprint are verified on this stream, beside the inner inserts;
	// forceExact is set while a build whose verification failed is redone with the exact insert.
	cudaStream_t aux5;
	cudaEvent_t evVerified;
	bool forceExact;
};

struct cpvs_minmax {
	cpvs_ctx* ctx;
	int n;
	int numLevels;
	float* ownedDepth;   // device copy when built from host memory
	float* levelStorage; // levels 1.. in one allocation
	const float* level[kMaxLevels];
	cudaEvent_t evStart, evBase, evStop;
	// Levels 1 and 2 are not needed by the leafmask builder and are only produced on first use.
	std::mutex lowLock;
	bool lowLevelsBuilt;
	// Roots of all z-slices of the column for the zTileNum last asked for (createShadowTiles builds them all
	// from this one pyramid): slices that miss the surface are answered from here without a device round trip.
	u32 columnSlices;
	std::vector<u64> columnRoots;  // [z] = hasNodes << 32 | root mask
};

struct cpvs_shadow {
	cpvs_ctx* ctx;
	u32* dag;
	cpvs_shadow_info info;
	// lookup shortcut over the top levels, built on the first lookup (see LookupDag::skip)
	std::mutex skipLock;
	u32* skip;
	u32 skipLevels;
};

struct ContainerCell {
	u32* words = nullptr;  // device copy owned by the container
	u64 count = 0;
	u32 numLevels = 0;
	int leafmasks = 0;
	u32 rootMask = 0;
	bool set = false;
};

struct cpvs_container {
	cpvs_ctx* ctx;
	u32 length;
	u32 filterSize;
	std::vector<ContainerCell> cells;
	u32* dag = nullptr;
	u32* grid = nullptr;
	u32* skip = nullptr;  // lookup shortcut over the top levels of every cell (see LookupDag::skip)
	u32 skipLevels = 0;
	u64 dagWords = 0;
	u32 dagLevels = 0, gridLevels = 0;
	int leafmasks = 0;
	bool finalized = false;
};

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

// Bump-pointer carving of the context arena; run once with base == nullptr to size it.
struct ArenaCarver {
	char* base;
	size_t offset = 0;
	explicit ArenaCarver(char* b) : base(b) {}
	template <typename T>
	T* take(u64 count) {
		offset = (offset + 255) & ~size_t(255);
		T* p = base ? reinterpret_cast<T*>(base + offset) : nullptr;
		offset += (count ? count : 1) * sizeof(T);
		return p;
	}
};

struct LevelArrays {
	u64 n = 0;
	u64* coords = nullptr;
	u16* masks = nullptr;
	u32* firstChild = nullptr;
	u32* uid = nullptr;
	u32* firstList = nullptr;
	u32* wordOffset = nullptr;
	u32* leafCodes = nullptr;
	u64* leafHash = nullptr;
	u32* leafAt = nullptr;  // leaf level, built per column: node index by column-order position
	u64* table = nullptr;      // merge table of this level (large levels own one; small levels share)
	u64 tableSlots = 0;
	u32* slotOffset = nullptr; // per table slot: word offset of the group's node
	unsigned char* sizeOf = nullptr;  // rank scratch, one byte per node (large levels)
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
	cpvs_ctx* ctx = new (std::nothrow) cpvs_ctx;
	if (!ctx) return fail(CPVS_ENOMEM, "cpvs_ctx_create: host allocation");
	ctx->device = device;
	ctx->launches = 0;
	ctx->arena = nullptr;
	ctx->arenaBytes = 0;
	ctx->scalars = nullptr;
	ctx->own = nullptr;
	cudaError_t e = cudaStreamCreateWithFlags(&ctx->own, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&ctx->scalars), 192 * sizeof(u64));
	ctx->hostScalars = nullptr;
	if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&ctx->hostScalars), 192 * sizeof(u64));
	ctx->aux = nullptr;
	ctx->evFork = ctx->evJoin = ctx->evAuxStart = nullptr;
	int prioLeast = 0, prioGreatest = 0;
	if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&prioLeast, &prioGreatest);
	if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->aux, cudaStreamNonBlocking, prioGreatest);
	if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming);
	if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->evJoin, cudaEventDisableTiming);
	if (e == cudaSuccess) e = cudaEventCreate(&ctx->evAuxStart);
	ctx->aux2 = ctx->aux3 = nullptr;
	ctx->evRankStart = ctx->evRankStop = ctx->evJoin3 = ctx->evClear = ctx->evCols = nullptr;
	{
		const char* v = std::getenv("CPVS_LEAF_COLUMNS");
		ctx->leafColumns = (v && v[0] >= '0' && v[0] <= '2') ? v[0] - '0' : 1;
		const char* x = std::getenv("CPVS_EXPERIMENTS");
		ctx->experiments = 0;
		if (x) {
			if (std::strstr(x, "expand-preload")) ctx->experiments |= kExperimentExpandPreload;
			if (std::strstr(x, "emit-gather")) ctx->experiments |= kExperimentEmitGather;
			if (std::strstr(x, "rank-preload")) ctx->experiments |= kExperimentRankPreload;
			if (std::strstr(x, "insert-witness")) ctx->experiments |= kExperimentInsertWitness;
			if (std::strstr(x, "early-bases")) ctx->experiments |= kExperimentEarlyBases;
			if (std::strstr(x, "leaf-fp64")) ctx->experiments |= kExperimentLeafFp64;
			if (std::strstr(x, "leaf-fp64-weak")) ctx->experiments |= kExperimentLeafFpWeak;
		}
	}
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->aux2, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->aux3, cudaStreamNonBlocking);
	ctx->aux4 = ctx->aux5 = nullptr;
	ctx->evBottomSized = ctx->evBottomRanked = ctx->evVerified = nullptr;
	ctx->forceExact = false;
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->aux5, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->evVerified, cudaEventDisableTiming);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->aux4, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->evBottomSized, cudaEventDisableTiming);
	if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->evBottomRanked, cudaEventDisableTiming);
	if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->evJoin3, cudaEventDisableTiming);
	if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->evClear, cudaEventDisableTiming);
	if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->evCols, cudaEventDisableTiming);
	if (e == cudaSuccess) e = cudaEventCreate(&ctx->evRankStart);
	if (e == cudaSuccess) e = cudaEventCreate(&ctx->evRankStop);
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
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	if (ctx->arena) cudaFreeAsync(ctx->arena, ctx->stream);
	cudaStreamSynchronize(ctx->stream);
	if (ctx->scalars) cudaFree(ctx->scalars);
	if (ctx->hostScalars) cudaFreeHost(ctx->hostScalars);
	if (ctx->aux) cudaStreamDestroy(ctx->aux);
	if (ctx->evFork) cudaEventDestroy(ctx->evFork);
	if (ctx->evJoin) cudaEventDestroy(ctx->evJoin);
	if (ctx->evAuxStart) cudaEventDestroy(ctx->evAuxStart);
	if (ctx->aux2) cudaStreamDestroy(ctx->aux2);
	if (ctx->aux3) cudaStreamDestroy(ctx->aux3);
	if (ctx->aux4) cudaStreamDestroy(ctx->aux4);
	if (ctx->aux5) cudaStreamDestroy(ctx->aux5);
	if (ctx->evVerified) cudaEventDestroy(ctx->evVerified);
	if (ctx->evBottomSized) cudaEventDestroy(ctx->evBottomSized);
	if (ctx->evBottomRanked) cudaEventDestroy(ctx->evBottomRanked);
	if (ctx->evJoin3) cudaEventDestroy(ctx->evJoin3);
	if (ctx->evClear) cudaEventDestroy(ctx->evClear);
	if (ctx->evCols) cudaEventDestroy(ctx->evCols);
	if (ctx->evRankStart) cudaEventDestroy(ctx->evRankStart);
	if (ctx->evRankStop) cudaEventDestroy(ctx->evRankStop);
	if (ctx->own) cudaStreamDestroy(ctx->own);
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

uint64_t cpvs_ctx_launch_count(const cpvs_ctx* ctx) { return ctx ? ctx->launches : 0; }

/* ---- MinMaxHierarchy ------------------------------------------------------------------------ */

int cpvs_minmax_build(cpvs_ctx* ctx, const float* depth, int n, int mem, cpvs_minmax** out) {
	if (!ctx || !depth || !out) return fail(CPVS_EINVAL, "cpvs_minmax_build: NULL argument");
	*out = nullptr;
	if (n < 2 || !isPow2((u64)n) || n > (1 << 19)) return fail(CPVS_EINVAL, "cpvs_minmax_build: side %d is not a power of two in [2, 2^19]", n);
	if (mem != CPVS_MEM_HOST && mem != CPVS_MEM_DEVICE) return fail(CPVS_EINVAL, "cpvs_minmax_build: mem %d", mem);
	CPVS_CUDA(cudaSetDevice(ctx->device));
	cpvs_minmax* mm = new (std::nothrow) cpvs_minmax();
	if (!mm) return fail(CPVS_ENOMEM, "cpvs_minmax_build: host allocation");
	mm->ownedDepth = nullptr;
	mm->levelStorage = nullptr;
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
	cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&mm->levelStorage), total * sizeof(float), ctx->stream);
	if (e == cudaSuccess && mem == CPVS_MEM_HOST) {
		e = cudaMallocAsync(reinterpret_cast<void**>(&mm->ownedDepth), (u64)n * n * sizeof(float), ctx->stream);
		if (e == cudaSuccess) e = cudaMemcpyAsync(mm->ownedDepth, depth, (u64)n * n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
	}
	if (e != cudaSuccess) {
		if (mm->levelStorage) cudaFreeAsync(mm->levelStorage, ctx->stream);
		if (mm->ownedDepth) cudaFreeAsync(mm->ownedDepth, ctx->stream);
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
	mm->lowLevelsBuilt = n < 128;  // small maps take the generic path, which writes every level
	ctx->launches += launchPyramid(mm->level[0], n, lv, levels, false, mm->evBase, ctx->stream);
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
	if (mm->levelStorage) cudaFreeAsync(mm->levelStorage, mm->ctx->stream);
	if (mm->ownedDepth) cudaFreeAsync(mm->ownedDepth, mm->ctx->stream);
	if (mm->evStart) {
		cudaEventDestroy(mm->evStart);
		cudaEventDestroy(mm->evBase);
		cudaEventDestroy(mm->evStop);
	}
	delete mm;
	return CPVS_OK;
}

}  // extern "C"

namespace {
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

// hasNodes << 32 | root mask of slice zTileIndex, computing the whole column on first use (one launch, one read-back).
int columnRoot(cpvs_ctx* ctx, const cpvs_minmax* cmm, const PyramidView& pyr, u32 zTileIndex, u32 zTileNum, u64* root) {
	cpvs_minmax* mm = const_cast<cpvs_minmax*>(cmm);
	std::lock_guard<std::mutex> guard(mm->lowLock);
	if (mm->columnSlices != zTileNum) {
		u64* dRoots = nullptr;
		CPVS_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&dRoots), zTileNum * sizeof(u64), ctx->stream));
		ctx->launches += launchColumnRoots(pyr, zTileNum, dRoots, ctx->stream);
		mm->columnRoots.assign(zTileNum, 0);
		cudaError_t e = cudaMemcpyAsync(mm->columnRoots.data(), dRoots, zTileNum * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
		cudaFreeAsync(dRoots, ctx->stream);
		if (e != cudaSuccess) return fail(CPVS_ECUDA, "column roots: %s", cudaGetErrorString(e));
		mm->columnSlices = zTileNum;
	}
	*root = mm->columnRoots[zTileIndex];
	return CPVS_OK;
}
}  // namespace

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
	PyramidView pyr;
	pyr.n = mm->n;
	pyr.numLevels = mm->numLevels;
	for (int k = 0; k < kMaxLevels; ++k) pyr.level[k] = k < mm->numLevels ? mm->level[k] : nullptr;
	u32* dOut = reinterpret_cast<u32*>(ctx->scalars + 191);
	ctx->launches += launchChildmask(pyr, (int)level, zTileNum, x, y, z, dOut, ctx->stream);
	CPVS_CUDA(cudaMemcpyAsync(out, dOut, sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
	CPVS_CUDA(cudaStreamSynchronize(ctx->stream));
	return CPVS_OK;
}

/* ---- CompressedShadow::create ------------------------------------------------------------------ */

int cpvs_shadow_create(cpvs_ctx* ctx, const cpvs_minmax* mm, uint32_t zTileIndex, uint32_t zTileNum, int leafmasks, cpvs_shadow** out) {
	if (!ctx || !mm || !out) return fail(CPVS_EINVAL, "cpvs_shadow_create: NULL argument");
	*out = nullptr;
	if (mm->ctx->device != ctx->device) return fail(CPVS_EINVAL, "cpvs_shadow_create: hierarchy lives on device %d, context on %d", mm->ctx->device, ctx->device);
	const int L = mm->numLevels;
	if (L <= 3) return fail(CPVS_EINVAL, "cpvs_shadow_create: needs more than 3 levels (side >= 8), got %d", L);  // src/CompressedShadow.cpp:46
	if (zTileNum == 0 || zTileIndex >= zTileNum) return fail(CPVS_EINVAL, "cpvs_shadow_create: z tile %u of %u", zTileIndex, zTileNum);
	if ((u64)mm->n * zTileNum > (1ull << 23))
		return fail(CPVS_EINVAL, "cpvs_shadow_create: side * zTileNum = %llu exceeds 2^23 (depth slices must stay exact in fp32)",
				(unsigned long long)((u64)mm->n * zTileNum));
	CPVS_CUDA(cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	// a hierarchy built by another context (createShadowTiles: one pyramid, one builder per z-slice) may
	// still be in flight on that context's stream
	if (mm->ctx != ctx && mm->evStop) CPVS_CUDA(cudaStreamWaitEvent(st, mm->evStop, 0));

	const int top = L - 2;
	const bool useLeaf = leafmasks && (L - 3) >= 2;  // src/CompressedShadow.cpp:20-27
	const int minLevel = useLeaf ? 2 : 0;            // src/CompressedShadow.cpp:30-32
	const int lastInner = useLeaf ? 3 : 0;
	if (!useLeaf)
		if (int rc = ensureLowLevels(mm, 1)) return rc;  // the leafmask-less octree descends through levels 2 and 1

	// One slice of a column (createShadowTiles, reference src/DeferredRenderer.cpp:150-163): most slices of a tall
	// grid miss the surface; their DAG is the root's mask word, known for the whole column after one launch.
	if (zTileNum > 1 && top - 1 >= minLevel) {
		PyramidView pyrTop;
		pyrTop.n = mm->n;
		pyrTop.numLevels = L;
		for (int k = 0; k < kMaxLevels; ++k) pyrTop.level[k] = k < L ? mm->level[k] : nullptr;
		u64 root = 0;
		if (int rc = columnRoot(ctx, mm, pyrTop, zTileIndex, zTileNum, &root)) return rc;
		if (!(root >> 32)) {
			cpvs_shadow* s = new (std::nothrow) cpvs_shadow();
			if (!s) return fail(CPVS_ENOMEM, "cpvs_shadow_create: host allocation");
			std::memset(&s->info, 0, sizeof(s->info));
			s->skip = nullptr;
			s->skipLevels = 0;
			s->ctx = ctx;
			s->dag = nullptr;
			const u32 rootMask = (u32)root;
			cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&s->dag), sizeof(u32), st);
			if (e != cudaSuccess) {
				delete s;
				return fail(e == cudaErrorMemoryAllocation ? CPVS_ENOMEM : CPVS_ECUDA, "cpvs_shadow_create: %s", cudaGetErrorString(e));
			}
			storeU32Kernel<<<1, 1, 0, st>>>(s->dag, rootMask);
			++ctx->launches;
			s->info.num_levels = (u32)L;
			s->info.leafmasks = useLeaf ? 1 : 0;
			s->info.total_visibility = rootMask == 0x5555u ? CPVS_VISIBLE : (rootMask == 0u ? CPVS_SHADOW : CPVS_PARTIAL);
			s->info.words = 1;
			s->info.svo_nodes[top] = s->info.dag_nodes[top] = s->info.dag_words[top] = 1;
			*out = s;
			return CPVS_OK;
		}
	}

	// ev[i] opens phase i (CPVS_PHASE_*), ev[CPVS_NUM_PHASES] closes the last one
	struct PhaseEvents {
		cudaEvent_t ev[CPVS_NUM_PHASES + 1];
		PhaseEvents() { std::memset(ev, 0, sizeof(ev)); }
		~PhaseEvents() {
			for (cudaEvent_t e : ev)
				if (e) cudaEventDestroy(e);
		}
	} phases;
	for (cudaEvent_t& e : phases.ev) CPVS_CUDA(cudaEventCreate(&e));
	CPVS_CUDA(cudaEventRecord(phases.ev[CPVS_PHASE_COUNT], st));

	std::unique_lock<std::mutex> buildGuard(ctx->buildLock);
	PyramidView pyr;
	pyr.n = mm->n;
	pyr.numLevels = L;
	for (int k = 0; k < kMaxLevels; ++k) pyr.level[k] = k < L ? mm->level[k] : nullptr;

	// device scalars: [0..31] SVO counts, [32..63] unique, [64..95] words, [96..127] bases, [128..159] child totals, [160] total words
	u64* dScalars = ctx->scalars;
	CPVS_CUDA(cudaMemsetAsync(dScalars, 0, 192 * sizeof(u64), st));
	u64 *dCounts = dScalars, *dUnique = dScalars + 32, *dWords = dScalars + 64, *dBases = dScalars + 96, *dChildTotal = dScalars + 128,
		*dTotal = dScalars + 160;

	HostTrace trace;
	trace.mark("setup");
	// 1. exact node counts of every level (closed form), so all buffers can be sized up front
	ctx->launches += launchCountNodes(pyr, zTileIndex, zTileNum, minLevel, dCounts, st);
	u64* hScalars = ctx->hostScalars;
	CPVS_CUDA(cudaMemcpyAsync(hScalars, dCounts, 32 * sizeof(u64), cudaMemcpyDeviceToHost, st));
	CPVS_CUDA(cudaEventRecord(phases.ev[CPVS_NUM_PHASES], st));  // end of a slice that stops here; re-recorded at the end otherwise
	CPVS_CUDA(cudaStreamSynchronize(st));
	trace.mark("count kernels + sync");
	LevelArrays lv[kMaxLevels];
	lv[top].n = 1;
	for (int l = top - 1; l >= minLevel; --l) lv[l].n = lv[l + 1].n ? hScalars[l] : 0;
	for (int l = minLevel; l <= top; ++l)
		if (lv[l].n >= (1ull << 30)) return fail(CPVS_EOVERFLOW, "level %d has %llu nodes (limit 2^30)", l, (unsigned long long)lv[l].n);

	// A z-slice that misses the surface altogether (most slices of a tall tile grid): the root has no PARTIAL
	// child, the DAG is its one mask word (0x5555 lit / 0x0000 shadow). One tiny kernel instead of the pipeline.
	if (top - 1 >= minLevel && lv[top - 1].n == 0) {
		cpvs_shadow* s = new (std::nothrow) cpvs_shadow();
		if (!s) return fail(CPVS_ENOMEM, "cpvs_shadow_create: host allocation");
		std::memset(&s->info, 0, sizeof(s->info));
		s->skip = nullptr;
		s->skipLevels = 0;
		s->ctx = ctx;
		s->dag = nullptr;
		// The count launch already reported the root's mask (kRootMaskScalar) and its read-back is complete: store
		// the word and return without another round trip. build_ms ends at the read-back.
		cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&s->dag), sizeof(u32), st);
		u32 rootMask = 0;
		u32* hMask = &rootMask;
		float ms = 0.f;
		if (e == cudaSuccess && (hScalars[kRootMaskScalar] >> 32) == 1ull) {
			rootMask = (u32)hScalars[kRootMaskScalar];
			storeU32Kernel<<<1, 1, 0, st>>>(s->dag, rootMask);
			++ctx->launches;
			e = cudaEventElapsedTime(&ms, phases.ev[CPVS_PHASE_COUNT], phases.ev[CPVS_NUM_PHASES]);
		} else if (e == cudaSuccess) {  // no level was counted (tiny maps): ask for the mask
			ctx->launches += launchChildmask(pyr, top, zTileNum, 0, 0, zTileIndex * 2, s->dag, st);
			e = cudaMemcpyAsync(hScalars + 190, s->dag, sizeof(u32), cudaMemcpyDeviceToHost, st);
			if (e == cudaSuccess) e = cudaEventRecord(phases.ev[CPVS_NUM_PHASES], st);
			if (e == cudaSuccess) e = cudaStreamSynchronize(st);
			rootMask = *reinterpret_cast<u32*>(hScalars + 190);
			if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, phases.ev[CPVS_PHASE_COUNT], phases.ev[CPVS_NUM_PHASES]);
		}
		if (e != cudaSuccess) {
			if (s->dag) cudaFreeAsync(s->dag, st);
			delete s;
			return fail(CPVS_ECUDA, "cpvs_shadow_create: %s", cudaGetErrorString(e));
		}
		s->info.num_levels = (u32)L;
		s->info.leafmasks = useLeaf ? 1 : 0;
		s->info.total_visibility = *hMask == 0x5555u ? CPVS_VISIBLE : (*hMask == 0u ? CPVS_SHADOW : CPVS_PARTIAL);
		s->info.words = 1;
		s->info.svo_nodes[top] = s->info.dag_nodes[top] = s->info.dag_words[top] = 1;
		s->info.build_ms = ms;
		s->info.phase_ms[CPVS_PHASE_COUNT] = ms;
		*out = s;
		return CPVS_OK;
	}

	// The top levels up to kSmallMaxNodes nodes each ("small": smallLow..top) are handled by single-CTA
	// kernels, one launch per phase instead of one or more per level.
	int smallLow = top + 1;
	for (int l = top; l >= lastInner && lv[l].n && lv[l].n <= kSmallMaxNodes; --l) smallLow = l;

	// 2. per-level arrays, carved out of the context arena
	u64 scanTiles = 0, scanLaunches = 0;
	const u64 maxTable = 2 * kSmallMaxNodes;  // shared by the small levels; large levels own their tables
	for (int l = top; l >= minLevel; --l) {
		const u64 n = lv[l].n;
		if (!n || l >= smallLow) continue;
		if (!(useLeaf && l == 2)) {
			scanTiles += (n + kExpandTileNodes - 1) / kExpandTileNodes;
			++scanLaunches;
		}
		if (n > 1) {
			scanTiles += (n + kScanTile - 1) / kScanTile;
			++scanLaunches;
		}
	}
	ScanTileState* dTiles = nullptr;
	u32* dTickets = nullptr;
	u64* dTable = nullptr;
	u32* dSketch = nullptr;
	const bool needSketch = useLeaf && lv[2].n > 0, haveLeaves = useLeaf && lv[2].n > 1;
	// leaves per column: one more scan (over the columns = texels of pyramid level 3)
	// Where it pays: whole-volume builds with 2..8 leaves per column of a depth map that does not fit in L2 (terrain-like
	// surfaces; measured at 16K^2: 0.42 ms against 0.50 ms per leaf, at 8192^2 0.121 against 0.137, at 4096^2 -- 64 MiB, L2
	// serves the re-reads -- 0.046 against 0.045). A z-slice of a tall grid leaves most columns empty; box edges make columns
	// of hundreds of leaves that a four-lane group walks alone (16K^2 city: 6.1 ms against 2.9 ms); a gentle plane has one
	// leaf per column and nothing to share (0.30 ms against 0.27 ms): those keep the per-leaf kernel.
	const u64 allCols = ((u64)mm->n >> 3) * ((u64)mm->n >> 3);
	const bool leafColumns = needSketch && (ctx->leafColumns == 2 || (ctx->leafColumns == 1 && zTileNum == 1 && mm->n >= 8192 &&
																   lv[2].n >= 2 * allCols && lv[2].n <= 8 * allCols));
	const u64 numCols = leafColumns ? ((u64)mm->n >> 3) * ((u64)mm->n >> 3) : 0;
	u32* dColBias = nullptr;
	if (leafColumns) {
		scanTiles += (numCols + kScanTile - 1) / kScanTile;
		++scanLaunches;
	}
	auto carve = [&](ArenaCarver& ar) {
		dTiles = ar.take<ScanTileState>(scanTiles);
		dTickets = ar.take<u32>(scanLaunches);
		dSketch = ar.take<u32>(needSketch ? kSketchWords : 0);
		dTable = ar.take<u64>(maxTable);
		dColBias = ar.take<u32>(numCols);
		for (int l = top; l >= minLevel; --l) {
			LevelArrays& a = lv[l];
			if (!a.n) continue;
			if (leafColumns && l == 2)
				a.leafAt = ar.take<u32>(a.n);
			else
				a.coords = ar.take<u64>(a.n);
			a.masks = ar.take<u16>(a.n);
			a.uid = ar.take<u32>(a.n);
			a.firstList = ar.take<u32>(a.n);
			a.wordOffset = ar.take<u32>(a.n);
			if (l >= smallLow) {
				a.table = dTable;
				a.tableSlots = 2 * kSmallMaxNodes;
			} else {
				a.tableSlots = pow2AtLeast(a.n * 2 < 1024 ? 1024 : a.n * 2);
				a.table = ar.take<u64>(a.tableSlots + kDirectSlots);
			}
			a.slotOffset = ar.take<u32>(a.tableSlots + kDirectSlots);
			if (l < smallLow) a.sizeOf = ar.take<unsigned char>(a.n + 4);
			if (useLeaf && l == 2) {
				a.leafCodes = ar.take<u32>(a.n * 8);
				if (!leafColumns) a.leafHash = ar.take<u64>(a.n);
			} else {
				a.firstChild = ar.take<u32>(a.n);
			}
		}
	};
	ArenaCarver sizing(nullptr);
	carve(sizing);
	if (sizing.offset > ctx->arenaBytes) {
		// Grow geometrically (a tile grid feeds builds of slowly increasing size) and from the stream-ordered
		// pool: a regrowth served from memory the pool already holds (cpvs_ctx_reserve, earlier frees) costs
		// microseconds, where cudaFree + cudaMalloc synchronise the device and, with peer access enabled by a
		// communication library, remap on every GPU (100+ ms). The previous build has completed on all streams.
		const size_t doubled = ctx->arenaBytes * 2;
		if (ctx->arena) CPVS_CUDA(cudaFreeAsync(ctx->arena, st));
		ctx->arena = nullptr;
		ctx->arenaBytes = 0;
		size_t want = sizing.offset + sizing.offset / 4;
		if (want < doubled) want = doubled;
		size_t freeBytes = 0, totalBytes = 0;
		if (cudaMemGetInfo(&freeBytes, &totalBytes) == cudaSuccess && want > freeBytes / 2) want = sizing.offset + sizing.offset / 8;
		cudaError_t ae = cudaMallocAsync(reinterpret_cast<void**>(&ctx->arena), want, st);
		if (ae != cudaSuccess) return fail(CPVS_ENOMEM, "scratch arena of %zu bytes: %s", want, cudaGetErrorString(ae));
		ctx->arenaBytes = want;
	}
	ArenaCarver real(ctx->arena);
	carve(real);
	trace.mark("arena carve");
	// tile states, tickets and the leaf sketch sit at the front of the arena: one memset clears them all
	CPVS_CUDA(cudaMemsetAsync(ctx->arena, 0, reinterpret_cast<char*>(dTable) - ctx->arena, st));
	// the large inner levels' tables are cleared up front (the leaf level sizes and clears its table on
	// the device, once the sketch is filled; the small levels clear theirs inside their kernel)
	// -- on a side stream, next to the expansion; the first inner insert waits for it.
	if (leafColumns) {
		CPVS_CUDA(cudaEventRecord(ctx->evFork, st));
		CPVS_CUDA(cudaStreamWaitEvent(ctx->aux2, ctx->evFork, 0));
		ScanLaunch colScan{dTickets + scanLaunches - 1, dTiles + scanTiles - (numCols + kScanTile - 1) / kScanTile};
		ctx->launches += launchColumnBias(pyr, zTileIndex, zTileNum, dColBias, colScan, ctx->aux2);
		CPVS_CUDA(cudaEventRecord(ctx->evCols, ctx->aux2));
	}
	bool tablesClearing = false;
	for (int l = minLevel; l < smallLow; ++l)
		if (lv[l].n > 1 && !(useLeaf && l == 2)) {
			if (!tablesClearing) {
				CPVS_CUDA(cudaEventRecord(ctx->evFork, st));  // the arena may still be in use by the previous build
				CPVS_CUDA(cudaStreamWaitEvent(ctx->aux3, ctx->evFork, 0));
				tablesClearing = true;
			}
			CPVS_CUDA(cudaMemsetAsync(lv[l].table, 0xFF, (lv[l].tableSlots + kDirectSlots) * sizeof(u64), ctx->aux3));
		}
	if (tablesClearing) CPVS_CUDA(cudaEventRecord(ctx->evClear, ctx->aux3));
	u64 *dSketchBits = dScalars + 161, *dLeafTableMask = dScalars + 162;
	u32* dErrorFlag = reinterpret_cast<u32*>(dScalars + 163);
	u32* dMismatchFlag = reinterpret_cast<u32*>(dScalars + 164);  // leaf-fp64: a fingerprint group held two different leaves
	const bool leafFingerprint = haveLeaves && (ctx->experiments & kExperimentLeafFp64) && !ctx->forceExact;
	bool verifyPending = false;
	u64 tileCursor = 0, launchCursor = 0;
	auto nextScan = [&](u64 n, u64 tileNodes = kScanTile) {
		ScanLaunch s{dTickets + launchCursor, dTiles + tileCursor};
		++launchCursor;
		tileCursor += (n + tileNodes - 1) / tileNodes;
		return s;
	};

	trace.mark("memsets");
	// 3. breadth-first construction, top level first (src/CompressedShadow.cpp:87-169)
	CPVS_CUDA(cudaEventRecord(phases.ev[CPVS_PHASE_EXPAND], st));
	storeU64Kernel<<<1, 1, 0, st>>>(lv[top].coords, packCoord(0, 0, zTileIndex * 2));
	++ctx->launches;
	{
		SmallExpandArgs sx;
		sx.count = 0;
		for (int l = top; l >= smallLow; --l) {
			SmallExpandLevel& e = sx.lv[sx.count++];
			e.side = (u32)mm->n >> l;
			e.tex = pyr.level[l];
			e.heightF = (float)(e.side * zTileNum);
			e.level0 = l == 0 ? 1 : 0;
			e.coords = lv[l].coords;
			e.masks = lv[l].masks;
			e.firstChild = lv[l].firstChild;
			e.childCoords = (l > minLevel && lv[l - 1].n) ? lv[l - 1].coords : nullptr;
			e.childTotal = dChildTotal + l;
			e.colBias = nullptr;
			e.leafAt = nullptr;
			e.numLeaves = 0;
			if (leafColumns && l == 3) {
				e.numLeaves = (u32)lv[2].n;
				e.colBias = dColBias;
				e.leafAt = lv[2].leafAt;
				CPVS_CUDA(cudaStreamWaitEvent(st, ctx->evCols, 0));
			}
		}
		ctx->launches += launchExpandSmallLevels(sx, st);
	}
	for (int l = smallLow - 1; l >= lastInner && lv[l].n; --l) {
		u64* childCoords = (l > minLevel && lv[l - 1].n) ? lv[l - 1].coords : nullptr;
		const bool toColumns = leafColumns && l == 3;
		if (toColumns) CPVS_CUDA(cudaStreamWaitEvent(st, ctx->evCols, 0));
		ctx->launches += launchExpandLevel(pyr, l, zTileNum, lv[l].coords, lv[l].n, lv[l].masks, lv[l].firstChild, childCoords,
				dChildTotal + l, nextScan(lv[l].n, kExpandTileNodes), toColumns ? dColBias : nullptr, toColumns ? lv[2].leafAt : nullptr, (u32)lv[2].n,
				(ctx->experiments & kExperimentExpandPreload) ? 1 : 0, st);
	}
	CPVS_CUDA(cudaEventRecord(phases.ev[CPVS_PHASE_LEAVES], st));
	if (leafColumns)  // constructLastLevels (src/CompressedShadow.cpp:171-190)
		ctx->launches += launchBuildLeafColumns(pyr, zTileIndex, zTileNum, dColBias, lv[2].leafAt, (u32)lv[2].n, lv[2].leafCodes, lv[2].masks,
				dSketch, st);
	else if (useLeaf && lv[2].n)
		ctx->launches += launchBuildLeaves(pyr, zTileNum, lv[2].coords, lv[2].n, lv[2].leafCodes, lv[2].leafHash, lv[2].masks,
				dSketch, st);

	// 4. bottom-up merge (src/CompressedShadow.cpp:215-241). The chain of inserts is the critical path and
	// runs on the high-priority stream `ms`, so that its CTAs are dispatched ahead of the queued CTAs of
	// the rank scans running beside it; the main stream rejoins before the bases.
	cudaStream_t mergeStream = ctx->aux;
	CPVS_CUDA(cudaEventRecord(ctx->evFork, st));
	CPVS_CUDA(cudaStreamWaitEvent(mergeStream, ctx->evFork, 0));
	CPVS_CUDA(cudaEventRecord(phases.ev[CPVS_PHASE_LEAF_TABLE], mergeStream));
	if (haveLeaves) {
		ctx->launches += launchSketchPopcount(dSketch, dSketchBits, mergeStream);
		ctx->launches += launchSizeLeafTable(lv[2].table, lv[2].tableSlots, dSketchBits, dLeafTableMask, leafFingerprint ? 1 : 0, mergeStream);
	}
	CPVS_CUDA(cudaEventRecord(phases.ev[CPVS_PHASE_LEAF_INSERT], mergeStream));
	if (!(useLeaf && lv[2].n)) {
		CPVS_CUDA(cudaEventRecord(phases.ev[CPVS_PHASE_LEAF_RESOLVE], mergeStream));
		CPVS_CUDA(cudaEventRecord(phases.ev[CPVS_PHASE_INNER_MERGE], mergeStream));
	}
	// Per level: insert on the main stream (gives every node its group id, all the next level needs),
	// rank on the side stream (orders the unique nodes; only the emission needs it).
	bool ranksPending = false;
	bool bottomRankSplit = false;  // early-bases: the bottom level's rank is on aux4, its write kernel not joined before the bases
	for (int l = minLevel; l < smallLow; ++l) {
		LevelArrays& a = lv[l];
		if (!a.n) continue;
		const bool leafLevel = useLeaf && l == 2;
		MergeLevelArgs m;
		m.n = a.n;
		m.leaf = leafLevel ? 1 : 0;
		m.leafCodes = a.leafCodes;
		m.leafHash = a.leafHash;
		m.masks = a.masks;
		m.firstChild = a.firstChild;
		m.childUid = l > minLevel ? lv[l - 1].uid : nullptr;
		m.table = a.table;
		m.tableSize = a.tableSlots;
		m.sketchBits = dSketchBits;
		m.tableMaskDev = dLeafTableMask;
		m.errorFlag = dErrorFlag;
		m.uid = a.uid;
		m.firstList = a.firstList;
		m.wordOffset = a.wordOffset;
		m.slotOffset = a.slotOffset;
		m.sizeOf = a.sizeOf;
		m.uniqueCount = dUnique + l;
		m.wordCount = dWords + l;
		m.rankPreload = (ctx->experiments & kExperimentRankPreload) ? 1 : 0;
		m.parallelWitness = (ctx->experiments & kExperimentInsertWitness) ? 1 : 0;
		m.fingerprint = (leafLevel && leafFingerprint) ? ((ctx->experiments & kExperimentLeafFpWeak) ? 2 : 1) : 0;
		if (!leafLevel && tablesClearing) {
			CPVS_CUDA(cudaStreamWaitEvent(mergeStream, ctx->evClear, 0));
			tablesClearing = false;
		}
		ctx->launches += launchInsertLevel(m, mergeStream);
		if (leafLevel) {
			CPVS_CUDA(cudaEventRecord(phases.ev[CPVS_PHASE_LEAF_RESOLVE], mergeStream));
			CPVS_CUDA(cudaEventRecord(phases.ev[CPVS_PHASE_INNER_MERGE], mergeStream));
			if (m.fingerprint) {  // the exact compare, off the critical path
				CPVS_CUDA(cudaEventRecord(ctx->evFork, mergeStream));
				CPVS_CUDA(cudaStreamWaitEvent(ctx->aux5, ctx->evFork, 0));
				ctx->launches += launchVerifyLeafGroups(m, dMismatchFlag, ctx->aux5);
				CPVS_CUDA(cudaEventRecord(ctx->evVerified, ctx->aux5));
				verifyPending = true;
			}
		}
		if (a.n > 1) {
			const bool split = (ctx->experiments & kExperimentEarlyBases) && l == minLevel;
			cudaStream_t rs = split ? ctx->aux4 : ((l & 1) ? ctx->aux3 : ctx->aux2);
			CPVS_CUDA(cudaEventRecord(ctx->evFork, mergeStream));
			CPVS_CUDA(cudaStreamWaitEvent(rs, ctx->evFork, 0));
			if (leafLevel) CPVS_CUDA(cudaEventRecord(ctx->evRankStart, rs));
			ctx->launches += launchRankLevel(m, nextScan(a.n), rs, split ? ctx->evBottomSized : nullptr);
			if (leafLevel) CPVS_CUDA(cudaEventRecord(ctx->evRankStop, rs));
			if (split) {
				CPVS_CUDA(cudaEventRecord(ctx->evBottomRanked, rs));
				bottomRankSplit = true;
			}
			ranksPending = true;
		}
	}
	{
		SmallMergeArgs sm;
		sm.count = 0;
		sm.table = dTable;
		sm.errorFlag = dErrorFlag;
		for (int l = smallLow; l <= top; ++l) {
			if (!lv[l].n) continue;
			SmallMergeLevel& m = sm.lv[sm.count++];
			m.n = (u32)lv[l].n;
			m.masks = lv[l].masks;
			m.firstChild = lv[l].firstChild;
			m.childUid = l > minLevel ? lv[l - 1].uid : nullptr;
			m.uid = lv[l].uid;
			m.firstList = lv[l].firstList;
			m.wordOffset = lv[l].wordOffset;
			m.slotOffset = lv[l].slotOffset;
			m.uniqueCount = dUnique + l;
			m.wordCount = dWords + l;
		}
		ctx->launches += launchMergeSmallLevels(sm, mergeStream);
	}
	if (ranksPending || tablesClearing) {  // join: the level sizes feed the bases
		CPVS_CUDA(cudaEventRecord(ctx->evJoin, ctx->aux2));
		CPVS_CUDA(cudaStreamWaitEvent(mergeStream, ctx->evJoin, 0));
		CPVS_CUDA(cudaEventRecord(ctx->evJoin3, ctx->aux3));
		CPVS_CUDA(cudaStreamWaitEvent(mergeStream, ctx->evJoin3, 0));
		if (bottomRankSplit) CPVS_CUDA(cudaStreamWaitEvent(mergeStream, ctx->evBottomSized, 0));  // its sizes, not its writes
	}
	if (verifyPending) CPVS_CUDA(cudaStreamWaitEvent(mergeStream, ctx->evVerified, 0));  // its verdict is read back with the sizes
	CPVS_CUDA(cudaEventRecord(ctx->evJoin, mergeStream));
	CPVS_CUDA(cudaStreamWaitEvent(st, ctx->evJoin, 0));
	CPVS_CUDA(cudaEventRecord(phases.ev[CPVS_PHASE_BASES], st));

	// 5. level bases and the total size (src/CompressedShadow.cpp:326-392)
	ctx->launches += launchLevelBases(dWords, dBases, top, minLevel, dTotal, st);
	trace.mark("enqueue expand..bases");
	CPVS_CUDA(cudaMemcpyAsync(hScalars, dScalars, 192 * sizeof(u64), cudaMemcpyDeviceToHost, st));
	CPVS_CUDA(cudaStreamSynchronize(st));
	trace.mark("sync after bases");
	CPVS_CUDA(cudaGetLastError());
	for (int l = top; l > minLevel; --l)
		if (lv[l].n && l >= lastInner && hScalars[128 + l] != lv[l - 1].n)
			return fail(CPVS_EINTERNAL, "level %d: expansion produced %llu nodes, count pass predicted %llu", l - 1,
					(unsigned long long)hScalars[128 + l], (unsigned long long)lv[l - 1].n);
	if (hScalars[163] != 0) return fail(CPVS_EINTERNAL, "merge table overflow (leaf table mask %llu)", (unsigned long long)hScalars[162]);
	if (verifyPending && (u32)hScalars[164] != 0) {
		// two different leaves shared a 64-bit fingerprint: nothing has been emitted yet -- once more, with the exact insert
		if (bottomRankSplit) CPVS_CUDA(cudaStreamSynchronize(ctx->aux4));  // the rank's last kernel still uses the arena
		ctx->forceExact = true;
		buildGuard.unlock();
		const int rc = cpvs_shadow_create(ctx, mm, zTileIndex, zTileNum, leafmasks, out);
		ctx->forceExact = false;
		return rc;
	}
	const u64 totalWords = hScalars[160];
	if (totalWords > (1ull << 32)) return fail(CPVS_EOVERFLOW, "DAG needs %llu words; offsets are 32-bit", (unsigned long long)totalWords);

	cpvs_shadow* s = new (std::nothrow) cpvs_shadow();
	if (!s) return fail(CPVS_ENOMEM, "cpvs_shadow_create: host allocation");
	std::memset(&s->info, 0, sizeof(s->info));
	s->dag = nullptr;
	s->skip = nullptr;
	s->skipLevels = 0;
	s->ctx = ctx;
	cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&s->dag), totalWords * sizeof(u32), st);
	if (e != cudaSuccess) {
		delete s;
		return fail(CPVS_ENOMEM, "DAG allocation of %llu words: %s", (unsigned long long)totalWords, cudaGetErrorString(e));
	}

	trace.mark("dag alloc");
	// 6. write every unique node once, in its final place. The levels are independent of each other
	// now; the leaf level (most of the words) stays on the main stream, the chain of inner levels runs on
	// the high-priority side stream next to it.
	// phase EMIT_INNER: fork .. join on the main stream; phase EMIT_LEAVES: the leaf kernel alone (they overlap).
	if (bottomRankSplit) cudaStreamWaitEvent(st, ctx->evBottomRanked, 0);  // the emission reads what the rank's last kernel wrote
	cudaEventRecord(phases.ev[CPVS_PHASE_EMIT_INNER], st);
	const bool leafEmit = useLeaf && lv[2].n;
	if (leafEmit) {
		cudaEventRecord(ctx->evFork, st);
		cudaStreamWaitEvent(ctx->aux, ctx->evFork, 0);
	}
	EmitMultiArgs inner;
	inner.count = 0;
	inner.gather = (ctx->experiments & kExperimentEmitGather) ? 1 : 0;
	for (int l = minLevel; l <= top; ++l) {
		const LevelArrays& a = lv[l];
		if (!a.n) continue;
		const bool isLeaf = useLeaf && l == 2;
		EmitLevelArgs em;
		em.n = hScalars[32 + l];  // unique nodes of the level (read back with the sizes): exact grid
		em.leaf = isLeaf ? 1 : 0;
		em.uniqueCount = dUnique + l;
		em.wordCount = dWords + l;
		em.firstList = a.firstList;
		em.wordOffset = a.wordOffset;
		em.levelBase = dBases + l;
		em.leafCodes = a.leafCodes;
		em.masks = a.masks;
		em.firstChild = a.firstChild;
		em.childUid = l > minLevel ? lv[l - 1].uid : nullptr;
		em.childSlotOffset = l > minLevel ? lv[l - 1].slotOffset : nullptr;
		em.childLevelBase = dBases + (l > minLevel ? l - 1 : l);
		em.dag = s->dag;
		if (isLeaf) {
			cudaEventRecord(ctx->evAuxStart, st);
			ctx->launches += launchEmitLevel(em, st);
			cudaEventRecord(phases.ev[CPVS_PHASE_EMIT_LEAVES], st);  // closes the leaf kernel
		} else if (inner.count < kMaxEmitLevels) {
			inner.lv[inner.count++] = em;
		}
	}
	ctx->launches += launchEmitInnerLevels(inner, leafEmit ? ctx->aux : st);
	if (leafEmit) {
		cudaEventRecord(ctx->evJoin, ctx->aux);
		cudaStreamWaitEvent(st, ctx->evJoin, 0);
	} else {
		cudaEventRecord(phases.ev[CPVS_PHASE_EMIT_LEAVES], st);
	}
	u32 rootMask = 0;
	e = cudaMemcpyAsync(&rootMask, s->dag, sizeof(u32), cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess) e = cudaEventRecord(phases.ev[CPVS_NUM_PHASES], st);
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	if (e == cudaSuccess) e = cudaGetLastError();
	trace.mark("emit + final sync");
	float ms = 0.f, phaseMs[CPVS_NUM_PHASES] = {0};
	if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, phases.ev[0], phases.ev[CPVS_NUM_PHASES]);
	for (int i = 0; i < CPVS_PHASE_EMIT_INNER && e == cudaSuccess; ++i) e = cudaEventElapsedTime(&phaseMs[i], phases.ev[i], phases.ev[i + 1]);
	if (e == cudaSuccess) e = cudaEventElapsedTime(&phaseMs[CPVS_PHASE_EMIT_INNER], phases.ev[CPVS_PHASE_EMIT_INNER], phases.ev[CPVS_NUM_PHASES]);
	if (e == cudaSuccess && leafEmit) e = cudaEventElapsedTime(&phaseMs[CPVS_PHASE_EMIT_LEAVES], ctx->evAuxStart, phases.ev[CPVS_PHASE_EMIT_LEAVES]);
	if (e == cudaSuccess && haveLeaves) e = cudaEventElapsedTime(&phaseMs[CPVS_PHASE_LEAF_RESOLVE], ctx->evRankStart, ctx->evRankStop);
	if (e != cudaSuccess) {
		cudaFreeAsync(s->dag, st);
		delete s;
		return fail(CPVS_ECUDA, "cpvs_shadow_create: %s", cudaGetErrorString(e));
	}
	s->info.num_levels = (u32)L;
	s->info.leafmasks = useLeaf ? 1 : 0;
	s->info.total_visibility = rootMask == 0x5555u ? CPVS_VISIBLE : (rootMask == 0u ? CPVS_SHADOW : CPVS_PARTIAL);
	s->info.words = totalWords;
	for (int l = minLevel; l <= top; ++l) {
		s->info.svo_nodes[l] = lv[l].n;
		s->info.dag_nodes[l] = hScalars[32 + l];
		s->info.dag_words[l] = hScalars[64 + l];
	}
	s->info.build_ms = ms;
	for (int i = 0; i < CPVS_NUM_PHASES; ++i) s->info.phase_ms[i] = phaseMs[i];
	*out = s;
	return CPVS_OK;
}

int cpvs_shadow_create_from_depth(cpvs_ctx* ctx, const float* depth, int n, int mem, uint32_t zTileIndex, uint32_t zTileNum, int leafmasks,
		cpvs_shadow** out) {
	cpvs_minmax* mm = nullptr;
	int rc = cpvs_minmax_build(ctx, depth, n, mem, &mm);
	if (rc != CPVS_OK) return rc;
	rc = cpvs_shadow_create(ctx, mm, zTileIndex, zTileNum, leafmasks, out);
	cpvs_minmax_destroy(mm);
	return rc;
}

int cpvs_shadow_destroy(cpvs_shadow* s) {
	if (!s) return CPVS_OK;
	cudaSetDevice(s->ctx->device);
	if (s->dag) cudaFreeAsync(s->dag, s->ctx->stream);
	if (s->skip) cudaFreeAsync(s->skip, s->ctx->stream);
	delete s;
	return CPVS_OK;
}

int cpvs_shadow_info_get(const cpvs_shadow* s, cpvs_shadow_info* info) {
	if (!s || !info) return fail(CPVS_EINVAL, "cpvs_shadow_info_get: NULL argument");
	*info = s->info;
	return CPVS_OK;
}

int cpvs_shadow_copy_dag(const cpvs_shadow* s, uint32_t* out_host) {
	if (!s || !out_host) return fail(CPVS_EINVAL, "cpvs_shadow_copy_dag: NULL argument");
	CPVS_CUDA(cudaSetDevice(s->ctx->device));
	CPVS_CUDA(cudaMemcpyAsync(out_host, s->dag, s->info.words * sizeof(u32), cudaMemcpyDeviceToHost, s->ctx->stream));
	CPVS_CUDA(cudaStreamSynchronize(s->ctx->stream));
	return CPVS_OK;
}

const uint32_t* cpvs_shadow_dag_device(const cpvs_shadow* s) { return s ? s->dag : nullptr; }

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
	if (tryLeafmasks && !s->info.leafmasks)
		return fail(CPVS_EINVAL, "cpvs_shadow_lookup_ndc: tryLeafmasks on a DAG built without leafmasks (SURVEY.md T2)");
	CPVS_CUDA(cudaSetDevice(s->ctx->device));
	cudaStream_t st = s->ctx->stream;
	LookupDag d{s->dag, nullptr, s->info.num_levels, 0, tryLeafmasks ? 1 : 0, nullptr, 0};
	if ((tryLeafmasks != 0) == (s->info.leafmasks != 0)) {  // shortcut grid, built once
		cpvs_shadow* ms = const_cast<cpvs_shadow*>(s);
		std::lock_guard<std::mutex> guard(ms->skipLock);
		if (!ms->skip) {
			const u32 g = skipLevelsFor(d.dagLevels, d.leafmasks, 0);
			if (g) {
				CPVS_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ms->skip), (sizeof(u32) << (3 * g)), st));
				d.skipLevels = g;
				s->ctx->launches += launchBuildSkipGrid(d, ms->skip, st);
				ms->skipLevels = g;
			}
		}
		d.skip = ms->skip;
		d.skipLevels = ms->skipLevels;
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

static void releaseContainerBuffers(cpvs_container* c) {
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
	if (s->ctx->device != c->ctx->device) return fail(CPVS_EINVAL, "cpvs_container_set: shadow lives on another device; use cpvs_container_set_dag");
	return cpvs_container_set_dag(c, s->dag, s->info.words, CPVS_MEM_DEVICE, s->info.num_levels, (int)s->info.leafmasks, x, y, z);
}

int cpvs_container_finalize(cpvs_container* c) {
	if (!c) return fail(CPVS_EINVAL, "cpvs_container_finalize: NULL argument");
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
	c->skipLevels = skipLevelsFor(c->dagLevels, c->leafmasks, c->gridLevels);
	if (c->skipLevels) {
		CPVS_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&c->skip), sizeof(u32) << (3 * (c->gridLevels + c->skipLevels)), st));
		LookupDag d{c->dag, c->grid, c->dagLevels, c->gridLevels, c->leafmasks, nullptr, c->skipLevels};
		c->ctx->launches += launchBuildSkipGrid(d, c->skip, st);
		CPVS_CUDA(cudaGetLastError());
	}
	c->finalized = true;
	return CPVS_OK;
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
	LookupDag d{c->dag, c->grid, c->dagLevels, c->gridLevels, c->leafmasks, c->skip, c->skipLevels};
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
	LookupDag d{c->dag, c->grid, c->dagLevels, c->gridLevels, c->leafmasks, c->skip, c->skipLevels};
	cudaStream_t st = c->ctx->stream;
	return runLookup(c->ctx, positions, (u64)count * 4, mem, visibilities, (u64)count,
			[&](const float* in, unsigned char* o) { return launchEvaluate(d, in, width, height, m, o, st); });
}

namespace {
struct ContainerFileHeader {
	char magic[8];
	u32 version, length, dagLevels, gridLevels, leafmasks, reserved;
	u64 dagWords, gridCells, fnv64;
	u64 pad;
};
static_assert(sizeof(ContainerFileHeader) == 64, "on-disk header is 64 bytes");
u64 fnv64Words(const u32* w, u64 n) {
	u64 h = 14695981039346656037ull;
	for (u64 i = 0; i < n; ++i) {
		h ^= w[i];
		h *= 1099511628211ull;
	}
	return h;
}
}  // namespace

int cpvs_container_save(const cpvs_container* c, const char* path) {
	if (!c || !c->finalized || !path) return fail(CPVS_EINVAL, "cpvs_container_save: container not finalized or NULL path");
	std::vector<u32> dag(c->dagWords), grid(c->cells.size());
	int rc = cpvs_container_copy(c, dag.data(), grid.data());
	if (rc != CPVS_OK) return rc;
	ContainerFileHeader h;
	std::memset(&h, 0, sizeof(h));
	std::memcpy(h.magic, "CPVSDAG1", 8);
	h.version = 1;
	h.length = c->length;
	h.dagLevels = c->dagLevels;
	h.gridLevels = c->gridLevels;
	h.leafmasks = (u32)c->leafmasks;
	h.dagWords = c->dagWords;
	h.gridCells = grid.size();
	h.fnv64 = fnv64Words(dag.data(), dag.size());
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
	bool ok = std::fread(&h, sizeof(h), 1, f) == 1 && std::memcmp(h.magic, "CPVSDAG1", 8) == 0 && h.version == 1 && isPow2(h.length) &&
			  h.length <= 64 && h.gridCells == (u64)h.length * h.length * h.length && h.dagWords > 0 && h.dagWords <= (1ull << 32) &&
			  h.dagLevels > 3 && h.dagLevels < kMaxLevels;
	if (ok) {
		grid.resize(h.gridCells);
		dag.resize(h.dagWords);
		ok = std::fread(grid.data(), sizeof(u32), grid.size(), f) == grid.size() && std::fread(dag.data(), sizeof(u32), dag.size(), f) == dag.size() &&
			 fnv64Words(dag.data(), dag.size()) == h.fnv64;
	}
	std::fclose(f);
	if (!ok) return fail(CPVS_EINVAL, "cpvs_container_load: %s is not a valid container file (header, size or checksum)", path);
	cpvs_container* c = nullptr;
	int rc = cpvs_container_create(ctx, h.length, &c);
	if (rc != CPVS_OK) return rc;
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
	c->skipLevels = skipLevelsFor(c->dagLevels, c->leafmasks, c->gridLevels);
	if (e == cudaSuccess && c->skipLevels) {
		e = cudaMallocAsync(reinterpret_cast<void**>(&c->skip), sizeof(u32) << (3 * (c->gridLevels + c->skipLevels)), st);
		if (e == cudaSuccess) {
			LookupDag d{c->dag, c->grid, c->dagLevels, c->gridLevels, c->leafmasks, nullptr, c->skipLevels};
			ctx->launches += launchBuildSkipGrid(d, c->skip, st);
			e = cudaGetLastError();
		}
	}
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // the host staging buffers go away
	if (e != cudaSuccess) {
		cpvs_container_destroy(c);
		return fail(CPVS_ECUDA, "cpvs_container_load: %s", cudaGetErrorString(e));
	}
	c->finalized = true;
	*out = c;
	return CPVS_OK;
}

int cpvs_container_set_filter_size(cpvs_container* c, uint32_t size) {
	if (!c) return fail(CPVS_EINVAL, "cpvs_container_set_filter_size: NULL argument");
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
