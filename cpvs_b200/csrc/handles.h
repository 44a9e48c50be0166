// Handle structs behind the opaque types of include/cpvs_b200.h and the helpers every translation unit of the C ABI
// shares (status codes, last-error text, small utilities). capi.cu owns contexts, hierarchies, lookups and containers;
// build.cu owns cpvs_shadow_create; grid.cu the multi-device tile-grid driver.
#pragma once
#include "../../include/cpvs_b200.h"

#include <cstdarg>
#include <cstdio>
#include <initializer_list>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "kernels.h"

namespace cpvs {

int fail(int code, const char* fmt, ...);  // records the message for cpvs_last_error() and returns `code`

#define CPVS_CUDA(expr)                                                                                            \
	do {                                                                                                           \
		cudaError_t _e = (expr);                                                                                   \
		if (_e != cudaSuccess)                                                                                     \
			return ::cpvs::fail(_e == cudaErrorMemoryAllocation ? CPVS_ENOMEM : CPVS_ECUDA, "%s: %s (%s:%d)", #expr, \
					cudaGetErrorString(_e), __FILE__, __LINE__);                                                   \
	} while (0)

inline bool isPow2(u64 v) { return v && !(v & (v - 1)); }
inline u64 pow2AtLeast(u64 v) {
	u64 p = 1;
	while (p < v) p <<= 1;
	return p;
}

// Sizes of the last build with the same shape (side, z tile, leafmasks): the next build of that shape carves its scratch
// arena, sizes its grids and allocates its DAG from these numbers plus head room, without asking the device first -- a
// light that moves a little changes the octree a little. Every kernel stays inside the capacities it was given and
// reports when one did not suffice; the build then runs again with exact counts.
struct SizeMemo {
	int n = 0;
	u32 zTileIndex = 0, zTileNum = 0;
	int leafmasks = 0;
	u64 nodes[kMaxLevels] = {0};   // SVO nodes per level
	u64 unique[kMaxLevels] = {0};  // nodes per level after merging
	u64 words = 0;                 // DAG words
};

}  // namespace cpvs

constexpr int kNumScalars = 256;  // device words per build (counters, sizes, flags); see build.cu

struct cpvs_ctx {
	int device;
	cudaStream_t own;
	cudaStream_t stream;
	cpvs::u64 launches;
	// Scratch arena for cpvs_shadow_create: one device allocation, grown when a build needs more and
	// kept between calls, carved by bump pointer -- no allocator traffic in steady state. Builds on one
	// context are serialised by `buildLock` (use one context per host thread for concurrent builds).
	std::mutex buildLock;
	char* arena;
	size_t arenaBytes;
	cpvs::u64* scalars;      // kNumScalars device words: counters, sizes and flags of the build in flight
	// Pinned read-back buffers (kNumScalars words each), one per build in flight: taken when a build is enqueued, returned
	// when it is finished; the pool grows when more builds are in flight than it has buffers.
	std::vector<cpvs::u64*> readbackFree, readbackAll;
	std::vector<cpvs::u64*> countBuffers;  // pinned, kPooledCountSlices * kMaxLevels words each (column counts in flight); under cacheLock
	// High-priority side stream for the chain of inserts (the critical path of the merge): its CTAs are dispatched
	// ahead of the queued CTAs of the rank scans and the leaf emission running beside it. Fork/join through events.
	cudaStream_t aux;
	cudaEvent_t evFork, evJoin;
	// Normal-priority side streams for the per-level rank scans, which only the final emission needs and
	// which therefore run next to the following levels' inserts (the ranks of different levels are independent:
	// they alternate between the two), and one for the leaf emission, which starts as soon as the leaf level is ranked.
	cudaStream_t aux2, aux3, aux4;
	cudaEvent_t evJoin3, evClear, evCols, evLeafRanked, evLeafEmitted;
	int leafColumns;  // leaves built per column: 1 = where it pays (default), 0 = never, 2 = always (CPVS_LEAF_COLUMNS; tests)
	int predictSizes; // 1 (default): size builds from the memo of the last build of the same shape; 0: always count first (CPVS_PREDICT=0; tests)
	unsigned headroomShift;  // capacities = predicted + (predicted >> headroomShift) + slack
	std::vector<cpvs::SizeMemo> memos;
	std::mutex memoLock;  // memos are also read by the other contexts of the family
	cpvs::u64 predictedBuilds, exactBuilds, overflowRebuilds, reemissions;  // statistics (cpvs_ctx_stats)
	// Sides of the depth maps whose last build did NOT use the per-column leaf builder (cities, planes): hierarchies of
	// that side skip the column residues, which only that builder reads.
	std::vector<int> noColumnSides;
	// Large per-build buffers (pyramid levels, residues, device copies of host depth maps, DAG allocations) are recycled inside
	// the context: a build of a shape seen before then allocates nothing, whatever the driver's pool would have to do to serve
	// several contexts and peers (with a communication library's peer mappings a fresh block costs milliseconds).
	std::mutex cacheLock;
	std::vector<std::pair<void*, size_t>> freeBlocks;  // oldest first
	std::unordered_map<void*, size_t> liveBlocks;
	size_t cachedBytes;
	// Staging buffers for DAGs whose size is only bounded when they are emitted (build.cu): (pointer, words), under cacheLock.
	std::vector<std::pair<cpvs::u32*, cpvs::u64>> stagingFree;
	cpvs::u64 stagingWords;  // the size new staging buffers get (kept in the family's first context, under sizeLock)
	cpvs::u64 stagingMaxWords;  // larger bounds are not staged
	size_t familyArenaBytes; // the largest arena of the family so far (likewise)
	std::mutex sizeLock;
	// Allocations of finished staged DAGs (>= 1 MB) that were released, kept for the next DAG of about that size (the same
	// grid built again asks for exactly these sizes): (pointer, bytes), oldest first, under cacheLock. They belong to the copy stream.
	std::vector<std::pair<cpvs::u32*, size_t>> dagFree;
	size_t dagFreeBytes;
	cudaStream_t copyStream;  // copies finished DAGs out of them, behind nothing else; their final allocations are made and released on it
	cudaEvent_t evCopyFree;
	// A second context on the same GPU, created on demand and kept (cpvs::siblingContext): independent builds -- the z-slices
	// of a tile -- alternate between the two so that their kernels run side by side.
	cpvs_ctx* sibling;
	cpvs_ctx* family;  // the first context of the chain of siblings this one belongs to (itself, if it was created by the caller)
	cpvs::u64 buildSerial;  // builds enqueued so far: a pending build whose serial is the latest still owns the arena's contents
};

struct cpvs_minmax {
	cpvs_ctx* ctx;
	int n;
	int numLevels;
	float* ownedDepth;    // device copy when built from host memory
	float* levelStorage;  // levels 1.. in one allocation
	// The depth map re-encoded per 8x8 column for the per-column leaf builder (pyramid.cu), valid for builds with
	// `residueTiles` z-slices; NULL when it was not produced.
	unsigned char* residue;
	cpvs::u32 residueTiles;
	const float* level[cpvs::kMaxLevels];
	cudaEvent_t evStart, evBase, evStop;
	// Levels 1 and 2 are not needed by the leafmask builder and are only produced on first use.
	std::mutex lowLock;
	bool lowLevelsBuilt;
	// Node counts of all z-slices of the column for the zTileNum last asked for (createShadowTiles builds them all
	// from this one pyramid): one launch and one read-back per hierarchy instead of one per slice, and slices that
	// miss the surface are answered from here without touching the device.
	cpvs::u32 columnSlices;
	int columnMinLevel;
	std::vector<cpvs::u64> columnCounts;  // [z * kMaxLevels + level]; [z * kMaxLevels + kRootMaskScalar] = 1 << 32 | root mask
	// Counts in flight (columnCountsBegin): the pinned buffer they are read back into, taken from `countsCtx`'s pool when
	// `countsPooled`, and the event that says they have arrived.
	cpvs::u64* countsPinned = nullptr;
	bool countsPooled = false;
	cpvs_ctx* countsCtx = nullptr;
	cudaEvent_t evCounts = nullptr;
	cpvs::u32 pendingSlices = 0;
	int pendingMinLevel = -1;
};

namespace cpvs {
// Private lookup-only copy of a DAG / container (lookup_index.cu): eight slots per inner node + one 32-byte k-code per leaf.
struct LookupIndex {
	u32* nodes = nullptr;  // 8 words per inner node: 0 shadow, 1 lit, 2 + id of the child node / leaf
	u32* grid = nullptr;   // container: cell table with root node ids
	u32* codes = nullptr;
	u32* skip = nullptr;   // shortcut over the top levels (see LookupDag::skip), holding node ids
	u32 skipLevels = 0;
	u64 numNodes = 0, numLeaves = 0;
	bool tried = false, valid = false;
};
void freeLookupIndex(cpvs_ctx* ctx, LookupIndex* ix);
int buildLookupIndex(cpvs_ctx* ctx, const u32* dag, u64 dagWords, const std::vector<u64>& cellStart, const std::vector<u64>& cellWords,
		const std::vector<u32>& hostGrid, u32 dagLevels, u32 gridLevels, LookupIndex* out);
}  // namespace cpvs

struct cpvs_pending_build;  // build.cu

struct cpvs_shadow {
	cpvs_ctx* ctx;
	// cpvs_shadow_create_async: the build is in flight; every accessor waits for it first (cpvs_shadow_wait)
	cpvs_pending_build* pending;
	int pendingLeafmasks;
	int status;              // result of the finished build
	std::string statusText;
	cpvs::u32* dag;       // first word of the DAG
	cpvs::u32* dagAlloc;  // the allocation it lives in (a predicted capacity; the DAG sits at its end)
	bool copyInFlight;    // the DAG is still being copied out of its staging buffer (between shadowWaitBegin and cpvs_shadow_wait)
	size_t dagAllocBytes; // size of dagAlloc when dagOnCopyStream
	bool dagOnCopyStream; // dagAlloc was allocated on the context's copy stream (a staged build) and is released there
	cudaEvent_t ready;    // recorded on the building stream once the words are written (consumers on other streams wait for it)
	cpvs_shadow_info info;
	// lookup shortcut over the top levels, built on the first lookup (see LookupDag::skip)
	std::mutex skipLock;
	cpvs::u32* skip;
	cpvs::u32 skipLevels;
	cpvs::LookupIndex index;  // built on the first lookup with leafmasks
};

struct ContainerCell {
	cpvs::u32* words = nullptr;  // device copy owned by the container
	cpvs::u64 count = 0;
	cpvs::u32 numLevels = 0;
	int leafmasks = 0;
	cpvs::u32 rootMask = 0;
	bool set = false;
};

struct cpvs_container {
	cpvs_ctx* ctx;
	cpvs::u32 length;
	cpvs::u32 filterSize;
	std::vector<ContainerCell> cells;
	cpvs::u32* dag = nullptr;
	cpvs::u32* grid = nullptr;
	cpvs::u32* skip = nullptr;  // lookup shortcut over the top levels of every cell (see LookupDag::skip)
	cpvs::u32 skipLevels = 0;
	cpvs::u64 dagWords = 0;
	cpvs::u32 dagLevels = 0, gridLevels = 0;
	int leafmasks = 0;
	bool finalized = false;
	bool loaded = false;  // came from cpvs_container_load: cells cannot be re-set
	cpvs::LookupIndex index;  // built when the container is finalized
	std::mutex indexLock;
};

namespace cpvs {
// capi.cu: stream-ordered blocks on the context's stream, recycled by size (see cpvs_ctx::freeBlocks).
cudaError_t ctxAlloc(cpvs_ctx* ctx, void** out, size_t bytes);
void ctxFree(cpvs_ctx* ctx, void* p);
void ctxAdopt(cpvs_ctx* ctx, void* p, size_t bytes);  // a pool allocation made elsewhere becomes one ctxFree may recycle
cpvs_ctx* siblingContext(cpvs_ctx* ctx);  // NULL if it cannot be created
// capi.cu: levels 1 and 2 of a hierarchy on demand.
int ensureLowLevels(const cpvs_minmax* mm, int level);
// Node counts of all z-slices of the hierarchy's column (build.cu): Begin enqueues the launch and its read-back on the context's
// stream, Of waits for them (and begins them if nobody has). counts: [z * kMaxLevels + level].
int columnCountsBegin(cpvs_ctx* ctx, const cpvs_minmax* mm, u32 zTileNum, int minLevel);
int columnCountsOf(cpvs_ctx* ctx, const cpvs_minmax* mm, u32 zTileNum, int minLevel, const u64** counts);
void releaseCountsBuffer(cpvs_minmax* mm);
// cpvs_shadow_wait in two halves, for callers that finish several builds at once: Begin does everything but wait for the copy of
// a staged DAG into its final allocation; cpvs_shadow_wait afterwards waits for it (one wait serves all the copies of a context).
int shadowWaitBegin(cpvs_shadow* s);
// build.cu: a new shadow handle around `words` device words (NULL: allocate one word and store `rootMask` there).
PyramidView pyramidView(const cpvs_minmax* mm);
}  // namespace cpvs
