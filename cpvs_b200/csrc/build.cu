// CompressedShadow::create (reference src/CompressedShadow.cpp:49-59: constructSvo -> mergeCommonSubtrees -> compress)
// as a chain of device stages:
//
//   plan      node capacities per level and the DAG capacity -- from the size memo of the last build of the same shape
//             (no device round trip), or from the closed-form count of the hierarchy's column (one read-back)
//   carve     per-level arrays out of the context's scratch arena
//   expand    breadth-first construction, top level first (src/CompressedShadow.cpp:87-169)
//   leaves    constructLastLevels (src/CompressedShadow.cpp:171-190)
//   merge     bottom-up inserts on a high-priority stream, rank scans beside them (src/CompressedShadow.cpp:215-304);
//             the leaf level is emitted as soon as it is ranked when the DAG's allocation already exists
//   bases     level bases and total size
//   emit      every unique node once, in its final place (src/CompressedShadow.cpp:326-392)
//   finish    ONE read-back of sizes and flags; a capacity that did not suffice sends the build through again with
//             exact numbers, so results never depend on the prediction
//
// Every kernel reads the sizes it needs on the device; the host only supplies capacities.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <new>

#include "handles.h"

using namespace cpvs;

namespace {

__global__ void storeU32Kernel(u32* dst, u32 value) { *dst = value; }
__global__ void storeCoordKernel(u64* dst, u64 value) { *dst = value; }

// CPVS_TRACE=1: host wall-clock between orchestration steps, to stderr.
struct HostTrace {
	bool on;
	std::chrono::steady_clock::time_point last;
	HostTrace() : on(std::getenv("CPVS_TRACE") && std::getenv("CPVS_TRACE")[0] == '1'), last(std::chrono::steady_clock::now()) {}
	void mark(const char* what) {
		if (!on) return;
		const auto now = std::chrono::steady_clock::now();
		fprintf(stderr, "[cpvs trace] %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(now - last).count());
		last = now;
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
	u64 cap = 0;   // capacity of the arrays (>= the nodes the level will have)
	u64 hint = 0;  // expected unique nodes (grid size of the emission)
	u64* coords = nullptr;
	u16* masks = nullptr;
	u32* firstChild = nullptr;
	u32* uid = nullptr;
	u32* firstList = nullptr;
	u32* wordOffset = nullptr;
	u32* leafCodes = nullptr;
	u64* leafHash = nullptr;
	u32* leafAt = nullptr;  // leaf level, built per column: node index by column-order position
	u64* table = nullptr;   // merge table of this level (large levels own one; small levels share)
	u64 tableSlots = 0;
	u32* slotOffset = nullptr;        // per table slot: word offset of the group's node
	unsigned char* sizeOf = nullptr;  // rank scratch, one byte per node (large levels)
};

// Device scalars of a build (cpvs_ctx::scalars), indices in 64-bit words.
enum : int {
	kSlotNodes = 0,      // [32] SVO nodes per level, written by the expansion of the level above
	kSlotUnique = 32,    // [32] nodes per level after merging
	kSlotWords = 64,     // [32] compressed words per level
	kSlotBases = 96,     // [32] word offset of each level in the DAG
	kSlotTotal = 160,    // words of the DAG
	kSlotSketchBits = 161,
	kSlotLeafTableMask = 162,
	kSlotTableError = 163,  // u32: a probe sequence wrapped a merge table
	kSlotOverflow = 164,    // u32: kOverflowNodes | kOverflowWords
	kSlotRootMask = 165,    // u32: first word of the DAG (the root's mask)
	kSlotTallColumns = 166, // u32: the residue leaf builder met columns it leaves to the depth-based one
};

struct PhaseEvents {
	cudaEvent_t ev[CPVS_NUM_PHASES + 1];  // ev[i] opens phase i, ev[CPVS_NUM_PHASES] closes the last one
	PhaseEvents() { std::memset(ev, 0, sizeof(ev)); }
	~PhaseEvents() {
		for (cudaEvent_t e : ev)
			if (e) cudaEventDestroy(e);
	}
};

struct Build {
	cpvs_ctx* ctx = nullptr;
	const cpvs_minmax* mm = nullptr;
	u32 zTileIndex, zTileNum;
	int L, top, minLevel, lastInner;
	bool useLeaf;
	PyramidView pyr;
	cudaStream_t st = nullptr;

	bool predicted = false;  // capacities come from the size memo, nothing was read back before the build
	bool exact = false;      // node capacities are the closed-form counts: the expansion must reproduce them
	LevelArrays lv[kMaxLevels];
	int smallLow = 0;  // levels smallLow..top are walked by the single-CTA kernels
	bool leafColumns = false, haveLeaves = false;
	u64 numCols = 0;
	u64 expectedDistinctLeaves = 0;  // > 0: the leaf table is sized from the memo, no distinct-count sketch is taken

	u64* dScalars = nullptr;
	ScanTileState* dTiles = nullptr;
	u32* dTickets = nullptr;
	u64* dTable = nullptr;  // shared by the small levels
	u32* dSketch = nullptr;
	u32* dColBias = nullptr;
	u64 scanTiles = 0, scanLaunches = 0, tileCursor = 0, launchCursor = 0;

	u32* dagAlloc = nullptr;  // capacity words; the DAG ends at its end (released here unless a shadow took it over)
	u64 dagCapacity = 0;
	// staged: dagAlloc is one of the context's staging buffers (stagingWords long), big enough for any DAG these node counts can
	// give; the finished DAG is copied out of it into an allocation of its size.
	bool staged = false;
	u64 stagingWords = 0;
	bool leavesEmitted = false;

	PhaseEvents phases;
	cudaEvent_t countStart = nullptr;                                                // opens phase COUNT when the column was counted for this build
	cudaEvent_t evRankStart = nullptr, evRankStop = nullptr, evLeafEmitStart = nullptr;  // timing of the concurrent leaf-level kernels
	cudaEvent_t done = nullptr;  // recorded behind the read-back of sizes and flags
	u64* h = nullptr;            // pinned read-back slot
	HostTrace trace;

	Build() = default;
	Build(const Build&) = delete;
	~Build() {  // (under the context's build lock, like everything that touches a Build)
		if (h) {
			if (done) cudaEventSynchronize(done);  // an abandoned build may still be writing it
			ctx->readbackFree.push_back(h);
		}
		if (dagAlloc && staged) {
			std::lock_guard<std::mutex> guard(ctx->cacheLock);
			ctx->stagingFree.emplace_back(dagAlloc, stagingWords);
		} else if (dagAlloc) {
			ctxFree(ctx, dagAlloc);
		}
		for (cudaEvent_t e : {evRankStart, evRankStop, evLeafEmitStart, done})
			if (e) cudaEventDestroy(e);
	}

	u64* dNodes() const { return dScalars + kSlotNodes; }
	u64* dUnique() const { return dScalars + kSlotUnique; }
	u64* dWords() const { return dScalars + kSlotWords; }
	u64* dBases() const { return dScalars + kSlotBases; }
	u64* dTotal() const { return dScalars + kSlotTotal; }
	u32* dTableError() const { return reinterpret_cast<u32*>(dScalars + kSlotTableError); }
	u32* dOverflow() const { return reinterpret_cast<u32*>(dScalars + kSlotOverflow); }
	bool leafLevel(int l) const { return useLeaf && l == 2; }
	ScanLaunch nextScan(u64 n, u64 tileNodes = kScanTile) {
		ScanLaunch s{dTickets + launchCursor, dTiles + tileCursor};
		++launchCursor;
		tileCursor += (n + tileNodes - 1) / tileNodes;
		return s;
	}
};

SizeMemo* ownMemo(cpvs_ctx* ctx, int n, u32 zTileIndex, u32 zTileNum, int leafmasks) {
	for (SizeMemo& m : ctx->memos)
		if (m.n == n && m.zTileIndex == zTileIndex && m.zTileNum == zTileNum && m.leafmasks == leafmasks) return &m;
	return nullptr;
}

// The sizes of the last build of this shape: the context's own, else what another context of its family (the contexts
// cpvs::siblingContext chains on one GPU, which take turns on the same stream of builds) has seen.
bool findMemo(cpvs_ctx* ctx, int n, u32 zTileIndex, u32 zTileNum, int leafmasks, SizeMemo* out) {
	auto lookIn = [&](cpvs_ctx* c) {
		std::lock_guard<std::mutex> guard(c->memoLock);
		const SizeMemo* m = ownMemo(c, n, zTileIndex, zTileNum, leafmasks);
		if (m) *out = *m;
		return m != nullptr;
	};
	if (lookIn(ctx)) return true;
	for (cpvs_ctx* c = ctx->family; c; c = c->sibling)
		if (c != ctx && lookIn(c)) return true;
	return false;
}

void rememberSizes(cpvs_ctx* ctx, const Build& b, const u64* h) {
	std::lock_guard<std::mutex> guard(ctx->memoLock);
	SizeMemo* m = ownMemo(ctx, b.mm->n, b.zTileIndex, b.zTileNum, b.useLeaf ? 1 : 0);
	if (!m) {
		if (ctx->memos.size() >= 256) ctx->memos.erase(ctx->memos.begin());
		ctx->memos.emplace_back();
		m = &ctx->memos.back();
		m->n = b.mm->n;
		m->zTileIndex = b.zTileIndex;
		m->zTileNum = b.zTileNum;
		m->leafmasks = b.useLeaf ? 1 : 0;
	}
	for (int l = 0; l < kMaxLevels; ++l) {
		m->nodes[l] = h[kSlotNodes + l];
		m->unique[l] = h[kSlotUnique + l];
	}
	m->words = h[kSlotTotal];
}

cpvs_shadow* newShadow(cpvs_ctx* ctx) {
	cpvs_shadow* s = new (std::nothrow) cpvs_shadow();
	if (!s) return nullptr;
	std::memset(&s->info, 0, sizeof(s->info));
	s->ctx = ctx;
	s->dag = s->dagAlloc = nullptr;
	s->ready = nullptr;
	s->skip = nullptr;
	s->skipLevels = 0;
	s->pending = nullptr;
	s->dagOnCopyStream = false;
	s->dagAllocBytes = 0;
	s->copyInFlight = false;
	s->pendingLeafmasks = 0;
	s->status = CPVS_OK;
	return s;
}

// A z-slice that misses the surface altogether (most slices of a tall tile grid): the root has no PARTIAL child, the DAG
// is its one mask word (0x5555 lit / 0x0000 shadow). The word is stored by a kernel on the context's stream; consumers on
// other streams wait for the handle's `ready` event.
int oneWordShadow(cpvs_ctx* ctx, int L, bool useLeaf, u32 rootMask, cpvs_shadow* s) {
	cudaStream_t st = ctx->stream;
	cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&s->dagAlloc), sizeof(u32), st);
	if (e == cudaSuccess) {
		s->dag = s->dagAlloc;
		storeU32Kernel<<<1, 1, 0, st>>>(s->dag, rootMask);
		++ctx->launches;
		e = cudaEventCreateWithFlags(&s->ready, cudaEventDisableTiming);
	}
	if (e == cudaSuccess) e = cudaEventRecord(s->ready, st);
	if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? CPVS_ENOMEM : CPVS_ECUDA, "cpvs_shadow_create: %s", cudaGetErrorString(e));
	const int top = L - 2;
	std::memset(&s->info, 0, sizeof(s->info));
	s->info.num_levels = (u32)L;
	s->info.leafmasks = useLeaf ? 1 : 0;
	s->info.total_visibility = rootMask == 0x5555u ? CPVS_VISIBLE : (rootMask == 0u ? CPVS_SHADOW : CPVS_PARTIAL);
	s->info.words = 1;
	s->info.svo_nodes[top] = s->info.dag_nodes[top] = s->info.dag_words[top] = 1;
	return CPVS_OK;
}

}  // namespace
namespace cpvs {
void releaseCountsBuffer(cpvs_minmax* mm) {
	if (!mm->countsPinned) return;
	if (mm->countsPooled) {
		std::lock_guard<std::mutex> poolGuard(mm->countsCtx->cacheLock);
		mm->countsCtx->countBuffers.push_back(mm->countsPinned);
	} else {
		cudaFreeHost(mm->countsPinned);
	}
	mm->countsPinned = nullptr;
}
}  // namespace cpvs
namespace {
constexpr u32 kPooledCountSlices = 64;

// Closed-form node counts of all z-slices of the hierarchy's column, computed once per (hierarchy, zTileNum, minLevel):
// one launch and one read-back, shared by every slice built from this pyramid. columnCountsBegin enqueues the two (the
// read-back into a pinned buffer of the context) and returns; columnCounts waits for them.
int columnCountsBegin(cpvs_ctx* ctx, const cpvs_minmax* cmm, const PyramidView& pyr, u32 zTileNum, int minLevel) {
	cpvs_minmax* mm = const_cast<cpvs_minmax*>(cmm);
	std::lock_guard<std::mutex> guard(mm->lowLock);
	if (mm->columnSlices == zTileNum && mm->columnMinLevel == minLevel) return CPVS_OK;
	if (mm->countsPinned && mm->pendingSlices == zTileNum && mm->pendingMinLevel == minLevel) return CPVS_OK;
	if (mm->countsPinned) CPVS_CUDA(cudaEventSynchronize(mm->evCounts));  // counts for another slicing in flight: superseded
	const size_t words = (size_t)zTileNum * kMaxLevels;
	if (mm->countsPinned && (!mm->countsPooled || zTileNum > kPooledCountSlices)) releaseCountsBuffer(mm);
	if (!mm->countsPinned) {
		if (zTileNum <= kPooledCountSlices) {
			std::lock_guard<std::mutex> poolGuard(ctx->cacheLock);
			if (ctx->countBuffers.empty()) {
				u64* buf = nullptr;
				CPVS_CUDA(cudaMallocHost(reinterpret_cast<void**>(&buf), (size_t)kPooledCountSlices * kMaxLevels * sizeof(u64)));
				ctx->countBuffers.push_back(buf);
			}
			mm->countsPinned = ctx->countBuffers.back();
			ctx->countBuffers.pop_back();
			mm->countsPooled = true;
		} else {
			CPVS_CUDA(cudaMallocHost(reinterpret_cast<void**>(&mm->countsPinned), words * sizeof(u64)));
			mm->countsPooled = false;
		}
		mm->countsCtx = ctx;
	}
	if (!mm->evCounts) CPVS_CUDA(cudaEventCreateWithFlags(&mm->evCounts, cudaEventDisableTiming));
	u64* dCounts = nullptr;
	cudaStream_t st = ctx->stream;
	CPVS_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&dCounts), words * sizeof(u64), st));
	cudaError_t e = cudaMemsetAsync(dCounts, 0, words * sizeof(u64), st);
	if (e == cudaSuccess) {
		ctx->launches += launchColumnCounts(pyr, zTileNum, minLevel, dCounts, st);
		e = cudaMemcpyAsync(mm->countsPinned, dCounts, words * sizeof(u64), cudaMemcpyDeviceToHost, st);
	}
	if (e == cudaSuccess) e = cudaEventRecord(mm->evCounts, st);
	cudaFreeAsync(dCounts, st);
	if (e != cudaSuccess) return fail(CPVS_ECUDA, "column counts: %s", cudaGetErrorString(e));
	mm->pendingSlices = zTileNum;
	mm->pendingMinLevel = minLevel;
	return CPVS_OK;
}

int columnCounts(cpvs_ctx* ctx, const cpvs_minmax* cmm, const PyramidView& pyr, u32 zTileNum, int minLevel, const u64** counts) {
	cpvs_minmax* mm = const_cast<cpvs_minmax*>(cmm);
	if (mm->columnSlices != zTileNum || mm->columnMinLevel != minLevel) {
		if (int rc = columnCountsBegin(ctx, cmm, pyr, zTileNum, minLevel)) return rc;
		std::lock_guard<std::mutex> guard(mm->lowLock);
		if (mm->columnSlices != zTileNum || mm->columnMinLevel != minLevel) {
			cudaError_t e = cudaEventSynchronize(mm->evCounts);
			if (e != cudaSuccess) return fail(CPVS_ECUDA, "column counts: %s", cudaGetErrorString(e));
			mm->columnCounts.assign(mm->countsPinned, mm->countsPinned + (size_t)zTileNum * kMaxLevels);
			releaseCountsBuffer(mm);
			mm->columnSlices = zTileNum;
			mm->columnMinLevel = minLevel;
		}
	}
	*counts = mm->columnCounts.data();
	return CPVS_OK;
}

// No DAG of an octree with these node counts has more words: every node its mask, every node below the root one pointer in
// its parent, every leaf two words per slice (emit.cu).
u64 upperBoundWords(const u64* counts, int top, int minLevel, bool useLeaf) {
	u64 words = 1;
	for (int l = top - 1; l >= minLevel && counts[l]; --l) words += 2 * counts[l] + (useLeaf && l == 2 ? 16 * counts[l] : 0);
	return words;
}
constexpr size_t kDagKeepMinBytes = 1u << 20, kDagKeepMaxBytes = 16ull << 30, kDagKeepMaxBlocks = 256;  // cpvs_ctx::dagFree
// Bounds above cpvs_ctx::stagingMaxWords (2 GiB of words; CPVS_STAGING_MAX_WORDS for tests) are not staged: the DAG's size is
// then predicted from the memo, or waited for.

// A staging buffer of at least `words` words. All buffers of a context have one size, the largest bound seen so far plus a
// quarter: once a grid's heaviest slice has been seen, nothing is allocated any more (buffers of an earlier, smaller size are
// released as they turn up).
cudaError_t takeStaging(cpvs_ctx* ctx, u64 words, u32** out, u64* have) {
	std::vector<u32*> drop;
	u64 size = 0;
	*out = nullptr;
	{
		{  // one size for the whole family: the lanes of a grid worker meet the same slices sooner or later
			std::lock_guard<std::mutex> familyGuard(ctx->family->sizeLock);
			if (words > ctx->family->stagingWords) ctx->family->stagingWords = words + (words >> 2);
			size = ctx->family->stagingWords;
		}
		std::lock_guard<std::mutex> guard(ctx->cacheLock);
		while (!ctx->stagingFree.empty() && !*out) {
			if (ctx->stagingFree.back().second >= size) {
				*out = ctx->stagingFree.back().first;
				*have = ctx->stagingFree.back().second;
			} else {
				drop.push_back(ctx->stagingFree.back().first);
			}
			ctx->stagingFree.pop_back();
		}
	}
	for (u32* p : drop) cudaFreeAsync(p, ctx->stream);
	if (*out) return cudaSuccess;
	*have = size;
	return cudaMallocAsync(reinterpret_cast<void**>(out), size * sizeof(u32), ctx->stream);
}

// ---- plan ----------------------------------------------------------------------------------------------------------------
// Fills lv[].cap / hint, smallLow, leafColumns. exactCounts: node counts per level (index = level) or NULL to predict from `memo`.
void planLevels(Build& b, const u64* exactCounts, const SizeMemo* memo) {
	cpvs_ctx* ctx = b.ctx;
	u64 expect[kMaxLevels] = {0};
	b.lv[b.top].cap = b.lv[b.top].hint = expect[b.top] = 1;
	for (int l = b.top - 1; l >= b.minLevel; --l) {
		if (exactCounts) {
			expect[l] = expect[l + 1] ? exactCounts[l] : 0;  // the octree stops below an empty level
			b.lv[l].cap = expect[l];
			b.lv[l].hint = memo ? memo->unique[l] + (memo->unique[l] >> 3) : expect[l];
			if (b.lv[l].hint > expect[l] || !b.lv[l].hint) b.lv[l].hint = expect[l];
		} else {
			expect[l] = memo->nodes[l];
			b.lv[l].cap = expect[l] ? expect[l] + (expect[l] >> ctx->headroomShift) + 256 : 0;
			b.lv[l].hint = memo->unique[l] + (memo->unique[l] >> ctx->headroomShift) + 256;
		}
	}
	// The top levels up to kSmallMaxNodes nodes each ("small": smallLow..top) are handled by single-CTA
	// kernels, one launch per phase instead of one or more per level.
	b.smallLow = b.top + 1;
	for (int l = b.top; l >= b.lastInner && b.lv[l].cap && b.lv[l].cap <= kSmallMaxNodes; --l) b.smallLow = l;
	b.haveLeaves = b.useLeaf && b.lv[2].cap > 0;
	// (the sizing kernel adds half again and rounds up to a power of two: no further head room here, or the table doubles
	// and drops out of L2; a build with many more distinct leaves than this is redone)
	if (b.haveLeaves && !exactCounts && memo->unique[2]) b.expectedDistinctLeaves = memo->unique[2];
	// Leaves per column -- where it pays: whole-volume builds with 2..8 leaves per column of a depth map that does not fit in L2 (terrain-like
	// surfaces; measured at 16K^2: 0.42 ms against 0.50 ms per leaf, at 8192^2 0.121 against 0.137, at 4096^2 -- 64 MiB, L2
	// serves the re-reads -- 0.046 against 0.045). A z-slice of a tall grid leaves most columns empty; box edges make columns
	// of hundreds of leaves that a four-lane group walks alone (16K^2 city: 6.1 ms against 2.9 ms); a gentle plane has one
	// leaf per column and nothing to share (0.30 ms against 0.27 ms): those keep the per-leaf kernel.
	const u64 allCols = ((u64)b.mm->n >> 3) * ((u64)b.mm->n >> 3);
	b.leafColumns = b.haveLeaves && (ctx->leafColumns == 2 || (ctx->leafColumns == 1 && b.zTileNum == 1 && b.mm->n >= 8192 &&
																 expect[2] >= 2 * allCols && expect[2] <= 8 * allCols));
	b.numCols = b.leafColumns ? allCols : 0;
	// remember for the hierarchies to come whether maps of this side take the per-column builder (cpvs_minmax_build)
	if (b.haveLeaves && b.zTileNum == 1) {
		std::vector<int>& no = ctx->noColumnSides;
		no.erase(std::remove(no.begin(), no.end(), b.mm->n), no.end());
		if (!b.leafColumns) no.push_back(b.mm->n);
	}
}

// ---- carve ---------------------------------------------------------------------------------------------------------------
int carveArena(Build& b) {
	cpvs_ctx* ctx = b.ctx;
	b.scanTiles = b.scanLaunches = 0;
	for (int l = b.top; l >= b.minLevel; --l) {
		const u64 n = b.lv[l].cap;
		if (!n || l >= b.smallLow) continue;
		if (!b.leafLevel(l)) {
			b.scanTiles += (n + kExpandTileNodes - 1) / kExpandTileNodes;
			++b.scanLaunches;
		}
		b.scanTiles += (n + kScanTile - 1) / kScanTile;
		++b.scanLaunches;
	}
	if (b.leafColumns) {  // one more scan, over the columns (= texels of pyramid level 3)
		b.scanTiles += (b.numCols + kScanTile - 1) / kScanTile;
		++b.scanLaunches;
	}
	const u64 maxTable = 2 * kSmallMaxNodes;  // shared by the small levels; large levels own their tables
	auto carve = [&](ArenaCarver& ar) {
		b.dTiles = ar.take<ScanTileState>(b.scanTiles);
		b.dTickets = ar.take<u32>(b.scanLaunches);
		b.dSketch = ar.take<u32>(b.haveLeaves && !b.expectedDistinctLeaves ? kSketchWords : 0);
		b.dTable = ar.take<u64>(maxTable);
		b.dColBias = ar.take<u32>(b.numCols);
		for (int l = b.top; l >= b.minLevel; --l) {
			LevelArrays& a = b.lv[l];
			if (!a.cap) continue;
			if (b.leafColumns && l == 2)
				a.leafAt = ar.take<u32>(a.cap);
			else
				a.coords = ar.take<u64>(a.cap);
			// (rounded up to whole rank tiles: the rank kernels fetch their vectors before they know the level's size)
			a.masks = ar.take<u16>(((a.cap + kScanTile - 1) / kScanTile) * kScanTile + 4);
			a.uid = ar.take<u32>(((a.cap + kScanTile - 1) / kScanTile) * kScanTile + 4);
			a.firstList = ar.take<u32>(a.cap);
			a.wordOffset = ar.take<u32>(a.cap);
			if (l >= b.smallLow) {
				a.table = b.dTable;
				a.tableSlots = 2 * kSmallMaxNodes;
			} else {
				a.tableSlots = pow2AtLeast(a.cap * 2 < 1024 ? 1024 : a.cap * 2);
				a.table = ar.take<u64>(a.tableSlots + kDirectSlots);
			}
			a.slotOffset = ar.take<u32>(a.tableSlots + kDirectSlots);
			if (l < b.smallLow) a.sizeOf = ar.take<unsigned char>(a.cap + 4);
			if (b.leafLevel(l)) {
				a.leafCodes = ar.take<u32>(a.cap * 8);
				if (!b.leafColumns) a.leafHash = ar.take<u64>(a.cap);
			} else {
				a.firstChild = ar.take<u32>(a.cap);
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
		if (ctx->arena) CPVS_CUDA(cudaFreeAsync(ctx->arena, b.st));
		ctx->arena = nullptr;
		ctx->arenaBytes = 0;
		size_t want = sizing.offset + sizing.offset / 4;
		if (want < doubled) want = doubled;
		{  // at least what another context of the family has already needed (the lanes of a grid worker take turns on the same slices)
			std::lock_guard<std::mutex> familyGuard(ctx->family->sizeLock);
			const size_t hint = ctx->family->familyArenaBytes < (8ull << 30) ? ctx->family->familyArenaBytes : (8ull << 30);  // (not a giant one-off build's)
			if (want < hint) want = hint;
		}
		size_t freeBytes = 0, totalBytes = 0;
		if (cudaMemGetInfo(&freeBytes, &totalBytes) == cudaSuccess && want > freeBytes / 2) want = sizing.offset + sizing.offset / 8;
		cudaError_t ae = cudaMallocAsync(reinterpret_cast<void**>(&ctx->arena), want, b.st);
		if (ae != cudaSuccess) return fail(CPVS_ENOMEM, "scratch arena of %zu bytes: %s", want, cudaGetErrorString(ae));
		ctx->arenaBytes = want;
		std::lock_guard<std::mutex> familyGuard(ctx->family->sizeLock);
		if (ctx->family->familyArenaBytes < want) ctx->family->familyArenaBytes = want;
	}
	ArenaCarver real(ctx->arena);
	carve(real);
	b.tileCursor = b.launchCursor = 0;
	// tile states, tickets and the leaf sketch sit at the front of the arena: one memset clears them all
	CPVS_CUDA(cudaMemsetAsync(ctx->arena, 0, reinterpret_cast<char*>(b.dTable) - ctx->arena, b.st));
	return CPVS_OK;
}

// ---- expand + leaves -------------------------------------------------------------------------------------------------------
int stageExpand(Build& b, bool& tablesClearing) {
	cpvs_ctx* ctx = b.ctx;
	cudaStream_t st = b.st;
	if (b.leafColumns) {  // the columns' scan runs beside the upper expansions
		CPVS_CUDA(cudaEventRecord(ctx->evFork, st));
		CPVS_CUDA(cudaStreamWaitEvent(ctx->aux2, ctx->evFork, 0));
		ScanLaunch colScan{b.dTickets + b.scanLaunches - 1, b.dTiles + b.scanTiles - (b.numCols + kScanTile - 1) / kScanTile};
		ctx->launches += launchColumnBias(b.pyr, b.zTileIndex, b.zTileNum, b.dColBias, colScan, ctx->aux2);
		CPVS_CUDA(cudaEventRecord(ctx->evCols, ctx->aux2));
	}
	// the large inner levels' tables are cleared up front (the leaf level sizes and clears its table on
	// the device, once the sketch is filled; the small levels clear theirs inside their kernel)
	// -- on a side stream, next to the expansion; the first inner insert waits for it.
	tablesClearing = false;
	for (int l = b.minLevel; l < b.smallLow; ++l)
		if (b.lv[l].cap && !b.leafLevel(l)) {
			if (!tablesClearing) {
				CPVS_CUDA(cudaEventRecord(ctx->evFork, st));  // the arena may still be in use by the previous build
				CPVS_CUDA(cudaStreamWaitEvent(ctx->aux3, ctx->evFork, 0));
				tablesClearing = true;
			}
			CPVS_CUDA(cudaMemsetAsync(b.lv[l].table, 0xFF, (b.lv[l].tableSlots + kDirectSlots) * sizeof(u64), ctx->aux3));
		}
	if (tablesClearing) CPVS_CUDA(cudaEventRecord(ctx->evClear, ctx->aux3));

	CPVS_CUDA(cudaEventRecord(b.phases.ev[CPVS_PHASE_EXPAND], st));
	storeCoordKernel<<<1, 1, 0, st>>>(b.lv[b.top].coords, packCoord(0, 0, b.zTileIndex * 2));
	++ctx->launches;
	{
		SmallExpandArgs sx;
		sx.count = 0;
		sx.rootN = b.dNodes() + b.top;
		sx.overflow = b.dOverflow();
		for (int l = b.top; l >= b.smallLow; --l) {
			SmallExpandLevel& e = sx.lv[sx.count++];
			e.side = (u32)b.mm->n >> l;
			e.tex = b.pyr.level[l];
			e.heightF = (float)(e.side * b.zTileNum);
			e.level0 = l == 0 ? 1 : 0;
			e.coords = b.lv[l].coords;
			e.masks = b.lv[l].masks;
			e.firstChild = b.lv[l].firstChild;
			const bool hasChildLevel = l > b.minLevel;
			e.childCoords = hasChildLevel ? b.lv[l - 1].coords : nullptr;
			e.childCap = hasChildLevel ? (u32)b.lv[l - 1].cap : 0u;
			e.childN = b.dNodes() + (hasChildLevel ? l - 1 : kMaxLevels - 1);  // (the bottom level has no children: a spare word)
			e.colBias = nullptr;
			e.leafAt = nullptr;
			if (b.leafColumns && l == 3) {
				e.colBias = b.dColBias;
				e.leafAt = b.lv[2].leafAt;
				CPVS_CUDA(cudaStreamWaitEvent(st, ctx->evCols, 0));
			}
		}
		ctx->launches += launchExpandSmallLevels(sx, st);
	}
	for (int l = b.smallLow - 1; l >= b.lastInner && b.lv[l].cap; --l) {
		const bool hasChildLevel = l > b.minLevel;
		const bool toColumns = b.leafColumns && l == 3;
		if (toColumns) CPVS_CUDA(cudaStreamWaitEvent(st, ctx->evCols, 0));
		ctx->launches += launchExpandLevel(b.pyr, l, b.zTileNum, b.lv[l].coords, b.dNodes() + l, b.lv[l].cap, b.lv[l].masks, b.lv[l].firstChild,
				hasChildLevel ? b.lv[l - 1].coords : nullptr, hasChildLevel ? b.lv[l - 1].cap : 0, b.dNodes() + (hasChildLevel ? l - 1 : kMaxLevels - 1),
				b.dOverflow(), b.nextScan(b.lv[l].cap, kExpandTileNodes), toColumns ? b.dColBias : nullptr, toColumns ? b.lv[2].leafAt : nullptr, st);
	}
	CPVS_CUDA(cudaEventRecord(b.phases.ev[CPVS_PHASE_LEAVES], st));
	if (b.leafColumns)  // constructLastLevels (src/CompressedShadow.cpp:171-190)
		ctx->launches += launchBuildLeafColumns(b.pyr, b.zTileIndex, b.zTileNum, b.dColBias, b.lv[2].leafAt, (u32)b.lv[2].cap, b.lv[2].leafCodes,
				b.lv[2].masks, b.expectedDistinctLeaves ? nullptr : b.dSketch, (b.mm->residue && b.mm->residueTiles == b.zTileNum) ? b.mm->residue : nullptr,
				reinterpret_cast<u32*>(b.dScalars + kSlotTallColumns), st);
	else if (b.haveLeaves)
		ctx->launches += launchBuildLeaves(b.pyr, b.zTileNum, b.lv[2].coords, b.dNodes() + 2, b.lv[2].cap, b.lv[2].leafCodes, b.lv[2].leafHash,
				b.lv[2].masks, b.expectedDistinctLeaves ? nullptr : b.dSketch, st);
	return CPVS_OK;
}

EmitLevelArgs emitArgs(const Build& b, int l) {
	const LevelArrays& a = b.lv[l];
	EmitLevelArgs em;
	em.n = a.hint;
	em.leaf = b.leafLevel(l) ? 1 : 0;
	em.fromEnd = 0;
	em.uniqueCount = b.dUnique() + l;
	em.wordCount = b.dWords() + l;
	em.firstList = a.firstList;
	em.wordOffset = a.wordOffset;
	em.levelBase = b.dBases() + l;
	em.totalWords = b.dTotal();
	em.leafCodes = a.leafCodes;
	em.masks = a.masks;
	em.firstChild = a.firstChild;
	em.childUid = l > b.minLevel ? b.lv[l - 1].uid : nullptr;
	em.childSlotOffset = l > b.minLevel ? b.lv[l - 1].slotOffset : nullptr;
	em.childLevelBase = b.dBases() + (l > b.minLevel ? l - 1 : l);
	em.dagAlloc = b.dagAlloc;
	em.capacity = b.dagCapacity;
	em.overflow = b.dOverflow();
	return em;
}

// ---- merge -----------------------------------------------------------------------------------------------------------------
// The chain of inserts is the critical path and runs on the high-priority stream, so that its CTAs are dispatched ahead of
// the queued CTAs of the rank scans running beside it; the main stream rejoins before the bases.
int stageMerge(Build& b, bool tablesClearing) {
	cpvs_ctx* ctx = b.ctx;
	cudaStream_t st = b.st, mergeStream = ctx->aux;
	u64 *dSketchBits = b.dScalars + kSlotSketchBits, *dLeafTableMask = b.dScalars + kSlotLeafTableMask;
	CPVS_CUDA(cudaEventRecord(ctx->evFork, st));
	CPVS_CUDA(cudaStreamWaitEvent(mergeStream, ctx->evFork, 0));
	CPVS_CUDA(cudaEventRecord(b.phases.ev[CPVS_PHASE_LEAF_TABLE], mergeStream));
	if (b.haveLeaves) {
		if (!b.expectedDistinctLeaves) ctx->launches += launchSketchPopcount(b.dSketch, dSketchBits, mergeStream);
		ctx->launches += launchSizeLeafTable(b.lv[2].table, b.lv[2].tableSlots, dSketchBits, dLeafTableMask, b.expectedDistinctLeaves, mergeStream);
	}
	CPVS_CUDA(cudaEventRecord(b.phases.ev[CPVS_PHASE_LEAF_INSERT], mergeStream));
	if (!b.haveLeaves) {
		CPVS_CUDA(cudaEventRecord(b.phases.ev[CPVS_PHASE_LEAF_RESOLVE], mergeStream));
		CPVS_CUDA(cudaEventRecord(b.phases.ev[CPVS_PHASE_INNER_MERGE], mergeStream));
	}
	// Per level: insert on the merge stream (gives every node its group id, all the next level needs),
	// rank on a side stream (orders the unique nodes; only the emission needs it).
	bool ranksPending = false;
	for (int l = b.minLevel; l < b.smallLow; ++l) {
		LevelArrays& a = b.lv[l];
		if (!a.cap) continue;
		const bool leafLevel = b.leafLevel(l);
		MergeLevelArgs m;
		m.nDev = b.dNodes() + l;
		m.cap = a.cap;
		m.leaf = leafLevel ? 1 : 0;
		m.leafCodes = a.leafCodes;
		m.leafHash = a.leafHash;
		m.masks = a.masks;
		m.firstChild = a.firstChild;
		m.childUid = l > b.minLevel ? b.lv[l - 1].uid : nullptr;
		m.table = a.table;
		m.tableSize = a.tableSlots;
		m.sketchBits = dSketchBits;
		m.tableMaskDev = dLeafTableMask;
		m.errorFlag = b.dTableError();
		m.overflow = b.dOverflow();
		m.uid = a.uid;
		m.firstList = a.firstList;
		m.wordOffset = a.wordOffset;
		m.slotOffset = a.slotOffset;
		m.sizeOf = a.sizeOf;
		m.uniqueCount = b.dUnique() + l;
		m.wordCount = b.dWords() + l;
		if (!leafLevel && tablesClearing) {
			CPVS_CUDA(cudaStreamWaitEvent(mergeStream, ctx->evClear, 0));
			tablesClearing = false;
		}
		ctx->launches += launchInsertLevel(m, mergeStream);
		if (leafLevel) {
			CPVS_CUDA(cudaEventRecord(b.phases.ev[CPVS_PHASE_LEAF_RESOLVE], mergeStream));
			CPVS_CUDA(cudaEventRecord(b.phases.ev[CPVS_PHASE_INNER_MERGE], mergeStream));
		}
		cudaStream_t rs = (l & 1) ? ctx->aux3 : ctx->aux2;
		CPVS_CUDA(cudaEventRecord(ctx->evFork, mergeStream));
		CPVS_CUDA(cudaStreamWaitEvent(rs, ctx->evFork, 0));
		if (leafLevel) {
			CPVS_CUDA(cudaEventCreate(&b.evRankStart));
			CPVS_CUDA(cudaEventCreate(&b.evRankStop));
			CPVS_CUDA(cudaEventCreate(&b.evLeafEmitStart));
			CPVS_CUDA(cudaEventRecord(b.evRankStart, rs));
		}
		ctx->launches += launchRankLevel(m, b.nextScan(a.cap), rs);
		if (leafLevel) {
			CPVS_CUDA(cudaEventRecord(b.evRankStop, rs));
			if (b.dagAlloc) {
				// The leaf level ends the DAG: with the allocation in hand it is written now, beside the inner levels' merge.
				CPVS_CUDA(cudaEventRecord(ctx->evLeafRanked, rs));
				CPVS_CUDA(cudaStreamWaitEvent(ctx->aux4, ctx->evLeafRanked, 0));
				EmitLevelArgs em = emitArgs(b, l);
				em.fromEnd = 1;
				CPVS_CUDA(cudaEventRecord(b.evLeafEmitStart, ctx->aux4));
				ctx->launches += launchEmitLevel(em, ctx->aux4);
				CPVS_CUDA(cudaEventRecord(b.phases.ev[CPVS_PHASE_EMIT_LEAVES], ctx->aux4));
				CPVS_CUDA(cudaEventRecord(ctx->evLeafEmitted, ctx->aux4));
				b.leavesEmitted = true;
			}
		}
		ranksPending = true;
	}
	{
		SmallMergeArgs sm;
		sm.count = 0;
		sm.table = b.dTable;
		sm.errorFlag = b.dTableError();
		sm.overflow = b.dOverflow();
		for (int l = b.smallLow; l <= b.top; ++l) {
			if (!b.lv[l].cap) continue;
			SmallMergeLevel& m = sm.lv[sm.count++];
			m.nDev = b.dNodes() + l;
			m.masks = b.lv[l].masks;
			m.firstChild = b.lv[l].firstChild;
			m.childUid = l > b.minLevel ? b.lv[l - 1].uid : nullptr;
			m.uid = b.lv[l].uid;
			m.firstList = b.lv[l].firstList;
			m.wordOffset = b.lv[l].wordOffset;
			m.slotOffset = b.lv[l].slotOffset;
			m.uniqueCount = b.dUnique() + l;
			m.wordCount = b.dWords() + l;
		}
		ctx->launches += launchMergeSmallLevels(sm, mergeStream);
	}
	if (ranksPending || tablesClearing) {  // join: the level sizes feed the bases
		CPVS_CUDA(cudaEventRecord(ctx->evJoin, ctx->aux2));
		CPVS_CUDA(cudaStreamWaitEvent(mergeStream, ctx->evJoin, 0));
		CPVS_CUDA(cudaEventRecord(ctx->evJoin3, ctx->aux3));
		CPVS_CUDA(cudaStreamWaitEvent(mergeStream, ctx->evJoin3, 0));
	}
	CPVS_CUDA(cudaEventRecord(ctx->evJoin, mergeStream));
	CPVS_CUDA(cudaStreamWaitEvent(st, ctx->evJoin, 0));
	return CPVS_OK;
}

// ---- emit ------------------------------------------------------------------------------------------------------------------
// Writes every unique node once, in its final place. The levels are independent of each other now; the leaf level (most of
// the words) stays on the main stream unless it was written during the merge, the inner levels go to the high-priority
// side stream next to it.
// phase EMIT_INNER: fork .. join on the main stream; phase EMIT_LEAVES: the leaf kernel alone (they overlap).
int stageEmit(Build& b) {
	cpvs_ctx* ctx = b.ctx;
	cudaStream_t st = b.st;
	CPVS_CUDA(cudaEventRecord(b.phases.ev[CPVS_PHASE_EMIT_INNER], st));
	const bool leafNow = b.haveLeaves && !b.leavesEmitted;
	if (leafNow) {
		CPVS_CUDA(cudaEventRecord(ctx->evFork, st));
		CPVS_CUDA(cudaStreamWaitEvent(ctx->aux, ctx->evFork, 0));
	}
	EmitMultiArgs inner;
	inner.count = 0;
	for (int l = b.minLevel; l <= b.top; ++l) {
		if (!b.lv[l].cap) continue;
		EmitLevelArgs em = emitArgs(b, l);
		if (b.leafLevel(l)) {
			if (!leafNow) continue;
			CPVS_CUDA(cudaEventRecord(b.evLeafEmitStart, st));
			ctx->launches += launchEmitLevel(em, st);
			CPVS_CUDA(cudaEventRecord(b.phases.ev[CPVS_PHASE_EMIT_LEAVES], st));  // closes the leaf kernel
		} else if (inner.count < kMaxEmitLevels) {
			inner.lv[inner.count++] = em;
		}
	}
	ctx->launches += launchEmitInnerLevels(inner, leafNow ? ctx->aux : st);
	if (leafNow) {
		CPVS_CUDA(cudaEventRecord(ctx->evJoin, ctx->aux));
		CPVS_CUDA(cudaStreamWaitEvent(st, ctx->evJoin, 0));
	} else if (b.leavesEmitted) {
		CPVS_CUDA(cudaStreamWaitEvent(st, ctx->evLeafEmitted, 0));
	} else {
		CPVS_CUDA(cudaEventRecord(b.phases.ev[CPVS_PHASE_EMIT_LEAVES], st));
	}
	return CPVS_OK;
}

// ---- enqueue / finish ----------------------------------------------------------------------------------------------------
// enqueueBuild puts one attempt at the whole build on the stream, up to the read-back of its sizes and flags; finishBuild waits
// for that read-back and turns it into the shadow's info -- or reports that a predicted capacity did not suffice (*redo).
// exactCounts == NULL: node capacities predicted from `memo`.
int enqueueBuild(Build& b, const u64* exactCounts, const SizeMemo* memo, cudaEvent_t countStart) {
	cpvs_ctx* ctx = b.ctx;
	b.L = b.mm->numLevels;
	b.top = b.L - 2;
	b.minLevel = b.useLeaf ? 2 : 0;  // src/CompressedShadow.cpp:30-32
	b.lastInner = b.useLeaf ? 3 : 0;
	b.pyr = pyramidView(b.mm);
	b.st = ctx->stream;
	b.predicted = exactCounts == nullptr;
	b.exact = exactCounts != nullptr;
	b.dScalars = ctx->scalars;
	b.countStart = countStart;
	cudaStream_t st = b.st;
	for (cudaEvent_t& e : b.phases.ev) CPVS_CUDA(cudaEventCreate(&e));
	CPVS_CUDA(cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming));
	if (!countStart) CPVS_CUDA(cudaEventRecord(b.phases.ev[CPVS_PHASE_COUNT], st));

	planLevels(b, exactCounts, memo);
	for (int l = b.minLevel; l <= b.top; ++l)
		if (b.lv[l].cap >= (1ull << 29)) return fail(CPVS_EOVERFLOW, "level %d has %llu nodes (limit 2^29)", l, (unsigned long long)b.lv[l].cap);
	CPVS_CUDA(cudaMemsetAsync(b.dScalars, 0, kNumScalars * sizeof(u64), st));
	if (int rc = carveArena(b)) return rc;
	b.trace.mark("plan + carve");

	// The DAG's allocation: with a memo of this shape its size is predicted and the leaf level can be written during the
	// merge; otherwise it is made once the sizes are known.
	if (exactCounts && ctx->predictSizes) {
		const u64 bound = upperBoundWords(exactCounts, b.top, b.minLevel, b.useLeaf);
		if (bound <= ctx->stagingMaxWords) {
			cudaError_t e = takeStaging(ctx, bound, &b.dagAlloc, &b.stagingWords);
			if (e != cudaSuccess) return fail(CPVS_ENOMEM, "DAG staging buffer of %llu words: %s", (unsigned long long)bound, cudaGetErrorString(e));
			b.staged = true;
			b.dagCapacity = bound;
		}
	}
	if (!b.dagAlloc && memo && ctx->predictSizes) {
		u64 expectWords = memo->words;
		if (exactCounts && memo->nodes[b.minLevel] && b.lv[b.minLevel].cap) {
			// this build's node counts are known: the words follow the bottom level (same scene, another tile or slice)
			const double ratio = (double)b.lv[b.minLevel].cap / (double)memo->nodes[b.minLevel];
			expectWords = (u64)((double)memo->words * ratio) + 4096;
		}
		// (scaled from another tile or slice: twice the head room -- a DAG that outgrows its allocation while later builds are
		// already using the arena costs a whole rebuild)
		const unsigned shift = exactCounts && ctx->headroomShift > 1 ? ctx->headroomShift - 1 : ctx->headroomShift;
		b.dagCapacity = expectWords + (expectWords >> shift) + 1024;
		cudaError_t e = ctxAlloc(ctx, reinterpret_cast<void**>(&b.dagAlloc), b.dagCapacity * sizeof(u32));
		if (e != cudaSuccess) return fail(CPVS_ENOMEM, "DAG allocation of %llu words: %s", (unsigned long long)b.dagCapacity, cudaGetErrorString(e));
	}
	bool tablesClearing = false;
	if (int rc = stageExpand(b, tablesClearing)) return rc;
	if (int rc = stageMerge(b, tablesClearing)) return rc;
	CPVS_CUDA(cudaEventRecord(b.phases.ev[CPVS_PHASE_BASES], st));
	ctx->launches += launchLevelBases(b.dWords(), b.dBases(), b.top, b.minLevel, b.dTotal(), b.dagAlloc ? b.dagCapacity : ~0ull, b.dOverflow(),
			b.lv[b.top].masks, reinterpret_cast<u32*>(b.dScalars + kSlotRootMask), st);
	b.trace.mark("enqueue expand..bases");

	if (ctx->readbackFree.empty()) {
		u64* slot = nullptr;
		CPVS_CUDA(cudaMallocHost(reinterpret_cast<void**>(&slot), kNumScalars * sizeof(u64)));
		ctx->readbackAll.push_back(slot);
		ctx->readbackFree.push_back(slot);
	}
	b.h = ctx->readbackFree.back();
	ctx->readbackFree.pop_back();
	if (!b.dagAlloc) {  // sizes first, then an exact allocation
		CPVS_CUDA(cudaMemcpyAsync(b.h, b.dScalars, kNumScalars * sizeof(u64), cudaMemcpyDeviceToHost, st));
		CPVS_CUDA(cudaStreamSynchronize(st));
		CPVS_CUDA(cudaGetLastError());
		b.trace.mark("sync after bases");
		if (!(u32)b.h[kSlotTableError] && !((u32)b.h[kSlotOverflow] & kOverflowNodes)) {  // (finishBuild reports those)
			b.dagCapacity = b.h[kSlotTotal];
			if (b.dagCapacity > (1ull << 32)) return fail(CPVS_EOVERFLOW, "DAG needs %llu words; offsets are 32-bit", (unsigned long long)b.dagCapacity);
			cudaError_t e = ctxAlloc(ctx, reinterpret_cast<void**>(&b.dagAlloc), b.dagCapacity * sizeof(u32));
			if (e != cudaSuccess) return fail(CPVS_ENOMEM, "DAG allocation of %llu words: %s", (unsigned long long)b.dagCapacity, cudaGetErrorString(e));
			for (int l = b.minLevel; l <= b.top; ++l) b.lv[l].hint = b.h[kSlotUnique + l];
		}
	}
	if (b.dagAlloc)
		if (int rc = stageEmit(b)) return rc;
	CPVS_CUDA(cudaMemcpyAsync(b.h, b.dScalars, kNumScalars * sizeof(u64), cudaMemcpyDeviceToHost, st));
	CPVS_CUDA(cudaEventRecord(b.phases.ev[CPVS_NUM_PHASES], st));
	CPVS_CUDA(cudaEventRecord(b.done, st));
	return CPVS_OK;
}

// arenaIntact: no later build has been enqueued on the context, so a DAG that outgrew its predicted allocation can be
// emitted again from the merge results still in the arena; otherwise that, too, asks for a rebuild.
int finishBuild(Build& b, cpvs_shadow* s, bool arenaIntact, bool* redo, bool deferCopy = false) {
	cpvs_ctx* ctx = b.ctx;
	cudaStream_t st = b.st;
	const u64* h = b.h;
	*redo = false;
	for (int attempt = 0;; ++attempt) {
		cudaError_t e = cudaEventSynchronize(b.done);
		if (e == cudaSuccess) e = cudaGetLastError();
		if (e != cudaSuccess) return fail(CPVS_ECUDA, "cpvs_shadow_create: %s", cudaGetErrorString(e));
		b.trace.mark("emit + final sync");
		if ((u32)h[kSlotTableError]) {
			if (b.predicted) {  // far more distinct leaves than the memo promised: the table was too small
				*redo = true;
				return CPVS_OK;
			}
			return fail(CPVS_EINTERNAL, "merge table overflow (leaf table mask %llu)", (unsigned long long)h[kSlotLeafTableMask]);
		}
		const u32 overflow = (u32)h[kSlotOverflow];
		if (overflow & kOverflowNodes) {
			if (b.predicted) {
				*redo = true;
				return CPVS_OK;
			}
			return fail(CPVS_EINTERNAL, "the expansion produced more nodes than the closed-form count predicted");
		}
		const u64 totalWords = h[kSlotTotal];
		if (totalWords > (1ull << 32)) return fail(CPVS_EOVERFLOW, "DAG needs %llu words; offsets are 32-bit", (unsigned long long)totalWords);
		if (b.exact)
			for (int l = b.top; l > b.minLevel; --l)
				if (b.lv[l].cap && l >= b.lastInner && h[kSlotNodes + l - 1] != b.lv[l - 1].cap)
					return fail(CPVS_EINTERNAL, "level %d: expansion produced %llu nodes, count pass predicted %llu", l - 1,
							(unsigned long long)h[kSlotNodes + l - 1], (unsigned long long)b.lv[l - 1].cap);
		if ((overflow & kOverflowWords) && b.staged) return fail(CPVS_EINTERNAL, "DAG of %llu words exceeds the bound of its node counts", (unsigned long long)totalWords);
		if (overflow & kOverflowWords) {
			// The predicted allocation was too small.
			if (attempt) return fail(CPVS_EINTERNAL, "DAG of %llu words did not fit an exact allocation", (unsigned long long)totalWords);
			if (!arenaIntact) {
				*redo = true;
				return CPVS_OK;
			}
			// The merge results are all still in the arena: emit again into an exact allocation.
			++ctx->reemissions;
			ctxFree(ctx, b.dagAlloc);
			b.dagAlloc = nullptr;
			b.dagCapacity = totalWords;
			e = ctxAlloc(ctx, reinterpret_cast<void**>(&b.dagAlloc), b.dagCapacity * sizeof(u32));
			if (e != cudaSuccess) return fail(CPVS_ENOMEM, "DAG allocation of %llu words: %s", (unsigned long long)b.dagCapacity, cudaGetErrorString(e));
			CPVS_CUDA(cudaMemsetAsync(b.dOverflow(), 0, sizeof(u32), st));
			b.leavesEmitted = false;
			for (int l = b.minLevel; l <= b.top; ++l) b.lv[l].hint = h[kSlotUnique + l];
			if (int rc = stageEmit(b)) return rc;
			CPVS_CUDA(cudaMemcpyAsync(b.h, b.dScalars, kNumScalars * sizeof(u64), cudaMemcpyDeviceToHost, st));
			CPVS_CUDA(cudaEventRecord(b.phases.ev[CPVS_NUM_PHASES], st));
			CPVS_CUDA(cudaEventRecord(b.done, st));
			continue;
		}
		const u32 rootMask = (u32)h[kSlotRootMask];

		float ms = 0.f, phaseMs[CPVS_NUM_PHASES] = {0};
		cudaEvent_t* ev = b.phases.ev;
		cudaEvent_t first = b.countStart ? b.countStart : ev[CPVS_PHASE_COUNT];
		e = cudaEventElapsedTime(&ms, first, ev[CPVS_NUM_PHASES]);
		if (e == cudaSuccess) e = cudaEventElapsedTime(&phaseMs[CPVS_PHASE_COUNT], first, ev[CPVS_PHASE_EXPAND]);
		for (int i = CPVS_PHASE_EXPAND; i < CPVS_PHASE_EMIT_INNER && e == cudaSuccess; ++i) e = cudaEventElapsedTime(&phaseMs[i], ev[i], ev[i + 1]);
		if (e == cudaSuccess) e = cudaEventElapsedTime(&phaseMs[CPVS_PHASE_EMIT_INNER], ev[CPVS_PHASE_EMIT_INNER], ev[CPVS_NUM_PHASES]);
		if (e == cudaSuccess && b.haveLeaves) e = cudaEventElapsedTime(&phaseMs[CPVS_PHASE_EMIT_LEAVES], b.evLeafEmitStart, ev[CPVS_PHASE_EMIT_LEAVES]);
		if (e == cudaSuccess && b.haveLeaves) e = cudaEventElapsedTime(&phaseMs[CPVS_PHASE_LEAF_RESOLVE], b.evRankStart, b.evRankStop);
		if (e != cudaSuccess) return fail(CPVS_ECUDA, "cpvs_shadow_create: %s", cudaGetErrorString(e));

		if (b.staged) {
			// out of the staging buffer into an allocation of the DAG's size, on the context's copy stream: nothing that was
			// enqueued on the build stream in the meantime (the next builds) is waited for
			u32* exact = nullptr;
			const u32* from = b.dagAlloc + (b.dagCapacity - totalWords);
			const u64 shift = (reinterpret_cast<uintptr_t>(from) & 15u) >> 2;  // same alignment on both sides: 16-byte copies
			size_t have = (totalWords + shift) * sizeof(u32);
			{  // a released allocation of about this size, if there is one (under memory pressure the pool remaps pages: ~1 ms per 200 MB)
				std::lock_guard<std::mutex> guard(ctx->cacheLock);
				size_t best = ctx->dagFree.size();
				for (size_t i = 0; i < ctx->dagFree.size(); ++i) {
					const size_t bytes = ctx->dagFree[i].second;
					if (bytes >= have && bytes - have <= have / 4 && (best == ctx->dagFree.size() || bytes < ctx->dagFree[best].second)) best = i;
				}
				if (best != ctx->dagFree.size()) {
					exact = ctx->dagFree[best].first;
					have = ctx->dagFree[best].second;
					ctx->dagFreeBytes -= have;
					ctx->dagFree.erase(ctx->dagFree.begin() + best);
				}
			}
			e = exact ? cudaSuccess : cudaMallocAsync(reinterpret_cast<void**>(&exact), have, ctx->copyStream);
			if (e != cudaSuccess) return fail(CPVS_ENOMEM, "DAG allocation of %llu words: %s", (unsigned long long)totalWords, cudaGetErrorString(e));
			s->dagAllocBytes = have;
			b.trace.mark("staged: allocation");
			ctx->launches += launchCopyWords(exact + shift, from, totalWords, ctx->copyStream);
			e = cudaGetLastError();
			if (e == cudaSuccess && !deferCopy) e = cudaStreamSynchronize(ctx->copyStream);
			s->copyInFlight = deferCopy;  // (the caller then keeps the Build, and with it the staging buffer, until the copy is done)
			if (e != cudaSuccess) {
				cudaFreeAsync(exact, ctx->copyStream);
				return fail(CPVS_ECUDA, "cpvs_shadow_create: %s", cudaGetErrorString(e));
			}
			b.trace.mark("staged: copy (+ wait)");
			s->dagAlloc = exact;
			s->dag = exact + shift;  // (the staging buffer goes back to the context with the Build)
			s->dagOnCopyStream = true;
		} else {
			s->dagAlloc = b.dagAlloc;
			s->dag = b.dagAlloc + (b.dagCapacity - totalWords);
			b.dagAlloc = nullptr;  // owned by the shadow now
		}
		std::memset(&s->info, 0, sizeof(s->info));
		s->info.num_levels = (u32)b.L;
		s->info.leafmasks = b.useLeaf ? 1 : 0;
		s->info.total_visibility = rootMask == 0x5555u ? CPVS_VISIBLE : (rootMask == 0u ? CPVS_SHADOW : CPVS_PARTIAL);
		s->info.words = totalWords;
		for (int l = b.minLevel; l <= b.top; ++l) {
			s->info.svo_nodes[l] = h[kSlotNodes + l];
			s->info.dag_nodes[l] = h[kSlotUnique + l];
			s->info.dag_words[l] = h[kSlotWords + l];
		}
		s->info.build_ms = ms;
		for (int i = 0; i < CPVS_NUM_PHASES; ++i) s->info.phase_ms[i] = phaseMs[i];
		s->info.predicted = b.predicted ? 1u : 0u;
		rememberSizes(ctx, b, h);
		return CPVS_OK;
	}
}

// The synchronous path with exact counts (first build of a shape, z-slices of a tile column, and every rebuild). The
// column's counts are computed once per hierarchy; a slice that misses the surface is its root's mask word.
int buildExact(cpvs_ctx* ctx, const cpvs_minmax* mm, u32 zTileIndex, u32 zTileNum, bool useLeaf, cpvs_shadow* s, bool async);

}  // namespace

// A build in flight (cpvs_shadow_create_async): everything finishBuild needs, and what a rebuild would.
struct cpvs_pending_build {
	Build b;
	bool finished = false;  // only the copy out of the staging buffer is still in flight (shadowWaitBegin)
	const cpvs_minmax* mm;
	u64 serial;  // the context's build counter when this one was enqueued
};

namespace {

// async: with a memo of this shape the DAG's allocation is predicted, nothing is read back inside the build, and the call may
// return with the build in flight (s->pending).
int buildExact(cpvs_ctx* ctx, const cpvs_minmax* mm, u32 zTileIndex, u32 zTileNum, bool useLeaf, cpvs_shadow* s, bool async) {
	const int L = mm->numLevels, top = L - 2, minLevel = useLeaf ? 2 : 0;
	cudaStream_t st = ctx->stream;
	const PyramidView pyr = pyramidView(mm);
	SizeMemo memoCopy;
	const SizeMemo* memo = findMemo(ctx, mm->n, zTileIndex, zTileNum, useLeaf ? 1 : 0, &memoCopy) ? &memoCopy : nullptr;
	async = async && ctx->predictSizes;  // (whether anything stands in the way is known once the counts are)
	struct EventGuard {
		cudaEvent_t ev = nullptr;
		~EventGuard() {
			if (ev) cudaEventDestroy(ev);
		}
	} countStart;
	u64 counts[kMaxLevels] = {0};
	const bool cached = mm->columnSlices == zTileNum && mm->columnMinLevel == minLevel;
	if (!cached && !async) {
		CPVS_CUDA(cudaEventCreate(&countStart.ev));
		CPVS_CUDA(cudaEventRecord(countStart.ev, st));
	}
	const u64* column = nullptr;
	if (int rc = columnCounts(ctx, mm, pyr, zTileNum, minLevel, &column)) return rc;
	const u64* mine = column + (size_t)zTileIndex * kMaxLevels;
	for (int l = 0; l < kMaxLevels; ++l) counts[l] = mine[l];
	if (counts[top - 1] == 0) return oneWordShadow(ctx, L, useLeaf, (u32)counts[kRootMaskScalar], s);
	++ctx->exactBuilds;
	++ctx->buildSerial;
	async = async && (memo || upperBoundWords(counts, top, minLevel, useLeaf) <= ctx->stagingMaxWords);
	if (async) {
		cpvs_pending_build* p = new (std::nothrow) cpvs_pending_build();
		if (!p) return fail(CPVS_ENOMEM, "cpvs_shadow_create: host allocation");
		p->mm = mm;
		p->b.ctx = ctx;
		p->b.mm = mm;
		p->b.zTileIndex = zTileIndex;
		p->b.zTileNum = zTileNum;
		p->b.useLeaf = useLeaf;
		p->serial = ctx->buildSerial;
		s->pending = p;
		return enqueueBuild(p->b, counts, memo, nullptr);
	}
	Build b;
	b.ctx = ctx;
	b.mm = mm;
	b.zTileIndex = zTileIndex;
	b.zTileNum = zTileNum;
	b.useLeaf = useLeaf;
	if (int rc = enqueueBuild(b, counts, memo, countStart.ev)) return rc;
	bool redo = false;
	return finishBuild(b, s, true, &redo);
}

}  // namespace

namespace cpvs {
int columnCountsBegin(cpvs_ctx* ctx, const cpvs_minmax* mm, u32 zTileNum, int minLevel) {
	CPVS_CUDA(cudaSetDevice(ctx->device));
	return ::columnCountsBegin(ctx, mm, pyramidView(mm), zTileNum, minLevel);
}
int columnCountsOf(cpvs_ctx* ctx, const cpvs_minmax* mm, u32 zTileNum, int minLevel, const u64** counts) {
	CPVS_CUDA(cudaSetDevice(ctx->device));
	return columnCounts(ctx, mm, pyramidView(mm), zTileNum, minLevel, counts);
}
PyramidView pyramidView(const cpvs_minmax* mm) {
	PyramidView pyr;
	pyr.n = mm->n;
	pyr.numLevels = mm->numLevels;
	for (int k = 0; k < kMaxLevels; ++k) pyr.level[k] = k < mm->numLevels ? mm->level[k] : nullptr;
	return pyr;
}
}  // namespace cpvs

extern "C" {

static int createImpl(cpvs_ctx* ctx, const cpvs_minmax* mm, uint32_t zTileIndex, uint32_t zTileNum, int leafmasks, bool async, cpvs_shadow** out) {
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
	const bool useLeaf = leafmasks && (L - 3) >= 2;  // src/CompressedShadow.cpp:20-27
	if (!useLeaf)
		if (int rc = ensureLowLevels(mm, 1)) return rc;  // the leafmask-less octree descends through levels 2 and 1

	cpvs_shadow* s = newShadow(ctx);
	if (!s) return fail(CPVS_ENOMEM, "cpvs_shadow_create: host allocation");
	std::unique_lock<std::mutex> buildGuard(ctx->buildLock);
	int rc = CPVS_OK;
	// Whole-volume builds of a shape seen before run on predicted sizes. Slices of a tile column always take the column's
	// exact counts: one launch and one read-back serve all of them, and most slices turn out to be a single word.
	SizeMemo memoCopy;
	const bool memo = findMemo(ctx, mm->n, zTileIndex, zTileNum, useLeaf ? 1 : 0, &memoCopy);
	if (memo && ctx->predictSizes && zTileNum == 1) {
		++ctx->predictedBuilds;
		cpvs_pending_build* p = new (std::nothrow) cpvs_pending_build();
		if (!p) {
			delete s;
			return fail(CPVS_ENOMEM, "cpvs_shadow_create: host allocation");
		}
		p->mm = mm;
		p->b.ctx = ctx;
		p->b.mm = mm;
		p->b.zTileIndex = zTileIndex;
		p->b.zTileNum = zTileNum;
		p->b.useLeaf = useLeaf;
		p->serial = ++ctx->buildSerial;
		s->pending = p;
		s->pendingLeafmasks = leafmasks;
		rc = enqueueBuild(p->b, nullptr, &memoCopy, nullptr);
		if (rc == CPVS_OK && async) {
			*out = s;
			return CPVS_OK;
		}
		buildGuard.unlock();
		if (rc == CPVS_OK) rc = cpvs_shadow_wait(s);
	} else {
		rc = buildExact(ctx, mm, zTileIndex, zTileNum, useLeaf, s, async);
		if (rc == CPVS_OK && s->pending) {  // in flight
			*out = s;
			return CPVS_OK;
		}
	}
	if (rc != CPVS_OK) {
		if (buildGuard.owns_lock()) buildGuard.unlock();
		cpvs_shadow_destroy(s);
		return rc;
	}
	*out = s;
	return CPVS_OK;
}

int cpvs_shadow_create(cpvs_ctx* ctx, const cpvs_minmax* mm, uint32_t zTileIndex, uint32_t zTileNum, int leafmasks, cpvs_shadow** out) {
	return createImpl(ctx, mm, zTileIndex, zTileNum, leafmasks, false, out);
}

int cpvs_shadow_create_async(cpvs_ctx* ctx, const cpvs_minmax* mm, uint32_t zTileIndex, uint32_t zTileNum, int leafmasks, cpvs_shadow** out) {
	return createImpl(ctx, mm, zTileIndex, zTileNum, leafmasks, true, out);
}

static int waitImpl(cpvs_shadow* s, bool deferCopy) {
	if (!s) return fail(CPVS_EINVAL, "cpvs_shadow_wait: NULL argument");
	if (!s->pending) return s->status;
	cpvs_ctx* ctx = s->ctx;
	CPVS_CUDA(cudaSetDevice(ctx->device));
	std::lock_guard<std::mutex> guard(ctx->buildLock);
	cpvs_pending_build* p = s->pending;
	if (!p) return s->status;
	int rc = s->status;
	if (!p->finished) {
		bool redo = false;
		rc = finishBuild(p->b, s, p->serial == ctx->buildSerial, &redo, deferCopy);
		if (rc == CPVS_OK && redo) {  // a predicted capacity did not suffice: once more, with exact counts
			++ctx->overflowRebuilds;
			const cpvs_minmax* mm = p->mm;
			const Build& b = p->b;
			s->pending = nullptr;
			rc = buildExact(ctx, mm, b.zTileIndex, b.zTileNum, b.useLeaf, s, false);
		}
		s->status = rc;
		if (rc != CPVS_OK) s->statusText = cpvs_last_error();
		if (rc == CPVS_OK && s->copyInFlight && deferCopy) {
			p->finished = true;
			s->pending = p;
			return rc;
		}
	}
	if (s->copyInFlight) {
		s->copyInFlight = false;
		const cudaError_t e = cudaStreamSynchronize(ctx->copyStream);
		if (e != cudaSuccess) {
			rc = s->status = fail(CPVS_ECUDA, "cpvs_shadow_create: %s", cudaGetErrorString(e));
			s->statusText = cpvs_last_error();
		}
	}
	s->pending = nullptr;
	delete p;
	return rc;
}

int cpvs_shadow_wait(cpvs_shadow* s) { return waitImpl(s, false); }

}  // extern "C"
namespace cpvs {
int shadowWaitBegin(cpvs_shadow* s) { return waitImpl(s, true); }
}  // namespace cpvs
extern "C" {

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
	if (s->pending) {  // abandon a build in flight: its kernels still use the arena and the allocation, in stream order
		std::lock_guard<std::mutex> guard(s->ctx->buildLock);
		cpvs_pending_build* p = s->pending;
		s->pending = nullptr;
		delete p;
	}
	if (s->copyInFlight) cudaStreamSynchronize(s->ctx->copyStream);
	if (s->dagOnCopyStream) {
		// allocated on the copy stream and given back there, so that the next such allocation finds it without depending on
		// what the build stream still has queued -- behind the work enqueued on the build stream so far, like every release
		cpvs_ctx* ctx = s->ctx;
		cudaEventRecord(ctx->evCopyFree, ctx->stream);
		cudaStreamWaitEvent(ctx->copyStream, ctx->evCopyFree, 0);
		std::vector<u32*> drop;
		{
			std::lock_guard<std::mutex> guard(ctx->cacheLock);
			if (s->dagAllocBytes >= kDagKeepMinBytes) {
				ctx->dagFree.emplace_back(s->dagAlloc, s->dagAllocBytes);
				ctx->dagFreeBytes += s->dagAllocBytes;
			} else {
				drop.push_back(s->dagAlloc);
			}
			while (!ctx->dagFree.empty() && (ctx->dagFreeBytes > kDagKeepMaxBytes || ctx->dagFree.size() > kDagKeepMaxBlocks)) {
				drop.push_back(ctx->dagFree.front().first);
				ctx->dagFreeBytes -= ctx->dagFree.front().second;
				ctx->dagFree.erase(ctx->dagFree.begin());
			}
		}
		for (u32* p : drop) cudaFreeAsync(p, ctx->copyStream);
	} else {
		ctxFree(s->ctx, s->dagAlloc);
	}
	if (s->skip) cudaFreeAsync(s->skip, s->ctx->stream);
	freeLookupIndex(s->ctx, &s->index);
	if (s->ready) cudaEventDestroy(s->ready);
	delete s;
	return CPVS_OK;
}

}  // extern "C"
