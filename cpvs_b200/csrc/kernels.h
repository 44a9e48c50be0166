// Host-callable launchers of the cpvs_b200 kernels. Every launcher enqueues on `stream` and returns
// the number of kernels it launched (for cpvs_ctx_launch_count).
#pragma once
#include "common.cuh"

namespace cpvs {

// ---- pyramid.cu: MinMaxHierarchy (reference src/MinMaxHierarchy.cpp:9-97) ----
// levels[k] (k >= 1) points at (n>>k)^2 float2 (min,max); levels[0] = depth.
// afterBase (optional) is recorded right after the fused base kernel. With writeLowLevels == false (and
// n >= 128) levels 1 and 2 are left unwritten; launchPyramidLowLevels produces them later if needed.
// residue (optional, n >= 128): n * n bytes, the depth map re-encoded per 8x8 column for the per-column leaf builder of
// builds with `residueTiles` z-slices (see pyramid.cu).
int launchPyramid(const float* depth, int n, float* const* levels, int numLevels, bool writeLowLevels, unsigned char* residue, unsigned residueTiles,
		cudaEvent_t afterBase, cudaStream_t stream);
int launchPyramidLowLevels(const float* depth, int n, float* const* levels, cudaStream_t stream);

// ---- svo.cu: constructSvo (reference src/CompressedShadow.cpp:87-190) ----
struct PyramidView {
	const float* level[kMaxLevels];  // level[0] = depth, level[k] = (min,max) pairs
	int n;
	int numLevels;
};

// cs::createChildmask for one node (known-answer tests): *out = 16-bit mask of the node at `level`.
int launchChildmask(const PyramidView& pyr, int level, u32 zTileNum, u32 x, u32 y, u32 z, u32* out, cudaStream_t stream);

// Closed-form node counts of every level for ALL z-slices of the column at once (createShadowTiles builds them from one
// pyramid): counts[z * kMaxLevels + l] += number of SVO nodes at level l of slice z, for l in [minLevel, numLevels-3];
// counts[z * kMaxLevels + kRootMaskScalar] = 1<<32 | the root's child mask of slice z (a slice without nodes below the
// root is that one word). counts must be zeroed (zTileNum * kMaxLevels words).
constexpr int kRootMaskScalar = 31;
int launchColumnCounts(const PyramidView& pyr, u32 zTileNum, int minLevel, u64* counts, cudaStream_t stream);

constexpr u64 kExpandTileNodes = 128 * kScanItems;  // nodes per look-back tile of launchExpandLevel
// One breadth-first step: masks of the *nDev nodes of `level`, index of each node's first child in the
// next level, and the next level's coordinate list. *childN receives the next level's node count.
// Every size is read on the device: `cap` (>= *nDev) only sizes the grid, `childCap` bounds the writes into the next
// level's arrays -- if the level turns out larger, *childN is clamped to childCap and bit 0 of *overflow is set (the
// host then rebuilds with exact counts).
// With leafAt != NULL the children are leaves built per column (launchBuildLeafColumns): instead of their
// coordinates, each child's index is stored at its column-order position leafAt[colBias[column] + z].
constexpr u32 kOverflowNodes = 1u, kOverflowWords = 2u;  // bits of the build's overflow word
int launchExpandLevel(const PyramidView& pyr, int level, u32 zTileNum, const u64* coords, const u64* nDev, u64 cap, u16* masks,
		u32* firstChild, u64* childCoords, u64 childCap, u64* childN, u32* overflow, ScanLaunch scan, const u32* colBias, u32* leafAt,
		cudaStream_t stream);

// The small top levels (<= kSmallMaxNodes nodes each, root first) expanded by a single CTA.
constexpr int kSmallThreads = 1024;
constexpr u32 kSmallMaxNodes = 4096;
struct SmallExpandLevel {
	const float* tex;  // pyramid level the children are classified against
	u32 side;
	float heightF;
	int level0;
	const u64* coords;
	u16* masks;
	u32* firstChild;
	u64* childCoords;  // next level's coordinate list (may be null when no child can exist)
	u64* childN;       // out: the next level's node count, clamped to childCap
	u32 childCap;      // capacity of the next level's arrays
	const u32* colBias;  // with leafAt: the children are leaves built per column (see launchExpandLevel)
	u32* leafAt;
};
struct SmallExpandArgs {
	SmallExpandLevel lv[kMaxLevels];
	int count;
	u64* rootN;     // out: node count of the first level (1)
	u32* overflow;  // see launchExpandLevel
};
int launchExpandSmallLevels(const SmallExpandArgs& a, cudaStream_t stream);

// Level-2 nodes: the leaf's k-code (codes[leaf*8 + row], nibble x = lit slices of texel (x,row)), a
// 64-bit hash of it and the 16-bit 1x1x8 childmask (2 bits per slice).
// Also sets one bit per leaf hash in `sketch` (kSketchWords zeroed words; NULL: skipped): a linear-counting estimate of
// the number of distinct leaves, used to size the merge table so that it stays resident in L2.
constexpr u32 kSketchWords = 1u << 22;  // 2^27 bits, 16 MiB
int launchBuildLeaves(const PyramidView& pyr, u32 zTileNum, const u64* coords, const u64* nDev, u64 cap, u32* codes, u64* hashes, u16* masks,
		u32* sketch, cudaStream_t stream);
// The same, one column of leaves (all z-blocks over an 8x8 texel block) at a time: every depth row is read once.
// launchColumnBias: colBias[c] = (leaves in the columns in front of c, row-major over the (n/8)^2 columns) - first
// z-block of c; one look-back scan over pyramid level 3. leafAt[colBias[c] + zb] = index of leaf (c, zb) in the
// level, written by the expansion of level 3.
int launchColumnBias(const PyramidView& pyr, u32 zTileIndex, u32 zTileNum, u32* colBias, ScanLaunch scan, cudaStream_t stream);
// (No hash array: the insert derives its hash from the code when MergeLevelArgs::leafHash is NULL.)
// residue: the hierarchy's re-encoded depth map for this zTileNum, or NULL (then, and for columns taller than 31 z-blocks,
// the depth map itself is read). tallFlag: one zeroed device word of scratch.
int launchBuildLeafColumns(const PyramidView& pyr, u32 zTileIndex, u32 zTileNum, const u32* colBias, const u32* leafAt, u32 numLeaves, u32* codes,
		u16* masks, u32* sketch, const unsigned char* residue, u32* tallFlag, cudaStream_t stream);
int launchSketchPopcount(const u32* sketch, u64* setBits, cudaStream_t stream);

// ---- merge.cu: mergeCommonSubtrees (reference src/CompressedShadow.cpp:215-304, Util.h:154-182) ----
// Inner-level tables carry kDirectSlots extra slots behind the tableSize hashed ones (merge.cu).
constexpr u32 kDirectSlots = 256;
struct MergeLevelArgs {
	const u64* nDev;       // device: nodes in this level
	u64 cap;               // host: upper bound of *nDev (sizes the grids)
	int leaf;              // 1: level of leafmask nodes
	const u32* leafCodes;  // leaf: k-code, 8 words per node
	const u64* leafHash;   // leaf: content hash per node, or NULL (then computed from the code)
	const u16* masks;      // inner: childmask per node
	const u32* firstChild; // inner: index of first child in the level below
	const u32* childUid;   // inner: unique id of every node of the level below
	u64* table;            // open-addressing table, pre-filled with 0xFF bytes (inner levels) or cleared by
	                       // the sizing kernel (leaves)
	u64 tableSize;         // allocated slots, power of two
	const u64* sketchBits; // leaf level: set bits of the distinct-count sketch (device)
	u64* tableMaskDev;     // leaf level: (chosen capacity - 1), written by the sizing kernel (device)
	u32* errorFlag;        // device: set if a probe sequence wraps the whole table
	const u32* overflow;   // device: kOverflowNodes set by the expansion = the level arrays are incomplete; merge nothing
	u32* uid;              // out (insert): group id per node = its group's table slot
	u32* firstList;        // out (rank): node index of the r-th unique node
	u32* wordOffset;       // out (rank): compressed word offset (inside the level) of the r-th unique node
	u32* slotOffset;       // out (rank): per table slot, the word offset of the group's node
	unsigned char* sizeOf; // scratch (rank): compressed size of node j if it is a first occurrence, else 0
	u64* uniqueCount;      // out: number of unique nodes
	u64* wordCount;        // out: compressed words of the level
};
// The small top levels merged bottom-up by a single CTA (same result as launchMergeLevel per level).
struct SmallMergeLevel {
	const u64* nDev;  // device: nodes in this level (<= kSmallMaxNodes)
	const u16* masks;
	const u32* firstChild;
	const u32* childUid;
	u32* uid;
	u32* firstList;
	u32* wordOffset;
	u32* slotOffset;  // 2 * kSmallMaxNodes entries
	u64* uniqueCount;
	u64* wordCount;
};
struct SmallMergeArgs {
	SmallMergeLevel lv[kMaxLevels];
	int count;
	u64* table;  // >= 2 * kSmallMaxNodes slots
	u32* errorFlag;
	const u32* overflow;  // see MergeLevelArgs
};
int launchMergeSmallLevels(const SmallMergeArgs& a, cudaStream_t stream);

// Leaf level only: picks the table capacity from the sketch's set-bit count -- or from expectedDistinct (> 0: the previous
// build's count of distinct leaves plus head room; no sketch then) -- and clears that many slots.
int launchSizeLeafTable(u64* table, u64 maxSlots, const u64* setBits, u64* tableMaskDev, u64 expectedDistinct, cudaStream_t stream);
// Insert assigns group ids (all the parent level needs); rank orders the unique nodes and may run
// concurrently with the next level's insert as long as this level's table is left alone.
int launchInsertLevel(const MergeLevelArgs& a, cudaStream_t stream);
int launchRankLevel(const MergeLevelArgs& a, ScanLaunch scan, cudaStream_t stream);

// ---- emit.cu: compress (reference src/CompressedShadow.cpp:326-392) ----
// The DAG is written at the END of an allocation of `capacity` words (the capacity may be a prediction made before the
// sizes were known): its first word is dagAlloc[capacity - *totalWords]. The leaf level ends the DAG, so it can be
// written as soon as it is ranked (fromEnd: level start = capacity - *wordCount), next to the merge of the inner levels.
// Nothing is written when the words do not fit (launchLevelBases reports that).
struct EmitLevelArgs {
	u64 n;                    // expected unique nodes of the level (sizes the grid; any value is correct)
	int leaf;
	int fromEnd;              // the level is the DAG's last: place it from the end of the allocation, totalWords not needed
	const u64* uniqueCount;   // device: unique nodes of this level
	const u32* firstList;
	const u32* wordOffset;
	const u64* wordCount;     // device: compressed words of this level
	const u64* levelBase;     // device: word offset of this level in the DAG
	const u64* totalWords;    // device: words of the whole DAG
	const u32* leafCodes;
	const u16* masks;
	const u32* firstChild;
	const u32* childUid;        // group ids (table slots) of the level below
	const u32* childSlotOffset; // word offset per slot of the level below
	const u64* childLevelBase;  // device: word offset of the level below in the DAG
	u32* dagAlloc;
	u64 capacity;
	const u32* overflow;  // device: kOverflowNodes set = the level arrays are incomplete; write nothing
};
int launchEmitLevel(const EmitLevelArgs& a, cudaStream_t stream);
// Several inner levels in one launch (block ranges per level).
constexpr int kMaxEmitLevels = 24;
struct EmitMultiArgs {
	EmitLevelArgs lv[kMaxEmitLevels];
	u32 blockStart[kMaxEmitLevels + 1];
	int count;
};
int launchEmitInnerLevels(EmitMultiArgs& a, cudaStream_t stream);
// bases[l] for l = top..minLevel from words[l]; total -> *totalWords; sets kOverflowWords in *overflow if the total
// exceeds `capacity`; *rootWord = the root's mask (the DAG's first word). One thread.
// dst[0..words) = src[0..words); both pointers equally aligned within 16 bytes.
int launchCopyWords(u32* dst, const u32* src, u64 words, cudaStream_t stream);
int launchLevelBases(const u64* words, u64* bases, int topLevel, int minLevel, u64* totalWords, u64 capacity, u32* overflow, const u16* rootMask,
		u32* rootWord, cudaStream_t stream);

// ---- lookup.cu: traverse (reference src/CompressedShadow.cpp:404-463, shader/traverse.cs) ----
struct LookupDag {
	const u32* dag;
	const u32* grid;   // NULL: single DAG
	u32 dagLevels;
	u32 gridLevels;
	int leafmasks;
	// Optional shortcut over the top `skipLevels` levels of every DAG: one entry per cell of the
	// (2^(gridLevels+skipLevels))^3 grid over the whole volume, holding what the descent would have reached
	// there -- kSkipShadow / kSkipVisible, or the word offset (relative to the cell's DAG) of the node to
	// continue from. Private to the lookup; the DAG words themselves are untouched.
	const u32* skip;
	u32 skipLevels;
	// Optional (lookup_index.cu): `dag` is the private lookup copy -- eight slots per inner node (0 shadow, 1 lit, 2 + id of the
	// child), `grid` and `skip` hold node ids -- and leafCodes the 32-byte k-code per leaf id (nibble x of word y = lit slices of
	// texel (x, y)).
	const u32* leafCodes = nullptr;
};
constexpr u32 kSkipShadow = 0xFFFFFFFFu, kSkipVisible = 0xFFFFFFFEu;
constexpr u32 kMaxSkipLevels = 6;
// Number of levels the shortcut can cover for a DAG of dagLevels (0: none).
inline u32 skipLevelsFor(u32 dagLevels, int leafmasks, u32 gridLevels) {
	const int inner = (int)dagLevels - 1 - (leafmasks ? 3 : 0);  // inner levels dagLevels-2 .. (3 | 0)
	int g = inner - 1 < (int)kMaxSkipLevels ? inner - 1 : (int)kMaxSkipLevels;
	while (g > 0 && 3 * (g + (int)gridLevels) > 21) --g;  // at most 2^21 entries (8 MiB)
	return g > 0 ? (u32)g : 0u;
}
// Fills d.skip (writable here) for all cells; d.skipLevels must be set.
int launchBuildSkipGrid(const LookupDag& d, u32* skip, cudaStream_t stream);
int launchLookupNdc(const LookupDag& d, const float* ndc, long long count, unsigned char* out, cudaStream_t stream);
// filterSize: percentage-closer filter over filterSize^2 voxels of the pixel's depth slice (1 = one lookup, 0 / 255).
int launchEvaluate(const LookupDag& d, const float* positions, unsigned width, unsigned height, const float* matrix, int filterSize, unsigned char* out,
		cudaStream_t stream);
// The same on CUDA surface objects: rgba32f positions, r8 visibilities (the renderer's G-buffer textures through CUDA-GL interop).
int launchEvaluateSurface(const LookupDag& d, unsigned long long positions, unsigned long long visibilities, unsigned width, unsigned height,
		const float* matrix, int filterSize, cudaStream_t stream);

// ---- synthgen.cu: device-resident synthetic depth tiles (scene definition: ../synth/scene.h)
struct CityBoxDev {
	int x0, y0, x1, y1;  // window texels [x0,x1) x [y0,y1)
	float z;
};
int launchPlaneDepth(float* out, int n, long long gx0, long long gy0, long long gn, cudaStream_t stream);
int launchTerrainDevDepth(float* out, int n, long long gx0, long long gy0, long long gn, cudaStream_t stream);
int launchCityDepth(float* out, int n, const CityBoxDev* boxes, int numBoxes, float farPlane, cudaStream_t stream);

}  // namespace cpvs
