// K4/K5 -- bottom-up merging of identical subtrees, one level at a time
// (reference CompressedShadow::mergeCommonSubtrees / updateParentPointers / removeUnusedNodes,
//  src/CompressedShadow.cpp:215-304, and cs::mergeLevel / isEqualSubtree,
//  src/CompressedShadowUtil.h:137-182).
//
// The reference keeps, for every distinct node tuple of a level, its FIRST occurrence, in order, and
// rewrites the parents' pointers. Here a node tuple is (childmask, unique ids of its PARTIAL
// children) -- equal tuples <=> equal 9-word reference nodes once the level below is merged -- or,
// for leaves, the eight 64-bit slice masks. Per level:
//   1. insert:   every node finds its group's slot in an open-addressing table (linear probing). A
//                slot is (32-bit fingerprint << 32 | smallest node index seen so far); a node joins a
//                slot only after comparing its full tuple against a member of the group, so grouping
//                is exact, never hash-trusting. atomicMin keeps the first occurrence.
//   2. resolve:  representative = slot's final index; nodes that are their own representative are
//                the unique nodes. One look-back scan ranks them (first-occurrence order = the
//                reference's layout) and prefix-sums their compressed sizes.
//   3. finalize: every node's unique id = rank of its representative (the parents' remap table).
#include "kernels.h"

namespace cpvs {

namespace {

constexpr u64 kEmpty = ~0ull;
constexpr u32 kFirstFlag = 0x80000000u;

template <typename Equal>
__device__ __forceinline__ u32 findGroupSlot(u64* __restrict__ table, u64 tableMask, u64 hash, u32 self, Equal sameTuple) {
	const u64 fp = hash >> 32;
	const u64 key = (fp << 32) | self;
	u64 slot = hash & tableMask;
	for (;;) {
		u64 v = ldRelaxed64(table + slot);
		if (v == kEmpty) {
			const u64 old = atomicCAS(reinterpret_cast<unsigned long long*>(table + slot), (unsigned long long)kEmpty, (unsigned long long)key);
			if (old == kEmpty) return (u32)slot;
			v = old;
		}
		if ((v >> 32) == fp) {
			const u32 other = (u32)v;
			if (other == self || sameTuple(other)) {
				atomicMin(reinterpret_cast<unsigned long long*>(table + slot), (unsigned long long)key);
				return (u32)slot;
			}
		}
		slot = (slot + 1) & tableMask;
	}
}

__global__ void __launch_bounds__(256) insertLeavesKernel(const u64* __restrict__ bits, const u64* __restrict__ hashes, u64 n, u64* __restrict__ table,
		u64 tableMask, u32* __restrict__ slotOf) {
	const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n) return;
	const ulonglong2* mine = reinterpret_cast<const ulonglong2*>(bits + j * 8);
	slotOf[j] = findGroupSlot(table, tableMask, hashes[j], (u32)j, [&](u32 other) {
		const ulonglong2* theirs = reinterpret_cast<const ulonglong2*>(bits + (u64)other * 8);
		bool same = true;
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const ulonglong2 a = mine[i], b = theirs[i];
			same = same && a.x == b.x && a.y == b.y;
		}
		return same;
	});
}

__global__ void __launch_bounds__(256) insertInnerKernel(const u16* __restrict__ masks, const u32* __restrict__ firstChild,
		const u32* __restrict__ childUid, u64 n, u64* __restrict__ table, u64 tableMask, u32* __restrict__ slotOf) {
	const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n) return;
	const u32 mask = masks[j];
	const u32 k = __popc(mask & 0xAAAAu);
	const u32* kids = childUid + firstChild[j];
	u32 uid[8];
	u64 h = mix64(0x51ED270B6F2D4A6Bull ^ mask);
#pragma unroll
	for (u32 c = 0; c < 8; ++c) {
		uid[c] = 0;
		if (c < k) {
			uid[c] = kids[c];
			h = mix64(h ^ ((u64)uid[c] + 0x9E3779B97F4A7C15ull * (c + 1)));
		}
	}
	slotOf[j] = findGroupSlot(table, tableMask, h, (u32)j, [&](u32 other) {
		if (masks[other] != mask) return false;
		const u32* theirs = childUid + firstChild[other];
		bool same = true;
#pragma unroll
		for (u32 c = 0; c < 8; ++c)
			if (c < k) same = same && theirs[c] == uid[c];
		return same;
	});
}

// slotOf[j] (in) -> uid[j] (out): rank | kFirstFlag for first occurrences, representative index otherwise.
__global__ void __launch_bounds__(kScanThreads) resolveKernel(const u64* __restrict__ table, const u16* __restrict__ masks, int leaf, u64 n,
		u32* __restrict__ uid, u32* __restrict__ firstList, u32* __restrict__ wordOffset, u64* __restrict__ uniqueCount,
		u64* __restrict__ wordCount, ScanLaunch scan, u32 numTiles) {
	const u32 tile = scanAcquireTile(scan);
	const u64 base = (u64)tile * kScanTile + (u64)threadIdx.x * kScanItems;
	u32 rep[kScanItems], words[kScanItems];
	u64 cnt = 0, wsum = 0;
#pragma unroll
	for (int i = 0; i < kScanItems; ++i) {
		rep[i] = 0xFFFFFFFFu;
		words[i] = 0;
		if (base + i < n) {
			rep[i] = (u32)table[uid[base + i]];
			if (rep[i] == (u32)(base + i)) {
				const u32 k = __popc(masks[base + i] & 0xAAAAu);
				words[i] = 1 + (leaf ? 2 * k : k);
				cnt += 1;
				wsum += words[i];
			}
		}
	}
	u64 preC = cnt, preW = wsum, totC, totW;
	blockExclusiveScan2(preC, preW, totC, totW);
	u64 tileC, tileW;
	scanLookback2(scan, tile, totC, totW, tileC, tileW);
	if (tile == numTiles - 1 && threadIdx.x == 0) {
		*uniqueCount = tileC + totC;
		*wordCount = tileW + totW;
	}
	u64 rank = tileC + preC, woff = tileW + preW;
#pragma unroll
	for (int i = 0; i < kScanItems; ++i) {
		if (base + i >= n) break;
		if (words[i]) {
			firstList[rank] = (u32)(base + i);
			wordOffset[rank] = (u32)woff;
			uid[base + i] = (u32)rank | kFirstFlag;
			++rank;
			woff += words[i];
		} else {
			uid[base + i] = rep[i];
		}
	}
}

// Safe in place: a first occurrence only ever loses its flag bit, which readers mask off.
__global__ void finalizeUidKernel(u32* __restrict__ uid, u64 n) {
	const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n) return;
	const u32 v = uid[j];
	uid[j] = (v & kFirstFlag) ? (v & ~kFirstFlag) : (reinterpret_cast<volatile u32*>(uid)[v] & ~kFirstFlag);
}

// A level with a single node (the root, which the reference never merges).
__global__ void singleNodeKernel(const u16* __restrict__ masks, int leaf, u32* uid, u32* firstList, u32* wordOffset, u64* uniqueCount,
		u64* wordCount) {
	const u32 k = __popc(masks[0] & 0xAAAAu);
	uid[0] = 0;
	firstList[0] = 0;
	wordOffset[0] = 0;
	*uniqueCount = 1;
	*wordCount = 1 + (leaf ? 2 * k : k);
}

}  // namespace

int launchMergeLevel(const MergeLevelArgs& a, ScanLaunch scan, cudaEvent_t afterInsert, cudaStream_t stream) {
	if (a.n == 1) {
		singleNodeKernel<<<1, 1, 0, stream>>>(a.masks, a.leaf, a.uid, a.firstList, a.wordOffset, a.uniqueCount, a.wordCount);
		if (afterInsert) cudaEventRecord(afterInsert, stream);
		return 1;
	}
	const unsigned blocks = (unsigned)((a.n + 255) / 256);
	if (a.leaf)
		insertLeavesKernel<<<blocks, 256, 0, stream>>>(a.leafBits, a.leafHash, a.n, a.table, a.tableSize - 1, a.uid);
	else
		insertInnerKernel<<<blocks, 256, 0, stream>>>(a.masks, a.firstChild, a.childUid, a.n, a.table, a.tableSize - 1, a.uid);
	if (afterInsert) cudaEventRecord(afterInsert, stream);
	const u32 tiles = (u32)((a.n + kScanTile - 1) / kScanTile);
	resolveKernel<<<tiles, kScanThreads, 0, stream>>>(a.table, a.masks, a.leaf, a.n, a.uid, a.firstList, a.wordOffset, a.uniqueCount,
			a.wordCount, scan, tiles);
	finalizeUidKernel<<<blocks, 256, 0, stream>>>(a.uid, a.n);
	return 3;
}

}  // namespace cpvs
