// K4/K5 -- bottom-up merging of identical subtrees, one level at a time
// (reference CompressedShadow::mergeCommonSubtrees / updateParentPointers / removeUnusedNodes,
//  src/CompressedShadow.cpp:215-304, and cs::mergeLevel / isEqualSubtree,
//  src/CompressedShadowUtil.h:137-182).
//
// The reference keeps, for every distinct node tuple of a level, its FIRST occurrence, in order, and
// rewrites the parents' pointers. Here a node tuple is (childmask, unique ids of its PARTIAL
// children) -- equal tuples <=> equal 9-word reference nodes once the level below is merged -- or,
// for leaves, the 32-byte k-code that determines the eight 64-bit slice masks (svo.cu). Per level:
//   0. (leaves)  the table is sized from a distinct-count sketch filled while the leaves were built.
//   1. insert:   every node finds its group's slot in an open-addressing table (linear probing). A
//                slot is (32-bit fingerprint << 32 | smallest node index seen so far); a node joins a
//                slot only after comparing its full tuple against a member of the group, so grouping
//                is exact, never hash-trusting. atomicMin keeps the first occurrence.
//                The slot index is the node's group id: equal nodes share it, which is all the parent
//                level needs to compare tuples, so the next level's insert can start right away.
//   2. rank:     (off the critical path, on a side stream) representative = slot's final index; nodes
//                that are their own representative are the unique nodes. One look-back scan ranks them
//                (first-occurrence order = the reference's layout), prefix-sums their compressed sizes
//                and leaves each group's word offset next to its slot for the parents' pointers.
#include <cstdlib>

#include "kernels.h"

namespace cpvs {

namespace {

constexpr u64 kEmpty = ~0ull;
constexpr u64 kMaxProbes = 4096;
static_assert(kDirectSlots == 256, "one direct slot per thread of the insert CTA");
// gid = table slot (< 2^31) | kCandidateFlag; consumers of the group id strip the flag (kGidMask).
constexpr u32 kCandidateFlag = 0x80000000u, kGidMask = 0x7FFFFFFFu;

// kShared: the table lives in shared memory (single-CTA kernels for the small levels).
template <bool kShared = false, typename Equal>
__device__ __forceinline__ u32 findGroupSlot(u64* __restrict__ table, u64 tableMask, u64 hash, u32 self, u32* errorFlag, Equal sameTuple) {
	const u64 fp = hash >> 32;
	const u64 key = (fp << 32) | self;
	u64 slot = hash & tableMask;
	// (a probe sequence this long means the table was sized for far fewer groups than there are -- a prediction that failed:
	// give up at once, the host rebuilds with a table sized from the data)
	for (u64 probes = 0; probes <= tableMask && probes < kMaxProbes; ++probes) {
		if ((probes & 63u) == 63u && ldRelaxed32(errorFlag)) return 0u;  // somebody already found the table too small
		u64 v = kShared ? *reinterpret_cast<volatile u64*>(table + slot) : ldRelaxed64(table + slot);
		if (v == kEmpty) {
			const u64 old = atomicCAS(reinterpret_cast<unsigned long long*>(table + slot), (unsigned long long)kEmpty, (unsigned long long)key);
			if (old == kEmpty) return (u32)slot | kCandidateFlag;
			v = old;
		}
		if ((v >> 32) == fp) {
			const u32 other = (u32)v;
			if (other == self || sameTuple(other)) {
				// the slot only ever decreases: nothing to do if an earlier node already holds it. A node
				// that never lowered its slot cannot be the first occurrence; the ones that did are marked
				// as candidates, which spares the rank scan the table look-up for everybody else.
				u32 lowered = 0;
				if (other > self)
					lowered = atomicMin(reinterpret_cast<unsigned long long*>(table + slot), (unsigned long long)key) > key ? kCandidateFlag : 0u;
				return (u32)slot | lowered;
			}
		}
		slot = (slot + 1) & tableMask;
	}
	atomicExch(errorFlag, 1u);  // table full: cannot happen with a sane size estimate; reported to the host
	return 0u;
}

// Leaf table capacity from the distinct-count sketch (linear counting: u ~ -m ln(zero fraction)), rounded
// up to a power of two with >= 1.5x headroom, then cleared. The result does not depend on the capacity,
// only the speed does: a table sized for the distinct leaves (not for all leaves) stays in L2.
__global__ void __launch_bounds__(256) sizeAndClearLeafTableKernel(u64* __restrict__ table, u64 maxSlots, const u64* __restrict__ setBits,
		u64* __restrict__ tableMaskDev, float expectedDistinct) {
	const float m = (float)kSketchWords * 32.0f;
	const float frac = fminf((float)*setBits / m, 0.999f);
	const float distinct = expectedDistinct > 0.f ? expectedDistinct : -m * log1pf(-frac);
	u64 want = (u64)(distinct * 1.5f) + 4096u;
	u64 cap = 4096;
	while (cap < want && cap < maxSlots) cap <<= 1;
	if (cap > maxSlots) cap = maxSlots;
	if (blockIdx.x == 0 && threadIdx.x == 0) *tableMaskDev = cap - 1;
	ulonglong2* t2 = reinterpret_cast<ulonglong2*>(table);
	const u64 pairs = cap / 2;
	for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += (u64)gridDim.x * blockDim.x) t2[i] = make_ulonglong2(kEmpty, kEmpty);
}

// Leaf insert, one leaf per thread. The chain per leaf is own code -> hash -> table probe -> witness code -> compare -> atomicMin:
// three to four dependent memory round trips; the kernel runs at 32 registers and full occupancy, with the loads that do not
// depend on the device-side size issued ahead of it. A staged variant (codes of 1024 leaves in shared memory through cp.async,
// several leaves per thread, every step of the chain issued for all of them before any result is looked at) was measured and
// removed: a warp then runs as many rounds as the longest probe chain among its leaves (profiles/r2_dropped.md).

__device__ __forceinline__ u64 hashLeafCode(const uint4& a0, const uint4& a1) {
	u64 h = 0x9E3779B97F4A7C15ull;
	h = (h ^ (((u64)a0.y << 32) | a0.x)) * 0xFF51AFD7ED558CCDull;
	h = (h ^ (h >> 32) ^ (((u64)a0.w << 32) | a0.z)) * 0xC4CEB9FE1A85EC53ull;
	h = (h ^ (h >> 32) ^ (((u64)a1.y << 32) | a1.x)) * 0xFF51AFD7ED558CCDull;
	h = (h ^ (h >> 32) ^ (((u64)a1.w << 32) | a1.z)) * 0xC4CEB9FE1A85EC53ull;
	return mix64(h);
}

__global__ void __launch_bounds__(256, 8) insertLeavesKernel(const u32* __restrict__ codes, const u64* __restrict__ hashes, const u64* __restrict__ nDev,
		u64 cap, u64* __restrict__ table, const u64* __restrict__ tableMaskDev, u32* __restrict__ slotOf, u32* errorFlag,
		const u32* __restrict__ overflow) {
	const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= cap) return;
	// own code (and hash) are fetched up front, together with the level's size: the arrays hold `cap` leaves, so the loads need not
	// wait for the size that only the device knows -- one round trip less in front of a chain of three
	const uint4* mine = reinterpret_cast<const uint4*>(codes + j * 8);
	const uint4 a0 = __ldcs(mine), a1 = __ldcs(mine + 1);
	u64 hash = hashes ? __ldcs(hashes + j) : 0;
	const u64 tableMask = *tableMaskDev;
	if (j >= *nDev || (*overflow & kOverflowNodes)) return;
	if (!hashes) hash = hashLeafCode(a0, a1);
	slotOf[j] = findGroupSlot(table, tableMask, hash, (u32)j, errorFlag, [&](u32 other) {
		const uint4* theirs = reinterpret_cast<const uint4*>(codes + (u64)other * 8);
		const uint4 b0 = theirs[0], b1 = theirs[1];
		return a0.x == b0.x && a0.y == b0.y && a0.z == b0.z && a0.w == b0.w && a1.x == b1.x && a1.y == b1.y && a1.z == b1.z && a1.w == b1.w;
	});
}

template <bool kShared = false>
__device__ __forceinline__ u32 insertInnerNode(u32 j, u32 mask, u32 first, const u16* __restrict__ masks, const u32* __restrict__ firstChild,
		const u32* __restrict__ childUid, u64* __restrict__ table, u64 tableMask, u32* errorFlag) {
	const u32 k = __popc(mask & 0xAAAAu);
	const u32* kids = childUid + first;
	u32 uid[8];
	u64 h = mix64(0x51ED270B6F2D4A6Bull ^ mask);
#pragma unroll
	for (u32 c = 0; c < 8; ++c) {
		uid[c] = 0;
		if (c < k) {
			uid[c] = kids[c] & kGidMask;
			h = mix64(h ^ ((u64)uid[c] + 0x9E3779B97F4A7C15ull * (c + 1)));
		}
	}
	return findGroupSlot<kShared>(table, tableMask, h, j, errorFlag, [&](u32 other) {
		if (masks[other] != mask) return false;
		const u32* theirs = childUid + firstChild[other];
		bool same = true;
#pragma unroll
		for (u32 c = 0; c < 8; ++c)
			if (c < k) same = same && (theirs[c] & kGidMask) == uid[c];
		return same;
	});
}

// Nodes without PARTIAL children are their mask: eight 1-bit child codes, 256 possible tuples. On
// repetitive maps (and on the bottom level of a leafmask-less octree, where no child is ever PARTIAL)
// millions of nodes share a handful of them, and probing would hammer a few table slots. They bypass
// the hash: each tuple owns one of kDirectSlots slots behind the hashed region, the CTA reduces its
// first occurrences in shared memory and touches each slot at most once.
__device__ __forceinline__ u32 compactLitBits(u32 mask) {
	u32 x = mask & 0x5555u;
	x = (x | (x >> 1)) & 0x3333u;
	x = (x | (x >> 2)) & 0x0F0Fu;
	return (x | (x >> 4)) & 0x00FFu;
}

// Held to 32 registers (8 instead of 6 CTAs of 256 threads per SM, one 4-byte spill): the kernel waits on dependent loads 62 % of
// the time, and the two extra CTAs are worth 0.03 ms on the 16K^2 terrain (inner merge 0.436 -> 0.409 ms, the leaf rank running
// beside it 0.375 -> 0.348 ms; profiles/r1_switch_probe.md).
__global__ void __launch_bounds__(256, 8) insertInnerKernel(const u16* __restrict__ masks, const u32* __restrict__ firstChild,
		const u32* __restrict__ childUid, const u64* __restrict__ nDev, u64 cap, u64* __restrict__ table, u64 tableMask, u32* __restrict__ slotOf,
		u32* errorFlag, const u32* __restrict__ overflow) {
	__shared__ u32 sFirst[kDirectSlots];
	const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	// (the mask is fetched together with the level's size: the arrays hold `cap` nodes)
	const u32 maskSpec = j < cap ? masks[j] : 0xAAAAu;
	const u32 firstSpec = j < cap ? firstChild[j] : 0u;
	const u64 n = *nDev;
	// the grid is sized for the level's capacity; a level cut short by a capacity (the child lists of its last nodes are
	// incomplete) is not merged at all -- the host rebuilds with exact counts
	if ((u64)blockIdx.x * blockDim.x >= n || (*overflow & kOverflowNodes)) return;
	sFirst[threadIdx.x] = 0xFFFFFFFFu;
	__syncthreads();
	const bool live = j < n;
	const u32 mask = live ? maskSpec : 0xAAAAu;
	const bool direct = live && (mask & 0xAAAAu) == 0;
	const u32 c = compactLitBits(mask);
	if (direct) atomicMin(&sFirst[c], (u32)j);
	// the direct nodes are settled before anybody starts probing: the barrier only ever waits for the mask loads, not for
	// the slowest probe sequence of the block (17 % of this kernel's stall samples when the probes came first)
	__syncthreads();
	const u32 mine = sFirst[threadIdx.x];
	if (mine != 0xFFFFFFFFu) {
		u64* slot = table + tableMask + 1 + threadIdx.x;  // key = (fingerprint 0, index): plain minimum, kEmpty is the maximum
		if ((u32)ldRelaxed64(slot) > mine) atomicMin(reinterpret_cast<unsigned long long*>(slot), (unsigned long long)mine);
	}
	if (direct)
		slotOf[j] = (u32)(tableMask + 1 + c) | (sFirst[c] == (u32)j ? kCandidateFlag : 0u);
	else if (live)
		slotOf[j] = insertInnerNode<false>((u32)j, mask, firstSpec, masks, firstChild, childUid, table, tableMask, errorFlag);
}

// gid[j] = slot of node j's group. Ranks the first occurrences (slot's final index == j) in order,
// prefix-sums their compressed sizes, and records firstList[rank], wordOffset[rank] and, per group,
// slotOffset[slot] = word offset of the group's node inside the level.
//
// Three plain kernels instead of one look-back scan: (1) per 1024-node tile, find the first occurrences
// (the only part with random accesses) and leave each node's compressed size in a byte plus the tile's
// totals; (2) one CTA prefix-sums the tile totals; (3) per tile, a block scan of the bytes and the writes.
// No CTA ever waits for another one, which matters more here than the extra byte per node of traffic.
// The four masks and group ids of a thread are fetched with one 8-byte / 16-byte load up front instead of one dependent load per
// first occurrence behind the table look-up (measured: 0.015 ms of the 16K^2 terrain build, profiles/r2_switch_probe.md).
__global__ void __launch_bounds__(kScanThreads, 8) rankCountKernel(const u64* __restrict__ table, const u16* __restrict__ masks, int leaf,
		const u64* __restrict__ nDev, const u32* __restrict__ gid, unsigned char* __restrict__ sizeOf, ScanTileState* __restrict__ tiles,
		const u32* __restrict__ overflow) {
	const u64 base = (u64)blockIdx.x * kScanTile + (u64)threadIdx.x * kScanItems;
	// (the arrays hold at least the grid's worth of nodes, rounded up to whole vectors: these loads do not wait for the size)
	const uint2 v = *reinterpret_cast<const uint2*>(masks + base);
	const uint4 w = *reinterpret_cast<const uint4*>(gid + base);
	const u64 n = (*overflow & kOverflowNodes) ? 0 : *nDev;
	if ((u64)blockIdx.x * kScanTile >= n) return;  // the grid is sized for the level's capacity
	u32 myMask[kScanItems] = {0, 0, 0, 0}, g[kScanItems] = {0, 0, 0, 0};
	if (base + kScanItems <= n) {
		myMask[0] = v.x & 0xFFFFu;
		myMask[1] = v.x >> 16;
		myMask[2] = v.y & 0xFFFFu;
		myMask[3] = v.y >> 16;
		g[0] = w.x;
		g[1] = w.y;
		g[2] = w.z;
		g[3] = w.w;
	} else {
#pragma unroll
		for (int i = 0; i < kScanItems; ++i)
			if (base + i < n) {
				myMask[i] = masks[base + i];
				g[i] = gid[base + i];
			}
	}
	u32 words[kScanItems];
	u64 cnt = 0, wsum = 0;
#pragma unroll
	for (int i = 0; i < kScanItems; ++i) {
		words[i] = 0;
		if (base + i < n && (g[i] & kCandidateFlag) && (u32)table[(u64)(g[i] & kGidMask)] == (u32)(base + i)) {
			const u32 k = __popc(myMask[i] & 0xAAAAu);
			words[i] = 1 + (leaf ? 2 * k : k);
			cnt += 1;
			wsum += words[i];
		}
	}
	if (base + kScanItems <= n) {
		static_assert(kScanItems == 4, "packed size store");
		*reinterpret_cast<u32*>(sizeOf + base) = words[0] | (words[1] << 8) | (words[2] << 16) | (words[3] << 24);
	} else {
#pragma unroll
		for (int i = 0; i < kScanItems; ++i)
			if (base + i < n) sizeOf[base + i] = (unsigned char)words[i];
	}
	u64 preC = cnt, preW = wsum, totC, totW;
	blockExclusiveScan2(preC, preW, totC, totW);
	if (threadIdx.x == 0) {
		tiles[blockIdx.x].a = totC;
		tiles[blockIdx.x].b = totW;
	}
}

// Exclusive prefix sums of the tile totals, in place; one CTA. Every thread owns a contiguous chunk of tiles and fetches it
// whole before anything is added up, so the kernel pays one memory round trip however many tiles there are (a loop of
// 1024-tile rounds paid one per round: 54 us for the 13.6 K tiles of the 16K^2 terrain's leaf level).
constexpr int kScanChunk = 16, kScanTilesThreads = 512;  // 8 K tiles per pass
__global__ void __launch_bounds__(kScanTilesThreads) rankScanTilesKernel(ScanTileState* __restrict__ tiles, const u64* __restrict__ nDev,
		u64* __restrict__ uniqueCount, u64* __restrict__ wordCount, const u32* __restrict__ overflow) {
	const u32 numTiles = (*overflow & kOverflowNodes) ? 0u : (u32)((*nDev + kScanTile - 1) / kScanTile);
	__shared__ u64 sA[32], sB[32];
	const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	u64 carryA = 0, carryB = 0;
	for (u32 base = 0; base < numTiles; base += (u32)kScanTilesThreads * kScanChunk) {
		const u32 first = base + threadIdx.x * kScanChunk;
		u64 a[kScanChunk], b[kScanChunk];
#pragma unroll
		for (int i = 0; i < kScanChunk; ++i) {
			const bool live = first + i < numTiles;
			a[i] = live ? tiles[first + i].a : 0;
			b[i] = live ? tiles[first + i].b : 0;
		}
		u64 sumA = 0, sumB = 0;
#pragma unroll
		for (int i = 0; i < kScanChunk; ++i) {
			sumA += a[i];
			sumB += b[i];
		}
		u64 ia = sumA, ib = sumB;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const u64 ua = __shfl_up_sync(0xFFFFFFFFu, ia, d), ub = __shfl_up_sync(0xFFFFFFFFu, ib, d);
			if ((int)lane >= d) {
				ia += ua;
				ib += ub;
			}
		}
		if (lane == 31) {
			sA[warp] = ia;
			sB[warp] = ib;
		}
		__syncthreads();
		u64 offA = 0, offB = 0, totA = 0, totB = 0;
#pragma unroll
		for (u32 w = 0; w < kScanTilesThreads / 32; ++w) {
			if (w < warp) {
				offA += sA[w];
				offB += sB[w];
			}
			totA += sA[w];
			totB += sB[w];
		}
		__syncthreads();
		u64 runA = carryA + offA + ia - sumA, runB = carryB + offB + ib - sumB;
#pragma unroll
		for (int i = 0; i < kScanChunk; ++i) {
			if (first + i < numTiles) {
				tiles[first + i].a = runA;
				tiles[first + i].b = runB;
			}
			runA += a[i];
			runB += b[i];
		}
		carryA += totA;
		carryB += totB;
	}
	if (threadIdx.x == 0) {
		*uniqueCount = carryA;
		*wordCount = carryB;
	}
}

__global__ void __launch_bounds__(kScanThreads, 8) rankWriteKernel(const unsigned char* __restrict__ sizeOf, const u32* __restrict__ gid,
		const u64* __restrict__ nDev, const ScanTileState* __restrict__ tiles, u32* __restrict__ firstList, u32* __restrict__ wordOffset,
		u32* __restrict__ slotOffset, const u32* __restrict__ overflow) {
	const u64 n = (*overflow & kOverflowNodes) ? 0 : *nDev;
	const u64 base = (u64)blockIdx.x * kScanTile + (u64)threadIdx.x * kScanItems;
	if ((u64)blockIdx.x * kScanTile >= n) return;  // the grid is sized for the level's capacity
	u32 g[kScanItems] = {0, 0, 0, 0};
	u32 words[kScanItems] = {0, 0, 0, 0};
	if (base + kScanItems <= n) {
		const uint4 v = *reinterpret_cast<const uint4*>(gid + base);
		g[0] = v.x;
		g[1] = v.y;
		g[2] = v.z;
		g[3] = v.w;
		const u32 packed = *reinterpret_cast<const u32*>(sizeOf + base);
		words[0] = packed & 0xFFu;
		words[1] = (packed >> 8) & 0xFFu;
		words[2] = (packed >> 16) & 0xFFu;
		words[3] = packed >> 24;
	} else {
#pragma unroll
		for (int i = 0; i < kScanItems; ++i)
			if (base + i < n) {
				g[i] = gid[base + i];
				words[i] = sizeOf[base + i];
			}
	}
	u64 cnt = 0, wsum = 0;
#pragma unroll
	for (int i = 0; i < kScanItems; ++i) {
		cnt += words[i] ? 1 : 0;
		wsum += words[i];
	}
	u64 preC = cnt, preW = wsum, totC, totW;
	blockExclusiveScan2(preC, preW, totC, totW);
	u64 rank = tiles[blockIdx.x].a + preC, woff = tiles[blockIdx.x].b + preW;
#pragma unroll
	for (int i = 0; i < kScanItems; ++i) {
		if (words[i]) {
			firstList[rank] = (u32)(base + i);
			wordOffset[rank] = (u32)woff;
			slotOffset[g[i] & kGidMask] = (u32)woff;
			++rank;
			woff += words[i];
		}
	}
}

// The small top levels, bottom-up, by one CTA: clear, insert, rank with block barriers in between.
// Same tuples, same first-occurrence rule, same outputs as the per-level kernels.
__global__ void __launch_bounds__(kSmallThreads) mergeSmallLevelsKernel(SmallMergeArgs a) {
	__shared__ u32 sWarpC[kSmallThreads / 32], sWarpW[kSmallThreads / 32];
	extern __shared__ __align__(16) u64 sTable[];  // 2 * kSmallMaxNodes slots: probes and atomics stay on chip
	const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	for (int s = 0; s < a.count; ++s) {
		SmallMergeLevel L = a.lv[s];
		const u32 n = (*a.overflow & kOverflowNodes) ? 0u : (u32)*L.nDev;
		if (n == 0) {
			if (threadIdx.x == 0) {
				*L.uniqueCount = 0;
				*L.wordCount = 0;
			}
			continue;
		}
		if (n == 1) {
			if (threadIdx.x == 0) {
				L.uid[0] = 0;
				L.slotOffset[0] = 0;
				L.firstList[0] = 0;
				L.wordOffset[0] = 0;
				*L.uniqueCount = 1;
				*L.wordCount = 1 + __popc(L.masks[0] & 0xAAAAu);
			}
			__syncthreads();
			continue;
		}
		// table sized to the level (power of two >= 2n), cleared and probed in shared memory
		u32 slots = 64;
		while (slots < 2 * n) slots <<= 1;
		for (u32 i = threadIdx.x; i < slots; i += kSmallThreads) sTable[i] = kEmpty;
		__syncthreads();
		for (u32 j = threadIdx.x; j < n; j += kSmallThreads)
			L.uid[j] = insertInnerNode<true>(j, L.masks[j], L.firstChild[j], L.masks, L.firstChild, L.childUid, sTable, slots - 1, a.errorFlag);
		__syncthreads();
		u32 carryC = 0, carryW = 0;
		for (u32 base = 0; base < n; base += kSmallThreads) {
			const u32 j = base + threadIdx.x;
			u32 slot = 0, words = 0;
			if (j < n) {
				slot = L.uid[j] & kGidMask;
				if ((u32)sTable[slot] == j) words = 1 + __popc(L.masks[j] & 0xAAAAu);
			}
			u32 inclC = words ? 1u : 0u, inclW = words;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const u32 uc = __shfl_up_sync(0xFFFFFFFFu, inclC, d), uw = __shfl_up_sync(0xFFFFFFFFu, inclW, d);
				if ((int)lane >= d) {
					inclC += uc;
					inclW += uw;
				}
			}
			if (lane == 31) {
				sWarpC[warp] = inclC;
				sWarpW[warp] = inclW;
			}
			__syncthreads();
			u32 beforeC = 0, beforeW = 0, totC = 0, totW = 0;
#pragma unroll
			for (u32 w = 0; w < kSmallThreads / 32; ++w) {
				const u32 c = sWarpC[w], v = sWarpW[w];
				if (w < warp) {
					beforeC += c;
					beforeW += v;
				}
				totC += c;
				totW += v;
			}
			__syncthreads();
			if (j < n && words) {
				const u32 rank = carryC + beforeC + inclC - 1, woff = carryW + beforeW + inclW - words;
				L.firstList[rank] = j;
				L.wordOffset[rank] = woff;
				L.slotOffset[slot] = woff;
			}
			carryC += totC;
			carryW += totW;
		}
		if (threadIdx.x == 0) {
			*L.uniqueCount = carryC;
			*L.wordCount = carryW;
		}
		__syncthreads();
	}
}

}  // namespace

int launchMergeSmallLevels(const SmallMergeArgs& a, cudaStream_t stream) {
	if (a.count <= 0) return 0;
	constexpr size_t kBytes = 2 * kSmallMaxNodes * sizeof(u64);
	// > 48 KiB of dynamic shared memory needs the opt-in, once per device (a process may hold one context per GPU)
	static bool configured[64] = {false};
	int device = 0;
	cudaGetDevice(&device);
	if (device < 0 || device >= 64 || !configured[device]) {
		cudaFuncSetAttribute(mergeSmallLevelsKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBytes);
		if (device >= 0 && device < 64) configured[device] = true;
	}
	mergeSmallLevelsKernel<<<1, kSmallThreads, kBytes, stream>>>(a);
	return 1;
}

int launchSizeLeafTable(u64* table, u64 maxSlots, const u64* setBits, u64* tableMaskDev, u64 expectedDistinct, cudaStream_t stream) {
	sizeAndClearLeafTableKernel<<<148 * 8, 256, 0, stream>>>(table, maxSlots, setBits, tableMaskDev, (float)expectedDistinct);
	return 1;
}

int launchInsertLevel(const MergeLevelArgs& a, cudaStream_t stream) {
	if (!a.cap) return 0;
	const unsigned blocks = (unsigned)((a.cap + 255) / 256);
	if (a.leaf)
		insertLeavesKernel<<<blocks, 256, 0, stream>>>(a.leafCodes, a.leafHash, a.nDev, a.cap, a.table, a.tableMaskDev, a.uid, a.errorFlag, a.overflow);
	else
		insertInnerKernel<<<blocks, 256, 0, stream>>>(a.masks, a.firstChild, a.childUid, a.nDev, a.cap, a.table, a.tableSize - 1, a.uid, a.errorFlag, a.overflow);
	return 1;
}

int launchRankLevel(const MergeLevelArgs& a, ScanLaunch scan, cudaStream_t stream) {
	if (!a.cap) return 0;
	const u32 tiles = (u32)((a.cap + kScanTile - 1) / kScanTile);
	rankCountKernel<<<tiles, kScanThreads, 0, stream>>>(a.table, a.masks, a.leaf, a.nDev, a.uid, a.sizeOf, scan.tiles, a.overflow);
	rankScanTilesKernel<<<1, kScanTilesThreads, 0, stream>>>(scan.tiles, a.nDev, a.uniqueCount, a.wordCount, a.overflow);
	rankWriteKernel<<<tiles, kScanThreads, 0, stream>>>(a.sizeOf, a.uid, a.nDev, scan.tiles, a.firstList, a.wordOffset, a.slotOffset, a.overflow);
	return 3;
}

}  // namespace cpvs
