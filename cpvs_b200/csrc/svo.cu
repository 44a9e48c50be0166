// K2/K3 -- sparse voxel octree construction from the depth pyramid
// (reference CompressedShadow::constructSvo / constructLastLevels, src/CompressedShadow.cpp:87-190,
//  cs::createChildmask / createLeafmask / createChildmask1x1x8, src/CompressedShadowUtil.cpp:20-99).
//
// The reference walks the octree breadth first on one thread and materialises 9-word nodes. Here a
// level is a structure of arrays -- packed node coordinates, 16-bit child masks, index of the first
// child in the next level -- and one kernel per level classifies all nodes, prefix-sums the PARTIAL
// child counts across the whole level in the same pass (decoupled look-back) and writes the next
// level's coordinate list in the reference's order: parent order, then child index x | y<<1 | z<<2.
// Leaves (level 2) are built by eight lanes per node straight from the depth map.
#include <cuda_fp16.h>

#include "kernels.h"

namespace cpvs {

namespace {

// ---- cs::createChildmask (src/CompressedShadowUtil.cpp:20-54) ---------------------------------------
// level >= 1: tex points at (min,max) pairs; level 0: at the depth map (absoluteVisible, never PARTIAL).
__device__ __forceinline__ u32 childmaskInner(const float2* __restrict__ tex, u32 side, float heightF, u32 ox, u32 oy, u32 oz) {
	const float4 r0 = *reinterpret_cast<const float4*>(tex + (size_t)oy * side + ox);        // texels (ox,oy),(ox+1,oy)
	const float4 r1 = *reinterpret_cast<const float4*>(tex + (size_t)(oy + 1) * side + ox);  // (ox,oy+1),(ox+1,oy+1)
	const float mn[4] = {__fmul_rn(r0.x, heightF), __fmul_rn(r0.z, heightF), __fmul_rn(r1.x, heightF), __fmul_rn(r1.z, heightF)};
	const float mx[4] = {__fmul_rn(r0.y, heightF), __fmul_rn(r0.w, heightF), __fmul_rn(r1.y, heightF), __fmul_rn(r1.w, heightF)};
	u32 mask = 0;
#pragma unroll
	for (u32 z = 0; z < 2; ++z) {
		const float z0 = __uint2float_rn(oz + z), z1 = __uint2float_rn(oz + z + 1);
#pragma unroll
		for (u32 xy = 0; xy < 4; ++xy) mask |= classifyRange(z0, z1, mn[xy], mx[xy]) << ((xy | (z << 2)) * 2);
	}
	return mask;
}
__device__ __forceinline__ u32 childmaskLevel0(const float* __restrict__ depth, u32 side, float heightF, u32 ox, u32 oy, u32 oz) {
	const float2 r0 = *reinterpret_cast<const float2*>(depth + (size_t)oy * side + ox);
	const float2 r1 = *reinterpret_cast<const float2*>(depth + (size_t)(oy + 1) * side + ox);
	const float d[4] = {__fmul_rn(r0.x, heightF), __fmul_rn(r0.y, heightF), __fmul_rn(r1.x, heightF), __fmul_rn(r1.y, heightF)};
	u32 mask = 0;
#pragma unroll
	for (u32 z = 0; z < 2; ++z) {
		const float z0 = __uint2float_rn(oz + z), z1 = __uint2float_rn(oz + z + 1);
#pragma unroll
		for (u32 xy = 0; xy < 4; ++xy) mask |= classifyPoint(z0, z1, d[xy]) << ((xy | (z << 2)) * 2);
	}
	return mask;
}

// ---- node counts per level, for exact allocation (closed form of the classification) ------------
// A level-l node exists for every voxel (x,y,z) of pyramid level l+1 that classifies PARTIAL:
// !(z+1 <= min*H) && !(z >= max*H)  <=>  floor(min*H) <= z <= ceil(max*H)-1. Slice s of the column owns the voxels
// s*side .. s*side+side-1 (side = that level's width), so one pass over the pyramid counts every slice at once.
struct CountLevels {
	const float2* texels[kMaxLevels];  // pyramid level l+1 for node level l
	u64 numTexels[kMaxLevels];
	float heightF[kMaxLevels];  // side * zTileNum
	u32 sideShift[kMaxLevels];  // log2(side)
	int minLevel;
	int topCounted;  // level of the root's children
	u32 zTileNum;
	u32 blockStart[kMaxLevels + 1];  // [minLevel + i] = first block of level minLevel + i; blocks per level follow its size
};
constexpr u32 kCountSharedSlices = 2048;
// All levels are counted by one launch; a level owns the blocks blockStart[level] .. blockStart[level + 1] - 1 (a grid of
// levels x the widest level's blocks spent more time dispatching the empty blocks of the small levels than counting).
__global__ void __launch_bounds__(256) countNodesKernel(CountLevels p, u64* __restrict__ counts) {
	__shared__ u32 sSlice[kCountSharedSlices];
	__shared__ u64 sWarp[8];
	// CTA 0 also reports every slice's root mask, which is the whole DAG of a slice that misses the surface.
	if (blockIdx.x == 0) {
		const float2* __restrict__ under = p.texels[p.topCounted];
		for (u32 z = threadIdx.x; z < p.zTileNum; z += blockDim.x)
			counts[(u64)z * kMaxLevels + kRootMaskScalar] = (1ull << 32) | childmaskInner(under, 2u, p.heightF[p.topCounted], 0u, 0u, z * 2u);
	}
	int level = p.minLevel;
	while (level < p.topCounted && blockIdx.x >= p.blockStart[level + 1]) ++level;
	const u32 block = blockIdx.x - p.blockStart[level], blocks = p.blockStart[level + 1] - p.blockStart[level];
	const float2* __restrict__ texels = p.texels[level];
	const u64 numTexels = p.numTexels[level];
	const float heightF = p.heightF[level], topF = __fadd_rn(heightF, -1.0f);
	const u32 shift = p.sideShift[level], side = 1u << shift;
	const bool single = p.zTileNum == 1, shared = p.zTileNum <= kCountSharedSlices;
	if (!single && shared) {
		for (u32 i = threadIdx.x; i < p.zTileNum; i += blockDim.x) sSlice[i] = 0;
		__syncthreads();
	}
	u64 local = 0;
	// two texels per 128-bit load (every counted level has an even number of texels), two loads in flight per thread
	const float4* __restrict__ pairs = reinterpret_cast<const float4*>(texels);
	const u64 numPairs = numTexels >> 1, stride = (u64)blocks * blockDim.x;
	auto add = [&](float mn, float mx) {
		const float a = __fmul_rn(mn, heightF), b = __fmul_rn(mx, heightF);
		// fmaxf/fminf drop a NaN operand: a NaN bound makes every z of the column PARTIAL, as in the reference
		const float lo = fmaxf(floorf(a), 0.0f);
		const float hi = fminf(__fadd_rn(ceilf(b), -1.0f), topF);
		if (!(hi >= lo)) return;
		if (single) {
			local += (u64)(hi - lo) + 1ull;
			return;
		}
		const u32 zl = (u32)lo, zh = (u32)hi;
		for (u32 s = zl >> shift; s <= (zh >> shift); ++s) {  // one or two slices for a surface, many across a box edge
			const u32 from = max(zl, s << shift), to = min(zh, (s << shift) + side - 1u);
			if (shared)
				atomicAdd(&sSlice[s], to - from + 1u);
			else
				atomicAdd(reinterpret_cast<unsigned long long*>(counts + (u64)s * kMaxLevels + level), (unsigned long long)(to - from + 1u));
		}
	};
	u64 i = (u64)block * blockDim.x + threadIdx.x;
	for (; i + stride < numPairs; i += 2 * stride) {
		const float4 t = pairs[i], u = pairs[i + stride];
		add(t.x, t.y);
		add(t.z, t.w);
		add(u.x, u.y);
		add(u.z, u.w);
	}
	if (i < numPairs) {
		const float4 t = pairs[i];
		add(t.x, t.y);
		add(t.z, t.w);
	}
	if (single) {
#pragma unroll
		for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xFFFFFFFFu, local, d);
		if ((threadIdx.x & 31) == 0) sWarp[threadIdx.x >> 5] = local;
		__syncthreads();
		if (threadIdx.x == 0) {
			u64 total = 0;
			for (int w = 0; w < (int)(blockDim.x >> 5); ++w) total += sWarp[w];
			if (total) atomicAdd(reinterpret_cast<unsigned long long*>(counts + level), (unsigned long long)total);
		}
	} else if (shared) {
		__syncthreads();
		for (u32 s = threadIdx.x; s < p.zTileNum; s += blockDim.x)
			if (sSlice[s]) atomicAdd(reinterpret_cast<unsigned long long*>(counts + (u64)s * kMaxLevels + level), (unsigned long long)sSlice[s]);
	}
}

// One tile = kExpandTile consecutive nodes of the level; thread t owns nodes [4t, 4t+4) of the tile.
// (128-thread tiles: more, shorter-lived CTAs per SM hide the per-tile load -> scan -> look-back -> store
// chain better than 256-thread ones.)
constexpr int kExpandThreads = 128;
constexpr int kExpandTile = kExpandThreads * kScanItems;
__global__ void __launch_bounds__(kExpandThreads) expandLevelKernel(const float* __restrict__ tex, u32 side, float heightF, int level0,
		const u64* __restrict__ coords, const u64* __restrict__ nDev, u16* __restrict__ masks, u32* __restrict__ firstChild,
		u64* __restrict__ childCoords, u64 childCap, u64* __restrict__ childN, u32* __restrict__ overflow, ScanLaunch scan,
		const u32* __restrict__ colBias, u32* __restrict__ leafAt) {
	const u64 n = *nDev;
	const u32 numTiles = (u32)((n + kExpandTile - 1) / kExpandTile);
	const u32 tile = scanAcquireTile(scan);
	if (tile >= numTiles) return;  // the grid is sized for the level's capacity
	const u64 base = (u64)tile * kExpandTile + (u64)threadIdx.x * kScanItems;
	u64 c[kScanItems];
	u32 m[kScanItems];
	u64 mine = 0;
#pragma unroll
	for (int i = 0; i < kScanItems; ++i) {
		m[i] = 0;
		c[i] = 0;
		if (base + i < n) {
			c[i] = coords[base + i];
			u32 x, y, z;
			unpackCoord(c[i], x, y, z);
			m[i] = level0 ? childmaskLevel0(tex, side, heightF, x, y, z)
						  : childmaskInner(reinterpret_cast<const float2*>(tex), side, heightF, x, y, z);
			mine += __popc(m[i] & 0xAAAAu);
		}
	}
	u64 pre = mine, dummy = 0, tot, totDummy;
	blockExclusiveScan2<kExpandThreads>(pre, dummy, tot, totDummy);
	u64 tilePre, tilePreB;
	scanLookback2(scan, tile, tot, 0, tilePre, tilePreB);
	if (tile == numTiles - 1 && threadIdx.x == 0) {
		const u64 total = tilePre + tot;
		*childN = total < childCap ? total : childCap;
		if (total > childCap) atomicOr(overflow, kOverflowNodes);
	}
	u64 pos = tilePre + pre;
#pragma unroll
	for (int i = 0; i < kScanItems; ++i) {
		if (base + i >= n) break;
		masks[base + i] = (u16)m[i];
		firstChild[base + i] = (u32)pos;
		u32 partial = m[i] & 0xAAAAu;
		if (partial) {
			u32 x, y, z;
			unpackCoord(c[i], x, y, z);
			if (leafAt) {  // the children are leaves: tell their column where they sit in the level (see leaf columns below)
				while (partial) {
					const u32 child = (__ffs(partial) - 1) >> 1;
					partial &= partial - 1;
					// (positions are in range whenever the pyramid is ordered; a NaN-ridden map fails the count check instead)
					const u32 at = colBias[(size_t)(y + ((child >> 1) & 1u)) * side + x + (child & 1u)] + z + (child >> 2);
					if (at < childCap && pos < childCap) leafAt[at] = (u32)pos;
					++pos;
				}
				continue;
			}
			while (partial) {  // ascending child index (cs::getChildCoordinates, Util.cpp:101-117)
				const u32 child = (__ffs(partial) - 1) >> 1;
				partial &= partial - 1;
				if (pos < childCap) childCoords[pos] = packCoord((x + (child & 1u)) * 2u, (y + ((child >> 1) & 1u)) * 2u, (z + (child >> 2)) * 2u);
				++pos;
			}
		}
	}
}

// The top of the octree: levels of at most kSmallMaxNodes nodes are a chain of tiny dependent steps.
// One CTA walks them all (barriers instead of kernel launches and look-back handshakes).
__global__ void __launch_bounds__(kSmallThreads) expandSmallLevelsKernel(SmallExpandArgs a) {
	__shared__ u32 sWarp[kSmallThreads / 32];
	const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	u32 n = 1;  // the first level is the root
	if (threadIdx.x == 0) *a.rootN = 1;
	for (int s = 0; s < a.count; ++s) {
		const SmallExpandLevel& L = a.lv[s];
		u32 carry = 0;
		for (u32 base = 0; base < n; base += kSmallThreads) {
			const u32 j = base + threadIdx.x;
			u64 c = 0;
			u32 m = 0;
			if (j < n) {
				c = L.coords[j];
				u32 x, y, z;
				unpackCoord(c, x, y, z);
				m = L.level0 ? childmaskLevel0(L.tex, L.side, L.heightF, x, y, z)
							 : childmaskInner(reinterpret_cast<const float2*>(L.tex), L.side, L.heightF, x, y, z);
			}
			const u32 cnt = __popc(m & 0xAAAAu);
			u32 incl = cnt;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const u32 up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
				if ((int)lane >= d) incl += up;
			}
			if (lane == 31) sWarp[warp] = incl;
			__syncthreads();
			u32 before = 0, total = 0;
#pragma unroll
			for (u32 w = 0; w < kSmallThreads / 32; ++w) {
				const u32 v = sWarp[w];
				if (w < warp) before += v;
				total += v;
			}
			__syncthreads();
			if (j < n) {
				u32 pos = carry + before + incl - cnt;
				L.masks[j] = (u16)m;
				L.firstChild[j] = pos;
				u32 partial = m & 0xAAAAu;
				if (partial) {
					u32 x, y, z;
					unpackCoord(c, x, y, z);
					while (partial) {
						const u32 child = (__ffs(partial) - 1) >> 1;
						partial &= partial - 1;
						if (L.leafAt) {
							const u32 at = L.colBias[(size_t)(y + ((child >> 1) & 1u)) * L.side + x + (child & 1u)] + z + (child >> 2);
							if (at < L.childCap && pos < L.childCap) L.leafAt[at] = pos;
						} else if (pos < L.childCap) {
							L.childCoords[pos] = packCoord((x + (child & 1u)) * 2u, (y + ((child >> 1) & 1u)) * 2u, (z + (child >> 2)) * 2u);
						}
						++pos;
					}
				}
			}
			carry += total;
		}
		if (carry > L.childCap) {
			if (threadIdx.x == 0) atomicOr(a.overflow, kOverflowNodes);
			carry = L.childCap;
		}
		if (threadIdx.x == 0) *L.childN = carry;
		n = carry;
		__syncthreads();  // the next level reads the coordinates written above
	}
}

// ---- leaves: cs::createChildmask1x1x8 + createLeafmask (src/CompressedShadowUtil.cpp:59-99) -------
// Eight lanes per level-2 node, lane r owns depth row r of the node's 8x8 texels.
//
// Slice z' of texel i is lit iff (z0+z') + 0.5 <= d_i*H0 (absoluteVisible with minZ=z, maxZ=z+1; exact
// because z < 2^23). Lit slices form a prefix 0..k_i-1 with k_i = clamp(floor(d_i*H0 - (z0 - 0.5)), 0, 8);
// the subtraction is rounded toward -inf so the floor is that of the exact difference. The 64 counts
// k_i determine the eight 64-bit slice masks and vice versa, so a leaf is kept as its "k-code": one
// 32-bit word per row, nibble x = k of texel (x, row). Merging compares k-codes (32 B instead of the
// reference's 68 B leaf); only the leaves that survive merging are expanded to slice masks (emit.cu).
//
// The 1x1x8 childmask needs min k / max k only; k is monotone in depth, so they come from the 8x8
// block's (min,max) texel of pyramid level 3.
// zc8 = (z0 - 0.5) / 8 (exact). With x = (q - (z0 - 0.5)) / 8 evaluated exactly inside the FMA and rounded toward
// -inf, RD(x) >= m/8 <=> x >= m/8 for m = 0..8 (m/8 is a float), so floor(8 * sat(RD(x))) = clamp(floor(q - (z0 - 0.5)), 0, 8).
// NaN saturates to 0: never lit, as midZ <= NaN is false. Four instructions per texel: FMUL, FFMA.RM.SAT, FFMA.RM, IMAD.
__device__ __forceinline__ u32 litCountBits(float depth, float heightF, float zc8) {
	const float q = __fmul_rn(depth, heightF);
	float s;
	asm("fma.rm.sat.f32 %0, %1, 0f3E000000, %2;" : "=f"(s) : "f"(q), "f"(-zc8));
	return (u32)__float_as_int(__fmaf_rd(s, 8.0f, 8388608.0f));  // 0x4B000000 + k
}

// 256 leaves per CTA, in two phases.
//  1. Each lane loads one 32-byte depth row with a single 256-bit load (a warp instruction covers four
//     consecutive leaves; leaves that are x-neighbours share their 128-byte lines, which is what bounds
//     this kernel: L1 tag wavefronts), turns its eight texels into eight nibbles and parks the row word
//     in shared memory. Each warp does its 32 leaves in 8 rounds of 4.
//  2. One thread per leaf reads the finished 32-byte k-code back, hashes it, derives the 1x1x8
//     childmask from the block's (min,max) pyramid texel, marks the distinct-count bitmap and stores.
constexpr int kLeavesPerCta = 256;

__global__ void __launch_bounds__(256, 8) buildLeavesKernel(const float* __restrict__ depth, u32 n, float heightF, const float2* __restrict__ level3,
		const u64* __restrict__ coords, const u64* __restrict__ nDev, u32* __restrict__ codes, u64* __restrict__ hashes, u16* __restrict__ masks,
		u32* __restrict__ bitmap, u32 bitmapWordMask) {
	__shared__ __align__(16) u32 sCode[kLeavesPerCta][8];
	__shared__ u64 sCoord[kLeavesPerCta];
	const u64 numLeaves = *nDev;
	const u64 ctaBase = (u64)blockIdx.x * kLeavesPerCta;
	if (ctaBase >= numLeaves) return;  // the grid is sized for the level's capacity
	const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const u32 row = lane & 7u, which = lane >> 3;
	sCoord[threadIdx.x] = coords[min(ctaBase + threadIdx.x, numLeaves - 1)];
	__syncthreads();
	// the loads of four rounds are issued before any of them is used
#pragma unroll
	for (u32 q0 = 0; q0 < 8; q0 += 4) {
		Float8 d[4];
		float zc[4];
#pragma unroll
		for (u32 i = 0; i < 4; ++i) {
			const u32 l = warp * 32u + (q0 + i) * 4u + which;
			u32 x, y, z;
			unpackCoord(sCoord[l], x, y, z);
			d[i] = ldSector256(depth + (size_t)(y * 4u + row) * n + x * 4u);
			zc[i] = __fmul_rn(__fadd_rn(__uint2float_rn(z * 4u), -0.5f), 0.125f);
		}
#pragma unroll
		for (u32 i = 0; i < 4; ++i) {
			const u32 l = warp * 32u + (q0 + i) * 4u + which;
			u32 code = 0;
#pragma unroll
			for (int t = 7; t >= 0; --t) code = code * 16u + litCountBits(d[i].v[t], heightF, zc[i]);
			sCode[l][row] = code - 0x4B000000u * 0x11111111u;  // strips the float exponent bits of all eight terms
		}
	}
	__syncthreads();

	const u64 leaf = ctaBase + threadIdx.x;
	if (leaf >= numLeaves) return;
	const uint4 c0 = *reinterpret_cast<const uint4*>(&sCode[threadIdx.x][0]), c1 = *reinterpret_cast<const uint4*>(&sCode[threadIdx.x][4]);
	u64 h = 0x9E3779B97F4A7C15ull;
	h = (h ^ (((u64)c0.y << 32) | c0.x)) * 0xFF51AFD7ED558CCDull;
	h = (h ^ (h >> 32) ^ (((u64)c0.w << 32) | c0.z)) * 0xC4CEB9FE1A85EC53ull;
	h = (h ^ (h >> 32) ^ (((u64)c1.y << 32) | c1.x)) * 0xFF51AFD7ED558CCDull;
	h = (h ^ (h >> 32) ^ (((u64)c1.w << 32) | c1.z)) * 0xC4CEB9FE1A85EC53ull;
	h = mix64(h);
	u32 ox, oy, oz;
	unpackCoord(sCoord[threadIdx.x], ox, oy, oz);
	const float zc = __fmul_rn(__fadd_rn(__uint2float_rn(oz * 4u), -0.5f), 0.125f);
	const float2 mm = level3[(size_t)(oy >> 1) * (n >> 3) + (ox >> 1)];
	const u32 kmin = litCountBits(mm.x, heightF, zc) & 15u, kmax = litCountBits(mm.y, heightF, zc) & 15u;
	// slices below kmin are lit (01), slices from kmax up are shadowed (00), the rest PARTIAL (10)
	const u32 below = (1u << (2u * kmin)) - 1u;
	masks[leaf] = (u16)((0x5555u & below) | (0xAAAAu & ((1u << (2u * kmax)) - 1u) & ~below));
	hashes[leaf] = h;
	uint4* dst = reinterpret_cast<uint4*>(codes + leaf * 8);
	dst[0] = c0;
	dst[1] = c1;
	// distinct-count sketch (linear counting): one bit per hash value, read back by sizeLeafTable
	// (read first: on repetitive maps nearly every leaf finds its bit already set, and the atomics of a
	// popular hash would otherwise serialise on one address)
	if (bitmap) {
		const u32 bit = (u32)(h >> 20);
		u32* word = bitmap + ((bit >> 5) & bitmapWordMask);
		if (!(__ldcg(word) & (1u << (bit & 31u)))) atomicOr(word, 1u << (bit & 31u));
	}
}

// ---- leaf columns ---------------------------------------------------------------------------------
// All leaves over one 8x8 texel block share its 64 depths, and by the closed form above they are the
// z-blocks lo..hi of the block's (min,max) texel of pyramid level 3. Building them per column reads every
// depth row once, fully coalesced, instead of once per leaf (3.3 leaves per column on the 16K^2 terrain).
// A leaf's place in its level is decided by the breadth-first expansion, so the expansion of level 3
// scatters that index into leafAt[] at the leaf's column-order position colBias[column] + zb, where
// colBias[c] = (leaves of the columns in front of c, row-major) - lo(c).
__device__ __forceinline__ void columnRange(float2 t, float heightF, float zLoF, float zHiF, float& lo, u32& cnt) {
	const float a = __fmul_rn(t.x, heightF), b = __fmul_rn(t.y, heightF);
	lo = fmaxf(floorf(a), zLoF);  // as countNodesKernel: fmaxf/fminf drop a NaN operand
	const float hi = fminf(__fadd_rn(ceilf(b), -1.0f), zHiF);
	cnt = hi >= lo ? (u32)(hi - lo) + 1u : 0u;
}

__global__ void __launch_bounds__(kScanThreads) columnBiasKernel(const float2* __restrict__ level3, u32 numCols, float heightF, float zLoF,
		float zHiF, u32* __restrict__ colBias, ScanLaunch scan) {
	const u32 tile = scanAcquireTile(scan);
	const u32 base = tile * kScanTile + threadIdx.x * kScanItems;
	float lo[kScanItems];
	u32 cnt[kScanItems];
	u64 mine = 0;
#pragma unroll
	for (int i = 0; i < kScanItems; ++i) {
		lo[i] = 0.f;
		cnt[i] = 0;
		if (base + i < numCols) columnRange(level3[base + i], heightF, zLoF, zHiF, lo[i], cnt[i]);
		mine += cnt[i];
	}
	u64 pre = mine, dummy = 0, tot, totDummy;
	blockExclusiveScan2(pre, dummy, tot, totDummy);
	u64 tilePre, tilePreB;
	scanLookback2(scan, tile, tot, 0, tilePre, tilePreB);
	u32 pos = (u32)(tilePre + pre);
#pragma unroll
	for (int i = 0; i < kScanItems; ++i) {
		if (base + i < numCols) colBias[base + i] = pos - (u32)lo[i];
		pos += cnt[i];
	}
}

// Lit slices of a texel, counted from slice 0 of the whole volume: slice z is lit iff z + 0.5 <= q with q = fl(depth * H0)
// (absoluteVisible), so T = floor(q + 0.5) of them are -- the sum rounded toward -inf never crosses an integer, and NaN
// ends up as 0 (never lit, as midZ <= NaN is false). A leaf at z-block zb then has k = clamp(T - 8 zb, 0, 8).
__device__ __forceinline__ float litSlicesBelow(float q) { return floorf(__fadd_rd(q, 0.5f)); }

// Four lanes per column, lane `sub` owns depth rows 2*sub and 2*sub+1. Per texel the lane keeps
// R = clamp(T - 8 * (first z-block of the chunk), 0, 2040) as an fp16 integer, two texels (x, x+4) per half2 register; a
// leaf's nibbles are then k = 8 * sat(R / 8 - i) for its block i inside the chunk -- two HFMA2 per texel pair, all exact
// (multiples of 1/8 below 256) -- and adding 1024 leaves k in the low bits of each half, so three shift-adds pack a row
// word. Chunks are 252 z-blocks (8 * 252 + 8 <= 2040); taller columns (box edges) re-base R per chunk.
// The column's leaves are walked in batches of four: every lane stores its two row words of each leaf (4 lanes x 8 bytes =
// the leaf's 32-byte k-code), the 64-bit content hash is a sum of per-lane products folded by two shuffles, and lane j
// finishes the j-th leaf of the batch (hash, 1x1x8 mask, sketch bit).
constexpr int kColumnsPerCta = 64;
constexpr u32 kChunkBlocks = 252;
constexpr u32 kNoLeaf = 0xFFFFFFFFu;

struct LeafColumn {  // per lane
	float lo;        // first z-block of the column
	u32 cnt;         // z-blocks (leaves) of the column
	u32 first;       // column-order position of the column's first leaf
	float tMin, tMax;
};
struct LeafSink {
	const u32* __restrict__ leafAt;
	u32 numLeaves;
	u32* __restrict__ codes;
	u16* __restrict__ masks;
	u32* __restrict__ bitmap;
	u32 bitmapWordMask;
};
// Distinct-count sketch update whose read is in flight: the bit is tested (and set if need be) when the next leaf of the
// lane comes along, so nobody waits for the random read. (Read first: on repetitive maps nearly every leaf finds its bit
// already set, and the atomics of a popular hash would otherwise serialise on one address.)
struct PendingSketch {
	u32* word = nullptr;
	u32 bit = 0, seen = 0;
	__device__ __forceinline__ void flush() {
		if (word && !(seen & bit)) atomicOr(word, bit);
		word = nullptr;
	}
	__device__ __forceinline__ void post(u32* w, u32 b) {
		flush();
		word = w;
		bit = b;
		seen = __ldcg(w);
	}
};

// R of the lane's two rows for the chunk starting at z-block `zb0`: clamp(floor(q + 0.5) - 8 zb0, 0, 2040). The sum
// q + (0.5 - 8 zb0) is rounded toward -inf (the addend is exact: 8 zb0 < 2^23), which never crosses an integer, so the floor is
// that of the exact value; the clamps run on the packed halves (below 2048 the conversion is exact, above it stays above 2040,
// NaN is dropped by max(.,0): never lit, as midZ <= NaN is false).
__device__ __forceinline__ void rowsToR(const Float8& a, const Float8& b, float heightF, float zb0, __half2 (&r0)[4], __half2 (&r1)[4]) {
	const float shift = __fsub_rn(0.5f, __fmul_rn(zb0, 8.0f));
	const __half2 zero2 = __float2half2_rn(0.f), cap2 = __float2half2_rn(2040.f);
#pragma unroll
	for (int p = 0; p < 4; ++p) {
		const float x0 = floorf(__fadd_rd(__fmul_rn(a.v[p], heightF), shift)), x1 = floorf(__fadd_rd(__fmul_rn(a.v[p + 4], heightF), shift));
		const float y0 = floorf(__fadd_rd(__fmul_rn(b.v[p], heightF), shift)), y1 = floorf(__fadd_rd(__fmul_rn(b.v[p + 4], heightF), shift));
		r0[p] = __hmin2(__hmax2(__floats2half2_rn(x0, x1), zero2), cap2);
		r1[p] = __hmin2(__hmax2(__floats2half2_rn(y0, y1), zero2), cap2);
	}
}

__constant__ u64 kLaneHashMul[4] = {0x9E3779B97F4A7C15ull, 0xC2B2AE3D27D4EB4Full, 0x165667B19E3779F9ull, 0xD6E8FEB86659FD93ull};

// The leaves chunk..chunkEnd-1 (warp-uniform bounds) of the eight columns of a warp. firstIdx: leafAt of the lane's leaf of
// the very first batch if the caller fetched it ahead (only looked at for chunk == 0).
// kSketch = false: the leaf table will be sized from the previous build's count of distinct leaves, so neither the hash nor
// the sketch is needed (a third of the instructions of a batch).
template <bool kSketch>
__device__ __forceinline__ void emitColumnLeaves(const LeafColumn& c, const __half2 (&r0)[4], const __half2 (&r1)[4], u32 chunk, u32 chunkEnd,
		u32 firstIdx, bool haveFirstIdx, const LeafSink& out, PendingSketch& pending) {
	const u32 lane = threadIdx.x & 31u, sub = lane & 3u, groupLane = lane & ~3u;
	const u64 mulK = kLaneHashMul[sub];
	const __half2 eighth2 = __float2half2_rn(0.125f), eight2 = __float2half2_rn(8.0f), bias2 = __float2half2_rn(1024.0f),
				  one2 = __float2half2_rn(1.0f);
	constexpr u32 kStrip = 0x64006400u * 0x1111u;  // the fp16 bits of 1024 under all eight nibbles of a row word
	__half2 negI2 = __float2half2_rn(0.f);         // -(block index inside the chunk)
	for (u32 m = chunk; m < chunkEnd; m += 4) {
		u32 mineIdx = kNoLeaf;
		if (haveFirstIdx && m == 0) {
			mineIdx = firstIdx;
		} else if (m + sub < c.cnt && c.first + m + sub < out.numLeaves) {
			mineIdx = out.leafAt[c.first + m + sub];
		}
		if (mineIdx >= out.numLeaves) mineIdx = kNoLeaf;  // cannot happen on an ordered pyramid (see the expansion)
		u64 hp[4] = {0, 0, 0, 0};
#pragma unroll
		for (u32 j = 0; j < 4; ++j) {
			if (m + j >= chunkEnd) break;  // warp-uniform
			const u32 leaf = __shfl_sync(0xFFFFFFFFu, mineIdx, groupLane + j);
			u32 w0 = 0, w1 = 0;
#pragma unroll
			for (int p = 3; p >= 0; --p) {
				const __half2 y0 = __hfma2(__hfma2_sat(r0[p], eighth2, negI2), eight2, bias2);
				const __half2 y1 = __hfma2(__hfma2_sat(r1[p], eighth2, negI2), eight2, bias2);
				w0 = w0 * 16u + *reinterpret_cast<const u32*>(&y0);
				w1 = w1 * 16u + *reinterpret_cast<const u32*>(&y1);
			}
			w0 -= kStrip;
			w1 -= kStrip;
			negI2 = __hsub2(negI2, one2);
			if (leaf != kNoLeaf) *reinterpret_cast<uint2*>(out.codes + (u64)leaf * 8u + sub * 2u) = make_uint2(w0, w1);
			if (kSketch) hp[j] = ((((u64)w1) << 32) | w0) * mulK;
		}
		u64 keep = 0;
		if (kSketch) {
			// lane `sub` ends up with the sum over the four lanes of hp[sub]: a 4 x 4 reduce-scatter in two exchanges
			const bool hi = (sub & 2u) != 0, odd = (sub & 1u) != 0;
			const u64 k0 = (hi ? hp[2] : hp[0]) + __shfl_xor_sync(0xFFFFFFFFu, hi ? hp[0] : hp[2], 2);
			const u64 k1 = (hi ? hp[3] : hp[1]) + __shfl_xor_sync(0xFFFFFFFFu, hi ? hp[1] : hp[3], 2);
			keep = (odd ? k1 : k0) + __shfl_xor_sync(0xFFFFFFFFu, odd ? k0 : k1, 1);
		}
		if (mineIdx != kNoLeaf) {
			const float z8 = __fmul_rn(__fadd_rn(c.lo, __uint2float_rn(m + sub)), 8.0f);
			const u32 kmin = (u32)fminf(fmaxf(__fsub_rn(c.tMin, z8), 0.f), 8.f), kmax = (u32)fminf(fmaxf(__fsub_rn(c.tMax, z8), 0.f), 8.f);
			// slices below kmin are lit (01), slices from kmax up are shadowed (00), the rest PARTIAL (10)
			const u32 below = (1u << (2u * kmin)) - 1u;
			out.masks[mineIdx] = (u16)((0x5555u & below) | (0xAAAAu & ((1u << (2u * kmax)) - 1u) & ~below));
			// the hash only feeds the sketch here: the insert recomputes its own from the code it reads anyway, which is
			// cheaper than 8-byte stores scattered over the level (partial sectors: a fill and a write-back each)
			if (kSketch) {
				const u64 h = mix64(keep);
				const u32 bit = (u32)(h >> 20);
				pending.post(out.bitmap + ((bit >> 5) & out.bitmapWordMask), 1u << (bit & 31u));
			}
		}
	}
}

// Persistent warps, each walking groups of eight x-adjacent columns (two 256-bit loads per lane; a warp instruction covers
// four rows of 256 contiguous bytes), software-pipelined in registers so that no load is waited for: a group's level-3
// texel and colBias are fetched two groups ahead; its depth rows and the leafAt of its first batch one group ahead, and only
// if the column has leaves (most columns of a z-slice of a tall grid have none). Warps never synchronise with each other.
// tallOnly != NULL: the residue kernel below has built every column of at most kResidueBlocks z-blocks; this launch only
// walks the taller ones (box edges), and not at all if *tallOnly says there are none.
constexpr u32 kResidueBlocks = 31;
template <bool kSketch>
__global__ void __launch_bounds__(256, 3) buildLeafColumnsKernel(const float* __restrict__ depth, u32 n, u32 colShift, float heightF,
		float height3F, float zLoF, float zHiF, const float2* __restrict__ level3, u32 numCols, const u32* __restrict__ colBias, LeafSink out,
		const u32* __restrict__ tallOnly) {
	if (tallOnly && *tallOnly == 0u) return;
	const float topF = __fadd_rn(height3F, -1.0f);
	const u32 lane = threadIdx.x & 31u, sub = lane & 3u;
	const u32 numWarps = gridDim.x * (blockDim.x >> 5), numGroups = (numCols + 7u) >> 3;
	u32 group = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	PendingSketch pending;

	// stage 2 (two groups ahead): level-3 texel and bias
	float2 mm2 = make_float2(0.f, 0.f);
	u32 bias2 = 0;
	auto fetchTexel = [&](u32 g) {
		const u32 col = g * 8u + (lane >> 2);
		mm2 = make_float2(0.f, 0.f);  // an empty column (lo > hi)
		bias2 = 0;
		if (g < numGroups && col < numCols) {
			mm2 = level3[col];
			bias2 = colBias[col];
		}
	};
	// stage 1 (one group ahead): column range, depth rows, first leafAt batch
	LeafColumn c1;
	float2 mm1 = make_float2(0.f, 0.f);
	Float8 a1, b1;
	u32 idx1 = kNoLeaf;
	auto fetchRows = [&](u32 g) {  // consumes stage 2
		mm1 = mm2;
		c1.lo = 0.f;
		c1.cnt = 0;
		const u32 col = g * 8u + (lane >> 2);
		if (g < numGroups && col < numCols) columnRange(mm1, height3F, zLoF, zHiF, c1.lo, c1.cnt);
		if (tallOnly && c1.cnt) {  // (the residue kernel decides by the column's extent over all z-slices)
			float loAll;
			u32 cntAll;
			columnRange(mm1, height3F, 0.0f, topF, loAll, cntAll);
			if (cntAll <= kResidueBlocks) c1.cnt = 0;
		}
		c1.first = bias2 + (u32)c1.lo;
		idx1 = kNoLeaf;
		if (c1.cnt) {
			const u32 cx = col & ((1u << colShift) - 1u), cy = col >> colShift;
			const float* p = depth + (size_t)(cy * 8u + sub * 2u) * n + cx * 8u;
			a1 = ldSector256(p);
			b1 = ldSector256(p + n);
			if (sub < c1.cnt && c1.first + sub < out.numLeaves) idx1 = out.leafAt[c1.first + sub];
		}
	};
	fetchTexel(group);
	fetchRows(group);
	fetchTexel(group + numWarps);
	for (; group < numGroups; group += numWarps) {
		// stage 0: this group. Its rows become R right away so that the row registers can take the next group's loads.
		LeafColumn c = c1;
		c.tMin = litSlicesBelow(__fmul_rn(mm1.x, heightF));
		c.tMax = litSlicesBelow(__fmul_rn(mm1.y, heightF));
		const u32 firstIdx = idx1;
		const u32 maxCnt = __reduce_max_sync(0xFFFFFFFFu, c.cnt);
		__half2 r0[4], r1[4];
		if (c.cnt) {
			rowsToR(a1, b1, heightF, c.lo, r0, r1);
		} else {
#pragma unroll
			for (int p = 0; p < 4; ++p) r0[p] = r1[p] = __float2half2_rn(0.f);
		}
		fetchRows(group + numWarps);
		fetchTexel(group + 2u * numWarps);
		if (maxCnt == 0) continue;
		emitColumnLeaves<kSketch>(c, r0, r1, 0, min(maxCnt, kChunkBlocks), firstIdx, true, out, pending);
		for (u32 chunk = kChunkBlocks; chunk < maxCnt; chunk += kChunkBlocks) {  // tall columns (box edges): re-read the rows
			if (c.cnt) {
				const u32 col = group * 8u + (lane >> 2);
				const u32 cx = col & ((1u << colShift) - 1u), cy = col >> colShift;
				const float* p = depth + (size_t)(cy * 8u + sub * 2u) * n + cx * 8u;
				rowsToR(ldSector256(p), ldSector256(p + n), heightF, __fadd_rn(c.lo, __uint2float_rn(chunk)), r0, r1);
			}
			emitColumnLeaves<kSketch>(c, r0, r1, chunk, min(maxCnt, chunk + kChunkBlocks), kNoLeaf, false, out, pending);
		}
	}
	pending.flush();
}

// ---- leaf columns from the hierarchy's residues -----------------------------------------------------------------------
// The same leaves from the re-encoded depth map the pyramid's base kernel leaves behind (pyramid.cu): byte (row, x) of a
// column = clamp(T - 8 * lo, 0, 255) with lo the column's first z-block over ALL z-slices -- the R of rowsToR above, already
// computed. A lane owns rows 2*sub and 2*sub+1: one 16-byte load instead of two 32-byte ones, and a byte permute plus a
// subtraction per texel pair instead of the float arithmetic (the bytes of a pair sit next to each other; 0x64 above a
// byte is the fp16 1024 + byte). Columns taller than kResidueBlocks are left to buildLeafColumnsKernel.
template <bool kSketch>
__global__ void __launch_bounds__(256, 4) buildLeafColumnsResidueKernel(const uint4* __restrict__ residue, float heightF, float height3F,
		float zLoF, float zHiF, const float2* __restrict__ level3, u32 numCols, const u32* __restrict__ colBias, LeafSink out, u32* __restrict__ tallFlag) {
	const u32 lane = threadIdx.x & 31u, sub = lane & 3u;
	const u32 numWarps = gridDim.x * (blockDim.x >> 5), numGroups = (numCols + 7u) >> 3;
	const float topF = __fadd_rn(height3F, -1.0f);
	u32 group = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	PendingSketch pending;
	bool sawTall = false;

	// stage 2 (two groups ahead): level-3 texel and bias
	float2 mm2 = make_float2(0.f, 0.f);
	u32 bias2 = 0;
	auto fetchTexel = [&](u32 g) {
		const u32 col = g * 8u + (lane >> 2);
		mm2 = make_float2(0.f, 0.f);  // an empty column (lo > hi)
		bias2 = 0;
		if (g < numGroups && col < numCols) {
			mm2 = level3[col];
			bias2 = colBias[col];
		}
	};
	// stage 1 (one group ahead): column range, residue rows, first leafAt batch
	LeafColumn c1;
	float delta1 = 0.f;
	u32 idx1 = kNoLeaf;
	uint4 r1 = make_uint4(0u, 0u, 0u, 0u);
	auto fetchRows = [&](u32 g) {  // consumes stage 2
		c1.lo = 0.f;
		c1.cnt = 0;
		delta1 = 0.f;
		const u32 col = g * 8u + (lane >> 2);
		if (g < numGroups && col < numCols) {
			float loAll;
			u32 cntAll;
			columnRange(mm2, height3F, 0.0f, topF, loAll, cntAll);
			columnRange(mm2, height3F, zLoF, zHiF, c1.lo, c1.cnt);
			if (cntAll > kResidueBlocks) {  // does not fit a byte: the depth-based kernel builds this column
				sawTall = sawTall || c1.cnt != 0;
				c1.cnt = 0;
			}
			delta1 = __fmul_rn(__fsub_rn(c1.lo, loAll), 8.0f);  // slices between the column's first block and this z-slice's first block
		}
		c1.tMin = litSlicesBelow(__fmul_rn(mm2.x, heightF));
		c1.tMax = litSlicesBelow(__fmul_rn(mm2.y, heightF));
		c1.first = bias2 + (u32)c1.lo;
		idx1 = kNoLeaf;
		if (c1.cnt) {
			r1 = residue[(size_t)col * 4u + sub];
			if (sub < c1.cnt && c1.first + sub < out.numLeaves) idx1 = out.leafAt[c1.first + sub];
		}
	};
	fetchTexel(group);
	fetchRows(group);
	fetchTexel(group + numWarps);
	for (; group < numGroups; group += numWarps) {
		// stage 0: this group. Its residues become R right away so that the registers can take the next group's loads.
		const LeafColumn c = c1;
		const u32 firstIdx = idx1;
		const u32 maxCnt = __reduce_max_sync(0xFFFFFFFFu, c.cnt);
		__half2 r0[4], r1h[4];
		{
			const u32 words[4] = {r1.x, r1.y, r1.z, r1.w};
			// 1024 + byte as fp16, then minus (1024 + the slice's offset), clamped at 0 (whole-volume builds: offset 0)
			const __half2 base2 = __float2half2_rn(__fadd_rn(1024.0f, fminf(delta1, 1016.0f))), zero2 = __float2half2_rn(0.f);
#pragma unroll
			for (int p = 0; p < 4; ++p) {
				const u32 a = __byte_perm(words[p >> 1], 0x64646464u, (p & 1) ? 0x4342u : 0x4140u);
				const u32 b = __byte_perm(words[2 + (p >> 1)], 0x64646464u, (p & 1) ? 0x4342u : 0x4140u);
				r0[p] = __hmax2(__hsub2(*reinterpret_cast<const __half2*>(&a), base2), zero2);
				r1h[p] = __hmax2(__hsub2(*reinterpret_cast<const __half2*>(&b), base2), zero2);
			}
		}
		fetchRows(group + numWarps);
		fetchTexel(group + 2u * numWarps);
		if (maxCnt == 0) continue;
		emitColumnLeaves<kSketch>(c, r0, r1h, 0, maxCnt, firstIdx, true, out, pending);
	}
	pending.flush();
	if (__any_sync(0xFFFFFFFFu, sawTall) && lane == 0) atomicOr(tallFlag, 1u);
}

// Number of set bits of the sketch -> *setBits (zeroed beforehand).
__global__ void __launch_bounds__(256) sketchPopcountKernel(const uint4* __restrict__ bitmap, u32 numVec, u64* __restrict__ setBits) {
	u32 local = 0;
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < numVec; i += gridDim.x * blockDim.x) {
		const uint4 v = bitmap[i];
		local += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
	}
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xFFFFFFFFu, local, d);
	if ((threadIdx.x & 31) == 0 && local) atomicAdd(reinterpret_cast<unsigned long long*>(setBits), (unsigned long long)local);
}

__global__ void childmaskKernel(const float* tex, u32 side, float heightF, int level0, u32 x, u32 y, u32 z, u32* out) {
	*out = level0 ? childmaskLevel0(tex, side, heightF, x, y, z) : childmaskInner(reinterpret_cast<const float2*>(tex), side, heightF, x, y, z);
}

}  // namespace

int launchChildmask(const PyramidView& pyr, int level, u32 zTileNum, u32 x, u32 y, u32 z, u32* out, cudaStream_t stream) {
	const u32 side = (u32)pyr.n >> level;
	childmaskKernel<<<1, 1, 0, stream>>>(pyr.level[level], side, (float)(side * zTileNum), level == 0 ? 1 : 0, x, y, z, out);
	return 1;
}

int launchColumnCounts(const PyramidView& pyr, u32 zTileNum, int minLevel, u64* counts, cudaStream_t stream) {
	const int numCounted = pyr.numLevels - 2 - minLevel;  // levels minLevel .. numLevels-3
	if (numCounted <= 0) return 0;
	CountLevels p;
	p.minLevel = minLevel;
	p.topCounted = pyr.numLevels - 3;
	p.zTileNum = zTileNum;
	u32 totalBlocks = 0;
	for (int level = minLevel; level <= pyr.numLevels - 3; ++level) {
		const u32 side = (u32)pyr.n >> (level + 1);
		p.texels[level] = reinterpret_cast<const float2*>(pyr.level[level + 1]);
		p.numTexels[level] = (u64)side * side;
		p.heightF[level] = (float)(side * zTileNum);
		p.sideShift[level] = 0;
		while ((1u << p.sideShift[level]) < side) ++p.sideShift[level];
		u64 blocks = (p.numTexels[level] + 256 * 8 - 1) / (256 * 8);  // two passes of two 2-texel loads per thread
		if (blocks > 148 * 8) blocks = 148 * 8;
		p.blockStart[level] = totalBlocks;
		totalBlocks += (u32)blocks;
	}
	p.blockStart[pyr.numLevels - 2] = totalBlocks;
	countNodesKernel<<<totalBlocks, 256, 0, stream>>>(p, counts);
	return 1;
}

int launchExpandSmallLevels(const SmallExpandArgs& a, cudaStream_t stream) {
	if (a.count <= 0) return 0;
	expandSmallLevelsKernel<<<1, kSmallThreads, 0, stream>>>(a);
	return 1;
}

int launchExpandLevel(const PyramidView& pyr, int level, u32 zTileNum, const u64* coords, const u64* nDev, u64 cap, u16* masks, u32* firstChild,
		u64* childCoords, u64 childCap, u64* childN, u32* overflow, ScanLaunch scan, const u32* colBias, u32* leafAt, cudaStream_t stream) {
	const u32 side = (u32)pyr.n >> level;
	const float heightF = (float)(side * zTileNum);
	const u32 tiles = (u32)((cap + kExpandTile - 1) / kExpandTile);
	if (!tiles) return 0;
	expandLevelKernel<<<tiles, kExpandThreads, 0, stream>>>(pyr.level[level], side, heightF, level == 0 ? 1 : 0, coords, nDev, masks, firstChild,
			childCoords, childCap, childN, overflow, scan, colBias, leafAt);
	return 1;
}

int launchBuildLeaves(const PyramidView& pyr, u32 zTileNum, const u64* coords, const u64* nDev, u64 cap, u32* codes, u64* hashes, u16* masks,
		u32* sketch, cudaStream_t stream) {
	const float heightF = (float)((u32)pyr.n * zTileNum);
	if (!cap) return 0;
	buildLeavesKernel<<<(unsigned)((cap + kLeavesPerCta - 1) / kLeavesPerCta), 256, 0, stream>>>(pyr.level[0], (u32)pyr.n, heightF,
			reinterpret_cast<const float2*>(pyr.level[3]), coords, nDev, codes, hashes, masks, sketch, kSketchWords - 1);
	return 1;
}

int launchColumnBias(const PyramidView& pyr, u32 zTileIndex, u32 zTileNum, u32* colBias, ScanLaunch scan, cudaStream_t stream) {
	const u32 side3 = (u32)pyr.n >> 3, numCols = side3 * side3;
	columnBiasKernel<<<(numCols + kScanTile - 1) / kScanTile, kScanThreads, 0, stream>>>(reinterpret_cast<const float2*>(pyr.level[3]), numCols,
			(float)(side3 * zTileNum), (float)(zTileIndex * side3), (float)(zTileIndex * side3 + side3 - 1), colBias, scan);
	return 1;
}

int launchBuildLeafColumns(const PyramidView& pyr, u32 zTileIndex, u32 zTileNum, const u32* colBias, const u32* leafAt, u32 numLeaves, u32* codes,
		u16* masks, u32* sketch, const unsigned char* residue, u32* tallFlag, cudaStream_t stream) {
	const u32 side3 = (u32)pyr.n >> 3, numCols = side3 * side3;
	u32 colShift = 0;
	while ((1u << colShift) < side3) ++colShift;
	const float heightF = (float)((u32)pyr.n * zTileNum), height3F = (float)(side3 * zTileNum);
	const float zLoF = (float)(zTileIndex * side3), zHiF = (float)(zTileIndex * side3 + side3 - 1);
	const float2* level3 = reinterpret_cast<const float2*>(pyr.level[3]);
	LeafSink out{leafAt, numLeaves, codes, masks, sketch, kSketchWords - 1};
	const u32 wanted = (numCols + kColumnsPerCta - 1) / kColumnsPerCta;  // 8 warps of 8 columns per CTA
	const u32 gridResidue = wanted < 148u * 4u ? wanted : 148u * 4u, gridDepth = wanted < 148u * 3u ? wanted : 148u * 3u;
	const u32* tallOnly = residue ? tallFlag : nullptr;
	int launches = 1;
	if (residue) {
		const uint4* res4 = reinterpret_cast<const uint4*>(residue);
		if (sketch)
			buildLeafColumnsResidueKernel<true><<<gridResidue, 256, 0, stream>>>(res4, heightF, height3F, zLoF, zHiF, level3, numCols, colBias, out, tallFlag);
		else
			buildLeafColumnsResidueKernel<false><<<gridResidue, 256, 0, stream>>>(res4, heightF, height3F, zLoF, zHiF, level3, numCols, colBias, out, tallFlag);
		++launches;
	}
	if (sketch)
		buildLeafColumnsKernel<true><<<gridDepth, 256, 0, stream>>>(pyr.level[0], (u32)pyr.n, colShift, heightF, height3F, zLoF, zHiF, level3, numCols, colBias,
				out, tallOnly);
	else
		buildLeafColumnsKernel<false><<<gridDepth, 256, 0, stream>>>(pyr.level[0], (u32)pyr.n, colShift, heightF, height3F, zLoF, zHiF, level3, numCols, colBias,
				out, tallOnly);
	return launches;
}

int launchSketchPopcount(const u32* sketch, u64* setBits, cudaStream_t stream) {
	sketchPopcountKernel<<<148 * 4, 256, 0, stream>>>(reinterpret_cast<const uint4*>(sketch), kSketchWords / 4, setBits);
	return 1;
}

}  // namespace cpvs
