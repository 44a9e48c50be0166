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
#include "kernels.h"

namespace cpvs {

namespace {

// ---- node counts per level, for exact allocation (closed form of the classification) ------------
// A level-l node exists for every voxel (x,y,z) of pyramid level l+1 that classifies PARTIAL:
// !(z+1 <= min*H) && !(z >= max*H)  <=>  floor(min*H) <= z <= ceil(max*H)-1, inside the z-tile.
__global__ void countNodesKernel(const float2* __restrict__ texels, u64 numTexels, float heightF, float zLoF, float zHiF,
		u64* __restrict__ count) {
	u64 local = 0;
	for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < numTexels; i += (u64)gridDim.x * blockDim.x) {
		const float2 t = texels[i];
		const float a = __fmul_rn(t.x, heightF), b = __fmul_rn(t.y, heightF);
		// fmaxf/fminf drop a NaN operand: a NaN bound makes every z of the tile PARTIAL, as in the reference
		const float lo = fmaxf(floorf(a), zLoF);
		const float hi = fminf(__fadd_rn(ceilf(b), -1.0f), zHiF);
		if (hi >= lo) local += (u64)(hi - lo) + 1ull;
	}
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xFFFFFFFFu, local, d);
	__shared__ u64 sWarp[8];
	if ((threadIdx.x & 31) == 0) sWarp[threadIdx.x >> 5] = local;
	__syncthreads();
	if (threadIdx.x == 0) {
		u64 total = 0;
		for (int w = 0; w < (int)(blockDim.x >> 5); ++w) total += sWarp[w];
		if (total) atomicAdd(reinterpret_cast<unsigned long long*>(count), (unsigned long long)total);
	}
}

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

// One tile = kScanTile consecutive nodes of the level; thread t owns nodes [4t, 4t+4) of the tile.
__global__ void __launch_bounds__(kScanThreads) expandLevelKernel(const float* __restrict__ tex, u32 side, float heightF, int level0,
		const u64* __restrict__ coords, u64 n, u16* __restrict__ masks, u32* __restrict__ firstChild, u64* __restrict__ childCoords,
		u64* __restrict__ childTotal, ScanLaunch scan, u32 numTiles) {
	const u32 tile = scanAcquireTile(scan);
	const u64 base = (u64)tile * kScanTile + (u64)threadIdx.x * kScanItems;
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
	blockExclusiveScan2(pre, dummy, tot, totDummy);
	u64 tilePre, tilePreB;
	scanLookback2(scan, tile, tot, 0, tilePre, tilePreB);
	if (tile == numTiles - 1 && threadIdx.x == 0) *childTotal = tilePre + tot;
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
			while (partial) {  // ascending child index (cs::getChildCoordinates, Util.cpp:101-117)
				const u32 child = (__ffs(partial) - 1) >> 1;
				partial &= partial - 1;
				childCoords[pos++] = packCoord((x + (child & 1u)) * 2u, (y + ((child >> 1) & 1u)) * 2u, (z + (child >> 2)) * 2u);
			}
		}
	}
}

// ---- leaves: cs::createChildmask1x1x8 + createLeafmask (src/CompressedShadowUtil.cpp:59-99) -------
// Eight lanes per level-2 node, lane r owns depth row r of the node's 8x8 texels.
//
// Slice z' of texel i is lit iff (z0+z') + 0.5 <= d_i*H0 (absoluteVisible with minZ=z, maxZ=z+1; exact
// because z < 2^23). Lit slices form a prefix 0..k_i-1 with k_i = clamp(floor(d_i*H0 - (z0 - 0.5)), 0, 8);
// the subtraction is rounded toward -inf so the floor is that of the exact difference. Byte i of `t`
// is the thermometer code (1<<k_i)-1; an 8x8 bit transpose turns texel-major into slice-major bytes,
// and a byte exchange across the eight lanes leaves lane z' with the 64-bit mask of slice z'.
__device__ __forceinline__ u64 transpose8x8(u64 x) {
	u64 t;
	t = (x ^ (x >> 7)) & 0x00AA00AA00AA00AAull;
	x = x ^ t ^ (t << 7);
	t = (x ^ (x >> 14)) & 0x0000CCCC0000CCCCull;
	x = x ^ t ^ (t << 14);
	t = (x ^ (x >> 28)) & 0x00000000F0F0F0F0ull;
	x = x ^ t ^ (t << 28);
	return x;
}

__global__ void __launch_bounds__(256) buildLeavesKernel(const float* __restrict__ depth, u32 n, float heightF, const u64* __restrict__ coords,
		u64 numLeaves, u64* __restrict__ bits, u64* __restrict__ hashes, u16* __restrict__ masks) {
	const u64 gid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	const u32 r = threadIdx.x & 7u;
	u64 leaf = gid >> 3;
	const bool live = leaf < numLeaves;
	if (!live) leaf = numLeaves - 1;  // keep the lane in the shuffles
	u32 ox, oy, oz;
	unpackCoord(coords[leaf], ox, oy, oz);
	const float* row = depth + (size_t)(oy * 4u + r) * n + ox * 4u;
	const float4 a = *reinterpret_cast<const float4*>(row), b = *reinterpret_cast<const float4*>(row + 4);
	const float zc = __fadd_rn(__uint2float_rn(oz * 4u), -0.5f);
	const float d[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
	u32 lo = 0, hi = 0;
#pragma unroll
	for (int i = 0; i < 8; ++i) {
		const float v = __fadd_rd(__fmul_rn(d[i], heightF), -zc);
		const float kf = fminf(fmaxf(v, 0.0f), 8.0f);  // NaN -> 0: never lit, as midZ <= NaN is false
		const u32 k = (u32)__float_as_int(__fadd_rd(kf, 8388608.0f)) & 15u;
		const u32 therm = (1u << k) - 1u;
		if (i < 4)
			lo |= therm << (8 * i);
		else
			hi |= therm << (8 * (i - 4));
	}
	const u64 rows = transpose8x8(((u64)hi << 32) | lo);  // byte z' = this row's 8 texels in slice z'
	const u32 rlo = (u32)rows, rhi = (u32)(rows >> 32);
	const int group = (threadIdx.x & 31) & ~7;
	u64 slice = 0;
#pragma unroll
	for (int s = 0; s < 8; ++s) {
		const u32 olo = __shfl_sync(0xFFFFFFFFu, rlo, group + s), ohi = __shfl_sync(0xFFFFFFFFu, rhi, group + s);
		const u32 byte = ((r < 4u ? olo : ohi) >> (8u * (r & 3u))) & 0xFFu;
		slice |= (u64)byte << (8 * s);
	}
	// lane r now holds the mask of slice z' = r
	u32 code = (slice == ~0ull) ? 1u : (slice == 0ull ? 0u : 2u);
	u32 mask = code << (2u * r);
	u64 h = mix64(slice + (u64)(r + 1u) * 0x9E3779B97F4A7C15ull);
#pragma unroll
	for (int dlt = 1; dlt < 8; dlt <<= 1) {
		mask |= __shfl_xor_sync(0xFFFFFFFFu, mask, dlt);
		h += __shfl_xor_sync(0xFFFFFFFFu, h, dlt);
	}
	if (live) {
		bits[leaf * 8 + r] = slice;
		if (r == 0) {
			hashes[leaf] = mix64(h);
			masks[leaf] = (u16)mask;
		}
	}
}

}  // namespace

int launchCountNodes(const PyramidView& pyr, u32 zTileIndex, u32 zTileNum, int minLevel, u64* counts, cudaStream_t stream) {
	int launches = 0;
	for (int level = minLevel; level <= pyr.numLevels - 3; ++level) {
		const u32 side = (u32)pyr.n >> (level + 1);
		const u64 texels = (u64)side * side;
		const float heightF = (float)(side * zTileNum);
		const float zLo = (float)(zTileIndex * side), zHi = (float)(zTileIndex * side + side - 1);
		u64 blocks = (texels + 255) / 256;
		if (blocks > 148 * 16) blocks = 148 * 16;
		countNodesKernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const float2*>(pyr.level[level + 1]), texels, heightF, zLo, zHi,
				counts + level);
		++launches;
	}
	return launches;
}

int launchExpandLevel(const PyramidView& pyr, int level, u32 zTileNum, const u64* coords, u64 n, u16* masks, u32* firstChild,
		u64* childCoords, u64* childTotal, ScanLaunch scan, cudaStream_t stream) {
	const u32 side = (u32)pyr.n >> level;
	const float heightF = (float)(side * zTileNum);
	const u32 tiles = (u32)((n + kScanTile - 1) / kScanTile);
	expandLevelKernel<<<tiles, kScanThreads, 0, stream>>>(pyr.level[level], side, heightF, level == 0 ? 1 : 0, coords, n, masks, firstChild,
			childCoords, childTotal, scan, tiles);
	return 1;
}

int launchBuildLeaves(const PyramidView& pyr, u32 zTileNum, const u64* coords, u64 n, u64* bits, u64* hashes, u16* masks,
		cudaStream_t stream) {
	const float heightF = (float)((u32)pyr.n * zTileNum);
	const u64 threads = n * 8;
	buildLeavesKernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(pyr.level[0], (u32)pyr.n, heightF, coords, n, bits, hashes, masks);
	return 1;
}

}  // namespace cpvs
