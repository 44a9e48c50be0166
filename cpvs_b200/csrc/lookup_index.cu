// A private, lookup-only copy of a finished DAG (or of a container's concatenated DAGs), built once from the DAG words alone.
//
// The wire format is made for size: an inner node is a mask word followed by one pointer per PARTIAL child, so a step of the
// descent is two dependent loads (the mask, then -- at a position the mask decides -- the pointer); a leaf is a mask word plus
// up to eight 64-bit slice masks at an arbitrary word offset, again two dependent loads that land in one of up to three 32-byte
// sectors of a 200 MB level (deduplicated leaves have no locality: DRAM accesses). The copy is made for lookups:
//   nodes   every inner node as EIGHT words, one per child, 32-byte aligned: 0 = shadow, 1 = lit, 2 + id = the child node (or,
//           in the level above the leaves, the leaf) -- the reference's uncompressed node (src/CompressedShadowUtil.h:9) with
//           the mask folded into the slots. A step of the descent is ONE load whose address the path alone decides.
//   codes   one 32-byte, 32-byte-aligned k-code per distinct leaf: nibble x of word y = lit slices of texel (x, y). The eight
//           slices of a leaf built from a depth map are nested (slice z lit => slice z-1 lit), so the counts describe it fully and
//           the last step is ONE load, independent of the leaf's mask: lit <=> (z & 7) < nibble.
//   grid    the container's cell table with the cells' root node ids
// About the size of the wire format (32 B per inner node instead of ~20, 32 B per leaf instead of ~80), half the dependent
// loads, one sector per access. The DAG words themselves are untouched (getDAG, save / load and parity tests see the
// reference's format); a leaf whose slices are not nested -- possible only in words that did not come from a depth map --
// makes the index invalid and lookups keep walking the wire format.
#include <algorithm>
#include <vector>

#include "handles.h"

namespace cpvs {

namespace {


__global__ void __launch_bounds__(256) markRootsKernel(const u64* __restrict__ roots, u32 n, u32* __restrict__ nodeBits) {
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const u32 s = (u32)roots[i];
	atomicOr(&nodeBits[s >> 5], 1u << (s & 31u));
}

// Node list entries: cell << 32 | absolute word offset of a node. Children not seen before (one bit per DAG word) are appended to the list.
__global__ void __launch_bounds__(256) expandFrontierKernel(const u32* __restrict__ dag, const u64* __restrict__ cellStart, const u64* __restrict__ in, u32 nIn,
		u32* __restrict__ visited, u64* __restrict__ out, u32* __restrict__ outCount, u32 outCapacity, u32* __restrict__ error) {
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nIn) return;
	const u64 e = in[i];
	const u32 cell = (u32)(e >> 32), s = (u32)e;
	const u64 base = cellStart[cell];
	const u32 mask = dag[s];
	u32 partial = mask & 0xAAAAu, slot = 1;
	while (partial) {
		partial &= partial - 1;
		const u64 child = base + dag[s + slot++];
		const u32 bit = 1u << (child & 31u);
		if (!(atomicOr(&visited[child >> 5], bit) & bit)) {
			const u32 at = atomicAdd(outCount, 1u);
			if (at < outCapacity)
				out[at] = ((u64)cell << 32) | child;
			else
				atomicExch(error, 1u);
		}
	}
}

// The nodes above the leaves: mark every leaf start.
__global__ void __launch_bounds__(256) markLeavesKernel(const u32* __restrict__ dag, const u64* __restrict__ cellStart, const u64* __restrict__ nodes, u32 n,
		u32* __restrict__ leafBits) {
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const u64 e = nodes[i];
	const u32 cell = (u32)(e >> 32), s = (u32)e;
	const u64 base = cellStart[cell];
	u32 partial = dag[s] & 0xAAAAu, slot = 1;
	while (partial) {
		partial &= partial - 1;
		const u64 leaf = base + dag[s + slot++];
		atomicOr(&leafBits[leaf >> 5], 1u << (leaf & 31u));
	}
}

// Exclusive prefix sums of the population counts of `bits` (ids of the leaves by position): tile sums, one CTA over the tiles, tile scans.
constexpr u32 kIndexTile = 1024;
__global__ void __launch_bounds__(256) tileCountKernel(const u32* __restrict__ bits, u64 words, u32* __restrict__ tileSum) {
	const u64 base = (u64)blockIdx.x * kIndexTile;
	u32 local = 0;
	for (u32 i = threadIdx.x; i < kIndexTile; i += 256)
		if (base + i < words) local += __popc(bits[base + i]);
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xFFFFFFFFu, local, d);
	__shared__ u32 s[8];
	if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = local;
	__syncthreads();
	if (threadIdx.x == 0) tileSum[blockIdx.x] = s[0] + s[1] + s[2] + s[3] + s[4] + s[5] + s[6] + s[7];
}
__global__ void __launch_bounds__(1024) scanTilesKernel(u32* __restrict__ tileSum, u32 tiles, u32* __restrict__ total) {
	__shared__ u32 sWarp[32];
	__shared__ u32 carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (u32 base = 0; base < tiles; base += 1024) {
		const u32 i = base + threadIdx.x;
		const u32 v = i < tiles ? tileSum[i] : 0u;
		u32 incl = v;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const u32 up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
			if ((int)(threadIdx.x & 31) >= d) incl += up;
		}
		if ((threadIdx.x & 31) == 31) sWarp[threadIdx.x >> 5] = incl;
		__syncthreads();
		u32 before = 0;
		for (u32 w = 0; w < (threadIdx.x >> 5); ++w) before += sWarp[w];
		const u32 start = carry;
		__syncthreads();
		if (i < tiles) tileSum[i] = start + before + incl - v;
		if (threadIdx.x == 1023) carry = start + before + incl;
		__syncthreads();
	}
	if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(256) tileScanKernel(const u32* __restrict__ bits, u64 words, const u32* __restrict__ tileSum, u32* __restrict__ prefix) {
	// a warp owns 128 consecutive words of the tile
	const u64 base = (u64)blockIdx.x * kIndexTile;
	__shared__ u32 sWarp[8];
	const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	u32 c[4], mine = 0;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const u64 i = base + warp * 128 + lane * 4 + k;
		c[k] = i < words ? __popc(bits[i]) : 0u;
		mine += c[k];
	}
	u32 incl = mine;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const u32 up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
		if ((int)lane >= d) incl += up;
	}
	if (lane == 31) sWarp[warp] = incl;
	__syncthreads();
	u32 run = tileSum[blockIdx.x] + incl - mine;
	for (u32 w = 0; w < warp; ++w) run += sWarp[w];
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const u64 i = base + warp * 128 + lane * 4 + k;
		if (i < words) prefix[i] = run;
		run += c[k];
	}
}

// Rank of the set bit at `position` (ids follow the order of the words: the DAG's own order).
__device__ __forceinline__ u32 leafIdAt(const u32* __restrict__ bits, const u32* __restrict__ prefix, u64 position) {
	return prefix[position >> 5] + __popc(bits[position >> 5] & ((1u << (position & 31u)) - 1u));
}

__global__ void __launch_bounds__(256) rootIdsKernel(const u64* __restrict__ roots, u32 n, const u32* __restrict__ nodeBits, const u32* __restrict__ nodePrefix,
		u32* __restrict__ ids) {
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) ids[i] = leafIdAt(nodeBits, nodePrefix, (u32)roots[i]);
}

// Every inner node -> eight slots. nodes: all inner nodes, the last `numAboveLeaves` of them are the ones whose children are leaves.
__global__ void __launch_bounds__(256) writeSlotsKernel(const u32* __restrict__ dag, const u64* __restrict__ cellStart, const u64* __restrict__ nodes, u32 n,
		u32 firstAboveLeaves, const u32* __restrict__ nodeBits, const u32* __restrict__ nodePrefix, const u32* __restrict__ leafBits,
		const u32* __restrict__ leafPrefix, u32* __restrict__ slots) {
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const u64 e = nodes[i];
	const u32 cell = (u32)(e >> 32), s = (u32)e;
	const u64 base = cellStart[cell];
	const bool toLeaves = i >= firstAboveLeaves;
	const u32 mask = dag[s];
	u32 out[8];
	u32 slot = 1;
#pragma unroll
	for (u32 c = 0; c < 8; ++c) {
		const u32 code = (mask >> (2 * c)) & 3u;
		out[c] = code & 1u;
		if (code == 2u) {
			const u64 child = base + dag[s + slot++];
			out[c] = 2u + (toLeaves ? leafIdAt(leafBits, leafPrefix, child) : leafIdAt(nodeBits, nodePrefix, child));
		}
	}
	uint4* dst = reinterpret_cast<uint4*>(slots + (u64)leafIdAt(nodeBits, nodePrefix, s) * 8);
	dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
	dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
}

// One bit per nibble: bit x of `row` -> bit 4x.
__device__ __forceinline__ u32 spreadToNibbles(u32 row) {
	u32 s = row & 0xFFu;
	s = (s | (s << 12)) & 0x000F000Fu;
	s = (s | (s << 6)) & 0x03030303u;
	s = (s | (s << 3)) & 0x11111111u;
	return s;
}

// Every distinct leaf -> its k-code. One thread per word of the leaf-start bitmap (leaves are several words apart: at most a few
// bits per word). *notNested is set if some leaf's slices are not nested (then the codes do not describe it).
__global__ void __launch_bounds__(256) leafCodesKernel(const u32* __restrict__ dag, const u32* __restrict__ leafBits, const u32* __restrict__ prefix, u64 words,
		u32* __restrict__ codes, u32* __restrict__ notNested) {
	const u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= words) return;
	u32 bits = leafBits[w];
	u32 id = prefix[w];
	while (bits) {
		const u32 pos = __ffs(bits) - 1;
		bits &= bits - 1;
		const u32* leaf = dag + w * 32 + pos;
		const u32 mask = leaf[0];
		u32 code[8] = {0, 0, 0, 0, 0, 0, 0, 0};
		u32 prevLo = 0xFFFFFFFFu, prevHi = 0xFFFFFFFFu, payload = 1;
		bool nested = true;
#pragma unroll
		for (u32 z = 0; z < 8; ++z) {
			const u32 c = (mask >> (2 * z)) & 3u;
			u32 lo = 0, hi = 0;
			if (c == 1u) {
				lo = hi = 0xFFFFFFFFu;
			} else if (c == 2u) {
				lo = leaf[payload];
				hi = leaf[payload + 1];
				payload += 2;
			}
			nested = nested && !(lo & ~prevLo) && !(hi & ~prevHi);
			prevLo = lo;
			prevHi = hi;
#pragma unroll
			for (u32 y = 0; y < 4; ++y) {
				code[y] += spreadToNibbles(lo >> (8 * y));
				code[4 + y] += spreadToNibbles(hi >> (8 * y));
			}
		}
		if (!nested) atomicExch(notNested, 1u);
		uint4* dst = reinterpret_cast<uint4*>(codes + (u64)id * 8);
		dst[0] = make_uint4(code[0], code[1], code[2], code[3]);
		dst[1] = make_uint4(code[4], code[5], code[6], code[7]);
		++id;
	}
}

template <typename T>
struct DeviceBuffer {  // stream-ordered scratch, released with the object
	T* p = nullptr;
	cudaStream_t st;
	explicit DeviceBuffer(cudaStream_t s) : st(s) {}
	~DeviceBuffer() {
		if (p) cudaFreeAsync(p, st);
	}
	cudaError_t alloc(u64 count) { return cudaMallocAsync(reinterpret_cast<void**>(&p), (count ? count : 1) * sizeof(T), st); }
};

}  // namespace

void freeLookupIndex(cpvs_ctx* ctx, LookupIndex* ix) {
	for (u32** p : {&ix->nodes, &ix->grid, &ix->codes, &ix->skip})
		if (*p) {
			cudaFreeAsync(*p, ctx->stream);
			*p = nullptr;
		}
	ix->valid = false;
	ix->skipLevels = 0;
}

namespace {
int prefixOfBits(cpvs_ctx* ctx, const u32* bits, u64 bitWords, u32* tileSum, u32* prefix, u32* total) {
	const u32 tiles = (u32)((bitWords + kIndexTile - 1) / kIndexTile);
	tileCountKernel<<<tiles, 256, 0, ctx->stream>>>(bits, bitWords, tileSum);
	scanTilesKernel<<<1, 1024, 0, ctx->stream>>>(tileSum, tiles, total);
	tileScanKernel<<<tiles, 256, 0, ctx->stream>>>(bits, bitWords, tileSum, prefix);
	ctx->launches += 3;
	return CPVS_OK;
}
}  // namespace

// cellStart[i] / cellWords[i]: first word and length of cell i's DAG inside `dag` (one cell at 0 for a single shadow);
// hostGrid: the container's cell table (offset or sentinel per cell), empty for a single shadow.
int buildLookupIndex(cpvs_ctx* ctx, const u32* dag, u64 dagWords, const std::vector<u64>& cellStart, const std::vector<u64>& cellWords,
		const std::vector<u32>& hostGrid, u32 dagLevels, u32 gridLevels, LookupIndex* out) {
	out->tried = true;
	out->valid = false;
	(void)gridLevels;
	if (dagLevels < 5 || dagWords >= (1ull << 32)) return CPVS_OK;  // no leaf level (or ids would not fit): lookups walk the wire format
	CPVS_CUDA(cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	const u32 numCells = (u32)cellStart.size();
	const u64 bitWords = (dagWords + 31) / 32;
	// roots of the cells with a DAG (a uniform cell is answered by its grid sentinel)
	std::vector<u64> roots;
	for (u32 c = 0; c < numCells; ++c) {
		const bool uniform = !hostGrid.empty() && (hostGrid[c] == CPVS_GRID_CELL_SHADOWED || hostGrid[c] == CPVS_GRID_CELL_VISIBLE);
		if (!uniform) roots.push_back(((u64)c << 32) | cellStart[c]);
	}
	if (roots.empty()) return CPVS_OK;
	DeviceBuffer<u64> dCellStart(st), nodes(st);
	DeviceBuffer<u32> nodeBits(st), leafBits(st), nodePrefix(st), leafPrefix(st), tileSum(st), counters(st);
	const u32 tiles = (u32)((bitWords + kIndexTile - 1) / kIndexTile);
	// the nodes above the leaves take about five words each and point to leaves of ten and more: an eighth of the words bounds the
	// number of inner nodes in practice (a DAG that breaks the bound gets no index)
	const u32 capacity = (u32)std::min<u64>(dagWords / 8 + roots.size() + 4096, 0xFFFFFFF0ull);
	cudaError_t e = dCellStart.alloc(numCells);
	if (e == cudaSuccess) e = nodes.alloc(capacity);
	if (e == cudaSuccess) e = nodeBits.alloc(bitWords);
	if (e == cudaSuccess) e = leafBits.alloc(bitWords);
	if (e == cudaSuccess) e = nodePrefix.alloc(bitWords);
	if (e == cudaSuccess) e = leafPrefix.alloc(bitWords);
	if (e == cudaSuccess) e = tileSum.alloc(tiles);
	if (e == cudaSuccess) e = counters.alloc(8);  // [0] nodes so far, [1] error, [2] inner nodes, [3] leaves, [4] not nested
	if (e != cudaSuccess) {
		cudaGetLastError();
		return CPVS_OK;  // not enough memory for the scratch: no index, lookups still work
	}
	CPVS_CUDA(cudaMemcpyAsync(dCellStart.p, cellStart.data(), numCells * sizeof(u64), cudaMemcpyHostToDevice, st));
	CPVS_CUDA(cudaMemcpyAsync(nodes.p, roots.data(), roots.size() * sizeof(u64), cudaMemcpyHostToDevice, st));
	CPVS_CUDA(cudaMemsetAsync(nodeBits.p, 0, bitWords * sizeof(u32), st));
	CPVS_CUDA(cudaMemsetAsync(leafBits.p, 0, bitWords * sizeof(u32), st));
	CPVS_CUDA(cudaMemsetAsync(counters.p, 0, 8 * sizeof(u32), st));
	// breadth first from the roots (level dagLevels-2) down to the nodes above the leaves (level 3); all inner nodes end up in
	// `nodes`, level after level
	u32 levelStart = 0, levelCount = (u32)roots.size(), total = levelCount;
	u32 hCount[8] = {0};
	{
		const u32 init[1] = {total};
		CPVS_CUDA(cudaMemcpyAsync(counters.p, init, sizeof(u32), cudaMemcpyHostToDevice, st));
	}
	markRootsKernel<<<(levelCount + 255) / 256, 256, 0, st>>>(nodes.p, levelCount, nodeBits.p);
	++ctx->launches;
	for (int level = (int)dagLevels - 2; level > 3 && levelCount; --level) {
		expandFrontierKernel<<<(levelCount + 255) / 256, 256, 0, st>>>(dag, dCellStart.p, nodes.p + levelStart, levelCount, nodeBits.p, nodes.p, counters.p,
				capacity, counters.p + 1);
		++ctx->launches;
		CPVS_CUDA(cudaMemcpyAsync(hCount, counters.p, 2 * sizeof(u32), cudaMemcpyDeviceToHost, st));
		CPVS_CUDA(cudaStreamSynchronize(st));
		if (hCount[1]) return CPVS_OK;  // more inner nodes than the bound: no index
		levelStart = total;
		levelCount = hCount[0] - total;
		total = hCount[0];
	}
	if (!levelCount) return CPVS_OK;
	const u32 firstAboveLeaves = levelStart, numInner = total;
	markLeavesKernel<<<(levelCount + 255) / 256, 256, 0, st>>>(dag, dCellStart.p, nodes.p + levelStart, levelCount, leafBits.p);
	++ctx->launches;
	prefixOfBits(ctx, nodeBits.p, bitWords, tileSum.p, nodePrefix.p, counters.p + 2);
	prefixOfBits(ctx, leafBits.p, bitWords, tileSum.p, leafPrefix.p, counters.p + 3);
	CPVS_CUDA(cudaMemcpyAsync(hCount, counters.p, 8 * sizeof(u32), cudaMemcpyDeviceToHost, st));
	CPVS_CUDA(cudaStreamSynchronize(st));
	const u32 numLeaves = hCount[3];
	if (!numLeaves || hCount[2] != numInner) return CPVS_OK;
	e = cudaMallocAsync(reinterpret_cast<void**>(&out->nodes), (u64)numInner * 8 * sizeof(u32), st);
	if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void**>(&out->codes), (u64)numLeaves * 8 * sizeof(u32), st);
	if (e == cudaSuccess && !hostGrid.empty()) e = cudaMallocAsync(reinterpret_cast<void**>(&out->grid), hostGrid.size() * sizeof(u32), st);
	if (e != cudaSuccess) {
		cudaGetLastError();
		freeLookupIndex(ctx, out);
		return CPVS_OK;
	}
	writeSlotsKernel<<<(numInner + 255) / 256, 256, 0, st>>>(dag, dCellStart.p, nodes.p, numInner, firstAboveLeaves, nodeBits.p, nodePrefix.p, leafBits.p,
			leafPrefix.p, out->nodes);
	leafCodesKernel<<<(unsigned)((bitWords + 255) / 256), 256, 0, st>>>(dag, leafBits.p, leafPrefix.p, bitWords, out->codes, counters.p + 4);
	ctx->launches += 2;
	if (!hostGrid.empty()) {  // the cell table with root node ids: the roots are the first nodes of the list, in cell order
		std::vector<u32> newGrid(hostGrid.size());
		std::vector<u32> rootIds(roots.size());
		CPVS_CUDA(cudaStreamSynchronize(st));
		// a root's id is its rank among all node starts: read the ranks back (one word per cell with a DAG)
		DeviceBuffer<u32> dIds(st);
		CPVS_CUDA(dIds.alloc(roots.size()));
		rootIdsKernel<<<((u32)roots.size() + 255) / 256, 256, 0, st>>>(nodes.p, (u32)roots.size(), nodeBits.p, nodePrefix.p, dIds.p);
		++ctx->launches;
		CPVS_CUDA(cudaMemcpyAsync(rootIds.data(), dIds.p, roots.size() * sizeof(u32), cudaMemcpyDeviceToHost, st));
		CPVS_CUDA(cudaStreamSynchronize(st));
		size_t r = 0;
		for (size_t c = 0; c < hostGrid.size(); ++c) {
			const bool uniform = hostGrid[c] == CPVS_GRID_CELL_SHADOWED || hostGrid[c] == CPVS_GRID_CELL_VISIBLE;
			newGrid[c] = uniform ? hostGrid[c] : rootIds[r++];
		}
		CPVS_CUDA(cudaMemcpyAsync(out->grid, newGrid.data(), newGrid.size() * sizeof(u32), cudaMemcpyHostToDevice, st));
		CPVS_CUDA(cudaStreamSynchronize(st));
	}
	CPVS_CUDA(cudaMemcpyAsync(hCount, counters.p, 8 * sizeof(u32), cudaMemcpyDeviceToHost, st));
	CPVS_CUDA(cudaStreamSynchronize(st));
	CPVS_CUDA(cudaGetLastError());
	if (hCount[4] || numInner >= 0x0FFFFFF0u) {  // slices that are not nested: these words did not come from a depth map
		freeLookupIndex(ctx, out);
		return CPVS_OK;
	}
	out->numNodes = numInner;
	out->numLeaves = numLeaves;
	out->valid = true;
	return CPVS_OK;
}

}  // namespace cpvs
