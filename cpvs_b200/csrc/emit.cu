// K6 -- pointer compression and final layout (reference CompressedShadow::compress,
// src/CompressedShadow.cpp:314-392).
//
// Final format (the lookup's wire contract): levels root first; inside a level the unique nodes in
// first-occurrence order; inner node = mask + one absolute word offset per PARTIAL child, leaf =
// mask + (lo32, hi32) per PARTIAL slice. The per-level word offsets were prefix-summed while
// merging; here level bases are chained top-down and every unique node is written exactly once.
#include "kernels.h"
#include "leafbits.cuh"

namespace cpvs {

namespace {

__global__ void levelBasesKernel(const u64* __restrict__ words, u64* __restrict__ bases, int topLevel, int minLevel, u64* totalWords,
		u64 capacity, u32* overflow, const u16* __restrict__ rootMask, u32* rootWord) {
	u64 running = 0;
	for (int level = topLevel; level >= minLevel; --level) {
		bases[level] = running;
		running += words[level];
	}
	*totalWords = running;
	*rootWord = rootMask[0];  // the DAG's first word: the root is never merged away
	if (running > capacity) atomicOr(overflow, kOverflowWords);
}

// Both emit kernels give one thread one unique node; the 256 nodes of a block occupy one contiguous run
// of the DAG, which is assembled in shared memory and written out with fully coalesced stores (a
// thread-per-node store of 1+k words would touch a different 32-byte sector per thread and word).
// Blocks are walked with a grid stride, so a grid sized from an estimate of the unique count is always correct.
constexpr int kEmitThreads = 256;

// First word of the level inside the allocation (see EmitLevelArgs), or NULL if the DAG does not fit.
__device__ __forceinline__ u32* levelStart(const EmitLevelArgs& a) {
	if (*a.overflow & kOverflowNodes) return nullptr;  // incomplete level arrays: the host rebuilds
	if (a.fromEnd) {
		const u64 words = *a.wordCount;
		return words <= a.capacity ? a.dagAlloc + (a.capacity - words) : nullptr;
	}
	const u64 total = *a.totalWords;
	return total <= a.capacity ? a.dagAlloc + (a.capacity - total) + *a.levelBase : nullptr;
}

__device__ __forceinline__ void emitRun(u32* __restrict__ level, const u32* sOut, u64 runStart, u32 runWords) {
	u32* out = level + runStart;
	for (u32 i = threadIdx.x; i < runWords; i += kEmitThreads) out[i] = sOut[i];
}

// The (up to) eight group ids of a node's PARTIAL children are loaded together, then the eight word offsets, then the
// stores: two round trips instead of 2k dependent ones.
__device__ __forceinline__ void emitInnerBlocks(const EmitLevelArgs& a, u32 firstBlock, u32 blockStride, u32* sOut) {
	const u64 unique = *a.uniqueCount;
	u32* level = levelStart(a);
	if (!level) return;
	for (u64 r0 = (u64)firstBlock * kEmitThreads; r0 < unique; r0 += (u64)blockStride * kEmitThreads) {
		const u64 r = r0 + threadIdx.x;
		const u32 runStart = a.wordOffset[r0];
		const u32 runEnd = (r0 + kEmitThreads < unique) ? a.wordOffset[r0 + kEmitThreads] : (u32)*a.wordCount;
		if (r < unique) {
			const u32 j = a.firstList[r];
			const u32 mask = a.masks[j];
			const u32 k = __popc(mask & 0xAAAAu);
			u32* out = sOut + (a.wordOffset[r] - runStart);
			out[0] = mask;
			if (k) {
				const u32* kids = a.childUid + a.firstChild[j];
				const u32 childBase = (u32)*a.childLevelBase;
				u32 kid[8], off[8];
#pragma unroll
				for (u32 c = 0; c < 8; ++c) kid[c] = c < k ? kids[c] & 0x7FFFFFFFu : 0u;
#pragma unroll
				for (u32 c = 0; c < 8; ++c) off[c] = c < k ? a.childSlotOffset[kid[c]] : 0u;
#pragma unroll
				for (u32 c = 0; c < 8; ++c)
					if (c < k) out[1 + c] = childBase + off[c];
			}
		}
		__syncthreads();
		emitRun(level, sOut, runStart, runEnd - runStart);
		__syncthreads();
	}
}

__global__ void __launch_bounds__(kEmitThreads) emitInnerKernel(EmitLevelArgs a) {
	__shared__ u32 sOut[kEmitThreads * 9];
	emitInnerBlocks(a, blockIdx.x, gridDim.x, sOut);
}

// All inner levels in one launch: blockStart[] maps a block to its level.
__global__ void __launch_bounds__(kEmitThreads) emitInnerLevelsKernel(EmitMultiArgs m) {
	__shared__ u32 sOut[kEmitThreads * 9];
	int s = 0;
	while (s + 1 < m.count && blockIdx.x >= m.blockStart[s + 1]) ++s;
	emitInnerBlocks(m.lv[s], blockIdx.x - m.blockStart[s], m.blockStart[s + 1] - m.blockStart[s], sOut);
}

// Leaves: expands the k-code (nibble x of word y = lit slices of texel (x,y)) into the 64-bit masks of
// the PARTIAL slices, bit x + 8y = lit (createLeafmask, src/CompressedShadowUtil.cpp:59-78) -- through bit planes
// (leafbits.cuh): the code is transposed once, every slice is then one or two logic instructions per half, and the loop over
// the slices is unrolled with predicated stores. About 330 executed instructions per terrain leaf (nearly all of its slices
// PARTIAL) against 850 for row-by-row nibble compares: 0.200 -> 0.174 ms at 16K^2 terrain (profiles/r1_switch_probe.md).

__global__ void __launch_bounds__(kEmitThreads) emitLeavesKernel(EmitLevelArgs a) {
	__shared__ u32 sOut[kEmitThreads * 17];
	const u64 unique = *a.uniqueCount;
	u32* level = levelStart(a);
	if (!level) return;
	for (u64 r0 = (u64)blockIdx.x * kEmitThreads; r0 < unique; r0 += (u64)gridDim.x * kEmitThreads) {
		const u64 r = r0 + threadIdx.x;
		const u32 runStart = a.wordOffset[r0];
		const u32 runEnd = (r0 + kEmitThreads < unique) ? a.wordOffset[r0 + kEmitThreads] : (u32)*a.wordCount;
		if (r < unique) {
			const u32 j = a.firstList[r];
			const u32 mask = a.masks[j];
			u32* out = sOut + (a.wordOffset[r] - runStart);
			*out++ = mask;
			const uint4* src = reinterpret_cast<const uint4*>(a.leafCodes + (u64)j * 8);
			const uint4 c0 = src[0], c1 = src[1];
			if (mask & 0xAAAAu) {
				const u32 code[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
				u32 lo[4], hi[4];
				codeToPlanes(code, lo, hi);
#pragma unroll
				for (u32 slice = 0; slice < 8; ++slice) {
					if (mask & (2u << (2u * slice))) {  // PARTIAL slices only, lowest first
						out[0] = sliceFromPlanes(lo, slice);
						out[1] = sliceFromPlanes(hi, slice);
						out += 2;
					}
				}
			}
		}
		__syncthreads();
		emitRun(level, sOut, runStart, runEnd - runStart);
		__syncthreads();
	}
}

}  // namespace

namespace {
// A finished DAG out of its staging buffer: dst and src have the same alignment within 16 bytes (the caller offsets dst).
__global__ void __launch_bounds__(256) copyWordsKernel(u32* __restrict__ dst, const u32* __restrict__ src, u64 words) {
	const u64 head = ((16u - (reinterpret_cast<uintptr_t>(src) & 15u)) & 15u) >> 2;  // words before the first 16-byte boundary
	const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x, threads = (u64)gridDim.x * blockDim.x;
	if (words <= head) {
		if (tid < words) dst[tid] = src[tid];
		return;
	}
	const u64 quads = (words - head) >> 2;
	const uint4* s4 = reinterpret_cast<const uint4*>(src + head);
	uint4* d4 = reinterpret_cast<uint4*>(dst + head);
	for (u64 i = tid; i < quads; i += threads) d4[i] = __ldcs(s4 + i);
	const u64 tail = head + (quads << 2);
	if (tid < head) dst[tid] = src[tid];
	if (tid < words - tail) dst[tail + tid] = src[tail + tid];
}

}  // namespace

int launchCopyWords(u32* dst, const u32* src, u64 words, cudaStream_t stream) {
	if (!words) return 0;
	const u64 want = (words / 4 + 255) / 256;
	const unsigned blocks = (unsigned)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
	copyWordsKernel<<<blocks, 256, 0, stream>>>(dst, src, words);
	return 1;
}

int launchLevelBases(const u64* words, u64* bases, int topLevel, int minLevel, u64* totalWords, u64 capacity, u32* overflow, const u16* rootMask,
		u32* rootWord, cudaStream_t stream) {
	levelBasesKernel<<<1, 1, 0, stream>>>(words, bases, topLevel, minLevel, totalWords, capacity, overflow, rootMask, rootWord);
	return 1;
}

int launchEmitInnerLevels(EmitMultiArgs& m, cudaStream_t stream) {
	if (m.count <= 0) return 0;
	u32 blocks = 0;
	for (int s = 0; s < m.count; ++s) {
		m.blockStart[s] = blocks;
		const u64 want = (m.lv[s].n + kEmitThreads - 1) / kEmitThreads;
		blocks += (u32)(want ? want : 1);
	}
	m.blockStart[m.count] = blocks;
	emitInnerLevelsKernel<<<blocks, kEmitThreads, 0, stream>>>(m);
	return 1;
}

int launchEmitLevel(const EmitLevelArgs& a, cudaStream_t stream) {
	const u64 want = (a.n + 255) / 256;
	const unsigned blocks = (unsigned)(want ? want : 1);
	if (a.leaf)
		emitLeavesKernel<<<blocks, 256, 0, stream>>>(a);
	else
		emitInnerKernel<<<blocks, 256, 0, stream>>>(a);
	return 1;
}

}  // namespace cpvs
