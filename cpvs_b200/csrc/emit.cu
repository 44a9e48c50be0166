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

__global__ void levelBasesKernel(const u64* __restrict__ words, u64* __restrict__ bases, int topLevel, int minLevel, u64* totalWords) {
	u64 running = 0;
	for (int level = topLevel; level >= minLevel; --level) {
		bases[level] = running;
		running += words[level];
	}
	*totalWords = running;
}

// Both emit kernels give one thread one unique node; the 256 nodes of a CTA occupy one contiguous run
// of the DAG, which is assembled in shared memory and written out with fully coalesced stores (a
// thread-per-node store of 1+k words would touch a different 32-byte sector per thread and word).
constexpr int kEmitThreads = 256;

__device__ __forceinline__ void emitRun(const EmitLevelArgs& a, const u32* sOut, u64 runStart, u32 runWords) {
	u32* out = a.dag + *a.levelBase + runStart;
	for (u32 i = threadIdx.x; i < runWords; i += kEmitThreads) out[i] = sOut[i];
}

// kGather (experimental, CPVS_EXPERIMENTS=emit-gather): the plain loop over the k PARTIAL children loads a child's group id,
// waits, loads that group's word offset, waits, stores, and only then turns to the next child -- 2k dependent round trips.
// The variant issues the (up to) eight id loads together, then the eight offset loads, then the stores: two round trips.
template <bool kGather>
__device__ __forceinline__ void emitInnerBlock(const EmitLevelArgs& a, u32 block, u32* sOut) {
	const u64 unique = *a.uniqueCount;
	const u64 r0 = (u64)block * kEmitThreads;
	if (r0 >= unique) return;
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
			if constexpr (kGather) {
				u32 kid[8], off[8];
#pragma unroll
				for (u32 c = 0; c < 8; ++c) kid[c] = c < k ? kids[c] & 0x7FFFFFFFu : 0u;
#pragma unroll
				for (u32 c = 0; c < 8; ++c) off[c] = c < k ? a.childSlotOffset[kid[c]] : 0u;
#pragma unroll
				for (u32 c = 0; c < 8; ++c)
					if (c < k) out[1 + c] = childBase + off[c];
			} else {
				for (u32 c = 0; c < k; ++c) out[1 + c] = childBase + a.childSlotOffset[kids[c] & 0x7FFFFFFFu];
			}
		}
	}
	__syncthreads();
	emitRun(a, sOut, runStart, runEnd - runStart);
}

__global__ void __launch_bounds__(kEmitThreads) emitInnerKernel(EmitLevelArgs a) {
	__shared__ u32 sOut[kEmitThreads * 9];
	emitInnerBlock<false>(a, blockIdx.x, sOut);
}

// All inner levels in one launch: blockStart[] maps a block to its level.
template <bool kGather>
__global__ void __launch_bounds__(kEmitThreads) emitInnerLevelsKernel(EmitMultiArgs m) {
	__shared__ u32 sOut[kEmitThreads * 9];
	int s = 0;
	while (s + 1 < m.count && blockIdx.x >= m.blockStart[s + 1]) ++s;
	emitInnerBlock<kGather>(m.lv[s], blockIdx.x - m.blockStart[s], sOut);
}

// Leaves: expands the k-code (nibble x of word y = lit slices of texel (x,y)) into the 64-bit masks of
// the PARTIAL slices, bit x + 8y = lit (createLeafmask, src/CompressedShadowUtil.cpp:59-78) -- through bit planes
// (leafbits.cuh): the code is transposed once, every slice is then one or two logic instructions per half, and the loop over
// the slices is unrolled with predicated stores. About 330 executed instructions per terrain leaf (nearly all of its slices
// PARTIAL) against 850 for row-by-row nibble compares: 0.200 -> 0.174 ms at 16K^2 terrain (profiles/r1_switch_probe.md).

__global__ void __launch_bounds__(kEmitThreads) emitLeavesKernel(EmitLevelArgs a) {
	__shared__ u32 sOut[kEmitThreads * 17];
	const u64 unique = *a.uniqueCount;
	const u64 r0 = (u64)blockIdx.x * kEmitThreads;
	if (r0 >= unique) return;
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
	emitRun(a, sOut, runStart, runEnd - runStart);
}

}  // namespace

int launchLevelBases(const u64* words, u64* bases, int topLevel, int minLevel, u64* totalWords, cudaStream_t stream) {
	levelBasesKernel<<<1, 1, 0, stream>>>(words, bases, topLevel, minLevel, totalWords);
	return 1;
}

int launchEmitInnerLevels(EmitMultiArgs& m, cudaStream_t stream) {
	if (m.count <= 0) return 0;
	u32 blocks = 0;
	for (int s = 0; s < m.count; ++s) {
		m.blockStart[s] = blocks;
		blocks += (u32)((m.lv[s].n + kEmitThreads - 1) / kEmitThreads);
	}
	m.blockStart[m.count] = blocks;
	if (m.gather)
		emitInnerLevelsKernel<true><<<blocks, kEmitThreads, 0, stream>>>(m);
	else
		emitInnerLevelsKernel<false><<<blocks, kEmitThreads, 0, stream>>>(m);
	return 1;
}

int launchEmitLevel(const EmitLevelArgs& a, cudaStream_t stream) {
	const unsigned blocks = (unsigned)((a.n + 255) / 256);
	if (a.leaf)
		emitLeavesKernel<<<blocks, 256, 0, stream>>>(a);
	else
		emitInnerKernel<<<blocks, 256, 0, stream>>>(a);
	return 1;
}

}  // namespace cpvs
