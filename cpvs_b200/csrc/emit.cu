// K6 -- pointer compression and final layout (reference CompressedShadow::compress,
// src/CompressedShadow.cpp:314-392).
//
// Final format (the lookup's wire contract): levels root first; inside a level the unique nodes in
// first-occurrence order; inner node = mask + one absolute word offset per PARTIAL child, leaf =
// mask + (lo32, hi32) per PARTIAL slice. The per-level word offsets were prefix-summed while
// merging; here level bases are chained top-down and every unique node is written exactly once.
#include "kernels.h"

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

__global__ void __launch_bounds__(256) emitInnerKernel(EmitLevelArgs a) {
	const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= *a.uniqueCount) return;
	const u32 j = a.firstList[r];
	const u32 mask = a.masks[j];
	const u32 k = __popc(mask & 0xAAAAu);
	u32* out = a.dag + (*a.levelBase + a.wordOffset[r]);
	out[0] = mask;
	if (k) {
		const u32* kids = a.childUid + a.firstChild[j];
		const u32 childBase = (u32)*a.childLevelBase;
		for (u32 c = 0; c < k; ++c) out[1 + c] = childBase + a.childWordOffset[kids[c]];
	}
}

// Two lanes... no: one thread per unique leaf; 64 B read, <= 68 B written.
__global__ void __launch_bounds__(256) emitLeavesKernel(EmitLevelArgs a) {
	const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= *a.uniqueCount) return;
	const u32 j = a.firstList[r];
	const u32 mask = a.masks[j];
	u32* out = a.dag + (*a.levelBase + a.wordOffset[r]);
	*out++ = mask;
	const ulonglong2* src = reinterpret_cast<const ulonglong2*>(a.leafBits + (u64)j * 8);
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const ulonglong2 v = src[i];
		if ((mask >> (4 * i)) & 2u) {
			*out++ = (u32)v.x;
			*out++ = (u32)(v.x >> 32);
		}
		if ((mask >> (4 * i + 2)) & 2u) {
			*out++ = (u32)v.y;
			*out++ = (u32)(v.y >> 32);
		}
	}
}

}  // namespace

int launchLevelBases(const u64* words, u64* bases, int topLevel, int minLevel, u64* totalWords, cudaStream_t stream) {
	levelBasesKernel<<<1, 1, 0, stream>>>(words, bases, topLevel, minLevel, totalWords);
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
