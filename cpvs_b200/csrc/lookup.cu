// K7 -- shadow lookup (reference CompressedShadow::traverse, src/CompressedShadow.cpp:404-463, and the
// GLSL compute shader shader/traverse.cs:41-149 it mirrors, including the top-level grid step).
//
// One thread per query. The path is the query's integer voxel coordinate; each level consumes one bit
// per axis, tests the 2-bit child code and follows the popcount-ranked pointer. Deviations from the
// reference, both documented in SURVEY.md: paths are clamped to the volume (N5) and the grid
// sentinels are the ones the C++ side writes (N3).
#include "kernels.h"

namespace cpvs {

namespace {

constexpr u32 kCellShadowed = 0x0FFFFFFFu;  // src/CompressedShadowContainer.cpp:8
constexpr u32 kCellVisible = 0x0FFFFFFEu;   // src/CompressedShadowContainer.cpp:9

// cs::getPathFromNDC (src/CompressedShadowUtil.h:70-75) / traverse.cs:43-48, clamped.
__device__ __forceinline__ int pathCoord(float ndc, int resolution) {
	float f = __fadd_rn(ndc, 1.0f);
	f = __fmul_rn(f, 0.5f);
	f = __fmul_rn(f, __int2float_rn(resolution));
	if (!(f > 0.0f)) return 0;  // also NaN
	if (f >= __int2float_rn(resolution)) return resolution;
	return min(__float2int_rz(f), resolution);
}

// getChildOffset (src/CompressedShadow.cpp:394-402): rank among PARTIAL children; childBits = 2*index.
__device__ __forceinline__ u32 childRank(u32 mask, u32 childBits) { return __popc(mask & (0xAAAAu >> (16u - childBits))); }

// Descends from the node at `offset`, whose children are picked by path bit `startLevel`.
__device__ __forceinline__ u32 descend(const u32* __restrict__ dag, u32 offset, int startLevel, bool leaf, int px, int py, int pz) {
	const int minLevel = leaf ? 3 : 0;
	for (int level = startLevel; level >= minLevel; --level) {
		const u32 idx = ((px >> level) & 1) | (((py >> level) & 1) << 1) | (((pz >> level) & 1) << 2);
		const u32 mask = __ldg(dag + offset);
		const u32 vis = (mask >> (idx * 2)) & 3u;
		if (vis != 2u) return vis & 1u;
		offset = __ldg(dag + offset + 1 + childRank(mask, idx * 2));
	}
	if (!leaf) return 2u;
	const u32 idx = pz & 7;
	const u32 mask = __ldg(dag + offset);
	const u32 vis = (mask >> (idx * 2)) & 3u;
	if (vis != 2u) return vis & 1u;
	const u32 bit = (px & 7) + 8 * (py & 7);
	const u32 word = __ldg(dag + offset + 1 + childRank(mask, idx * 2) * 2 + (bit >> 5));
	return (word >> (bit & 31)) & 1u;
}

__device__ __forceinline__ u32 lookupOne(const LookupDag& d, float x, float y, float z) {
	const int resolution = (1 << (d.dagLevels + d.gridLevels - 1)) - 1;
	const int px = pathCoord(x, resolution), py = pathCoord(y, resolution), pz = pathCoord(z, resolution);
	const u32* dag = d.dag;
	if (d.grid) {  // traverse.cs:78-88
		const u32 shift = d.dagLevels - 1, res = 1u << d.gridLevels;
		const u32 cell = __ldg(d.grid + ((u32)(pz >> shift) * res + (u32)(py >> shift)) * res + (u32)(px >> shift));
		if (cell == kCellShadowed) return 0u;
		if (cell == kCellVisible) return 1u;
		dag += cell;
	}
	int startLevel = (int)d.dagLevels - 2;
	u32 offset = 0;
	if (d.skip) {  // the first skipLevels steps of the descent, precomputed per cell
		const u32 shift = d.dagLevels - 1 - d.skipLevels, res = 1u << (d.gridLevels + d.skipLevels);
		const u32 entry = __ldg(d.skip + ((u32)(pz >> shift) * res + (u32)(py >> shift)) * res + (u32)(px >> shift));
		if (entry == kSkipShadow) return 0u;
		if (entry == kSkipVisible) return 1u;
		offset = entry;
		startLevel -= (int)d.skipLevels;
	}
	return descend(dag, offset, startLevel, d.leafmasks != 0, px, py, pz);
}

// One thread per shortcut cell: runs the first skipLevels steps of the descent for the cell's path prefix.
__global__ void __launch_bounds__(256) buildSkipGridKernel(LookupDag d, u32* __restrict__ skip) {
	const u32 bits = d.gridLevels + d.skipLevels, res = 1u << bits;
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= res * res * res) return;
	const u32 cx = i & (res - 1), cy = (i >> bits) & (res - 1), cz = i >> (2 * bits);
	const u32* dag = d.dag;
	if (d.grid) {
		const u32 gres = 1u << d.gridLevels;
		const u32 cell = d.grid[((cz >> d.skipLevels) * gres + (cy >> d.skipLevels)) * gres + (cx >> d.skipLevels)];
		if (cell == kCellShadowed || cell == kCellVisible) {
			skip[i] = cell == kCellShadowed ? kSkipShadow : kSkipVisible;
			return;
		}
		dag += cell;
	}
	u32 offset = 0;
	for (int step = (int)d.skipLevels - 1; step >= 0; --step) {  // path bit dagLevels-2-k of the voxel = bit `step` of the cell
		const u32 idx = ((cx >> step) & 1u) | (((cy >> step) & 1u) << 1) | (((cz >> step) & 1u) << 2);
		const u32 mask = dag[offset];
		const u32 vis = (mask >> (idx * 2)) & 3u;
		if (vis != 2u) {
			skip[i] = vis ? kSkipVisible : kSkipShadow;
			return;
		}
		offset = dag[offset + 1 + childRank(mask, idx * 2)];
	}
	skip[i] = offset;
}

__global__ void __launch_bounds__(256) lookupNdcKernel(LookupDag d, const float* __restrict__ ndc, long long count, unsigned char* __restrict__ out) {
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	out[i] = (unsigned char)lookupOne(d, ndc[3 * i], ndc[3 * i + 1], ndc[3 * i + 2]);
}

struct Mat4 {
	float m[16];
};

// traverse.cs main() (:135-149) with glm's evaluation order for mat4*vec4 and the divide by w.
// 8x4-pixel warps (blockDim 8x32): neighbouring pixels share most of their path through the DAG, so a
// compact footprint per warp means fewer distinct nodes per load instruction than a 32x1 strip.
__global__ void __launch_bounds__(256) evaluateKernel(LookupDag d, const float4* __restrict__ pos, u32 width, u32 height, Mat4 mat,
		unsigned char* __restrict__ out) {
	const u32 x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= width || y >= height) return;
	const size_t i = (size_t)y * width + x;
	const float4 p = pos[i];
	float v[4];
#pragma unroll
	for (int r = 0; r < 4; ++r)
		v[r] = __fadd_rn(__fadd_rn(__fmul_rn(mat.m[r], p.x), __fmul_rn(mat.m[4 + r], p.y)),
				__fadd_rn(__fmul_rn(mat.m[8 + r], p.z), mat.m[12 + r]));
	const u32 vis = lookupOne(d, __fdiv_rn(v[0], v[3]), __fdiv_rn(v[1], v[3]), __fdiv_rn(v[2], v[3]));
	out[i] = vis == 1u ? 255 : 0;
}

}  // namespace

int launchBuildSkipGrid(const LookupDag& d, u32* skip, cudaStream_t stream) {
	const u32 bits = 3 * (d.gridLevels + d.skipLevels);
	const u32 cells = 1u << bits;
	buildSkipGridKernel<<<(cells + 255) / 256, 256, 0, stream>>>(d, skip);
	return 1;
}

int launchLookupNdc(const LookupDag& d, const float* ndc, long long count, unsigned char* out, cudaStream_t stream) {
	if (count <= 0) return 0;
	lookupNdcKernel<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(d, ndc, count, out);
	return 1;
}

int launchEvaluate(const LookupDag& d, const float* positions, unsigned width, unsigned height, const float* matrix, unsigned char* out,
		cudaStream_t stream) {
	if (!width || !height) return 0;
	Mat4 m;
	for (int i = 0; i < 16; ++i) m.m[i] = matrix[i];
	const dim3 block(8, 32), grid((width + 7) / 8, (height + 31) / 32);
	evaluateKernel<<<grid, block, 0, stream>>>(d, reinterpret_cast<const float4*>(positions), width, height, m, out);
	return 1;
}

}  // namespace cpvs
