// K7 -- shadow lookup (reference CompressedShadow::traverse, src/CompressedShadow.cpp:404-463, and the
// GLSL compute shader shader/traverse.cs:41-149 it mirrors, including the top-level grid step).
//
// Two queries per thread (the single-tap kernels) or one (PCF taps). The path is the query's integer voxel coordinate; each level consumes one bit
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

// The same descent on the private lookup copy (lookup_index.cu): one load per level at an address the path alone decides, one
// load for the leaf -- lit <=> the slice lies below the texel's count of lit slices.
__device__ __forceinline__ u32 descendSlots(const u32* __restrict__ nodes, const u32* __restrict__ codes, u32 node, int startLevel, int px, int py, int pz) {
	for (int level = startLevel; level >= 3; --level) {
		const u32 idx = ((px >> level) & 1) | (((py >> level) & 1) << 1) | (((pz >> level) & 1) << 2);
		const u32 slot = __ldg(nodes + (size_t)node * 8u + idx);
		if (slot < 2u) return slot;
		node = slot - 2u;
	}
	const u32 row = __ldg(codes + (size_t)node * 8u + (u32)(py & 7));
	return (u32)(pz & 7) < ((row >> (4 * (px & 7))) & 15u) ? 1u : 0u;
}

// The voxel (px, py, pz) of the whole virtual volume.
__device__ __forceinline__ u32 lookupPath(const LookupDag& d, int px, int py, int pz) {
	const u32* dag = d.dag;
	u32 root = 0;  // lookup copy: the root's node id
	if (d.grid) {  // traverse.cs:78-88
		const u32 shift = d.dagLevels - 1, res = 1u << d.gridLevels;
		const u32 cell = __ldg(d.grid + ((u32)(pz >> shift) * res + (u32)(py >> shift)) * res + (u32)(px >> shift));
		if (cell == kCellShadowed) return 0u;
		if (cell == kCellVisible) return 1u;
		if (d.leafCodes)
			root = cell;
		else
			dag += cell;
	}
	int startLevel = (int)d.dagLevels - 2;
	u32 offset = root;
	if (d.skip) {  // the first skipLevels steps of the descent, precomputed per cell
		const u32 shift = d.dagLevels - 1 - d.skipLevels, res = 1u << (d.gridLevels + d.skipLevels);
		const u32 entry = __ldg(d.skip + ((u32)(pz >> shift) * res + (u32)(py >> shift)) * res + (u32)(px >> shift));
		if (entry == kSkipShadow) return 0u;
		if (entry == kSkipVisible) return 1u;
		offset = entry;
		startLevel -= (int)d.skipLevels;
	}
	if (d.leafCodes) return descendSlots(dag, d.leafCodes, offset, startLevel, px, py, pz);
	return descend(dag, offset, startLevel, d.leafmasks != 0, px, py, pz);
}

// One thread per shortcut cell: runs the first skipLevels steps of the descent for the cell's path prefix.
__global__ void __launch_bounds__(256) buildSkipGridKernel(LookupDag d, u32* __restrict__ skip) {
	const u32 bits = d.gridLevels + d.skipLevels, res = 1u << bits;
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= res * res * res) return;
	const u32 cx = i & (res - 1), cy = (i >> bits) & (res - 1), cz = i >> (2 * bits);
	const u32* dag = d.dag;
	u32 offset = 0;
	if (d.grid) {
		const u32 gres = 1u << d.gridLevels;
		const u32 cell = d.grid[((cz >> d.skipLevels) * gres + (cy >> d.skipLevels)) * gres + (cx >> d.skipLevels)];
		if (cell == kCellShadowed || cell == kCellVisible) {
			skip[i] = cell == kCellShadowed ? kSkipShadow : kSkipVisible;
			return;
		}
		if (d.leafCodes)
			offset = cell;  // lookup copy: the root's node id
		else
			dag += cell;
	}
	for (int step = (int)d.skipLevels - 1; step >= 0; --step) {  // path bit dagLevels-2-k of the voxel = bit `step` of the cell
		const u32 idx = ((cx >> step) & 1u) | (((cy >> step) & 1u) << 1) | (((cz >> step) & 1u) << 2);
		if (d.leafCodes) {
			const u32 slot = dag[(size_t)offset * 8u + idx];
			if (slot < 2u) {
				skip[i] = slot ? kSkipVisible : kSkipShadow;
				return;
			}
			offset = slot - 2u;
			continue;
		}
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

struct Mat4 {
	float m[16];
};

// traverse.cs main() (:135-149) with glm's evaluation order for mat4*vec4 and the divide by w.
// 8x4-pixel warps (blockDim 8x32): neighbouring pixels share most of their path through the DAG, so a
// compact footprint per warp means fewer distinct nodes per load instruction than a 32x1 strip.
//
// filterSize (setFilterSize, src/CompressedShadowContainer.h:71-73; `uniform int filterSize`, shader/traverse.cs:16-17 -- plumbed
// through the reference but never used by its shader): percentage-closer filtering over filterSize x filterSize voxels of the
// pixel's depth slice, centred on its voxel and clamped to the volume. 1 = the single lookup of the reference (0 / 255); larger
// sizes write round(255 * lit taps / taps). The taps of a pixel share the upper part of their paths: their loads hit L1.
struct PixelSource {  // linear rgba32f / r8 buffers, or CUDA surfaces (the G-buffer textures of the renderer mapped through CUDA-GL interop)
	const float4* pos;
	unsigned char* out;
	cudaSurfaceObject_t posSurface, outSurface;
};
template <bool kSurface>
__global__ void __launch_bounds__(256) evaluateKernel(LookupDag d, PixelSource io, u32 width, u32 height, Mat4 mat, int filterSize) {
	const u32 x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= width || y >= height) return;
	const size_t i = (size_t)y * width + x;
	float4 p;
	if (kSurface)
		p = surf2Dread<float4>(io.posSurface, (int)(x * sizeof(float4)), (int)y);
	else
		p = io.pos[i];
	float v[4];
#pragma unroll
	for (int r = 0; r < 4; ++r)
		v[r] = __fadd_rn(__fadd_rn(__fmul_rn(mat.m[r], p.x), __fmul_rn(mat.m[4 + r], p.y)),
				__fadd_rn(__fmul_rn(mat.m[8 + r], p.z), mat.m[12 + r]));
	const int resolution = (1 << (d.dagLevels + d.gridLevels - 1)) - 1;
	const int px = pathCoord(__fdiv_rn(v[0], v[3]), resolution), py = pathCoord(__fdiv_rn(v[1], v[3]), resolution),
			  pz = pathCoord(__fdiv_rn(v[2], v[3]), resolution);
	unsigned char result;
	if (filterSize <= 1) {
		result = lookupPath(d, px, py, pz) == 1u ? 255 : 0;
	} else {
		const int lo = -(filterSize / 2), hi = lo + filterSize;  // even sizes lean towards the lower texels
		u32 lit = 0;
		for (int dy = lo; dy < hi; ++dy)
			for (int dx = lo; dx < hi; ++dx)
				lit += lookupPath(d, min(max(px + dx, 0), resolution), min(max(py + dy, 0), resolution), pz) == 1u ? 1u : 0u;
		const u32 taps = (u32)(filterSize * filterSize);
		result = (unsigned char)((255u * lit + taps / 2u) / taps);
	}
	if (kSurface)
		surf2Dwrite(result, io.outSurface, (int)x, (int)y);
	else
		io.out[i] = result;
}

// Two horizontally adjacent pixels per thread, for the single-tap lookup on the lookup copy. The descent is a chain of dependent
// loads (shortcut grid, one node per level, the leaf); the two chains of a thread advance in lockstep, so every step has two
// loads in flight, and neighbouring pixels share most of their path: the second load usually hits the line the first one fetched.
struct PathState {
	u32 node;
	u32 result;  // 0 shadow, 1 lit once `done`
	bool done;
};
__device__ __forceinline__ PathState beginSlots(const LookupDag& d, bool live, int px, int py, int pz) {
	PathState st{0u, 0u, !live};
	if (!live) return st;
	if (d.grid) {  // traverse.cs:78-88
		const u32 shift = d.dagLevels - 1, res = 1u << d.gridLevels;
		const u32 cell = __ldg(d.grid + ((u32)(pz >> shift) * res + (u32)(py >> shift)) * res + (u32)(px >> shift));
		if (cell == kCellShadowed || cell == kCellVisible) {
			st.done = true;
			st.result = cell == kCellVisible ? 1u : 0u;
			return st;
		}
		st.node = cell;
	}
	if (d.skip) {
		const u32 shift = d.dagLevels - 1 - d.skipLevels, res = 1u << (d.gridLevels + d.skipLevels);
		const u32 entry = __ldg(d.skip + ((u32)(pz >> shift) * res + (u32)(py >> shift)) * res + (u32)(px >> shift));
		if (entry == kSkipShadow || entry == kSkipVisible) {
			st.done = true;
			st.result = entry == kSkipVisible ? 1u : 0u;
			return st;
		}
		st.node = entry;
	}
	return st;
}

// kQueries descents on the lookup copy in lockstep; result[q] = 0 shadow / 1 lit (dead queries: 0).
template <int kQueries>
__device__ __forceinline__ void lookupSlotsTogether(const LookupDag& d, const bool* live, const int* px, const int* py, const int* pz, u32* result) {
	PathState st[kQueries];
#pragma unroll
	for (int q = 0; q < kQueries; ++q) st[q] = beginSlots(d, live[q], px[q], py[q], pz[q]);
	const int startLevel = (int)d.dagLevels - 2 - (d.skip ? (int)d.skipLevels : 0);
	const u32* __restrict__ nodes = d.dag;
	for (int level = startLevel; level >= 3; --level) {
		bool all = true;
#pragma unroll
		for (int q = 0; q < kQueries; ++q) all = all && st[q].done;
		if (all) break;
		u32 slot[kQueries];
#pragma unroll
		for (int q = 0; q < kQueries; ++q) {
			const u32 idx = ((px[q] >> level) & 1) | (((py[q] >> level) & 1) << 1) | (((pz[q] >> level) & 1) << 2);
			slot[q] = st[q].done ? 0u : __ldg(nodes + (size_t)st[q].node * 8u + idx);
		}
#pragma unroll
		for (int q = 0; q < kQueries; ++q)
			if (!st[q].done) {
				if (slot[q] < 2u) {
					st[q].done = true;
					st[q].result = slot[q];
				} else {
					st[q].node = slot[q] - 2u;
				}
			}
	}
	u32 row[kQueries];
#pragma unroll
	for (int q = 0; q < kQueries; ++q) row[q] = st[q].done ? 0u : __ldg(d.leafCodes + (size_t)st[q].node * 8u + (u32)(py[q] & 7));
#pragma unroll
	for (int q = 0; q < kQueries; ++q)
		result[q] = st[q].done ? st[q].result : ((u32)(pz[q] & 7) < ((row[q] >> (4 * (px[q] & 7))) & 15u) ? 1u : 0u);
}

// The same on the wire format (DAGs without leafmasks, containers whose leaves are not nested): two dependent loads per level --
// mask, then the popcount-ranked pointer -- for all queries of the thread before either is used.
template <int kQueries>
__device__ __forceinline__ void lookupWireTogether(const LookupDag& d, const bool* live, const int* px, const int* py, const int* pz, u32* result) {
	u32 base[kQueries], offset[kQueries];
	bool done[kQueries];
#pragma unroll
	for (int q = 0; q < kQueries; ++q) {
		base[q] = 0u;
		offset[q] = 0u;
		done[q] = !live[q];
		result[q] = 0u;
		if (done[q]) continue;
		if (d.grid) {  // traverse.cs:78-88
			const u32 shift = d.dagLevels - 1, res = 1u << d.gridLevels;
			const u32 cell = __ldg(d.grid + ((u32)(pz[q] >> shift) * res + (u32)(py[q] >> shift)) * res + (u32)(px[q] >> shift));
			if (cell == kCellShadowed || cell == kCellVisible) {
				done[q] = true;
				result[q] = cell == kCellVisible ? 1u : 0u;
				continue;
			}
			base[q] = cell;
		}
		if (d.skip) {
			const u32 shift = d.dagLevels - 1 - d.skipLevels, res = 1u << (d.gridLevels + d.skipLevels);
			const u32 entry = __ldg(d.skip + ((u32)(pz[q] >> shift) * res + (u32)(py[q] >> shift)) * res + (u32)(px[q] >> shift));
			if (entry == kSkipShadow || entry == kSkipVisible) {
				done[q] = true;
				result[q] = entry == kSkipVisible ? 1u : 0u;
				continue;
			}
			offset[q] = entry;
		}
	}
	const bool leaf = d.leafmasks != 0;
	const int startLevel = (int)d.dagLevels - 2 - (d.skip ? (int)d.skipLevels : 0), minLevel = leaf ? 3 : 0;
	const u32* __restrict__ dag = d.dag;
	for (int level = startLevel; level >= minLevel; --level) {
		bool all = true;
#pragma unroll
		for (int q = 0; q < kQueries; ++q) all = all && done[q];
		if (all) break;
		u32 mask[kQueries], next[kQueries];
#pragma unroll
		for (int q = 0; q < kQueries; ++q) mask[q] = done[q] ? 0u : __ldg(dag + base[q] + offset[q]);
#pragma unroll
		for (int q = 0; q < kQueries; ++q) {
			next[q] = 0u;
			if (done[q]) continue;
			const u32 idx = ((px[q] >> level) & 1) | (((py[q] >> level) & 1) << 1) | (((pz[q] >> level) & 1) << 2);
			const u32 vis = (mask[q] >> (idx * 2)) & 3u;
			if (vis != 2u) {
				done[q] = true;
				result[q] = vis & 1u;
			} else {
				next[q] = base[q] + offset[q] + 1 + childRank(mask[q], idx * 2);
			}
		}
#pragma unroll
		for (int q = 0; q < kQueries; ++q)
			if (!done[q]) offset[q] = __ldg(dag + next[q]);
	}
	if (!leaf) {
#pragma unroll
		for (int q = 0; q < kQueries; ++q)
			if (!done[q]) result[q] = 2u;  // PARTIAL at the last level (src/CompressedShadow.cpp:460-462)
		return;
	}
	u32 mask[kQueries], at[kQueries];
#pragma unroll
	for (int q = 0; q < kQueries; ++q) mask[q] = done[q] ? 0u : __ldg(dag + base[q] + offset[q]);
#pragma unroll
	for (int q = 0; q < kQueries; ++q) {
		at[q] = 0u;
		if (done[q]) continue;
		const u32 idx = pz[q] & 7;
		const u32 vis = (mask[q] >> (idx * 2)) & 3u;
		if (vis != 2u) {
			done[q] = true;
			result[q] = vis & 1u;
		} else {
			const u32 bit = (px[q] & 7) + 8 * (py[q] & 7);
			at[q] = base[q] + offset[q] + 1 + childRank(mask[q], idx * 2) * 2 + (bit >> 5);
		}
	}
#pragma unroll
	for (int q = 0; q < kQueries; ++q)
		if (!done[q]) {
			const u32 bit = (px[q] & 7) + 8 * (py[q] & 7);
			result[q] = (__ldg(dag + at[q]) >> (bit & 31)) & 1u;
		}
}

template <int kQueries>
__device__ __forceinline__ void lookupTogether(const LookupDag& d, const bool* live, const int* px, const int* py, const int* pz, u32* result) {
	if (d.leafCodes)
		lookupSlotsTogether<kQueries>(d, live, px, py, pz, result);
	else
		lookupWireTogether<kQueries>(d, live, px, py, pz, result);
}

constexpr int kPixelsPerThread = 2;  // measured on the 4K surface G-buffer: 1 -> 60.8, 2 -> 81.1, 3 -> 78.3, 4 -> 76.4 G lookups/s

template <bool kSurface>
__global__ void __launch_bounds__(256) evaluatePairsKernel(LookupDag d, PixelSource io, u32 width, u32 height, Mat4 mat) {
	constexpr int kPixels = kPixelsPerThread;
	const u32 x0 = (blockIdx.x * blockDim.x + threadIdx.x) * (u32)kPixels, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x0 >= width || y >= height) return;
	const size_t i = (size_t)y * width + x0;
	const int resolution = (1 << (d.dagLevels + d.gridLevels - 1)) - 1;
	int px[kPixels], py[kPixels], pz[kPixels];
	bool live[kPixels];
#pragma unroll
	for (int q = 0; q < kPixels; ++q) {
		live[q] = x0 + (u32)q < width;
		const u32 xq = live[q] ? x0 + (u32)q : x0;
		float4 p;
		if (kSurface)
			p = surf2Dread<float4>(io.posSurface, (int)(xq * sizeof(float4)), (int)y);
		else
			p = io.pos[(size_t)y * width + xq];
		float v[4];
#pragma unroll
		for (int r = 0; r < 4; ++r)
			v[r] = __fadd_rn(__fadd_rn(__fmul_rn(mat.m[r], p.x), __fmul_rn(mat.m[4 + r], p.y)), __fadd_rn(__fmul_rn(mat.m[8 + r], p.z), mat.m[12 + r]));
		px[q] = pathCoord(__fdiv_rn(v[0], v[3]), resolution);
		py[q] = pathCoord(__fdiv_rn(v[1], v[3]), resolution);
		pz[q] = pathCoord(__fdiv_rn(v[2], v[3]), resolution);
	}
	u32 result[kPixels];
	lookupTogether<kPixels>(d, live, px, py, pz, result);
#pragma unroll
	for (int q = 0; q < kPixels; ++q) {
		if (!live[q]) continue;
		const unsigned char vis = result[q] == 1u ? 255 : 0;
		if (kSurface)
			surf2Dwrite(vis, io.outSurface, (int)(x0 + (u32)q), (int)y);
		else
			io.out[i + q] = vis;
	}
}

// The same for plain NDC points: two consecutive points per thread.
__global__ void __launch_bounds__(256) lookupNdcPairsKernel(LookupDag d, const float* __restrict__ ndc, long long count, unsigned char* __restrict__ out) {
	constexpr int kPoints = 2;
	const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * kPoints;
	if (i0 >= count) return;
	const int resolution = (1 << (d.dagLevels + d.gridLevels - 1)) - 1;
	int px[kPoints], py[kPoints], pz[kPoints];
	bool live[kPoints];
#pragma unroll
	for (int q = 0; q < kPoints; ++q) {
		live[q] = i0 + q < count;
		const long long i = live[q] ? i0 + q : i0;
		px[q] = pathCoord(ndc[3 * i], resolution);
		py[q] = pathCoord(ndc[3 * i + 1], resolution);
		pz[q] = pathCoord(ndc[3 * i + 2], resolution);
	}
	u32 result[kPoints];
	lookupTogether<kPoints>(d, live, px, py, pz, result);
#pragma unroll
	for (int q = 0; q < kPoints; ++q)
		if (live[q]) out[i0 + q] = (unsigned char)result[q];
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
	lookupNdcPairsKernel<<<(unsigned)((count + 511) / 512), 256, 0, stream>>>(d, ndc, count, out);
	return 1;
}

int launchEvaluate(const LookupDag& d, const float* positions, unsigned width, unsigned height, const float* matrix, int filterSize, unsigned char* out,
		cudaStream_t stream) {
	if (!width || !height) return 0;
	Mat4 m;
	for (int i = 0; i < 16; ++i) m.m[i] = matrix[i];
	const dim3 block(8, 32), grid((width + 7) / 8, (height + 31) / 32);
	PixelSource io{reinterpret_cast<const float4*>(positions), out, 0, 0};
	if (filterSize <= 1) {
		evaluatePairsKernel<false><<<dim3((width + 8 * kPixelsPerThread - 1) / (8 * kPixelsPerThread), (height + 31) / 32), block, 0, stream>>>(d, io, width, height, m);
	} else {
		evaluateKernel<false><<<grid, block, 0, stream>>>(d, io, width, height, m, filterSize);
	}
	return 1;
}

int launchEvaluateSurface(const LookupDag& d, unsigned long long positions, unsigned long long visibilities, unsigned width, unsigned height,
		const float* matrix, int filterSize, cudaStream_t stream) {
	if (!width || !height) return 0;
	Mat4 m;
	for (int i = 0; i < 16; ++i) m.m[i] = matrix[i];
	const dim3 block(8, 32), grid((width + 7) / 8, (height + 31) / 32);
	PixelSource io{nullptr, nullptr, (cudaSurfaceObject_t)positions, (cudaSurfaceObject_t)visibilities};
	if (filterSize <= 1) {
		evaluatePairsKernel<true><<<dim3((width + 8 * kPixelsPerThread - 1) / (8 * kPixelsPerThread), (height + 31) / 32), block, 0, stream>>>(d, io, width, height, m);
	} else {
		evaluateKernel<true><<<grid, block, 0, stream>>>(d, io, width, height, m, filterSize);
	}
	return 1;
}

}  // namespace cpvs
