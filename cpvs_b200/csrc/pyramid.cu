// K1 -- min/max depth pyramid (reference MinMaxHierarchy, src/MinMaxHierarchy.cpp:9-97).
//
// Level 1 holds (min,max) of every 2x2 block of the depth map, level k of every 2x2 block of level
// k-1, reduced as pred(pred(a,b),pred(c,d)) with a=(x,y) b=(x+1,y) c=(x,y+1) d=(x+1,y+1) and
// std::min/std::max comparison order (src/MinMaxHierarchy.cpp:29-33,46-47), so the stored bits match
// the reference's even for signed zeros.
//
// HBM-bound: reads 4*N^2 bytes once, writes (8/3)*N^2. The base kernel keeps a 128x32 depth tile in
// registers/shared memory and emits levels 1..5 from it in one pass (128-bit loads and stores); the
// remaining levels (1/1024 of the data) go through a generic one-level kernel.
#include "kernels.h"

namespace cpvs {

namespace {

__device__ __forceinline__ float2 reduce4(float2 a, float2 b, float2 c, float2 d) {
	return make_float2(stdMin(stdMin(a.x, b.x), stdMin(c.x, d.x)), stdMax(stdMax(a.y, b.y), stdMax(c.y, d.y)));
}

// 256 threads, one 128 (x) by 32 (y) depth tile per CTA. Requires n >= 128.
// kWriteLow = false skips the stores of levels 1 and 2: the leafmask builder never reads them (it
// classifies against levels >= 3 and builds leaves from level 0), and they are 37% of this kernel's
// traffic. They are produced on demand (launchPyramidLowLevels) for accessors and the leafmask-less mode.
// kResidue: also writes the depth map re-encoded for the per-column leaf builder (svo.cu buildLeafColumnsResidueKernel): per
// 8x8-texel column 64 bytes, 8 per row in the order x = 0 4 1 5 2 6 3 7, byte = clamp(T - 8*lo, 0, 255), where T = floor(fl(depth * H) + 0.5) is the number of
// lit slices below the texel (H = n * zTileNum) and lo the column's first z-block, floor(fl(min * H/8)) of its level-3 texel.
// A leaf's k-code is min(residue - 8 * block, 8) per texel, so the builder reads 1 byte per texel instead of 4 and
// needs no float arithmetic; columns taller than 31 z-blocks (box edges) do not fit a byte and keep the depth path.
template <bool kWriteLow, bool kResidue>
__global__ void __launch_bounds__(256) pyramidBaseKernel(const float* __restrict__ depth, int n, float2* __restrict__ l1,
		float2* __restrict__ l2, float2* __restrict__ l3, float2* __restrict__ l4, float2* __restrict__ l5, unsigned char* __restrict__ residue,
		float heightF, float height3F) {
	__shared__ float2 s2[8][32];
	__shared__ float2 s3[4][16];
	__shared__ float2 s4[2][8];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int x0 = blockIdx.x * 128 + lane * 4, y0 = blockIdx.y * 32 + warp * 4;

	float4 r[4];
#pragma unroll
	for (int i = 0; i < 4; ++i) r[i] = __ldcs(reinterpret_cast<const float4*>(depth + (size_t)(y0 + i) * n + x0));

	// level 1: two rows of two texels
	float2 m[2][2];
#pragma unroll
	for (int i = 0; i < 2; ++i) {
		const float4 t = r[2 * i], b = r[2 * i + 1];
		m[i][0] = make_float2(stdMin(stdMin(t.x, t.y), stdMin(b.x, b.y)), stdMax(stdMax(t.x, t.y), stdMax(b.x, b.y)));
		m[i][1] = make_float2(stdMin(stdMin(t.z, t.w), stdMin(b.z, b.w)), stdMax(stdMax(t.z, t.w), stdMax(b.z, b.w)));
	}
	const int n1 = n >> 1, n2 = n >> 2, n3 = n >> 3, n4 = n >> 4, n5 = n >> 5;
	if (kWriteLow) {
#pragma unroll
		for (int i = 0; i < 2; ++i)
			*reinterpret_cast<float4*>(l1 + (size_t)(y0 / 2 + i) * n1 + x0 / 2) = make_float4(m[i][0].x, m[i][0].y, m[i][1].x, m[i][1].y);
	}

	// level 2: one texel per thread
	const float2 v2 = reduce4(m[0][0], m[0][1], m[1][0], m[1][1]);
	if (kWriteLow) l2[(size_t)(y0 / 4) * n2 + x0 / 4] = v2;
	s2[warp][lane] = v2;
	__syncthreads();

	const int bx3 = blockIdx.x * 16, by3 = blockIdx.y * 4;
	if (threadIdx.x < 64) {
		const int x = threadIdx.x & 15, y = threadIdx.x >> 4;
		const float2 v = reduce4(s2[2 * y][2 * x], s2[2 * y][2 * x + 1], s2[2 * y + 1][2 * x], s2[2 * y + 1][2 * x + 1]);
		l3[(size_t)(by3 + y) * n3 + bx3 + x] = v;
		s3[y][x] = v;
	}
	__syncthreads();
	if (kResidue) {
		// the thread's 4 x 4 texels are rows (warp & 1) * 4 .. + 3, bytes (lane & 1) * 4 .. + 3 of column (lane >> 1, warp >> 1)
		const float lo3 = fmaxf(floorf(__fmul_rn(s3[warp >> 1][lane >> 1].x, height3F)), 0.0f);  // NaN: 0, as columnRange
		const float shift = __fsub_rn(0.5f, __fmul_rn(lo3, 8.0f));                             // exact
		unsigned char* col = residue + ((size_t)(by3 + (warp >> 1)) * n3 + bx3 + (lane >> 1)) * 64 + (warp & 1) * 32 + (lane & 1) * 4;
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const float d[4] = {r[i].x, r[i].y, r[i].z, r[i].w};
			unsigned int b[4];
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				// floor of the exact sum (rounded toward -inf, which never crosses an integer), clamped to a byte. NaN becomes
				// residue 0 = never lit, as midZ <= NaN is false.
				const float x = __fadd_rd(__fmul_rn(d[k], heightF), shift);
				// (one conversion that floors and saturates to 0..255, NaN -> 0; it runs on the otherwise idle conversion pipe)
				asm("cvt.rmi.sat.u8.f32 %0, %1;" : "=r"(b[k]) : "f"(x));
			}
			const unsigned int lo = __byte_perm(b[0], b[1], 0x0040), hi = __byte_perm(b[2], b[3], 0x0040);
			// A row is stored with texels x and x + 4 next to each other (bytes 0 4 1 5 | 2 6 3 7): the builder works on such
			// pairs. The even lane holds texels 0..3 of the row, its odd neighbour 4..7; each writes one of the two words.
			const unsigned int mine = __byte_perm(lo, hi, 0x5410), other = __shfl_xor_sync(0xFFFFFFFFu, mine, 1);
			*reinterpret_cast<unsigned int*>(col + i * 8) = (lane & 1) ? __byte_perm(mine, other, 0x3726) : __byte_perm(mine, other, 0x5140);
		}
	}
	if (threadIdx.x < 16) {
		const int x = threadIdx.x & 7, y = threadIdx.x >> 3;
		const float2 v = reduce4(s3[2 * y][2 * x], s3[2 * y][2 * x + 1], s3[2 * y + 1][2 * x], s3[2 * y + 1][2 * x + 1]);
		l4[(size_t)(by3 / 2 + y) * n4 + bx3 / 2 + x] = v;
		s4[y][x] = v;
	}
	__syncthreads();
	if (threadIdx.x < 4) {
		const int x = threadIdx.x;
		const float2 v = reduce4(s4[0][2 * x], s4[0][2 * x + 1], s4[1][2 * x], s4[1][2 * x + 1]);
		l5[(size_t)(by3 / 4) * n5 + bx3 / 4 + x] = v;
	}
}

// One level from the one below; srcChannels = 1 for the depth map, 2 for (min,max) levels.
__global__ void pyramidLevelKernel(const float* __restrict__ src, int srcChannels, int outSide, float2* __restrict__ dst) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= outSide || y >= outSide) return;
	const size_t s = (size_t)outSide * 2;
	float2 a, b, c, d;
	if (srcChannels == 1) {
		const float* p = src + (size_t)(2 * y) * s + 2 * x;
		a = make_float2(p[0], p[0]);
		b = make_float2(p[1], p[1]);
		c = make_float2(p[s], p[s]);
		d = make_float2(p[s + 1], p[s + 1]);
	} else {
		const float2* p = reinterpret_cast<const float2*>(src) + (size_t)(2 * y) * s + 2 * x;
		a = p[0];
		b = p[1];
		c = p[s];
		d = p[s + 1];
	}
	dst[(size_t)y * outSide + x] = reduce4(a, b, c, d);
}

// All remaining levels in one CTA once a level fits in shared memory (side <= 64): the top of the
// pyramid is a chain of tiny dependent steps, cheaper as barriers than as kernel launches.
struct TailLevels {
	float2* level[kMaxLevels];  // level[k] for k in (first, last]
};
__global__ void __launch_bounds__(1024) pyramidTailKernel(const float2* __restrict__ src, int srcSide, TailLevels out, int firstLevel, int lastLevel) {
	__shared__ float2 sA[64 * 64], sB[32 * 32];
	for (int i = threadIdx.x; i < srcSide * srcSide; i += blockDim.x) sA[i] = src[i];
	__syncthreads();
	float2* cur = sA;
	float2* nxt = sB;
	int side = srcSide;
	for (int k = firstLevel + 1; k <= lastLevel; ++k) {
		const int o = side >> 1;
		for (int i = threadIdx.x; i < o * o; i += blockDim.x) {
			const int x = i % o, y = i / o;
			const float2 v = reduce4(cur[(2 * y) * side + 2 * x], cur[(2 * y) * side + 2 * x + 1], cur[(2 * y + 1) * side + 2 * x],
					cur[(2 * y + 1) * side + 2 * x + 1]);
			nxt[i] = v;
			out.level[k][i] = v;
		}
		__syncthreads();
		float2* t = cur;
		cur = nxt;
		nxt = t;
		side = o;
	}
}

}  // namespace

int launchPyramidLowLevels(const float* depth, int n, float* const* levels, cudaStream_t stream) {
	for (int k = 1; k <= 2; ++k) {
		const int side = n >> k;
		dim3 block(16, 16), grid((side + 15) / 16, (side + 15) / 16);
		pyramidLevelKernel<<<grid, block, 0, stream>>>(k == 1 ? depth : levels[1], k == 1 ? 1 : 2, side, reinterpret_cast<float2*>(levels[k]));
	}
	return 2;
}

int launchPyramid(const float* depth, int n, float* const* levels, int numLevels, bool writeLowLevels, unsigned char* residue, unsigned residueTiles,
		cudaEvent_t afterBase, cudaStream_t stream) {
	int launches = 0;
	int next = 1;
	if (n >= 128) {
		dim3 grid(n / 128, n / 32);
		float2 *l1 = reinterpret_cast<float2*>(levels[1]), *l2 = reinterpret_cast<float2*>(levels[2]), *l3 = reinterpret_cast<float2*>(levels[3]),
			   *l4 = reinterpret_cast<float2*>(levels[4]), *l5 = reinterpret_cast<float2*>(levels[5]);
		const float heightF = (float)((unsigned)n * residueTiles), height3F = (float)(((unsigned)n >> 3) * residueTiles);
		if (writeLowLevels && residue)
			pyramidBaseKernel<true, true><<<grid, 256, 0, stream>>>(depth, n, l1, l2, l3, l4, l5, residue, heightF, height3F);
		else if (writeLowLevels)
			pyramidBaseKernel<true, false><<<grid, 256, 0, stream>>>(depth, n, l1, l2, l3, l4, l5, nullptr, 0.f, 0.f);
		else if (residue)
			pyramidBaseKernel<false, true><<<grid, 256, 0, stream>>>(depth, n, l1, l2, l3, l4, l5, residue, heightF, height3F);
		else
			pyramidBaseKernel<false, false><<<grid, 256, 0, stream>>>(depth, n, l1, l2, l3, l4, l5, nullptr, 0.f, 0.f);
		++launches;
		next = 6;
	}
	if (afterBase) cudaEventRecord(afterBase, stream);
	for (int k = next; k < numLevels; ++k) {
		const int srcSide = n >> (k - 1);
		if (k >= 2 && srcSide <= 64) {  // levels[k-1] is a (min,max) level that fits in shared memory
			TailLevels t;
			for (int j = 0; j < kMaxLevels; ++j) t.level[j] = j < numLevels ? reinterpret_cast<float2*>(levels[j]) : nullptr;
			pyramidTailKernel<<<1, 1024, 0, stream>>>(reinterpret_cast<const float2*>(levels[k - 1]), srcSide, t, k - 1, numLevels - 1);
			++launches;
			break;
		}
		const int side = n >> k;
		dim3 block(side >= 16 ? 16 : side, side >= 16 ? 16 : side);
		dim3 grid((side + block.x - 1) / block.x, (side + block.y - 1) / block.y);
		pyramidLevelKernel<<<grid, block, 0, stream>>>(levels[k - 1], k == 1 ? 1 : 2, side, reinterpret_cast<float2*>(levels[k]));
		++launches;
	}
	return launches;
}

}  // namespace cpvs
