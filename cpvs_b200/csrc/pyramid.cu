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
template <bool kWriteLow>
__global__ void __launch_bounds__(256) pyramidBaseKernel(const float* __restrict__ depth, int n, float2* __restrict__ l1,
		float2* __restrict__ l2, float2* __restrict__ l3, float2* __restrict__ l4, float2* __restrict__ l5) {
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

int launchPyramid(const float* depth, int n, float* const* levels, int numLevels, bool writeLowLevels, cudaEvent_t afterBase, cudaStream_t stream) {
	int launches = 0;
	int next = 1;
	if (n >= 128) {
		dim3 grid(n / 128, n / 32);
		float2 *l1 = reinterpret_cast<float2*>(levels[1]), *l2 = reinterpret_cast<float2*>(levels[2]), *l3 = reinterpret_cast<float2*>(levels[3]),
			   *l4 = reinterpret_cast<float2*>(levels[4]), *l5 = reinterpret_cast<float2*>(levels[5]);
		if (writeLowLevels)
			pyramidBaseKernel<true><<<grid, 256, 0, stream>>>(depth, n, l1, l2, l3, l4, l5);
		else
			pyramidBaseKernel<false><<<grid, 256, 0, stream>>>(depth, n, l1, l2, l3, l4, l5);
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
