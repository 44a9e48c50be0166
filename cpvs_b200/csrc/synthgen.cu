// Device-resident depth source for the synthetic scenes: stands in for the shadow-map
// render + read-back of the reference (src/ShadowMap.cpp:23-30, src/DeferredRenderer.cpp:173-177) so a
// tile-grid build never leaves the GPU. The scenes are defined once in synth/scene.h; the bytes
// written here equal the host generator's (plane: IEEE mul/div/add in the same order, no FMA; city:
// a min over the same box list; terrain_dev: synth/scene.h's terrainDevDepth compiled for both sides). SURVEY.md's own
// terrain uses the host libm and has no device twin.
#include "../synth/scene.h"
#include "kernels.h"

namespace cpvs {
namespace {

constexpr int kGenTileW = 128, kGenTileH = 32;  // texels per CTA: 256 threads x (4 x 4)
constexpr int kGenThreads = 256;

__global__ void __launch_bounds__(kGenThreads) planeDepthKernel(float* __restrict__ out, int n, long long gx0, long long gy0, float fN) {
	const int x = (blockIdx.x * kGenTileW) + (threadIdx.x & 31) * 4;
	const int yBase = blockIdx.y * kGenTileH + (threadIdx.x >> 5) * 4;
	if (x >= n) return;
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		const int y = yBase + j;
		if (y >= n) break;
		const float fy = __fdiv_rn(__fmul_rn(0.013f, (float)(gy0 + y)), fN);
		float4 v;
		v.x = __fadd_rn(__fadd_rn(0.3f, __fdiv_rn(__fmul_rn(0.4f, (float)(gx0 + x + 0)), fN)), fy);
		v.y = __fadd_rn(__fadd_rn(0.3f, __fdiv_rn(__fmul_rn(0.4f, (float)(gx0 + x + 1)), fN)), fy);
		v.z = __fadd_rn(__fadd_rn(0.3f, __fdiv_rn(__fmul_rn(0.4f, (float)(gx0 + x + 2)), fN)), fy);
		v.w = __fadd_rn(__fadd_rn(0.3f, __fdiv_rn(__fmul_rn(0.4f, (float)(gx0 + x + 3)), fN)), fy);
		__stcs(reinterpret_cast<float4*>(out + (size_t)y * n + x), v);
	}
}

__global__ void __launch_bounds__(kGenThreads) terrainDevDepthKernel(float* __restrict__ out, int n, long long gx0, long long gy0, float fN) {
	const int x = (blockIdx.x * kGenTileW) + (threadIdx.x & 31) * 4;
	const int yBase = blockIdx.y * kGenTileH + (threadIdx.x >> 5) * 4;
	if (x >= n) return;
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		const int y = yBase + j;
		if (y >= n) break;
		const float v = __fdiv_rn((float)(gy0 + y), fN);
		float4 d;
		d.x = cpvs_synth::terrainDevDepth(__fdiv_rn((float)(gx0 + x + 0), fN), v);
		d.y = cpvs_synth::terrainDevDepth(__fdiv_rn((float)(gx0 + x + 1), fN), v);
		d.z = cpvs_synth::terrainDevDepth(__fdiv_rn((float)(gx0 + x + 2), fN), v);
		d.w = cpvs_synth::terrainDevDepth(__fdiv_rn((float)(gx0 + x + 3), fN), v);
		__stcs(reinterpret_cast<float4*>(out + (size_t)y * n + x), d);
	}
}

// One CTA rasterises a 128 x 32 region: the window's box list is scanned in chunks of 256; boxes that
// cover the region entirely only lower a region-wide scalar, the few that cut it are kept in shared
// memory and tested per texel. Depths are positive floats, so their bit patterns order like ints.
__global__ void __launch_bounds__(kGenThreads) cityDepthKernel(float* __restrict__ out, int n, const CityBoxDev* __restrict__ boxes,
		int numBoxes, float farPlane) {
	__shared__ CityBoxDev sCut[kGenThreads];
	__shared__ int sCutCount;
	__shared__ int sCoverBits;

	const int rx0 = blockIdx.x * kGenTileW, ry0 = blockIdx.y * kGenTileH;
	const int rx1 = min(rx0 + kGenTileW, n), ry1 = min(ry0 + kGenTileH, n);
	const int x = rx0 + (threadIdx.x & 31) * 4;
	const int yBase = ry0 + (threadIdx.x >> 5) * 4;

	if (threadIdx.x == 0) sCoverBits = __float_as_int(farPlane);
	float d[4][4];
#pragma unroll
	for (int j = 0; j < 4; ++j)
#pragma unroll
		for (int i = 0; i < 4; ++i) d[j][i] = farPlane;

	for (int base = 0; base < numBoxes; base += kGenThreads) {
		if (threadIdx.x == 0) sCutCount = 0;
		__syncthreads();
		const int b = base + threadIdx.x;
		if (b < numBoxes) {
			const CityBoxDev box = boxes[b];
			const bool touches = box.x0 < rx1 && box.x1 > rx0 && box.y0 < ry1 && box.y1 > ry0;
			if (touches) {
				const bool covers = box.x0 <= rx0 && box.x1 >= rx1 && box.y0 <= ry0 && box.y1 >= ry1;
				if (covers)
					atomicMin(&sCoverBits, __float_as_int(box.z));
				else
					sCut[atomicAdd(&sCutCount, 1)] = box;
			}
		}
		__syncthreads();
		const int cut = sCutCount;
		for (int c = 0; c < cut; ++c) {
			const CityBoxDev box = sCut[c];
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				const int y = yBase + j;
				const bool inY = y >= box.y0 && y < box.y1;
#pragma unroll
				for (int i = 0; i < 4; ++i)
					if (inY && x + i >= box.x0 && x + i < box.x1) d[j][i] = fminf(d[j][i], box.z);
			}
		}
		__syncthreads();
	}
	__syncthreads();  // also orders thread 0's initialisation when the box list is empty
	const float cover = __int_as_float(sCoverBits);
	if (x >= n) return;
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		const int y = yBase + j;
		if (y >= n) break;
		const float4 v = make_float4(fminf(d[j][0], cover), fminf(d[j][1], cover), fminf(d[j][2], cover), fminf(d[j][3], cover));
		__stcs(reinterpret_cast<float4*>(out + (size_t)y * n + x), v);
	}
}

}  // namespace

int launchPlaneDepth(float* out, int n, long long gx0, long long gy0, long long gn, cudaStream_t stream) {
	const dim3 grid((n + kGenTileW - 1) / kGenTileW, (n + kGenTileH - 1) / kGenTileH);
	planeDepthKernel<<<grid, kGenThreads, 0, stream>>>(out, n, gx0, gy0, (float)gn);
	return 1;
}

int launchTerrainDevDepth(float* out, int n, long long gx0, long long gy0, long long gn, cudaStream_t stream) {
	const dim3 grid((n + kGenTileW - 1) / kGenTileW, (n + kGenTileH - 1) / kGenTileH);
	terrainDevDepthKernel<<<grid, kGenThreads, 0, stream>>>(out, n, gx0, gy0, (float)gn);
	return 1;
}

int launchCityDepth(float* out, int n, const CityBoxDev* boxes, int numBoxes, float farPlane, cudaStream_t stream) {
	const dim3 grid((n + kGenTileW - 1) / kGenTileW, (n + kGenTileH - 1) / kGenTileH);
	cityDepthKernel<<<grid, kGenThreads, 0, stream>>>(out, n, boxes, numBoxes, farPlane);
	return 1;
}

}  // namespace cpvs
