// evaluate() on CUDA surfaces (cpvs_container_evaluate_surface): the headless half of the CUDA-GL interop path. The G-buffer's
// rgba32f position texture and the r8 visibility texture of the reference (src/DeferredRenderer.cpp:18,80) are stood in for
// by cudaArrays; with GL they are the arrays cudaGraphicsSubResourceGetMappedArray hands out. The result must equal
// cpvs_container_evaluate on linear buffers, byte for byte, with and without filtering.
//
//   surface_test <n> <width> <height>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cpvs_b200.h"

#define CHECK(call)                                                                      \
	do {                                                                                 \
		int rc_ = (call);                                                                \
		if (rc_ != CPVS_OK) {                                                            \
			std::fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, cpvs_last_error()); \
			return 1;                                                                    \
		}                                                                                \
	} while (0)
#define CUDA(call)                                                                        \
	do {                                                                                  \
		cudaError_t e_ = (call);                                                          \
		if (e_ != cudaSuccess) {                                                          \
			std::fprintf(stderr, "%s failed: %s\n", #call, cudaGetErrorString(e_));       \
			return 1;                                                                     \
		}                                                                                 \
	} while (0)

int main(int argc, char** argv) {
	const int n = argc > 1 ? std::atoi(argv[1]) : 512;
	const int width = argc > 2 ? std::atoi(argv[2]) : 640, height = argc > 3 ? std::atoi(argv[3]) : 360;
	cpvs_ctx* ctx = nullptr;
	CHECK(cpvs_ctx_create(0, &ctx));
	std::vector<float> depth((size_t)n * n);
	for (int y = 0; y < n; ++y)
		for (int x = 0; x < n; ++x) depth[(size_t)y * n + x] = 0.3f + 0.4f * (float)x / (float)n + 0.2f * (float)((x / 37 + y / 23) & 1);
	cpvs_shadow* shadow = nullptr;
	CHECK(cpvs_shadow_create_from_depth(ctx, depth.data(), n, CPVS_MEM_HOST, 0, 1, 1, &shadow));
	cpvs_container* cont = nullptr;
	CHECK(cpvs_container_create(ctx, 1, &cont));
	CHECK(cpvs_container_set(cont, shadow, 0, 0, 0));
	CHECK(cpvs_container_finalize(cont));

	std::vector<float> pos((size_t)width * height * 4);
	uint32_t s = 4242;
	for (size_t i = 0; i < pos.size(); ++i) {
		s ^= s << 13;
		s ^= s >> 17;
		s ^= s << 5;
		pos[i] = (i & 3) == 3 ? 1.0f : (s % 20001) / 10000.f - 1.f;
	}
	const float m[16] = {0.9f, 0.05f, 0.f, 0.f, -0.05f, 0.9f, 0.f, 0.f, 0.f, 0.f, 0.8f, 0.f, 0.02f, -0.03f, 0.1f, 1.f};  // column-major

	cudaArray_t posArray = nullptr, visArray = nullptr;
	const cudaChannelFormatDesc f4 = cudaCreateChannelDesc<float4>(), u8 = cudaCreateChannelDesc<unsigned char>();
	CUDA(cudaMallocArray(&posArray, &f4, width, height, cudaArraySurfaceLoadStore));
	CUDA(cudaMallocArray(&visArray, &u8, width, height, cudaArraySurfaceLoadStore));
	CUDA(cudaMemcpy2DToArray(posArray, 0, 0, pos.data(), (size_t)width * 16, (size_t)width * 16, height, cudaMemcpyHostToDevice));
	cudaResourceDesc desc = {};
	desc.resType = cudaResourceTypeArray;
	cudaSurfaceObject_t posSurf = 0, visSurf = 0;
	desc.res.array.array = posArray;
	CUDA(cudaCreateSurfaceObject(&posSurf, &desc));
	desc.res.array.array = visArray;
	CUDA(cudaCreateSurfaceObject(&visSurf, &desc));

	for (uint32_t filter : {1u, 3u, 4u}) {
		CHECK(cpvs_container_set_filter_size(cont, filter));
		std::vector<uint8_t> linear((size_t)width * height), viaSurface((size_t)width * height, 7);
		CHECK(cpvs_container_evaluate(cont, pos.data(), width, height, CPVS_MEM_HOST, m, linear.data()));
		CHECK(cpvs_container_evaluate_surface(cont, posSurf, visSurf, width, height, m));
		CHECK(cpvs_ctx_synchronize(ctx));
		CUDA(cudaMemcpy2DFromArray(viaSurface.data(), width, visArray, 0, 0, width, height, cudaMemcpyDeviceToHost));
		size_t lit = 0, grey = 0;
		for (size_t i = 0; i < linear.size(); ++i) {
			if (linear[i] != viaSurface[i]) {
				std::fprintf(stderr, "filter %u: pixel %zu differs (%u vs %u)\n", filter, i, linear[i], viaSurface[i]);
				return 1;
			}
			lit += linear[i] == 255;
			grey += linear[i] != 0 && linear[i] != 255;
		}
		if ((filter == 1 && grey != 0) || lit == 0 || lit == linear.size()) {
			std::fprintf(stderr, "filter %u: implausible result (%zu lit, %zu grey)\n", filter, lit, grey);
			return 1;
		}
		std::printf("filter %u: %zu lit, %zu partially lit of %zu pixels; surfaces == linear buffers\n", filter, lit, grey, linear.size());
	}
	cudaDestroySurfaceObject(posSurf);
	cudaDestroySurfaceObject(visSurf);
	cudaFreeArray(posArray);
	cudaFreeArray(visArray);
	cpvs_container_destroy(cont);
	cpvs_shadow_destroy(shadow);
	cpvs_ctx_destroy(ctx);
	std::printf("surface_test ok\n");
	return 0;
}
