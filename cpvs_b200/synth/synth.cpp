/* Synthetic depth maps and lookup points for tests and bench.py.
 *
 * Host-side workload generators: no CUDA, no oracle code. The same bytes are fed to the CUDA path,
 * to the CPU oracle and to the compiled reference, so the maps only have to be deterministic. They
 * stand in for what the reference renders with GL and reads back in ShadowMap::createImageF
 * (reference src/ShadowMap.cpp:23-30): one float32 depth per texel, row-major, values in (0,1).
 *
 *   plane   - tilted ground as seen from the default light direction (reference src/main.cpp:35)
 *   terrain - smooth height field with three octaves (sinf/cosf: results depend on the host libm)
 *   terrain_dev - the same height field on the deterministic sin/cos of scene.h: byte-identical to the CUDA generator
 *   city    - far plane at 0.9 with N/8 random axis-aligned boxes (xorshift32, seed 12345; libm-free)
 *
 * A tile (tx,ty) of a tilesPerSide x tilesPerSide virtual map samples the same function at the
 * global coordinate (tx*N+x)/(tilesPerSide*N), which is how DeferredRenderer::renderWithTiles
 * (reference src/DeferredRenderer.cpp:165-187) cuts the light frustum with getSubProjection.
 *
 * Build: g++ -O2 -ffp-contract=off -fPIC -shared synth.cpp -o libcpvs_synth.so -lpthread
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

#include "scene.h"

using cpvs_synth::XorShift32;

namespace {

template <typename F>
void parallelRows(int n, int threads, F rowFn) {
	threads = std::max(1, std::min(threads, n));
	if (threads == 1) {
		for (int y = 0; y < n; ++y) rowFn(y);
		return;
	}
	std::vector<std::thread> pool;
	for (int t = 0; t < threads; ++t)
		pool.emplace_back([=]() {
			for (int y = t; y < n; y += threads) rowFn(y);
		});
	for (auto& th : pool) th.join();
}

}  // namespace

extern "C" {

enum {
	CPVS_SYNTH_PLANE = cpvs_synth::kPlane,
	CPVS_SYNTH_TERRAIN = cpvs_synth::kTerrain,
	CPVS_SYNTH_CITY = cpvs_synth::kCity,
	CPVS_SYNTH_TERRAIN_DEV = cpvs_synth::kTerrainDev
};

/* out: n*n floats. (tx,ty,tilesPerSide) select a window of the virtual map; (0,0,1) is the whole map. */
int cpvs_synth_depth(int kind, int n, int tx, int ty, int tilesPerSide, int threads, float* out) {
	if (n <= 0 || tilesPerSide <= 0 || !out) return -1;
	const long gn = static_cast<long>(n) * tilesPerSide;  // virtual side in texels
	const long gx0 = static_cast<long>(tx) * n, gy0 = static_cast<long>(ty) * n;
	const float fN = static_cast<float>(gn);

	if (kind == CPVS_SYNTH_PLANE) {
		parallelRows(n, threads, [=](int y) {
			float* row = out + static_cast<size_t>(y) * n;
			for (int x = 0; x < n; ++x)
				row[x] = 0.3f + 0.4f * (float)(gx0 + x) / fN + 0.013f * (float)(gy0 + y) / fN;
		});
		return 0;
	}
	if (kind == CPVS_SYNTH_TERRAIN) {
		parallelRows(n, threads, [=](int y) {
			float* row = out + static_cast<size_t>(y) * n;
			const float v = (float)(gy0 + y) / fN;
			for (int x = 0; x < n; ++x) {
				const float u = (float)(gx0 + x) / fN;
				row[x] = 0.5f + 0.15f * sinf(9.1f * u) * cosf(7.3f * v) + 0.05f * sinf(41.f * u + 3.f * v) +
						 0.02f * cosf(97.f * v - 11.f * u);
			}
		});
		return 0;
	}
	if (kind == CPVS_SYNTH_TERRAIN_DEV) {
		parallelRows(n, threads, [=](int y) {
			float* row = out + static_cast<size_t>(y) * n;
			const float v = (float)(gy0 + y) / fN;
			for (int x = 0; x < n; ++x) row[x] = cpvs_synth::terrainDevDepth((float)(gx0 + x) / fN, v);
		});
		return 0;
	}
	if (kind == CPVS_SYNTH_CITY) {
		std::vector<cpvs_synth::CityBox> boxes;
		cpvs_synth::forEachCityBox(gn, gx0, gy0, n, [&](const cpvs_synth::CityBox& box) { boxes.push_back(box); });
		parallelRows(n, threads, [&boxes, out, n](int y) {
			float* row = out + static_cast<size_t>(y) * n;
			std::fill(row, row + n, cpvs_synth::kCityFarPlane);
			for (const cpvs_synth::CityBox& box : boxes) {
				if (y < box.y0 || y >= box.y1) continue;
				for (int x = box.x0; x < box.x1; ++x) row[x] = std::min(row[x], box.z);
			}
		});
		return 0;
	}
	return -1;
}

/* count points in [-1,1]^3 as x,y,z triples; xorshift32 with the given seed (777 in the survey). */
void cpvs_synth_lookups(uint32_t seed, long count, float* out) {
	XorShift32 rng(seed);
	for (long i = 0; i < 3 * count; ++i) out[i] = (rng.next() % 20001) / 10000.f - 1.f;
}

/* The survey's 64-bit FNV-1a-style digest over 32-bit words (its basis is one digit short of the
 * standard one; kept so the SURVEY.md 8c table reproduces). */
uint64_t cpvs_synth_fnv64(const uint32_t* words, long count) {
	uint64_t h = 1469598103934665603ull;
	for (long i = 0; i < count; ++i) {
		h ^= words[i];
		h *= 1099511628211ull;
	}
	return h;
}

}  // extern "C"
