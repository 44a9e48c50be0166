/* The synthetic scenes of SURVEY.md 8d as one definition shared by the host generator (synth.cpp) and
 * the device generator of the CUDA library (csrc/synthgen.cu), so both produce the same bytes.
 * Workload definition only: no oracle code, nothing from the reference.
 */
#pragma once
#include <algorithm>
#include <cstdint>

namespace cpvs_synth {

enum { kPlane = 0, kTerrain = 1, kCity = 2 };
constexpr uint32_t kMapSeed = 12345u;
constexpr float kCityFarPlane = 0.9f;

struct XorShift32 {
	uint32_t s;
	explicit XorShift32(uint32_t seed) : s(seed) {}
	uint32_t next() {
		s ^= s << 13;
		s ^= s >> 17;
		s ^= s << 5;
		return s;
	}
};

/* One occluder of the city scene, clipped to a window of the virtual map: texels [x0,x1) x [y0,y1) in
 * window coordinates get min(depth, z). */
struct CityBox {
	int x0, y0, x1, y1;
	float z;
};

/* Calls fn(CityBox) for every box of the gn x gn city that touches the n x n window at (gx0, gy0), in
 * generation order. The random sequence is consumed for every box, touching or not. */
template <typename F>
inline void forEachCityBox(long gn, long gx0, long gy0, long n, F fn) {
	XorShift32 rng(kMapSeed);
	const uint32_t ugn = static_cast<uint32_t>(gn);
	const long nb = gn / 8;
	for (long b = 0; b < nb; ++b) {
		const long w = 8 + rng.next() % (ugn / 16 + 1);
		const long h = 8 + rng.next() % (ugn / 16 + 1);
		const long x0 = rng.next() % ugn;
		const long y0 = rng.next() % ugn;
		const float z = 0.2f + 0.6f * (rng.next() % 1024) / 1024.f;
		const long xa = std::max(x0, gx0), xb = std::min(std::min(x0 + w, gn), gx0 + n);
		const long ya = std::max(y0, gy0), yb = std::min(std::min(y0 + h, gn), gy0 + n);
		if (xa >= xb || ya >= yb) continue;
		fn(CityBox{static_cast<int>(xa - gx0), static_cast<int>(ya - gy0), static_cast<int>(xb - gx0),
				   static_cast<int>(yb - gy0), z});
	}
}

}  // namespace cpvs_synth
