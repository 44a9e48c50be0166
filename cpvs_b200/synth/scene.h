/* The synthetic scenes of SURVEY.md 8d as one definition shared by the host generator (synth.cpp) and
 * the device generator of the CUDA library (csrc/synthgen.cu), so both produce the same bytes.
 * Workload definition only: no oracle code, nothing from the reference.
 */
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define CPVS_SYNTH_HD __host__ __device__
#else
#define CPVS_SYNTH_HD
#endif

namespace cpvs_synth {

/* kTerrain is SURVEY.md 8d's height field verbatim (sinf/cosf of the host libm: host generator only, its digests are tied
 * to the build container's glibc). kTerrainDev is the same formula on sinDet/cosDet below -- plain IEEE float arithmetic in
 * a fixed order, no FMA -- so the host generator (g++ -ffp-contract=off) and the CUDA generator (--fmad=false) write
 * the same bytes and a tile grid never has to be copied to the device. */
enum { kPlane = 0, kTerrain = 1, kCity = 2, kTerrainDev = 3 };

/* sin and cos for |x| < 2^15: Cody-Waite reduction by pi/2 split into three floats (the first with 8 significant bits, so
 * k * kPiHalf1 is exact), then the cephes single-precision polynomials on [-pi/4, pi/4]. Accurate to a few ulp, which is
 * all a synthetic scene needs; what matters is that every operation is a single correctly rounded IEEE one. */
CPVS_SYNTH_HD inline void sinCosDet(float x, float* s, float* c) {
	const float kf = floorf(x * 0.636619772f + 0.5f);
	float r = x - kf * 1.5703125f;
	r = r - kf * 4.837512969970703125e-4f;
	r = r - kf * 7.54978995489188216e-8f;
	const float z = r * r;
	const float sp = r + r * z * (-1.6666654611e-1f + z * (8.3321608736e-3f + z * -1.9515295891e-4f));
	const float cp = (1.0f - 0.5f * z) + z * z * (4.166664568298827e-2f + z * (-1.388731625493765e-3f + z * 2.443315711809948e-5f));
	const int q = static_cast<int>(kf) & 3;
	*s = q == 0 ? sp : (q == 1 ? cp : (q == 2 ? -sp : -cp));
	*c = q == 0 ? cp : (q == 1 ? -sp : (q == 2 ? -cp : sp));
}

/* Depth of the device-generatable terrain at (u, v) in [0,1)^2 of the virtual map. */
CPVS_SYNTH_HD inline float terrainDevDepth(float u, float v) {
	float s1, c1, s2, c2, s3, c3, s4, c4;
	sinCosDet(9.1f * u, &s1, &c1);
	sinCosDet(7.3f * v, &s2, &c2);
	sinCosDet(41.f * u + 3.f * v, &s3, &c3);
	sinCosDet(97.f * v - 11.f * u, &s4, &c4);
	return 0.5f + 0.15f * s1 * c2 + 0.05f * s3 + 0.02f * c4;
}
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
