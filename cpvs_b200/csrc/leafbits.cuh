// k-code -> 64-bit slice masks through bit planes (createLeafmask, reference src/CompressedShadowUtil.cpp:59-78).
//
// A leaf's k-code is one 32-bit word per row y, nibble x = k(x,y) = number of lit slices (0..8) of texel (x,y); slice s of
// the leaf is the 64-bit mask with bit x + 8y = (k(x,y) > s). Comparing nibble by nibble costs ~11 integer instructions per
// row and slice (8 x 8 x 11 per leaf). Here the code is first turned into its four bit planes P0..P3 (bit x + 8y of Pj =
// bit j of k(x,y)): one 8 x 4 bit transpose per row (four delta swaps) and two 4 x 4 byte transposes (PRMT). Every slice is
// then a boolean function of the planes, one or two LOP3 per 32-bit half:
//   k > 0: P3|P2|P1|P0   k > 1: P3|P2|P1     k > 2: P3|P2|(P1&P0)   k > 3: P3|P2
//   k > 4: P3|(P2&(P1|P0))   k > 5: P3|(P2&P1)   k > 6: P3|(P2&P1&P0)   k > 7: P3
// Integer logic only, so the host build of this header (tests/test_leaf_bits.py) checks exactly what the device runs.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CPVS_HD __host__ __device__ __forceinline__
#else
#define CPVS_HD inline
#endif

namespace cpvs {

// The direct way (what the emission did before the planes; kept as the second opinion of tests/test_leaf_bits.py): one row of
// one slice. nibble k > slice <=> bit 3 of (k + 7 - slice); nibbles are <= 8 so nothing carries.
CPVS_HD uint32_t rowBits(uint32_t code, uint32_t slice) {
	uint32_t y = ((code + (7u - slice) * 0x11111111u) >> 3) & 0x11111111u;
	y = (y | (y >> 3)) & 0x03030303u;
	y = (y | (y >> 6)) & 0x000F000Fu;
	return (y | (y >> 12)) & 0xFFu;
}

CPVS_HD uint32_t deltaSwap(uint32_t x, uint32_t mask, unsigned shift) {
	const uint32_t t = ((x >> shift) ^ x) & mask;
	return x ^ t ^ (t << shift);
}

// bit 4x + j  ->  bit 8j + x: byte j of the result holds bit j of the eight nibbles (a rotation of the five index bits,
// done as four transpositions of index bits = four delta swaps).
CPVS_HD uint32_t nibblesToPlaneBytes(uint32_t code) {
	uint32_t x = code;
	x = deltaSwap(x, 0x22222222u, 1);
	x = deltaSwap(x, 0x0A0A0A0Au, 3);
	x = deltaSwap(x, 0x00CC00CCu, 6);
	x = deltaSwap(x, 0x0000F0F0u, 12);
	return x;
}

// __byte_perm without the sign-replication mode: result byte i = byte (sel >> 4i) & 7 of the eight bytes (hi:lo).
CPVS_HD uint32_t bytePerm(uint32_t lo, uint32_t hi, uint32_t sel) {
#if defined(__CUDA_ARCH__)
	return __byte_perm(lo, hi, sel);
#else
	const uint64_t both = ((uint64_t)hi << 32) | lo;
	uint32_t r = 0;
	for (int i = 0; i < 4; ++i) r |= (uint32_t)((both >> (8u * ((sel >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
	return r;
#endif
}

// Four rows' plane bytes t[y] = [P3 P2 P1 P0] -> four planes' row bytes out[j] = [row3 row2 row1 row0].
CPVS_HD void transposeBytes4x4(uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3, uint32_t (&out)[4]) {
	const uint32_t a = bytePerm(t0, t1, 0x5140u), b = bytePerm(t0, t1, 0x7362u);
	const uint32_t c = bytePerm(t2, t3, 0x5140u), d = bytePerm(t2, t3, 0x7362u);
	out[0] = bytePerm(a, c, 0x5410u);
	out[1] = bytePerm(a, c, 0x7632u);
	out[2] = bytePerm(b, d, 0x5410u);
	out[3] = bytePerm(b, d, 0x7632u);
}

// lo[j] / hi[j]: rows 0..3 / 4..7 of bit plane j.
CPVS_HD void codeToPlanes(const uint32_t (&code)[8], uint32_t (&lo)[4], uint32_t (&hi)[4]) {
	transposeBytes4x4(nibblesToPlaneBytes(code[0]), nibblesToPlaneBytes(code[1]), nibblesToPlaneBytes(code[2]), nibblesToPlaneBytes(code[3]), lo);
	transposeBytes4x4(nibblesToPlaneBytes(code[4]), nibblesToPlaneBytes(code[5]), nibblesToPlaneBytes(code[6]), nibblesToPlaneBytes(code[7]), hi);
}

// One 32-bit half of slice s (0..7) from the same half of the four planes: bit set <=> k > s.
CPVS_HD uint32_t sliceFromPlanes(const uint32_t (&p)[4], unsigned s) {
	switch (s) {
		case 0: return p[3] | p[2] | p[1] | p[0];
		case 1: return p[3] | p[2] | p[1];
		case 2: return p[3] | p[2] | (p[1] & p[0]);
		case 3: return p[3] | p[2];
		case 4: return p[3] | (p[2] & (p[1] | p[0]));
		case 5: return p[3] | (p[2] & p[1]);
		case 6: return p[3] | (p[2] & p[1] & p[0]);
		default: return p[3];
	}
}

}  // namespace cpvs
