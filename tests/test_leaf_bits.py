"""CPU checks of the integer bit manipulation the leaf emission runs on the device (cpvs_b200/csrc/leafbits.cuh): the header
is plain integer logic marked __host__ __device__, so it is compiled here with g++ and compared with the definition of a
leafmask -- slice s of a leaf has bit x + 8y set iff texel (x, y) has more than s lit slices (createLeafmask, reference
src/CompressedShadowUtil.cpp:59-78). Covers the bit-plane expansion the emission kernel runs (codeToPlanes /
sliceFromPlanes) and the direct per-row one it replaced (rowBits, kept in the header as the second opinion)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SRC = r"""
#include "leafbits.cuh"
using namespace cpvs;
extern "C" void expand_rows(const uint32_t* codes, uint64_t n, uint64_t* out) {
	for (uint64_t i = 0; i < n; ++i)
		for (uint32_t s = 0; s < 8; ++s) {
			const uint32_t* c = codes + i * 8;
			const uint32_t lo = rowBits(c[0], s) | (rowBits(c[1], s) << 8) | (rowBits(c[2], s) << 16) | (rowBits(c[3], s) << 24);
			const uint32_t hi = rowBits(c[4], s) | (rowBits(c[5], s) << 8) | (rowBits(c[6], s) << 16) | (rowBits(c[7], s) << 24);
			out[i * 8 + s] = ((uint64_t)hi << 32) | lo;
		}
}
extern "C" void expand_planes(const uint32_t* codes, uint64_t n, uint64_t* out) {
	for (uint64_t i = 0; i < n; ++i) {
		uint32_t code[8], lo[4], hi[4];
		for (int y = 0; y < 8; ++y) code[y] = codes[i * 8 + y];
		codeToPlanes(code, lo, hi);
		for (uint32_t s = 0; s < 8; ++s) out[i * 8 + s] = ((uint64_t)sliceFromPlanes(hi, s) << 32) | sliceFromPlanes(lo, s);
	}
}
extern "C" uint32_t plane_bytes(uint32_t code) { return nibblesToPlaneBytes(code); }
extern "C" uint32_t byte_perm(uint32_t lo, uint32_t hi, uint32_t sel) { return bytePerm(lo, hi, sel); }
"""


@pytest.fixture(scope="module")
def bits(tmp_path_factory):
    d = tmp_path_factory.mktemp("leafbits")
    src = d / "leafbits_host.cpp"
    src.write_text(_SRC)
    lib = d / "libleafbits_host.so"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-I", os.path.join(ROOT, "cpvs_b200", "csrc"), str(src), "-o", str(lib)])
    h = ctypes.CDLL(str(lib))
    for f in (h.expand_rows, h.expand_planes):
        f.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p]
        f.restype = None
    h.plane_bytes.argtypes = [ctypes.c_uint32]
    h.plane_bytes.restype = ctypes.c_uint32
    h.byte_perm.argtypes = [ctypes.c_uint32] * 3
    h.byte_perm.restype = ctypes.c_uint32
    return h


def _codes_from_counts(k):
    """k: (n, 8, 8) lit-slice counts [leaf, y, x] -> (n, 8) row words, nibble x = k."""
    w = np.zeros(k.shape[:2], np.uint32)
    for x in range(8):
        w |= k[:, :, x].astype(np.uint32) << np.uint32(4 * x)
    return np.ascontiguousarray(w)


def _definition(k):
    """(n, 8) uint64 slice masks straight from the definition."""
    out = np.zeros((k.shape[0], 8), np.uint64)
    for s in range(8):
        lit = (k > s)
        for y in range(8):
            for x in range(8):
                out[:, s] |= lit[:, y, x].astype(np.uint64) << np.uint64(x + 8 * y)
    return out


def _counts():
    rng = np.random.default_rng(2)
    k = [rng.integers(0, 9, size=(4000, 8, 8)),  # anything
         np.clip(rng.integers(0, 9, size=(500, 1, 1)) + rng.integers(-1, 2, size=(500, 8, 8)), 0, 8),  # surfaces: neighbours differ by one
         np.stack([np.full((8, 8), v) for v in range(9)])]  # uniform leaves, k = 0 and k = 8 included
    one = np.zeros((64 * 9, 8, 8), np.int64)  # a single texel at every position and every count
    for i in range(64 * 9):
        one[i, (i // 9) // 8, (i // 9) % 8] = i % 9
    k.append(one)
    return np.concatenate(k).astype(np.int64)


@pytest.mark.parametrize("which", ["expand_rows", "expand_planes"])
def test_slice_masks_follow_the_definition(bits, which):
    k = _counts()
    codes = _codes_from_counts(k)
    out = np.zeros((k.shape[0], 8), np.uint64)
    getattr(bits, which)(codes.ctypes.data, k.shape[0], out.ctypes.data)
    want = _definition(k)
    bad = np.argwhere(out != want)
    assert bad.size == 0, (which, bad[0], hex(int(out[tuple(bad[0])])), hex(int(want[tuple(bad[0])])))


def test_plane_bytes_is_the_8x4_bit_transpose(bits):
    rng = np.random.default_rng(4)
    for code in [1 << i for i in range(32)] + [int(v) for v in rng.integers(0, 1 << 32, size=2000, dtype=np.uint64)]:
        want = 0
        for x in range(8):
            for j in range(4):
                if (code >> (4 * x + j)) & 1:
                    want |= 1 << (8 * j + x)
        assert bits.plane_bytes(code) == want, hex(code)


def test_byte_perm_host_stand_in(bits):
    """The host stand-in of __byte_perm for the selectors the transposes use (no sign-replication mode)."""
    lo, hi = 0x33221100, 0x77665544
    for sel, want in ((0x5140, 0x55114400), (0x7362, 0x77336622), (0x5410, 0x55441100), (0x7632, 0x77663322), (0x3210, lo), (0x7654, hi)):
        assert bits.byte_perm(lo, hi, sel) == want, hex(sel)
