"""CPU checks of the arithmetic the per-column leaf builder relies on (cpvs_b200/csrc/svo.cu, DESIGN.md "Why the result is
bit-exact"): the reference decides a texel's slice by `(z + (z+1)) * 0.5 <= d*H` in fp32 (cs::absoluteVisible, reference
src/CompressedShadowUtil.h:51-57); the kernel counts T = floor(RD(q + (0.5 - 8 zb))) lit slices above a z-block in fp32 and
turns R = clamp(T, 0, 2040) into the leaf's nibbles with two fp16 FMAs. Everything here is plain numpy."""
import numpy as np


def _rd_add_f32(a, b):
    """fp32 addition rounded toward -inf (add.rm.f32) of fp32 values, emulated exactly in float64."""
    exact = a.astype(np.float64) + b.astype(np.float64)  # exact: both operands have 24-bit significands within 2^-30 .. 2^24
    near = exact.astype(np.float32)
    too_big = near.astype(np.float64) > exact
    return np.where(too_big, np.nextafter(near, np.float32(-np.inf)), near).astype(np.float32)


def _reference_lit_count(q, z0):
    """Number of lit slices among z0 .. z0+7, slice by slice as the reference does."""
    k = np.zeros(q.shape, np.int64)
    for dz in range(8):
        z = np.float32(z0 + dz)
        mid = (z + (z + np.float32(1.0))) * np.float32(0.5)
        k += (mid <= q)
    return k


def test_lit_slices_below_equals_reference_predicate():
    rng = np.random.default_rng(7)
    for height in (16.0, 1024.0, 16384.0, 3 * 4096.0, float(1 << 23)):
        d = np.concatenate([rng.random(20000, dtype=np.float32),
                            (np.arange(0, 4000, dtype=np.float32) * np.float32(0.125) + np.float32(0.5)) / np.float32(height),  # exact halves
                            np.array([0.0, 1.0, 1e-8, 0.49999997, 0.5, 0.50000006, 0.99999994], np.float32)])
        q = (d * np.float32(height)).astype(np.float32)  # fl(d * H), the product the reference compares against
        top = int(height) // 8
        for zb in sorted({0, 1, 2, top // 3, top // 2, max(top - 2, 0), max(top - 1, 0)}):
            shift = np.float32(0.5) - np.float32(zb) * np.float32(8.0)  # exact: 8 zb < 2^23
            t = np.floor(_rd_add_f32(q, np.full(q.shape, shift, np.float32)))
            k = np.clip(t, 0, 8).astype(np.int64)
            assert np.array_equal(k, _reference_lit_count(q, 8 * zb)), (height, zb)


def test_nearest_rounding_would_be_wrong():
    """Why the sum is rounded down: q just below a half-integer must not reach the next slice."""
    q = np.nextafter(np.float32(0.5), np.float32(0.0))  # 0.49999997: slice 0 is NOT lit (0.5 <= q is false)
    assert np.floor(np.float32(q + np.float32(0.5))) == 1.0  # round-to-nearest lands on 1.0
    assert np.floor(_rd_add_f32(np.array([q]), np.array([0.5], np.float32)))[0] == 0.0
    assert _reference_lit_count(np.array([q]), 0)[0] == 0


def test_fp16_nibbles_are_exact():
    """k = 8 * sat(R/8 - i) + 1024 in fp16 for every R the kernel can hold and every block of a 252-block chunk."""
    r = np.arange(0, 2041, dtype=np.float64)
    assert np.array_equal(r.astype(np.float16).astype(np.float64), r)  # R itself
    for i in range(252):
        x = r / 8.0 - i  # the FMA's exact value
        x16 = x.astype(np.float16)  # its single rounding
        inside = (x >= 0) & (x <= 1)
        assert np.array_equal(x16[inside].astype(np.float64), x[inside])  # multiples of 1/8: representable
        assert (x16[x < 0] <= 0).all() and (x16[x > 1] >= 1).all()  # rounding never crosses the saturation bounds
        s = np.clip(x16.astype(np.float64), 0.0, 1.0)
        y = (s * 8.0 + 1024.0).astype(np.float16)
        k = y.view(np.uint16).astype(np.int64) - 0x6400  # the kernel strips the bits of 1024
        assert np.array_equal(k, np.clip(r - 8 * i, 0, 8).astype(np.int64)), i


def test_row_word_packing_strips_the_exponent_bits():
    """Four half2 registers (texels p and p+4 each) -> one row word, nibble x = k of texel x (svo.cu emitColumnLeaves)."""
    rng = np.random.default_rng(3)
    for _ in range(200):
        k = rng.integers(0, 9, size=8)
        regs = [((0x6400 + int(k[p + 4])) << 16) | (0x6400 + int(k[p])) for p in range(4)]
        w = 0
        for p in (3, 2, 1, 0):
            w = (w * 16 + regs[p]) & 0xFFFFFFFF
        w = (w - ((0x64006400 * 0x1111) & 0xFFFFFFFF)) & 0xFFFFFFFF
        assert w == sum(int(k[x]) << (4 * x) for x in range(8))
