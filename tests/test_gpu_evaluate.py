"""GPU suite (-m gpu): CompressedShadowContainer::evaluate (reference src/CompressedShadowContainer.cpp:93-124 + shader/traverse.cs)
pinned against an evaluation that shares no code with the library or the oracle: the light transform in numpy float32 in glm's
operation order, the path and the grid cell from shader/traverse.cs:41-48,78-88, and the expected visibility straight from the
depth tiles (a voxel is lit iff z + 0.5 <= depth * H, reference src/CompressedShadowUtil.h:47-57). Also the percentage-closer
filter of setFilterSize and the CUDA-surface entry point (the headless half of the CUDA-GL interop)."""
import os
import subprocess

import numpy as np
import pytest

import cpvs_b200
from cpvs_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F = np.float32


def _container(ctx, kind, tile, length):
    cont = cpvs_b200.CompressedShadowContainer(length, ctx)
    depth = np.empty((length * tile, length * tile), np.float32)
    for y in range(length):
        for x in range(length):
            d = synth.depth_map(kind, tile, (x, y), length)
            depth[y * tile:(y + 1) * tile, x * tile:(x + 1) * tile] = d
            mm = cpvs_b200.MinMaxHierarchy(d, ctx, zTileNum=length)
            for z in range(length):
                cont.set(cpvs_b200.CompressedShadow.create(mm, z, length), x, y, z)
    cont.copyToGPU()
    return cont, depth


def _paths(pos, m, res):
    """glm's mat4 * vec4 (Mul0 + Mul1) + (Mul2 + Mul3) with w = 1, the divide by w, and the truncating path of traverse.cs,
    clamped to the volume -- all in float32, one rounding per operation."""
    m = np.asarray(m, F).reshape(4, 4)  # column-major: m[c][r]
    x, y, z = pos[..., 0].astype(F), pos[..., 1].astype(F), pos[..., 2].astype(F)
    v = [(m[0][r] * x + m[1][r] * y) + (m[2][r] * z + m[3][r]) for r in range(4)]
    out = []
    for c in range(3):
        ndc = (v[c] / v[3]).astype(F)
        f = ((ndc + F(1)) * F(0.5)).astype(F) * F(res)
        p = np.where(f > 0, np.minimum(np.trunc(f), res), 0).astype(np.int64)
        out.append(p)
    return out


def _expected_lit(depth, px, py, pz, height):
    return (pz.astype(F) + F(0.5)) <= depth[py, px] * F(height)


@pytest.mark.parametrize("kind,tile,length", [("terrain", 128, 1), ("terrain_dev", 128, 2), ("city", 64, 4)])
def test_evaluate_against_independent_transform(gpu_ctx, kind, tile, length):
    cont, depth = _container(gpu_ctx, kind, tile, length)
    side = tile * length
    rng = np.random.default_rng(7)
    h, w = 270, 480
    pos = np.empty((h, w, 4), np.float32)
    pos[..., :3] = rng.uniform(-1.3, 1.3, (h, w, 3)).astype(np.float32)
    pos[..., 3] = 1
    mats = [np.eye(4, dtype=np.float32),
            np.array([[0.8, 0.1, 0, 0], [-0.1, 0.7, 0.05, 0], [0, 0.02, 0.9, 0], [0.03, -0.02, 0.05, 1]], np.float32),  # columns
            np.array([[1.1, 0, 0, 0.1], [0, 0.9, 0, -0.05], [0, 0, 0.7, 0.2], [0, 0.1, 0, 1.5]], np.float32)]      # perspective: w varies
    for m in mats:
        px, py, pz = _paths(pos, m, side - 1)
        want = _expected_lit(depth, px, py, pz, side)
        got = cont.evaluate(pos, m)
        assert np.array_equal(got != 0, want), float((got != 0).mean())
        assert set(np.unique(got)) <= {0, 255}
    # filtered: the mean over size x size voxels of the slice, clamped at the borders
    for size in (2, 3, 5):
        cont.setFilterSize(size)
        m = mats[1]
        px, py, pz = _paths(pos, m, side - 1)
        lo = -(size // 2)
        lit = np.zeros(px.shape, np.int64)
        for dy in range(lo, lo + size):
            for dx in range(lo, lo + size):
                lit += _expected_lit(depth, np.clip(px + dx, 0, side - 1), np.clip(py + dy, 0, side - 1), pz, side)
        want = (255 * lit + (size * size) // 2) // (size * size)
        got = cont.evaluate(pos, m)
        assert np.array_equal(got.astype(np.int64), want), size
    cont.setFilterSize(1)
    with pytest.raises(cpvs_b200.CpvsError):
        cont.setFilterSize(0)


def test_cpp_caller_evaluates_on_cuda_surfaces(tmp_path):
    """tests/cpp/surface_test.cpp: cudaArrays + surface objects standing in for the GL textures of the renderer."""
    from cpvs_b200 import build
    lib = build.build()
    exe = str(tmp_path / "surface_test")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
                           os.path.join(ROOT, "tests", "cpp", "surface_test.cpp"), "-o", exe, "-L", os.path.dirname(lib), "-lcpvs_b200",
                           "-L", "/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + os.path.dirname(lib) + ":/usr/local/cuda/lib64"])
    out = subprocess.run([exe, "512", "640", "360"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "surface_test ok" in out.stdout, out.stdout + out.stderr
