"""C++ facade (include/cpvs/*.h): the reference's class API on top of the C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROGRAM = r"""
#include "CompressedShadowContainer.h"
#include "CompressedShadowUtil.h"
#include <cstdio>
int main() {
    const int n = 64;
    ImageF img(n, n, 1);
    for (int y = 0; y < n; ++y) for (int x = 0; x < n; ++x) img.set(x, y, 0, 0.3f + 0.4f * x / n + 0.013f * y / n);
    MinMaxHierarchy mm(img);
    if (mm.getNumLevels() != 7) return 1;
    auto cs = CompressedShadow::create(mm);
    std::printf("words %zu levels %u vis %d\n", cs->getDAG().size(), cs->getNumLevels(), (int)cs->getTotalVisibility());
    if (cs->traverse(vec3(-1.f, -1.f, -1.f)) != CompressedShadow::VISIBLE) return 2;
    if (cs->traverse(vec3(1.f, 1.f, 1.f)) != CompressedShadow::SHADOW) return 3;
    // create(const ShadowMap*, ...) (reference src/CompressedShadow.h:55-56): same words as through an explicit hierarchy
    auto depth = std::make_shared<ImageF>(n, n, 1);
    depth->setAll(img.data());
    ShadowMap shadowMap(depth);
    auto fromMap = CompressedShadow::create(&shadowMap);
    if (fromMap->getDAG() != cs->getDAG()) return 5;
    auto slice = CompressedShadow::create(&shadowMap, 1, 2);
    if (slice->getDAG() != CompressedShadow::create(mm, 1, 2)->getDAG()) return 6;
    CompressedShadowContainer box(std::move(cs));
    box.setFilterSize(1);
    box.moveToGPU();
    const float p[3] = {-1.f, -1.f, -1.f};
    uint8_t out = 9;
    box.lookupNdc(p, 1, &out);
    return out == 1 ? 0 : 4;
}
"""


def _compile(tmp_path, extra):
    src = tmp_path / "facade.cpp"
    src.write_text(PROGRAM)
    exe = tmp_path / "facade"
    from cpvs_b200 import build
    lib = build.build()
    subprocess.check_call(["g++", "-std=c++14", "-Wall", "-I", os.path.join(ROOT, "include", "cpvs")] + extra +
                          [str(src), "-o", str(exe), "-L", os.path.dirname(lib), "-lcpvs_b200", "-Wl,-rpath," + os.path.dirname(lib)])
    return exe


def test_facade_compiles_without_glm(tmp_path):
    _compile(tmp_path, [])


@pytest.mark.gpu
def test_facade_runs(tmp_path):
    exe = _compile(tmp_path, [])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "levels 7" in out.stdout


@pytest.mark.gpu
def test_reference_gtests_against_cuda_library():
    """The reference's own 18 gtests (test/*.cpp, unmodified) compiled against include/cpvs and linked with
    libcpvs_b200.so by `make -C oracle facade-tests` in the build container; here they run on the B200."""
    exe = os.path.join(ROOT, "oracle", "_ref", "runUnitTests_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/runUnitTests_b200 not built (needs /root/reference at build time)")
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "[  PASSED  ] 18 tests." in out.stdout
