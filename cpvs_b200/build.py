"""In-tree build of libcpvs_b200.so (hand-written CUDA for sm_100a behind the C ABI of include/cpvs_b200.h).

    python -m cpvs_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcpvs_b200.so")
SOURCES = ["capi.cu", "build.cu", "grid.cu", "pyramid.cu", "svo.cu", "merge.cu", "emit.cu", "lookup.cu", "lookup_index.cu", "synthgen.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",  # float products feeding comparisons must round like the reference's
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--shared",
]


def _newest_source():
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "cpvs_b200.h"),
                                                                os.path.join(HERE, "synth", "scene.h")]
    return max(os.path.getmtime(p) for p in paths)


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
