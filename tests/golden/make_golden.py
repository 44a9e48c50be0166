"""Regenerates tests/golden/*.npz|json from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Needs /root/reference and oracle/_ref (``make -C oracle ref``). Nothing here runs on the GPU box: the
outputs are committed. Sources of the vectors:
  * depths8x8 / depths16x16 / depths32x32: the fixtures of reference test/TestImages.cpp:3-66, dumped
    by compiling that file into a throw-away binary (the file itself is not copied);
  * minmax8x8 / minmax4x4: the literal images of reference test/MinMaxTest.cpp:15-23,58-64;
  * every DAG / pyramid / lookup result: produced by oracle/_ref/libcpvs_ref*.so, i.e. the reference's
    own MinMaxHierarchy, CompressedShadow::create and CompressedShadow::traverse.
"""
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cpvs_b200 import synth  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

DUMPER = r"""
#include "TestImages.h"
#include <cstdio>
static void dump(const char* name, const vector<float>& v) {
    FILE* f = fopen(name, "wb"); fwrite(v.data(), sizeof(float), v.size(), f); fclose(f);
}
int main() { dump("d8.bin", getDepths8x8()); dump("d16.bin", getDepths16x16()); dump("d32.bin", getDepths32x32()); }
"""


def reference_test_images():
    with tempfile.TemporaryDirectory() as tmp:
        with open(os.path.join(tmp, "dump.cpp"), "w") as f:
            f.write(DUMPER)
        subprocess.check_call(["g++", "-std=c++14", "-w", "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + REF + "/glm",
                               "-I" + REF + "/src", "-I" + REF + "/test", "dump.cpp", REF + "/test/TestImages.cpp", "-o", "dump"], cwd=tmp)
        subprocess.check_call(["./dump"], cwd=tmp)
        imgs = {}
        for n in (8, 16, 32):
            imgs["depths%dx%d" % (n, n)] = np.fromfile(os.path.join(tmp, "d%d.bin" % n), np.float32).reshape(n, n)
    text = open(REF + "/test/MinMaxTest.cpp").read()
    blocks = re.findall(r"vector<float>\s*\{([^}]*)\}", text)
    vals = [np.array([float(t) for t in re.findall(r"[-+]?\d*\.?\d+", b)], np.float32) for b in blocks]
    imgs["minmax8x8"] = [v for v in vals if v.size == 64][0].reshape(8, 8)
    imgs["minmax4x4"] = [v for v in vals if v.size == 16][0].reshape(4, 4)
    return imgs


def main():
    assert os.path.isdir(REF) and O.have_ref(), "needs /root/reference and `make -C oracle ref`"
    arrays, meta = {}, {}
    for name, img in reference_test_images().items():
        arrays[name] = img
        mm = O.MinMax(img, "ref")
        for lvl in range(1, mm.num_levels()):
            arrays["%s.minmax%d" % (name, lvl)] = mm.level(lvl)
        if img.shape[0] >= 8:
            arrays[name + ".dag"] = O.Shadow(mm).dag()
            arrays[name + ".dag_noleaf"] = O.Shadow(O.MinMax(img, "ref_noleaf"), leafmasks=False).dag()
        if img.shape[0] == 16:
            for t in (0, 1):
                arrays["%s.dag_z%dof2" % (name, t)] = O.Shadow(mm, t, 2).dag()
    # constant maps: 1.0 -> [0x5555], 0.0 -> [0x0], 0.5 -> one word
    for val in (0.0, 0.5, 1.0):
        arrays["const%.1f.dag" % val] = O.Shadow(O.MinMax(np.full((64, 64), val, np.float32), "ref")).dag()

    pts = synth.lookups(100000)
    table = []
    for kind in ("plane", "terrain", "city"):
        for n in (64, 256, 1024):
            d = synth.depth_map(kind, n)
            for zt, zn in ((0, 1), (1, 2), (3, 4)):
                if n == 1024 and (zt, zn) != (0, 1) and kind == "terrain":
                    continue
                sh = O.Shadow(O.MinMax(d, "ref"), zt, zn)
                dag = sh.dag()
                svo, offs = O.MinMax(d, "ref").svo(zt, zn)
                vis = sh.traverse(pts)
                table.append({"kind": kind, "n": n, "z_tile": zt, "z_num": zn, "leafmasks": True, "words": int(dag.size),
                              "fnv64": "%016x" % synth.fnv64(dag), "svo_words": int(svo.size),
                              "svo_fnv64": "%016x" % synth.fnv64(svo), "lit": int((vis == 1).sum()),
                              "vis_fnv64": "%016x" % synth.fnv64(vis.astype(np.uint32))})
            if n <= 256:
                sh = O.Shadow(O.MinMax(d, "ref_noleaf"), leafmasks=False)
                dag = sh.dag()
                vis = sh.traverse(pts, False)
                table.append({"kind": kind, "n": n, "z_tile": 0, "z_num": 1, "leafmasks": False, "words": int(dag.size),
                              "fnv64": "%016x" % synth.fnv64(dag), "lit": int((vis == 1).sum()),
                              "vis_fnv64": "%016x" % synth.fnv64(vis.astype(np.uint32))})
    meta["synthetic"] = table
    meta["lookups"] = {"count": 100000, "seed": 777}
    np.savez_compressed(os.path.join(OUT, "reference_vectors.npz"), **arrays)
    with open(os.path.join(OUT, "reference_synthetic.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", len(arrays), "arrays,", len(table), "synthetic rows")


if __name__ == "__main__":
    main()
