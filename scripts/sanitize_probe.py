"""Dev helper: small builds / lookups / containers for compute-sanitizer (memcheck, racecheck)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import cpvs_b200
from cpvs_b200 import synth

ctx = cpvs_b200.Context(0)
pts = synth.lookups(5000)
for kind, n in (("terrain", 256), ("city", 128), ("plane", 64), ("terrain", 16), ("city", 8)):
    d = synth.depth_map(kind, n)
    mm = cpvs_b200.MinMaxHierarchy(d, ctx)
    for leaf in (True, False):
        sh = cpvs_b200.CompressedShadow.create(mm, 0, 1, leaf)
        use_leaf = bool(sh.info.leafmasks)
        vis = sh.traverse(pts, use_leaf)
        print(kind, n, leaf, int(sh.info.words), int(vis.sum()))
cont = cpvs_b200.CompressedShadowContainer(2, ctx)
for y in range(2):
    for x in range(2):
        mm = cpvs_b200.MinMaxHierarchy(synth.depth_map("terrain", 64, (x, y), 2), ctx)
        for z in range(2):
            cont.set(cpvs_b200.CompressedShadow.create(mm, z, 2), x, y, z)
cont.copyToGPU()
print("container", cont.info(), int(cont.lookup_ndc(pts).sum()))
