"""Dev helper: one 2048^2 build (large-level kernels: look-back expansion, inserts, rank, emit) for compute-sanitizer."""
import sys
sys.path.insert(0, ".")
import cpvs_b200
from cpvs_b200 import synth

ctx = cpvs_b200.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
d = synth.depth_map("terrain", n)
mm = cpvs_b200.MinMaxHierarchy(d, ctx)
for leaf in (True, False):
    sh = cpvs_b200.CompressedShadow.create(mm, 0, 1, leaf)
    print(n, leaf, int(sh.info.words), synth.fnv64(sh.getDAG()) == synth.fnv64(sh.getDAG()), int(sh.traverse(synth.lookups(20000), leaf).sum()))
