"""Dev helper (no torch): a few builds of one resident depth map, for ncu launch lists and `--set full` captures.

    python scripts/one_build.py [size] [kind] [builds]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpvs_b200  # noqa: E402
from cpvs_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
kind = sys.argv[2] if len(sys.argv) > 2 else "terrain"
builds = int(sys.argv[3]) if len(sys.argv) > 3 else 3
lib = cpvs_b200.load_library()
ctx = cpvs_b200.Context(0)
mm0 = cpvs_b200.MinMaxHierarchy(synth.depth_map(kind, n), ctx)
ctx.synchronize()
dptr = int(lib.cpvs_minmax_level_device(mm0.handle, 0))
for i in range(builds):
    mm = cpvs_b200.MinMaxHierarchy(dptr, ctx, n=n)
    sh = cpvs_b200.CompressedShadow.create(mm)
    print("build %d: %s create %.3f pyramid %s" % (i, {k: round(v, 3) for k, v in sh.phase_ms().items()}, sh.info.build_ms, mm.timing()), flush=True)
    sh.close()
    mm.close()
