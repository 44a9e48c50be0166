"""Dev probe (no torch): the z-slices of one tile of a tile grid -- per-slice device time and phases, repeated so that the
later rounds run on warm memos.   python scripts/column_probe.py [kind] [length] [tile x] [tile y] [size]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpvs_b200  # noqa: E402
from cpvs_b200 import synth  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "terrain_dev"
length = int(sys.argv[2]) if len(sys.argv) > 2 else 4
tx, ty = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1, 2)
n = int(sys.argv[5]) if len(sys.argv) > 5 else 16384
lib = cpvs_b200.load_library()
ctx = cpvs_b200.Context(0)
mm0 = cpvs_b200.MinMaxHierarchy(synth.depth_map(kind, n, (tx, ty), length), ctx)
ctx.synchronize()
dptr = int(lib.cpvs_minmax_level_device(mm0.handle, 0))
for rep in range(4):
    t0 = time.perf_counter()
    mm = cpvs_b200.MinMaxHierarchy(dptr, ctx, n=n, zTileNum=length)
    rows = []
    for z in range(length):
        sh = cpvs_b200.CompressedShadow.create(mm, z, length)
        rows.append((z, int(sh.info.words), sh.info.build_ms, {k: round(v, 3) for k, v in sh.phase_ms().items() if v > 0.004}))
        sh.close()
    ctx.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    print("round %d: wall %.3f ms, pyramid %.3f, slices %.3f" % (rep, wall, mm.timing()[0], sum(r[2] for r in rows)))
    if rep == 3:
        for r in rows:
            print("   z=%d words=%d build=%.3f %s" % r)
    mm.close()

# the same slices alternating between two contexts, all in flight (what the grid worker does)
ctx2 = cpvs_b200.Context(0)
for rep in range(4):
    t0 = time.perf_counter()
    mm = cpvs_b200.MinMaxHierarchy(dptr, ctx, n=n, zTileNum=length)
    shs = [cpvs_b200.CompressedShadow.create(mm, z, length, ctx=(ctx2 if z & 1 else ctx), wait=False) for z in range(length)]
    t1 = time.perf_counter()
    for sh in shs:
        sh.wait()
    ctx.synchronize()
    ctx2.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    print("two contexts, round %d: wall %.3f ms (enqueue %.3f ms), sum of builds %.3f, stats %s %s" % (
        rep, wall, (t1 - t0) * 1e3, sum(sh.info.build_ms for sh in shs), ctx.stats(), ctx2.stats()))
    for sh in shs:
        sh.close()
    mm.close()
