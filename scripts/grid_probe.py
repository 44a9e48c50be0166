"""Dev probe: one GPU builds a whole grid a few times through GridWorker; device time, wall time and the context's statistics."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpvs_b200  # noqa: E402
from cpvs_b200 import grid as cgrid, tiling  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "terrain_dev"
length = int(sys.argv[2]) if len(sys.argv) > 2 else 4
tile = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ctx = cpvs_b200.Context(0)
ctx.reserve(int(min(length * length * 320e6, 8 * 2.0 ** 30) + 6e9))
tiles = tiling.xy_tiles(length)
for rep in range(reps):
    w = cgrid.GridWorker(ctx, length, tile, kind)
    t0 = time.perf_counter()
    w.build(tiles)
    wall = (time.perf_counter() - t0) * 1e3
    st = ctx.stats()
    print("rep %d: device %.2f ms, depth %.2f ms, wall %.2f ms; predicted %d exact %d rebuilds %d reemissions %d" % (
        rep, w.device_ms(), w.depth_ms(), wall, st["predicted_builds"], st["exact_builds"], st["overflow_rebuilds"], st["reemissions"]))
    w.close()
w = cgrid.GridWorker(ctx, length, tile, kind)
w.build(tiles)
for c in w.cells():
    if c.words > 1:
        z, rest = divmod(int(c.index), length * length)
        y, x = divmod(rest, length)
        print("cell x%d y%d z%d: words %10d svo %10d dag %9d  words/svo %.3f" % (x, y, z, c.words, c.svo_nodes, c.dag_nodes, c.words / c.svo_nodes))
w.close()
