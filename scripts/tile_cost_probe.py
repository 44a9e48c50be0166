"""Dev probe: estimated cost of every xy tile of a grid (cpvs_grid_worker_estimate) next to the device time its build takes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpvs_b200  # noqa: E402
from cpvs_b200 import grid as cgrid, tiling  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "terrain_dev"
length = int(sys.argv[2]) if len(sys.argv) > 2 else 4
tile = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
ctx = cpvs_b200.Context(0)
ctx.reserve(int(12e9 * (tile / 16384.0) ** 2))
w = cgrid.GridWorker(ctx, length, tile, kind)
tiles = tiling.xy_tiles(length)
w.build(tiles[:1])
w.close()
w = cgrid.GridWorker(ctx, length, tile, kind)
costs = w.estimate(tiles)
est_ms = w.device_ms()
rows = []
for t, c in zip(tiles, costs):
    before = w.device_ms()
    w.build([t])
    rows.append((t, c, w.device_ms() - before))
tot_c, tot_ms = sum(r[1] for r in rows), sum(r[2] for r in rows)
print("estimates: %.3f ms for %d tiles" % (est_ms, len(tiles)))
for t, c, ms in rows:
    print("tile %s  cost share %.4f  time share %.4f  (%.3f ms)  ratio %.3f" % (t, c / tot_c, ms / tot_ms, ms, (c / tot_c) / (ms / tot_ms)))
