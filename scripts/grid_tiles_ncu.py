"""Dev probe for ncu launch lists: one worker builds a handful of tiles of a grid (after one untimed tile)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpvs_b200  # noqa: E402
from cpvs_b200 import grid as cgrid  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "city"
length = int(sys.argv[2]) if len(sys.argv) > 2 else 16
tile = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
tiles = [(int(a), int(b)) for a, b in (p.split(",") for p in sys.argv[4:])] or [(3, 5), (8, 8), (12, 1), (0, 0)]
ctx = cpvs_b200.Context(0)
ctx.reserve(int(12e9))
w = cgrid.GridWorker(ctx, length, tile, kind)
w.build(tiles)
print("device %.3f ms for %d tiles, depth %.3f ms, %d launches" % (w.device_ms(), len(tiles), w.depth_ms(), ctx.launch_count))
w.close()
