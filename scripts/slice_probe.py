"""Dev probe: the z-slices of a few tiles of a grid one by one (synchronous builds): words, device time, phases."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpvs_b200  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "city"
length = int(sys.argv[2]) if len(sys.argv) > 2 else 16
tile = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
ctx = cpvs_b200.Context(0)
ctx.reserve(int(12e9))
depth = torch.empty((tile, tile), dtype=torch.float32, device="cuda")
names = ["count", "expand", "leaves", "leaf_ins", "inner", "bases", "emit_in", "emit_lf", "leaf_res"]
for rep, (x, y) in enumerate([(3, 5), (3, 5), (8, 8), (12, 1)]):
    cpvs_b200.generate_depth(kind, tile, depth, tile=(x, y), tiles_per_side=length, ctx=ctx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mm = cpvs_b200.MinMaxHierarchy(depth, ctx=ctx, zTileNum=length)
    e1.record()
    torch.cuda.synchronize()
    print("tile (%d, %d): pyramid %.3f ms (events on torch's stream: host-side bracket)" % (x, y, e0.elapsed_time(e1)))
    total = 0.0
    for z in range(length):
        s = cpvs_b200.CompressedShadow.create(mm, z, length, ctx=ctx)
        info = s.info
        if info.words > 1:
            total += info.build_ms
            print("  z=%2d words %9d  %.3f ms  " % (z, info.words, info.build_ms) + " ".join("%s %.3f" % (n, p) for n, p in zip(names, info.phase_ms)))
    print("  sum of slices %.3f ms" % total)
