"""Where the time of one tile column goes (a 16K^2 city tile of the 256K^2 map, 16 z-slices): per slice the
device time between the first and last event of the build, the host wall time of the call, and the phases."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import cpvs_b200

n, length = 16384, 16
tile = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (5, 5)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = cpvs_b200.Context(0, stream=stream.cuda_stream)
depth = torch.empty((n, n), dtype=torch.float32, device="cuda")
cpvs_b200.generate_depth("city", n, depth, tile, length, ctx)
ctx.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    mm = cpvs_b200.MinMaxHierarchy(depth, ctx, n=n)
    ctx.synchronize()
    t_mm = (time.perf_counter() - t0) * 1e3
    rows = []
    for z in range(length):
        t0 = time.perf_counter()
        sh = cpvs_b200.CompressedShadow.create(mm, z, length)
        wall = (time.perf_counter() - t0) * 1e3
        rows.append((z, int(sh.info.words), int(sum(sh.info.svo_nodes)), float(sh.info.build_ms), wall, sh.phase_ms()))
        sh.close()
    mm.close()
    if rep == 2:
        print("pyramid wall %.3f ms" % t_mm)
        for z, words, nodes, dev_ms, wall, ph in rows:
            print("z=%2d words=%8d svo_nodes=%9d device %.3f ms wall %.3f ms %s" % (
                z, words, nodes, dev_ms, wall, {k: round(v, 3) for k, v in ph.items() if v > 0.0005} if words > 1 else ""))
        print("column: device %.3f ms, wall %.3f ms" % (sum(r[3] for r in rows), sum(r[4] for r in rows)))
