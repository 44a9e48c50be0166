"""Dev helper for compute-sanitizer: the round-2 paths -- predicted and async builds on two contexts, z-slices through the staging
buffers, a pipelined grid worker on four lanes with tiles pulled from a queue, the lookup copy, PCF and surface-less evaluate."""
import sys
sys.path.insert(0, ".")
import numpy as np
import cpvs_b200
from cpvs_b200 import grid as cgrid, synth

ctx = cpvs_b200.Context(0)
ctx2 = cpvs_b200.Context(0)
pts = synth.lookups(5000)
# predicted + async builds alternating on two contexts
frames = [synth.depth_map("terrain", 256, (i, 0), 4) for i in range(4)]
pending = []
for i, d in enumerate(frames * 2):
    c = (ctx, ctx2)[i & 1]
    mm = cpvs_b200.MinMaxHierarchy(d, c)
    pending.append((mm, cpvs_b200.CompressedShadow.create(mm, 0, 1, True, ctx=c, wait=False)))
    if len(pending) > 2:
        mm0, s0 = pending.pop(0)
        s0.wait()
        print("async", int(s0.info.words), int(s0.info.predicted))
for mm0, s0 in pending:
    s0.wait()
# z-slices of one hierarchy in flight together (staging buffers)
mm = cpvs_b200.MinMaxHierarchy(synth.depth_map("terrain", 512), ctx, zTileNum=4)
slices = [cpvs_b200.CompressedShadow.create(mm, z, 4, True, ctx=(ctx, ctx2)[z & 1], wait=False) for z in range(4)]
print("slices", [int(s.info.words) for s in slices], int(sum(s.traverse(pts, True).sum() for s in slices)))
# grid worker: list, then queue
for kind, length, tile in (("city", 4, 128), ("terrain_dev", 2, 256)):
    tiles = [(x, y) for y in range(length) for x in range(length)]
    w = cgrid.GridWorker(ctx, length, tile, kind)
    w.build(tiles[: len(tiles) // 2])
    rest = list(tiles[len(tiles) // 2:])
    w.build_from(lambda: rest.pop(0) if rest else None)
    cells = w.cells()
    print(kind, len(cells), sum(int(c.words) for c in cells))
    parts = sorted(((c.index, c.words, c.root_mask, c.device, c.words_device) for c in cells))
    cont = cgrid.assemble(ctx, length, cells[0].num_levels, True, [p[1:] for p in parts])
    cont.setFilterSize(3)
    print("lookups", int(cont.lookup_ndc(pts).sum()))
    cont.close()
    w.close()
print("stats", ctx.stats())
