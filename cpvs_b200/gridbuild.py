"""Whole tile-grid builds with ONE PROCESS PER GPU (torchrun): BASELINE configs[2] (64K^2 virtual map, 4x4x4 cells) and
configs[4] (256K^2 virtual map, 16x16x16 cells, city occluders).

What the reference does per precomputation in ``DeferredRenderer::renderWithTiles`` / ``createShadowTiles`` /
``precomputeShadows`` (reference ``src/DeferredRenderer.cpp:150-235``): for every xy tile render a depth map, build one
MinMaxHierarchy and ``length`` z-slice DAGs from it, drop them into the cubic container, then ``moveToGPU``. The work is done
by the C++ grid worker of the CUDA library (``cpvs_grid_worker_*``, csrc/grid.cu) -- the same one ``cpvs_grid_build`` runs on
one host thread per GPU inside a single process; here every rank drives the worker of its own GPU:

  1. ownership. Few tiles per GPU: cost of the xy tiles (closed-form node counts over each tile resampled at 1/8 resolution)
     for a rotated round-robin share, gathered on the host (gloo), then assigned longest first (``cpvs_grid_assign``). Many
     tiles per GPU: every rank pulls its next tile from a shared counter on the rendezvous store.
  2. every rank builds its tiles: depth tile generated on its GPU (or handed over by the caller), hierarchy, cells.
  3. host-side gather (gloo) of (words, root mask) per cell -- the only thing that crosses ranks for the build --, the
     exclusive scan of ``createTopLevelGrid`` (reference ``src/CompressedShadowContainer.cpp:71-91``) on the host.
  4. for lookups every rank writes its finished words to a file in the host's shared memory, reads the others', and puts its
     own container together (``cpvs_container_assemble``); the query batch is split by rows. (The single-process driver
     ``cpvs_grid_build`` replicates with cudaMemcpyPeerAsync instead.)

No NCCL: ``group`` is any torch.distributed group with a CPU (gloo) backend, or None for a single process. Timing is on the
device (CUDA events on the worker's stream); the figure of a multi-rank run is the maximum over ranks. No CPU fallback.
"""
import os
import time

import numpy as np

from . import grid as cgrid
from . import tiling


def _gather(group, obj, world):
    if world == 1:
        return [obj]
    import torch.distributed as dist
    out = [None] * world
    dist.all_gather_object(out, obj, group=group)
    return out


_QUEUE_RUNS = {}  # (every rank calls run() the same number of times: the counters' names agree)


def _barrier(group, world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier(group=group)


def run(ctx, tile, length, kind, rank=0, world=1, group=None, lookups=3840 * 2160, lookup_iters=32, verify=True, leafmasks=True,
        fetch=None, replicate=True, log=None):
    """Builds the ``length^3`` container of a ``(tile*length)^2`` virtual map and runs ``lookups`` random NDC lookups through
    it. ``kind``: a scene with a device generator; or ``fetch(x, y, out)`` for caller-provided depth tiles. Returns a dict of
    measurements (identical on every rank except the rank-local ones)."""
    import cpvs_b200
    from cpvs_b200 import synth

    n, res = tile, tile * length
    tiles = tiling.xy_tiles(length)
    start = [(y * length + (x + y) % length) % world for (x, y) in tiles]  # rotated round-robin (tiling.tiles_of_rank)
    worker = cgrid.GridWorker(ctx, length, tile, kind if fetch is None else None, fetch, leafmasks)
    scale = (n / 16384.0) ** 2
    ctx.reserve(int(min(len(tiles) / world * 320e6 * scale, 8 * 2.0 ** 30) + 6e9 * scale))

    # untimed warm-up: one tile of this rank's share (module load, scratch arena growth, size memos, clocks)
    mine0 = [t for t, o in zip(tiles, start) if o == rank]
    if mine0:
        warm = cgrid.GridWorker(ctx, length, tile, kind if fetch is None else None, fetch, leafmasks)
        warm.build(mine0[:1])
        warm.close()
    ctx.synchronize()
    _barrier(group, world)
    launches0 = ctx.launch_count
    wall0 = time.perf_counter()

    # 1. ownership + 2. build. With many tiles per GPU every rank pulls its next tile from a shared counter (an atomic add on the
    # rendezvous store -- a host-side integer, nothing on the data path), which evens out tiles of very different cost (the
    # 256K^2 city: box edges) without knowing the costs. With few tiles per GPU the tiles are costed first (each rank a
    # round-robin share, gathered on the host) and assigned longest first.
    cost_aware = world > 1 and len(tiles) <= 4 * world
    moved = 0
    if cost_aware:
        # few tiles per GPU: ownership by cost, longest first, decided before anybody builds (no traffic on the store while
        # the few milliseconds of builds run)
        costs_mine = dict(zip(mine0, worker.estimate(mine0)))
        costs = {}
        for part in _gather(group, costs_mine, world):
            costs.update(part)
        owners = cgrid.assign([costs[t] for t in tiles], world, start)
        mine = [t for t, o in zip(tiles, owners) if o == rank]
        worker.release([t for t in mine0 if t not in mine])
        worker.build(mine)
    elif world > 1:
        import torch.distributed as dist
        store = dist.distributed_c10d._get_default_store()
        key = "cpvs_grid_queue_%d_%d_%s" % (length, tile, kind)
        _QUEUE_RUNS[key] = _QUEUE_RUNS.get(key, 0) + 1
        key += "_%d" % _QUEUE_RUNS[key]
        mine = []

        def next_tile():
            k = store.add(key, 1) - 1
            if k >= len(tiles):
                return None
            mine.append(tiles[k])
            return tiles[k]

        worker.build_from(next_tile)
    else:
        mine = list(tiles)
        worker.build(mine)
    build_ms = worker.device_ms()
    depth_ms = worker.depth_ms()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    launches = ctx.launch_count - launches0
    if log:
        log("rank %d: %d tiles, %.2f ms on the device" % (rank, len(mine), build_ms))

    # 3. host-side gather of sizes
    t0 = time.perf_counter()
    cells = worker.cells()
    offsets = worker.copy_cells(None)
    my_words = (offsets[-1] + int(cells[-1].words)) if cells else 0
    mine_info = {"rank": rank, "device": ctx.device, "build_ms": build_ms, "depth_ms": depth_ms, "wall_ms": wall_ms, "launches": launches,
                 "tiles": len(mine), "words": my_words, "moved": sum(1 for t in mine if t not in mine0),
                 "cells": [(c.index, int(c.words), int(c.root_mask), int(c.num_levels), int(off), int(c.svo_nodes), int(c.dag_nodes))
                           for c, off in zip(cells, offsets)]}
    everyone = _gather(group, mine_info, world)
    moved = sum(i["moved"] for i in everyone)
    table = {}
    for info in everyone:
        for (index, words, mask, levels, off, svo, dagn) in info["cells"]:
            table[index] = (words, mask, levels, off, info["rank"], svo, dagn)
    assert len(table) == length ** 3, "some cell was never built"
    ordered = [table[i] for i in range(length ** 3)]
    grid_host, total_words = tiling.top_level_grid([(w, m) for (w, m, *_rest) in ordered], length)
    gather_ms = (time.perf_counter() - t0) * 1e3

    # 4. replication for the lookups (after the build, not on its data path). One process per GPU: the finished words go
    # through files in the host's shared memory -- every rank writes its cells once and reads the others' -- because mapping
    # eight processes' device memory into each other (CUDA IPC + peer access) costs seconds of set-up per process.
    t0 = time.perf_counter()
    cont = None
    if replicate:
        import torch
        local = {c.index: c for c in cells}
        staged = {}
        if world > 1:
            tag = "/dev/shm/cpvs_grid_%s_%d" % (os.environ.get("MASTER_PORT", "0"), _QUEUE_RUNS.get("shm", 0))
            _QUEUE_RUNS["shm"] = _QUEUE_RUNS.get("shm", 0) + 1
            mine_file = np.lib.format.open_memmap("%s_%d.npy" % (tag, rank), mode="w+", dtype=np.uint32, shape=(max(1, my_words),))
            worker.copy_cells(mine_file)
            mine_file.flush()
            _barrier(group, world)
            dev = torch.device("cuda", ctx.device)
            for info in everyone:
                if info["rank"] != rank and info["words"]:
                    words = np.load("%s_%d.npy" % (tag, info["rank"]), mmap_mode="r")
                    staged[info["rank"]] = torch.from_numpy(np.ascontiguousarray(words).view(np.int32)).to(dev)
            torch.cuda.synchronize(dev)
        parts = []
        for i, (words, mask, levels, off, owner, _svo, _dagn) in enumerate(ordered):
            if owner == rank:
                parts.append((words, mask, ctx.device, local[i].words_device))
            elif words == 1 and mask in (0, 0x5555):
                parts.append((words, mask, ctx.device, 0))
            else:
                parts.append((words, mask, ctx.device, staged[owner].data_ptr() + 4 * off))
        cont = cgrid.assemble(ctx, length, ordered[0][2], leafmasks, parts)
        ctx.synchronize()
        staged.clear()
        if world > 1:
            _barrier(group, world)  # nobody removes its file while a peer still reads it
            del mine_file
            try:
                os.remove("%s_%d.npy" % (tag, rank))
            except OSError:
                pass
    assemble_ms = (time.perf_counter() - t0) * 1e3

    result = {
        "virtual_side": res, "tile": n, "length": length, "kind": kind if fetch is None else "caller-provided tiles", "leafmasks": bool(leafmasks),
        "n_gpus": world, "depth_source": "device generator (cpvs_depth_generate)" if fetch is None else "host callback + H2D copy",
        "xy_tiles": length * length, "cells": length ** 3, "one_word_cells": sum(1 for c in ordered if c[0] == 1), "samples": res * res,
        "ownership": ("by cost (closed-form node counts of the tiles at 1/8 resolution), longest first" if cost_aware
                      else "shared queue (atomic counter on the rendezvous store)" if world > 1 else "single GPU"), "moved_tiles": moved,
        "build_ms_max_rank": max(i["build_ms"] for i in everyone), "build_ms_per_rank": [i["build_ms"] for i in everyone],
        "depth_ms_per_rank": [i["depth_ms"] for i in everyone], "timing": "device time from depth tiles resident in device memory (SURVEY.md 8d); "
                                                                          "producing them is clocked separately (depth_ms_per_rank)",
        "tiles_per_rank": [i["tiles"] for i in everyone], "wall_ms_max_rank": max(i["wall_ms"] for i in everyone),
        "gather_sizes_ms": gather_ms, "replicate_and_finalize_ms": assemble_ms,
        "dag_words": int(total_words), "dag_mbytes": 4.0 * total_words / 1e6,
        "svo_nodes": int(sum(c[5] for c in ordered)), "dag_nodes": int(sum(c[6] for c in ordered)),
        "grid_cells_with_dag": int(((grid_host != tiling.GRID_CELL_SHADOWED) & (grid_host != tiling.GRID_CELL_VISIBLE)).sum()),
        "gpu_launches": int(sum(i["launches"] for i in everyone)),
    }
    result["build_msamples_per_s"] = res * res / (result["build_ms_max_rank"] * 1e-3) / 1e6

    if cont is not None:
        info = cont.info()
        assert info["dag_words"] == total_words, (info, total_words)
        if total_words <= (1 << 26):  # small enough to read back: the device grid must be the host scan's
            _, cgrid_dev = cont.dag_and_grid()
            assert np.array_equal(cgrid_dev, grid_host), "top-level grid differs from the host-side scan"
        # lookups: the batch is split by rows
        import torch
        dev = torch.device("cuda", ctx.device)
        width = 3840
        rows = max(1, lookups // width)
        lo, hi = rows * rank // world, rows * (rank + 1) // world
        pts_all = synth.lookups(rows * width)  # the same points on every rank (seed 777)
        pts = torch.from_numpy(pts_all[lo * width:hi * width]).to(dev)
        out = torch.empty(pts.shape[0], dtype=torch.uint8, device=dev)
        torch.cuda.synchronize(dev)
        for _ in range(3):
            cont.lookup_ndc(pts, out)
        ctx.synchronize()
        _barrier(group, world)
        # (wall clock around back-to-back launches: at eight ranks a rank's share is a 15 us kernel, so many launches per reading)
        t0 = time.perf_counter()
        for _ in range(lookup_iters):
            cont.lookup_ndc(pts, out)
        ctx.synchronize()
        lookup_ms = (time.perf_counter() - t0) * 1e3 / lookup_iters
        lit_local = int(out.sum().item())
        verified = 0
        if verify and fetch is None:
            # every point against the depth tile it falls into, checked by the rank that owns the tile -- on the container
            # this rank assembled from everybody's words
            allp = torch.from_numpy(pts_all).to(dev)
            res_all = torch.empty(allp.shape[0], dtype=torch.uint8, device=dev)
            torch.cuda.synchronize(dev)
            cont.lookup_ndc(allp, res_all)
            ctx.synchronize()
            path = (((allp + 1.0) * 0.5) * float(res - 1)).to(torch.int32)
            depth = torch.empty((n, n), dtype=torch.float32, device=dev)
            for (x, y) in mine:
                sel = ((path[:, 0] // n) == x) & ((path[:, 1] // n) == y)
                if not bool(sel.any()):
                    continue
                cpvs_b200.generate_depth(kind, n, depth, (x, y), length, ctx)
                ctx.synchronize()
                p = path[sel]
                d = depth[p[:, 1] % n, p[:, 0] % n]
                want = ((p[:, 2].to(torch.float32) + 0.5) <= d * float(res)).to(torch.uint8)
                if not torch.equal(res_all[sel], want):
                    raise RuntimeError("container lookups over tile (%d,%d) do not decode to its depth" % (x, y))
                verified += int(sel.sum().item())
        parts_l = _gather(group, (lookup_ms, lit_local, verified), world)
        lookup_max = max(p[0] for p in parts_l)
        result.update({"lookups": rows * width, "lookup_ms_max_rank": lookup_max, "lookups_g_per_s": rows * width / (lookup_max * 1e-3) / 1e9,
                       "lookups_lit": sum(p[1] for p in parts_l),
                       "verified": {"container_lookups_vs_depth": sum(p[2] for p in parts_l)} if verify else None})
        cont.close()
    worker.close()
    return result
