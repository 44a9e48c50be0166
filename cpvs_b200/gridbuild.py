"""Whole tile-grid builds on the GPUs of one box: BASELINE configs[2] (64K^2 virtual map, 4x4x4 cells) and
configs[4] (256K^2 virtual map, 16x16x16 cells, city occluders).

What the reference does per frame-of-precomputation in ``DeferredRenderer::renderWithTiles`` /
``createShadowTiles`` / ``precomputeShadows`` (reference ``src/DeferredRenderer.cpp:150-235``): for every xy
tile render a depth map, build one MinMaxHierarchy and ``length`` z-slice DAGs from it, drop them into the
cubic container, then ``moveToGPU``. Here every rank owns a rotated round-robin share of the xy tiles
(``cpvs_b200.tiling.tiles_of_rank``), produces each depth tile in device memory (CUDA generator for the libm-free scenes,
host generator + copy otherwise), builds its cells, and only the per-cell sizes cross ranks on the host.
For lookups the finished DAG words are then replicated to every GPU (NCCL broadcast over NVLink -- after the
build, not on its data path) and the query batch is split by screen rows.

Timing is on the device (CUDA events on the context's stream); the figure of a multi-rank run is the maximum
over ranks. No CPU fallback: everything below drives the CUDA library.
"""
import time

import numpy as np

from . import tiling


class _DeviceWords:
    """Zero-copy view of a finished DAG (device pointer owned by a CompressedShadow) for torch."""

    def __init__(self, ptr, words):
        self.__cuda_array_interface__ = {"shape": (int(words),), "typestr": "<i4", "data": (int(ptr), False), "version": 2}


def _local_expected(torch, pts, sel_path, depth, res):
    """z + 0.5 <= d * H straight from the depth tile (reference src/CompressedShadowUtil.h:47-57)."""
    n = depth.shape[0]
    d = depth[sel_path[:, 1] % n, sel_path[:, 0] % n]
    return ((sel_path[:, 2].to(torch.float32) + 0.5) <= d * float(res)).to(torch.uint8)


def run(ctx, stream, tile, length, kind, rank=0, world=1, dist=None, lookups=3840 * 2160, lookup_iters=8, verify=True,
        leafmasks=True, log=None, reserve_bytes=None):
    """Builds the ``length^3`` container of a ``(tile*length)^2`` virtual ``kind`` map and runs ``lookups`` random
    NDC lookups through it. Returns a dict of measurements (identical on every rank except the rank-local ones)."""
    import torch
    import cpvs_b200
    from cpvs_b200 import synth

    dev = torch.device("cuda", ctx.device)
    n, res = tile, tile * length
    if res > 2 ** 23:
        raise ValueError("virtual z resolution %d exceeds 2^23" % res)
    on_device = kind in cpvs_b200.SCENES
    mine = tiling.tiles_of_rank(length, rank, world)
    # Depth tiles. Device-generated scenes are produced tile by tile into one buffer. Host-generated ones (terrain:
    # host libm) are ALL produced and copied before the first timed build -- the metric starts from "depth resident
    # in device memory", and another rank's generator threads would otherwise steal the CPU from this rank's builds.
    slots = 1 if on_device else max(1, len(mine))
    depth_all = torch.empty((slots, n, n), dtype=torch.float32, device=dev)
    host = None if on_device else torch.empty((n, n), dtype=torch.float32, pin_memory=True)
    resident = {}

    def produce(xy):
        """Depth tile `xy` in device memory -> (tensor, ms spent producing it)."""
        if xy in resident:
            return resident[xy], 0.0
        t0 = time.perf_counter()
        if on_device:
            buf = depth_all[0]
            cpvs_b200.generate_depth(kind, n, buf, xy, length, ctx)
        else:
            buf = depth_all[len(resident)]
            synth.depth_map(kind, n, xy, length, out=host.numpy())
            buf.copy_(host, non_blocking=True)
            resident[xy] = buf
        torch.cuda.synchronize(dev)
        return buf, (time.perf_counter() - t0) * 1e3

    produce_ms = 0.0
    if not on_device:
        for xy in mine:
            produce_ms += produce(xy)[1]
    if reserve_bytes is None:
        # kept DAG words (the largest tile seen so far, 16K^2 terrain, keeps 243 MB) + room for the scratch arena to
        # regrow inside the pool (4 GB covers the 16K^2 scenes of the survey)
        scale = (n / 16384.0) ** 2
        reserve_bytes = int(min(len(mine) * 320e6 * scale, 8 * 2.0 ** 30) + 4e9 * scale)
    ctx.reserve(reserve_bytes)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cells = {}  # cell index -> CompressedShadow (kept: the DAG words stay in HBM)
    build_ms, tile_ms = 0.0, []
    svo_nodes = np.zeros(32, np.int64)
    dag_nodes = np.zeros(32, np.int64)
    checked = 0
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    if world > 1:
        dist.barrier()  # nobody starts its timed builds while another rank still generates depth tiles on the host
        torch.cuda.synchronize(dev)
    if mine:  # untimed warm-up on the first owned tile, right before the timed builds: module load, scratch arena
        depth, _ = produce(mine[0])  # growth, and clocks back up after the wait at the barrier
        for _ in range(2):
            mm = cpvs_b200.MinMaxHierarchy(depth, ctx, n=n)
            for z in range(length):
                cpvs_b200.CompressedShadow.create(mm, z, length, leafmasks).close()
            mm.close()
        ctx.synchronize()
    launches0 = ctx.launch_count
    wall0 = time.perf_counter()
    for (x, y) in mine:
        depth, ms = produce((x, y))
        produce_ms += ms
        ev0.record(stream)
        mm = cpvs_b200.MinMaxHierarchy(depth, ctx, n=n)
        column = [cpvs_b200.CompressedShadow.create(mm, z, length, leafmasks) for z in range(length)]
        ev1.record(stream)
        torch.cuda.synchronize(dev)
        ms = ev0.elapsed_time(ev1)
        build_ms += ms
        tile_ms.append(ms)
        for z, sh in enumerate(column):
            cells[tiling.cell_index(x, y, z, length)] = sh
            s, d, _ = sh.level_counts()
            svo_nodes[: len(s)] += s.astype(np.int64)
            dag_nodes[: len(d)] += d.astype(np.int64)
            if verify:  # every cell against its depth tile, on the device
                pts = torch.rand((1 << 16, 3), generator=gen, device=dev, dtype=torch.float32) * 2 - 1
                out = torch.empty(pts.shape[0], dtype=torch.uint8, device=dev)
                torch.cuda.synchronize(dev)
                sh.traverse(pts, True, out)
                ctx.synchronize()
                path = (((pts + 1.0) * 0.5) * float(n - 1)).to(torch.int32)
                path[:, 2] += z * n
                if not torch.equal(out, _local_expected(torch, pts, path, depth, res)):
                    raise RuntimeError("cell (%d,%d,%d): lookups do not decode to the depth tile" % (x, y, z))
                checked += pts.shape[0]
        mm.close()
        if log:
            log("rank %d tile (%d,%d): build %.2f ms" % (rank, x, y, ms))
    wall_ms = (time.perf_counter() - wall0) * 1e3
    launches = ctx.launch_count - launches0

    # ---- host-side gather of sizes (the only cross-rank step of the build) -------------------------
    t0 = time.perf_counter()
    root_of = {cpvs_b200.SHADOW: 0x0000, cpvs_b200.VISIBLE: 0x5555, cpvs_b200.PARTIAL: 0xAAAA}  # what the grid scan tests
    sizes = {i: (int(sh.info.words), int(sh.info.num_levels), root_of[sh.getTotalVisibility()]) for i, sh in cells.items()}
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, sizes)
        stats = [None] * world
        dist.all_gather_object(stats, (build_ms, produce_ms, wall_ms, svo_nodes, dag_nodes, checked, launches))
    else:
        parts, stats = [sizes], [(build_ms, produce_ms, wall_ms, svo_nodes, dag_nodes, checked, launches)]
    all_sizes = {}
    for p in parts:
        all_sizes.update(p)
    assert len(all_sizes) == length ** 3, "some cell was never built"
    ordered = [all_sizes[i] for i in range(length ** 3)]
    grid, total_words = tiling.top_level_grid([(w, m) for (w, _, m) in ordered], length)
    gather_ms = (time.perf_counter() - t0) * 1e3

    # ---- replicate the DAG words for lookups --------------------------------------------------------
    t0 = time.perf_counter()
    cont = cpvs_b200.CompressedShadowContainer(length, ctx)
    if world == 1:
        for i, sh in cells.items():
            z, rem = divmod(i, length * length)
            y, x = divmod(rem, length)
            cont.set(sh, x, y, z)
    else:
        packed = []
        for r in range(world):
            idx = sorted(parts[r])
            buf = torch.empty(max(1, sum(parts[r][i][0] for i in idx)), dtype=torch.int32, device=dev)
            if r == rank:
                off = 0
                with torch.cuda.stream(stream):
                    for i in idx:
                        w = parts[r][i][0]
                        buf[off:off + w].copy_(torch.as_tensor(_DeviceWords(cells[i].dag_device_ptr, w), device=dev))
                        off += w
                torch.cuda.synchronize(dev)
            dist.broadcast(buf, src=r)
            packed.append((idx, buf))
        torch.cuda.synchronize(dev)
        for r, (idx, buf) in enumerate(packed):
            off = 0
            for i in idx:
                w, levels, _ = parts[r][i]
                z, rem = divmod(i, length * length)
                y, x = divmod(rem, length)
                cont.set_dag(buf[off:off + w], levels, leafmasks, x, y, z)
                off += w
        del packed
    cont.copyToGPU()
    ctx.synchronize()
    assemble_ms = (time.perf_counter() - t0) * 1e3
    info = cont.info()
    assert info["dag_words"] == total_words, (info, total_words)
    if total_words <= (1 << 26):  # small enough to read back: the device grid must be the host scan's
        _, cgrid = cont.dag_and_grid()
        assert np.array_equal(cgrid, grid), "top-level grid differs from the host-side scan"
    for sh in cells.values():
        sh.close()
    cells.clear()

    # ---- lookups: the batch is split by screen rows -------------------------------------------------
    width = 3840
    rows = max(1, lookups // width)
    lo, hi = rows * rank // world, rows * (rank + 1) // world
    pts_all = synth.lookups(rows * width)  # the same points on every rank (seed 777)
    pts = torch.from_numpy(pts_all[lo * width:hi * width]).to(dev)
    out = torch.empty(pts.shape[0], dtype=torch.uint8, device=dev)
    torch.cuda.synchronize(dev)
    for _ in range(3):
        cont.lookup_ndc(pts, out)
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    l0.record(stream)
    for _ in range(lookup_iters):
        cont.lookup_ndc(pts, out)
    l1.record(stream)
    torch.cuda.synchronize(dev)
    lookup_ms = l0.elapsed_time(l1) / lookup_iters
    lit_local = int(out.sum().item())

    # ---- container lookups against the depth tiles, every point checked by the rank that owns its tile
    # (the container on this rank holds ALL cells: this checks the replicated words and the grid) ----
    verified = 0
    if verify:
        allp = torch.from_numpy(pts_all).to(dev)
        res_all = torch.empty(allp.shape[0], dtype=torch.uint8, device=dev)
        torch.cuda.synchronize(dev)
        cont.lookup_ndc(allp, res_all)
        ctx.synchronize()
        path = (((allp + 1.0) * 0.5) * float(res - 1)).to(torch.int32)
        for (x, y) in mine:
            sel = ((path[:, 0] // n) == x) & ((path[:, 1] // n) == y)
            if not bool(sel.any()):
                continue
            depth, _ = produce((x, y))
            if not torch.equal(res_all[sel], _local_expected(torch, allp[sel], path[sel], depth, res)):
                raise RuntimeError("container lookups over tile (%d,%d) do not decode to its depth" % (x, y))
            verified += int(sel.sum().item())
        del allp, res_all, path

    agg = torch.tensor([build_ms, lookup_ms, wall_ms], dtype=torch.float64, device=dev)
    cnt = torch.tensor([lit_local, verified, checked], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    build_max, lookup_max, wall_max = [float(v) for v in agg.tolist()]
    lit, verified_all, checked_all = [int(v) for v in cnt.tolist()]
    svo_total = sum(s[3] for s in stats)
    dag_total = sum(s[4] for s in stats)
    levels = int(np.log2(n)) + 1
    trivial = sum(1 for (w, _, _) in ordered if w == 1)
    result = {
        "virtual_side": res, "tile": n, "length": length, "kind": kind, "leafmasks": bool(leafmasks), "n_gpus": world,
        "depth_source": "device generator (cpvs_depth_generate)" if on_device else "host generator + H2D copy; all owned tiles resident before the timed builds",
        "xy_tiles": length * length, "cells": length ** 3, "one_word_cells": trivial,
        "samples": res * res,
        "build_ms_max_rank": build_max, "build_ms_per_rank": [s[0] for s in stats],
        "build_msamples_per_s": res * res / (build_max * 1e-3) / 1e6,
        "tile_ms_mean_rank0": float(np.mean(tile_ms)) if tile_ms else None,
        "depth_produce_ms_per_rank": [s[1] for s in stats], "wall_ms_max_rank": wall_max,
        "gather_sizes_ms": gather_ms, "replicate_and_finalize_ms": assemble_ms,
        "dag_words": int(total_words), "dag_mbytes": 4.0 * total_words / 1e6,
        "svo_nodes_per_level": {str(l): int(svo_total[l]) for l in range(levels - 2, 1 if leafmasks else -1, -1)},
        "dag_nodes_per_level": {str(l): int(dag_total[l]) for l in range(levels - 2, 1 if leafmasks else -1, -1)},
        "grid_cells_with_dag": int(((grid != tiling.GRID_CELL_SHADOWED) & (grid != tiling.GRID_CELL_VISIBLE)).sum()),
        "lookups": rows * width, "lookup_ms_max_rank": lookup_max, "lookups_g_per_s": rows * width / (lookup_max * 1e-3) / 1e9,
        "lookups_lit": lit, "gpu_launches": int(sum(s[6] for s in stats)),
        "verified": {"cell_lookups_vs_depth": checked_all, "container_lookups_vs_depth": verified_all} if verify else None,
    }
    cont.close()
    return result
