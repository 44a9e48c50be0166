#!/usr/bin/env python
"""bench.py -- headline benchmark of the cpvs hot path on B200 (BASELINE.json: DAG build Msamples/s at
16K^2, shadow lookups G/s, % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl own|reference]

One step = one pass of the hot path over one depth map: MinMaxHierarchy + CompressedShadow::create on the device (BASELINE
configs[1]: one 16K^2 terrain map, leafmasks on, single DAG). At N > 1 every rank runs the same unit on its own map (weak
scaling, no collective on the data path). `value` is throughput with the depth map resident in HBM and two builds in flight
through the public API (cpvs_shadow_create_async on two contexts); `one_build_at_a_time` is the same loop with plain
synchronous calls; `e2e` goes through the same C-ABI calls with the depth map in pinned host memory. Next to it, in the same
line: 1 M random lookups, BASELINE configs[3] (4K G-buffer, leafmasks on vs off), and the tiled configs built whole --
configs[2] (64K^2, 4x4x4 cells) and configs[4] (256K^2, 16x16x16 cells), xy tiles sharded over the N ranks (strong scaling).
`--impl reference` times the reference's own CPU implementation (oracle/_ref, the unmodified sources compiled by
oracle/Makefile) on bounded samples of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "dag_build_msamples_per_s"
UNIT = "Msamples/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--size", type=int, default=16384, help="side of one depth map")
    ap.add_argument("--kind", default="terrain", choices=["plane", "terrain", "terrain_dev", "city"])
    ap.add_argument("--lookups", type=int, default=1000000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="do not run the nvidia-smi sampler (debugging)")
    ap.add_argument("--ref-sample", type=int, default=1024, help="side of one reference sample window")
    ap.add_argument("--contexts", type=int, default=2, help="contexts of the GPU the timed loops alternate between")
    ap.add_argument("--in-flight", type=int, default=2,
                    help="builds in flight in the timed loops (cpvs_shadow_create_async on two contexts of the GPU); 0 = one synchronous build at a time")
    ap.add_argument("--no-configs3", action="store_true", help="skip the leafmasks on/off comparison (BASELINE configs[3])")
    ap.add_argument("--no-grids", action="store_true", help="skip the whole tile grids (BASELINE configs[2] and configs[4])")
    ap.add_argument("--no-port-16k", action="store_true", help="skip the single-thread oracle port on the whole bench map (about 30 s)")
    ap.add_argument("--grid", default=None, choices=["64k", "256k"],
                    help="whole tile-grid build instead of the step bench: 64k = BASELINE configs[2] (4x4x4 cells of 16K^2 terrain "
                         "tiles), 256k = configs[4] (16x16x16 cells of 16K^2 city tiles); tiles are sharded over --gpus ranks")
    ap.add_argument("--grid-tile", type=int, default=16384, help="side of one depth tile of --grid (reduce for a quick run)")
    ap.add_argument("--grid-passes", type=int, default=2, help="--grid builds the grid this many times and reports the last pass")
    ap.add_argument("--grid-length", type=int, default=None, help="override the grid length of --grid")
    ap.add_argument("--grid-kind", default=None, choices=["plane", "terrain_dev", "city"])
    ap.add_argument("--no-verify", action="store_true", help="--grid: skip the lookups-decode-to-depth checks")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML while the timed region runs (the same counters
    as the nvidia-smi line of B200_PROFILING.md; NVML in-process because a concurrently starting
    nvidia-smi process stalls CUDA API calls of the timed steps for tens of milliseconds)."""

    def __init__(self, device):
        self.device = device
        self.samples = []
        self.reasons = set()
        self.stop_flag = threading.Event()
        self.thread = None
        self.handle = None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = int(visible.split(",")[device]) if visible and visible.split(",")[device].isdigit() else device
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception as exc:  # noqa: BLE001
            self.error = str(exc)
            self.handle = None

    def _poll(self):
        nv = self.nvml
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake_slowdown"}
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self.stop_flag.wait(0.004)

    def start(self):
        if self.handle is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % getattr(self, "error", "not started")]}
        self.stop_flag.set()
        self.thread.join()
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.sm_max,
                "samples": len(self.samples), "reasons": sorted(self.reasons), "source": "nvml"}


# ---- reference arm / cpu baseline -----------------------------------------------------------------

def reference_sample(size, kind, window, threads, steps, warmup):
    """Unmodified reference (oracle/_ref) on `threads` independent windows of the workload per step, one
    host thread each (the reference's create is single-threaded; its tile driver runs one create per
    thread, src/DeferredRenderer.cpp:150-163). Returns (Msamples/s, ms per step, description, kind)."""
    from oracle import pyoracle as O
    from cpvs_b200 import synth
    kind_used = "reference" if O.have_ref() else "port"
    tiles = size // window
    maps = [synth.depth_map(kind, window, (t % tiles, (t // tiles) % tiles), tiles, threads=1) for t in range(threads)]

    def one(d):
        if kind_used == "reference":
            O.ref_time_build(d)
        else:
            O.Shadow(O.MinMax(d, "port")).dag()

    def step():
        ths = [threading.Thread(target=one, args=(m,)) for m in maps]
        for t in ths:
            t.start()
        for t in ths:
            t.join()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    samples = threads * window * window * steps
    desc = ("%d windows of %dx%d texels of the %dx%d %s map per step, one %s MinMaxHierarchy+CompressedShadow::create per host thread"
            % (threads, window, window, size, size, kind, "reference" if kind_used == "reference" else "oracle-port"))
    return samples / dt / 1e6, dt / steps * 1e3, desc, kind_used


def reference_lookups(size, kind, window, threads, count):
    """CompressedShadow::traverse of the reference (or the port) on `count` random NDC points against the DAG of one
    window: (M lookups/s on 1 thread, M lookups/s on `threads` threads)."""
    from oracle import pyoracle as O
    from cpvs_b200 import synth
    tiles = size // window
    d = synth.depth_map(kind, window, (0, 0), tiles, threads=1)
    backend = "ref" if O.have_ref() else "port"
    sh = O.Shadow(O.MinMax(d, backend))
    pts = synth.lookups(count)
    t0 = time.perf_counter()
    sh.traverse(pts)
    one = count / (time.perf_counter() - t0) / 1e6
    t0 = time.perf_counter()
    sh.traverse(pts, threads=threads)
    many = count / (time.perf_counter() - t0) / 1e6
    return one, many


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 64))
    value, ms, desc, kind_used = reference_sample(args.size, args.kind, args.ref_sample, threads, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32/u32", "data": "synthetic", "config": workload_config(args, args.gpus, reference=True),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind_used, "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, n_gpus, reference=False):
    cfg = {"workload": "configs[1]: %dx%d synthetic %s depth map, leafmasks on, single DAG (MinMaxHierarchy + CompressedShadow::create) "
                       "+ %d random NDC lookups" % (args.size, args.size, args.kind, args.lookups),
           "depth_map": "%dx%d f32 (%.0f MiB) > L2: every build streams it from HBM; no explicit L2 flush needed" % (args.size, args.size, args.size * args.size * 4 / 2**20),
           "tiles_per_rank": 1, "z_slices": 1}
    if n_gpus > 1:
        cfg["sharding"] = ("rank r builds the %dx%d map at offset (r %% 4, r // 4) of the same scene: the unit of work of N=1 on different data, "
                           "no collective on the data path" % (args.size, args.size))
    if reference:
        cfg["proxy"] = ("the reference's create is O(nodes x unique nodes) per level: the whole %dx%d map would take hours, so every step builds "
                        "`cores` independent %dx%d windows of it, one per host thread -- less work per sample than the single DAG the GPU arm "
                        "builds, i.e. a ratio against this line understates the speed-up" % (args.size, args.size, args.ref_sample, args.ref_sample))
    else:
        cfg["in_flight"] = ("%d builds in flight (cpvs_shadow_create_async, two contexts of the GPU)" % args.in_flight) if args.in_flight else "one synchronous build at a time"
    return cfg


# ---- whole tile grids (BASELINE configs[2] and configs[4]) ---------------------------------------

GRIDS = {"64k": (2, 4, "terrain_dev"), "256k": (4, 16, "city")}  # name -> (BASELINE configs index, grid length, scene)


def grid_line(res, cfg_index):
    """The part of a grid result that goes into the bench line."""
    keep = ("virtual_side", "tile", "length", "kind", "n_gpus", "cells", "one_word_cells", "ownership", "moved_tiles", "build_ms_max_rank",
            "build_ms_per_rank", "depth_ms_per_rank", "timing", "tiles_per_rank", "wall_ms_max_rank", "gather_sizes_ms", "replicate_and_finalize_ms", "dag_words", "dag_mbytes",
            "svo_nodes", "dag_nodes", "gpu_launches", "lookups", "lookups_g_per_s", "verified", "depth_source")
    out = {k: res[k] for k in keep if k in res}
    out["workload"] = ("configs[%d]: %dK^2 virtual %s shadow map as a %dx%dx%d CompressedShadowContainer of %dx%d depth tiles generated on the "
                       "owning GPU, xy tiles sharded over %d GPU(s), host-side gather of sizes only (no NCCL)"
                       % (cfg_index, res["virtual_side"] // 1024, res["kind"], res["length"], res["length"], res["length"], res["tile"], res["tile"],
                          res["n_gpus"]))
    out["value"] = res["build_msamples_per_s"]
    out["unit"] = UNIT
    out["scaling"] = "strong"
    spread = res["build_ms_per_rank"]
    out["rank_time_spread"] = (max(spread) - min(spread)) / max(spread) if spread else 0.0
    return out


def run_grid(args):
    import torch
    import torch.distributed as dist
    import cpvs_b200
    from cpvs_b200 import build as cbuild, gridbuild

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback")
    torch.cuda.set_device(local)
    group = None
    if world > 1:
        dist.init_process_group("gloo")  # host-side gathers only
        group = dist.group.WORLD
    if rank == 0:
        cbuild.build()
    if world > 1:
        dist.barrier()
    cfg_index, length, kind = GRIDS[args.grid]
    length = args.grid_length or length
    kind = args.grid_kind or kind
    ctx = cpvs_b200.Context(local)
    sampler = ClockSampler(local)
    if not args.no_clocks:
        sampler.start()
    first = None
    for r in range(max(1, args.grid_passes)):  # the last pass is reported; the first one warms the contexts up (see configs[2]/[4] below)
        res = gridbuild.run(ctx, args.grid_tile, length, kind, rank, world, group, verify=not args.no_verify and r == 0,
                            log=(lambda m: print(m, file=sys.stderr, flush=True)) if os.environ.get("CPVS_GRID_LOG") else None)
        first = first or res
    res["first_pass_build_ms_max_rank"] = first["build_ms_max_rank"]
    clocks = sampler.stop()
    if rank == 0:
        line = {"metric": METRIC, "value": res["build_msamples_per_s"], "unit": UNIT, "n_gpus": world, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32/u32", "data": "synthetic",
                "config": {"workload": grid_line(res, cfg_index)["workload"]}, "ms_total": res["build_ms_max_rank"],
                "gpu_launches": res["gpu_launches"], "clocks": clocks, "grid": res}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---- own arm -----------------------------------------------------------------------------------------

def build_bytes(n, info, leaf):
    """SURVEY.md 8(d) algorithmic bytes of one build from the actual per-level node counts."""
    nl = int(info.num_levels)
    w_svo = w_merged = nodes = 0
    for lvl in range(nl - 1):
        size = 17 if (leaf and lvl == 2) else 9
        w_svo += size * int(info.svo_nodes[lvl])
        w_merged += size * int(info.dag_nodes[lvl])
        nodes += int(info.svo_nodes[lvl])
    return (40.0 / 3.0) * n * n + 8.0 * w_svo + 4.0 * w_merged + 8.0 * nodes + 4.0 * int(info.words)


def golden_digest(kind, n):
    """(words, fnv64) of the port's DAG for a whole-volume map of the bench, from tests/golden/port_large.json, or None."""
    path = os.path.join(ROOT, "tests", "golden", "port_large.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        for row in json.load(f)["maps"]:
            if row["kind"] == kind and row["n"] == n and row["tiles_per_side"] == 1 and row["z_num"] == 1:
                return row["words"], row["fnv64"]
    return None


class Pipeline:
    """Back-to-back builds through the public API with several in flight: cpvs_shadow_create_async on alternating contexts of
    the same GPU (the kernels of one build fill the gaps in the latency-bound phases of the other). depth == 0 is the plain
    synchronous call on one context."""

    def __init__(self, ctxs, n, in_flight):
        self.ctxs, self.n, self.in_flight = ctxs, n, in_flight
        self.flying = []
        self.timings = []
        self.k = 0

    def _finish(self, item):
        import cpvs_b200  # noqa: F401
        mm, sh = item
        info = sh.info  # waits
        self.timings.append((mm.timing(), sh.phase_ms(), float(info.build_ms), bool(info.predicted)))
        sh.close()
        mm.close()

    def step(self, src):
        import cpvs_b200
        ctx = self.ctxs[self.k % len(self.ctxs)] if self.in_flight else self.ctxs[0]
        self.k += 1
        mm = cpvs_b200.MinMaxHierarchy(src, ctx, n=self.n)
        sh = cpvs_b200.CompressedShadow.create(mm, wait=self.in_flight == 0)
        self.flying.append((mm, sh))
        while len(self.flying) > self.in_flight:
            self._finish(self.flying.pop(0))

    def drain(self):
        while self.flying:
            self._finish(self.flying.pop(0))
        for c in self.ctxs:
            c.synchronize()


def run_own(args):
    import torch
    import torch.distributed as dist
    import cpvs_b200
    from cpvs_b200 import build as cbuild, gridbuild, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback")
    torch.cuda.set_device(local)
    gloo = None
    if world > 1:
        # NCCL only carries the barriers and the max-over-ranks of the timing the bench contract asks for; everything the tile
        # grids exchange goes through the host (gloo).
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        gloo = dist.new_group(backend="gloo")
    if rank == 0:
        cbuild.build()
    if world > 1:
        dist.barrier()
    n, K, W = args.size, args.steps, args.warmup
    # two contexts on real (non-default) streams: the library enqueues on them and the CUDA events below are ordered against them
    streams = [torch.cuda.Stream() for _ in range(max(1, args.contexts))]
    ctxs = [cpvs_b200.Context(local, stream=s.cuda_stream) for s in streams]
    clock = torch.cuda.Stream()
    torch.cuda.set_stream(streams[0])
    ctx = ctxs[0]

    # workload: the same unit of work at every N (weak scaling): one 16K^2 depth map -> pyramid + one DAG. Rank r takes the map
    # whose origin is shifted by (r % 4, r // 4) maps in the same analytic scene at the same texel density: statistically the
    # work of rank 0's configs[1] map, different data.
    tile, tps = (rank % 4, (rank // 4) % 4), 1
    host = torch.empty((n, n), dtype=torch.float32, pin_memory=True)
    depth_np = host.numpy()
    synth.depth_map(args.kind, n, tile, tps, out=depth_np)
    depth = host.to("cuda", non_blocking=True)
    torch.cuda.synchronize()

    # parity on the bench workload itself: the words against the committed digest of the CPU port (rank 0's map), and every
    # looked-up voxel must decode to z + 0.5 <= d * H. These handles stay alive: the lookups below run on them.
    mm = cpvs_b200.MinMaxHierarchy(depth, ctx, n=n)
    main_shadow = cpvs_b200.CompressedShadow.create(mm)
    info0 = main_shadow.info
    leaf = bool(info0.leafmasks)
    pts_np = synth.lookups(args.lookups)
    path = (((pts_np + np.float32(1)) * np.float32(0.5)) * np.float32(n - 1)).astype(np.int32)
    lit = (path[:, 2].astype(np.float32) + np.float32(0.5)) <= depth_np[path[:, 1], path[:, 0]] * np.float32(n)
    if not np.array_equal(main_shadow.traverse(pts_np), lit.astype(np.uint8)):
        raise SystemExit("bench.py: lookup results do not decode to the depth map")
    digest = {"fnv64": "%016x" % synth.fnv64(main_shadow.getDAG()), "golden": None, "golden_match": None}
    want = golden_digest(args.kind, n) if rank == 0 else None
    if want:
        digest["golden"] = {"words": want[0], "fnv64": want[1], "source": "tests/golden/port_large.json (oracle port)"}
        digest["golden_match"] = bool(want[0] == int(info0.words) and want[1] == digest["fnv64"])
        if not digest["golden_match"]:
            raise SystemExit("bench.py: the %dx%d %s DAG differs from the oracle's digest (%s vs %s)" % (n, n, args.kind, digest["fnv64"], want[1]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(pipeline, src, steps):
        """Device time of `steps` builds: an event on a clock stream the contexts' streams wait for, another one that waits for them."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(clock)
        for s in streams:
            s.wait_event(ev0)
        t0 = time.perf_counter()
        for _ in range(steps):
            pipeline.step(src)
        pipeline.drain()
        for s in streams:
            e = torch.cuda.Event()
            e.record(s)
            clock.wait_event(e)
        ev1.record(clock)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        barrier()
        return ev0.elapsed_time(ev1), wall

    def over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    in_flight = max(0, args.in_flight)
    pipe = Pipeline(ctxs, n, in_flight)
    for _ in range(max(W, 3) + 2 * len(ctxs) + 2 * in_flight):  # also grows the memory pool to what this many builds in flight need
        pipe.step(depth)
    pipe.drain()

    # ---- timed: device-resident input ----
    sampler = ClockSampler(local)
    launches0 = sum(c.launch_count for c in ctxs)
    if not args.no_clocks:
        sampler.start()
    pipe.timings = []
    ms_total, _ = timed(pipe, depth, K)
    clocks = sampler.stop()
    launches = sum(c.launch_count for c in ctxs) - launches0
    timings = pipe.timings
    ms_step = over_ranks(ms_total) / K
    value = world * n * n / (ms_step * 1e-3) / 1e6

    # the same builds one at a time on one context (what a single cpvs_shadow_create call costs)
    solo = Pipeline(ctxs[:1], n, 0)
    for _ in range(2):
        solo.step(depth)
    solo.timings = []
    solo_total, _ = timed(solo, depth, K)
    solo_ms = over_ranks(solo_total) / K
    solo_timings = solo.timings

    # ---- timed: end to end from pinned host memory through the C ABI (H2D copy of the depth map inside, sizes read back) ----
    e2e_pipe = Pipeline(ctxs, n, in_flight)
    for _ in range(2 + in_flight):
        e2e_pipe.step(depth_np)
    e2e_pipe.drain()
    _, e2e_wall = timed(e2e_pipe, depth_np, K)
    e2e_ms = over_ranks(e2e_wall) / K
    e2e_value = world * n * n / (e2e_ms * 1e-3) / 1e6
    d2h = 256 * 8  # size scalars read back per create

    # ---- lookups on the resident DAG ----
    # 16 different batches (seeds 777..792) are cycled so that no iteration finds its points in L2
    torch.cuda.set_stream(streams[0])
    stream = streams[0]
    batches = [torch.from_numpy(pts_np if b == 0 else synth.lookups(args.lookups, seed=777 + b)).cuda() for b in range(16)]
    out = torch.empty(args.lookups, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    for i in range(max(W, 3)):
        main_shadow.traverse(batches[i % 16], True, out)
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    l0.record(stream)
    for i in range(4 * K):
        main_shadow.traverse(batches[i % 16], True, out)
    l1.record(stream)
    torch.cuda.synchronize()
    lookup_ms = l0.elapsed_time(l1) / (4 * K)
    t0 = time.perf_counter()
    for _ in range(K):
        main_shadow.traverse(pts_np)
    lookup_e2e_ms = (time.perf_counter() - t0) * 1e3 / K

    # ---- BASELINE configs[3]: a 4K G-buffer of world positions on / just off the surface (deepest descent), through
    # CompressedShadowContainer::evaluate (light transform + grid step + DAG descent), leafmasks on vs off ----
    configs3 = None
    if world == 1 and not args.no_configs3:
        gw, gh = 3840, 2160
        u = (np.arange(gw, dtype=np.float32) + np.float32(0.5)) / np.float32(gw)
        v = (np.arange(gh, dtype=np.float32) + np.float32(0.5)) / np.float32(gh)
        tex = depth_np[np.minimum((v * n).astype(np.int64), n - 1)[:, None], np.minimum((u * n).astype(np.int64), n - 1)[None, :]]
        eps = np.where((np.add.outer(np.arange(gh), np.arange(gw)) & 1) == 0, np.float32(1.5), np.float32(-1.5)) / np.float32(n)
        pos_np = np.empty((gh, gw, 4), np.float32)
        pos_np[..., 0] = (u * 2 - 1)[None, :]
        pos_np[..., 1] = (v * 2 - 1)[:, None]
        pos_np[..., 2] = (tex + eps) * 2 - 1
        pos_np[..., 3] = 1
        path3 = (((pos_np[..., :3] + np.float32(1)) * np.float32(0.5)) * np.float32(n - 1)).astype(np.int32)
        want3 = ((path3[..., 2].astype(np.float32) + np.float32(0.5)) <= depth_np[path3[..., 1], path3[..., 0]] * np.float32(n))
        frames = [torch.from_numpy(pos_np).cuda() for _ in range(4)]  # 4 x 133 MB, cycled: every frame comes from HBM
        vis = torch.zeros((gh, gw), dtype=torch.uint8, device="cuda")
        ident = np.eye(4, dtype=np.float32)
        configs3 = {"what": "3840x2160 rgba32f positions within 1.5 texels of the %dx%d %s surface, identity lightViewProj, through "
                            "CompressedShadowContainer::evaluate; 4 frames cycled (532 MB > L2); build = MinMaxHierarchy + create, synchronous" % (n, n, args.kind),
                    "pixels": gw * gh, "lit_fraction": float(want3.mean())}
        for label, use_leaf in (("leafmasks_on", True), ("leafmasks_off", False)):
            builds = []
            sh3 = None
            for rep in range(4):
                if sh3 is not None:
                    sh3.close()
                m3 = cpvs_b200.MinMaxHierarchy(depth, ctx, n=n)
                sh3 = cpvs_b200.CompressedShadow.create(m3, leafmasks=use_leaf)
                if rep:
                    builds.append(m3.timing()[0] + float(sh3.info.build_ms))
                m3.close()
            cont = cpvs_b200.CompressedShadowContainer(sh3, ctx)
            cont.copyToGPU()
            for i in range(3):
                cont.evaluate(frames[i % 4], ident, vis)
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            vis.zero_()
            s0.record(stream)
            for i in range(K):
                cont.evaluate(frames[i % 4], ident, vis)
            s1.record(stream)
            torch.cuda.synchronize()
            surf_ms = s0.elapsed_time(s1) / K
            if not np.array_equal(vis.cpu().numpy() != 0, want3):
                raise SystemExit("bench.py: evaluate() results do not decode to the depth map (%s)" % label)
            configs3[label] = {"build_ms": statistics.median(builds), "dag_words": int(sh3.info.words), "dag_mbytes": int(sh3.info.words) * 4 / 1e6,
                               "svo_nodes": int(sum(sh3.info.svo_nodes[:])), "evaluate_ms": surf_ms,
                               "evaluate_glookups_per_s": gw * gh / (surf_ms * 1e-3) / 1e9, "stream_bytes": gw * gh * 17}
            cont.close()
            sh3.close()
        del frames

    # ---- the tiled configs, whole: configs[2] (64K^2, 4x4x4) and configs[4] (256K^2, 16x16x16), tiles sharded over the ranks ----
    grids = {}
    if not args.no_grids:
        for name in ("64k", "256k"):
            cfg_index, length, kind = GRIDS[name]
            # Like the steps of the headline, a grid is built again and again (every frame the light moves): the first pass is the
            # warm-up -- the scratch arenas, staging buffers and size memos of every context grow to the grid's heaviest slice -- and
            # is listed, not counted. The 64K^2 grid is a few milliseconds of work per GPU at N = 8: three timed passes, the median
            # is reported and all are listed (one descheduled host thread is a fifth of such a build). The 256K^2 grid hands its
            # tiles out through the shared queue, so a rank meets other tiles in every pass: its second pass is still warming up
            # (N = 8: 307, 97, 61 ms), hence three timed passes there too.
            passes = 4
            reps = [gridbuild.run(ctx, args.grid_tile, length, kind, rank, world, gloo, lookups=3840 * 2160, lookup_iters=32, verify=(r == 0),
                                  replicate=(r == 0)) for r in range(passes)]
            by_time = sorted(reps[1:], key=lambda g: g["build_ms_max_rank"])
            g = dict(reps[0])
            for key in ("build_ms_max_rank", "build_ms_per_rank", "depth_ms_per_rank", "tiles_per_rank", "wall_ms_max_rank", "build_msamples_per_s", "moved_tiles"):
                g[key] = by_time[len(by_time) // 2][key]
            entry = grid_line(g, cfg_index)
            entry["first_pass_build_ms_max_rank"] = reps[0]["build_ms_max_rank"]
            entry["repetitions_build_ms_max_rank"] = [r["build_ms_max_rank"] for r in reps[1:]]
            grids["configs%d" % cfg_index] = entry
            barrier()

    if rank == 0:
        peak, peak_src = peaks()
        names = cpvs_b200.PHASE_NAMES
        phase = {nm: statistics.mean(tm[1][nm] for tm in timings) for nm in names}
        pyr_total = statistics.mean(tm[0][0] for tm in timings)
        pyr_base = statistics.mean(tm[0][1] for tm in timings)
        create_ms = statistics.mean(tm[2] for tm in timings)
        solo_phase = {nm: statistics.mean(tm[1][nm] for tm in solo_timings) for nm in names}
        solo_pyr = statistics.mean(tm[0][0] for tm in solo_timings)
        solo_pyr_base = statistics.mean(tm[0][1] for tm in solo_timings)
        solo_create = statistics.mean(tm[2] for tm in solo_timings)
        n_leaves = int(info0.svo_nodes[2]) if leaf else 0
        u_leaves = int(info0.dag_nodes[2]) if leaf else 0
        # algorithmic (compulsory HBM) bytes per launch of the single-kernel phases -- DESIGN.md "Kernels" (timed alone: the
        # synchronous builds, where no other build's kernels share the GPU):
        #   pyramid_base  depth read once + levels 3..5 written (levels 1 and 2 are produced lazily, not by this launch) + one
        #                 residue byte per texel where the per-column leaf builder will read it
        #   leaves        per column (whole-volume builds of maps >= 8192^2 with 2..8 leaves per column, the library's own rule): one
        #                 residue byte per texel + per column 8 B level-3 texel and 4 B bias in + per leaf 4 B index in, 32 B k-code
        #                 and 2 B mask out; per leaf (otherwise): depth read once (L2 serves the z-block re-reads) + per leaf 8 B
        #                 coordinate in, 32 B k-code, 8 B hash, 2 B mask out
        #   leaf_insert   per leaf: 32 B own k-code (+ 8 B hash on the per-leaf path) in, 4 B slot out; per duplicate: 32 B
        #                 representative k-code (the table itself is sized to stay in L2)
        #   emit_leaves   per unique leaf: 4 B index + 4 B offset + 2 B mask + 32 B k-code in; compressed words out
        cols = (n // 8) * (n // 8)
        per_column = (leaf and n >= 8192 and 2 * cols <= n_leaves <= 8 * cols
                      and os.environ.get("CPVS_LEAF_COLUMNS", "1") != "0") or os.environ.get("CPVS_LEAF_COLUMNS") == "2"
        leaves_bytes = (1.0 * n * n + cols * 12.0 + n_leaves * (4.0 + 32 + 2)) if per_column else (4.0 * n * n + n_leaves * (8.0 + 32 + 8 + 2))
        kernels = {
            "pyramid_base": ((4.0 + 8.0 * (1 / 64 + 1 / 256 + 1 / 1024) + (1.0 if per_column else 0.0)) * n * n, solo_pyr_base),
            "leaves": (leaves_bytes, solo_phase["leaves"]),
            "leaf_insert": (n_leaves * ((0.0 if per_column else 8.0) + 32 + 4) + (n_leaves - u_leaves) * 32.0, solo_phase["leaf_insert"]),
            "emit_leaves": (u_leaves * (4.0 + 4 + 2 + 32) + 4.0 * int(info0.dag_words[2] if leaf else 0), solo_phase["emit_leaves"]),
        }
        dom = max(kernels, key=lambda k: kernels[k][1])
        dom_bytes, dom_ms = kernels[dom]
        # measured DRAM bytes per launch of the same kernel from the committed `ncu --set full` capture (profiles/)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "top_kernel_traffic.json")
        if os.path.exists(tpath) and n == 16384 and args.kind == "terrain":
            kname = {"leaves": "buildLeafColumnsResidueKernel" if per_column else "buildLeavesKernel", "leaf_insert": "insertLeavesKernel",
                     "emit_leaves": "emitLeavesKernel", "pyramid_base": "pyramidBaseKernel"}
            with open(tpath) as f:
                table = json.load(f)["dram_bytes_per_launch"]
            traffic = next((v for k, v in table.items() if k == kname[dom] or k.startswith(kname[dom] + "<")), None)
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        bbytes = build_bytes(n, info0, leaf)
        roof_ms = bbytes / (peak * 1e9) * 1e3
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/u32", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": n * n * 4, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes, "ms_per_launch": dom_ms,
                         "timed": "CUDA events around the kernel in the synchronous builds (one build on the GPU at a time)",
                         "kernels": {k: {"algorithmic_bytes": v[0], "ms": v[1], "gbs": (v[0] / (v[1] * 1e-3) / 1e9 if v[1] > 0 else 0.0)}
                                     for k, v in kernels.items()}},
            "build_roofline": {"algorithmic_bytes": bbytes, "roofline_ms": roof_ms, "formula": "SURVEY.md 8(d)",
                               "ms_per_step": ms_step, "frac": roof_ms / ms_step,
                               "one_build_at_a_time": {"ms_per_step": solo_ms, "frac": roof_ms / solo_ms, "device_ms": solo_pyr + solo_create}},
            "one_build_at_a_time": {"ms_per_step": solo_ms, "value": world * n * n / (solo_ms * 1e-3) / 1e6, "unit": UNIT,
                                    "what": "cpvs_minmax_build + cpvs_shadow_create, one context, each call waited for",
                                    "phases_ms": dict(solo_phase, pyramid=solo_pyr, pyramid_base=solo_pyr_base, create_total=solo_create)},
            "phases_ms": dict(phase, pyramid=pyr_total, pyramid_base=pyr_base, create_total=create_ms,
                              note="per build, while %d builds share the GPU" % max(1, in_flight)),
            "predicted_builds": sum(1 for tm in timings if tm[3]), "context_stats": [c.stats() for c in ctxs],
            "dag": {"words": int(info0.words), "mbytes": int(info0.words) * 4 / 1e6, "num_levels": int(info0.num_levels),
                    "svo_nodes": [int(v) for v in info0.svo_nodes[:info0.num_levels - 1]],
                    "dag_nodes": [int(v) for v in info0.dag_nodes[:info0.num_levels - 1]], **digest},
            "lookups": {"count": args.lookups, "value": args.lookups / (lookup_ms * 1e-3) / 1e9, "unit": "Glookups/s", "ms": lookup_ms,
                        "e2e_value": args.lookups / (lookup_e2e_ms * 1e-3) / 1e9, "e2e_ms": lookup_e2e_ms,
                        "stream_bytes": args.lookups * 13},
            "configs3": configs3,
        }
        line.update(grids)
        if world == 1 and not args.no_cpu_baseline:
            cores = max(1, min(os.cpu_count() or 1, 64))
            v, ms, desc, kind_used = reference_sample(n, args.kind, args.ref_sample, cores, 3, 1)
            lk1, lkn = reference_lookups(n, args.kind, args.ref_sample, cores, args.lookups)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind_used, "sample": desc, "ms_per_step": ms,
                                    "lookups_mps_1_thread": lk1, "lookups_mps_all_threads": lkn,
                                    "lookups_sample": "%d random NDC points, traverse() on the DAG of one %dx%d window" % (args.lookups, args.ref_sample, args.ref_sample)}
            if not args.no_port_16k:
                line["cpu_baseline"]["port_full_size"] = port_full_size(depth_np, n, args.kind, int(info0.words), digest["fnv64"])
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def port_full_size(depth_np, n, kind, words, fnv):
    """The whole bench map on ONE host thread through the oracle's port (hash-based merge, word-identical to the reference where
    the reference finishes; NOT the reference, whose O(n*u) merge needs hours at this size). About half a minute at 16K^2."""
    from oracle import pyoracle as O
    from cpvs_b200 import synth
    t0 = time.perf_counter()
    sh = O.Shadow(O.MinMax(depth_np, "port"))
    dag = sh.dag()
    dt = time.perf_counter() - t0
    return {"kind": "port", "note": "not the reference: same algorithm with a hash-based mergeLevel", "cores": 1, "seconds": dt,
            "value": n * n / dt / 1e6, "unit": UNIT, "sample": "the whole %dx%d %s map, MinMaxHierarchy + create" % (n, n, kind),
            "words": int(dag.size), "words_equal_gpu": bool(int(dag.size) == words and "%016x" % synth.fnv64(dag) == fnv)}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.grid:
        run_grid(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
