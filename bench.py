#!/usr/bin/env python
"""bench.py -- headline benchmark of the cpvs hot path on B200 (BASELINE.json: DAG build Msamples/s at
16K^2, shadow lookups G/s, % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl own|reference]

One step = one pass of the hot path over one depth map: MinMaxHierarchy + CompressedShadow::create on
the device (N=1: BASELINE configs[1], one 16K^2 terrain map, leafmasks on, single DAG; N>1: configs[2],
every rank builds one 16K^2 xy-tile of the 4x4 virtual 64K^2 map with its 4 z-slices, no collective on
the data path). `value` is measured with the depth map resident in HBM; `e2e` goes through the same
C-ABI calls with the depth map in pinned host memory. Lookups (1M random NDC points) are timed next to
it. `--impl reference` times the reference's own CPU implementation (oracle/_ref, the unmodified
sources compiled by oracle/Makefile) on bounded samples of the same workload.
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
    ap.add_argument("--kind", default="terrain", choices=["plane", "terrain", "city"])
    ap.add_argument("--lookups", type=int, default=1000000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="do not run the nvidia-smi sampler (debugging)")
    ap.add_argument("--ref-sample", type=int, default=1024, help="side of one reference sample window")
    ap.add_argument("--z-slices", type=int, default=1,
                    help="z-slice DAGs built per depth map (createShadowTiles); 4 = the cubic 4x4x4 container of BASELINE configs[2]")
    ap.add_argument("--grid", default=None, choices=["64k", "256k"],
                    help="whole tile-grid build instead of the step bench: 64k = BASELINE configs[2] (4x4x4 cells of 16K^2 terrain "
                         "tiles), 256k = configs[4] (16x16x16 cells of 16K^2 city tiles); tiles are sharded over --gpus ranks")
    ap.add_argument("--grid-tile", type=int, default=16384, help="side of one depth tile of --grid (reduce for a quick run)")
    ap.add_argument("--grid-length", type=int, default=None, help="override the grid length of --grid")
    ap.add_argument("--grid-kind", default=None, choices=["plane", "terrain", "city"])
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
            "dtype": "f32/u32", "data": "synthetic", "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind_used, "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, n_gpus):
    cfg = _workload_config(args, n_gpus)
    if os.environ.get("CPVS_EXPERIMENTS"):  # unmeasured kernel variants switched on for this run
        cfg["experiments"] = os.environ["CPVS_EXPERIMENTS"]
    return cfg


def _workload_config(args, n_gpus):
    if n_gpus == 1:
        return {"workload": "configs[1]: %dx%d synthetic %s depth map, leafmasks on, single DAG (MinMaxHierarchy + CompressedShadow::create) "
                            "+ %d random NDC lookups" % (args.size, args.size, args.kind, args.lookups),
                "depth_map": "%dx%d f32 (%.0f MiB) > L2, regenerated state per step; no explicit L2 flush needed" % (args.size, args.size, args.size * args.size * 4 / 2**20),
                "tiles_per_rank": 1, "z_slices": args.z_slices}
    return {"workload": "configs[2] sharding: a 4x4 grid of %dx%d %s xy-tiles at the texel density of configs[1] (64K^2 texels in all); rank r "
                        "builds tile r (1 pyramid + %d z-slice DAG(s)) per step -- the same unit of work as at N=1 --, host-side gather "
                        "of sizes only" % (args.size, args.size, args.kind, args.z_slices),
            "depth_map": "%dx%d f32 per rank > L2" % (args.size, args.size), "tiles_per_rank": 1, "z_slices": args.z_slices}


# ---- whole tile grids (BASELINE configs[2] and configs[4]) ---------------------------------------

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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        cbuild.build()
    if world > 1:
        dist.barrier()
    length = args.grid_length or (4 if args.grid == "64k" else 16)
    kind = args.grid_kind or ("terrain" if args.grid == "64k" else "city")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = cpvs_b200.Context(local, stream=stream.cuda_stream)
    sampler = ClockSampler(local)
    if not args.no_clocks:
        sampler.start()
    res = gridbuild.run(ctx, stream, args.grid_tile, length, kind, rank, world, dist if world > 1 else None,
                        verify=not args.no_verify,
                        log=(lambda m: print(m, file=sys.stderr, flush=True)) if os.environ.get("CPVS_GRID_LOG") else None)
    clocks = sampler.stop()
    if rank == 0:
        line = {"metric": METRIC, "value": res["build_msamples_per_s"], "unit": UNIT, "n_gpus": world, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32/u32", "data": "synthetic",
                "config": {"workload": "configs[%d]: %dK^2 virtual %s shadow map as a %dx%dx%d CompressedShadowContainer of %dx%d depth "
                                       "tiles, xy tiles sharded round-robin over %d GPU(s), host-side gather of sizes only"
                                       % (2 if args.grid == "64k" else 4, res["virtual_side"] // 1024, kind, length, length, length,
                                          args.grid_tile, args.grid_tile, world)},
                "ms_total": res["build_ms_max_rank"], "gpu_launches": res["gpu_launches"], "clocks": clocks, "grid": res}
        if not args.no_cpu_baseline:
            # "build time vs reference": the reference itself on a bounded sample of the same virtual map, extrapolated
            # per sample (its merge is O(n*u) per DAG, so whole 16K^2 tiles would only be slower than this)
            cores = max(1, min(os.cpu_count() or 1, 64))
            v, ms, desc, kind_used = reference_sample(res["virtual_side"], kind, args.ref_sample, cores, 2, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind_used, "sample": desc, "ms_per_step": ms,
                                    "extrapolated_build_s": res["samples"] / (v * 1e6),
                                    "note": "single z-slice per window; extrapolated linearly in samples"}
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


def run_own(args):
    import torch
    import torch.distributed as dist
    import cpvs_b200
    from cpvs_b200 import build as cbuild, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        cbuild.build()
    if world > 1:
        dist.barrier()
    n, K, W = args.size, args.steps, args.warmup
    # a real (non-default) stream: the library enqueues on it and the CUDA events below are recorded on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = cpvs_b200.Context(local, stream=stream.cuda_stream)

    # workload: N=1 whole map, one DAG; N>1 xy-tile `rank` of the 4x4 virtual map, 4 z-slices
    # the unit of work is the same at every N (weak scaling): one 16K^2 depth map -> pyramid + z_slices DAGs
    z_slices = args.z_slices
    # rank r takes the 16K^2 map whose origin is shifted by (r % 4, r // 4) maps in the same analytic scene at
    # the same texel density: statistically the same work as rank 0's configs[1] map, different data
    tile, tps = (rank % 4, (rank // 4) % 4), 1
    host = torch.empty((n, n), dtype=torch.float32, pin_memory=True)
    depth_np = host.numpy()
    synth.depth_map(args.kind, n, tile, tps, out=depth_np)
    depth = host.to("cuda", non_blocking=True)
    torch.cuda.synchronize()

    def step(src, keep=False):
        mm = cpvs_b200.MinMaxHierarchy(src, ctx, n=n)
        shadows = [cpvs_b200.CompressedShadow.create(mm, z, z_slices) for z in range(z_slices)]
        if keep:
            return mm, shadows
        infos = [s.info for s in shadows]
        timing = (mm.timing(), [s.phase_ms() for s in shadows], [s.info.build_ms for s in shadows])
        for s in shadows:
            s.close()
        mm.close()
        return infos, timing

    # parity property on the bench workload itself: every looked-up voxel decodes to z + 0.5 <= d * H.
    # These handles stay alive (the lookups below run on them), so they are built before the warm-up.
    mm, shadows = step(depth, keep=True)
    pts_np = synth.lookups(args.lookups)
    res = n * z_slices  # z resolution of the whole tile column
    path = (((pts_np + np.float32(1)) * np.float32(0.5)) * np.float32(n - 1)).astype(np.int32)
    for z, sh in enumerate(shadows):
        vis = sh.traverse(pts_np)
        zz = path[:, 2] + z * n
        lit = (zz.astype(np.float32) + np.float32(0.5)) <= depth_np[path[:, 1], path[:, 0]] * np.float32(res)
        if not np.array_equal(vis, lit.astype(np.uint8)):
            raise SystemExit("bench.py: lookup results do not decode to the depth map (z-slice %d)" % z)
    main_idx = max(range(len(shadows)), key=lambda i: int(shadows[i].info.words))
    main_shadow = shadows[main_idx]
    info0 = main_shadow.info
    leaf = bool(info0.leafmasks)
    for _ in range(W):
        step(depth)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed: device-resident input ----
    sampler = ClockSampler(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = ctx.launch_count
    if not args.no_clocks:
        sampler.start()
    ev0.record(stream)
    timings = []
    for _ in range(K):
        timings.append(step(depth)[1])
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / K
    value = world * n * n / (ms_step * 1e-3) / 1e6

    # ---- timed: end to end from pinned host memory through the C ABI ----
    for _ in range(min(W, 2)):
        step(depth_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        step(depth_np)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / K
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = world * n * n / (e2e_ms * 1e-3) / 1e6
    d2h = z_slices * (192 * 8 + 32 * 8 + 4)  # size scalars read back per create

    # ---- lookups on the resident DAG ----
    # 16 different batches (seeds 777..792) are cycled so that no iteration finds its points in L2
    batches = [torch.from_numpy(pts_np if b == 0 else synth.lookups(args.lookups, seed=777 + b)).cuda() for b in range(16)]
    out = torch.empty(args.lookups, dtype=torch.uint8, device="cuda")
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

    # ---- BASELINE configs[3]: a 4K G-buffer of world positions on / just off the surface (deepest descent),
    # through CompressedShadowContainer::evaluate (light transform + grid step + DAG descent) ----
    surface = None
    if world == 1 and z_slices == 1:
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
        cont = cpvs_b200.CompressedShadowContainer(main_shadow, ctx)
        cont.copyToGPU()
        frames = [torch.from_numpy(pos_np).cuda() for _ in range(4)]  # 4 x 133 MB, cycled: every frame comes from HBM
        vis = torch.zeros((gh, gw), dtype=torch.uint8, device="cuda")
        ident = np.eye(4, dtype=np.float32)
        for i in range(max(W, 3)):
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
        # decode check: lit iff z + 0.5 <= d * N for the texel the path lands in
        path = (((pos_np[..., :3] + np.float32(1)) * np.float32(0.5)) * np.float32(n - 1)).astype(np.int32)
        want = ((path[..., 2].astype(np.float32) + np.float32(0.5)) <= depth_np[path[..., 1], path[..., 0]] * np.float32(n))
        if not np.array_equal(vis.cpu().numpy() != 0, want):
            raise SystemExit("bench.py: evaluate() results do not decode to the depth map")
        surface = {"pixels": gw * gh, "value": gw * gh / (surf_ms * 1e-3) / 1e9, "unit": "Glookups/s", "ms": surf_ms,
                   "stream_bytes": gw * gh * 17, "lit_fraction": float(want.mean()),
                   "what": "3840x2160 rgba32f positions within 1.5 texels of the surface, identity lightViewProj, leafmasks on; "
                           "4 frames cycled (532 MB > L2)"}
        cont.close()

    # ---- gather of sizes on the host (the only cross-rank step of a tiled build) ----
    sizes = [(int(s.info.words), int(s.info.total_visibility)) for s in shadows]
    gathered = [sizes]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, sizes)

    if rank == 0:
        peak, peak_src = peaks()
        # per-phase device time, averaged over the timed steps (sum over z-slices within a step)
        names = cpvs_b200.PHASE_NAMES
        phase = {nm: statistics.mean(sum(p[nm] for p in tm[1]) for tm in timings) for nm in names}
        main_phase = {nm: statistics.mean(tm[1][main_idx][nm] for tm in timings) for nm in names}  # the z-slice `dag` describes
        pyr_total = statistics.mean(tm[0][0] for tm in timings)
        pyr_base = statistics.mean(tm[0][1] for tm in timings)
        create_ms = statistics.mean(sum(tm[2]) for tm in timings)
        n_leaves = int(info0.svo_nodes[2]) if leaf else 0
        u_leaves = int(info0.dag_nodes[2]) if leaf else 0
        # algorithmic (compulsory HBM) bytes per launch of the single-kernel phases -- DESIGN.md "Kernels":
        #   pyramid_base  depth read once + levels 1..5 written
        #   leaves        per column (whole-volume builds of maps >= 8192^2 with 2..8 leaves per column, the library's own rule): depth read once
        #                 + per column 8 B level-3 texel and 4 B bias in + per leaf 4 B index in, 32 B k-code and 2 B mask out;
        #                 per leaf (otherwise): depth read once (L2 serves the z-block re-reads) + per leaf 8 B coordinate in,
        #                 32 B k-code, 8 B hash, 2 B mask out
        #   leaf_insert   per leaf: 32 B own k-code (+ 8 B hash on the per-leaf path) in, 4 B slot out; per duplicate: 32 B
        #                 representative k-code (the table itself is sized to stay in L2)
        #   emit_leaves   per unique leaf: 4 B index + 4 B offset + 2 B mask + 32 B k-code in; compressed words out
        cols = (n // 8) * (n // 8)
        per_column = (leaf and args.z_slices == 1 and n >= 8192 and 2 * cols <= n_leaves <= 8 * cols
                      and os.environ.get("CPVS_LEAF_COLUMNS", "1") != "0") or os.environ.get("CPVS_LEAF_COLUMNS") == "2"
        leaves_bytes = (4.0 * n * n + cols * 12.0 + n_leaves * (4.0 + 32 + 2)) if per_column else (4.0 * n * n + n_leaves * (8.0 + 32 + 8 + 2))
        kernels = {
            "pyramid_base": ((4.0 + 8.0 * (1 / 4 + 1 / 16 + 1 / 64 + 1 / 256 + 1 / 1024)) * n * n, pyr_base),
            "leaves": (leaves_bytes, main_phase["leaves"]),
            "leaf_insert": (n_leaves * ((0.0 if per_column else 8.0) + 32 + 4) + (n_leaves - u_leaves) * 32.0, main_phase["leaf_insert"]),
            "emit_leaves": (u_leaves * (4.0 + 4 + 2 + 32) + 4.0 * int(info0.dag_words[2] if leaf else 0), main_phase["emit_leaves"]),
        }
        dom = max(kernels, key=lambda k: kernels[k][1])
        dom_bytes, dom_ms = kernels[dom]
        # measured DRAM bytes per launch of the same kernel from the committed `ncu --set full` capture (profiles/)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "top_kernel_traffic.json")
        if os.path.exists(tpath) and n == 16384 and args.kind == "terrain" and world == 1:
            names = {"leaves": "buildLeafColumnsKernel" if per_column else "buildLeavesKernel", "leaf_insert": "insertLeavesKernel", "emit_leaves": "emitLeavesKernel",
                     "pyramid_base": "pyramidBaseKernel<0>"}
            with open(tpath) as f:
                traffic = json.load(f)["dram_bytes_per_launch"].get(names[dom])
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        bbytes = build_bytes(n, info0, leaf)
        step_ms_device = pyr_total + create_ms
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/u32", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": n * n * 4, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes, "ms_per_launch": dom_ms,
                         "kernels": {k: {"algorithmic_bytes": v[0], "ms": v[1], "gbs": (v[0] / (v[1] * 1e-3) / 1e9 if v[1] > 0 else 0.0)}
                                     for k, v in kernels.items()}},
            "build_roofline": {"algorithmic_bytes": bbytes, "roofline_ms": bbytes / (peak * 1e9) * 1e3, "device_ms": step_ms_device,
                               "frac": (bbytes / (peak * 1e9) * 1e3) / step_ms_device, "formula": "SURVEY.md 8(d)"},
            "phases_ms": dict(phase, pyramid=pyr_total, pyramid_base=pyr_base, create_total=create_ms),
            "dag": {"words": int(info0.words), "mbytes": int(info0.words) * 4 / 1e6, "num_levels": int(info0.num_levels),
                    "svo_nodes": [int(v) for v in info0.svo_nodes[:info0.num_levels - 1]],
                    "dag_nodes": [int(v) for v in info0.dag_nodes[:info0.num_levels - 1]]},
            "lookups": {"count": args.lookups, "value": args.lookups / (lookup_ms * 1e-3) / 1e9, "unit": "Glookups/s", "ms": lookup_ms,
                        "e2e_value": args.lookups / (lookup_e2e_ms * 1e-3) / 1e9, "e2e_ms": lookup_e2e_ms,
                        "stream_bytes": args.lookups * 13, "surface_gbuffer": surface},
            "grid_gather": {"cells": sum(len(g) for g in gathered), "words": sum(w for g in gathered for w, _ in g)},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = max(1, min(os.cpu_count() or 1, 64))
            v, ms, desc, kind_used = reference_sample(n, args.kind, args.ref_sample, cores, 3, 1)
            lk1, lkn = reference_lookups(n, args.kind, args.ref_sample, cores, args.lookups)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind_used, "sample": desc, "ms_per_step": ms,
                                    "lookups_mps_1_thread": lk1, "lookups_mps_all_threads": lkn,
                                    "lookups_sample": "%d random NDC points, traverse() on the DAG of one %dx%d window" % (args.lookups, args.ref_sample, args.ref_sample)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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
