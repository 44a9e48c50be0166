"""One short process (no torch) that runs the default path and every experimental kernel variant (CPVS_EXPERIMENTS, see
cpvs_b200/csrc/kernels.h) on the same resident 16K^2 terrain depth map: median per-phase device times of a few builds and a word-for-word comparison of the DAG with the default path's,
plus a sweep of small maps with the per-column leaf builder forced. The switches are read when a context is created, so one
process can walk through them. Results are appended to gpurun_out/switch_probe.txt line by line (a run that is cut short
keeps what it has). Risky settings come last: a faulting kernel poisons the CUDA context for everything after it.

    python scripts/switch_probe.py [size] [builds] [time budget in s]
"""
import os
import statistics
import sys
import time

import numpy as np

T0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpvs_b200  # noqa: E402
from cpvs_b200 import synth  # noqa: E402

KEYS = ["CPVS_EXPERIMENTS", "CPVS_LEAF_COLUMNS"]
NAMES = ["expand-preload", "emit-gather", "rank-preload", "insert-witness", "early-bases", "leaf-fp64"]
SETTINGS = [{}] + [{"CPVS_EXPERIMENTS": name} for name in NAMES] + [{"CPVS_EXPERIMENTS": ",".join(NAMES)}, {}]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    builds = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    budget = float(sys.argv[3]) if len(sys.argv) > 3 else 17.0
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "switch_probe.txt"), "a")

    def say(text):
        line = "[%5.1f s] %s" % (time.time() - T0, text)
        print(line, flush=True)
        log.write(line + "\n")
        log.flush()
        os.fsync(log.fileno())

    def context(env):
        for k in KEYS:
            os.environ.pop(k, None)
        os.environ.update(env)
        return cpvs_b200.Context(0)

    say("switch probe: %dx%d terrain, median of %d builds after 1 warm-up" % (n, n, builds))
    lib = cpvs_b200.load_library()
    ctx0 = context({})
    depth = synth.depth_map("terrain", n)
    say("depth map generated")
    mm0 = cpvs_b200.MinMaxHierarchy(depth, ctx0)  # the one host-to-device copy; its level 0 is the resident depth map
    ctx0.synchronize()
    dptr = int(lib.cpvs_minmax_level_device(mm0.handle, 0))
    del depth
    say("depth map resident")

    rng = np.random.default_rng(5)
    small = [("terrain 512", synth.depth_map("terrain", 512), 0, 1), ("city 512 z1/2", synth.depth_map("city", 512), 1, 2),
             ("plane 256", synth.depth_map("plane", 256), 0, 1), ("random 128", rng.random((128, 128), dtype=np.float32), 0, 1),
             ("terrain 256 z2/4", synth.depth_map("terrain", 256), 2, 4), ("city 1024", synth.depth_map("city", 1024), 0, 1)]

    def small_dags(ctx):
        out = []
        for _, d, zt, zn in small:
            mm = cpvs_b200.MinMaxHierarchy(d, ctx)
            sh = cpvs_b200.CompressedShadow.create(mm, zt, zn)
            out.append(sh.getDAG())
            sh.close()
            mm.close()
        return out

    small_base = small_dags(ctx0)
    base_dag = None
    for env in SETTINGS:
        tag = ",".join("%s=%s" % (k[5:], v) for k, v in sorted(env.items())) or "default"
        if time.time() - T0 > budget:
            say("%-60s skipped (time budget)" % tag)
            continue
        try:
            ctx = context(env)
            rows, dag = [], None
            for i in range(builds + 1):
                mm = cpvs_b200.MinMaxHierarchy(dptr, ctx, n=n)
                sh = cpvs_b200.CompressedShadow.create(mm)
                if i > 0:
                    rows.append(dict(sh.phase_ms(), create=sh.info.build_ms, pyramid=mm.timing()[0]))
                if i == builds:
                    dag = sh.getDAG()
                sh.close()
                mm.close()
            med = {k: statistics.median(r[k] for r in rows) for k in rows[0]}
            if base_dag is None:
                base_dag = dag
            same = dag.size == base_dag.size and np.array_equal(dag, base_dag)
            ctx.close()
            # small maps with the per-column builder forced (the switches that touch the leaf path only act there)
            ctx2 = context(dict(env, CPVS_LEAF_COLUMNS="2"))
            got = small_dags(ctx2)
            ctx2.close()
            bad = [small[i][0] for i in range(len(small)) if got[i].size != small_base[i].size or not np.array_equal(got[i], small_base[i])]
            say("%-60s build %.3f ms (pyramid %.3f + create %.3f) %s%s" % (
                tag, med["pyramid"] + med["create"], med["pyramid"], med["create"], "words identical" if same else "WORDS DIFFER (%d vs %d)" % (dag.size, base_dag.size),
                "" if not bad else "; SMALL MAPS DIFFER: " + ", ".join(bad)))
            say("        " + " ".join("%s %.3f" % (k, v) for k, v in med.items() if k not in ("create", "pyramid")))
        except Exception as exc:  # noqa: BLE001 -- keep going: later settings may still work
            say("%-60s FAILED: %s" % (tag, exc))
    say("done")


if __name__ == "__main__":
    main()
