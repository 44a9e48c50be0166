"""Dev probe (no torch): throughput of back-to-back 16K^2 builds -- synchronous, asynchronous on one context (the next
build enqueued while the previous one runs), and on several contexts of the same GPU (independent builds whose kernels
overlap). Wall clock over `steps` builds after a warm-up, one resident depth map.

    python scripts/pipeline_probe.py [size] [kind] [steps]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpvs_b200  # noqa: E402
from cpvs_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
kind = sys.argv[2] if len(sys.argv) > 2 else "terrain"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
lib = cpvs_b200.load_library()
ctx0 = cpvs_b200.Context(0)
mm0 = cpvs_b200.MinMaxHierarchy(synth.depth_map(kind, n), ctx0)
ctx0.synchronize()
dptr = int(lib.cpvs_minmax_level_device(mm0.handle, 0))


def run(num_ctx, depth_in_flight, label):
    ctxs = [cpvs_b200.Context(0) for _ in range(num_ctx)]
    for c in ctxs:  # first build of the shape on every context (exact), then one predicted
        for _ in range(2):
            mm = cpvs_b200.MinMaxHierarchy(dptr, c, n=n)
            cpvs_b200.CompressedShadow.create(mm).close()
            mm.close()
    for c in ctxs:
        c.synchronize()
    flying = []
    t0 = time.perf_counter()
    for k in range(steps + 12):
        if k == 12:  # the first rounds grow the memory pool to what this many builds in flight need
            for c in ctxs:
                c.synchronize()
            t0 = time.perf_counter()
        c = ctxs[k % num_ctx]
        mm = cpvs_b200.MinMaxHierarchy(dptr, c, n=n)
        sh = cpvs_b200.CompressedShadow.create(mm, wait=depth_in_flight == 0)
        flying.append((mm, sh))
        while len(flying) > depth_in_flight:
            m, s = flying.pop(0)
            s.wait()
            s.close()
            m.close()
    while flying:
        m, s = flying.pop(0)
        s.wait()
        words = int(s.info.words)
        s.close()
        m.close()
    for c in ctxs:
        c.synchronize()
    dt = (time.perf_counter() - t0) / steps * 1e3
    print("%-44s %.3f ms per build  (%.1f Gsamples/s)" % (label, dt, n * n / dt / 1e6), flush=True)
    for c in ctxs:
        c.close()


run(1, 0, "synchronous, 1 context")
run(1, 1, "async, 1 context, 1 in flight")
run(1, 2, "async, 1 context, 2 in flight")
run(2, 2, "async, 2 contexts, 2 in flight")
run(2, 4, "async, 2 contexts, 4 in flight")
run(3, 3, "async, 3 contexts, 3 in flight")
run(4, 4, "async, 4 contexts, 4 in flight")
