"""CPU side of SURVEY.md 8(d) / BASELINE.md plan items 2-3, on the host this runs on: the unmodified reference (oracle/_ref)
at the sizes it finishes in minutes, the growth exponent of its create(), and the word-identical hash-merging port
(oracle/oracle_port.cpp -- NOT the reference) up to 16K^2 as the best-effort CPU figure. Writes a markdown table.

    python scripts/cpu_scaling.py [--max-port 16384] [--out profiles/r1_cpu_scaling.md]
"""
import argparse
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cpvs_b200 import synth  # noqa: E402
from oracle import pyoracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-port", type=int, default=16384)
    ap.add_argument("--max-ref-terrain", type=int, default=2048, help="4096 adds ~3.5 minutes")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r1_cpu_scaling.md"))
    a = ap.parse_args()
    O.build(ref=True)
    pts = synth.lookups(1000000)
    cores = os.cpu_count()
    lines = ["# CPU baseline scaling -- r1", "",
             "Host: %d cores (this container; the GPU box's host differs -- `bench.py` times the reference there on bounded windows)." % cores,
             "Reference = unmodified sources compiled by `oracle/Makefile` (`-O2 -DNDEBUG`), `MinMaxHierarchy` (4 threads) +",
             "`CompressedShadow::create` (1 thread). Port = `oracle/oracle_port.cpp`, word-identical, hash-based `mergeLevel` -- **not the reference**.",
             "", "| N | map | reference pyramid ms | reference create ms | reference Msamples/s | words | port build ms | port Msamples/s | 1 M traverse ms, 1 thread (reference) | %d threads |" % cores,
             "|---|---|---|---|---|---|---|---|---|---|"]
    ref_create = {}
    for n in (1024, 2048, 4096, 8192, 16384):
        for kind in ("plane", "terrain", "city"):
            run_ref = n <= 2048 or (n == 4096 and (kind != "terrain" or a.max_ref_terrain >= 4096))
            run_port = n <= a.max_port and (kind == "terrain" or n <= 4096)
            if not (run_ref or run_port):
                continue
            d = synth.depth_map(kind, n)
            row = ["%d" % n, kind]
            words = None
            if run_ref:
                pyr, cre, words = O.ref_time_build(d)
                ref_create.setdefault(kind, []).append((n, cre))
                row += ["%.1f" % pyr, "%.1f" % cre, "%.2f" % (n * n / 1e3 / (pyr + cre))]
            else:
                row += ["--", "not run", "--"]
            if run_port:
                t0 = time.perf_counter()
                s = O.Shadow(O.MinMax(d, "port"))
                port_ms = (time.perf_counter() - t0) * 1e3
                w = int(s.dag().size)
                assert words is None or w == words, (n, kind, w, words)
                words = w
                del s
                row += ["%d" % words, "%.0f" % port_ms, "%.2f" % (n * n / 1e3 / port_ms)]
            else:
                row += ["%d" % words, "--", "--"]
            if run_ref and n <= 2048:
                s = O.Shadow(O.MinMax(d, "ref"))
                t0 = time.perf_counter()
                r1 = s.traverse(pts)
                t1 = time.perf_counter()
                r8 = s.traverse(pts, threads=cores)
                t2 = time.perf_counter()
                assert np.array_equal(r1, r8)
                row += ["%.1f (%.1f M/s)" % ((t1 - t0) * 1e3, len(pts) / (t1 - t0) / 1e6),
                        "%.1f (%.1f M/s)" % ((t2 - t1) * 1e3, len(pts) / (t2 - t1) / 1e6)]
                del s
            else:
                row += ["--", "--"]
            lines.append("| " + " | ".join(row) + " |")
            print(lines[-1], flush=True)
            del d
    lines.append("")
    for kind, ptsk in ref_create.items():
        if len(ptsk) >= 2:
            (n0, c0), (n1, c1) = ptsk[0], ptsk[-1]
            e = math.log(c1 / c0) / math.log(n1 / n0)
            ext = c1 * (16384 / n1) ** e
            lines.append("* reference `create`, %s: %.1f ms at %d^2 -> %.1f ms at %d^2, i.e. time ~ N^%.2f (%.1fx per doubling); "
                         "extrapolated to 16384^2: %.3g s (**extrapolation**, not measured)." % (kind, c0, n0, c1, n1, e, 2 ** e, ext / 1e3))
    with open(a.out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[-4:]))


if __name__ == "__main__":
    main()
