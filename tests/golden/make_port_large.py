"""Regenerates tests/golden/port_large.json: the DAGs of the headline-size maps (BASELINE configs[1] and the tiles of
configs[2] / configs[4]) from the CPU restatement oracle/oracle_port.cpp, which tests/test_oracle.py pins word for word
to the compiled reference at every size the reference finishes (<= 2048^2 terrain, 4096^2 plane / city).

    python tests/golden/make_port_large.py            (build container only: minutes of CPU time, tens of GB of RAM)

Per map: word count, the survey's FNV-64 digest of the words (synth.fnv64), per-level SVO / DAG node counts, and the
digest of 1 M lookups (seed 777). The GPU suite rebuilds the same maps and compares (tests/test_gpu_large.py).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cpvs_b200 import synth  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "port_large.json")
CASES = [("terrain", 4096, (0, 0), 1, 0, 1), ("terrain", 8192, (0, 0), 1, 0, 1), ("terrain", 16384, (0, 0), 1, 0, 1),
         ("terrain_dev", 16384, (0, 0), 1, 0, 1), ("city", 16384, (0, 0), 1, 0, 1),
         # one tile column of configs[2] (64K^2 terrain_dev as 4 x 4 tiles, 4 z-slices) and a tile of configs[4] (256K^2 city, 16 slices)
         ("terrain_dev", 16384, (1, 2), 4, 1, 4), ("terrain_dev", 16384, (1, 2), 4, 2, 4), ("city", 16384, (5, 9), 16, 8, 16), ("city", 16384, (5, 9), 16, 10, 16), ("city", 16384, (0, 0), 16, 4, 16)]


def main():
    only = sys.argv[1:]
    rows = json.load(open(OUT))["maps"] if os.path.exists(OUT) else []
    done = {(r["kind"], r["n"], tuple(r["tile"]), r["tiles_per_side"], r["z_tile"], r["z_num"]) for r in rows}
    pts = synth.lookups(1000000)
    for kind, n, tile, tps, zt, zn in CASES:
        if (kind, n, tile, tps, zt, zn) in done or (only and kind not in only):
            continue
        t0 = time.time()
        d = synth.depth_map(kind, n, tile, tps)
        sh = O.Shadow(O.MinMax(d), zt, zn)
        dag = sh.dag()
        svo, uniq = sh.level_counts()
        vis = sh.traverse(pts)
        rows.append({"kind": kind, "n": n, "tile": list(tile), "tiles_per_side": tps, "z_tile": zt, "z_num": zn, "words": int(dag.size),
                     "fnv64": "%016x" % synth.fnv64(dag), "svo_nodes": [int(v) for v in svo], "dag_nodes": [int(v) for v in uniq],
                     "lit": int((vis == 1).sum()), "vis_fnv64": "%016x" % synth.fnv64(vis.astype(np.uint32)),
                     "port_seconds": round(time.time() - t0, 1)})
        print(rows[-1], flush=True)
        del sh, dag, d
        with open(OUT, "w") as f:
            json.dump({"source": "oracle/oracle_port.cpp (pinned to the compiled reference by tests/test_oracle.py); lookups: 1000000 points, seed 777",
                       "maps": rows}, f, indent=1)


if __name__ == "__main__":
    main()
