"""Per-phase device times of 16K^2 builds under each experimental switch, next to the default path, in one process per
setting (the switches are read when a context is created). Writes gpurun_out/switch_compare.txt.

    python scripts/switch_compare.py [size] [kind]

Round 2, first GPU call:
    CPVS_TEST_EXPERIMENTAL=1 python -m pytest tests -m gpu -k experimental -q && python scripts/switch_compare.py
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SETTINGS = [{}, {"CPVS_EMIT_PLANES": "1"}, {"CPVS_LEAF_ORDER": "1", "CPVS_LEAF_CTAS": "2"}, {"CPVS_LEAF_ORDER": "1", "CPVS_LEAF_CTAS": "3"},
            {"CPVS_LEAF_ORDER": "1", "CPVS_LEAF_CTAS": "1"}, {"CPVS_LEAF_CTAS": "2"}, {"CPVS_LEAF_CTAS": "4"}, {"CPVS_LEAF_ORDER": "1", "CPVS_LEAF_CTAS": "4"}, {"CPVS_EXPAND_BLOCKS": "12"}, {"CPVS_INNER_BLOCKS": "8"},
            {"CPVS_INSERT_HINTS": "1"},
            {"CPVS_LEAF_ORDER": "1", "CPVS_EMIT_PLANES": "1", "CPVS_INSERT_HINTS": "1"}]

CHILD = r"""
import json, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch
import cpvs_b200
from cpvs_b200 import synth
n, kind = %(n)d, %(kind)r
ctx = cpvs_b200.Context(0)
d = torch.from_numpy(synth.depth_map(kind, n)).cuda()
torch.cuda.synchronize()
rows, words, digest = [], None, None
for i in range(9):
    mm = cpvs_b200.MinMaxHierarchy(d, ctx, n=n)
    sh = cpvs_b200.CompressedShadow.create(mm)
    if i >= 3:
        rows.append(dict(sh.phase_ms(), create_total=sh.info.build_ms, pyramid=mm.timing()[0]))
    if i == 8:
        dag = sh.getDAG()
        words, digest = int(dag.size), int(np.bitwise_xor.reduce(dag.astype(np.uint64) * np.arange(1, dag.size + 1, dtype=np.uint64)))
    sh.close(); mm.close()
med = {k: float(np.median([r[k] for r in rows])) for k in rows[0]}
print(json.dumps({"median_ms": med, "words": words, "digest": digest}))
"""


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    kind = sys.argv[2] if len(sys.argv) > 2 else "terrain"
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    lines, base = [], None
    for env in SETTINGS:
        e = dict(os.environ, **env)
        tag = ",".join("%s=%s" % kv for kv in sorted(env.items())) or "default"
        try:
            out = subprocess.check_output([sys.executable, "-c", CHILD % {"root": ROOT, "n": n, "kind": kind}], env=e, text=True, timeout=600)
            res = json.loads(out.strip().splitlines()[-1])
        except Exception as ex:  # a switch that fails must not hide the others
            lines.append("%-45s FAILED: %s" % (tag, ex))
            continue
        if base is None:
            base = res
        m = res["median_ms"]
        same = res["words"] == base["words"] and res["digest"] == base["digest"]
        lines.append("%-45s build %.3f ms (pyramid %.3f + create %.3f)  words %s  %s" % (
            tag, m["pyramid"] + m["create_total"], m["pyramid"], m["create_total"], res["words"], "same words" if same else "WORDS DIFFER"))
        lines.append("    " + "  ".join("%s %.3f" % (k, v) for k, v in m.items() if k not in ("pyramid", "create_total")))
    text = "\n".join(["%dx%d %s, median of 6 builds after 3 warm-ups" % (n, n, kind)] + lines)
    print(text)
    with open(os.path.join(ROOT, "gpurun_out", "switch_compare.txt"), "w") as f:
        f.write(text + "\n")


if __name__ == "__main__":
    main()
