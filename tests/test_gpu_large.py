"""GPU suite, headline sizes (-m gpu): the 8192^2 / 16384^2 maps of BASELINE configs[1], configs[2] and configs[4] against
tests/golden/port_large.json -- word counts, FNV-64 digests of all words, per-level node counts and the digest of 1 M
lookups, produced by the CPU restatement (oracle/oracle_port.cpp), which the CPU suite pins word for word to the compiled
reference at every size the reference finishes. Nothing here reads /root/reference or runs the oracle."""
import json
import os

import numpy as np
import pytest

import cpvs_b200
from cpvs_b200 import synth

pytestmark = pytest.mark.gpu

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "port_large.json")) as f:
    LARGE = json.load(f)["maps"]


def _id(row):
    tag = "%s-%d" % (row["kind"], row["n"])
    if row["tiles_per_side"] > 1:
        tag += "-tile%d.%dof%d" % (row["tile"][0], row["tile"][1], row["tiles_per_side"])
    if row["z_num"] > 1:
        tag += "-z%dof%d" % (row["z_tile"], row["z_num"])
    return tag


@pytest.mark.parametrize("row", LARGE, ids=_id)
def test_headline_maps_equal_port_words(gpu_ctx, row):
    n = row["n"]
    d = synth.depth_map(row["kind"], n, tuple(row["tile"]), row["tiles_per_side"])
    pts = synth.lookups(1000000)
    dags = []
    # exact build, then the same shape again (sizes predicted, leaves emitted during the merge), then with a hierarchy
    # prepared for the slicing (column residues)
    for rep, tiles in enumerate((1, 1, row["z_num"])):
        mm = cpvs_b200.MinMaxHierarchy(d, gpu_ctx, zTileNum=tiles)
        g = cpvs_b200.CompressedShadow.create(mm, row["z_tile"], row["z_num"])
        dag = g.getDAG()
        assert dag.size == row["words"], (rep, dag.size, row["words"])
        assert "%016x" % synth.fnv64(dag) == row["fnv64"], rep
        svo, dagn, _ = g.level_counts()
        assert [int(v) for v in svo] == row["svo_nodes"] and [int(v) for v in dagn] == row["dag_nodes"], rep
        if rep == 0:
            vis = g.traverse(pts)
            assert int((vis == 1).sum()) == row["lit"]
            assert "%016x" % synth.fnv64(vis.astype(np.uint32)) == row["vis_fnv64"]
        dags.append(dag)
        g.close()
        mm.close()
    assert np.array_equal(dags[0], dags[1]) and np.array_equal(dags[0], dags[2])
