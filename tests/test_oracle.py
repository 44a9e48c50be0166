"""CPU suite: pins the oracle (oracle/oracle_port.cpp) against the reference's own known-answer tests
and against golden vectors produced by the unmodified reference (tests/golden/make_golden.py).

Where oracle/_ref exists (the compiled reference, built from /root/reference or prebuilt and shipped),
the port is additionally compared word-for-word with the reference itself on fresh inputs.
"""
import numpy as np
import pytest

from cpvs_b200 import synth

SH, VI = 0, 1  # CompressedShadow::SHADOW / VISIBLE


def _dag_hex(words):
    return " ".join("%x" % w for w in words)


# ---- reference test/MinMaxTest.cpp ---------------------------------------------------------------

def test_minmax_get(oracle, golden):
    """MinMaxTest.get (reference test/MinMaxTest.cpp:41-55)."""
    vec, _ = golden
    mm = oracle.MinMax(vec["minmax8x8"])
    assert mm.num_levels() == 4
    f = np.float32
    assert mm.level(0)[0, 0] == f(0.0)
    assert mm.level(1)[0, 0, 0] == f(0.0) and mm.level(1)[0, 0, 1] == f(0.9)
    assert mm.level(2)[0, 0, 1] == f(1.0)
    assert mm.level(3)[0, 0, 1] == f(1.0) and mm.level(3)[0, 0, 0] == f(0.0)


def test_minmax_create4x4(oracle, golden):
    """MinMaxTest.create4x4 (reference test/MinMaxTest.cpp:57-71)."""
    vec, _ = golden
    l1 = oracle.MinMax(vec["minmax4x4"]).level(1)
    assert l1[0, 0, 0] == 0.0 and l1[0, 0, 1] == 0.0 and l1[0, 1, 1] == 1.0


def test_minmax_32x32(oracle, golden):
    """MinMaxTest.test32x32 (reference test/MinMaxTest.cpp:77-90)."""
    vec, _ = golden
    mm = oracle.MinMax(vec["depths32x32"])
    assert mm.num_levels() == 6
    l4, l2 = mm.level(4), mm.level(2)
    assert abs(l4[1, 0, 0] - 0.673203) < 1e-6 and abs(l4[0, 0, 0] - 0.63008) < 1e-6
    assert abs(l4[0, 1, 0] - 0.63008) < 1e-6 and abs(l4[1, 1, 0] - 0.700469) < 1e-6
    assert abs(l2[1, 1, 0] - 0.63008) < 1e-6


@pytest.mark.parametrize("name", ["depths8x8", "depths16x16", "depths32x32", "minmax8x8", "minmax4x4"])
def test_minmax_levels_match_reference(oracle, golden, name):
    vec, _ = golden
    mm = oracle.MinMax(vec[name])
    for lvl in range(1, mm.num_levels()):
        assert np.array_equal(mm.level(lvl).view(np.uint32), vec["%s.minmax%d" % (name, lvl)].view(np.uint32))


# ---- reference test/CompressedShadowUtilTest.cpp ---------------------------------------------------

def test_create_childmask_8x8(oracle, golden):
    """testCreateChildmask.test8x8 (reference test/CompressedShadowUtilTest.cpp:14-19): 0x88aa."""
    vec, _ = golden
    assert oracle.MinMax(vec["depths8x8"]).childmask(1, 2, 0, 0) == 0x88AA


def test_merge_level_simple(oracle):
    """mergeLevelTest.testSimpleLevel (reference test/CompressedShadowUtilTest.cpp:83-97)."""
    level = np.array([0xAAAA, 10, 42, 0, 0, 1, 2, 3, 4,
                      0xAAA0, 10, 0, 0, 0, 0, 0, 0, 0,
                      0xAAAA, 10, 42, 0, 0, 1, 2, 3, 4], np.uint32)
    kept, merged, mapping = oracle.merge_level(level, 9)
    assert kept == 2
    assert mapping[0] == 0 and mapping[1] == 9 and mapping[2] == 0
    assert merged[9] == 0xAAA0


def test_merge_level_random_first_occurrence(oracle):
    rng = np.random.default_rng(5)
    pool = rng.integers(0, 50, size=(40, 17), dtype=np.uint32)
    picks = rng.integers(0, 40, size=2000)
    level = pool[picks].reshape(-1)
    kept, merged, mapping = oracle.merge_level(level, 17)
    first = {}
    for i, p in enumerate(picks):
        key = pool[p].tobytes()
        first.setdefault(key, len(first))
        assert mapping[i] == first[key] * 17
    assert kept == len(first)


# ---- reference test/CompressedShadowTest.cpp ---------------------------------------------------------

TRAVERSE_8 = [((-1, 1, -1), SH), ((-1, -1, 0), VI), ((1, 1, 0), SH), ((0, 0, 0), SH), ((0.5, 0.5, 0.5), SH),
              ((1, 1, -0.75), VI), ((1, 1, 0.5), SH), ((1, -1, 0.99), VI), ((0.7, -1, 0.99), SH),
              ((.45, -1, -0.9), SH), ((0.1, -1.0, -0.9), SH)]  # test/CompressedShadowTest.cpp:50-88
_S16 = 2.0 / 16.0
TRAVERSE_16 = [((1, 1, 0), VI), ((-1, -1, -1.0 + 3 * _S16), VI), ((-1, -_S16, -1.0 + 3 * _S16), VI), ((-1, 1, -1), SH),
               ((0, 0, -.9), VI), ((0 + _S16, 0, .9), SH), ((0 + 3 * _S16, 0, -.9), VI)]  # :90-117
_S32 = 2.0 / 32.0
TRAVERSE_32 = [((1, 1, 0), VI), ((-1.0 + 7 * _S32, 1.0 - 7 * _S32, 0.5), VI), ((-1.0 + 7 * _S32, 1.0 - 7 * _S32, 0.6), SH),
               ((-1.0 + 7 * _S32, -1.0 + 8 * _S32, 0.3), SH)]  # :119-136


def sweep_points_32():
    """The two sweeps of testTraverse32x32 (reference test/CompressedShadowTest.cpp:138-151)."""
    f = np.float32
    vis = [(f(x / 32.0) * f(2) - f(1), f(y / 32.0) * f(2) - f(1), f(0.59) * f(2) - f(1)) for y in range(31) for x in range(31)]
    sha = [(f(x / 32.0) * f(2) - f(1), f(y / 32.0) * f(2) - f(1), f(0.79) * f(2) - f(1)) for y in range(8, 26) for x in range(7, 27)]
    return np.array(vis, np.float32), np.array(sha, np.float32)


@pytest.mark.parametrize("name,table,leaf", [("depths8x8", TRAVERSE_8, False), ("depths16x16", TRAVERSE_16, True),
                                             ("depths32x32", TRAVERSE_32, True)])
def test_traverse_known_answers(oracle, golden, name, table, leaf):
    vec, _ = golden
    sh = oracle.Shadow(oracle.MinMax(vec[name]))
    pts = np.array([p for p, _ in table], np.float32)
    assert list(sh.traverse(pts, leaf)) == [v for _, v in table]
    if name == "depths32x32":
        vis, sha = sweep_points_32()
        assert (sh.traverse(vis) == VI).all() and (sh.traverse(sha) == SH).all()


@pytest.mark.parametrize("name", ["depths8x8", "depths16x16", "depths32x32", "minmax8x8"])
def test_dag_words_match_reference(oracle, golden, name):
    vec, _ = golden
    mm = oracle.MinMax(vec[name])
    assert _dag_hex(oracle.Shadow(mm).dag()) == _dag_hex(vec[name + ".dag"])
    assert _dag_hex(oracle.Shadow(mm, leafmasks=False).dag()) == _dag_hex(vec[name + ".dag_noleaf"])
    if name == "depths16x16":
        for t in (0, 1):
            assert _dag_hex(oracle.Shadow(mm, t, 2).dag()) == _dag_hex(vec["%s.dag_z%dof2" % (name, t)])


def test_survey_goldens(oracle, golden):
    """Spot values recorded in SURVEY.md 8c from the reference."""
    vec, _ = golden
    assert _dag_hex(vec["depths8x8.dag"]).startswith("a8a 6 7 e 6 14 1111 88aa")
    assert _dag_hex(vec["depths16x16.dag"][:9]) == "aaaa 9 14 14 25 36 14 14 14"
    assert vec["depths32x32.dag"].size == 247 and _dag_hex(vec["depths32x32.dag"][:5]) == "aa55 5 d 15 1c"
    assert list(vec["const1.0.dag"]) == [0x5555] and list(vec["const0.0.dag"]) == [0] and vec["const0.5.dag"].size == 1
    for val in (0.0, 0.5, 1.0):
        d = np.full((64, 64), val, np.float32)
        assert np.array_equal(oracle.Shadow(oracle.MinMax(d)).dag(), vec["const%.1f.dag" % val])


def test_synthetic_table(oracle, golden):
    """Port == reference on every synthetic generator (words, FNV digest, lookup results)."""
    _, meta = golden
    pts = synth.lookups(meta["lookups"]["count"], meta["lookups"]["seed"])
    for row in meta["synthetic"]:
        d = synth.depth_map(row["kind"], row["n"])
        sh = oracle.Shadow(oracle.MinMax(d), row["z_tile"], row["z_num"], row["leafmasks"])
        dag = sh.dag()
        assert dag.size == row["words"], row
        assert "%016x" % synth.fnv64(dag) == row["fnv64"], row
        vis = sh.traverse(pts, row["leafmasks"])
        assert int((vis == 1).sum()) == row["lit"], row
        assert "%016x" % synth.fnv64(vis.astype(np.uint32)) == row["vis_fnv64"], row


def test_survey_synthetic_digests(oracle):
    """SURVEY.md 8c table, 1024^2: words and FNV digests recorded from the reference during the survey."""
    expect = {"plane": (6078, "66d2a9503bd42163"), "terrain": (662932, "4c8c05979424c852"), "city": (33073, "2f5d3f2103efeca3")}
    for kind, (words, digest) in expect.items():
        dag = oracle.Shadow(oracle.MinMax(synth.depth_map(kind, 1024))).dag()
        assert dag.size == words and "%016x" % synth.fnv64(dag) == digest
    svo, uniq = oracle.Shadow(oracle.MinMax(synth.depth_map("terrain", 1024))).level_counts()
    assert list(svo[2:10][::-1]) == [1, 8, 32, 164, 784, 3452, 14068, 54314]
    assert list(uniq[2:10][::-1]) == [1, 8, 32, 164, 784, 3406, 13045, 39551]


def test_decoded_visibility_property(oracle):
    """Every voxel decodes to (z + 0.5 <= depth * H): the DAG is lossless (size-independent check)."""
    pts = synth.lookups(50000, seed=99)
    for kind in ("plane", "terrain", "city"):
        for n, zt, zn in ((256, 0, 1), (128, 1, 2)):
            d = synth.depth_map(kind, n)
            sh = oracle.Shadow(oracle.MinMax(d), zt, zn)
            path = (((pts + np.float32(1)) * np.float32(0.5)) * np.float32(n - 1)).astype(np.int32)
            z = path[:, 2] + zt * n
            lit = (z.astype(np.float32) + np.float32(0.5)) <= d[path[:, 1], path[:, 0]] * np.float32(n * zn)
            assert np.array_equal(sh.traverse(pts), lit.astype(np.uint8)), (kind, n, zt, zn)


def test_container_grid(oracle):
    """combineDAGs / createTopLevelGrid (reference src/CompressedShadowContainer.cpp:52-91)."""
    n, length = 32, 2
    cont = oracle.Container(length)
    shadows, mms = {}, {}
    for y in range(length):
        for x in range(length):
            mms[x, y] = oracle.MinMax(synth.depth_map("terrain", n, (x, y), length))
            for z in range(length):
                shadows[x, y, z] = oracle.Shadow(mms[x, y], z, length)
                cont.set(shadows[x, y, z], x, y, z)
    cont.finalize()
    dag, grid = cont.dag_and_grid()
    offset = 0
    for z in range(length):
        for y in range(length):
            for x in range(length):
                words = shadows[x, y, z].dag()
                tv = shadows[x, y, z].total_visibility()
                want = {0: 0x0FFFFFFF, 1: 0x0FFFFFFE}.get(tv, offset)
                assert grid[(z * length + y) * length + x] == want
                assert np.array_equal(dag[offset:offset + words.size], words)
                offset += words.size
    # lookups over the virtual volume decode to the depth of the tile the point falls in
    pts = synth.lookups(20000, seed=3)
    res = n * length
    path = (((pts + np.float32(1)) * np.float32(0.5)) * np.float32(res - 1)).astype(np.int32)
    full = np.block([[synth.depth_map("terrain", n, (x, y), length) for x in range(length)] for y in range(length)])
    lit = (path[:, 2].astype(np.float32) + np.float32(0.5)) <= full[path[:, 1], path[:, 0]] * np.float32(n * length)
    assert np.array_equal(cont.lookup_ndc(pts), lit.astype(np.uint8))


# ---- the port against the compiled reference itself (only where oracle/_ref exists) -------------------

def _need_ref(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("kind", ["plane", "terrain", "city"])
def test_port_equals_reference(oracle, kind):
    _need_ref(oracle)
    pts = synth.lookups(20000, seed=11)
    for n in (16, 64, 256):
        d = synth.depth_map(kind, n)
        mr, mp = oracle.MinMax(d, "ref"), oracle.MinMax(d, "port")
        for lvl in range(mr.num_levels()):
            assert np.array_equal(mr.level(lvl).view(np.uint32), mp.level(lvl).view(np.uint32))
        for zt, zn in ((0, 1), (1, 2), (1, 3)):
            assert np.array_equal(mr.svo(zt, zn)[0], mp.svo(zt, zn)[0])
            sr, sp = oracle.Shadow(mr, zt, zn), oracle.Shadow(mp, zt, zn)
            assert np.array_equal(sr.dag(), sp.dag())
            assert np.array_equal(sr.traverse(pts), sp.traverse(pts))
        mn = oracle.MinMax(d, "ref_noleaf")
        assert np.array_equal(oracle.Shadow(mn, leafmasks=False).dag(), oracle.Shadow(mp, leafmasks=False).dag())


def test_port_equals_reference_random_maps(oracle):
    _need_ref(oracle)
    rng = np.random.default_rng(2024)
    for n in (8, 16, 32, 64):
        for trial in range(4):
            d = rng.random((n, n), dtype=np.float32)
            if trial % 2:
                d = np.round(d * 8) / np.float32(8) * np.float32(0.999) + np.float32(0.0004)  # plateaus -> many duplicates
            a = oracle.Shadow(oracle.MinMax(d, "ref")).dag()
            b = oracle.Shadow(oracle.MinMax(d, "port")).dag()
            assert np.array_equal(a, b), (n, trial)


def test_reference_gtests_pass(oracle):
    """The reference's own 18 gtests, compiled unmodified by oracle/Makefile."""
    import os
    import subprocess
    exe = os.path.join(oracle.HERE, "_ref", "runUnitTests")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/runUnitTests not built")
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "[  PASSED  ] 18 tests." in out.stdout
