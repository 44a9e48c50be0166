"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle and the
committed golden vectors. Bit-exact throughout: pyramid floats, DAG words, per-level node counts,
lookup results. Nothing here reads /root/reference."""
import os

import numpy as np
import pytest

import cpvs_b200
from cpvs_b200 import synth
from test_oracle import TRAVERSE_8, TRAVERSE_16, TRAVERSE_32, sweep_points_32

pytestmark = pytest.mark.gpu


def _build(ctx, depth, zt=0, zn=1, leaf=True):
    mm = cpvs_b200.MinMaxHierarchy(depth, ctx)
    return mm, cpvs_b200.CompressedShadow.create(mm, zt, zn, leaf)


def _assert_same_dag(gpu_shadow, ora_shadow, tag):
    a, b = gpu_shadow.getDAG(), ora_shadow.dag()
    assert a.size == b.size, (tag, a.size, b.size)
    bad = np.nonzero(a != b)[0]
    assert bad.size == 0, (tag, "first differing word", int(bad[0]), hex(a[bad[0]]), hex(b[bad[0]]))


# ---- pyramid -------------------------------------------------------------------------------------

@pytest.mark.parametrize("n", [2, 4, 8, 16, 64, 128, 256, 1024])
def test_pyramid_bits(gpu_ctx, oracle, n):
    rng = np.random.default_rng(n)
    d = rng.random((n, n), dtype=np.float32)
    d[rng.random((n, n)) < 0.05] = 0.0
    d[rng.random((n, n)) < 0.02] = -0.0
    mm = cpvs_b200.MinMaxHierarchy(d, gpu_ctx)
    om = oracle.MinMax(d)
    assert mm.getNumLevels() == om.num_levels()
    for lvl in range(mm.getNumLevels()):
        assert np.array_equal(mm.getLevel(lvl).view(np.uint32), om.level(lvl).view(np.uint32)), lvl


def test_pyramid_reference_vectors(gpu_ctx, golden):
    vec, _ = golden
    for name in ("depths8x8", "depths16x16", "depths32x32", "minmax8x8", "minmax4x4"):
        mm = cpvs_b200.MinMaxHierarchy(vec[name], gpu_ctx)
        for lvl in range(1, mm.getNumLevels()):
            assert np.array_equal(mm.getLevel(lvl).view(np.uint32), vec["%s.minmax%d" % (name, lvl)].view(np.uint32))
    mm = cpvs_b200.MinMaxHierarchy(vec["minmax8x8"], gpu_ctx)  # MinMaxTest.get (reference test/MinMaxTest.cpp:41-55)
    assert mm.getNumLevels() == 4
    assert mm.getMin(0, 0, 0) == 0.0 and mm.getMax(1, 0, 0) == np.float32(0.9) and mm.getMax(2, 0, 0) == 1.0
    assert mm.getMax(3, 0, 0) == 1.0 and mm.getMin(3, 0, 0) == 0.0


# ---- DAG words -----------------------------------------------------------------------------------

@pytest.mark.parametrize("name", ["depths8x8", "depths16x16", "depths32x32", "minmax8x8"])
def test_reference_fixture_dags(gpu_ctx, golden, name):
    vec, _ = golden
    _, sh = _build(gpu_ctx, vec[name])
    assert np.array_equal(sh.getDAG(), vec[name + ".dag"])
    _, sh = _build(gpu_ctx, vec[name], leaf=False)
    assert np.array_equal(sh.getDAG(), vec[name + ".dag_noleaf"])
    if name == "depths16x16":
        for t in (0, 1):
            _, sh = _build(gpu_ctx, vec[name], t, 2)
            assert np.array_equal(sh.getDAG(), vec["%s.dag_z%dof2" % (name, t)])


def test_constant_maps(gpu_ctx, golden):
    vec, _ = golden
    for val, vis in ((0.0, cpvs_b200.SHADOW), (1.0, cpvs_b200.VISIBLE), (0.5, cpvs_b200.PARTIAL)):
        _, sh = _build(gpu_ctx, np.full((64, 64), val, np.float32))
        assert np.array_equal(sh.getDAG(), vec["const%.1f.dag" % val])
        assert sh.getTotalVisibility() == vis


def test_golden_synthetic_table(gpu_ctx, golden):
    """Words, FNV digests and lookup digests the unmodified reference produced (tests/golden)."""
    _, meta = golden
    pts = synth.lookups(meta["lookups"]["count"], meta["lookups"]["seed"])
    for row in meta["synthetic"]:
        _, sh = _build(gpu_ctx, synth.depth_map(row["kind"], row["n"]), row["z_tile"], row["z_num"], row["leafmasks"])
        dag = sh.getDAG()
        assert dag.size == row["words"], row
        assert "%016x" % synth.fnv64(dag) == row["fnv64"], row
        vis = sh.traverse(pts, row["leafmasks"])
        assert int((vis == 1).sum()) == row["lit"], row
        assert "%016x" % synth.fnv64(vis.astype(np.uint32)) == row["vis_fnv64"], row


@pytest.mark.parametrize("kind", ["plane", "terrain", "city"])
@pytest.mark.parametrize("n", [8, 16, 32, 128, 512, 2048])
def test_dag_equals_oracle(gpu_ctx, oracle, kind, n):
    d = synth.depth_map(kind, n)
    om = oracle.MinMax(d)
    mm = cpvs_b200.MinMaxHierarchy(d, gpu_ctx)
    cases = [(0, 1, True), (0, 1, False)] if n <= 128 else [(0, 1, True)]
    if n >= 16:
        cases += [(1, 2, True), (2, 4, True), (1, 3, True)]
    for zt, zn, leaf in cases:
        g = cpvs_b200.CompressedShadow.create(mm, zt, zn, leaf)
        o = oracle.Shadow(om, zt, zn, leaf)
        _assert_same_dag(g, o, (kind, n, zt, zn, leaf))
        svo, dagn, _ = g.level_counts()
        osvo, ouniq = o.level_counts()
        assert np.array_equal(svo, osvo), (kind, n, zt, zn, leaf)
        assert np.array_equal(dagn, ouniq), (kind, n, zt, zn, leaf)
        assert g.getTotalVisibility() == o.total_visibility()


def test_random_maps_equal_oracle(gpu_ctx, oracle):
    """High-entropy and plateau maps: many unique leaves, many duplicates, edge depths."""
    rng = np.random.default_rng(7)
    for n in (8, 16, 32, 64, 256):
        for trial in range(4):
            d = rng.random((n, n), dtype=np.float32)
            if trial == 1:
                d = (np.round(d * 8) / np.float32(8) * np.float32(0.999) + np.float32(0.0004)).astype(np.float32)
            if trial == 2:  # depths sitting exactly on slice mid-points and slice boundaries
                k = rng.integers(0, n, size=(n, n))
                d = ((k + rng.choice([0.0, 0.5], size=(n, n))) / n).astype(np.float32)
                d = np.clip(d + np.float32(1e-3) * (rng.random((n, n)) < 0.5), 0.001, 0.999).astype(np.float32)
            if trial == 3:
                d = np.repeat(np.repeat(rng.random((n // 8, n // 8), dtype=np.float32), 8, 0), 8, 1) * np.float32(0.97) + np.float32(0.013)
            for leaf in (True, False) if n <= 64 else (True,):
                _, g = _build(gpu_ctx, d, leaf=leaf)
                o = oracle.Shadow(oracle.MinMax(d), leafmasks=leaf)
                _assert_same_dag(g, o, (n, trial, leaf))


def test_z_boundary_depths(gpu_ctx, oracle):
    """Depths whose d*H lands exactly on integers / half-integers, and tiny depths (k rounding)."""
    n = 64
    vals = np.array([(i + f) / n for i in range(0, n, 3) for f in (0.0, 0.5, 0.4999999, 0.5000001)] +
                    [1e-8, 0.0078124995, 0.007812501, 0.99999994, 0.2499999, 0.25000003], np.float32)
    rng = np.random.default_rng(1)
    d = rng.choice(vals, size=(n, n)).astype(np.float32)
    d[:8, :8] = np.float32(0.3)  # keep the root PARTIAL but avoid SURVEY.md N2 (fully dyadic levels)
    d[8:16, :8] = np.float32(0.71)
    for zt, zn in ((0, 1), (1, 2)):
        _, g = _build(gpu_ctx, d, zt, zn)
        o = oracle.Shadow(oracle.MinMax(d), zt, zn)
        _assert_same_dag(g, o, (zt, zn))


# ---- lookups ---------------------------------------------------------------------------------------

@pytest.mark.parametrize("name,table,leaf", [("depths8x8", TRAVERSE_8, False), ("depths16x16", TRAVERSE_16, True),
                                             ("depths32x32", TRAVERSE_32, True)])
def test_traverse_known_answers(gpu_ctx, golden, name, table, leaf):
    """The reference's own traverse tests (test/CompressedShadowTest.cpp:50-152) on the CUDA lookup."""
    vec, _ = golden
    _, sh = _build(gpu_ctx, vec[name])
    pts = np.array([p for p, _ in table], np.float32)
    assert list(sh.traverse(pts, leaf)) == [v for _, v in table]
    if name == "depths32x32":
        vis, sha = sweep_points_32()
        assert (sh.traverse(vis) == 1).all() and (sh.traverse(sha) == 0).all()


@pytest.mark.parametrize("kind", ["plane", "terrain", "city"])
def test_lookups_equal_oracle(gpu_ctx, oracle, kind):
    pts = synth.lookups(300000)
    pts[:16] = np.array([[-1, -1, -1], [1, 1, 1], [1, -1, 1], [-1, 1, -1]] * 4, np.float32)
    for n, zt, zn, leaf in ((1024, 0, 1, True), (256, 1, 2, True), (64, 0, 1, False)):
        d = synth.depth_map(kind, n)
        _, g = _build(gpu_ctx, d, zt, zn, leaf)
        o = oracle.Shadow(oracle.MinMax(d), zt, zn, leaf)
        assert np.array_equal(g.traverse(pts, leaf), o.traverse(pts, leaf)), (kind, n)


def test_lookup_clamps_out_of_range(gpu_ctx, oracle):
    d = synth.depth_map("terrain", 128)
    _, g = _build(gpu_ctx, d)
    o = oracle.Shadow(oracle.MinMax(d))
    pts = np.array([[-3, 0, 0], [3, 0, 0], [0, -2, 5], [0, 0, -9], [np.nan, 0, 0], [np.inf, -np.inf, 0]], np.float32)
    assert np.array_equal(g.traverse(pts), o.traverse(pts))


def test_empty_lookup(gpu_ctx):
    _, g = _build(gpu_ctx, synth.depth_map("plane", 64))
    assert g.traverse(np.zeros((0, 3), np.float32)).size == 0


# ---- container / top-level grid -----------------------------------------------------------------------

@pytest.mark.parametrize("kind,n,length", [("terrain", 64, 2), ("city", 128, 2), ("plane", 32, 4)])
def test_container_equals_oracle(gpu_ctx, oracle, kind, n, length):
    cont = cpvs_b200.CompressedShadowContainer(length, gpu_ctx)
    ocont = oracle.Container(length)
    keep = []
    for y in range(length):
        for x in range(length):
            d = synth.depth_map(kind, n, (x, y), length)
            mm = cpvs_b200.MinMaxHierarchy(d, gpu_ctx)
            om = oracle.MinMax(d)
            for z in range(length):  # createShadowTiles (reference src/DeferredRenderer.cpp:150-163)
                cont.set(cpvs_b200.CompressedShadow.create(mm, z, length), x, y, z)
                osh = oracle.Shadow(om, z, length)
                keep.append(osh)
                ocont.set(osh, x, y, z)
    cont.copyToGPU()
    ocont.finalize()
    dag, grid = cont.dag_and_grid()
    odag, ogrid = ocont.dag_and_grid()
    assert np.array_equal(grid, ogrid) and np.array_equal(dag, odag)
    info = cont.info()
    assert info["grid_levels"] == {2: 1, 4: 2}[length] and info["grid_cells"] == length ** 3
    pts = synth.lookups(200000, seed=5)
    assert np.array_equal(cont.lookup_ndc(pts), ocont.lookup_ndc(pts))
    # evaluate(): world positions through an orthographic light matrix (column-major, glm order)
    rng = np.random.default_rng(0)
    pos = np.concatenate([rng.uniform(-10, 10, (48, 64, 3)), np.ones((48, 64, 1))], axis=2).astype(np.float32)
    m = np.array([[0.09, 0, 0, 0], [0, 0.1, 0, 0], [0.01, 0.02, 0.095, 0], [0.03, -0.02, 0.01, 1]], np.float32)  # m[col][row]
    assert np.array_equal(cont.evaluate(pos, m), ocont.evaluate(pos, m))


def test_single_shadow_container(gpu_ctx, oracle):
    """CompressedShadowContainer(unique_ptr<CompressedShadow>) (reference src/CompressedShadowContainer.h:28-32)."""
    d = synth.depth_map("city", 256)
    _, g = _build(gpu_ctx, d)
    cont = cpvs_b200.CompressedShadowContainer(g)
    cont.copyToGPU()
    pts = synth.lookups(50000, seed=8)
    assert np.array_equal(cont.lookup_ndc(pts), g.traverse(pts))
    dag, grid = cont.dag_and_grid()
    assert np.array_equal(dag, g.getDAG()) and list(grid) == [0]


# ---- full-size, size-independent properties ----------------------------------------------------------

@pytest.mark.parametrize("kind", ["terrain", "city"])
def test_large_map_decodes_to_depth(gpu_ctx, kind):
    """4096^2 (the oracle would take minutes): every looked-up voxel must decode to z+0.5 <= d*H, the
    per-level counts must be consistent, and a rebuild must give identical words (determinism)."""
    n = 4096
    d = synth.depth_map(kind, n)
    _, g = _build(gpu_ctx, d)
    pts = synth.lookups(1000000)
    path = (((pts + np.float32(1)) * np.float32(0.5)) * np.float32(n - 1)).astype(np.int32)
    lit = (path[:, 2].astype(np.float32) + np.float32(0.5)) <= d[path[:, 1], path[:, 0]] * np.float32(n)
    assert np.array_equal(g.traverse(pts), lit.astype(np.uint8))
    svo, dagn, words = g.level_counts()
    assert (dagn <= svo).all() and int(words.sum()) == int(g.info.words) and dagn[g.getNumLevels() - 2] == 1
    _, g2 = _build(gpu_ctx, d)
    assert synth.fnv64(g.getDAG()) == synth.fnv64(g2.getDAG())


def test_survey_4096_digests(gpu_ctx):
    """words / FNV of the reference at 4096^2 (libm-free generators only)."""
    for kind, words, digest in (("plane", 17495, "6cebb7d538e96084"), ("city", 176032, "ea44abc8910b0f2c")):
        _, g = _build(gpu_ctx, synth.depth_map(kind, 4096))
        dag = g.getDAG()
        assert dag.size == words and "%016x" % synth.fnv64(dag) == digest


# ---- errors ----------------------------------------------------------------------------------------

def test_argument_errors(gpu_ctx):
    with pytest.raises(cpvs_b200.CpvsError) as e:
        cpvs_b200.MinMaxHierarchy(np.zeros((12, 12), np.float32), gpu_ctx)
    assert e.value.code == cpvs_b200.EINVAL
    with pytest.raises(cpvs_b200.CpvsError):
        cpvs_b200.MinMaxHierarchy(np.zeros((8, 16), np.float32), gpu_ctx)
    mm = cpvs_b200.MinMaxHierarchy(np.zeros((4, 4), np.float32), gpu_ctx)
    with pytest.raises(cpvs_b200.CpvsError):  # numLevels > 3 (reference src/CompressedShadow.cpp:46)
        cpvs_b200.CompressedShadow.create(mm)
    mm = cpvs_b200.MinMaxHierarchy(synth.depth_map("plane", 64), gpu_ctx)
    with pytest.raises(cpvs_b200.CpvsError):
        cpvs_b200.CompressedShadow.create(mm, 2, 2)
    sh = cpvs_b200.CompressedShadow.create(cpvs_b200.MinMaxHierarchy(synth.depth_map("plane", 8), gpu_ctx))
    with pytest.raises(cpvs_b200.CpvsError):
        sh.traverse(np.zeros((1, 3), np.float32), True)
    cont = cpvs_b200.CompressedShadowContainer(2, gpu_ctx)
    with pytest.raises(cpvs_b200.CpvsError):
        cont.copyToGPU()


# ---- an octree that stops early (the reference reads out of bounds there) ----------------

@pytest.mark.parametrize("value", [0.25, 0.75, 0.5, 0.125])
def test_early_terminating_octree_decodes(gpu_ctx, oracle, value):
    """Dyadic constant planes make d*H an exact integer on every level: some level has nodes but no PARTIAL
    child. Not a parity target against the reference (undefined behaviour there); the defined result is a
    valid, shorter DAG -- identical to the oracle port's -- whose every voxel decodes to z + 0.5 <= d*H."""
    n = 64
    d = np.full((n, n), value, np.float32)
    for leaf in (True, False):
        _, g = _build(gpu_ctx, d, leaf=leaf)
        o = oracle.Shadow(oracle.MinMax(d), leafmasks=leaf)
        _assert_same_dag(g, o, (value, leaf))
        zs = (np.arange(n, dtype=np.float32) + np.float32(0.5)) / np.float32(n) * 2 - 1
        pts = np.array([[x, y, z] for z in zs for (x, y) in ((-0.9, -0.9), (0.3, 0.7))], np.float32)
        path = (((pts + np.float32(1)) * np.float32(0.5)) * np.float32(n - 1)).astype(np.int32)
        lit = (path[:, 2].astype(np.float32) + np.float32(0.5)) <= np.float32(value) * np.float32(n)
        assert np.array_equal(g.traverse(pts, leaf), lit.astype(np.uint8)), (value, leaf)


def test_mixed_map_with_flat_regions(gpu_ctx, oracle):
    """Half the map is a dyadic constant (subtrees that stop early), half is terrain."""
    n = 256
    d = synth.depth_map("terrain", n)
    d[:, : n // 2] = np.float32(0.5)
    d[: n // 4, :] = np.float32(1.0)
    for zt, zn in ((0, 1), (1, 2)):
        _, g = _build(gpu_ctx, d, zt, zn)
        o = oracle.Shadow(oracle.MinMax(d), zt, zn)
        _assert_same_dag(g, o, (zt, zn))
        pts = synth.lookups(100000, seed=21)
        assert np.array_equal(g.traverse(pts), o.traverse(pts))


def test_lazy_low_levels_and_childmask(gpu_ctx, oracle, golden):
    """Levels 1 and 2 are only materialised on demand (n >= 128): accessors, createChildmask and a
    leafmask-less build after a leafmask build must all see them."""
    d = synth.depth_map("city", 256)
    mm = cpvs_b200.MinMaxHierarchy(d, gpu_ctx)
    a = cpvs_b200.CompressedShadow.create(mm)  # leafmask build first: levels 1-2 not built yet
    om = oracle.MinMax(d)
    assert mm.createChildmask(1, 10, 20, 30) == om.childmask(1, 10, 20, 30)
    for lvl in (1, 2):
        assert np.array_equal(mm.getLevel(lvl).view(np.uint32), om.level(lvl).view(np.uint32))
    b = cpvs_b200.CompressedShadow.create(mm, leafmasks=False)
    assert np.array_equal(b.getDAG(), oracle.Shadow(om, leafmasks=False).dag())
    assert np.array_equal(a.getDAG(), oracle.Shadow(om).dag())
    vec, _ = golden  # testCreateChildmask.test8x8 (reference test/CompressedShadowUtilTest.cpp:14-19)
    assert cpvs_b200.MinMaxHierarchy(vec["depths8x8"], gpu_ctx).createChildmask(1, 2, 0, 0) == 0x88AA


def test_many_z_tiles(gpu_ctx, oracle):
    """16 z-slices of one pyramid (createShadowTiles with numSlices = 16): most slices are trivial."""
    n, zn = 128, 16
    d = synth.depth_map("terrain", n)
    mm = cpvs_b200.MinMaxHierarchy(d, gpu_ctx)
    om = oracle.MinMax(d)
    trivial = 0
    for zt in range(zn):
        g = cpvs_b200.CompressedShadow.create(mm, zt, zn)
        o = oracle.Shadow(om, zt, zn)
        _assert_same_dag(g, o, zt)
        trivial += int(g.info.words == 1)
    assert trivial >= 8


def test_concurrent_z_slices_share_one_pyramid(gpu_ctx, oracle):
    """createShadowTiles (reference src/DeferredRenderer.cpp:150-163): one thread per z-slice, all reading the
    same MinMaxHierarchy. Here every thread has its own context (stream + scratch arena)."""
    import threading
    n, zn = 256, 4
    d = synth.depth_map("terrain", n)
    mm = cpvs_b200.MinMaxHierarchy(d, gpu_ctx)
    om = oracle.MinMax(d)
    results, errors = {}, []

    def work(z):
        try:
            ctx = cpvs_b200.Context(0)
            for _ in range(3):
                results[z] = cpvs_b200.CompressedShadow.create(mm, z, zn, ctx=ctx).getDAG()
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(z,)) for z in range(zn)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for z in range(zn):
        assert np.array_equal(results[z], oracle.Shadow(om, z, zn).dag()), z


def test_container_file_round_trip(gpu_ctx, tmp_path):
    """On-disk container format: save, load, identical words and lookups; corrupt
    files are rejected."""
    n, length = 64, 2
    cont = cpvs_b200.CompressedShadowContainer(length, gpu_ctx)
    for y in range(length):
        for x in range(length):
            mm = cpvs_b200.MinMaxHierarchy(synth.depth_map("city", n, (x, y), length), gpu_ctx)
            for z in range(length):
                cont.set(cpvs_b200.CompressedShadow.create(mm, z, length), x, y, z)
    cont.copyToGPU()
    path = str(tmp_path / "shadow.cpvs")
    cont.save(path)
    back = cpvs_b200.CompressedShadowContainer.load(path, gpu_ctx)
    assert back.info() == cont.info()
    dag, grid = cont.dag_and_grid()
    dag2, grid2 = back.dag_and_grid()
    assert np.array_equal(dag, dag2) and np.array_equal(grid, grid2)
    pts = synth.lookups(50000, seed=4)
    assert np.array_equal(cont.lookup_ndc(pts), back.lookup_ndc(pts))
    good = open(path, "rb").read()
    bad = str(tmp_path / "bad.cpvs")

    def rejected(raw):
        open(bad, "wb").write(bytes(raw))
        with pytest.raises(cpvs_b200.CpvsError) as e:
            cpvs_b200.CompressedShadowContainer.load(bad, gpu_ctx)
        return e.value.code == cpvs_b200.EINVAL

    raw = bytearray(good)
    raw[-5] ^= 0x40  # a DAG word
    assert rejected(raw)
    raw = bytearray(good)
    raw[64 + 2] ^= 0x01  # a grid word: covered by the checksum too
    assert rejected(raw)
    raw = bytearray(good)
    raw[20] = 5  # gridLevels no longer log2(length): lookups would index past the grid
    assert rejected(raw)
    raw = bytearray(good)
    raw[24] = 7  # leafmasks must be 0 or 1
    assert rejected(raw)
    assert rejected(good[:-8])  # truncated: the size is checked against the header before anything is allocated
    raw = bytearray(good)
    raw[32:40] = (1 << 31).to_bytes(8, "little")  # dagWords far beyond the file
    assert rejected(raw)
    with pytest.raises(cpvs_b200.CpvsError):  # a loaded container is final
        mm = cpvs_b200.MinMaxHierarchy(synth.depth_map("city", n, (0, 0), length), gpu_ctx)
        back.set(cpvs_b200.CompressedShadow.create(mm, 0, length), 0, 0, 0)
    assert np.array_equal(cont.lookup_ndc(pts), back.lookup_ndc(pts))


def test_config2_tile_grid_4x4x4(gpu_ctx, oracle):
    """BASELINE configs[2] at reduced size: a 4x4 grid of depth tiles, 4 z-slices each = 64 DAGs in one cubic
    container (renderWithTiles / createShadowTiles, reference src/DeferredRenderer.cpp:150-187). Most outer
    z-slices miss the surface (one-word DAGs, grid sentinels); words, grid and lookups must equal the oracle's."""
    n, length = 256, 4
    cont = cpvs_b200.CompressedShadowContainer(length, gpu_ctx)
    ocont = oracle.Container(length)
    keep, trivial = [], 0
    for (x, y) in [(x, y) for y in range(length) for x in range(length)]:
        d = synth.depth_map("terrain", n, (x, y), length)
        mm = cpvs_b200.MinMaxHierarchy(d, gpu_ctx)
        om = oracle.MinMax(d)
        for z in range(length):
            g = cpvs_b200.CompressedShadow.create(mm, z, length)
            o = oracle.Shadow(om, z, length)
            keep.append(o)
            assert g.getTotalVisibility() == o.total_visibility()
            trivial += int(g.info.words == 1)
            cont.set(g, x, y, z)
            ocont.set(o, x, y, z)
    assert trivial >= 16
    cont.copyToGPU()
    ocont.finalize()
    dag, grid = cont.dag_and_grid()
    odag, ogrid = ocont.dag_and_grid()
    assert np.array_equal(grid, ogrid) and np.array_equal(dag, odag)
    pts = synth.lookups(300000, seed=12)
    assert np.array_equal(cont.lookup_ndc(pts), ocont.lookup_ndc(pts))


# ---- device-resident depth source ----------------------------------------------------

@pytest.mark.parametrize("kind,n,tiles", [("plane", 1024, 1), ("plane", 512, 4), ("city", 1024, 1), ("city", 512, 4),
                                          ("city", 256, 64), ("city", 16, 2), ("terrain_dev", 1024, 1), ("terrain_dev", 512, 4),
                                          ("terrain_dev", 256, 64), ("terrain_dev", 2048, 8)])
def test_device_generated_tiles_equal_host_bytes(gpu_ctx, kind, n, tiles):
    """The CUDA generator must write the same bytes as the host generator the oracle is fed with. 64 tiles per
    side of 256 texels = a 16K^2 city with 2048 boxes of up to 1032 texels: most of them cover whole 128x32
    regions of a tile (the scalar path of the kernel), the rest cut through regions."""
    import torch
    out = torch.empty((n, n), dtype=torch.float32, device="cuda:0")
    picks = [(x, y) for y in range(tiles) for x in range(tiles)]
    if len(picks) > 16:
        picks = picks[:: len(picks) // 16]
    for tile in picks:
        out.fill_(-1.0)
        torch.cuda.synchronize()  # the context runs on its own stream
        cpvs_b200.generate_depth(kind, n, out, tile, tiles, gpu_ctx)
        gpu_ctx.synchronize()
        host = synth.depth_map(kind, n, tile, tiles)
        assert np.array_equal(out.cpu().numpy().view(np.uint32), host.view(np.uint32)), (kind, tile)


def test_device_generated_tile_grid_equals_oracle(gpu_ctx, oracle):
    """BASELINE configs[4] at reduced size: city tiles generated on the device, built without leaving it,
    gathered into a cubic container; the oracle gets the host generator's bytes."""
    import torch
    n, length = 128, 4
    cont = cpvs_b200.CompressedShadowContainer(length, gpu_ctx)
    ocont = oracle.Container(length)
    keep = []
    depth = torch.empty((n, n), dtype=torch.float32, device="cuda:0")
    for (x, y) in [(x, y) for y in range(length) for x in range(length)]:
        cpvs_b200.generate_depth("city", n, depth, (x, y), length, gpu_ctx)
        mm = cpvs_b200.MinMaxHierarchy(depth, gpu_ctx)
        om = oracle.MinMax(synth.depth_map("city", n, (x, y), length))
        for z in range(length):
            g = cpvs_b200.CompressedShadow.create(mm, z, length)
            o = oracle.Shadow(om, z, length)
            keep.append(o)
            cont.set(g, x, y, z)
            ocont.set(o, x, y, z)
        mm.close()
    cont.copyToGPU()
    ocont.finalize()
    dag, grid = cont.dag_and_grid()
    odag, ogrid = ocont.dag_and_grid()
    assert np.array_equal(grid, ogrid) and np.array_equal(dag, odag)
    pts = synth.lookups(200000, seed=5)
    assert np.array_equal(cont.lookup_ndc(pts), ocont.lookup_ndc(pts))


def test_depth_generate_errors(gpu_ctx):
    import torch
    out = torch.empty((64, 64), dtype=torch.float32, device="cuda:0")
    lib = cpvs_b200.load_library()
    assert lib.cpvs_depth_generate(gpu_ctx.handle, 1, 64, 0, 0, 1, out.data_ptr()) == cpvs_b200.EINVAL  # terrain: host libm
    assert lib.cpvs_depth_generate(gpu_ctx.handle, 0, 64, 2, 0, 2, out.data_ptr()) == cpvs_b200.EINVAL
    assert lib.cpvs_depth_generate(gpu_ctx.handle, 0, 64, 0, 0, 1, None) == cpvs_b200.EINVAL


# ---- whole tile grids (cpvs_b200.gridbuild: what bench.py --grid runs) -------------------------------

@pytest.mark.parametrize("kind,tile,length", [("city", 256, 4), ("terrain_dev", 128, 4), ("plane", 64, 8)])
def test_gridbuild_small(kind, tile, length):
    """The one-process-per-GPU grid driver at a small size (a single rank here; tests/test_gpu_grid_ranks.py runs two): every
    container lookup is checked against the depth tiles inside run(); the grid must equal the host scan."""
    from cpvs_b200 import gridbuild
    ctx = cpvs_b200.Context(0)
    res = gridbuild.run(ctx, tile, length, kind, lookups=3840 * 64, lookup_iters=2)
    ctx.close()
    assert res["cells"] == length ** 3
    assert res["verified"]["container_lookups_vs_depth"] == res["lookups"] == 3840 * 64
    assert res["grid_cells_with_dag"] >= res["cells"] - res["one_word_cells"]  # one-word cells may still mix lit and shadow
    assert 0 < res["lookups_lit"] < res["lookups"] and res["dag_nodes"] >= length ** 3


def test_reserve_and_kept_dags(gpu_ctx, oracle):
    """cpvs_ctx_reserve only moves where the memory comes from: words stay identical while many DAGs are kept alive
    and the scratch arena regrows (a larger map after a smaller one)."""
    gpu_ctx.reserve(64 << 20)
    gpu_ctx.reserve(0)
    kept = []
    for n, kind in ((64, "terrain"), (256, "city"), (1024, "terrain"), (128, "plane"), (2048, "city")):
        d = synth.depth_map(kind, n)
        _, g = _build(gpu_ctx, d)
        kept.append((g, oracle.Shadow(oracle.MinMax(d))))
    for g, o in kept:
        _assert_same_dag(g, o, "kept")


# ---- leaves built per column (svo.cu buildLeafColumnsKernel) -------------------------------------------

@pytest.mark.parametrize("mode", ["0", "2"])
def test_leaf_kernels_forced(oracle, mode, monkeypatch):
    """Both leaf builders on everything: CPVS_LEAF_COLUMNS=2 forces the per-column kernel onto the cases the default
    heuristic keeps away from it (columns of hundreds of leaves that cross the 252-block chunks, z-slices with mostly empty
    columns, a 2x2-column map), =0 forces the per-leaf kernel onto smooth surfaces. Same words either way."""
    monkeypatch.setenv("CPVS_LEAF_COLUMNS", mode)
    ctx = cpvs_b200.Context(0)  # the switch is read when the context is created
    rng = np.random.default_rng(5)
    cases = [("terrain", synth.depth_map("terrain", 512), 0, 1), ("plane", synth.depth_map("plane", 256), 0, 1),
             ("city", synth.depth_map("city", 1024), 0, 1), ("city z1/2", synth.depth_map("city", 512), 1, 2),
             ("terrain z2/4", synth.depth_map("terrain", 256), 2, 4), ("random 16", rng.random((16, 16), dtype=np.float32), 0, 1),
             ("random 128", rng.random((128, 128), dtype=np.float32), 0, 1)]
    wall = np.full((4096, 4096), 0.97, np.float32)  # a cliff: columns of ~450 leaves, in x- and in y-direction
    wall[:, 2001:] = 0.05
    wall[3000:, :] = 0.5
    wall += (rng.random((4096, 4096), dtype=np.float32) * np.float32(1e-3))
    cases.append(("cliff", wall, 0, 1))
    steps = synth.depth_map("terrain", 512).copy()  # terraces: residues that sit exactly on block boundaries, and NaN-free extremes
    steps = np.round(steps * np.float32(64)) / np.float32(64)
    steps[:64, :64] = 0.0
    steps[-64:, -64:] = 1.0
    cases += [("terraces", steps, 0, 1), ("terraces z1/3", steps, 1, 3), ("terrain z0/4", synth.depth_map("terrain", 1024), 0, 4),
              ("terrain z3/4", synth.depth_map("terrain", 1024), 3, 4), ("city z5/16", synth.depth_map("city", 512), 5, 16)]
    for tag, d, zt, zn in cases:
        o = oracle.Shadow(oracle.MinMax(d), zt, zn)
        # hierarchy prepared for whole-volume builds (its column residues do not fit a sliced build: depth path) and for this slicing
        for tiles in sorted({1, zn}):
            mm = cpvs_b200.MinMaxHierarchy(d, ctx, zTileNum=tiles)
            g = cpvs_b200.CompressedShadow.create(mm, zt, zn)
            _assert_same_dag(g, o, (mode, tag, tiles))


# ---- builds sized from the previous build of the same shape (cpvs_ctx_set_prediction) ---------------------------------------

def _terrain_like(n, seed, amp):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:n, 0:n].astype(np.float32) / np.float32(n)
    d = 0.5 + amp * np.sin(9.1 * x) * np.cos(7.3 * y) + 0.05 * np.sin(41.0 * x + 3.0 * y)
    return (d + rng.random((n, n)) * 1e-3).astype(np.float32)


@pytest.mark.parametrize("columns", ["1", "2"])
def test_predicted_builds_keep_the_words(oracle, columns, monkeypatch):
    """The second build of a shape runs on predicted sizes (no count pass, the DAG allocated up front, leaves emitted during the
    merge); the words must be those of the exact build and of the oracle -- also when the map changed in between."""
    monkeypatch.setenv("CPVS_LEAF_COLUMNS", columns)
    ctx = cpvs_b200.Context(0)
    maps = [_terrain_like(1024, 1, 0.15), _terrain_like(1024, 2, 0.16), _terrain_like(1024, 3, 0.14), synth.depth_map("terrain", 1024)]
    for i, d in enumerate(maps):
        mm = cpvs_b200.MinMaxHierarchy(d, ctx)
        g = cpvs_b200.CompressedShadow.create(mm)
        if i < 3:  # (the last map may outgrow what its noisy predecessors predicted: rebuilt exactly, same words)
            assert bool(g.info.predicted) == (i > 0), i
        _assert_same_dag(g, oracle.Shadow(oracle.MinMax(d)), ("predicted", columns, i))
        pts = synth.lookups(5000, seed=i + 1)
        assert np.array_equal(g.traverse(pts), oracle.Shadow(oracle.MinMax(d)).traverse(pts))
    st = ctx.stats()
    assert st["predicted_builds"] == 3 and st["overflow_rebuilds"] <= 1, st
    for leaf in (True, False):  # leafmask-less builds keep their own memo
        for rep in range(2):
            d = synth.depth_map("city", 256)
            mm = cpvs_b200.MinMaxHierarchy(d, ctx)
            g = cpvs_b200.CompressedShadow.create(mm, leafmasks=leaf)
            _assert_same_dag(g, oracle.Shadow(oracle.MinMax(d), 0, 1, leaf), ("predicted city", leaf, rep))


@pytest.mark.parametrize("staging", ["default", "0"])
def test_z_slices_in_flight_staged_and_unstaged(oracle, staging, monkeypatch):
    """Builds with exact node counts (the z-slices of a tile) emit their DAG into a staging buffer bounded by those counts and
    copy it into an allocation of its size; with staging switched off they take the older paths (capacity scaled from the memo
    of another tile, re-emission or rebuild when it does not suffice; count-then-allocate without a memo). Several slices in
    flight on two contexts, tiles of different weight one after the other: the words are the oracle's either way."""
    if staging != "default":
        monkeypatch.setenv("CPVS_STAGING_MAX_WORDS", staging)
    ctxs = [cpvs_b200.Context(0), cpvs_b200.Context(0)]
    n, zn = 512, 4
    maps = [synth.depth_map("plane", n), synth.depth_map("terrain", n), synth.depth_map("city", n), synth.depth_map("terrain", n, (1, 0), 2)]
    for i, d in enumerate(maps):
        mm = cpvs_b200.MinMaxHierarchy(d, ctxs[0], zTileNum=zn)
        om = oracle.MinMax(d)
        flying = [cpvs_b200.CompressedShadow.create(mm, z, zn, ctx=ctxs[z & 1], wait=False) for z in range(zn)]
        for z in reversed(range(zn)):  # finished out of order
            _assert_same_dag(flying[z], oracle.Shadow(om, z, zn), ("slices", staging, i, z))
    stats = [c.stats() for c in ctxs]
    if staging == "default":
        assert all(s["overflow_rebuilds"] == 0 and s["reemissions"] == 0 for s in stats), stats


def test_prediction_overflow_falls_back_to_exact(oracle):
    """A map that outgrows the capacities predicted from its predecessor is rebuilt with exact counts; one that only outgrows
    the predicted DAG allocation is emitted again. Either way the words are the oracle's."""
    ctx = cpvs_b200.Context(0)
    small, big = synth.depth_map("plane", 512), synth.depth_map("terrain", 512)
    for d in (small, big, small, big):
        mm = cpvs_b200.MinMaxHierarchy(d, ctx)
        _assert_same_dag(cpvs_b200.CompressedShadow.create(mm), oracle.Shadow(oracle.MinMax(d)), "overflow")
    st = ctx.stats()
    assert st["overflow_rebuilds"] >= 1, st
    # no head room at all: the slightest growth overflows -- nodes (rebuild) or only words (re-emission)
    ctx2 = cpvs_b200.Context(0)
    ctx2.set_prediction(True, 40)
    rng = np.random.default_rng(3)
    base = synth.depth_map("plane", 512)  # nearly all of its leaves are duplicates: every disturbed texel adds a distinct one
    for i in range(6):
        d = base.copy()
        ys, xs = rng.integers(0, 512, 1500 * i), rng.integers(0, 512, 1500 * i)
        d[ys, xs] += np.float32(0.0005)  # a quarter of a slice: more distinct leaves (words), hardly any more nodes
        mm = cpvs_b200.MinMaxHierarchy(d, ctx2)
        _assert_same_dag(cpvs_b200.CompressedShadow.create(mm), oracle.Shadow(oracle.MinMax(d)), ("tight words", i))
    st2 = ctx2.stats()
    assert st2["overflow_rebuilds"] + st2["reemissions"] >= 1, st2
    for i in range(4):
        d = base.copy()
        ys, xs = rng.integers(0, 512, 3000 * i), rng.integers(0, 512, 3000 * i)
        d[ys, xs] += np.float32(0.05)  # spikes: more nodes on every level
        mm = cpvs_b200.MinMaxHierarchy(d, ctx2)
        _assert_same_dag(cpvs_b200.CompressedShadow.create(mm), oracle.Shadow(oracle.MinMax(d)), ("tight nodes", i))
    st3 = ctx2.stats()
    assert st3["overflow_rebuilds"] > st2["overflow_rebuilds"], (st2, st3)
    # prediction switched off: every build counts first
    ctx3 = cpvs_b200.Context(0)
    ctx3.set_prediction(False)
    for _ in range(2):
        mm = cpvs_b200.MinMaxHierarchy(big, ctx3)
        g = cpvs_b200.CompressedShadow.create(mm)
        assert not g.info.predicted
        _assert_same_dag(g, oracle.Shadow(oracle.MinMax(big)), "unpredicted")
    assert ctx3.stats()["predicted_builds"] == 0


def test_lookup_leafmask_mismatch_is_an_error(gpu_ctx):
    """traverse(..., tryLeafmasks=False) on a leafmask DAG walked leaf words as pointers."""
    _, with_leaf = _build(gpu_ctx, synth.depth_map("terrain", 64))
    _, without = _build(gpu_ctx, synth.depth_map("terrain", 64), leaf=False)
    pts = synth.lookups(100)
    with pytest.raises(cpvs_b200.CpvsError) as e:
        with_leaf.traverse(pts, False)
    assert e.value.code == cpvs_b200.EINVAL
    with pytest.raises(cpvs_b200.CpvsError):
        without.traverse(pts, True)
    assert np.array_equal(with_leaf.traverse(pts, True), without.traverse(pts, False))


def test_async_builds_in_flight(oracle):
    """cpvs_shadow_create_async: several builds in flight on one context and on two contexts of the same GPU; every handle
    ends up with the oracle's words, also when a prediction fails while later builds are already enqueued."""
    ctxs = [cpvs_b200.Context(0), cpvs_b200.Context(0)]
    maps = [_terrain_like(512, s, 0.15) for s in (1, 2, 3, 4)]
    spiky = maps[0].copy()
    spiky[::7, ::5] += np.float32(0.07)  # outgrows whatever the smooth maps predicted
    want = [oracle.Shadow(oracle.MinMax(d)).dag() for d in maps + [spiky]]
    for ctx in ctxs:  # first build of the shape: exact, fills the memo
        mm = cpvs_b200.MinMaxHierarchy(maps[0], ctx)
        assert np.array_equal(cpvs_b200.CompressedShadow.create(mm).getDAG(), want[0])
    order = [0, 1, 2, 4, 3, 0, 4, 1]
    flying = []
    for k, idx in enumerate(order):
        ctx = ctxs[k % 2]
        d = spiky if idx == 4 else maps[idx]
        mm = cpvs_b200.MinMaxHierarchy(d, ctx)
        flying.append((idx, mm, cpvs_b200.CompressedShadow.create(mm, wait=False)))
    for idx, mm, sh in reversed(flying):  # finished out of order
        assert np.array_equal(sh.getDAG(), want[idx]), idx
    stats = [c.stats() for c in ctxs]
    assert sum(s["overflow_rebuilds"] for s in stats) >= 1, stats
    pts = synth.lookups(2000)
    idx, mm, sh = flying[0]
    assert np.array_equal(sh.traverse(pts), oracle.Shadow(oracle.MinMax(maps[idx])).traverse(pts))
