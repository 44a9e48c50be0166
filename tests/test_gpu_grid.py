"""GPU suite (-m gpu): the C++ tile-grid driver (cpvs_grid_* of include/cpvs_b200.h) -- against the CPU oracle's container at
small sizes, and a C++ program (tests/cpp/grid_test.cpp) that builds grids on one and on several workers."""
import os
import subprocess

import numpy as np
import pytest

import cpvs_b200
from cpvs_b200 import grid as cgrid, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_container(oracle, kind, tile, length):
    cont = oracle.Container(length)
    keep = []
    for y in range(length):
        for x in range(length):
            mm = oracle.MinMax(synth.depth_map(kind, tile, (x, y), length))
            for z in range(length):
                sh = oracle.Shadow(mm, z, length)
                keep.append((mm, sh))
                cont.set(sh, x, y, z)
    cont.finalize()
    return cont, keep


@pytest.mark.parametrize("kind,tile,length,workers", [("terrain_dev", 128, 4, 1), ("terrain_dev", 128, 4, 3), ("city", 256, 4, 2), ("plane", 64, 8, 2),
                                                     ("city", 64, 16, 3)])
def test_grid_build_equals_oracle_container(gpu_ctx, oracle, kind, tile, length, workers):
    """combineDAGs + createTopLevelGrid over cells built by several workers (contexts of GPU 0 when the box has one GPU):
    the container's words and grid are the oracle's, whoever built which tile."""
    import torch
    devices = [d % torch.cuda.device_count() for d in range(workers)]
    g = cgrid.Grid.build(devices, length, tile, kind)
    st = g.stats()
    assert st["cells"] == length ** 3 and sum(st["tiles"]) == length * length
    want, keep = _oracle_container(oracle, kind, tile, length)
    wdag, wgrid = want.dag_and_grid()
    pts = synth.lookups(100000, seed=5)
    for i in range(len(devices)):
        dag, grid = g.container(i).dag_and_grid()
        assert np.array_equal(dag, wdag) and np.array_equal(grid, wgrid), (kind, i)
    assert np.array_equal(g.lookup_ndc(pts), want.lookup_ndc(pts))
    assert st["dag_words"] == wdag.size
    g.close()


def test_grid_fetch_callback_and_assemble(gpu_ctx, oracle):
    """Depth tiles handed over by a host callback (the reference reads them back from GL, src/ShadowMap.cpp:23-30), and
    cpvs_container_assemble on cells collected from two workers by hand (the one-process-per-GPU path)."""
    tile, length = 128, 2
    maps = {(x, y): synth.depth_map("terrain", tile, (x, y), length) for x in range(length) for y in range(length)}

    def fetch(x, y, out):
        out[...] = maps[x, y]

    g = cgrid.Grid.build([0, 0], length, tile, fetch=fetch)
    want, keep = _oracle_container(oracle, "terrain", tile, length)
    wdag, wgrid = want.dag_and_grid()
    dag, grid = g.container(0).dag_and_grid()
    assert np.array_equal(dag, wdag) and np.array_equal(grid, wgrid)
    g.close()

    ctxs = [cpvs_b200.Context(0), cpvs_b200.Context(0)]
    workers = [cgrid.GridWorker(c, length, tile, fetch=fetch) for c in ctxs]
    tiles = [(x, y) for y in range(length) for x in range(length)]
    costs = [workers[i % 2].estimate([t])[0] for i, t in enumerate(tiles)]
    owners = cgrid.assign(costs, 2, [i % 2 for i in range(len(tiles))])
    for i, t in enumerate(tiles):
        if owners[i] != i % 2:
            workers[i % 2].release([t])
    for w, worker in enumerate(workers):
        worker.build([t for i, t in enumerate(tiles) if owners[i] == w])
    cells = {c.index: c for worker in workers for c in worker.cells()}
    assert sorted(cells) == list(range(length ** 3))
    parts = [(cells[i].words, cells[i].root_mask, cells[i].device, cells[i].words_device) for i in range(length ** 3)]
    cont = cgrid.assemble(ctxs[0], length, cells[0].num_levels, True, parts)
    dag, grid = cont.dag_and_grid()
    assert np.array_equal(dag, wdag) and np.array_equal(grid, wgrid)
    for worker in workers:
        worker.close()


def _cells_by_index(worker):
    cells = worker.cells()
    offsets = worker.copy_cells(None)
    total = (offsets[-1] + int(cells[-1].words)) if cells else 0
    words = np.zeros(max(1, total), dtype=np.uint32)
    worker.copy_cells(words)
    return {int(c.index): words[o:o + int(c.words)].copy() for c, o in zip(cells, offsets)}


@pytest.mark.parametrize("kind,length,tile", [("city", 4, 256), ("terrain_dev", 2, 512)])
def test_worker_pulls_tiles_from_a_queue(kind, length, tile):
    """cpvs_grid_worker_build_from: two workers of one GPU pull their tiles from one shared queue (each asks one tile ahead
    of its build); the cells equal those of a worker that was handed the whole list."""
    ctxs = [cpvs_b200.Context(0), cpvs_b200.Context(0)]
    tiles = [(x, y) for y in range(length) for x in range(length)]
    whole = cgrid.GridWorker(ctxs[0], length, tile, kind)
    whole.build(tiles)
    want = _cells_by_index(whole)
    assert sorted(want) == list(range(length ** 3))

    queue = list(tiles)
    taken = [[], []]

    def puller(i):
        def next_tile():
            if not queue:
                return None
            taken[i].append(queue.pop(0))
            return taken[i][-1]
        return next_tile

    workers = [cgrid.GridWorker(c, length, tile, kind) for c in ctxs]
    # alternate: a few tiles through worker 0, the rest through worker 1, then an empty pull
    limited = iter(range(3))
    workers[0].build_from(lambda: puller(0)() if next(limited, None) is not None else None)
    workers[1].build_from(puller(1))
    workers[0].build_from(puller(0))
    assert len(taken[0]) == 3 and len(taken[0]) + len(taken[1]) == len(tiles)
    got = {}
    for worker in workers:
        got.update(_cells_by_index(worker))
    assert sorted(got) == sorted(want)
    for i in want:
        assert np.array_equal(got[i], want[i]), i
    st = ctxs[0].stats()
    assert st["overflow_rebuilds"] == 0 and st["reemissions"] == 0  # slices are emitted into staging buffers that cannot overflow

    def failing():
        raise RuntimeError("queue broke")

    # give the recycled memory back in the middle: the next build allocates again, same words
    ctxs[0].trim()
    again = cgrid.GridWorker(ctxs[0], length, tile, kind)
    again.build(tiles[:2] + tiles[:1])  # (a tile named twice is built once)
    for i, words in _cells_by_index(again).items():
        assert np.array_equal(words, want[i]), i
    again.close()

    with pytest.raises(RuntimeError, match="queue broke"):
        workers[1].build_from(failing)
    for worker in workers + [whole]:
        worker.close()


@pytest.mark.parametrize("lanes", ["1", "3", "8"])
def test_worker_lanes(lanes, monkeypatch):
    """CPVS_GRID_LANES: however many contexts a worker spreads a tile's slices over, the cells are the same."""
    length, tile, kind = 4, 256, "city"
    tiles = [(x, y) for y in range(length) for x in range(length)][:6]
    ctx = cpvs_b200.Context(0)
    ref = cgrid.GridWorker(ctx, length, tile, kind)  # the default number of lanes
    ref.build(tiles)
    want = _cells_by_index(ref)
    monkeypatch.setenv("CPVS_GRID_LANES", lanes)
    ctx2 = cpvs_b200.Context(0)
    w = cgrid.GridWorker(ctx2, length, tile, kind)
    w.build(tiles)
    got = _cells_by_index(w)
    assert sorted(got) == sorted(want)
    for i in want:
        assert np.array_equal(got[i], want[i]), (lanes, i)
    w.close()
    ref.close()


def test_cpp_caller_builds_grids_on_several_workers(tmp_path):
    """tests/cpp/grid_test.cpp: a plain C++ program against include/cpvs_b200.h + libcpvs_b200.so."""
    import torch
    from cpvs_b200 import build
    lib = build.build()
    exe = str(tmp_path / "grid_test")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "grid_test.cpp"),
                           "-o", exe, "-L", os.path.dirname(lib), "-lcpvs_b200", "-Wl,-rpath," + os.path.dirname(lib)])
    count = torch.cuda.device_count()
    devices = [str(d) for d in range(min(count, 4))] if count > 1 else ["0", "0", "0"]
    out = subprocess.run([exe, "256", "4"] + devices, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "grid_test ok" in out.stdout, out.stdout + out.stderr
