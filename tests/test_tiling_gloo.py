"""Host-side logic of the multi-GPU tile grid, on CPU: two gloo ranks each build their share of the
cells (the CPU oracle stands in for the CUDA builder -- test infrastructure), gather on the host, and
the assembled container must equal the single-process container word for word."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, length, n, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cpvs_b200 import synth, tiling
    from oracle import pyoracle as O

    def build_cells(x, y):
        mm = O.MinMax(synth.depth_map("terrain", n, (x, y), length, threads=1))
        return [O.Shadow(mm, z, length).dag() for z in range(length)]

    def gather(obj):
        parts = [None] * world
        dist.all_gather_object(parts, obj)
        return parts

    dags, grid, total = tiling.build_distributed(length, rank, world, build_cells, gather)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), dag=np.concatenate(dags), grid=grid, total=total)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_two_rank_tile_grid_matches_single_process(tmp_path, oracle, world):
    from cpvs_b200 import synth, tiling
    length, n = 2, 32
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, length, n, str(tmp_path)), nprocs=world, join=True)
    cont = oracle.Container(length)
    keep = []
    for (x, y) in tiling.xy_tiles(length):
        mm = oracle.MinMax(synth.depth_map("terrain", n, (x, y), length))
        for z in range(length):
            sh = oracle.Shadow(mm, z, length)
            keep.append(sh)
            cont.set(sh, x, y, z)
    cont.finalize()
    dag, grid = cont.dag_and_grid()
    for rank in range(world):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        assert np.array_equal(got["grid"], grid) and np.array_equal(got["dag"], dag) and int(got["total"]) == dag.size


def test_tile_ownership_partitions_the_grid():
    from cpvs_b200 import tiling
    for length in (1, 2, 4, 16):
        for world in (1, 2, 3, 8):
            owned = [t for r in range(world) for t in tiling.tiles_of_rank(length, r, world)]
            assert sorted(owned) == sorted(tiling.xy_tiles(length))
            sizes = [len(tiling.tiles_of_rank(length, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_top_level_grid_sentinels():
    from cpvs_b200 import tiling
    grid, total = tiling.top_level_grid([(1, 0x5555), (10, 0xAAAA), (1, 0), (7, 0x2)] * 2, 2)
    assert list(grid) == [0x0FFFFFFE, 1, 0x0FFFFFFF, 12, 0x0FFFFFFE, 20, 0x0FFFFFFF, 31] and total == 38
