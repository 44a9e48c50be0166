"""GPU suite (-m gpu): the tile grid with one process per GPU -- two ranks (on two GPUs when the box has them, else both on
GPU 0), gloo for the host-side gathers, CUDA IPC + peer copies for the replication, no NCCL. Both ranks must end up with the
container a single process builds."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rank(rank, world, port, kind, tile, length, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cpvs_b200
    from cpvs_b200 import gridbuild
    device = rank % torch.cuda.device_count()
    torch.cuda.set_device(device)
    ctx = cpvs_b200.Context(device)
    res = gridbuild.run(ctx, tile, length, kind, rank, world, dist.group.WORLD, lookups=3840 * 32, lookup_iters=2)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.array([res["dag_words"], res["lookups_lit"], res["moved_tiles"], sum(res["tiles_per_rank"]),
                                                                 res["verified"]["container_lookups_vs_depth"]], np.int64))
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind,tile,length", [("terrain_dev", 256, 2), ("city", 128, 8)])
def test_two_ranks_build_one_grid(tmp_path, kind, tile, length):
    import torch.multiprocessing as mp
    import cpvs_b200
    from cpvs_b200 import gridbuild
    port = 29700 + (os.getpid() % 2000)
    mp.spawn(_rank, args=(2, port, kind, tile, length, str(tmp_path)), nprocs=2, join=True)
    ctx = cpvs_b200.Context(0)
    one = gridbuild.run(ctx, tile, length, kind, lookups=3840 * 32, lookup_iters=1)
    ctx.close()
    for rank in range(2):
        words, lit, moved, tiles, verified = np.load(os.path.join(str(tmp_path), "rank%d.npy" % rank))
        assert words == one["dag_words"] and lit == one["lookups_lit"] and tiles == length * length
        assert verified == 3840 * 32  # every lookup was checked against its depth tile by the rank owning the tile
