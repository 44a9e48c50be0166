"""BASELINE configs[3], second half: the same surface-hugging G-buffer lookups with leafmasks on vs off.
Leafmask-less octrees descend three more levels of 9-word nodes, so they are built at 4096^2 here.
Appends what it prints to gpurun_out/leafmask_compare.txt (copy it under profiles/ for the record)."""
import os
import sys
import time
sys.path.insert(0, ".")
import numpy as np
import torch
import cpvs_b200
from cpvs_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = cpvs_b200.Context(0)
depth_np = synth.depth_map("terrain", n)
d = torch.from_numpy(depth_np).cuda()
gw, gh = 3840, 2160
u = (np.arange(gw, dtype=np.float32) + np.float32(0.5)) / np.float32(gw)
v = (np.arange(gh, dtype=np.float32) + np.float32(0.5)) / np.float32(gh)
tex = depth_np[np.minimum((v * n).astype(np.int64), n - 1)[:, None], np.minimum((u * n).astype(np.int64), n - 1)[None, :]]
eps = np.where((np.add.outer(np.arange(gh), np.arange(gw)) & 1) == 0, np.float32(1.5), np.float32(-1.5)) / np.float32(n)
pos_np = np.empty((gh, gw, 4), np.float32)
pos_np[..., 0] = (u * 2 - 1)[None, :]
pos_np[..., 1] = (v * 2 - 1)[:, None]
pos_np[..., 2] = (tex + eps) * 2 - 1
pos_np[..., 3] = 1
frames = [torch.from_numpy(pos_np).cuda() for _ in range(4)]
vis = torch.zeros((gh, gw), dtype=torch.uint8, device="cuda")
ident = np.eye(4, dtype=np.float32)
results = {}
os.makedirs("gpurun_out", exist_ok=True)
record = open(os.path.join("gpurun_out", "leafmask_compare.txt"), "a")


def say(text):
    print(text)
    record.write(text + "\n")
    record.flush()


for leaf in (True, False):
    mm = cpvs_b200.MinMaxHierarchy(d, ctx, n=n)
    sh = cpvs_b200.CompressedShadow.create(mm, leafmasks=leaf)
    cont = cpvs_b200.CompressedShadowContainer(sh, ctx)
    cont.copyToGPU()
    for i in range(4):
        cont.evaluate(frames[i], ident, vis)
    ctx.synchronize()
    t = time.perf_counter()
    for i in range(20):
        cont.evaluate(frames[i % 4], ident, vis)
    ctx.synchronize()
    dt = (time.perf_counter() - t) / 20
    results[leaf] = vis.cpu().numpy().copy()
    svo, dagn, _ = sh.level_counts()
    say("n=%d leafmasks=%s: build %.2f ms, %d words (%.1f MB), svo nodes %d, lookups %.3f ms = %.1f G/s"
          % (n, leaf, sh.info.build_ms, sh.info.words, sh.info.words * 4 / 1e6, int(svo.sum()), dt * 1e3, gw * gh / dt / 1e9))
say("identical visibility with and without leafmasks: %s" % bool(np.array_equal(results[True], results[False])))
