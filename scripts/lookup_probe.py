"""Dev helper: surface-coherent G-buffer lookups through evaluate() on a 16K^2 terrain DAG."""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
import cpvs_b200
from cpvs_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ctx = cpvs_b200.Context(0)
depth_np = synth.depth_map("terrain", n)
d = torch.from_numpy(depth_np).cuda()
mm = cpvs_b200.MinMaxHierarchy(d, ctx, n=n)
sh = cpvs_b200.CompressedShadow.create(mm)
cont = cpvs_b200.CompressedShadowContainer(sh, ctx)
cont.copyToGPU()
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
bufs = [torch.from_numpy(pos_np).cuda() for _ in range(4)]
vis = torch.zeros((gh, gw), dtype=torch.uint8, device="cuda")
ident = np.eye(4, dtype=np.float32)
for b in bufs:
    cont.evaluate(b, ident, vis)
ctx.synchronize()
import time
t = time.perf_counter()
for i in range(20):
    cont.evaluate(bufs[i % 4], ident, vis)
ctx.synchronize()
dt = (time.perf_counter() - t) / 20
print("evaluate: %.3f ms per 8.29M lookups = %.1f Glookups/s, lit %.3f" % (dt * 1e3, gw * gh / dt / 1e9, float((vis != 0).float().mean())))
