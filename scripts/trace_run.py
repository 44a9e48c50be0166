"""Dev helper: host-side trace (CPVS_TRACE=1) and per-phase device times of a few 16K^2 builds."""
import sys
sys.path.insert(0, ".")
import torch
import cpvs_b200
from cpvs_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
mode = sys.argv[2] if len(sys.argv) > 2 else "own"
kind = sys.argv[3] if len(sys.argv) > 3 else "terrain"
if mode == "own":
    ctx = cpvs_b200.Context(0)
elif mode == "legacy":
    ctx = cpvs_b200.Context(0, stream=torch.cuda.current_stream().cuda_stream)
else:
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    ctx = cpvs_b200.Context(0, stream=s.cuda_stream)
print("mode", mode, "stream", torch.cuda.current_stream().cuda_stream, file=sys.stderr)
d = torch.from_numpy(synth.depth_map(kind, n)).cuda()
torch.cuda.synchronize()
for i in range(4):
    print("--- step", i, file=sys.stderr)
    mm = cpvs_b200.MinMaxHierarchy(d, ctx, n=n)
    sh = cpvs_b200.CompressedShadow.create(mm)
    print({k: round(v, 3) for k, v in sh.phase_ms().items()}, round(sh.info.build_ms, 3), mm.timing(), file=sys.stderr)
    sh.close()
    mm.close()
