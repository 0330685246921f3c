"""Runs the render kernel a few times on one workload (for ncu captures): python tools/run_k1.py [workload] [reps] [--f64] [--no-cull]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import rusty_marcher_b200 as rm  # noqa: E402
from bench import workload_of  # noqa: E402
from rusty_marcher_b200 import tiled, workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "cornell_4k"
reps = int(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("-") else 5
scene_name, w, h, depth, kw, accel = workload_of(name)
rm.init(0)
dev = torch.device("cuda:0")
scene = workloads.scene(scene_name, **kw)
r = rm.create_renderer(1.5, h, w)
r.max_depth = depth
r.accel = accel
r.cull_backfacing = "--no-cull" not in sys.argv
if "--f64" in sys.argv:
    r.precision = rm.RM_FP64
be = tiled.CudaBackend(scene, r, w, h, dev)
dt = torch.float64 if "--f64" in sys.argv else torch.float32
rgb = torch.zeros((h, w, 3), dtype=dt, device=dev)
dmax = torch.zeros(1, dtype=dt, device=dev)
rgb8 = torch.zeros((h, w, 3), dtype=torch.uint8, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
for i in range(reps):
    ev[i].record()
    be.render_rows((0, h // 32), rgb, dmax)
ev[reps].record()
be.tonemap_rows((0, h // 32), rgb, dmax, rgb8)
torch.cuda.synchronize()
print(name, "K1 ms:", ["%.3f" % ev[i].elapsed_time(ev[i + 1]) for i in range(reps)], "max", float(dmax.item()))
