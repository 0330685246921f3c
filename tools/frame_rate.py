"""Back-to-back frames of the re-render loop (no L2 flush, no synchronisation between frames): frames per second the HOST can
issue and the device can retire, as one CUDA graph launch per frame (RM_B200_GRAPH=1) and as the two plain launches.
    python tools/frame_rate.py [workload] [frames]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import rusty_marcher_b200 as rm  # noqa: E402
from bench import workload_of  # noqa: E402
from rusty_marcher_b200 import _abi, tiled, workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cornell_4k"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
scene_name, w, h, depth, kw, accel = workload_of(name)
rm.init(0)
L = _abi.load()
scene = workloads.scene(scene_name, **kw)
r = rm.create_renderer(1.5, h, w)
r.max_depth, r.accel = depth, accel
be = tiled.CudaBackend(scene, r, w, h, dev)
tr = tiled.TiledRenderer(be, w, h, dev, exchange="peer")
for _ in range(20):
    tr.render()
torch.cuda.synchronize()
for block in range(3):
    for arm, label in (("1", "one graph launch per frame"), ("0", "two launches per frame   ")):
        os.environ["RM_B200_GRAPH"] = arm
        for _ in range(20):
            tr.render()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(frames):
            tr.render()
        t_issue = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_done = time.perf_counter() - t0
        print("%s: host issues a frame every %.1f us, %d frames retired in %.1f ms = %.1f us per frame (%.0f frames/s)"
              % (label, t_issue / frames * 1e6, frames, t_done * 1e3, t_done / frames * 1e6, frames / t_done), flush=True)
tr.close()
