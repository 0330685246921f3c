"""K0/K1 times of one rank's share of a frame for several band strides (single process, single GPU):
python tools/run_bands.py [workload] [reps]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import rusty_marcher_b200 as rm  # noqa: E402
from bench import workload_of  # noqa: E402
from rusty_marcher_b200 import _abi, tiled, workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cornell_4k"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
scene_name, w, h, depth, kw, accel = workload_of(name)
rm.init(0)
L = _abi.load()
dev = torch.device("cuda:0")
scene = workloads.scene(scene_name, **kw)
r = rm.create_renderer(1.5, h, w)
r.max_depth = int(os.environ.get("BANDS_DEPTH", depth))
r.accel = accel
be = tiled.CudaBackend(scene, r, w, h, dev)
rgb = torch.zeros((h, w, 3), dtype=torch.float32, device=dev)
dmax = torch.zeros(1, dtype=torch.float32, device=dev)
rgb8 = torch.zeros((h, w, 3), dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
_abi.check(L.rm_set_profiling(1))
P = h // 32
strides = [int(x) for x in os.environ.get("BANDS_STRIDES", "1,2,4,8").split(",")]
for stride in strides:
    for first in range(stride if os.environ.get("BANDS_ALL") else min(stride, 2)):
        rows = (first, P, stride)
        for flushed in ((True,) if os.environ.get("BANDS_ALL") else (False, True)):
            for i in range(reps):
                if flushed:
                    flush.fill_(0)
                be.render_rows(rows, rgb, dmax, rgb8=rgb8)
            torch.cuda.synchronize()
            k0, k1 = [], []
            a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
            for back in range(reps - 2):
                _abi.check(L.rm_kernel_times(back, C.byref(a), C.byref(b), C.byref(c)))
                k0.append(a.value)
                k1.append(b.value)
            print("%s stride %d first %d %s: K0 %.4f ms  K1 %.4f ms (min %.4f)" % (name, stride, first, "L2 flushed" if flushed else "warm", sum(k0) / len(k0), sum(k1) / len(k1), min(k1)))
