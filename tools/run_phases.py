"""Phase breakdown of the frame kernel per rank from its in-kernel %globaltimer stamps (rm_peer_stamps).
torchrun --nproc-per-node N tools/run_phases.py [workload] [frames]   (or plain python for one GPU)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import rusty_marcher_b200 as rm  # noqa: E402
from bench import workload_of  # noqa: E402
from rusty_marcher_b200 import _abi, tiled, workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cornell_4k"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 8
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
scene_name, w, h, depth, kw, accel = workload_of(name)
rm.init(lr)
L = _abi.load()
scene = workloads.scene(scene_name, **kw)
r = rm.create_renderer(1.5, h, w)
r.max_depth = depth
r.accel = accel
be = tiled.CudaBackend(scene, r, w, h, dev)
tr = tiled.TiledRenderer(be, w, h, dev, exchange="peer")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
_abi.check(L.rm_set_profiling(1))
rows = []
for i in range(frames):
    flush.fill_(0)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    tr.render()
    torch.cuda.synchronize()
    st = (C.c_uint64 * 7)()
    _abi.check(L.rm_peer_stamps(C.byref(tr.peer.x), st))
    k0, k1, k4 = C.c_double(0), C.c_double(0), C.c_double(0)
    _abi.check(L.rm_kernel_times(0, C.byref(k0), C.byref(k1), C.byref(k4)))
    t = [int(x) for x in st]
    rows.append(((t[6] - t[5]) / 1e3 if t[5] else 0.0, (t[0] - t[6]) / 1e3 if t[5] else 0.0, k1.value * 1e3, (t[1] - t[0]) / 1e3, (t[2] - t[1]) / 1e3, (t[3] - t[2]) / 1e3, (t[4] - t[3]) / 1e3 if rank == 0 and world > 1 else 0.0))
for rk in range(world):
    if world > 1:
        dist.barrier()
    if rk == rank:
        for i, x in enumerate(rows[2:]):
            print("rank %d frame %d: K0 %.1f us, gap %.1f us to K1's work; K0+K1 events %.1f us; in K1: render %.1f + wait-max %.1f + tone %.1f + wait-done %.1f" % ((rank, i + 2) + x), flush=True)
tr.close()
if world > 1:
    dist.destroy_process_group()
