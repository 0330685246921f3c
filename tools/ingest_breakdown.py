"""Where the FIRST frame of a large scene goes (ingest at the speed of the frame) -- configs[4]'s
stress scene, 104k primitives at 7680x4320 through the hierarchy: fingerprint and marshalling in the Python mirror, the
library's upload (content hash, packing, hierarchy build, one H2D copy; RM_B200_PACK_TRACE=1 prints the packer's own
phases), the first frame delivered to host memory, and the steady state.

    python tools/ingest_breakdown.py [workload]          (default stress_8k_bvh)
"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("RM_B200_PACK_TRACE", "1")
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import rusty_marcher_b200 as rm  # noqa: E402
from rusty_marcher_b200 import _abi, workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "stress_8k_bvh"
scene_name, w, h, depth, kw, accel = bench.workload_of(name)
rm.init(0)
L = _abi.load()
out, real = open(os.devnull, "w"), sys.stdout


def ms(t0):
    return (time.perf_counter() - t0) * 1e3


def fresh_scene():
    """a scene the process has never seen (fresh objects; distinct content, so the library's pack cache does not know it)"""
    global seed
    seed += 1
    scene = workloads.scene(scene_name, seed=seed, **kw) if scene_name == "stress" else workloads.scene(scene_name, **kw)
    r = rm.create_renderer(1.5, h, w)
    r.max_depth, r.accel = depth, accel
    fb = rm.create_frame_buffer(32, 32)
    fb.width, fb.height = w, h
    fb.buffer = np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_float)), shape=(h, w, 3))
    torch.cuda.synchronize()
    return scene, r, fb


def first_frame(label):
    """the call a user makes, on a scene never seen: Renderer.render(frame, scene)"""
    scene, r, fb = fresh_scene()
    t0 = time.perf_counter()
    sys.stdout = out
    r.render(fb, scene)
    sys.stdout = real
    print("%s: first frame %.1f ms (Renderer.render on a scene never seen, float frame in host memory)" % (label, ms(t0)))
    return scene, r, fb


def first_frame_parts(label):
    """the same, step by step (the fingerprint is taken once here, as inside the call)"""
    scene, r, fb = fresh_scene()
    t0 = time.perf_counter()
    fp = scene._fp()
    t_fp = ms(t0)
    t0 = time.perf_counter()
    flat = scene.flatten()
    t_flat = ms(t0)
    t0 = time.perf_counter()
    handle = C.c_int64(0)
    _abi.check(L.rm_scene_upload(C.byref(flat.c), C.byref(handle)))
    t_upload = ms(t0)
    scene._flat, scene._flat_fingerprint, scene._handle, scene._fingerprint = flat, fp, handle.value, fp
    t0 = time.perf_counter()
    sys.stdout = out
    r.render(fb, scene)
    sys.stdout = real
    t_render = ms(t0)
    print("%s: fingerprint %.1f + marshal %.1f + rm_scene_upload %.1f (hash, pack, hierarchy, H2D) + render and deliver %.1f (of it a fingerprint again) = %.1f ms"
          % (label, t_fp, t_flat, t_upload, t_render, t_fp + t_flat + t_upload + t_render))
    return scene, r, fb


pin = L.rm_host_alloc(h * w * 12)
seed = 0x5EED
first_frame("cold process")              # includes CUDA module load, pinned staging growth, pool start-up
for _ in range(3):
    first_frame("new scene")
for _ in range(2):
    scene, r, fb = first_frame_parts("new scene, by parts")


def both():
    scene.release()
    sys.stdout = out
    r.render(fb, scene)
    sys.stdout = real


def render():
    sys.stdout = out
    r.render(fb, scene)
    sys.stdout = real


for label, fn in (("steady state: upload of a known scene + frame (bench e2e)", both), ("steady state: frame, scene resident", render)):
    ts = []
    for i in range(8):
        t0 = time.perf_counter()
        fn()
        ts.append(ms(t0))
    ts.sort()
    print("%s: median %.1f ms, min %.1f" % (label, ts[len(ts) // 2], ts[0]))
print("library ms_total (events) of the last call: %.2f, d2h %.1f MB, %d host threads" % (r.last_stats.ms_total, r.last_stats.d2h_bytes / 1e6, os.cpu_count()))
L.rm_host_free(pin)
