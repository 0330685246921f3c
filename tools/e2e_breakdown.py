"""Where the end-to-end time of Renderer.render goes (cornell 4K): scene upload, the call with the scene resident, retained,
the bare C call; with RM_B200_DELIVERY_TRACE=1 the library prints its own host-side phases."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import rusty_marcher_b200 as rm  # noqa: E402
from rusty_marcher_b200 import _abi, workloads  # noqa: E402

w, h = 3840, 2160
rm.init(0)
L = _abi.load()
scene = workloads.scene("cornell_box")
r = rm.create_renderer(1.5, h, w)
fb = rm.create_frame_buffer(32, 32)
fb.width, fb.height = w, h
pin = L.rm_host_alloc(h * w * 12)
fb.buffer = np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_float)), shape=(h, w, 3))
out = open(os.devnull, "w")
real = sys.stdout


def timed(fn, n=60, warm=5):
    ts = []
    for i in range(n + warm):
        t0 = time.perf_counter()
        fn()
        if i >= warm:
            ts.append(time.perf_counter() - t0)
    ts.sort()
    return 1e3 * ts[len(ts) // 2], 1e3 * ts[0]


def upload():
    scene.release()
    scene.device_handle()


def render():
    sys.stdout = out
    r.render(fb, scene)
    sys.stdout = real


def both():
    scene.release()
    render()


import torch  # noqa: E402


def upload_sync():
    scene.release()
    scene.device_handle()
    torch.cuda.synchronize()


def mark(msg):
    sys.stderr.write("== %s\n" % msg)
    sys.stderr.flush()


p = r.params(fb, scene)
print("upload + device synchronise           median %.3f  min %.3f ms" % timed(upload_sync))
st = _abi.RmStats()
print("upload (release + device_handle)      median %.3f  min %.3f ms" % timed(upload))
print("render, scene resident, fresh         median %.3f  min %.3f ms" % timed(render))
print("upload + render (bench e2e)           median %.3f  min %.3f ms" % timed(both))
print("bare rm_render (ctypes), fresh        median %.3f  min %.3f ms" % timed(lambda: L.rm_render(scene.device_handle(), C.byref(p), fb.buffer.ctypes.data, None, None, C.byref(st))))
r.retained = True
mark("retained resident")
print("render, scene resident, retained      median %.3f  min %.3f ms" % timed(render))
mark("retained with upload")
print("upload + render, retained             median %.3f  min %.3f ms" % timed(both))
fb64 = rm.create_frame_buffer(w, h, dtype=np.float64)
r.retained = False


def render64():
    sys.stdout = out
    r.render(fb64, scene)
    sys.stdout = real


mark("f64")
print("f64 rows, fresh                       median %.3f  min %.3f ms" % timed(render64, 30))
r.retained = True
print("f64 rows, retained                    median %.3f  min %.3f ms" % timed(render64, 30))
print("library ms_total (events) of the last call: %.3f, d2h %d bytes" % (r.last_stats.ms_total, r.last_stats.d2h_bytes))
