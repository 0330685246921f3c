"""GPU diagnostic: primary ids of the hierarchy path against brute force, saved for inspection on the host."""
import contextlib, io, sys, time
import numpy as np
sys.path.insert(0, ".")
import rusty_marcher_b200 as rm
from rusty_marcher_b200 import workloads

w, h, depth = (int(x) for x in sys.argv[1:4])
rm.init(0)
scene = workloads.build_scene(workloads.describe("stress", n_spheres=1024, grid=64))
scene.offset_camera((7.5, -3.25, 20.0))

def render(accel):
    r = rm.create_renderer(1.5, h, w)
    r.max_depth, r.accel = depth, accel
    fb = rm.create_frame_buffer(w, h, dtype=np.float32)
    ids = np.full((h, w), -1, dtype=np.int32)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        r.render(fb, scene, prim_id=ids)
    print("accel", accel, "%.3f s kernel %.3f ms hits %d" % (time.perf_counter() - t0, r.last_stats.ms_render, int((ids >= 0).sum())), flush=True)
    return fb.buffer.copy(), ids

a = render(False)
b = render(True)
c = render(True)
print("accel deterministic:", bool(np.array_equal(b[1], c[1]) and np.array_equal(b[0], c[0])), "ids differ from brute at", int((a[1] != b[1]).sum()), "rgb differ at", int((a[0] != b[0]).any(axis=2).sum()), flush=True)
ok = bool(np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0]) and np.array_equal(b[1], c[1]) and np.array_equal(b[0], c[0]))
tag = sys.argv[4] if len(sys.argv) > 4 else ""
np.savez_compressed("gpurun_out/diag_ids%s_%dx%d_d%d.npz" % (tag, w, h, depth), brute=a[1], accel=b[1], accel2=c[1], brute_rgb=a[0], accel_rgb=b[0])
sys.exit(0 if ok else 3)
