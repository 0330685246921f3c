"""Small frames through every production kernel, for compute-sanitizer (tools/sanitize.sh): host call, frame-level call
(fused exchange + K4), tile-scheduled and unscheduled scenes, brute force and hierarchy, glass and opaque instantiations."""
import sys

import numpy as np
import torch

import rusty_marcher_b200 as rm
from rusty_marcher_b200 import tiled, workloads

rm.init(0)
dev = torch.device("cuda:0")
CASES = [("demo", {}, 256, 160, 3, False), ("cornell_box", {}, 320, 192, 3, False), ("dodecahedron", {}, 256, 128, 3, True),
         ("stress", dict(n_spheres=64, grid=7), 160, 128, 6, True), ("stress", dict(n_spheres=64, grid=7), 160, 128, 6, False),
         ("ngons", dict(n_polygons=48, n_spheres=8), 256, 96, 4, False)]
for name, kw, w, h, depth, accel in CASES:
    scene = workloads.scene(name, **kw)
    r = rm.create_renderer(1.5, h, w)
    r.max_depth, r.accel = depth, accel
    fb = rm.create_frame_buffer(w, h, dtype=np.float32)
    ids = np.full((h, w), -1, dtype=np.int32)
    rgb8 = np.zeros((h, w, 3), dtype=np.uint8)
    r.render(fb, scene, prim_id=ids, rgb8=rgb8)
    tr = tiled.TiledRenderer(tiled.CudaBackend(scene, r, w, h, dev), w, h, dev)
    try:
        for cam in ((0., 0., 0.), (2., -1., 1.)):
            tr.set_camera(cam)
            f = tr.render()
            torch.cuda.synchronize()
            tr.peer.status()
        tr.set_camera((0., 0., 0.))
        f = tr.render().cpu().numpy()
        assert np.array_equal(f, rgb8), name
        assert np.array_equal(tr.rgb.cpu().numpy(), fb.buffer), name
    finally:
        tr.close()
    print("ok", name, kw, "accel" if accel else "brute", int((ids >= 0).sum()), "hits")
    sys.stdout.flush()
print("sanitize target done")
