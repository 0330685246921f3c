"""Generates the committed fixtures from the reference tree (run in the build container only):

  rusty_marcher_b200/scenes/{cornell_box,dodecahedron}.npz  parsed OBJ models (f32 vertices per model)
  tests/golden/out_ppm.json          sha256 + size + sampled bytes of the reference's engine/out.ppm
  tests/golden/oracle_demo_800x600.json   oracle known answers (counters, pixels) for the demo scene

    python tests/golden/make_fixtures.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import oracle as O  # noqa: E402
from rusty_marcher_b200 import obj  # noqa: E402


def main():
    out_dir = os.path.join(ROOT, "rusty_marcher_b200", "scenes")
    os.makedirs(out_dir, exist_ok=True)
    for name in ("cornell_box", "dodecahedron"):
        models = obj.load(os.path.join(REF, "test_data", name + ".obj"))
        arrays = {"names": np.array([m.name for m in models])}
        for i, m in enumerate(models):
            v = m.triangles[:, 0:9].reshape(-1, 3, 3)
            v32 = v.astype(np.float32)
            assert np.array_equal(v32.astype(np.float64), v), "OBJ vertices must be exact f32 values"
            arrays["model_%d" % i] = v32
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrays)
        print(name, [(m.name, m.triangles.shape[0]) for m in models])

    ppm = open(os.path.join(REF, "engine", "out.ppm"), "rb").read()
    idx = list(range(0, len(ppm), 9973))
    json.dump({"sha256": hashlib.sha256(ppm).hexdigest(), "size": len(ppm), "header": ppm[:15].decode("latin1"),
               "sample_stride": 9973, "sample": [ppm[i] for i in idx]},
              open(os.path.join(ROOT, "tests", "golden", "out_ppm.json"), "w"))

    sc = O.Scene.create_default()
    r = O.render(sc, 800, 600)
    px = {"%d,%d" % (x, y): [float(v) for v in r["rgb"][y, x]] for x, y in ((400, 300), (100, 100), (700, 300), (400, 500), (799, 575))}
    json.dump({"counters": r["counters"], "pixels": px, "mean": float(r["rgb"].mean()), "max": float(r["rgb"].max())},
              open(os.path.join(ROOT, "tests", "golden", "oracle_demo_800x600.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
