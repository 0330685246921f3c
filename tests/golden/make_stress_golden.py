"""The oracle's frame of BASELINE.json configs[4]'s FULL scene (4096 spheres + the 224 x 224 quad grid: 104,448 primitives,
depth cap 6) at 384 x 224, committed as a fixture: the reference's brute-force traversal needs about a minute of CPU per
frame at this size (10^5 primitive tests per ray segment), too long for a test run on the GPU box.

    python tests/golden/make_stress_golden.py        (build container; writes tests/golden/stress_full_384x224.npz)

rgb is stored as float32 (the f64 oracle value rounded once: 6e-8 relative, three orders below the 1e-4 the parity test
allows), prim_id as int32, the near-tie mask as uint8."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from rusty_marcher_b200 import workloads  # noqa: E402
from tests.oracle_scenes import build_oracle_scene  # noqa: E402

W, H, DEPTH = 384, 224, 6

if __name__ == "__main__":
    desc = workloads.describe("stress")
    t0 = time.time()
    ref = O.render(build_oracle_scene(desc), W, H, max_depth=DEPTH)
    print("oracle: %.1f s, counters %s" % (time.time() - t0, ref["counters"]))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "stress_full_%dx%d.npz" % (W, H)),
                        rgb=ref["rgb"].astype(np.float32), rgb_max=np.float64(ref["rgb"].max()), prim_id=ref["prim_id"], fragile=ref["fragile"],
                        counters=np.array([ref["counters"][k] for k in O.COUNTER_FIELDS], dtype=np.uint64), depth=DEPTH)
