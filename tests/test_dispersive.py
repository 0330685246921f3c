"""Extension mode (BASELINE.json configs[3]): per-channel refractive indices as three scalar-index
passes whose R/G/B channels are recombined.  The reference has no such behaviour (one scalar index per material,
shapes.rs:21-32), so there is nothing of the reference to be bit-faithful to: the oracle (oracle.render_dispersive) and
the CUDA path (rm_render_dispersive) are extended identically, and the parity criteria are the usual ones between them.
CPU part: the kernel code in the host emulation against the oracle.  GPU part: tests/test_zz_dispersive_gpu.py."""
import numpy as np
import pytest

from oracle import oracle as O
from rusty_marcher_b200 import workloads
from tests import parity
from tests.emu import emu
from tests.oracle_scenes import build_oracle_scene

INDICES = (1.50, 1.52, 1.54)


def glass_dodecahedron():
    """configs[3]: the dodecahedron mesh as glass (reflection 0.2), on both sides."""
    desc = workloads.describe("dodecahedron")
    scene = workloads.build_scene(desc)
    osc = build_oracle_scene(desc)
    for k, shape in enumerate(scene.shapes):
        shape.make_glass(reflection=0.2, refractive_index=INDICES[0], diffusion=1.)
        assert osc.make_glass(k, reflection=0.2, refractive_index=INDICES[0], diffusion=1.) == shape.triangles.shape[0]
    return scene, osc


def demo():
    return workloads.scene("demo"), O.Scene.create_default()


def emu_dispersive(scene, w, h, precision, depth):
    out = None
    for c, n in enumerate(INDICES):
        r = emu.render(scene, w, h, precision, max_depth=depth, glass_index=n)
        if out is None:
            out = {"rgb": np.zeros_like(r["rgb"]), "prim_id": r["prim_id"]}
        out["rgb"][..., c] = r["rgb"][..., c]
    return out


@pytest.mark.parametrize("make,w,h,depth", [(demo, 256, 160, 3), (demo, 128, 96, 5), (glass_dodecahedron, 320, 256, 3)])
def test_dispersive_kernel_code_against_the_oracle(make, w, h, depth):
    scene, osc = make()
    ref = O.render_dispersive(osc, w, h, INDICES, max_depth=depth)
    plain = O.render(make()[1], w, h, max_depth=depth)
    assert np.array_equal(ref["prim_id"], plain["prim_id"])                  # primary rays do not see the index
    assert (ref["rgb"] != plain["rgb"]).any(axis=2).sum() > 50                # ... but the glass does
    got64 = emu_dispersive(scene, w, h, "f64", depth)
    assert np.array_equal(got64["prim_id"], ref["prim_id"]) and np.array_equal(got64["rgb"], ref["rgb"])
    got = emu_dispersive(scene, w, h, "fast", depth)
    m = np.float32(got["rgb"].max())
    rgb8 = (np.float32(255) * np.clip(got["rgb"] * (np.float32(1) / m), 0, 1)).astype(np.uint8)
    parity.check_fp32(got, ref, (h // 32) * 32, rgb8)


def test_equal_indices_reduce_to_the_plain_frame():
    scene, osc = demo()
    a = O.render(osc, 128, 96)
    b = O.render_dispersive(O.Scene.create_default(), 128, 96, (1.5, 1.5, 1.5))   # the demo's glass has index 1.5 (scene.rs)
    assert np.array_equal(a["rgb"], b["rgb"]) and np.array_equal(a["prim_id"], b["prim_id"])
    flat = scene.flatten()
    assert flat.set_glass_index(1.5) == 2                                     # the two glass shapes of the demo scene
