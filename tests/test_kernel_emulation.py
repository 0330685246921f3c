"""The product's kernel code (csrc/rm_trace.cuh -- the very templates the CUDA kernels instantiate)
compiled for the host by tests/emu and checked against the oracle.  Keeps the kernel LOGIC under
test in the GPU-less container; the real parity tests (through the C ABI, on a B200) are in
test_gpu_parity.py.  The emulation is a test tool, not a fallback: the product cannot reach it."""
import numpy as np
import pytest

from oracle import oracle as O
from rusty_marcher_b200 import workloads
from tests import parity
from tests.emu import emu
from tests.oracle_scenes import build_oracle_scene

CASES = [("demo", 256, 160, 3, {}), ("cornell_box", 160, 128, 3, {}), ("dodecahedron", 160, 128, 3, {}),
         ("stress", 160, 128, 6, dict(n_spheres=64, grid=7))]


@pytest.fixture(scope="module", params=CASES, ids=[c[0] for c in CASES])
def case(request):
    name, w, h, depth, kw = request.param
    desc = workloads.describe(name, **kw)
    ref = O.render(build_oracle_scene(desc), w, h, max_depth=depth)
    return name, w, h, depth, workloads.build_scene(desc), ref


@pytest.mark.parametrize("cull", [False, True])
def test_fp64_kernel_code_is_bit_exact(case, cull):
    name, w, h, depth, scene, ref = case
    got = emu.render(scene, w, h, "f64", max_depth=depth, cull=cull)
    parity.check_exact(got, ref)
    if not cull:
        assert got["counters"] == ref["counters"]
        assert got["max"] == ref["rgb"].max()
    else:
        # culling removes primitives that can never be hit: same control flow, fewer tests
        for k in ("pixels", "closest_segments", "anyhit_segments", "hits", "light_evals", "lit_lights", "glass_hits",
                  "reflections", "refractions", "sphere_tests", "sphere_hits"):
            assert got["counters"][k] == ref["counters"][k], k
        assert got["counters"]["plane_tests"] <= ref["counters"]["plane_tests"]


@pytest.mark.parametrize("path", ["f32", "fast"])   # generic FP32 kernel code / production fast path (rm_fast.cuh)
@pytest.mark.parametrize("cull", [False, True])
def test_fp32_kernel_code_within_tolerance(case, cull, path):
    name, w, h, depth, scene, ref = case
    got = emu.render(scene, w, h, path, max_depth=depth, cull=cull)
    m = np.float32(got["rgb"].max())
    rgb8 = (np.float32(255) * np.clip(got["rgb"] * (np.float32(1) / m), 0, 1)).astype(np.uint8)
    parity.check_fp32(got, ref, (h // 32) * 32, rgb8)


@pytest.mark.parametrize("cull", [False, True])
def test_strip_bound_changes_no_pixel(case, cull):
    """Stage A drops, per 32x4 strip, the triangles whose raster functions cannot be positive anywhere in the strip
    (exact corner bound, rm_fast.cuh tri_may_touch): the frame must be bit-identical to walking every triangle."""
    name, w, h, depth, scene, ref = case
    a = emu.render(scene, w, h, "fast", max_depth=depth, cull=cull, strip_bound=True)
    b = emu.render(scene, w, h, "fast", max_depth=depth, cull=cull, strip_bound=False)
    assert np.array_equal(a["prim_id"], b["prim_id"]) and np.array_equal(a["rgb"], b["rgb"])


def test_strip_bound_off_centre_camera():
    scene = workloads.scene("cornell_box")
    scene.offset_camera((35., -20., 15.))
    a = emu.render(scene, 416, 224, "fast", cull=False, strip_bound=True)
    b = emu.render(scene, 416, 224, "fast", cull=False, strip_bound=False)
    assert (a["prim_id"] >= 0).any()
    assert np.array_equal(a["prim_id"], b["prim_id"]) and np.array_equal(a["rgb"], b["rgb"])


def test_row_tiles_reassemble_to_the_full_frame():
    scene = workloads.scene("demo")
    full = emu.render(scene, 128, 160, "fast")
    parts = [emu.render(scene, 128, 160, "fast", patch_rows=r) for r in ((0, 2), (2, 3), (3, 5))]
    acc = np.zeros_like(full["rgb"])
    for p, (a, b) in zip(parts, ((0, 2), (2, 3), (3, 5))):
        assert np.all(p["rgb"][:a * 32] == 0) and np.all(p["rgb"][b * 32:] == 0)
        acc += p["rgb"]
    assert np.array_equal(acc, full["rgb"])


def test_depth_cap_semantics():
    """renderer.rs:262-264: max_depth 0 returns the background for every pixel, deeper caps add light."""
    scene = workloads.scene("demo")
    d0 = emu.render(scene, 64, 64, "f64", max_depth=0)
    assert np.all(d0["rgb"] == 0.1) and np.all(d0["prim_id"] == -1)
    osc = O.Scene.create_default()
    for depth in (1, 2, 5):
        parity.check_exact(emu.render(scene, 64, 64, "f64", max_depth=depth), O.render(osc, 64, 64, max_depth=depth))


def test_empty_scene_and_camera_offset():
    import rusty_marcher_b200 as rm
    empty = rm.Scene.new()
    r = emu.render(empty, 64, 32, "fast")
    assert np.all(r["rgb"] == 0) and np.all(r["prim_id"] == -1)
    scene = workloads.scene("demo")
    scene.offset_camera((5., 0., -5.))                          # main.rs:124-171 moves by +-5
    osc = O.Scene.create_default()
    osc.offset_camera((5., 0., -5.))
    parity.check_exact(emu.render(scene, 96, 64, "f64"), O.render(osc, 96, 64))


@pytest.mark.parametrize("cull", [False, True])
def test_hierarchy_changes_no_pixel(case, cull):
    """RmParams.accel: queries through the bounding-volume hierarchy (rm_bvh.cuh) test a superset of the primitives a
    ray can hit with the same routines and pick the winner by the same (distance, id) order -- the frame must be
    bit-identical to the brute-force FP32 frame, primary ids included."""
    name, w, h, depth, scene, ref = case
    a = emu.render(scene, w, h, "fast", max_depth=depth, cull=cull)
    b = emu.render(scene, w, h, "fast", max_depth=depth, cull=cull, accel=True)
    assert np.array_equal(a["prim_id"], b["prim_id"]) and np.array_equal(a["rgb"], b["rgb"])


def test_hierarchy_on_a_large_scene_and_a_moved_camera():
    """1024 spheres + 8192 triangles (the bench's stress_4k scene), depth cap 6, camera off the origin."""
    scene = workloads.build_scene(workloads.describe("stress", n_spheres=1024, grid=64))
    scene.offset_camera((7.5, -3.25, 20.0))
    a = emu.render(scene, 320, 192, "fast", max_depth=6)
    b = emu.render(scene, 320, 192, "fast", max_depth=6, accel=True)
    assert (a["prim_id"] >= 0).sum() > 5000
    assert np.array_equal(a["prim_id"], b["prim_id"]) and np.array_equal(a["rgb"], b["rgb"])


def _with_lights(desc, n):
    """The stress scene lit by n lights (the reference's two, main.rs:293-315, plus more of the same kind)."""
    extra = [((-40., 30., 10.), (.5, 1., .5), .6), ((35., -25., -20.), (.6, .6, 1.), .7), ((0., 60., -90.), (1., 1., .4), .5),
             ((-70., -10., -60.), (1., .3, 1.), .4), ((10., 5., -100.), (.9, .9, .9), .3), ((55., 40., -140.), (.3, 1., 1.), .5),
             ((-5., -45., -40.), (1., .7, .2), .6)]
    desc = dict(desc)
    desc["lights"] = (list(desc["lights"]) + extra)[:n]
    return desc


@pytest.mark.parametrize("n_lights", [1, 3, 4, 8, 9])
def test_hierarchy_with_more_lights(n_lights):
    """The hierarchy kernel shades its lights as unrolled pairs (up to 8; a 9-light scene stays on the brute-force
    kernel on the device): odd counts, several pairs."""
    scene = workloads.build_scene(_with_lights(workloads.describe("stress", n_spheres=128, grid=12), n_lights))
    a = emu.render(scene, 224, 128, "fast", max_depth=4)
    b = emu.render(scene, 224, 128, "fast", max_depth=4, accel=True)
    assert (a["prim_id"] >= 0).sum() > 2000
    assert np.array_equal(a["prim_id"], b["prim_id"]) and np.array_equal(a["rgb"], b["rgb"])


@pytest.mark.parametrize("name,kw", [("demo", {}), ("cornell_box", {}), ("stress", dict(n_spheres=512, grid=32)),
                                     ("stress", dict(n_spheres=1024, grid=64))])     # last: >= 8192 primitives, built on a thread pool
def test_hierarchy_invariants(name, kw):
    """Builder (rm_bvh.cpp): every hittable primitive sits in exactly one leaf, leaves hold at most four, a child's box
    lies inside the box its parent records for it, and the depth fits the traversal stack."""
    scene = workloads.build_scene(workloads.describe(name, **kw))
    nodes, prims, depth = emu.bvh(scene)
    assert 1 <= depth < 64 and len(nodes) >= 1
    assert len(set(prims.tolist())) == len(prims)              # no primitive twice
    kinds = prims.view(np.uint32) >> 30
    flat = scene.flatten().c
    assert (kinds == 0).sum() == flat.n_spheres
    assert (kinds > 0).sum() <= flat.n_polygons + flat.n_triangles
    seen = np.zeros(len(prims), dtype=np.int32)

    def boxes(n):
        a, b, z = n[0:4], n[4:8], n[8:12]
        return ((a[0], a[1], a[2], a[3], z[0], z[1]), (b[0], b[1], b[2], b[3], z[2], z[3]))

    def inside(inner, outer):
        return all(inner[k] >= outer[k] for k in (0, 2, 4)) and all(inner[k] <= outer[k] for k in (1, 3, 5))

    stack = [(0, None, 1)]
    deepest = 0
    while stack:
        i, bound, d = stack.pop()
        deepest = max(deepest, d)
        n = nodes[i]
        b0, b1 = boxes(n)
        for box, child in ((b0, int(n[12:13].view(np.int32)[0])), (b1, int(n[13:14].view(np.int32)[0]))):
            if bound is not None:
                assert inside(box, bound)
            if child >= 0:
                stack.append((child, box, d + 1))
            else:
                code = ~child
                first, cnt = code >> 3, code & 7
                assert cnt <= 4
                seen[first:first + cnt] += 1
    assert np.all(seen == 1)
    assert deepest <= depth
    nodes2, prims2, depth2 = emu.bvh(scene)                    # the build does not depend on thread timing
    assert np.array_equal(nodes.view(np.uint32), nodes2.view(np.uint32)) and np.array_equal(prims, prims2) and depth == depth2


def test_hierarchy_walk_cost():
    """Algorithmic cost of the hierarchy on the bench's stress scene (emulation-only counters, -DRM_EMU_STATS): a walk
    visits a few dozen of the 2594 nodes and tests a handful of the 9216 primitives.  A guard on the builder's quality
    (measured: 16.7 node visits and 5.3 primitive tests per walk)."""
    scene = workloads.build_scene(workloads.describe("stress", n_spheres=1024, grid=64))
    emu.walk_stats()
    r = emu.render(scene, 320, 192, "fast", max_depth=6, accel=True)
    walks, nodes, prims = emu.walk_stats()
    assert walks >= 320 * 192 and (r["prim_id"] >= 0).sum() > 5000
    assert nodes / walks < 25 and prims / walks < 8


def test_full_stress_scene_kernel_code_against_the_golden_frame():
    """The production kernel code (host emulation, hierarchy walk, f64 ray geometry on glass paths) on configs[4]'s full
    scene against the oracle's committed frame -- the case that FP32 ray geometry fails (99.5 % of the pixels within 1e-4)."""
    from tests.test_gpu_parity import load_stress_golden
    ref, depth, _mx = load_stress_golden()
    h, w = ref["prim_id"].shape
    scene = workloads.scene("stress")
    got = emu.render(scene, w, h, "fast", max_depth=depth, accel=True)
    assert got["glass_mode"] == 2
    rep = parity.check_fp32(got, ref, h)
    assert rep["frac_within_tol"] > 0.9995
    fp32 = emu.render(scene, w, h, "fast", max_depth=depth, accel=True, glass_mode=1)
    a, b = fp32["rgb"].astype(np.float64), ref["rgb"]
    rel = (np.abs(a - b) / np.maximum(np.abs(b), 1e-12)).max(axis=2)
    assert (rel <= parity.REL_TOL).mean() < parity.GOOD_FRACTION      # documents why the f64 geometry exists
