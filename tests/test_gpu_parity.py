"""Parity of the CUDA path against the oracle, through the C ABI, on a real B200 (-m gpu).

  * RM_FP64 validation kernels: primitive ids and event counters bit-exact, colours bit-exact up to
    the ulps by which CUDA's pow differs from glibc's (1e-13 relative);
  * RM_FP32 production kernels: the north_star criteria of tests/parity.py;
  * at BASELINE.json's full sizes: size-independent properties (tiles == whole frame, idempotence,
    culling on == off, FP32 agrees with FP64 on device, counters consistent).
"""
import ctypes as C
import hashlib

import numpy as np
import pytest

from oracle import oracle as O
from rusty_marcher_b200 import _abi, workloads
from tests import parity
from tests.oracle_scenes import build_oracle_scene

pytestmark = pytest.mark.gpu


def gpu_render(rm, scene, w, h, precision="f32", depth=3, cull=True, counters=False, rows=(0, -1), want_rgb8=True, accel=False):
    r = rm.create_renderer(1.5, h, w)
    r.max_depth, r.cull_backfacing, r.accel = depth, cull, accel
    r.precision = rm.RM_FP64 if precision == "f64" else rm.RM_FP32
    fb = rm.create_frame_buffer(w, h, dtype=np.float64 if precision == "f64" else np.float32)
    ids = np.full((h, w), -1, dtype=np.int32)
    rgb8 = np.zeros((h, w, 3), dtype=np.uint8) if want_rgb8 else None
    msg = r.render(fb, scene, prim_id=ids, rgb8=rgb8, counters=counters, patch_rows=rows)
    assert msg.startswith("Scene rendered in ")
    st = r.last_stats
    return {"rgb": fb.buffer, "prim_id": ids, "rgb8": rgb8, "counters": st.counters() if counters else None,
            "max": st.max_value, "stats": st}


CASES = [("demo", 800, 600, 3, {}), ("cornell_box", 640, 480, 3, {}), ("dodecahedron", 640, 480, 3, {}),
         ("stress", 160, 128, 6, dict(n_spheres=64, grid=7))]


@pytest.fixture(scope="module", params=CASES, ids=[c[0] for c in CASES])
def case(request, rm_gpu):
    name, w, h, depth, kw = request.param
    desc = workloads.describe(name, **kw)
    ref = O.render(build_oracle_scene(desc), w, h, max_depth=depth)
    return name, w, h, depth, workloads.build_scene(desc), ref


@pytest.mark.parametrize("cull", [False, True])
def test_fp64_kernels_match_the_oracle(rm_gpu, case, cull):
    name, w, h, depth, scene, ref = case
    got = gpu_render(rm_gpu, scene, w, h, "f64", depth, cull, counters=True)
    parity.check_exact(got, ref, rel=1e-13)
    assert abs(got["max"] - ref["rgb"].max()) <= 1e-13 * ref["rgb"].max()
    if not cull:
        assert got["counters"] == ref["counters"]
    else:
        for k in ("pixels", "closest_segments", "anyhit_segments", "hits", "light_evals", "lit_lights", "glass_hits",
                  "reflections", "refractions", "sphere_tests", "sphere_hits"):
            assert got["counters"][k] == ref["counters"][k], k
    rows = (h // 32) * 32
    d = np.abs(got["rgb8"][:rows].astype(np.int16) - parity.oracle_rgb8(ref["rgb"])[:rows].astype(np.int16))
    assert (d > 0).mean() < 1e-4 and d.max() <= 1              # only pow-ulp flips at a quantisation boundary


@pytest.mark.parametrize("cull", [False, True])
def test_fp32_kernels_meet_the_north_star_tolerance(rm_gpu, case, cull):
    name, w, h, depth, scene, ref = case
    got = gpu_render(rm_gpu, scene, w, h, "f32", depth, cull)
    rep = parity.check_fp32(got, ref, (h // 32) * 32, got["rgb8"])
    print(name, "cull", cull, rep)
    assert np.all(got["rgb"][(h // 32) * 32:] == 0)            # renderer.rs:47-55: rows never rendered stay untouched


@pytest.mark.parametrize("cull", [False, True])
def test_hierarchy_changes_no_pixel(rm_gpu, case, cull):
    """RmParams.accel = 1 (scene queries walk the bounding-volume hierarchy, rm_bvh.cuh): the same tests per (ray,
    primitive) and the same winner as the brute-force traversal, so the FP32 frame -- floats, primary ids, max, 8-bit
    bytes -- is bit-identical, and meets the north-star tolerance against the oracle like the brute-force frame."""
    name, w, h, depth, scene, ref = case
    a = gpu_render(rm_gpu, scene, w, h, "f32", depth, cull)
    b = gpu_render(rm_gpu, scene, w, h, "f32", depth, cull, accel=True)
    assert np.array_equal(a["prim_id"], b["prim_id"])
    assert np.array_equal(a["rgb"], b["rgb"]) and np.array_equal(a["rgb8"], b["rgb8"]) and a["max"] == b["max"]
    parity.check_fp32(b, ref, (h // 32) * 32, b["rgb8"])
    # the kernel's own query count: rendered pixels + queries behind the primary rays = the oracle's segments, up to the
    # handful of near-tie pixels where FP32 and f64 take different branches
    L = _abi.load()
    q = C.c_uint64(0)
    _abi.check(L.rm_scene_query_count(scene.device_handle(), C.byref(q), 1))
    gpu_render(rm_gpu, scene, w, h, "f32", depth, cull, accel=True, want_rgb8=False)
    _abi.check(L.rm_scene_query_count(scene.device_handle(), C.byref(q), 1))
    want = ref["counters"]["closest_segments"] + ref["counters"]["anyhit_segments"]
    assert abs((h // 32) * 32 * w + q.value - want) <= 2e-3 * want
    words = (C.c_int32 * 16)()
    assert L.rm_scene_accel_status(scene.device_handle(), words) == 0 and words[0] == 0


def test_hierarchy_walk_statistics(rm_gpu):
    """RmParams.accel = 2: the counting instantiation renders the very same frame and reports the walks' work -- node
    visits and leaf tests -- which must agree with the counts of the host emulation of the same kernel code."""
    from tests.emu import emu
    w, h, depth = 640, 352, 6
    scene = workloads.build_scene(workloads.describe("stress", n_spheres=512, grid=32))
    a = gpu_render(rm_gpu, scene, w, h, "f32", depth=depth, accel=True)
    L = _abi.load()
    ws = (C.c_uint64 * 3)()
    _abi.check(L.rm_scene_walk_stats(scene.device_handle(), ws, 1))
    r = rm_gpu.create_renderer(1.5, h, w)
    r.max_depth, r.accel = depth, True
    fb = rm_gpu.create_frame_buffer(w, h, dtype=np.float32)
    ids = np.full((h, w), -1, dtype=np.int32)
    p = r.params(fb, scene)
    p.accel = 2
    st = _abi.RmStats()
    _abi.check(L.rm_render(scene.device_handle(), C.byref(p), fb.buffer.ctypes.data, ids.ctypes.data, None, C.byref(st)))
    _abi.check(L.rm_scene_walk_stats(scene.device_handle(), ws, 1))
    assert np.array_equal(fb.buffer, a["rgb"]) and np.array_equal(ids, a["prim_id"])
    nodes, sph, pln = int(ws[0]), int(ws[1]), int(ws[2])
    assert nodes > 10 * w * h and sph > 0 and pln > 0
    emu.lib().emu_walk_stats((C.c_ulonglong * 3)())
    emu.render(scene, w, h, "fast", max_depth=depth, accel=True)
    ew = (C.c_ulonglong * 3)()
    emu.lib().emu_walk_stats(ew)
    assert abs(nodes - int(ew[1])) <= 0.01 * nodes and abs(sph + pln - int(ew[2])) <= 0.01 * (sph + pln), (nodes, sph, pln, list(ew))
    _abi.check(L.rm_scene_walk_stats(scene.device_handle(), ws, 0))
    assert int(ws[0]) == 0


def test_hierarchy_on_the_stress_scene(rm_gpu):
    """The bench's stress scene (1024 spheres + 8192 triangles, depth cap 6) at 1280x704, camera off the origin: hierarchy
    against brute force bit for bit, whole frame and interleaved bands, host call and frame-level call."""
    import torch
    from rusty_marcher_b200 import tiled
    w, h = 1280, 704
    scene = workloads.build_scene(workloads.describe("stress", n_spheres=1024, grid=64))
    scene.offset_camera((7.5, -3.25, 20.0))
    a = gpu_render(rm_gpu, scene, w, h, "f32", depth=6)
    b = gpu_render(rm_gpu, scene, w, h, "f32", depth=6, accel=True)
    assert (a["prim_id"] >= 0).sum() > 50000
    assert np.array_equal(a["prim_id"], b["prim_id"]) and np.array_equal(a["rgb"], b["rgb"]) and np.array_equal(a["rgb8"], b["rgb8"])
    acc = np.zeros_like(a["rgb"])
    for k in range(3):
        t = gpu_render(rm_gpu, scene, w, h, "f32", depth=6, accel=True, rows=(k, -1, 3), want_rgb8=False)
        for pr in range(k, h // 32, 3):
            acc[pr * 32:(pr + 1) * 32] = t["rgb"][pr * 32:(pr + 1) * 32]
    assert np.array_equal(acc, a["rgb"])
    dev = torch.device("cuda:0")
    r = rm_gpu.create_renderer(1.5, h, w)
    r.max_depth, r.accel = 6, "auto"
    tr = tiled.TiledRenderer(tiled.CudaBackend(scene, r, w, h, dev), w, h, dev)
    try:
        assert tr.params.accel == 1
        f = tr.render()
        torch.cuda.synchronize()
        tr.peer.status()
        assert np.array_equal(f.cpu().numpy(), a["rgb8"]) and np.array_equal(tr.rgb.cpu().numpy(), a["rgb"])
    finally:
        tr.close()


@pytest.mark.parametrize("n_lights", [3, 4, 8, 9])
def test_hierarchy_with_more_lights(rm_gpu, n_lights):
    """Several pairs of lights through the hierarchy kernel (unrolled pairs up to 8 lights; 9 lights: the call stays on
    the brute-force kernel) against brute force, bit for bit, twice (the frame must also be reproducible)."""
    from tests.test_kernel_emulation import _with_lights
    w, h = 640, 352
    scene = workloads.build_scene(_with_lights(workloads.describe("stress", n_spheres=512, grid=32), n_lights))
    scene.offset_camera((7.5, -3.25, 20.0))
    a = gpu_render(rm_gpu, scene, w, h, "f32", depth=4)
    b = gpu_render(rm_gpu, scene, w, h, "f32", depth=4, accel=True)
    c = gpu_render(rm_gpu, scene, w, h, "f32", depth=4, accel=True)
    assert (a["prim_id"] >= 0).sum() > 10000
    assert np.array_equal(a["prim_id"], b["prim_id"]) and np.array_equal(a["rgb"], b["rgb"]) and np.array_equal(a["rgb8"], b["rgb8"])
    assert np.array_equal(b["prim_id"], c["prim_id"]) and np.array_equal(b["rgb"], c["rgb"])
    words = (C.c_int32 * 16)()
    assert _abi.load().rm_scene_accel_status(scene.device_handle(), words) == 0 and words[0] == 0


def test_demo_golden_image_on_gpu(rm_gpu):
    """The GPU's FP64 render of the demo scene against the reference's golden engine/out.ppm."""
    scene = workloads.scene("demo")
    got = gpu_render(rm_gpu, scene, 800, 600, "f64", cull=False)
    rgb = got["rgb"].copy()
    O.normalize(rgb)
    ppm = O.ppm_bytes(rgb)
    golden = hashlib.sha256(ppm).hexdigest() == "82d51afaaf4a644547728dde89478e484c245d2e3eb1e40da8b928ebd7584797"
    if not golden:      # a pow ulp at a quantisation boundary may flip single bytes; nothing more
        ref = O.render(O.Scene.create_default(), 800, 600, want_ids=False, want_fragile=False, want_counters=False)
        a = np.frombuffer(ppm[15:], dtype=np.uint8).astype(np.int16)
        b = parity.oracle_rgb8(ref["rgb"]).ravel().astype(np.int16)
        assert np.abs(a - b).max() <= 1 and (a != b).mean() < 1e-5
    # and the device-side normalize + to_vec (K4) reproduces the same bytes from the device max
    assert np.abs(got["rgb8"].ravel().astype(np.int16) - np.frombuffer(ppm[15:], dtype=np.uint8).astype(np.int16)).max() <= 1


def test_cornell_1080p_config2(rm_gpu):
    """BASELINE.json config 2 at full size against the oracle (the oracle needs about a second)."""
    desc = workloads.describe("cornell_box")
    ref = O.render(build_oracle_scene(desc), 1920, 1080)
    scene = workloads.build_scene(desc)
    got = gpu_render(rm_gpu, scene, 1920, 1080, "f32")
    rep = parity.check_fp32(got, ref, 1056, got["rgb8"])
    assert rep["id_mismatch"] == 0
    hit = got["prim_id"] >= 0
    assert hit.sum() == (ref["prim_id"] >= 0).sum() == 279591
    assert set(np.unique(got["prim_id"][hit]).tolist()) == set(np.unique(ref["prim_id"][ref["prim_id"] >= 0]).tolist()) == {24, 25, 33}
    g64 = gpu_render(rm_gpu, scene, 1920, 1080, "f64", counters=True, cull=False)
    parity.check_exact(g64, ref, rel=1e-13)
    assert g64["counters"] == ref["counters"]


@pytest.mark.parametrize("name,w,h", [("cornell_box", 3840, 2160), ("dodecahedron", 3840, 2160), ("demo", 1600, 1280)])
def test_full_size_against_the_oracle_and_properties(rm_gpu, name, w, h):
    """Configs 1, 3, 4 at BASELINE.json's full sizes: the FP32 production frame against the ORACLE's frame of the same
    size (north-star criteria, fragile mask and all; the oracle needs one to two seconds for a 4K frame), the FP64
    validation kernel bit-exact against it, then the size-independent properties."""
    desc = workloads.describe(name)
    ref = O.render(build_oracle_scene(desc), w, h)
    scene = workloads.build_scene(desc)
    whole = gpu_render(rm_gpu, scene, w, h, "f32")
    rep = parity.check_fp32(whole, ref, (h // 32) * 32, whole["rgb8"])
    print(name, w, h, rep)
    assert abs(whole["max"] - ref["rgb"].max()) <= 1e-4 * ref["rgb"].max()
    # the instrumented launch uses the generic FP32 kernel (other roundings than the production one): counters only
    counted = gpu_render(rm_gpu, scene, w, h, "f32", counters=True)
    rows = (h // 32) * 32
    assert (counted["prim_id"] != whole["prim_id"]).mean() < 2e-4
    n_patch = h // 32
    # idempotence
    again = gpu_render(rm_gpu, scene, w, h, "f32")
    assert np.array_equal(whole["rgb"], again["rgb"]) and np.array_equal(whole["prim_id"], again["prim_id"])
    # row tiles of 2, 4, 8 ranks reassemble to the whole frame, and the tile maxima combine to the frame max
    for g in (2, 8):
        acc = np.zeros_like(whole["rgb"])
        ids = np.full_like(whole["prim_id"], -1)
        mx = 0.
        for k in range(g):
            a, b = (k * n_patch) // g, ((k + 1) * n_patch) // g
            t = gpu_render(rm_gpu, scene, w, h, "f32", rows=(a, b), want_rgb8=False)
            acc[a * 32:b * 32] = t["rgb"][a * 32:b * 32]
            ids[a * 32:b * 32] = t["prim_id"][a * 32:b * 32]
            assert np.all(t["rgb"][:a * 32] == 0) and np.all(t["rgb"][b * 32:] == 0)
            mx = max(mx, t["max"])
        assert np.array_equal(acc, whole["rgb"]) and np.array_equal(ids, whole["prim_id"])
        assert mx == whole["max"] == float(whole["rgb"].max())
    # culling never changes a pixel
    nocull = gpu_render(rm_gpu, scene, w, h, "f32", cull=False)
    assert np.array_equal(nocull["prim_id"], whole["prim_id"])
    assert np.array_equal(nocull["rgb"], whole["rgb"])
    # the device's FP64 validation kernel at full size: bit-exact ids, counters equal to the oracle's (culling on: the
    # control-flow counters), colours up to pow ulps -- and the FP32 frame against it
    g64 = gpu_render(rm_gpu, scene, w, h, "f64", counters=True)
    parity.check_exact(g64, ref, rel=1e-13)
    for k in ("pixels", "closest_segments", "anyhit_segments", "hits", "light_evals", "lit_lights", "glass_hits", "reflections", "refractions"):
        assert g64["counters"][k] == ref["counters"][k], k
    mism = whole["prim_id"] != g64["prim_id"]
    assert mism.mean() < 2e-4
    rel = (np.abs(whole["rgb"][:rows].astype(np.float64) - g64["rgb"][:rows]) / np.maximum(np.abs(g64["rgb"][:rows]), 1e-12)).max(axis=2)
    assert (rel <= parity.REL_TOL).mean() >= parity.GOOD_FRACTION
    d = np.abs(whole["rgb8"][:rows].astype(np.int16) - g64["rgb8"][:rows].astype(np.int16)).max(axis=2)
    assert (d <= 1).mean() >= parity.GOOD_FRACTION
    # counters: one closest segment per pixel at least, shadow rays only from hits, same control flow in both precisions
    c32, c64 = counted["counters"], g64["counters"]
    assert c32["pixels"] == rows * w == c64["pixels"]
    assert c32["closest_segments"] >= c32["pixels"] and c32["anyhit_segments"] == c32["light_evals"] == 2 * c32["hits"]
    for k in ("closest_segments", "hits", "glass_hits"):
        assert abs(c32[k] - c64[k]) <= 2e-4 * max(c64[k], 1), k
    # rows below the last patch row are untouched, rgb8 there is zero
    assert np.all(whole["rgb"][rows:] == 0) and np.all(whole["rgb8"][rows:] == 0)


STRESS_PARITY = dict(n_spheres=2048, grid=32)                   # 2048 spheres + 2048 triangles
STRESS_PARITY_CAMERA = (-20., 8., -90.)                         # inside the cloud of spheres: a quarter of them glass


def test_stress_scene_against_the_oracle(rm_gpu):
    """The stress generator (BASELINE.json configs[4]) at a size that exercises what the full scene exercises -- glass
    spheres recursing to depth 6 over thousands of primitives: 2048 spheres + a 32x32 quad grid (2048 triangles), 640x352,
    camera inside the cloud of spheres, depth cap 6 (52,729 glass hits, 1.3 closest-hit segments per pixel) -- against the
    oracle: brute force and the hierarchy, FP32 production kernels; FP64 validation kernel bit-exact with equal counters."""
    w, h, depth = 640, 352, 6
    desc = workloads.describe("stress", **STRESS_PARITY)
    osc = build_oracle_scene(desc)
    osc.offset_camera(STRESS_PARITY_CAMERA)
    ref = O.render(osc, w, h, max_depth=depth)
    assert ref["counters"]["glass_hits"] > 40000 and ref["counters"]["closest_segments"] > 1.25 * ref["counters"]["pixels"]
    scene = workloads.build_scene(desc)
    scene.offset_camera(STRESS_PARITY_CAMERA)
    a = gpu_render(rm_gpu, scene, w, h, "f32", depth=depth)
    print("stress brute force", parity.check_fp32(a, ref, h, a["rgb8"]))
    b = gpu_render(rm_gpu, scene, w, h, "f32", depth=depth, accel=True)
    print("stress hierarchy  ", parity.check_fp32(b, ref, h, b["rgb8"]))
    assert np.array_equal(a["prim_id"], b["prim_id"]) and np.array_equal(a["rgb"], b["rgb"]) and np.array_equal(a["rgb8"], b["rgb8"])
    g64 = gpu_render(rm_gpu, scene, w, h, "f64", depth=depth, cull=False, counters=True)
    parity.check_exact(g64, ref, rel=1e-12)
    assert g64["counters"] == ref["counters"]


def load_stress_golden():
    """The oracle's frame of configs[4]'s full scene (tests/golden/make_stress_golden.py: 86 s of CPU here)."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stress_full_384x224.npz"))
    ref = {"rgb": z["rgb"].astype(np.float64), "prim_id": z["prim_id"], "fragile": z["fragile"],
           "counters": dict(zip(O.COUNTER_FIELDS, (int(v) for v in z["counters"])))}
    return ref, int(z["depth"]), float(z["rgb_max"])


def test_full_stress_scene_against_the_oracles_golden_frame(rm_gpu):
    """BASELINE.json configs[4]'s scene at FULL size -- 4096 spheres + 100,352 triangles, depth cap 6 -- through the hierarchy,
    against the oracle's frame of the same scene (a committed fixture: the reference's brute force needs 10^5 primitive
    tests per segment).  North-star criteria; the FP64 validation kernel (brute force, a few seconds on the GPU) must
    reproduce the oracle's ids and counters exactly."""
    ref, depth, ref_max = load_stress_golden()
    h, w = ref["prim_id"].shape
    assert ref["counters"]["glass_hits"] > 10000 and ref["counters"]["plane_tests"] > 1e10
    scene = workloads.scene("stress")
    assert scene.num_prims == 104448
    got = gpu_render(rm_gpu, scene, w, h, "f32", depth=depth, accel=True)
    rep = parity.check_fp32(got, ref, h)
    print("full stress scene, hierarchy:", rep)
    assert abs(got["max"] - ref_max) <= 1e-4 * ref_max
    g64 = gpu_render(rm_gpu, scene, w, h, "f64", depth=depth, cull=False, counters=True)
    assert np.array_equal(g64["prim_id"], ref["prim_id"])
    assert g64["counters"] == ref["counters"]
    err = np.abs(g64["rgb"] - ref["rgb"]) / np.maximum(np.abs(ref["rgb"]), 1e-30)
    assert err.max() < 1e-6                                    # the fixture stores the oracle's colours as float32


@pytest.mark.parametrize("cull", [False, True])
def test_many_ngons_against_the_oracle(rm_gpu, cull):
    """320 convex n-gons (n = 4 ... 8, one in ten wound clockwise = never hittable, a fifth glass-like) + 48 spheres, depth
    cap 4: the n-gon routines of the production kernel (stage A's primary_rest, plane_intersect in the shading stage)
    on far more than the demo scene's two polygons -- against the oracle (polygon.rs:60-98), brute force and hierarchy;
    the frame must also be reproducible run to run."""
    w, h, depth = 640, 352, 4
    desc = workloads.describe("ngons")
    ref = O.render(build_oracle_scene(desc), w, h, max_depth=depth)
    assert len(np.unique(ref["prim_id"])) > 250 and ref["counters"]["glass_hits"] > 10000
    scene = workloads.build_scene(desc)
    a = gpu_render(rm_gpu, scene, w, h, "f32", depth=depth, cull=cull)
    print("ngons brute force", parity.check_fp32(a, ref, h, a["rgb8"]))
    a2 = gpu_render(rm_gpu, scene, w, h, "f32", depth=depth, cull=cull)
    assert np.array_equal(a["prim_id"], a2["prim_id"]) and np.array_equal(a["rgb"], a2["rgb"])
    b = gpu_render(rm_gpu, scene, w, h, "f32", depth=depth, cull=cull, accel=True)
    parity.check_fp32(b, ref, h, b["rgb8"])
    assert np.array_equal(a["prim_id"], b["prim_id"]) and np.array_equal(a["rgb"], b["rgb"]) and np.array_equal(a["rgb8"], b["rgb8"])
    g64 = gpu_render(rm_gpu, scene, w, h, "f64", depth=depth, cull=cull, counters=True)
    parity.check_exact(g64, ref, rel=1e-12)
    if not cull:
        assert g64["counters"] == ref["counters"]
    # a camera off the axis and a 4K-wide strip of the same scene (many strips per warp: the schedule's other regime)
    scene.offset_camera((3., -2., 6.))
    osc = build_oracle_scene(desc)
    osc.offset_camera((3., -2., 6.))
    ref2 = O.render(osc, 3840, 160, max_depth=depth)
    c = gpu_render(rm_gpu, scene, 3840, 160, "f32", depth=depth, cull=cull)
    parity.check_fp32(c, ref2, 160, c["rgb8"])


def test_error_behaviour(rm_gpu):
    rm = rm_gpu
    L = _abi.load()
    scene = workloads.scene("demo")
    fb = rm.create_frame_buffer(100, 64)                        # width not a multiple of 32: renderer.rs:107 panics
    with pytest.raises(rm.RmError) as e:
        rm.create_renderer(1.5, 64, 100).render(fb, scene)
    assert e.value.code == -4 and "Dimensions mismatch" in str(e.value)
    p = _abi.RmParams()
    L.rm_params_default(C.byref(p), 64, 64)
    assert L.rm_render(987654, C.byref(p), None, None, None, None) == -3       # unknown handle
    assert b"unknown scene handle" in L.rm_last_error()
    p.max_depth = 99
    assert L.rm_render(scene.device_handle(), C.byref(p), None, None, None, None) == -3
    # malformed flat scene
    flat = scene.flatten()
    flat.c.shapes[0].index = 77
    h = C.c_int64(0)
    assert L.rm_scene_upload(C.byref(flat.c), C.byref(h)) == -5
    # empty scene renders black, misses everywhere
    empty = rm.Scene.new()
    got = gpu_render(rm, empty, 64, 64, "f32")
    assert np.all(got["rgb"] == 0) and np.all(got["prim_id"] == -1) and np.all(got["rgb8"] == 0)
    # height not a multiple of 32 only prints (renderer.rs:49-51) and renders floor(H/32) patch rows
    got = gpu_render(rm, scene, 64, 70, "f32")
    assert np.any(got["rgb"][:64] > 0) and np.all(got["rgb"][64:] == 0)


def test_depth_and_camera(rm_gpu):
    scene = workloads.scene("demo")
    osc = O.Scene.create_default()
    for depth in (0, 1, 2, 5):
        parity.check_exact(gpu_render(rm_gpu, scene, 96, 64, "f64", depth=depth, cull=False),
                           O.render(osc, 96, 64, max_depth=depth), rel=1e-13)
    scene.offset_camera((5., 0., -5.))                          # main.rs:124-171
    osc.offset_camera((5., 0., -5.))
    ref = O.render(osc, 256, 160)
    parity.check_exact(gpu_render(rm_gpu, scene, 256, 160, "f64"), ref, rel=1e-13)
    parity.check_fp32(gpu_render(rm_gpu, scene, 256, 160, "f32"), ref, 160)


def test_device_api_with_torch_buffers(rm_gpu):
    """rm_render_device / rm_tonemap_device on caller-owned device memory and stream."""
    import torch
    from rusty_marcher_b200 import tiled
    w, h = 640, 480
    scene = workloads.scene("cornell_box")
    dev = torch.device("cuda:0")
    r = rm_gpu.create_renderer(1.5, h, w)
    tr = tiled.TiledRenderer(tiled.CudaBackend(scene, r, w, h, dev), w, h, dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        frame = tr.render()
    s.synchronize()
    host = gpu_render(rm_gpu, scene, w, h, "f32")
    assert np.array_equal(tr.rgb.cpu().numpy(), host["rgb"])
    assert np.array_equal(frame.cpu().numpy(), host["rgb8"])
    assert float(tr.dmax.item()) == host["max"]


@pytest.mark.parametrize("name,w,h", [("cornell_box", 1920, 1080), ("dodecahedron", 640, 480), ("demo", 800, 600)])
def test_fused_rgb8_and_interleaved_bands_on_device(rm_gpu, name, w, h):
    """Device API: (1) rm_render_device_rgb8 + rm_tonemap_device_busy (render kernel zero-fills the 8-bit frame, tone map
    converts the busy tiles only) gives the very bytes of rm_render_device + rm_tonemap_device; (2) patch rows dealt
    round-robin (patch_row_stride) to 2 and 3 'ranks' reassemble to the frame rendered in one piece, bit for bit."""
    import torch
    from rusty_marcher_b200 import tiled
    dev = torch.device("cuda:0")
    scene = workloads.scene(name)
    r = rm_gpu.create_renderer(1.5, h, w)
    be = tiled.CudaBackend(scene, r, w, h, dev)
    n_patch = h // 32

    def frame(rows_list, fused):
        rgb = torch.zeros((h, w, 3), dtype=torch.float32, device=dev)
        rgb8 = torch.full((h, w, 3), 7, dtype=torch.uint8, device=dev)      # poisoned: every rendered byte must be written
        ids = torch.full((h, w), -2, dtype=torch.int32, device=dev)
        dmax = torch.zeros(1, dtype=torch.float32, device=dev)
        for rows in rows_list:
            be.render_rows(rows, rgb, dmax, prim=ids, rgb8=rgb8 if fused else None)
            if fused:           # the schedule of a scene handle describes its LAST render: convert before the next one
                be.tonemap_rows(rows, rgb, dmax, rgb8, busy=True, normalise=False)
        if not fused:
            for rows in rows_list:
                be.tonemap_rows(rows, rgb, dmax, rgb8, normalise=False)
        torch.cuda.synchronize()
        return rgb.cpu().numpy(), rgb8.cpu().numpy(), ids.cpu().numpy(), float(dmax.item())

    whole = frame([(0, n_patch)], fused=False)
    rows = n_patch * 32
    assert np.all(whole[1][rows:] == 7) and np.all(whole[2][rows:] == -2)       # rows below the last patch row stay untouched
    for rows_list, fused in (([(0, n_patch)], True), ([(k, n_patch, 2) for k in range(2)], False),
                             ([(k, n_patch, 3) for k in range(3)], True), ([(0, 1), (1, n_patch)], True)):
        got = frame(rows_list, fused)
        assert np.array_equal(got[0], whole[0]), (rows_list, fused)
        assert np.array_equal(got[1], whole[1]), (rows_list, fused)
        assert np.array_equal(got[2], whole[2]) and got[3] == whole[3]


def test_host_api_interleaved_bands(rm_gpu):
    scene = workloads.scene("cornell_box")
    w, h = 640, 480
    whole = gpu_render(rm_gpu, scene, w, h, "f32")
    acc_rgb, acc8, acc_id = np.zeros_like(whole["rgb"]), np.zeros_like(whole["rgb8"]), np.full_like(whole["prim_id"], -1)
    for k in range(3):
        r = rm_gpu.create_renderer(1.5, h, w)
        fb = rm_gpu.create_frame_buffer(w, h, dtype=np.float32)
        ids = np.full((h, w), -1, dtype=np.int32)
        rgb8 = np.zeros((h, w, 3), dtype=np.uint8)
        p = r.params(fb, scene, (k, -1))
        p.patch_row_stride = 3
        st = _abi.RmStats()
        _abi.check(_abi.load().rm_render(scene.device_handle(), C.byref(p), fb.buffer.ctypes.data, ids.ctypes.data, rgb8.ctypes.data, C.byref(st)))
        band = np.zeros(h, dtype=bool)
        for pr in range(k, h // 32, 3):
            band[pr * 32:(pr + 1) * 32] = True
        assert np.all(fb.buffer[~band] == 0) and np.all(ids[~band] == -1)     # other ranks' rows are not touched
        acc_rgb[band], acc_id[band] = fb.buffer[band], ids[band]
    assert np.array_equal(acc_rgb, whole["rgb"]) and np.array_equal(acc_id, whole["prim_id"])


@pytest.mark.parametrize("name,w,h,depth", [("cornell_box", 1920, 1080, 3), ("dodecahedron", 640, 480, 3), ("demo", 800, 600, 3)])
def test_host_delivery_of_rows(rm_gpu, name, w, h, depth):
    """What Renderer::render hands back (renderer.rs:92-108): the float frame in HOST memory.  The delivery of busy tiles
    only (staged: packed on the device, chunked copies, host threads scatter; pinned frames: written by the device itself;
    host threads clear the black tiles either way) against the plain copy of every row: float32 frame,
    the reference's f64 rows (contiguous and one allocation per row, framebuffer.rs:6-10), the retained mode over a moving
    camera, interleaved bands of several 'ranks' into one frame -- all bit-identical, on poisoned buffers."""
    rm = rm_gpu
    L = _abi.load()
    cams = [(0., 0., 0.), (30., -20., 10.), (-45., 15., 0.), (0., 0., 0.)]
    if name == "demo":
        cams = [(0., 0., 0.), (3., -2., 1.), (-4., 1.5, 0.), (0., 0., 0.)]
    rows = (h // 32) * 32
    want = []
    for cam in cams:                                           # the plain path: every row copied (prim_id requested)
        scene = workloads.scene(name)
        scene.offset_camera(cam)
        want.append(gpu_render(rm, scene, w, h, "f32", depth)["rgb"])
    assert not np.array_equal(want[0], want[1])
    scene = workloads.scene(name)
    r = rm.create_renderer(1.5, h, w)
    r.max_depth = depth
    # float32 frame, not retained: poisoned before every call
    fb = rm.create_frame_buffer(w, h, dtype=np.float32)
    for cam, ref in zip(cams, want):
        scene.camera = type(scene.camera)(*cam)
        fb.buffer[:] = 7.
        r.render(fb, scene)
        assert np.array_equal(fb.buffer[:rows], ref[:rows]) and np.all(fb.buffer[rows:] == 7.)
        assert r.last_stats.max_value == float(ref.max())
    # the reference's frame: f64 rows; poisoned once, then retained over the moving camera
    fb64 = rm.create_frame_buffer(w, h, dtype=np.float64)
    fb64.buffer[:] = 7.
    r.retained = True
    for cam, ref in zip(cams, want):
        scene.camera = type(scene.camera)(*cam)
        r.render(fb64, scene)
        assert fb64.buffer.dtype == np.float64
        assert np.array_equal(fb64.buffer[:rows], ref[:rows].astype(np.float64)) and np.all(fb64.buffer[rows:] == 7.)
    # retained float32, and a caller that scribbles in between without saying so gets what it asked for only when it says so
    fb32 = rm.create_frame_buffer(w, h, dtype=np.float32)
    fb32.buffer[:] = 7.
    for cam, ref in zip(cams, want):
        scene.camera = type(scene.camera)(*cam)
        r.render(fb32, scene)
        assert np.array_equal(fb32.buffer[:rows], ref[:rows])
    r.retained = False
    fb32.buffer[:] = 3.
    r.render(fb32, scene)
    assert np.array_equal(fb32.buffer[:rows], want[-1][:rows])
    # a float32 frame in PINNED host memory (rm_host_alloc): the device writes the busy tiles straight into it (no staging,
    # no scatter; three launches instead of four) -- fresh, retained, and as interleaved bands of three 'ranks'
    pin = L.rm_host_alloc(h * w * 12)
    assert pin
    try:
        fbp = rm.create_frame_buffer(32, 32)
        fbp.width, fbp.height = w, h
        fbp.buffer = np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_float)), shape=(h, w, 3))
        r.render(fb, scene)
        staged_launches = r.last_stats.kernel_launches
        for retained in (False, True):                           # (retained: the frame as the previous delivery left it)
            r.retained = retained
            for cam, ref in zip(cams, want):
                scene.camera = type(scene.camera)(*cam)
                if not retained:
                    fbp.buffer[:] = 7.
                r.render(fbp, scene)
                assert np.array_equal(fbp.buffer[:rows], ref[:rows]) and np.all(fbp.buffer[rows:] == 7.)
                assert r.last_stats.max_value == float(ref.max())
                if name == "cornell_box":
                    assert r.last_stats.kernel_launches == staged_launches - 1 == 3
        r.retained = False
        fbp.buffer[:] = 7.
        scene.camera = type(scene.camera)(*cams[2])
        for k in range(3):
            p = r.params(fbp, scene, (k, -1, 3))
            st = _abi.RmStats()
            _abi.check(L.rm_render(scene.device_handle(), C.byref(p), fbp.buffer.ctypes.data, None, None, C.byref(st)))
        assert np.array_equal(fbp.buffer[:rows], want[2][:rows]) and np.all(fbp.buffer[rows:] == 7.)
        fbp.buffer = None
    finally:
        L.rm_host_free(pin)
    # one allocation per row (Vec<Vec<Vec3f>>), bands of three 'ranks' into the same frame
    row_arrays = [np.full((w, 3), 7., dtype=np.float64) for _ in range(h)]
    ptrs = (C.c_void_p * h)(*[a.ctypes.data for a in row_arrays])
    scene.camera = type(scene.camera)(*cams[1])
    for k in range(3):
        p = r.params(fb, scene, (k, -1, 3))
        st = _abi.RmStats()
        _abi.check(L.rm_render_rows_f64(scene.device_handle(), C.byref(p), ptrs, 0, C.byref(st)))
    got = np.stack(row_arrays)
    assert np.array_equal(got[:rows], want[1][:rows].astype(np.float64)) and np.all(got[rows:] == 7.)
    assert L.rm_render_rows_f64(scene.device_handle(), C.byref(p), None, 0, None) == -3


def test_refused_host_register_does_not_poison_later_calls(rm_gpu):
    """A recoverable CUDA failure reported through the ABI (registering the same range twice) is not reported again by
    the next call that checks for launch errors."""
    L = _abi.load()
    buf = np.zeros(1 << 20, dtype=np.uint8)
    assert L.rm_host_register(buf.ctypes.data, buf.nbytes) == 0
    try:
        assert L.rm_host_register(buf.ctypes.data, buf.nbytes) != 0
        assert b"cudaHostRegister" in L.rm_last_error()
        scene = workloads.scene("cornell_box")
        got = gpu_render(rm_gpu, scene, 256, 160, "f32")
        assert (got["prim_id"] >= 0).any()
    finally:
        assert L.rm_host_unregister(buf.ctypes.data) == 0


def test_kernel_profiling_events(rm_gpu):
    import torch
    from rusty_marcher_b200 import tiled
    L = _abi.load()
    a, b = C.c_double(0), C.c_double(0)
    assert L.rm_set_profiling(1) == 0
    assert L.rm_last_kernel_times(C.byref(a), C.byref(b)) != 0                  # nothing rendered yet
    dev = torch.device("cuda:0")
    w, h = 640, 480
    be = tiled.CudaBackend(workloads.scene("cornell_box"), rm_gpu.create_renderer(1.5, h, w), w, h, dev)
    rgb = torch.zeros((h, w, 3), dtype=torch.float32, device=dev)
    dmax = torch.zeros(1, dtype=torch.float32, device=dev)
    be.render_rows((0, h // 32), rgb, dmax)
    _abi.check(L.rm_last_kernel_times(C.byref(a), C.byref(b)))
    assert 0 < a.value < 50 and 0 < b.value < 50
    assert L.rm_set_profiling(0) == 0


def test_fp32_peak_probe(rm_gpu):
    t, ms = C.c_double(0), C.c_double(0)
    _abi.check(_abi.load().rm_measure_fp32_peak(C.byref(t), C.byref(ms)))
    assert 20. < t.value < 90., t.value                        # 148 SMs x 128 lanes x 2 x ~1.9 GHz = 72 TFLOP/s nominal


@pytest.mark.parametrize("name,w,h", [("cornell_box", 1920, 1080), ("demo", 800, 600), ("dodecahedron", 640, 480)])
def test_frame_level_call_equals_the_two_step_path(rm_gpu, name, w, h):
    """rm_render_frame (K0 + K1 + K4, the exchange done by the kernels through the mailbox) with world = 1 against
    rm_render: same floats, same ids, same bytes; consecutive frames alternate between the two 8-bit buffers."""
    import torch
    from rusty_marcher_b200 import tiled
    dev = torch.device("cuda:0")
    scene = workloads.scene(name)
    whole = gpu_render(rm_gpu, scene, w, h, "f32")
    be = tiled.CudaBackend(scene, rm_gpu.create_renderer(1.5, h, w), w, h, dev)
    tr = tiled.TiledRenderer(be, w, h, dev)
    assert tr.exchange == "peer" and tr.world == 1
    import os
    graphs_before = _abi.load().rm_graph_launch_count()
    try:
        frames, ptrs = [], []
        for i in range(5):
            # frames 3..5 as ONE CUDA graph launch each (RM_B200_GRAPH is read per call; torch's legacy default stream here:
            # captured on a stream of the library's own, launched on the caller's), camera parameters pushed by graph update
            os.environ["RM_B200_GRAPH"] = "1" if i >= 2 else "0"
            f = tr.render()
            torch.cuda.synchronize()
            tr.peer.status()
            ptrs.append(f.data_ptr())
            frames.append(f.cpu().numpy().copy())
            assert np.array_equal(tr.rgb.cpu().numpy(), whole["rgb"])
            assert float(tr.dmax.item()) == np.float32(whole["max"])
        assert ptrs[0] != ptrs[1] and ptrs[0] == ptrs[2]
        for f in frames:
            assert np.array_equal(f, whole["rgb8"])
        assert _abi.load().rm_graph_launch_count() - graphs_before == 3
    finally:
        os.environ.pop("RM_B200_GRAPH", None)
        tr.close()


def test_frame_level_call_argument_checks(rm_gpu):
    import torch
    from rusty_marcher_b200 import tiled
    dev = torch.device("cuda:0")
    w, h = 64, 64
    scene = workloads.scene("demo")
    be = tiled.CudaBackend(scene, rm_gpu.create_renderer(1.5, h, w), w, h, dev)
    tr = tiled.TiledRenderer(be, w, h, dev)
    L = _abi.load()
    try:
        stream = torch.cuda.current_stream().cuda_stream
        args = (be.handle, C.byref(tr.params), tr.rgb.data_ptr(), None, tr.dmax.data_ptr())
        assert L.rm_render_frame(*args, C.byref(tr.peer.x), 0, 1, stream) == -3            # sequence numbers start at 1
        assert L.rm_render_frame(*args, None, 1, 1, stream) == -3
        bad = _abi.RmExchange()
        bad.rank, bad.world = 0, 2                                                           # second mailbox missing
        bad.mailbox[0] = tr.peer.x.mailbox[0]
        bad.frame8[0], bad.frame8[1] = tr.peer.x.frame8[0], tr.peer.x.frame8[1]
        assert L.rm_render_frame(*args, C.byref(bad), 1, 1, stream) == -3
        bad.world, bad.rank = 1, 1
        assert L.rm_render_frame(*args, C.byref(bad), 1, 1, stream) == -3
        p64 = be.frame_params(tr.rows)
        p64.precision = _abi.RM_FP64
        assert L.rm_render_frame(be.handle, C.byref(p64), tr.rgb.data_ptr(), None, tr.dmax.data_ptr(), C.byref(tr.peer.x), 1, 1, stream) == -3
        h64 = C.create_string_buffer(64)
        ptr = C.c_void_p()
        assert L.rm_peer_alloc(0, C.byref(ptr), h64) == -3
        assert L.rm_peer_close(None) == 0 and L.rm_peer_free(None) == 0
    finally:
        tr.close()


CAMERAS = [(0., 0., 0.), (30., -20., 10.), (-45., 15., 0.)]      # the frames of the multi-GPU test alternate between them


def _two_rank_worker(rank, world, port, out_path, name, w, h):
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import rusty_marcher_b200 as rm
    from rusty_marcher_b200 import tiled
    try:
        rm.init(rank)
        scene = workloads.scene(name)
        be = tiled.CudaBackend(scene, rm.create_renderer(1.5, h, w), w, h, dev)
        tr = tiled.TiledRenderer(be, w, h, dev, exchange="peer")
        frames = []
        for i in range(7):                                     # back to back, the camera moving: exercises the two buffers
            tr.set_camera(CAMERAS[i % len(CAMERAS)])
            f = tr.render()
            if rank == 0:
                frames.append(f.clone())                       # (stream-ordered: valid until the next render is issued)
        torch.cuda.synchronize()
        tr.peer.status()
        if rank == 0:
            np.save(out_path, torch.stack(frames).cpu().numpy())
        tr.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,w,h", [("cornell_box", 1920, 1080), ("demo", 800, 600), ("demo", 64, 32)])   # last: ranks without rows
def test_peer_exchange_across_gpus_equals_single_gpu_frame(rm_gpu, tmp_path, name, w, h):
    """world = all visible GPUs (>= 2): the frame assembled on rank 0 by the kernels' peer stores is byte-identical to the
    single-GPU frame.  Skipped on a one-GPU box (the driver's round-end run); run with gpurun --gpus 2."""
    import socket
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    whole = []
    for cam in CAMERAS:
        scene = workloads.scene(name)
        scene.offset_camera(cam)
        whole.append(gpu_render(rm_gpu, scene, w, h, "f32")["rgb8"])
    assert not np.array_equal(whole[0], whole[1])
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "frames.npy")
    mp.spawn(_two_rank_worker, args=(world, port, out, name, w, h), nprocs=world, join=True)
    frames = np.load(out)
    assert len(frames) == 7
    for i, f in enumerate(frames):
        assert np.array_equal(f, whole[i % len(CAMERAS)]), "frame %d differs from the single-GPU frame" % i


def test_headless_cpp_harness_writes_the_golden_ppm(rm_gpu, tmp_path):
    """examples/rm_headless -- the C++ stand-in for `cargo run` + its buttons, over the C ABI only: default scene at the golden
    file's size, "Save to file".  With the f64 kernels the P6 stream is the reference's engine/out.ppm (up to a pow ulp at a
    quantisation boundary); the FP32 production kernels and the f64-rows delivery are within 1 LSB."""
    import json
    import os
    import subprocess
    from rusty_marcher_b200 import build as b
    exe = b.build_examples()
    meta = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "out_ppm.json")))
    ref = O.render(O.Scene.create_default(), 800, 600, want_ids=False, want_fragile=False, want_counters=False)
    want = parity.oracle_rgb8(ref["rgb"]).ravel().astype(np.int16)
    for flags in (["--f64"], [], ["--rows-f64", "--frames", "3"]):
        out = str(tmp_path / ("out_%d.ppm" % len(flags)))
        r = subprocess.run([exe, "--width", "800", "--height", "600", "--out", out] + flags, capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "Rendering using patches of size 32, using 450 patches overall" in r.stdout and "Scene rendered in " in r.stdout
        data = open(out, "rb").read()
        assert len(data) == meta["size"] and data[:15].decode("latin1") == meta["header"]
        got = np.frombuffer(data[15:], dtype=np.uint8).astype(np.int16)
        d = np.abs(got - want)
        assert d.max() <= 1 and (d > 0).mean() < (1e-5 if flags == ["--f64"] else 2e-3), (flags, d.max(), (d > 0).mean())
        if flags == ["--f64"] and d.max() == 0:
            assert hashlib.sha256(data).hexdigest() == meta["sha256"]
    # "Open file": a two-triangle OBJ, moved by (0, 0, -500) like main.rs:278-288
    objp = tmp_path / "quad.obj"
    objp.write_text("o quad\nv -200 -150 0\nv 200 -150 0\nv 200 150 0\nv -200 150 0\nf 1 2 3 4\n")
    out = str(tmp_path / "quad.ppm")
    r = subprocess.run([exe, "--obj", str(objp), "--width", "256", "--height", "160", "--out", out], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "2 primitives resident" in r.stdout, r.stdout + r.stderr
    scene = rm_gpu.Scene.from_obj(str(objp))
    py = gpu_render(rm_gpu, scene, 256, 160, "f32")
    assert np.array_equal(np.frombuffer(open(out, "rb").read()[15:], dtype=np.uint8).reshape(160, 256, 3), py["rgb8"])
    assert (py["prim_id"] >= 0).any()
