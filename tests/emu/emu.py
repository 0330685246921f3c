"""ctypes binding of the dev/test host emulation of the kernel code (tests/emu/rm_emu.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

from rusty_marcher_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "librm_emu.so")
ROOT = os.path.dirname(os.path.dirname(HERE))
_lib = None


def build(force=False):
    srcs = [os.path.join(HERE, "rm_emu.cpp")] + [os.path.join(ROOT, "rusty_marcher_b200", "csrc", f) for f in
                                                 ("rm_trace.cuh", "rm_fast.cuh", "rm_math.cuh", "rm_scene.cpp", "rm_scene.h", "rm_host.cpp", "rm_bvh.cuh", "rm_bvh.cpp")]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(s) for s in srcs):
        return LIB
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-fPIC", "-shared", "-pthread",
                    "-o", LIB, srcs[0]], check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        for name in ("emu_render_f32", "emu_render_f64", "emu_render_fast"):
            fn = getattr(_lib, name)
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(_abi.RmFlatScene), C.POINTER(_abi.RmParams), C.c_void_p, C.c_void_p,
                           C.POINTER(_abi.RmStats), C.c_int]
    return _lib


def render(scene, width, height, precision="f32", fov=1.5, max_depth=3, cull=True, threads=8, patch_rows=(0, -1), strip_bound=True, accel=False, glass_index=None,
           glass_mode=None):
    """glass_mode (fast path): None = as the device decides per scene and camera (rm::glass_mode), 1 = FP32 ray geometry on
    glass paths, 2 = f64 ray geometry."""
    lib().emu_set_strip_bound(int(strip_bound))
    lib().emu_set_glass_mode(-1 if glass_mode is None else int(glass_mode))
    flat = scene.flatten()
    if glass_index is not None:
        flat.set_glass_index(float(glass_index))     # one pass of the per-channel dispersion (extension mode)
    p = _abi.RmParams()
    p.width, p.height, p.fov = width, height, fov
    p.camera[:] = list(scene.camera)
    p.max_depth, p.background, p.patch_size = max_depth, 0.1, 32
    p.patch_row_begin, p.patch_row_end = patch_rows
    p.cull_backfacing = int(cull)
    p.accel = int(accel)
    dt = np.float32 if precision in ("f32", "fast") else np.float64
    rgb = np.zeros((height, width, 3), dtype=dt)
    ids = np.full((height, width), -1, dtype=np.int32)
    st = _abi.RmStats()
    fn = {"f32": lib().emu_render_f32, "f64": lib().emu_render_f64, "fast": lib().emu_render_fast}[precision]
    rc = fn(C.byref(flat.c), C.byref(p), rgb.ctypes.data, ids.ctypes.data, C.byref(st), threads)
    if rc != 0:
        raise RuntimeError("emu rc %d" % rc)
    return {"rgb": rgb, "prim_id": ids, "counters": st.counters(), "max": st.max_value, "resident": st.resident_prims,
            "glass_mode": lib().emu_last_glass_mode() if precision == "fast" else None}


def bvh(scene):
    """The hierarchy rm_scene.cpp builds for the scene's FP32 pack: (nodes (n, 16) float32, leaf entries int32, depth)."""
    L = lib()
    flat = scene.flatten()
    cap = 4 * (flat.c.n_spheres + flat.c.n_polygons + flat.c.n_triangles) + 16
    nodes = np.zeros((cap, 16), dtype=np.float32)
    prims = np.zeros(cap, dtype=np.int32)
    n_nodes, n_prims, depth = C.c_int(), C.c_int(), C.c_int()
    L.emu_bvh.restype = C.c_int
    L.emu_bvh.argtypes = [C.POINTER(_abi.RmFlatScene), C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                          C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    rc = L.emu_bvh(C.byref(flat.c), nodes.ctypes.data, cap, prims.ctypes.data, cap, C.byref(n_nodes), C.byref(n_prims), C.byref(depth))
    if rc != 0:
        raise RuntimeError("emu_bvh rc %d" % rc)
    return nodes[:n_nodes.value], prims[:n_prims.value], depth.value


def pack_digest(scene):
    """(FP32, FP64) digests of everything the library's packer produces for the scene (blob, records, materials, lists,
    hierarchy)."""
    L = lib()
    flat = scene.flatten()
    out = (C.c_ulonglong * 2)()
    L.emu_pack_digest.restype = C.c_int
    L.emu_pack_digest.argtypes = [C.POINTER(_abi.RmFlatScene), C.POINTER(C.c_ulonglong)]
    rc = L.emu_pack_digest(C.byref(flat.c), out)
    if rc != 0:
        raise RuntimeError("emu_pack_digest rc %d" % rc)
    return int(out[0]), int(out[1])


def walk_stats():
    """(walks, node visits, primitive tests) of the hierarchy walks of the accel renders since the last call."""
    out = (C.c_ulonglong * 3)()
    lib().emu_walk_stats(out)
    return tuple(int(x) for x in out)


def render_costs(scene, width, height, max_depth=3, threads=8):
    """Accel render that also returns the per-pixel walk cost (node visits + 0.7 x primitive tests) of stage A and of
    the shading stage, (H, W, 2) float32 -- the input of tools/tail_model.py."""
    cost = np.zeros((height, width, 2), dtype=np.float32)
    L = lib()
    L.emu_set_cost_buffer.argtypes = [C.c_void_p]
    L.emu_set_cost_buffer(cost.ctypes.data)
    try:
        r = render(scene, width, height, "fast", max_depth=max_depth, accel=True, threads=threads)
    finally:
        L.emu_set_cost_buffer(None)
    r["cost"] = cost
    return r
