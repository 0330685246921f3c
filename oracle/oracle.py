"""ctypes binding of the CPU oracle (oracle/rm_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under rusty_marcher_b200/ imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "librm_oracle.so")


class Reflectance(C.Structure):
    """engine/src/shapes.rs:21-32"""
    _fields_ = [
        ("diffusion", C.c_double),
        ("diffuse_color", C.c_double * 3),
        ("specular", C.c_double),
        ("specular_exponent", C.c_double),
        ("is_glass_like", C.c_int32),
        ("reflection", C.c_double),
        ("refractive_index", C.c_double),
    ]


COUNTER_FIELDS = [
    "pixels", "closest_segments", "anyhit_segments", "sphere_tests", "sphere_disc", "sphere_hits",
    "plane_tests", "plane_dist", "plane_point", "edge_tests", "cand_dist", "hits", "light_evals",
    "lit_lights", "glass_hits", "reflections", "refractions",
]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in COUNTER_FIELDS]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n in COUNTER_FIELDS}


def build(force=False):
    """Compile the oracle with its Makefile (g++ -O2 -ffp-contract=off)."""
    src = os.path.join(_HERE, "rm_oracle.cpp")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "rm_oracle.h")))):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    d3 = C.POINTER(C.c_double)
    L.orc_scene_new.restype = C.c_void_p
    L.orc_scene_create_default.restype = C.c_void_p
    L.orc_scene_free.argtypes = [C.c_void_p]
    L.orc_scene_set_camera.argtypes = [C.c_void_p, d3]
    L.orc_scene_offset_camera.argtypes = [C.c_void_p, d3]
    L.orc_reflectance_default.argtypes = [C.POINTER(Reflectance)]
    L.orc_scene_add_sphere.argtypes = [C.c_void_p, d3, C.c_double, C.POINTER(Reflectance)]
    L.orc_scene_add_polygon.argtypes = [C.c_void_p, d3, C.c_int, C.POINTER(Reflectance)]
    L.orc_scene_add_obj_file.argtypes = [C.c_void_p, C.c_char_p, d3]
    L.orc_scene_add_mesh.argtypes = [C.c_void_p, d3, C.c_int, d3]
    L.orc_scene_add_light.argtypes = [C.c_void_p, d3, d3, C.c_double]
    L.orc_scene_make_glass.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
    L.orc_scene_set_glass_index.argtypes = [C.c_void_p, C.c_double]
    L.orc_scene_num_shapes.argtypes = [C.c_void_p]
    L.orc_scene_num_prims.argtypes = [C.c_void_p]
    L.orc_scene_obj_triangles.argtypes = [C.c_void_p, C.c_int, d3, C.c_int]
    L.orc_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                             d3, C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.POINTER(Counters)]
    L.orc_normalize.argtypes = [d3, C.c_int, C.c_int]
    L.orc_normalize.restype = C.c_double
    L.orc_to_vec.argtypes = [d3, C.c_int, C.c_int, C.POINTER(C.c_uint8)]
    L.orc_write_ppm.argtypes = [C.c_char_p, d3, C.c_int, C.c_int]
    L.orc_vec_normalized.argtypes = [d3, d3]
    L.orc_vec_normalized_l0.argtypes = [d3, d3]
    L.orc_vec_dot.argtypes = [d3, d3]
    L.orc_vec_dot.restype = C.c_double
    L.orc_vec_cross.argtypes = [d3, d3, d3]
    L.orc_vec_scaled.argtypes = [d3, C.c_double, d3]
    L.orc_reflect.argtypes = [d3, d3, d3]
    L.orc_reflect_ray.argtypes = [d3, d3, d3, C.c_double, d3, d3]
    L.orc_refract_ray.argtypes = [d3, d3, d3, C.c_double, d3, d3]
    L.orc_triangle_intersect.argtypes = [d3, d3, d3, d3, d3]
    L.orc_last_degenerate_hits.restype = C.c_uint64
    L.orc_sphere_intersect.argtypes = [d3, C.c_double, d3, d3, d3, d3]
    _lib = L
    return L


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def default_reflectance():
    r = Reflectance()
    lib().orc_reflectance_default(C.byref(r))
    return r


def make_reflectance(diffusion=1., diffuse_color=(1., 1., 1.), specular=1., specular_exponent=30.,
                     is_glass_like=False, reflection=0.95, refractive_index=1.):
    r = Reflectance()
    r.diffusion = diffusion
    r.diffuse_color[:] = [float(c) for c in diffuse_color]
    r.specular = specular
    r.specular_exponent = specular_exponent
    r.is_glass_like = int(bool(is_glass_like))
    r.reflection = reflection
    r.refractive_index = refractive_index
    return r


class Scene:
    """engine/src/scene.rs:9-27 on the oracle side."""

    def __init__(self, handle=None):
        self._h = C.c_void_p(handle if handle is not None else lib().orc_scene_new())

    @classmethod
    def create_default(cls):
        return cls(lib().orc_scene_create_default())

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.orc_scene_free(self._h)
            self._h = None

    def set_camera(self, xyz):
        lib().orc_scene_set_camera(self._h, _d3(xyz))

    def offset_camera(self, xyz):
        lib().orc_scene_offset_camera(self._h, _d3(xyz))

    def add_sphere(self, center, radius, reflectance=None):
        r = reflectance if reflectance is not None else default_reflectance()
        return lib().orc_scene_add_sphere(self._h, _d3(center), float(radius), C.byref(r))

    def add_polygon(self, vertices, reflectance=None):
        r = reflectance if reflectance is not None else default_reflectance()
        v = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
        return lib().orc_scene_add_polygon(self._h, _dptr(v), v.shape[0], C.byref(r))

    def add_obj_file(self, path, offset=(0., 0., -500.)):
        n = lib().orc_scene_add_obj_file(self._h, os.fsencode(path), _d3(offset))
        if n < 0:
            raise IOError("oracle: could not load obj from %s" % path)
        return n

    def add_mesh(self, tri_verts, offset=(0., 0., 0.)):
        v = np.ascontiguousarray(tri_verts, dtype=np.float64).reshape(-1, 9)
        return lib().orc_scene_add_mesh(self._h, _dptr(v), v.shape[0], _d3(offset))

    def add_light(self, position, color, intensity):
        lib().orc_scene_add_light(self._h, _d3(position), _d3(color), float(intensity))

    # ---- extension mode (no reference counterpart) ----
    def make_glass(self, shape, reflection=0.2, refractive_index=1.5, diffusion=1.):
        """Every primitive of shape `shape` becomes glass-like (an OBJ mesh is opaque in the reference, obj.rs:125-138)."""
        n = lib().orc_scene_make_glass(self._h, int(shape), float(reflection), float(refractive_index), float(diffusion))
        if n < 0:
            raise IndexError("oracle: no shape %d" % shape)
        return n

    def set_glass_index(self, refractive_index):
        """Every glass-like material gets this refractive index (one pass of the per-channel dispersion)."""
        return lib().orc_scene_set_glass_index(self._h, float(refractive_index))

    def add_default_lights(self):
        """engine/src/main.rs:293-315 (the same two lights as scene.rs:178-198)."""
        self.add_light((0., 0., 0.), (1., 1., 1.), 1.)
        self.add_light((20., 20., 20.), (1., .5, .5), .8)

    @property
    def num_shapes(self):
        return lib().orc_scene_num_shapes(self._h)

    @property
    def num_prims(self):
        return lib().orc_scene_num_prims(self._h)

    def obj_triangles(self, shape):
        n = lib().orc_scene_obj_triangles(self._h, shape, None, 0)
        out = np.zeros((n, 3, 3), dtype=np.float64)
        if n:
            lib().orc_scene_obj_triangles(self._h, shape, _dptr(out), n)
        return out


def hardware_threads():
    return lib().orc_hardware_threads()


def render(scene, width, height, fov=1.5, max_depth=3, threads=None, patch_rows=(0, -1), patch_stride=1,
           want_ids=True, want_fragile=True, want_counters=True, out=None):
    """Renderer::render (engine/src/renderer.rs:36-126).  Returns dict(rgb, prim_id, fragile, counters)."""
    if threads is None:
        threads = hardware_threads()
    rgb = out if out is not None else np.zeros((height, width, 3), dtype=np.float64)
    ids = np.full((height, width), -1, dtype=np.int32) if want_ids else None
    frag = np.zeros((height, width), dtype=np.uint8) if want_fragile else None
    cnt = Counters() if want_counters else None
    rc = lib().orc_render(scene._h, width, height, float(fov), int(max_depth), int(threads),
                          int(patch_rows[0]), int(patch_rows[1]), int(patch_stride), _dptr(rgb),
                          ids.ctypes.data_as(C.POINTER(C.c_int32)) if want_ids else None,
                          frag.ctypes.data_as(C.POINTER(C.c_uint8)) if want_fragile else None,
                          C.byref(cnt) if want_counters else None)
    if rc != 0:
        raise ValueError("oracle: width must be a positive multiple of 32 (renderer.rs:107)")
    return {"rgb": rgb, "prim_id": ids, "fragile": frag, "counters": cnt.as_dict() if want_counters else None,
            "degenerate_hits": int(lib().orc_last_degenerate_hits())}


def render_dispersive(scene, width, height, indices=(1.50, 1.52, 1.54), **kw):
    """EXTENSION MODE (BASELINE.json configs[3]): per-channel refractive indices as three passes of
    the unchanged restatement, pass c with every glass-like material at indices[c]; channel c of the frame is channel c
    of pass c.  Primary ids are those of the first pass (the primary ray does not depend on the index), the fragile mask
    is the union, the counters are summed.  No reference behaviour to compare with -- the oracle and the CUDA path are
    extended identically.  The scene is left with indices[2]."""
    out = None
    for c, n in enumerate(indices):
        scene.set_glass_index(n)
        r = render(scene, width, height, **kw)
        if out is None:
            out = {"rgb": np.zeros_like(r["rgb"]), "prim_id": r["prim_id"], "fragile": r["fragile"],
                   "counters": dict(r["counters"]) if r["counters"] else None, "degenerate_hits": r["degenerate_hits"]}
        else:
            if out["fragile"] is not None:
                out["fragile"] = out["fragile"] | r["fragile"]
            if out["counters"] is not None:
                for k, v in r["counters"].items():
                    out["counters"][k] += v
        out["rgb"][..., c] = r["rgb"][..., c]
    return out


def normalize(rgb):
    """FrameBuffer::normalize (framebuffer.rs:58-77), in place.  Returns the max."""
    h, w, _ = rgb.shape
    return lib().orc_normalize(_dptr(rgb), w, h)


def to_vec(rgb):
    """FrameBuffer::to_vec (framebuffer.rs:40-55)."""
    h, w, _ = rgb.shape
    out = np.zeros((h, w, 3), dtype=np.uint8)
    lib().orc_to_vec(_dptr(rgb), w, h, out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


def ppm_bytes(rgb):
    """FrameBuffer::write_ppm byte stream (framebuffer.rs:26-38)."""
    h, w, _ = rgb.shape
    return b"P6\n%d %d\n255\n" % (w, h) + to_vec(rgb).tobytes()


# weights (flops, min FP32 issue slots) of each counted event.
WEIGHTS = {
    "pixels": (22, 14), "sphere_tests": (15, 10), "sphere_disc": (4, 5), "sphere_hits": (19, 13),
    "plane_tests": (5, 3), "plane_dist": (9, 8), "plane_point": (6, 3), "edge_tests": (7, 6),
    "cand_dist": (8, 6), "hits": (6, 6), "light_evals": (24, 16), "lit_lights": (57, 40),
    "glass_hits": (40, 30), "reflections": (38, 24), "refractions": (58, 38),
}


def algorithmic_work(counters):
    flops = sum(counters[k] * w[0] for k, w in WEIGHTS.items())
    slots = sum(counters[k] * w[1] for k, w in WEIGHTS.items())
    return flops, slots
