"""Scene (engine/src/scene.rs:9-211): lights, shapes and the camera position, plus the flattening
of the shape list into the RmFlatScene PODs the C ABI consumes."""
import ctypes as C

import numpy as np

from . import _abi, lights, polygon, sphere
from .geometry import Vec3f
from .obj import REFL_DTYPE, Obj
from .shapes import Reflectance


class _Flat:
    """Accumulates the arrays of an RmFlatScene and keeps them alive."""

    def __init__(self):
        self.shapes, self.sphere_rows, self.polygons, self.polygon_vertices, self.objs = [], [], [], [], []
        self.triangle_chunks, self.reflectance_chunks = [], []
        self.n_triangles = 0
        self.n_prims = 0
        self.lights = []

    def finish(self):
        def arr(ctype, items):
            a = (ctype * max(len(items), 1))()
            for i, it in enumerate(items):
                a[i] = it
            return a

        self._shapes_np = np.ascontiguousarray(self.shapes if self.shapes else [(0, 0)], dtype=np.int32)     # RmShapeRef {kind, index}
        self._shapes = C.cast(self._shapes_np.ctypes.data, C.POINTER(_abi.RmShapeRef))
        # RmSphere = 13 doubles, the 11th an int32 flag (+ padding)
        assert C.sizeof(_abi.RmSphere) == 104 and _abi.RmSphere.reflectance.offset + _abi.RmReflectance.is_glass_like.offset == 80
        rows = np.array(self.sphere_rows if self.sphere_rows else [(0.,) * 13], dtype=np.float64).reshape(-1, 13)
        flags = rows[:, 10] != 0.
        rows[:, 10] = 0.
        rows.view(np.int32)[:, 20] = flags
        self._spheres_np = rows
        self._spheres = (_abi.RmSphere * rows.shape[0]).from_buffer(rows)
        self._polygons = arr(_abi.RmPolygon, self.polygons)
        self._pv = np.ascontiguousarray(self.polygon_vertices if self.polygon_vertices else [0.], dtype=np.float64)
        self._objs = arr(_abi.RmObj, [_abi.RmObj(f, n) for f, n in self.objs])
        # a scene with ONE mesh hands its arrays over as they are (rm_scene_upload copies what it keeps): no 20 MB
        # concatenation for a 10^5-triangle mesh; set_glass_index() copies before it writes
        def joined(chunks, empty):
            if not chunks:
                return empty
            return np.ascontiguousarray(chunks[0] if len(chunks) == 1 else np.concatenate(chunks))
        self._tris = joined(self.triangle_chunks, np.zeros((1, 15)))
        self._refl = joined(self.reflectance_chunks, np.zeros(1, dtype=REFL_DTYPE))
        self._refl_shared = len(self.reflectance_chunks) == 1
        self._lights = arr(_abi.RmLight, self.lights)
        fs = _abi.RmFlatScene()
        fs.n_shapes, fs.shapes = len(self.shapes), self._shapes
        fs.n_spheres, fs.spheres = len(self.sphere_rows), self._spheres
        fs.n_polygons, fs.polygons = len(self.polygons), self._polygons
        fs.n_polygon_vertices = len(self.polygon_vertices) // 3
        fs.polygon_vertices = self._pv.ctypes.data_as(C.POINTER(C.c_double))
        fs.n_objs, fs.objs = len(self.objs), self._objs
        fs.n_triangles = self.n_triangles
        fs.triangles = C.cast(self._tris.ctypes.data, C.POINTER(_abi.RmTriangle))
        fs.triangle_reflectances = C.cast(self._refl.ctypes.data, C.POINTER(_abi.RmReflectance))
        fs.n_lights, fs.lights = len(self.lights), self._lights
        self.c = fs
        return self

    def set_glass_index(self, refractive_index):
        """Extension mode (rm_render_dispersive): every glass-like material of the flattened scene gets this refractive
        index.  Returns the number of materials changed."""
        n = 0
        for i in range(self.c.n_spheres):
            if self._spheres[i].reflectance.is_glass_like:
                self._spheres[i].reflectance.refractive_index = refractive_index
                n += 1
        for i in range(self.c.n_polygons):
            if self._polygons[i].reflectance.is_glass_like:
                self._polygons[i].reflectance.refractive_index = refractive_index
                n += 1
        if self.n_triangles:
            if self._refl_shared:                      # still the mesh's own array
                self._refl, self._refl_shared = self._refl.copy(), False
                self.c.triangle_reflectances = C.cast(self._refl.ctypes.data, C.POINTER(_abi.RmReflectance))
            glass = self._refl["is_glass_like"] != 0
            self._refl["refractive_index"][glass] = refractive_index
            n += int(glass.sum())
        return n


class Scene:
    def __init__(self):                                         # Scene::new, scene.rs:16-23
        self.lights = []
        self.shapes = []
        self.camera = Vec3f.zero()
        self._handle = None
        self._fingerprint = None
        self._flat = None                 # flatten() of the fingerprint below, reused by device_handle()
        self._flat_fingerprint = None

    @staticmethod
    def new():
        return Scene()

    def offset_camera(self, offset):                            # scene.rs:25-27
        self.camera = self.camera + Vec3f.of(offset)

    @staticmethod
    def create_default():
        """Scene::create_default (scene.rs:28-211).  The reference mutates one Reflectance value
        between shapes, so fields carry over; the same mutation sequence is kept here."""
        r = Reflectance.create_default()
        r.diffuse_color = Vec3f(0.8, 0., 0.)
        r.specular_exponent = 100.
        sphere_red = sphere.create(Vec3f(-5., 0., -16.), 4., r)
        r.diffuse_color = Vec3f(0.6, 0., 0.7)
        triangle = polygon.ConvexPolygon.create([Vec3f(7., -4., -8.), Vec3f(15., 0., -9.), Vec3f(6., 3., -8.)], r)
        r.diffusion = 1.0
        r.specular = 1.
        r.is_glass_like = True
        r.refractive_index = 1.5
        r.reflection = 0.5
        r.diffuse_color = Vec3f(0.3, 0.9, 0.9)
        square = polygon.ConvexPolygon.create(
            [Vec3f(20., -3., -50.), Vec3f(-20., -3., -50.), Vec3f(-15., -6., -3.), Vec3f(15., -6., -3.)], r)
        r.specular = 1.0
        r.diffusion = 0.1
        r.diffuse_color = Vec3f(0., 0., 0.2)
        r.is_glass_like = True
        r.refractive_index = 1.5
        r.reflection = 0.2
        sphere_blue = sphere.create(Vec3f(-0.5, -1.5, -5.), 2., r)
        r.diffusion = 1.
        r.reflection = 1.
        r.is_glass_like = False
        r.specular = 0.8
        r.diffuse_color = Vec3f(0., 1., 0.)
        sphere_green = sphere.create(Vec3f(6., -0.5, -18.), 3., r)
        r.diffuse_color = Vec3f(0.9, 0.9, 0.9)
        sphere_white = sphere.create(Vec3f(-10., 6., -14.), 4., r)
        s = Scene()
        s.lights = [lights.create_light(Vec3f(0., 0., 0.), Vec3f.ones(), 1.),
                    lights.create_light(Vec3f(20., 20., 20.), Vec3f(1., 0.5, 0.5), 0.8)]
        s.shapes = [sphere_blue, sphere_green, sphere_red, sphere_white, triangle, square]   # scene.rs:201-208
        return s

    @staticmethod
    def from_obj(path, offset=(0., 0., -500.)):
        """Win::open_obj (main.rs:261-315): load, move every model by `offset`, add the two lights."""
        from . import obj
        objects = obj.load(path)
        s = Scene()
        if objects is not None:
            for o in objects:
                o.offset(offset)
                s.shapes.append(o)
        s.lights.append(lights.create_light(Vec3f(0., 0., 0.), Vec3f.ones(), 1.))
        s.lights.append(lights.create_light(Vec3f(20., 20., 20.), Vec3f(1., 0.5, 0.5), 0.8))
        return s

    # ---- flattening / device residency -------------------------------------------------------
    def flatten(self):
        flat = _Flat()
        for sh in self.shapes:
            sh.flatten(flat)
        for lg in self.lights:
            c = _abi.RmLight()
            c.position[:] = list(lg.position)
            c.color[:] = list(lg.color)
            c.intensity = lg.intensity
            flat.lights.append(c)
        return flat.finish()

    @staticmethod
    def _checksum(a):
        """Content key of an array: the library's 64-bit content hash of its raw bytes (rm_content_hash, memory speed: a mesh
        of 10^5 triangles, 12 MB, costs about 2 ms; zlib's CRC-32 took five times that).  Every bit and every position
        counts -- a float sum misses an int field seen as a denormal, a permutation of the triangles (order sets the colour
        gradient and the tie-breaks, obj.rs:125-138, 198) and any sum-preserving edit."""
        b = np.ascontiguousarray(a)
        return (a.shape, _abi.load().rm_content_hash(b.ctypes.data, b.nbytes, 0))

    def _fp(self):
        def key(s):
            if isinstance(s, Obj):
                return (id(s), s.triangles.shape[0], self._checksum(s.triangles), self._checksum(s.reflectances))
            r = s.reflectance
            cx, cy, cz = r.diffuse_color                          # a Vec3f or any 3-sequence
            rk = (r.diffusion, cx, cy, cz, r.specular, r.specular_exponent, r.is_glass_like, r.reflection, r.refractive_index)
            if isinstance(s, sphere.Sphere):
                return (id(s), s.center.x, s.center.y, s.center.z, s.radius_square, rk)
            return (id(s), tuple((v.x, v.y, v.z) for v in s.vertices), rk)
        return (tuple(key(s) for s in self.shapes),
                tuple((tuple(l.position), tuple(l.color), l.intensity) for l in self.lights))

    def device_handle(self):
        """Uploads the scene (once; again only after it changed) and returns the RmScene handle."""
        L = _abi.load()
        fp = self._fp()
        if self._handle is not None and fp == self._fingerprint:
            return self._handle
        self.release()
        # the flattened host arrays are kept with the fingerprint they were made from: a scene that has to be uploaded
        # again unchanged (after release(), on another device) is not marshalled again
        if self._flat is None or self._flat_fingerprint != fp:
            self._flat, self._flat_fingerprint = self.flatten(), fp
        flat = self._flat
        h = C.c_int64(0)
        _abi.check(L.rm_scene_upload(C.byref(flat.c), C.byref(h)))
        self._handle, self._fingerprint = h.value, fp
        return self._handle

    @property
    def num_prims(self):
        return sum(s.triangles.shape[0] if isinstance(s, Obj) else 1 for s in self.shapes)

    def release(self):
        if self._handle is not None and _abi._lib is not None:
            _abi._lib.rm_scene_free(self._handle)
        self._handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass
