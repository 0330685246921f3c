"""Obj + load (engine/src/obj.rs:13-151): triangle meshes from Wavefront files."""
import ctypes as C
import os

import numpy as np

from . import _abi
from .geometry import Vec3f
from .shapes import Shape

REFL_DTYPE = np.dtype(_abi.RmReflectance)


def triangles_from_vertices(v):
    """Triangle::create (triangle.rs:33-47) for an (n, 3, 3) f64 array -> (n, 15) RmTriangle rows:
    9 vertex coordinates, normal = normalized((v1-v0) x (v2-v1)), center = (v0+v1+v2)*(1/3).
    Plain numpy f64 elementwise arithmetic in the reference's evaluation order (no FMA)."""
    v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1, 3, 3)
    a, b, c = v[:, 0], v[:, 1], v[:, 2]
    center = ((a + b) + c) * (1. / 3.)
    e1, e2 = b - a, c - b
    nx = e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1]
    ny = e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2]
    nz = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
    norm = np.sqrt((nx * nx + ny * ny) + nz * nz)
    inv = np.ones_like(norm)
    np.divide(1., norm, out=inv, where=norm > 0.)
    out = np.empty((v.shape[0], 15), dtype=np.float64)
    out[:, 0:9] = v.reshape(-1, 9)
    out[:, 9], out[:, 10], out[:, 11] = nx * inv, ny * inv, nz * inv
    out[:, 12:15] = center
    return out


def gradient_reflectances(n):
    """obj.rs:125-138: default material with diffuse colour (1 - t/n, t/n, 1)."""
    r = np.zeros(n, dtype=REFL_DTYPE)
    t = np.arange(n, dtype=np.float64)
    r["diffusion"] = 1.
    r["specular"] = 1.
    r["specular_exponent"] = 30.
    r["is_glass_like"] = 0
    r["reflection"] = 0.95
    r["refractive_index"] = 1.
    if n:
        r["diffuse_color"][:, 0] = 1. - t / float(n)
        r["diffuse_color"][:, 1] = t / float(n)
        r["diffuse_color"][:, 2] = 1.
    return r


class Obj(Shape):
    def __init__(self, triangles, reflectances=None, name="mesh"):
        self.triangles = np.ascontiguousarray(triangles, dtype=np.float64).reshape(-1, 15)
        n = self.triangles.shape[0]
        self.reflectances = gradient_reflectances(n) if reflectances is None else np.ascontiguousarray(reflectances, dtype=REFL_DTYPE)
        assert self.reflectances.shape[0] == n
        self.name = name

    @staticmethod
    def from_vertices(vertices, name="mesh"):
        return Obj(triangles_from_vertices(vertices), None, name)

    def offset(self, off):                                      # obj.rs:24-29 / triangle.rs:19-24
        off = np.array(list(Vec3f.of(off)), dtype=np.float64)
        self.triangles[:, 12:15] += off                         # centre
        self.triangles[:, 0:9] += np.tile(off, 3)               # vertices; the normal is left alone

    def make_glass(self, reflection=0.2, refractive_index=1.5, diffusion=1.):
        """Extension mode: the reference loads meshes opaque (obj.rs:125-138); this turns every
        triangle's material glass-like, the gradient colour stays."""
        self.reflectances["is_glass_like"] = 1
        self.reflectances["reflection"] = reflection
        self.reflectances["refractive_index"] = refractive_index
        self.reflectances["diffusion"] = diffusion

    def flatten(self, flat):
        flat.objs.append((flat.n_triangles, self.triangles.shape[0]))
        flat.triangle_chunks.append(self.triangles)
        flat.reflectance_chunks.append(self.reflectances)
        flat.n_triangles += self.triangles.shape[0]
        flat.shapes.append((2, len(flat.objs) - 1))
        flat.n_prims += self.triangles.shape[0]


def load(path):
    """obj::load (obj.rs:44-151): Some(vec of Obj, one per model) or None when the file cannot be
    read.  Parsing (the `tobj` behaviour) is done by the C++ host library."""
    L = _abi.load()
    b = L.rm_builder_new()
    try:
        n = L.rm_builder_add_obj_file(b, os.fsencode(path), None)
        if n < 0:
            print("Could not load obj from %s" % path)          # obj.rs:53-56
            return None
        fs = L.rm_builder_flatten(b).contents
        tris = np.ctypeslib.as_array(C.cast(fs.triangles, C.POINTER(C.c_double)), shape=(fs.n_triangles, 15)).copy() \
            if fs.n_triangles else np.zeros((0, 15))
        out = []
        for s in range(fs.n_shapes):
            o = fs.objs[fs.shapes[s].index]
            out.append(Obj(tris[o.first_triangle:o.first_triangle + o.n_triangles].copy(), None,
                           L.rm_builder_shape_name(b, s).decode()))
        return out
    finally:
        L.rm_builder_free(b)
