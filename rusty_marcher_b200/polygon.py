"""ConvexPolygon (engine/src/polygon.rs:6-52)."""
from . import _abi
from .geometry import Vec3f
from .shapes import Reflectance, Shape


class ConvexPolygon(Shape):
    def __init__(self, vertices, reflectance):
        vs = [Vec3f.of(v) for v in vertices]
        assert len(vs) > 2                                      # polygon.rs:18
        mean = Vec3f.zero()
        for v in vs:                                            # polygon.rs:25-28
            mean = mean + v
        mean = mean.scaled(1. / float(len(vs)))                 # polygon.rs:29
        edge_1 = vs[1] - vs[0]
        edge_2 = vs[2] - vs[1]
        self.vertices = vs
        self.reflectance = reflectance.copy()
        self.plane_normal = edge_1.cross(edge_2).normalized()   # polygon.rs:38
        self.plane_point = mean

    @staticmethod
    def create(vertices, reflectance=None):
        return ConvexPolygon(vertices, reflectance or Reflectance.create_default())

    def offset(self, off):                                      # polygon.rs:44-49
        off = Vec3f.of(off)
        self.plane_point = self.plane_point + off
        self.vertices = [v + off for v in self.vertices]

    def flatten(self, flat):
        p = _abi.RmPolygon()
        p.first_vertex = len(flat.polygon_vertices) // 3
        p.n_vertices = len(self.vertices)
        p.plane_normal[:] = list(self.plane_normal)
        p.plane_point[:] = list(self.plane_point)
        p.reflectance = self.reflectance.to_c()
        for v in self.vertices:
            flat.polygon_vertices.extend(v)
        flat.polygons.append(p)
        flat.shapes.append((1, len(flat.polygons) - 1))
        flat.n_prims += 1
