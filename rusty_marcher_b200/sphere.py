"""Sphere (engine/src/sphere.rs:6-25)."""
from . import _abi
from .geometry import Vec3f
from .shapes import Reflectance, Shape


class Sphere(Shape):
    def __init__(self, center, radius, reflectance):
        self.center = Vec3f.of(center)
        self.radius_square = float(radius) * float(radius)
        self.reflectance = reflectance.copy()

    def flatten(self, flat):
        s = _abi.RmSphere()
        s.center[:] = list(self.center)
        s.radius_square = self.radius_square
        s.reflectance = self.reflectance.to_c()
        flat.spheres.append(s)
        flat.shapes.append((0, len(flat.spheres) - 1))
        flat.n_prims += 1


def create(center, radius, reflectance=None):
    return Sphere(center, radius, reflectance or Reflectance.create_default())
