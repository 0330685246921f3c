"""Sphere (engine/src/sphere.rs:6-25)."""
from .geometry import Vec3f
from .shapes import Reflectance, Shape


class Sphere(Shape):
    def __init__(self, center, radius, reflectance):
        self.center = Vec3f.of(center)
        self.radius_square = float(radius) * float(radius)
        self.reflectance = reflectance.copy()

    def flatten(self, flat):
        # one row of 13 doubles = the RmSphere POD (center, radius_square, reflectance; the int flag is patched in by
        # finish()): a scene of thousands of spheres is marshalled as ONE array, not one ctypes object per sphere
        c, r = self.center, self.reflectance
        dr, dg, db = r.diffuse_color
        flat.sphere_rows.append((c.x, c.y, c.z, self.radius_square, r.diffusion, dr, dg, db, r.specular, r.specular_exponent,
                                 1. if r.is_glass_like else 0., r.reflection, r.refractive_index))
        flat.shapes.append((0, len(flat.sphere_rows) - 1))
        flat.n_prims += 1


def create(center, radius, reflectance=None):
    return Sphere(center, radius, reflectance or Reflectance.create_default())
