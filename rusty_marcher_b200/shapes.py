"""Reflectance / Shape (engine/src/shapes.rs:21-61)."""
from dataclasses import dataclass, field, replace

from . import _abi
from .geometry import Vec3f


@dataclass
class Reflectance:
    diffusion: float = 1.
    diffuse_color: Vec3f = field(default_factory=Vec3f.ones)
    specular: float = 1.
    specular_exponent: float = 30.
    is_glass_like: bool = False
    reflection: float = 0.95
    refractive_index: float = 1.

    @staticmethod
    def create_default():
        return Reflectance()

    def copy(self):
        return replace(self, diffuse_color=Vec3f.of(self.diffuse_color))

    def to_c(self):
        r = _abi.RmReflectance()
        r.diffusion = self.diffusion
        r.diffuse_color[:] = list(Vec3f.of(self.diffuse_color))
        r.specular = self.specular
        r.specular_exponent = self.specular_exponent
        r.is_glass_like = int(bool(self.is_glass_like))
        r.reflection = self.reflection
        r.refractive_index = self.refractive_index
        return r

    @staticmethod
    def from_c(r):
        return Reflectance(r.diffusion, Vec3f(*r.diffuse_color), r.specular, r.specular_exponent,
                           bool(r.is_glass_like), r.reflection, r.refractive_index)


class Shape:
    """trait Shape (shapes.rs:40-47).  On the GPU path a shape does not intersect rays itself: it
    flattens into the POD arrays of RmFlatScene (the `Shape::flatten` addition of INTEGRATION.md)."""

    def flatten(self, flat):
        raise NotImplementedError
