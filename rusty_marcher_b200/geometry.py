"""Vec3f (engine/src/geometry.rs:4-182) -- host-side f64 3-vector used to describe scenes."""
import math
from dataclasses import dataclass


@dataclass
class Vec3f:
    x: float = 0.0
    y: float = 0.0
    z: float = 0.0

    @staticmethod
    def zero():
        return Vec3f(0., 0., 0.)

    @staticmethod
    def ones():
        return Vec3f(1., 1., 1.)

    @staticmethod
    def of(v):
        if isinstance(v, Vec3f):
            return Vec3f(v.x, v.y, v.z)
        x, y, z = v
        return Vec3f(float(x), float(y), float(z))

    def __iter__(self):
        return iter((self.x, self.y, self.z))

    def __add__(self, o):
        return Vec3f(self.x + o.x, self.y + o.y, self.z + o.z)

    def __sub__(self, o):
        return Vec3f(self.x - o.x, self.y - o.y, self.z - o.z)

    def __mul__(self, o):
        return Vec3f(self.x * o.x, self.y * o.y, self.z * o.z)

    def __neg__(self):
        return Vec3f(-self.x, -self.y, -self.z)

    def scaled(self, s):
        return Vec3f(self.x * s, self.y * s, self.z * s)

    def dot(self, o):
        return self.x * o.x + self.y * o.y + self.z * o.z

    def cross(self, o):
        return Vec3f(self.y * o.z - self.z * o.y, self.z * o.x - self.x * o.z, self.x * o.y - self.y * o.x)

    def squared_norm(self):
        return self.dot(self)

    def normalized(self):
        n = math.sqrt(self.dot(self))
        return self.scaled(1. / n) if n > 0. else Vec3f.of(self)

    def normalized_l0(self):
        n = max(max(self.x, self.y), self.z)
        return self.scaled(1. / n) if n > 0. else Vec3f.of(self)

    def max(self):
        return max(max(self.x, self.y), self.z)
