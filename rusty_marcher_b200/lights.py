"""Light (engine/src/lights.rs:4-16)."""
from .geometry import Vec3f


class Light:
    def __init__(self, position, color, intensity):
        self.position = Vec3f.of(position)
        self.color = Vec3f.of(color)
        self.intensity = float(intensity)


def create_light(position, color, intensity):
    return Light(position, Vec3f.of(color).normalized_l0(), intensity)   # lights.rs:13
