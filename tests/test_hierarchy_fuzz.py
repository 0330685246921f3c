"""Randomised scenes for the bounding-volume hierarchy (RmParams.accel, csrc/rm_bvh.cuh) in the host emulation of the
kernel code: spheres, OBJ triangle soups and convex n-gons of any orientation, 2 or 3 lights, culling on and off, scene
scales from 0.1 to 10^4 and cameras up to 10^3 scene sizes away from the origin -- the walk must test a superset of the
primitives every ray can hit, i.e. reproduce the brute-force FP32 frame bit for bit.  (The margins that make the FP32
box test conservative are relative to the scene's coordinates; this is the test that they hold far from the origin.)"""
import numpy as np
import pytest

from rusty_marcher_b200 import lights, polygon, sphere
from rusty_marcher_b200.geometry import Vec3f
from rusty_marcher_b200.obj import Obj
from rusty_marcher_b200.scene import Scene
from rusty_marcher_b200.shapes import Reflectance
from tests.emu import emu


def random_scene(rng):
    scale = float(10 ** rng.uniform(-1, 4))
    cam = (rng.random(3) - 0.5) * scale * float(10 ** rng.uniform(-2, 3))
    dist = scale * float(10 ** rng.uniform(-0.2, 1.0))
    centre = cam + np.array([0., 0., -dist])                     # the camera looks down -z (renderer.rs:128-135)
    s = Scene()

    def refl():
        glass = rng.random() < 0.3
        return Reflectance(0.1 if glass else 1.0, Vec3f(*rng.random(3)), 1.0, 30. if rng.random() < .5 else 100., glass,
                           0.2 + 0.3 * rng.random(), 1.5 if glass else 1.0)

    for _ in range(int(rng.integers(0, 120))):
        c = (rng.random(3) - 0.5) * scale + centre
        s.shapes.append(sphere.create(Vec3f(*c), float(scale * (0.01 + 0.12 * rng.random())), refl()))
    n_tri = int(rng.integers(0, 400))
    if n_tri:
        c = (rng.random((n_tri, 1, 3)) - 0.5) * scale + centre
        tris = c + (rng.random((n_tri, 3, 3)) - 0.5) * scale * 0.35 * rng.random((n_tri, 1, 1))
        s.shapes.append(Obj.from_vertices(tris.astype(np.float32).astype(np.float64), "soup"))     # obj.rs:102-106: f32 positions
    for _ in range(int(rng.integers(0, 10))):
        c = (rng.random(3) - 0.5) * scale + centre
        ang = np.sort(rng.random(int(rng.integers(4, 7))) * 2 * np.pi)
        r, (a, b) = scale * 0.15, rng.random(2) - 0.5
        s.shapes.append(polygon.ConvexPolygon.create(
            [Vec3f(c[0] + r * np.cos(t), c[1] + r * np.sin(t), c[2] + a * r * np.cos(t) + b * r * np.sin(t)) for t in ang], refl()))
    s.lights.append(lights.create_light(tuple(centre + np.array([0., 0., dist])), (1., 1., 1.), 1.))
    s.lights.append(lights.create_light(tuple(centre + 0.6 * scale), (1., .5, .5), .8))
    if rng.random() < 0.5:
        s.lights.append(lights.create_light(tuple(centre + np.array([-0.7, 0.3, 0.2]) * scale), (.5, .5, 1.), .5))
    s.offset_camera(tuple(cam))
    return s


@pytest.mark.parametrize("seed", range(16))
def test_hierarchy_equals_brute_force_on_random_scenes(seed):
    rng = np.random.default_rng(1000 + seed)
    scene = random_scene(rng)
    cull = bool(rng.random() < 0.7)
    depth = int(rng.integers(1, 6))
    a = emu.render(scene, 128, 96, "fast", max_depth=depth, cull=cull)
    b = emu.render(scene, 128, 96, "fast", max_depth=depth, cull=cull, accel=True)
    assert np.array_equal(a["prim_id"], b["prim_id"])
    assert np.array_equal(a["rgb"], b["rgb"], equal_nan=True)
