"""The benchmark / parity workloads of BASELINE.json, as plain descriptions that can
be instantiated on the GPU path (build_scene) and, in tests, on the oracle.

  demo          Scene::create_default (engine/src/scene.rs:28-211)
  cornell_box   test_data/cornell_box.obj, every model moved by (0,0,-500), default lights (main.rs:261-315)
  dodecahedron  test_data/dodecahedron.obj, same treatment
  stress        synthetic: 4096 random spheres + a 224x224 quad grid (100,352 triangles), seed 0x5EED
  ngons         synthetic: convex n-gons (n = 4..8, engine/src/polygon.rs) of random tilt + spheres, seed 0x90A5 -- the
                primitives the demo scene has only two of; exercises the n-gon routines of the production kernel

The OBJ meshes are shipped as parsed vertex arrays (scenes/*.npz, made from the reference's
test_data by tests/golden/make_fixtures.py) because the reference tree is not present on the GPU box.
"""
import os

import numpy as np

from . import lights, polygon, sphere
from .geometry import Vec3f
from .obj import Obj
from .scene import Scene
from .shapes import Reflectance

SCENES_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scenes")
OBJ_OFFSET = (0., 0., -500.)                                   # main.rs:278-282
DEFAULT_LIGHTS = [((0., 0., 0.), (1., 1., 1.), 1.), ((20., 20., 20.), (1., .5, .5), .8)]   # main.rs:293-315


def load_models(name):
    """[(model name, (n,3,3) f64 vertices widened from the f32 the OBJ reader produced)]"""
    z = np.load(os.path.join(SCENES_DIR, name + ".npz"))
    names = [str(s) for s in z["names"]]
    return [(n, z["model_%d" % i].astype(np.float64)) for i, n in enumerate(names)]


def _splitmix64(seed):
    state = seed & 0xFFFFFFFFFFFFFFFF
    mask = 0xFFFFFFFFFFFFFFFF

    def nxt():
        nonlocal state
        state = (state + 0x9E3779B97F4A7C15) & mask
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & mask
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & mask
        z = z ^ (z >> 31)
        return (z >> 11) * (1.0 / 9007199254740992.0)     # [0,1) with 53 bits
    return nxt


def describe(name, n_spheres=4096, grid=224, seed=0x5EED, n_polygons=320):
    """A workload as plain data: {'default': bool, 'spheres': [...], 'polygons': [...], 'meshes': [...], 'lights': [...]}
    (shape order of the scene: spheres, polygons, meshes)"""
    if name == "demo":
        return {"name": name, "default": True}
    if name in ("cornell_box", "dodecahedron"):
        return {"name": name, "default": False, "spheres": [],
                "meshes": [(n, v, OBJ_OFFSET) for n, v in load_models(name)], "lights": DEFAULT_LIGHTS}
    if name == "stress":
        u = _splitmix64(seed)
        spheres = []
        for _ in range(n_spheres):
            c = (-60. + 120. * u(), -35. + 70. * u(), -160. + 130. * u())
            radius = 0.3 + 0.9 * u()
            glass = u() < 0.25
            if glass:
                r = dict(diffusion=0.1, diffuse_color=(u(), u(), u()), specular=1., specular_exponent=30. if u() < .5 else 100.,
                         is_glass_like=True, reflection=0.2 + 0.3 * u(), refractive_index=1.5)
            else:
                r = dict(diffusion=1., diffuse_color=(u(), u(), u()), specular=1., specular_exponent=30. if u() < .5 else 100.,
                         is_glass_like=False, reflection=0.95, refractive_index=1.)
            spheres.append((c, radius, r))
        xs = np.linspace(-80., 80., grid + 1)
        ys = np.linspace(-45., 45., grid + 1)
        X, Y = np.meshgrid(xs, ys, indexing="xy")
        Z = -170. + 4. * np.sin(X / 7.) * np.cos(Y / 5.)
        P = np.stack([X, Y, Z], axis=-1)                     # (grid+1, grid+1, 3), [iy, ix]
        a, b, c, d = P[:-1, :-1], P[:-1, 1:], P[1:, 1:], P[1:, :-1]   # CCW in XY: (x0,y0) (x1,y0) (x1,y1) (x0,y1)
        tris = np.concatenate([np.stack([a, b, c], axis=2), np.stack([a, c, d], axis=2)], axis=2)   # (g,g,6,3)
        tris = tris.reshape(grid, grid, 2, 3, 3).reshape(-1, 3, 3)
        # widen through f32 like a tobj-loaded mesh (obj.rs:102-106)
        tris = tris.astype(np.float32).astype(np.float64)
        return {"name": name, "default": False, "spheres": spheres, "meshes": [("grid", tris, (0., 0., 0.))],
                "lights": DEFAULT_LIGHTS}
    if name == "ngons":
        u = _splitmix64(0x90A5 if seed == 0x5EED else seed)

        def refl():
            glass = u() < 0.2
            return dict(diffusion=0.1 if glass else 1., diffuse_color=(u(), u(), u()), specular=1.,
                        specular_exponent=30. if u() < .5 else 100., is_glass_like=glass,
                        reflection=0.2 + 0.3 * u() if glass else 0.95, refractive_index=1.5 if glass else 1.)
        polygons = []
        for k in range(n_polygons):
            n = 4 + k % 5                                          # 4, 5, 6, 7, 8 vertices in turn
            z = -15. - 45. * u()
            cx, cy = (2. * u() - 1.) * 1.2 * -z, (2. * u() - 1.) * 0.7 * -z
            rad, a, b = 0.4 + 2.6 * u(), 1.6 * u() - 0.8, 1.6 * u() - 0.8
            ccw = u() < 0.9                                        # one in ten wound clockwise in XY: never hittable (polygon.rs:83-91)
            ts = [2. * np.pi * (i + 0.6 * u()) / n for i in range(n)]
            if not ccw:
                ts = ts[::-1]
            verts = [(cx + rad * np.cos(t), cy + rad * np.sin(t), z + rad * (a * np.cos(t) + b * np.sin(t))) for t in ts]
            polygons.append((verts, refl()))
        spheres = []
        for _ in range(n_spheres if n_spheres != 4096 else 48):
            z = -12. - 50. * u()
            c = ((2. * u() - 1.) * 1.1 * -z, (2. * u() - 1.) * 0.65 * -z, z)
            spheres.append((c, 0.3 + 1.7 * u(), refl()))
        return {"name": name, "default": False, "spheres": spheres, "polygons": polygons, "meshes": [], "lights": DEFAULT_LIGHTS}
    raise KeyError(name)


def n_prims(desc):
    if desc.get("default"):
        return 6
    return len(desc["spheres"]) + len(desc.get("polygons", ())) + sum(len(v) for _n, v, _o in desc["meshes"])


def build_scene(desc):
    """Instantiates a description with the engine-mirroring classes of this package."""
    if desc["default"]:
        return Scene.create_default()
    s = Scene()
    for c, radius, r in desc["spheres"]:
        refl = Reflectance(r["diffusion"], Vec3f(*r["diffuse_color"]), r["specular"], r["specular_exponent"],
                           r["is_glass_like"], r["reflection"], r["refractive_index"])
        s.shapes.append(sphere.create(Vec3f(*c), radius, refl))
    for verts, r in desc.get("polygons", ()):
        refl = Reflectance(r["diffusion"], Vec3f(*r["diffuse_color"]), r["specular"], r["specular_exponent"],
                           r["is_glass_like"], r["reflection"], r["refractive_index"])
        s.shapes.append(polygon.ConvexPolygon.create([Vec3f(*v) for v in verts], refl))
    for name, verts, offset in desc["meshes"]:
        o = Obj.from_vertices(verts, name)
        o.offset(offset)
        s.shapes.append(o)
    for pos, col, inten in desc["lights"]:
        s.lights.append(lights.create_light(pos, col, inten))
    return s


def scene(name, **kw):
    return build_scene(describe(name, **kw))
