"""Instantiates rusty_marcher_b200.workloads descriptions on the ORACLE (test helper)."""
from oracle import oracle as O


def build_oracle_scene(desc):
    if desc["default"]:
        return O.Scene.create_default()
    s = O.Scene()
    for c, radius, r in desc["spheres"]:
        s.add_sphere(c, radius, O.make_reflectance(**r))
    for verts, r in desc.get("polygons", ()):
        s.add_polygon(verts, O.make_reflectance(**r))
    for _name, verts, offset in desc["meshes"]:
        s.add_mesh(verts, offset)
    for pos, col, inten in desc["lights"]:
        s.add_light(pos, col, inten)
    return s
