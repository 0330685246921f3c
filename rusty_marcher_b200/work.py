"""Algorithmic work of the hot path from its event counters.

Each counted event carries (flops, minimum FP32 issue slots with FMA fusion) of the REFERENCE
arithmetic it stands for; the roofline numerator is sum(count * weight), independent of how the
kernels actually evaluate the predicate."""

WEIGHTS = {
    "pixels": (22, 14),          # primary ray generation, renderer.rs:128-135
    "sphere_tests": (15, 10),    # sphere.rs:28-38
    "sphere_disc": (4, 5),       # sphere.rs:40-51
    "sphere_hits": (19, 13),     # sphere.rs:54-60
    "plane_tests": (5, 3),       # triangle.rs:56 / polygon.rs:65
    "plane_dist": (9, 8),        # triangle.rs:62
    "plane_point": (6, 3),       # triangle.rs:69
    "edge_tests": (7, 6),        # triangle.rs:13-15
    "cand_dist": (8, 6),         # shapes.rs:128, obj.rs:197
    "hits": (6, 6),              # renderer.rs:272,192
    "light_evals": (24, 16),     # renderer.rs:166-172
    "lit_lights": (57, 40),      # renderer.rs:180-189
    "glass_hits": (40, 30),      # optics.rs:15-35,56-76
    "reflections": (38, 24),     # optics.rs:38-47
    "refractions": (58, 38),     # optics.rs:78-88
}


def algorithmic_work(counters):
    """(flops, issue slots) of one frame from its event counters."""
    flops = sum(counters[k] * w[0] for k, w in WEIGHTS.items())
    slots = sum(counters[k] * w[1] for k, w in WEIGHTS.items())
    return flops, slots


def segments(counters):
    """Ray segments = scene queries: find_closest_intersect + intersect_shape_set calls."""
    return counters["closest_segments"] + counters["anyhit_segments"]
