"""The scene packer (csrc/rm_scene.cpp, rm_bvh.cpp): what it writes -- hot blob, triangle records, f64 refinement sources,
materials, scene-order lists, hierarchy -- is pinned by digest against the values its first, single-threaded, one-allocation-
per-primitive version produced (taken before it was rewritten for speed: views instead of copies, counting sort by class,
records and hierarchy sweeps on the host pool), and does not depend on the number of threads it runs on.  Through the
emulation library, which compiles the same sources."""
import ctypes as C
import os

import numpy as np
import pytest

from rusty_marcher_b200 import _abi, workloads
from tests.emu import emu

# (workload, kwargs) -> (FP32 digest, FP64 digest), produced by the round-1 packer (commit cb593c4) on these scenes
PINNED = [
    ("demo", {}, 0xbe5a6e7552ff1391, 0x9c6380af77e1e9ef),
    ("cornell_box", {}, 0xee23335113d147e2, 0xd7057b8cadda8a7a),
    ("dodecahedron", {}, 0x8245a1f1149e0ad6, 0x0412ee46b6bb80fe),                       # degenerate-projection class
    ("ngons", {}, 0x65dcd287fefd5d68, 0x6e158843636b049b),                              # n-gons, back-facing class
    ("ngons", dict(n_spheres=500, grid=0), 0x5e69d8b482e109d6, 0x259b42ddda97b723),
    ("stress", dict(grid=40, n_spheres=64), 0x1b0fe30209311504, 0x47bde3a69d3f48b3),
    ("stress", dict(grid=96, n_spheres=300), 0x261e43a54e886bf4, 0x6e8d9f8f87dcda3c),   # 18.7k primitives: the threaded path
]


@pytest.mark.parametrize("name,kw,want32,want64", PINNED, ids=["%s%s" % (p[0], "-".join("%s%d" % kv for kv in p[1].items())) for p in PINNED])
def test_packed_scene_is_what_the_first_packer_produced(name, kw, want32, want64):
    got32, got64 = emu.pack_digest(workloads.scene(name, **kw))
    assert (got32, got64) == (want32, want64)


def test_packed_scene_does_not_depend_on_the_number_of_threads(monkeypatch):
    scene = workloads.scene("stress", grid=96, n_spheres=300)
    digests = set()
    for threads in ("1", "2", "3", "7"):
        monkeypatch.setenv("RM_B200_HOST_THREADS", threads)       # read by the packer at every call (rm_pool.h)
        digests.add(emu.pack_digest(scene))
    assert len(digests) == 1 and digests.pop() == (0x261e43a54e886bf4, 0x6e8d9f8f87dcda3c)


def test_content_hash_sees_every_byte_and_position():
    L = _abi.load()
    a = np.arange(100_003, dtype=np.float64)

    def h(x, seed=0):
        x = np.ascontiguousarray(x)
        return L.rm_content_hash(x.ctypes.data, x.nbytes, seed)
    base = h(a)
    assert base == h(a.copy()) and base != h(a, seed=1)
    for i in (0, 1, 3, 4, 50_000, 100_002):                      # every lane of a 32-byte block, first and last element
        b = a.copy()
        b.view(np.uint64)[i] ^= 1                                 # one bit
        assert h(b) != base
    b = a.copy()
    b[[10, 20]] = b[[20, 10]]                                     # a permutation keeps every sum
    assert h(b) != base
    assert h(a[:-1]) != base
    assert L.rm_content_hash(None, 0, 0) == L.rm_content_hash(None, 0, 0)
    assert h(a.view(np.uint8)[:-3]) != h(a.view(np.uint8)[:-2])  # ragged tails
