"""GPU part of the extension-mode tests (see tests/test_dispersive.py): rm_render_dispersive through the Python mirror
against the oracle, and against three plain renders of the same kernels merged on the host (bit for bit)."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O
from rusty_marcher_b200 import _abi
from rusty_marcher_b200.obj import Obj
from tests import parity
from tests.test_dispersive import INDICES, demo, glass_dodecahedron

pytestmark = pytest.mark.gpu


def set_index(scene, n):
    for s in scene.shapes:
        if isinstance(s, Obj):
            s.reflectances["refractive_index"][s.reflectances["is_glass_like"] != 0] = n
        elif s.reflectance.is_glass_like:
            s.reflectance.refractive_index = n


@pytest.mark.parametrize("make,w,h", [(demo, 800, 600), (glass_dodecahedron, 640, 480)])
def test_dispersive_frame(rm_gpu, make, w, h):
    rm = rm_gpu
    scene, osc = make()
    ref = O.render_dispersive(osc, w, h, INDICES)
    r = rm.create_renderer(1.5, h, w)
    fb = rm.create_frame_buffer(w, h, dtype=np.float32)
    ids = np.full((h, w), -1, dtype=np.int32)
    msg = r.render_dispersive(fb, scene, INDICES, prim_id=ids)
    assert msg.startswith("Scene rendered in ") and r.last_stats.kernel_launches >= 6
    rows = (h // 32) * 32
    m = np.float32(fb.buffer.max())
    rgb8 = (np.float32(255) * np.clip(fb.buffer * (np.float32(1) / m), 0, 1)).astype(np.uint8)
    parity.check_fp32({"rgb": fb.buffer, "prim_id": ids}, ref, rows, rgb8)
    assert np.all(fb.buffer[rows:] == 0)
    # the same three passes as plain renders, merged here: identical bits
    merged = np.zeros_like(fb.buffer)
    for c, n in enumerate(INDICES):
        set_index(scene, n)
        one = rm.create_frame_buffer(w, h, dtype=np.float32)
        ids1 = np.full((h, w), -1, dtype=np.int32)
        rm.create_renderer(1.5, h, w).render(one, scene, prim_id=ids1)
        merged[..., c] = one.buffer[..., c]
        assert np.array_equal(ids1, ids)
    assert np.array_equal(merged, fb.buffer)


def test_dispersive_argument_checks(rm_gpu):
    rm = rm_gpu
    scene, _ = demo()
    L = _abi.load()
    r = rm.create_renderer(1.5, 64, 64)
    fb = rm.create_frame_buffer(64, 64, dtype=np.float32)
    p = r.params(fb, scene)
    h = scene.device_handle()
    handles = (C.c_int64 * 3)(h, h, h)
    assert L.rm_render_dispersive(handles, C.byref(p), fb.buffer.ctypes.data, None, None) == 0
    plain = rm.create_frame_buffer(64, 64, dtype=np.float32)
    r.render(plain, scene)
    assert np.array_equal(plain.buffer, fb.buffer)              # three times the same scene: the plain frame
    p.patch_row_stride = 2
    assert L.rm_render_dispersive(handles, C.byref(p), fb.buffer.ctypes.data, None, None) == -3
    p.patch_row_stride = 1
    p.precision = _abi.RM_FP64
    assert L.rm_render_dispersive(handles, C.byref(p), fb.buffer.ctypes.data, None, None) == -3
    p.precision = _abi.RM_FP32
    handles[1] = 987654
    assert L.rm_render_dispersive(handles, C.byref(p), fb.buffer.ctypes.data, None, None) == -3
    assert L.rm_render_dispersive(None, C.byref(p), fb.buffer.ctypes.data, None, None) == -3
