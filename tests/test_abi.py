"""The C-ABI shared library: builds, loads, exports every symbol include/*.h declares, struct
layouts agree between the header (gcc) and the ctypes mirror, and -- without a GPU -- every
compute entry point fails loudly instead of falling back to the CPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

from rusty_marcher_b200 import _abi
from tests.conftest import ROOT, has_gpu

HEADERS = [os.path.join(ROOT, "include", "rm_b200.h"), os.path.join(ROOT, "include", "rm_b200_host.h")]


def declared_functions():
    names = []
    for h in HEADERS:
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b(rm_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_builds_and_exports_every_declared_symbol():
    L = _abi.load()
    decl = declared_functions()
    assert len(decl) >= 35
    for name in decl:
        assert hasattr(L, name), "librm_b200.so does not export %s" % name
    assert sorted(_abi.SYMBOLS) == decl, "ctypes table and headers disagree"
    assert L.rm_abi_version() == 5


def test_struct_layouts_match_the_header(tmp_path):
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "rm_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(RmReflectance),sizeof(RmSphere),sizeof(RmPolygon),sizeof(RmTriangle),sizeof(RmObj),"
                   "sizeof(RmLight),sizeof(RmShapeRef),sizeof(RmFlatScene),sizeof(RmParams),sizeof(RmStats),sizeof(RmExchange));return 0;}\n")
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(t) for t in (_abi.RmReflectance, _abi.RmSphere, _abi.RmPolygon, _abi.RmTriangle, _abi.RmObj,
                                  _abi.RmLight, _abi.RmShapeRef, _abi.RmFlatScene, _abi.RmParams, _abi.RmStats, _abi.RmExchange)]
    assert got == want


def test_params_default_are_the_reference_constants():
    p = _abi.RmParams()
    _abi.load().rm_params_default(C.byref(p), 1600, 1280)
    assert (p.width, p.height, p.fov, p.max_depth, p.background, p.patch_size) == (1600, 1280, 1.5, 3, 0.1, 32)
    assert (p.patch_row_begin, p.patch_row_end, p.precision) == (0, -1, _abi.RM_FP32)
    r = _abi.RmReflectance()
    _abi.load().rm_reflectance_default(C.byref(r))
    assert (r.diffusion, tuple(r.diffuse_color), r.specular, r.specular_exponent, r.is_glass_like, r.reflection,
            r.refractive_index) == (1., (1., 1., 1.), 1., 30., 0, 0.95, 1.)


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_cpu_fallback():
    L = _abi.load()
    assert L.rm_init(0) == -1                                   # RM_ERR_NO_DEVICE
    assert b"no CPU fallback" in L.rm_last_error()
    import rusty_marcher_b200 as rm
    sc = rm.Scene.create_default()
    fb = rm.create_frame_buffer(64, 64)
    with pytest.raises(rm.RmError) as e:
        rm.create_renderer(1.5, 64, 64).render(fb, sc)
    assert e.value.code in (-1, -2)
    h = C.c_int64(0)
    flat = sc.flatten()
    assert L.rm_scene_upload(C.byref(flat.c), C.byref(h)) == -2  # RM_ERR_NOT_INITIALISED
    p = _abi.RmParams()
    L.rm_params_default(C.byref(p), 64, 64)
    assert L.rm_render(1, C.byref(p), None, None, None, None) == -2
    assert L.rm_render_device(1, C.byref(p), 16, None, 16, None) == -2
    assert L.rm_measure_fp32_peak(None, None) == -2


def test_product_never_references_the_oracle():
    """The product path must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "rusty_marcher_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "rm_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f
    out = subprocess.run(["nm", "-D", "--defined-only", _abi.lib_path()], capture_output=True, text=True).stdout
    assert "orc_" not in out and "emu_" not in out
