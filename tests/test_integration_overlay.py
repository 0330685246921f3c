"""The Rust overlay (integration/engine): `ffi.rs` must be a faithful transcription of include/rm_b200.h -- every struct
field for field (names, order, types), every function (name, argument count and types, return type) -- and the patches
must apply to the reference's files.  No Rust toolchain exists in this image, so this is the check that keeps the glue
honest."""
import os
import re
import shutil
import subprocess

import pytest

from tests.conftest import ROOT

HEADER = os.path.join(ROOT, "include", "rm_b200.h")
FFI = os.path.join(ROOT, "integration", "engine", "src", "ffi.rs")
PATCHES = os.path.join(ROOT, "integration", "engine", "patches")

C_TO_RUST = {"double": "f64", "float": "f32", "int32_t": "i32", "uint32_t": "u32", "int64_t": "i64", "uint64_t": "u64",
             "uint8_t": "u8", "int": "c_int", "size_t": "usize", "char": "c_char", "void": "c_void", "unsigned char": "u8",
             "long long": "i64", "RmScene": "RmScene"}


def strip_comments(src):
    return re.sub(r"//[^\n]*", "", re.sub(r"/\*.*?\*/", "", src, flags=re.S))


def c_type_to_rust(t, name_suffix=""):
    """'const double*' -> '*const f64'; arrays via name_suffix '[3]'."""
    t = t.strip()
    const = False
    ptr = 0
    while t.endswith("*") or t.endswith("const"):
        if t.endswith("*"):
            ptr += 1
            t = t[:-1].strip()
        else:
            t = t[:-5].strip()
            const_after = True  # noqa: F841  (T* const: constness of the pointer itself, irrelevant)
    if t.startswith("const "):
        const = True
        t = t[6:].strip()
    base = C_TO_RUST.get(t, t)
    out = base
    for i in range(ptr):
        out = ("*const " if (const and i == 0) else "*mut ") + out
    m = re.fullmatch(r"\[(\w+)\]", name_suffix or "")
    if m:
        out = "[%s; %s]" % (out, m.group(1))
    return out


def header_structs():
    src = strip_comments(open(HEADER).read())
    out = {}
    for body, name in re.findall(r"typedef struct \w+ \{(.*?)\}\s*(\w+);", src, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            m = re.match(r"(.*?)(\w+(?:\s*,\s*\w+)*)\s*((?:\[\w+\])?)$", decl)
            ctype, names, arr = m.group(1), m.group(2), m.group(3)
            # 'const T* a' parses as type 'const T*' + name; 'uint64_t a, b, c' as several names
            for n in [x.strip() for x in names.split(",")]:
                fields.append((n, c_type_to_rust(ctype, arr)))
        out[name] = fields
    return out


def header_functions():
    src = strip_comments(open(HEADER).read())
    src = re.sub(r"typedef struct \w+ \{.*?\}\s*\w+;", "", src, flags=re.S)
    src = re.sub(r"typedef enum \w+ \{.*?\}\s*\w+;", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)              # preprocessor lines
    src = re.sub(r"typedef \w+ \w+;", "", src)
    out = {}
    for ret, name, args in re.findall(r"([\w\s\*]+?)\b(rm_\w+)\s*\(([^)]*)\)\s*;", src):
        ret = " ".join(ret.split())
        params = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                m = re.match(r"(.*?)(\w+)\s*((?:\[\w*\])?)$", a)
                ctype, arr = m.group(1).strip(), m.group(3)
                if arr:                                      # an array parameter is a pointer
                    ctype += "*"
                params.append(c_type_to_rust(ctype))
        out[name] = (None if ret == "void" else c_type_to_rust(ret), params)
    return out


def rust_items():
    src = strip_comments(open(FFI).read())
    src = re.sub(r"///[^\n]*", "", src)
    structs = {}
    for name, body in re.findall(r"pub struct (\w+) \{(.*?)\n\}", src, flags=re.S):
        structs[name] = [(n, " ".join(t.split())) for n, t in re.findall(r"pub (\w+):\s*([^,\n]+),", body)]
    fns = {}
    block = re.search(r'extern "C" \{(.*?)\n\}', src, flags=re.S).group(1)
    for name, args, ret in re.findall(r"pub fn (\w+)\((.*?)\)\s*(?:->\s*([^;]+))?;", block, flags=re.S):
        params = [" ".join(a.split(":", 1)[1].split()) for a in args.split(",") if ":" in a]
        fns[name] = (ret.strip() if ret else None, params)
    consts = dict(re.findall(r"pub const (\w+): \w+ = (-?\w+);", src))
    return structs, fns, consts


def norm(t):
    """equivalences that are the same ABI: i32 / c_int, *mut / *const of the pointed-to data in out-parameters of handles"""
    t = t.replace("c_int", "i32").replace("RM_MAX_RANKS", "16")
    return t


def test_every_struct_matches_the_header_field_for_field():
    hs, (rs, _f, _c) = header_structs(), rust_items()
    assert set(hs) >= {"RmReflectance", "RmSphere", "RmPolygon", "RmTriangle", "RmObj", "RmLight", "RmShapeRef", "RmFlatScene",
                       "RmParams", "RmStats", "RmExchange"}
    for name, fields in hs.items():
        assert name in rs, "ffi.rs lacks struct %s" % name
        got = rs[name]
        assert [n for n, _ in got] == [n for n, _ in fields], "%s: field order differs: %s vs %s" % (name, got, fields)
        for (n, tr), (_n, tc) in zip(got, fields):
            assert norm(tr) == norm(tc), "%s.%s: %s in ffi.rs, %s in the header" % (name, n, tr, tc)


def test_every_function_of_the_header_is_declared_with_the_same_signature():
    hf, (_s, rf, _c) = header_functions(), rust_items()
    assert len(hf) >= 37 and "rm_render_rows_f64" in hf and "rm_render_frame" in hf and "rm_peer_alloc" in hf
    assert sorted(rf) == sorted(hf), "ffi.rs and rm_b200.h declare different functions: %s" % sorted(set(rf) ^ set(hf))
    for name, (ret, params) in hf.items():
        rret, rparams = rf[name]
        assert (rret is None) == (ret is None) and (ret is None or norm(rret) == norm(ret)), (name, rret, ret)
        assert len(rparams) == len(params), "%s: %d parameters in ffi.rs, %d in the header" % (name, len(rparams), len(params))
        for i, (a, b) in enumerate(zip(rparams, params)):
            a, b = norm(a), norm(b)
            # *mut vs *const on the outermost level of double pointers / handle bytes is not part of the C ABI
            assert a == b or a.replace("*const", "*mut") == b.replace("*const", "*mut"), "%s arg %d: %s vs %s" % (name, i, a, b)


def test_constants_agree():
    src = open(HEADER).read()
    _s, _f, consts = rust_items()
    for name in ("RM_ABI_VERSION", "RM_MAX_RANKS", "RM_IPC_HANDLE_BYTES", "RM_MAILBOX_BYTES", "RM_ROWS_RETAINED"):
        m = re.search(r"#define %s (\d+)" % name, src)
        assert m and consts[name] == m.group(1), name
    for name, val in re.findall(r"\b(RM_(?:OK|ERR_\w+|FP32|FP64|SHAPE_\w+)) = (-?\d+)", src):
        assert consts[name] == val, name


def test_patches_apply_to_the_reference(reference_dir, tmp_path):
    dst = tmp_path / "ref"
    shutil.copytree(os.path.join(reference_dir, "engine", "src"), dst / "engine" / "src")
    shutil.copy(os.path.join(reference_dir, "engine", "Cargo.toml"), dst / "engine" / "Cargo.toml")
    names = sorted(os.listdir(PATCHES))
    assert len(names) == 8
    for p in names:
        r = subprocess.run(["patch", "-d", str(dst), "-p0", "--forward", "--fuzz=0", "-i", os.path.join(PATCHES, p)],
                           capture_output=True, text=True)
        assert r.returncode == 0, "%s does not apply: %s" % (p, r.stdout + r.stderr)
    out = (dst / "engine" / "src" / "renderer.rs").read_text()
    assert "rm_render_rows_f64" in out and "into_par_iter" not in out
    assert "pub fn render(&self, frame: &mut FrameBuffer, scene: &Scene) -> String" in out
    assert 'format!(\n            "Scene rendered in {} ms ({} fps, {:.2} MP/s)"' in out
    assert "fn flatten" in (dst / "engine" / "src" / "shapes.rs").read_text()
