"""The C-ABI library loads and exports every symbol include/rl_b200.h declares (no GPU needed);
ctypes struct layouts match the C header."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from rendering_learning_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rl_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(rl_[a-z0-9_]+)\s*\(", src))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = A.load_library()
    names = _declared()
    assert names == set(A.SYMBOLS), names ^ set(A.SYMBOLS)
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.rl_abi_version() == A.RL_B200_ABI_VERSION


def test_struct_layouts_match_header():
    structs = ["rl_node", "rl_material", "rl_texture", "rl_image", "rl_perlin", "rl_light", "rl_scene_desc",
               "rl_rtc_camera", "rl_ow_camera", "rl_ray", "rl_hit", "rl_stats", "rl_scene_info",
               "rl_lbvh_host", "rl_job"]
    prog = '#include <stdio.h>\n#include "rl_b200.h"\nint main(){' + "".join(
        f'printf("{s} %zu\\n", sizeof({s}));' for s in structs) + "return 0;}"
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "s")
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.dirname(HEADER), c, "-o", exe])
        out = subprocess.check_output([exe]).decode().split("\n")
    sizes = dict(l.split() for l in out if l)
    for s in structs:
        assert int(sizes[s]) == C.sizeof(getattr(A, s)), s


def test_no_cpu_fallback_without_device():
    """Without a GPU rl_create must fail loudly with RL_E_NO_DEVICE; nothing routes to a CPU path."""
    lib = A.load_library()
    h = C.c_void_p()
    rc = lib.rl_create(0, C.byref(h))
    if rc == A.RL_OK:  # running on a GPU box
        lib.rl_destroy(h)
        pytest.skip("a GPU is visible")
    assert rc == A.RL_E_NO_DEVICE
    assert b"no CPU fallback" in lib.rl_last_error(None)
    from rendering_learning_b200 import Context, RlError
    with pytest.raises(RlError):
        Context(0)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under rendering_learning_b200/ may reference it."""
    pkg = os.path.join(ROOT, "rendering_learning_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
                assert "orc_" not in txt, f


# ---- rust/rl-b200-sys (uncompilable here: no cargo/rustc) is kept in step with the header by parsing both ----
_C2RUST = {"int32_t": "i32", "int64_t": "i64", "uint64_t": "u64", "uint32_t": "u32", "float": "f32", "double": "f64",
           "int": "c_int", "void": "c_void", "char": "c_char"}


def _c_structs():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    out = {}
    for body, name in re.findall(r"typedef struct \w+ \{(.*?)\}\s*(\w+);", src, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.match(r"(const\s+)?(\w+)\s*(\*?)\s*(.*)", decl)
            const, ctype, ptr0, rest = m.groups()
            for item in rest.split(","):
                item = item.strip()
                ptr = bool(ptr0) or item.startswith("*")
                item = item.lstrip("* ")
                am = re.match(r"(\w+)((?:\[\d+\])+)$", item)
                base = _C2RUST.get(ctype, ctype)
                if am:
                    ty = base
                    for dim in reversed(re.findall(r"\[(\d+)\]", am.group(2))):  # C a[256][3] == Rust [[T; 3]; 256]
                        ty = f"[{ty}; {dim}]"
                    fields.append((am.group(1), ty))
                elif ptr:
                    fields.append((item, ("*const " if const else "*mut ") + base))
                else:
                    fields.append((item, base))
        out[name] = fields
    return out


def _rust_structs():
    src = open(os.path.join(ROOT, "rust", "rl-b200-sys", "src", "lib.rs")).read()
    out = {}
    for name, body in re.findall(r"pub struct (\w+) \{(.*?)\n\}", src, flags=re.S):
        out[name] = [(a, b.strip()) for a, b in re.findall(r"pub (\w+): ([^,\n]+),", body)]
    return src, out


def test_rust_sys_matches_header():
    cs = _c_structs()
    src, rs = _rust_structs()
    for name, fields in cs.items():
        assert name in rs, f"{name} missing from rl-b200-sys"
        assert rs[name] == fields, (name, rs[name], fields)
    m = re.search(r'extern "C" \{(.*?)\n\}', src, flags=re.S)
    rust_fns = set(re.findall(r"pub fn (rl_\w+)\(", m.group(1)))
    assert rust_fns == _declared(), rust_fns ^ _declared()
    consts = dict(re.findall(r"pub const (RL_\w+): \w+ = (-?\d+);", src))
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for k, v in re.findall(r"\b(RL_[A-Z0-9_]+)\s*=\s*(-?\d+)", hdr):
        assert consts.get(k) == v, (k, v, consts.get(k))
    assert consts["RL_B200_ABI_VERSION"] == re.search(r"#define RL_B200_ABI_VERSION (\d+)", hdr).group(1)
