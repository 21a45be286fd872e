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
    structs = ["rl_node", "rl_material", "rl_texture", "rl_image", "rl_light", "rl_scene_desc",
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
