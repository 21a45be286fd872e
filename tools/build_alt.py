"""Build an EXPERIMENT variant of the library next to the product one:
    python tools/build_alt.py NAME -DFLAG=VALUE ...   ->  rendering_learning_b200/librl_b200_NAME.so
tools/time_ow.py / time_rtc.py pick it up with `lib=NAME`.  The product never loads these."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g

name, defs = sys.argv[1], sys.argv[2:]
odir = os.path.join(g.CSRC, "_build_" + name)
os.makedirs(odir, exist_ok=True)
lib = g.LIB.replace(".so", f"_{name}.so")
jobs, objs = [], []
for src in g.SOURCES:
    obj = os.path.join(odir, os.path.splitext(src)[0] + ".o")
    objs.append(obj)
    jobs.append((src, subprocess.Popen(["nvcc"] + g.NVCC_FLAGS + defs + ["-ccbin", "/usr/bin/g++", "-c", src, "-o", obj], cwd=g.CSRC)))
for src, p in jobs:
    if p.wait() != 0:
        raise SystemExit(f"nvcc failed on {src}")
subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs, cwd=g.CSRC)
print(lib)
