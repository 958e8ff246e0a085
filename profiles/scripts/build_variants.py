#!/usr/bin/env python
"""Builds experiment variants of libhsearch_b200.so: one translation unit recompiled with extra
-D flags, linked with the standard objects into hsearch_b200/variants/lib_<name>.so (selected at
run time with HS_LIBRARY=...).  Usage: build_variants.py <file.cu> name[:flags] ..."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hsearch_b200 import build as hb  # noqa: E402

hb.build()
src = sys.argv[1]
out_dir = os.path.join(ROOT, "hsearch_b200", "variants")
os.makedirs(out_dir, exist_ok=True)
objdir = os.path.join(ROOT, "hsearch_b200", "build")
procs = []
for spec in sys.argv[2:]:
    name, _, flags = spec.partition(":")
    obj = os.path.join(out_dir, f"{src}.{name}.o")
    cmd = ["/usr/local/cuda/bin/nvcc"] + hb.NVCC_FLAGS + flags.split() + ["-x", "cu", "-c", os.path.join(hb.CSRC, src), "-o", obj]
    procs.append((name, obj, subprocess.Popen(cmd)))
for name, obj, p in procs:
    assert p.wait() == 0, name
    objs = [obj if o.endswith(os.sep + src + ".o") else o for o in
            (os.path.join(objdir, f + ".o") for f in hb.CU + hb.CPP)]
    lib = os.path.join(out_dir, f"lib_{name}.so")
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-shared", "-cudart", "shared", "-o", lib] + objs +
                          ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl", "-Xlinker", "-rpath,/usr/local/cuda/lib64"])
    os.remove(obj)
    print(lib)
