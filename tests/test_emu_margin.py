"""The tensor filter's margins on the CPU (DESIGN.md "Filter margins"): the FP16 / FP32 decision
D = <x~, q~> + c_q >= rowthr must keep every pair whose FP64 distance is within R.  Everything that determines the
decision runs unchanged, cut out of csrc/filter_mma.cu -- the library's FP16 embedding rows and rounded-down norms
(mma_upload_tables), its error bound (mma_geometry / mma_beta), its query rows with the split constant c_q
(build_qb_*_kernel, write_cq) -- and the kernel's row-threshold and threshold statements are restated in the
harness (this test fails if the source's differ).  D is accumulated in FP32 in four orders and the smallest must
pass for every pair the oracle's brute force finds within R; the filter must also stay tight."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "hsearch_b200", "csrc")

# statements of filter_mma_kernel / launch_filter_mma that tests/emu/margin_emu.cpp restates
RESTATED = [
    "float rt = 0.5f * (nx[i] * (1.0f - 4e-6f) * (1.0f - a.beta) - a.thr);",
    "rt -= (nx[i] + a.thr) * 2.4e-7f + 1e-6f;",
    "const double r2 = rr * (1.0 + 1e-12) + 1e-30;",
    "if ((double)thr < r2) thr = nextafterf(thr, INFINITY);",
    "a.beta = (float)(g.beta * 1.0001);",
    "nx[i] += sh.nx32[c];",
]


def cut(src, start, end):
    a = src.index(start)
    return src[a:src.index(end, a)]


def kernel_text():
    cuh = open(os.path.join(CSRC, "verify.cuh")).read()
    cu = open(os.path.join(CSRC, "filter_mma.cu")).read()
    for stmt in RESTATED:
        assert stmt in cu, f"filter_mma.cu no longer contains: {stmt}"
    text = cut(cuh, "struct MmaGeometry {", "struct MmaItemHost")
    text += cut(cu, "__device__ __forceinline__ bool write_cq", "// queries = the members at positions pos0")
    text += cut(cu, "static inline double mma_beta(int kp) {", "// The tensor path needs every table entry representable")
    text += cut(cu, "int mma_upload_tables(hs_ctx *ctx) {", "int launch_build_qb_points(")
    assert "asm" not in text and "<<<" not in text
    return text


@pytest.mark.skipif(os.uname().machine != "x86_64", reason="the emulation's fiber switch is x86-64 assembly")
def test_tensor_filter_margins_keep_every_pair_within_r(tmp_path):
    (tmp_path / "margin_kernels.inc").write_text(kernel_text())
    obj = tmp_path / "hs_oracle.o"
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-D_GNU_SOURCE", "-c",
                           os.path.join(ROOT, "oracle", "hs_oracle.c"), "-o", str(obj)])
    exe = tmp_path / "margin_emu"
    emu = os.path.join(ROOT, "tests", "emu")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", f"-I{tmp_path}", f"-I{os.path.join(emu, 'stub')}",
                           f"-I{emu}", "-o", str(exe), os.path.join(emu, "margin_emu.cpp"), str(obj), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stdout + out.stderr
    results = re.findall(r" -> (\w+)$", out.stdout, flags=re.M)
    assert len(results) == 4 and all(r == "ok" for r in results), out.stdout
