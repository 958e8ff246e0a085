"""CPU check of the scalar candidate filter without a GPU: build_tq_points_kernel / build_tq_int_kernel and
filter_kernel (csrc/verify.cu) with the library's own filter_threshold (csrc/api.cu) are compiled unchanged over
tests/emu/cuda_emu.h (-ffp-contract=off) and run over ALL (query, fragment) pairs of a seeded set with planted
neighbours: no pair the oracle's brute force finds within R may be dropped (the exact stage decides, the filter
only has to be safe), none may be reported twice, and the filter must be tight (integer metric: exact)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "hsearch_b200", "csrc")


def cut(src, start, end):
    a = src.index(start)
    return src[a:src.index(end, a)]


def kernel_text():
    cuh = open(os.path.join(CSRC, "verify.cuh")).read()
    cu = open(os.path.join(CSRC, "verify.cu")).read()
    api = open(os.path.join(CSRC, "api.cu")).read()
    text = cut(cuh, "struct WorkItem {", "int launch_probe(")
    text += cut(api, "float filter_threshold(const hs_ctx *ctx) {", "// Device tables of per-table pointers")
    text += cut(cu, "__global__ void build_tq_points_kernel", "// A dense query whose every 8-vector is bit-identical")
    body = cut(cu, "template <int MODE, int LENB>\n__global__ void __launch_bounds__(kFilterThreads)\nfilter_kernel",
               "template <int MODE>\nstatic int launch_filter_mode")
    decl = "extern __shared__ __align__(16) float s_tq[];"
    assert decl in body
    body = body.replace(decl, "float *s_tq = reinterpret_cast<float *>(emu_dyn_smem);")
    text += body
    assert "asm" not in text and "<<<" not in text and "extern __shared__" not in text
    return text


@pytest.mark.skipif(os.uname().machine != "x86_64", reason="the emulation's fiber switch is x86-64 assembly")
def test_scalar_filter_never_drops_a_pair_within_r(tmp_path):
    (tmp_path / "filter_kernels.inc").write_text(kernel_text())
    obj = tmp_path / "hs_oracle.o"
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-D_GNU_SOURCE", "-c",
                           os.path.join(ROOT, "oracle", "hs_oracle.c"), "-o", str(obj)])
    exe = tmp_path / "filter_emu"
    emu = os.path.join(ROOT, "tests", "emu")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", f"-I{tmp_path}", f"-I{os.path.join(emu, 'stub')}",
                           f"-I{emu}", "-o", str(exe), os.path.join(emu, "filter_emu.cpp"), str(obj), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stdout + out.stderr
    results = re.findall(r" -> (\w+)$", out.stdout, flags=re.M)
    assert len(results) == 3 and all(r == "ok" for r in results), out.stdout
