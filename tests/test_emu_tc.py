"""The one-hot tensor filter (csrc/filter_tc.cu: the tcgen05 fallback of the integer metric for len > 30) on the CPU:
tq_to_half_kernel and filter_tc_kernel run unchanged over the emulated mbarrier / tensor memory / tcgen05.mma of
tests/emu/tc_emu.cpp with the library's own geometry and threshold; the survivor set must equal a direct evaluation of
the work list's pairs and keep every pair the oracle's brute force finds within R."""
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
    tc = open(os.path.join(CSRC, "filter_tc.cu")).read()
    text = cut(cuh, "struct WorkItem {", "int launch_probe(")
    text += cut(api, "float filter_threshold(const hs_ctx *ctx) {", "// Device tables of per-table pointers")
    text += cut(cu, "__global__ void build_tq_points_kernel", "// A dense query whose every 8-vector is bit-identical")
    text += cut(tc, "constexpr int kTcThreads = 256;", "__device__ __forceinline__ uint32_t smem_u32")
    text += cut(tc, "// K-major, SWIZZLE_NONE shared-memory matrix descriptor", "__device__ __forceinline__ void umma_f16")
    body = cut(tc, "// ---- FP16 query tables ---", "int tc_min_queries() {")
    body, n1 = re.subn(r'asm volatile\("tcgen05\.alloc.*?"memory"\);', "emu_tmem_alloc(&s_tmem);", body, flags=re.S)
    body, n2 = re.subn(r'asm volatile\("tcgen05\.relinquish_alloc_permit.*?"memory"\);', ";", body, flags=re.S)
    body, n3 = re.subn(r'asm volatile\("tcgen05\.dealloc.*?"memory"\);', ";", body, flags=re.S)
    body, n4 = re.subn(r'asm volatile\("fence\.mbarrier_init\.release\.cluster;" ::: "memory"\);', ";", body)
    assert (n1, n2, n3, n4) == (1, 1, 1, 1), (n1, n2, n3, n4)
    decl = "extern __shared__ __align__(1024) unsigned char tc_smem[];"
    assert decl in body
    body = body.replace(decl, "unsigned char *tc_smem = emu_dyn_smem;")
    text += body
    assert "asm" not in text and "<<<" not in text and "extern __shared__" not in text
    return text


@pytest.mark.skipif(os.uname().machine != "x86_64", reason="the emulation's fiber switch is x86-64 assembly")
def test_one_hot_tensor_filter_under_cpu_emulation(tmp_path):
    (tmp_path / "tc_kernels.inc").write_text(kernel_text())
    obj = tmp_path / "hs_oracle.o"
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-D_GNU_SOURCE", "-c",
                           os.path.join(ROOT, "oracle", "hs_oracle.c"), "-o", str(obj)])
    exe = tmp_path / "tc_emu"
    emu = os.path.join(ROOT, "tests", "emu")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", f"-I{tmp_path}", f"-I{os.path.join(emu, 'stub')}",
                           f"-I{emu}", "-o", str(exe), os.path.join(emu, "tc_emu.cpp"), str(obj), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stdout + out.stderr
    results = re.findall(r" -> (\w+)$", out.stdout, flags=re.M)
    assert len(results) == 3 and all(r == "ok" for r in results), out.stdout
