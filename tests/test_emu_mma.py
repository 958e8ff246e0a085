"""The pipelined tensor filter's kernel on the CPU: filter_mma_kernel (csrc/filter_mma.cu) runs unchanged -- 22
warps in four roles -- over an emulation of what the hardware provides (mbarriers with transaction bytes, bulk
copies, tensor memory, tcgen05.mma on the kernel's own shared-memory descriptors, commit, tcgen05.ld; tests/emu/
mma_emu.cpp).  Its survivor set must equal a direct evaluation of every (query, member) pair of the work list and
keep every pair the oracle's brute force finds within R.  This checks the pipeline's logic; the GPU tests check it
on the real tensor cores."""
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
    cu = open(os.path.join(CSRC, "filter_mma.cu")).read()
    text = cut(cuh, "struct Survivor {", "constexpr int kFilterThreads")
    text += cut(cuh, "struct MmaGeometry {", "struct MmaItemHost")
    text += "#define HS_MMA_EVENTS 0\n"
    text += cut(cu, "constexpr int kMmaEpiWarps = 16;", "__device__ __forceinline__ uint32_t smem_addr")
    text += cut(cu, "// K-major, SWIZZLE_NONE shared-memory matrix descriptor", "__device__ __forceinline__ void mma_f16_ss")
    body = cut(cu, "// Append the survivors the lanes of a warp hold in their private slots", "// One event per lane: the column test")
    # the three statements of the kernel body that talk to the hardware directly
    body, n1 = re.subn(r'asm volatile\("tcgen05\.alloc.*?"memory"\);', "emu_tmem_alloc(&sh.tmem_base);", body, flags=re.S)
    body, n2 = re.subn(r'asm volatile\("tcgen05\.relinquish_alloc_permit.*?"memory"\);', ";", body, flags=re.S)
    body, n3 = re.subn(r'asm volatile\("tcgen05\.dealloc.*?"memory"\);', ";", body, flags=re.S)
    body, n4 = re.subn(r'asm volatile\("fence\.mbarrier_init\.release\.cluster;" ::: "memory"\);', ";", body)
    assert (n1, n2, n3, n4) == (1, 1, 1, 1), (n1, n2, n3, n4)
    decl = "extern __shared__ __align__(1024) unsigned char mma_smem[];"
    assert decl in body
    body = body.replace(decl, "unsigned char *mma_smem = emu_dyn_smem;")
    text += body
    text += cut(cu, "__device__ __forceinline__ bool write_cq", "// queries = the members at positions pos0")
    text += cut(cu, "static inline double mma_beta(int kp) {", "// The tensor path needs every table entry representable")
    text += cut(cu, "int mma_upload_tables(hs_ctx *ctx) {", "int launch_build_qb_points(")
    assert "asm" not in text and "<<<" not in text and "extern __shared__" not in text
    return text


@pytest.mark.skipif(os.uname().machine != "x86_64", reason="the emulation's fiber switch is x86-64 assembly")
def test_pipelined_tensor_filter_under_cpu_emulation(tmp_path):
    (tmp_path / "mma_kernels.inc").write_text(kernel_text())
    obj = tmp_path / "hs_oracle.o"
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-D_GNU_SOURCE", "-c",
                           os.path.join(ROOT, "oracle", "hs_oracle.c"), "-o", str(obj)])
    exe = tmp_path / "mma_emu"
    emu = os.path.join(ROOT, "tests", "emu")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", f"-I{tmp_path}", f"-I{os.path.join(emu, 'stub')}",
                           f"-I{emu}", "-o", str(exe), os.path.join(emu, "mma_emu.cpp"), str(obj), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stdout + out.stderr
    results = re.findall(r" -> (\w+)$", out.stdout, flags=re.M)
    assert len(results) == 3 and all(r == "ok" for r in results), out.stdout
