"""CPU check of the production hash against the oracle, without a GPU: hash_fast_kernel (FP32 partial sums, FP64
guard band, dense bucket ranks, fragment records; csrc/hash.cu) is compiled unchanged over tests/emu/cuda_emu.h
and launched in the library's configurations -- among them the headline one, 1024 threads in eight groups with
replicated tables -- on projection data prepared by the library's own host code (setup_projection / setup_ranks,
run unchanged over a stand-in CUDA runtime).  Bucket ints must equal the oracle's bit for bit (lsh.hpp:33-59),
ranks must be the positions of the key strings, records the codes plus ranks; one setting puts nearly every
projection inside the guard band."""
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
    cuh = open(os.path.join(CSRC, "hash.cuh")).read()
    cu = open(os.path.join(CSRC, "hash.cu")).read()
    text = cut(cuh, "template <int KW>\nstruct KeyBuilder", "constexpr uint64_t kHashRangeAlign")
    text += cut(cu, "__device__ __noinline__ int exact_bucket_codes_cold", "// Fragment records without ranks")
    bar = 'asm volatile("bar.sync %0, %1;" ::"r"(gid + 1), "r"(GT) : "memory");'
    assert bar in text      # the groups' named barrier
    text = text.replace(bar, "emu_named_barrier(gid + 1, GT);")
    decl = "extern __shared__ __align__(128) unsigned char smem_raw[];"
    assert decl in text
    text = text.replace(decl, "unsigned char *smem_raw = emu_dyn_smem;")
    assert "asm" not in text and "<<<" not in text and "extern __shared__" not in text
    return text


def host_text():
    cu = open(os.path.join(CSRC, "hash.cu")).read()
    a = cu.index("// std::to_string(int) strings of a bucket tuple, packed like KeyBuilder does")
    text = cu[a:cu.rindex("}  // namespace hs")]
    assert "int setup_projection(hs_ctx *ctx" in text and "static int setup_ranks(hs_ctx *ctx)" in text
    assert "<<<" not in text
    return text


@pytest.mark.skipif(os.uname().machine != "x86_64", reason="the emulation's fiber switch is x86-64 assembly")
def test_production_hash_under_cpu_emulation(tmp_path):
    (tmp_path / "hashfast_kernels.inc").write_text(kernel_text())
    (tmp_path / "hashfast_host.inc").write_text(host_text())
    obj = tmp_path / "hs_oracle.o"
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-D_GNU_SOURCE", "-c",
                           os.path.join(ROOT, "oracle", "hs_oracle.c"), "-o", str(obj)])
    exe = tmp_path / "hashfast_emu"
    emu = os.path.join(ROOT, "tests", "emu")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", f"-I{tmp_path}", f"-I{os.path.join(emu, 'stub')}",
                           f"-I{emu}", "-o", str(exe), os.path.join(emu, "hashfast_emu.cpp"), str(obj), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stdout + out.stderr
    results = re.findall(r" -> (\w+)", out.stdout)
    assert len(results) == 7 and all(r == "ok" for r in results), out.stdout
