"""CPU check of the exact stage against the oracle, without a GPU: exact_kernel (FP64 distance in the reference's
operation order, predicate, first-table-wins, hit emission; csrc/verify.cu) is compiled unchanged over
tests/emu/cuda_emu.h (-ffp-contract=off) and handed every member of every query's buckets as survivors; the hits
must equal the oracle's Search() / brute force -- set, first table and FP64 distance bit for bit -- for residue-
string and dense queries, the packed-key and the rank path of the first-table rule, both layouts of the residue-pair table, len 10 / 20 / 25 and the integer metric."""
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
    text = cut(cuh, "struct Survivor {", "constexpr int kFilterThreads")
    text += cut(cuh, "enum FilterMode", "struct FilterArgs")
    text += cut(cuh, "// ---- lock-free union-find (device)", "int launch_exact(")
    body = cut(cu, "// ---- exact stage ---", "int launch_exact(hs_ctx")
    pf = 'asm volatile("prefetch.global.L2 [%0];" ::"l"(a.rec + (uint64_t)nx.pos * a.rec_stride));'
    assert pf in body      # (a cache hint: no effect on the result)
    body = body.replace(pf, "(void)nx;")
    decl = "extern __shared__ __align__(16) unsigned char exact_smem[];"
    assert decl in body
    body = body.replace(decl, "unsigned char *exact_smem = emu_dyn_smem;")
    text += body
    assert "asm" not in text and "<<<" not in text and "extern __shared__" not in text
    return text


@pytest.mark.skipif(os.uname().machine != "x86_64", reason="the emulation's fiber switch is x86-64 assembly")
def test_exact_stage_under_cpu_emulation(tmp_path):
    (tmp_path / "exact_kernels.inc").write_text(kernel_text())
    obj = tmp_path / "hs_oracle.o"
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-D_GNU_SOURCE", "-c",
                           os.path.join(ROOT, "oracle", "hs_oracle.c"), "-o", str(obj)])
    exe = tmp_path / "exact_emu"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", f"-I{tmp_path}",
                           f"-I{os.path.join(ROOT, 'tests', 'emu')}", "-o", str(exe),
                           os.path.join(ROOT, "tests", "emu", "exact_emu.cpp"), str(obj), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=1200)
    assert out.returncode == 0, out.stdout + out.stderr
    results = re.findall(r" -> (\w+)$", out.stdout, flags=re.M)
    assert len(results) == 9 and all(r == "ok" for r in results), out.stdout
