"""CPU check of the greedy centre clustering kernel (CL1: Clustering() of hclust2.cpp:86-151) without a GPU:
greedy_round_kernel (csrc/cluster.cu, one warp per bucket, 32 members at a time) is compiled unchanged over
tests/emu/cuda_emu.h (-ffp-contract=off) and must leave every point in the state and with the centre the
reference's sequential procedure gives it, on seeded families of near-duplicates (bucket keys and distances
from oracle/hs_oracle.c).  The GPU suite compares the hclust2 / hclust3 programs with the reference binaries."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "hsearch_b200", "csrc")


def kernel_text():
    cu = open(os.path.join(CSRC, "cluster.cu")).read()
    a = cu.index("constexpr int kGreedyThreads")
    text = cu[a:cu.index("__global__ void greedy_init_kernel", a)]
    assert "asm" not in text and "<<<" not in text and "greedy_round_kernel" in text
    return text


@pytest.mark.skipif(os.uname().machine != "x86_64", reason="the emulation's fiber switch is x86-64 assembly")
def test_greedy_clustering_kernel_under_cpu_emulation(tmp_path):
    (tmp_path / "greedy_kernels.inc").write_text(kernel_text())
    obj = tmp_path / "hs_oracle.o"
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-D_GNU_SOURCE", "-c",
                           os.path.join(ROOT, "oracle", "hs_oracle.c"), "-o", str(obj)])
    exe = tmp_path / "greedy_emu"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", f"-I{tmp_path}",
                           f"-I{os.path.join(ROOT, 'tests', 'emu')}", "-o", str(exe),
                           os.path.join(ROOT, "tests", "emu", "greedy_emu.cpp"), str(obj), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=1200)
    assert out.returncode == 0, out.stdout + out.stderr
    results = re.findall(r" -> (\w+)$", out.stdout, flags=re.M)
    assert len(results) == 3 and all(r == "ok" for r in results), out.stdout
