"""CPU check of the recall join (R1: evaulate() and weight() of motif_both_points.cpp:67-87,100-165) without a
GPU: the recall kernels of csrc/evaluate.cu are compiled unchanged over tests/emu/cuda_emu.h and compared with the
oracle's sequential restatement (pinned bit-exact against the reference's own function): matched / missed / extra
counts and the distance bins equal, the weighted sums within 1e-12 relative, an unordered list reported."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "hsearch_b200", "csrc")


def kernel_text():
    cu = open(os.path.join(CSRC, "evaluate.cu")).read()
    a = cu.index("constexpr int kRecallThreads")
    text = cu[a:cu.index("int evaluate_recall_dev(", a)]
    assert "asm" not in text and "<<<" not in text and "recall_join_kernel" in text
    return text


@pytest.mark.skipif(os.uname().machine != "x86_64", reason="the emulation's fiber switch is x86-64 assembly")
def test_recall_join_under_cpu_emulation(tmp_path):
    (tmp_path / "recall_kernels.inc").write_text(kernel_text())
    obj = tmp_path / "hs_oracle.o"
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-D_GNU_SOURCE", "-c",
                           os.path.join(ROOT, "oracle", "hs_oracle.c"), "-o", str(obj)])
    exe = tmp_path / "recall_emu"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", f"-I{tmp_path}",
                           f"-I{os.path.join(ROOT, 'tests', 'emu')}", "-o", str(exe),
                           os.path.join(ROOT, "tests", "emu", "recall_emu.cpp"), str(obj), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=1200)
    assert out.returncode == 0, out.stdout + out.stderr
    results = re.findall(r" -> (\w+)$", out.stdout, flags=re.M)
    assert len(results) == 3 and all(r == "ok" for r in results), out.stdout
