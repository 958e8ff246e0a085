"""CPU check of the multi-GPU hit merge without any GPU: the kernels of csrc/comm.cu that count a rank's (query,
table) segments, scan all ranks' counts into every segment's final position and write the rank's hits to those
positions of rank 0's list are compiled unchanged over tests/emu/cuda_emu.h and run for 2, 3 and 8 ranks (the NCCL
all-gather of the counts is a copy here); the merged list must be the sorted union of the ranks' lists -- the
reference's order over the whole database -- byte for byte, and a receive buffer that is too small must be reported
and left untouched."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "hsearch_b200", "csrc")


def kernel_text():
    cu = open(os.path.join(CSRC, "comm.cu")).read()
    a = cu.index("// off[s] = first index of the sorted key list whose segment")
    text = cu[a:cu.index("static int bits_for(uint64_t nvalues)", a)]
    assert "asm" not in text and "<<<" not in text and "scatter_merged_kernel" in text
    return text


@pytest.mark.skipif(os.uname().machine != "x86_64", reason="the emulation's fiber switch is x86-64 assembly")
def test_multi_gpu_hit_merge_under_cpu_emulation(tmp_path):
    (tmp_path / "merge_kernels.inc").write_text(kernel_text())
    exe = tmp_path / "merge_emu"
    subprocess.check_call(["g++", "-O1", "-std=c++17", f"-I{tmp_path}", f"-I{os.path.join(ROOT, 'tests', 'emu')}",
                           "-o", str(exe), os.path.join(ROOT, "tests", "emu", "merge_emu.cpp")])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=1200)
    assert out.returncode == 0, out.stdout + out.stderr
    results = re.findall(r" -> (\w+)$", out.stdout, flags=re.M)
    assert len(results) == 4 and all(r == "ok" for r in results), out.stdout
