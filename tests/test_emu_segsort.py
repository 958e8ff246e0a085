"""CPU check of the segmented hit sort's kernels (hsearch_b200/csrc/hitsort.cu) without a GPU: the
kernel text is compiled unchanged over a small fiber-based emulation of the CUDA execution model
(tests/emu/cuda_emu.h: blocks of 1024 threads, __syncthreads, warp ballots and shuffles, shared
memory) and run against std::sort on seeded hit lists -- plain and compact output, bins of several
ranking steps, the range path of bins beyond the shared-memory buffer, the hand-back.  It checks
indices, barrier uniformity and data flow; the GPU tests (test_gpu_segsort.py) check the real thing."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernel_text():
    src = open(os.path.join(ROOT, "hsearch_b200", "csrc", "hitsort.cu")).read()
    a = src.index("constexpr int kSegThreads")
    b = src.index("static int seg_bits_for")
    text = src[a:b]
    # dynamic shared memory: one static arena in the emulation
    for name, decl in (("seg_buf", "extern __shared__ __align__(16) uint32_t seg_buf[];"),
                       ("seg_cnt", "extern __shared__ uint32_t seg_cnt[];"),
                       ("seg_cur", "extern __shared__ uint32_t seg_cur[];")):
        assert decl in text, decl
        text = text.replace(decl, f"uint32_t *{name} = reinterpret_cast<uint32_t *>(emu_dyn_smem);")
    assert "extern __shared__" not in text
    return text


@pytest.mark.skipif(os.uname().machine != "x86_64", reason="the emulation's fiber switch is x86-64 assembly")
def test_segsort_kernels_under_cpu_emulation(tmp_path):
    (tmp_path / "segsort_kernels.inc").write_text(kernel_text())
    exe = tmp_path / "segsort_emu"
    # few bins (2^6): the same kernels with a short per-bin loop, so that the run takes well under a minute
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-DHS_SEG_BIN_BITS=6", f"-I{tmp_path}",
                           f"-I{os.path.join(ROOT, 'tests', 'emu')}", "-o", str(exe),
                           os.path.join(ROOT, "tests", "emu", "segsort_emu.cpp")])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    cases = re.findall(r"^case \d+: .* -> (\w+)$", out.stdout, flags=re.M)
    assert len(cases) >= 10 and all(c == "ok" for c in cases), out.stdout
    assert "handed back" in out.stdout      # the buffer-0 case took the hand-back path
