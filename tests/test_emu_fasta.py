"""CPU check of the device FASTA parser (E4: ProteinDB::ReadFASTAFile, pcluster/src/pcluster/read_proteins.cpp:6-41)
without a GPU: the kernels of csrc/fasta.cu are compiled unchanged over tests/emu/cuda_emu.h and compared with a
sequential restatement of the reference's reader on seeded texts -- headers with and without descriptions, empty
sequences and lines, lower-case and non-amino-acid letters (replaced with rand(), one per letter in file order: the
state of rand() afterwards must agree too), junk bytes, lines of several chunks, texts without a final newline."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "hsearch_b200", "csrc")


def kernel_text():
    cu = open(os.path.join(CSRC, "fasta.cu")).read()
    a = cu.index("constexpr int kFaThreads")
    text = cu[a:cu.index("int parse_fasta_gpu_impl(", a)]
    assert "asm" not in text and "<<<" not in text and "fasta_bound_scatter_kernel" in text
    return text


@pytest.mark.skipif(os.uname().machine != "x86_64", reason="the emulation's fiber switch is x86-64 assembly")
def test_device_fasta_parser_under_cpu_emulation(tmp_path):
    (tmp_path / "fasta_kernels.inc").write_text(kernel_text())
    exe = tmp_path / "fasta_emu"
    subprocess.check_call(["g++", "-O1", "-std=c++17", f"-I{tmp_path}", f"-I{os.path.join(ROOT, 'tests', 'emu')}",
                           "-o", str(exe), os.path.join(ROOT, "tests", "emu", "fasta_emu.cpp")])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=1200)
    assert out.returncode == 0, out.stdout + out.stderr
    results = re.findall(r" -> (\w+)$", out.stdout, flags=re.M)
    assert len(results) == 6 and all(r == "ok" for r in results), out.stdout
