"""CPU check of the front-end and union-find kernels against the oracle, without a GPU: window extraction
and ProteinID (csrc/extract.cu), six-frame translation, 3-mer histograms and KLSH bits (csrc/sequence.cu) and the lock-free union-find
(csrc/verify.cuh, csrc/cluster.cu) are compiled unchanged over the fiber-based emulation of
tests/emu/cuda_emu.h and compared with oracle/hs_oracle.c on seeded inputs (ragged proteins, sequences
of every short length, random edge lists with chains and self loops, the label merge of two forests)."""
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
    verify = open(os.path.join(CSRC, "verify.cuh")).read()
    cluster = open(os.path.join(CSRC, "cluster.cu")).read()
    extract = open(os.path.join(CSRC, "extract.cu")).read()
    sequence = open(os.path.join(CSRC, "sequence.cu")).read()
    text = cut(verify, "// ---- lock-free union-find (device)", "struct ExactArgs")
    text += cut(cluster, "__global__ void iota32_kernel", "// warp per bucket")
    text += cut(cluster, "__global__ void uf_flatten_kernel", "int comm_allgather_u32")
    text += cut(cluster, "__global__ void uf_edges_kernel", "// FindRoot / JoinUnion (union_find.cpp:16-33) over ne edges")
    text += cut(extract, "// one thread per fragment; frag_start[p]", "int extract_windows_impl")
    text += cut(extract, "__global__ void protein_id_kernel", "int protein_id_impl")
    text += cut(sequence, "constexpr int kFeat = 512;", "// GetHashValue on the host")
    text += cut(sequence, "__constant__ char c_codon_aa[65]", "}  // namespace hs")
    assert "asm" not in text and "<<<" not in text
    return text


@pytest.mark.skipif(os.uname().machine != "x86_64", reason="the emulation's fiber switch is x86-64 assembly")
def test_frontend_and_union_find_kernels_under_cpu_emulation(tmp_path):
    (tmp_path / "frontend_kernels.inc").write_text(kernel_text())
    obj = tmp_path / "hs_oracle.o"
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-D_GNU_SOURCE", "-c",
                           os.path.join(ROOT, "oracle", "hs_oracle.c"), "-o", str(obj)])
    exe = tmp_path / "frontend_emu"
    subprocess.check_call(["g++", "-O1", "-std=c++17", f"-I{tmp_path}", f"-I{os.path.join(ROOT, 'tests', 'emu')}",
                           "-ffp-contract=off", "-o", str(exe), os.path.join(ROOT, "tests", "emu", "frontend_emu.cpp"), str(obj), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    results = re.findall(r" -> (\w+)$", out.stdout, flags=re.M)
    assert len(results) == 8 and all(r == "ok" for r in results), out.stdout
