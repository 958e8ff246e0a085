"""CPU check of the FP64 hash kernels against the oracle, without a GPU: hash_exact_kernel (the all-FP64 hash and
the audit, in the reference's operation order, lsh.hpp:33-59), hash_queries_kernel (motif_both_points.cpp:227) and
the packed digit-string key of hash.cuh are compiled unchanged over tests/emu/cuda_emu.h (-ffp-contract=off) and
compared with oracle/hs_oracle.c: bucket ints bit-equal, keys equal to the nibble packing of the concatenated
std::to_string strings, the audit silent on untouched keys and exact on a tampered one."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "hsearch_b200", "csrc")


def kernel_text():
    cuh = open(os.path.join(CSRC, "hash.cuh")).read()
    cu = open(os.path.join(CSRC, "hash.cu")).read()
    a = cuh.index("template <int KW>\nstruct KeyBuilder")
    text = cuh[a:cuh.index("struct HashChunkArgs", a)]
    a = cuh.index("__device__ __forceinline__ bool rank_tuple_push")
    text += cuh[a:cuh.index("constexpr uint64_t kHashRangeAlign", a)]
    k = cu.index("hash_exact_kernel(const uint8_t")
    a = cu.rindex("template <int KW>", 0, k)
    text += cu[a:cu.index("// ---- host side", a)]
    assert "asm" not in text and "<<<" not in text and "hash_queries_kernel" in text
    return text


@pytest.mark.skipif(os.uname().machine != "x86_64", reason="the emulation's fiber switch is x86-64 assembly")
def test_fp64_hash_kernels_under_cpu_emulation(tmp_path):
    (tmp_path / "hash_kernels.inc").write_text(kernel_text())
    obj = tmp_path / "hs_oracle.o"
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-D_GNU_SOURCE", "-c",
                           os.path.join(ROOT, "oracle", "hs_oracle.c"), "-o", str(obj)])
    exe = tmp_path / "hash_emu"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", f"-I{tmp_path}",
                           f"-I{os.path.join(ROOT, 'tests', 'emu')}", "-o", str(exe),
                           os.path.join(ROOT, "tests", "emu", "hash_emu.cpp"), str(obj), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    results = re.findall(r" -> (\w+)$", out.stdout, flags=re.M)
    assert len(results) == 6 and all(r == "ok" for r in results), out.stdout
