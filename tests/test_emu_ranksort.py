"""CPU check of the index-build sort kernels (hsearch_b200/csrc/radix_sort.cu: rank upsweep / downsweep /
bounds, the device-wide scan, one pass of the general radix sort; csrc/verify.cu: the probe of a query's key into the
sorted slots and into the hashed-key index; the L2-blocked gather of the bucket-ordered code stores) under the fiber-based emulation of
tests/emu/cuda_emu.h: the kernel text is compiled unchanged and run against std::stable_sort -- ids in
bucket order must be the stable order (ascending id inside a bucket = the reference's insertion order,
motif_both_points.cpp:212-218), slot boundaries the lower bounds of the ranks."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cut(src, start, end):
    a = src.index(start)
    return src[a:src.index(end, a)]


def kernel_text():
    src = open(os.path.join(ROOT, "hsearch_b200", "csrc", "radix_sort.cu")).read()
    text = cut(src, "constexpr int kSortThreads", "int exclusive_scan_u32(")
    text += cut(src, "constexpr size_t kDownsweepSmem", "// ---- bucket grouping")
    text += cut(src, "constexpr int kRkThreads", "// Rank path of build_table_index")
    decl = "extern __shared__ __align__(16) unsigned char sort_smem[];"
    assert decl in text
    text = text.replace(decl, "unsigned char *sort_smem = emu_dyn_smem;")
    assert "extern __shared__" not in text and "asm" not in text
    return text


def probe_text():
    common = open(os.path.join(ROOT, "hsearch_b200", "csrc", "common.cuh")).read()
    verify = open(os.path.join(ROOT, "hsearch_b200", "csrc", "verify.cu")).read()
    a = common.index("template <int NW>\n__host__ __device__ __forceinline__ uint64_t key_hash")
    text = common[a:common.index("struct TableIndex", a)]
    text += cut(verify, "// ---- probe: query key -> bucket range", "int launch_probe(")
    assert "asm" not in text and "<<<" not in text
    return text


@pytest.mark.skipif(os.uname().machine != "x86_64", reason="the emulation's fiber switch is x86-64 assembly")
def test_index_sort_kernels_under_cpu_emulation(tmp_path):
    (tmp_path / "radixsort_kernels.inc").write_text(kernel_text())
    (tmp_path / "probe_kernels.inc").write_text(probe_text())
    src = open(os.path.join(ROOT, "hsearch_b200", "csrc", "radix_sort.cu")).read()
    (tmp_path / "gather_kernels.inc").write_text(cut(src, "constexpr uint32_t kGatherBlockBytes", "// Builds codes_sorted of every table."))
    exe = tmp_path / "ranksort_emu"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-DHS_GATHER_PART=64", f"-I{tmp_path}", f"-I{os.path.join(ROOT, 'tests', 'emu')}",
                           "-o", str(exe), os.path.join(ROOT, "tests", "emu", "ranksort_emu.cpp")])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    results = re.findall(r" -> (\w+)$", out.stdout, flags=re.M)
    assert len(results) == 15 and all(r == "ok" for r in results), out.stdout
