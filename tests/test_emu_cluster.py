"""The near-pair clustering path on the CPU against the oracle's cluster composition (in-bucket pairs
within R -> UnionFind, pcluster/src/pcluster/union_find.cpp:3-33 -> smallest-id labels): small_bucket_pairs_kernel,
the tiled scalar self-join filter_kernel<SelfJoin> with cluster_impl's work items, exact_kernel in self-join mode
(FP64, sqrt predicate, lock-free union) and uf_flatten run unchanged over tests/emu/cuda_emu.h; labels and the edge
count must equal orc_cluster's, also when the pair work is split over two emulated ranks whose forests are merged
through their labels as on a communicator."""
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
    api = open(os.path.join(CSRC, "api.cu")).read()
    cl = open(os.path.join(CSRC, "cluster.cu")).read()
    text = cut(cuh, "struct WorkItem {", "int launch_probe(")
    text += cut(cuh, "// ---- lock-free union-find (device)", "int launch_exact(")
    text += cut(api, "float filter_threshold(const hs_ctx *ctx) {", "// Device tables of per-table pointers")
    f = cut(cu, "template <int MODE, int LENB>\n__global__ void __launch_bounds__(kFilterThreads)\nfilter_kernel",
            "template <int MODE>\nstatic int launch_filter_mode")
    f = f.replace("extern __shared__ __align__(16) float s_tq[];", "float *s_tq = reinterpret_cast<float *>(emu_dyn_smem);")
    text += f
    e = cut(cu, "// ---- exact stage ---", "int launch_exact(hs_ctx")
    e = e.replace('asm volatile("prefetch.global.L2 [%0];" ::"l"(a.rec + (uint64_t)nx.pos * a.rec_stride));', "(void)nx;")
    e = e.replace("extern __shared__ __align__(16) unsigned char exact_smem[];", "unsigned char *exact_smem = emu_dyn_smem;")
    text += e
    text += cut(cl, "constexpr uint32_t kSmallBucket = 64;", "int comm_allgather_u32")
    assert "asm" not in text and "<<<" not in text and "extern __shared__" not in text
    return text


@pytest.mark.skipif(os.uname().machine != "x86_64", reason="the emulation's fiber switch is x86-64 assembly")
def test_cluster_path_under_cpu_emulation(tmp_path):
    (tmp_path / "cluster_kernels.inc").write_text(kernel_text())
    obj = tmp_path / "hs_oracle.o"
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-D_GNU_SOURCE", "-c",
                           os.path.join(ROOT, "oracle", "hs_oracle.c"), "-o", str(obj)])
    exe = tmp_path / "cluster_emu"
    emu = os.path.join(ROOT, "tests", "emu")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", f"-I{tmp_path}", f"-I{os.path.join(emu, 'stub')}",
                           f"-I{emu}", "-o", str(exe), os.path.join(emu, "cluster_emu.cpp"), str(obj), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stdout + out.stderr
    results = re.findall(r" -> (\w+)$", out.stdout, flags=re.M)
    assert len(results) == 2 and all(r == "ok" for r in results), out.stdout
