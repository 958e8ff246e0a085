"""The whole hot path on the CPU, stage after stage, against the oracle's Search() -- no GPU: the production hash
(hash_fast_kernel in the headline launch configuration, on the library's own setup_projection), the rank sort and
slot boundaries, the L2-blocked gather of the code stores, the query hash and probe, the pipelined tcgen05 filter
(over emulated mbarriers / bulk copies / tensor memory / MMA), the exact FP64 stage with the rank-path first-table
rule, and the segmented hit sort run unchanged over tests/emu/cuda_emu.h, each consuming what the previous one
produced; the final hit list must equal the oracle's -- order, first table, ids and FP64 distances bit for bit
(motif_both_points.cpp:195-250)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "hsearch_b200", "csrc")


def cut(src, start, end):
    a = src.index(start)
    return src[a:src.index(end, a)]


def read(name):
    return open(os.path.join(CSRC, name)).read()


def kernel_text():
    hash_cuh, hash_cu = read("hash.cuh"), read("hash.cu")
    sort_cu, verify_cuh, verify_cu = read("radix_sort.cu"), read("verify.cuh"), read("verify.cu")
    mma_cu, hitsort_cu = read("filter_mma.cu"), read("hitsort.cu")
    t = []
    # K1: hash
    t.append(cut(hash_cuh, "template <int KW>\nstruct KeyBuilder", "constexpr uint64_t kHashRangeAlign"))
    k = cut(hash_cu, "__device__ __noinline__ int exact_bucket_codes_cold", "// Fragment records without ranks")
    bar = 'asm volatile("bar.sync %0, %1;" ::"r"(gid + 1), "r"(GT) : "memory");'
    assert bar in k
    k = k.replace(bar, "emu_named_barrier(gid + 1, GT);")
    k = k.replace("extern __shared__ __align__(128) unsigned char smem_raw[];", "unsigned char *smem_raw = emu_dyn_smem;")
    t.append(k)
    q = hash_cu.index("__global__ void hash_queries_kernel")
    t.append(hash_cu[hash_cu.rindex("template <int KW>", 0, q):hash_cu.index("// ---- host side", q)])
    a = hash_cu.index("// std::to_string(int) strings of a bucket tuple, packed like KeyBuilder does")
    t.append(hash_cu[a:hash_cu.rindex("}  // namespace hs")])
    # K2: index
    t.append(cut(sort_cu, "constexpr int kSortThreads", "int exclusive_scan_u32("))
    t.append(cut(sort_cu, "constexpr int kRkThreads", "// Rank path of build_table_index"))
    t.append(cut(sort_cu, "constexpr uint32_t kGatherBlockBytes", "// Builds codes_sorted of every table."))
    # probe
    t.append(cut(verify_cu, "// ---- probe: query key -> bucket range", "// Hashed-key path (radix_sort.cu): binary search"))
    # K3: tensor filter
    t.append(cut(verify_cuh, "struct Survivor {", "constexpr int kFilterThreads"))
    t.append(cut(verify_cuh, "enum FilterMode", "struct FilterArgs"))
    t.append(cut(verify_cuh, "struct MmaGeometry {", "struct MmaItemHost"))
    t.append("#define HS_MMA_EVENTS 0\n")
    t.append(cut(mma_cu, "constexpr int kMmaEpiWarps = 16;", "__device__ __forceinline__ uint32_t smem_addr"))
    t.append(cut(mma_cu, "// K-major, SWIZZLE_NONE shared-memory matrix descriptor", "__device__ __forceinline__ void mma_f16_ss"))
    body = cut(mma_cu, "// Append the survivors the lanes of a warp hold in their private slots", "// One event per lane: the column test")
    body = re.sub(r'asm volatile\("tcgen05\.alloc.*?"memory"\);', "emu_tmem_alloc(&sh.tmem_base);", body, flags=re.S)
    body = re.sub(r'asm volatile\("tcgen05\.relinquish_alloc_permit.*?"memory"\);', ";", body, flags=re.S)
    body = re.sub(r'asm volatile\("tcgen05\.dealloc.*?"memory"\);', ";", body, flags=re.S)
    body = body.replace('asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");', ";")
    body = body.replace("extern __shared__ __align__(1024) unsigned char mma_smem[];", "unsigned char *mma_smem = emu_dyn_smem;")
    t.append(body)
    t.append(cut(mma_cu, "__device__ __forceinline__ bool write_cq", "// queries = DB fragments q0 .. q0+nq-1"))
    t.append(cut(mma_cu, "static inline double mma_beta(int kp) {", "// The tensor path needs every table entry representable"))
    t.append(cut(mma_cu, "int mma_upload_tables(hs_ctx *ctx) {", "int launch_build_qb_points("))
    # K4: exact stage
    t.append(cut(verify_cuh, "// ---- lock-free union-find (device)", "int launch_exact("))
    e = cut(verify_cu, "// ---- exact stage ---", "int launch_exact(hs_ctx")
    e = e.replace('asm volatile("prefetch.global.L2 [%0];" ::"l"(a.rec + (uint64_t)nx.pos * a.rec_stride));', "(void)nx;")
    e = e.replace("extern __shared__ __align__(16) unsigned char exact_smem[];", "unsigned char *exact_smem = emu_dyn_smem;")
    t.append(e)
    # hit order
    h = cut(hitsort_cu, "constexpr int kSegThreads", "static int seg_bits_for")
    for name, decl in (("seg_buf", "extern __shared__ __align__(16) uint32_t seg_buf[];"),
                       ("seg_cnt", "extern __shared__ uint32_t seg_cnt[];"), ("seg_cur", "extern __shared__ uint32_t seg_cur[];")):
        h = h.replace(decl, f"uint32_t *{name} = reinterpret_cast<uint32_t *>(emu_dyn_smem);")
    t.append(h)
    text = "\n".join(t)
    assert "asm" not in text and "<<<" not in text and "extern __shared__" not in text
    return text


@pytest.mark.skipif(os.uname().machine != "x86_64", reason="the emulation's fiber switch is x86-64 assembly")
def test_whole_hot_path_under_cpu_emulation(tmp_path):
    (tmp_path / "pipeline_kernels.inc").write_text(kernel_text())
    obj = tmp_path / "hs_oracle.o"
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-D_GNU_SOURCE", "-c",
                           os.path.join(ROOT, "oracle", "hs_oracle.c"), "-o", str(obj)])
    exe = tmp_path / "pipeline_emu"
    emu = os.path.join(ROOT, "tests", "emu")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-DHS_GATHER_PART=256", f"-I{tmp_path}",
                           f"-I{os.path.join(emu, 'stub')}", f"-I{emu}", "-o", str(exe),
                           os.path.join(emu, "pipeline_emu.cpp"), str(obj), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=1800)
    assert out.returncode == 0, out.stdout + out.stderr
    results = re.findall(r" -> (\w+)$", out.stdout, flags=re.M)
    assert len(results) >= 2 and all(r == "ok" for r in results), out.stdout
