"""Build libhsearch_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so
travels with the repo snapshot to the GPU box)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhsearch_b200.so")
CU = ["api.cu", "hash.cu", "radix_sort.cu", "verify.cu", "filter_tc.cu", "filter_mma.cu", "cluster.cu", "extract.cu", "sequence.cu", "evaluate.cu", "hits.cu", "comm.cu", "fasta.cu", "hitsort.cu"]
CPP = ["tables.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, ptxas_verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "hsearch_b200.h"))
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in CU + CPP:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src + ".o")
        objs.append(o)
        if force or _newer(o, [s] + hdrs):
            extra = os.environ.get("HS_NVCC_EXTRA", "").split()
            cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if ptxas_verbose else []) + ["-x", "cu", "-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd))
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- {src} ---\n{out}\n")
        elif (verbose or ptxas_verbose) and out.strip():
            print(f"--- {src} ---\n{out}")
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _newer(LIB, objs):
        # -cudart shared: the library binds to libcudart.so.12 at load time (the copy torch already
        # mapped, or /usr/local/cuda/lib64 for the CLI programs) instead of embedding a static one
        cmd = [nvcc, "-shared", "-cudart", "shared", "-o", LIB] + objs + [
            "-gencode", "arch=compute_100a,code=sm_100a", "-ldl", "-Xlinker", "-rpath,/usr/local/cuda/lib64"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    build_cli(force=force or bool(procs), verbose=verbose)
    return LIB


CLI = ["motif_both_points", "motif_both_points_noLSH", "protein2datapoints", "hclust2", "hclust3", "evaluate2", "orf", "pcluster"]
BIN = os.path.join(HERE, "bin")


def build_cli(force=False, verbose=False):
    """The drop-in command-line programs (host C++ over the C ABI): hsearch_b200/bin/<name>."""
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    src_dir = os.path.join(HERE, "cli")
    hdrs = [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith(".hpp")]
    hdrs.append(os.path.join(HERE, "..", "include", "hsearch_b200.h"))
    os.makedirs(BIN, exist_ok=True)
    procs = []
    for name in CLI:
        # hclust3 is hclust2 with one more progress line (hclust3.cpp:77)
        src = os.path.join(src_dir, ("hclust2" if name == "hclust3" else name) + ".cpp")
        out = os.path.join(BIN, name)
        if force or _newer(out, [src, LIB] + hdrs):
            cmd = [cxx, "-O2", "-std=c++17", "-Wall"] + (["-DHCLUST3"] if name == "hclust3" else []) + [
                   "-o", out, src, "-L" + HERE, "-lhsearch_b200", "-pthread",
                   "-Wl,-rpath,$ORIGIN/..", "-Wl,-rpath," + "/usr/local/cuda/lib64"]
            if verbose:
                print(" ".join(cmd))
            procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for name, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- {name} ---\n{out}\n")
        elif verbose and out.strip():
            print(f"--- {name} ---\n{out}")
    if failed:
        raise RuntimeError("g++ failed building the CLI programs")
    return [os.path.join(BIN, n) for n in CLI]


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True, ptxas_verbose="--ptxas" in sys.argv)
    print(LIB)
