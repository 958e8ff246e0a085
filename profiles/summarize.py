"""Turn gpurun_out/launches.csv and gpurun_out/prof_*.ncu-rep into the tracked
summaries under profiles/ (run here, no GPU needed):
    python profiles/summarize.py r01a
"""
import collections
import csv
import glob
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size", "lts__t_bytes.sum",
        # tensor pipe activity (whatever of these the sm_100 ncu exposes)
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.sum",
        "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.sum", "smsp__inst_executed_pipe_tmem.sum"]


TENSOR = ["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__ops_path_tensor_src_fp16_dst_fp32.sum", "sm__ops_path_tensor_src_fp16_dst_fp32.sum.pct_of_peak_sustained_elapsed",
          "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.sum",
          "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active"]


def launches(tag):
    path = os.path.join(ROOT, "gpurun_out", "launches.csv")
    if not os.path.exists(path):
        return
    lines = open(path).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    agg, tot = collections.OrderedDict(), 0.0
    for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r["Metric Unit"]]
        a = agg.setdefault(r["Kernel Name"].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n"
                "Command: see profiles/{tag[:3]}_ncu_commands.sh.  Times are cold-cache and serialised: compare SHARES.\n\n"
                f"total {tot:.1f} ms over {sum(a[0] for a in agg.values())} launches\n\n"
                "| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:80]}` | {n} | {ms:.3f} | {100 * ms / tot:.1f}% |\n")


def full(tag):
    out = {}
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "prof_*.ncu-rep"))):
        name = os.path.basename(rep)[5:-8]
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        if len(rows) < 3:
            continue
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                d[w] = f"{vals[i]} {units[i]}".strip()
        for i, name_i in enumerate(hdr):   # tensor-pipe / TMEM activity
            if name_i in TENSOR and i < len(vals) and vals[i] not in ("0", ""):
                d[name_i] = f"{vals[i]} {units[i]}".strip()
        out[name] = d
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full.json"), "w") as f:
        json.dump(out, f, indent=1)
    # DRAM bytes (read + write) per launch of every captured kernel: bench.py copies the dominant
    # kernel's entry into roofline.traffic
    traffic = {}
    for k, d in out.items():
        def gb(x):
            v, u = x.split()[:2]
            return float(v.replace(",", "")) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
        if "dram__bytes_read.sum" in d and "dram__bytes_write.sum" in d:
            traffic[k] = gb(d["dram__bytes_read.sum"]) + gb(d["dram__bytes_write.sum"])
    if traffic:
        with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)
    return out


def sass_histogram(tag):
    """Opcode histogram of the dominant kernel's SASS (cuobjdump of the built library, no GPU needed)."""
    lib = os.path.join(ROOT, "hsearch_b200", "libhsearch_b200.so")
    fun = "_ZN2hs17filter_mma_kernelILi10EEEvNS_7MmaArgsE"
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, lib], capture_output=True, text=True).stdout
    hist = collections.Counter()
    for line in txt.splitlines():
        parts = line.split("*/")
        if len(parts) < 2 or "/*" not in parts[0]:
            continue
        ins = parts[1].strip().split(";")[0].split()
        if not ins:
            continue
        op = ins[1] if ins[0].startswith("@") and len(ins) > 1 else ins[0]
        hist[op.split(".")[0]] += 1
    if not hist:
        return
    with open(os.path.join(ROOT, "profiles", f"{tag}_filter_sass_histogram.md"), "w") as f:
        f.write(f"# {tag}: SASS opcode histogram of `filter_mma_kernel<10>` (cuobjdump -sass of libhsearch_b200.so)\n\n"
                f"{sum(hist.values())} instructions.  UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, "
                "UBLKCP = cp.async.bulk (TMA engine), SYNCS = mbarrier operations.\n\n| opcode | count |\n|---|---:|\n")
        for op, n in hist.most_common():
            f.write(f"| `{op}` | {n} |\n")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    launches(tag)
    sass_histogram(tag)
    print(json.dumps(full(tag), indent=1)[:400])
