#!/bin/bash
# round 2, final call (one GPU, <= 7.5 minutes): segmented hit sort -- its tests again (the large-bin test now
# prints what happened), the partition's block count swept with per-kernel times, a full bench line of the best
# setting if it beats the radix passes (7.9 ms), and the default GPU suite with whatever time is left.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
T0=$(date +%s)
timeout 120 python -m pytest tests/test_gpu_segsort.py -q -x -s > gpurun_out/r02ak_segsort_tests.log 2>&1; echo "segsort tests rc=$? t=$(( $(date +%s) - T0 ))"
grep -E "segsort \(lists|passed|failed|Error|assert" gpurun_out/r02ak_segsort_tests.log | head -12
for CFG in 296:1024 48:512 12:1024 96:512; do
  NB=${CFG%%:*}; NT=${CFG##*:}
  HS_SEGSORT=1 HS_SEGSORT_PROF=1 HS_SEGSORT_NBLK=$NB HS_SEGSORT_THREADS=$NT timeout 120 python bench.py --steps 3 --warmup 2 --no-e2e --no-recall --no-cpu-baseline --no-subset-check \
    > gpurun_out/r02ak_bench_nblk$NB.json 2> gpurun_out/r02ak_bench_nblk$NB.err
  echo "nblk=$NB threads=$NT rc=$? t=$(( $(date +%s) - T0 ))"
  grep "^segsort:" gpurun_out/r02ak_bench_nblk$NB.err | tail -2
done
BEST=$(python - <<'PY'
import json, re
part, sort = {}, {}
for nb, nt in ((296, 1024), (48, 512), (12, 1024), (96, 512)):
    try:
        d = json.loads(open(f"gpurun_out/r02ak_bench_nblk{nb}.json").read().strip().splitlines()[-1])
        lines = [l for l in open(f"gpurun_out/r02ak_bench_nblk{nb}.err") if l.startswith("segsort:")]
        m = re.search(r"hist ([\d.]+)  scan ([\d.]+)  scatter ([\d.]+)  sort ([\d.]+) ms  flags (\d+) (\d+)", lines[-1])
        h, sc, st, so = (float(m.group(i)) for i in range(1, 5))
        ok = d["checks"]["reference_order"] and m.group(5) == "0" and m.group(6) == "0"
        print(f"# nblk={nb} threads={nt}: hitsort {d['stages_ms']['hitsort']:.3f} ms (hist {h} scan {sc} scatter {st} sort {so}) step {d['ms_per_step']:.2f} order_ok {ok}", flush=True)
        if ok:
            part[nb] = h + sc + st
            sort[nt] = min(sort.get(nt, 1e9), so)
    except Exception as e:
        print(f"# nblk={nb} failed: {e}")
if part and sort:
    nb = min(part, key=part.get)
    nt = min(sort, key=sort.get)
    print(f"# predicted best: nblk={nb} threads={nt}: {part[nb] + sort[nt]:.3f} ms against 7.9 for the radix passes")
    print(f"{nb}:{nt}" if part[nb] + sort[nt] < 7.3 else "")
else:
    print("")
PY
)
echo "$BEST" | grep "^#"
BEST=$(echo "$BEST" | tail -1)
echo "best: '$BEST'"
if [ -n "$BEST" ]; then
  NB=${BEST%%:*}; NT=${BEST##*:}
  HS_SEGSORT=1 HS_SEGSORT_PROF=1 HS_SEGSORT_NBLK=$NB HS_SEGSORT_THREADS=$NT timeout 200 python bench.py --steps 10 --warmup 3 --no-recall --no-cpu-baseline > gpurun_out/r02ak_bench_full_seg.json 2> gpurun_out/r02ak_bench_full_seg.err
  echo "full bench (segsort, nblk=$NB threads=$NT) rc=$? t=$(( $(date +%s) - T0 ))"
  python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02ak_bench_full_seg.json").read().strip().splitlines()[-1])
    print({k: round(d[k], 3) for k in ("value", "ms_per_step")}, "e2e", round(d["e2e"]["ms_per_step"], 2), d["e2e"].get("hits_equal_device_run_after_expansion"))
    print("  e2e stages", json.dumps(d["e2e"].get("search_stages_ms")))
    print("  stages", json.dumps(d["stages_ms"]))
    print("  checks", json.dumps(d["checks"])[:500])
except Exception as e:
    print("parse failed", e)
PY
  grep "^segsort:" gpurun_out/r02ak_bench_full_seg.err | tail -6
fi
LEFT=$(( 425 - ( $(date +%s) - T0 ) ))
echo "left for the suite: $LEFT s"
if [ "$LEFT" -gt 60 ]; then
  timeout $LEFT python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r02ak_suite.log 2>&1; echo "default suite rc=$? t=$(( $(date +%s) - T0 ))"
  tail -4 gpurun_out/r02ak_suite.log
fi
