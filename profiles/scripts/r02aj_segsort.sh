#!/bin/bash
# round 2, last call (one GPU, ~9 minutes): the segmented hit sort on hardware -- its parity tests, the bench
# line with it on (incl. the exact slice comparison with the reference) and off, then the GPU suite with every
# hit list forced through it (HS_SEGSORT_MIN=0) for as long as the budget lasts.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
T0=$(date +%s)
timeout 150 python -m pytest tests/test_gpu_segsort.py -q -x --durations=5 > gpurun_out/r02aj_segsort_tests.log 2>&1; echo "segsort tests rc=$? t=$(( $(date +%s) - T0 ))"
tail -15 gpurun_out/r02aj_segsort_tests.log
HS_SEGSORT=1 timeout 240 python bench.py --steps 5 --warmup 3 --no-recall --no-cpu-baseline > gpurun_out/r02aj_bench_seg.json 2> gpurun_out/r02aj_bench_seg.err; echo "bench seg rc=$? t=$(( $(date +%s) - T0 ))"
tail -c 300 gpurun_out/r02aj_bench_seg.err
HS_SEGSORT=0 timeout 180 python bench.py --steps 5 --warmup 3 --no-recall --no-cpu-baseline --no-subset-check > gpurun_out/r02aj_bench_radix.json 2> gpurun_out/r02aj_bench_radix.err; echo "bench radix rc=$? t=$(( $(date +%s) - T0 ))"
python - <<'PY'
import json
for tag in ("seg", "radix"):
    try:
        d = json.loads(open(f"gpurun_out/r02aj_bench_{tag}.json").read().strip().splitlines()[-1])
        print(tag, {k: round(d[k], 3) for k in ("value", "ms_per_step")}, "e2e", round(d["e2e"]["ms_per_step"], 2),
              d["e2e"].get("hits_equal_device_run_after_expansion"), "e2e stages", json.dumps(d["e2e"].get("search_stages_ms")))
        print("  stages", json.dumps(d["stages_ms"]))
        print("  checks", json.dumps(d["checks"])[:600])
        print("  counts", json.dumps(d["counts"])[:400])
    except Exception as e:
        print(tag, "parse failed", e)
PY
LEFT=$(( 520 - ( $(date +%s) - T0 ) ))
if [ "$LEFT" -gt 40 ]; then
  HS_SEGSORT=1 HS_SEGSORT_MIN=0 timeout $LEFT python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_segsort.py -p no:cacheprovider > gpurun_out/r02aj_suite_seg.log 2>&1; echo "suite (segsort forced) rc=$? t=$(( $(date +%s) - T0 ))"
  tail -6 gpurun_out/r02aj_suite_seg.log
fi
