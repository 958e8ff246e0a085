#!/bin/bash
# round 2, last GPU seconds: the segmented hit sort with the per-bin bucket sort -- its tests, the bench with
# per-kernel times (buckets / radix passes per bin), a full bench line with the exact slice comparison, and the
# GPU suite with every hit list forced through it for as long as the budget lasts.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
T0=$(date +%s)
timeout 60 python -m pytest tests/test_gpu_segsort.py -q -x -s > gpurun_out/r02al_segsort_tests.log 2>&1; echo "segsort tests rc=$? t=$(( $(date +%s) - T0 ))"
grep -E "segsort \(lists|passed|failed|Error|assert" gpurun_out/r02al_segsort_tests.log | head -12
for RX in 0 1; do
  HS_SEGSORT=1 HS_SEGSORT_PROF=1 HS_SEGSORT_RADIX=$RX timeout 60 python bench.py --steps 3 --warmup 2 --no-e2e --no-recall --no-cpu-baseline --no-subset-check \
    > gpurun_out/r02al_bench_rx$RX.json 2> gpurun_out/r02al_bench_rx$RX.err
  echo "radix-per-bin=$RX rc=$? t=$(( $(date +%s) - T0 ))"
  grep "^segsort:" gpurun_out/r02al_bench_rx$RX.err | tail -1
  python -c "
import json
d = json.loads(open('gpurun_out/r02al_bench_rx$RX.json').read().strip().splitlines()[-1])
print('  hitsort', d['stages_ms']['hitsort'], 'step', round(d['ms_per_step'], 2), 'order', d['checks']['reference_order'])"
done
HS_SEGSORT=1 HS_SEGSORT_PROF=1 timeout 120 python bench.py --steps 10 --warmup 3 --no-recall --no-cpu-baseline > gpurun_out/r02al_bench_full_seg.json 2> gpurun_out/r02al_bench_full_seg.err
echo "full bench (segsort) rc=$? t=$(( $(date +%s) - T0 ))"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02al_bench_full_seg.json").read().strip().splitlines()[-1])
    print({k: round(d[k], 3) for k in ("value", "ms_per_step")}, "e2e", round(d["e2e"]["ms_per_step"], 2), d["e2e"].get("hits_equal_device_run_after_expansion"))
    print("  e2e stages", json.dumps(d["e2e"].get("search_stages_ms")))
    print("  stages", json.dumps(d["stages_ms"]))
    print("  checks", json.dumps(d["checks"])[:500])
except Exception as e:
    print("parse failed", e)
PY
grep "^segsort:" gpurun_out/r02al_bench_full_seg.err | tail -5
LEFT=$(( 228 - ( $(date +%s) - T0 ) ))
echo "left for the suite: $LEFT s"
if [ "$LEFT" -gt 30 ]; then
  HS_SEGSORT=1 HS_SEGSORT_MIN=0 timeout $LEFT python -m pytest tests -m gpu -q -x -p no:cacheprovider --deselect tests/test_gpu_segsort.py > gpurun_out/r02al_suite_seg.log 2>&1; echo "suite (segsort forced) rc=$? (124 = out of time) t=$(( $(date +%s) - T0 ))"
  tail -4 gpurun_out/r02al_suite_seg.log
fi
