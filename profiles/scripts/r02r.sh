#!/bin/bash
# round 2: survivor binning by fragment-id block (exact stage locality): tests, then N = 1 bench with and without
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q --maxfail=5 -k "not c1_one and not two_gpu" > gpurun_out/r02r_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r02r_tests.log
for nb in 0 1; do
HS_NO_SURV_BINS=$nb timeout 600 python bench.py --steps 10 --warmup 3 --no-subset-check --no-recall --no-cpu-baseline > gpurun_out/r02r_bench_nobins$nb.json 2> gpurun_out/r02r_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/r02r_bench_nobins$nb.json').read().strip().splitlines()[-1])
print('no_bins=$nb', {k: d[k] for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['ms_per_step'], d['e2e']['sequential_ms_per_step'])
print(' stages', json.dumps(d['stages_ms']))
print(' e2e stages', json.dumps(d['e2e']['search_stages_ms']))
PY
done
