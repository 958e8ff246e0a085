#!/bin/bash
# round 2, call 1: new parity tests (C1 at 1 M x 1 k vs the reference), whole GPU suite, bench with the
# two-context end-to-end leg, the audit and the exact subset check
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_baseline.py -q -x --durations=8 > gpurun_out/r02a_new_tests.log 2>&1; echo "new tests rc=$?"
tail -25 gpurun_out/r02a_new_tests.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02a_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r02a_bench.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step')})
    print('e2e', json.dumps(d['e2e'])[:900])
    print('checks', json.dumps(d['checks']))
    print('counts', json.dumps(d['counts']))
    print('stages', json.dumps(d['stages_ms']))
except Exception as e:
    print('parse failed', e)
PY
timeout 1500 python -m pytest tests -m gpu -q --maxfail=5 --deselect tests/test_gpu_baseline.py > gpurun_out/r02a_tests.log 2>&1; echo "suite rc=$?"
tail -15 gpurun_out/r02a_tests.log
