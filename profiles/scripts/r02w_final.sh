#!/bin/bash
# round 2 final verification on one GPU: whole suite, smoke, reference arm, full bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 1800 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/r02w_tests.log 2>&1; echo "suite rc=$?"
tail -5 gpurun_out/r02w_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02w_bench_reference.json 2> gpurun_out/r02w_ref.err; echo "ref rc=$?"
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02w_bench_full.json 2> gpurun_out/r02w_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r02w_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02w_bench_full.json').read().strip().splitlines()[-1])
r = json.loads(open('gpurun_out/r02w_bench_reference.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, 'ref', r['value'])
print('e2e', {k: d['e2e'][k] for k in ('ms_per_step', 'sequential_ms_per_step', 'value', 'hits_equal_device_run_after_expansion')})
print('roofline', json.dumps({k: d['roofline'][k] for k in ('kernel', 'achieved', 'peak', 'frac', 'traffic')}))
print('checks', json.dumps(d['checks']))
print('counts', json.dumps(d['counts']))
print('cpu', json.dumps(d['cpu_baseline'])[:300])
print('recall', json.dumps(d['recall'])[:300])
print('stages', json.dumps(d['stages_ms']))
PY
