#!/bin/bash
# round 2: integer metric on the pipelined tensor filter, replicated pair table in the exact stage,
# lazy code stores + filter bypass; GPU suite, C5 length sweep, K = 16 sweep points, bench A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r02aa_tests.log 2>&1; echo "tests rc=$?"
tail -25 gpurun_out/r02aa_tests.log
timeout 600 python profiles/scripts/configs_bench.py allpairs 16,32,20 16,4,50 8,16,20 > gpurun_out/r02aa_configs.jsonl 2> gpurun_out/r02aa_configs.err; echo "configs rc=$?"
cat gpurun_out/r02aa_configs.jsonl; tail -c 600 gpurun_out/r02aa_configs.err
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-recall --no-subset-check"
timeout 600 $B > gpurun_out/r02aa_bench_rep8.json 2> gpurun_out/r02aa_bench.err; echo "bench rc=$?"
HS_EXACT_REP=0 timeout 600 $B --no-e2e > gpurun_out/r02aa_bench_rep1.json 2>> gpurun_out/r02aa_bench.err; echo "bench rep1 rc=$?"
tail -c 600 gpurun_out/r02aa_bench.err
python - <<'PY'
import json
for f in ('gpurun_out/r02aa_bench_rep8.json', 'gpurun_out/r02aa_bench_rep1.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d[k] for k in ('value', 'ms_per_step')})
        if d.get('e2e'): print(' e2e', {k: d['e2e'].get(k) for k in ('ms_per_step', 'sequential_ms_per_step', 'hits_equal_device_run_after_expansion')})
        print(' stages', json.dumps(d['stages_ms']))
        print(' checks', json.dumps(d['checks'])[:300])
    except Exception as e:
        print(f, 'parse failed', e)
PY
