#!/bin/bash
# round 2, final verification on one B200: GPU suite, smoke, reference arm, full bench line, ncu captures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/r02ah_tests.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/r02ah_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ah_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02ah_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02ah_bench_reference.json 2> gpurun_out/r02ah_ref.err; echo "reference rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02ah_bench_full.json 2> gpurun_out/r02ah_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r02ah_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r02ah_bench_full.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'])
    print(' roofline', d['roofline']['frac'], d['roofline']['avg_launch_ms'])
    print(' stages', json.dumps(d['stages_ms']))
    print(' checks', json.dumps(d['checks'])[:500], d['counts'].get('residual_flips'))
    r = json.loads(open('gpurun_out/r02ah_bench_reference.json').read().strip().splitlines()[-1])
    print(' reference', r['value'], r['cpu_baseline']['cores'])
except Exception as e:
    print('parse failed', e)
PY
bash profiles/r02b_ncu_commands.sh > gpurun_out/r02ah_ncu.log 2>&1; grep "rc=" gpurun_out/r02ah_ncu.log
