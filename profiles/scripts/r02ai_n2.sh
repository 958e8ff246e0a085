#!/bin/bash
# round 2, final state: N = 2 bench line (100 M fragments per GPU) and the 2-GPU tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r02ai_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02ai_tests.log
timeout 900 $TR --nproc-per-node 2 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 --no-subset-check --no-recall --no-cpu-baseline > gpurun_out/r02ai_bench_n2.json 2> gpurun_out/r02ai_bench_n2.err; echo "bench n2 rc=$?"
tail -c 400 gpurun_out/r02ai_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r02ai_bench_n2.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, 'e2e', d['e2e']['ms_per_step'])
    print(' multi', json.dumps(d.get('multi_gpu_checks')))
    print(' stages', json.dumps(d['stages_ms']))
except Exception as e:
    print('parse failed', e)
PY
