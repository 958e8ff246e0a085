#!/bin/bash
# round 2, call 2 (2 GPUs): the library's multi-GPU merge against the one-GPU search, the remaining
# new tests, the N = 2 bench line, and the N = 1 bench with kernel-driven control transfers
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
nvidia-smi --query-gpu=index,name --format=csv,noheader
nvidia-smi topo -m | head -6
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_baseline.py -q -x --durations=5 -k "not c1_one" > gpurun_out/r02b_tests.log 2>&1; echo "tests rc=$?"
tail -20 gpurun_out/r02b_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-subset-check > gpurun_out/r02b_bench_n2.json 2> gpurun_out/r02b_bench_n2.err; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/r02b_bench_n2.err
timeout 900 python bench.py --steps 10 --warmup 3 --no-subset-check --no-recall --no-cpu-baseline > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench n1 rc=$?"
tail -c 600 gpurun_out/r02b_bench_n1.err
python - <<'PY'
import json
for f in ('gpurun_out/r02b_bench_n2.json', 'gpurun_out/r02b_bench_n1.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')})
        print(' e2e', {k: d['e2e'][k] for k in ('ms_per_step', 'sequential_ms_per_step', 'value')})
        print(' multi', json.dumps(d.get('multi_gpu_checks')))
        print(' stages', json.dumps(d['stages_ms']))
        print(' checks', json.dumps(d['checks'])[:400])
    except Exception as e:
        print(f, 'parse failed', e)
PY
