#!/bin/bash
# round 2: whole GPU suite, N = 1 bench and (2 GPUs) N = 2 bench after the hash / gather / filter epilogue changes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q --maxfail=5 -k "not c1_one" > gpurun_out/r02n_tests.log 2>&1; echo "suite rc=$?"
tail -12 gpurun_out/r02n_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-subset-check --no-recall > gpurun_out/r02n_bench_n2.json 2> gpurun_out/r02n_bench_n2.err; echo "bench n2 rc=$?"
tail -c 600 gpurun_out/r02n_bench_n2.err
timeout 900 python bench.py --steps 10 --warmup 3 --no-subset-check --no-recall --no-cpu-baseline > gpurun_out/r02n_bench_n1.json 2> gpurun_out/r02n_bench_n1.err; echo "bench n1 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r02n_bench_n2.json', 'gpurun_out/r02n_bench_n1.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')})
        print(' e2e', {k: d['e2e'][k] for k in ('ms_per_step', 'sequential_ms_per_step', 'value')})
        print(' multi', json.dumps(d.get('multi_gpu_checks')))
        print(' stages', json.dumps(d['stages_ms']))
    except Exception as e:
        print(f, 'parse failed', e)
PY
