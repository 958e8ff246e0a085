#!/bin/bash
# round 2 (8 GPUs): host copy ceiling with 8 ranks, then the bench at 125 M fragments per GPU = the 1 B of north_star
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
nproc; free -g | head -2; nvidia-smi topo -m | head -12
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29521 profiles/scripts/pcie_ceiling.py 2>/dev/null | tail -1 | tee gpurun_out/r02q_pcie8.json
timeout 120 python profiles/scripts/pcie_ceiling.py 2>/dev/null | tail -1 | tee gpurun_out/r02q_pcie1.json
timeout 900 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --n-db 125000000 --no-subset-check --no-recall > gpurun_out/r02q_bench_n8_1B.json 2> gpurun_out/r02q_bench_n8.err; echo "bench n8 rc=$?"
tail -c 800 gpurun_out/r02q_bench_n8.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r02q_bench_n8_1B.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')})
    print(' e2e', {k: d['e2e'][k] for k in ('ms_per_step', 'sequential_ms_per_step', 'value')})
    print(' multi', json.dumps(d.get('multi_gpu_checks')))
    print(' stages', json.dumps(d['stages_ms']))
    print(' clocks', json.dumps(d['clocks']))
except Exception as e:
    print('parse failed', e)
PY
