set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "blocked_gather or index or search" > gpurun_out/t_gather.log 2>&1; echo rc=$?
tail -12 gpurun_out/t_gather.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b19.log 2> gpurun_out/b19.err; echo rc=$?
tail -2 gpurun_out/b19.err
tail -1 gpurun_out/b19.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stages_ms'], d['counts']['hits_total'])"
