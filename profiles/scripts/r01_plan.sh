set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_cli.py -x -q -m gpu -k "search or pipelined or tensor" > gpurun_out/t_p.log 2>&1; echo rc=$?
tail -3 gpurun_out/t_p.log
HS_PLAN_STATS=1 timeout 600 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-recall > gpurun_out/b27.log 2> gpurun_out/b27.err; echo rc=$?
grep "host planning" gpurun_out/b27.err | tail -1
tail -1 gpurun_out/b27.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stages_ms'], d['counts']['hits_total'])"
