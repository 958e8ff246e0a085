set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_rank.log 2>&1; echo all_rc=$?
tail -15 gpurun_out/t_rank.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b3.log 2> gpurun_out/b3.err; echo bench_rc=$?
tail -3 gpurun_out/b3.err
tail -1 gpurun_out/b3.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stages_ms'], d['kernels'], d['counts'])"
