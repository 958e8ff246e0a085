set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu -k "index or rank_path or search or hash" > gpurun_out/t_b.log 2>&1; echo rc=$?
tail -3 gpurun_out/t_b.log
timeout 600 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-recall > gpurun_out/b26.log 2> gpurun_out/b26.err; echo rc=$?
tail -1 gpurun_out/b26.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stages_ms'])"
