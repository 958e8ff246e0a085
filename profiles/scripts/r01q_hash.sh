set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu -k "hash or index or rank_path" > gpurun_out/t_hash.log 2>&1; echo hash_rc=$?
tail -4 gpurun_out/t_hash.log
timeout 600 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/b10.log 2> gpurun_out/b10.err; echo rc=$?
tail -1 gpurun_out/b10.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stages_ms'])"
