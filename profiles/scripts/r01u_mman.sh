set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu -k "tensor_filter or bruteforce or search or rank_path" > gpurun_out/t_mma3.log 2>&1; echo mma_rc=$?
tail -3 gpurun_out/t_mma3.log
timeout 600 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/b13.log 2> gpurun_out/b13.err; echo rc=$?
tail -1 gpurun_out/b13.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N128', d['ms_per_step'], d['stages_ms']['filter_tc'])"
