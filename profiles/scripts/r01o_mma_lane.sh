set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu -k "tensor_filter or bruteforce or search or rank_path" > gpurun_out/t_mma2.log 2>&1; echo mma_rc=$?
tail -4 gpurun_out/t_mma2.log
timeout 600 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/b6.log 2> gpurun_out/b6.err; echo rc=$?
tail -1 gpurun_out/b6.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stages_ms'], d['counts'])"
HS_NVCC_EXTRA=-DHS_MMA_PROF python hsearch_b200/build.py --force > gpurun_out/build_prof.log 2>&1; echo build_rc=$?
timeout 600 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/b7.log 2> gpurun_out/b7.err; echo rc=$?
grep "mma prof" gpurun_out/b7.err | tail -1
