set -x
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_filter or bruteforce" > gpurun_out/t_mma.log 2>&1; echo mma_rc=$?
tail -25 gpurun_out/t_mma.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo all_rc=$?
tail -8 gpurun_out/t_all.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/b_mma.log 2>&1; echo bench_rc=$?
tail -1 gpurun_out/b_mma.log | head -c 3000
