set -x
timeout 240 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_filter" > gpurun_out/t_tc.log 2>&1; echo tc_rc=$?
tail -15 gpurun_out/t_tc.log
HS_TC_SWAP_LBO_SBO=1 timeout 240 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_filter" > gpurun_out/t_tc_swap.log 2>&1; echo tc_swap_rc=$?
tail -5 gpurun_out/t_tc_swap.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo all_rc=$?
tail -5 gpurun_out/t_all.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/b_tc.log 2>&1; echo bench_rc=$?
tail -1 gpurun_out/b_tc.log | head -c 3000
