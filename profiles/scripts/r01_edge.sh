set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_sequence.py -x -q -m gpu -k "greedy or pipelined or edge_cases" > gpurun_out/t_edge.log 2>&1; echo rc=$?
tail -15 gpurun_out/t_edge.log
