set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "long_fragments" > gpurun_out/t_long.log 2>&1; echo rc=$?
tail -12 gpurun_out/t_long.log
