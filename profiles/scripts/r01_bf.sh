set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bruteforce" > gpurun_out/t_bf.log 2>&1; echo rc=$?
tail -8 gpurun_out/t_bf.log
