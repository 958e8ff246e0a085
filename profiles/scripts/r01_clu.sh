set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_dist_gloo.py -x -q -m gpu -k "cluster" > gpurun_out/t_clu.log 2>&1; echo rc=$?
tail -14 gpurun_out/t_clu.log
