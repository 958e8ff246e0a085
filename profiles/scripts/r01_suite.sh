set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_final.log 2>&1; echo tests_rc=$?
tail -4 gpurun_out/t_final.log
