set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sequence.py tests/test_cli.py -x -q -m gpu > gpurun_out/t_cov.log 2>&1; echo cov_rc=$?
tail -30 gpurun_out/t_cov.log
