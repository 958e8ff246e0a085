set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "evaluate or capi or export" > gpurun_out/t_eval.log 2>&1; echo tests_rc=$?
tail -15 gpurun_out/t_eval.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b_eval.log 2> gpurun_out/b_eval.err; echo rc=$?
tail -1 gpurun_out/b_eval.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], json.dumps(d['recall']))"
tail -3 gpurun_out/b_eval.err
