set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/b23.log 2> gpurun_out/b23.err; echo rc=$?
tail -3 gpurun_out/b23.err
tail -1 gpurun_out/b23.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['recall'], d['checks'])"
