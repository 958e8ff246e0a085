set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b2.log 2> gpurun_out/b2.err; echo bench_rc=$?
tail -1 gpurun_out/b2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stages_ms'], d['kernels'])"
