set -x
mkdir -p gpurun_out
timeout 900 python bench.py --no-recall --no-cpu-baseline --no-e2e > gpurun_out/b24.log 2> gpurun_out/b24.err; echo rc=$?
tail -2 gpurun_out/b24.err
tail -1 gpurun_out/b24.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['clocks'])"
