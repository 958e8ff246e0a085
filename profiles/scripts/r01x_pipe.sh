set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipelined or search" > gpurun_out/t_pipe.log 2>&1; echo rc=$?
tail -12 gpurun_out/t_pipe.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/b16.log 2> gpurun_out/b16.err; echo rc=$?
tail -2 gpurun_out/b16.err
tail -1 gpurun_out/b16.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['cpu_baseline'], d['roofline']['frac'], d['kernels'])"
