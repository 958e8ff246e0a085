set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "blocked_gather" > gpurun_out/t_gather.log 2>&1; echo rc=$?
tail -5 gpurun_out/t_gather.log
for MB in 4 8 16; do
HS_GATHER_MB=$MB timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/b20_$MB.log 2> gpurun_out/b20_$MB.err; echo rc=$?
tail -1 gpurun_out/b20_$MB.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('MB=$MB', d['ms_per_step'], d['stages_ms']['permute'], d['counts']['hits_total'])"
done
