set -x
mkdir -p gpurun_out
for MB in 16 32; do
HS_GATHER_MB=$MB timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/b20_$MB.log 2> gpurun_out/b20_$MB.err; echo rc=$?
tail -1 gpurun_out/b20_$MB.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('MB=$MB', d['ms_per_step'], d['stages_ms']['permute'])"
done
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:gather_blocked_kernel -s 1 -c 1 -f -o gpurun_out/prof_gather_blocked_kernel $CMD > gpurun_out/ncu_gather.log 2>&1; echo rc=$?
