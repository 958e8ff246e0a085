set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
HS_GATHER_MB=16 ncu --set full --clock-control none --import-source on -k regex:gather_blocked_kernel -s 1 -c 1 -f -o gpurun_out/prof_gather_blocked_kernel $CMD > gpurun_out/ncu_gather.log 2>&1; echo rc=$?
HS_GATHER_MB=16 ncu --set full --clock-control none -k regex:gather_runs_kernel -s 1 -c 1 -f -o gpurun_out/prof_gather_runs_kernel $CMD > gpurun_out/ncu_gather2.log 2>&1; echo rc=$?
