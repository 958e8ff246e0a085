set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --n-db 20000000"
timeout 300 $CMD > gpurun_out/plain_mma.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_mma_kernel -s 1 -c 1 -f -o gpurun_out/prof_filter_mma_v3 $CMD > gpurun_out/ncu_filter_mma_v3.log 2>&1
echo ncu_rc=$?
