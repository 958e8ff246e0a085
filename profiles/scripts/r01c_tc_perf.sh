set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo all_rc=$?
tail -4 gpurun_out/t_all.log
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/b_tc.log 2>&1; rc=$?; echo bench_rc=$rc
tail -1 gpurun_out/b_tc.log | head -c 2500
if [ $rc -eq 0 ]; then
for K in filter_tc_kernel exact_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_$K $CMD > gpurun_out/ncu_$K.log 2>&1
  echo $K rc=$?
done
fi
