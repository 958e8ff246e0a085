# Round-2 final profile capture (run under gpurun on one B200; outputs land in gpurun_out/ and are
# summarised into profiles/ by `python profiles/summarize.py r02b`).
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches.csv
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-recall --no-subset-check"
# 1. the command exits 0 without ncu
$CMD > gpurun_out/plain.log 2>&1; echo plain_rc=$?
# 2. launch list: per-launch durations (cold cache, serialised -- compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo list_rc=$?
# 3. one full capture per hot kernel (second launch of each: the first is the sizing pass)
for K in filter_mma_kernel exact_kernel gather_blocked_kernel hash_fast_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_$K $CMD > gpurun_out/ncu_$K.log 2>&1
  echo $K rc=$?
done
tail -1 gpurun_out/plain.log | head -c 300
