set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo all_rc=$?
tail -4 gpurun_out/t_all.log
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/b_mma.log 2>&1; rc=$?; echo bench_rc=$rc
python - <<PY
import json
d=json.loads(open("gpurun_out/b_mma.log").read().strip().splitlines()[-1])
print("FULL ms_step", round(d["ms_per_step"],2), d["value"], {k:v["ms"] for k,v in d["kernels"].items()}, d["counts"]["survivors"], d["counts"]["hits_total"])
PY
if [ $rc -eq 0 ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1; echo list_rc=$?
fi
