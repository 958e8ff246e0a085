set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --n-db 20000000"
for D in 0 1 2 4 7 3; do
  HS_MMA_DEBUG=$D timeout 300 $B > gpurun_out/b_dbg$D.log 2>&1; echo rc=$?
  python - <<PY
import json
d=json.loads(open("gpurun_out/b_dbg$D.log").read().strip().splitlines()[-1])
print("DEBUG=$D", "ms_step", round(d["ms_per_step"],2), {k:v["ms"] for k,v in d["kernels"].items() if "filter" in k or "exact" in k}, d["counts"]["survivors"], d["counts"]["hits_total"])
PY
done
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --n-db 20000000"
timeout 300 $CMD > gpurun_out/plain_mma.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_mma_kernel -s 1 -c 1 -f -o gpurun_out/prof_filter_mma_kernel $CMD > gpurun_out/ncu_filter_mma.log 2>&1
echo ncu_rc=$?
