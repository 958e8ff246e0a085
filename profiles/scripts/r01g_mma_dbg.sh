set -x
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_filter or bruteforce" > gpurun_out/t_mma.log 2>&1; echo mma_rc=$?
tail -5 gpurun_out/t_mma.log
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
for D in ${DBG_LIST:-0 1 3 7 15}; do
  HS_MMA_DEBUG=$D timeout 300 $B --n-db 20000000 > gpurun_out/b_dbg$D.log 2>&1; echo rc=$?
  python - <<PY
import json
d=json.loads(open("gpurun_out/b_dbg$D.log").read().strip().splitlines()[-1])
print("DEBUG=$D", "ms_step", round(d["ms_per_step"],2), {k:v["ms"] for k,v in d["kernels"].items() if "filter" in k or "exact" in k}, d["counts"]["survivors"], d["counts"]["hits_total"])
PY
done
