set -x
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_filter or bruteforce or search" > gpurun_out/t_mma.log 2>&1; echo mma_rc=$?
tail -5 gpurun_out/t_mma.log
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
for D in 0 1; do
  HS_MMA_DEBUG=$D timeout 300 $B --n-db 20000000 > gpurun_out/b_dbg$D.log 2>&1; echo rc=$?
  python - <<PY
import json
d=json.loads(open("gpurun_out/b_dbg$D.log").read().strip().splitlines()[-1])
print("DEBUG=$D", "ms_step", round(d["ms_per_step"],2), {k:v["ms"] for k,v in d["kernels"].items() if "filter" in k or "exact" in k}, d["counts"]["survivors"], d["counts"]["hits_total"])
PY
done
timeout 600 $B > gpurun_out/b_mma.log 2>&1; echo bench_rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/b_mma.log").read().strip().splitlines()[-1])
print("FULL ms_step", round(d["ms_per_step"],2), d["value"], {k:v["ms"] for k,v in d["kernels"].items()}, d["counts"]["survivors"], d["counts"]["hits_total"])
PY
