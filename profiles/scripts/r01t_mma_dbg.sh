set -x
mkdir -p gpurun_out
HS_NVCC_EXTRA=-DHS_MMA_PROF python hsearch_b200/build.py --force > gpurun_out/build_prof.log 2>&1; echo build_rc=$?
for D in 0 1 2 4 3 7; do
HS_MMA_DEBUG=$D timeout 600 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --n-db 50000000 > gpurun_out/b12_$D.log 2> gpurun_out/b12_$D.err; echo rc=$?
echo "DEBUG=$D"; grep "mma prof" gpurun_out/b12_$D.err | tail -1
tail -1 gpurun_out/b12_$D.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['stages_ms']['filter_tc'], d['counts']['survivors'])"
done
