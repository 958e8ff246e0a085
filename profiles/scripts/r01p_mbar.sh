set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/b8.log 2> gpurun_out/b8.err; echo rc=$?
tail -1 gpurun_out/b8.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('NOHINT', d['ms_per_step'], d['stages_ms']['filter_tc'])"
HS_NVCC_EXTRA=-DHS_MMA_PROF python hsearch_b200/build.py --force > gpurun_out/build_prof.log 2>&1; echo build_rc=$?
timeout 600 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/b9.log 2> gpurun_out/b9.err; echo rc=$?
grep "mma prof" gpurun_out/b9.err | tail -1
