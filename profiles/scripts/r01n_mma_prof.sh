set -x
mkdir -p gpurun_out
HS_PLAN_STATS=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/b4.log 2> gpurun_out/b4.err; echo rc=$?
grep "\[plan\]" gpurun_out/b4.err | tail -2
HS_NVCC_EXTRA=-DHS_MMA_PROF python hsearch_b200/build.py --force > gpurun_out/build_prof.log 2>&1; echo build_rc=$?
timeout 600 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/b5.log 2> gpurun_out/b5.err; echo rc=$?
grep "mma prof" gpurun_out/b5.err | tail -2
