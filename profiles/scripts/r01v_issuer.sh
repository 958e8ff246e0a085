set -x
mkdir -p gpurun_out
HS_NVCC_EXTRA=-DHS_MMA_PROF python hsearch_b200/build.py --force > gpurun_out/build_prof.log 2>&1; echo build_rc=$?
for D in 0 3; do
HS_MMA_DEBUG=$D timeout 600 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --n-db 50000000 > gpurun_out/b15_$D.log 2> gpurun_out/b15_$D.err; echo rc=$?
echo "DEBUG=$D"; grep "mma prof" gpurun_out/b15_$D.err | tail -1
done
