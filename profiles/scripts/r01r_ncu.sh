set -x
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1; echo plain_rc=$?
for K in ${KERNELS:-hash_fast_kernel exact_kernel permute_rec_kernel rank_downsweep_kernel}; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_$K $CMD > gpurun_out/ncu_$K.log 2>&1
  echo $K rc=$?
done
ls -la gpurun_out | head -30
