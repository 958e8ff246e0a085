set -x
mkdir -p gpurun_out
for P in 0 1; do
HS_NO_PIPELINE=$P timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b17_$P.log 2> gpurun_out/b17_$P.err; echo rc=$?
tail -1 gpurun_out/b17_$P.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('NO_PIPELINE=$P', d['ms_per_step'], d['e2e'])"
done
