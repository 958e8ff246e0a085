set -x
mkdir -p gpurun_out
for V in ${VARIANTS:-GP8k GP2k}; do
HS_LIBRARY=$PWD/hsearch_b200/libhs_$V.so timeout 600 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-recall > gpurun_out/b29_$V.log 2> gpurun_out/b29_$V.err; echo rc=$?
tail -1 gpurun_out/b29_$V.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$V', d['ms_per_step'], d['stages_ms']['permute'], d['stages_ms']['filter_tc'], d['counts']['hits_total'])"
done
