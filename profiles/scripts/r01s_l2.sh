set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu -k "hash or index or rank_path" > gpurun_out/t_hash.log 2>&1; echo hash_rc=$?
tail -3 gpurun_out/t_hash.log
for G in 0 32 64 128; do
HS_L2_FETCH=$G timeout 600 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/b11_$G.log 2> gpurun_out/b11_$G.err; echo rc=$?
tail -1 gpurun_out/b11_$G.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('L2FETCH=$G', d['ms_per_step'], d['stages_ms'])"
done
