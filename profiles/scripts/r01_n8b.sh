set -x
mkdir -p gpurun_out
for AG in 1 0; do
HS_GATHER_ALLGATHER=$AG timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-8} --master-addr 127.0.0.1 --master-port 2951$AG bench.py --gpus ${NG:-8} --steps 4 --warmup 3 --no-e2e > gpurun_out/bench_ag$AG.json 2> gpurun_out/bench_ag$AG.err; echo rc=$?
tail -2 gpurun_out/bench_ag$AG.err | cut -c1-300
tail -1 gpurun_out/bench_ag$AG.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('AG=$AG', d['n_gpus'], d['value'], d['ms_per_step'], d['counts']['hits_total'])"
done
