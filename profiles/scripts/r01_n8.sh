set -x
mkdir -p gpurun_out
nvidia-smi -L | head -10
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-8} --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus ${NG:-8} --steps 3 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo rc=$?
tail -3 gpurun_out/bench_n8.err
tail -1 gpurun_out/bench_n8.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['counts']['hits_total'], d['checks'])"
