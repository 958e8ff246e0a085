set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cluster_sharded or cluster" > gpurun_out/t_cs.log 2>&1; echo rc=$?
tail -5 gpurun_out/t_cs.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/scripts/cluster_2gpu.py 2000000 > gpurun_out/cluster_2gpu.log 2>&1; echo rc=$?
grep cluster_sharded gpurun_out/cluster_2gpu.log || tail -20 gpurun_out/cluster_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?
tail -1 gpurun_out/bench_n2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['counts']['hits_total'])"
