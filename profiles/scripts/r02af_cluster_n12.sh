#!/bin/bash
# round 2 (2 GPUs): configs[3] at 50 M fragments on one and on two GPUs (same generator as the 8-GPU run)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python bench.py --workload cluster --n-db 50000000 --steps 1 --warmup 0 > gpurun_out/r02af_cluster_n1_50M.json 2> gpurun_out/r02af_cluster_n1.err; echo "cluster n1 rc=$?"
tail -c 300 gpurun_out/r02af_cluster_n1.err; cat gpurun_out/r02af_cluster_n1_50M.json
timeout 600 $TR --nproc-per-node 2 --master-port 29551 bench.py --gpus 2 --workload cluster --n-db 50000000 --steps 1 --warmup 0 > gpurun_out/r02af_cluster_n2_50M.json 2> gpurun_out/r02af_cluster_n2.err; echo "cluster n2 rc=$?"
tail -c 300 gpurun_out/r02af_cluster_n2.err; cat gpurun_out/r02af_cluster_n2_50M.json
