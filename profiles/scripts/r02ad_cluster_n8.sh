#!/bin/bash
# round 2 (8 GPUs): configs[3] at its stated size, 50 M fragments, pair work split over 8 ranks
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
NG=${NG:-8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node $NG --master-port 29541 bench.py --gpus $NG --workload cluster --n-db 50000000 --steps 1 --warmup 0 > gpurun_out/r02ad_cluster_n${NG}_50M.json 2> gpurun_out/r02ad_cluster_n${NG}.err; echo "cluster n$NG rc=$?"
tail -c 500 gpurun_out/r02ad_cluster_n${NG}.err; cat gpurun_out/r02ad_cluster_n${NG}_50M.json
