#!/bin/bash
# round 2 (2 GPUs): hs_cluster on a communicator (replicated DB, split pair work, labels exchanged)
# against the one-GPU labels; cluster bench at 20 M fragments, N = 1 and N = 2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x --durations=5 > gpurun_out/r02y_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/r02y_tests.log
timeout 300 python bench.py --workload cluster > gpurun_out/r02y_cluster_n1.json 2> gpurun_out/r02y_cluster_n1.err; echo "cluster n1 rc=$?"
tail -c 600 gpurun_out/r02y_cluster_n1.err; cat gpurun_out/r02y_cluster_n1.json
timeout 300 $TR --nproc-per-node 2 --master-port 29531 bench.py --gpus 2 --workload cluster > gpurun_out/r02y_cluster_n2.json 2> gpurun_out/r02y_cluster_n2.err; echo "cluster n2 rc=$?"
tail -c 600 gpurun_out/r02y_cluster_n2.err; cat gpurun_out/r02y_cluster_n2.json
