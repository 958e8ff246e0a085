#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
export FT_REPS=3
bash profiles/scripts/r02c_filter_variants.sh ev1 ev0
timeout 1500 python -m pytest tests -m gpu -q --maxfail=5 -k "not c1_one and not two_gpu" > gpurun_out/r02p_tests.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/r02p_tests.log
timeout 900 python profiles/scripts/cluster_bench.py 50000000 > gpurun_out/r02p_cluster.jsonl 2> gpurun_out/r02p_cluster.err; echo "cluster rc=$?"
cat gpurun_out/r02p_cluster.jsonl; tail -3 gpurun_out/r02p_cluster.err
