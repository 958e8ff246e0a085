#!/bin/bash
# round 2: new tests (protein id, orf, multi-device CLI, chunked cluster self-join) and the C4 clustering at 20 M / 50 M
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q --maxfail=5 -k "protein_id or orf or workflow or cli or cluster or reuse or residue" > gpurun_out/r02o_tests.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/r02o_tests.log
timeout 900 python profiles/scripts/cluster_bench.py 20000000 50000000 > gpurun_out/r02o_cluster.jsonl 2> gpurun_out/r02o_cluster.err; echo "cluster rc=$?"
cat gpurun_out/r02o_cluster.jsonl; tail -3 gpurun_out/r02o_cluster.err
