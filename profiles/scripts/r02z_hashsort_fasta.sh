#!/bin/bash
# round 2: hashed-key sort (K >= 8), FASTA parse on the device, whole GPU suite, K = 16 sweep points
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r02z_tests.log 2>&1; echo "tests rc=$?"
tail -25 gpurun_out/r02z_tests.log
timeout 600 python profiles/scripts/configs_bench.py 16,32,20 16,4,50 8,16,20 > gpurun_out/r02z_sweep.jsonl 2> gpurun_out/r02z_sweep.err; echo "sweep rc=$?"
cat gpurun_out/r02z_sweep.jsonl; tail -c 600 gpurun_out/r02z_sweep.err
HS_NO_HASH_SORT=1 timeout 600 python profiles/scripts/configs_bench.py 16,32,20 > gpurun_out/r02z_sweep_nohash.jsonl 2>> gpurun_out/r02z_sweep.err; echo "sweep nohash rc=$?"
cat gpurun_out/r02z_sweep_nohash.jsonl
