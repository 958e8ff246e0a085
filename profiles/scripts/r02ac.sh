#!/bin/bash
# round 2: probe-side key fetch on the hashed path, lane-per-projection query hash, one-shift push_int; GPU suite, sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/r02ac_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/r02ac_tests.log
timeout 600 python profiles/scripts/configs_bench.py 16,32,20 16,4,50 8,16,20 8,4,50 4,4,10 4,4,50 > gpurun_out/r02ac_configs.jsonl 2> gpurun_out/r02ac_configs.err; echo "configs rc=$?"
cat gpurun_out/r02ac_configs.jsonl; tail -c 600 gpurun_out/r02ac_configs.err
