#!/bin/bash
# round 2: fix of the lazy code stores (records), len 30 on the pipelined tensor filter; GPU suite, sweep points, C5 sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r02ab_tests.log 2>&1; echo "tests rc=$?"
tail -25 gpurun_out/r02ab_tests.log
timeout 600 python profiles/scripts/configs_bench.py 16,32,20 16,4,50 8,16,20 8,4,50 4,4,10 allpairs > gpurun_out/r02ab_configs.jsonl 2> gpurun_out/r02ab_configs.err; echo "configs rc=$?"
cat gpurun_out/r02ab_configs.jsonl; tail -c 600 gpurun_out/r02ab_configs.err
