#!/bin/bash
# round 2: hashed-sort decision once per build; GPU suite; whole configs sweep for the record
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/r02ae_tests.log 2>&1; echo "tests rc=$?"
tail -12 gpurun_out/r02ae_tests.log
timeout 900 python profiles/scripts/configs_bench.py > gpurun_out/r02ae_configs.jsonl 2> gpurun_out/r02ae_configs.err; echo "configs rc=$?"
cat gpurun_out/r02ae_configs.jsonl | cut -c1-420; tail -c 600 gpurun_out/r02ae_configs.err
